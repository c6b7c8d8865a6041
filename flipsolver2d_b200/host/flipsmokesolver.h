// Host mirror of FlipSolver2dLib/flipsmokesolver.h: smoke = FlipSolver with temperature / soot fields,
// buoyancy body force and pressure rows for every non-solid cell. The differing stages
// (flipsmokesolver.cpp) are selected on the device by sim_type; this class carries the extra
// parameters, property columns, seeding rule and accessors.
#ifndef FS2D_HOST_FLIPSMOKESOLVER_H
#define FS2D_HOST_FLIPSMOKESOLVER_H

#include "flipsolver2d.h"

struct SmokeSolverParameters : FlipSolverParameters
{
    float ambientTemperature;
    float temperatureDecayRate;
    float concentrationDecayRate;
    float buoyancyFactor;
    float sootFactor;
};

class FlipSmokeSolver : public FlipSolver
{
public:
    explicit FlipSmokeSolver(const SmokeSolverParameters *p);

    const Grid2d<float> smokeConcentration() const;
    const Grid2d<float> temperature() const;
    void initAdditionalParameters() override;

protected:
    fs2d_params deviceParameters() const override;
    void seedInitialFluid() override;

    mutable Grid2d<float> m_temperature;
    mutable Grid2d<float> m_smokeConcentration;
    size_t m_temperatureIndex = 0, m_concentrationIndex = 0;
    float m_ambientTemperature, m_temperatureDecayRate, m_concentrationDecayRate, m_buoyancyFactor, m_sootFactor;
};

#endif
