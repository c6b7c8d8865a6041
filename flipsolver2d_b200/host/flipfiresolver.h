// Host mirror of FlipSolver2dLib/flipfiresolver.h: smoke plus a fuel field / particle column.
#ifndef FS2D_HOST_FLIPFIRESOLVER_H
#define FS2D_HOST_FLIPFIRESOLVER_H

#include "flipsmokesolver.h"

struct FireSolverParameters : SmokeSolverParameters
{
    float ignitionTemperature;
    float burnRate;
    float smokeProportion;
    float heatProportion;
    float divergenceProportion;
};

class FlipFireSolver : public FlipSmokeSolver
{
public:
    explicit FlipFireSolver(const FireSolverParameters *p);
    void initAdditionalParameters() override;

protected:
    fs2d_params deviceParameters() const override;

    size_t m_fuelPropertyIndex = 0;
    float m_ignitionTemperature, m_burnRate, m_smokeProportion, m_heatProportion, m_divergenceProportion;
};

#endif
