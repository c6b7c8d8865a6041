// Host mirror of FlipSolver2dLib/nbflipsolver.h: narrow-band FLIP. Particles live only in a band below
// the surface; U, V, the level set and viscosity are also advected on the grid and combined with the
// particle fields (nbflipsolver.cpp:366-427). The substep order differs from FlipSolver::step
// (nbflipsolver.cpp:26-64) and frame 0 builds the level set from the initial-fluid polygons.
#ifndef FS2D_HOST_NBFLIPSOLVER_H
#define FS2D_HOST_NBFLIPSOLVER_H

#include "flipsolver2d.h"

struct NBFlipParameters : FlipSolverParameters
{
};

class NBFlipSolver : public FlipSolver
{
public:
    explicit NBFlipSolver(const NBFlipParameters *p);

protected:
    void step() override;
    void advect() override;
    void buildScene() override;
    void gridUpdate() override;
    void uploadScene() override;
    void initialFluidSeed();
    void fluidSdfFromInitialFluid();
    void sourceLevelset();

    Grid2d<float> m_sourceSdf;
    Grid2d<int> m_sourceSdfId;
};

#endif
