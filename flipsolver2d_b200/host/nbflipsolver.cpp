#include "nbflipsolver.h"

#include <limits>

NBFlipSolver::NBFlipSolver(const NBFlipParameters *p)
    : FlipSolver(p), m_sourceSdf(p->gridSizeI, p->gridSizeJ, std::numeric_limits<float>::max()),
      m_sourceSdfId(p->gridSizeI, p->gridSizeJ, -1)
{
    m_projectTolerance = 1e-6;  // nbflipsolver.cpp:22
}

// NBFlipSolver::step (nbflipsolver.cpp:26-64)
void NBFlipSolver::step()
{
    advect();
    endStage(ADVECTION);
    pruneParticles();
    rebinParticles();
    endStage(PARTICLE_REBIN);
    particleToGrid();
    extrapolateVelocity(10);
    saveVelocity();
    endStage(PARTICLE_TO_GRID);
    gridUpdate();
    endStage(GRID_UPDATE);
    buildPressureSystem();
    endStage(DECOMPOSITION);
    project();
    endStage(PRESSURE);
    updateVelocityFromSolids();
    if (m_viscosityEnabled)
    {
        applyViscosity();
        endStage(VISCOSITY);
        project();
        endStage(REPRESSURE);
    }
    extrapolateVelocity(10);
    particleUpdate();
    endStage(PARTICLE_UPDATE);
    countParticles();
    reseedParticles();
    endStage(PARTICLE_RESEED);
}

// particles, then pruneNarrowBand + semi-Lagrangian U, V, sdf, viscosity (nbflipsolver.cpp:66-109)
void NBFlipSolver::advect()
{
    FlipSolver::advect();
    check(fs2d_nbflip_advect_grids(device()), "fs2d_nbflip_advect_grids");
}

// nbflipsolver.cpp:213-225
void NBFlipSolver::gridUpdate()
{
    updateSdf();
    extrapolateLevelsetOutside();
    afterTransfer();
    extrapolateLevelsetInside();
    endStage(AFTER_TRANSFER);
    updateMaterials();
    applyBodyForces();
}

// nbflipsolver.cpp:203-211
void NBFlipSolver::buildScene()
{
    updateSinks();
    updateSources();
    updateSolids();
    fluidSdfFromInitialFluid();
    m_fluidParticleCounts.fill(0);
    initialFluidSeed();
    sourceLevelset();
}

// nbflipsolver.cpp:297-327
void NBFlipSolver::fluidSdfFromInitialFluid()
{
    const float dx = static_cast<float>(m_dx);
    for (ssize_t i = 0; i < m_sizeI; i++)
        for (ssize_t j = 0; j < m_sizeJ; j++)
        {
            float dist = std::numeric_limits<float>::max();
            int fluidId = -1;
            for (size_t k = 0; k < m_initialFluid.size(); k++)
            {
                const float px = static_cast<float>((static_cast<float>(i) + 0.5) * dx);
                const float py = static_cast<float>((static_cast<float>(j) + 0.5) * dx);
                const float sdf = m_initialFluid[k].geometry().signedDistance(px, py) / dx;
                if (sdf < dist)
                {
                    dist = sdf;
                    fluidId = static_cast<int>(k);
                }
            }
            m_fluidSdf.at(i, j) = dist;
            if (dist < 0)
            {
                m_materialGrid.at(i, j) = FluidMaterial::FLUID;
                if (fluidId != -1) m_viscosityGrid.at(i, j) = m_initialFluid[fluidId].viscosity();
            }
        }
}

// initialFluidSeed (nbflipsolver.cpp:255-295) adds NO particles -- the adds are commented out in the
// reference -- but it still draws one jittered position per candidate, which advances the solver's
// mt19937 stream; later reseeding depends on that.
void NBFlipSolver::initialFluidSeed()
{
    m_seedProps.assign(m_markerParticles.propertyCount(), std::vector<float>());
    for (ssize_t i = 0; i < m_sizeI; i++)
        for (ssize_t j = 0; j < m_sizeJ; j++)
            if (m_fluidSdf.at(i, j) < 0.f)
                for (int p = 0; p < m_particlesPerCell - m_fluidParticleCounts.at(i, j); p++) (void)jitteredPosInCell(i, j);
}

// The per-cell source distance that updateGridFromSources recomputes every substep
// (nbflipsolver.cpp:329-364) is static, so it is rasterised once: min over sources of
// signedDistance(i*dx, j*dx)/dx and the arg-min.
void NBFlipSolver::sourceLevelset()
{
    const float dx = static_cast<float>(m_dx);
    for (ssize_t i = 0; i < m_sizeI; i++)
        for (ssize_t j = 0; j < m_sizeJ; j++)
        {
            float dist = std::numeric_limits<float>::max();
            int id = -1;
            for (size_t k = 0; k < m_sources.size(); k++)
            {
                const float sdf = m_sources[k].geometry().signedDistance(static_cast<float>(i) * dx, static_cast<float>(j) * dx) / dx;
                if (sdf < dist)
                {
                    dist = sdf;
                    id = static_cast<int>(k);
                }
            }
            m_sourceSdf.at(i, j) = dist;
            m_sourceSdfId.at(i, j) = id;
        }
}

void NBFlipSolver::uploadScene()
{
    FlipSolver::uploadScene();
    check(fs2d_upload_grid(device(), FS2D_GRID_SOURCE_SDF, m_sourceSdf.data().data(), linearSize() * 4), "upload sourceSdf");
    check(fs2d_upload_grid(device(), FS2D_GRID_SOURCE_SDF_ID, m_sourceSdfId.data().data(), linearSize() * 4), "upload sourceSdfId");
    check(fs2d_upload_grid(device(), FS2D_GRID_COUNTS, m_fluidParticleCounts.data().data(), linearSize() * 4), "upload counts");
}
