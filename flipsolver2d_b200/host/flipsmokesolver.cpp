#include "flipsmokesolver.h"

#include <cmath>

FlipSmokeSolver::FlipSmokeSolver(const SmokeSolverParameters *p)
    : FlipSolver(p), m_temperature(p->gridSizeI, p->gridSizeJ, p->ambientTemperature, OOB_CONST, p->ambientTemperature),
      m_smokeConcentration(p->gridSizeI, p->gridSizeJ, 0.f, OOB_CONST, 0.f), m_ambientTemperature(p->ambientTemperature),
      m_temperatureDecayRate(p->temperatureDecayRate), m_concentrationDecayRate(p->concentrationDecayRate),
      m_buoyancyFactor(p->buoyancyFactor), m_sootFactor(p->sootFactor)
{
    m_projectTolerance = 1e-6;   // flipsmokesolver.cpp:19-20
    m_viscosityEnabled = false;
}

void FlipSmokeSolver::initAdditionalParameters()
{
    m_concentrationIndex = m_markerParticles.addParticleProperty<float>();
    m_temperatureIndex = m_markerParticles.addParticleProperty<float>();
}

fs2d_params FlipSmokeSolver::deviceParameters() const
{
    fs2d_params q = FlipSolver::deviceParameters();
    q.viscosity_property = -1;
    q.temperature_property = static_cast<int32_t>(m_temperatureIndex);
    q.concentration_property = static_cast<int32_t>(m_concentrationIndex);
    q.ambient_temperature = m_ambientTemperature;
    q.temperature_decay = m_temperatureDecayRate;
    q.concentration_decay = m_concentrationDecayRate;
    q.buoyancy_factor = m_buoyancyFactor;
    q.soot_factor = m_sootFactor;
    return q;
}

// flipsmokesolver.cpp:324-352
void FlipSmokeSolver::seedInitialFluid()
{
    m_seedProps.assign(m_markerParticles.propertyCount(), std::vector<float>());
    int rowLo = 0, rowHi = 0;
    seedRows(rowLo, rowHi);  // row slabs: the whole jitter stream is drawn, the particles of the own rows are kept
    for (ssize_t i = 0; i < m_sizeI; i++)
        for (ssize_t j = 0; j < m_sizeJ; j++)
        {
            if (!m_materialGrid.isStrictFluid(i, j)) continue;
            for (int p = 0; p < m_particlesPerCell; p++)
            {
                const Vec3 pos = jitteredPosInCell(i, j);
                const int row = static_cast<int>(std::floor(pos.x()));
                if (row < rowLo || row >= rowHi) continue;
                const Vec3 velocity = m_fluidVelocityGrid.velocityAt(pos);
                const float conc = m_smokeConcentration.interpolateAt(pos);
                const float temp = m_temperature.interpolateAt(pos);
                m_seedPos.push_back(pos.x());
                m_seedPos.push_back(pos.y());
                m_seedVel.push_back(velocity.x());
                m_seedVel.push_back(velocity.y());
                for (size_t c = 0; c < m_seedProps.size(); c++)
                    m_seedProps[c].push_back(c == m_temperatureIndex ? temp : (c == m_concentrationIndex ? conc : 0.f));
            }
        }
}

const Grid2d<float> FlipSmokeSolver::smokeConcentration() const
{
    fetchGrid(FS2D_GRID_CONCENTRATION, m_smokeConcentration.data().data(), linearSize() * 4);
    return m_smokeConcentration;
}

const Grid2d<float> FlipSmokeSolver::temperature() const
{
    fetchGrid(FS2D_GRID_TEMPERATURE, m_temperature.data().data(), linearSize() * 4);
    return m_temperature;
}
