#include "flipfiresolver.h"

FlipFireSolver::FlipFireSolver(const FireSolverParameters *p)
    : FlipSmokeSolver(p), m_ignitionTemperature(p->ignitionTemperature), m_burnRate(p->burnRate),
      m_smokeProportion(p->smokeProportion), m_heatProportion(p->heatProportion), m_divergenceProportion(p->divergenceProportion)
{
}

void FlipFireSolver::initAdditionalParameters()
{
    FlipSmokeSolver::initAdditionalParameters();
    m_fuelPropertyIndex = m_markerParticles.addParticleProperty<float>();
}

fs2d_params FlipFireSolver::deviceParameters() const
{
    fs2d_params q = FlipSmokeSolver::deviceParameters();
    q.fuel_property = static_cast<int32_t>(m_fuelPropertyIndex);
    q.ignition_temperature = m_ignitionTemperature;
    q.burn_rate = m_burnRate;
    q.smoke_proportion = m_smokeProportion;
    q.heat_proportion = m_heatProportion;
    q.divergence_proportion = m_divergenceProportion;
    return q;
}
