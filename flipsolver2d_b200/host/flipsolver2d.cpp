#include "flipsolver2d.h"

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <limits>
#include <sstream>
#include <stdexcept>
#include <thread>

namespace
{
bool g_quiet = std::getenv("FS2D_QUIET") != nullptr;
int g_device = std::getenv("FS2D_DEVICE") ? std::atoi(std::getenv("FS2D_DEVICE")) : 0;
int g_convergenceThreads = std::getenv("FS2D_CONVERGENCE_THREADS") ? std::atoi(std::getenv("FS2D_CONVERGENCE_THREADS")) : 0;
int g_slabRank = 0, g_slabWorld = 1, g_slabShare = 1;

// Frame-0 rasterisation runs once on the host; rows are independent.
template <class F> void parallelRows(ssize_t rows, F f)
{
    unsigned int n = std::max(1u, std::min<unsigned int>(std::thread::hardware_concurrency(), 32u));
    if (rows < 256) n = 1;
    std::vector<std::thread> pool;
    for (unsigned int t = 0; t < n; t++)
    {
        const ssize_t b = rows * t / n, e = rows * (t + 1) / n;
        pool.emplace_back([=]() { for (ssize_t i = b; i < e; i++) f(i); });
    }
    for (std::thread &t : pool) t.join();
}
}  // namespace

void FlipSolver::setQuiet(bool q) { g_quiet = q; }
void FlipSolver::setDevice(int ordinal) { g_device = ordinal; }
void FlipSolver::setConvergenceThreads(int t) { g_convergenceThreads = t; }
void FlipSolver::setSlab(int rank, int world, int deviceShare)
{
    g_slabRank = rank;
    g_slabWorld = world;
    g_slabShare = deviceShare;
}

FlipSolver::FlipSolver(const FlipSolverParameters *p)
    : LinearIndexable2d(p->gridSizeI, p->gridSizeJ), m_randEngine(p->seed), m_markerParticles(p->gridSizeI, p->gridSizeJ, 3),
      m_fluidVelocityGrid(p->gridSizeI, p->gridSizeJ), m_materialGrid(p->gridSizeI, p->gridSizeJ, FluidMaterial::SINK),
      m_solidSdf(p->gridSizeI, p->gridSizeJ), m_fluidSdf(p->gridSizeI, p->gridSizeJ),
      m_viscosityGrid(p->gridSizeI, p->gridSizeJ, 0.f, OOB_EXTEND, 0.f, Vec3(0.5f, 0.5f)),
      m_emitterId(p->gridSizeI, p->gridSizeJ, -1), m_solidId(p->gridSizeI, p->gridSizeJ, -1),
      m_fluidParticleCounts(p->gridSizeI, p->gridSizeJ), m_divergenceControl(p->gridSizeI, p->gridSizeJ, 0.f, OOB_CONST, 0.f),
      m_testGrid(p->gridSizeI, p->gridSizeJ), m_stepDt(1.f / p->fps), m_frameDt(1.f / p->fps), m_dx(p->dx),
      m_fluidDensity(p->fluidDensity), m_seed(p->seed), m_particlesPerCell(p->particlesPerCell),
      m_globalAcceleration(p->globalAcceleration), m_resolution(p->resolution), m_fps(p->fps), m_maxSubsteps(p->maxSubsteps),
      m_picRatio(p->picRatio), m_cflNumber(p->cflNumber), m_particleScale(p->particleScale), m_pcgIterLimit(p->pcgIterLimit),
      m_domainSizeI(p->domainSizeI), m_domainSizeJ(p->domainSizeJ), m_sceneScale(p->sceneScale),
      m_viscosityEnabled(p->viscosityEnabled), m_simulationMethod(p->simulationMethod),
      m_parameterHandlingMethod(p->parameterHandlingMethod)
{
    m_testValuePropertyIndex = m_markerParticles.addParticleProperty<float>();
    m_projectTolerance = m_viscosityEnabled ? 1e-6 : 1e-2;  // flipsolver2d.cpp:73
    m_useHeavyViscosity = p->useHeavyViscosity;  // flipsolver2d.cpp:77-85
}

FlipSolver::~FlipSolver()
{
    if (m_device) fs2d_destroy(m_device);
}

void FlipSolver::initAdditionalParameters() { m_viscosityPropertyIndex = m_markerParticles.addParticleProperty<float>(); }

void FlipSolver::check(int rc, const char *what) const
{
    if (rc == FS2D_OK) return;
    std::string msg = std::string(what) + " failed (" + std::to_string(rc) + ")";
    if (m_device) msg += std::string(": ") + fs2d_last_error(m_device);
    // The reference has no error path on step(); a CUDA failure here is unrecoverable, fail loudly.
    throw std::runtime_error(msg);
}

fs2d_params FlipSolver::deviceParameters() const
{
    fs2d_params q;
    std::memset(&q, 0, sizeof(q));
    q.size_i = static_cast<int32_t>(m_sizeI);
    q.size_j = static_cast<int32_t>(m_sizeJ);
    q.num_properties = static_cast<int32_t>(m_markerParticles.propertyCount());
    q.particles_per_cell = m_particlesPerCell;
    q.pcg_iter_limit = m_pcgIterLimit;
    q.sim_type = m_simulationMethod;
    q.parameter_handling = m_parameterHandlingMethod;
    q.viscosity_enabled = m_viscosityEnabled ? 1 : 0;
    q.heavy_viscosity = m_useHeavyViscosity ? 1 : 0;
    q.convergence_threads = g_convergenceThreads;
    q.device = g_device;
    q.viscosity_property = m_viscosityPropertyIndex == static_cast<size_t>(-1) ? -1 : static_cast<int32_t>(m_viscosityPropertyIndex);
    q.temperature_property = q.concentration_property = q.fuel_property = -1;
    q.test_property = static_cast<int32_t>(m_testValuePropertyIndex);
    q.dx = m_dx;
    q.fluid_density = m_fluidDensity;
    q.project_tolerance = m_projectTolerance;
    q.gravity_x = m_globalAcceleration.x();
    q.gravity_y = m_globalAcceleration.y();
    q.pic_ratio = m_picRatio;
    q.particle_scale = m_particleScale;
    q.ambient_temperature = 273.f;
    q.buoyancy_factor = q.soot_factor = 1.f;
    return q;
}

fs2d_handle FlipSolver::device()
{
    if (!m_device)
    {
        const fs2d_params q = deviceParameters();
        const int rc = fs2d_create(&q, &m_device);
        if (rc != FS2D_OK)
            throw std::runtime_error("fs2d_create failed (" + std::to_string(rc) +
                                     "): this solver needs a CUDA device, there is no CPU path");
        if (g_slabWorld > 1)
        {
            m_slabRank = g_slabRank;
            m_slabWorld = g_slabWorld;
            const std::vector<int32_t> bounds = slabBounds(m_slabWorld);
            check(fs2d_slab_configure_rows(m_device, m_slabRank, m_slabWorld, g_slabShare, bounds.data()), "fs2d_slab_configure_rows");
        }
    }
    return m_device;
}

// Slab boundaries balanced by work instead of by rows: the seed particles per 16-row tile row (every rank rasterises
// the same scene, so every rank computes the same table) plus a small cost per row for the dense grid passes. In a dam break the fluid fills the lower half of the tank only: equal row counts would leave half of
// the GPUs without a single particle or matrix row.
std::vector<int32_t> FlipSolver::slabBounds(int world)
{
    prepareHost();
    return slabBoundsFromMaterial(world);
}

// seedInitialFluid puts exactly particlesPerCell particles into every strict-FLUID cell (flipsolver2d.cpp:682-707), so the
// work per tile row is known from the rasterised material grid alone -- before (and without) drawing a single particle.
std::vector<int32_t> FlipSolver::slabBoundsFromMaterial(int world) const
{
    const int tileRows = static_cast<int>((m_sizeI + 15) / 16);
    std::vector<double> w(static_cast<size_t>(tileRows), 0.0);
    double n = 0.0;
    for (ssize_t i = 0; i < m_sizeI; i++)
    {
        size_t fluid = 0;
        for (ssize_t j = 0; j < m_sizeJ; j++) fluid += m_materialGrid.isStrictFluid(i, j) ? 1 : 0;
        w[static_cast<size_t>(i / 16)] += static_cast<double>(fluid) * m_particlesPerCell;
        n += static_cast<double>(fluid) * m_particlesPerCell;
    }
    const double base = std::max(1.0, 0.02 * n / tileRows);
    double total = 0.0;
    for (double &x : w)
    {
        x += base;
        total += x;
    }
    std::vector<int32_t> b(static_cast<size_t>(world) + 1, 0);
    const int minTiles = 2;  // a slab holds at least 32 rows (the halo width)
    if (tileRows < minTiles * world) throw std::runtime_error("the grid has too few rows for this many slabs");
    double prefix = 0.0;
    int t = 0;
    for (int r = 1; r < world; r++)
    {
        const double target = total * r / world;
        while (t < tileRows && prefix + w[static_cast<size_t>(t)] <= target)
        {
            prefix += w[static_cast<size_t>(t)];
            t++;
        }
        int cut = t;
        cut = std::max(cut, b[static_cast<size_t>(r) - 1] / 16 + minTiles);
        cut = std::min(cut, tileRows - minTiles * (world - r));
        while (t < cut)
        {
            prefix += w[static_cast<size_t>(t)];
            t++;
        }
        b[static_cast<size_t>(r)] = 16 * cut;
    }
    b[static_cast<size_t>(world)] = static_cast<int32_t>(m_sizeI);
    return b;
}

void FlipSolver::slabExport(void *blob) { check(fs2d_slab_export(device(), blob), "fs2d_slab_export"); }
void FlipSolver::slabConnect(int peerRank, const void *blob) { check(fs2d_slab_connect(device(), peerRank, blob), "fs2d_slab_connect"); }

size_t FlipSolver::globalParticleCount()
{
    if (m_slabWorld == 1) return particleCount();
    int64_t v[4] = {static_cast<int64_t>(particleCount()), 0, 0, 0}, all[4 * 8];
    check(fs2d_slab_allgather(device(), v, all), "fs2d_slab_allgather");
    size_t total = 0;
    for (int r = 0; r < m_slabWorld; r++) total += static_cast<size_t>(all[4 * r]);
    return total;
}

int64_t FlipSolver::kernelLaunches() { return m_device ? fs2d_launch_count(m_device) : 0; }

void FlipSolver::endStage(SolverStage s)
{
    check(fs2d_synchronize(device()), "fs2d_synchronize");
    m_stats.endStage(s);
}

// ------------------------------------------------------------------ frame loop (flipsolver2d.cpp:464-500)
void FlipSolver::prepareHost()
{
    if (m_frameNumber == 0 && !m_sceneBuilt)
    {
        buildScene();
        m_sceneBuilt = true;
    }
}

void FlipSolver::prepare()
{
    if (m_frameNumber == 0 && !m_prepared)
    {
        firstFrameInit();
        m_prepared = true;
    }
}

// stepFrame = CFL sub-stepping until the frame time is used up. The loop body is exposed as
// stepSubstep() so that a caller (bench.py) can advance by exactly one substep; the arithmetic and
// the order of operations are those of flipsolver2d.cpp:464-500.
bool FlipSolver::stepSubstep()
{
    if (!m_inFrame)
    {
        prepare();
        check(fs2d_clear_grid(device(), FS2D_GRID_TEST), "fs2d_clear_grid");
        m_substepTime = 0.f;
        m_substepCount = 0;
        m_stats.reset();
        m_inFrame = true;
    }
    bool finished = false;
    const float vel = maxParticleVelocity();
    float maxSubstepSize = m_cflNumber / (vel + 1e-15f);
    if (m_substepTime + maxSubstepSize >= m_frameDt || m_substepCount == (m_maxSubsteps - 1))
    {
        maxSubstepSize = m_frameDt - m_substepTime;
        finished = true;
    }
    else if (m_substepTime + 2.f * maxSubstepSize >= m_frameDt)
    {
        maxSubstepSize = 0.5f * (m_frameDt - m_substepTime);
    }
    m_stepDt = maxSubstepSize;
    if (!g_quiet) std::cout << "Substep " << m_substepCount << " substep dt: " << m_stepDt << " vel " << vel << std::endl;
    check(fs2d_set_step_dt(device(), m_stepDt), "fs2d_set_step_dt");
    step();
    m_stats.addSubstep();
    m_substepTime += maxSubstepSize;
    m_substepCount++;
    if (m_substepCount > 50) finished = true;
    if (finished)
    {
        m_stats.endFrame();
        m_frameNumber++;
        m_inFrame = false;
    }
    invalidateMirrors();
    return finished;
}

void FlipSolver::stepFrame()
{
    while (!stepSubstep()) {}
}

bool FlipSolver::stepSubstepStreamed(void *hostBuf, int64_t capacityRecords, int64_t countIn, int64_t *countOut)
{
    prepare();
    check(fs2d_particle_stream_begin(device(), hostBuf, countIn, capacityRecords), "fs2d_particle_stream_begin");
    m_streamBuf = hostBuf;
    m_streamCapacity = capacityRecords;
    bool finished = false;
    try
    {
        finished = stepSubstep();
    }
    catch (...)
    {
        m_streamBuf = nullptr;
        throw;
    }
    m_streamBuf = nullptr;
    int64_t n = 0;
    check(fs2d_particle_stream_end(device(), hostBuf, capacityRecords, &n), "fs2d_particle_stream_end");
    if (countOut) *countOut = n;
    return finished;
}

// ------------------------------------------------------------------ state dump / restore
namespace
{
struct HostStateHeader
{
    char magic[8];  // "FS2DHOST"
    uint32_t version;
    int32_t frameNumber, inFrame, substepCount;
    float substepTime, stepDt;
    int32_t statSubsteps, pressureIters, densityIters, viscosityIters;
    float stageMs[SOLVER_STAGE_COUNT];
    uint64_t rngBytes, deviceBytes;
};
}  // namespace

void FlipSolver::saveState(const std::string &path)
{
    prepare();
    size_t need = 0;
    check(fs2d_state_bytes(device(), &need), "fs2d_state_bytes");
    std::vector<unsigned char> blob(need + 64);
    size_t written = 0;
    check(fs2d_state_save(device(), blob.data(), blob.size(), &written), "fs2d_state_save");
    std::ostringstream rng;
    rng << m_randEngine;
    const std::string rngText = rng.str();
    HostStateHeader h;
    std::memset(&h, 0, sizeof(h));
    std::memcpy(h.magic, "FS2DHOST", 8);
    h.version = 1;
    h.frameNumber = m_frameNumber;
    h.inFrame = m_inFrame ? 1 : 0;
    h.substepCount = m_substepCount;
    h.substepTime = m_substepTime;
    h.stepDt = m_stepDt;
    h.statSubsteps = m_stats.substepCount();
    h.pressureIters = m_stats.pressureIterations();
    h.densityIters = m_stats.densityIterations();
    h.viscosityIters = m_stats.viscosityIterations();
    const SolverStats::StageTimings t = m_stats.timings();
    for (int k = 0; k < SOLVER_STAGE_COUNT; k++) h.stageMs[k] = t[static_cast<size_t>(k)];
    h.rngBytes = rngText.size();
    h.deviceBytes = written;
    std::ofstream f(path, std::ios::binary | std::ios::trunc);
    f.write(reinterpret_cast<const char *>(&h), sizeof(h));
    f.write(rngText.data(), static_cast<std::streamsize>(rngText.size()));
    f.write(reinterpret_cast<const char *>(blob.data()), static_cast<std::streamsize>(written));
    if (!f) throw std::runtime_error("saveState: cannot write " + path);
}

void FlipSolver::loadState(const std::string &path)
{
    std::ifstream f(path, std::ios::binary);
    HostStateHeader h;
    f.read(reinterpret_cast<char *>(&h), sizeof(h));
    if (!f || std::memcmp(h.magic, "FS2DHOST", 8) != 0 || h.version != 1) throw std::runtime_error("loadState: not a state file: " + path);
    std::string rngText(h.rngBytes, '\0');
    f.read(rngText.data(), static_cast<std::streamsize>(h.rngBytes));
    std::vector<unsigned char> blob(h.deviceBytes);
    f.read(reinterpret_cast<char *>(blob.data()), static_cast<std::streamsize>(h.deviceBytes));
    if (!f) throw std::runtime_error("loadState: truncated state file: " + path);
    prepare();  // device, scene tables (sources, obstacles) and the static grids; everything dynamic is overwritten below
    check(fs2d_state_load(device(), blob.data(), blob.size()), "fs2d_state_load");
    std::istringstream rng(rngText);
    rng >> m_randEngine;
    m_frameNumber = h.frameNumber;
    m_prepared = true;
    m_inFrame = h.inFrame != 0;
    m_substepCount = h.substepCount;
    m_substepTime = h.substepTime;
    m_stepDt = h.stepDt;
    SolverStats::StageTimings t;
    for (int k = 0; k < SOLVER_STAGE_COUNT; k++) t[static_cast<size_t>(k)] = h.stageMs[k];
    m_stats.restore(t, h.statSubsteps, h.pressureIters, h.densityIters, h.viscosityIters);
    invalidateMirrors();
}

// ------------------------------------------------------------------ one substep (flipsolver2d.cpp:412-462)
void FlipSolver::step()
{
    advect();
    endStage(ADVECTION);
    buildPressureSystem();
    endStage(DECOMPOSITION);
    pruneParticles();
    rebinParticles();
    endStage(PARTICLE_REBIN);
    if (!m_viscosityEnabled)
    {
        densityCorrection();
        endStage(DENSITY);
    }
    gridUpdate();
    endStage(GRID_UPDATE);
    afterTransfer();
    extrapolateLevelsetInside();
    extrapolateVelocity(10);
    saveVelocity();
    applyBodyForces();
    endStage(AFTER_TRANSFER);
    // Streamed particle state: since the density correction no stage moves or reorders the existing records (reseeding
    // appends, the count cap flags) and the liquid solver never rewrites property columns, so positions and columns can
    // leave for the host. They do so HERE, under the pressure solve -- a latency-bound kernel working out of L2 that the
    // copy does not disturb -- rather than under the gathers right after the density correction, which it slowed down by
    // ~20 % (P2G 0.94 -> 1.27 ms at 4096^2).
    if (m_streamBuf) check(fs2d_particle_stream_positions_final(device(), m_streamBuf, m_streamCapacity, 1), "fs2d_particle_stream_positions_final");
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
    // streamed particle state: the velocities of the existing records are final (the count cap flags, reseeding appends)
    if (m_streamBuf) check(fs2d_particle_stream_velocities_final(device(), m_streamBuf, m_streamCapacity), "fs2d_particle_stream_velocities_final");
    endStage(PARTICLE_UPDATE);
    countParticles();
    reseedParticles();
    endStage(PARTICLE_RESEED);
}

void FlipSolver::advect() { check(fs2d_advect(device()), "fs2d_advect"); }
void FlipSolver::buildPressureSystem() { check(fs2d_build_matrix(device()), "fs2d_build_matrix"); }
void FlipSolver::pruneParticles() {}
void FlipSolver::rebinParticles() { check(fs2d_sort_particles(device()), "fs2d_sort_particles"); }

void FlipSolver::densityCorrection()
{
    int iters = 0;
    check(fs2d_density_correction(device(), &iters), "fs2d_density_correction");
    m_stats.setDensityIters(iters);
    if (iters >= m_pcgIterLimit)
    {
        if (!g_quiet) std::cout << "Density solver solving failed!\n";
        return;
    }
    if (!g_quiet) std::cout << "Density correction done\n";
}

void FlipSolver::gridUpdate()
{
    particleToGrid();
    endStage(PARTICLE_TO_GRID);
    updateSdf();
    updateMaterials();
}

void FlipSolver::particleToGrid() { check(fs2d_particle_to_grid(device()), "fs2d_particle_to_grid"); }
void FlipSolver::updateSdf() { check(fs2d_update_sdf(device()), "fs2d_update_sdf"); }
void FlipSolver::updateMaterials() { check(fs2d_update_materials(device()), "fs2d_update_materials"); }
void FlipSolver::afterTransfer() { check(fs2d_after_transfer(device()), "fs2d_after_transfer"); }
void FlipSolver::extrapolateLevelsetInside() { check(fs2d_extrapolate_sdf_inside(device()), "fs2d_extrapolate_sdf_inside"); }
void FlipSolver::extrapolateLevelsetOutside() { check(fs2d_extrapolate_sdf_outside(device()), "fs2d_extrapolate_sdf_outside"); }
void FlipSolver::extrapolateVelocity(int radius) { check(fs2d_extrapolate_velocity(device(), radius), "fs2d_extrapolate_velocity"); }
void FlipSolver::saveVelocity() { check(fs2d_save_velocity(device()), "fs2d_save_velocity"); }
void FlipSolver::applyBodyForces() { check(fs2d_apply_body_forces(device()), "fs2d_apply_body_forces"); }

void FlipSolver::project()
{
    int iters = 0;
    check(fs2d_project(device(), &iters), "fs2d_project");
    m_stats.setPressureIterations(iters);
    if (iters >= m_pcgIterLimit && !g_quiet) std::cout << "Pressure solver solving failed! Expect bogus pressures\n";
}

void FlipSolver::updateVelocityFromSolids() { check(fs2d_velocity_from_solids(device()), "fs2d_velocity_from_solids"); }

void FlipSolver::applyViscosity()
{
    int iters = 0;
    check(fs2d_apply_viscosity(device(), &iters), "fs2d_apply_viscosity");
    m_stats.setViscosityIterations(iters);
}

void FlipSolver::particleUpdate() { check(fs2d_particle_update(device()), "fs2d_particle_update"); }
void FlipSolver::countParticles() { check(fs2d_count_particles(device()), "fs2d_count_particles"); }

// reseedParticles: the device plans how many jittered positions every cell needs, the host draws
// them from the solver's mt19937 in row-major cell order (x first, then y) and the device places them.
void FlipSolver::reseedParticles()
{
    int64_t candidates = 0;
    check(fs2d_reseed_plan(device(), &candidates), "fs2d_reseed_plan");
    static std::uniform_real_distribution<float> dist(0.f, 1.f);
    if (m_slabWorld > 1)
    {
        // every rank plans its own rows; the draws of all ranks strung together in rank order are the reference's
        // row-major stream, so each rank draws the whole frame's stream (same seed everywhere) and uses its slice
        int64_t v[4] = {candidates, 0, 0, 0}, all[4 * 8];
        check(fs2d_slab_allgather(device(), v, all), "fs2d_slab_allgather");
        int64_t before = 0, total = 0;
        for (int r = 0; r < m_slabWorld; r++)
        {
            if (r < m_slabRank) before += all[4 * r];
            total += all[4 * r];
        }
        if (total == 0) return;
        std::vector<float> u(static_cast<size_t>(2 * total));
        for (float &x : u) x = dist(m_randEngine);
        check(fs2d_reseed_apply(device(), candidates, u.data() + 2 * before), "fs2d_reseed_apply");
        return;
    }
    if (candidates == 0) return;
    std::vector<float> u(static_cast<size_t>(2 * candidates));
    for (float &v : u) v = dist(m_randEngine);
    check(fs2d_reseed_apply(device(), candidates, u.data()), "fs2d_reseed_apply");
}

float FlipSolver::maxParticleVelocity()
{
    float v = 0.f;
    check(fs2d_max_particle_velocity(device(), &v), "fs2d_max_particle_velocity");
    return v;
}

Vec3 FlipSolver::jitteredPosInCell(size_t i, size_t j)
{
    static std::uniform_real_distribution<float> dist(0.f, 1.f);
    const float x = static_cast<float>(i) + dist(m_randEngine);
    const float y = static_cast<float>(j) + dist(m_randEngine);
    return Vec3(x, y);
}

// ------------------------------------------------------------------ frame-0 scene rasterisation
// flipsolver2d.cpp:502-590, restated including the float/double mix of every sample coordinate.
void FlipSolver::updateSolids()
{
    const float dx = static_cast<float>(m_dx);
    parallelRows(m_sizeI, [&](ssize_t i) {
        for (ssize_t j = 0; j < m_sizeJ; j++)
        {
            float dist = std::numeric_limits<float>::max();
            int minIdx = 0;
            for (size_t k = 0; k < m_obstacles.size(); k++)
            {
                const float px = static_cast<float>((static_cast<float>(i) + 0.5) * dx);
                const float py = static_cast<float>((static_cast<float>(j) + 0.5) * dx);
                const float sdf = m_obstacles[k].geometry().signedDistance(px, py) / dx;
                if (sdf < dist)
                {
                    minIdx = static_cast<int>(k);
                    dist = sdf;
                }
            }
            m_solidSdf.at(i, j) = dist;
            if (dist < 0)
            {
                m_materialGrid.at(i, j) = FluidMaterial::SOLID;
                m_solidId.at(i, j) = minIdx;
            }
        }
    });
}

void FlipSolver::updateSources()
{
    const float dx = static_cast<float>(m_dx);
    const float hdx = static_cast<float>(dx / 2.0);
    parallelRows(m_sizeI, [&](ssize_t i) {
        for (ssize_t j = 0; j < m_sizeJ; j++)
            for (size_t k = 0; k < m_sources.size(); k++)
            {
                Emitter &e = m_sources[k];
                if (e.geometry().signedDistance(static_cast<size_t>(i) * dx + hdx, static_cast<size_t>(j) * dx + hdx) <= 0.f)
                {
                    m_materialGrid.at(i, j) = FluidMaterial::SOURCE;
                    m_emitterId.at(i, j) = static_cast<int>(k);
                    m_divergenceControl.at(i, j) = e.divergence();
                }
            }
    });
}

void FlipSolver::updateSinks()
{
    const float dx = static_cast<float>(m_dx);
    parallelRows(m_sizeI, [&](ssize_t i) {
        for (ssize_t j = 0; j < m_sizeJ; j++)
            for (Sink &s : m_sinks)
                if (s.geo().signedDistance((static_cast<float>(i) + 0.5f) * dx, (static_cast<float>(j) + 0.5f) * dx) <= 0.f)
                {
                    m_materialGrid.at(i, j) = FluidMaterial::SINK;
                    m_divergenceControl.at(i, j) = s.divergence();
                }
    });
}

void FlipSolver::updateInitialFluid()
{
    const float dx = static_cast<float>(m_dx);
    parallelRows(m_sizeI, [&](ssize_t i) {
        for (ssize_t j = 0; j < m_sizeJ; j++)
            for (Emitter &e : m_initialFluid)
                if (e.geometry().signedDistance((static_cast<float>(i) + 0.5f) * dx, (static_cast<float>(j) + 0.5f) * dx) <= 0.f)
                {
                    m_materialGrid.at(i, j) = FluidMaterial::FLUID;
                    m_viscosityGrid.at(i, j) = e.viscosity();
                }
    });
}

// The cell rows whose seed particles this solver keeps (all rows unless it owns a row slab).
void FlipSolver::seedRows(int &rowLo, int &rowHi) const
{
    rowLo = 0;
    rowHi = static_cast<int>(m_sizeI);
    if (g_slabWorld > 1)
    {
        const std::vector<int32_t> b = slabBoundsFromMaterial(g_slabWorld);
        rowLo = b[static_cast<size_t>(g_slabRank)];
        rowHi = b[static_cast<size_t>(g_slabRank) + 1];
    }
}

// seedInitialFluid (flipsolver2d.cpp:682-707): ppc jittered particles per strict-FLUID cell, row major;
// velocity and viscosity sampled from the (frame-0) grids.
void FlipSolver::seedInitialFluid()
{
    m_seedProps.assign(m_markerParticles.propertyCount(), std::vector<float>());
    // Row slabs: every rank draws the WHOLE mt19937 stream (the jitter of a particle depends on all particles before
    // it) but keeps only the particles of its own rows -- at 8192^2 with a full tank that is 50 M instead of 400 M
    // records of host memory per rank.
    int rowLo = 0, rowHi = 0;
    seedRows(rowLo, rowHi);
    for (ssize_t i = 0; i < m_sizeI; i++)
        for (ssize_t j = 0; j < m_sizeJ; j++)
        {
            if (!m_materialGrid.isStrictFluid(i, j)) continue;
            for (int p = 0; p < m_particlesPerCell; p++)
            {
                const Vec3 pos = jitteredPosInCell(i, j);
                const int row = static_cast<int>(std::floor(pos.x()));  // the cell row the device files it under
                if (row < rowLo || row >= rowHi) continue;
                const Vec3 velocity = m_fluidVelocityGrid.velocityAt(pos);
                const float viscosity = m_viscosityGrid.interpolateAt(pos);
                m_seedPos.push_back(pos.x());
                m_seedPos.push_back(pos.y());
                m_seedVel.push_back(velocity.x());
                m_seedVel.push_back(velocity.y());
                for (size_t c = 0; c < m_seedProps.size(); c++)
                    m_seedProps[c].push_back(c == m_viscosityPropertyIndex ? viscosity : 0.f);
            }
        }
}

void FlipSolver::uploadSeed()
{
    fs2d_handle h = device();
    if (m_slabWorld > 1)
    {
        // every rank seeded the whole scene (same mt19937 stream); keep the particles whose cell row is ours
        int lo = 0, hi = 0;
        check(fs2d_slab_rows(h, &lo, &hi, nullptr), "fs2d_slab_rows");
        size_t kept = 0;
        const size_t all = m_seedPos.size() / 2;
        for (size_t p = 0; p < all; p++)
        {
            const int row = static_cast<int>(std::floor(m_seedPos[2 * p]));
            if (row < lo || row >= hi) continue;
            m_seedPos[2 * kept] = m_seedPos[2 * p];
            m_seedPos[2 * kept + 1] = m_seedPos[2 * p + 1];
            m_seedVel[2 * kept] = m_seedVel[2 * p];
            m_seedVel[2 * kept + 1] = m_seedVel[2 * p + 1];
            for (auto &c : m_seedProps) c[kept] = c[p];
            kept++;
        }
        m_seedPos.resize(2 * kept);
        m_seedVel.resize(2 * kept);
        for (auto &c : m_seedProps) c.resize(kept);
    }
    const size_t n = m_seedPos.size() / 2;
    std::vector<float> props;
    for (auto &c : m_seedProps) props.insert(props.end(), c.begin(), c.end());
    check(fs2d_upload_particles(device(), static_cast<int64_t>(n), m_seedPos.data(), m_seedVel.data(), props.empty() ? nullptr : props.data()),
          "fs2d_upload_particles");
    std::vector<float>().swap(m_seedPos);
    std::vector<float>().swap(m_seedVel);
    m_seedProps.clear();
}

void FlipSolver::uploadScene()
{
    fs2d_handle h = device();
    const size_t N = linearSize();
    check(fs2d_upload_grid(h, FS2D_GRID_MATERIAL, m_materialGrid.data().data(), N), "upload material");
    check(fs2d_upload_grid(h, FS2D_GRID_SOLID_SDF, m_solidSdf.data().data(), N * 4), "upload solidSdf");
    check(fs2d_upload_grid(h, FS2D_GRID_FLUID_SDF, m_fluidSdf.data().data(), N * 4), "upload fluidSdf");
    check(fs2d_upload_grid(h, FS2D_GRID_VISCOSITY, m_viscosityGrid.data().data(), N * 4), "upload viscosity");
    check(fs2d_upload_grid(h, FS2D_GRID_EMITTER_ID, m_emitterId.data().data(), N * 4), "upload emitterId");
    check(fs2d_upload_grid(h, FS2D_GRID_SOLID_ID, m_solidId.data().data(), N * 4), "upload solidId");
    check(fs2d_upload_grid(h, FS2D_GRID_DIVERGENCE_CONTROL, m_divergenceControl.data().data(), N * 4), "upload divergenceControl");
    std::vector<float> friction;
    for (Obstacle &o : m_obstacles) friction.push_back(o.friction());
    check(fs2d_set_obstacles(h, static_cast<int>(friction.size()), friction.data()), "fs2d_set_obstacles");
    std::vector<fs2d_source> src;
    for (Emitter &e : m_sources)
    {
        fs2d_source s;
        s.viscosity = e.viscosity();
        s.temperature = e.temperature();
        s.concentration = e.concentrartion();
        s.fuel = e.fuel();
        s.divergence = e.divergence();
        s.velocity_x = e.velocity().x();
        s.velocity_y = e.velocity().y();
        s.transfer_velocity = e.velocityTransfer() ? 1 : 0;
        src.push_back(s);
    }
    check(fs2d_set_sources(h, static_cast<int>(src.size()), src.data()), "fs2d_set_sources");
}

// firstFrameInit (flipsolver2d.cpp:788-795)
void FlipSolver::buildScene()
{
    updateSinks();
    updateSources();
    updateSolids();
    updateInitialFluid();
    seedInitialFluid();
}

void FlipSolver::firstFrameInit()
{
    prepareHost();
    uploadScene();
    uploadSeed();
    invalidateMirrors();
}

// ------------------------------------------------------------------ accessors (host copies)
void FlipSolver::fetchGrid(int grid, void *dst, size_t bytes) const
{
    if (!m_device || m_gridEpoch[grid] == m_mirrorEpoch) return;
    if (m_slabWorld > 1) check(fs2d_slab_gather_grid(m_device, grid), "fs2d_slab_gather_grid");  // collective
    check(fs2d_download_grid(m_device, grid, dst, bytes), "fs2d_download_grid");
    m_gridEpoch[grid] = m_mirrorEpoch;
}

size_t FlipSolver::particleCount() { return m_device ? static_cast<size_t>(fs2d_particle_count(m_device)) : m_seedPos.size() / 2; }

MarkerParticleSystem &FlipSolver::markerParticles()
{
    if (m_device && m_particleEpoch != m_mirrorEpoch)
    {
        const size_t n = static_cast<size_t>(fs2d_particle_count(m_device));
        const size_t k = m_markerParticles.propertyCount();
        std::vector<float> pos(2 * n), vel(2 * n), props(k * n);
        check(fs2d_download_particles(m_device, pos.data(), vel.data(), props.data()), "fs2d_download_particles");
        m_markerParticles.assign(n, pos.data(), vel.data(), props.data());
        m_particleEpoch = m_mirrorEpoch;
    }
    return m_markerParticles;
}

const MaterialGrid &FlipSolver::materialGrid() const
{
    fetchGrid(FS2D_GRID_MATERIAL, m_materialGrid.data().data(), linearSize());
    return m_materialGrid;
}

const StaggeredVelocityGrid &FlipSolver::fluidVelocityGrid() const
{
    Grid2d<float> &u = m_fluidVelocityGrid.velocityGridU(), &v = m_fluidVelocityGrid.velocityGridV();
    fetchGrid(FS2D_GRID_U, u.data().data(), u.data().size() * 4);
    fetchGrid(FS2D_GRID_V, v.data().data(), v.data().size() * 4);
    fetchGrid(FS2D_GRID_U_VALID, m_fluidVelocityGrid.uSampleValidityGrid().data().data(), u.data().size());
    fetchGrid(FS2D_GRID_V_VALID, m_fluidVelocityGrid.vSampleValidityGrid().data().data(), v.data().size());
    return m_fluidVelocityGrid;
}

const SdfGrid &FlipSolver::fluidSdf() const
{
    fetchGrid(FS2D_GRID_FLUID_SDF, m_fluidSdf.data().data(), linearSize() * 4);
    return m_fluidSdf;
}

const SdfGrid &FlipSolver::solidSdf() const
{
    fetchGrid(FS2D_GRID_SOLID_SDF, m_solidSdf.data().data(), linearSize() * 4);
    return m_solidSdf;
}

const Grid2d<float> &FlipSolver::testGrid() const
{
    fetchGrid(FS2D_GRID_TEST, m_testGrid.data().data(), linearSize() * 4);
    return m_testGrid;
}

const Grid2d<float> &FlipSolver::viscosityGrid() const
{
    fetchGrid(FS2D_GRID_VISCOSITY, m_viscosityGrid.data().data(), linearSize() * 4);
    return m_viscosityGrid;
}

const Grid2d<int> &FlipSolver::fluidParticleCounts() const
{
    fetchGrid(FS2D_GRID_COUNTS, m_fluidParticleCounts.data().data(), linearSize() * 4);
    return m_fluidParticleCounts;
}
