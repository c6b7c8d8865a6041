// Host mirror of FlipSolver2dLib/flipsolver2d.h: the same solver API (parameters, SolverStats,
// stepFrame(), const accessors, protected virtual stage hooks) on top of the fs2d C ABI
// (include/fs2d.h). All per-substep state lives on the GPU; this class owns the scene description,
// the frame / CFL substep loop (flipsolver2d.cpp:464-500), the std::mt19937 stream that seeds and
// reseeds particles (flipsolver2d.cpp:1013-1019) and lazily refreshed host copies of the grids and
// particle bins for the accessors.
#ifndef FS2D_HOST_FLIPSOLVER2D_H
#define FS2D_HOST_FLIPSOLVER2D_H

#include <array>
#include <chrono>
#include <memory>
#include <random>
#include <string>
#include <vector>

#include "../../include/fs2d.h"
#include "grid2d.h"
#include "markerparticlesystem.h"
#include "sceneobjects.h"

enum SimulationMethod : char { SIMULATION_LIQUID, SIMULATION_SMOKE, SIMULATION_FIRE, SIMULATION_NBFLIP };
enum ParameterHandlingMethod : char { PARTICLE, HYBRID, GRID };

// flipsolver2d.h:33-56
struct FlipSolverParameters
{
    double fluidDensity;
    unsigned int seed;
    double dx;
    int particlesPerCell;
    Vec3 globalAcceleration;
    float resolution;
    int fps;
    int maxSubsteps;
    float picRatio;
    float cflNumber;
    float particleScale;
    int pcgIterLimit;
    float domainSizeI;
    float domainSizeJ;
    int gridSizeI;
    int gridSizeJ;
    float sceneScale;
    bool viscosityEnabled;
    bool useHeavyViscosity;
    SimulationMethod simulationMethod;
    ParameterHandlingMethod parameterHandlingMethod;
};

// flipsolver2d.h:58-72
enum SolverStage
{
    ADVECTION = 0,
    DECOMPOSITION,
    DENSITY,
    PARTICLE_REBIN,
    PARTICLE_TO_GRID,
    GRID_UPDATE,
    AFTER_TRANSFER,
    PRESSURE,
    VISCOSITY,
    REPRESSURE,
    PARTICLE_UPDATE,
    PARTICLE_RESEED,
    SOLVER_STAGE_COUNT
};

// flipsolver2d.h:88-181. Stage time is host wall clock between endStage() calls; FlipSolver
// synchronises the device stream before each call, so a slot holds the GPU time of its stage plus
// the host work that belongs to it.
class SolverStats
{
public:
    using Clock = std::chrono::high_resolution_clock;
    using StageTimings = std::array<float, SOLVER_STAGE_COUNT>;

    SolverStats() { reset(); }
    int pressureIterations() const { return m_pressureIters; }
    int densityIterations() const { return m_densityIters; }
    int viscosityIterations() const { return m_viscosityIters; }
    void setPressureIterations(int v) { m_pressureIters = std::max(v, m_pressureIters); }
    void setDensityIters(int v) { m_densityIters = std::max(v, m_densityIters); }
    void setViscosityIterations(int v) { m_viscosityIters = std::max(v, m_viscosityIters); }
    void reset()
    {
        m_times.fill(0.f);
        m_last = m_frameStart = Clock::now();
        m_substeps = m_pressureIters = m_densityIters = m_viscosityIters = 0;
    }
    void endStage(SolverStage s)
    {
        const Clock::time_point now = Clock::now();
        m_times.at(s) += std::chrono::duration<float, std::milli>(now - m_last).count();
        m_last = Clock::now();
    }
    void endFrame() { m_total = std::chrono::duration<float, std::milli>(Clock::now() - m_frameStart).count(); }
    void addSubstep() { m_substeps++; }
    // checkpoint support (FlipSolver::loadState): the counters of the frame in progress
    void restore(const StageTimings &times, int substeps, int pressureIters, int densityIters, int viscosityIters)
    {
        m_times = times;
        m_substeps = substeps;
        m_pressureIters = pressureIters;
        m_densityIters = densityIters;
        m_viscosityIters = viscosityIters;
        m_last = Clock::now();
    }
    int substepCount() const { return m_substeps; }
    StageTimings timings() const { return m_times; }
    float frameTime() const { return m_total; }

private:
    StageTimings m_times;
    Clock::time_point m_last, m_frameStart;
    int m_substeps, m_pressureIters, m_densityIters, m_viscosityIters;
    float m_total = 0.f;
};

class FlipSolver : public LinearIndexable2d
{
public:
    explicit FlipSolver(const FlipSolverParameters *p);
    virtual ~FlipSolver();

    size_t particleCount();
    size_t cellCount() { return linearSize(); }
    int pcgIterationLimit() { return m_pcgIterLimit; }

    void stepFrame();
    // One CFL substep of the current frame (starts a new frame when the last one is complete);
    // returns true when this substep finished its frame. stepFrame() == loop until true.
    bool stepSubstep();
    // One substep whose particle state comes from and returns to a (pinned) host buffer in the sectioned layout of
    // fs2d_particle_stream_begin; the copies overlap the stages (include/fs2d.h). Returns what stepSubstep() returns.
    bool stepSubstepStreamed(void *hostBuf, int64_t capacityRecords, int64_t countIn, int64_t *countOut);
    // Frame-0 initialisation (scene rasterisation, seeding, upload) without stepping; stepFrame()
    // calls it when needed. Lets callers keep set-up out of a timed region.
    void prepare();
    // The host-only half of frame 0 (rasterise scene objects, draw the seed particles); needs no GPU.
    void prepareHost();
    size_t seedParticleCount() const { return m_seedPos.size() / 2; }
    const std::vector<float> &seedPositions() const { return m_seedPos; }
    const std::vector<float> &seedVelocities() const { return m_seedVel; }
    const std::vector<std::vector<float>> &seedProperties() const { return m_seedProps; }
    const Grid2d<int> &solidIdGrid() const { return m_solidId; }
    const Grid2d<int> &emitterIdGrid() const { return m_emitterId; }
    const Grid2d<float> &divergenceControlGrid() const { return m_divergenceControl; }

    void updateSolids();
    void updateSources();
    void updateSinks();
    void updateInitialFluid();

    size_t gridSizeI() { return m_sizeI; }
    size_t gridSizeJ() { return m_sizeJ; }

    void addGeometry(Obstacle &o) { m_obstacles.push_back(o); }
    void addSource(Emitter &e) { m_sources.push_back(e); }
    void addSink(Sink &s) { m_sinks.push_back(s); }
    void addInitialFluid(Emitter &e) { m_initialFluid.push_back(e); }

    int frameNumber() { return m_frameNumber; }
    std::vector<Obstacle> &geometryObjects() { return m_obstacles; }
    std::vector<Emitter> &sourceObjects() { return m_sources; }
    std::vector<Sink> &sinkObjects() { return m_sinks; }

    // Host copies, refreshed from the device when stale.
    MarkerParticleSystem &markerParticles();
    const MaterialGrid &materialGrid() const;
    const StaggeredVelocityGrid &fluidVelocityGrid() const;
    const SdfGrid &fluidSdf() const;
    const SdfGrid &solidSdf() const;
    const Grid2d<float> &testGrid() const;
    const Grid2d<float> &viscosityGrid() const;
    const Grid2d<int> &fluidParticleCounts() const;
    const SolverStats &timeStats() const { return m_stats; }

    float stepDt() const { return m_stepDt; }
    double dx() const { return m_dx; }
    SimulationMethod simulationMethod() const { return m_simulationMethod; }
    float domainSizeI() const { return m_domainSizeI; }
    float domainSizeJ() const { return m_domainSizeJ; }
    double fluidDensity() const { return m_fluidDensity; }
    int particlesPerCell() const { return m_particlesPerCell; }
    float picRatio() const { return m_picRatio; }
    float cflNumber() const { return m_cflNumber; }
    int maxSubsteps() const { return m_maxSubsteps; }
    int fps() const { return m_fps; }
    float frameDt() const { return m_frameDt; }
    Vec3 globalAcceleration() const { return m_globalAcceleration; }
    float sceneScale() const { return m_sceneScale; }
    float lastFrameTime() const { return m_stats.frameTime(); }
    size_t testValuePropertyIndex() { return m_testValuePropertyIndex; }

    virtual void initAdditionalParameters();

    // ---- additions of this implementation
    fs2d_handle device();                       // the C-ABI handle (created on first use)
    static void setQuiet(bool q);               // silence the per-substep stdout lines (flipsolver2d.cpp:491)
    static void setDevice(int ordinal);         // CUDA device for solvers created afterwards (env FS2D_DEVICE)
    // Convergence test of the PCG: T > 0 reproduces the reference's thread-count dependent test for a
    // pool of T threads (vmath.cpp:100-136), 0 = true max-norm (env FS2D_CONVERGENCE_THREADS).
    static void setConvergenceThreads(int t);
    int64_t kernelLaunches();
    // State dump / restore, doubling as checkpoint (the reference has no serialisation). saveState writes
    // the device state (fs2d_state_save: all grids, particle records) plus the host side of the stepping loop: frame
    // and substep counters, substep time, the mt19937 stream, the stage counters of the frame in progress. loadState on
    // a solver loaded from the SAME scene continues bit-identically to the uninterrupted run. Single GPU only.
    void saveState(const std::string &path);
    void loadState(const std::string &path);
    // Row slabs over several GPUs (include/fs2d.h "row slabs"; the reference's ThreadPool splits the same loops over
    // row ranges, threadpool.cpp:41-76): one process -- one solver -- per GPU, every rank loads the SAME scene.
    // setSlab before the solver touches the device; then exchange the 256-byte blobs (slabExport -> every other
    // rank's slabConnect, by any transport) before prepare()/stepFrame(). In slab mode stepFrame(), the grid
    // accessors (which gather all rows) and globalParticleCount() are collective: every rank must call them.
    // markerParticles() / particleCount() return the rank's own particles.
    static void setSlab(int rank, int world, int deviceShare = 1);
    void slabExport(void *blob /* FS2D_SLAB_HANDLE_BYTES */);
    void slabConnect(int peerRank, const void *blob);
    int slabRank() const { return m_slabRank; }
    int slabWorld() const { return m_slabWorld; }
    size_t globalParticleCount();
    std::vector<int32_t> slabBounds(int world);  // row boundaries balanced by seed particles per tile row
    std::vector<int32_t> slabBoundsFromMaterial(int world) const;
    void seedRows(int &rowLo, int &rowHi) const;

protected:
    virtual fs2d_params deviceParameters() const;
    void check(int rc, const char *what) const;
    void endStage(SolverStage s);
    void invalidateMirrors() { m_mirrorEpoch++; }

    virtual void firstFrameInit();
    virtual void buildScene();
    virtual void uploadScene();
    virtual void seedInitialFluid();
    void uploadSeed();

    virtual void step();
    virtual void advect();
    virtual void buildPressureSystem();          // getPressureProjectionMatrix + getIPPCoefficients
    void pruneParticles();                       // folded into rebinParticles on the device
    void rebinParticles();
    void densityCorrection();
    virtual void gridUpdate();
    virtual void particleToGrid();
    virtual void updateSdf();
    virtual void updateMaterials();
    virtual void afterTransfer();
    void extrapolateLevelsetInside();
    void extrapolateLevelsetOutside();
    void extrapolateVelocity(int radius);
    void saveVelocity();
    virtual void applyBodyForces();
    virtual void project();
    void updateVelocityFromSolids();
    void applyViscosity();
    virtual void particleUpdate();
    virtual void countParticles();
    virtual void reseedParticles();
    float maxParticleVelocity();
    Vec3 jitteredPosInCell(size_t i, size_t j);

    void fetchGrid(int grid, void *dst, size_t bytes) const;

    int m_frameNumber = 0;
    bool m_prepared = false;
    bool m_sceneBuilt = false;
    bool m_inFrame = false;
    void *m_streamBuf = nullptr;       // set during stepSubstepStreamed: where the particle state of this substep goes
    int64_t m_streamCapacity = 0;
    float m_substepTime = 0.f;
    int m_substepCount = 0;
    std::mt19937 m_randEngine;
    std::vector<Obstacle> m_obstacles;
    std::vector<Emitter> m_sources;
    std::vector<Sink> m_sinks;
    std::vector<Emitter> m_initialFluid;

    // host copies (authoritative only during frame-0 initialisation)
    mutable MarkerParticleSystem m_markerParticles;
    mutable StaggeredVelocityGrid m_fluidVelocityGrid;
    mutable MaterialGrid m_materialGrid;
    mutable SdfGrid m_solidSdf;
    mutable SdfGrid m_fluidSdf;
    mutable Grid2d<float> m_viscosityGrid;
    mutable Grid2d<int> m_emitterId;
    mutable Grid2d<int> m_solidId;
    mutable Grid2d<int> m_fluidParticleCounts;
    mutable Grid2d<float> m_divergenceControl;
    mutable Grid2d<float> m_testGrid;
    mutable uint64_t m_mirrorEpoch = 1;
    mutable uint64_t m_gridEpoch[32] = {0};
    mutable uint64_t m_particleEpoch = 0;

    // frame-0 seed buffers
    std::vector<float> m_seedPos, m_seedVel;
    std::vector<std::vector<float>> m_seedProps;

    float m_stepDt, m_frameDt;
    double m_dx, m_fluidDensity;
    unsigned int m_seed;
    int m_particlesPerCell;
    Vec3 m_globalAcceleration;
    float m_resolution;
    int m_fps, m_maxSubsteps;
    float m_picRatio, m_cflNumber, m_particleScale;
    int m_pcgIterLimit;
    float m_domainSizeI, m_domainSizeJ, m_sceneScale;
    double m_projectTolerance;
    bool m_viscosityEnabled;
    bool m_useHeavyViscosity = false;
    SimulationMethod m_simulationMethod;
    ParameterHandlingMethod m_parameterHandlingMethod;
    SolverStats m_stats;

    size_t m_testValuePropertyIndex = 0;
    size_t m_viscosityPropertyIndex = static_cast<size_t>(-1);

    fs2d_handle m_device = nullptr;
    int m_slabRank = 0, m_slabWorld = 1;
};

#endif
