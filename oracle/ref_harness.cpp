// TEST INFRASTRUCTURE ONLY -- never linked into, imported by or executed from the
// product path (flipsolver2d_b200/). See oracle/ref_api.h.
//
// Harness around the UNMODIFIED reference sources (compiled from where they lie under
// /root/reference by oracle/build_ref.py). It subclasses the reference's solver
// classes to reach their `protected` stages/state (flipsolver2d.h:280-431) and
// JsonSceneReader to reach its `protected static` helpers (Utils/jsonscenereader.h:18-33),
// and exports a small C ABI so tests can drive one stage at a time and read state.
//
// The ThreadPool size is `std::thread::hardware_concurrency()`
// (threading/threadpool.cpp:12) and several results depend on it (reduction
// grouping vmath.cpp:28-44, convergence test vmath.cpp:100-136). To pin it without
// touching the reference, this file defines hardware_concurrency() itself; the
// library is linked with -Bsymbolic-functions so the reference's call binds here.
// FS2D_ORACLE_THREADS=<T> selects T (default: online CPUs).

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <memory>
#include <thread>
#include <type_traits>
#include <unistd.h>

#include "jsonscenereader.h"
#include "solvers.h"
#include "vmath.h"

#include "ref_api.h"

unsigned int std::thread::hardware_concurrency() noexcept
{
    const char *env = std::getenv("FS2D_ORACLE_THREADS");
    if (env != nullptr)
    {
        int t = std::atoi(env);
        if (t > 0) return static_cast<unsigned int>(t);
    }
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? static_cast<unsigned int>(n) : 1u;
}

namespace
{
struct IOracle
{
    virtual ~IOracle() = default;
    virtual FlipSolver *solver() = 0;
    virtual void runStage(int stage) = 0;
    virtual void setStepDt(float dt) = 0;
    virtual float maxVelocity() = 0;
    virtual void bumpFrame() = 0;
    virtual void getParams(double *out) = 0;
    virtual int propertyCount() = 0;
    virtual int64_t gridSize(int grid) = 0;
    virtual int getGrid(int grid, void *out) = 0;
    virtual int setGrid(int grid, const void *in) = 0;
    virtual double getMatrix(uint8_t *isUnit, uint8_t *mask, uint8_t *count, double *coef) = 0;
    virtual void spmv(const double *in, double *out) = 0;
    virtual void precond(const double *in, double *out) = 0;
    virtual int pcg(const double *rhs, double *x, int iterLimit, double tol) = 0;
    virtual void pressureRhs(double *rhs) = 0;
    virtual void densityRhs(double *rhs) = 0;
    virtual void applyPressure(const double *p) = 0;
};

template <class T> void copyOut(const std::vector<T> &v, void *out)
{
    std::memcpy(out, v.data(), v.size() * sizeof(T));
}

inline void copyOutBool(const std::vector<bool> &v, void *out)
{
    uint8_t *o = static_cast<uint8_t *>(out);
    for (size_t i = 0; i < v.size(); i++) o[i] = v[i] ? 1 : 0;
}

template <class T> void copyIn(std::vector<T> &v, const void *in)
{
    std::memcpy(v.data(), in, v.size() * sizeof(T));
}

inline void copyInBool(std::vector<bool> &v, const void *in)
{
    const uint8_t *p = static_cast<const uint8_t *>(in);
    for (size_t i = 0; i < v.size(); i++) v[i] = p[i] != 0;
}

template <class Base> class Exposed : public Base, public IOracle
{
public:
    template <class P> explicit Exposed(const P *p) : Base(p) {}

    FlipSolver *solver() override { return this; }

    void setStepDt(float dt) override { this->m_stepDt = dt; }
    float maxVelocity() override { return this->maxParticleVelocity(); }
    void bumpFrame() override { this->m_frameNumber++; }
    int propertyCount() override { return static_cast<int>(this->m_markerParticles.bins().data()[0].properties().size()); }

    void getParams(double *out) override
    {
        out[0] = this->m_stepDt;
        out[1] = this->m_frameDt;
        out[2] = this->m_dx;
        out[3] = this->m_fluidDensity;
        out[4] = this->m_particlesPerCell;
        out[5] = this->m_globalAcceleration.x();
        out[6] = this->m_globalAcceleration.y();
        out[7] = this->m_picRatio;
        out[8] = this->m_cflNumber;
        out[9] = this->m_particleScale;
        out[10] = this->m_pcgIterLimit;
        out[11] = this->m_projectTolerance;
        out[12] = this->m_maxSubsteps;
        out[13] = this->m_viscosityEnabled ? 1.0 : 0.0;
        out[14] = static_cast<int>(this->m_parameterHandlingMethod);
        out[15] = this->m_fps;
    }

    void runStage(int stage) override
    {
        switch (stage)
        {
        case REF_STAGE_ADVECT: this->advect(); break;
        case REF_STAGE_BUILD_MATRIX:
            this->m_pressureMatrix = this->getPressureProjectionMatrix();
            this->m_pressurePrecond = this->getIPPCoefficients(this->m_pressureMatrix);
            break;
        case REF_STAGE_PRUNE_REBIN:
            this->pruneParticles();
            this->m_markerParticles.rebinParticles();
            break;
        case REF_STAGE_DENSITY_CORRECTION: this->densityCorrection(); break;
        case REF_STAGE_P2G: this->particleToGrid(); break;
        case REF_STAGE_UPDATE_SDF: this->updateSdf(); break;
        case REF_STAGE_UPDATE_MATERIALS: this->updateMaterials(); break;
        case REF_STAGE_AFTER_TRANSFER: this->afterTransfer(); break;
        case REF_STAGE_EXTRAPOLATE_SDF_IN: this->extrapolateLevelsetInside(this->m_fluidSdf); break;
        case REF_STAGE_EXTRAPOLATE_SDF_OUT: this->extrapolateLevelsetOutside(this->m_fluidSdf); break;
        case REF_STAGE_EXTRAPOLATE_VEL: this->m_fluidVelocityGrid.extrapolate(10); break;
        case REF_STAGE_SAVE_VELOCITY: this->m_savedFluidVelocityGrid = this->m_fluidVelocityGrid; break;
        case REF_STAGE_BODY_FORCES: this->applyBodyForces(); break;
        case REF_STAGE_PROJECT: this->project(); break;
        case REF_STAGE_VELOCITY_FROM_SOLIDS: this->updateVelocityFromSolids(); break;
        case REF_STAGE_VISCOSITY: this->applyViscosity(); break;
        case REF_STAGE_PARTICLE_UPDATE: this->particleUpdate(); break;
        case REF_STAGE_COUNT_PARTICLES: this->countParticles(); break;
        case REF_STAGE_RESEED: this->reseedParticles(); break;
        case REF_STAGE_GRID_UPDATE: this->gridUpdate(); break;
        case REF_STAGE_FULL_STEP: this->step(); break;
        case REF_STAGE_UPDATE_DENSITY_GRID: this->updateDensityGrid(); break;
        case REF_STAGE_FIRST_FRAME_INIT: this->firstFrameInit(); break;
        default: std::fprintf(stderr, "ref_run_stage: unknown stage %d\n", stage); break;
        }
    }

    int64_t gridSize(int grid) override
    {
        const int64_t I = this->m_sizeI, J = this->m_sizeJ;
        switch (grid)
        {
        case REF_GRID_U:
        case REF_GRID_U_VALID:
        case REF_GRID_SAVED_U: return (I + 1) * J;
        case REF_GRID_V:
        case REF_GRID_V_VALID:
        case REF_GRID_SAVED_V: return I * (J + 1);
        case REF_GRID_TEMPERATURE:
        case REF_GRID_CONCENTRATION: return std::is_base_of<FlipSmokeSolver, Base>::value ? I * J : 0;
        case REF_GRID_FUEL: return std::is_base_of<FlipFireSolver, Base>::value ? I * J : 0;
        default: return I * J;
        }
    }

    int getGrid(int grid, void *out) override
    {
        switch (grid)
        {
        case REF_GRID_U: copyOut(this->m_fluidVelocityGrid.velocityGridU().data(), out); return 0;
        case REF_GRID_V: copyOut(this->m_fluidVelocityGrid.velocityGridV().data(), out); return 0;
        case REF_GRID_U_VALID: copyOutBool(this->m_fluidVelocityGrid.uSampleValidityGrid().data(), out); return 0;
        case REF_GRID_V_VALID: copyOutBool(this->m_fluidVelocityGrid.vSampleValidityGrid().data(), out); return 0;
        case REF_GRID_SAVED_U: copyOut(this->m_savedFluidVelocityGrid.velocityGridU().data(), out); return 0;
        case REF_GRID_SAVED_V: copyOut(this->m_savedFluidVelocityGrid.velocityGridV().data(), out); return 0;
        case REF_GRID_MATERIAL: copyOut(this->m_materialGrid.data(), out); return 0;
        case REF_GRID_FLUID_SDF: copyOut(this->m_fluidSdf.data(), out); return 0;
        case REF_GRID_SOLID_SDF: copyOut(this->m_solidSdf.data(), out); return 0;
        case REF_GRID_VISCOSITY: copyOut(this->m_viscosityGrid.data(), out); return 0;
        case REF_GRID_DENSITY: copyOut(this->m_densityGrid.data(), out); return 0;
        case REF_GRID_COUNTS: copyOut(this->m_fluidParticleCounts.data(), out); return 0;
        case REF_GRID_EMITTER_ID: copyOut(this->m_emitterId.data(), out); return 0;
        case REF_GRID_SOLID_ID: copyOut(this->m_solidId.data(), out); return 0;
        case REF_GRID_DIVERGENCE_CONTROL: copyOut(this->m_divergenceControl.data(), out); return 0;
        case REF_GRID_TEST: copyOut(this->m_testGrid.data(), out); return 0;
        case REF_GRID_KNOWN_CENTERED: copyOutBool(this->m_knownCenteredParams.data(), out); return 0;
        default: break;
        }
        if constexpr (std::is_base_of<FlipSmokeSolver, Base>::value)
        {
            if (grid == REF_GRID_TEMPERATURE) { copyOut(this->m_temperature.data(), out); return 0; }
            if (grid == REF_GRID_CONCENTRATION) { copyOut(this->m_smokeConcentration.data(), out); return 0; }
        }
        if constexpr (std::is_base_of<FlipFireSolver, Base>::value)
        {
            if (grid == REF_GRID_FUEL) { copyOut(this->m_fuel.data(), out); return 0; }
        }
        return -1;
    }

    int setGrid(int grid, const void *in) override
    {
        switch (grid)
        {
        case REF_GRID_U: copyIn(this->m_fluidVelocityGrid.velocityGridU().data(), in); return 0;
        case REF_GRID_V: copyIn(this->m_fluidVelocityGrid.velocityGridV().data(), in); return 0;
        case REF_GRID_U_VALID: copyInBool(this->m_fluidVelocityGrid.uSampleValidityGrid().data(), in); return 0;
        case REF_GRID_V_VALID: copyInBool(this->m_fluidVelocityGrid.vSampleValidityGrid().data(), in); return 0;
        case REF_GRID_SAVED_U: copyIn(this->m_savedFluidVelocityGrid.velocityGridU().data(), in); return 0;
        case REF_GRID_SAVED_V: copyIn(this->m_savedFluidVelocityGrid.velocityGridV().data(), in); return 0;
        case REF_GRID_MATERIAL: copyIn(this->m_materialGrid.data(), in); return 0;
        case REF_GRID_FLUID_SDF: copyIn(this->m_fluidSdf.data(), in); return 0;
        case REF_GRID_SOLID_SDF: copyIn(this->m_solidSdf.data(), in); return 0;
        case REF_GRID_VISCOSITY: copyIn(this->m_viscosityGrid.data(), in); return 0;
        case REF_GRID_DENSITY: copyIn(this->m_densityGrid.data(), in); return 0;
        case REF_GRID_COUNTS: copyIn(this->m_fluidParticleCounts.data(), in); return 0;
        case REF_GRID_EMITTER_ID: copyIn(this->m_emitterId.data(), in); return 0;
        case REF_GRID_SOLID_ID: copyIn(this->m_solidId.data(), in); return 0;
        case REF_GRID_DIVERGENCE_CONTROL: copyIn(this->m_divergenceControl.data(), in); return 0;
        case REF_GRID_TEST: copyIn(this->m_testGrid.data(), in); return 0;
        case REF_GRID_KNOWN_CENTERED: copyInBool(this->m_knownCenteredParams.data(), in); return 0;
        default: break;
        }
        if constexpr (std::is_base_of<FlipSmokeSolver, Base>::value)
        {
            if (grid == REF_GRID_TEMPERATURE) { copyIn(this->m_temperature.data(), in); return 0; }
            if (grid == REF_GRID_CONCENTRATION) { copyIn(this->m_smokeConcentration.data(), in); return 0; }
        }
        if constexpr (std::is_base_of<FlipFireSolver, Base>::value)
        {
            if (grid == REF_GRID_FUEL) { copyIn(this->m_fuel.data(), in); return 0; }
        }
        return -1;
    }

    double getMatrix(uint8_t *isUnit, uint8_t *mask, uint8_t *count, double *coef) override
    {
        const size_t n = this->linearSize();
        if (isUnit) std::memset(isUnit, 0, n);
        if (mask) std::memset(mask, 0, n);
        if (count) std::memset(count, 0, n);
        if (coef) std::memset(coef, 0, 4 * n * sizeof(double));
        for (const IndexedPressureParameterUnit &u : this->m_pressureMatrix.data())
        {
            if (isUnit) isUnit[u.unitIndex] = 1;
            if (mask) mask[u.unitIndex] = u.fluidNeighborMask;
            if (count) count[u.unitIndex] = u.nonsolidNeighborCount;
        }
        if (coef)
        {
            for (const IndexedIPPCoefficientUnit &u : this->m_pressurePrecond.data())
            {
                coef[0 * n + u.unitIndex] = u.iNeg;
                coef[1 * n + u.unitIndex] = u.iPos;
                coef[2 * n + u.unitIndex] = u.jNeg;
                coef[3 * n + u.unitIndex] = u.jPos;
            }
        }
        return this->m_stepDt / (this->m_fluidDensity * this->m_dx * this->m_dx);
    }

    void spmv(const double *in, double *out) override
    {
        const size_t n = this->linearSize();
        std::vector<double> vin(in, in + n), vout(n, 0.0);
        this->m_pressureMatrix.multiply(vin, vout);
        std::memcpy(out, vout.data(), n * sizeof(double));
    }

    void precond(const double *in, double *out) override
    {
        const size_t n = this->linearSize();
        std::vector<double> vin(in, in + n), vout(n, 0.0);
        this->m_pressurePrecond.multiply(vin, vout);
        std::memcpy(out, vout.data(), n * sizeof(double));
    }

    int pcg(const double *rhs, double *x, int iterLimit, double tol) override
    {
        const size_t n = this->linearSize();
        std::vector<double> vrhs(rhs, rhs + n), vx(n, 0.0);
        int iters = this->m_pressureSolver.solve(this->m_pressureMatrix, this->m_pressurePrecond, vx, vrhs, iterLimit, tol);
        std::memcpy(x, vx.data(), n * sizeof(double));
        return iters;
    }

    void pressureRhs(double *rhs) override
    {
        std::vector<double> v(this->linearSize(), 0.0);
        this->calcPressureRhs(v);
        std::memcpy(rhs, v.data(), v.size() * sizeof(double));
    }

    void densityRhs(double *rhs) override
    {
        std::vector<double> v(this->linearSize(), 0.0);
        this->calcDensityCorrectionRhs(v);
        std::memcpy(rhs, v.data(), v.size() * sizeof(double));
    }

    void applyPressure(const double *p) override
    {
        std::vector<double> v(p, p + this->linearSize());
        this->applyPressuresToVelocityField(v);
    }
};

// Same dispatch as JsonSceneReader::loadJson (Utils/jsonscenereader.cpp:8-77) but
// instantiating the exposing subclasses.
class OracleReader : public JsonSceneReader
{
public:
    static IOracle *load(const std::string &fileName)
    {
        using json = nlohmann::json;
        IOracle *out = nullptr;
        try
        {
            std::ifstream sceneFile(fileName);
            if (!sceneFile.is_open())
            {
                std::cerr << "ref_load_scene: cannot open " << fileName << "\n";
                return nullptr;
            }
            json sceneJson;
            sceneFile >> sceneJson;
            json settingsJson = sceneJson["settings"];
            SimulationMethod method = simMethodFromName(settingsJson["simType"].get<std::string>());
            std::shared_ptr<FlipSolver> alias;
            switch (method)
            {
            case SIMULATION_LIQUID:
            {
                FlipSolverParameters p;
                populateFlipSolverParamsFromJson(&p, settingsJson);
                auto *s = new Exposed<FlipSolver>(&p);
                out = s;
                alias.reset(static_cast<FlipSolver *>(s), [](FlipSolver *) {});
            }
            break;
            case SIMULATION_SMOKE:
            {
                SmokeSolverParameters p;
                populateFlipSolverParamsFromJson(&p, settingsJson);
                populateSmokeSolverParamsFromJson(&p, settingsJson);
                auto *s = new Exposed<FlipSmokeSolver>(&p);
                out = s;
                alias.reset(static_cast<FlipSolver *>(s), [](FlipSolver *) {});
            }
            break;
            case SIMULATION_FIRE:
            {
                FireSolverParameters p;
                populateFlipSolverParamsFromJson(&p, settingsJson);
                populateFireSolverParamsFromJson(&p, settingsJson);
                auto *s = new Exposed<FlipFireSolver>(&p);
                out = s;
                alias.reset(static_cast<FlipSolver *>(s), [](FlipSolver *) {});
            }
            break;
            case SIMULATION_NBFLIP:
            {
                NBFlipParameters p;
                populateFlipSolverParamsFromJson(&p, settingsJson);
                populateNBFlipSolverParamsFromJson(&p, settingsJson);
                auto *s = new Exposed<NBFlipSolver>(&p);
                out = s;
                alias.reset(static_cast<FlipSolver *>(s), [](FlipSolver *) {});
            }
            break;
            }
            out->solver()->initAdditionalParameters();
            objectsFromJson(sceneJson["solver"], alias);
        }
        catch (std::exception &e)
        {
            std::cerr << "ref_load_scene: " << e.what() << "\n";
            delete out;
            return nullptr;
        }
        return out;
    }
};

IOracle *O(ref_handle h) { return static_cast<IOracle *>(h); }

std::streambuf *g_savedCout = nullptr;
std::ofstream g_devNull;
}  // namespace

extern "C" {

ref_handle ref_load_scene(const char *json_path) { return OracleReader::load(json_path); }

void ref_destroy(ref_handle h) { delete O(h); }

int ref_thread_count(void) { return static_cast<int>(ThreadPool::i()->threadCount()); }

void ref_set_quiet(int quiet)
{
    if (quiet && g_savedCout == nullptr)
    {
        g_devNull.open("/dev/null");
        g_savedCout = std::cout.rdbuf(g_devNull.rdbuf());
    }
    else if (!quiet && g_savedCout != nullptr)
    {
        std::cout.rdbuf(g_savedCout);
        g_savedCout = nullptr;
        g_devNull.close();
    }
}

int ref_size_i(ref_handle h) { return static_cast<int>(O(h)->solver()->gridSizeI()); }
int ref_size_j(ref_handle h) { return static_cast<int>(O(h)->solver()->gridSizeJ()); }
int ref_sim_type(ref_handle h) { return static_cast<int>(O(h)->solver()->simulationMethod()); }
int64_t ref_particle_count(ref_handle h) { return static_cast<int64_t>(O(h)->solver()->particleCount()); }
int ref_property_count(ref_handle h) { return O(h)->propertyCount(); }
int ref_frame_number(ref_handle h) { return O(h)->solver()->frameNumber(); }
void ref_get_params(ref_handle h, double *out16) { O(h)->getParams(out16); }

void ref_step_frame(ref_handle h) { O(h)->solver()->stepFrame(); }

void ref_get_stats(ref_handle h, float *timings12, float *misc5)
{
    const SolverStats &s = O(h)->solver()->timeStats();
    SolverStats::StageTimings t = s.timings();
    for (int i = 0; i < SOLVER_STAGE_COUNT; i++) timings12[i] = t[static_cast<size_t>(i)];
    misc5[0] = s.frameTime();
    misc5[1] = static_cast<float>(s.substepCount());
    misc5[2] = static_cast<float>(s.pressureIterations());
    misc5[3] = static_cast<float>(s.densityIterations());
    misc5[4] = static_cast<float>(s.viscosityIterations());
}

void ref_set_step_dt(ref_handle h, float dt) { O(h)->setStepDt(dt); }
float ref_max_particle_velocity(ref_handle h) { return O(h)->maxVelocity(); }
void ref_run_stage(ref_handle h, int stage) { O(h)->runStage(stage); }
void ref_bump_frame_number(ref_handle h) { O(h)->bumpFrame(); }

void ref_get_particles(ref_handle h, float *pos, float *vel, float *props, int32_t *bin_of)
{
    MarkerParticleSystem &ps = O(h)->solver()->markerParticles();
    const int64_t total = static_cast<int64_t>(ps.particleCount());
    const int k = O(h)->propertyCount();
    int64_t at = 0;
    std::vector<ParticleBin> &bins = ps.bins().data();
    for (size_t b = 0; b < bins.size(); b++)
    {
        ParticleBin &bin = bins[b];
        for (size_t p = 0; p < bin.size(); p++, at++)
        {
            if (pos)
            {
                pos[2 * at] = bin.particlePosition(p).x();
                pos[2 * at + 1] = bin.particlePosition(p).y();
            }
            if (vel)
            {
                vel[2 * at] = bin.particleVelocity(p).x();
                vel[2 * at + 1] = bin.particleVelocity(p).y();
            }
            if (props)
                for (int c = 0; c < k; c++) props[c * total + at] = bin.particleProperties<float>(static_cast<size_t>(c))[p];
            if (bin_of) bin_of[at] = static_cast<int32_t>(b);
        }
    }
}

void ref_set_particles(ref_handle h, int64_t count, const float *pos, const float *vel, const float *props)
{
    MarkerParticleSystem &ps = O(h)->solver()->markerParticles();
    const int k = O(h)->propertyCount();
    for (ParticleBin &bin : ps.bins().data()) bin.clear();
    for (int64_t i = 0; i < count; i++)
    {
        Vec3 p(pos[2 * i], pos[2 * i + 1]);
        Vec3 v(vel ? vel[2 * i] : 0.f, vel ? vel[2 * i + 1] : 0.f);
        ParticleBin &bin = ps.binForGridPosition(p);
        size_t idx = bin.addMarkerParticle(p, v);
        if (props)
            for (int c = 0; c < k; c++) bin.particleProperties<float>(static_cast<size_t>(c))[idx] = props[c * count + i];
    }
}

int64_t ref_grid_size(ref_handle h, int grid) { return O(h)->gridSize(grid); }
int ref_get_grid(ref_handle h, int grid, void *out) { return O(h)->getGrid(grid, out); }
int ref_set_grid(ref_handle h, int grid, const void *in) { return O(h)->setGrid(grid, in); }

double ref_get_matrix(ref_handle h, uint8_t *is_unit, uint8_t *mask, uint8_t *count, double *coef)
{
    return O(h)->getMatrix(is_unit, mask, count, coef);
}
void ref_spmv(ref_handle h, const double *in, double *out) { O(h)->spmv(in, out); }
void ref_precond_apply(ref_handle h, const double *in, double *out) { O(h)->precond(in, out); }
int ref_pcg_solve(ref_handle h, const double *rhs, double *x, int iter_limit, double tol)
{
    return O(h)->pcg(rhs, x, iter_limit, tol);
}
void ref_pressure_rhs(ref_handle h, double *rhs) { O(h)->pressureRhs(rhs); }
void ref_density_rhs(ref_handle h, double *rhs) { O(h)->densityRhs(rhs); }
void ref_apply_pressure(ref_handle h, const double *p) { O(h)->applyPressure(p); }

double ref_vops_dot(const double *a, const double *b, int64_t n)
{
    std::vector<double> va(a, a + n), vb(b, b + n);
    return VOps::i().dot(va, vb);
}

double ref_vops_max_abs(const double *a, int64_t n)
{
    std::vector<double> va(a, a + n);
    return VOps::i().maxAbs(va);
}

}  // extern "C"
