// C shim over the host classes so that non-C++ callers (bench.py, the tests) drive the same
// JsonSceneReader / FlipSolver objects a C++ application links against. Declared in include/fs2d_host.h.
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>

#include "../../include/fs2d_host.h"
#include "benchruntable.h"
#include "jsonscenereader.h"

namespace
{
struct Holder
{
    std::shared_ptr<FlipSolver> solver;
    std::string error;
};

template <class F> int guarded(Holder *h, F f)
{
    if (!h || !h->solver) return FS2D_ERR_ARG;
    try
    {
        f();
        return FS2D_OK;
    }
    catch (std::exception &e)
    {
        h->error = e.what();
        std::cerr << "fs2d host: " << e.what() << std::endl;
        return FS2D_ERR_STATE;
    }
}
}  // namespace

extern "C" {

void fs2dh_set_quiet(int quiet) { FlipSolver::setQuiet(quiet != 0); }
void fs2dh_set_device(int ordinal) { FlipSolver::setDevice(ordinal); }
void fs2dh_set_convergence_threads(int threads) { FlipSolver::setConvergenceThreads(threads); }

void fs2dh_set_slab(int rank, int world, int device_share) { FlipSolver::setSlab(rank, world, device_share); }

int fs2dh_slab_export(fs2dh_solver s, void *blob)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() { h->solver->slabExport(blob); });
}

int fs2dh_slab_connect(fs2dh_solver s, int peer_rank, const void *blob)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() { h->solver->slabConnect(peer_rank, blob); });
}

int fs2dh_write_stats_xlsx(const char *path, int scenes, const char *const *scene_names, const int *rows_per_scene, const double *rows18)
{
    if (!path || scenes < 0 || (scenes > 0 && (!scene_names || !rows_per_scene || !rows18))) return FS2D_ERR_ARG;
    BenchRunTable table;
    table.setOutputFile(path);
    const double *r = rows18;
    for (int k = 0; k < scenes; k++)
    {
        for (int f = 0; f < rows_per_scene[k]; f++, r += BenchRunTable::TABLE_COLUMN_COUNT)
        {
            BenchRunTable::Row row{};
            for (int c = 0; c < BenchRunTable::TABLE_COLUMN_COUNT; c++) row[static_cast<size_t>(c)] = r[c];
            table.addRow(row);
        }
        table.finishScene(scene_names[k]);
    }
    return table.save() ? FS2D_OK : FS2D_ERR_STATE;
}

int fs2dh_slab_bounds(fs2dh_solver s, int world, int32_t *row_bounds)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() {
        const std::vector<int32_t> b = h->solver->slabBounds(world);
        for (size_t k = 0; k < b.size(); k++) row_bounds[k] = b[k];
    });
}

int64_t fs2dh_global_particle_count(fs2dh_solver s)
{
    Holder *h = static_cast<Holder *>(s);
    int64_t n = -1;
    guarded(h, [&]() { n = static_cast<int64_t>(h->solver->globalParticleCount()); });
    return n;
}

fs2dh_solver fs2dh_load_scene(const char *json_path)
{
    if (!json_path) return nullptr;
    std::shared_ptr<FlipSolver> s = JsonSceneReader::loadJson(json_path);
    if (!s) return nullptr;
    Holder *h = new Holder();
    h->solver = s;
    return h;
}

void fs2dh_destroy(fs2dh_solver s) { delete static_cast<Holder *>(s); }
const char *fs2dh_last_error(fs2dh_solver s) { return s ? static_cast<Holder *>(s)->error.c_str() : "null solver"; }

int fs2dh_prepare_host(fs2dh_solver s)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() { h->solver->prepareHost(); });
}

int64_t fs2dh_seed_count(fs2dh_solver s) { return s ? static_cast<int64_t>(static_cast<Holder *>(s)->solver->seedParticleCount()) : 0; }

int fs2dh_seed_particles(fs2dh_solver s, float *pos, float *vel, float *props)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() {
        const size_t n = h->solver->seedParticleCount();
        if (pos) std::memcpy(pos, h->solver->seedPositions().data(), n * 2 * sizeof(float));
        if (vel) std::memcpy(vel, h->solver->seedVelocities().data(), n * 2 * sizeof(float));
        if (props)
        {
            size_t c = 0;
            for (const std::vector<float> &col : h->solver->seedProperties()) std::memcpy(props + (c++) * n, col.data(), n * sizeof(float));
        }
    });
}

int fs2dh_host_grid(fs2dh_solver s, int grid, void *out)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() {
        FlipSolver &f = *h->solver;
        const size_t N = f.linearSize();
        switch (grid)
        {
        case FS2D_GRID_MATERIAL: std::memcpy(out, f.materialGrid().data().data(), N); break;
        case FS2D_GRID_SOLID_SDF: std::memcpy(out, f.solidSdf().data().data(), N * 4); break;
        case FS2D_GRID_FLUID_SDF: std::memcpy(out, f.fluidSdf().data().data(), N * 4); break;
        case FS2D_GRID_VISCOSITY: std::memcpy(out, f.viscosityGrid().data().data(), N * 4); break;
        case FS2D_GRID_SOLID_ID: std::memcpy(out, f.solidIdGrid().data().data(), N * 4); break;
        case FS2D_GRID_EMITTER_ID: std::memcpy(out, f.emitterIdGrid().data().data(), N * 4); break;
        case FS2D_GRID_DIVERGENCE_CONTROL: std::memcpy(out, f.divergenceControlGrid().data().data(), N * 4); break;
        default: throw std::runtime_error("fs2dh_host_grid: grid has no host copy");
        }
    });
}

int fs2dh_prepare(fs2dh_solver s)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() { h->solver->prepare(); });
}

int fs2dh_step_frame(fs2dh_solver s)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() { h->solver->stepFrame(); });
}

int fs2dh_step_substep(fs2dh_solver s, int *frame_finished)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() {
        const bool done = h->solver->stepSubstep();
        if (frame_finished) *frame_finished = done ? 1 : 0;
    });
}

int fs2dh_step_substep_streamed(fs2dh_solver s, void *host_buf, int64_t capacity_records, int64_t count_in, int64_t *count_out,
                                int *frame_finished)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() {
        const bool done = h->solver->stepSubstepStreamed(host_buf, capacity_records, count_in, count_out);
        if (frame_finished) *frame_finished = done ? 1 : 0;
    });
}

int fs2dh_save_state(fs2dh_solver s, const char *path)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() { h->solver->saveState(path); });
}

int fs2dh_load_state(fs2dh_solver s, const char *path)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() { h->solver->loadState(path); });
}

int fs2dh_get_stats(fs2dh_solver s, float *timings12, float *misc5)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() {
        const SolverStats &st = h->solver->timeStats();
        const SolverStats::StageTimings t = st.timings();
        for (int k = 0; k < SOLVER_STAGE_COUNT; k++) timings12[k] = t[static_cast<size_t>(k)];
        misc5[0] = st.frameTime();
        misc5[1] = static_cast<float>(st.substepCount());
        misc5[2] = static_cast<float>(st.pressureIterations());
        misc5[3] = static_cast<float>(st.densityIterations());
        misc5[4] = static_cast<float>(st.viscosityIterations());
    });
}

int fs2dh_size_i(fs2dh_solver s) { return s ? static_cast<int>(static_cast<Holder *>(s)->solver->gridSizeI()) : 0; }
int fs2dh_size_j(fs2dh_solver s) { return s ? static_cast<int>(static_cast<Holder *>(s)->solver->gridSizeJ()) : 0; }
int fs2dh_sim_type(fs2dh_solver s) { return s ? static_cast<int>(static_cast<Holder *>(s)->solver->simulationMethod()) : 0; }
int fs2dh_frame_number(fs2dh_solver s) { return s ? static_cast<Holder *>(s)->solver->frameNumber() : 0; }
int64_t fs2dh_particle_count(fs2dh_solver s) { return s ? static_cast<int64_t>(static_cast<Holder *>(s)->solver->particleCount()) : 0; }
int64_t fs2dh_kernel_launches(fs2dh_solver s) { return s ? static_cast<Holder *>(s)->solver->kernelLaunches() : 0; }

fs2d_handle fs2dh_device(fs2dh_solver s)
{
    Holder *h = static_cast<Holder *>(s);
    fs2d_handle out = nullptr;
    guarded(h, [&]() { out = h->solver->device(); });
    return out;
}

// Accessor round trip: the host copies the GUI reads (materialGrid().data(), bins()...).
int fs2dh_material(fs2dh_solver s, int8_t *out)
{
    Holder *h = static_cast<Holder *>(s);
    return guarded(h, [&]() {
        const MaterialGrid &g = h->solver->materialGrid();
        std::memcpy(out, g.data().data(), g.data().size());
    });
}

int64_t fs2dh_bin_sizes(fs2dh_solver s, int32_t *out, int64_t capacity)
{
    Holder *h = static_cast<Holder *>(s);
    int64_t n = 0;
    guarded(h, [&]() {
        std::vector<ParticleBin> &bins = h->solver->markerParticles().bins().data();
        n = static_cast<int64_t>(bins.size());
        for (int64_t k = 0; k < n && k < capacity; k++) out[k] = static_cast<int32_t>(bins[static_cast<size_t>(k)].size());
    });
    return n;
}

}  // extern "C"
