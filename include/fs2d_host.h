/* fs2d_host.h -- C view of libfs2d_host.so, the C++ host mirror of the reference's solver API
 * (flipsolver2d_b200/host/: JsonSceneReader, FlipSolver, NBFlipSolver, FlipSmokeSolver,
 * FlipFireSolver, SolverStats -- same class and method names as FlipSolver2dLib/flipsolver2d.h:183-278
 * and Utils/jsonscenereader.h:16). C++ applications (AutoBench, a viewer) link the classes directly;
 * this shim exists for callers without a C++ toolchain (bench.py, tests) and replaces exactly the calls
 * AutoBench makes: loadJson -> stepFrame -> timeStats (AutoBench/benchmarkrunnerapplication.cpp:74-97).
 */
#ifndef FS2D_HOST_H
#define FS2D_HOST_H

#include "fs2d.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef void *fs2dh_solver;

/* process-wide switches; call before fs2dh_load_scene */
void fs2dh_set_quiet(int quiet);                      /* silence the per-substep stdout lines (flipsolver2d.cpp:491) */
void fs2dh_set_device(int ordinal);                   /* CUDA device of solvers created afterwards */
void fs2dh_set_convergence_threads(int threads);      /* fs2d_params.convergence_threads */

/* Row slabs over several GPUs (fs2d.h "row slabs"): one process and one solver per GPU, all loading the same scene.
 * fs2dh_set_slab applies to solvers whose device is created afterwards; exchange the FS2D_SLAB_HANDLE_BYTES blobs of
 * fs2dh_slab_export between the ranks and hand each to fs2dh_slab_connect before fs2dh_prepare / stepping. Stepping,
 * fs2dh_material and fs2dh_global_particle_count are then collective. */
void fs2dh_set_slab(int rank, int world, int device_share);
int fs2dh_slab_export(fs2dh_solver s, void *blob);
int fs2dh_slab_connect(fs2dh_solver s, int peer_rank, const void *blob);
int64_t fs2dh_global_particle_count(fs2dh_solver s);
/* The slab boundaries the solver would use for `world` ranks (world + 1 rows, balanced by seed particles per
 * 16-row tile row; host only, no GPU needed). Every rank computes the same table from the same scene. */
int fs2dh_slab_bounds(fs2dh_solver s, int world, int32_t *row_bounds);

fs2dh_solver fs2dh_load_scene(const char *json_path); /* JsonSceneReader::loadJson; NULL on failure */
void fs2dh_destroy(fs2dh_solver s);
const char *fs2dh_last_error(fs2dh_solver s);

/* Host-only half of frame 0: rasterise the scene polygons (updateSinks/Sources/Solids/InitialFluid,
 * flipsolver2d.cpp:502-590) and draw the seed particles (seedInitialFluid :682-707). No GPU needed. */
int fs2dh_prepare_host(fs2dh_solver s);
int64_t fs2dh_seed_count(fs2dh_solver s);
int fs2dh_seed_particles(fs2dh_solver s, float *pos, float *vel, float *props);
/* Host copy of a scene grid (MATERIAL, SOLID_SDF, FLUID_SDF, VISCOSITY, SOLID_ID, EMITTER_ID, DIVERGENCE_CONTROL). */
int fs2dh_host_grid(fs2dh_solver s, int grid, void *out);
int fs2dh_prepare(fs2dh_solver s);                    /* frame-0 init (firstFrameInit) without stepping */
int fs2dh_step_frame(fs2dh_solver s);                 /* FlipSolver::stepFrame */
/* One CFL substep of the current frame (the body of stepFrame's loop, flipsolver2d.cpp:476-497);
 * *frame_finished = 1 when it completed the frame. */
int fs2dh_step_substep(fs2dh_solver s, int *frame_finished);
/* The same substep with the particle state in a caller-owned (pinned) host buffer: count_in records of host_buf in the
 * sectioned layout of fs2d_particle_stream_begin (include/fs2d.h) are the particles the substep starts from, the
 * buffer holds *count_out records when the call returns; uploads and downloads overlap the stages. */
int fs2dh_step_substep_streamed(fs2dh_solver s, void *host_buf, int64_t capacity_records, int64_t count_in, int64_t *count_out,
                                int *frame_finished);
/* timings12: ms per SolverStage; misc5: frameTime ms, substeps, pressure/density/viscosity iterations */
int fs2dh_get_stats(fs2dh_solver s, float *timings12, float *misc5);

/* FlipSolver::saveState / loadState: checkpoint of the device state (fs2d_state_save) plus the host side of the
 * stepping loop (frame / substep counters, mt19937 stream). Load into a solver created from the same scene. */
int fs2dh_save_state(fs2dh_solver s, const char *path);
int fs2dh_load_state(fs2dh_solver s, const char *path);

int fs2dh_size_i(fs2dh_solver s);
int fs2dh_size_j(fs2dh_solver s);
int fs2dh_sim_type(fs2dh_solver s);
int fs2dh_frame_number(fs2dh_solver s);
int64_t fs2dh_particle_count(fs2dh_solver s);
int64_t fs2dh_kernel_launches(fs2dh_solver s);

fs2d_handle fs2dh_device(fs2dh_solver s);             /* the fs2d handle behind the solver (state download) */
int fs2dh_material(fs2dh_solver s, int8_t *out);      /* materialGrid().data() */
int64_t fs2dh_bin_sizes(fs2dh_solver s, int32_t *out, int64_t capacity); /* markerParticles().bins()[k].size() */

/* BenchRunTable::save (AutoBench/benchruntable.cpp:15-48): Stats.xlsx with one worksheet per scene and the 18 columns of
 * AutoBench/benchruntable.h:28-49, written without OpenXLSX. rows18 = all rows of all scenes back to back, 18 doubles
 * each, in the order of the TableColumn enum; no GPU needed. The Autobench driver fills the same table from SolverStats. */
int fs2dh_write_stats_xlsx(const char *path, int scenes, const char *const *scene_names, const int *rows_per_scene, const double *rows18);

#ifdef __cplusplus
}
#endif
#endif /* FS2D_HOST_H */
