/* TEST INFRASTRUCTURE ONLY -- C ABI of the reference oracle (oracle/_ref/libfs2d_ref*.so).
 *
 * The library behind this header is the UNMODIFIED ArtNlk/FlipSolver2d source tree
 * (/root/reference, FlipSolver2dLib + Utils/jsonscenereader) compiled by
 * oracle/build_ref.py together with oracle/ref_harness.cpp, which subclasses the
 * reference's solver classes to reach their `protected` stages and state
 * (FlipSolver2dLib/flipsolver2d.h:280-431).  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may load it.
 */
#ifndef FS2D_REF_API_H
#define FS2D_REF_API_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Stage ids accepted by ref_run_stage(); each is one call the reference makes
 * inside FlipSolver::step() (flipsolver2d.cpp:412-462) or NBFlipSolver::step()
 * (nbflipsolver.cpp:26-64). */
enum RefStage
{
    REF_STAGE_ADVECT = 0,             /* advect()                                   */
    REF_STAGE_BUILD_MATRIX = 1,       /* getPressureProjectionMatrix + getIPPCoefficients */
    REF_STAGE_PRUNE_REBIN = 2,        /* pruneParticles(); rebinParticles()          */
    REF_STAGE_DENSITY_CORRECTION = 3, /* densityCorrection()                         */
    REF_STAGE_P2G = 4,                /* particleToGrid()                            */
    REF_STAGE_UPDATE_SDF = 5,         /* updateSdf()                                 */
    REF_STAGE_UPDATE_MATERIALS = 6,   /* updateMaterials()                           */
    REF_STAGE_AFTER_TRANSFER = 7,     /* afterTransfer()                             */
    REF_STAGE_EXTRAPOLATE_SDF_IN = 8, /* extrapolateLevelsetInside(m_fluidSdf)       */
    REF_STAGE_EXTRAPOLATE_VEL = 9,    /* m_fluidVelocityGrid.extrapolate(10)         */
    REF_STAGE_SAVE_VELOCITY = 10,     /* m_savedFluidVelocityGrid = m_fluidVelocityGrid */
    REF_STAGE_BODY_FORCES = 11,       /* applyBodyForces()                           */
    REF_STAGE_PROJECT = 12,           /* project()                                   */
    REF_STAGE_VELOCITY_FROM_SOLIDS = 13, /* updateVelocityFromSolids()               */
    REF_STAGE_VISCOSITY = 14,         /* applyViscosity()                            */
    REF_STAGE_PARTICLE_UPDATE = 15,   /* particleUpdate()                            */
    REF_STAGE_COUNT_PARTICLES = 16,   /* countParticles()                            */
    REF_STAGE_RESEED = 17,            /* reseedParticles()                           */
    REF_STAGE_GRID_UPDATE = 18,       /* gridUpdate() (virtual; nbflip differs)      */
    REF_STAGE_FULL_STEP = 19,         /* step() (virtual)                            */
    REF_STAGE_UPDATE_DENSITY_GRID = 20, /* updateDensityGrid()                       */
    REF_STAGE_EXTRAPOLATE_SDF_OUT = 21, /* extrapolateLevelsetOutside(m_fluidSdf)    */
    REF_STAGE_FIRST_FRAME_INIT = 22   /* firstFrameInit() (virtual)                  */
};

/* Grid ids for ref_get_grid()/ref_set_grid(). dtype in brackets. */
enum RefGrid
{
    REF_GRID_U = 0,           /* f32 (I+1)*J */
    REF_GRID_V = 1,           /* f32 I*(J+1) */
    REF_GRID_U_VALID = 2,     /* u8  (I+1)*J */
    REF_GRID_V_VALID = 3,     /* u8  I*(J+1) */
    REF_GRID_SAVED_U = 4,     /* f32 (I+1)*J */
    REF_GRID_SAVED_V = 5,     /* f32 I*(J+1) */
    REF_GRID_MATERIAL = 6,    /* i8  N */
    REF_GRID_FLUID_SDF = 7,   /* f32 N */
    REF_GRID_SOLID_SDF = 8,   /* f32 N */
    REF_GRID_VISCOSITY = 9,   /* f32 N */
    REF_GRID_DENSITY = 10,    /* f32 N */
    REF_GRID_COUNTS = 11,     /* i32 N */
    REF_GRID_EMITTER_ID = 12, /* i32 N */
    REF_GRID_SOLID_ID = 13,   /* i32 N */
    REF_GRID_DIVERGENCE_CONTROL = 14, /* f32 N */
    REF_GRID_TEST = 15,       /* f32 N */
    REF_GRID_KNOWN_CENTERED = 16, /* u8 N */
    REF_GRID_TEMPERATURE = 17,    /* f32 N (smoke/fire only) */
    REF_GRID_CONCENTRATION = 18,  /* f32 N (smoke/fire only) */
    REF_GRID_FUEL = 19            /* f32 N (fire only) */
};

typedef void *ref_handle;

/* JsonSceneReader::loadJson (Utils/jsonscenereader.cpp:8-77) through an exposing
 * subclass. Returns NULL on failure. */
ref_handle ref_load_scene(const char *json_path);
void ref_destroy(ref_handle h);

int ref_thread_count(void);        /* ThreadPool::i()->threadCount() */
void ref_set_quiet(int quiet);     /* 1: redirect std::cout of the reference to /dev/null */

int ref_size_i(ref_handle h);
int ref_size_j(ref_handle h);
int ref_sim_type(ref_handle h);    /* SimulationMethod */
int64_t ref_particle_count(ref_handle h);
int ref_property_count(ref_handle h);
int ref_frame_number(ref_handle h);

/* Scalars of FlipSolverParameters as the solver stores them. out[16]:
 * 0 stepDt 1 frameDt 2 dx 3 fluidDensity 4 ppc 5 gx 6 gy 7 picRatio 8 cfl
 * 9 particleScale 10 pcgIterLimit 11 projectTolerance 12 maxSubsteps
 * 13 viscosityEnabled 14 parameterHandlingMethod 15 fps */
void ref_get_params(ref_handle h, double *out16);

void ref_step_frame(ref_handle h);                        /* FlipSolver::stepFrame */
/* timings[12] (ms per SolverStage), misc[5] = frameTime, substeps, pressureIters,
 * densityIters, viscosityIters */
void ref_get_stats(ref_handle h, float *timings12, float *misc5);

void ref_set_step_dt(ref_handle h, float dt);
float ref_max_particle_velocity(ref_handle h);
void ref_run_stage(ref_handle h, int stage);
void ref_bump_frame_number(ref_handle h);                 /* m_frameNumber++ */

/* Particles, concatenated in bin order (bin 0.., in-bin order as stored).
 * pos/vel: 2 floats per particle (x, y). props: property-major [k][count].
 * bin_of: the bin each particle is stored in. Any pointer may be NULL. */
void ref_get_particles(ref_handle h, float *pos, float *vel, float *props, int32_t *bin_of);
/* Replace all particles; each goes to binForGridPosition(pos) in the given order. */
void ref_set_particles(ref_handle h, int64_t count, const float *pos, const float *vel, const float *props);

int64_t ref_grid_size(ref_handle h, int grid);
int ref_get_grid(ref_handle h, int grid, void *out);
int ref_set_grid(ref_handle h, int grid, const void *in);

/* Pressure system built by REF_STAGE_BUILD_MATRIX. Dense per-cell export:
 * is_unit[N] (1 where a matrix row exists), mask[N], count[N],
 * coef[4*N] = iNeg, iPos, jNeg, jPos planes. Returns scale. */
double ref_get_matrix(ref_handle h, uint8_t *is_unit, uint8_t *mask, uint8_t *count, double *coef);
void ref_spmv(ref_handle h, const double *in, double *out);           /* IndexedPressureParameters::multiply */
void ref_precond_apply(ref_handle h, const double *in, double *out);  /* IndexedIPPCoefficients::multiply   */
int ref_pcg_solve(ref_handle h, const double *rhs, double *x, int iter_limit, double tol); /* LinearSolver::solve */
void ref_pressure_rhs(ref_handle h, double *rhs);                     /* calcPressureRhs */
void ref_density_rhs(ref_handle h, double *rhs);                      /* calcDensityCorrectionRhs */
void ref_apply_pressure(ref_handle h, const double *p);               /* applyPressuresToVelocityField */

/* VOps (vmath.cpp) */
double ref_vops_dot(const double *a, const double *b, int64_t n);
double ref_vops_max_abs(const double *a, int64_t n);

#ifdef __cplusplus
}
#endif
#endif
