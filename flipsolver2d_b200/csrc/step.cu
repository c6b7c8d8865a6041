// One whole substep on the device: FlipSolver::step() (flipsolver2d.cpp:412-462) or
// NBFlipSolver::step() (nbflipsolver.cpp:26-64), stage boundaries and SolverStage slots as in the
// reference (flipsolver2d.h:58-72), timed with CUDA events on the handle's stream.
//
// Scenes with SOURCE cells need the host's std::mt19937 stream for reseeding
// (flipsolver2d.cpp:1013-1019); for those the host mirror drives the stage entry points one by one
// and calls fs2d_reseed_plan/apply itself. fs2d_substep covers everything up to and including
// countParticles and refuses scenes whose reseed plan is non-empty.
#include "fs2d_internal.h"

namespace
{
enum Stage
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
    STAGE_COUNT
};

struct StageClock
{
    Ctx *ctx;
    int used = 0;
    int stageOf[15];
    explicit StageClock(Ctx *c) : ctx(c) { cudaEventRecord(ctx->ev[0], ctx->stream); }
    void end(int stage)
    {
        if (used >= 15) return;
        stageOf[used] = stage;
        used++;
        cudaEventRecord(ctx->ev[used], ctx->stream);
    }
    void resolve(float *ms)
    {
        cudaEventSynchronize(ctx->ev[used]);
        for (int k = 0; k < STAGE_COUNT; k++) ms[k] = 0.f;
        for (int k = 0; k < used; k++)
        {
            float t = 0.f;
            cudaEventElapsedTime(&t, ctx->ev[k], ctx->ev[k + 1]);
            ms[stageOf[k]] += t;
        }
    }
};

int projectStage(Ctx *ctx, int *iters)
{
    FS2D_TRY(gridPressureRhs(ctx));
    FS2D_TRY(pcgSolveDevice(ctx, ctx->p.pcg_iter_limit, ctx->p.project_tolerance));
    FS2D_TRY(slabExchangePressure(ctx));
    FS2D_TRY(gridApplyPressure(ctx));
    if (iters) FS2D_TRY(fs2d_pcg_last_iterations(ctx, iters));
    return FS2D_OK;
}

int stepFlip(Ctx *ctx, StageClock &clk, int *iters)
{
    FS2D_TRY(fs2d_advect(ctx));
    clk.end(ADVECTION);
    FS2D_TRY(gridBuildMatrix(ctx));
    clk.end(DECOMPOSITION);
    FS2D_TRY(particlesRebin(ctx));
    clk.end(PARTICLE_REBIN);
    if (!ctx->p.viscosity_enabled)
    {
        FS2D_TRY(fs2d_density_correction(ctx, iters ? iters + 1 : nullptr));
        clk.end(DENSITY);
    }
    FS2D_TRY(fs2d_particle_to_grid(ctx));
    clk.end(PARTICLE_TO_GRID);
    FS2D_TRY(transferSdf(ctx));
    FS2D_TRY(gridUpdateMaterials(ctx));
    clk.end(GRID_UPDATE);
    FS2D_TRY(gridAfterTransfer(ctx));
    FS2D_TRY(gridExtrapolateSdf(ctx, true));
    FS2D_TRY(gridExtrapolateVelocity(ctx, 10));
    FS2D_TRY(gridSaveVelocity(ctx));
    FS2D_TRY(gridBodyForces(ctx));
    clk.end(AFTER_TRANSFER);
    // streamed particle state with an announced output buffer: no stage since the density correction, and none below, moves
    // or reorders the existing records; the copy runs under the pressure solve
    if (ctx->pstream.outHost)
        FS2D_TRY(fs2d_particle_stream_positions_final(ctx, ctx->pstream.outHost, ctx->pstream.outCapacity, ctx->p.sim_type == FS2D_SIM_LIQUID ? 1 : 0));
    FS2D_TRY(projectStage(ctx, iters));
    clk.end(PRESSURE);
    FS2D_TRY(gridVelocityFromSolids(ctx));
    if (ctx->p.viscosity_enabled)
    {
        FS2D_TRY(gridViscosity(ctx, iters ? iters + 2 : nullptr));
        clk.end(VISCOSITY);
        FS2D_TRY(projectStage(ctx, iters));
        clk.end(REPRESSURE);
    }
    FS2D_TRY(gridExtrapolateVelocity(ctx, 10));
    FS2D_TRY(particlesUpdate(ctx));
    if (ctx->pstream.outHost) FS2D_TRY(fs2d_particle_stream_velocities_final(ctx, ctx->pstream.outHost, ctx->pstream.outCapacity));
    clk.end(PARTICLE_UPDATE);
    FS2D_TRY(particlesCount(ctx));
    return FS2D_OK;
}

int stepNbflip(Ctx *ctx, StageClock &clk, int *iters)
{
    FS2D_TRY(fs2d_advect(ctx));
    FS2D_TRY(fs2d_nbflip_advect_grids(ctx));
    clk.end(ADVECTION);
    FS2D_TRY(particlesRebin(ctx));
    clk.end(PARTICLE_REBIN);
    FS2D_TRY(fs2d_particle_to_grid(ctx));
    FS2D_TRY(gridExtrapolateVelocity(ctx, 10));
    FS2D_TRY(gridSaveVelocity(ctx));
    clk.end(PARTICLE_TO_GRID);
    // NBFlipSolver::gridUpdate (nbflipsolver.cpp:213-225)
    FS2D_TRY(transferSdf(ctx));
    FS2D_TRY(gridExtrapolateSdf(ctx, false));
    FS2D_TRY(gridAfterTransfer(ctx));
    FS2D_TRY(gridExtrapolateSdf(ctx, true));
    clk.end(AFTER_TRANSFER);
    FS2D_TRY(gridUpdateMaterials(ctx));
    FS2D_TRY(gridBodyForces(ctx));
    clk.end(GRID_UPDATE);
    FS2D_TRY(gridBuildMatrix(ctx));
    clk.end(DECOMPOSITION);
    FS2D_TRY(projectStage(ctx, iters));
    clk.end(PRESSURE);
    FS2D_TRY(gridVelocityFromSolids(ctx));
    if (ctx->p.viscosity_enabled)
    {
        FS2D_TRY(gridViscosity(ctx, iters ? iters + 2 : nullptr));
        clk.end(VISCOSITY);
        FS2D_TRY(projectStage(ctx, iters));
        clk.end(REPRESSURE);
    }
    FS2D_TRY(gridExtrapolateVelocity(ctx, 10));
    FS2D_TRY(particlesUpdate(ctx));
    clk.end(PARTICLE_UPDATE);
    FS2D_TRY(particlesCount(ctx));
    return FS2D_OK;
}
}  // namespace

int stepSubstep(Ctx *ctx, float dt, float *stageMs, int *iters)
{
    ctx->stepDt = dt;
    if (iters) iters[0] = iters[1] = iters[2] = 0;
    StageClock clk(ctx);
    if (ctx->p.sim_type == FS2D_SIM_NBFLIP)
        FS2D_TRY(stepNbflip(ctx, clk, iters));
    else
        FS2D_TRY(stepFlip(ctx, clk, iters));
    int64_t candidates = 0;
    FS2D_TRY(particlesReseedPlan(ctx, &candidates));
    clk.end(PARTICLE_RESEED);
    if (stageMs) clk.resolve(stageMs);
    if (candidates != 0)
    {
        ctx->lastError = "fs2d_substep: the scene reseeds particles; drive the stages and fs2d_reseed_apply from the host";
        return FS2D_ERR_STATE;
    }
    return FS2D_OK;
}
