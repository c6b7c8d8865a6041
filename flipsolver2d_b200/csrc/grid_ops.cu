// Kernel group 3 (classification / extrapolation) and the grid-side pieces of group 4
// (matrix rows, right-hand sides, pressure application) and group 5 (semi-Lagrangian
// advection). All are one-thread-per-sample streaming kernels over the dense row-major
// grids; each sample is read/written once, neighbours come from L1/L2.
#include "fs2d_internal.h"
#include "fs2d_device.cuh"

namespace
{
constexpr int NT = 256;

// ------------------------------------------------------------------ matrix rows
// getPressureProjectionMatrix (flipsolver2d.cpp:797-887; smoke flipsmokesolver.cpp:354-444)
// and getIPPCoefficients (flipsolver2d.cpp:889-945) as per-cell bit fields.
__global__ void __launch_bounds__(NT) buildMatrixKernel(const int8_t *__restrict__ mat, int I, int J, int smokeRows,
                                                        uint8_t *__restrict__ rowInfo, uint16_t *__restrict__ preInfo)
{
    const long long N = static_cast<long long>(I) * J;
    const long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n >= N) return;
    const int i = static_cast<int>(n / J), j = static_cast<int>(n - static_cast<long long>(i) * J);
    const int8_t m = mat[n];
    const bool hasRow = smokeRows ? !matSolid(m) : matFluid(m);
    if (!hasRow)
    {
        rowInfo[n] = 0;
        preInfo[n] = 0;
        return;
    }
    unsigned int cnt = 0, mask = 0;
    {
        const int8_t a = matAt(mat, I, J, i - 1, j);
        if (matFluid(a)) { cnt++; if (n - J >= 0) mask |= 1u; } else if (matEmpty(a)) cnt++;
        const int8_t b = matAt(mat, I, J, i + 1, j);
        if (matFluid(b)) { cnt++; if (n + J < N) mask |= 2u; } else if (matEmpty(b)) cnt++;
        const int8_t c = matAt(mat, I, J, i, j - 1);
        if (matFluid(c)) { cnt++; if (n - 1 >= 0) mask |= 4u; } else if (matEmpty(c)) cnt++;
        const int8_t d = matAt(mat, I, J, i, j + 1);
        if (matFluid(d)) { cnt++; if (n + 1 < N) mask |= 8u; } else if (matEmpty(d)) cnt++;
    }
    rowInfo[n] = static_cast<uint8_t>(FS2D_ROW_UNIT | (cnt << 4) | mask);
    // nonsolidNeighborCount(linear index of the neighbour) -> index2d with truncating
    // division (linearindexable2d.h:44-52), then OOB_EXTEND look-ups (materialgrid.cpp:130-140)
    const long long nb[4] = {n - J, n + J, n - 1, n + 1};
    unsigned int pre = FS2D_PRE_UNIT;
#pragma unroll
    for (int k = 0; k < 4; k++)
    {
        const long long ni = nb[k] / J;  // C++ division truncates toward zero, as in the reference
        const long long nj = nb[k] - ni * J;
        const unsigned int c = nonsolidCount(mat, I, J, static_cast<int>(ni), static_cast<int>(nj));
        pre |= c << (3 * k);
    }
    preInfo[n] = static_cast<uint16_t>(pre);
}
}  // namespace

int gridBuildMatrix(Ctx *ctx)
{
    const int smokeRows = (ctx->p.sim_type == FS2D_SIM_SMOKE || ctx->p.sim_type == FS2D_SIM_FIRE) ? 1 : 0;
    buildMatrixKernel<<<divUp(ctx->N, NT), NT, 0, ctx->stream>>>(ctx->material, ctx->I, ctx->J, smokeRows, ctx->rowInfo,
                                                                  ctx->preInfo);
    ctx->launches++;
    // scale = dt / (rho dx^2) (flipsolver2d.cpp:799), float dt promoted to double
    ctx->matrixScale = static_cast<double>(ctx->stepDt) / (ctx->p.fluid_density * ctx->p.dx * ctx->p.dx);
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}
