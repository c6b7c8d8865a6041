// Implicit viscosity: LightViscosityModel::apply (viscositymodel.cpp:4-162) as a matrix-free Jacobi-preconditioned
// conjugate gradient on the device.
//
// The reference assembles an N x N (N = I*J, cell indexed) Eigen::SparseMatrix with coeffRef ASSIGNMENTS
// (viscositymodel.cpp:65-138) and hands it to Eigen::ConjugateGradient<SparseMatrix<double>, Upper>
// (viscositymodel.h:25), which reads the diagonal and the UPPER triangle only. With the write order of getMatrix
// (row-major cells, the later writer wins) the operator every solve sees is
//     d(c)      = 1                      if c is SOLID, else 1 + 4*mu(c)*dt
//     E(c, c+1) = E(c, c+J) = mu(c)*dt   if c is not SOLID
//               = mu(nb)*dt              if c is SOLID and the neighbour nb is not
//               = absent                 if both are SOLID (or the neighbour is outside the grid)
// applied symmetrically, all products in double with the float viscosity and float dt widened first. The same
// matrix serves U and V: rhs[i*J + j] = density * field.at(i, j) for i < I, j < J, the solution divided by the
// density is written back to those samples (:140-162) -- U's last row and V's last column stay untouched.
//
// Solver = Eigen 3.4.0's conjugate_gradient (third-party, not in the reference tree; DESIGN.md section 2 says how
// the tests restate it): x0 = 0, preconditioner 1/diag, stop when |r|^2 < max(tol^2 |b|^2, DBL_MIN) with
// tol = 1e-4 (viscositymodel.cpp:24), at most 2N iterations; iterations() counts the completed loop bodies before
// the one that broke out. apply() returns the iteration count of the V solve (:52).
//
// Three streaming kernels per iteration with device-resident scalars; the host looks at the "done" flag every few
// iterations only. Reductions: per-CTA partials, fixed-order final sum by the last CTA -> run-to-run deterministic.
// This stage takes a handful of iterations (6 at 128^2), far from the PCG's share of a substep.
#include <algorithm>
#include <cfloat>
#include <cmath>

#include "fs2d_internal.h"

namespace
{
constexpr int NT = 256;

struct ViscScalars
{
    double rhsNorm2, threshold, absNew, pAp, alpha, beta, resNorm2;
    int iter;      // completed loop bodies (Eigen's `i`)
    int done;      // 1 once the solve has finished (converged, zero rhs, or iteration cap)
    int applied;   // loop bodies that ran (diagnostics)
    unsigned int ticket;
};

struct ViscArgs
{
    const int8_t *material;
    const float *mu;
    int I, J;
    long long N;
    double dt;
    double *x, *r, *p, *tmp, *z;
    double *partials;
    ViscScalars *sc;
    long long maxIters;
};

__device__ __forceinline__ double warpSum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// Sum over the CTA of up to two values; results valid in thread 0.
__device__ void blockSum2(double &a, double &b, double *scratch /* 16 */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    a = warpSum(a);
    b = warpSum(b);
    __syncthreads();
    if (lane == 0)
    {
        scratch[warp] = a;
        scratch[8 + warp] = b;
    }
    __syncthreads();
    if (warp == 0)
    {
        a = lane < (blockDim.x >> 5) ? scratch[lane] : 0.0;
        b = lane < (blockDim.x >> 5) ? scratch[8 + lane] : 0.0;
        a = warpSum(a);
        b = warpSum(b);
    }
}

// Per-CTA partials -> the last CTA sums them in a fixed order; returns true (in every thread of that CTA) with the
// totals in thread 0.
__device__ bool gridSum2(double &a, double &b, double *partials, unsigned int *ticket, double *scratch, int *isLast)
{
    const int nb = gridDim.x;
    blockSum2(a, b, scratch);
    if (threadIdx.x == 0)
    {
        partials[blockIdx.x] = a;
        partials[nb + blockIdx.x] = b;
        __threadfence();
        *isLast = atomicAdd(ticket, 1u) == static_cast<unsigned int>(nb - 1);
    }
    __syncthreads();
    if (!*isLast) return false;
    __threadfence();
    double ta = 0.0, tb = 0.0;
    for (int k = threadIdx.x; k < nb; k += blockDim.x)
    {
        ta += __ldcg(partials + k);
        tb += __ldcg(partials + nb + k);
    }
    blockSum2(ta, tb, scratch);
    a = ta;
    b = tb;
    if (threadIdx.x == 0) *ticket = 0;
    return true;
}

__device__ __forceinline__ double diagAt(const ViscArgs &a, long long n)
{
    if (matSolid(a.material[n])) return 1.0;
    double d = 4.0;
    d *= static_cast<double>(a.mu[n]) * a.dt;  // diag *= viscosityGrid.at(i,j) * scale
    d += 1.0;
    return d;
}

// Coupling between cell n and its neighbour m > n (m = n + 1 or n + J), as Eigen's Upper view sees it.
__device__ __forceinline__ double upperCoupling(const ViscArgs &a, long long n, long long m)
{
    const bool sn = matSolid(a.material[n]);
    if (!sn) return static_cast<double>(a.mu[n]) * a.dt;
    if (!matSolid(a.material[m])) return static_cast<double>(a.mu[m]) * a.dt;
    return 0.0;
}

// y = selfadjointView<Upper>(A) * v at cell (i, j)
__device__ __forceinline__ double applyRow(const ViscArgs &a, const double *__restrict__ v, int i, int j)
{
    const long long J = a.J, n = static_cast<long long>(i) * J + j;
    double y = diagAt(a, n) * v[n];
    if (i > 0) y += upperCoupling(a, n - J, n) * v[n - J];
    if (j > 0) y += upperCoupling(a, n - 1, n) * v[n - 1];
    if (j + 1 < a.J) y += upperCoupling(a, n, n + 1) * v[n + 1];
    if (i + 1 < a.I) y += upperCoupling(a, n, n + J) * v[n + J];
    return y;
}

// rhs = density * field (float product, viscositymodel.cpp:140-150); x = 0; r = rhs; p = r / diag;
// rhsNorm2 = |rhs|^2, absNew = r.p; zero / already-converged exits of Eigen's conjugate_gradient.
__global__ void __launch_bounds__(NT) viscInitKernel(ViscArgs a, const float *__restrict__ field, int fieldStride, float density)
{
    __shared__ double scratch[16];
    __shared__ int isLast;
    double s0 = 0.0, s1 = 0.0;
    for (long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; n < a.N; n += static_cast<long long>(gridDim.x) * NT)
    {
        const long long i = n / a.J, j = n - i * a.J;
        const double b = static_cast<double>(__fmul_rn(density, field[i * fieldStride + j]));
        a.x[n] = 0.0;
        a.r[n] = b;
        const double p = b * (1.0 / diagAt(a, n));  // DiagonalPreconditioner: m_invdiag(j) * b(j)
        a.p[n] = p;
        s0 += b * b;
        s1 += b * p;
    }
    if (!gridSum2(s0, s1, a.partials, &a.sc->ticket, scratch, &isLast)) return;
    if (threadIdx.x == 0)
    {
        ViscScalars *sc = a.sc;
        sc->rhsNorm2 = s0;
        sc->threshold = fmax(1e-4 * 1e-4 * s0, DBL_MIN);
        sc->absNew = s1;
        sc->resNorm2 = s0;
        sc->alpha = sc->beta = sc->pAp = 0.0;
        sc->iter = 0;
        sc->applied = 0;
        sc->done = (s0 == 0.0 || s0 < sc->threshold) ? 1 : 0;
    }
}

// tmp = A p; pAp -> alpha
__global__ void __launch_bounds__(NT) viscApplyKernel(ViscArgs a)
{
    __shared__ double scratch[16];
    __shared__ int isLast;
    if (a.sc->done) return;
    double s0 = 0.0, s1 = 0.0;
    for (long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; n < a.N; n += static_cast<long long>(gridDim.x) * NT)
    {
        const int i = static_cast<int>(n / a.J), j = static_cast<int>(n - static_cast<long long>(i) * a.J);
        const double y = applyRow(a, a.p, i, j);
        a.tmp[n] = y;
        s0 += a.p[n] * y;
    }
    if (!gridSum2(s0, s1, a.partials, &a.sc->ticket, scratch, &isLast)) return;
    if (threadIdx.x == 0)
    {
        a.sc->pAp = s0;
        a.sc->alpha = a.sc->absNew / s0;
    }
}

// x += alpha p; r -= alpha tmp; |r|^2; z = r / diag; r.z -> convergence test / beta
__global__ void __launch_bounds__(NT) viscUpdateKernel(ViscArgs a)
{
    __shared__ double scratch[16];
    __shared__ int isLast;
    if (a.sc->done) return;
    const double alpha = a.sc->alpha;
    double s0 = 0.0, s1 = 0.0;
    for (long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; n < a.N; n += static_cast<long long>(gridDim.x) * NT)
    {
        a.x[n] += alpha * a.p[n];
        const double r = a.r[n] - alpha * a.tmp[n];
        a.r[n] = r;
        const double z = r * (1.0 / diagAt(a, n));
        a.z[n] = z;
        s0 += r * r;
        s1 += r * z;
    }
    if (!gridSum2(s0, s1, a.partials, &a.sc->ticket, scratch, &isLast)) return;
    if (threadIdx.x == 0)
    {
        ViscScalars *sc = a.sc;
        sc->resNorm2 = s0;
        sc->applied++;
        if (s0 < sc->threshold)
            sc->done = 1;  // break before i++
        else
        {
            sc->beta = s1 / sc->absNew;
            sc->absNew = s1;
            sc->iter++;    // the direction update that follows completes this loop body
            if (sc->iter >= a.maxIters) sc->done = 1;  // while (i < maxIters); p is not needed any more
        }
    }
}

// p = z + beta p
__global__ void __launch_bounds__(NT) viscDirectionKernel(ViscArgs a)
{
    if (a.sc->done) return;
    const double beta = a.sc->beta;
    for (long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; n < a.N; n += static_cast<long long>(gridDim.x) * NT)
        a.p[n] = a.z[n] + beta * a.p[n];
}

// field(i, j) = x / density (viscositymodel.cpp:152-162)
__global__ void __launch_bounds__(NT) viscWriteBackKernel(ViscArgs a, float *__restrict__ field, int fieldStride, float density)
{
    for (long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; n < a.N; n += static_cast<long long>(gridDim.x) * NT)
    {
        const long long i = n / a.J, j = n - i * a.J;
        field[i * fieldStride + j] = static_cast<float>(a.x[n] / static_cast<double>(density));
    }
}
}  // namespace

// One Eigen-style solve for `field` (U: stride J; V: stride J + 1). *iters = Eigen's iterations(); returns
// FS2D_OK with *failed = 1 when the solver did not reach the tolerance (the reference prints and gives up).
static int viscSolve(Ctx *ctx, ViscArgs &a, float *field, int stride, float density, int *iters, int *failed)
{
    cudaStream_t st = ctx->stream;
    const int blocks = std::min<long long>(divUp(ctx->N, NT), static_cast<long long>(ctx->smCount) * 8);
    viscInitKernel<<<blocks, NT, 0, st>>>(a, field, stride, density);
    ctx->launches++;
    ViscScalars sc;
    const int batch = 8;
    for (;;)
    {
        for (int k = 0; k < batch; k++)
        {
            viscApplyKernel<<<blocks, NT, 0, st>>>(a);
            viscUpdateKernel<<<blocks, NT, 0, st>>>(a);
            viscDirectionKernel<<<blocks, NT, 0, st>>>(a);
        }
        ctx->launches += 3 * batch;
        FS2D_CUDA(cudaGetLastError());
        FS2D_CUDA(fs2dCopyToHost(ctx, &sc, a.sc, sizeof(sc)));
        if (sc.done) break;
    }
    const double err = sc.rhsNorm2 > 0.0 ? std::sqrt(sc.resNorm2 / sc.rhsNorm2) : 0.0;
    *failed = (sc.rhsNorm2 > 0.0 && !(err <= 1e-4)) ? 1 : 0;
    *iters = sc.iter;
    return FS2D_OK;
}

int gridViscosity(Ctx *ctx, int *iters)
{
    if (ctx->slab.enabled && ctx->slab.world > 1)
    {
        ctx->lastError = "applyViscosity is not slab-aware yet (only FS2D_SIM_LIQUID without viscosity runs on several GPUs)";
        return FS2D_ERR_STATE;
    }
    if (!ctx->viscScalars) FS2D_CUDA(cudaMalloc(&ctx->viscScalars, 256));
    FS2D_CUDA(cudaMemsetAsync(ctx->viscScalars, 0, 256, ctx->stream));
    ViscArgs a;
    a.material = ctx->material;
    a.mu = ctx->viscosity;
    a.I = ctx->I;
    a.J = ctx->J;
    a.N = ctx->N;
    a.dt = static_cast<double>(ctx->stepDt);  // const double scale = dt (float argument)
    // the Krylov vectors of the pressure solve are idle during this stage
    a.x = ctx->x;
    a.r = ctx->r[0];
    a.p = ctx->s[0];
    a.tmp = ctx->q;
    a.z = ctx->z;
    a.partials = ctx->partials;
    a.sc = static_cast<ViscScalars *>(ctx->viscScalars);
    a.maxIters = 2 * ctx->N;
    const float density = static_cast<float>(ctx->p.fluid_density);  // apply(..., float density)
    const int blocks = std::min<long long>(divUp(ctx->N, NT), static_cast<long long>(ctx->smCount) * 8);
    int itU = 0, itV = 0, failed = 0;
    FS2D_TRY(viscSolve(ctx, a, ctx->U, ctx->J, density, &itU, &failed));
    if (failed)
    {
        if (iters) *iters = -1;  // "Viscosity solver U solving failed!" -> return -1 before anything is applied
        return FS2D_OK;
    }
    viscWriteBackKernel<<<blocks, NT, 0, ctx->stream>>>(a, ctx->U, ctx->J, density);
    ctx->launches++;
    FS2D_TRY(viscSolve(ctx, a, ctx->V, ctx->J + 1, density, &itV, &failed));
    if (failed)
    {
        if (iters) *iters = -1;
        return FS2D_OK;
    }
    viscWriteBackKernel<<<blocks, NT, 0, ctx->stream>>>(a, ctx->V, ctx->J + 1, density);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    if (iters) *iters = itV;
    return FS2D_OK;
}
