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

#include "fs2d_device.cuh"
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
    const int *box;  // {iMin, iMax, jMin, jMax} of the cells with a non-zero right-hand side or viscosity (viscBoxKernel)
};

// The rows of the system outside the box B of cells with a non-zero right-hand side or a non-zero viscosity are identity
// rows with b = 0, and so are their couplings: their r, p, z and x stay exactly zero in Eigen's CG (a solid cell next to
// a viscous one has a non-zero coupling -- it lies within one cell of B). The solve therefore runs on B grown by one
// cell and keeps zeros on a second ring for the operator to read; everything outside contributes exact zeros to the dot
// products and keeps its (zero) velocity. At 4096^2 the box of the narrow-band dam break is 9 % of the grid.
struct ViscBox
{
    int i0, j0, h, w;  // first row / column, rows, columns
    __device__ __forceinline__ long long cells() const { return static_cast<long long>(h) * w; }
};

__device__ __forceinline__ ViscBox viscBox(const ViscArgs &a, int grow)
{
    ViscBox b;
    const int iMin = a.box[0], iMax = a.box[1], jMin = a.box[2], jMax = a.box[3];
    if (iMax < iMin)
    {
        b.i0 = b.j0 = b.h = b.w = 0;
        return b;
    }
    b.i0 = max(iMin - grow, 0);
    b.j0 = max(jMin - grow, 0);
    b.h = min(iMax + grow, a.I - 1) - b.i0 + 1;
    b.w = min(jMax + grow, a.J - 1) - b.j0 + 1;
    return b;
}

// cell t of the box -> (i, j); the boxes of the grids this runs on have fewer than 2^32 cells
__device__ __forceinline__ void viscCell(const ViscBox &b, long long t, int *i, int *j)
{
    const unsigned int q = static_cast<unsigned int>(t) / static_cast<unsigned int>(b.w);
    *i = b.i0 + static_cast<int>(q);
    *j = b.j0 + static_cast<int>(static_cast<unsigned int>(t) - q * static_cast<unsigned int>(b.w));
}

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

// n / J: a 32-bit division whenever the grid allows it (the 64-bit one is ~100 instructions and was most of the
// per-cell work of these streaming kernels)
__device__ __forceinline__ long long rowOf(long long n, long long J, long long N)
{
    return N <= 0x7fffffffll ? static_cast<long long>(static_cast<unsigned int>(n) / static_cast<unsigned int>(J)) : n / J;
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

__global__ void viscBoxResetKernel(int *box)
{
    box[0] = 0x7fffffff;
    box[1] = -1;
    box[2] = 0x7fffffff;
    box[3] = -1;
}

// Bounding box of the cells whose right-hand side (density * field, zero exactly when the field is) or viscosity is not zero.
__global__ void __launch_bounds__(NT) viscBoxKernel(const float *__restrict__ field, int fieldStride, const float *__restrict__ mu, int rowLo, int rowHi,
                                                    int J, int *__restrict__ box)
{
    // a CTA takes whole rows (of [rowLo, rowHi)), its threads stride over the columns (no index division)
    int iMin = 0x7fffffff, iMax = -1, jMin = 0x7fffffff, jMax = -1;
    for (int i = rowLo + blockIdx.x; i < rowHi; i += gridDim.x)
        for (int j = threadIdx.x; j < J; j += NT)
            if (field[static_cast<long long>(i) * fieldStride + j] != 0.f || mu[static_cast<long long>(i) * J + j] != 0.f)
            {
                iMin = min(iMin, i);
                iMax = max(iMax, i);
                jMin = min(jMin, j);
                jMax = max(jMax, j);
            }
    if (!__any_sync(0xffffffffu, iMax >= 0)) return;
    iMin = __reduce_min_sync(0xffffffffu, iMin);
    iMax = __reduce_max_sync(0xffffffffu, iMax);
    jMin = __reduce_min_sync(0xffffffffu, jMin);
    jMax = __reduce_max_sync(0xffffffffu, jMax);
    if ((threadIdx.x & 31) == 0)
    {
        atomicMin(box + 0, iMin);
        atomicMax(box + 1, iMax);
        atomicMin(box + 2, jMin);
        atomicMax(box + 3, jMax);
    }
}

// rhs = density * field (float product, viscositymodel.cpp:140-150); x = 0; r = rhs; p = r / diag;
// rhsNorm2 = |rhs|^2, absNew = r.p; zero / already-converged exits of Eigen's conjugate_gradient.
__global__ void __launch_bounds__(NT) viscInitKernel(ViscArgs a, const float *__restrict__ field, int fieldStride, float density)
{
    __shared__ double scratch[16];
    __shared__ int isLast;
    double s0 = 0.0, s1 = 0.0;
    const ViscBox bx = viscBox(a, 2);
    for (long long t = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; t < bx.cells(); t += static_cast<long long>(gridDim.x) * NT)
    {
        int i, j;
        viscCell(bx, t, &i, &j);
        const long long n = static_cast<long long>(i) * a.J + j;
        const double b = static_cast<double>(__fmul_rn(density, field[static_cast<long long>(i) * fieldStride + j]));
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
    const ViscBox bx = viscBox(a, 1);
    for (long long t = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; t < bx.cells(); t += static_cast<long long>(gridDim.x) * NT)
    {
        int i, j;
        viscCell(bx, t, &i, &j);
        const long long n = static_cast<long long>(i) * a.J + j;
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
    const ViscBox bx = viscBox(a, 1);
    for (long long t = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; t < bx.cells(); t += static_cast<long long>(gridDim.x) * NT)
    {
        int i, j;
        viscCell(bx, t, &i, &j);
        const long long n = static_cast<long long>(i) * a.J + j;
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
    const ViscBox bx = viscBox(a, 1);
    for (long long t = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; t < bx.cells(); t += static_cast<long long>(gridDim.x) * NT)
    {
        int i, j;
        viscCell(bx, t, &i, &j);
        const long long n = static_cast<long long>(i) * a.J + j;
        a.p[n] = a.z[n] + beta * a.p[n];
    }
}

// field(i, j) = x / density (viscositymodel.cpp:152-162)
__global__ void __launch_bounds__(NT) viscWriteBackKernel(ViscArgs a, float *__restrict__ field, int fieldStride, float density)
{
    const ViscBox bx = viscBox(a, 1);
    for (long long t = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; t < bx.cells(); t += static_cast<long long>(gridDim.x) * NT)
    {
        int i, j;
        viscCell(bx, t, &i, &j);
        field[static_cast<long long>(i) * fieldStride + j] = static_cast<float>(a.x[static_cast<long long>(i) * a.J + j] / static_cast<double>(density));
    }
}

// ------------------------------------------------------------------ HeavyViscosityModel (viscositymodel.cpp:164-470)
// One coupled system over all U and all V samples (n = (I+1)J + I(J+1) unknowns, U block first). getMatrix accumulates
// (+=) a non-symmetric matrix; Eigen's ConjugateGradient<.., Upper> reads the diagonal and the strict upper triangle and
// mirrors it, so the operator every solve sees is S = D + triu(A) + triu(A)^T. With
//     T(a,b) = scaleTwoDt * viscosity.getAt(a,b)                 scaleTwoDt = 2 dt / dx^2      (float arithmetic)
//     C(a,b) = scaleTwoDx * viscosity.interpolateAt(a-1/2, b-1/2) scaleTwoDx = dt / (2 dx^2)
// and uval / vval = "the sample exists", the strict upper entries are (viscositymodel.cpp:232-385)
//   row U(i,j): U(i+1,j): -T(i,j) | if U(i,j+1),V(i,j+1),V(i-1,j+1) exist: U(i,j+1): -C(i,j+1), V(i,j+1): -C(i,j+1),
//               V(i-1,j+1): +C(i,j+1) | if U(i,j-1),V(i,j),V(i-1,j) exist: V(i,j): +C(i,j), V(i-1,j): -C(i,j)
//   row V(i,j): V(i,j+1): -T(i,j) | if U(i+1,j),V(i+1,j-1),V(i+1,j) exist: V(i+1,j): -C(i+1,j)
// (every V-row entry that points into the U block -- including the one the reference addresses with a V index,
// :364,377-379 -- lies below the diagonal and is never read). The diagonal takes every += in code order. The formulas
// below were checked entry by entry against a scipy assembly of the reference's loops, which in turn reproduces the
// oracle's result bit for bit (tests/test_stages_gpu.py::test_heavy_viscosity_stage).
struct HeavyArgs
{
    const float *mu;       // viscosity grid (I x J, OOB_EXTEND, sample offset 1/2)
    const float *corner;   // C(a,b), (I+2) x (J+2)
    int I, J;
    long long NU, n;
    float s2dt, rho;
    double *diag, *x, *r, *p, *tmp, *z;
    double *partials;
    ViscScalars *sc;
    long long maxIters;
};

__device__ __forceinline__ bool uval(const HeavyArgs &a, int i, int j) { return i >= 0 && i <= a.I && j >= 0 && j < a.J; }
__device__ __forceinline__ bool vval(const HeavyArgs &a, int i, int j) { return i >= 0 && i < a.I && j >= 0 && j <= a.J; }
__device__ __forceinline__ double heavyT(const HeavyArgs &a, int i, int j)
{
    i = i < 0 ? 0 : (i > a.I - 1 ? a.I - 1 : i);
    j = j < 0 ? 0 : (j > a.J - 1 ? a.J - 1 : j);
    return static_cast<double>(__fmul_rn(a.s2dt, a.mu[static_cast<long long>(i) * a.J + j]));
}
__device__ __forceinline__ double heavyC(const HeavyArgs &a, int i, int j) { return static_cast<double>(a.corner[static_cast<long long>(i) * (a.J + 2) + j]); }

__global__ void __launch_bounds__(NT) heavyCornerKernel(GridView visc, int I, int J, float s2dx, float *__restrict__ corner)
{
    const long long t = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (t >= static_cast<long long>(I + 2) * (J + 2)) return;
    const int a = static_cast<int>(t / (J + 2)), b = static_cast<int>(t - static_cast<long long>(a) * (J + 2));
    corner[t] = __fmul_rn(s2dx, gridLerp(visc, __fsub_rn(static_cast<float>(a), 0.5f), __fsub_rn(static_cast<float>(b), 0.5f)));
}

__device__ __forceinline__ double heavyDiag(const HeavyArgs &a, long long row)
{
    double d = static_cast<double>(a.rho);
    if (row < a.NU)
    {
        const int i = rowOfCell(row, a.J), j = static_cast<int>(row - static_cast<long long>(i) * a.J);
        if (uval(a, i - 1, j)) d += heavyT(a, i - 1, j);
        if (uval(a, i + 1, j)) d += heavyT(a, i, j);
        if (uval(a, i, j - 1) && vval(a, i, j) && vval(a, i - 1, j)) d += heavyC(a, i, j);
        if (uval(a, i, j + 1) && vval(a, i, j + 1) && vval(a, i - 1, j + 1)) d += heavyC(a, i, j + 1);
    }
    else
    {
        const long long m = row - a.NU;
        const int i = static_cast<int>(m / (a.J + 1)), j = static_cast<int>(m - static_cast<long long>(i) * (a.J + 1));
        if (vval(a, i, j - 1)) d += heavyT(a, i, j - 1);
        if (vval(a, i, j + 1)) d += heavyT(a, i, j);
        if (uval(a, i, j) && uval(a, i, j - 1) && vval(a, i - 1, j)) d += heavyC(a, i, j);
        if (uval(a, i + 1, j) && vval(a, i + 1, j - 1) && vval(a, i + 1, j)) d += heavyC(a, i + 1, j);
    }
    return d;
}

// (S v)[row]: diagonal, the row's own strict upper entries, and the entries other rows hold in this column
__device__ __forceinline__ double heavyApplyRow(const HeavyArgs &a, const double *__restrict__ v, long long row)
{
    const long long J = a.J, J1 = a.J + 1;
    const double *vU = v, *vV = v + a.NU;
    double y = a.diag[row] * v[row];
    if (row < a.NU)
    {
        const int i = rowOfCell(row, J), j = static_cast<int>(row - static_cast<long long>(i) * J);
        const bool jm = uval(a, i, j - 1) && vval(a, i, j) && vval(a, i - 1, j);
        const bool jp = uval(a, i, j + 1) && vval(a, i, j + 1) && vval(a, i - 1, j + 1);
        if (uval(a, i + 1, j)) y += -heavyT(a, i, j) * vU[(i + 1) * J + j];
        if (jp)
        {
            const double c = heavyC(a, i, j + 1);
            y += -c * vU[i * J + j + 1];
            y += -c * vV[i * J1 + j + 1];
            y += c * vV[(i - 1) * J1 + j + 1];
        }
        if (jm)
        {
            const double c = heavyC(a, i, j);
            y += c * vV[i * J1 + j];
            y += -c * vV[(i - 1) * J1 + j];
        }
        if (uval(a, i - 1, j)) y += -heavyT(a, i - 1, j) * vU[(i - 1) * J + j];  // U(i-1,j) -> U(i,j)
        if (jm) y += -heavyC(a, i, j) * vU[i * J + j - 1];                       // U(i,j-1) -> U(i,j), same existence test
    }
    else
    {
        const long long m = row - a.NU;
        const int i = rowOfCell(m, J1), j = static_cast<int>(m - static_cast<long long>(i) * J1);
        if (vval(a, i, j + 1)) y += -heavyT(a, i, j) * vV[i * J1 + j + 1];
        if (uval(a, i + 1, j) && vval(a, i + 1, j - 1) && vval(a, i + 1, j)) y += -heavyC(a, i + 1, j) * vV[(i + 1) * J1 + j];
        if (vval(a, i, j - 1)) y += -heavyT(a, i, j - 1) * vV[i * J1 + j - 1];                                   // V(i,j-1)
        if (vval(a, i - 1, j) && uval(a, i, j) && vval(a, i, j - 1)) y += -heavyC(a, i, j) * vV[(i - 1) * J1 + j];  // V(i-1,j)
        if (uval(a, i, j) && uval(a, i, j - 1) && vval(a, i - 1, j)) y += heavyC(a, i, j) * vU[i * J + j];                    // U(i,j)
        if (uval(a, i + 1, j) && uval(a, i + 1, j - 1) && vval(a, i + 1, j)) y += -heavyC(a, i + 1, j) * vU[(i + 1) * J + j]; // U(i+1,j)
        if (uval(a, i, j - 1) && uval(a, i, j) && vval(a, i - 1, j)) y += -heavyC(a, i, j) * vU[i * J + j - 1];               // U(i,j-1)
        if (uval(a, i + 1, j - 1) && uval(a, i + 1, j) && vval(a, i + 1, j)) y += heavyC(a, i + 1, j) * vU[(i + 1) * J + j - 1];  // U(i+1,j-1)
    }
    return y;
}

// getRhs (viscositymodel.cpp:400-432): rhs = density * velocity (float product) for every U and V sample; then the
// start of Eigen's conjugate_gradient as in viscInitKernel
__global__ void __launch_bounds__(NT) heavyInitKernel(HeavyArgs a, const float *__restrict__ U, const float *__restrict__ V)
{
    __shared__ double scratch[16];
    __shared__ int isLast;
    double s0 = 0.0, s1 = 0.0;
    for (long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; n < a.n; n += static_cast<long long>(gridDim.x) * NT)
    {
        const double b = static_cast<double>(__fmul_rn(a.rho, n < a.NU ? U[n] : V[n - a.NU]));
        const double d = heavyDiag(a, n);
        a.diag[n] = d;
        a.x[n] = 0.0;
        a.r[n] = b;
        const double p = b * (1.0 / d);
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

__global__ void __launch_bounds__(NT) heavyApplyKernel(HeavyArgs a)
{
    __shared__ double scratch[16];
    __shared__ int isLast;
    if (a.sc->done) return;
    double s0 = 0.0, s1 = 0.0;
    for (long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; n < a.n; n += static_cast<long long>(gridDim.x) * NT)
    {
        const double y = heavyApplyRow(a, a.p, n);
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

__global__ void __launch_bounds__(NT) heavyUpdateKernel(HeavyArgs a)
{
    __shared__ double scratch[16];
    __shared__ int isLast;
    if (a.sc->done) return;
    const double alpha = a.sc->alpha;
    double s0 = 0.0, s1 = 0.0;
    for (long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; n < a.n; n += static_cast<long long>(gridDim.x) * NT)
    {
        a.x[n] += alpha * a.p[n];
        const double r = a.r[n] - alpha * a.tmp[n];
        a.r[n] = r;
        const double z = r * (1.0 / a.diag[n]);
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
            sc->done = 1;
        else
        {
            sc->beta = s1 / sc->absNew;
            sc->absNew = s1;
            sc->iter++;
            if (sc->iter >= a.maxIters) sc->done = 1;
        }
    }
}

__global__ void __launch_bounds__(NT) heavyDirectionKernel(HeavyArgs a)
{
    if (a.sc->done) return;
    const double beta = a.sc->beta;
    for (long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; n < a.n; n += static_cast<long long>(gridDim.x) * NT)
        a.p[n] = a.z[n] + beta * a.p[n];
}

// applyResult (viscositymodel.cpp:434-470): every U and V sample takes the solution (no division by the density)
__global__ void __launch_bounds__(NT) heavyWriteBackKernel(HeavyArgs a, float *__restrict__ U, float *__restrict__ V)
{
    for (long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; n < a.n; n += static_cast<long long>(gridDim.x) * NT)
    {
        const float v = static_cast<float>(a.x[n]);
        if (n < a.NU)
            U[n] = v;
        else
            V[n - a.NU] = v;
    }
}
}  // namespace

// One Eigen-style solve for `field` (U: stride J; V: stride J + 1). *iters = Eigen's iterations(); returns
// FS2D_OK with *failed = 1 when the solver did not reach the tolerance (the reference prints and gives up).
static int viscSolve(Ctx *ctx, ViscArgs &a, float *field, int stride, float density, int *iters, int *failed, int rowLo, int rowHi)
{
    cudaStream_t st = ctx->stream;
    const int blocks = std::min<long long>(divUp(ctx->N, NT), static_cast<long long>(ctx->smCount) * 8);
    int *box = reinterpret_cast<int *>(static_cast<unsigned char *>(ctx->viscScalars) + 128);
    viscBoxResetKernel<<<1, 1, 0, st>>>(box);
    // row slabs: only the rows [rowLo, rowHi) gathered from all ranks are exact copies; every row outside them has a zero
    // right-hand side and a zero viscosity on the rank that owns it (that is how the band was chosen)
    viscBoxKernel<<<std::max(1, std::min(rowHi - rowLo, ctx->smCount * 8)), NT, 0, st>>>(field, stride, a.mu, rowLo, rowHi, ctx->J, box);
    a.box = box;
    viscInitKernel<<<blocks, NT, 0, st>>>(a, field, stride, density);
    ctx->launches += 3;
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

GridView viscosityView(const Ctx *c);

// HeavyViscosityModel::apply (viscositymodel.cpp:164-200)
static int gridViscosityHeavy(Ctx *ctx, int *iters)
{
    cudaStream_t st = ctx->stream;
    const long long n = ctx->NU + ctx->NV;
    const long long cornerCount = static_cast<long long>(ctx->I + 2) * (ctx->J + 2);
    if (!ctx->heavyBuf)
    {
        const size_t bytes = sizeof(double) * 6 * static_cast<size_t>(n) + sizeof(float) * static_cast<size_t>(cornerCount) + 256;
        FS2D_CUDA(cudaMalloc(&ctx->heavyBuf, bytes));
    }
    if (!ctx->viscScalars) FS2D_CUDA(cudaMalloc(&ctx->viscScalars, 256));
    FS2D_CUDA(cudaMemsetAsync(ctx->viscScalars, 0, 256, st));
    double *base = static_cast<double *>(ctx->heavyBuf);
    HeavyArgs a;
    a.mu = ctx->viscosity;
    a.I = ctx->I;
    a.J = ctx->J;
    a.NU = ctx->NU;
    a.n = n;
    a.diag = base;
    a.x = base + n;
    a.r = base + 2 * n;
    a.p = base + 3 * n;
    a.tmp = base + 4 * n;
    a.z = base + 5 * n;
    float *corner = reinterpret_cast<float *>(base + 6 * n);
    a.corner = corner;
    a.partials = ctx->partials;
    a.sc = static_cast<ViscScalars *>(ctx->viscScalars);
    a.maxIters = 2 * n;
    // apply(..., float dt, float dx, float density): scaleTwoDt = 2*dt / (dx*dx), scaleTwoDx = dt / (2*dx*dx) in float
    const float dt = ctx->stepDt, dx = static_cast<float>(ctx->p.dx);
    a.rho = static_cast<float>(ctx->p.fluid_density);
    a.s2dt = 2 * dt / (dx * dx);
    const float s2dx = dt / (2 * dx * dx);
    const int blocks = std::min<long long>(divUp(n, NT), static_cast<long long>(ctx->smCount) * 8);
    heavyCornerKernel<<<divUp(cornerCount, NT), NT, 0, st>>>(viscosityView(ctx), ctx->I, ctx->J, s2dx, corner);
    heavyInitKernel<<<blocks, NT, 0, st>>>(a, ctx->U, ctx->V);
    ctx->launches += 2;
    ViscScalars sc;
    for (;;)
    {
        for (int k = 0; k < 8; k++)
        {
            heavyApplyKernel<<<blocks, NT, 0, st>>>(a);
            heavyUpdateKernel<<<blocks, NT, 0, st>>>(a);
            heavyDirectionKernel<<<blocks, NT, 0, st>>>(a);
        }
        ctx->launches += 24;
        FS2D_CUDA(cudaGetLastError());
        FS2D_CUDA(fs2dCopyToHost(ctx, &sc, a.sc, sizeof(sc)));
        if (sc.done) break;
    }
    const double err = sc.rhsNorm2 > 0.0 ? std::sqrt(sc.resNorm2 / sc.rhsNorm2) : 0.0;
    if (sc.rhsNorm2 > 0.0 && !(err <= 1e-4))
    {
        if (iters) *iters = -1;  // "Viscosity solver U solving failed!": nothing is applied
        return FS2D_OK;
    }
    heavyWriteBackKernel<<<blocks, NT, 0, st>>>(a, ctx->U, ctx->V);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    if (iters) *iters = sc.iter;
    return FS2D_OK;
}

int gridViscosity(Ctx *ctx, int *iters)
{
    int bandLo = 0, bandHi = ctx->I;  // rows the solve may look at (row slabs: the gathered band)
    if (ctx->slab.enabled && ctx->slab.world > 1)
    {
        // Row slabs: the solve is REPLICATED. Every rank pushes its rows of U, V, the viscosity grid and the material grid
        // to all the others (one collective) and then solves the whole system itself -- same kernels, same reduction
        // order, hence bit-identical to a single handle on every rank. The stage is a handful of iterations of grid-only
        // passes; distributing it would add two all-reduces and a halo exchange per iteration for a
        // stage that is ~2 ms at 4096^2.
        void *arr[4] = {ctx->U, ctx->V, ctx->viscosity, ctx->material};
        const size_t rb[4] = {sizeof(float) * ctx->J, sizeof(float) * (ctx->J + 1), sizeof(float) * ctx->J, static_cast<size_t>(ctx->J)};
        const int rt[4] = {ctx->I + 1, ctx->I, ctx->I, ctx->I};
        if (ctx->p.heavy_viscosity)
            FS2D_TRY(slabGatherMany(ctx, arr, rb, rt, 4));  // the coupled model runs on every sample: all rows
        else
        {
            // The light model only touches the box of cells with a non-zero right-hand side or viscosity, grown by two
            // cells (viscBox): only the band of rows around it travels -- 44 % of the rows of the 4096^2 dam break, and none
            // of the 2500 empty rows the first slab of a balanced split owns (the full gather cost 3.8 ms of NVLink
            // pushes per substep on 4 GPUs). Each rank looks at its own rows; the opening handshake carries the result.
            if (!ctx->viscScalars) FS2D_CUDA(cudaMalloc(&ctx->viscScalars, 256));
            int *box = reinterpret_cast<int *>(static_cast<unsigned char *>(ctx->viscScalars) + 128);
            const SlabRows own = slabOwn(ctx);
            viscBoxResetKernel<<<1, 1, 0, ctx->stream>>>(box);
            const int grid = std::max(1, std::min(own.hi - own.lo, ctx->smCount * 8));
            viscBoxKernel<<<grid, NT, 0, ctx->stream>>>(ctx->U, ctx->J, ctx->viscosity, own.lo, own.hi, ctx->J, box);
            viscBoxKernel<<<grid, NT, 0, ctx->stream>>>(ctx->V, ctx->J + 1, ctx->viscosity, own.lo, own.hi, ctx->J, box);
            ctx->launches += 3;
            int rows[2] = {0, -1};
            FS2D_CUDA(fs2dCopyToHost(ctx, rows, box, sizeof(rows)));
            FS2D_TRY(slabGatherBand(ctx, arr, rb, rt, 4, rows[0], rows[1], 3, &bandLo, &bandHi));
        }
    }
    if (ctx->p.heavy_viscosity) return gridViscosityHeavy(ctx, iters);
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
    FS2D_TRY(viscSolve(ctx, a, ctx->U, ctx->J, density, &itU, &failed, bandLo, bandHi));
    if (failed)
    {
        if (iters) *iters = -1;  // "Viscosity solver U solving failed!" -> return -1 before anything is applied
        return FS2D_OK;
    }
    viscWriteBackKernel<<<blocks, NT, 0, ctx->stream>>>(a, ctx->U, ctx->J, density);
    ctx->launches++;
    FS2D_TRY(viscSolve(ctx, a, ctx->V, ctx->J + 1, density, &itV, &failed, bandLo, bandHi));
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
