// Device-side restatement of the small numerics helpers every stage shares:
// Grid2d<T>::getAt / lerp (grid2d.h:122-143,187-216), simmath (mathfuncs.cpp:17-116),
// MaterialGrid neighbour predicates (materialgrid.cpp:130-164).
//
// Float arithmetic is written with explicit round-to-nearest intrinsics in the reference's
// evaluation order (no FMA contraction) so that positions, weights and the decisions that
// hang on them (floor(x), sdf < 0, w > 1e-9) reproduce the strict (-ffp-contract=off) oracle.
#ifndef FS2D_DEVICE_CUH
#define FS2D_DEVICE_CUH

#include "fs2d_internal.h"

#define FS2D_OOB_EXTEND 0
#define FS2D_OOB_CONST 1

__device__ __forceinline__ float fmulr(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float faddr(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float fsubr(float a, float b) { return __fsub_rn(a, b); }

// n / J for a linear cell index: one 32-bit division whenever the index allows it (every grid up to 65536^2 cells... in
// practice always); the 64-bit division the plain expression compiles to is a ~100-instruction routine and made the
// streaming grid kernels instruction-bound.
__device__ __forceinline__ int rowOfCell(long long n, int J)
{
    return static_cast<unsigned long long>(n) <= 0xffffffffull ? static_cast<int>(static_cast<unsigned int>(n) / static_cast<unsigned int>(J))
                                                                 : static_cast<int>(n / J);
}

__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// MaterialGrid is OOB_EXTEND (materialgrid.cpp:5-8): out-of-domain look-ups replicate the border.
__device__ __forceinline__ int8_t matAt(const int8_t *__restrict__ mat, int I, int J, int i, int j)
{
    i = clampi(i, 0, I - 1);
    j = clampi(j, 0, J - 1);
    return mat[static_cast<long long>(i) * J + j];
}

// MaterialGrid::nonsolidNeighborCount (materialgrid.cpp:137-140)
__device__ __forceinline__ unsigned int nonsolidCount(const int8_t *__restrict__ mat, int I, int J, int i, int j)
{
    return (matSolid(matAt(mat, I, J, i - 1, j)) ? 0u : 1u) + (matSolid(matAt(mat, I, J, i + 1, j)) ? 0u : 1u) +
           (matSolid(matAt(mat, I, J, i, j - 1)) ? 0u : 1u) + (matSolid(matAt(mat, I, J, i, j + 1)) ? 0u : 1u);
}

// Bin of a position: gridToBinIdx(Vec3) truncates the coordinates to ssize_t, then divides by the bin
// size 3 (markerparticlesystem.cpp:97-102,146-159).
__device__ __forceinline__ int2 positionBin(float2 p) { return make_int2(static_cast<int>(p.x) / 3, static_cast<int>(p.y) / 3); }

// Is a particle with storage code `mis` among the particles the reference visits for cell (ci, cj), i.e.
// is its storage bin inside the 3x3 block of bins around the cell's bin (binsForGridCell,
// markerparticlesystem.cpp:109-128)?
__device__ __forceinline__ bool storageVisible(unsigned int mis, float2 p, int ci, int cj)
{
    if (mis == FS2D_MIS_LOST) return false;
    const int2 pb = positionBin(p);
    const int di = pb.x + static_cast<int>(mis / 5u) - 2 - ci / 3;
    const int dj = pb.y + static_cast<int>(mis % 5u) - 2 - cj / 3;
    return di >= -1 && di <= 1 && dj >= -1 && dj <= 1;
}

__device__ __forceinline__ unsigned int storageCode(int di, int dj)
{
    if (di < -2 || di > 2 || dj < -2 || dj > 2) return FS2D_MIS_LOST;
    return static_cast<unsigned int>((di + 2) * 5 + (dj + 2));
}

// A float grid with the reference's out-of-bounds policy and sample offset.
struct GridView
{
    const float *data;
    int sizeI, sizeJ;
    float offX, offY;
    int oob;       // FS2D_OOB_EXTEND / FS2D_OOB_CONST
    float oobVal;
};

__host__ __device__ inline GridView makeView(const float *d, int sI, int sJ, float ox, float oy, int oob = FS2D_OOB_EXTEND,
                                             float oobVal = 0.f)
{
    GridView g;
    g.data = d;
    g.sizeI = sI;
    g.sizeJ = sJ;
    g.offX = ox;
    g.offY = oy;
    g.oob = oob;
    g.oobVal = oobVal;
    return g;
}

// Grid2d<T>::getAt(i, j) (grid2d.h:122-143)
__device__ __forceinline__ float gridAt(const GridView &g, int i, int j)
{
    if (g.oob == FS2D_OOB_EXTEND)
    {
        i = clampi(i, 0, g.sizeI - 1);
        j = clampi(j, 0, g.sizeJ - 1);
    }
    else if (i < 0 || i >= g.sizeI || j < 0 || j >= g.sizeJ)
    {
        return g.oobVal;
    }
    return __ldg(g.data + static_cast<long long>(i) * g.sizeJ + j);
}

// simmath::lerp (mathfuncs.cpp:27-30)
__device__ __forceinline__ float lerpf(float a, float b, float f) { return faddr(fmulr(a, fsubr(1.0f, f)), fmulr(b, f)); }

// Grid2d::lerp (grid2d.h:187-216): cell-centred bilinear interpolation with |frac - 1/2| factors.
__device__ __forceinline__ float gridLerp(const GridView &g, float i, float j)
{
    i = faddr(i, g.offX);
    j = faddr(j, g.offY);
    i = fminf(fmaxf(i, 0.f), static_cast<float>(g.sizeI - 1));
    j = fminf(fmaxf(j, 0.f), static_cast<float>(g.sizeJ - 1));
    // i, j are clamped to [0, size - 1] above, so floor = truncation and one 32-bit conversion each way gives the
    // reference's std::floor / static_cast<ssize_t> values (the 64-bit conversions and the two roundings ran on the
    // 16-lane XU pipe: 8 of them per sample, 9 samples per advected particle)
    const int ci = __float2int_rz(i), cj = __float2int_rz(j);
    const float fi = fsubr(i, __int2float_rn(ci));
    const float fj = fsubr(j, __int2float_rn(cj));
    const int ni = fi >= 0.5f ? ci + 1 : ci - 1;
    const int nj = fj >= 0.5f ? cj + 1 : cj - 1;
    const float iF = fi < 0.5f ? fsubr(0.5f, fi) : fsubr(fi, 0.5f);
    const float jF = fj < 0.5f ? fsubr(0.5f, fj) : fsubr(fj, 0.5f);
    const float v1 = lerpf(gridAt(g, ci, cj), gridAt(g, ni, cj), iF);
    const float v2 = lerpf(gridAt(g, ci, nj), gridAt(g, ni, nj), iF);
    return lerpf(v1, v2, jF);
}

// simmath::bSpline (mathfuncs.cpp:32-37). The reference evaluates both pieces, multiplies each by its 0 / 1 range flag
// and adds them: (0.75 - v^2)*[v < 0.5] + 0.5*(1.5 - v)^2*[0.5 <= v < 1.5]. For finite v the piece with flag 0
// contributes +-0 and the sum IS the other piece (x + +-0 = x; both flags 0: -0 + +0 = +0), so selecting the piece gives
// the same bits with half the arithmetic -- this function runs four times per particle-cell pair of the P2G gather.
__device__ __forceinline__ float bSpline(float v)
{
    v = fabsf(v);
    const float a = fsubr(0.75f, fmulr(v, v));
    const float h = fsubr(1.5f, v);
    const float b = fmulr(fmulr(0.5f, h), h);
    return v < 0.5f ? a : (v < 1.5f ? b : 0.f);  // selects, no branch: the gather loop stays convergent
}

// simmath::quadraticBSpline (mathfuncs.cpp:39-43); bSpline(0) = 0.75
__device__ __forceinline__ float quadraticBSpline(float x, float y) { return fmulr(fmulr(bSpline(x), bSpline(y)), 0.75f); }

// simmath::linearHat / bilinearHat (mathfuncs.cpp:100-116); linearHat(0) = 1
__device__ __forceinline__ float linearHat(float v)
{
    if (v >= 0.f && v <= 1.f) return fsubr(1.f, v);
    if (v < 0.f && v >= -1.f) return faddr(1.f, v);
    return 0.f;
}
__device__ __forceinline__ float bilinearHat(float x, float y) { return fmulr(fmulr(linearHat(x), linearHat(y)), 1.f); }

// StaggeredVelocityGrid::velocityAt (staggeredvelocitygrid.cpp:158-162); U offset (1/2, 0), V offset (0, 1/2)
struct VelocityView
{
    GridView u, v;
};

__host__ __device__ inline VelocityView makeVelocityView(const float *U, const float *V, int I, int J)
{
    VelocityView w;
    w.u = makeView(U, I + 1, J, 0.5f, 0.f);
    w.v = makeView(V, I, J + 1, 0.f, 0.5f);
    return w;
}

__device__ __forceinline__ float2 velocityAt(const VelocityView &w, float x, float y)
{
    return make_float2(gridLerp(w.u, x, y), gridLerp(w.v, x, y));
}

// FlipSolver::rk4Integrate (flipsolver2d.cpp:1195-1203)
__device__ __forceinline__ float2 rk4(const VelocityView &w, float2 p, float dt)
{
    float2 v = velocityAt(w, p.x, p.y);
    const float2 k1 = make_float2(fmulr(dt, v.x), fmulr(dt, v.y));
    v = velocityAt(w, faddr(p.x, fmulr(0.5f, k1.x)), faddr(p.y, fmulr(0.5f, k1.y)));
    const float2 k2 = make_float2(fmulr(dt, v.x), fmulr(dt, v.y));
    v = velocityAt(w, faddr(p.x, fmulr(0.5f, k2.x)), faddr(p.y, fmulr(0.5f, k2.y)));
    const float2 k3 = make_float2(fmulr(dt, v.x), fmulr(dt, v.y));
    v = velocityAt(w, faddr(p.x, k3.x), faddr(p.y, k3.y));
    const float2 k4 = make_float2(fmulr(dt, v.x), fmulr(dt, v.y));
    // currentPosition + (1/6)*(k1 + 2*k2 + 2*k3 + k4), left to right
    const float sx = faddr(faddr(faddr(k1.x, fmulr(2.f, k2.x)), fmulr(2.f, k3.x)), k4.x);
    const float sy = faddr(faddr(faddr(k1.y, fmulr(2.f, k2.y)), fmulr(2.f, k3.y)), k4.y);
    const float sixth = 1.0f / 6.0f;
    return make_float2(faddr(p.x, fmulr(sixth, sx)), faddr(p.y, fmulr(sixth, sy)));
}

#endif
