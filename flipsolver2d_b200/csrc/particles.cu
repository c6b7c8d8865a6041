// Kernel groups 1 and 5: marker-particle storage, binning / counting sort by cell, RK4
// advection, G2P PIC/FLIP blend, per-cell counting + cap, reseeding.
//
// The reference keeps particles in per-bin std::vectors (3x3-cell ParticleBin,
// markerparticlesystem.h:59-149) and moves them between bins with an O(#bins x #moved) scan
// (markerparticlesystem.cpp:260-269). Here all particles live in one SoA (float2 pos, float2
// vel, K float property columns) kept sorted by CELL (key = floor(x)*J + floor(y)) with a
// cellStart[N+1] table; the reference's bin of a particle is (cell_i/3, cell_j/3), so bin
// membership/counts derive from the key. Inside a cell particles are ordered by their
// position bits, which makes the device order a pure function of the particle set
// (run-to-run and decomposition independent).
#include <algorithm>
#include <cfloat>

#include "fs2d_device.cuh"
#include "fs2d_internal.h"

namespace
{
constexpr int NT = 256;

inline int gridFor(int64_t n) { return std::max(1, divUp(n, NT)); }

// ------------------------------------------------------------------ exclusive scan (int32)
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = 1024 * SCAN_ITEMS;

__global__ void __launch_bounds__(1024) scanTileKernel(const int32_t *__restrict__ in, int32_t *__restrict__ out,
                                                       int32_t *__restrict__ tileSums, long long n)
{
    __shared__ int32_t warpSums[32];
    const long long base = blockIdx.x * static_cast<long long>(SCAN_TILE) + threadIdx.x * SCAN_ITEMS;
    int32_t v[SCAN_ITEMS];
    int32_t sum = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
    {
        v[k] = (base + k < n) ? in[base + k] : 0;
        sum += v[k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t incl = sum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1)
    {
        int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) warpSums[warp] = incl;
    __syncthreads();
    if (warp == 0)
    {
        int32_t w = warpSums[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            int32_t t = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += t;
        }
        warpSums[lane] = w;
    }
    __syncthreads();
    int32_t excl = incl - sum + (warp > 0 ? warpSums[warp - 1] : 0);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
    {
        if (base + k < n) out[base + k] = excl;
        excl += v[k];
    }
    if (threadIdx.x == 1023) tileSums[blockIdx.x] = excl;
}

__global__ void __launch_bounds__(1024) scanSumsKernel(int32_t *tileSums, int tiles)
{
    __shared__ int32_t warpSums[32];
    __shared__ int32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < tiles; base += 1024)
    {
        const int idx = base + threadIdx.x;
        const int32_t v = idx < tiles ? tileSums[idx] : 0;
        int32_t incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warpSums[warp] = incl;
        __syncthreads();
        if (warp == 0)
        {
            int32_t w = warpSums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                int32_t t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warpSums[lane] = w;
        }
        __syncthreads();
        const int32_t excl = incl - v + (warp > 0 ? warpSums[warp - 1] : 0) + carry;
        if (idx < tiles) tileSums[idx] = excl;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
}

__global__ void __launch_bounds__(1024) scanAddKernel(int32_t *__restrict__ out, const int32_t *__restrict__ tileSums,
                                                      long long n)
{
    const int32_t add = tileSums[blockIdx.x];
    const long long base = blockIdx.x * static_cast<long long>(SCAN_TILE) + threadIdx.x * SCAN_ITEMS;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; k++)
        if (base + k < n) out[base + k] += add;
}

// out[0..n) = exclusive scan of in[0..n); in and out may alias.
void exclusiveScan(Ctx *ctx, const int32_t *in, int32_t *out, int64_t n)
{
    const int tiles = divUp(n, SCAN_TILE);
    scanTileKernel<<<tiles, 1024, 0, ctx->stream>>>(in, out, ctx->scanBlock, n);
    scanSumsKernel<<<1, 1024, 0, ctx->stream>>>(ctx->scanBlock, tiles);
    scanAddKernel<<<tiles, 1024, 0, ctx->stream>>>(out, ctx->scanBlock, n);
    ctx->launches += 3;
}

// ------------------------------------------------------------------ sort by cell
// Cell key of an alive particle. advectThread kills everything whose floor() leaves the grid
// (flipsolver2d.cpp:323-330), so alive particles always map to a valid cell; the clamp only
// guards state uploaded by a caller.
__device__ __forceinline__ uint32_t cellKey(float2 p, int I, int J)
{
    const int i = clampi(static_cast<int>(floorf(p.x)), 0, I - 1);
    const int j = clampi(static_cast<int>(floorf(p.y)), 0, J - 1);
    return static_cast<uint32_t>(i) * static_cast<uint32_t>(J) + static_cast<uint32_t>(j);
}

// Slab mode: a rank updates only the particles it OWNS -- those whose cell row at the last sort (key / J) lies in
// its rows; the others are ghost copies of a neighbour's particles. own.key == nullptr: everything is owned.
struct OwnedRows
{
    const uint32_t *key;
    uint32_t J;
    int rowBegin, rowEnd;
};
__device__ __forceinline__ bool ownedParticle(const OwnedRows &o, long long p)
{
    if (!o.key) return true;
    const int r = static_cast<int>(o.key[p] / o.J);
    return r >= o.rowBegin && r < o.rowEnd;
}

__global__ void __launch_bounds__(NT) cellKeyKernel(const float2 *__restrict__ pos, long long begin, long long end, int I, int J,
                                                    uint32_t *__restrict__ key);

__global__ void __launch_bounds__(NT) histogramKernel(const float2 *__restrict__ pos, const uint8_t *__restrict__ dead,
                                                      long long count, int I, int J, uint32_t cLo, uint32_t cHi,
                                                      uint32_t *__restrict__ key, int32_t *__restrict__ cellCount)
{
    const long long p = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (p >= count) return;
    if (dead[p])
    {
        key[p] = 0xFFFFFFFFu;
        return;
    }
    const uint32_t c = cellKey(pos[p], I, J);
    if (c < cLo || c >= cHi)  // slab mode: outside the rows this rank keeps (cannot happen after a particle exchange)
    {
        key[p] = 0xFFFFFFFFu;
        return;
    }
    key[p] = c;
    atomicAdd(cellCount + c, 1);
}

__global__ void __launch_bounds__(NT) scatterKernel(const uint32_t *__restrict__ key, long long count,
                                                    const int32_t *__restrict__ cellStart, int32_t *__restrict__ cellFill,
                                                    uint32_t *__restrict__ perm)
{
    const long long p = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (p >= count) return;
    const uint32_t c = key[p];
    if (c == 0xFFFFFFFFu) return;
    const int32_t slot = cellStart[c] + atomicAdd(cellFill + c, 1);
    perm[slot] = static_cast<uint32_t>(p);
}

__device__ __forceinline__ bool particleLess(const float2 *__restrict__ pos, uint32_t a, uint32_t b)
{
    const float2 pa = pos[a], pb = pos[b];
    if (pa.x != pb.x) return pa.x < pb.x;
    if (pa.y != pb.y) return pa.y < pb.y;
    return a < b;
}

// Canonical order inside each cell: ascending (x, y, old index). One thread per cell; cells hold
// O(particlesPerCell) entries so an insertion sort on the permutation is enough.
__global__ void __launch_bounds__(NT) cellOrderKernel(const float2 *__restrict__ pos, const int32_t *__restrict__ cellStart,
                                                      long long cBegin, long long N, uint32_t *__restrict__ perm)
{
    const long long c = cBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (c >= N) return;
    const int32_t b = cellStart[c], e = cellStart[c + 1];
    for (int32_t a = b + 1; a < e; a++)
    {
        const uint32_t v = perm[a];
        int32_t k = a - 1;
        while (k >= b && particleLess(pos, v, perm[k]))
        {
            perm[k + 1] = perm[k];
            k--;
        }
        perm[k + 1] = v;
    }
}

__global__ void __launch_bounds__(NT) gatherKernel(const uint32_t *__restrict__ perm, long long alive,
                                                   const float2 *__restrict__ posIn, const float2 *__restrict__ velIn,
                                                   const float *__restrict__ propsIn, long long capIn,
                                                   const uint32_t *__restrict__ keyIn, const uint8_t *__restrict__ misIn,
                                                   float2 *__restrict__ posOut, float2 *__restrict__ velOut,
                                                   float *__restrict__ propsOut, long long capOut,
                                                   uint32_t *__restrict__ keyOut, uint8_t *__restrict__ misOut, int numProps,
                                                   uint8_t *__restrict__ dead)
{
    const long long s = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (s >= alive) return;
    const uint32_t p = perm[s];
    posOut[s] = posIn[p];
    velOut[s] = velIn[p];
    keyOut[s] = keyIn[p];
    misOut[s] = misIn[p];
    for (int k = 0; k < numProps; k++) propsOut[k * capOut + s] = propsIn[k * capIn + p];
    dead[s] = 0;
}

// The property columns of a sort whose input columns were still on their way from the host (streamed upload).
__global__ void __launch_bounds__(NT) gatherPropsKernel(const uint32_t *__restrict__ perm, long long alive, const float *__restrict__ propsIn,
                                                        long long capIn, float *__restrict__ propsOut, long long capOut, int numProps)
{
    const long long s = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (s >= alive) return;
    const uint32_t p = perm[s];
    for (int k = 0; k < numProps; k++) propsOut[k * capOut + s] = propsIn[k * capIn + p];
}

__global__ void __launch_bounds__(NT) cellKeyKernel(const float2 *__restrict__ pos, long long begin, long long end, int I, int J,
                                                    uint32_t *__restrict__ key)
{
    const long long p = begin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (p < end) key[p] = cellKey(pos[p], I, J);
}

// ------------------------------------------------------------------ CFL velocity
// maxParticleVelocity (flipsolver2d.cpp:1560-1574): max of vx*vx + vy*vy, initial value FLT_MIN.
__global__ void __launch_bounds__(NT) maxVelocityKernel(const float2 *__restrict__ vel, const uint8_t *__restrict__ dead,
                                                        long long count, OwnedRows own, unsigned int *__restrict__ outBits)
{
    float m = 0.f;
    for (long long p = blockIdx.x * static_cast<long long>(NT) + threadIdx.x; p < count;
         p += static_cast<long long>(gridDim.x) * NT)
    {
        if (dead[p] || !ownedParticle(own, p)) continue;
        const float2 v = vel[p];
        const float s = faddr(fmulr(v.x, v.x), fmulr(v.y, v.y));
        m = fmaxf(m, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_down_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(outBits, __float_as_uint(m));  // non-negative floats order as uints
}

// ------------------------------------------------------------------ advection
// SdfGrid::closestSurfacePoint (sdfgrid.cpp:15-46)
__device__ float2 closestSurfacePoint(const GridView &sdf, float2 pos)
{
    float2 closest = pos;
    float value = gridLerp(sdf, pos.x, pos.y);
    // gradCenteredGrid takes ssize_t arguments: the float coordinates truncate (mathfuncs.cpp:87-98)
    int gi = static_cast<int>(pos.x), gj = static_cast<int>(pos.y);
    float gradX = fsubr(gridAt(sdf, gi + 1, gj), gridAt(sdf, gi - 1, gj)) / 2.f;
    float gradY = fsubr(gridAt(sdf, gi, gj + 1), gridAt(sdf, gi, gj - 1)) / 2.f;
    for (int it = 0; it < 100; it++)
    {
        float alpha = 1.f;
        for (int in = 0; in < 10; in++)
        {
            const float av = fmulr(alpha, value);
            const float2 q = make_float2(fsubr(closest.x, fmulr(av, gradX)), fsubr(closest.y, fmulr(av, gradY)));
            const float qv = gridLerp(sdf, q.x, q.y);
            if (fabsf(qv) < fabsf(value))
            {
                closest = q;
                value = qv;
                gi = static_cast<int>(q.x);
                gj = static_cast<int>(q.y);
                gradX = fsubr(gridAt(sdf, gi + 1, gj), gridAt(sdf, gi - 1, gj)) / 2.f;
                gradY = fsubr(gridAt(sdf, gi, gj + 1), gridAt(sdf, gi, gj - 1)) / 2.f;
                if (fabsf(value) < 1e-5f) return closest;
            }
            else
            {
                alpha = fmulr(alpha, 0.7f);
            }
        }
    }
    return closest;
}

// advectThread (flipsolver2d.cpp:305-338)
__global__ void __launch_bounds__(NT) advectKernel(float2 *__restrict__ pos, uint8_t *__restrict__ dead,
                                                   uint8_t *__restrict__ mis, long long count,
                                                   VelocityView vel, GridView solidSdf, const int8_t *__restrict__ mat,
                                                   int I, int J, float dt, OwnedRows own, unsigned long long *__restrict__ killed)
{
    const long long p = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (p >= count || dead[p] || !ownedParticle(own, p)) return;
    const float2 before = pos[p];
    float2 x = rk4(vel, before, dt);
    if (gridLerp(solidSdf, x.x, x.y) < 0.f) x = closestSurfacePoint(solidSdf, x);
    pos[p] = x;
    const float fi = floorf(x.x), fj = floorf(x.y);
    const bool inb = fi >= 0.f && fi < static_cast<float>(I) && fj >= 0.f && fj < static_cast<float>(J);
    if (!inb || matSink(mat[static_cast<long long>(fi) * J + static_cast<long long>(fj)]))
    {
        dead[p] = 1;
        atomicAdd(killed, 1ull);
        return;
    }
    // re-filed only when the bin of the POSITION changed (oldBinIdx is computed from the position, not
    // from the bin the particle is stored in)
    const int2 ob = positionBin(before), nb = positionBin(x);
    if (ob.x != nb.x || ob.y != nb.y) mis[p] = FS2D_MIS_HOME;
}

// particleUpdate (flipsolver2d.cpp:361-388) + smoke decay (flipsmokesolver.cpp:104-130)
__global__ void __launch_bounds__(NT) particleUpdateKernel(const float2 *__restrict__ pos, float2 *__restrict__ vel,
                                                           const uint8_t *__restrict__ dead, long long count,
                                                           VelocityView cur, VelocityView saved, float pic,
                                                           float *__restrict__ temperature, float *__restrict__ concentration,
                                                           float ambient, float tempFactor, float concFactor, OwnedRows own)
{
    const long long p = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (p >= count || dead[p] || !ownedParticle(own, p)) return;
    const float2 x = pos[p];
    const float2 oldV = velocityAt(saved, x.x, x.y);
    const float2 newV = velocityAt(cur, x.x, x.y);
    const float2 v = vel[p];
    const float om = fsubr(1.f, pic);
    vel[p] = make_float2(faddr(fmulr(pic, newV.x), fmulr(om, fsubr(faddr(v.x, newV.x), oldV.x))),
                         faddr(fmulr(pic, newV.y), fmulr(om, fsubr(faddr(v.y, newV.y), oldV.y))));
    if (temperature) temperature[p] = faddr(ambient, fmulr(fsubr(temperature[p], ambient), tempFactor));
    if (concentration) concentration[p] = fmulr(concentration[p], concFactor);
}

// adjustParticlesByDensityThread (flipsolver2d.cpp:261-303)
__global__ void __launch_bounds__(NT) densityAdjustKernel(float2 *__restrict__ pos, const uint8_t *__restrict__ dead,
                                                          uint8_t *__restrict__ mis, long long count, const double *__restrict__ pressure,
                                                          const int8_t *__restrict__ mat, int I, int J, float scale, OwnedRows own)
{
    const long long p = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (p >= count || dead[p] || !ownedParticle(own, p)) return;
    float2 x = pos[p];
    const int iCorr = clampi(static_cast<int>(fsubr(x.x, 0.5f)), 0, I);
    const int jCorr = clampi(static_cast<int>(fsubr(x.y, 0.5f)), 0, J);
    const int i = clampi(static_cast<int>(x.x), 0, I);
    const int j = clampi(static_cast<int>(x.y), 0, J);
    // linearIndex() is -1 out of range and the reference then reads out of bounds; every scene
    // with walls keeps these in range, the clamp below only avoids a fault.
    auto P = [&](int a, int b) -> float {
        a = clampi(a, 0, I - 1);
        b = clampi(b, 0, J - 1);
        return static_cast<float>(pressure[static_cast<long long>(a) * J + b]);
    };
    const float pCurI = P(iCorr, j), pCurJ = P(i, jCorr);
    float pI = P(iCorr + 1, j), pJ = P(i, jCorr + 1);
    if (matSolid(matAt(mat, I, J, iCorr + 1, j)) || matSolid(matAt(mat, I, J, iCorr, j))) pI = pCurI;
    if (matSolid(matAt(mat, I, J, i, jCorr + 1)) || matSolid(matAt(mat, I, J, i, jCorr))) pJ = pCurJ;
    const int2 ob = positionBin(x);
    x.x = faddr(x.x, fmulr(fsubr(pI, pCurI), scale));
    x.y = faddr(x.y, fmulr(fsubr(pJ, pCurJ), scale));
    pos[p] = x;
    // the particle stays filed where it was: keep (storage bin - position bin) up to date
    const int2 nb = positionBin(x);
    const unsigned int m = mis[p];
    if ((ob.x != nb.x || ob.y != nb.y) && m != FS2D_MIS_LOST)
        mis[p] = static_cast<uint8_t>(storageCode(rowOfCell(m, 5u) - 2 + ob.x - nb.x, static_cast<int>(m % 5u) - 2 + ob.y - nb.y));
}

// countParticles (flipsolver2d.cpp:1021-1051): per-cell count, particles beyond 2*ppc die.
__global__ void __launch_bounds__(NT) countCapKernel(const int32_t *__restrict__ cellStart, const uint8_t *__restrict__ mis,
                                                     long long cBegin, long long N, int cap, int32_t *__restrict__ counts,
                                                     uint8_t *__restrict__ dead, unsigned long long *__restrict__ killed)
{
    const long long c = cBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (c >= N) return;
    const int32_t b = cellStart[c], e = cellStart[c + 1];
    // only particles filed in this cell's own bin are seen by the reference's loop (binForGridIdx(i2d))
    int32_t n = 0, over = 0;
    for (int32_t s = b; s < e; s++)
    {
        if (mis[s] != FS2D_MIS_HOME) continue;
        if (n >= cap)
        {
            dead[s] = 1;
            over++;
        }
        else
        {
            n++;
        }
    }
    counts[c] = n;
    if (over > 0) atomicAdd(killed, static_cast<unsigned long long>(over));
}

// pruneNarrowBand (nbflipsolver.cpp:227-253)
__global__ void __launch_bounds__(NT) pruneBandKernel(const float2 *__restrict__ pos, uint8_t *__restrict__ dead,
                                                      long long count, GridView fluidSdf, const int8_t *__restrict__ mat,
                                                      const int32_t *__restrict__ counts, int I, int J, int ppc,
                                                      float narrowBand, float resamplingBand, OwnedRows own,
                                                      unsigned long long *__restrict__ killed)
{
    const long long p = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (p >= count || dead[p] || !ownedParticle(own, p)) return;  // ghost copies belong to a neighbour
    const float2 x = pos[p];
    const int i = clampi(static_cast<int>(x.x), 0, I - 1), j = clampi(static_cast<int>(x.y), 0, J - 1);
    const float sdf = gridLerp(fluidSdf, x.x, x.y);
    const bool src = matSource(mat[static_cast<long long>(i) * J + j]);
    bool kill = !src && sdf < narrowBand;
    if (!kill && (src || sdf < resamplingBand) && counts[static_cast<long long>(i) * J + j] > 2 * ppc) kill = true;
    if (kill)
    {
        dead[p] = 1;
        atomicAdd(killed, 1ull);
    }
}

// ------------------------------------------------------------------ reseeding
// Candidate count per cell. Water/smoke/fire: SOURCE cells short of ppc/2 (flipsolver2d.cpp:632-667,
// flipsmokesolver.cpp:155-176). NBFlip: SOURCE or band cells short of ppc (nbflipsolver.cpp:143-160).
__global__ void __launch_bounds__(NT) reseedPlanKernel(const int8_t *__restrict__ mat, const int32_t *__restrict__ counts,
                                                       const float *__restrict__ fluidSdf, long long cBegin, long long N, int ppc,
                                                       int nbflip, float narrowBand, float resamplingBand,
                                                       int32_t *__restrict__ want)
{
    const long long c = cBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (c > N) return;
    int32_t w = 0;
    if (c < N)
    {
        const bool src = matSource(mat[c]);
        if (nbflip)
        {
            const float s = fluidSdf[c];
            if (src || (s < resamplingBand && s > narrowBand)) w = ppc - counts[c];
        }
        else if (src)
        {
            w = ppc / 2 - counts[c];
        }
        if (w < 0) w = 0;
    }
    want[c] = w;
}

struct ReseedArgs
{
    int I, J, numProps;
    int simType;
    int viscosityProp, temperatureProp, concentrationProp, fuelProp, testProp;
    float narrowBand, resamplingBand;
    int numSources;
};

__global__ void __launch_bounds__(NT) reseedApplyKernel(const int32_t *__restrict__ offset, long long N,
                                                        const float *__restrict__ uniform, const int8_t *__restrict__ mat,
                                                        const int32_t *__restrict__ emitterId,
                                                        const fs2d_source *__restrict__ sources, VelocityView vel,
                                                        GridView fluidSdf, GridView viscosity, GridView temperature,
                                                        GridView concentration, GridView fuel, ReseedArgs a,
                                                        float2 *__restrict__ pos, float2 *__restrict__ velOut,
                                                        float *__restrict__ props, long long cap, long long base,
                                                        uint8_t *__restrict__ dead, uint8_t *__restrict__ mis,
                                                        uint32_t *__restrict__ key, long long cBegin,
                                                        unsigned long long *__restrict__ rejected)
{
    const long long c = cBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (c >= N) return;
    const int32_t b = offset[c], e = offset[c + 1];
    if (e == b) return;
    const int i = rowOfCell(c, a.J), j = static_cast<int>(c - static_cast<long long>(i) * a.J);
    const bool src = matSource(mat[c]);
    // a SOURCE cell without a (valid) emitter -- a material grid uploaded through the C ABI without its emitter grid --
    // behaves like an emitter with zero velocity / viscosity instead of indexing the table out of bounds
    const int emRaw = emitterId[c];
    const int em = (emRaw >= 0 && emRaw < a.numSources) ? emRaw : -1;
    for (int32_t k = b; k < e; k++)
    {
        // jitteredPosInCell (flipsolver2d.cpp:1013-1019): x drawn first, then y
        const float2 x = make_float2(faddr(static_cast<float>(i), uniform[2 * k]), faddr(static_cast<float>(j), uniform[2 * k + 1]));
        const long long slot = base + k;
        bool reject = false;
        float2 v = make_float2(0.f, 0.f);
        for (int q = 0; q < a.numProps; q++) props[q * cap + slot] = 0.f;
        if (a.simType == FS2D_SIM_NBFLIP)
        {
            const float newSdf = gridLerp(fluidSdf, x.x, x.y);
            if ((src && fluidSdf.data[c] > a.resamplingBand) || newSdf < a.narrowBand) reject = true;
            if (em != -1 && sources[em].transfer_velocity) v = velocityAt(vel, x.x, x.y);
            const float visc = em != -1 ? sources[em].viscosity : gridLerp(viscosity, x.x, x.y);
            if (a.viscosityProp >= 0) props[a.viscosityProp * cap + slot] = visc;
            if (a.testProp >= 0) props[a.testProp * cap + slot] = visc;
        }
        else
        {
            if (em != -1 && sources[em].transfer_velocity) v = velocityAt(vel, x.x, x.y);
            if (a.simType == FS2D_SIM_LIQUID)
            {
                if (a.viscosityProp >= 0) props[a.viscosityProp * cap + slot] = em != -1 ? sources[em].viscosity : 0.f;
            }
            else
            {
                // smoke: grid values at the new position (flipsmokesolver.cpp:188-193);
                // fire adds fuel (flipfiresolver.cpp reseedParticles)
                if (a.concentrationProp >= 0) props[a.concentrationProp * cap + slot] = gridLerp(concentration, x.x, x.y);
                if (a.temperatureProp >= 0) props[a.temperatureProp * cap + slot] = gridLerp(temperature, x.x, x.y);
                if (a.simType == FS2D_SIM_FIRE && a.fuelProp >= 0) props[a.fuelProp * cap + slot] = gridLerp(fuel, x.x, x.y);
            }
        }
        pos[slot] = x;
        velOut[slot] = v;
        dead[slot] = reject ? 1 : 0;
        mis[slot] = FS2D_MIS_HOME;
        key[slot] = static_cast<uint32_t>(c);
        if (reject) atomicAdd(rejected, 1ull);
    }
}

int fetchKilled(Ctx *ctx)
{
    if (!ctx->killedDirty) return FS2D_OK;
    unsigned long long k = 0;
    FS2D_CUDA(fs2dCopyToHost(ctx, &k, ctx->d_counter, sizeof(k)));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    FS2D_CUDA(cudaMemsetAsync(ctx->d_counter, 0, sizeof(k), ctx->stream));
    ctx->deadCount += static_cast<int64_t>(k);
    ctx->killedDirty = false;
    return FS2D_OK;
}
}  // namespace

GridView solidSdfView(const Ctx *c) { return makeView(c->solidSdf, c->I, c->J, 0.f, 0.f, FS2D_OOB_EXTEND); }
GridView fluidSdfView(const Ctx *c) { return makeView(c->fluidSdf, c->I, c->J, 0.f, 0.f, FS2D_OOB_EXTEND); }
GridView viscosityView(const Ctx *c) { return makeView(c->viscosity, c->I, c->J, 0.5f, 0.5f, FS2D_OOB_EXTEND); }
GridView temperatureView(const Ctx *c)
{
    // ctor: OOB_CONST(ambient), offset 0; after the first grid-mode advection the member is replaced by
    // a grid with OOB_EXTEND and offset (1/2, 1/2) (flipsmokesolver.cpp:11-12 vs :214-228)
    return c->smokeGridsAdvected ? makeView(c->temperature, c->I, c->J, 0.5f, 0.5f, FS2D_OOB_EXTEND)
                                 : makeView(c->temperature, c->I, c->J, 0.f, 0.f, FS2D_OOB_CONST, c->p.ambient_temperature);
}
GridView concentrationView(const Ctx *c)
{
    return c->smokeGridsAdvected ? makeView(c->concentration, c->I, c->J, 0.5f, 0.5f, FS2D_OOB_EXTEND)
                                 : makeView(c->concentration, c->I, c->J, 0.f, 0.f, FS2D_OOB_CONST, 0.f);
}
GridView fuelView(const Ctx *c)
{
    return c->smokeGridsAdvected ? makeView(c->fuel, c->I, c->J, 0.5f, 0.5f, FS2D_OOB_EXTEND)
                                 : makeView(c->fuel, c->I, c->J, 0.f, 0.f, FS2D_OOB_CONST, 0.f);
}

static OwnedRows ownedRows(const Ctx *ctx)
{
    OwnedRows o;
    const bool slab = ctx->slab.enabled && ctx->slab.world > 1;
    o.key = slab ? ctx->pb[ctx->cur].key : nullptr;
    o.J = static_cast<uint32_t>(ctx->J);
    o.rowBegin = ctx->slab.rowBegin;
    o.rowEnd = ctx->slab.rowEnd;
    return o;
}

// Cell keys of particles [begin, end) of the current buffer (uploads and appends; the sort recomputes them).
int particlesKeyRange(Ctx *ctx, int64_t begin, int64_t end)
{
    if (end <= begin) return FS2D_OK;
    cellKeyKernel<<<gridFor(end - begin), NT, 0, ctx->stream>>>(ctx->pb[ctx->cur].pos, begin, end, ctx->I, ctx->J, ctx->pb[ctx->cur].key);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int particlesGatherProps(Ctx *ctx, int from, int64_t count)
{
    if (count == 0 || ctx->p.num_properties == 0) return FS2D_OK;
    const ParticleBuffers &in = ctx->pb[from];
    ParticleBuffers &out = ctx->pb[ctx->cur];
    gatherPropsKernel<<<gridFor(count), NT, 0, ctx->stream>>>(ctx->perm, count, in.props, in.capacity, out.props, out.capacity,
                                                             ctx->p.num_properties);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int particlesReserve(Ctx *ctx, int64_t capacity)
{
    if (capacity <= ctx->pb[0].capacity) return FS2D_OK;
    FS2D_TRY(particleStreamSettleAll(ctx));  // the copies below move whole records
    const int64_t newCap = std::max<int64_t>(capacity + capacity / 4, 1024);
    const int K = ctx->p.num_properties;
    // Stream-ordered allocation and copies: growing the buffers never synchronises the device (another rank sharing
    // this GPU may have a kernel waiting for this rank's next launch) and never touches the legacy default stream.
    cudaStream_t st = ctx->stream;
    auto alloc = [&](void **p, size_t bytes) { return cudaMallocAsync(p, bytes, st); };
    for (int b = 0; b < 2; b++)
    {
        ParticleBuffers nb;
        nb.capacity = newCap;
        FS2D_CUDA(alloc(reinterpret_cast<void **>(&nb.pos), sizeof(float2) * newCap));
        FS2D_CUDA(alloc(reinterpret_cast<void **>(&nb.vel), sizeof(float2) * newCap));
        FS2D_CUDA(alloc(reinterpret_cast<void **>(&nb.props), sizeof(float) * newCap * std::max(K, 1)));
        FS2D_CUDA(alloc(reinterpret_cast<void **>(&nb.key), sizeof(uint32_t) * newCap));
        FS2D_CUDA(alloc(reinterpret_cast<void **>(&nb.mis), newCap));
        FS2D_CUDA(cudaMemsetAsync(nb.mis, FS2D_MIS_HOME, newCap, st));
        ParticleBuffers &ob = ctx->pb[b];
        if (b == ctx->cur && ctx->count > 0)
        {
            FS2D_CUDA(cudaMemcpyAsync(nb.pos, ob.pos, sizeof(float2) * ctx->count, cudaMemcpyDeviceToDevice, st));
            FS2D_CUDA(cudaMemcpyAsync(nb.vel, ob.vel, sizeof(float2) * ctx->count, cudaMemcpyDeviceToDevice, st));
            FS2D_CUDA(cudaMemcpyAsync(nb.key, ob.key, sizeof(uint32_t) * ctx->count, cudaMemcpyDeviceToDevice, st));
            FS2D_CUDA(cudaMemcpyAsync(nb.mis, ob.mis, ctx->count, cudaMemcpyDeviceToDevice, st));
            for (int k = 0; k < K; k++)
                FS2D_CUDA(cudaMemcpyAsync(nb.props + k * newCap, ob.props + k * ob.capacity, sizeof(float) * ctx->count,
                                          cudaMemcpyDeviceToDevice, st));
        }
        if (ob.pos) cudaFreeAsync(ob.pos, st);
        if (ob.vel) cudaFreeAsync(ob.vel, st);
        if (ob.props) cudaFreeAsync(ob.props, st);
        if (ob.key) cudaFreeAsync(ob.key, st);
        if (ob.mis) cudaFreeAsync(ob.mis, st);
        ob = nb;
    }
    uint8_t *nd = nullptr;
    uint32_t *np = nullptr;
    FS2D_CUDA(alloc(reinterpret_cast<void **>(&nd), newCap));
    FS2D_CUDA(cudaMemsetAsync(nd, 0, newCap, st));
    FS2D_CUDA(alloc(reinterpret_cast<void **>(&np), sizeof(uint32_t) * newCap));
    if (ctx->dead && ctx->count > 0) FS2D_CUDA(cudaMemcpyAsync(nd, ctx->dead, ctx->count, cudaMemcpyDeviceToDevice, st));
    if (ctx->dead) cudaFreeAsync(ctx->dead, st);
    if (ctx->perm) cudaFreeAsync(ctx->perm, st);
    ctx->dead = nd;
    ctx->perm = np;
    return FS2D_OK;
}

int particlesAliveCount(Ctx *ctx, int64_t *out)
{
    FS2D_TRY(fetchKilled(ctx));
    *out = ctx->count - ctx->deadCount - ctx->slab.ghostCount;  // slab mode: ghost copies belong to a neighbour
    return FS2D_OK;
}

int particlesMaxVelocity(Ctx *ctx, float *out)
{
    unsigned int *bits = reinterpret_cast<unsigned int *>(ctx->d_fscratch);
    const unsigned int init = 0x00800000u;  // FLT_MIN
    FS2D_CUDA(cudaMemcpyAsync(bits, &init, sizeof(init), cudaMemcpyHostToDevice, ctx->stream));
    if (ctx->count > 0)
    {
        const int blocks = std::min(gridFor(ctx->count), ctx->smCount * 16);
        maxVelocityKernel<<<blocks, NT, 0, ctx->stream>>>(ctx->pb[ctx->cur].vel, ctx->dead, ctx->count, ownedRows(ctx), bits);
        ctx->launches++;
    }
    unsigned int r = 0;
    FS2D_CUDA(fs2dCopyToHost(ctx, &r, bits, sizeof(r)));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->slab.enabled && ctx->slab.world > 1)
    {
        // the CFL step is global: max over ranks (non-negative floats order as their bit patterns)
        long long v[4] = {static_cast<long long>(r), 0, 0, 0}, all[4 * FS2D_MAX_RANKS];
        FS2D_TRY(slabAllGather(ctx, v, all));
        for (int k = 0; k < ctx->slab.world; k++) r = std::max(r, static_cast<unsigned int>(all[4 * k]));
    }
    float sq;
    memcpy(&sq, &r, sizeof(sq));
    *out = sqrtf(sq);
    return FS2D_OK;
}

int particlesAdvect(Ctx *ctx)
{
    particleStreamPositionsChanged(ctx);
    KernelGroupTimer kgt(ctx, FS2D_KGROUP_ADVECT);
    if (ctx->pstream.posPending && ctx->count > 0)
    {
        // streamed upload (one handle): the position section arrives in chunks and the advection follows it chunk by chunk
        ParticleBuffers &b = ctx->pb[ctx->cur];
        int64_t lo = 0, hi = 0;
        for (int c = 0; particleStreamNextPosChunk(ctx, c, &lo, &hi); c++)
        {
            if (hi <= lo) continue;
            advectKernel<<<gridFor(hi - lo), NT, 0, ctx->stream>>>(b.pos + lo, ctx->dead + lo, b.mis + lo, hi - lo,
                                                                 makeVelocityView(ctx->U, ctx->V, ctx->I, ctx->J), solidSdfView(ctx),
                                                                 ctx->material, ctx->I, ctx->J, ctx->stepDt, ownedRows(ctx),
                                                                 reinterpret_cast<unsigned long long *>(ctx->d_counter));
            ctx->launches++;
        }
        FS2D_TRY(particleStreamSettlePos(ctx));  // the whole section has arrived: cell keys
        ctx->killedDirty = true;
        ctx->sorted = false;
        FS2D_CUDA(cudaGetLastError());
        return FS2D_OK;
    }
    FS2D_TRY(particleStreamSettlePos(ctx));
    if (ctx->count == 0) return FS2D_OK;
    advectKernel<<<gridFor(ctx->count), NT, 0, ctx->stream>>>(ctx->pb[ctx->cur].pos, ctx->dead, ctx->pb[ctx->cur].mis, ctx->count,
                                                             makeVelocityView(ctx->U, ctx->V, ctx->I, ctx->J),
                                                             solidSdfView(ctx), ctx->material, ctx->I, ctx->J, ctx->stepDt, ownedRows(ctx),
                                                             reinterpret_cast<unsigned long long *>(ctx->d_counter));
    ctx->launches++;
    ctx->killedDirty = true;
    ctx->sorted = false;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

// Local counting sort of everything in the particle arrays (in slab mode: owned particles and ghosts). Cells
// outside the rows [rows.lo, rows.hi) hold no particle, so the histogram / scan / order passes only cover those.
int particlesSort(Ctx *ctx)
{
    // streamed upload: the sort itself needs positions only; property columns still in flight are gathered later
    // (particleStreamSettleAll), from the input buffer and the permutation this sort leaves behind
    if (ctx->pstream.propsGatherPending)
        FS2D_TRY(particleStreamSettleAll(ctx));
    else
        FS2D_TRY(particleStreamSettlePos(ctx));
    particleStreamPositionsChanged(ctx);
    const bool propsLater = ctx->pstream.propsPending;
    KernelGroupTimer kgt(ctx, FS2D_KGROUP_SORT);
    const bool slab = ctx->slab.enabled && ctx->slab.world > 1;
    const SlabRows rows = slab ? slabExt(ctx, ctx->slab.ghost) : SlabRows{0, ctx->I};
    const int64_t cLo = static_cast<int64_t>(rows.lo) * ctx->J, cHi = static_cast<int64_t>(rows.hi) * ctx->J;
    const int64_t cells = cHi - cLo;
    cudaStream_t st = ctx->stream;
    FS2D_CUDA(cudaMemsetAsync(ctx->cellCursor + cLo, 0, sizeof(int32_t) * (cells + 1), st));
    ParticleBuffers &in = ctx->pb[ctx->cur];
    ParticleBuffers &out = ctx->pb[ctx->cur ^ 1];
    if (ctx->count > 0)
    {
        histogramKernel<<<gridFor(ctx->count), NT, 0, st>>>(in.pos, ctx->dead, ctx->count, ctx->I, ctx->J, static_cast<uint32_t>(cLo),
                                                            static_cast<uint32_t>(cHi), in.key, ctx->cellCursor);
        ctx->launches++;
    }
    exclusiveScan(ctx, ctx->cellCursor + cLo, ctx->cellStart + cLo, cells + 1);
    int32_t alive = 0, ownedRange[2] = {0, 0};
    FS2D_CUDA(fs2dCopyToHost(ctx, &alive, ctx->cellStart + cHi, sizeof(alive)));
    if (slab)
    {
        FS2D_CUDA(fs2dCopyToHost(ctx, &ownedRange[0], ctx->cellStart + static_cast<int64_t>(ctx->slab.rowBegin) * ctx->J, sizeof(int32_t)));
        FS2D_CUDA(fs2dCopyToHost(ctx, &ownedRange[1], ctx->cellStart + static_cast<int64_t>(ctx->slab.rowEnd) * ctx->J, sizeof(int32_t)));
    }
    FS2D_CUDA(cudaMemsetAsync(ctx->cellCursor + cLo, 0, sizeof(int32_t) * (cells + 1), st));
    if (ctx->count > 0)
    {
        scatterKernel<<<gridFor(ctx->count), NT, 0, st>>>(in.key, ctx->count, ctx->cellStart, ctx->cellCursor, ctx->perm);
        cellOrderKernel<<<gridFor(cells), NT, 0, st>>>(in.pos, ctx->cellStart, cLo, cHi, ctx->perm);
        ctx->launches += 2;
    }
    FS2D_CUDA(cudaStreamSynchronize(st));
    if (alive > 0)
    {
        gatherKernel<<<gridFor(alive), NT, 0, st>>>(ctx->perm, alive, in.pos, in.vel, in.props, in.capacity, in.key, in.mis, out.pos,
                                                    out.vel, out.props, out.capacity, out.key, out.mis, propsLater ? 0 : ctx->p.num_properties,
                                                    ctx->dead);
        ctx->launches++;
    }
    if (propsLater)
    {
        ctx->pstream.propsGatherPending = true;
        ctx->pstream.propsFrom = ctx->cur;
        ctx->pstream.gatherCount = alive;
    }
    FS2D_CUDA(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), st));
    ctx->cur ^= 1;
    ctx->count = alive;
    ctx->deadCount = 0;
    ctx->killedDirty = false;
    ctx->sorted = true;
    if (slab)
    {
        ctx->slab.ownedBegin = ownedRange[0];
        ctx->slab.ownedEnd = ownedRange[1];
        ctx->slab.ghostCount = alive - (ownedRange[1] - ownedRange[0]);
    }
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

// pruneParticles + rebinParticles; in slab mode preceded by the exchange of migrants and ghosts with the row
// neighbours (collective: every rank calls it at the same points of the substep).
int particlesRebin(Ctx *ctx)
{
    FS2D_TRY(slabExchangeParticles(ctx));
    return particlesSort(ctx);
}

// FlipFireSolver::combustionUpdate (flipfiresolver.cpp:35-106), run after particleUpdate (:149-153). Above the ignition
// temperature a particle (PARTICLE mode) or a cell (GRID / HYBRID mode) burns min(dt * burnRate, fuel) and turns it into
// soot and heat. particleCombustionUpdateThread ignores the range it is handed and walks every bin (:58-86), so the
// reference burns every particle once per ThreadPool range, i.e. T times per substep for a pool of T threads;
// `repeats` reproduces that (convergence_threads = T; 1 when the knob is 0).
__global__ void __launch_bounds__(NT) particleCombustionKernel(float *__restrict__ temperature, float *__restrict__ concentration,
                                                               float *__restrict__ fuel, int64_t count, float ignition, float burn,
                                                               float smokeProportion, float heatProportion, int repeats)
{
    const int64_t p = blockIdx.x * static_cast<int64_t>(NT) + threadIdx.x;
    if (p >= count) return;
    float t = temperature[p], c = concentration[p], f = fuel[p];
    for (int k = 0; k < repeats; k++)
    {
        if (t > ignition && f > 0.f)
        {
            const float burnt = fminf(burn, f);
            f = __fsub_rn(f, burnt);
            c = __fadd_rn(c, __fmul_rn(smokeProportion, burnt));
            t = __fadd_rn(t, __fmul_rn(heatProportion, burnt));
        }
    }
    temperature[p] = t;
    concentration[p] = c;
    fuel[p] = f;
}

__global__ void __launch_bounds__(NT) gridCombustionKernel(float *__restrict__ temperature, float *__restrict__ concentration,
                                                           float *__restrict__ fuel, float *__restrict__ testGrid, int64_t n,
                                                           float ignition, float burn, float smokeProportion, float heatProportion)
{
    const int64_t k = blockIdx.x * static_cast<int64_t>(NT) + threadIdx.x;
    if (k >= n) return;
    const float t = temperature[k], f = fuel[k];
    if (t > ignition && f > 0.f)
    {
        const float burnt = fminf(burn, f);
        const float left = __fsub_rn(f, burnt);
        fuel[k] = left;
        testGrid[k] = left;
        concentration[k] = __fadd_rn(concentration[k], __fmul_rn(smokeProportion, burnt));
        temperature[k] = __fadd_rn(t, __fmul_rn(heatProportion, burnt));
    }
}

static int combustionUpdate(Ctx *ctx)
{
    const float burn = ctx->stepDt * ctx->p.burn_rate;  // m_stepDt * m_burnRate (float)
    if (ctx->p.parameter_handling == FS2D_PARAMS_PARTICLE)
    {
        if (ctx->count == 0) return FS2D_OK;
        if (ctx->p.temperature_property < 0 || ctx->p.concentration_property < 0 || ctx->p.fuel_property < 0)
        {
            ctx->lastError = "fire: temperature / concentration / fuel property columns are not set";
            return FS2D_ERR_STATE;
        }
        ParticleBuffers &b = ctx->pb[ctx->cur];
        particleCombustionKernel<<<gridFor(ctx->count), NT, 0, ctx->stream>>>(
            b.props + ctx->p.temperature_property * b.capacity, b.props + ctx->p.concentration_property * b.capacity,
            b.props + ctx->p.fuel_property * b.capacity, ctx->count, ctx->p.ignition_temperature, burn, ctx->p.smoke_proportion,
            ctx->p.heat_proportion, std::max(1, ctx->p.convergence_threads));
    }
    else
    {
        gridCombustionKernel<<<gridFor(ctx->N), NT, 0, ctx->stream>>>(ctx->temperature, ctx->concentration, ctx->fuel, ctx->testGrid, ctx->N,
                                                                      ctx->p.ignition_temperature, burn, ctx->p.smoke_proportion,
                                                                      ctx->p.heat_proportion);
    }
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int particlesUpdate(Ctx *ctx)
{
    FS2D_TRY(particleStreamSettleAll(ctx));
    ctx->pstream.earlyVelCount = -1;
    if (ctx->p.sim_type == FS2D_SIM_SMOKE || ctx->p.sim_type == FS2D_SIM_FIRE) ctx->pstream.earlyProps = false;  // decay / combustion rewrite the columns
    KernelGroupTimer kgt(ctx, FS2D_KGROUP_G2P);
    if (ctx->count == 0) return ctx->p.sim_type == FS2D_SIM_FIRE ? combustionUpdate(ctx) : FS2D_OK;
    const bool smoke = ctx->p.sim_type == FS2D_SIM_SMOKE || ctx->p.sim_type == FS2D_SIM_FIRE;
    ParticleBuffers &b = ctx->pb[ctx->cur];
    float *t = nullptr, *c = nullptr;
    float tf = 1.f, cf = 1.f;
    if (smoke)
    {
        if (ctx->p.temperature_property >= 0) t = b.props + ctx->p.temperature_property * b.capacity;
        if (ctx->p.concentration_property >= 0) c = b.props + ctx->p.concentration_property * b.capacity;
        // std::exp(-rate * dt) evaluated in float on the host, like the reference does per particle
        tf = std::exp(-ctx->p.temperature_decay * ctx->stepDt);
        cf = std::exp(-ctx->p.concentration_decay * ctx->stepDt);
    }
    particleUpdateKernel<<<gridFor(ctx->count), NT, 0, ctx->stream>>>(
        b.pos, b.vel, ctx->dead, ctx->count, makeVelocityView(ctx->U, ctx->V, ctx->I, ctx->J),
        makeVelocityView(ctx->savedU, ctx->savedV, ctx->I, ctx->J), ctx->p.pic_ratio, t, c, ctx->p.ambient_temperature, tf, cf,
        ownedRows(ctx));
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    if (ctx->p.sim_type == FS2D_SIM_FIRE) FS2D_TRY(combustionUpdate(ctx));
    return FS2D_OK;
}

int particlesAdjustByDensity(Ctx *ctx)
{
    FS2D_TRY(particleStreamSettlePos(ctx));
    particleStreamPositionsChanged(ctx);
    if (ctx->count == 0) return FS2D_OK;
    // (dt*dt) in float, denominator in double, result narrowed to float (flipsolver2d.cpp:263)
    const float scale = static_cast<float>(static_cast<double>(ctx->stepDt * ctx->stepDt) /
                                           (ctx->p.fluid_density * ctx->p.dx * ctx->p.dx));
    densityAdjustKernel<<<gridFor(ctx->count), NT, 0, ctx->stream>>>(ctx->pb[ctx->cur].pos, ctx->dead, ctx->pb[ctx->cur].mis, ctx->count, ctx->x,
                                                                    ctx->material, ctx->I, ctx->J, scale, ownedRows(ctx));
    ctx->launches++;
    ctx->sorted = false;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int particlesCount(Ctx *ctx)
{
    FS2D_TRY(particleStreamSettlePos(ctx));
    if (!ctx->sorted) FS2D_TRY(particlesSort(ctx));
    const SlabRows own = slabOwn(ctx);
    const int64_t cLo = static_cast<int64_t>(own.lo) * ctx->J, cHi = static_cast<int64_t>(own.hi) * ctx->J;
    countCapKernel<<<gridFor(cHi - cLo), NT, 0, ctx->stream>>>(ctx->cellStart, ctx->pb[ctx->cur].mis, cLo, cHi,
                                                           2 * ctx->p.particles_per_cell, ctx->counts, ctx->dead,
                                                           reinterpret_cast<unsigned long long *>(ctx->d_counter));
    ctx->launches++;
    ctx->killedDirty = true;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int particlesPruneNarrowBand(Ctx *ctx)
{
    FS2D_TRY(particleStreamSettlePos(ctx));
    if (ctx->count == 0) return FS2D_OK;
    pruneBandKernel<<<gridFor(ctx->count), NT, 0, ctx->stream>>>(ctx->pb[ctx->cur].pos, ctx->dead, ctx->count, fluidSdfView(ctx),
                                                                ctx->material, ctx->counts, ctx->I, ctx->J,
                                                                ctx->p.particles_per_cell, -3.f, -1.f, ownedRows(ctx),
                                                                reinterpret_cast<unsigned long long *>(ctx->d_counter));
    ctx->launches++;
    ctx->killedDirty = true;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int particlesReseedPlan(Ctx *ctx, int64_t *candidates)
{
    const int nb = ctx->p.sim_type == FS2D_SIM_NBFLIP ? 1 : 0;
    if (nb)
    {
        // NBFlip gives a new particle the viscosity GRID value at its position (nbflipsolver.cpp:183-190): a sample one row
        // beyond the slab for particles in its first / last row. P2G and the combine pass rewrote the owned rows since the
        // halo was last refreshed.
        const void *arr[1] = {ctx->viscosity};
        const size_t rb[1] = {sizeof(float) * ctx->J};
        FS2D_TRY(slabExchangeFields(ctx, arr, rb, 1));
    }
    // slab mode: every rank plans its own rows; the host mirror strings the ranks' draws together in rank order,
    // which IS the reference's row-major cell order
    const SlabRows own = slabOwn(ctx);
    const int64_t cLo = static_cast<int64_t>(own.lo) * ctx->J, cHi = static_cast<int64_t>(own.hi) * ctx->J;
    reseedPlanKernel<<<gridFor(cHi - cLo + 1), NT, 0, ctx->stream>>>(ctx->material, ctx->counts, ctx->fluidSdf, cLo, cHi,
                                                                    ctx->p.particles_per_cell, nb, -3.f, -1.f, ctx->reseedOffset);
    ctx->launches++;
    exclusiveScan(ctx, ctx->reseedOffset + cLo, ctx->reseedOffset + cLo, cHi - cLo + 1);
    int32_t total = 0;
    FS2D_CUDA(fs2dCopyToHost(ctx, &total, ctx->reseedOffset + cHi, sizeof(total)));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->reseedCandidates = total;
    *candidates = total;
    return FS2D_OK;
}

int particlesReseedApply(Ctx *ctx, int64_t candidates, const float *hostUniform)
{
    if (candidates != ctx->reseedCandidates)
    {
        ctx->lastError = "fs2d_reseed_apply: candidate count does not match the last fs2d_reseed_plan";
        return FS2D_ERR_STATE;
    }
    if (candidates == 0) return FS2D_OK;
    if (!hostUniform) return FS2D_ERR_ARG;
    FS2D_TRY(particleStreamSettleAll(ctx));
    if (ctx->numSources == 0 && ctx->p.sim_type != FS2D_SIM_NBFLIP)
    {
        ctx->lastError = "fs2d_reseed_apply: SOURCE cells exist but no source table was set";
        return FS2D_ERR_STATE;
    }
    FS2D_TRY(particlesReserve(ctx, ctx->count + candidates));
    if (2 * candidates > ctx->reseedUniformCapacity)  // persistent: cudaFree synchronises the device, keep it off the step path
    {
        if (ctx->reseedUniform) cudaFree(ctx->reseedUniform);
        ctx->reseedUniform = nullptr;
        ctx->reseedUniformCapacity = std::max<int64_t>(4 * candidates, 1 << 16);
        FS2D_CUDA(cudaMalloc(reinterpret_cast<void **>(&ctx->reseedUniform), sizeof(float) * ctx->reseedUniformCapacity));
    }
    float *du = ctx->reseedUniform;
    FS2D_CUDA(cudaMemcpyAsync(du, hostUniform, sizeof(float) * 2 * candidates, cudaMemcpyHostToDevice, ctx->stream));
    ParticleBuffers &b = ctx->pb[ctx->cur];
    ReseedArgs a;
    a.I = ctx->I;
    a.J = ctx->J;
    a.numProps = ctx->p.num_properties;
    a.simType = ctx->p.sim_type;
    a.viscosityProp = ctx->p.viscosity_property;
    a.temperatureProp = ctx->p.temperature_property;
    a.concentrationProp = ctx->p.concentration_property;
    a.fuelProp = ctx->p.fuel_property;
    a.testProp = ctx->p.test_property;
    a.narrowBand = -3.f;
    a.resamplingBand = -1.f;
    a.numSources = ctx->numSources;
    const bool smoke = ctx->p.sim_type == FS2D_SIM_SMOKE || ctx->p.sim_type == FS2D_SIM_FIRE;
    GridView none = makeView(nullptr, 1, 1, 0.f, 0.f);
    const SlabRows own = slabOwn(ctx);
    const int64_t cLo = static_cast<int64_t>(own.lo) * ctx->J, cHi = static_cast<int64_t>(own.hi) * ctx->J;
    reseedApplyKernel<<<gridFor(cHi - cLo), NT, 0, ctx->stream>>>(
        ctx->reseedOffset, cHi, du, ctx->material, ctx->emitterId, ctx->sources,
        makeVelocityView(ctx->U, ctx->V, ctx->I, ctx->J), fluidSdfView(ctx), viscosityView(ctx),
        smoke ? temperatureView(ctx) : none, smoke ? concentrationView(ctx) : none,
        ctx->p.sim_type == FS2D_SIM_FIRE ? fuelView(ctx) : none, a, b.pos, b.vel, b.props, b.capacity, ctx->count, ctx->dead, b.mis,
        b.key, cLo, reinterpret_cast<unsigned long long *>(ctx->d_counter));
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->count += candidates;
    ctx->killedDirty = true;
    ctx->sorted = false;
    ctx->reseedCandidates = 0;
    return FS2D_OK;
}

namespace
{
__global__ void __launch_bounds__(NT) setStorageBinsKernel(const float2 *__restrict__ pos, const int32_t *__restrict__ bins,
                                                           long long count, int binsJ, uint8_t *__restrict__ mis)
{
    const long long p = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (p >= count) return;
    const int2 pb = positionBin(pos[p]);
    const int sb = bins[p];
    mis[p] = static_cast<uint8_t>(storageCode(sb / binsJ - pb.x, sb % binsJ - pb.y));
}

__global__ void __launch_bounds__(NT) getStorageBinsKernel(const float2 *__restrict__ pos, const uint8_t *__restrict__ mis,
                                                           long long count, int binsJ, int32_t *__restrict__ bins)
{
    const long long p = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (p >= count) return;
    const int2 pb = positionBin(pos[p]);
    const unsigned int m = mis[p];
    bins[p] = m == FS2D_MIS_LOST ? -1 : (pb.x + rowOfCell(m, 5u) - 2) * binsJ + pb.y + static_cast<int>(m % 5u) - 2;
}
}  // namespace

// Storage bin (linear index in the ceil(I/3) x ceil(J/3) bin grid) of every particle, in device order.
int particlesSetStorageBins(Ctx *ctx, const int32_t *hostBins)
{
    FS2D_TRY(particleStreamSettleAll(ctx));  // ctx->perm is the scratch array here
    if (ctx->count == 0) return FS2D_OK;
    FS2D_CUDA(cudaMemcpyAsync(ctx->perm, hostBins, sizeof(int32_t) * ctx->count, cudaMemcpyHostToDevice, ctx->stream));
    setStorageBinsKernel<<<gridFor(ctx->count), NT, 0, ctx->stream>>>(ctx->pb[ctx->cur].pos, reinterpret_cast<const int32_t *>(ctx->perm),
                                                                     ctx->count, (ctx->J + 2) / 3, ctx->pb[ctx->cur].mis);
    ctx->launches++;
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    return FS2D_OK;
}

int particlesGetStorageBins(Ctx *ctx, int32_t *hostBins)
{
    FS2D_TRY(particleStreamSettleAll(ctx));
    if (ctx->count == 0) return FS2D_OK;
    getStorageBinsKernel<<<gridFor(ctx->count), NT, 0, ctx->stream>>>(ctx->pb[ctx->cur].pos, ctx->pb[ctx->cur].mis, ctx->count,
                                                                     (ctx->J + 2) / 3, reinterpret_cast<int32_t *>(ctx->perm));
    ctx->launches++;
    FS2D_CUDA(fs2dCopyToHost(ctx, hostBins, ctx->perm, sizeof(int32_t) * ctx->count));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    return FS2D_OK;
}
