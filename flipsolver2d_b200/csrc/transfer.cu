// Kernel group 2 (P2G) and the particle-reading part of group 3 (fluid SDF).
//
// The reference's P2G is a gather: every grid cell loops over all particles of the 3x3 block of
// 3x3-cell bins around it and sums in storage order (flipsolver2d.cpp:1329-1377). With the
// particles sorted by cell (particles.cu) the same gather needs only the cells inside the kernel
// support, and each row of those cells is ONE contiguous particle range
// [cellStart[r*J + j0], cellStart[r*J + j1 + 1]). A CTA owns a TILE_I x TILE_J tile of cells and
// first stages the particle records of the tile plus its halo into shared memory (one coalesced
// pass over the sorted arrays), then every thread gathers its cell from the staged records in a
// fixed order -- no atomics, so the result is deterministic and independent of launch geometry.
// When a tile's particles do not fit the staging buffer the threads read the same ranges
// straight from global memory (L1/L2 hits), same order, same result.
//
// Weights are evaluated with the reference's expressions and thresholds (mathfuncs.cpp:32-43,
// :100-116) using non-contracted float arithmetic; only the summation ORDER differs from the
// reference (its in-bin order is history dependent), so sums agree to float rounding, not bitwise.
#include <algorithm>
#include <cfloat>

#include "fs2d_device.cuh"
#include "fs2d_internal.h"

namespace
{
constexpr int TI = 8;          // tile rows (cells)
constexpr int TJ = 32;         // tile columns (cells)
constexpr int NT = TI * TJ;    // one thread per cell
constexpr int STAGE = 4352;    // staged particle records per CTA (pos + 2 payload floats + storage byte = 17 B each): 72 KB, 3 CTAs per SM

// PAY = payload floats per record: 2 (velocity), 1 (one property column), 0 (positions only: density). The shared
// memory request follows the payload -- 72 / 55 / 38 KB, i.e. 3 / 4 / 5 CTAs per SM: the gathers are bound by instruction
// issue at low occupancy (ncu: 25 % warps active with the former 96 KB for every kernel), not by staging traffic.
template <int PAY> struct TileStage
{
    float2 pos[STAGE];
    float pay[PAY > 0 ? STAGE * PAY : 1];
    unsigned char mis[STAGE];   // storage-bin code (FS2D_MIS_*)
    int rowBegin[TI + 4];       // global particle index where each staged row starts
    int rowOffset[TI + 5];      // offset of each staged row inside pos/pay
    __device__ __forceinline__ float2 payload(int k) const
    {
        if (PAY == 2) return make_float2(pay[2 * k], pay[2 * k + 1]);
        if (PAY == 1) return make_float2(pay[k], 0.f);
        return make_float2(0.f, 0.f);
    }
};
template <int PAY> constexpr size_t stageBytes() { return sizeof(TileStage<PAY>); }
template <int PAY> constexpr int stageCtasPerSm() { return PAY == 2 ? 3 : (PAY == 1 ? 4 : 5); }

// Stage rows [i0-haloLo, i0+TI-1+haloHi] x columns [j0-haloLo, j0+TJ-1+haloHi] of the sorted
// particle arrays. Returns (for the whole CTA) 1 = staged, 0 = the records do not fit (read them from global memory),
// 2 = there is no particle at all in the tile and its halo.
template <int PAY>
__device__ int stageTile(TileStage<PAY> &s, const int32_t *__restrict__ cellStart, const float2 *__restrict__ pos,
                          const float2 *__restrict__ pay2, const float *__restrict__ pay1, const uint8_t *__restrict__ mis, int I,
                          int J, int i0, int j0, int haloLo, int haloHi)
{
    const int rows = TI + haloLo + haloHi;
    __shared__ int total;
    // one thread per staged row fetches its particle range (the loads run in parallel: with one CTA per 8 x 32 cells
    // and 91 % of the tiles empty in a dam break, a serial walk over the row table dominated the whole kernel)
    if (threadIdx.x < rows)
    {
        const int r = threadIdx.x;
        const int ja = max(j0 - haloLo, 0), jb = min(j0 + TJ - 1 + haloHi, J - 1);
        const int gi = i0 - haloLo + r;
        int b = 0, e = 0;
        if (gi >= 0 && gi < I && ja <= jb)
        {
            b = cellStart[static_cast<long long>(gi) * J + ja];
            e = cellStart[static_cast<long long>(gi) * J + jb + 1];
        }
        s.rowBegin[r] = b;
        s.rowOffset[r] = e - b;
    }
    __syncthreads();
    if (threadIdx.x == 0)
    {
        int acc = 0;
        for (int r = 0; r < rows; r++)
        {
            const int c = s.rowOffset[r];
            s.rowOffset[r] = acc;
            acc += c;
        }
        s.rowOffset[rows] = acc;
        total = acc;
    }
    __syncthreads();
    if (total == 0) return 2;  // nothing within reach of this tile: the callers write their "no particle" values
    if (total > STAGE) return 0;
    for (int r = 0; r < rows; r++)
    {
        const int n = s.rowOffset[r + 1] - s.rowOffset[r];
        const int gb = s.rowBegin[r], so = s.rowOffset[r];
        for (int k = threadIdx.x; k < n; k += NT)
        {
            s.pos[so + k] = pos[gb + k];
            s.mis[so + k] = mis[gb + k];
            if (PAY == 2)
            {
                const float2 v = pay2[gb + k];
                s.pay[2 * (so + k)] = v.x;
                s.pay[2 * (so + k) + 1] = v.y;
            }
            else if (PAY == 1)
                s.pay[so + k] = pay1 ? pay1[gb + k] : 0.f;
        }
    }
    __syncthreads();
    return 1;
}

// Iterate the particles of cells [ja..jb] of row gi, staged or global. F(pos, payload).
template <int PAY, class F>
__device__ __forceinline__ void forRowRange(const TileStage<PAY> &s, bool staged, const int32_t *__restrict__ cellStart,
                                            const float2 *__restrict__ pos, const float2 *__restrict__ pay2,
                                            const float *__restrict__ pay1, const uint8_t *__restrict__ mis, int J, int gi,
                                            int ja, int jb, int stagedRow, int ci, int cj, F f)
{
    const int b = cellStart[static_cast<long long>(gi) * J + ja];
    const int e = cellStart[static_cast<long long>(gi) * J + jb + 1];
    // A particle filed in the bin of its position is always inside the 3x3 bins of a cell whose kernel
    // support reaches it; only re-filed-late particles (FS2D_MIS_*) need the explicit bin test.
    if (staged)
    {
        const int shift = s.rowOffset[stagedRow] - s.rowBegin[stagedRow];
        for (int k = b; k < e; k++)
        {
            const unsigned int m = s.mis[k + shift];
            const float2 p = s.pos[k + shift];
            if (m != FS2D_MIS_HOME && !storageVisible(m, p, ci, cj)) continue;
            f(p, s.payload(k + shift));
        }
    }
    else
    {
        for (int k = b; k < e; k++)
        {
            const unsigned int m = mis[k];
            const float2 p = pos[k];
            if (m != FS2D_MIS_HOME && !storageVisible(m, p, ci, cj)) continue;
            f(p, PAY == 2 ? pay2[k] : make_float2((PAY == 1 && pay1) ? pay1[k] : 0.f, 0.f));
        }
    }
}

// particleVelocityToGridThread (flipsolver2d.cpp:1329-1377). Cell (i,j) receives particles with
// weightU > 1e-9 && weightV > 1e-9, i.e. floor(pos) within one cell of (i,j).
__global__ void __launch_bounds__(NT) p2gVelocityKernel(const int32_t *__restrict__ cellStart, const float2 *__restrict__ pos,
                                                        const float2 *__restrict__ vel, const uint8_t *__restrict__ mis, int I,
                                                        int J, int tilesJ, const int *__restrict__ tileList,
                                                        const int *__restrict__ tileCount,
                                                        float *__restrict__ U, float *__restrict__ V,
                                                        uint8_t *__restrict__ uValid, uint8_t *__restrict__ vValid)
{
    extern __shared__ __align__(16) unsigned char stageRaw[];
    TileStage<2> &s = *reinterpret_cast<TileStage<2> *>(stageRaw);
    const int count = *tileCount;
    for (int t = blockIdx.x; t < count; t += gridDim.x)  // persistent CTAs walk the tiles that have particles in reach
    {
        const int tile = tileList[t];
        const int ti = tile / tilesJ, tj = tile - ti * tilesJ;
        const int i0 = ti * TI, j0 = tj * TJ;
        const int stageMode = stageTile<2>(s, cellStart, pos, vel, nullptr, mis, I, J, i0, j0, 1, 1);
        const bool staged = stageMode != 0, empty = stageMode == 2;
        const int li = threadIdx.x / TJ, lj = threadIdx.x - li * TJ;
        const int i = i0 + li, j = j0 + lj;
        if (i < I && j < J)
        {
            const float ci = static_cast<float>(i), cj = static_cast<float>(j);
            const float cjh = faddr(cj, 0.5f), cih = faddr(ci, 0.5f);
            float uW = 1e-10f, vW = 1e-10f, uAcc = 0.f, vAcc = 0.f;
            bool any = false;
            const int ja = max(j - 1, 0), jb = min(j + 1, J - 1);
            for (int gi = max(i - 1, 0); !empty && gi <= min(i + 1, I - 1); gi++)
            {
                forRowRange<2>(s, staged, cellStart, pos, vel, nullptr, mis, J, gi, ja, jb, gi - (i0 - 1), i, j,
                                  [&](float2 p, float2 v)
                                  {
                                      const float wU = quadraticBSpline(fsubr(p.x, ci), fsubr(p.y, cjh));
                                      const float wV = quadraticBSpline(fsubr(p.x, cih), fsubr(p.y, cj));
                                      if (wU > 1e-9f && wV > 1e-9f)
                                      {
                                          uW = faddr(uW, wU);
                                          uAcc = faddr(uAcc, fmulr(wU, v.x));
                                          vW = faddr(vW, wV);
                                          vAcc = faddr(vAcc, fmulr(wV, v.y));
                                          any = true;
                                      }
                                  });
            }
            U[static_cast<long long>(i) * J + j] = __fdiv_rn(uAcc, uW);
            uValid[static_cast<long long>(i) * J + j] = any ? 1 : 0;
            V[static_cast<long long>(i) * (J + 1) + j] = __fdiv_rn(vAcc, vW);
            vValid[static_cast<long long>(i) * (J + 1) + j] = any ? 1 : 0;
        }
        __syncthreads();  // the stage is rewritten for the next tile
    }
}

// centeredParamsToGridThread. MODE 0: water (flipsolver2d.cpp:1394-1431: threshold w > 1e-9, value
// = sum/(1e-10 + sum w) always written). MODE 1: nbflip/smoke/fire (nbflipsolver.cpp:463-504,
// flipsmokesolver.cpp:509-558: threshold |w| > 1e-6, divided only when known, grids pre-zeroed).
// The support of B(px - i) spans cells i-2 .. i+1.
template <int MODE>
__global__ void __launch_bounds__(NT) p2gCenteredKernel(const int32_t *__restrict__ cellStart, const float2 *__restrict__ pos,
                                                        const float *__restrict__ prop, const uint8_t *__restrict__ mis, int I,
                                                        int J, int tilesJ, const int *__restrict__ tileList,
                                                        const int *__restrict__ tileCount,
                                                        float *__restrict__ out, uint8_t *__restrict__ known)
{
    extern __shared__ __align__(16) unsigned char stageRaw[];
    TileStage<1> &s = *reinterpret_cast<TileStage<1> *>(stageRaw);
    const int count = *tileCount;
    for (int t = blockIdx.x; t < count; t += gridDim.x)
    {
        const int tile = tileList[t];
        const int ti = tile / tilesJ, tj = tile - ti * tilesJ;
        const int i0 = ti * TI, j0 = tj * TJ;
        const int stageMode = stageTile<1>(s, cellStart, pos, nullptr, prop, mis, I, J, i0, j0, 2, 1);
        const bool staged = stageMode != 0, empty = stageMode == 2;
        const int li = threadIdx.x / TJ, lj = threadIdx.x - li * TJ;
        const int i = i0 + li, j = j0 + lj;
        if (i < I && j < J)
        {
            const float ci = static_cast<float>(i), cj = static_cast<float>(j);
            float wSum = 1e-10f, acc = 0.f;
            bool any = false;
            const int ja = max(j - 2, 0), jb = min(j + 1, J - 1);
            for (int gi = max(i - 2, 0); !empty && gi <= min(i + 1, I - 1); gi++)
            {
                forRowRange<1>(s, staged, cellStart, pos, nullptr, prop, mis, J, gi, ja, jb, gi - (i0 - 2), i, j,
                                   [&](float2 p, float2 v)
                                   {
                                       const float w = quadraticBSpline(fsubr(p.x, ci), fsubr(p.y, cj));
                                       const bool take = MODE == 0 ? (w > 1e-9f) : (fabsf(w) > 1e-6f);
                                       if (take)
                                       {
                                           wSum = faddr(wSum, w);
                                           acc = faddr(acc, fmulr(w, v.x));
                                           any = true;
                                       }
                                   });
            }
            const long long n = static_cast<long long>(i) * J + j;
            if (MODE == 0)
                out[n] = __fdiv_rn(acc, wSum);
            else
                out[n] = any ? __fdiv_rn(acc, wSum) : 0.f;
            if (known) known[n] = any ? 1 : 0;
        }
        __syncthreads();
    }
}

// The tiles a P2G kernel has to visit: those with at least one particle inside the tile grown by the widest kernel
// support (2 cells up/left, 1 down/right). Everything else keeps the cleared value -- "no particle in reach" is 0 for
// the velocity samples, their validity flags and the centred parameters alike (uAcc / 1e-10 = 0, known = false). One
// warp per tile; the list order is arbitrary (tiles are independent and each cell's sum has a fixed order).
__global__ void __launch_bounds__(256) p2gTileListKernel(const int32_t *__restrict__ cellStart, int I, int J, int tilesJ, int tileBase,
                                                         int tiles, int *__restrict__ list, int *__restrict__ count)
{
    const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (w >= tiles) return;
    const int tile = tileBase + w;
    const int ti = tile / tilesJ, tj = tile - ti * tilesJ;
    const int i0 = ti * TI, j0 = tj * TJ;
    const int ja = max(j0 - 2, 0), jb = min(j0 + TJ, J - 1);
    int n = 0;
    const int gi = i0 - 2 + lane;
    if (lane < TI + 3 && gi >= 0 && gi < I && ja <= jb)
        n = cellStart[static_cast<long long>(gi) * J + jb + 1] - cellStart[static_cast<long long>(gi) * J + ja];
    n = __reduce_add_sync(0xffffffffu, n);
    if (lane == 0 && n > 0) list[atomicAdd(count, 1)] = tile;
}

// updateDensityGridThread (flipsolver2d.cpp:201-249)
// The value a cell without any particle in reach gets: 0 / cellVolume, clamped up to the rest density when the cell is
// FLUID and touches an EMPTY cell (flipsolver2d.cpp:236-247).
__device__ __forceinline__ float densityClamp(const int8_t *__restrict__ mat, int I, int J, int i, int j, float d, float restDensity)
{
    if (matFluid(mat[static_cast<long long>(i) * J + j]) &&
        (matEmpty(matAt(mat, I, J, i + 1, j)) || matEmpty(matAt(mat, I, J, i - 1, j)) || matEmpty(matAt(mat, I, J, i, j + 1)) ||
         matEmpty(matAt(mat, I, J, i, j - 1))))
        d = fminf(fmaxf(d, restDensity), FLT_MAX);
    return d;
}

// all cells of [nBegin, nEnd): the "no particle in reach" value; the tiles with particles overwrite theirs afterwards
__global__ void __launch_bounds__(256) densityBackgroundKernel(const int8_t *__restrict__ mat, int I, int J, float cellVolume, float restDensity,
                                                               float *__restrict__ density, long long nBegin, long long nEnd)
{
    const long long n = nBegin + blockIdx.x * 256ll + threadIdx.x;
    if (n >= nEnd) return;
    const int i = rowOfCell(n, J), j = static_cast<int>(n - static_cast<long long>(i) * J);
    density[n] = densityClamp(mat, I, J, i, j, __fdiv_rn(0.f, cellVolume), restDensity);
}

__global__ void __launch_bounds__(NT) densityKernel(const int32_t *__restrict__ cellStart, const float2 *__restrict__ pos,
                                                    const uint8_t *__restrict__ mis, const int8_t *__restrict__ mat, int I, int J,
                                                    int tilesJ, const int *__restrict__ tileList, const int *__restrict__ tileCount,
                                                    float particleMass, float cellVolume, float restDensity, float *__restrict__ density)
{
    extern __shared__ __align__(16) unsigned char stageRaw[];
    TileStage<0> &s = *reinterpret_cast<TileStage<0> *>(stageRaw);
    const int count = *tileCount;
    for (int t = blockIdx.x; t < count; t += gridDim.x)
    {
        const int tile = tileList[t];
        const int ti = tile / tilesJ, tj = tile - ti * tilesJ;
        const int i0 = ti * TI, j0 = tj * TJ;
        const int stageMode = stageTile<0>(s, cellStart, pos, nullptr, nullptr, mis, I, J, i0, j0, 1, 1);
        const bool staged = stageMode != 0, empty = stageMode == 2;
        const int li = threadIdx.x / TJ, lj = threadIdx.x - li * TJ;
        const int i = i0 + li, j = j0 + lj;
        if (i < I && j < J)
        {
            const float ci = static_cast<float>(i), cj = static_cast<float>(j);
            float acc = 0.f;
            const int ja = max(j - 1, 0), jb = min(j + 1, J - 1);
            for (int gi = max(i - 1, 0); !empty && gi <= min(i + 1, I - 1); gi++)
            {
                forRowRange<0>(s, staged, cellStart, pos, nullptr, nullptr, mis, J, gi, ja, jb, gi - (i0 - 1), i, j,
                                   [&](float2 p, float2)
                                   {
                                       const float w = bilinearHat(fsubr(fsubr(p.x, ci), 0.5f), fsubr(fsubr(p.y, cj), 0.5f));
                                       if (fabsf(w) > 1e-6f) acc = faddr(acc, fmulr(w, particleMass));
                                   });
            }
            density[static_cast<long long>(i) * J + j] = densityClamp(mat, I, J, i, j, __fdiv_rn(acc, cellVolume), restDensity);
        }
        __syncthreads();
    }
}

// updateSdfThread (flipsolver2d.cpp:1276-1311): min squared distance to the particles FILED in the 3x3
// block of bins around the cell's bin, minus the particle radius. The minimum is order independent, so
// rows are visited outwards from the cell's own row and the walk stops once no remaining row can hold a
// closer particle; columns are clipped the same way -- the value is identical to the full scan. A
// particle can be filed one bin away from its position (FS2D_MIS_*), so positions are scanned over the
// 5x5 bins around the cell and every candidate is tested against the reference's 3x3 bin window.
__global__ void __launch_bounds__(256) sdfKernel(const int32_t *__restrict__ cellStart, const float2 *__restrict__ pos,
                                                 const uint8_t *__restrict__ mis, int I, int J, float radius,
                                                 float *__restrict__ sdf, long long nBegin, long long nEnd)
{
    const long long n = nBegin + blockIdx.x * 256ll + threadIdx.x;
    const bool valid = n < nEnd;
    const long long nc = valid ? n : nEnd - 1;
    const int i = rowOfCell(nc, J), j = static_cast<int>(nc - static_cast<long long>(i) * J);
    const int bi = i / 3, bj = j / 3;
    const int iLo = max(3 * (bi - 2), 0), iHi = min(3 * (bi + 2) + 2, I - 1);
    const int jLo = max(3 * (bj - 2), 0), jHi = min(3 * (bj + 2) + 2, J - 1);
    const float cx = faddr(static_cast<float>(i), 0.5f), cy = faddr(static_cast<float>(j), 0.5f);
    float best = FLT_MAX;
    {
        // Most of the grid is far from any particle (91 % of the cells in the dam break). The warp counts, with one
        // row per lane, the particles inside the union of its 32 search windows; when there is none every cell of the
        // warp gets the "no particle" value without walking its own window.
        const unsigned int all = 0xffffffffu;
        const int wiLo = __reduce_min_sync(all, iLo), wiHi = __reduce_max_sync(all, iHi);
        const int wjLo = __reduce_min_sync(all, jLo), wjHi = __reduce_max_sync(all, jHi);
        int cnt = 0;
        for (int gi = wiLo + static_cast<int>(threadIdx.x & 31u); gi <= wiHi; gi += 32)
            cnt += cellStart[static_cast<long long>(gi) * J + wjHi + 1] - cellStart[static_cast<long long>(gi) * J + wjLo];
        if (__reduce_add_sync(all, cnt) == 0)
        {
            if (valid) sdf[n] = fsubr(__fsqrt_rn(best), radius);
            return;
        }
    }
    if (!valid) return;
    const int reach = max(i - iLo, iHi - i);
    for (int d = 0; d <= reach; d++)
    {
        // every particle in a row (column) at cell distance d is at least d - 1/2 away from the centre
        const float lower = static_cast<float>(d) - 0.5f;
        if (d > 0 && lower * lower >= best) break;
        int ja = jLo, jb = jHi;
        if (best < 64.f)
        {
            const int w = static_cast<int>(sqrtf(best) + 0.5f) + 1;
            ja = max(jLo, j - w);
            jb = min(jHi, j + w);
        }
        for (int sgn = 0; sgn < (d == 0 ? 1 : 2); sgn++)
        {
            const int gi = sgn == 0 ? i - d : i + d;
            if (gi < iLo || gi > iHi) continue;
            const int b = cellStart[static_cast<long long>(gi) * J + ja];
            const int e = cellStart[static_cast<long long>(gi) * J + jb + 1];
            for (int k = b; k < e; k++)
            {
                const float2 p = pos[k];
                if (!storageVisible(mis[k], p, i, j)) continue;
                const float dx = fsubr(p.x, cx), dy = fsubr(p.y, cy);
                const float d2 = faddr(fmulr(dx, dx), fmulr(dy, dy));
                if (d2 < best) best = d2;
            }
        }
    }
    sdf[n] = fsubr(__fsqrt_rn(best), radius);
}

template <int PAY, class K> void allowStage(K kernel)
{
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(stageBytes<PAY>()));
}

// Tiles covering the rows `r` (slab boundaries are multiples of 16, hence of TI); *tileBase = first tile.
int tileCount(const Ctx *ctx, SlabRows r, int *tilesJ, int *tileBase)
{
    *tilesJ = divUp(ctx->J, TJ);
    const int t0 = r.lo / TI, t1 = divUp(r.hi, TI);
    *tileBase = t0 * *tilesJ;
    return (t1 - t0) * *tilesJ;
}
}  // namespace

int particlesSort(Ctx *ctx);

// Builds the list of tiles with particles in reach for the rows `r`; returns the persistent grid size to launch.
static int buildTileList(Ctx *ctx, SlabRows r, int *tilesJ, int ctasPerSm)
{
    int tileBase = 0;
    const int tiles = tileCount(ctx, r, tilesJ, &tileBase);
    if (!ctx->p2gTileList)
    {
        const int all = divUp(ctx->I, TI) * divUp(ctx->J, TJ);
        if (cudaMalloc(reinterpret_cast<void **>(&ctx->p2gTileList), sizeof(int) * (static_cast<size_t>(all) + 1)) != cudaSuccess) return -1;
    }
    int *count = ctx->p2gTileList, *list = ctx->p2gTileList + 1;
    cudaMemsetAsync(count, 0, sizeof(int), ctx->stream);
    p2gTileListKernel<<<divUp(tiles, 8), 256, 0, ctx->stream>>>(ctx->cellStart, ctx->I, ctx->J, *tilesJ, tileBase, tiles, list, count);
    ctx->launches++;
    (void)list;
    return std::max(1, std::min(tiles, ctasPerSm * ctx->smCount));
}

static int ensureSorted(Ctx *ctx)
{
    if (!ctx->sorted) return particlesSort(ctx);
    return particleStreamSettlePos(ctx);
}

int transferVelocity(Ctx *ctx)
{
    KernelGroupTimer kgt(ctx, FS2D_KGROUP_P2G);
    FS2D_TRY(ensureSorted(ctx));
    cudaStream_t st = ctx->stream;
    // fill(0)/fill(false) of all four arrays (flipsolver2d.cpp:1315-1318); row I of U and column J of V keep 0
    FS2D_CUDA(cudaMemsetAsync(ctx->U + ctx->N, 0, sizeof(float) * (ctx->NU - ctx->N), st));
    FS2D_CUDA(cudaMemsetAsync(ctx->uValid + ctx->N, 0, ctx->NU - ctx->N, st));
    const SlabRows own = slabOwn(ctx);
    const size_t vOff = static_cast<size_t>(own.lo) * (ctx->J + 1), vCnt = static_cast<size_t>(own.hi - own.lo) * (ctx->J + 1);
    FS2D_CUDA(cudaMemsetAsync(ctx->V + vOff, 0, sizeof(float) * vCnt, st));
    FS2D_CUDA(cudaMemsetAsync(ctx->vValid + vOff, 0, vCnt, st));
    // the owned rows of U and its flags: cells no particle reaches keep this 0 (the kernel only visits the other tiles)
    const size_t uOff = static_cast<size_t>(own.lo) * ctx->J, uCnt = static_cast<size_t>(own.hi - own.lo) * ctx->J;
    FS2D_CUDA(cudaMemsetAsync(ctx->U + uOff, 0, sizeof(float) * uCnt, st));
    FS2D_CUDA(cudaMemsetAsync(ctx->uValid + uOff, 0, uCnt, st));
    int tilesJ;
    const int grid = buildTileList(ctx, own, &tilesJ, stageCtasPerSm<2>());
    if (grid < 0) return FS2D_ERR_CUDA;
    ParticleBuffers &b = ctx->pb[ctx->cur];
    allowStage<2>(p2gVelocityKernel);
    p2gVelocityKernel<<<grid, NT, stageBytes<2>(), st>>>(ctx->cellStart, b.pos, b.vel, b.mis, ctx->I, ctx->J, tilesJ, ctx->p2gTileList + 1,
                                                     ctx->p2gTileList, ctx->U, ctx->V, ctx->uValid, ctx->vValid);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int transferCentered(Ctx *ctx)
{
    KernelGroupTimer kgt(ctx, FS2D_KGROUP_P2G);
    FS2D_TRY(ensureSorted(ctx));
    FS2D_TRY(particleStreamSettleAll(ctx));  // streamed upload: the property columns are needed from here on
    cudaStream_t st = ctx->stream;
    const SlabRows own = slabOwn(ctx);
    int tilesJ;
    const int grid = buildTileList(ctx, own, &tilesJ, stageCtasPerSm<1>());
    if (grid < 0) return FS2D_ERR_CUDA;
    const int *tl = ctx->p2gTileList + 1, *tc = ctx->p2gTileList;
    ParticleBuffers &b = ctx->pb[ctx->cur];
    // m_divergenceControl.fill(0.f) in all three variants (flipsolver2d.cpp:1382, nbflipsolver.cpp:450,
    // flipsmokesolver.cpp:60); slab mode clears the rows the right-hand side is evaluated on
    const SlabRows e1 = slabExt(ctx, 1);
    FS2D_CUDA(cudaMemsetAsync(ctx->divergenceControl + static_cast<size_t>(e1.lo) * ctx->J, 0,
                              sizeof(float) * static_cast<size_t>(e1.hi - e1.lo) * ctx->J, st));
    // cells no particle reaches: value 0, not known (the kernels only visit tiles with particles in reach)
    const size_t cOff = static_cast<size_t>(own.lo) * ctx->J, cCnt = static_cast<size_t>(own.hi - own.lo) * ctx->J;
    auto clear = [&](float *grid) { return cudaMemsetAsync(grid + cOff, 0, sizeof(float) * cCnt, st); };
    FS2D_CUDA(cudaMemsetAsync(ctx->knownCentered + cOff, 0, cCnt, st));
    allowStage<1>(p2gCenteredKernel<0>);
    allowStage<1>(p2gCenteredKernel<1>);
    auto column = [&](int prop) -> const float * { return prop >= 0 ? b.props + static_cast<int64_t>(prop) * b.capacity : nullptr; };
    switch (ctx->p.sim_type)
    {
    case FS2D_SIM_LIQUID:
        FS2D_CUDA(clear(ctx->viscosity));
        p2gCenteredKernel<0><<<grid, NT, stageBytes<1>(), st>>>(ctx->cellStart, b.pos, column(ctx->p.viscosity_property), b.mis, ctx->I, ctx->J,
                                                  tilesJ, tl, tc, ctx->viscosity, ctx->knownCentered);
        ctx->launches++;
        break;
    case FS2D_SIM_NBFLIP:
        FS2D_CUDA(clear(ctx->viscosity));
        p2gCenteredKernel<1><<<grid, NT, stageBytes<1>(), st>>>(ctx->cellStart, b.pos, column(ctx->p.viscosity_property), b.mis, ctx->I, ctx->J,
                                                  tilesJ, tl, tc, ctx->viscosity, ctx->knownCentered);
        ctx->launches++;
        break;
    default:  // smoke / fire: temperature and concentration (fire's fuel column has no P2G in the reference)
        FS2D_CUDA(clear(ctx->temperature));
        FS2D_CUDA(clear(ctx->concentration));
        p2gCenteredKernel<1><<<grid, NT, stageBytes<1>(), st>>>(ctx->cellStart, b.pos, column(ctx->p.temperature_property), b.mis, ctx->I, ctx->J,
                                                  tilesJ, tl, tc, ctx->temperature, ctx->knownCentered);
        p2gCenteredKernel<1><<<grid, NT, stageBytes<1>(), st>>>(ctx->cellStart, b.pos, column(ctx->p.concentration_property), b.mis, ctx->I, ctx->J,
                                                  tilesJ, tl, tc, ctx->concentration, nullptr);
        ctx->launches += 2;
        break;
    }
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int transferDensity(Ctx *ctx)
{
    KernelGroupTimer kgt(ctx, FS2D_KGROUP_DENSITY);
    FS2D_TRY(ensureSorted(ctx));
    // slab mode: one extra tile row each side, the density right-hand side needs one halo row (ghost particles
    // reach 12 rows, the tile halo needs 9)
    const SlabRows rows = slabExt(ctx, ctx->slab.enabled ? TI : 0);
    int tilesJ;
    const int grid = buildTileList(ctx, rows, &tilesJ, stageCtasPerSm<0>());
    if (grid < 0) return FS2D_ERR_CUDA;
    // float cellVolume = dx*dx*dx; float particleMass = (rho * cellVolume) / float(ppc) (flipsolver2d.cpp:203-204)
    const float cellVolume = static_cast<float>(ctx->p.dx * ctx->p.dx * ctx->p.dx);
    const float particleMass =
        static_cast<float>((ctx->p.fluid_density * cellVolume) / static_cast<float>(ctx->p.particles_per_cell));
    const int rLo = (rows.lo / TI) * TI, rHi = std::min(ctx->I, divUp(rows.hi, TI) * TI);
    const long long nBegin = static_cast<long long>(rLo) * ctx->J, nEnd = static_cast<long long>(rHi) * ctx->J;
    densityBackgroundKernel<<<divUp(nEnd - nBegin, 256), 256, 0, ctx->stream>>>(ctx->material, ctx->I, ctx->J, cellVolume,
                                                                               static_cast<float>(ctx->p.fluid_density), ctx->density, nBegin, nEnd);
    allowStage<0>(densityKernel);
    densityKernel<<<grid, NT, stageBytes<0>(), ctx->stream>>>(ctx->cellStart, ctx->pb[ctx->cur].pos, ctx->pb[ctx->cur].mis, ctx->material, ctx->I, ctx->J,
                                                         tilesJ, ctx->p2gTileList + 1, ctx->p2gTileList, particleMass, cellVolume,
                                                         static_cast<float>(ctx->p.fluid_density), ctx->density);
    ctx->launches++;
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int transferSdf(Ctx *ctx)
{
    KernelGroupTimer kgt(ctx, FS2D_KGROUP_SDF);
    FS2D_TRY(ensureSorted(ctx));
    ctx->sdfInsidePending = false;  // the whole level set is rewritten below
    const SlabRows own = slabOwn(ctx);
    const long long nBegin = static_cast<long long>(own.lo) * ctx->J, nEnd = static_cast<long long>(own.hi) * ctx->J;
    sdfKernel<<<divUp(nEnd - nBegin, 256), 256, 0, ctx->stream>>>(ctx->cellStart, ctx->pb[ctx->cur].pos, ctx->pb[ctx->cur].mis, ctx->I, ctx->J,
                                                                 ctx->p.particle_scale, ctx->fluidSdf, nBegin, nEnd);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}
