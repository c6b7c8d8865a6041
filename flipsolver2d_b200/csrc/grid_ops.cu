// Kernel group 3 (classification / extrapolation) and the grid-side pieces of group 4
// (matrix rows, right-hand sides, pressure application) and group 5 (semi-Lagrangian
// advection). All are one-thread-per-sample streaming kernels over the dense row-major
// grids; each sample is read/written once, neighbours come from L1/L2.
#include <algorithm>
#include <cooperative_groups.h>

#include <cstdlib>

#include "fs2d_internal.h"
#include "fs2d_device.cuh"

namespace cg = cooperative_groups;

namespace
{
constexpr int NT = 256;

// Linear cell range of a row range (slab mode restricts every sweep to the rows a rank needs).
struct CellRange
{
    long long begin, end;
    long long count() const { return end > begin ? end - begin : 0; }
};
inline CellRange cellRange(const Ctx *ctx, SlabRows r) { return {static_cast<long long>(r.lo) * ctx->J, static_cast<long long>(r.hi) * ctx->J}; }

// ------------------------------------------------------------------ matrix rows
// getPressureProjectionMatrix (flipsolver2d.cpp:797-887; smoke flipsmokesolver.cpp:354-444)
// and getIPPCoefficients (flipsolver2d.cpp:889-945) as per-cell bit fields.
__global__ void __launch_bounds__(NT) buildMatrixKernel(const int8_t *__restrict__ mat, int I, int J, int smokeRows,
                                                        uint8_t *__restrict__ rowInfo, uint16_t *__restrict__ preInfo,
                                                        long long nBegin, long long nEnd)
{
    const long long N = static_cast<long long>(I) * J;
    const long long n = nBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n >= nEnd) return;
    const int i = rowOfCell(n, J), j = static_cast<int>(n - static_cast<long long>(i) * J);
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

// ------------------------------------------------------------------ materials / sources
// updateMaterials (flipsolver2d.cpp:1053-1075)
__global__ void __launch_bounds__(NT) updateMaterialsKernel(const float *__restrict__ sdf, int8_t *__restrict__ mat, long long nBegin,
                                                            long long N)
{
    const long long n = nBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n >= N) return;
    const int8_t m = mat[n];
    if (sdf[n] < 0.f)
    {
        if (matEmpty(m)) mat[n] = FS2D_FLUID;
    }
    else if (matStrictFluid(m))
    {
        mat[n] = FS2D_EMPTY;
    }
}

// afterTransfer: FlipSolver (flipsolver2d.cpp:390-410), smoke (flipsmokesolver.cpp:132-148), fire
// (flipfiresolver.cpp:18-33). U(i,j)/V(i,j) of a SOURCE cell only belong to that cell, so this is race free.
__global__ void __launch_bounds__(NT) afterTransferKernel(const int8_t *__restrict__ mat, const int32_t *__restrict__ emitterId,
                                                          const fs2d_source *__restrict__ sources, int I, int J, double dx,
                                                          float *__restrict__ viscosity, float *__restrict__ U,
                                                          float *__restrict__ V, uint8_t *__restrict__ uValid,
                                                          uint8_t *__restrict__ vValid, float *__restrict__ temperature,
                                                          float *__restrict__ concentration, float *__restrict__ fuel,
                                                          long long nBegin, long long nEnd)
{
    const long long n = nBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n >= nEnd || !matSource(mat[n])) return;
    const int i = rowOfCell(n, J), j = static_cast<int>(n - static_cast<long long>(i) * J);
    const fs2d_source s = sources[emitterId[n]];
    viscosity[n] = s.viscosity;
    if (s.transfer_velocity)
    {
        U[n] = static_cast<float>(static_cast<double>(s.velocity_x) / dx);
        V[static_cast<long long>(i) * (J + 1) + j] = static_cast<float>(static_cast<double>(s.velocity_y) / dx);
        uValid[n] = 1;
        vValid[static_cast<long long>(i) * (J + 1) + j] = 1;
    }
    if (concentration) concentration[n] = s.concentration;
    if (temperature) temperature[n] = s.temperature;
    if (fuel) fuel[n] = s.fuel;
}

// ------------------------------------------------------------------ BFS extrapolation
// simmath::breadthFirstExtrapolate (mathfuncs.cpp:152-215) as layer-synchronous sweeps: layer k
// (8-neighbour BFS distance from the valid samples) averages, in a double, the neighbours of
// smaller layer in the reference's neighbour order (linearindexable2d.h:69-78). Those are final
// before layer k starts, so one in-place sweep per layer reproduces the queue order exactly.
// Markers (0 = valid sample, 255 = unknown) and the bounding box {iMin, iMax, jMin, jMax} of the valid
// samples of both grids: layers can only appear within radius+1 samples of it, so the layer sweeps are
// restricted to that box instead of the whole grid.
// Rows [rowLo, rowHiU) of U and [rowLo, rowHiV) of V are swept (the whole grids without slabs).
__global__ void __launch_bounds__(NT) bfsInitKernel(const uint8_t *__restrict__ uValid, const uint8_t *__restrict__ vValid, int I,
                                                    int J, uint8_t *__restrict__ marker, int *__restrict__ bbox, int rowLo, int rowHiU,
                                                    int rowHiV)
{
    const long long NU = static_cast<long long>(I + 1) * J;
    const long long cntU = static_cast<long long>(rowHiU - rowLo) * J, cntV = static_cast<long long>(rowHiV - rowLo) * (J + 1);
    const long long t = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    int i = -1, j = -1;
    if (t < cntU)
    {
        const long long n = static_cast<long long>(rowLo) * J + t;
        const bool v = uValid[n] != 0;
        marker[n] = v ? 0 : 255;
        if (v)
        {
            i = rowOfCell(n, J);
            j = static_cast<int>(n - static_cast<long long>(i) * J);
        }
    }
    else if (t < cntU + cntV)
    {
        const long long m = static_cast<long long>(rowLo) * (J + 1) + (t - cntU);
        const long long n = NU + m;
        const bool v = vValid[m] != 0;
        marker[n] = v ? 0 : 255;
        if (v)
        {
            i = static_cast<int>(m / (J + 1));
            j = static_cast<int>(m - static_cast<long long>(i) * (J + 1));
        }
    }
    // one atomic per warp and bound
    int iMin = i < 0 ? 0x7fffffff : i, iMax = i, jMin = j < 0 ? 0x7fffffff : j, jMax = j;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
    {
        iMin = min(iMin, __shfl_xor_sync(0xffffffffu, iMin, o));
        iMax = max(iMax, __shfl_xor_sync(0xffffffffu, iMax, o));
        jMin = min(jMin, __shfl_xor_sync(0xffffffffu, jMin, o));
        jMax = max(jMax, __shfl_xor_sync(0xffffffffu, jMax, o));
    }
    // warps -> CTA in shared memory, then one global atomic per CTA and bound (hundreds of thousands of warp-level
    // atomics on four addresses were most of this kernel's time)
    __shared__ int box[4];
    if (threadIdx.x == 0)
    {
        box[0] = 0x7fffffff;
        box[1] = -1;
        box[2] = 0x7fffffff;
        box[3] = -1;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0 && iMax >= 0)
    {
        atomicMin(box + 0, iMin);
        atomicMax(box + 1, iMax);
        atomicMin(box + 2, jMin);
        atomicMax(box + 3, jMax);
    }
    __syncthreads();
    if (threadIdx.x == 0 && box[1] >= 0)
    {
        atomicMin(bbox + 0, box[0]);
        atomicMax(bbox + 1, box[1]);
        atomicMin(bbox + 2, box[2]);
        atomicMax(bbox + 3, box[3]);
    }
}

__device__ __forceinline__ void bfsLayerSampleAt(float *__restrict__ g, uint8_t *__restrict__ valid, uint8_t *__restrict__ marker,
                                                 int sI, int sJ, int i, int j, int k)
{
    const long long n = static_cast<long long>(i) * sJ + j;
    if (marker[n] != 255) return;
    bool hit = false;
    double avg = 0.0;
    int cnt = 0;
#pragma unroll
    for (int di = -1; di <= 1; di++)
#pragma unroll
        for (int dj = -1; dj <= 1; dj++)
        {
            if (di == 0 && dj == 0) continue;
            const int ni = i + di, nj = j + dj;
            if (ni < 0 || ni >= sI || nj < 0 || nj >= sJ) continue;
            const long long nn = static_cast<long long>(ni) * sJ + nj;
            const int m = marker[nn];
            if (m == k - 1) hit = true;
            if (m < k)
            {
                avg += static_cast<double>(g[nn]);
                cnt++;
            }
        }
    if (!hit) return;
    g[n] = static_cast<float>(avg / cnt);
    valid[n] = 1;
    marker[n] = static_cast<uint8_t>(k);
}

__device__ __forceinline__ void bfsLayerSample(float *__restrict__ g, uint8_t *__restrict__ valid, uint8_t *__restrict__ marker,
                                               int sI, int sJ, long long n, int k)
{
    const int i = rowOfCell(n, sJ);
    bfsLayerSampleAt(g, valid, marker, sI, sJ, i, static_cast<int>(n - static_cast<long long>(i) * sJ), k);
}

// One layer over the box [i0, i0+h) x [j0, j0+w) of both sample grids (clipped to each grid's extent).
__global__ void __launch_bounds__(NT) bfsLayerKernel(float *U, float *V, uint8_t *uValid, uint8_t *vValid, uint8_t *marker, int I,
                                                     int J, int k, int i0, int j0, int h, int w)
{
    const long long NU = static_cast<long long>(I + 1) * J;
    const long long box = static_cast<long long>(h) * w;
    const long long t = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (t >= 2 * box) return;
    const bool second = t >= box;
    const long long q = second ? t - box : t;
    const int i = i0 + rowOfCell(q, w), j = j0 + static_cast<int>(q % w);
    if (!second)
    {
        if (i <= I && j < J) bfsLayerSample(U, uValid, marker, I + 1, J, static_cast<long long>(i) * J + j, k);
    }
    else if (i < I && j <= J)
    {
        bfsLayerSample(V, vValid, marker + NU, I, J + 1, static_cast<long long>(i) * (J + 1) + j, k);
    }
}

// All layers in ONE cooperative launch: the bounding box is read on the device (no host round trip between
// bfsInitKernel and the sweeps) and a grid-wide barrier separates the layers. The box is swept ONCE, to list its unknown
// samples (at 4096^2: 3.1 M samples in the box, ~0.1 M of them unknown -- the fluid interior is all valid); the layers
// then only visit the list. (History at 4096^2, per call: 11 launches sweeping the box + a D2H synchronisation 0.35 ms;
// the same sweeps inside one cooperative kernel 0.31-0.34 ms whatever the index arithmetic, load batching or grid size
// -- a layer costs its grid barrier plus a chain of dependent L2 round trips over mostly known samples.)
constexpr int BFSV_THREADS = 1024;
__global__ void __launch_bounds__(BFSV_THREADS) bfsLayersKernel(float *U, float *V, uint8_t *uValid, uint8_t *vValid, uint8_t *marker, int I,
                                                                int J, int layers, const int *__restrict__ bbox, int grow, int rowLo,
                                                                int rowHiU, int32_t *__restrict__ queue, unsigned int *__restrict__ ctl,
                                                                unsigned int capacity)
{
    cg::grid_group grid = cg::this_grid();
    const int b0 = bbox[0], b1 = bbox[1], b2 = bbox[2], b3 = bbox[3];
    if (b1 < b0) return;  // no valid sample anywhere: the BFS has no seed (mathfuncs.cpp:171-189); uniform over the grid
    const int i0 = max(max(b0 - grow, 0), rowLo), i1 = min(min(b1 + grow, I), rowHiU - 1);
    const int j0 = max(b2 - grow, 0), j1 = min(b3 + grow, J);
    const int h = i1 - i0 + 1, w = j1 - j0 + 1;
    const unsigned int NU = static_cast<unsigned int>(I + 1) * static_cast<unsigned int>(J);
    // ---- the unknown samples of the box, as indices into the marker array (U samples first, V samples from NU on)
    for (int r = blockIdx.x; r < 2 * h; r += gridDim.x)
    {
        const bool second = r >= h;
        const int i = i0 + (second ? r - h : r);
        const int sI = second ? I : I + 1, sJ = second ? J + 1 : J;
        if (i >= sI) continue;
        const unsigned int rowBase = (second ? NU : 0u) + static_cast<unsigned int>(i) * sJ;
        const int jEnd = min(j0 + w, sJ);
        for (int jb = j0; jb < jEnd; jb += BFSV_THREADS)  // warp-uniform trip count (ballot below)
        {
            const int j = jb + static_cast<int>(threadIdx.x);
            const bool unknown = j < jEnd && marker[rowBase + j] == 255;
            const unsigned int votes = __ballot_sync(0xffffffffu, unknown);
            if (!votes) continue;
            const int lane = threadIdx.x & 31;
            unsigned int base = 0;
            if (lane == 0) base = atomicAdd(ctl, static_cast<unsigned int>(__popc(votes)));
            base = __shfl_sync(0xffffffffu, base, 0);
            const unsigned int slot = base + __popc(votes & ((1u << lane) - 1u));
            if (unknown && slot < capacity) queue[slot] = static_cast<int32_t>(rowBase + j);
        }
    }
    grid.sync();
    const unsigned int count = *reinterpret_cast<volatile unsigned int *>(ctl);
    const bool listed = count <= capacity;
    const unsigned int stride = gridDim.x * static_cast<unsigned int>(BFSV_THREADS);
    for (int k = 1; k <= layers; k++)
    {
        if (listed)
        {
            for (unsigned int t = blockIdx.x * static_cast<unsigned int>(BFSV_THREADS) + threadIdx.x; t < count; t += stride)
            {
                const unsigned int id = static_cast<unsigned int>(queue[t]);
                if (id < NU)
                {
                    const unsigned int i = id / static_cast<unsigned int>(J);
                    bfsLayerSampleAt(U, uValid, marker, I + 1, J, static_cast<int>(i), static_cast<int>(id - i * J), k);
                }
                else
                {
                    const unsigned int m = id - NU, i = m / static_cast<unsigned int>(J + 1);
                    bfsLayerSampleAt(V, vValid, marker + NU, I, J + 1, static_cast<int>(i), static_cast<int>(m - i * (J + 1)), k);
                }
            }
        }
        else
        {
            // more unknown samples than the list holds: sweep the box rows
            for (int r = blockIdx.x; r < 2 * h; r += gridDim.x)
            {
                const bool second = r >= h;
                const int i = i0 + (second ? r - h : r);
                const int sI = second ? I : I + 1, sJ = second ? J + 1 : J;
                if (i >= sI) continue;
                const int jEnd = min(j0 + w, sJ);
                for (int j = j0 + static_cast<int>(threadIdx.x); j < jEnd; j += BFSV_THREADS)
                {
                    if (!second)
                        bfsLayerSampleAt(U, uValid, marker, I + 1, J, i, j, k);
                    else
                        bfsLayerSampleAt(V, vValid, marker + NU, I, J + 1, i, j, k);
                }
            }
        }
        if (k < layers) grid.sync();
    }
}

// bfsInitKernel, sixteen samples per thread (gridSizeJ a multiple of 16, fewer than 2^31 samples, rowLo a multiple of 16):
// markers by byte-wise compare, the bounding-box reduction only in warps that saw a valid sample.
__global__ void __launch_bounds__(NT) bfsInitVecKernel(const uint8_t *__restrict__ uValid, const uint8_t *__restrict__ vValid, int I, int J,
                                                       uint8_t *__restrict__ marker, int *__restrict__ bbox, int rowLo, int rowHiU, int rowHiV)
{
    const unsigned int NU = static_cast<unsigned int>(I + 1) * static_cast<unsigned int>(J);
    const unsigned int uBegin = static_cast<unsigned int>(rowLo) * J, gU = static_cast<unsigned int>(rowHiU - rowLo) * J / 16u;
    const unsigned int vBegin = static_cast<unsigned int>(rowLo) * (J + 1), cntV = static_cast<unsigned int>(rowHiV - rowLo) * (J + 1);
    const unsigned int gV = (cntV + 15u) / 16u;
    const unsigned int t = blockIdx.x * static_cast<unsigned int>(NT) + threadIdx.x;
    int iMin = 0x7fffffff, iMax = -1, jMin = 0x7fffffff, jMax = -1;
    if (t < gU + gV)
    {
        const bool second = t >= gU;
        const unsigned int n = second ? vBegin + 16u * (t - gU) : uBegin + 16u * t;   // first sample of the group
        const unsigned int end = second ? vBegin + cntV : uBegin + gU * 16u;
        const uint8_t *src = second ? vValid : uValid;
        uint8_t *dst = second ? marker + NU : marker;
        const unsigned int sJ = second ? J + 1 : J;
        if (n + 16u <= end)
        {
            const uint4 v = *reinterpret_cast<const uint4 *>(src + n);
            uint4 m;  // 0xff where the flag byte is zero (unknown), 0 where the sample is valid
            m.x = __vcmpeq4(v.x, 0u);
            m.y = __vcmpeq4(v.y, 0u);
            m.z = __vcmpeq4(v.z, 0u);
            m.w = __vcmpeq4(v.w, 0u);
            *reinterpret_cast<uint4 *>(dst + n) = m;
            if ((m.x & m.y & m.z & m.w) != 0xffffffffu)
            {
                const unsigned int words[4] = {m.x, m.y, m.z, m.w};
#pragma unroll
                for (int b = 0; b < 16; b++)
                    if (!((words[b >> 2] >> (8 * (b & 3))) & 0xffu))
                    {
                        const unsigned int q = n + b;
                        const int i = rowOfCell(q, sJ), j = static_cast<int>(q - static_cast<unsigned int>(i) * sJ);
                        iMin = min(iMin, i);
                        iMax = max(iMax, i);
                        jMin = min(jMin, j);
                        jMax = max(jMax, j);
                    }
            }
        }
        else
            for (unsigned int q = n; q < end; q++)
            {
                const bool valid = src[q] != 0;
                dst[q] = valid ? 0 : 255;
                if (valid)
                {
                    const int i = rowOfCell(q, sJ), j = static_cast<int>(q - static_cast<unsigned int>(i) * sJ);
                    iMin = min(iMin, i);
                    iMax = max(iMax, i);
                    jMin = min(jMin, j);
                    jMax = max(jMax, j);
                }
            }
    }
    if (!__any_sync(0xffffffffu, iMax >= 0)) return;
    iMin = __reduce_min_sync(0xffffffffu, iMin);
    iMax = __reduce_max_sync(0xffffffffu, iMax);
    jMin = __reduce_min_sync(0xffffffffu, jMin);
    jMax = __reduce_max_sync(0xffffffffu, jMax);
    if ((threadIdx.x & 31) == 0)
    {
        atomicMin(bbox + 0, iMin);
        atomicMax(bbox + 1, iMax);
        atomicMin(bbox + 2, jMin);
        atomicMax(bbox + 3, jMax);
    }
}

// extrapolateLevelsetInside / Outside (flipsolver2d.cpp:1433-1558): same BFS with unbounded radius and -1 / +1 per
// layer: a cell of layer k takes the mean (double) of its 8-neighbours in lower layers plus the step. Layers are
// processed one after the other (their values depend on the previous layer) but each cell is visited ONCE: layer k is a
// list of cells; processing a cell claims its still unmarked neighbours for layer k + 1 (atomicCAS on the marker) and
// appends them to the list. One cooperative launch walks all layers with one grid-wide barrier per layer -- at 4096^2
// the level set has thousands of layers; sweeping the bounding box of the unmarked cells once per layer, as the first
// version did, cost 0.5 s per nbflip substep. The value of a cell does not depend on the order inside its layer
// (SURVEY appendix A-9), so the arbitrary list order is harmless: results are bit-identical to the reference.
// queue: one slot per cell; ctl[0] = tail (cells appended so far).
__global__ void __launch_bounds__(NT) sdfMarkKernel(const float *__restrict__ sdf, int32_t *__restrict__ marker, long long nBegin,
                                                    long long nEnd, int inside, float maxSdf)
{
    const long long n = nBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n >= nEnd) return;
    const float v = sdf[n];
    const bool known = inside ? (v > 0.f) : (v < maxSdf);
    marker[n] = known ? 0 : 0x7fffffff;
}

// layer 1: unmarked cells with a known 8-neighbour (flipsolver2d.cpp:1450-1468). Rows [rowLo, rowHi) only (the whole grid
// without slabs); cells outside that range do not exist for the walk.
__global__ void __launch_bounds__(NT) sdfFirstLayerKernel(int32_t *__restrict__ marker, int J, int rowLo, int rowHi,
                                                          int32_t *__restrict__ queue, unsigned int *__restrict__ ctl)
{
    const long long n = static_cast<long long>(rowLo) * J + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n >= static_cast<long long>(rowHi) * J) return;
    if (marker[n] == 0) return;
    const int i = rowOfCell(n, J), j = static_cast<int>(n - static_cast<long long>(i) * J);
    bool hit = false;
#pragma unroll
    for (int di = -1; di <= 1; di++)
#pragma unroll
        for (int dj = -1; dj <= 1; dj++)
        {
            const int ni = i + di, nj = j + dj;
            if ((di == 0 && dj == 0) || ni < rowLo || ni >= rowHi || nj < 0 || nj >= J) continue;
            if (marker[static_cast<long long>(ni) * J + nj] == 0) hit = true;
        }
    if (!hit) return;
    // neighbours only test for == 0 in this kernel, so writing 1 here cannot change what another thread decides
    marker[n] = 1;
    queue[atomicAdd(ctl, 1u)] = static_cast<int32_t>(n);
}

// Band mode, inside walk: cells the walk did not reach within `layers` layers (still unmarked, or claimed for the next
// layer) get one value below everything the band holds.
__global__ void __launch_bounds__(NT) sdfClampKernel(float *__restrict__ sdf, const int32_t *__restrict__ marker, long long nBegin,
                                                     long long nEnd, int layers, float value)
{
    const long long n = nBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n >= nEnd) return;
    if (marker[n] > layers) sdf[n] = value;
}

// A cooperative grid of one CTA per SM walks the layers (measured at 4096^2 nbflip, ~3400 layers per substep: one
// 8-CTA cluster with the hardware cluster barrier 41 ms, this grid 2x faster -- a frontier is several thousand cells
// and each visit is a chain of dependent L2 / DRAM round trips, so the number of threads in flight matters more than
// the barrier).
constexpr int BFS_THREADS = 256;

__global__ void __launch_bounds__(BFS_THREADS) sdfExtrapolateKernel(float *sdf, int32_t *marker, int rowLo, int rowHi, int J, float step,
                                                                    int32_t *queue, unsigned int *ctl, int maxLayers)
{
    cg::grid_group grid = cg::this_grid();
    const unsigned int stride = gridDim.x * BFS_THREADS;
    const unsigned int me = blockIdx.x * BFS_THREADS + threadIdx.x;
    unsigned int begin = 0, end = *reinterpret_cast<volatile unsigned int *>(ctl);
    for (int k = 1; begin < end && (maxLayers <= 0 || k <= maxLayers); k++)
    {
        for (unsigned int t0 = begin + (me & ~31u); t0 < end; t0 += stride)  // warp-uniform trip count (ballots below)
        {
            const unsigned int t = t0 + (threadIdx.x & 31u);
            const bool valid = t < end;
            const long long n = valid ? queue[t] : 0;
            const int i = rowOfCell(n, J), j = static_cast<int>(n - static_cast<long long>(i) * J);
            // markers and values of all eight neighbours in flight together; the values of cells that turn out not to
            // be in a lower layer are simply not used
            int m[8];
            float v[8];
            long long nn[8];
#pragma unroll
            for (int q = 0; q < 8; q++)
            {
                // (i-1,j-1) (i-1,j) (i-1,j+1) (i,j-1) (i,j+1) (i+1,j-1) (i+1,j) (i+1,j+1): the order of getNeighborhood
                const int di = (q < 3) ? -1 : (q < 5 ? 0 : 1);
                const int dj = (q < 3) ? q - 1 : (q == 3 ? -1 : (q == 4 ? 1 : q - 6));
                const int ni = i + di, nj = j + dj;
                const bool in = valid && ni >= rowLo && ni < rowHi && nj >= 0 && nj < J;
                nn[q] = in ? static_cast<long long>(ni) * J + nj : -1;
                m[q] = in ? marker[nn[q]] : -1;
                v[q] = in ? sdf[nn[q]] : 0.f;
            }
            // claim unmarked neighbours for the next layer: all eight compare-and-swaps in flight together, exactly one
            // claimant wins each cell; then the winners of the whole warp reserve their queue slots with ONE atomic (a
            // chain of dependent atomics per neighbour, or thousands of single appends to the same counter, would set
            // the time per layer)
            unsigned int wonMask = 0;
#pragma unroll
            for (int q = 0; q < 8; q++)
                if (nn[q] >= 0 && m[q] == 0x7fffffff && atomicCAS(marker + nn[q], 0x7fffffff, k + 1) == 0x7fffffff) wonMask |= 1u << q;
            const unsigned int lane = threadIdx.x & 31u;
            const unsigned int mine = __popc(wonMask);
            unsigned int incl = mine;  // inclusive warp scan of the win counts
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const unsigned int up = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= static_cast<unsigned int>(o)) incl += up;
            }
            const unsigned int total = __shfl_sync(0xffffffffu, incl, 31);
            unsigned int base = 0;
            if (total)
            {
                if (lane == 0) base = atomicAdd(ctl, total);
                base = __shfl_sync(0xffffffffu, base, 0) + incl - mine;
            }
            double avg = 0.0;
            int cnt = 0;
#pragma unroll
            for (int q = 0; q < 8; q++)
            {
                if (wonMask & (1u << q)) queue[base++] = static_cast<int32_t>(nn[q]);
                if (nn[q] >= 0 && m[q] != 0x7fffffff && m[q] < k)
                {
                    avg += static_cast<double>(v[q]);  // a lower layer: final since the last barrier
                    cnt++;
                }
            }
            if (valid) sdf[n] = static_cast<float>(avg / cnt + static_cast<double>(step));
        }
        __threadfence();
        grid.sync();
        begin = end;
        end = *reinterpret_cast<volatile unsigned int *>(ctl);
    }
}

// ------------------------------------------------------------------ body forces
// applyBodyForces (flipsolver2d.cpp:1222-1241): every U and V sample gets factor * g
// U samples [uBegin, uBegin + NU) and V samples [vBegin, vBegin + NV) (whole grids without slabs)
__global__ void __launch_bounds__(NT) bodyForceKernel(float *__restrict__ U, float *__restrict__ V, long long uBegin, long long NU,
                                                      long long vBegin, long long NV, float addU, float addV)
{
    const long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n < NU)
        U[uBegin + n] = faddr(U[uBegin + n], addU);
    else if (n < NU + NV)
        V[vBegin + n - NU] = faddr(V[vBegin + n - NU], addV);
}

// FlipSmokeSolver::applyBodyForces (flipsmokesolver.cpp:23-52)
__global__ void __launch_bounds__(NT) smokeBodyForceKernel(float *__restrict__ U, float *__restrict__ V, int J, long long uBegin, long long NU,
                                                           long long vBegin, long long NV,
                                                           GridView temperature, GridView concentration, float sootWeight,
                                                           float buoyancyInfluence, float ambient, float gx, float gy, float factor)
{
    // samples [uBegin, uBegin + NU) of U and [vBegin, vBegin + NV) of V (row slabs: the rows with valid inputs)
    long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n < NU)
    {
        n += uBegin;
        const int i = rowOfCell(n, J), j = static_cast<int>(n - static_cast<long long>(i) * J);
        const float x = static_cast<float>(i), y = faddr(static_cast<float>(j), 0.5f);
        const float td = fsubr(gridLerp(temperature, x, y), ambient);
        const float c = gridLerp(concentration, x, y);
        const float a = fmulr(fmulr(fsubr(fmulr(sootWeight, c), fmulr(buoyancyInfluence, td)), gx), factor);
        U[n] = faddr(U[n], a);
    }
    else if (n < NU + NV)
    {
        const long long m = vBegin + n - NU;
        const int i = static_cast<int>(m / (J + 1)), j = static_cast<int>(m - static_cast<long long>(i) * (J + 1));
        const float x = faddr(static_cast<float>(i), 0.5f), y = static_cast<float>(j);
        const float td = fsubr(gridLerp(temperature, x, y), ambient);
        const float c = gridLerp(concentration, x, y);
        const float a = fmulr(fmulr(fsubr(fmulr(sootWeight, c), fmulr(buoyancyInfluence, td)), gy), factor);
        V[m] = faddr(V[m], a);
    }
}

// ------------------------------------------------------------------ right-hand sides
// calcPressureRhs (flipsolver2d.cpp:962-991; smoke: flipsmokesolver.cpp:73-102) with divergenceAt (:759-764)
__global__ void __launch_bounds__(NT) pressureRhsKernel(const float *__restrict__ U, const float *__restrict__ V,
                                                        const float *__restrict__ divCtl, const int8_t *__restrict__ mat, int I,
                                                        int J, int smokeRows, double scale, double *__restrict__ rhs, long long nBegin,
                                                        long long nEnd)
{
    const long long n = nBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n >= nEnd) return;
    const int i = rowOfCell(n, J), j = static_cast<int>(n - static_cast<long long>(i) * J);
    const int8_t m = mat[n];
    const bool row = smokeRows ? !matSolid(m) : matFluid(m);
    if (!row)
    {
        rhs[n] = 0.0;
        return;
    }
    const float u0 = U[n], u1 = U[n + J];
    const long long vn = static_cast<long long>(i) * (J + 1) + j;
    const float v0 = V[vn], v1 = V[vn + 1];
    const float div = faddr(fsubr(faddr(fsubr(u1, u0), v1), v0), divCtl[n]);
    double val = __dmul_rn(-scale, static_cast<double>(div));
    const double sIm = matSolid(matAt(mat, I, J, i - 1, j)) ? 1.0 : 0.0;
    const double sIp = matSolid(matAt(mat, I, J, i + 1, j)) ? 1.0 : 0.0;
    const double sJm = matSolid(matAt(mat, I, J, i, j - 1)) ? 1.0 : 0.0;
    const double sJp = matSolid(matAt(mat, I, J, i, j + 1)) ? 1.0 : 0.0;
    val = __dsub_rn(val, __dmul_rn(__dmul_rn(sIm, scale), static_cast<double>(u0)));
    val = __dadd_rn(val, __dmul_rn(__dmul_rn(sIp, scale), static_cast<double>(u1)));
    val = __dsub_rn(val, __dmul_rn(__dmul_rn(sJm, scale), static_cast<double>(v0)));
    val = __dadd_rn(val, __dmul_rn(__dmul_rn(sJp, scale), static_cast<double>(v1)));
    rhs[n] = val;
}

// calcDensityCorrectionRhs (flipsolver2d.cpp:993-1011)
__global__ void __launch_bounds__(NT) densityRhsKernel(const float *__restrict__ density, const int8_t *__restrict__ mat,
                                                       long long nBegin, long long N, double scale, double restDensity,
                                                       double *__restrict__ rhs)
{
    const long long n = nBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n >= N) return;
    if (!matFluid(mat[n]))
    {
        rhs[n] = 0.0;
        return;
    }
    double r = __ddiv_rn(static_cast<double>(density[n]), restDensity);
    r = r < 0.5 ? 0.5 : (r > 1.5 ? 1.5 : r);
    rhs[n] = __dmul_rn(scale, __dsub_rn(1.0, r));
}

// applyPressureThreadU/V (flipsolver2d.cpp:1133-1193; smoke: flipsmokesolver.cpp:262-322)
__global__ void __launch_bounds__(NT) applyPressureKernel(const double *__restrict__ p, const int8_t *__restrict__ mat, int I,
                                                          int J, int smokeRows, double scale, float *__restrict__ U,
                                                          float *__restrict__ V, uint8_t *__restrict__ uValid,
                                                          uint8_t *__restrict__ vValid, float *__restrict__ testGrid, long long nBegin,
                                                          long long nEnd)
{
    const long long n = nBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n >= nEnd) return;
    const int i = rowOfCell(n, J), j = static_cast<int>(n - static_cast<long long>(i) * J);
    const double pc = p[n];
    const int8_t mc = mat[n];
    {
        const int8_t mn = matAt(mat, I, J, i - 1, j);
        const double pn = i > 0 ? p[n - J] : 0.0;
        const bool active = smokeRows ? (!matSolid(mn) || !matSolid(mc)) : (matFluid(mn) || matFluid(mc));
        if (active)
        {
            if (matSolid(mn) || matSolid(mc))
                U[n] = 0.f;
            else
                U[n] = static_cast<float>(__dsub_rn(static_cast<double>(U[n]), __dmul_rn(scale, __dsub_rn(pc, pn))));
        }
        else
        {
            uValid[n] = 0;
        }
    }
    {
        const long long vn = static_cast<long long>(i) * (J + 1) + j;
        const int8_t mn = matAt(mat, I, J, i, j - 1);
        const double pn = j > 0 ? p[n - 1] : 0.0;
        const bool active = smokeRows ? (!matSolid(mn) || !matSolid(mc)) : (matFluid(mn) || matFluid(mc));
        if (active)
        {
            if (matSolid(mn) || matSolid(mc))
                V[vn] = 0.f;
            else
                V[vn] = static_cast<float>(__dsub_rn(static_cast<double>(V[vn]), __dmul_rn(scale, __dsub_rn(pc, pn))));
        }
        else
        {
            vValid[vn] = 0;
        }
    }
    if (testGrid) testGrid[n] = static_cast<float>(pc / 100.0);  // flipsmokesolver.cpp:247-250
}

// updateVelocityFromSolids (flipsolver2d.cpp:1077-1104) with validSolidNeighborIds (:766-786) and
// u/vSampleAffectedBySolid (materialgrid.cpp:152-164)
__global__ void __launch_bounds__(NT) solidFrictionKernel(const int32_t *__restrict__ solidId, const float *__restrict__ friction,
                                                          const int8_t *__restrict__ mat, int I, int J, float *__restrict__ U,
                                                          float *__restrict__ V, long long nBegin, long long nEnd)
{
    const long long n = nBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n >= nEnd) return;
    const int i = rowOfCell(n, J), j = static_cast<int>(n - static_cast<long long>(i) * J);
    float avg = 0.f;
    int cnt = 0;
    for (int di = -1; di <= 1; di++)
        for (int dj = -1; dj <= 1; dj++)
        {
            const int ni = i + di, nj = j + dj;
            if (ni < 0 || ni >= I || nj < 0 || nj >= J) continue;
            const int id = solidId[static_cast<long long>(ni) * J + nj];
            if (id != -1)
            {
                avg = faddr(avg, friction[id]);
                cnt++;
            }
        }
    if (cnt == 0) return;
    avg = __fdiv_rn(avg, static_cast<float>(cnt));
    const float keep = fsubr(1.f, avg);
    auto S = [&](int a, int b) { return matSolid(matAt(mat, I, J, a, b)); };
    if (S(i, j + 1) || S(i - 1, j + 1) || S(i, j) || S(i - 1, j) || S(i, j - 1) || S(i - 1, j - 1)) U[n] = fmulr(U[n], keep);
    if (S(i + 1, j) || S(i + 1, j - 1) || S(i, j) || S(i, j - 1) || S(i - 1, j) || S(i - 1, j - 1))
    {
        const long long vn = static_cast<long long>(i) * (J + 1) + j;
        V[vn] = fmulr(V[vn], keep);
    }
}

// ------------------------------------------------------------------ semi-Lagrangian advection
// eulerAdvectionThread (flipsolver2d.cpp:340-351): back-trace RK4(-dt) from the INTEGER sample index
// (not the staggered sample position) and interpolate the input grid with its own offset / OOB policy.
// (One value = gridLerp(grid, rk4(velocity, (i, j), -dt)); the kernels below evaluate the trace once for all the grids
// that are advected from the same start point.)
// The smoke / fire parameter grids (soot, temperature, fuel) in one pass: same cells, same integer start points, one trace.
__global__ void __launch_bounds__(NT) smokeAdvectKernel(GridView concentration, GridView temperature, GridView fuel, int withFuel, VelocityView vel,
                                                        float dt, float *__restrict__ outC, float *__restrict__ outT, float *__restrict__ outF,
                                                        long long nBegin, long long nEnd)
{
    const long long n = nBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n >= nEnd) return;
    const int i = rowOfCell(n, concentration.sizeJ), j = static_cast<int>(n - static_cast<long long>(i) * concentration.sizeJ);
    const float2 prev = rk4(vel, make_float2(static_cast<float>(i), static_cast<float>(j)), -dt);
    outC[n] = gridLerp(concentration, prev.x, prev.y);
    outT[n] = gridLerp(temperature, prev.x, prev.y);
    if (withFuel) outF[n] = gridLerp(fuel, prev.x, prev.y);
}

// NBFlipSolver::advect, the four grids in one pass (nbflipsolver.cpp:83-103). eulerAdvectionThread back-traces from the INTEGER
// sample index whatever the offset of the grid it is called for, so the U, V, level-set and viscosity samples with the same
// (i, j) share one and the same RK4 trace -- 16 bilinear velocity look-ups, against one look-up per advected value. Sample
// (i, j) of the (I + 1) x (J + 1) index space: U for j < J, V / level set / viscosity for cell rows below rowHi. Same
// arithmetic per value as one pass per grid, hence the same bits.
__global__ void __launch_bounds__(NT) nbflipAdvectKernel(VelocityView vel, GridView sdf, GridView visc, float dt, int J, int rowLo,
                                                         int rowHiU, int rowHi, float *__restrict__ advU, float *__restrict__ advV,
                                                         float *__restrict__ advSdf, float *__restrict__ advVisc)
{
    const long long t = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    const int W = J + 1;
    if (t >= static_cast<long long>(rowHiU - rowLo) * W) return;
    const int r = rowOfCell(t, W), j = static_cast<int>(t - static_cast<long long>(r) * W), i = rowLo + r;
    const bool cellRow = i < rowHi;
    if (j == J && !cellRow) return;  // the corner sample belongs to no grid
    const float2 prev = rk4(vel, make_float2(static_cast<float>(i), static_cast<float>(j)), -dt);
    if (j < J) advU[static_cast<long long>(i) * J + j] = gridLerp(vel.u, prev.x, prev.y);
    if (cellRow)
    {
        advV[static_cast<long long>(i) * W + j] = gridLerp(vel.v, prev.x, prev.y);
        if (j < J)
        {
            const long long n = static_cast<long long>(i) * J + j;
            advSdf[n] = gridLerp(sdf, prev.x, prev.y);
            advVisc[n] = gridLerp(visc, prev.x, prev.y);
        }
    }
}

// NBFlipSolver::updateGridFromSources + combineLevelset (nbflipsolver.cpp:329-376)
__global__ void __launch_bounds__(NT) nbSourcesLevelsetKernel(const float *__restrict__ sourceSdf, const int32_t *__restrict__ sourceId,
                                                              const fs2d_source *__restrict__ sources,
                                                              const float *__restrict__ advSdf, long long nBegin, long long N,
                                                              float *__restrict__ fluidSdf, float *__restrict__ viscosity,
                                                              float *__restrict__ advViscosity)
{
    const long long n = nBegin + blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n >= N) return;
    const float d = sourceSdf[n];
    float f = fminf(d, fluidSdf[n]);
    if (d < 0.f && sourceId[n] != -1)
    {
        const float v = sources[sourceId[n]].viscosity;
        viscosity[n] = v;
        advViscosity[n] = v;
    }
    fluidSdf[n] = fminf(faddr(advSdf[n], 1.f), f);
}

// combineVelocityGrid + combineCenteredGrids (nbflipsolver.cpp:378-427), rule nbCombine (nbflipsolver.h:53-63)
__global__ void __launch_bounds__(NT) nbCombineKernel(GridView fluidSdf, int I, int J, const float *__restrict__ advU,
                                                      const float *__restrict__ advV, const float *__restrict__ advViscosity,
                                                      float *__restrict__ U, float *__restrict__ V, float *__restrict__ viscosity,
                                                      float *__restrict__ testGrid, float band, int rowLo, int rowHiU, int rowHi)
{
    // rows [rowLo, rowHiU) of U, [rowLo, rowHi) of V and of the centred grids (the whole grids without slabs)
    const long long NU = static_cast<long long>(rowHiU - rowLo) * J, NV = static_cast<long long>(rowHi - rowLo) * (J + 1),
                    N = static_cast<long long>(rowHi - rowLo) * J;
    const long long n = blockIdx.x * static_cast<long long>(NT) + threadIdx.x;
    if (n < NU)
    {
        const long long g = n + static_cast<long long>(rowLo) * J;
        const int i = rowOfCell(g, J), j = static_cast<int>(g - static_cast<long long>(i) * J);
        const float sdf = gridLerp(fluidSdf, static_cast<float>(i), faddr(0.5f, static_cast<float>(j)));
        if (!(sdf > band)) U[g] = advU[g];
    }
    else if (n < NU + NV)
    {
        const long long m = n - NU + static_cast<long long>(rowLo) * (J + 1);
        const int i = static_cast<int>(m / (J + 1)), j = static_cast<int>(m - static_cast<long long>(i) * (J + 1));
        const float sdf = gridLerp(fluidSdf, faddr(0.5f, static_cast<float>(i)), static_cast<float>(j));
        if (!(sdf > band)) V[m] = advV[m];
    }
    else if (n < NU + NV + N)
    {
        const long long m = n - NU - NV + static_cast<long long>(rowLo) * J;
        const int i = rowOfCell(m, J), j = static_cast<int>(m - static_cast<long long>(i) * J);
        // Vec3(0.5f + i, j + 0.5): the second component is evaluated in double and narrowed
        const float sdf = gridLerp(fluidSdf, faddr(0.5f, static_cast<float>(i)), static_cast<float>(static_cast<double>(j) + 0.5));
        if (!(sdf > band)) viscosity[m] = advViscosity[m];
        testGrid[m] = viscosity[m];
    }
}
}  // namespace

int gridBuildMatrix(Ctx *ctx)
{
    // NBFlip builds the system AFTER gridUpdate (nbflipsolver.cpp:43-47), which rewrote materials, combined the advected
    // velocities in and applied the body forces on the owned rows: in slab mode the matrix rows, the right-hand side and
    // the pressure gradient at a slab boundary need those on the halo rows too. (Water builds it from the previous
    // substep's materials, whose halo the last velocity extrapolation refreshed.)
    if (ctx->p.sim_type == FS2D_SIM_NBFLIP) FS2D_TRY(slabExchangeVelocity(ctx, true));
    const int smokeRows = (ctx->p.sim_type == FS2D_SIM_SMOKE || ctx->p.sim_type == FS2D_SIM_FIRE) ? 1 : 0;
    const CellRange cr = cellRange(ctx, slabOwn(ctx));
    buildMatrixKernel<<<divUp(cr.count(), NT), NT, 0, ctx->stream>>>(ctx->material, ctx->I, ctx->J, smokeRows, ctx->rowInfo,
                                                                      ctx->preInfo, cr.begin, cr.end);
    ctx->launches++;
    // scale = dt / (rho dx^2) (flipsolver2d.cpp:799), float dt promoted to double
    ctx->matrixScale = static_cast<double>(ctx->stepDt) / (ctx->p.fluid_density * ctx->p.dx * ctx->p.dx);
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

GridView solidSdfView(const Ctx *c);
GridView fluidSdfView(const Ctx *c);
GridView viscosityView(const Ctx *c);
GridView temperatureView(const Ctx *c);
GridView concentrationView(const Ctx *c);
GridView fuelView(const Ctx *c);

static bool isSmoke(const Ctx *ctx) { return ctx->p.sim_type == FS2D_SIM_SMOKE || ctx->p.sim_type == FS2D_SIM_FIRE; }

// Row slabs, smoke / fire: halo rows of the temperature, soot (and fuel) grids from the row neighbours.
static int smokeHalo(Ctx *ctx)
{
    if (!ctx->slab.enabled || ctx->slab.world == 1) return FS2D_OK;
    const void *arr[3] = {ctx->temperature, ctx->concentration, ctx->fuel};
    const size_t rb[3] = {sizeof(float) * ctx->J, sizeof(float) * ctx->J, sizeof(float) * ctx->J};
    return slabExchangeFields(ctx, arr, rb, ctx->p.sim_type == FS2D_SIM_FIRE ? 3 : 2);
}


int gridUpdateMaterials(Ctx *ctx)
{
    const CellRange cr = cellRange(ctx, slabOwn(ctx));
    updateMaterialsKernel<<<divUp(cr.count(), NT), NT, 0, ctx->stream>>>(ctx->fluidSdf, ctx->material, cr.begin, cr.end);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int gridAfterTransfer(Ctx *ctx)
{
    cudaStream_t st = ctx->stream;
    if (ctx->numSources > 0)
    {
        const CellRange cr = cellRange(ctx, slabOwn(ctx));
        afterTransferKernel<<<divUp(cr.count(), NT), NT, 0, st>>>(ctx->material, ctx->emitterId, ctx->sources, ctx->I, ctx->J, ctx->p.dx,
                                                                ctx->viscosity, ctx->U, ctx->V, ctx->uValid, ctx->vValid,
                                                                isSmoke(ctx) ? ctx->temperature : nullptr,
                                                                isSmoke(ctx) ? ctx->concentration : nullptr,
                                                                ctx->p.sim_type == FS2D_SIM_FIRE ? ctx->fuel : nullptr, cr.begin, cr.end);
        ctx->launches++;
    }
    if (isSmoke(ctx)) FS2D_CUDA(cudaMemsetAsync(ctx->divergenceControl, 0, sizeof(float) * ctx->N, st));  // flipsmokesolver.cpp:135
    if (ctx->p.sim_type == FS2D_SIM_NBFLIP)
    {
        // NBFlipSolver::afterTransfer (nbflipsolver.cpp:111-116). Slab mode: the owned rows; combineAdvectedGrids samples
        // the level set combineLevelset has just rewritten at staggered positions, i.e. one row beyond the slab, so the
        // halo rows of the level set are refreshed in between.
        const SlabRows own = slabOwn(ctx);
        const CellRange cr = cellRange(ctx, own);
        nbSourcesLevelsetKernel<<<divUp(cr.count(), NT), NT, 0, st>>>(ctx->sourceSdf, ctx->sourceSdfId, ctx->sources, ctx->advSdf, cr.begin,
                                                                    cr.end, ctx->fluidSdf, ctx->viscosity, ctx->advViscosity);
        ctx->launches++;
        {
            const void *arr[1] = {ctx->fluidSdf};
            const size_t rb[1] = {sizeof(float) * ctx->J};
            FS2D_TRY(slabExchangeFields(ctx, arr, rb, 1));
        }
        const int rowHiU = own.hi == ctx->I ? ctx->I + 1 : own.hi;
        const long long samples = static_cast<long long>(rowHiU - own.lo) * ctx->J + static_cast<long long>(own.hi - own.lo) * (2ll * ctx->J + 1);
        nbCombineKernel<<<divUp(samples, NT), NT, 0, st>>>(fluidSdfView(ctx), ctx->I, ctx->J, ctx->advU, ctx->advV, ctx->advViscosity, ctx->U,
                                                         ctx->V, ctx->viscosity, ctx->testGrid, -2.f, own.lo, rowHiU, own.hi);
        ctx->launches++;
    }
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int gridExtrapolateVelocity(Ctx *ctx, int radius)
{
    KernelGroupTimer kgt(ctx, FS2D_KGROUP_EXTRAPOLATE);
    // slab mode: both call sites of a substep (after the transfer, after the projection) follow stages that
    // rewrote U / V / validity / material on the owned rows only -> refresh the halo copies first
    FS2D_TRY(slabExchangeVelocity(ctx, true));
    uint8_t *marker = reinterpret_cast<uint8_t *>(ctx->markers);
    cudaStream_t st = ctx->stream;
    int *bbox = reinterpret_cast<int *>(ctx->d_counter) + 16;  // ints 16..19 of the scalar scratch
    const int init[4] = {0x7fffffff, -1, 0x7fffffff, -1};
    FS2D_CUDA(cudaMemcpyAsync(bbox, init, sizeof(init), cudaMemcpyHostToDevice, st));
    // slab mode: the sweep covers the owned rows plus the exchanged halo. Values on the outer halo rows lack
    // their outside neighbours; that error moves one row per layer, so after radius+1 layers the owned rows and
    // the inner halo - (radius+2) rows are exact.
    const SlabRows reg = slabExt(ctx, ctx->slab.halo);
    const int rowHiU = reg.hi == ctx->I ? ctx->I + 1 : reg.hi;
    const long long samples = static_cast<long long>(rowHiU - reg.lo) * ctx->J + static_cast<long long>(reg.hi - reg.lo) * (ctx->J + 1);
    if (ctx->J % 16 == 0 && reg.lo % 16 == 0 && ctx->NU + ctx->NV < (1ll << 31))
        bfsInitVecKernel<<<divUp(samples / 16 + 2, NT), NT, 0, st>>>(ctx->uValid, ctx->vValid, ctx->I, ctx->J, marker, bbox, reg.lo, rowHiU, reg.hi);
    else
        bfsInitKernel<<<divUp(samples, NT), NT, 0, st>>>(ctx->uValid, ctx->vValid, ctx->I, ctx->J, marker, bbox, reg.lo, rowHiU, reg.hi);
    ctx->launches++;
    const int grow = radius + 2;
    static const bool stepwiseBfs = std::getenv("FS2D_BFS_STEPWISE") != nullptr;  // A/B: one launch per layer, bounding box via the host
    if (!stepwiseBfs && ctx->NU + ctx->NV < (1ll << 31))
    {
        if (!ctx->bfsQueue)
        {
            FS2D_CUDA(cudaMalloc(reinterpret_cast<void **>(&ctx->bfsQueue), sizeof(int32_t) * static_cast<size_t>(ctx->N)));
            FS2D_CUDA(cudaMalloc(reinterpret_cast<void **>(&ctx->bfsCtl), 64));
        }
        FS2D_CUDA(cudaMemsetAsync(ctx->bfsCtl, 0, 64, st));
        const int share = (ctx->slab.enabled && ctx->slab.world > 1) ? std::max(ctx->slab.share, 1) : 1;
        float *U = ctx->U, *V = ctx->V;
        uint8_t *uv = ctx->uValid, *vv = ctx->vValid;
        int I = ctx->I, J = ctx->J, layers = radius + 1, g = grow, rowLo = reg.lo, rowHi = rowHiU;
        const int *bb = bbox;
        int32_t *queue = ctx->bfsQueue;
        unsigned int *ctl = ctx->bfsCtl;
        unsigned int capacity = static_cast<unsigned int>(ctx->N);
        void *args[] = {&U, &V, &uv, &vv, &marker, &I, &J, &layers, &bb, &g, &rowLo, &rowHi, &queue, &ctl, &capacity};
        const cudaError_t le = cudaLaunchCooperativeKernel(reinterpret_cast<void *>(bfsLayersKernel), dim3(std::max(1, ctx->smCount / share)),
                                                           dim3(BFSV_THREADS), args, 0, st);
        if (le == cudaSuccess)
        {
            ctx->launches++;
            return FS2D_OK;
        }
        cudaGetLastError();  // the grid cannot be co-resident here: one launch per layer below
    }
    int box[4];
    FS2D_CUDA(fs2dCopyToHost(ctx, box, bbox, sizeof(box)));
    FS2D_CUDA(cudaStreamSynchronize(st));
    if (box[1] < box[0]) return FS2D_OK;  // no valid sample anywhere: the BFS has no seed (mathfuncs.cpp:171-189)
    const int i0 = std::max({box[0] - grow, 0, reg.lo}), i1 = std::min({box[1] + grow, ctx->I, rowHiU - 1});
    const int j0 = std::max(box[2] - grow, 0), j1 = std::min(box[3] + grow, ctx->J);
    const int h = i1 - i0 + 1, w = j1 - j0 + 1;
    const int blocks = divUp(2ll * h * w, NT);
    for (int k = 1; k <= radius + 1; k++)
        bfsLayerKernel<<<blocks, NT, 0, st>>>(ctx->U, ctx->V, ctx->uValid, ctx->vValid, marker, ctx->I, ctx->J, k, i0, j0, h, w);
    ctx->launches += radius + 1;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int gridExtrapolateSdfNow(Ctx *ctx, bool inside, float *field, int band, SlabRows rows);

// ---- how far the level-set walks go
// extrapolateLevelsetInside / Outside have unbounded radius: thousands of strictly dependent layers at 4096^2, and
// nothing to shard (a layer is a thin front).
//  * Water / smoke / fire: extrapolateLevelsetInside only rewrites the level set BELOW the surface and nothing in the
//    substep reads it (updateMaterials has already run, the next updateSdf overwrites it), so the walk is deferred
//    until somebody asks for the grid (fs2d_download_grid / fs2d_grid_device_ptr) -- same values, off the hot path.
//  * NBFlip reads the level set every substep, but only near the interface: pruneNarrowBand / reseedParticles compare
//    it with -3 and -1 at particle positions (nbflipsolver.cpp:227-253,118-201), combineAdvectedGrids with -2
//    (:378-427), updateMaterials with 0, and the semi-Lagrangian step re-samples it at most cflNumber + 1 = 6 cells
//    away; everything else only has to keep its sign and stay beyond those thresholds. A layer's values depend on lower
//    layers only, so walking `sdfBand` layers (default 24) leaves every cell within 24 layers of the interface
//    bit-identical to the unbounded walk. Beyond the band the INSIDE walk writes -(band + 1) (the unbounded walk:
//    between -layer and -layer + 1.5) and the OUTSIDE walk leaves updateSdf's "no particle" value, sqrt(FLT_MAX) -
//    radius (the unbounded walk: about +layer). tests/test_nbflip_band_gpu.py steps both variants side by side:
//    particles, materials, velocities and pressure stay bit-identical, and the level set is identical inside the band.
//    fs2d_set_sdf_band(h, 0) (env FS2D_SDF_BAND=0) selects the unbounded walks; a download of FLUID_SDF completes the
//    inside walk on a copy, so host readers see the reference's interior values either way. 4096^2 nbflip substep:
//    23.7 ms -> under 1 ms for the two walks, and only the band makes the walks local enough for row slabs
//    (band + 2 <= halo rows).
static int sdfBandOf(const Ctx *ctx) { return ctx->p.sim_type == FS2D_SIM_NBFLIP ? ctx->sdfBand : 0; }

int gridFlushSdf(Ctx *ctx)
{
    if (!ctx->sdfInsidePending) return FS2D_OK;
    if (ctx->slab.enabled && ctx->slab.world > 1)
    {
        // the walk has unbounded radius: it needs every rank's rows, which only a collective call can bring here
        ctx->lastError = "the fluid level set below the surface is computed lazily and needs all slabs: call "
                         "fs2d_slab_gather_grid(FS2D_GRID_FLUID_SDF) on every rank first";
        return FS2D_ERR_STATE;
    }
    ctx->sdfInsidePending = false;
    return gridExtrapolateSdfNow(ctx, true, ctx->fluidSdf, 0, SlabRows{0, ctx->I});
}

// After fs2d_slab_gather_grid(FS2D_GRID_FLUID_SDF): every rank holds all rows and runs the walk on the whole grid
// (deferred for water; for NBFlip the unbounded completion of the banded interior).
int gridFlushSdfGathered(Ctx *ctx)
{
    if (!ctx->sdfInsidePending && !(ctx->p.sim_type == FS2D_SIM_NBFLIP && sdfBandOf(ctx) > 0)) return FS2D_OK;
    ctx->sdfInsidePending = false;
    return gridExtrapolateSdfNow(ctx, true, ctx->fluidSdf, 0, SlabRows{0, ctx->I});
}

// The level set a host reader should see: for NBFlip in band mode a COPY with the interior completed by the unbounded
// walk (the solver's own field keeps its band values, so stepping does not depend on who looked at the grid).
int gridSdfForRead(Ctx *ctx, const float **field)
{
    *field = ctx->fluidSdf;
    if (ctx->p.sim_type != FS2D_SIM_NBFLIP || sdfBandOf(ctx) <= 0) return gridFlushSdf(ctx);
    if (ctx->slab.enabled && ctx->slab.world > 1) return FS2D_OK;  // slabs: band values unless fs2d_slab_gather_grid ran just before
    FS2D_CUDA(cudaMemcpyAsync(ctx->scratchA, ctx->fluidSdf, sizeof(float) * ctx->N, cudaMemcpyDeviceToDevice, ctx->stream));
    FS2D_TRY(gridExtrapolateSdfNow(ctx, true, ctx->scratchA, 0, SlabRows{0, ctx->I}));
    *field = ctx->scratchA;
    return FS2D_OK;
}

int gridExtrapolateSdf(Ctx *ctx, bool inside)
{
    if (inside && ctx->p.sim_type != FS2D_SIM_NBFLIP && !ctx->eagerSdf)
    {
        ctx->sdfInsidePending = true;
        return FS2D_OK;
    }
    const int band = sdfBandOf(ctx);
    if (ctx->slab.enabled && ctx->slab.world > 1)
    {
        if (band <= 0 || band + 2 > ctx->slab.halo)
        {
            ctx->lastError = "extrapolateLevelset over row slabs needs the banded walk (0 < fs2d_set_sdf_band <= halo rows - 2)";
            return FS2D_ERR_STATE;
        }
        // the walk covers the owned rows plus the halo; what the cut at the outer halo row gets wrong moves one row per
        // layer, so after `band` layers the owned rows (and halo - band rows around them) are exact
        const void *arr[1] = {ctx->fluidSdf};
        const size_t rb[1] = {sizeof(float) * ctx->J};
        FS2D_TRY(slabExchangeFields(ctx, arr, rb, 1));
        return gridExtrapolateSdfNow(ctx, inside, ctx->fluidSdf, band, slabExt(ctx, ctx->slab.halo));
    }
    FS2D_TRY(gridFlushSdf(ctx));
    return gridExtrapolateSdfNow(ctx, inside, ctx->fluidSdf, band, SlabRows{0, ctx->I});
}

int gridExtrapolateSdfNow(Ctx *ctx, bool inside, float *field, int band, SlabRows rows)
{
    cudaStream_t st = ctx->stream;
    if (!ctx->bfsQueue)
    {
        FS2D_CUDA(cudaMalloc(reinterpret_cast<void **>(&ctx->bfsQueue), sizeof(int32_t) * static_cast<size_t>(ctx->N)));
        FS2D_CUDA(cudaMalloc(reinterpret_cast<void **>(&ctx->bfsCtl), 64));
    }
    FS2D_CUDA(cudaMemsetAsync(ctx->bfsCtl, 0, 64, st));
    const float maxSdf = static_cast<float>(static_cast<size_t>(ctx->I) * static_cast<size_t>(ctx->J));
    const CellRange cr = cellRange(ctx, rows);
    sdfMarkKernel<<<divUp(cr.count(), NT), NT, 0, st>>>(field, ctx->markers, cr.begin, cr.end, inside ? 1 : 0, maxSdf);
    sdfFirstLayerKernel<<<divUp(cr.count(), NT), NT, 0, st>>>(ctx->markers, ctx->J, rows.lo, rows.hi, ctx->bfsQueue, ctx->bfsCtl);
    {
        static const int perSm = std::getenv("FS2D_BFS_CTAS_PER_SM") ? std::atoi(std::getenv("FS2D_BFS_CTAS_PER_SM")) : 1;
        const int share = (ctx->slab.enabled && ctx->slab.world > 1) ? std::max(ctx->slab.share, 1) : 1;
        float *sdf = field;
        int32_t *markers = ctx->markers;
        int rowLo = rows.lo, rowHi = rows.hi, J = ctx->J;
        float step = inside ? -1.f : 1.f;
        int32_t *queue = ctx->bfsQueue;
        unsigned int *ctl = ctx->bfsCtl;
        int maxLayers = band;
        void *args[] = {&sdf, &markers, &rowLo, &rowHi, &J, &step, &queue, &ctl, &maxLayers};
        FS2D_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(sdfExtrapolateKernel),
                                              dim3(std::max(1, ctx->smCount * std::max(perSm, 1) / share)), dim3(BFS_THREADS), args, 0, st));
    }
    ctx->launches += 3;
    if (inside && band > 0)
    {
        sdfClampKernel<<<divUp(cr.count(), NT), NT, 0, st>>>(field, ctx->markers, cr.begin, cr.end, band, -static_cast<float>(band + 1));
        ctx->launches++;
    }
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int gridSaveVelocity(Ctx *ctx)
{
    const SlabRows reg = slabExt(ctx, ctx->slab.halo);
    const int rowHiU = reg.hi == ctx->I ? ctx->I + 1 : reg.hi;
    const size_t uOff = static_cast<size_t>(reg.lo) * ctx->J, vOff = static_cast<size_t>(reg.lo) * (ctx->J + 1);
    FS2D_CUDA(cudaMemcpyAsync(ctx->savedU + uOff, ctx->U + uOff, sizeof(float) * static_cast<size_t>(rowHiU - reg.lo) * ctx->J,
                              cudaMemcpyDeviceToDevice, ctx->stream));
    FS2D_CUDA(cudaMemcpyAsync(ctx->savedV + vOff, ctx->V + vOff, sizeof(float) * static_cast<size_t>(reg.hi - reg.lo) * (ctx->J + 1),
                              cudaMemcpyDeviceToDevice, ctx->stream));
    return FS2D_OK;
}

int gridBodyForces(Ctx *ctx)
{
    // const float factor = m_stepDt / m_dx (float / double, narrowed)
    const float factor = static_cast<float>(static_cast<double>(ctx->stepDt) / ctx->p.dx);
    if (isSmoke(ctx))
    {
        // Row slabs: the buoyancy samples temperature and soot one row around a velocity sample. The owned rows of both
        // grids were rewritten since their halo copies were last refreshed (P2G, or the semi-Lagrangian step in grid
        // mode), so the halo rows travel first; the force then goes to every velocity sample this rank holds a valid copy
        // of and has valid inputs for (owned rows + halo - 1).
        FS2D_TRY(smokeHalo(ctx));
        const bool slab = ctx->slab.enabled && ctx->slab.world > 1;
        const SlabRows reg = slabExt(ctx, slab ? ctx->slab.halo - 1 : 0);
        const int rowHiU = reg.hi == ctx->I ? ctx->I + 1 : reg.hi;
        const long long uBegin = static_cast<long long>(reg.lo) * ctx->J, nu = static_cast<long long>(rowHiU - reg.lo) * ctx->J;
        const long long vBegin = static_cast<long long>(reg.lo) * (ctx->J + 1), nv = static_cast<long long>(reg.hi - reg.lo) * (ctx->J + 1);
        smokeBodyForceKernel<<<divUp(nu + nv, NT), NT, 0, ctx->stream>>>(ctx->U, ctx->V, ctx->J, uBegin, nu, vBegin, nv, temperatureView(ctx),
                                                                        concentrationView(ctx), ctx->p.soot_factor,
                                                                        ctx->p.buoyancy_factor / ctx->p.ambient_temperature,
                                                                        ctx->p.ambient_temperature, ctx->p.gravity_x, ctx->p.gravity_y, factor);
    }
    else
    {
        // slab mode: every sample a rank holds a valid copy of (owned rows + exchanged halo) gets the force once
        const SlabRows reg = slabExt(ctx, ctx->slab.halo);
        const int rowHiU = reg.hi == ctx->I ? ctx->I + 1 : reg.hi;
        const long long uBegin = static_cast<long long>(reg.lo) * ctx->J, nu = static_cast<long long>(rowHiU - reg.lo) * ctx->J;
        const long long vBegin = static_cast<long long>(reg.lo) * (ctx->J + 1), nv = static_cast<long long>(reg.hi - reg.lo) * (ctx->J + 1);
        bodyForceKernel<<<divUp(nu + nv, NT), NT, 0, ctx->stream>>>(ctx->U, ctx->V, uBegin, nu, vBegin, nv, factor * ctx->p.gravity_x,
                                                                   factor * ctx->p.gravity_y);
    }
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int gridPressureRhs(Ctx *ctx)
{
    const double scale = 1.f / ctx->p.dx;  // const double scale = 1.f/m_dx
    const CellRange cr = cellRange(ctx, slabExt(ctx, 1));  // + one halo row: r0 = rhs there feeds the slab PCG
    pressureRhsKernel<<<divUp(cr.count(), NT), NT, 0, ctx->stream>>>(ctx->U, ctx->V, ctx->divergenceControl, ctx->material, ctx->I, ctx->J,
                                                                    isSmoke(ctx) ? 1 : 0, scale, ctx->rhs, cr.begin, cr.end);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int gridDensityRhs(Ctx *ctx)
{
    const double scale = 1.0 / ctx->stepDt;
    const CellRange cr = cellRange(ctx, slabExt(ctx, 1));
    densityRhsKernel<<<divUp(cr.count(), NT), NT, 0, ctx->stream>>>(ctx->density, ctx->material, cr.begin, cr.end, scale, ctx->p.fluid_density,
                                                                   ctx->rhs);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int gridApplyPressure(Ctx *ctx)
{
    const double scale = ctx->stepDt / (ctx->p.fluid_density * ctx->p.dx);
    const CellRange cr = cellRange(ctx, slabOwn(ctx));
    applyPressureKernel<<<divUp(cr.count(), NT), NT, 0, ctx->stream>>>(ctx->x, ctx->material, ctx->I, ctx->J, isSmoke(ctx) ? 1 : 0, scale,
                                                                      ctx->U, ctx->V, ctx->uValid, ctx->vValid,
                                                                      isSmoke(ctx) ? ctx->testGrid : nullptr, cr.begin, cr.end);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int gridVelocityFromSolids(Ctx *ctx)
{
    if (ctx->numObstacles == 0) return FS2D_OK;
    // every coefficient 0 (the default of a solid in the scene file): the average is 0 / count = 0 and every affected sample
    // is multiplied by 1 - 0 = 1, which leaves every float as it is -- the pass (9 look-ups per cell over the whole grid) is
    // skipped
    if (ctx->obstaclesFrictionless) return FS2D_OK;
    const CellRange cr = cellRange(ctx, slabOwn(ctx));
    solidFrictionKernel<<<divUp(cr.count(), NT), NT, 0, ctx->stream>>>(ctx->solidId, ctx->obstacleFriction, ctx->material, ctx->I, ctx->J,
                                                                      ctx->U, ctx->V, cr.begin, cr.end);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

// FlipSmokeSolver / FlipFireSolver::eulerAdvectParameters (flipsmokesolver.cpp:211-233, flipfiresolver.cpp:155-172)
int gridEulerAdvectParameters(Ctx *ctx)
{
    if (!isSmoke(ctx)) return FS2D_OK;  // FlipSolver::eulerAdvectParameters is empty for water
    // Row slabs: the samples of the owned rows; a back-trace reaches cflNumber + 1 rows beyond them, inside the halo (U and V
    // were refreshed by the extrapolation that ended the previous substep, the advected grids travel here). Every rank
    // swaps the same arrays, so a grid keeps one offset in the symmetric heap on all ranks.
    FS2D_TRY(smokeHalo(ctx));
    const VelocityView vel = makeVelocityView(ctx->U, ctx->V, ctx->I, ctx->J);
    const CellRange cr = cellRange(ctx, slabOwn(ctx));
    const int blocks = divUp(cr.count(), NT);
    const bool fire = ctx->p.sim_type == FS2D_SIM_FIRE;
    smokeAdvectKernel<<<blocks, NT, 0, ctx->stream>>>(concentrationView(ctx), temperatureView(ctx), fire ? fuelView(ctx) : concentrationView(ctx),
                                                      fire ? 1 : 0, vel, ctx->stepDt, ctx->scratchA, ctx->scratchB, ctx->scratchC, cr.begin, cr.end);
    ctx->launches++;
    if (fire) std::swap(ctx->fuel, ctx->scratchC);
    std::swap(ctx->concentration, ctx->scratchA);
    std::swap(ctx->temperature, ctx->scratchB);
    ctx->smokeGridsAdvected = true;  // the members now carry OOB_EXTEND and offset (1/2,1/2)
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

// Slab mode: halo rows of the level set and of the viscosity grid before pruneNarrowBand and the semi-Lagrangian step.
int gridNbflipHalo(Ctx *ctx)
{
    if (ctx->p.sim_type != FS2D_SIM_NBFLIP) return FS2D_OK;
    const void *arr[2] = {ctx->fluidSdf, ctx->viscosity};
    const size_t rb[2] = {sizeof(float) * ctx->J, sizeof(float) * ctx->J};
    return slabExchangeFields(ctx, arr, rb, 2);
}

// NBFlipSolver::advect, the grid part (nbflipsolver.cpp:66-109)
int gridNbflipAdvect(Ctx *ctx)
{
    if (ctx->p.sim_type != FS2D_SIM_NBFLIP) return FS2D_OK;
    // Slab mode: the samples of the owned rows; a back-trace reaches cflNumber + 1 rows beyond them, inside the halo
    // (U and V: refreshed by the extrapolation that ended the previous substep; level set and viscosity: by
    // gridNbflipHalo, which fs2d_nbflip_advect_grids calls first).
    const SlabRows own = slabOwn(ctx);
    const int rowHiU = own.hi == ctx->I ? ctx->I + 1 : own.hi;
    const long long J = ctx->J;
    const VelocityView vel = makeVelocityView(ctx->U, ctx->V, ctx->I, ctx->J);
    cudaStream_t st = ctx->stream;
    nbflipAdvectKernel<<<divUp((rowHiU - own.lo) * (J + 1), NT), NT, 0, st>>>(vel, fluidSdfView(ctx), viscosityView(ctx), ctx->stepDt, ctx->J, own.lo,
                                                                             rowHiU, own.hi, ctx->advU, ctx->advV, ctx->advSdf, ctx->advViscosity);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}
