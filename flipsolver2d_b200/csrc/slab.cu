// Row-slab decomposition over several GPUs: configuration, peer mapping and the
// exchange primitives every slab-aware stage uses.
//
// One handle per GPU owns the cell rows [rowBegin, rowEnd) (a multiple of the 16-row PCG tile). Dense
// arrays keep their GLOBAL size and indexing on every rank -- at 8192^2 all grids and PCG vectors are
// ~11 GB of the 180 GB -- so a kernel only needs a row range, and a halo row lives at the same offset
// on the neighbour. Ranks never call a library for the exchange: they PUSH rows (or packed particle
// records) into the neighbour's memory through peer-mapped pointers (CUDA IPC between processes, plain
// pointers inside one process) and raise a sequence number the neighbour spins on in its OWN memory.
// Measured on 2 x B200 (profiles/r1e_ipc_probe_2gpu.jsonl): 1.2 us one way, against 18 us for a 16-byte
// NCCL all-reduce -- the PCG needs two such reductions per iteration (pcg.cu fuses them into its kernels).
//
// Exchange protocol (one kernel launch per exchange, both neighbours at once):
//   1. tell both neighbours "ready #seq" -- everything I launched before has finished reading my halo;
//   2. wait for their "ready #seq", then copy my boundary rows into their arrays;
//   3. system fence, tell them "data #seq", wait for their "data #seq".
// Every rank issues the same sequence of exchanges, so #seq needs no negotiation.
#include <unistd.h>

#include <algorithm>
#include <cstring>

#include "fs2d_internal.h"

namespace
{
constexpr long long SPIN_LIMIT = 8000000000ll;  // cycles (~4 s): a lost peer turns into an error, not a hang

struct ExportBlob
{
    cudaIpcMemHandle_t heap;
    cudaIpcMemHandle_t xchg;
    unsigned long long pid;
    unsigned long long heapPtr;
    unsigned long long xchgPtr;
    unsigned long long heapBytes;
    unsigned long long xchgBytes;
    int device;
    int rank;
};
static_assert(sizeof(ExportBlob) <= FS2D_SLAB_HANDLE_BYTES, "export blob does not fit the ABI buffer");

__device__ __forceinline__ bool spinAtLeast(const unsigned long long *flag, unsigned long long want, int *err, int site = 9)
{
    const volatile unsigned long long *f = flag;
    const long long t0 = clock64();
    while (*f < want)
    {
        if (clock64() - t0 > SPIN_LIMIT)
        {
            *err = site;
            return false;
        }
    }
    return true;
}

__device__ __forceinline__ void storeFlag(unsigned long long *flag, unsigned long long v)
{
    *reinterpret_cast<volatile unsigned long long *>(flag) = v;
}

__device__ __forceinline__ void copyBytes16(unsigned char *dst, const unsigned char *src, unsigned long long bytes)
{
    // both ends are 16-byte aligned (row ranges start at multiples of 16 rows)
    const unsigned long long n16 = bytes >> 4;
    const uint4 *s = reinterpret_cast<const uint4 *>(src);
    uint4 *d = reinterpret_cast<uint4 *>(dst);
    for (unsigned long long k = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; k < n16;
         k += static_cast<unsigned long long>(gridDim.x) * blockDim.x)
        d[k] = s[k];
    if (blockIdx.x == 0)
        for (unsigned long long k = (n16 << 4) + threadIdx.x; k < bytes; k += blockDim.x) dst[k] = src[k];
}

struct HaloArgs
{
    int nFields;
    SlabField f[8];
    unsigned char *heap, *loHeap, *hiHeap;
    SlabMail *mail, *loMail, *hiMail;
    unsigned long long seq;
    long long rowBegin, rowEnd, halo;
    unsigned int *ticket;
};

__device__ void exchangeOpen(SlabMail *mail, SlabMail *loMail, SlabMail *hiMail, unsigned long long seq)
{
    if (threadIdx.x == 0)
    {
        if (blockIdx.x == 0)
        {
            if (loMail) storeFlag(&loMail->ready[1], seq);
            if (hiMail) storeFlag(&hiMail->ready[0], seq);
        }
        if (loMail) spinAtLeast(&mail->ready[0], seq, &mail->error, 5);
        if (hiMail) spinAtLeast(&mail->ready[1], seq, &mail->error, 5);
    }
    __syncthreads();
}

__device__ void exchangeClose(SlabMail *mail, SlabMail *loMail, SlabMail *hiMail, unsigned long long seq, unsigned int *ticket)
{
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        if (atomicAdd(ticket, 1u) == gridDim.x - 1)
        {
            *ticket = 0;
            __threadfence_system();
            if (loMail) storeFlag(&loMail->data[1], seq);
            if (hiMail) storeFlag(&hiMail->data[0], seq);
            if (loMail) spinAtLeast(&mail->data[0], seq, &mail->error, 6);
            if (hiMail) spinAtLeast(&mail->data[1], seq, &mail->error, 6);
            __threadfence_system();
        }
    }
}

__global__ void __launch_bounds__(256) slabHaloKernel(HaloArgs a)
{
    exchangeOpen(a.mail, a.loMail, a.hiMail, a.seq);
    const long long loEnd = a.rowBegin + a.halo < a.rowEnd ? a.rowBegin + a.halo : a.rowEnd;
    const long long hiBegin = a.rowEnd - a.halo > a.rowBegin ? a.rowEnd - a.halo : a.rowBegin;
    for (int k = 0; k < a.nFields; k++)
    {
        const unsigned long long rb = a.f[k].rowBytes, off = a.f[k].offset;
        if (a.loHeap)
            copyBytes16(a.loHeap + off + a.rowBegin * rb, a.heap + off + a.rowBegin * rb, (loEnd - a.rowBegin) * rb);
        if (a.hiHeap) copyBytes16(a.hiHeap + off + hiBegin * rb, a.heap + off + hiBegin * rb, (a.rowEnd - hiBegin) * rb);
    }
    exchangeClose(a.mail, a.loMail, a.hiMail, a.seq, a.ticket);
}

// ---------------------------------------------------------------- particle records
// Exchange buffers (one allocation, mapped by the neighbours): send[lo], send[hi], recv[from lo], recv[from hi];
// each holds `cap` records as SoA: pos (8 B), vel (8 B), K property columns (4 B), storage code (1 B).
struct RecordView
{
    float2 *pos, *vel;
    float *props;
    uint8_t *mis;
};

__host__ __device__ inline size_t recordBufferBytes(long long cap, int K)
{
    size_t b = static_cast<size_t>(cap) * (16 + 4 * static_cast<size_t>(K) + 1);
    return (b + 255) & ~static_cast<size_t>(255);
}

__host__ __device__ inline RecordView recordView(unsigned char *base, int which, long long cap, int K)
{
    unsigned char *b = base + which * recordBufferBytes(cap, K);
    RecordView v;
    v.pos = reinterpret_cast<float2 *>(b);
    v.vel = reinterpret_cast<float2 *>(b + 8 * cap);
    v.props = reinterpret_cast<float *>(b + 16 * cap);
    v.mis = b + (16 + 4 * static_cast<size_t>(K)) * cap;
    return v;
}

// Owned particles (cell key inside my rows at the last sort) are routed by the row their position is in now:
// still mine -> stays, copied to a neighbour as a ghost when within `ghost` rows of that boundary; left my rows
// -> record goes to the neighbour, and it stays here as a ghost while within `ghost` rows. Ghosts of the previous
// exchange are dropped.
__global__ void __launch_bounds__(256) slabClassifyKernel(const float2 *__restrict__ pos, const float2 *__restrict__ vel,
                                                          const float *__restrict__ props, long long propStride,
                                                          const uint32_t *__restrict__ key, const uint8_t *__restrict__ mis,
                                                          uint8_t *__restrict__ dead, long long count, int I, int J, int K,
                                                          int rowBegin, int rowEnd, int ghost, int hasLo, int hasHi,
                                                          RecordView sendLo, RecordView sendHi, long long cap,
                                                          unsigned long long *__restrict__ counters /* [0] lo, [1] hi, [2] overflow */)
{
    const long long p = blockIdx.x * 256ll + threadIdx.x;
    if (p >= count || dead[p]) return;
    const int keyRow = static_cast<int>(key[p] / static_cast<uint32_t>(J));
    if (keyRow < rowBegin || keyRow >= rowEnd)
    {
        dead[p] = 1;  // ghost of the previous exchange
        return;
    }
    const float2 x = pos[p];
    int i = static_cast<int>(floorf(x.x));
    i = i < 0 ? 0 : (i > I - 1 ? I - 1 : i);
    bool toLo, toHi, drop = false;
    if (i < rowBegin)
    {
        toLo = true;
        toHi = false;
        drop = i < rowBegin - ghost;
    }
    else if (i >= rowEnd)
    {
        toLo = false;
        toHi = true;
        drop = i >= rowEnd + ghost;
    }
    else
    {
        toLo = hasLo && i < rowBegin + ghost;
        toHi = hasHi && i >= rowEnd - ghost;
    }
#pragma unroll
    for (int side = 0; side < 2; side++)
    {
        if (!(side == 0 ? toLo : toHi)) continue;
        const unsigned long long slot = atomicAdd(counters + side, 1ull);
        if (slot >= static_cast<unsigned long long>(cap))
        {
            counters[2] = 1;
            continue;
        }
        const RecordView &s = side == 0 ? sendLo : sendHi;
        s.pos[slot] = x;
        s.vel[slot] = vel[p];
        s.mis[slot] = mis[p];
        for (int k = 0; k < K; k++) s.props[k * cap + slot] = props[k * propStride + p];
    }
    if (drop) dead[p] = 1;
}

struct ParticleXArgs
{
    unsigned char *xchg, *loXchg, *hiXchg;
    SlabMail *mail, *loMail, *hiMail;
    unsigned long long seq;
    long long cap;
    int K;
    const unsigned long long *counters;
    unsigned int *ticket;
};

__device__ void pushRecords(const RecordView &src, const RecordView &dst, long long n, long long cap, int K)
{
    const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
    for (long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; k < n; k += stride)
    {
        dst.pos[k] = src.pos[k];
        dst.vel[k] = src.vel[k];
        dst.mis[k] = src.mis[k];
        for (int q = 0; q < K; q++) dst.props[q * cap + k] = src.props[q * cap + k];
    }
}

__global__ void __launch_bounds__(256) slabParticleKernel(ParticleXArgs a)
{
    exchangeOpen(a.mail, a.loMail, a.hiMail, a.seq);
    long long nLo = static_cast<long long>(a.counters[0]), nHi = static_cast<long long>(a.counters[1]);
    if (nLo > a.cap) nLo = a.cap;
    if (nHi > a.cap) nHi = a.cap;
    if (a.loXchg)
    {
        // my send[lo] -> the lower neighbour's recv[from hi]
        pushRecords(recordView(a.xchg, 0, a.cap, a.K), recordView(a.loXchg, 3, a.cap, a.K), nLo, a.cap, a.K);
        if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile long long *>(&a.loMail->counts[1][0]) = nLo;
    }
    if (a.hiXchg)
    {
        pushRecords(recordView(a.xchg, 1, a.cap, a.K), recordView(a.hiXchg, 2, a.cap, a.K), nHi, a.cap, a.K);
        if (blockIdx.x == 0 && threadIdx.x == 0) *reinterpret_cast<volatile long long *>(&a.hiMail->counts[0][0]) = nHi;
    }
    exchangeClose(a.mail, a.loMail, a.hiMail, a.seq, a.ticket);
}

__global__ void __launch_bounds__(256) slabAppendKernel(RecordView src, long long n, long long cap, int K, int I, int J,
                                                        float2 *__restrict__ pos, float2 *__restrict__ vel,
                                                        float *__restrict__ props, long long propStride,
                                                        uint32_t *__restrict__ key, uint8_t *__restrict__ mis,
                                                        uint8_t *__restrict__ dead, long long base)
{
    const long long k = blockIdx.x * 256ll + threadIdx.x;
    if (k >= n) return;
    const float2 x = src.pos[k];
    pos[base + k] = x;
    vel[base + k] = src.vel[k];
    mis[base + k] = src.mis[k];
    for (int q = 0; q < K; q++) props[q * propStride + base + k] = src.props[q * cap + k];
    dead[base + k] = 0;
    int i = static_cast<int>(floorf(x.x)), j = static_cast<int>(floorf(x.y));
    i = i < 0 ? 0 : (i > I - 1 ? I - 1 : i);
    j = j < 0 ? 0 : (j > J - 1 ? J - 1 : j);
    key[base + k] = static_cast<uint32_t>(i) * static_cast<uint32_t>(J) + static_cast<uint32_t>(j);
}

// ---------------------------------------------------------------- small all-gather
struct GatherArgs
{
    SlabMail *mail;
    SlabMail *peerMail[FS2D_MAX_RANKS];
    int rank, world;
    unsigned long long seq;
    long long v[4];
};

__global__ void slabGatherKernel(GatherArgs a)
{
    const int r = threadIdx.x;
    const int par = static_cast<int>(a.seq & 1ull);
    if (r < a.world)
    {
        SlabGatherSlot *slot = &a.peerMail[r]->gather[par][a.rank];
        for (int k = 0; k < 4; k++) *reinterpret_cast<volatile long long *>(&slot->v[k]) = a.v[k];
        __threadfence_system();
        storeFlag(&slot->tag, a.seq);
        spinAtLeast(&a.mail->gather[par][r].tag, a.seq, &a.mail->error, 7);
        __threadfence_system();
    }
}

// ---------------------------------------------------------------- all-gather of the rows of one dense array
struct GatherRowsArgs
{
    unsigned char *heap;
    unsigned char *peerHeap[FS2D_MAX_RANKS];
    int rank, world;
    unsigned long long offset, bytesBegin, bytes;
};

__global__ void __launch_bounds__(256) slabGatherRowsKernel(GatherRowsArgs a)
{
    for (int r = 0; r < a.world; r++)
        if (r != a.rank) copyBytes16(a.peerHeap[r] + a.offset + a.bytesBegin, a.heap + a.offset + a.bytesBegin, a.bytes);
    __threadfence_system();
}

SlabMail *peerMailOf(Ctx *ctx, int r)
{
    if (r < 0 || r >= ctx->slab.world) return nullptr;
    return reinterpret_cast<SlabMail *>(ctx->slab.peerHeap[r] + (reinterpret_cast<unsigned char *>(ctx->mail) - ctx->heap));
}

unsigned int *ticketOf(Ctx *ctx) { return reinterpret_cast<unsigned int *>(ctx->d_counter + 13); }

int requireConnected(Ctx *ctx)
{
    if (!ctx->slab.enabled) return FS2D_OK;
    if (ctx->slab.connected != ctx->slab.world - 1)
    {
        ctx->lastError = "slab: not all peers are connected (fs2d_slab_connect)";
        return FS2D_ERR_COMM;
    }
    return FS2D_OK;
}
}  // namespace

SlabRows slabOwn(const Ctx *ctx)
{
    if (!ctx->slab.enabled) return {0, ctx->I};
    return {ctx->slab.rowBegin, ctx->slab.rowEnd};
}

SlabRows slabExt(const Ctx *ctx, int k)
{
    if (!ctx->slab.enabled) return {0, ctx->I};
    return {std::max(ctx->slab.rowBegin - k, 0), std::min(ctx->slab.rowEnd + k, ctx->I)};
}

void slabRelease(Ctx *ctx)
{
    SlabState &s = ctx->slab;
    for (int r = 0; r < FS2D_MAX_RANKS; r++)
    {
        if (s.peerMapped[r])
        {
            if (s.peerHeap[r]) cudaIpcCloseMemHandle(s.peerHeap[r]);
            if (s.peerXchg[r]) cudaIpcCloseMemHandle(s.peerXchg[r]);
        }
        s.peerHeap[r] = s.peerXchg[r] = nullptr;
        s.peerMapped[r] = false;
    }
    if (s.xchg) cudaFree(s.xchg);
    s.xchg = nullptr;
    s.enabled = false;
}

// Sticky: once a peer was given up on, this rank's halos are stale and every later check fails too.
int slabCheckError(Ctx *ctx)
{
    if (!ctx->slab.enabled || ctx->slab.world == 1) return FS2D_OK;
    int err = 0;
    FS2D_CUDA(fs2dCopyToHost(ctx, &err, &ctx->mail->error, sizeof(err)));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    if (err)
    {
        // which wait gave up: 1 PCG phase-0 partials, 2 PCG barrier words, 3 solve mode of a row neighbour, 4 LL halo row,
        // 5 / 6 halo exchange open / close, 7 all-gather, 9 other
        ctx->lastError = "slab: a peer did not answer within the spin limit (wait site " + std::to_string(err) + ")";
        return FS2D_ERR_COMM;
    }
    return FS2D_OK;
}

int slabExchangeFields(Ctx *ctx, const void *const *arrays, const size_t *rowBytes, int count)
{
    SlabState &s = ctx->slab;
    if (!s.enabled || s.world == 1) return FS2D_OK;
    FS2D_TRY(requireConnected(ctx));
    if (count > 8) return FS2D_ERR_ARG;
    HaloArgs a;
    a.nFields = count;
    for (int k = 0; k < count; k++)
    {
        a.f[k].offset = static_cast<unsigned long long>(static_cast<const unsigned char *>(arrays[k]) - ctx->heap);
        a.f[k].rowBytes = rowBytes[k];
    }
    a.heap = ctx->heap;
    a.loHeap = s.rank > 0 ? s.peerHeap[s.rank - 1] : nullptr;
    a.hiHeap = s.rank + 1 < s.world ? s.peerHeap[s.rank + 1] : nullptr;
    a.mail = ctx->mail;
    a.loMail = peerMailOf(ctx, s.rank - 1);
    a.hiMail = peerMailOf(ctx, s.rank + 1);
    a.seq = ++s.seq;
    a.rowBegin = s.rowBegin;
    a.rowEnd = s.rowEnd;
    a.halo = s.halo;
    a.ticket = ticketOf(ctx);
    const int blocks = std::max(1, std::min(64, ctx->smCount / (2 * s.share)));
    slabHaloKernel<<<blocks, 256, 0, ctx->stream>>>(a);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int slabExchangeVelocity(Ctx *ctx, bool withMaterial)
{
    const void *arr[5] = {ctx->U, ctx->V, ctx->uValid, ctx->vValid, ctx->material};
    const size_t rb[5] = {sizeof(float) * ctx->J, sizeof(float) * (ctx->J + 1), static_cast<size_t>(ctx->J),
                          static_cast<size_t>(ctx->J + 1), static_cast<size_t>(ctx->J)};
    return slabExchangeFields(ctx, arr, rb, withMaterial ? 5 : 4);
}

int slabExchangePressure(Ctx *ctx)
{
    const void *arr[1] = {ctx->x};
    const size_t rb[1] = {sizeof(double) * ctx->J};
    return slabExchangeFields(ctx, arr, rb, 1);
}

int particlesReserve(Ctx *ctx, int64_t capacity);

// Migrants and ghosts: classify + pack, push to the neighbours, append what they pushed here. The caller sorts.
int slabExchangeParticles(Ctx *ctx)
{
    SlabState &s = ctx->slab;
    if (!s.enabled || s.world == 1) return FS2D_OK;
    FS2D_TRY(requireConnected(ctx));
    cudaStream_t st = ctx->stream;
    const int K = ctx->p.num_properties;
    unsigned long long *counters = reinterpret_cast<unsigned long long *>(ctx->d_counter + 10);
    FS2D_CUDA(cudaMemsetAsync(counters, 0, 3 * sizeof(unsigned long long), st));
    ParticleBuffers &b = ctx->pb[ctx->cur];
    if (ctx->count > 0)
    {
        slabClassifyKernel<<<divUp(ctx->count, 256), 256, 0, st>>>(
            b.pos, b.vel, b.props, b.capacity, b.key, b.mis, ctx->dead, ctx->count, ctx->I, ctx->J, K, s.rowBegin, s.rowEnd, s.ghost,
            s.rank > 0 ? 1 : 0, s.rank + 1 < s.world ? 1 : 0, recordView(s.xchg, 0, s.xchgCapacity, K),
            recordView(s.xchg, 1, s.xchgCapacity, K), s.xchgCapacity, counters);
        ctx->launches++;
    }
    long long recvLo = 0, recvHi = 0;
    if (s.world > 1)
    {
        ParticleXArgs a;
        a.xchg = s.xchg;
        a.loXchg = s.rank > 0 ? s.peerXchg[s.rank - 1] : nullptr;
        a.hiXchg = s.rank + 1 < s.world ? s.peerXchg[s.rank + 1] : nullptr;
        a.mail = ctx->mail;
        a.loMail = peerMailOf(ctx, s.rank - 1);
        a.hiMail = peerMailOf(ctx, s.rank + 1);
        a.seq = ++s.seq;
        a.cap = s.xchgCapacity;
        a.K = K;
        a.counters = counters;
        a.ticket = ticketOf(ctx);
        const int blocks = std::max(1, std::min(64, ctx->smCount / (2 * s.share)));
        slabParticleKernel<<<blocks, 256, 0, st>>>(a);
        ctx->launches++;
        FS2D_CUDA(cudaGetLastError());
        long long counts[2][4];
        unsigned long long overflow = 0;
        FS2D_CUDA(fs2dCopyToHost(ctx, counts, ctx->mail->counts, sizeof(counts)));
        FS2D_CUDA(fs2dCopyToHost(ctx, &overflow, counters + 2, sizeof(overflow)));
        FS2D_CUDA(cudaStreamSynchronize(st));
        // the host is synchronised here anyway: a spin loop of any slab kernel since the last check that gave up on a
        // peer (halo exchange, particle exchange, PCG barrier) is reported now instead of stepping on with stale halos
        FS2D_TRY(slabCheckError(ctx));
        if (overflow)
        {
            ctx->lastError = "slab: particle exchange buffer overflow";
            return FS2D_ERR_STATE;
        }
        recvLo = s.rank > 0 ? counts[0][0] : 0;
        recvHi = s.rank + 1 < s.world ? counts[1][0] : 0;
    }
    const long long recv = recvLo + recvHi;
    if (recv > 0)
    {
        FS2D_TRY(particlesReserve(ctx, ctx->count + recv));
        ParticleBuffers &nb = ctx->pb[ctx->cur];
        long long base = ctx->count;
        for (int side = 0; side < 2; side++)
        {
            const long long n = side == 0 ? recvLo : recvHi;
            if (n == 0) continue;
            slabAppendKernel<<<divUp(n, 256), 256, 0, st>>>(recordView(s.xchg, 2 + side, s.xchgCapacity, K), n, s.xchgCapacity, K, ctx->I,
                                                            ctx->J, nb.pos, nb.vel, nb.props, nb.capacity, nb.key, nb.mis, ctx->dead, base);
            ctx->launches++;
            base += n;
        }
        ctx->count += recv;
    }
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int slabAllGather(Ctx *ctx, const long long v[4], long long *out)
{
    SlabState &s = ctx->slab;
    if (!s.enabled || s.world == 1)
    {
        for (int k = 0; k < 4; k++) out[k] = v[k];
        return FS2D_OK;
    }
    FS2D_TRY(requireConnected(ctx));
    GatherArgs a;
    a.mail = ctx->mail;
    for (int r = 0; r < FS2D_MAX_RANKS; r++) a.peerMail[r] = r < s.world ? peerMailOf(ctx, r) : nullptr;
    a.rank = s.rank;
    a.world = s.world;
    a.seq = ++s.gatherSeq;
    for (int k = 0; k < 4; k++) a.v[k] = v[k];
    slabGatherKernel<<<1, 32, 0, ctx->stream>>>(a);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    SlabGatherSlot slots[FS2D_MAX_RANKS];
    FS2D_CUDA(fs2dCopyToHost(ctx, slots, ctx->mail->gather[a.seq & 1ull], sizeof(slots)));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    FS2D_TRY(slabCheckError(ctx));
    for (int r = 0; r < s.world; r++)
    {
        if (slots[r].tag != a.seq)
        {
            ctx->lastError = "slab: all-gather timed out";
            return FS2D_ERR_COMM;
        }
        for (int k = 0; k < 4; k++) out[4 * r + k] = slots[r].v[k];
    }
    return FS2D_OK;
}

// Every rank pushes the rows it owns of `array` into every other rank's copy. Bracketed by two all-gathers (host
// synchronous): nobody is overwritten before it has finished what it was doing, nobody reads before all pushes landed.
int slabGatherRows(Ctx *ctx, void *array, size_t rowBytes, int rowsTotal)
{
    void *arrays[1] = {array};
    const size_t rb[1] = {rowBytes};
    const int rt[1] = {rowsTotal};
    return slabGatherMany(ctx, arrays, rb, rt, 1);
}

// Several arrays inside one pair of all-gathers.
int slabGatherMany(Ctx *ctx, void *const *arrays, const size_t *rowBytes, const int *rowsTotal, int count)
{
    SlabState &s = ctx->slab;
    if (!s.enabled || s.world == 1) return FS2D_OK;
    FS2D_TRY(requireConnected(ctx));
    const long long zero[4] = {0, 0, 0, 0};
    long long all[4 * FS2D_MAX_RANKS];
    FS2D_TRY(slabAllGather(ctx, zero, all));
    for (int k = 0; k < count; k++)
    {
        GatherRowsArgs a;
        a.heap = ctx->heap;
        for (int r = 0; r < FS2D_MAX_RANKS; r++) a.peerHeap[r] = r < s.world ? s.peerHeap[r] : nullptr;
        a.rank = s.rank;
        a.world = s.world;
        a.offset = static_cast<unsigned long long>(static_cast<unsigned char *>(arrays[k]) - ctx->heap);
        const int rowEnd = s.rank == s.world - 1 ? rowsTotal[k] : s.rowEnd;  // the extra U row belongs to the last slab
        a.bytesBegin = static_cast<unsigned long long>(s.rowBegin) * rowBytes[k];
        a.bytes = static_cast<unsigned long long>(rowEnd - s.rowBegin) * rowBytes[k];
        const int blocks = std::max(1, ctx->smCount / s.share);
        slabGatherRowsKernel<<<blocks, 256, 0, ctx->stream>>>(a);
        ctx->launches++;
    }
    FS2D_CUDA(cudaGetLastError());
    FS2D_TRY(slabAllGather(ctx, zero, all));
    return FS2D_OK;
}

int slabGatherBand(Ctx *ctx, void *const *arrays, const size_t *rowBytes, const int *rowsTotal, int count, int localMin, int localMax,
                   int margin, int *rowLo, int *rowHi)
{
    SlabState &s = ctx->slab;
    *rowLo = 0;
    *rowHi = ctx->I;
    if (!s.enabled || s.world == 1) return FS2D_OK;
    FS2D_TRY(requireConnected(ctx));
    // the opening all-gather (nobody is overwritten before it has finished what it was doing) carries the rows
    const long long mine[4] = {localMax >= localMin ? localMin : 0x7fffffff, localMax >= localMin ? localMax : -1, 0, 0};
    long long all[4 * FS2D_MAX_RANKS];
    FS2D_TRY(slabAllGather(ctx, mine, all));
    long long gmin = 0x7fffffff, gmax = -1;
    for (int r = 0; r < s.world; r++)
    {
        gmin = std::min(gmin, all[4 * r]);
        gmax = std::max(gmax, all[4 * r + 1]);
    }
    const long long zero[4] = {0, 0, 0, 0};
    if (gmax < gmin)
    {
        *rowLo = *rowHi = 0;
        return slabAllGather(ctx, zero, all);  // every rank takes this branch: keep the pair of handshakes
    }
    const int lo = static_cast<int>(std::max<long long>(0, gmin - margin)), hi = static_cast<int>(std::min<long long>(ctx->I, gmax + 1 + margin));
    *rowLo = lo;
    *rowHi = hi;
    for (int k = 0; k < count; k++)
    {
        // the extra U row belongs to the last slab and travels when the band reaches the last cell row
        const int ownEnd = s.rank == s.world - 1 ? rowsTotal[k] : s.rowEnd;
        const int bandEnd = hi == ctx->I ? rowsTotal[k] : hi;
        const int b = std::max(s.rowBegin, lo), e = std::min(ownEnd, bandEnd);
        if (e <= b) continue;
        GatherRowsArgs a;
        a.heap = ctx->heap;
        for (int r = 0; r < FS2D_MAX_RANKS; r++) a.peerHeap[r] = r < s.world ? s.peerHeap[r] : nullptr;
        a.rank = s.rank;
        a.world = s.world;
        a.offset = static_cast<unsigned long long>(static_cast<unsigned char *>(arrays[k]) - ctx->heap);
        a.bytesBegin = static_cast<unsigned long long>(b) * rowBytes[k];
        a.bytes = static_cast<unsigned long long>(e - b) * rowBytes[k];
        const int blocks = std::max(1, ctx->smCount / s.share);
        slabGatherRowsKernel<<<blocks, 256, 0, ctx->stream>>>(a);
        ctx->launches++;
    }
    FS2D_CUDA(cudaGetLastError());
    return slabAllGather(ctx, zero, all);
}

extern "C" {

int fs2d_slab_configure(fs2d_handle ctx, int rank, int world, int device_share)
{
    if (!ctx || world < 1 || world > FS2D_MAX_RANKS) return FS2D_ERR_ARG;
    // equal numbers of 16-row tile rows per rank
    int32_t bounds[FS2D_MAX_RANKS + 1];
    const int tileRows = divUp(ctx->I, 16);
    for (int r = 0; r <= world; r++) bounds[r] = std::min(ctx->I, 16 * static_cast<int>(static_cast<long long>(r) * tileRows / world));
    return fs2d_slab_configure_rows(ctx, rank, world, device_share, bounds);
}

int fs2d_slab_configure_rows(fs2d_handle ctx, int rank, int world, int device_share, const int32_t *row_bounds)
{
    if (!ctx || !row_bounds || world < 1 || world > FS2D_MAX_RANKS || rank < 0 || rank >= world || device_share < 1) return FS2D_ERR_ARG;
    SlabState &s = ctx->slab;
    if (s.enabled)
    {
        ctx->lastError = "fs2d_slab_configure: already configured";
        return FS2D_ERR_STATE;
    }
    if (ctx->J % 2 != 0 || ctx->p.convergence_threads > 0)
    {
        ctx->lastError = "fs2d_slab_configure: needs an even gridSizeJ and convergence_threads = 0 (true max-norm test)";
        return FS2D_ERR_ARG;
    }
    if (row_bounds[0] != 0 || row_bounds[world] != ctx->I)
    {
        ctx->lastError = "fs2d_slab_configure_rows: the boundaries must start at 0 and end at gridSizeI";
        return FS2D_ERR_ARG;
    }
    for (int r = 0; r < world; r++)
    {
        const bool aligned = r == 0 || row_bounds[r] % 16 == 0;
        if (!aligned || row_bounds[r + 1] <= row_bounds[r] || (world > 1 && row_bounds[r + 1] - row_bounds[r] < s.halo))
        {
            ctx->lastError = "fs2d_slab_configure: slab boundaries must be multiples of 16 and a slab must hold at least 32 rows";
            return FS2D_ERR_ARG;
        }
    }
    s.rank = rank;
    s.world = world;
    s.share = device_share;
    s.rowBegin = row_bounds[rank];
    s.rowEnd = row_bounds[rank + 1];
    // migrants (CFL <= 5 cells -> a few rows) + ghost rows, every cell at the 2*ppc cap
    s.xchgCapacity = std::max<int64_t>(static_cast<int64_t>(ctx->J) * (s.ghost + 8) * 2 * std::max(ctx->p.particles_per_cell, 1), 1 << 16);
    s.xchgBytes = 4 * recordBufferBytes(s.xchgCapacity, ctx->p.num_properties);
    FS2D_CUDA(cudaMalloc(reinterpret_cast<void **>(&s.xchg), s.xchgBytes));
    FS2D_CUDA(cudaMemset(s.xchg, 0, s.xchgBytes));
    {
        cudaFuncAttributes at;
        cudaFuncGetAttributes(&at, slabHaloKernel);
        cudaFuncGetAttributes(&at, slabClassifyKernel);
        cudaFuncGetAttributes(&at, slabParticleKernel);
        cudaFuncGetAttributes(&at, slabAppendKernel);
        cudaFuncGetAttributes(&at, slabGatherKernel);
        cudaFuncGetAttributes(&at, slabGatherRowsKernel);
        cudaGetLastError();
        pcgPreloadSlabKernels();
    }
    s.peerHeap[rank] = ctx->heap;
    s.peerXchg[rank] = s.xchg;
    s.connected = 0;
    s.enabled = true;
    return FS2D_OK;
}

int fs2d_slab_rows(fs2d_handle ctx, int *row_begin, int *row_end, int *halo_rows)
{
    if (!ctx) return FS2D_ERR_ARG;
    const SlabRows r = slabOwn(ctx);
    if (row_begin) *row_begin = r.lo;
    if (row_end) *row_end = r.hi;
    if (halo_rows) *halo_rows = ctx->slab.enabled ? ctx->slab.halo : 0;
    return FS2D_OK;
}

int fs2d_slab_export(fs2d_handle ctx, void *handle_out)
{
    if (!ctx || !handle_out) return FS2D_ERR_ARG;
    if (!ctx->slab.enabled)
    {
        ctx->lastError = "fs2d_slab_export: call fs2d_slab_configure first";
        return FS2D_ERR_STATE;
    }
    ExportBlob b;
    memset(&b, 0, sizeof(b));
    FS2D_CUDA(cudaIpcGetMemHandle(&b.heap, ctx->heap));
    FS2D_CUDA(cudaIpcGetMemHandle(&b.xchg, ctx->slab.xchg));
    b.pid = static_cast<unsigned long long>(getpid());
    b.heapPtr = reinterpret_cast<unsigned long long>(ctx->heap);
    b.xchgPtr = reinterpret_cast<unsigned long long>(ctx->slab.xchg);
    b.heapBytes = ctx->heapBytes;
    b.xchgBytes = ctx->slab.xchgBytes;
    b.device = ctx->device;
    b.rank = ctx->slab.rank;
    memset(handle_out, 0, FS2D_SLAB_HANDLE_BYTES);
    memcpy(handle_out, &b, sizeof(b));
    return FS2D_OK;
}

int fs2d_slab_connect(fs2d_handle ctx, int peer_rank, const void *handle)
{
    if (!ctx || !handle) return FS2D_ERR_ARG;
    SlabState &s = ctx->slab;
    if (!s.enabled || peer_rank < 0 || peer_rank >= s.world || peer_rank == s.rank || s.peerHeap[peer_rank])
    {
        ctx->lastError = "fs2d_slab_connect: bad peer rank or state";
        return FS2D_ERR_ARG;
    }
    ExportBlob b;
    memcpy(&b, handle, sizeof(b));
    if (b.rank != peer_rank || b.heapBytes != ctx->heapBytes || b.xchgBytes != s.xchgBytes)
    {
        ctx->lastError = "fs2d_slab_connect: the peer was created with different parameters";
        return FS2D_ERR_ARG;
    }
    FS2D_CUDA(cudaSetDevice(ctx->device));
    if (b.pid == static_cast<unsigned long long>(getpid()))
    {
        // same process: the peer's pointers are valid here; another device needs peer access switched on
        if (b.device != ctx->device)
        {
            cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
            {
                ctx->lastError = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
                return FS2D_ERR_COMM;
            }
            cudaGetLastError();
        }
        s.peerHeap[peer_rank] = reinterpret_cast<unsigned char *>(b.heapPtr);
        s.peerXchg[peer_rank] = reinterpret_cast<unsigned char *>(b.xchgPtr);
    }
    else
    {
        void *ph = nullptr, *px = nullptr;
        FS2D_CUDA(cudaIpcOpenMemHandle(&ph, b.heap, cudaIpcMemLazyEnablePeerAccess));
        FS2D_CUDA(cudaIpcOpenMemHandle(&px, b.xchg, cudaIpcMemLazyEnablePeerAccess));
        s.peerHeap[peer_rank] = static_cast<unsigned char *>(ph);
        s.peerXchg[peer_rank] = static_cast<unsigned char *>(px);
        s.peerMapped[peer_rank] = true;
    }
    s.connected++;
    return FS2D_OK;
}

int fs2d_slab_allgather(fs2d_handle ctx, const int64_t value[4], int64_t *out)
{
    if (!ctx || !value || !out) return FS2D_ERR_ARG;
    long long v[4] = {value[0], value[1], value[2], value[3]};
    long long o[4 * FS2D_MAX_RANKS];
    FS2D_TRY(slabAllGather(ctx, v, o));
    const int n = ctx->slab.enabled ? ctx->slab.world : 1;
    for (int k = 0; k < 4 * n; k++) out[k] = o[k];
    return FS2D_OK;
}

}  // extern "C"
