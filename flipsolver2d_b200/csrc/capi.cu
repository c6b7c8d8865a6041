// C ABI glue (include/fs2d.h): handle lifetime, state transfer and the stage entry points.
#include <algorithm>
#include <cstring>
#include <new>

#include <cstdlib>

#include "fs2d_internal.h"

int pcgTileBlocks(const Ctx *ctx);
void slabRelease(Ctx *ctx);

namespace
{
struct GridDesc
{
    void **ptr;
    int elemSize;
    int64_t count;
};

GridDesc gridDesc(Ctx *c, int grid)
{
    switch (grid)
    {
    case FS2D_GRID_U: return {reinterpret_cast<void **>(&c->U), 4, c->NU};
    case FS2D_GRID_V: return {reinterpret_cast<void **>(&c->V), 4, c->NV};
    case FS2D_GRID_U_VALID: return {reinterpret_cast<void **>(&c->uValid), 1, c->NU};
    case FS2D_GRID_V_VALID: return {reinterpret_cast<void **>(&c->vValid), 1, c->NV};
    case FS2D_GRID_SAVED_U: return {reinterpret_cast<void **>(&c->savedU), 4, c->NU};
    case FS2D_GRID_SAVED_V: return {reinterpret_cast<void **>(&c->savedV), 4, c->NV};
    case FS2D_GRID_MATERIAL: return {reinterpret_cast<void **>(&c->material), 1, c->N};
    case FS2D_GRID_FLUID_SDF: return {reinterpret_cast<void **>(&c->fluidSdf), 4, c->N};
    case FS2D_GRID_SOLID_SDF: return {reinterpret_cast<void **>(&c->solidSdf), 4, c->N};
    case FS2D_GRID_VISCOSITY: return {reinterpret_cast<void **>(&c->viscosity), 4, c->N};
    case FS2D_GRID_DENSITY: return {reinterpret_cast<void **>(&c->density), 4, c->N};
    case FS2D_GRID_COUNTS: return {reinterpret_cast<void **>(&c->counts), 4, c->N};
    case FS2D_GRID_EMITTER_ID: return {reinterpret_cast<void **>(&c->emitterId), 4, c->N};
    case FS2D_GRID_SOLID_ID: return {reinterpret_cast<void **>(&c->solidId), 4, c->N};
    case FS2D_GRID_DIVERGENCE_CONTROL: return {reinterpret_cast<void **>(&c->divergenceControl), 4, c->N};
    case FS2D_GRID_TEST: return {reinterpret_cast<void **>(&c->testGrid), 4, c->N};
    case FS2D_GRID_KNOWN_CENTERED: return {reinterpret_cast<void **>(&c->knownCentered), 1, c->N};
    case FS2D_GRID_TEMPERATURE: return {reinterpret_cast<void **>(&c->temperature), 4, c->N};
    case FS2D_GRID_CONCENTRATION: return {reinterpret_cast<void **>(&c->concentration), 4, c->N};
    case FS2D_GRID_FUEL: return {reinterpret_cast<void **>(&c->fuel), 4, c->N};
    case FS2D_GRID_PRESSURE: return {reinterpret_cast<void **>(&c->x), 8, c->N};
    case FS2D_GRID_RHS: return {reinterpret_cast<void **>(&c->rhs), 8, c->N};
    case FS2D_GRID_SOURCE_SDF: return {reinterpret_cast<void **>(&c->sourceSdf), 4, c->N};
    case FS2D_GRID_SOURCE_SDF_ID: return {reinterpret_cast<void **>(&c->sourceSdfId), 4, c->N};
    case FS2D_GRID_ADVECTED_U: return {reinterpret_cast<void **>(&c->advU), 4, c->NU};
    case FS2D_GRID_ADVECTED_V: return {reinterpret_cast<void **>(&c->advV), 4, c->NV};
    case FS2D_GRID_ADVECTED_SDF: return {reinterpret_cast<void **>(&c->advSdf), 4, c->N};
    case FS2D_GRID_ADVECTED_VISCOSITY: return {reinterpret_cast<void **>(&c->advViscosity), 4, c->N};
    default: return {nullptr, 0, 0};
    }
}

// Every dense array is carved from one allocation (the "symmetric heap"): same parameters -> same offsets
// on every rank, so a peer that maps the heap through one IPC handle finds each array where its own is.
struct HeapPlan
{
    struct Item
    {
        void **ptr;
        size_t bytes;
        int fill;
    };
    std::vector<Item> items;
    size_t total = 0;
    template <class T> void add(T **p, int64_t count, int fillByte = 0)
    {
        size_t bytes = static_cast<size_t>(std::max<int64_t>(count, 1)) * sizeof(T);
        bytes = (bytes + 255) & ~static_cast<size_t>(255);
        items.push_back({reinterpret_cast<void **>(p), bytes, fillByte});
        total += bytes;
    }
};

template <class T> int devAlloc(HeapPlan &plan, T **p, int64_t count, int fillByte = 0)
{
    plan.add(p, count, fillByte);
    return FS2D_OK;
}

__global__ void fillFloatKernel(float *p, long long n, float v)
{
    for (long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; k < n;
         k += static_cast<long long>(gridDim.x) * blockDim.x)
        p[k] = v;
}

// Packed particle transfers: the storage-bin byte of a record doubles as its dead flag (254).
#define FS2D_PACKED_DEAD 254
__global__ void packStorageKernel(const uint8_t *__restrict__ mis, const uint8_t *__restrict__ dead, long long n, unsigned char *__restrict__ out)
{
    const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (p < n) out[p] = dead[p] ? FS2D_PACKED_DEAD : mis[p];
}

__global__ void unpackStorageKernel(const unsigned char *__restrict__ in, long long n, uint8_t *__restrict__ mis, uint8_t *__restrict__ dead,
                                    unsigned long long *__restrict__ killed)
{
    const long long p = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const bool isDead = p < n && in[p] == FS2D_PACKED_DEAD;
    if (p < n)
    {
        mis[p] = isDead ? FS2D_MIS_HOME : in[p];
        dead[p] = isDead ? 1 : 0;
    }
    const unsigned int votes = __popc(__ballot_sync(0xffffffffu, isDead));
    if ((threadIdx.x & 31) == 0 && votes) atomicAdd(killed, static_cast<unsigned long long>(votes));
}

__global__ void fillIntKernel(int32_t *p, long long n, int32_t v)
{
    for (long long k = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; k < n;
         k += static_cast<long long>(gridDim.x) * blockDim.x)
        p[k] = v;
}

int allocAll(Ctx *realCtx)
{
    HeapPlan plan;
    HeapPlan *ctxPlan = &plan;
    Ctx *ctx = realCtx;
    const int64_t N = ctx->N, NU = ctx->NU, NV = ctx->NV;
    const bool smoke = ctx->p.sim_type == FS2D_SIM_SMOKE || ctx->p.sim_type == FS2D_SIM_FIRE;
    const bool fire = ctx->p.sim_type == FS2D_SIM_FIRE;
    const bool nb = ctx->p.sim_type == FS2D_SIM_NBFLIP;
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->U, NU));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->V, NV));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->savedU, NU));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->savedV, NV));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->uValid, NU));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->vValid, NV));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->material, N, FS2D_EMPTY));  // MaterialGrid init value (materialgrid.cpp:5-8)
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->fluidSdf, N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->solidSdf, N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->viscosity, N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->density, N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->counts, N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->emitterId, N, 0xFF));       // -1 (flipsolver2d.cpp:38-39)
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->solidId, N, 0xFF));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->divergenceControl, N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->testGrid, N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->knownCentered, N));
    if (smoke)
    {
        FS2D_TRY(devAlloc(*ctxPlan, &ctx->temperature, N));
        FS2D_TRY(devAlloc(*ctxPlan, &ctx->concentration, N));
    }
    if (fire) FS2D_TRY(devAlloc(*ctxPlan, &ctx->fuel, N));
    if (nb)
    {
        FS2D_TRY(devAlloc(*ctxPlan, &ctx->sourceSdf, N));
        FS2D_TRY(devAlloc(*ctxPlan, &ctx->sourceSdfId, N, 0xFF));
        FS2D_TRY(devAlloc(*ctxPlan, &ctx->advU, NU));
        FS2D_TRY(devAlloc(*ctxPlan, &ctx->advV, NV));
        FS2D_TRY(devAlloc(*ctxPlan, &ctx->advSdf, N));
        FS2D_TRY(devAlloc(*ctxPlan, &ctx->advViscosity, N));
    }
    const int64_t big = std::max(NU, NV);
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->scratchA, big));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->scratchB, big));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->scratchC, big));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->markers, big));

    FS2D_TRY(devAlloc(*ctxPlan, &ctx->rhs, N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->x, N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->r[0], N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->r[1], N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->s[0], N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->s[1], N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->q, N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->z, N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->rowInfo, N));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->preInfo, N));
    ctx->maxBlocks = std::max({pcgTileBlocks(ctx), ctx->smCount * 8, ctx->p.convergence_threads, 1024});
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->partials, 3 * static_cast<int64_t>(ctx->maxBlocks)));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->scalars, 1));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->tileFlags, pcgTileBlocks(ctx)));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->activeTiles, pcgTileBlocks(ctx)));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->activeCount, 1));
    ctx->traceCapacity = std::max(ctx->p.pcg_iter_limit, 16) + 8;
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->trace, 4 * static_cast<int64_t>(ctx->traceCapacity)));

    FS2D_TRY(devAlloc(*ctxPlan, &ctx->cellStart, N + 1));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->cellCursor, N + 1));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->reseedOffset, N + 1));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->scanBlock, divUp(N + 1, 1024) + 1024));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->d_counter, 16));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->d_fscratch, 4096));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->mail, 1));
    FS2D_TRY(devAlloc(*ctxPlan, &ctx->haloLL, 8 * static_cast<int64_t>(ctx->J)));
    FS2D_CUDA(cudaMalloc(reinterpret_cast<void **>(&ctx->heap), plan.total));
    ctx->heapBytes = plan.total;
    {
        size_t off = 0;
        for (const HeapPlan::Item &it : plan.items)
        {
            *it.ptr = ctx->heap + off;
            FS2D_CUDA(cudaMemsetAsync(ctx->heap + off, it.fill, it.bytes, ctx->stream));
            off += it.bytes;
        }
    }
    if (smoke) fillFloatKernel<<<ctx->smCount * 4, 256, 0, ctx->stream>>>(ctx->temperature, N, ctx->p.ambient_temperature);
    for (int k = 0; k < 16; k++) FS2D_CUDA(cudaEventCreate(&ctx->ev[k]));
    ctx->eventsReady = true;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

void freeAll(Ctx *c)
{
    slabRelease(c);
    if (c->heap) cudaFree(c->heap);  // every dense array lives inside the heap
    if (c->viscScalars) cudaFree(c->viscScalars);
    if (c->heavyBuf) cudaFree(c->heavyBuf);
    if (c->bfsQueue) cudaFree(c->bfsQueue);
    if (c->p2gTileList) cudaFree(c->p2gTileList);
    if (c->bfsCtl) cudaFree(c->bfsCtl);
    if (c->solveMaps) std::free(c->solveMaps);
    void *ptrs[] = {c->rangeLast, c->dead, c->perm, c->obstacleFriction, c->sources, c->reseedUniform, c->stage};
    for (void *p : ptrs)
        if (p) cudaFree(p);
    for (int b = 0; b < 2; b++)
    {
        if (c->pb[b].pos) cudaFree(c->pb[b].pos);
        if (c->pb[b].vel) cudaFree(c->pb[b].vel);
        if (c->pb[b].props) cudaFree(c->pb[b].props);
        if (c->pb[b].key) cudaFree(c->pb[b].key);
        if (c->pb[b].mis) cudaFree(c->pb[b].mis);
    }
    if (c->eventsReady)
        for (int k = 0; k < 16; k++) cudaEventDestroy(c->ev[k]);
    for (cudaEvent_t e : c->profEvents) cudaEventDestroy(e);
    if (c->pstream.copy)
    {
        cudaStreamDestroy(c->pstream.copy);
        cudaEvent_t ev[10] = {c->pstream.evByte, c->pstream.evVel, c->pstream.evPos, c->pstream.evProps, c->pstream.evMain,
                              c->pstream.evStart, c->pstream.evEarly0, c->pstream.evEarly1, c->pstream.evEnd0, c->pstream.evEnd1};
        for (cudaEvent_t e : ev)
            if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : c->pstream.evPosChunk)
            if (e) cudaEventDestroy(e);
    }
    if (c->stream) cudaStreamDestroy(c->stream);
}
}  // namespace

// Kernels of several ranks may wait for each other on the device (slab.cu, pcg.cu). CUDA's default lazy module
// loading synchronises the context at the first use of a kernel, which deadlocks (until the spin limit) when that
// first use happens on one host thread while a kernel of the same context spins for a launch another host thread
// has not made yet. Ask for eager loading before the driver is initialised; fs2d_slab_configure additionally
// touches the kernels that only exist in slab mode.
__attribute__((constructor)) static void fs2dEagerModuleLoading() { setenv("CUDA_MODULE_LOADING", "EAGER", 0); }

// Streamed uploads (fs2d_particle_stream_begin): the solver's stream waits for a section of the host buffer where the
// substep first needs it.
bool particleStreamNextPosChunk(Ctx *ctx, int idx, int64_t *begin, int64_t *end)
{
    Ctx::ParticleStream &ps = ctx->pstream;
    if (!ps.posPending || idx < 0 || idx >= Ctx::ParticleStream::POS_CHUNKS) return false;
    if (cudaStreamWaitEvent(ctx->stream, ps.evPosChunk[idx], 0) != cudaSuccess) return false;
    *begin = idx == 0 ? 0 : ps.posChunkEnd[idx - 1];
    *end = ps.posChunkEnd[idx];
    return true;
}

int particleStreamSettleSlow(Ctx *ctx, bool all)
{
    Ctx::ParticleStream &ps = ctx->pstream;
    if (ps.posPending)
    {
        FS2D_CUDA(cudaStreamWaitEvent(ctx->stream, ps.evPos, 0));
        ps.posPending = false;
        FS2D_TRY(particlesKeyRange(ctx, 0, ctx->count));
    }
    if (!all) return FS2D_OK;
    if (ps.propsPending)
    {
        FS2D_CUDA(cudaStreamWaitEvent(ctx->stream, ps.evProps, 0));
        ps.propsPending = false;
    }
    if (ps.propsGatherPending)
    {
        ps.propsGatherPending = false;
        FS2D_TRY(particlesGatherProps(ctx, ps.propsFrom, ps.gatherCount));
    }
    return FS2D_OK;
}

extern "C" {

int fs2d_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
    return n;
}

int fs2d_create(const fs2d_params *params, fs2d_handle *out)
{
    if (!params || !out || params->size_i <= 0 || params->size_j <= 0 || params->num_properties < 0) return FS2D_ERR_ARG;
    *out = nullptr;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) return FS2D_ERR_NO_DEVICE;  // no CPU fallback, by design
    if (params->device < 0 || params->device >= n) return FS2D_ERR_ARG;
    Ctx *ctx = new (std::nothrow) Ctx();
    if (!ctx) return FS2D_ERR_ARG;
    ctx->p = *params;
    ctx->I = params->size_i;
    ctx->J = params->size_j;
    ctx->N = static_cast<int64_t>(ctx->I) * ctx->J;
    ctx->NU = static_cast<int64_t>(ctx->I + 1) * ctx->J;
    ctx->NV = static_cast<int64_t>(ctx->I) * (ctx->J + 1);
    ctx->device = params->device;
    ctx->stepDt = 0.f;
    int rc = FS2D_OK;
    do
    {
        if (cudaSetDevice(ctx->device) != cudaSuccess) { rc = FS2D_ERR_CUDA; break; }
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, ctx->device) != cudaSuccess) { rc = FS2D_ERR_CUDA; break; }
        ctx->smCount = prop.multiProcessorCount;
        if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = FS2D_ERR_CUDA; break; }
        rc = allocAll(ctx);
        if (rc == FS2D_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) rc = FS2D_ERR_CUDA;
    } while (0);
    if (rc != FS2D_OK)
    {
        fprintf(stderr, "fs2d_create failed: %s\n", ctx->lastError.c_str());
        freeAll(ctx);
        delete ctx;
        return rc;
    }
    *out = ctx;
    return FS2D_OK;
}

int fs2d_destroy(fs2d_handle h)
{
    if (!h) return FS2D_ERR_ARG;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
    freeAll(h);
    delete h;
    return FS2D_OK;
}

const char *fs2d_last_error(fs2d_handle h) { return h ? h->lastError.c_str() : "null handle"; }

int fs2d_synchronize(fs2d_handle ctx)
{
    if (!ctx) return FS2D_ERR_ARG;
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    return FS2D_OK;
}

void *fs2d_stream(fs2d_handle h) { return h ? static_cast<void *>(h->stream) : nullptr; }

int64_t fs2d_launch_count(fs2d_handle h) { return h ? h->launches : 0; }

int64_t fs2d_grid_elements(fs2d_handle h, int grid) { return h ? gridDesc(h, grid).count : 0; }

int fs2d_grid_element_size(int grid)
{
    Ctx dummy;
    return gridDesc(&dummy, grid).elemSize;
}

void *fs2d_grid_device_ptr(fs2d_handle h, int grid)
{
    if (!h) return nullptr;
    if (grid == FS2D_GRID_FLUID_SDF && gridFlushSdf(h) != FS2D_OK) return nullptr;
    GridDesc d = gridDesc(h, grid);
    return d.ptr ? *d.ptr : nullptr;
}

int fs2d_upload_grid(fs2d_handle ctx, int grid, const void *host_data, size_t bytes)
{
    if (!ctx || !host_data) return FS2D_ERR_ARG;
    GridDesc d = gridDesc(ctx, grid);
    if (!d.ptr || !*d.ptr || bytes != static_cast<size_t>(d.count) * d.elemSize)
    {
        ctx->lastError = "fs2d_upload_grid: unknown grid or size mismatch";
        return FS2D_ERR_ARG;
    }
    if (grid == FS2D_GRID_FLUID_SDF) ctx->sdfInsidePending = false;
    FS2D_CUDA(cudaMemcpyAsync(*d.ptr, host_data, bytes, cudaMemcpyHostToDevice, ctx->stream));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    return FS2D_OK;
}

int fs2d_download_grid(fs2d_handle ctx, int grid, void *host_data, size_t bytes)
{
    if (!ctx || !host_data) return FS2D_ERR_ARG;
    GridDesc d = gridDesc(ctx, grid);
    if (!d.ptr || !*d.ptr || bytes != static_cast<size_t>(d.count) * d.elemSize)
    {
        ctx->lastError = "fs2d_download_grid: unknown grid or size mismatch";
        return FS2D_ERR_ARG;
    }
    const void *src = *d.ptr;
    if (grid == FS2D_GRID_FLUID_SDF)
    {
        const float *field = nullptr;
        FS2D_TRY(gridSdfForRead(ctx, &field));  // deferred (water) / banded (NBFlip) level-set walks completed for the reader
        src = field;
    }
    FS2D_CUDA(fs2dCopyToHost(ctx, host_data, src, bytes));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    return FS2D_OK;
}

int fs2d_slab_gather_grid(fs2d_handle ctx, int grid)
{
    if (!ctx) return FS2D_ERR_ARG;
    GridDesc d = gridDesc(ctx, grid);
    if (!d.ptr || !*d.ptr)
    {
        ctx->lastError = "fs2d_slab_gather_grid: unknown grid";
        return FS2D_ERR_ARG;
    }
    if (!ctx->slab.enabled || ctx->slab.world == 1) return grid == FS2D_GRID_FLUID_SDF ? gridFlushSdf(ctx) : FS2D_OK;
    const bool uLike = grid == FS2D_GRID_U || grid == FS2D_GRID_U_VALID || grid == FS2D_GRID_SAVED_U || grid == FS2D_GRID_ADVECTED_U;
    const bool vLike = grid == FS2D_GRID_V || grid == FS2D_GRID_V_VALID || grid == FS2D_GRID_SAVED_V || grid == FS2D_GRID_ADVECTED_V;
    const size_t rowBytes = static_cast<size_t>(d.elemSize) * (vLike ? ctx->J + 1 : ctx->J);
    FS2D_TRY(slabGatherRows(ctx, *d.ptr, rowBytes, uLike ? ctx->I + 1 : ctx->I));
    if (grid == FS2D_GRID_FLUID_SDF) FS2D_TRY(gridFlushSdfGathered(ctx));
    return FS2D_OK;
}

int fs2d_clear_grid(fs2d_handle ctx, int grid)
{
    if (!ctx) return FS2D_ERR_ARG;
    GridDesc d = gridDesc(ctx, grid);
    if (!d.ptr || !*d.ptr)
    {
        ctx->lastError = "fs2d_clear_grid: unknown grid";
        return FS2D_ERR_ARG;
    }
    FS2D_CUDA(cudaMemsetAsync(*d.ptr, 0, static_cast<size_t>(d.count) * d.elemSize, ctx->stream));
    return FS2D_OK;
}

int fs2d_set_obstacles(fs2d_handle ctx, int count, const float *host_friction)
{
    if (!ctx || count < 0 || (count > 0 && !host_friction)) return FS2D_ERR_ARG;
    if (ctx->obstacleFriction) cudaFree(ctx->obstacleFriction);
    ctx->obstacleFriction = nullptr;
    ctx->numObstacles = count;
    ctx->obstaclesFrictionless = true;
    for (int k = 0; k < count; k++) ctx->obstaclesFrictionless = ctx->obstaclesFrictionless && host_friction[k] == 0.f;
    FS2D_CUDA(cudaMalloc(reinterpret_cast<void **>(&ctx->obstacleFriction), sizeof(float) * std::max(count, 1)));
    if (count > 0)
        FS2D_CUDA(cudaMemcpy(ctx->obstacleFriction, host_friction, sizeof(float) * count, cudaMemcpyHostToDevice));
    return FS2D_OK;
}

int fs2d_set_sources(fs2d_handle ctx, int count, const fs2d_source *host_sources)
{
    if (!ctx || count < 0 || (count > 0 && !host_sources)) return FS2D_ERR_ARG;
    if (ctx->sources) cudaFree(ctx->sources);
    ctx->sources = nullptr;
    ctx->numSources = count;
    ctx->hostSources.assign(host_sources, host_sources + count);
    FS2D_CUDA(cudaMalloc(reinterpret_cast<void **>(&ctx->sources), sizeof(fs2d_source) * std::max(count, 1)));
    if (count > 0)
        FS2D_CUDA(cudaMemcpy(ctx->sources, host_sources, sizeof(fs2d_source) * count, cudaMemcpyHostToDevice));
    return FS2D_OK;
}

// ---------------------------------------------------------------- particles
int64_t fs2d_particle_count(fs2d_handle h)
{
    if (!h) return 0;
    int64_t n = 0;
    if (particlesAliveCount(h, &n) != FS2D_OK) return -1;
    return n;
}

int fs2d_upload_particles(fs2d_handle ctx, int64_t count, const float *host_pos, const float *host_vel,
                          const float *host_props)
{
    if (!ctx || count < 0) return FS2D_ERR_ARG;
    FS2D_TRY(particleStreamSettleAll(ctx));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    FS2D_CUDA(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), ctx->stream));
    ctx->pstream.earlyCount = ctx->pstream.earlyVelCount = -1;
    ctx->count = 0;
    ctx->deadCount = 0;
    ctx->killedDirty = false;
    ctx->sorted = false;
    ctx->slab.ghostCount = 0;  // slab mode: the caller uploads the particles this rank owns
    ctx->slab.ownedBegin = ctx->slab.ownedEnd = 0;
    if (count == 0)
    {
        FS2D_CUDA(cudaMemsetAsync(ctx->cellStart, 0, sizeof(int32_t) * (ctx->N + 1), ctx->stream));
        ctx->sorted = true;
    }
    if (ctx->slab.enabled && ctx->slab.world > 1) FS2D_TRY(particlesReserve(ctx, count + 2 * ctx->slab.xchgCapacity));
    return fs2d_append_particles(ctx, count, host_pos, host_vel, host_props);
}

int fs2d_append_particles(fs2d_handle ctx, int64_t count, const float *host_pos, const float *host_vel,
                          const float *host_props)
{
    if (!ctx || count < 0 || (count > 0 && !host_pos)) return FS2D_ERR_ARG;
    if (count == 0) return FS2D_OK;
    FS2D_TRY(particleStreamSettleAll(ctx));
    const int64_t base = ctx->count;
    // slab mode: room for the ghosts and migrants of the neighbours up front, so that no exchange has to grow the
    // buffers (cudaFree synchronises the device, which must not happen while a peer spins on this rank)
    const int64_t slack = (ctx->slab.enabled && ctx->slab.world > 1) ? 2 * ctx->slab.xchgCapacity : 0;
    FS2D_TRY(particlesReserve(ctx, base + count + slack));
    ParticleBuffers &b = ctx->pb[ctx->cur];
    FS2D_CUDA(cudaMemcpyAsync(b.pos + base, host_pos, sizeof(float2) * count, cudaMemcpyHostToDevice, ctx->stream));
    if (host_vel)
        FS2D_CUDA(cudaMemcpyAsync(b.vel + base, host_vel, sizeof(float2) * count, cudaMemcpyHostToDevice, ctx->stream));
    else
        FS2D_CUDA(cudaMemsetAsync(b.vel + base, 0, sizeof(float2) * count, ctx->stream));
    for (int k = 0; k < ctx->p.num_properties; k++)
    {
        float *dst = b.props + static_cast<int64_t>(k) * b.capacity + base;
        if (host_props)
            FS2D_CUDA(cudaMemcpyAsync(dst, host_props + static_cast<int64_t>(k) * count, sizeof(float) * count,
                                      cudaMemcpyHostToDevice, ctx->stream));
        else
            FS2D_CUDA(cudaMemsetAsync(dst, 0, sizeof(float) * count, ctx->stream));
    }
    FS2D_CUDA(cudaMemsetAsync(ctx->dead + base, 0, count, ctx->stream));
    FS2D_CUDA(cudaMemsetAsync(b.mis + base, FS2D_MIS_HOME, count, ctx->stream));  // filed in the bin of its position
    FS2D_TRY(particlesKeyRange(ctx, base, base + count));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->count = base + count;
    ctx->sorted = false;
    return FS2D_OK;
}

int fs2d_download_particles(fs2d_handle ctx, float *host_pos, float *host_vel, float *host_props)
{
    if (!ctx) return FS2D_ERR_ARG;
    FS2D_TRY(particleStreamSettleAll(ctx));
    // dead-flagged particles (cap, sinks, narrow band) are dropped by the sort; do it now so the
    // caller sees exactly fs2d_particle_count() records in device order
    int64_t alive = 0;
    FS2D_TRY(particlesAliveCount(ctx, &alive));
    const bool slab = ctx->slab.enabled && ctx->slab.world > 1;
    if (alive != ctx->count - ctx->slab.ghostCount || !ctx->sorted) FS2D_TRY(particlesSort(ctx));
    // slab mode: only the particles this rank owns (the sorted arrays also hold ghost copies of the neighbours')
    const int64_t first = slab ? ctx->slab.ownedBegin : 0;
    const int64_t n = slab ? ctx->slab.ownedEnd - ctx->slab.ownedBegin : ctx->count;
    if (n == 0) return FS2D_OK;
    ParticleBuffers &b = ctx->pb[ctx->cur];
    if (host_pos) FS2D_CUDA(fs2dCopyToHost(ctx, host_pos, b.pos + first, sizeof(float2) * n));
    if (host_vel) FS2D_CUDA(fs2dCopyToHost(ctx, host_vel, b.vel + first, sizeof(float2) * n));
    if (host_props)
        for (int k = 0; k < ctx->p.num_properties; k++)
            FS2D_CUDA(fs2dCopyToHost(ctx, host_props + static_cast<int64_t>(k) * n, b.props + static_cast<int64_t>(k) * b.capacity + first,
                                      sizeof(float) * n));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    return FS2D_OK;
}

// ---- packed particle state: ONE copy per direction, lossless (the storage-bin byte travels with the record)
static int stageReserve(Ctx *ctx, size_t bytes)
{
    if (bytes <= ctx->stageBytes) return FS2D_OK;
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ctx->stage) cudaFree(ctx->stage);
    ctx->stage = nullptr;
    ctx->stageBytes = 0;
    const size_t want = bytes + bytes / 4 + 4096;
    FS2D_CUDA(cudaMalloc(reinterpret_cast<void **>(&ctx->stage), want));
    ctx->stageBytes = want;
    return FS2D_OK;
}

size_t fs2d_packed_particle_bytes(fs2d_handle ctx, int64_t count)
{
    if (!ctx || count < 0) return 0;
    return static_cast<size_t>(count) * (16u + 4u * static_cast<size_t>(ctx->p.num_properties) + 1u);
}

int fs2d_download_particles_packed(fs2d_handle ctx, void *host_buf, size_t capacity_bytes, int64_t *count)
{
    if (!ctx || !host_buf || !count) return FS2D_ERR_ARG;
    FS2D_TRY(particleStreamSettleAll(ctx));
    const bool slab = ctx->slab.enabled && ctx->slab.world > 1;
    int64_t first = 0, n = ctx->count;
    if (slab)
    {
        // row slabs: the arrays also hold ghost copies of the neighbours' particles; after a sort the owned records
        // are one contiguous range
        int64_t alive = 0;
        FS2D_TRY(particlesAliveCount(ctx, &alive));
        if (alive != ctx->count - ctx->slab.ghostCount || !ctx->sorted) FS2D_TRY(particlesSort(ctx));
        first = ctx->slab.ownedBegin;
        n = ctx->slab.ownedEnd - ctx->slab.ownedBegin;
    }
    // one handle: every record in the CURRENT device order, flagged-dead ones included (storage byte 254) -- no sort,
    // no host synchronisation before the copy, and the round trip restores the arrays exactly as they are
    *count = n;
    const size_t bytes = fs2d_packed_particle_bytes(ctx, n);
    if (bytes > capacity_bytes)
    {
        ctx->lastError = "fs2d_download_particles_packed: host buffer too small";
        return FS2D_ERR_ARG;
    }
    if (n == 0) return FS2D_OK;
    FS2D_TRY(stageReserve(ctx, bytes));
    const ParticleBuffers &b = ctx->pb[ctx->cur];
    const int K = ctx->p.num_properties;
    cudaStream_t st = ctx->stream;
    unsigned char *d = ctx->stage;
    FS2D_CUDA(cudaMemcpyAsync(d, b.pos + first, 8u * n, cudaMemcpyDeviceToDevice, st));
    FS2D_CUDA(cudaMemcpyAsync(d + 8u * n, b.vel + first, 8u * n, cudaMemcpyDeviceToDevice, st));
    for (int k = 0; k < K; k++)
        FS2D_CUDA(cudaMemcpyAsync(d + (16u + 4u * k) * n, b.props + static_cast<int64_t>(k) * b.capacity + first, 4u * n,
                                  cudaMemcpyDeviceToDevice, st));
    packStorageKernel<<<divUp(n, 256), 256, 0, st>>>(b.mis + first, ctx->dead + first, n, d + (16u + 4u * K) * n);
    ctx->launches++;
    FS2D_CUDA(cudaMemcpyAsync(host_buf, d, bytes, cudaMemcpyDeviceToHost, st));
    FS2D_CUDA(cudaStreamSynchronize(st));
    return FS2D_OK;
}

int fs2d_upload_particles_packed(fs2d_handle ctx, const void *host_buf, int64_t count)
{
    if (!ctx || count < 0 || (count > 0 && !host_buf)) return FS2D_ERR_ARG;
    FS2D_TRY(particleStreamSettleAll(ctx));
    const bool slab = ctx->slab.enabled && ctx->slab.world > 1;
    cudaStream_t st = ctx->stream;
    FS2D_CUDA(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), st));
    ctx->pstream.earlyCount = ctx->pstream.earlyVelCount = -1;
    ctx->count = 0;
    ctx->deadCount = 0;
    ctx->killedDirty = false;
    ctx->sorted = false;
    ctx->slab.ghostCount = 0;
    ctx->slab.ownedBegin = ctx->slab.ownedEnd = 0;
    if (count == 0)
    {
        FS2D_CUDA(cudaMemsetAsync(ctx->cellStart, 0, sizeof(int32_t) * (ctx->N + 1), st));
        ctx->sorted = true;
        return FS2D_OK;
    }
    FS2D_TRY(particlesReserve(ctx, count + (slab ? 2 * ctx->slab.xchgCapacity : 0)));
    const size_t bytes = fs2d_packed_particle_bytes(ctx, count);
    FS2D_TRY(stageReserve(ctx, bytes));
    ParticleBuffers &b = ctx->pb[ctx->cur];
    const int K = ctx->p.num_properties;
    const size_t n = static_cast<size_t>(count);
    unsigned char *d = ctx->stage;
    FS2D_CUDA(cudaMemcpyAsync(d, host_buf, bytes, cudaMemcpyHostToDevice, st));
    FS2D_CUDA(cudaMemcpyAsync(b.pos, d, 8u * n, cudaMemcpyDeviceToDevice, st));
    FS2D_CUDA(cudaMemcpyAsync(b.vel, d + 8u * n, 8u * n, cudaMemcpyDeviceToDevice, st));
    for (int k = 0; k < K; k++)
        FS2D_CUDA(cudaMemcpyAsync(b.props + static_cast<int64_t>(k) * b.capacity, d + (16u + 4u * k) * n, 4u * n, cudaMemcpyDeviceToDevice, st));
    // storage byte -> storage-bin code + dead flag; the dead ones are counted on the device (folded into the host's view
    // at the next particle count, like the kills of a stage)
    unpackStorageKernel<<<divUp(count, 256), 256, 0, st>>>(d + (16u + 4u * K) * n, count, b.mis, ctx->dead,
                                                           reinterpret_cast<unsigned long long *>(ctx->d_counter));
    ctx->launches++;
    ctx->count = count;
    ctx->killedDirty = true;
    FS2D_TRY(particlesKeyRange(ctx, 0, count));
    // the host buffer may be reused as soon as this returns
    FS2D_CUDA(cudaStreamSynchronize(st));
    return FS2D_OK;
}

// ---- streamed particle state: the copies of a substep's particle state overlap the substep itself
static int streamReady(Ctx *ctx)
{
    Ctx::ParticleStream &ps = ctx->pstream;
    if (ps.copy) return FS2D_OK;
    FS2D_CUDA(cudaStreamCreateWithFlags(&ps.copy, cudaStreamNonBlocking));
    cudaEvent_t *ev[10] = {&ps.evByte, &ps.evVel, &ps.evPos, &ps.evProps, &ps.evMain, &ps.evStart, &ps.evEarly0, &ps.evEarly1, &ps.evEnd0, &ps.evEnd1};
    for (cudaEvent_t *e : ev) FS2D_CUDA(cudaEventCreate(e));
    for (cudaEvent_t &e : ps.evPosChunk) FS2D_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    return FS2D_OK;
}

size_t fs2d_particle_stream_bytes(fs2d_handle ctx, int64_t capacity_records) { return fs2d_packed_particle_bytes(ctx, capacity_records); }

int fs2d_particle_stream_begin(fs2d_handle ctx, const void *host_in, int64_t count, int64_t capacity_records)
{
    if (!ctx || count < 0 || capacity_records < count || (count > 0 && !host_in)) return FS2D_ERR_ARG;
    if (ctx->slab.enabled && ctx->slab.world > 1)
    {
        ctx->lastError = "fs2d_particle_stream_begin: one handle only (row slabs use the packed transfers)";
        return FS2D_ERR_STATE;
    }
    FS2D_TRY(particleStreamSettleAll(ctx));
    FS2D_TRY(streamReady(ctx));
    Ctx::ParticleStream &ps = ctx->pstream;
    cudaStream_t st = ctx->stream;
    FS2D_CUDA(cudaMemsetAsync(ctx->d_counter, 0, sizeof(unsigned long long), st));
    ctx->count = 0;
    ctx->deadCount = 0;
    ctx->killedDirty = false;
    ctx->sorted = false;
    ps.earlyCount = ps.earlyVelCount = -1;
    if (count == 0)
    {
        FS2D_CUDA(cudaMemsetAsync(ctx->cellStart, 0, sizeof(int32_t) * (ctx->N + 1), st));
        ctx->sorted = true;
        return FS2D_OK;
    }
    FS2D_TRY(particlesReserve(ctx, count));
    FS2D_TRY(stageReserve(ctx, static_cast<size_t>(count)));
    ParticleBuffers &b = ctx->pb[ctx->cur];
    const int K = ctx->p.num_properties;
    const size_t n = static_cast<size_t>(count), cap = static_cast<size_t>(capacity_records);
    const unsigned char *h = static_cast<const unsigned char *>(host_in);
    // the copy stream starts when everything queued on the solver's stream (readers of the arrays about to be
    // overwritten) is done; sections in the order the substep needs them: dead flags and velocities (CFL maximum),
    // positions (advection, sort, density correction), property columns (centred P2G)
    FS2D_CUDA(cudaEventRecord(ps.evMain, st));
    FS2D_CUDA(cudaStreamWaitEvent(ps.copy, ps.evMain, 0));
    FS2D_CUDA(cudaEventRecord(ps.evStart, ps.copy));
    ps.timedUpload = true;
    ps.timedEarly = ps.timedEnd = false;
    FS2D_CUDA(cudaMemcpyAsync(ctx->stage, h + (16u + 4u * K) * cap, n, cudaMemcpyHostToDevice, ps.copy));
    FS2D_CUDA(cudaEventRecord(ps.evByte, ps.copy));
    FS2D_CUDA(cudaMemcpyAsync(b.vel, h + 8u * cap, 8u * n, cudaMemcpyHostToDevice, ps.copy));
    FS2D_CUDA(cudaEventRecord(ps.evVel, ps.copy));
    {
        // in chunks of whole 4096-record blocks: the advection of a chunk starts while the next one travels
        const int64_t per = ((count + Ctx::ParticleStream::POS_CHUNKS - 1) / Ctx::ParticleStream::POS_CHUNKS + 4095) / 4096 * 4096;
        int64_t done = 0;
        for (int c = 0; c < Ctx::ParticleStream::POS_CHUNKS; c++)
        {
            const int64_t end = std::min<int64_t>(count, done + per);
            if (end > done)
                FS2D_CUDA(cudaMemcpyAsync(b.pos + done, h + 8u * static_cast<size_t>(done), 8u * static_cast<size_t>(end - done), cudaMemcpyHostToDevice, ps.copy));
            FS2D_CUDA(cudaEventRecord(ps.evPosChunk[c], ps.copy));
            ps.posChunkEnd[c] = end;
            done = end;
        }
    }
    FS2D_CUDA(cudaEventRecord(ps.evPos, ps.copy));
    for (int k = 0; k < K; k++)
        FS2D_CUDA(cudaMemcpyAsync(b.props + static_cast<int64_t>(k) * b.capacity, h + (16u + 4u * k) * cap, 4u * n, cudaMemcpyHostToDevice, ps.copy));
    FS2D_CUDA(cudaEventRecord(ps.evProps, ps.copy));
    FS2D_CUDA(cudaStreamWaitEvent(st, ps.evByte, 0));
    unpackStorageKernel<<<divUp(count, 256), 256, 0, st>>>(ctx->stage, count, b.mis, ctx->dead, reinterpret_cast<unsigned long long *>(ctx->d_counter));
    ctx->launches++;
    FS2D_CUDA(cudaStreamWaitEvent(st, ps.evVel, 0));
    ctx->count = count;
    ctx->killedDirty = true;
    ps.posPending = true;
    ps.propsPending = K > 0;
    ps.propsGatherPending = false;
    FS2D_CUDA(cudaGetLastError());
    return FS2D_OK;
}

int fs2d_particle_stream_positions_final(fs2d_handle ctx, void *host_out, int64_t capacity_records, int props_final)
{
    if (!ctx || !host_out) return FS2D_ERR_ARG;
    if (ctx->slab.enabled && ctx->slab.world > 1) return FS2D_OK;  // nothing leaves early over row slabs
    FS2D_TRY(particleStreamSettleAll(ctx));
    FS2D_TRY(streamReady(ctx));
    Ctx::ParticleStream &ps = ctx->pstream;
    const int64_t n = ctx->count;
    if (n > capacity_records)
    {
        ctx->lastError = "fs2d_particle_stream_positions_final: host buffer too small";
        return FS2D_ERR_ARG;
    }
    ps.earlyCount = -1;
    if (n == 0) return FS2D_OK;
    const ParticleBuffers &b = ctx->pb[ctx->cur];
    const int K = ctx->p.num_properties;
    const size_t cap = static_cast<size_t>(capacity_records);
    unsigned char *h = static_cast<unsigned char *>(host_out);
    FS2D_CUDA(cudaEventRecord(ps.evMain, ctx->stream));
    FS2D_CUDA(cudaStreamWaitEvent(ps.copy, ps.evMain, 0));
    FS2D_CUDA(cudaEventRecord(ps.evEarly0, ps.copy));
    FS2D_CUDA(cudaMemcpyAsync(h, b.pos, 8u * static_cast<size_t>(n), cudaMemcpyDeviceToHost, ps.copy));
    if (props_final)
        for (int k = 0; k < K; k++)
            FS2D_CUDA(cudaMemcpyAsync(h + (16u + 4u * k) * cap, b.props + static_cast<int64_t>(k) * b.capacity, 4u * static_cast<size_t>(n),
                                      cudaMemcpyDeviceToHost, ps.copy));
    FS2D_CUDA(cudaEventRecord(ps.evEarly1, ps.copy));
    ps.timedEarly = true;
    ps.earlyCount = n;
    ps.earlyProps = props_final != 0;
    ps.earlyHost = host_out;
    ps.earlyCapacity = capacity_records;
    return FS2D_OK;
}

int fs2d_particle_stream_velocities_final(fs2d_handle ctx, void *host_out, int64_t capacity_records)
{
    if (!ctx || !host_out) return FS2D_ERR_ARG;
    if (ctx->slab.enabled && ctx->slab.world > 1) return FS2D_OK;
    FS2D_TRY(particleStreamSettleAll(ctx));
    FS2D_TRY(streamReady(ctx));
    Ctx::ParticleStream &ps = ctx->pstream;
    const int64_t n = ctx->count;
    if (n > capacity_records)
    {
        ctx->lastError = "fs2d_particle_stream_velocities_final: host buffer too small";
        return FS2D_ERR_ARG;
    }
    ps.earlyVelCount = -1;
    if (n == 0) return FS2D_OK;
    FS2D_CUDA(cudaEventRecord(ps.evMain, ctx->stream));
    FS2D_CUDA(cudaStreamWaitEvent(ps.copy, ps.evMain, 0));
    FS2D_CUDA(cudaMemcpyAsync(static_cast<unsigned char *>(host_out) + 8u * static_cast<size_t>(capacity_records), ctx->pb[ctx->cur].vel,
                              8u * static_cast<size_t>(n), cudaMemcpyDeviceToHost, ps.copy));
    ps.earlyVelCount = n;
    ps.earlyVelHost = host_out;
    ps.earlyVelCapacity = capacity_records;
    return FS2D_OK;
}

int fs2d_particle_stream_set_output(fs2d_handle ctx, void *host_out, int64_t capacity_records)
{
    if (!ctx || capacity_records < 0) return FS2D_ERR_ARG;
    ctx->pstream.outHost = host_out;
    ctx->pstream.outCapacity = capacity_records;
    return FS2D_OK;
}

int fs2d_particle_stream_end(fs2d_handle ctx, void *host_out, int64_t capacity_records, int64_t *count)
{
    if (!ctx || !host_out || !count) return FS2D_ERR_ARG;
    ctx->pstream.outHost = nullptr;
    if (ctx->slab.enabled && ctx->slab.world > 1)
    {
        ctx->lastError = "fs2d_particle_stream_end: one handle only (row slabs use the packed transfers)";
        return FS2D_ERR_STATE;
    }
    FS2D_TRY(particleStreamSettleAll(ctx));
    Ctx::ParticleStream &ps = ctx->pstream;
    const int64_t n = ctx->count;
    *count = n;
    if (n > capacity_records)
    {
        ctx->lastError = "fs2d_particle_stream_end: host buffer too small";
        return FS2D_ERR_ARG;
    }
    // what left early (fs2d_particle_stream_positions_final into the same buffer, nothing moved since)
    int64_t early = (ps.earlyCount >= 0 && ps.earlyHost == host_out && ps.earlyCapacity == capacity_records) ? std::min(ps.earlyCount, n) : 0;
    const bool earlyProps = early > 0 && ps.earlyProps;
    const int64_t earlyVel = (ps.earlyVelCount >= 0 && ps.earlyVelHost == host_out && ps.earlyVelCapacity == capacity_records) ? std::min(ps.earlyVelCount, n) : 0;
    cudaStream_t st = ctx->stream;
    if (n > 0)
    {
        FS2D_TRY(stageReserve(ctx, static_cast<size_t>(n)));
        const ParticleBuffers &b = ctx->pb[ctx->cur];
        const int K = ctx->p.num_properties;
        const size_t cap = static_cast<size_t>(capacity_records), un = static_cast<size_t>(n), ue = static_cast<size_t>(early);
        unsigned char *h = static_cast<unsigned char *>(host_out);
        if (ps.evEnd0)
        {
            FS2D_CUDA(cudaEventRecord(ps.evEnd0, st));
            ps.timedEnd = true;
        }
        packStorageKernel<<<divUp(n, 256), 256, 0, st>>>(b.mis, ctx->dead, n, ctx->stage);
        ctx->launches++;
        const size_t uv = static_cast<size_t>(earlyVel);
        if (un > uv) FS2D_CUDA(cudaMemcpyAsync(h + 8u * cap + 8u * uv, b.vel + earlyVel, 8u * (un - uv), cudaMemcpyDeviceToHost, st));
        FS2D_CUDA(cudaMemcpyAsync(h + (16u + 4u * K) * cap, ctx->stage, un, cudaMemcpyDeviceToHost, st));
        if (un > ue) FS2D_CUDA(cudaMemcpyAsync(h + 8u * ue, b.pos + early, 8u * (un - ue), cudaMemcpyDeviceToHost, st));
        const size_t pe = earlyProps ? ue : 0u;
        if (un > pe)
            for (int k = 0; k < K; k++)
                FS2D_CUDA(cudaMemcpyAsync(h + (16u + 4u * k) * cap + 4u * pe, b.props + static_cast<int64_t>(k) * b.capacity + pe, 4u * (un - pe),
                                          cudaMemcpyDeviceToHost, st));
    }
    ps.earlyCount = ps.earlyVelCount = -1;
    if (ps.timedEnd) FS2D_CUDA(cudaEventRecord(ps.evEnd1, st));
    if (ps.copy) FS2D_CUDA(cudaStreamSynchronize(ps.copy));
    FS2D_CUDA(cudaStreamSynchronize(st));
    return FS2D_OK;
}

int fs2d_particle_stream_timing(fs2d_handle ctx, float *ms6)
{
    if (!ctx || !ms6) return FS2D_ERR_ARG;
    Ctx::ParticleStream &ps = ctx->pstream;
    for (int k = 0; k < 6; k++) ms6[k] = 0.f;
    if (!ps.copy) return FS2D_OK;
    FS2D_CUDA(cudaStreamSynchronize(ps.copy));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    if (ps.timedUpload)
    {
        cudaEventElapsedTime(&ms6[0], ps.evStart, ps.evByte);
        cudaEventElapsedTime(&ms6[1], ps.evByte, ps.evVel);
        cudaEventElapsedTime(&ms6[2], ps.evVel, ps.evPos);
        cudaEventElapsedTime(&ms6[3], ps.evPos, ps.evProps);
    }
    if (ps.timedEarly) cudaEventElapsedTime(&ms6[4], ps.evEarly0, ps.evEarly1);
    if (ps.timedEnd) cudaEventElapsedTime(&ms6[5], ps.evEnd0, ps.evEnd1);
    return FS2D_OK;
}

int fs2d_set_particle_storage_bins(fs2d_handle h, const int32_t *host_bins)
{
    return (h && host_bins) ? particlesSetStorageBins(h, host_bins) : FS2D_ERR_ARG;
}

int fs2d_get_particle_storage_bins(fs2d_handle h, int32_t *host_bins)
{
    return (h && host_bins) ? particlesGetStorageBins(h, host_bins) : FS2D_ERR_ARG;
}

// ---------------------------------------------------------------- PCG
int fs2d_pcg_solve(fs2d_handle ctx, const double *host_rhs, double *host_x, int iter_limit, double tol, int *iters)
{
    if (!ctx || !host_rhs || !host_x || iter_limit < 0) return FS2D_ERR_ARG;
    const size_t bytes = static_cast<size_t>(ctx->N) * sizeof(double);
    FS2D_CUDA(cudaMemcpyAsync(ctx->rhs, host_rhs, bytes, cudaMemcpyHostToDevice, ctx->stream));
    FS2D_TRY(pcgSolveDevice(ctx, iter_limit, tol));
    FS2D_CUDA(fs2dCopyToHost(ctx, host_x, ctx->x, bytes));
    return fs2d_pcg_last_iterations(ctx, iters);
}

int fs2d_pcg_solve_device(fs2d_handle ctx, int iter_limit, double tol)
{
    if (!ctx || iter_limit < 0) return FS2D_ERR_ARG;
    return pcgSolveDevice(ctx, iter_limit, tol);
}

int fs2d_pcg_last_iterations(fs2d_handle ctx, int *iters)
{
    if (!ctx) return FS2D_ERR_ARG;
    PcgScalars sc;
    FS2D_CUDA(fs2dCopyToHost(ctx, &sc, ctx->scalars, sizeof(sc)));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    ctx->lastPcgIters = sc.result;
    if (iters) *iters = sc.result;
    return slabCheckError(ctx);  // slab mode: a barrier of the solve that gave up on a peer (host is synchronised here anyway)
}

int fs2d_pcg_trace(fs2d_handle ctx, double *host_trace, int max_iterations, int *written)
{
    if (!ctx || !host_trace || max_iterations < 0) return FS2D_ERR_ARG;
    PcgScalars sc;
    FS2D_CUDA(fs2dCopyToHost(ctx, &sc, ctx->scalars, sizeof(sc)));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    int n = std::min({sc.iter, max_iterations, ctx->traceCapacity});
    if (n > 0) FS2D_CUDA(fs2dCopyToHost(ctx, host_trace, ctx->trace, sizeof(double) * 4 * n));
    if (written) *written = n;
    return FS2D_OK;
}

// Debug aid, not part of the ABI header: copies the FS2D_MG_DEBUG&8 timeline (1024 x 8 globaltimer stamps) out.
int fs2d_debug_mg_timeline(fs2d_handle ctx, unsigned long long *host_out)
{
    if (!ctx || !host_out || !ctx->mgTimeline) return FS2D_ERR_STATE;
    FS2D_CUDA(fs2dCopyToHost(ctx, host_out, ctx->mgTimeline, 1024 * 8 * sizeof(unsigned long long)));
    return FS2D_OK;
}

int fs2d_pcg_set_dense(fs2d_handle ctx, int dense)
{
    if (!ctx) return FS2D_ERR_ARG;
    ctx->densePcg = dense != 0;
    return FS2D_OK;
}

int fs2d_pcg_active_cells(fs2d_handle ctx, int64_t *cells)
{
    if (!ctx || !cells) return FS2D_ERR_ARG;
    int n = 0;
    FS2D_CUDA(fs2dCopyToHost(ctx, &n, ctx->activeCount, sizeof(n)));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    *cells = static_cast<int64_t>(n) * 16 * 128;
    return FS2D_OK;
}

int fs2d_pcg_profile(fs2d_handle ctx, int enable)
{
    if (!ctx) return FS2D_ERR_ARG;
    ctx->profilePcg = enable != 0;
    ctx->profMs[0] = ctx->profMs[1] = 0.0;
    ctx->profLaunches[0] = ctx->profLaunches[1] = 0;
    ctx->profSolveMs = 0.0;
    ctx->profSolves = 0;
    return FS2D_OK;
}

int fs2d_pcg_profile_solves(fs2d_handle ctx, double *ms, int64_t *solves)
{
    if (!ctx || !ms || !solves) return FS2D_ERR_ARG;
    *ms = ctx->profSolveMs;
    *solves = ctx->profSolves;
    return FS2D_OK;
}

int fs2d_pcg_set_stepwise(fs2d_handle ctx, int stepwise)
{
    if (!ctx) return FS2D_ERR_ARG;
    ctx->stepwisePcg = stepwise != 0;
    return FS2D_OK;
}

int fs2d_pcg_set_grid_limit(fs2d_handle ctx, int max_ctas)
{
    if (!ctx || max_ctas < 0) return FS2D_ERR_ARG;
    ctx->pcgGridLimit = max_ctas;
    return FS2D_OK;
}

int fs2d_kernel_profile(fs2d_handle ctx, int enable)
{
    if (!ctx) return FS2D_ERR_ARG;
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto &it : ctx->kprofPending)
    {
        cudaEventDestroy(it.e0);
        cudaEventDestroy(it.e1);
    }
    ctx->kprofPending.clear();
    for (int g = 0; g < FS2D_KGROUP_COUNT_; g++)
    {
        ctx->kprofMs[g] = 0.0;
        ctx->kprofCalls[g] = 0;
    }
    ctx->kprofOn = enable != 0;
    return FS2D_OK;
}

int fs2d_kernel_profile_read(fs2d_handle ctx, double *ms, int64_t *calls)
{
    if (!ctx || !ms || !calls) return FS2D_ERR_ARG;
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    for (auto &it : ctx->kprofPending)
    {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, it.e0, it.e1) == cudaSuccess && it.group >= 0 && it.group < FS2D_KGROUP_COUNT_)
        {
            ctx->kprofMs[it.group] += t;
            ctx->kprofCalls[it.group]++;
        }
        cudaEventDestroy(it.e0);
        cudaEventDestroy(it.e1);
    }
    ctx->kprofPending.clear();
    for (int g = 0; g < FS2D_KGROUP_COUNT_; g++)
    {
        ms[g] = ctx->kprofMs[g];
        calls[g] = ctx->kprofCalls[g];
    }
    return FS2D_OK;
}

int fs2d_pcg_set_resident(fs2d_handle ctx, int resident)
{
    if (!ctx) return FS2D_ERR_ARG;
    ctx->residentPcg = resident != 0;
    ctx->pagedPcg = resident == 1 && !(std::getenv("FS2D_PCG_PAGED") && std::atoi(std::getenv("FS2D_PCG_PAGED")) == 0);
    return FS2D_OK;
}

int fs2d_pcg_last_kernel(fs2d_handle ctx, int *kind)
{
    if (!ctx || !kind) return FS2D_ERR_ARG;
    int pad = 0;
    FS2D_CUDA(fs2dCopyToHost(ctx, &pad, &ctx->scalars->pad, sizeof(pad)));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    *kind = pad;
    return FS2D_OK;
}

int fs2d_pcg_set_tile_kernels(fs2d_handle ctx, int tile)
{
    if (!ctx) return FS2D_ERR_ARG;
    ctx->forceTileKernels = tile != 0;
    return FS2D_OK;
}

int fs2d_pcg_profile_read(fs2d_handle ctx, double *ms2, int64_t *launches2)
{
    if (!ctx || !ms2 || !launches2) return FS2D_ERR_ARG;
    for (int k = 0; k < 2; k++)
    {
        ms2[k] = ctx->profMs[k];
        launches2[k] = ctx->profLaunches[k];
    }
    return FS2D_OK;
}

int fs2d_spmv(fs2d_handle ctx, const double *host_in, double *host_out)
{
    if (!ctx || !host_in || !host_out) return FS2D_ERR_ARG;
    return pcgSpmvHost(ctx, host_in, host_out, false);
}

int fs2d_precond_apply(fs2d_handle ctx, const double *host_in, double *host_out)
{
    if (!ctx || !host_in || !host_out) return FS2D_ERR_ARG;
    return pcgSpmvHost(ctx, host_in, host_out, true);
}

int fs2d_download_matrix(fs2d_handle ctx, uint8_t *host_is_unit, uint8_t *host_mask, uint8_t *host_count,
                         uint8_t *host_precond_counts)
{
    if (!ctx) return FS2D_ERR_ARG;
    const size_t n = static_cast<size_t>(ctx->N);
    std::vector<uint8_t> row(n);
    std::vector<uint16_t> pre(n);
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    FS2D_CUDA(fs2dCopyToHost(ctx, row.data(), ctx->rowInfo, n));
    FS2D_CUDA(fs2dCopyToHost(ctx, pre.data(), ctx->preInfo, n * 2));
    for (size_t k = 0; k < n; k++)
    {
        if (host_is_unit) host_is_unit[k] = (row[k] & FS2D_ROW_UNIT) ? 1 : 0;
        if (host_mask) host_mask[k] = row[k] & 0xF;
        if (host_count) host_count[k] = (row[k] >> 4) & 7;
        if (host_precond_counts)
            for (int d = 0; d < 4; d++) host_precond_counts[d * n + k] = (pre[k] >> (3 * d)) & 7;
    }
    return FS2D_OK;
}

// ---------------------------------------------------------------- stages
// ---- state dump / restore ("a state dump/restore that doubles as checkpoint"; the reference has none,
// its viewer reads the live solver object: Liquid2dRender/fluidrenderer.cpp:497-986)
namespace
{
struct StateHeader
{
    uint32_t magic, version;
    int32_t I, J, simType, numProperties;
    int64_t records;          // particle records (flagged-dead ones included, see fs2d_download_particles_packed)
    uint64_t gridMask;        // bit g: grid g of the FS2D_GRID_* table follows
    float stepDt;
    int32_t flags;            // bit 0 sdfInsidePending, bit 1 smokeGridsAdvected
    double matrixScale;
    uint64_t reserved[4];
};
constexpr uint32_t STATE_MAGIC = 0x44325346u;  // "FS2D"
size_t stateGridBytes(Ctx *ctx, uint64_t *mask)
{
    size_t total = 0;
    uint64_t m = 0;
    for (int g = 0; g < FS2D_GRID_COUNT_; g++)
    {
        GridDesc d = gridDesc(ctx, g);
        if (!d.ptr || !*d.ptr) continue;
        m |= 1ull << g;
        total += (static_cast<size_t>(d.count) * d.elemSize + 15) & ~static_cast<size_t>(15);
    }
    if (mask) *mask = m;
    return total;
}
}  // namespace

int fs2d_state_bytes(fs2d_handle ctx, size_t *bytes)
{
    if (!ctx || !bytes) return FS2D_ERR_ARG;
    *bytes = sizeof(StateHeader) + stateGridBytes(ctx, nullptr) + fs2d_packed_particle_bytes(ctx, ctx->count);
    return FS2D_OK;
}

int fs2d_state_save(fs2d_handle ctx, void *host_buf, size_t capacity_bytes, size_t *written)
{
    if (!ctx || !host_buf || !written) return FS2D_ERR_ARG;
    if (ctx->slab.enabled && ctx->slab.world > 1)
    {
        ctx->lastError = "fs2d_state_save: not available over row slabs";
        return FS2D_ERR_STATE;
    }
    size_t need = 0;
    FS2D_TRY(fs2d_state_bytes(ctx, &need));
    if (need > capacity_bytes)
    {
        ctx->lastError = "fs2d_state_save: buffer too small";
        return FS2D_ERR_ARG;
    }
    StateHeader h;
    memset(&h, 0, sizeof(h));
    h.magic = STATE_MAGIC;
    h.version = 1;
    h.I = ctx->I;
    h.J = ctx->J;
    h.simType = ctx->p.sim_type;
    h.numProperties = ctx->p.num_properties;
    h.stepDt = ctx->stepDt;
    h.flags = (ctx->sdfInsidePending ? 1 : 0) | (ctx->smokeGridsAdvected ? 2 : 0);
    h.matrixScale = ctx->matrixScale;
    stateGridBytes(ctx, &h.gridMask);
    unsigned char *out = static_cast<unsigned char *>(host_buf) + sizeof(StateHeader);
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    for (int g = 0; g < FS2D_GRID_COUNT_; g++)
    {
        if (!((h.gridMask >> g) & 1ull)) continue;
        GridDesc d = gridDesc(ctx, g);
        const size_t bytes = static_cast<size_t>(d.count) * d.elemSize;
        // the arrays as they are: a deferred level-set walk stays deferred (flag above)
        FS2D_CUDA(cudaMemcpyAsync(out, *d.ptr, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        out += (bytes + 15) & ~static_cast<size_t>(15);
    }
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    int64_t records = 0;
    FS2D_TRY(fs2d_download_particles_packed(ctx, out, capacity_bytes - static_cast<size_t>(out - static_cast<unsigned char *>(host_buf)), &records));
    h.records = records;
    out += fs2d_packed_particle_bytes(ctx, records);
    memcpy(host_buf, &h, sizeof(h));
    *written = static_cast<size_t>(out - static_cast<unsigned char *>(host_buf));
    return FS2D_OK;
}

int fs2d_state_load(fs2d_handle ctx, const void *host_buf, size_t bytes)
{
    if (!ctx || !host_buf || bytes < sizeof(StateHeader)) return FS2D_ERR_ARG;
    if (ctx->slab.enabled && ctx->slab.world > 1)
    {
        ctx->lastError = "fs2d_state_load: not available over row slabs";
        return FS2D_ERR_STATE;
    }
    StateHeader h;
    memcpy(&h, host_buf, sizeof(h));
    uint64_t mask = 0;
    const size_t gridBytes = stateGridBytes(ctx, &mask);
    if (h.magic != STATE_MAGIC || h.version != 1 || h.I != ctx->I || h.J != ctx->J || h.simType != ctx->p.sim_type ||
        h.numProperties != ctx->p.num_properties || h.gridMask != mask || h.records < 0 ||
        bytes < sizeof(StateHeader) + gridBytes + fs2d_packed_particle_bytes(ctx, h.records))
    {
        ctx->lastError = "fs2d_state_load: the blob was not written by a handle with these parameters";
        return FS2D_ERR_ARG;
    }
    const unsigned char *in = static_cast<const unsigned char *>(host_buf) + sizeof(StateHeader);
    for (int g = 0; g < FS2D_GRID_COUNT_; g++)
    {
        if (!((mask >> g) & 1ull)) continue;
        GridDesc d = gridDesc(ctx, g);
        const size_t gb = static_cast<size_t>(d.count) * d.elemSize;
        FS2D_CUDA(cudaMemcpyAsync(*d.ptr, in, gb, cudaMemcpyHostToDevice, ctx->stream));
        in += (gb + 15) & ~static_cast<size_t>(15);
    }
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    FS2D_TRY(fs2d_upload_particles_packed(ctx, in, h.records));
    ctx->stepDt = h.stepDt;
    ctx->matrixScale = h.matrixScale;
    ctx->sdfInsidePending = (h.flags & 1) != 0;
    ctx->smokeGridsAdvected = (h.flags & 2) != 0;
    return FS2D_OK;
}

int fs2d_set_step_dt(fs2d_handle h, float dt)
{
    if (!h) return FS2D_ERR_ARG;
    h->stepDt = dt;
    return FS2D_OK;
}

int fs2d_max_particle_velocity(fs2d_handle h, float *out) { return (h && out) ? particlesMaxVelocity(h, out) : FS2D_ERR_ARG; }

int fs2d_advect(fs2d_handle h)
{
    if (!h) return FS2D_ERR_ARG;
    FS2D_TRY(particlesAdvect(h));
    if (h->p.parameter_handling == FS2D_PARAMS_GRID) FS2D_TRY(gridEulerAdvectParameters(h));  // flipsolver2d.cpp:157-160
    return FS2D_OK;
}

int fs2d_build_matrix(fs2d_handle h) { return h ? gridBuildMatrix(h) : FS2D_ERR_ARG; }
int fs2d_sort_particles(fs2d_handle h) { return h ? particlesRebin(h) : FS2D_ERR_ARG; }
int fs2d_update_density_grid(fs2d_handle h) { return h ? transferDensity(h) : FS2D_ERR_ARG; }
int fs2d_density_rhs(fs2d_handle h) { return h ? gridDensityRhs(h) : FS2D_ERR_ARG; }

int fs2d_density_correction(fs2d_handle h, int *iters)
{
    if (!h) return FS2D_ERR_ARG;
    // densityCorrection (flipsolver2d.cpp:164-186)
    FS2D_TRY(transferDensity(h));
    FS2D_TRY(gridDensityRhs(h));
    FS2D_TRY(pcgSolveDevice(h, h->p.pcg_iter_limit, h->p.project_tolerance));
    int it = 0;
    FS2D_TRY(fs2d_pcg_last_iterations(h, &it));
    if (iters) *iters = it;
    if (it >= h->p.pcg_iter_limit) return FS2D_OK;  // "Density solver solving failed!": result discarded (:179-182)
    FS2D_TRY(slabExchangePressure(h));               // slab mode: the gradient reads p one row outside the slab
    FS2D_TRY(particlesAdjustByDensity(h));
    // The reference leaves adjusted particles in their old bins ("adjusted not enough to require
    // rebinning", flipsolver2d.cpp:427); the cell-sorted layout is re-keyed instead so that the
    // gathers that follow see every particle in the cell its position says.
    FS2D_TRY(particlesRebin(h));
    return FS2D_OK;
}

int fs2d_particle_to_grid(fs2d_handle h)
{
    if (!h) return FS2D_ERR_ARG;
    FS2D_TRY(transferVelocity(h));                                                          // particleVelocityToGrid
    if (h->p.parameter_handling != FS2D_PARAMS_GRID) FS2D_TRY(transferCentered(h));         // flipsolver2d.cpp:1216-1219
    return FS2D_OK;
}

int fs2d_update_sdf(fs2d_handle h) { return h ? transferSdf(h) : FS2D_ERR_ARG; }
int fs2d_update_materials(fs2d_handle h) { return h ? gridUpdateMaterials(h) : FS2D_ERR_ARG; }
int fs2d_after_transfer(fs2d_handle h) { return h ? gridAfterTransfer(h) : FS2D_ERR_ARG; }
int fs2d_extrapolate_velocity(fs2d_handle h, int radius) { return h ? gridExtrapolateVelocity(h, radius) : FS2D_ERR_ARG; }
int fs2d_extrapolate_sdf_inside(fs2d_handle h) { return h ? gridExtrapolateSdf(h, true) : FS2D_ERR_ARG; }
int fs2d_extrapolate_sdf_outside(fs2d_handle h) { return h ? gridExtrapolateSdf(h, false) : FS2D_ERR_ARG; }
int fs2d_save_velocity(fs2d_handle h) { return h ? gridSaveVelocity(h) : FS2D_ERR_ARG; }
int fs2d_apply_body_forces(fs2d_handle h) { return h ? gridBodyForces(h) : FS2D_ERR_ARG; }
int fs2d_pressure_rhs(fs2d_handle h) { return h ? gridPressureRhs(h) : FS2D_ERR_ARG; }
int fs2d_apply_pressure(fs2d_handle h) { return h ? gridApplyPressure(h) : FS2D_ERR_ARG; }

int fs2d_project(fs2d_handle h, int *iters)
{
    if (!h) return FS2D_ERR_ARG;
    // project (flipsolver2d.cpp:93-126)
    FS2D_TRY(gridPressureRhs(h));
    FS2D_TRY(pcgSolveDevice(h, h->p.pcg_iter_limit, h->p.project_tolerance));
    FS2D_TRY(slabExchangePressure(h));  // slab mode: U(i,j) needs p(i-1,j) of the row neighbour
    FS2D_TRY(gridApplyPressure(h));
    if (iters) FS2D_TRY(fs2d_pcg_last_iterations(h, iters));
    return FS2D_OK;
}

int fs2d_velocity_from_solids(fs2d_handle h) { return h ? gridVelocityFromSolids(h) : FS2D_ERR_ARG; }
int fs2d_apply_viscosity(fs2d_handle h, int *iters) { return h ? gridViscosity(h, iters) : FS2D_ERR_ARG; }
int fs2d_particle_update(fs2d_handle h) { return h ? particlesUpdate(h) : FS2D_ERR_ARG; }
int fs2d_count_particles(fs2d_handle h) { return h ? particlesCount(h) : FS2D_ERR_ARG; }
int fs2d_reseed_plan(fs2d_handle h, int64_t *candidates) { return (h && candidates) ? particlesReseedPlan(h, candidates) : FS2D_ERR_ARG; }
int fs2d_reseed_apply(fs2d_handle h, int64_t candidates, const float *host_uniform_xy)
{
    return h ? particlesReseedApply(h, candidates, host_uniform_xy) : FS2D_ERR_ARG;
}

int fs2d_nbflip_advect_grids(fs2d_handle h)
{
    if (!h) return FS2D_ERR_ARG;
    FS2D_TRY(gridNbflipHalo(h));  // row slabs: level set and viscosity on the halo rows
    FS2D_TRY(particlesPruneNarrowBand(h));
    return gridNbflipAdvect(h);
}

int fs2d_set_sdf_band(fs2d_handle ctx, int layers)
{
    if (!ctx || layers < 0) return FS2D_ERR_ARG;
    ctx->sdfBand = layers;
    return FS2D_OK;
}

int fs2d_substep(fs2d_handle h, float dt, float *stage_ms, int *iters)
{
    if (!h) return FS2D_ERR_ARG;
    return stepSubstep(h, dt, stage_ms, iters);
}

}  // extern "C"
