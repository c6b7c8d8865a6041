// Internal state of one fs2d handle: every per-substep array lives on the device.
// Layout notes are in DESIGN.md ("Data layout in HBM").
#ifndef FS2D_INTERNAL_H
#define FS2D_INTERNAL_H

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/fs2d.h"

#define FS2D_CUDA(call)                                                                             \
    do                                                                                              \
    {                                                                                               \
        cudaError_t e__ = (call);                                                                   \
        if (e__ != cudaSuccess)                                                                     \
        {                                                                                           \
            char b__[512];                                                                          \
            snprintf(b__, sizeof(b__), "%s:%d %s -> %s", __FILE__, __LINE__, #call,                 \
                     cudaGetErrorString(e__));                                                      \
            ctx->lastError = b__;                                                                   \
            return FS2D_ERR_CUDA;                                                                   \
        }                                                                                           \
    } while (0)

#define FS2D_TRY(expr)                  \
    do                                  \
    {                                   \
        int rc__ = (expr);              \
        if (rc__ != FS2D_OK) return rc__; \
    } while (0)

// Material predicates (materialgrid.h:15-20).
__host__ __device__ inline bool matFluid(int8_t m) { return (m & 0x40) != 0; }
__host__ __device__ inline bool matStrictFluid(int8_t m) { return m == 0x40; }
__host__ __device__ inline bool matEmpty(int8_t m) { return (m & 0x10) != 0; }
__host__ __device__ inline bool matSolid(int8_t m) { return (m & 0x20) != 0; }
__host__ __device__ inline bool matSource(int8_t m) { return m == 0x41; }
__host__ __device__ inline bool matSink(int8_t m) { return m == 0x12; }

// Per-cell pressure-row byte (replaces IndexedPressureParameterUnit, pressuredata.h:97-151):
// bit 7 = the cell has a matrix row, bits 4-6 = nonsolidNeighborCount (0..4),
// bits 0-3 = fluidNeighborMask (I_NEG 1, I_POS 2, J_NEG 4, J_POS 8).
#define FS2D_ROW_UNIT 0x80u
// Per-cell preconditioner half-word (replaces IndexedIPPCoefficientUnit,
// PressureIPPCoeficients.h:8-48): bit 15 = row exists, then four 3-bit fields with
// nonsolidNeighborCount of the iNeg, iPos, jNeg, jPos neighbour (bits 0-2, 3-5, 6-8, 9-11).
#define FS2D_PRE_UNIT 0x8000u
// Storage-bin code of a particle. The reference files a particle in a 3x3-cell bin and re-files it only
// when ADVECTION moves it to another bin (flipsolver2d.cpp:317-336); the density correction moves
// particles without re-filing them (:427), so a particle can sit in a bin that is not the bin of its
// position. Gathers only see a particle through its storage bin and countParticles only counts
// particles filed in the cell's own bin, so the offset (storage bin - position bin) per axis is kept:
// code = (di+2)*5 + (dj+2); 12 = filed where it is; 255 = further than two bins away.
#define FS2D_MIS_HOME 12
#define FS2D_MIS_LOST 255

struct PcgScalars
{
    double sigma;     // r.z of the current iteration
    double alpha;     // step of the last executed iteration (pending x update)
    double beta;
    double err;
    double gamma;
    int iter;         // iterations completed
    int done;         // 1 once converged / zero rhs
    int result;       // value LinearSolver::solve returns
    int pad;
    unsigned int ticketA;
    unsigned int ticketB;
    unsigned int ticketC;
    unsigned int ticketS;           // whole-solve kernel: monotonic barrier arrivals
    unsigned long long phaseNs[2];  // whole-solve kernel: device time spent in the K1 / K2 phases (globaltimer, CTA 0)
    unsigned int phaseLaunches;     // iterations those times cover
    unsigned int pad3;
};

// ------------------------------------------------------------------ row-slab decomposition (slab.cu)
// One process (or one handle) per GPU owns the cell rows [rowBegin, rowEnd); every dense array keeps the
// GLOBAL size and indexing, a rank computes its own rows plus what it needs of the halo, and ranks talk
// through peer-mapped device memory only: data is PUSHED into the peer's copy of an array and announced
// with a tag the peer spins on in its own memory (1.2 us one way over NVLink, profiles/r1e_ipc_probe*).
#define FS2D_MAX_RANKS 8
#define FS2D_PCG_RING 16

struct SlabPcgSlot
{
    double v0, v1;             // partial dot product, partial max
    unsigned long long tag;    // (solve sequence << 20) + phase + 1, written after the payload
    unsigned long long pad;
};

// Whole-solve kernel: a reduction result travels as four self-validating 8-byte words {32 data bits, 32-bit tag}
// (the layout NCCL's LL protocol uses): 8-byte stores are atomic, so the reader needs no fence between payload and
// flag -- it polls until all four words carry the tag it expects.
struct SlabLLSlot
{
    unsigned long long w[4];   // v0 low, v0 high, v1 low, v1 high
};

struct SlabGatherSlot
{
    long long v[4];
    unsigned long long tag;
    unsigned long long pad[3];
};

struct SlabMail
{
    SlabPcgSlot pcg[FS2D_PCG_RING][FS2D_MAX_RANKS];  // written by every rank (slot [.][writer])
    SlabLLSlot ll[FS2D_PCG_RING][FS2D_MAX_RANKS];    // the same ring for the whole-solve kernel
    unsigned long long ready[2];   // [0] written by the lower neighbour (rank-1), [1] by the upper: "I am done reading
                                   //     what exchange #seq overwrites"
    unsigned long long data[2];    // "my data of exchange #seq has landed in your memory"
    long long counts[2][4];        // particle exchange: records pushed by neighbour [from] {owned, ghosts}
    SlabGatherSlot gather[2][FS2D_MAX_RANKS];      // double buffered by the parity of the gather sequence
    unsigned long long mode[2][FS2D_MAX_RANKS];    // whole-solve kernels: {tag, mode} of rank r for the solve of this parity
                                                   // (1 = resident kernel: talks LL halo rows, 2 = streaming kernel)
    unsigned long long edgeTiles[2][FS2D_MAX_RANKS][2];  // with mode 1: bit tj set = tile tj of rank r's first / last tile row
                                                         // is active (a skipped tile pushes nothing: its halo values are 0)
    int error;                     // set when a spin loop timed out
    int pad[3];
};

struct SlabField                   // one dense array taking part in a halo exchange
{
    unsigned long long offset;     // byte offset inside the symmetric heap
    unsigned long long rowBytes;   // bytes per grid row
};

struct SlabState
{
    bool enabled = false;
    int rank = 0, world = 1, share = 1;
    int rowBegin = 0, rowEnd = 0;          // owned cell rows
    int halo = 32;                         // rows kept valid on either side after a halo exchange
    int ghost = 12;                        // rows of neighbour particles mirrored as ghosts
    unsigned char *peerHeap[FS2D_MAX_RANKS] = {};   // mapped base address of every rank's heap (own: local)
    unsigned char *peerXchg[FS2D_MAX_RANKS] = {};   // mapped particle receive buffers
    bool peerMapped[FS2D_MAX_RANKS] = {};           // opened through cudaIpcOpenMemHandle (to be closed)
    unsigned long long seq = 0;            // generic exchange sequence (same on every rank)
    unsigned long long gatherSeq = 0;
    unsigned long long solveSeq = 0;
    unsigned char *xchg = nullptr;         // local particle exchange buffers: send[2], recv[2]
    size_t xchgBytes = 0;
    int64_t xchgCapacity = 0;              // records per buffer
    int64_t ghostCount = 0;                // ghost records inside the particle arrays (host view)
    int64_t ownedBegin = 0, ownedEnd = 0;  // owned range of the sorted particle arrays
    int connected = 0;
};

struct ParticleBuffers
{
    float2 *pos = nullptr;
    float2 *vel = nullptr;
    float *props = nullptr;      // [numProps][capacity]
    uint32_t *key = nullptr;     // cell key at sort time
    uint8_t *mis = nullptr;      // storage-bin offset code (FS2D_MIS_*), see particles.cu
    int64_t capacity = 0;
};

struct fs2d_context
{
    fs2d_params p;
    int I = 0, J = 0;
    int64_t N = 0, NU = 0, NV = 0;
    int device = 0;
    int smCount = 148;
    cudaStream_t stream = nullptr;
    std::string lastError;
    int64_t launches = 0;
    float stepDt = 0.f;

    // ---- grids
    float *U = nullptr, *V = nullptr, *savedU = nullptr, *savedV = nullptr;
    uint8_t *uValid = nullptr, *vValid = nullptr;
    int8_t *material = nullptr;
    float *fluidSdf = nullptr, *solidSdf = nullptr, *viscosity = nullptr, *density = nullptr;
    int32_t *counts = nullptr, *emitterId = nullptr, *solidId = nullptr;
    float *divergenceControl = nullptr, *testGrid = nullptr;
    uint8_t *knownCentered = nullptr;
    float *temperature = nullptr, *concentration = nullptr, *fuel = nullptr;
    float *sourceSdf = nullptr;
    int32_t *sourceSdfId = nullptr;
    float *advU = nullptr, *advV = nullptr, *advSdf = nullptr, *advViscosity = nullptr;
    float *scratchA = nullptr, *scratchB = nullptr, *scratchC = nullptr;  // N+max(I,J)+1 floats each
    int32_t *markers = nullptr;                                           // BFS layer markers, max(NU,NV)

    // ---- PCG
    double *rhs = nullptr, *x = nullptr, *r[2] = {nullptr, nullptr}, *s[2] = {nullptr, nullptr};
    double *q = nullptr, *z = nullptr;
    uint8_t *rowInfo = nullptr;       // N
    uint16_t *preInfo = nullptr;      // N
    double matrixScale = 0.0;
    double *partials = nullptr;       // reduction scratch (3 * maxBlocks doubles)
    int maxBlocks = 0;
    PcgScalars *scalars = nullptr;    // device
    double *trace = nullptr;          // device, 4 doubles per iteration
    int *p2gTileList = nullptr;       // [0] = count, [1..] = tiles with particles in reach (transfer.cu), allocated on first use
    int32_t *bfsQueue = nullptr;      // level-set BFS: cells in layer order (one slot per cell), allocated on first use
    unsigned int *bfsCtl = nullptr;   // level-set BFS: [0] = queue tail
    void *heavyBuf = nullptr;         // heavy viscosity model: diag, x, r, p, tmp, z over all U and V samples + corner viscosities
    void *viscScalars = nullptr;      // device scalars of the viscosity CG (viscosity.cu), allocated on first use
    unsigned long long *mgTimeline = nullptr;  // debug (FS2D_MG_DEBUG & 8): globaltimer stamps of the slab PCG kernels
    int traceCapacity = 0;
    int64_t *rangeLast = nullptr;     // device, convergence_threads entries
    int lastPcgIters = 0;
    // optional per-kernel timing of the PCG iteration kernels (fs2d_pcg_profile)
    bool densePcg = std::getenv("FS2D_PCG_DENSE") != nullptr;  // walk every tile, not only the active ones
    int *tileFlags = nullptr, *activeTiles = nullptr, *activeCount = nullptr;
    bool forceTileKernels = std::getenv("FS2D_PCG_TILE") != nullptr;  // A/B switch: plain tiled kernels
    bool stepwisePcg = std::getenv("FS2D_PCG_STEPWISE") != nullptr;  // A/B switch: two kernels per iteration instead of the whole-solve kernel
    bool rowCopyPcg = std::getenv("FS2D_PCG_ROWCOPY") != nullptr;    // A/B switch: per-row bulk copies instead of tensor copies
    // test knob (fs2d_pcg_set_grid_limit / FS2D_PCG_GRID): cap on the CTAs of the persistent PCG grids, so that a small
    // grid already gives every CTA several tiles to walk (the pipelined k >= 1 path the 4096^2 runs live on); 0 = no cap
    int pcgGridLimit = std::getenv("FS2D_PCG_GRID") ? std::atoi(std::getenv("FS2D_PCG_GRID")) : 0;
    bool residentPcg = !(std::getenv("FS2D_PCG_RESIDENT") && std::atoi(std::getenv("FS2D_PCG_RESIDENT")) == 0);  // A/B switch (fs2d_pcg_set_resident)
    bool pagedPcg = !(std::getenv("FS2D_PCG_PAGED") && std::atoi(std::getenv("FS2D_PCG_PAGED")) == 0);          // resident + paged tiles (pcgResidentKernel<false, true>)
    int pcgOccupancy = 0;                 // CTAs of pcgSolveKernel one SM holds (occupancy query, cached)
    void *solveMaps = nullptr;            // host copy of the tensor maps of the Krylov vectors (pcg.cu)
    bool solveMapsTried = false, solveMapsOk = false;
    bool profilePcg = false;
    double profSolveMs = 0.0;             // whole-solve kernel: accumulated launch durations and their number
    int64_t profSolves = 0;
    std::vector<cudaEvent_t> profEvents;
    double profMs[2] = {0.0, 0.0};        // accumulated device time of K1 / K2 launches
    int64_t profLaunches[2] = {0, 0};

    // ---- particles (double buffered for the sort)
    ParticleBuffers pb[2];
    int cur = 0;
    int64_t count = 0;
    uint8_t *dead = nullptr;          // capacity bytes
    uint32_t *perm = nullptr;         // capacity
    int32_t *cellStart = nullptr;     // N+1, valid after fs2d_sort_particles
    int32_t *cellCursor = nullptr;    // N
    int32_t *scanBlock = nullptr;     // scan scratch
    bool sorted = false;
    int64_t deadCount = 0;            // particles flagged dead since the last sort (host view)
    bool killedDirty = false;         // d_counter[0] holds kills not yet folded into deadCount
    bool sdfInsidePending = false;    // extrapolateLevelsetInside deferred until the grid is read (grid_ops.cu)
    bool eagerSdf = std::getenv("FS2D_EAGER_SDF") != nullptr;
    int sdfBand = std::getenv("FS2D_SDF_BAND") ? std::atoi(std::getenv("FS2D_SDF_BAND")) : 24;  // NBFlip level-set walks: layers (0 = unbounded), grid_ops.cu
    bool smokeGridsAdvected = false;  // temperature/concentration/fuel replaced by advected grids (App. A-13)
    int64_t *d_counter = nullptr;     // device scalar scratch (8 int64)
    float *d_fscratch = nullptr;      // device float scratch

    unsigned char *stage = nullptr;   // device staging buffer of the packed particle transfers (capi.cu)
    size_t stageBytes = 0;

    // ---- streamed particle state (fs2d_particle_stream_*, capi.cu): the sections of the host buffer travel on a copy
    // stream of their own and the solver's stream waits for each section where it is first needed
    struct ParticleStream
    {
        cudaStream_t copy = nullptr;
        cudaEvent_t evByte = nullptr, evVel = nullptr, evPos = nullptr, evProps = nullptr, evMain = nullptr;
        cudaEvent_t evStart = nullptr, evEarly0 = nullptr, evEarly1 = nullptr, evEnd0 = nullptr, evEnd1 = nullptr;  // timing only
        bool timedUpload = false, timedEarly = false, timedEnd = false;
        bool posPending = false;          // positions (and cell keys) not yet waited for / derived
        static constexpr int POS_CHUNKS = 4;   // the position section travels in chunks so that advection can follow it
        cudaEvent_t evPosChunk[POS_CHUNKS] = {};
        int64_t posChunkEnd[POS_CHUNKS] = {};
        bool propsPending = false;        // property columns not yet waited for
        bool propsGatherPending = false;  // a sort ran meanwhile: the columns still lie in buffer `propsFrom`, unsorted
        int propsFrom = 0;
        int64_t gatherCount = 0;
        int64_t earlyCount = -1;          // records whose positions already left for the host (-1: none)
        int64_t earlyVelCount = -1;       // records whose velocities already left for the host (-1: none)
        void *earlyVelHost = nullptr;
        int64_t earlyVelCapacity = 0;
        bool earlyProps = false;          // ... and whose property columns did
        void *earlyHost = nullptr;
        int64_t earlyCapacity = 0;
        void *outHost = nullptr;          // fs2d_particle_stream_set_output: where the composite fs2d_substep sends early sections
        int64_t outCapacity = 0;
    } pstream;

    // ---- scene tables
    float *obstacleFriction = nullptr;
    int numObstacles = 0;
    bool obstaclesFrictionless = true;  // every friction coefficient is 0: updateVelocityFromSolids multiplies by exactly 1
    fs2d_source *sources = nullptr;
    int numSources = 0;
    std::vector<fs2d_source> hostSources;

    // ---- reseed plan
    int32_t *reseedOffset = nullptr;  // N+1 exclusive scan of per-cell candidate counts
    int64_t reseedCandidates = 0;
    float *reseedUniform = nullptr;   // device copy of the host-drawn jitter
    int64_t reseedUniformCapacity = 0;

    // ---- symmetric heap: every dense array above is carved from ONE allocation so that a peer can map it
    // with a single IPC handle and address the same array at the same offset
    unsigned char *heap = nullptr;
    size_t heapBytes = 0;
    SlabMail *mail = nullptr;         // inside the heap
    unsigned long long *haloLL = nullptr;  // inside the heap: halo rows of q / z received as self-validating words,
                                           // [from lower / upper neighbour][q / z][J][lo, hi] (pcgResidentKernel)
    SlabState slab;

    // ---- per-group kernel timing (fs2d_kernel_profile)
    struct KprofItem { int group; cudaEvent_t e0, e1; };
    bool kprofOn = false;
    std::vector<KprofItem> kprofPending;
    double kprofMs[FS2D_KGROUP_COUNT_] = {};
    int64_t kprofCalls[FS2D_KGROUP_COUNT_] = {};

    // ---- timing
    cudaEvent_t ev[16];
    bool eventsReady = false;
};

typedef fs2d_context Ctx;

// Measurement aid (fs2d_kernel_profile): CUDA events on the handle's stream around the transfer-kernel groups of
// SURVEY 8(d), so that bench.py can state a roofline fraction for each from live timings.
struct KernelGroupTimer
{
    Ctx *ctx;
    int group;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    KernelGroupTimer(Ctx *c, int g) : ctx(c), group(g)
    {
        if (!ctx->kprofOn) return;
        if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) { e0 = e1 = nullptr; return; }
        cudaEventRecord(e0, ctx->stream);
    }
    ~KernelGroupTimer()
    {
        if (!e0 || !e1) return;
        cudaEventRecord(e1, ctx->stream);
        ctx->kprofPending.push_back({group, e0, e1});
    }
};

inline int divUp(int64_t a, int64_t b) { return static_cast<int>((a + b - 1) / b); }

// pcg.cu
int pcgSolveDevice(Ctx *ctx, int iterLimit, double tol);
int pcgSpmvHost(Ctx *ctx, const double *in, double *out, bool precond);
// particles.cu
int particlesReserve(Ctx *ctx, int64_t capacity);
int particlesKeyRange(Ctx *ctx, int64_t begin, int64_t end);
int particlesMaxVelocity(Ctx *ctx, float *out);
int particlesAdvect(Ctx *ctx);
int particlesSort(Ctx *ctx);
int particlesRebin(Ctx *ctx);
int particlesUpdate(Ctx *ctx);
int particlesCount(Ctx *ctx);
int particlesAdjustByDensity(Ctx *ctx);
int particlesReseedPlan(Ctx *ctx, int64_t *candidates);
int particlesReseedApply(Ctx *ctx, int64_t candidates, const float *hostUniform);
int particlesPruneNarrowBand(Ctx *ctx);
int particlesAliveCount(Ctx *ctx, int64_t *out);
int particlesSetStorageBins(Ctx *ctx, const int32_t *hostBins);
int particlesGetStorageBins(Ctx *ctx, int32_t *hostBins);
// Streamed uploads (capi.cu): make the solver's stream wait for the position section (SettlePos) or for every section
// (SettleAll, which also runs the property gather a sort had to postpone). No-ops unless an upload is in flight.
int particleStreamSettleSlow(Ctx *ctx, bool all);
inline int particleStreamSettlePos(Ctx *ctx)
{
    return (ctx->pstream.posPending) ? particleStreamSettleSlow(ctx, false) : FS2D_OK;
}
inline int particleStreamSettleAll(Ctx *ctx)
{
    return (ctx->pstream.posPending || ctx->pstream.propsPending || ctx->pstream.propsGatherPending) ? particleStreamSettleSlow(ctx, true) : FS2D_OK;
}
inline void particleStreamPositionsChanged(Ctx *ctx) { ctx->pstream.earlyCount = ctx->pstream.earlyVelCount = -1; }
int particlesGatherProps(Ctx *ctx, int from, int64_t count);  // particles.cu
// Streamed upload, advection only: makes the solver's stream wait for chunk `idx` of the position section and returns its
// record range; false when there is no such chunk. After the last chunk the caller calls particleStreamSettlePos.
bool particleStreamNextPosChunk(Ctx *ctx, int idx, int64_t *begin, int64_t *end);
// transfer.cu
int transferVelocity(Ctx *ctx);
int transferCentered(Ctx *ctx);
int transferDensity(Ctx *ctx);
int transferSdf(Ctx *ctx);
// grid_ops.cu
int gridBuildMatrix(Ctx *ctx);
int gridUpdateMaterials(Ctx *ctx);
int gridAfterTransfer(Ctx *ctx);
int gridExtrapolateVelocity(Ctx *ctx, int radius);
int gridExtrapolateSdf(Ctx *ctx, bool inside);
int gridFlushSdf(Ctx *ctx);
int gridFlushSdfGathered(Ctx *ctx);   // after every rank pushed its rows of the level set to every other rank
int gridSdfForRead(Ctx *ctx, const float **field);  // the level set a host reader sees (deferred / banded walks completed)
int gridNbflipHalo(Ctx *ctx);
int gridSaveVelocity(Ctx *ctx);
int gridBodyForces(Ctx *ctx);
int gridPressureRhs(Ctx *ctx);
int gridDensityRhs(Ctx *ctx);
int gridApplyPressure(Ctx *ctx);
int gridVelocityFromSolids(Ctx *ctx);
int gridEulerAdvectParameters(Ctx *ctx);
int gridNbflipAdvect(Ctx *ctx);
int gridViscosity(Ctx *ctx, int *iters);
// Device -> host copy into pageable memory: wait for the stream FIRST. cudaMemcpyAsync towards pageable memory blocks
// inside the driver until the preceding work of the stream has finished, and other host threads cannot launch
// meanwhile -- fatal when the preceding work is a kernel waiting for a launch of another rank that shares this
// process (slab tests run several ranks on one GPU). cudaStreamSynchronize waits without that side effect.
inline cudaError_t fs2dCopyToHost(Ctx *ctx, void *dst, const void *src, size_t bytes)
{
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, ctx->stream);
    if (e != cudaSuccess) return e;
    return cudaStreamSynchronize(ctx->stream);
}

// slab.cu
struct SlabRows { int lo, hi; };                       // half-open cell-row range
SlabRows slabOwn(const Ctx *ctx);                      // owned rows (whole grid without slabs)
SlabRows slabExt(const Ctx *ctx, int k);               // owned rows grown by k, clipped to the grid
int slabExchangeFields(Ctx *ctx, const void *const *arrays, const size_t *rowBytes, int count);
int slabExchangeVelocity(Ctx *ctx, bool withMaterial);
int slabExchangePressure(Ctx *ctx);
int slabExchangeParticles(Ctx *ctx);
int slabAllGather(Ctx *ctx, const long long v[4], long long *out /* world x 4 */);
int slabCheckError(Ctx *ctx);
int slabGatherRows(Ctx *ctx, void *array, size_t rowBytes, int rowsTotal);  // collective: own rows -> every rank
int slabGatherMany(Ctx *ctx, void *const *arrays, const size_t *rowBytes, const int *rowsTotal, int count);
// The same for a band of rows only: every rank contributes the first / last row it finds interesting among its own
// (localMax < localMin: none), the union over the ranks grown by `margin` rows is what travels; *rowLo / *rowHi return it
// (half open; rowHi <= rowLo: no rank found anything, nothing was sent).
int slabGatherBand(Ctx *ctx, void *const *arrays, const size_t *rowBytes, const int *rowsTotal, int count, int localMin, int localMax,
                   int margin, int *rowLo, int *rowHi);
void pcgPreloadSlabKernels();           // pcg.cu
// step.cu
int stepSubstep(Ctx *ctx, float dt, float *stageMs, int *iters);

#endif
