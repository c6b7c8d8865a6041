// Kernel group 4: the pressure-projection PCG.
//
// Replaces LinearSolver::solve (linearsolver.cpp:25-73), IndexedPressureParameters::multiply
// (pressuredata.h:132-238), IndexedIPPCoefficients::multiply (PressureIPPCoeficients.h:23-132)
// and the VOps BLAS-1 passes (vmath.cpp:25-136) with two fused, bandwidth-bound phases per
// iteration and scalars that never leave the device:
//
//   K1(i):  s_i = z + beta*s_{i-1}          (addMul, linearsolver.cpp:68)
//           x  += alpha_{i-1}*s_{i-1}       (deferred addMul of :51 -- s_{i-1} is in flight anyway)
//           q   = A*s_i                     (:49)      gamma = q.s_i -> alpha_i (:50)
//   K2(i):  r  -= alpha_i*q                 (:52)
//           z   = M*r   sigma' = z.r        (:63-66)   err = max|r| (:53)
//           convergence / beta (:59-69)
//
// Per iteration each phase streams every vector it touches exactly once: K1 reads z, s, x
// and writes s, q, x; K2 reads r, q and writes r, z -> 10 fp64 passes + 3 bytes of per-cell
// row info = 83 B/cell (the reference makes 18 passes). Stencil neighbours come from a
// shared-memory tile of the *derived* vector (s_i resp. the updated r), so the 5-point
// operators never re-read HBM. Vectors stay fp64 like the reference's std::vector<double>
// (linearsolver.h:15-20); arithmetic uses explicit non-contracted mul/add in the reference's
// evaluation order so a single operator application is bit-identical to the strict oracle.
//
// Three generations of the same arithmetic live here (all produce bit-identical iterates):
//   pcgSolveKernel   the default: ONE persistent cooperative launch per solve, tensor-map TMA tile loads, the two
//                    reductions of an iteration carried by in-kernel grid barriers (which, with row slabs, ARE the
//                    all-reduce across GPUs) -- see "whole-solve kernel" below;
//   pcgPipeKernel    two launches per iteration with a bulk-copy pipeline; still used for the reference-compatible
//                    convergence test (convergence_threads > 0) and as the A/B baseline (fs2d_pcg_set_stepwise);
//   pcgTileKernel    plain tiles, for odd gridSizeJ and the stand-alone operator applications.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "fs2d_internal.h"

namespace
{
constexpr int TR = 16;        // tile rows
constexpr int TC = 128;       // tile columns
constexpr int NT = 256;       // threads per CTA
constexpr int SW = TC + 2;    // shared row stride (doubles)

enum Mode { MODE_K1 = 0, MODE_K2 = 1, MODE_APPLY_A = 2, MODE_APPLY_M = 3 };

struct PcgArgs
{
    int I, J;
    long long N;
    int tilesJ;
    double scale;
    double pre[8];            // 1/(k*scale), k = 0..4 (getIPPCoefficients, flipsolver2d.cpp:924-927)
    const uint8_t *rowInfo;
    const uint16_t *preInfo;
    const double *in0;        // K1: z      K2: r_in   APPLY: input vector
    const double *in1;        // K1: s_old  K2: q
    double *out0;             // K1: s_new  K2: r_out
    double *out1;             // K1: q      K2: z      APPLY: output vector
    double *x;                // K1 only
    double *partials;
    PcgScalars *sc;
    double *trace;
    int traceCapacity;
    double tol;
    int compat;               // 1: convergence decided by the range kernel
    const int *activeTiles;   // ordered list of tiles that can hold non-zero entries (nullptr: all tiles)
    const int *activeCount;
};

// Slab mode (several GPUs, slab.cu): the two reductions of an iteration become all-reduces and the operators need
// one halo row from each row neighbour. Both ride on the iteration kernels themselves:
//   * the last CTA of a kernel PUSHES its rank's partial sums into a mail slot of every rank (peer-mapped
//     store + tag), and every CTA of the NEXT kernel starts by waiting for the tags in its own memory and adding
//     the partials in rank order -- every CTA of every rank derives bit-identical alpha / beta / err;
//   * the CTAs that own a slab-boundary row store their q (K1) or z (K2) values of that row straight into the
//     neighbour's array as well; the tag that carries the partial sum also publishes those rows.
// The derived vectors s and r are kept up to date on the halo rows redundantly (same inputs, same beta/alpha
// -> same bits as on the owner). Phase p of a solve: 0 = init, 2i+1 = K1(i), 2i+2 = K2(i); slot (p & 7) of the
// half of the ring selected by the parity of the solve.
struct MgArgs
{
    int rank, world;
    int rowBegin, rowEnd;                 // owned cell rows
    int tileBase;                         // first owned tile (dense walk)
    int phase;                            // phase this kernel PRODUCES
    int iter;                             // iteration index i
    int ringBase;                         // 0 or 8
    unsigned long long solveTag;          // solve sequence << 20
    SlabMail *mail;
    SlabMail *peerMail[FS2D_MAX_RANKS];
    double *loOut1, *hiOut1;              // the row neighbours' copies of out1 (nullptr at the domain ends)
    int iterLimit;
    unsigned long long *haloLL, *loHaloLL, *hiHaloLL;  // LL halo-row buffers: own, lower neighbour's, upper neighbour's
    unsigned long long *timeline;         // debug: 8 globaltimer stamps per phase (nullptr normally)
    int debug;                            // FS2D_MG_DEBUG bit mask (timing experiments only; results are wrong when set):
                                          // 1 = do not wait for the peers' partials, 2 = no halo-row stores into the peers,
                                          // 4 = device-scope instead of system-scope fence in every CTA
};

constexpr long long MG_SPIN_LIMIT = 8000000000ll;

// Sum (v0) and max (v1) over ranks of the partials published for `phase`; spins until every rank's tag is there.
__device__ __forceinline__ bool mgCollect(const MgArgs &m, int phase, bool wait, double *sum, double *mx)
{
    const unsigned long long want = m.solveTag + static_cast<unsigned long long>(phase) + 1ull;
    double s = 0.0, x = 0.0;
    for (int r = 0; r < m.world; r++)
    {
        const SlabPcgSlot *slot = &m.mail->pcg[m.ringBase + (phase & 7)][r];
        if (wait && !((m.debug & 1) && phase > 0))
        {
            const volatile unsigned long long *tag = &slot->tag;
            const long long t0 = clock64();
            while (*tag != want)
            {
                if (clock64() - t0 > MG_SPIN_LIMIT)
                {
                    m.mail->error = 1;
                    return false;
                }
            }
            __threadfence_system();
        }
        const double v0 = *reinterpret_cast<const volatile double *>(&slot->v0);
        const double v1 = *reinterpret_cast<const volatile double *>(&slot->v1);
        s += v0;
        x = fmax(x, v1);
    }
    *sum = s;
    *mx = x;
    return true;
}

// Publish this rank's partials for m.phase in every rank's mail (threads 0..world-1 of the last CTA).
__device__ __forceinline__ void mgPublishPhase(const MgArgs &m, int phase, double v0, double v1)
{
    const int r = threadIdx.x;
    if (r < m.world)
    {
        SlabPcgSlot *slot = &m.peerMail[r]->pcg[m.ringBase + (phase & 7)][m.rank];
        *reinterpret_cast<volatile double *>(&slot->v0) = v0;
        *reinterpret_cast<volatile double *>(&slot->v1) = v1;
        __threadfence_system();
        *reinterpret_cast<volatile unsigned long long *>(&slot->tag) = m.solveTag + static_cast<unsigned long long>(phase) + 1ull;
    }
}

__device__ __forceinline__ void mgPublish(const MgArgs &m, double v0, double v1) { mgPublishPhase(m, m.phase, v0, v1); }

__device__ __forceinline__ unsigned long long globalTimerNs()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}

__device__ __forceinline__ void mgStamp(const MgArgs *m, int point)
{
    if (m && m->timeline) m->timeline[(m->phase & 1023) * 8 + point] = globalTimerNs();
}

__device__ __forceinline__ void mgStampAt(const MgArgs &m, int phase, int point)
{
    if (m.timeline) m.timeline[(phase & 1023) * 8 + point] = globalTimerNs();
}

struct MgScalars
{
    double coef, alphaPrev;
    int done;
};

__device__ __forceinline__ double warpSum(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ double warpMax(double v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_down_sync(0xffffffffu, v, o));
    return v;
}

// Sum (or max) over the CTA; result valid in thread 0.
template <bool IS_MAX> __device__ double blockReduce(double v, double *scratch /* >= 8 doubles */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = IS_MAX ? warpMax(v) : warpSum(v);
    __syncthreads();
    if (lane == 0) scratch[warp] = v;
    __syncthreads();
    if (warp == 0)
    {
        v = (lane < (blockDim.x >> 5)) ? scratch[lane] : (IS_MAX ? 0.0 : 0.0);
        v = IS_MAX ? warpMax(v) : warpSum(v);
    }
    return v;
}

// Deterministic final reduction by the last CTA: fixed assignment of partials to threads.
template <bool IS_MAX> __device__ double finalReduce(const double *partials, int count, double *scratch)
{
    double v = 0.0;
    for (int k = threadIdx.x; k < count; k += blockDim.x)
    {
        double p = __ldcg(partials + k);
        v = IS_MAX ? fmax(v, p) : v + p;
    }
    return blockReduce<IS_MAX>(v, scratch);
}

// One row of A (IndexedPressureParameterUnit::multiply, pressuredata.h:132-146): centre,
// iNeg, iPos, jNeg, jPos in that order, products and sums individually rounded.
__device__ __forceinline__ double rowA(uint8_t info, double scale, double c, double im, double ip, double jm, double jp)
{
    if (!(info & FS2D_ROW_UNIT)) return c;  // identity row (pressuredata.h:227-236)
    // The reference multiplies by static_cast<double>(count) and by the 0 / 1 neighbour flags. Integer -> double
    // conversions run on the XU pipe (16 lanes per SM; ncu showed it saturated: five conversions per cell), so the same
    // values are SELECTED instead: double(k) for k = 0 .. 4 is one of five constants, and -scale * double(bit) is
    // -scale (bit 1) or -scale * 0.0 (bit 0) -- identical bits, no conversion.
    const unsigned int k = (info >> 4) & 7u;
    const double cnt = k < 2u ? (k ? 1.0 : 0.0) : (k == 2u ? 2.0 : (k == 3u ? 3.0 : 4.0));
    const double ns = -scale, nz = __dmul_rn(ns, 0.0);
    double acc = __dmul_rn(__dmul_rn(scale, cnt), c);
    acc = __dadd_rn(acc, __dmul_rn((info & 1u) ? ns : nz, im));
    acc = __dadd_rn(acc, __dmul_rn((info & 2u) ? ns : nz, ip));
    acc = __dadd_rn(acc, __dmul_rn((info & 4u) ? ns : nz, jm));
    acc = __dadd_rn(acc, __dmul_rn((info & 8u) ? ns : nz, jp));
    return acc;
}

// One row of M (IndexedIPPCoefficientUnit::multiply, PressureIPPCoeficients.h:23-40): centre,
// jNeg, jPos, iNeg, iPos in that order.
__device__ __forceinline__ double rowM(uint16_t info, const double *pre, double c, double im, double ip, double jm, double jp)
{
    if (!(info & FS2D_PRE_UNIT)) return c;
    const double iNeg = pre[info & 7u], iPos = pre[(info >> 3) & 7u];
    const double jNeg = pre[(info >> 6) & 7u], jPos = pre[(info >> 9) & 7u];
    const double diag = __dadd_rn(__dadd_rn(1.0, __dmul_rn(iNeg, iNeg)), __dmul_rn(jNeg, jNeg));
    double acc = __dmul_rn(diag, c);
    acc = __dadd_rn(acc, __dmul_rn(jNeg, jm));
    acc = __dadd_rn(acc, __dmul_rn(jPos, jp));
    acc = __dadd_rn(acc, __dmul_rn(iNeg, im));
    acc = __dadd_rn(acc, __dmul_rn(iPos, ip));
    return acc;
}


// Per-CTA partial sums -> the last CTA to arrive reduces them in a fixed order and updates the
// device-resident scalars (alpha after K1; sigma', err, convergence decision and beta after K2).
template <int MODE>
__device__ void finishReductions(const PcgArgs &a, double accDot, double accMax, double *red, int *isLastShared, const MgArgs *mg = nullptr)
{
    const int tid = threadIdx.x;
    const int nb = gridDim.x;
    double bs = blockReduce<false>(accDot, red);
    double bm = 0.0;
    if (MODE == MODE_K2) bm = blockReduce<true>(accMax, red);
    if (tid == 0)
    {
        a.partials[blockIdx.x] = bs;
        if (MODE == MODE_K2) a.partials[nb + blockIdx.x] = bm;
        if (mg && !(mg->debug & 4))
            __threadfence_system();  // also orders this CTA's halo-row stores into the neighbours' arrays
        else
            __threadfence();
        unsigned int *ticket = (MODE == MODE_K1) ? &a.sc->ticketA : &a.sc->ticketB;
        *isLastShared = (atomicAdd(ticket, 1u) == static_cast<unsigned int>(nb - 1));
    }
    __syncthreads();
    if (!*isLastShared) return;
    if (tid == 0) mgStamp(mg, 3);
    __threadfence();
    double total = finalReduce<false>(a.partials, nb, red);
    double emax = 0.0;
    if (MODE == MODE_K2) emax = finalReduce<true>(a.partials + nb, nb, red);
    if (mg)
    {
        // slab mode: the scalars are derived by the consumers (mgPrologue); only publish the partials
        __shared__ double pub[2];
        if (tid == 0)
        {
            pub[0] = total;
            pub[1] = emax;
            if (MODE == MODE_K1)
                a.sc->ticketA = 0;
            else
                a.sc->ticketB = 0;
        }
        __syncthreads();
        if (tid == 0) mgStamp(mg, 4);
        mgPublish(*mg, pub[0], pub[1]);
        if (tid == 0) mgStamp(mg, 5);
        return;
    }
    if (tid == 0)
    {
        PcgScalars *sc = a.sc;
        if (MODE == MODE_K1)
        {
            sc->gamma = total;
            sc->alpha = sc->sigma / (total + 1e-8);  // linearsolver.cpp:50
            sc->ticketA = 0;
        }
        else
        {
            sc->ticketB = 0;
            sc->gamma = total;  // sigma' parked until the decision
            sc->err = emax;
            if (!a.compat)
            {
                const int it = sc->iter;
                double beta = 0.0;
                if (emax <= a.tol)  // linearsolver.cpp:59-61
                {
                    sc->done = 1;
                    sc->result = it;
                }
                else
                {
                    beta = total / sc->sigma;  // :66-67
                    sc->beta = beta;
                    sc->sigma = total;
                }
                if (a.trace && it < a.traceCapacity)
                {
                    a.trace[4 * it + 0] = sc->alpha;
                    a.trace[4 * it + 1] = beta;
                    a.trace[4 * it + 2] = total;
                    a.trace[4 * it + 3] = emax;
                }
                sc->iter = it + 1;
            }
        }
    }
}

template <int MODE> __global__ void __launch_bounds__(NT) pcgTileKernel(PcgArgs a)
{
    __shared__ double tile[(TR + 2) * SW];
    __shared__ double red[8];
    __shared__ double preTbl[8];
    __shared__ int isLast;

    if (MODE == MODE_K1 || MODE == MODE_K2)
    {
        if (a.sc->done) return;
    }
    const int tid = threadIdx.x;
    if (MODE == MODE_K2 || MODE == MODE_APPLY_M)
    {
        if (tid < 8) preTbl[tid] = a.pre[tid];
    }

    const int ti = blockIdx.x / a.tilesJ, tj = blockIdx.x - ti * a.tilesJ;
    const int i0 = ti * TR, j0 = tj * TC;
    const long long J = a.J, N = a.N;

    double coef = 0.0, alphaPrev = 0.0;
    if (MODE == MODE_K1)
    {
        coef = a.sc->beta;
        alphaPrev = a.sc->alpha;
    }
    else if (MODE == MODE_K2)
    {
        coef = a.sc->alpha;
    }

    // ---- load phase: derived vector into shared memory, by LINEAR index so that the
    // j = 0 / J-1 neighbours wrap to the adjacent row exactly as the reference's
    // linearIdxOfOffset(idx, 0, +-1) does (pressuredata.h:135-145).
    // main body: (TR+2) rows x TC columns, each warp reads 32 consecutive doubles
    for (int e = tid; e < (TR + 2) * TC; e += NT)
    {
        const int ar = e / TC, bc = e - ar * TC;  // ar 0..TR+1 (halo rows 0 and TR+1), bc 0..TC-1
        const long long gi = i0 - 1 + ar, gj = j0 + bc;
        const long long n = gi * J + gj;
        double v = 0.0;
        if (n >= 0 && n < N)
        {
            if (MODE == MODE_K1)
            {
                const double zv = a.in0[n], so = a.in1[n];
                v = __dadd_rn(zv, __dmul_rn(so, coef));
                if (ar >= 1 && ar <= TR && gi < a.I && gj < J)
                {
                    a.out0[n] = v;
                    a.x[n] = __dadd_rn(a.x[n], __dmul_rn(so, alphaPrev));
                }
            }
            else if (MODE == MODE_K2)
            {
                v = __dsub_rn(a.in0[n], __dmul_rn(a.in1[n], coef));
                if (ar >= 1 && ar <= TR && gi < a.I && gj < J) a.out0[n] = v;
            }
            else
            {
                v = a.in0[n];
            }
        }
        tile[ar * SW + bc + 1] = v;
    }
    // halo columns (left j0-1, right j0+TC) for the TR interior rows
    if (tid < 2 * TR)
    {
        const int side = tid / TR, ar = 1 + (tid - side * TR);
        const long long gi = i0 - 1 + ar, gj = side ? (j0 + TC) : (j0 - 1);
        const long long n = gi * J + gj;
        double v = 0.0;
        if (n >= 0 && n < N)
        {
            if (MODE == MODE_K1)
                v = __dadd_rn(a.in0[n], __dmul_rn(a.in1[n], coef));
            else if (MODE == MODE_K2)
                v = __dsub_rn(a.in0[n], __dmul_rn(a.in1[n], coef));
            else
                v = a.in0[n];
        }
        tile[ar * SW + (side ? TC + 1 : 0)] = v;
    }
    __syncthreads();

    // ---- stencil phase: thread owns one column and marches TR/2 rows
    const int bc = tid % TC, rg = tid / TC;  // rg 0..1
    const long long gj = j0 + bc;
    double accDot = 0.0, accMax = 0.0;
    if (gj < J)
    {
#pragma unroll 4
        for (int k = 0; k < TR / 2; k++)
        {
            const int ar = 1 + rg * (TR / 2) + k;
            const long long gi = i0 - 1 + ar;
            if (gi >= a.I) break;
            const long long n = gi * J + gj;
            const double *t = tile + ar * SW + bc + 1;
            const double c = t[0], im = t[-SW], ip = t[SW], jm = t[-1], jp = t[1];
            double o;
            if (MODE == MODE_K1 || MODE == MODE_APPLY_A)
                o = rowA(a.rowInfo[n], a.scale, c, im, ip, jm, jp);
            else
                o = rowM(a.preInfo[n], preTbl, c, im, ip, jm, jp);
            a.out1[n] = o;
            accDot += o * c;
            accMax = fmax(accMax, fabs(c));
        }
    }
    if (MODE == MODE_APPLY_A || MODE == MODE_APPLY_M) return;

    // ---- reductions: per-CTA partials, finished by the last CTA in a fixed order
    finishReductions<MODE>(a, accDot, accMax, red, &isLast);
}


// ------------------------------------------------------------------ pipelined iteration kernels
// Same arithmetic as pcgTileKernel<K1/K2>, restructured for HBM throughput: a persistent grid (2 CTAs
// per SM) walks the tiles; the raw input rows of the NEXT tile are fetched by the bulk-copy engine
// (cp.async.bulk global->shared, completion on an mbarrier) while the threads compute the current one,
// so bytes stay in flight during the stencil / store phases instead of only during a load phase.
// Rows are fetched by LINEAR index (segment [gi*J + j0 - 2, +132) clipped to [0, N)), which keeps the
// reference's wrap of the j = 0 / J-1 neighbours (pressuredata.h:135-145); needs J even for the 16-byte
// alignment of the bulk copies (the host falls back to pcgTileKernel otherwise).
constexpr int PSW = TC + 4;               // staged row: 2 pad + TC + 2 pad doubles (1056 B)
constexpr int PROWS = TR + 2;
constexpr int PTILE = PROWS * PSW;        // 2376 doubles per staged array
constexpr int PTILE_PAD = (PTILE + 15) / 16 * 16;  // arrays padded to 128 bytes: tensor copies need 128-byte aligned targets
constexpr int PSTAGES = 2;

template <int MODE> struct PipeStage
{
    double a[PTILE_PAD];                  // K1: z -> s_new (in place)    K2: r -> r_new (in place)
    double b[PTILE_PAD];                  // K1: s_old                    K2: q
    double x[MODE == MODE_K1 ? TR * TC : 2];  // K1: x tile
};

template <int MODE> struct PipeSmem
{
    PipeStage<MODE> st[PSTAGES];
    unsigned long long full[PSTAGES];
    double red[8];
    double preTbl[8];
    int isLast;
};

__device__ __forceinline__ unsigned int smemAddr(const void *p) { return static_cast<unsigned int>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbarInit(unsigned long long *bar, unsigned int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(count) : "memory");
}

__device__ __forceinline__ void mbarArriveExpectTx(unsigned long long *bar, unsigned int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbarExpectTx(unsigned long long *bar, unsigned int bytes)  // more bytes expected, no arrival
{
    asm volatile("mbarrier.expect_tx.relaxed.cta.shared::cta.b64 [%0], %1;" ::"r"(smemAddr(bar)), "r"(bytes) : "memory");
}

__device__ __forceinline__ void mbarArrive(unsigned long long *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smemAddr(bar)) : "memory");
}

__device__ __forceinline__ void mbarWait(unsigned long long *bar, unsigned int parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "FS2D_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra FS2D_DONE;\n"
        "bra FS2D_WAIT;\n"
        "FS2D_DONE:\n"
        "}\n" ::"r"(smemAddr(bar)),
        "r"(parity)
        : "memory");
}

__device__ __forceinline__ void bulkLoad(void *dstSmem, const void *srcGlobal, unsigned int bytes, unsigned long long *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smemAddr(dstSmem)),
                 "l"(srcGlobal), "r"(bytes), "r"(smemAddr(bar))
                 : "memory");
}

__device__ __forceinline__ void fenceProxyAsync() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// shared -> global bulk copy (bulk-group completion), used by the paged tiles of pcgResidentKernel
__device__ __forceinline__ void bulkStore(void *dstGlobal, const void *srcSmem, unsigned int bytes)
{
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dstGlobal), "r"(smemAddr(srcSmem)), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void bulkWaitRead0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }  // sources may be overwritten
__device__ __forceinline__ void bulkWaitAllButLast() { asm volatile("cp.async.bulk.wait_group 1;" ::: "memory"); }  // all but the newest group complete
__device__ __forceinline__ void bulkWaitAll() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// Warp 0 fetches one tile: every lane owns whole rows. A row segment clipped by the ends of the vector is
// zero-filled with ordinary stores (visible to the consumers through the __syncthreads that separates
// the issue from the use of a stage).
template <int MODE>
__device__ __forceinline__ void pipeIssueV(PipeStage<MODE> &st, unsigned long long *bar, const PcgArgs &a, const double *in0,
                                          const double *in1, const double *xv, int tile, int lane);

template <int MODE>
__device__ __forceinline__ void pipeIssue(PipeStage<MODE> &st, unsigned long long *bar, const PcgArgs &a, int tile, int lane)
{
    pipeIssueV<MODE>(st, bar, a, a.in0, a.in1, a.x, tile, lane);
}

template <int MODE>
__device__ __forceinline__ void pipeIssueV(PipeStage<MODE> &st, unsigned long long *bar, const PcgArgs &a, const double *in0,
                                          const double *in1, const double *xv, int tile, int lane)
{
    const int ti = tile / a.tilesJ, tj = tile - ti * a.tilesJ;
    const long long J = a.J, N = a.N;
    const long long i0 = static_cast<long long>(ti) * TR, j0 = static_cast<long long>(tj) * TC;
    // bytes this lane will request
    unsigned int bytes = 0;
    long long lo[2], hi[2];  // valid linear range of the (up to) two row kinds this lane handles
    const int r = lane;      // PROWS = 18 <= 32: one halo-extended row per lane
    long long segLo = 0, segHi = 0, xLo = 0, xHi = 0;
    if (r < PROWS)
    {
        const long long n0 = (i0 - 1 + r) * J + j0 - 2;
        segLo = n0 < 0 ? 0 : n0;
        segHi = n0 + PSW > N ? N : n0 + PSW;
        if (segHi < segLo) segHi = segLo;
        bytes += 2u * static_cast<unsigned int>(segHi - segLo) * 8u;
        if (MODE == MODE_K1 && r >= 1 && r <= TR)
        {
            const long long m0 = (i0 - 1 + r) * J + j0;
            xLo = m0 < 0 ? 0 : m0;
            xHi = m0 + TC > N ? N : m0 + TC;
            if (xHi < xLo) xHi = xLo;
            bytes += static_cast<unsigned int>(xHi - xLo) * 8u;
        }
    }
    (void)lo;
    (void)hi;
    unsigned int total = bytes;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) total += __shfl_xor_sync(0xffffffffu, total, o);
    if (lane == 0) mbarArriveExpectTx(bar, total);
    __syncwarp();
    if (r < PROWS)
    {
        const long long n0 = (i0 - 1 + r) * J + j0 - 2;
        double *da = st.a + r * PSW, *db = st.b + r * PSW;
        if (segHi - segLo < PSW)
        {
            for (int c = 0; c < PSW; c++)
            {
                const long long n = n0 + c;
                if (n < segLo || n >= segHi)
                {
                    da[c] = 0.0;
                    db[c] = 0.0;
                }
            }
        }
        if (segHi > segLo)
        {
            const unsigned int nb = static_cast<unsigned int>(segHi - segLo) * 8u;
            bulkLoad(da + (segLo - n0), in0 + segLo, nb, bar);
            bulkLoad(db + (segLo - n0), in1 + segLo, nb, bar);
        }
        if (MODE == MODE_K1 && r >= 1 && r <= TR && xHi > xLo)
        {
            const long long m0 = (i0 - 1 + r) * J + j0;
            bulkLoad(st.x + (r - 1) * TC + (xLo - m0), xv + xLo, static_cast<unsigned int>(xHi - xLo) * 8u, bar);
        }
    }
}

// Slab mode: thread 0 of every CTA waits for the partials of the previous phase from all ranks and derives the
// scalars of this kernel; block 0 keeps the bookkeeping (iteration count, trace, convergence) in PcgScalars.
template <int MODE> __device__ void mgPrologue(const PcgArgs &a, const MgArgs &m, MgScalars *out)
{
    MgScalars r;
    r.coef = 0.0;
    r.alphaPrev = 0.0;
    r.done = 0;
    const int i = m.iter;
    double sA = 0.0, mA = 0.0, sB = 0.0, sC = 0.0, dummy = 0.0;
    if (MODE == MODE_K1)
    {
        // waits for phase 2i: init (i = 0) or K2(i-1)
        if (!mgCollect(m, 2 * i, true, &sA, &mA))
        {
            r.done = 1;
            a.sc->done = 1;  // a lost peer: let the remaining launches of this solve fall through
        }
        if (i == 0)
        {
            if (!(mA > 1.0e-15)) r.done = 1;  // VOps::isZero (vmath.cpp:47-58): x = 0, zero iterations
            if (blockIdx.x == 0)
            {
                PcgScalars *sc = a.sc;
                sc->sigma = sA;
                sc->alpha = sc->beta = sc->gamma = sc->err = 0.0;
                sc->iter = 0;
                sc->result = 0;
                if (r.done) sc->done = 1;
            }
        }
        else
        {
            mgCollect(m, 2 * i - 1, false, &sB, &dummy);  // gamma_{i-1}
            mgCollect(m, 2 * i - 2, false, &sC, &dummy);  // sigma_{i-1}
            r.alphaPrev = sC / (sB + 1e-8);                // linearsolver.cpp:50
            double beta = 0.0;
            if (mA <= a.tol)                               // :59-61
                r.done = 1;
            else
                beta = sA / sC;                            // :66-67
            r.coef = beta;
            if (blockIdx.x == 0)
            {
                PcgScalars *sc = a.sc;
                const int it = i - 1;
                sc->alpha = r.alphaPrev;
                sc->gamma = sA;
                sc->err = mA;
                if (r.done)
                {
                    sc->done = 1;
                    sc->result = it;
                }
                else
                {
                    sc->beta = beta;
                    sc->sigma = sA;
                }
                if (a.trace && it < a.traceCapacity)
                {
                    a.trace[4 * it + 0] = r.alphaPrev;
                    a.trace[4 * it + 1] = beta;
                    a.trace[4 * it + 2] = sA;
                    a.trace[4 * it + 3] = mA;
                }
                sc->iter = i;
            }
        }
    }
    else
    {
        // K2(i) waits for phase 2i+1 (gamma_i); sigma_i was published in phase 2i
        if (!mgCollect(m, 2 * i + 1, true, &sB, &dummy))
        {
            r.done = 1;
            a.sc->done = 1;
        }
        mgCollect(m, 2 * i, false, &sC, &dummy);
        r.coef = sC / (sB + 1e-8);
    }
    *out = r;
}

template <int MODE, bool MG> __global__ void __launch_bounds__(NT, 2) pcgPipeKernel(PcgArgs a, int numTiles, MgArgs mg)
{
    extern __shared__ __align__(128) unsigned char pipeRaw[];
    PipeSmem<MODE> &sm = *reinterpret_cast<PipeSmem<MODE> *>(pipeRaw);
    if (a.sc->done) return;
    __shared__ MgScalars mgs;
    if (MG)
    {
        if (threadIdx.x == 0 && blockIdx.x == 0) mgStamp(&mg, 0);
        if (threadIdx.x == 0) mgPrologue<MODE>(a, mg, &mgs);
        if (threadIdx.x == 0 && blockIdx.x == 0) mgStamp(&mg, 1);
        __syncthreads();
        if (mgs.done) return;
        asm volatile("fence.proxy.async;" ::: "memory");  // halo rows written by a peer are read by bulk copies below
    }
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (MODE == MODE_K2 && tid < 8) sm.preTbl[tid] = a.pre[tid];
    if (tid == 0)
    {
#pragma unroll
        for (int s = 0; s < PSTAGES; s++) mbarInit(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    fenceProxyAsync();
    __syncthreads();

    double coef = 0.0, alphaPrev = 0.0;
    if (MG)
    {
        coef = mgs.coef;
        alphaPrev = mgs.alphaPrev;
    }
    else if (MODE == MODE_K1)
    {
        coef = a.sc->beta;
        alphaPrev = a.sc->alpha;
    }
    else
    {
        coef = a.sc->alpha;
    }
    const long long J = a.J;
    if (a.activeCount) numTiles = *a.activeCount;
    int myTiles = (numTiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    if (myTiles < 0) myTiles = 0;
    auto tileAt = [&](int k) -> int {
        const int t = blockIdx.x + k * gridDim.x;
        return a.activeTiles ? a.activeTiles[t] : (MG ? mg.tileBase + t : t);
    };
    if (warp == 0 && myTiles > 0) pipeIssue<MODE>(sm.st[0], &sm.full[0], a, tileAt(0), lane);

    double accDot = 0.0, accMax = 0.0;
    const int bc = tid % TC, rg = tid / TC;  // stencil phase: one column, TR/2 rows per thread
    for (int k = 0; k < myTiles; k++)
    {
        const int s = k & 1;
        const unsigned int parity = static_cast<unsigned int>(k >> 1) & 1u;
        const int tile = tileAt(k);
        if (warp == 0 && k + 1 < myTiles) pipeIssue<MODE>(sm.st[s ^ 1], &sm.full[s ^ 1], a, tileAt(k + 1), lane);
        const int ti = tile / a.tilesJ, tj = tile - ti * a.tilesJ;
        const int i0 = ti * TR, j0 = tj * TC;
        PipeStage<MODE> &st = sm.st[s];

        // row info of this thread's cells: issued before the wait so the latency overlaps it
        const long long gj = j0 + bc;
        unsigned int info[TR / 2];
#pragma unroll
        for (int q = 0; q < TR / 2; q++)
        {
            const long long gi = i0 + rg * (TR / 2) + q;
            info[q] = 0;
            if (gi < a.I && gj < J)
            {
                const long long n = gi * J + gj;
                info[q] = (MODE == MODE_K1) ? static_cast<unsigned int>(a.rowInfo[n]) : static_cast<unsigned int>(a.preInfo[n]);
            }
        }

        mbarWait(&sm.full[s], parity);

        // ---- phase 1: derived vector in place, interior written back (and x advanced in K1)
        for (int e = tid; e < PTILE; e += NT)
        {
            const int ar = e / PSW, c = e - ar * PSW;
            const double av = st.a[e], bv = st.b[e];
            double v;
            if (MODE == MODE_K1)
                v = __dadd_rn(av, __dmul_rn(bv, coef));
            else
                v = __dsub_rn(av, __dmul_rn(bv, coef));
            st.a[e] = v;
            const long long gi = i0 - 1 + ar, gjj = j0 - 2 + c;
            if (ar >= 1 && ar <= TR && c >= 2 && c < TC + 2 && gi < a.I && gjj < J)
            {
                const long long n = gi * J + gjj;
                a.out0[n] = v;
                if (MODE == MODE_K1) a.x[n] = __dadd_rn(st.x[(ar - 1) * TC + (c - 2)], __dmul_rn(bv, alphaPrev));
            }
            else if (MG && c >= 2 && c < TC + 2 && gjj < J && ((ar == 0 && gi == mg.rowBegin - 1 && gi >= 0) || (gi == mg.rowEnd && gi < a.I && ar <= TR + 1)))
            {
                // halo row of the slab: the derived vector is kept current here too (the owner computes the same bits)
                a.out0[gi * J + gjj] = v;
            }
        }
        __syncthreads();

        // ---- phase 2: 5-point operator on the staged derived vector
        if (gj < J)
        {
#pragma unroll
            for (int q = 0; q < TR / 2; q++)
            {
                const int ar = 1 + rg * (TR / 2) + q;
                const long long gi = i0 - 1 + ar;
                if (gi < a.I)
                {
                    const long long n = gi * J + gj;
                    const double *t = st.a + ar * PSW + bc + 2;
                    const double c = t[0], im = t[-PSW], ip = t[PSW], jm = t[-1], jp = t[1];
                    double o;
                    if (MODE == MODE_K1)
                        o = rowA(static_cast<uint8_t>(info[q]), a.scale, c, im, ip, jm, jp);
                    else
                        o = rowM(static_cast<uint16_t>(info[q]), sm.preTbl, c, im, ip, jm, jp);
                    a.out1[n] = o;
                    if (MG && !(mg.debug & 2))
                    {
                        if (gi == mg.rowBegin && mg.loOut1) mg.loOut1[n] = o;
                        if (gi == mg.rowEnd - 1 && mg.hiOut1) mg.hiOut1[n] = o;
                    }
                    accDot += o * c;
                    accMax = fmax(accMax, fabs(c));
                }
            }
        }
        fenceProxyAsync();   // generic writes to this stage are ordered before the next bulk copy into it
        __syncthreads();
    }
    if (MG && threadIdx.x == 0 && blockIdx.x == 0) mgStamp(&mg, 2);
    finishReductions<MODE>(a, accDot, accMax, sm.red, &sm.isLast, MG ? &mg : nullptr);
}

// ------------------------------------------------------------------ whole-solve kernel
// One persistent, cooperatively launched kernel runs ALL iterations of a solve. The two kernel boundaries per
// iteration of the pipelined path (launch gap + cold pipeline + last-CTA reduction, ~10 us each on one GPU, ~20 us
// with peers) become two grid-wide barriers that carry the reductions:
//   every CTA: block partial -> partials[cta], arrive on a monotonic ticket;
//   the last CTA to arrive reduces the partials in a fixed order and PUBLISHES the rank's sum / max into the mail slot
//   of every rank (its own included; peers through peer-mapped memory), tagged with the phase;
//   every CTA of every rank waits for the tags of all ranks in its OWN memory and adds the values in rank order.
// So the all-reduce across GPUs is the barrier itself (one NVLink hop), every CTA everywhere derives bit-identical
// alpha / beta / err, and the scalars never leave registers. Tiles are assigned to CTAs statically; the tile walk
// (bulk-copy pipeline, in-place derived vector, 5-point operator from shared memory) is the one of pcgPipeKernel.
// Phase numbering, mail ring and halo-row pushes are those of the slab path above (phase 0 = pcgInitKernel<true>).
__device__ __forceinline__ unsigned int llTag(const MgArgs &m, int phase)
{
    return static_cast<unsigned int>((m.solveTag >> 4) + static_cast<unsigned long long>(phase) + 1ull);  // solveSeq * 65536 + phase + 1
}

// ------------------------------------------------------------------ LL halo rows between resident kernels
// A slab-boundary row of q (K1) / z (K2) pushed into the neighbour's array has to be ordered before the barrier that
// publishes it: a system-scope fence behind remote stores waits for their acknowledgement (~3 us per phase, measured on
// 2 B200: 13.4 -> 10.7 us). When BOTH neighbours run pcgResidentKernel the row travels instead as self-validating
// 8-byte words {32 payload bits, 32-bit phase tag} (the LL idea again): no fence on the producer, the consumer's ring
// threads poll the two words of their cell. Which kernel a rank runs is decided on the device (tile count), so the
// kernels tell their row neighbours at the start of every solve (SlabMail::mode) and fall back to the plain push +
// fence towards a neighbour that streams.
enum { SOLVE_MODE_RESIDENT = 1, SOLVE_MODE_STREAMING = 2 };

template <bool MG> __device__ __forceinline__ void solveModePublish(const MgArgs &m, int mode, unsigned long long firstRowTiles = 0,
                                                                    unsigned long long lastRowTiles = 0)
{
    if (!MG) return;
    const unsigned long long word = (static_cast<unsigned long long>(llTag(m, 0)) << 32) | static_cast<unsigned long long>(mode);
    const int par = m.ringBase ? 1 : 0;
    for (int d = -1; d <= 1; d += 2)
    {
        const int r = m.rank + d;
        if (r < 0 || r >= m.world) continue;
        SlabMail *pm = m.peerMail[r];
        if (mode == SOLVE_MODE_RESIDENT)
        {
            *reinterpret_cast<volatile unsigned long long *>(&pm->edgeTiles[par][m.rank][0]) = firstRowTiles;
            *reinterpret_cast<volatile unsigned long long *>(&pm->edgeTiles[par][m.rank][1]) = lastRowTiles;
            __threadfence_system();  // once per solve: the masks are in place before the word that announces them
        }
        *reinterpret_cast<volatile unsigned long long *>(&pm->mode[par][m.rank]) = word;
    }
}

// mode of the row neighbour `r` for this solve (0 when there is none, -1 when it never answered)
template <bool MG> __device__ __forceinline__ int solveModeOf(const MgArgs &m, int r)
{
    if (!MG || r < 0 || r >= m.world) return 0;
    const int par = m.ringBase ? 1 : 0;
    const unsigned int want = llTag(m, 0);
    const long long t0 = clock64();
    for (;;)
    {
        unsigned long long w;
        asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(&m.mail->mode[par][r]) : "memory");
        if (static_cast<unsigned int>(w >> 32) == want) return static_cast<int>(w & 0xffffffffull);
        if (clock64() - t0 > MG_SPIN_LIMIT)
        {
            m.mail->error = 3;
            return -1;
        }
    }
}

__device__ __forceinline__ void llHaloStore(unsigned long long *buf, int side, int kind, long long J, long long gj, double v, unsigned int tag)
{
    const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(v));
    unsigned long long *dst = buf + ((static_cast<long long>(side) * 2 + kind) * J + gj) * 2;
    const unsigned long long t = static_cast<unsigned long long>(tag) << 32;
    asm volatile("st.relaxed.sys.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"((bits & 0xffffffffull) | t), "l"((bits >> 32) | t) : "memory");
}

__device__ __forceinline__ double llHaloLoad(const unsigned long long *buf, int side, int kind, long long J, long long gj, unsigned int tag,
                                             int *error)
{
    const unsigned long long *src = buf + ((static_cast<long long>(side) * 2 + kind) * J + gj) * 2;
    const long long t0 = clock64();
    for (;;)
    {
        unsigned long long a, b;
        asm volatile("ld.relaxed.sys.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(src) : "memory");
        if (static_cast<unsigned int>(a >> 32) == tag && static_cast<unsigned int>(b >> 32) == tag)
            return __longlong_as_double(static_cast<long long>((a & 0xffffffffull) | (b << 32)));
        if (clock64() - t0 > MG_SPIN_LIMIT)
        {
            *error = 4;
            return 0.0;
        }
    }
}

// Threads 0 .. 4*world-1 of the calling CTA store one word each into rank (t >> 2)'s mail.
template <bool MG> __device__ __forceinline__ void llPublish(const MgArgs &m, int phase, double v0, double v1)
{
    const int t = threadIdx.x;
    if (t < 4 * m.world)
    {
        const int r = t >> 2, word = t & 3;
        const unsigned long long bits = static_cast<unsigned long long>(__double_as_longlong(word < 2 ? v0 : v1));
        const unsigned long long half = (word & 1) ? (bits >> 32) : (bits & 0xffffffffull);
        unsigned long long *dst = &m.peerMail[r]->ll[m.ringBase + (phase & 7)][m.rank].w[word];
        // Everything this rank wrote before the barrier is ordered before the word the consumers wait for -- at GPU scope:
        // the tile outputs are only ever read by CTAs of this rank; halo rows stored into a neighbour's arrays were fenced
        // at system scope by the CTA that stored them, before its ticket (resBarrier / solveBarrier, `remoteStores`), and
        // this CTA has seen every ticket; halo rows sent as self-validating words need no ordering at all. A system-scope
        // fence here costs 1.3 us per barrier (one rank in slab mode: 4.5-4.9 -> 3.2-3.6 us; FS2D_MG_DEBUG & 64 restores it).
        if (MG && (m.debug & 64))
            __threadfence_system();
        else
            __threadfence();
        *reinterpret_cast<volatile unsigned long long *>(dst) = half | (static_cast<unsigned long long>(llTag(m, phase)) << 32);
    }
}

// Warp 0 of the calling CTA: lane l < 4*world polls word (l & 3) of rank (l >> 2) in this rank's OWN mail; then the
// values are combined in rank order (identical bits in every CTA of every rank). Returns false on a lost peer.
template <bool MG> __device__ __forceinline__ bool llCollect(const MgArgs &m, int phase, double *sum, double *mx)
{
    const int lane = threadIdx.x & 31;
    const unsigned int want = llTag(m, phase);
    unsigned long long w = 0;
    bool ok = true;
    if (lane < 4 * m.world)
    {
        const unsigned long long *src = &m.mail->ll[m.ringBase + (phase & 7)][lane >> 2].w[lane & 3];
        const long long t0 = clock64();
        const bool relaxedPoll = MG && (m.debug & (32 | 128));  // A/B: poll with relaxed loads; 32: one fence after the last word arrived, 128: none
        for (;;)
        {
            if (relaxedPoll)
                asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
            else if (MG)
                asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
            else
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
            if (static_cast<unsigned int>(w >> 32) == want) break;
            if (clock64() - t0 > MG_SPIN_LIMIT)
            {
                ok = false;
                m.mail->error = 2;
                break;
            }
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    if (MG && (m.debug & 32)) __threadfence_system();
    double s = 0.0, x = 0.0;
    for (int r = 0; r < m.world; r++)
    {
        const unsigned long long a0 = __shfl_sync(0xffffffffu, w, 4 * r), a1 = __shfl_sync(0xffffffffu, w, 4 * r + 1);
        const unsigned long long b0 = __shfl_sync(0xffffffffu, w, 4 * r + 2), b1 = __shfl_sync(0xffffffffu, w, 4 * r + 3);
        s += __longlong_as_double(static_cast<long long>((a0 & 0xffffffffull) | (a1 << 32)));
        x = fmax(x, __longlong_as_double(static_cast<long long>((b0 & 0xffffffffull) | (b1 << 32))));
    }
    *sum = s;
    *mx = x;
    return ok;
}

// The same for a rank that already knows its own sums (lane 0 holds them): only the words of the OTHER ranks are waited for.
template <bool MG> __device__ __forceinline__ bool llCollectPeers(const MgArgs &m, int phase, double ownSum, double ownMax, double *sum, double *mx)
{
    const int lane = threadIdx.x & 31;
    const unsigned int want = llTag(m, phase);
    unsigned long long w = 0;
    bool ok = true;
    if (lane < 4 * m.world && (lane >> 2) != m.rank)
    {
        const unsigned long long *src = &m.mail->ll[m.ringBase + (phase & 7)][lane >> 2].w[lane & 3];
        const long long t0 = clock64();
        for (;;)
        {
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(w) : "l"(src) : "memory");
            if (static_cast<unsigned int>(w >> 32) == want) break;
            if (clock64() - t0 > MG_SPIN_LIMIT)
            {
                ok = false;
                m.mail->error = 2;
                break;
            }
        }
    }
    ok = __all_sync(0xffffffffu, ok);
    ownSum = __shfl_sync(0xffffffffu, ownSum, 0);
    ownMax = __shfl_sync(0xffffffffu, ownMax, 0);
    double s = 0.0, x = 0.0;
    for (int r = 0; r < m.world; r++)
    {
        const unsigned long long a0 = __shfl_sync(0xffffffffu, w, 4 * r), a1 = __shfl_sync(0xffffffffu, w, 4 * r + 1);
        const unsigned long long b0 = __shfl_sync(0xffffffffu, w, 4 * r + 2), b1 = __shfl_sync(0xffffffffu, w, 4 * r + 3);
        const bool mine = r == m.rank;
        s += mine ? ownSum : __longlong_as_double(static_cast<long long>((a0 & 0xffffffffull) | (a1 << 32)));
        x = fmax(x, mine ? ownMax : __longlong_as_double(static_cast<long long>((b0 & 0xffffffffull) | (b1 << 32))));
    }
    *sum = s;
    *mx = x;
    return ok;
}

// Sum and max over the CTA in one pass (same operation order as blockReduce); result valid in thread 0.
__device__ __forceinline__ void blockReduce2(double &s, double &m, double *scratch /* >= 16 doubles */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s = warpSum(s);
    m = warpMax(m);
    __syncthreads();
    if (lane == 0)
    {
        scratch[warp] = s;
        scratch[8 + warp] = m;
    }
    __syncthreads();
    if (warp == 0)
    {
        s = (lane < (blockDim.x >> 5)) ? scratch[lane] : 0.0;
        m = (lane < (blockDim.x >> 5)) ? scratch[8 + lane] : 0.0;
        s = warpSum(s);
        m = warpMax(m);
    }
}

// Tensor maps of the Krylov vectors seen as I x J matrices of doubles: one cp.async.bulk.tensor.2d (SASS UTMALDG)
// brings a whole halo-extended tile -- 18 rows of 132 doubles -- instead of 18 row copies, and out-of-range rows /
// columns arrive as zeros. (The row copies of pcgPipeKernel cost ~50 TMA instructions per tile; at 2 us per tile and
// SM that instruction stream is what bounded the K2 phase and the start of every phase.)
enum { TM_Z = 0, TM_Q, TM_S0, TM_S1, TM_R0, TM_R1, TM_X, TM_COUNT };

struct SolveMaps
{
    CUtensorMap m[TM_COUNT];
};

// L2 cache policies per operand role (0 = none). In the active-tile walk the seven vectors of the walked cells
// (~100 MB at 4096^2) cycle through the L2 once per iteration and, untreated, evict each other before they are used
// again: the walk then runs at HBM speed although the set nearly fits. x is pure streaming (read-modify-write once per
// iteration, never used by the recurrences) and is marked evict-first; q and z, written by one phase and read by the
// next, evict-last. FS2D_PCG_HINTS selects the combination (A/B switch).
struct WalkHints
{
    unsigned long long in0, in1, x, out0, out1;
    unsigned long long xStore = 0;  // policy of the x store when it differs from the x load (0: same as x)
};

__device__ __forceinline__ unsigned long long policyEvictFirst()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ unsigned long long policyEvictLast()
{
    unsigned long long p;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
    return p;
}

__device__ __forceinline__ void storeHint2(double *p, double2 v, unsigned long long pol)
{
    if (pol)
        asm volatile("st.global.L2::cache_hint.v2.f64 [%0], {%1, %2}, %3;" ::"l"(p), "d"(v.x), "d"(v.y), "l"(pol) : "memory");
    else
        *reinterpret_cast<double2 *>(p) = v;
}

__device__ __forceinline__ void storeHint1(double *p, double v, unsigned long long pol)
{
    if (pol)
        asm volatile("st.global.L2::cache_hint.f64 [%0], %1, %2;" ::"l"(p), "d"(v), "l"(pol) : "memory");
    else
        *p = v;
}

__device__ __forceinline__ void tmaLoad2DHint(void *dstSmem, const CUtensorMap *map, int col, int row, unsigned long long *bar,
                                              unsigned long long pol)
{
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(
            smemAddr(dstSmem)),
        "l"(reinterpret_cast<unsigned long long>(map)), "r"(smemAddr(bar)), "r"(col), "r"(row), "l"(pol)
        : "memory");
}

__device__ __forceinline__ void tmaLoad2D(void *dstSmem, const CUtensorMap *map, int col, int row, unsigned long long *bar,
                                          unsigned long long pol = 0)
{
    if (pol)
    {
        tmaLoad2DHint(dstSmem, map, col, row, bar, pol);
        return;
    }
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                     smemAddr(dstSmem)),
                 "l"(reinterpret_cast<unsigned long long>(map)), "r"(smemAddr(bar)), "r"(col), "r"(row)
                 : "memory");
}

// One thread fetches one tile: the two halo-extended input boxes (and the x tile in K1).
template <int MODE>
__device__ __forceinline__ void pipeIssueT(PipeStage<MODE> &st, unsigned long long *bar, const CUtensorMap *in0, const CUtensorMap *in1,
                                          const CUtensorMap *xm, int tilesJ, int tile, const WalkHints &h)
{
    const int ti = tile / tilesJ, tj = tile - ti * tilesJ;
    const int i0 = ti * TR, j0 = tj * TC;
    constexpr unsigned int boxBytes = PTILE * 8u, xBytes = TR * TC * 8u;
    mbarArriveExpectTx(bar, 2u * boxBytes + (MODE == MODE_K1 ? xBytes : 0u));
    tmaLoad2D(st.a, in0, j0 - 2, i0 - 1, bar, h.in0);
    tmaLoad2D(st.b, in1, j0 - 2, i0 - 1, bar, h.in1);
    if (MODE == MODE_K1) tmaLoad2D(st.x, xm, j0, i0, bar, h.x);
}

// The first tile of a phase in two halves. EARLY (issued BEFORE the barrier that ends the previous phase, so that it
// overlaps the barrier): the operands the previous phase did not write -- K1: s_old and x (written a whole iteration
// ago), K2: r_old. LATE (after the barrier): the vector the previous phase produced -- K1: z, K2: q. The early half only
// raises the transaction count; the single arrival of the stage comes with the late half.
template <int MODE>
__device__ __forceinline__ void pipeIssueEarly(PipeStage<MODE> &st, unsigned long long *bar, const CUtensorMap *early, const CUtensorMap *xm,
                                              int tilesJ, int tile, const WalkHints &h)
{
    const int ti = tile / tilesJ, tj = tile - ti * tilesJ;
    const int i0 = ti * TR, j0 = tj * TC;
    constexpr unsigned int boxBytes = PTILE * 8u, xBytes = TR * TC * 8u;
    mbarExpectTx(bar, boxBytes + (MODE == MODE_K1 ? xBytes : 0u));
    tmaLoad2D(MODE == MODE_K1 ? st.b : st.a, early, j0 - 2, i0 - 1, bar, MODE == MODE_K1 ? h.in1 : h.in0);
    if (MODE == MODE_K1) tmaLoad2D(st.x, xm, j0, i0, bar, h.x);
}

template <int MODE>
__device__ __forceinline__ void pipeIssueLate(PipeStage<MODE> &st, unsigned long long *bar, const CUtensorMap *late, int tilesJ, int tile,
                                             const WalkHints &h)
{
    const int ti = tile / tilesJ, tj = tile - ti * tilesJ;
    const int i0 = ti * TR, j0 = tj * TC;
    mbarArriveExpectTx(bar, PTILE * 8u);
    tmaLoad2D(MODE == MODE_K1 ? st.a : st.b, late, j0 - 2, i0 - 1, bar, MODE == MODE_K1 ? h.in0 : h.in1);
}

struct SolveSmem
{
    PipeStage<MODE_K1> st[PSTAGES];       // phase B views each stage as a (smaller) PipeStage<MODE_K2>
    unsigned long long full[PSTAGES];
    double red[16];
    double preTbl[8];
    double pub[2];
    double bc[2];
    int isLast;
    int ok;
};

struct SolveArgs
{
    PcgArgs a;                            // operator tables, partials, scalars, trace, tol, active-tile list
    double *z, *q, *x, *s[2], *r[2];
    double *loQ, *hiQ, *loZ, *hiZ;        // the row neighbours' copies of q and z (slab mode)
    const int *tileFlags;                 // activity flag of each of the numTiles tiles this rank walks (nullptr: dense walk)
    int numTiles;                         // tiles of the dense walk (active walk: *a.activeCount)
    int iterLimit;
    unsigned int *ticket;                 // zeroed before the launch
};

template <int MODE, bool MG>
__device__ __forceinline__ bool pipeWalk(PipeStage<MODE> *st0, PipeStage<MODE> *st1, unsigned long long *full, const double *preTbl,
                                         const PcgArgs &a, const MgArgs &mg, const double *__restrict__ in0,
                                         const double *__restrict__ in1, double *__restrict__ out0, double *__restrict__ out1,
                                         double *__restrict__ xv, double *loOut1, double *hiOut1, double coef, double alphaPrev,
                                         int numTiles, unsigned int &use0, unsigned int &use1, double &accDot, double &accMax,
                                         int phase = 0, const CUtensorMap *tm0 = nullptr, const CUtensorMap *tm1 = nullptr,
                                         const CUtensorMap *tmx = nullptr, bool earlyIssued = false, WalkHints h = WalkHints{0, 0, 0, 0, 0, 0})
{
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const long long J = a.J;
    if (blockIdx.x == 0 && tid == 0) mgStampAt(mg, phase, 0);
    int myTiles = (numTiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    if (myTiles < 0) myTiles = 0;
    auto tileAt = [&](int k) -> int {
        const int t = blockIdx.x + k * gridDim.x;
        return a.activeTiles ? a.activeTiles[t] : (MG ? mg.tileBase + t : t);
    };
    const bool tensor = tm0 != nullptr;
    bool remote = false;  // (uniform over the CTA) one of its tiles holds a slab-boundary row pushed into a neighbour
    if (myTiles > 0)
    {
        if (tensor)
        {
            if (tid == 0)
            {
                if (earlyIssued)
                    pipeIssueLate<MODE>(*st0, &full[0], MODE == MODE_K1 ? tm0 : tm1, a.tilesJ, tileAt(0), h);
                else
                    pipeIssueT<MODE>(*st0, &full[0], tm0, tm1, tmx, a.tilesJ, tileAt(0), h);
            }
        }
        else if (warp == 0)
            pipeIssueV<MODE>(*st0, &full[0], a, in0, in1, xv, tileAt(0), lane);
    }
    const int bc = tid % TC, rg = tid / TC;
    for (int k = 0; k < myTiles; k++)
    {
        const int s = k & 1;
        const unsigned int parity = (s ? use1 : use0) & 1u;
        if (s)
            use1++;
        else
            use0++;
        const int tile = tileAt(k);
        if (k + 1 < myTiles)
        {
            if (tensor)
            {
                if (tid == 0) pipeIssueT<MODE>(s ? *st0 : *st1, &full[s ^ 1], tm0, tm1, tmx, a.tilesJ, tileAt(k + 1), h);
            }
            else if (warp == 0)
                pipeIssueV<MODE>(s ? *st0 : *st1, &full[s ^ 1], a, in0, in1, xv, tileAt(k + 1), lane);
        }
        const int ti = tile / a.tilesJ, tj = tile - ti * a.tilesJ;
        const int i0 = ti * TR, j0 = tj * TC;
        PipeStage<MODE> &st = s ? *st1 : *st0;
        const long long gj = j0 + bc;
        if (MG)
            remote = remote || (loOut1 && mg.rowBegin >= i0 && mg.rowBegin < i0 + TR) || (hiOut1 && mg.rowEnd - 1 >= i0 && mg.rowEnd - 1 < i0 + TR);
        // Tensor copies zero-fill what lies outside the I x J matrix; the reference's operators address the j = -1 /
        // j = J neighbours by LINEAR index, i.e. the last / first element of the adjacent row (pressuredata.h:135-145).
        // Threads 0..15 (left edge tiles) and 16..31 (right edge tiles) fetch those values while the tile is in flight.
        const bool leftEdge = tensor && tj == 0, rightEdge = tensor && j0 + TC >= a.J;
        double wrapA = 0.0, wrapB = 0.0;
        int wrapAt = -1;
        if ((leftEdge && tid < TR) || (rightEdge && tid >= TR && tid < 2 * TR))
        {
            const int ar = 1 + (tid & (TR - 1));
            const long long gi = i0 - 1 + ar;
            const bool left = tid < TR;
            const long long n = left ? gi * J - 1 : gi * J + J;
            wrapAt = ar * PSW + (left ? 1 : static_cast<int>(J - j0) + 2);
            if (n >= 0 && n < a.N && gi < a.I)
            {
                wrapA = in0[n];
                wrapB = in1[n];
            }
        }
        unsigned int info[TR / 2];
#pragma unroll
        for (int q = 0; q < TR / 2; q++)
        {
            const long long gi = i0 + rg * (TR / 2) + q;
            info[q] = 0;
            if (gi < a.I && gj < J)
            {
                const long long n = gi * J + gj;
                info[q] = (MODE == MODE_K1) ? static_cast<unsigned int>(a.rowInfo[n]) : static_cast<unsigned int>(a.preInfo[n]);
            }
        }
        mbarWait(&full[s], parity);
        if (k == 0 && blockIdx.x == 0 && tid == 0) mgStampAt(mg, phase, 1);
        if (leftEdge || rightEdge)
        {
            if (wrapAt >= 0)
            {
                st.a[wrapAt] = wrapA;
                st.b[wrapAt] = wrapB;
            }
            __syncthreads();
        }
        // phase 1, two elements per thread and step: the staged rows hold an even number of doubles, the interior
        // starts at an even column and J is even, so a pair never straddles a row, the interior or the matrix edge;
        // 16-byte shared and global accesses halve the instruction count of this phase
        for (int e = 2 * tid; e < PTILE; e += 2 * NT)
        {
            const int ar = e / PSW, c = e - ar * PSW;
            const double2 av = *reinterpret_cast<const double2 *>(st.a + e), bv = *reinterpret_cast<const double2 *>(st.b + e);
            double2 v;
            if (MODE == MODE_K1)
            {
                v.x = __dadd_rn(av.x, __dmul_rn(bv.x, coef));
                v.y = __dadd_rn(av.y, __dmul_rn(bv.y, coef));
            }
            else
            {
                v.x = __dsub_rn(av.x, __dmul_rn(bv.x, coef));
                v.y = __dsub_rn(av.y, __dmul_rn(bv.y, coef));
            }
            *reinterpret_cast<double2 *>(st.a + e) = v;
            const long long gi = i0 - 1 + ar, gjj = j0 - 2 + c;
            if (ar >= 1 && ar <= TR && c >= 2 && c < TC + 2 && gi < a.I && gjj < J)
            {
                const long long n = gi * J + gjj;
                storeHint2(out0 + n, v, h.out0);
                if (MODE == MODE_K1)
                {
                    const double2 xo = *reinterpret_cast<const double2 *>(st.x + (ar - 1) * TC + (c - 2));
                    double2 xn;
                    xn.x = __dadd_rn(xo.x, __dmul_rn(bv.x, alphaPrev));
                    xn.y = __dadd_rn(xo.y, __dmul_rn(bv.y, alphaPrev));
                    storeHint2(xv + n, xn, h.xStore ? h.xStore : h.x);
                }
            }
            else if (MG && c >= 2 && c < TC + 2 && gjj < J && ((ar == 0 && gi == mg.rowBegin - 1 && gi >= 0) || (gi == mg.rowEnd && gi < a.I && ar <= TR + 1)))
            {
                *reinterpret_cast<double2 *>(out0 + gi * J + gjj) = v;  // halo row of the slab, kept current redundantly (same bits as on the owner)
            }
        }
        __syncthreads();
        if (gj < J)
        {
            // phase 2: the thread walks down its column; the vertical neighbours slide through registers
            const double *t = st.a + (1 + rg * (TR / 2)) * PSW + bc + 2;
            double im = t[-PSW], c = t[0];
#pragma unroll
            for (int q = 0; q < TR / 2; q++)
            {
                const int ar = 1 + rg * (TR / 2) + q;
                const long long gi = i0 - 1 + ar;
                const double ip = t[PSW], jm = t[-1], jp = t[1];
                if (gi < a.I)
                {
                    const long long n = gi * J + gj;
                    double o;
                    if (MODE == MODE_K1)
                        o = rowA(static_cast<uint8_t>(info[q]), a.scale, c, im, ip, jm, jp);
                    else
                        o = rowM(static_cast<uint16_t>(info[q]), preTbl, c, im, ip, jm, jp);
                    storeHint1(out1 + n, o, h.out1);
                    if (MG)
                    {
                        if (gi == mg.rowBegin && loOut1) loOut1[n] = o;
                        if (gi == mg.rowEnd - 1 && hiOut1) hiOut1[n] = o;
                    }
                    accDot += o * c;
                    accMax = fmax(accMax, fabs(c));
                }
                im = c;
                c = ip;
                t += PSW;
            }
        }
        fenceProxyAsync();
        __syncthreads();
    }
    return remote;
}

// Grid-wide (and, with slabs, machine-wide) barrier that all-reduces (v0: sum, v1: max). Returns false when a peer
// did not answer within the spin limit.
template <bool MG>
__device__ __forceinline__ bool solveBarrier(const SolveArgs &g, const MgArgs &m, int phase, unsigned int barrierIndex, double v0,
                                             double v1, SolveSmem &sm, double *sum, double *mx, bool remoteStores = true)
{
    const int tid = threadIdx.x;
    const unsigned int nb = gridDim.x;
    if (blockIdx.x == 0 && tid == 0) mgStampAt(m, phase, 2);
    blockReduce2(v0, v1, sm.red);
    if (!MG)
    {
        // One GPU: no publication step. Every CTA waits for the ticket to reach this barrier's target and then sums the
        // partials itself, all in the same fixed order -> identical bits in every CTA, one L2 round trip less than
        // "last CTA reduces and publishes". The partials are double-buffered by barrier parity: a CTA can run at most
        // one barrier ahead of the slowest reader.
        double *part = g.a.partials + (barrierIndex & 1u) * 2u * nb;
        if (tid == 0)
        {
            part[blockIdx.x] = v0;
            part[nb + blockIdx.x] = v1;
            __threadfence();
            atomicAdd(g.ticket, 1u);
            if (blockIdx.x == 0) mgStampAt(m, phase, 3);
            const unsigned int target = (barrierIndex + 1u) * nb;
            unsigned int seen;
            do
            {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(g.ticket) : "memory");
            } while (seen < target);
            if (blockIdx.x == 0) mgStampAt(m, phase, 4);
        }
        __syncthreads();
        double ts = 0.0, tm = 0.0;
        for (unsigned int k = tid; k < nb; k += blockDim.x)
        {
            ts += __ldcg(part + k);
            tm = fmax(tm, __ldcg(part + nb + k));
        }
        blockReduce2(ts, tm, sm.red);
        if (tid == 0)
        {
            sm.bc[0] = ts;
            sm.bc[1] = tm;
            if (blockIdx.x == 0) mgStampAt(m, phase, 5);
        }
        __syncthreads();
        *sum = sm.bc[0];
        *mx = sm.bc[1];
        asm volatile("fence.proxy.async;" ::: "memory");  // what other CTAs wrote is read by bulk copies next
        if (blockIdx.x == 0 && tid == 0) mgStampAt(m, phase, 6);
        return true;
    }
    if (tid == 0)
    {
        g.a.partials[blockIdx.x] = v0;
        g.a.partials[nb + blockIdx.x] = v1;
        // A CTA that pushed a halo row into a neighbour's array orders those stores at system scope before it arrives
        // (the publication by the last CTA then covers them); for everybody else the data only has to be visible on this
        // GPU. A system-scope fence behind a tile's worth of stores costs 3-5 us, a device-scope one ~1 us.
        if (MG && remoteStores)
            __threadfence_system();
        else
            __threadfence();
        sm.isLast = (atomicAdd(g.ticket, 1u) == (barrierIndex + 1u) * nb - 1u);
        if (blockIdx.x == 0) mgStampAt(m, phase, 3);
    }
    __syncthreads();
    if (sm.isLast)
    {
        if (tid == 0) mgStampAt(m, phase, 4);
        // fixed assignment of partials to threads, then the fixed block tree: deterministic, same order as finalReduce
        double ts = 0.0, tm = 0.0;
        for (unsigned int k = tid; k < nb; k += blockDim.x)
        {
            ts += __ldcg(g.a.partials + k);
            tm = fmax(tm, __ldcg(g.a.partials + nb + k));
        }
        blockReduce2(ts, tm, sm.red);
        if (tid == 0)
        {
            sm.pub[0] = ts;
            sm.pub[1] = tm;
        }
        __syncthreads();
        llPublish<MG>(m, phase, sm.pub[0], sm.pub[1]);
        if (tid == 0) mgStampAt(m, phase, 5);
    }
    if (tid < 32)
    {
        double s = 0.0, x = 0.0;
        const bool ok = llCollect<MG>(m, phase, &s, &x);
        if (tid == 0)
        {
            sm.ok = ok ? 1 : 0;
            sm.bc[0] = s;
            sm.bc[1] = x;
        }
    }
    __syncthreads();
    *sum = sm.bc[0];
    *mx = sm.bc[1];
    asm volatile("fence.proxy.async;" ::: "memory");  // what other CTAs / peers wrote is read by bulk copies next
    if (blockIdx.x == 0 && tid == 0) mgStampAt(m, phase, 6);
    return sm.ok != 0;
}

template <bool MG>
__global__ void __launch_bounds__(NT, 2) pcgSolveKernel(SolveArgs g, MgArgs mg, const __grid_constant__ SolveMaps tm, int useTensor)
{
    extern __shared__ __align__(128) unsigned char solveRaw[];
    SolveSmem &sm = *reinterpret_cast<SolveSmem *>(solveRaw);
    const int tid = threadIdx.x;
    const bool scribe = blockIdx.x == 0 && tid == 0;  // keeps PcgScalars / the trace for the host
    PcgScalars *sc = g.a.sc;
    if (sc->pad) return;  // pcgResidentKernel (launched just before) took this solve
    if (MG && scribe) solveModePublish<MG>(mg, SOLVE_MODE_STREAMING);  // row neighbours running the resident kernel push plain rows + fence
    if (tid < 8) sm.preTbl[tid] = g.a.pre[tid];
    if (tid == 0)
    {
#pragma unroll
        for (int s = 0; s < PSTAGES; s++) mbarInit(&sm.full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        double s0 = 0.0, m0 = 0.0;
        sm.ok = mgCollect(mg, 0, true, &s0, &m0) ? 1 : 0;  // phase 0: rhs.rhs and max|rhs| from pcgInitKernel<true>
        sm.bc[0] = s0;
        sm.bc[1] = m0;
    }
    fenceProxyAsync();
    __syncthreads();
    asm volatile("fence.proxy.async;" ::: "memory");
    double sigma = sm.bc[0];
    const double max0 = sm.bc[1];
    const bool lost0 = sm.ok == 0;
    __syncthreads();
    if (lost0 || !(max0 > 1.0e-15))  // a lost peer, or VOps::isZero (vmath.cpp:47-58): x = 0, zero iterations
    {
        if (scribe)
        {
            sc->sigma = sigma;
            sc->alpha = sc->beta = sc->gamma = sc->err = 0.0;
            sc->iter = 0;
            sc->result = 0;
            sc->done = 1;
        }
        return;
    }
    const int numTiles = g.a.activeCount ? *g.a.activeCount : g.numTiles;
    PipeStage<MODE_K1> *a0 = &sm.st[0], *a1 = &sm.st[1];
    PipeStage<MODE_K2> *b0 = reinterpret_cast<PipeStage<MODE_K2> *>(&sm.st[0]), *b1 = reinterpret_cast<PipeStage<MODE_K2> *>(&sm.st[1]);
    unsigned int use0 = 0, use1 = 0, bar = 0;
    double alpha = 0.0, beta = 0.0, alphaPrev = 0.0, gamma = 0.0, err = 0.0;
    int result = g.iterLimit, executed = 0;
    unsigned long long tA = 0, tB = 0, t0 = scribe ? globalTimerNs() : 0ull;
    // first tile of this CTA (the one every walk starts with) and whether its early half is in flight
    const bool hasTile = static_cast<int>(blockIdx.x) < numTiles;
    const int firstTile = !hasTile ? 0 : (g.a.activeTiles ? g.a.activeTiles[blockIdx.x] : (MG ? mg.tileBase + static_cast<int>(blockIdx.x) : static_cast<int>(blockIdx.x)));
    const bool canEarly = (useTensor & 1) && hasTile;
    bool early = false;
    const int hintMask = useTensor >> 8;
    const unsigned long long pf = hintMask ? policyEvictFirst() : 0ull, pl = hintMask ? policyEvictLast() : 0ull;
    const unsigned long long hx = (hintMask & 1) ? pf : 0ull;                               // x
    const unsigned long long hqz = (hintMask & 2) ? pl : 0ull;                              // q and z
    const unsigned long long hsr = (hintMask & 4) ? pf : ((hintMask & 8) ? pl : 0ull);      // s and r
    // bit 16: every element is written once and read once per iteration, so every load is a LAST use (evict-first) and
    // every store will be read exactly once (evict-last); x loads / stores follow bit 32 (0: like the rest, 1: streaming)
    const bool lastUse = (hintMask & 16) != 0;
    const unsigned long long hxs = (hintMask & 32) ? pf : pl;
    const WalkHints hA = lastUse ? WalkHints{pf, pf, pf, pl, pl} : WalkHints{hqz, hsr, hx, hsr, hqz};  // K1: in0 = z, in1 = s_old, x, out0 = s_new, out1 = q
    const WalkHints hB = lastUse ? WalkHints{pf, pf, 0ull, pl, pl} : WalkHints{hsr, hqz, 0ull, hsr, hqz};  // K2: in0 = r_old, in1 = q, out0 = r_new, out1 = z
    WalkHints hA2 = hA;
    hA2.xStore = lastUse ? hxs : 0ull;
    for (int i = 0; i < g.iterLimit; i++)
    {
        // K1(i): s_i = z + beta s_{i-1}; x += alpha_{i-1} s_{i-1}; q = A s_i; gamma = q.s_i
        double accDot = 0.0, accMax = 0.0, unused = 0.0;
        const bool remoteA = pipeWalk<MODE_K1, MG>(a0, a1, sm.full, sm.preTbl, g.a, mg, g.z, g.s[i & 1], g.s[(i + 1) & 1], g.q, g.x, g.loQ, g.hiQ, beta,
                              alphaPrev, numTiles, use0, use1, accDot, accMax, 2 * i + 1, (useTensor & 1) ? &tm.m[TM_Z] : nullptr,
                              &tm.m[TM_S0 + (i & 1)], &tm.m[TM_X], early, hA2);
        early = canEarly;  // r_old of K2(i) was written an iteration ago: fetch it while the barrier runs
        if (early && tid == 0) pipeIssueEarly<MODE_K2>(*b0, &sm.full[0], &tm.m[TM_R0 + (i & 1)], nullptr, g.a.tilesJ, firstTile, hB);
        if (!solveBarrier<MG>(g, mg, 2 * i + 1, bar++, accDot, 0.0, sm, &gamma, &unused, remoteA)) break;
        alpha = sigma / (gamma + 1e-8);  // linearsolver.cpp:50
        if (scribe)
        {
            const unsigned long long t = globalTimerNs();
            tA += t - t0;
            t0 = t;
        }
        // K2(i): r -= alpha q; z = M r; sigma' = z.r; err = max|r|
        accDot = 0.0;
        accMax = 0.0;
        const bool remoteB = pipeWalk<MODE_K2, MG>(b0, b1, sm.full, sm.preTbl, g.a, mg, g.r[i & 1], g.q, g.r[(i + 1) & 1], g.z, nullptr, g.loZ, g.hiZ, alpha, 0.0,
                              numTiles, use0, use1, accDot, accMax, 2 * i + 2, (useTensor & 1) ? &tm.m[TM_R0 + (i & 1)] : nullptr, &tm.m[TM_Q],
                              nullptr, early, hB);
        early = canEarly && i + 1 < g.iterLimit;  // s_old and x of K1(i+1), unless this was the last iteration
        if (early && tid == 0)
            pipeIssueEarly<MODE_K1>(*a0, &sm.full[0], &tm.m[TM_S0 + ((i + 1) & 1)], &tm.m[TM_X], g.a.tilesJ, firstTile, hA);
        double sigmaNew = 0.0;
        if (!solveBarrier<MG>(g, mg, 2 * i + 2, bar++, accDot, accMax, sm, &sigmaNew, &err, remoteB)) break;
        executed = i + 1;
        const bool converged = err <= g.a.tol;  // linearsolver.cpp:59-61
        const double betaNew = converged ? 0.0 : sigmaNew / sigma;  // :66-67
        if (scribe)
        {
            const unsigned long long t = globalTimerNs();
            tB += t - t0;
            t0 = t;
            if (g.a.trace && i < g.a.traceCapacity)
            {
                g.a.trace[4 * i + 0] = alpha;
                g.a.trace[4 * i + 1] = betaNew;
                g.a.trace[4 * i + 2] = sigmaNew;
                g.a.trace[4 * i + 3] = err;
            }
        }
        if (converged)
        {
            result = i;
            break;
        }
        beta = betaNew;
        sigma = sigmaNew;
        alphaPrev = alpha;
        if (i + 1 >= g.iterLimit) early = false;
    }
    if (early)
    {
        // left the loop with the early half of a first tile in flight (convergence, or a lost peer): complete the
        // stage before the CTA gives its shared memory back
        if (tid == 0) mbarArrive(&sm.full[0]);
        mbarWait(&sm.full[0], use0 & 1u);
    }
    if (scribe)
    {
        sc->alpha = alpha;  // pending x += alpha s of the last executed iteration (pcgFinalizeKernel)
        sc->beta = beta;
        sc->sigma = sigma;
        sc->gamma = gamma;
        sc->err = err;
        sc->iter = executed;
        sc->result = result;
        sc->done = 1;
        sc->phaseNs[0] += tA;
        sc->phaseNs[1] += tB;
        sc->phaseLaunches += static_cast<unsigned int>(executed);
    }
}

// ------------------------------------------------------------------ resident whole-solve kernel
// When the active set is small -- at most RES_TPC tiles per SM: every grid <= 2048^2 with a dam-break fill, and every
// rank of a 4096^2 run over >= 2 GPUs -- the Krylov vectors never leave the SMs. A CTA (512 threads, one per SM) owns
// up to RES_TPC tiles for the WHOLE solve and keeps, per tile, in shared memory: the halo-extended search vector s and
// residual r (18 x 132 doubles each) and the interior of the vector that travels between the two phases (z into K1, q
// into K2; 16 x 128); the solution x accumulates in registers (4 cells per thread and tile). Per phase a tile then only
//   * writes its 16 x 128 q (K1) or z (K2) to the global array -- the only thing neighbours need --, and
//   * after the barrier reads the 288 ring values of its neighbours' z / q back (L2, ld.cg) to advance s / r on its halo
//     ring redundantly (same inputs, same scalars -> same bits as on the owner; what the slab path already does for the
//     halo rows of a rank),
// i.e. ~18 KB of L2 traffic instead of the 100 KB a streamed tile moves, and no load pipeline to start: a phase is the
// barrier plus ~1 us of shared-memory arithmetic. Arithmetic, barrier protocol (phases, tags, mail ring), halo-row pushes
// into the row neighbours' q / z and convergence logic are those of pcgSolveKernel, so a rank running this kernel
// interoperates with a rank streaming its tiles. x is written once, at the end (pending alpha * s included).
//
// PAGED = true (one GPU, RES_TPC * CTAs < tiles <= RES_PAGED_TPC * CTAs: the 4096^2 dam break has 904 tiles for 148 SMs): a
// CTA keeps RES_TPC - 1 tiles resident as above and treats its other tiles as *paged* resident tiles: the same
// halo-extended private boxes of s and r, but stored in global memory (L2) -- in the s[0] / r[0] arrays, which nothing else
// uses while this kernel runs, box k of the active list at k * PTILE_PAD -- and brought through two shared-memory scratch
// boxes (the storage of the fourth resident tile) by 19 KB bulk copies: cp.async.bulk in (the box of the NEXT paged tile,
// or of the next phase's first one, flies while the CTA computes; mbarrier completion), in-place update and stencil in
// shared memory, cp.async.bulk out (bulk group). The inter-phase values (z into K1, q into K2) are the ones the thread
// itself wrote to the global arrays one phase earlier, x is updated in place in global memory. Same arithmetic, same ring
// protocol, so resident and paged tiles are neighbours without knowing it.
constexpr int RNT = 512;          // threads per CTA
constexpr int RES_TPC = 4;        // tiles a CTA can hold
constexpr int RES_PAGED_TPC = 12; // tiles per CTA (resident + paged) beyond which the streaming kernel takes the solve
constexpr int RES_CELLS = TR * TC / RNT;  // interior cells per thread and tile (4)
constexpr unsigned int RES_BOX_BYTES = PTILE_PAD * 8u;

struct ResTile
{
    double S[PTILE_PAD];
    double R[PTILE_PAD];
    double T[TR * TC];
};

struct ResSmem
{
    ResTile t[RES_TPC];
    double red[32];
    double preTbl[8];
    double pub[2];
    double bc[2];
    unsigned long long full[2];  // paged tiles: arrival of a box in scratch 0 / 1
    int pagedI0[RES_PAGED_TPC], pagedJ0[RES_PAGED_TPC];  // origins of the CTA's paged tiles
    int isLast;
    int ok;
};

__device__ __forceinline__ void blockReduce2W16(double &s, double &m, double *scratch /* >= 32 doubles */)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    s = warpSum(s);
    m = warpMax(m);
    __syncthreads();
    if (lane == 0)
    {
        scratch[warp] = s;
        scratch[16 + warp] = m;
    }
    __syncthreads();
    if (warp == 0)
    {
        s = (lane < (RNT >> 5)) ? scratch[lane] : 0.0;
        m = (lane < (RNT >> 5)) ? scratch[16 + lane] : 0.0;
        s = warpSum(s);
        m = warpMax(m);
    }
}

// (Tried and measured slower: a ticket-less barrier where every CTA publishes its partials as self-validating words and
// thread k of every CTA polls the words of CTA k -- one L2 round trip in theory, but 148 x 148 pollers instead of 148
// turn the poll traffic into the bottleneck: 8.2 -> 15 us per iteration at 1024^2.)
// solveBarrier for `nb` participating CTAs of RNT threads (same protocol; see there).
template <bool MG>
__device__ __forceinline__ bool resBarrier(const SolveArgs &g, const MgArgs &m, int phase, unsigned int barrierIndex, unsigned int nb,
                                           double v0, double v1, ResSmem &sm, double *sum, double *mx, bool remoteStores)
{
    const int tid = threadIdx.x;
    blockReduce2W16(v0, v1, sm.red);
    if (!MG)
    {
        double *part = g.a.partials + (barrierIndex & 1u) * 2u * nb;
        if (tid == 0)
        {
            part[blockIdx.x] = v0;
            part[nb + blockIdx.x] = v1;
            __threadfence();
            atomicAdd(g.ticket, 1u);
            const unsigned int target = (barrierIndex + 1u) * nb;
            unsigned int seen;
            do
            {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(g.ticket) : "memory");
            } while (seen < target);
        }
        __syncthreads();
        double ts = 0.0, tm = 0.0;
        for (unsigned int k = tid; k < nb; k += RNT)
        {
            ts += __ldcg(part + k);
            tm = fmax(tm, __ldcg(part + nb + k));
        }
        blockReduce2W16(ts, tm, sm.red);
        if (tid == 0)
        {
            sm.bc[0] = ts;
            sm.bc[1] = tm;
        }
        __syncthreads();
        *sum = sm.bc[0];
        *mx = sm.bc[1];
        return true;
    }
    // Row slabs: the local half is the one-GPU barrier above -- ticket, and EVERY CTA sums the partials of the rank's CTAs
    // itself (same order everywhere, hence the same bits) -- so no CTA waits for its own rank's result to come back
    // through the mail (the last-CTA reduction + publish + collect chain cost 0.9 us per barrier more, measured with one
    // rank). CTA 0 publishes the rank's sums to the other ranks; warp 0 of every CTA polls the words of the OTHER ranks in
    // its own mail and combines all of them in rank order.
    {
        double *part = g.a.partials + (barrierIndex & 1u) * 2u * nb;
        if (tid == 0)
        {
            part[blockIdx.x] = v0;
            part[nb + blockIdx.x] = v1;
            if (remoteStores)
                __threadfence_system();  // halo rows stored into a neighbour's arrays: acknowledged before the ticket
            else
                __threadfence();
            atomicAdd(g.ticket, 1u);
            const unsigned int target = (barrierIndex + 1u) * nb;
            unsigned int seen;
            do
            {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(g.ticket) : "memory");
            } while (seen < target);
        }
        __syncthreads();
        double ts = 0.0, tm = 0.0;
        for (unsigned int k = tid; k < nb; k += RNT)
        {
            ts += __ldcg(part + k);
            tm = fmax(tm, __ldcg(part + nb + k));
        }
        blockReduce2W16(ts, tm, sm.red);  // thread 0: the rank's sums
        if (blockIdx.x == 0)
        {
            if (tid == 0)
            {
                sm.pub[0] = ts;
                sm.pub[1] = tm;
            }
            __syncthreads();
            llPublish<MG>(m, phase, sm.pub[0], sm.pub[1]);
        }
        if (tid < 32)
        {
            double s = 0.0, x = 0.0;
            const bool ok = llCollectPeers<MG>(m, phase, ts, tm, &s, &x);
            if (tid == 0)
            {
                sm.ok = ok ? 1 : 0;
                sm.bc[0] = s;
                sm.bc[1] = x;
            }
        }
    }
    __syncthreads();
    *sum = sm.bc[0];
    *mx = sm.bc[1];
    return sm.ok != 0;
}

// Ring cell `tid` (< 288) of a tile: position inside the halo-extended 18 x 132 array and the LINEAR global index of the
// value (the reference addresses the j = -1 / j = J neighbours by linear index: pressuredata.h:135-145); n = -1 when the
// value lies outside the vector (reads as zero).
__device__ __forceinline__ void resRingCell(int tid, int i0, int j0, int I, long long J, long long N, int *pos, long long *n)
{
    *pos = -1;
    *n = -1;
    if (tid < 2 * TC)
    {
        const int bottom = tid >= TC ? 1 : 0, col = tid - bottom * TC;
        const long long gi = bottom ? i0 + TR : i0 - 1, gj = j0 + col;
        if (gj < J)
        {
            *pos = (bottom ? TR + 1 : 0) * PSW + col + 2;
            if (gi >= 0 && gi < I) *n = gi * J + gj;
        }
    }
    else if (tid < 2 * TC + 2 * TR)
    {
        const int e = tid - 2 * TC, right = e >= TR ? 1 : 0, ar = 1 + (e - right * TR);
        const long long gi = i0 - 1 + ar;
        const long long span = J - j0;                                        // valid columns of this tile (<= TC for edge tiles)
        const int c = right ? static_cast<int>(span < TC ? span : TC) + 2 : 1;  // right neighbour of the last valid column
        const long long gj = j0 - 2 + c;
        *pos = ar * PSW + c;
        const long long lin = gi * J + gj;
        if (gi < I && lin >= 0 && lin < N) *n = lin;
    }
}

// The same with 32-bit indices (paged tiles: N < 2^31 there), a third of the instructions.
__device__ __forceinline__ void resRingCell32(int tid, int i0, int j0, int I, int J, int N, int *pos, int *n)
{
    *pos = -1;
    *n = -1;
    if (tid < 2 * TC)
    {
        const int bottom = tid >= TC ? 1 : 0, col = tid - bottom * TC;
        const int gi = bottom ? i0 + TR : i0 - 1, gj = j0 + col;
        if (gj < J)
        {
            *pos = (bottom ? TR + 1 : 0) * PSW + col + 2;
            if (gi >= 0 && gi < I) *n = gi * J + gj;
        }
    }
    else if (tid < 2 * TC + 2 * TR)
    {
        const int e = tid - 2 * TC, right = e >= TR ? 1 : 0, ar = 1 + (e - right * TR);
        const int gi = i0 - 1 + ar;
        const int span = J - j0;
        const int c = right ? (span < TC ? span : TC) + 2 : 1;
        const int gj = j0 - 2 + c;
        *pos = ar * PSW + c;
        const long long lin = static_cast<long long>(gi) * J + gj;
        if (gi < I && lin >= 0 && lin < N) *n = static_cast<int>(lin);
    }
}

template <bool MG, bool PAGED> __global__ void __launch_bounds__(RNT, 1) pcgResidentKernel(SolveArgs g, MgArgs mg)
{
    static_assert(!(MG && PAGED), "paged tiles: one GPU only");
    constexpr int RES = PAGED ? RES_TPC - 1 : RES_TPC;  // resident tiles of a CTA (paged mode: the last tile's storage is scratch)
    extern __shared__ __align__(128) unsigned char resRaw[];
    ResSmem &sm = *reinterpret_cast<ResSmem *>(resRaw);
    const int tid = threadIdx.x;
    PcgScalars *sc = g.a.sc;
    const int count = *g.a.activeCount;
    const unsigned int P = static_cast<unsigned int>(min(static_cast<int>(gridDim.x), count));  // participating CTAs
    // too many tiles to hold (or nothing to do): PcgScalars::pad stays 0 and the next kernel in line runs
    if (count <= 0) return;
    if (!PAGED && (count > RES_TPC * static_cast<int>(gridDim.x) || g.a.N >= (1ll << 31))) return;  // 32-bit cell indices below
    if (PAGED && (count <= RES_TPC * static_cast<int>(gridDim.x) || count > RES_PAGED_TPC * static_cast<int>(gridDim.x) ||
                  static_cast<long long>(count) * PTILE_PAD > g.a.N || g.a.N >= (1ll << 31)))
        return;
    if (blockIdx.x >= P) return;
    const bool scribe = blockIdx.x == 0 && tid == 0;
    if (scribe)
    {
        sc->pad = PAGED ? 2 : 1;  // tells the streaming kernel and pcgFinalizeKernel that this solve is done here
        if (MG)
        {
            // which tiles of my first / last tile row are walked (one bit per tile column; more than 64 columns: no LL)
            unsigned long long first = 0, last = 0;
            const int tj = g.a.tilesJ;
            if (tj <= 64 && g.tileFlags)
                for (int k = 0; k < tj; k++)
                {
                    if (g.tileFlags[k]) first |= 1ull << k;
                    if (g.tileFlags[g.numTiles - tj + k]) last |= 1ull << k;
                }
            solveModePublish<MG>(mg, (tj <= 64 && g.tileFlags) ? SOLVE_MODE_RESIDENT : SOLVE_MODE_STREAMING, first, last);
        }
    }
    if (tid < 8) sm.preTbl[tid] = g.a.pre[tid];
    if (tid == 0)
    {
        if (PAGED)
        {
            mbarInit(&sm.full[0], 1);
            mbarInit(&sm.full[1], 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        double s0 = 0.0, m0 = 0.0;
        sm.ok = mgCollect(mg, 0, true, &s0, &m0) ? 1 : 0;  // phase 0: rhs.rhs and max|rhs| from pcgInitKernel<true>
        sm.bc[0] = s0;
        sm.bc[1] = m0;
        sm.isLast = 0;
        if (MG && sm.ok && s0 == s0 && m0 > 1.0e-15 && !(mg.debug & 16))
        {
            // both row neighbours must talk LL for a boundary to use it (bit 0: lower, bit 1: upper)
            const int lo = solveModeOf<MG>(mg, mg.rank - 1), hi = solveModeOf<MG>(mg, mg.rank + 1);
            if (lo < 0 || hi < 0) sm.ok = 0;
            const bool mine = g.a.tilesJ <= 64 && g.tileFlags;
            sm.isLast = (mine && lo == SOLVE_MODE_RESIDENT ? 1 : 0) | (mine && hi == SOLVE_MODE_RESIDENT ? 2 : 0);
            __threadfence_system();
            const int par = mg.ringBase ? 1 : 0;
            // the lower neighbour's LAST tile row feeds my halo row below rowBegin, the upper neighbour's FIRST my row rowEnd
            sm.pub[0] = 0.0;
            sm.pub[1] = 0.0;
            unsigned long long mlo = 0, mhi = 0;
            if (sm.isLast & 1) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(mlo) : "l"(&mg.mail->edgeTiles[par][mg.rank - 1][1]) : "memory");
            if (sm.isLast & 2) asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(mhi) : "l"(&mg.mail->edgeTiles[par][mg.rank + 1][0]) : "memory");
            sm.pub[0] = __longlong_as_double(static_cast<long long>(mlo));
            sm.pub[1] = __longlong_as_double(static_cast<long long>(mhi));
        }
    }
    __syncthreads();
    double sigma = sm.bc[0];
    const double max0 = sm.bc[1];
    const bool lost0 = sm.ok == 0;
    const bool loLL = MG && (sm.isLast & 1), hiLL = MG && (sm.isLast & 2);
    const unsigned long long loTiles = loLL ? static_cast<unsigned long long>(__double_as_longlong(sm.pub[0])) : 0ull;
    const unsigned long long hiTiles = hiLL ? static_cast<unsigned long long>(__double_as_longlong(sm.pub[1])) : 0ull;
    __syncthreads();
    if (lost0 || !(max0 > 1.0e-15))  // a lost peer, or VOps::isZero (vmath.cpp:47-58): x = 0 (pcgInitKernel), zero iterations
    {
        if (scribe)
        {
            sc->sigma = sigma;
            sc->alpha = sc->beta = sc->gamma = sc->err = 0.0;
            sc->iter = 0;
            sc->result = 0;
            sc->done = 1;
        }
        return;
    }
    const int I = g.a.I;
    const long long J = g.a.J, N = g.a.N;
    const int allTiles = (count - static_cast<int>(blockIdx.x) + static_cast<int>(P) - 1) / static_cast<int>(P);
    const int myTiles = min(allTiles, RES);              // resident tiles: list entries blockIdx.x + t * P, t < myTiles
    const int nPaged = PAGED ? allTiles - myTiles : 0;   // paged tiles: list entries blockIdx.x + (RES + k) * P
    const int lr = tid >> 7, lc = tid & (TC - 1);  // this thread's interior cells of a tile: rows lr + 4k, column lc
    double *const scratch0 = sm.t[RES_TPC - 1].S, *const scratch1 = sm.t[RES_TPC - 1].R;  // PAGED only
    double *const boxS = g.s[0], *const boxR = g.r[0];                                    // PAGED only: private boxes of the paged tiles
    unsigned int use0 = 0, use1 = 0;  // completed uses of the two scratch boxes (mbarrier parity)
    int pb = 0;                       // scratch box the next paged tile arrives in

    // ---- per-tile constants in registers: origin, ring cell, operator bytes of the thread's cells
    int ti0[RES], tj0[RES], ringPos[RES];
    int ringSrc[RES];                      // 0: the local q / z array, 1 / 2: LL halo row from the lower / upper neighbour
    int ringN[RES];                        // linear index of the ring value (N < 2^31 here), -1: outside the vector
    unsigned int edge[RES];                // MG: bit k: cell k lies on my first row and is pushed to the lower neighbour; bit 4 + k: last row, upper
    unsigned int rowBits[RES];             // 4 x rowInfo byte
    unsigned int validBits[RES];           // bit k: cell k of this thread lies inside the matrix
    unsigned long long preBits[RES];       // 4 x preInfo half-word
    double xacc[RES][RES_CELLS];
    int cell0[RES];                            // linear index of the thread's first cell of the tile (rows advance by rowStep)
    const int J32 = static_cast<int>(J), N32 = static_cast<int>(N);
    const int rowStep = 4 * J32;
    const int q0 = (lr + 1) * PSW + lc + 2, t0i = lr * TC + lc;  // the same cell inside the halo-extended box / the interior array
    bool remote = false;                       // (uniform over the CTA) a tile holds a slab-boundary row pushed to a neighbour
#pragma unroll
    for (int t = 0; t < RES; t++)
    {
        ti0[t] = tj0[t] = 0;
        ringPos[t] = -1;
        ringSrc[t] = 0;
        ringN[t] = -1;
        edge[t] = 0;
        cell0[t] = 0;
        rowBits[t] = 0;
        preBits[t] = 0;
        validBits[t] = 0;
#pragma unroll
        for (int k = 0; k < RES_CELLS; k++) xacc[t][k] = 0.0;
        if (t < myTiles)
        {
            const int tile = g.a.activeTiles[blockIdx.x + t * P];
            const int ti = tile / g.a.tilesJ, tj = tile - ti * g.a.tilesJ;
            ti0[t] = ti * TR;
            tj0[t] = tj * TC;
            cell0[t] = (ti0[t] + lr) * J32 + tj0[t] + lc;
            resRingCell32(tid, ti0[t], tj0[t], I, J32, N32, &ringPos[t], &ringN[t]);
            if (MG && !(mg.debug & 2))
            {
                // which of the thread's cells sit on a slab-boundary row that is pushed to a row neighbour
#pragma unroll
                for (int k = 0; k < RES_CELLS; k++)
                {
                    const int gi = ti0[t] + lr + 4 * k;
                    if (gi == mg.rowBegin && g.loQ) edge[t] |= 1u << k;
                    if (gi == mg.rowEnd - 1 && g.hiQ) edge[t] |= 16u << k;
                }
            }
            if (MG && ringN[t] >= 0)
            {
                // a ring value that lives on a neighbour's row (wrap columns included: the linear index decides)
                const int srcRow = ringN[t] / J32;
                const int srcTile = (ringN[t] - srcRow * J32) / TC;
                if (loLL && srcRow == mg.rowBegin - 1) ringSrc[t] = 1;
                if (hiLL && srcRow == mg.rowEnd) ringSrc[t] = 2;
                // a tile the neighbour skips pushes nothing: every vector is identically zero there
                if (ringSrc[t] && !(((ringSrc[t] == 1 ? loTiles : hiTiles) >> srcTile) & 1ull))
                {
                    ringSrc[t] = 0;
                    ringN[t] = -1;
                }
            }
            if (MG && !(mg.debug & 2))
                remote = remote || (g.loQ && !loLL && mg.rowBegin >= ti0[t] && mg.rowBegin < ti0[t] + TR) ||
                         (g.hiQ && !hiLL && mg.rowEnd - 1 >= ti0[t] && mg.rowEnd - 1 < ti0[t] + TR);
            // s = 0, r = rhs on the halo-extended tile (r0 = z0 = rhs, linearsolver.cpp:32-46), T = z0 interior
            ResTile &rt = sm.t[t];
            for (int e = tid; e < PTILE; e += RNT)
            {
                const int ar = e / PSW, c = e - ar * PSW;
                const long long gi = ti0[t] - 1 + ar, gj = tj0[t] - 2 + c;
                double v = 0.0;
                if (gi >= 0 && gi < I && gj >= -1 && gj <= J)
                {
                    const long long n = gi * J + gj;
                    if (n >= 0 && n < N) v = g.z[n];  // pcgInitKernel left z = rhs; linear index: wrap columns included
                }
                rt.S[e] = 0.0;
                rt.R[e] = v;
            }
#pragma unroll
            for (int k = 0; k < RES_CELLS; k++)
            {
                const int row = lr + 4 * k;
                const long long gi = ti0[t] + row, gj = tj0[t] + lc;
                double v = 0.0;
                if (gi < I && gj < J)
                {
                    const long long n = gi * J + gj;
                    v = g.z[n];
                    validBits[t] |= 1u << k;
                    rowBits[t] |= static_cast<unsigned int>(g.a.rowInfo[n]) << (8 * k);
                    preBits[t] |= static_cast<unsigned long long>(g.a.preInfo[n]) << (16 * k);
                }
                rt.T[row * TC + lc] = v;
            }
        }
    }
    if (PAGED)
    {
        // private boxes of the paged tiles: s = 0, r = rhs on the halo-extended tile (as above), written with ordinary
        // stores and read by bulk copies from here on
        for (int k = 0; k < nPaged; k++)
        {
            const long long a = static_cast<long long>(blockIdx.x) + static_cast<long long>(RES + k) * P;
            const int tile = g.a.activeTiles[a];
            const int ti = tile / g.a.tilesJ, tj = tile - ti * g.a.tilesJ;
            if (tid == 0)
            {
                sm.pagedI0[k] = ti * TR;
                sm.pagedJ0[k] = tj * TC;
            }
            for (int e = tid; e < PTILE_PAD; e += RNT)
            {
                const int ar = e / PSW, c = e - ar * PSW;
                const long long gi = static_cast<long long>(ti) * TR - 1 + ar, gj = static_cast<long long>(tj) * TC - 2 + c;
                double v = 0.0;
                if (e < PTILE && gi >= 0 && gi < I && gj >= -1 && gj <= J)
                {
                    const long long n = gi * J + gj;
                    if (n >= 0 && n < N) v = g.z[n];
                }
                __stcg(boxS + a * PTILE_PAD + e, 0.0);
                __stcg(boxR + a * PTILE_PAD + e, v);
            }
        }
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
    bool inFlight = false;  // a box is on its way into scratch `pb` (uniform over the CTA)
    if (PAGED && nPaged > 0)
    {
        if (tid == 0)
        {
            asm volatile("fence.proxy.async;" ::: "memory");
            mbarArriveExpectTx(&sm.full[0], RES_BOX_BYTES);
            bulkLoad(scratch0, boxS + (static_cast<long long>(blockIdx.x) + static_cast<long long>(RES) * P) * PTILE_PAD, RES_BOX_BYTES, &sm.full[0]);
        }
        inFlight = true;
    }

    unsigned int bar = 0;
    double alpha = 0.0, beta = 0.0, alphaPrev = 0.0, gamma = 0.0, err = 0.0;
    int result = g.iterLimit, executed = 0;
    unsigned long long tA = 0, tB = 0, t0 = scribe ? globalTimerNs() : 0ull;
    for (int i = 0; i < g.iterLimit; i++)
    {
        // ---- K1(i): s = z + beta s (interior from T, ring from the neighbours' z); x += alpha_{i-1} s_{i-1}; q = A s; gamma = q.s
        double accDot = 0.0, accMax = 0.0, unused = 0.0;
        if (scribe) mgStampAt(mg, 2 * i + 1, 0);
        double ringV[RES];
#pragma unroll
        for (int t = 0; t < RES; t++)
        {
            ringV[t] = 0.0;
            if (t < myTiles && ringN[t] >= 0)
            {
                // z of K2(i - 1) = phase 2i; before the first iteration z = rhs, which every rank computed for its halo rows itself
                if (MG && ringSrc[t] && i > 0)
                    ringV[t] = llHaloLoad(mg.haloLL, ringSrc[t] - 1, 1, J, ringN[t] % J32, llTag(mg, 2 * i), &mg.mail->error);
                else
                    ringV[t] = __ldcg(g.z + ringN[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < RES; t++)
            if (t < myTiles)
            {
                ResTile &rt = sm.t[t];
#pragma unroll
                for (int k = 0; k < RES_CELLS; k++)
                {
                    // cells outside the matrix stay zero; in an edge tile the slot right of the last valid column is the
                    // ring cell (wrap neighbour) and belongs to the ring thread
                    if (!((validBits[t] >> k) & 1u)) continue;
                    const int p = q0 + k * 4 * PSW;
                    const double so = rt.S[p];
                    rt.S[p] = __dadd_rn(rt.T[t0i + k * 4 * TC], __dmul_rn(so, beta));
                    xacc[t][k] = __dadd_rn(xacc[t][k], __dmul_rn(so, alphaPrev));
                }
                if (ringPos[t] >= 0) rt.S[ringPos[t]] = __dadd_rn(ringV[t], __dmul_rn(rt.S[ringPos[t]], beta));
            }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < RES; t++)
            if (t < myTiles)
            {
                ResTile &rt = sm.t[t];
#pragma unroll
                for (int k = 0; k < RES_CELLS; k++)
                {
                    const int p = q0 + k * 4 * PSW;
                    if ((validBits[t] >> k) & 1u)
                    {
                        const int n = cell0[t] + k * rowStep;
                        const double c = rt.S[p];
                        const double o = rowA(static_cast<uint8_t>(rowBits[t] >> (8 * k)), g.a.scale, c, rt.S[p - PSW], rt.S[p + PSW], rt.S[p - 1], rt.S[p + 1]);
                        rt.T[t0i + k * 4 * TC] = o;
                        g.q[n] = o;
                        if (MG && edge[t])
                        {
                            // my first row is the lower neighbour's halo row "from above" (its side 1), my last row the
                            // upper neighbour's halo row "from below" (its side 0)
                            if ((edge[t] >> k) & 1u)
                            {
                                if (loLL)
                                    llHaloStore(mg.loHaloLL, 1, 0, J, tj0[t] + lc, o, llTag(mg, 2 * i + 1));
                                else
                                    g.loQ[n] = o;
                            }
                            if ((edge[t] >> (4 + k)) & 1u)
                            {
                                if (hiLL)
                                    llHaloStore(mg.hiHaloLL, 0, 0, J, tj0[t] + lc, o, llTag(mg, 2 * i + 1));
                                else
                                    g.hiQ[n] = o;
                            }
                        }
                        accDot += o * c;
                    }
                }
            }
        if (scribe) mgStampAt(mg, 2 * i + 1, 1);
        if (PAGED)
        {
            // 32-bit indices and kernel-invariant offsets throughout: the paged tiles are bound by instruction issue
            const int J32 = static_cast<int>(J), N32 = static_cast<int>(N), rowStep = 4 * J32;
            const int q0 = (lr + 1) * PSW + lc + 2;
            for (int k = 0; k < nPaged; k++)
            {
                const long long a = static_cast<long long>(blockIdx.x) + static_cast<long long>(RES + k) * P;
                const int i0 = sm.pagedI0[k], j0 = sm.pagedJ0[k];
                double *box = pb ? scratch1 : scratch0;
                // operands from global memory: z and x of the thread's cells, the ring value of the neighbours' z
                const int row0 = i0 + lr, col = j0 + lc;
                const int n0 = row0 * J32 + col;
                const bool colOk = col < J32;
                double zv[RES_CELLS], xv[RES_CELLS];
                unsigned int info[RES_CELLS];
#pragma unroll
                for (int c = 0; c < RES_CELLS; c++)
                {
                    const bool ok = colOk && row0 + 4 * c < I;
                    const int n = n0 + c * rowStep;
                    zv[c] = ok ? __ldcg(g.z + n) : 0.0;
                    xv[c] = ok ? __ldcg(g.x + n) : 0.0;
                    info[c] = ok ? static_cast<unsigned int>(g.a.rowInfo[n]) : 0u;
                }
                int rpos, rn;
                resRingCell32(tid, i0, j0, I, J32, N32, &rpos, &rn);
                const double rv = rn >= 0 ? __ldcg(g.z + rn) : 0.0;
                if (tid == 0 && k + 1 < nPaged)
                {
                    // the next paged tile's box flies into the other scratch while this one is worked on; that scratch
                    // was the source of the previous tile's write-back
                    bulkWaitRead0();
                    mbarArriveExpectTx(&sm.full[pb ^ 1], RES_BOX_BYTES);
                    bulkLoad(pb ? scratch0 : scratch1, boxS + (a + P) * PTILE_PAD, RES_BOX_BYTES, &sm.full[pb ^ 1]);
                }
                mbarWait(&sm.full[pb], (pb ? use1 : use0) & 1u);
                if (pb) use1++; else use0++;
#pragma unroll
                for (int c = 0; c < RES_CELLS; c++)
                    if (colOk && row0 + 4 * c < I)
                    {
                        const int q = q0 + c * 4 * PSW;
                        const double so = box[q];
                        box[q] = __dadd_rn(zv[c], __dmul_rn(so, beta));
                        g.x[n0 + c * rowStep] = __dadd_rn(xv[c], __dmul_rn(so, alphaPrev));
                    }
                if (rpos >= 0) box[rpos] = __dadd_rn(rv, __dmul_rn(box[rpos], beta));
                __syncthreads();
#pragma unroll
                for (int c = 0; c < RES_CELLS; c++)
                    if (colOk && row0 + 4 * c < I)
                    {
                        const int q = q0 + c * 4 * PSW;
                        const double cv = box[q];
                        const double o = rowA(static_cast<uint8_t>(info[c]), g.a.scale, cv, box[q - PSW], box[q + PSW], box[q - 1], box[q + 1]);
                        g.q[n0 + c * rowStep] = o;
                        accDot += o * cv;
                    }
                fenceProxyAsync();
                __syncthreads();
                if (tid == 0) bulkStore(boxS + a * PTILE_PAD, box, RES_BOX_BYTES);
                pb ^= 1;
            }
            if (nPaged > 0)
            {
                // the r box of the first paged tile, for K2(i), travels during the barrier
                if (tid == 0)
                {
                    bulkWaitAllButLast();
                    mbarArriveExpectTx(&sm.full[pb], RES_BOX_BYTES);
                    bulkLoad(pb ? scratch1 : scratch0, boxR + (static_cast<long long>(blockIdx.x) + static_cast<long long>(RES) * P) * PTILE_PAD,
                             RES_BOX_BYTES, &sm.full[pb]);
                }
                inFlight = true;
            }
        }
        if (scribe) mgStampAt(mg, 2 * i + 1, 2);
        if (!resBarrier<MG>(g, mg, 2 * i + 1, bar++, P, accDot, 0.0, sm, &gamma, &unused, remote)) break;
        if (scribe) mgStampAt(mg, 2 * i + 1, 3);
        alpha = sigma / (gamma + 1e-8);  // linearsolver.cpp:50
        if (scribe)
        {
            const unsigned long long t = globalTimerNs();
            tA += t - t0;
            t0 = t;
        }
        // ---- K2(i): r -= alpha q (interior from T, ring from the neighbours' q); z = M r; sigma' = z.r; err = max|r|
        accDot = 0.0;
        if (scribe) mgStampAt(mg, 2 * i + 2, 0);
#pragma unroll
        for (int t = 0; t < RES; t++)
        {
            ringV[t] = 0.0;
            if (t < myTiles && ringN[t] >= 0)
            {
                if (MG && ringSrc[t])
                    ringV[t] = llHaloLoad(mg.haloLL, ringSrc[t] - 1, 0, J, ringN[t] % J32, llTag(mg, 2 * i + 1), &mg.mail->error);
                else
                    ringV[t] = __ldcg(g.q + ringN[t]);
            }
        }
#pragma unroll
        for (int t = 0; t < RES; t++)
            if (t < myTiles)
            {
                ResTile &rt = sm.t[t];
#pragma unroll
                for (int k = 0; k < RES_CELLS; k++)
                {
                    if (!((validBits[t] >> k) & 1u)) continue;
                    const int p = q0 + k * 4 * PSW;
                    rt.R[p] = __dsub_rn(rt.R[p], __dmul_rn(rt.T[t0i + k * 4 * TC], alpha));
                }
                if (ringPos[t] >= 0) rt.R[ringPos[t]] = __dsub_rn(rt.R[ringPos[t]], __dmul_rn(ringV[t], alpha));
            }
        __syncthreads();
#pragma unroll
        for (int t = 0; t < RES; t++)
            if (t < myTiles)
            {
                ResTile &rt = sm.t[t];
#pragma unroll
                for (int k = 0; k < RES_CELLS; k++)
                {
                    const int p = q0 + k * 4 * PSW;
                    if ((validBits[t] >> k) & 1u)
                    {
                        const int n = cell0[t] + k * rowStep;
                        const double c = rt.R[p];
                        const double o = rowM(static_cast<uint16_t>(preBits[t] >> (16 * k)), sm.preTbl, c, rt.R[p - PSW], rt.R[p + PSW], rt.R[p - 1], rt.R[p + 1]);
                        rt.T[t0i + k * 4 * TC] = o;
                        g.z[n] = o;
                        if (MG && edge[t])
                        {
                            if ((edge[t] >> k) & 1u)
                            {
                                if (loLL)
                                    llHaloStore(mg.loHaloLL, 1, 1, J, tj0[t] + lc, o, llTag(mg, 2 * i + 2));
                                else
                                    g.loZ[n] = o;
                            }
                            if ((edge[t] >> (4 + k)) & 1u)
                            {
                                if (hiLL)
                                    llHaloStore(mg.hiHaloLL, 0, 1, J, tj0[t] + lc, o, llTag(mg, 2 * i + 2));
                                else
                                    g.hiZ[n] = o;
                            }
                        }
                        accDot += o * c;
                        accMax = fmax(accMax, fabs(c));
                    }
                }
            }
        if (scribe) mgStampAt(mg, 2 * i + 2, 1);
        if (PAGED)
        {
            const int J32 = static_cast<int>(J), N32 = static_cast<int>(N), rowStep = 4 * J32;
            const int q0 = (lr + 1) * PSW + lc + 2;
            for (int k = 0; k < nPaged; k++)
            {
                const long long a = static_cast<long long>(blockIdx.x) + static_cast<long long>(RES + k) * P;
                const int i0 = sm.pagedI0[k], j0 = sm.pagedJ0[k];
                double *box = pb ? scratch1 : scratch0;
                const int row0 = i0 + lr, col = j0 + lc;
                const int n0 = row0 * J32 + col;
                const bool colOk = col < J32;
                double qv[RES_CELLS];
                unsigned int info[RES_CELLS];
#pragma unroll
                for (int c = 0; c < RES_CELLS; c++)
                {
                    const bool ok = colOk && row0 + 4 * c < I;
                    const int n = n0 + c * rowStep;
                    qv[c] = ok ? __ldcg(g.q + n) : 0.0;
                    info[c] = ok ? static_cast<unsigned int>(g.a.preInfo[n]) : 0u;
                }
                int rpos, rn;
                resRingCell32(tid, i0, j0, I, J32, N32, &rpos, &rn);
                const double rv = rn >= 0 ? __ldcg(g.q + rn) : 0.0;
                if (tid == 0 && k + 1 < nPaged)
                {
                    bulkWaitRead0();
                    mbarArriveExpectTx(&sm.full[pb ^ 1], RES_BOX_BYTES);
                    bulkLoad(pb ? scratch0 : scratch1, boxR + (a + P) * PTILE_PAD, RES_BOX_BYTES, &sm.full[pb ^ 1]);
                }
                mbarWait(&sm.full[pb], (pb ? use1 : use0) & 1u);
                if (pb) use1++; else use0++;
#pragma unroll
                for (int c = 0; c < RES_CELLS; c++)
                    if (colOk && row0 + 4 * c < I)
                    {
                        const int q = q0 + c * 4 * PSW;
                        box[q] = __dsub_rn(box[q], __dmul_rn(qv[c], alpha));
                    }
                if (rpos >= 0) box[rpos] = __dsub_rn(box[rpos], __dmul_rn(rv, alpha));
                __syncthreads();
#pragma unroll
                for (int c = 0; c < RES_CELLS; c++)
                    if (colOk && row0 + 4 * c < I)
                    {
                        const int q = q0 + c * 4 * PSW;
                        const double cv = box[q];
                        const double o = rowM(static_cast<uint16_t>(info[c]), sm.preTbl, cv, box[q - PSW], box[q + PSW], box[q - 1], box[q + 1]);
                        g.z[n0 + c * rowStep] = o;
                        accDot += o * cv;
                        accMax = fmax(accMax, fabs(cv));
                    }
                fenceProxyAsync();
                __syncthreads();
                if (tid == 0) bulkStore(boxR + a * PTILE_PAD, box, RES_BOX_BYTES);
                pb ^= 1;
            }
            if (nPaged > 0)
            {
                // the s box of the first paged tile, for K1(i + 1) (drained after the loop if there is none)
                if (tid == 0)
                {
                    bulkWaitAllButLast();
                    mbarArriveExpectTx(&sm.full[pb], RES_BOX_BYTES);
                    bulkLoad(pb ? scratch1 : scratch0, boxS + (static_cast<long long>(blockIdx.x) + static_cast<long long>(RES) * P) * PTILE_PAD,
                             RES_BOX_BYTES, &sm.full[pb]);
                }
                inFlight = true;
            }
        }
        double sigmaNew = 0.0;
        if (scribe) mgStampAt(mg, 2 * i + 2, 2);
        if (!resBarrier<MG>(g, mg, 2 * i + 2, bar++, P, accDot, accMax, sm, &sigmaNew, &err, remote)) break;
        if (scribe) mgStampAt(mg, 2 * i + 2, 3);
        executed = i + 1;
        const bool converged = err <= g.a.tol;                      // linearsolver.cpp:59-61
        const double betaNew = converged ? 0.0 : sigmaNew / sigma;  // :66-67
        if (scribe)
        {
            const unsigned long long t = globalTimerNs();
            tB += t - t0;
            t0 = t;
            if (g.a.trace && i < g.a.traceCapacity)
            {
                g.a.trace[4 * i + 0] = alpha;
                g.a.trace[4 * i + 1] = betaNew;
                g.a.trace[4 * i + 2] = sigmaNew;
                g.a.trace[4 * i + 3] = err;
            }
        }
        if (converged)
        {
            result = i;
            break;
        }
        beta = betaNew;
        sigma = sigmaNew;
        alphaPrev = alpha;
    }
    if (PAGED && nPaged > 0)
    {
        // the box requested for a phase that never came lands before the CTA gives its shared memory back; all write-backs
        // complete; then the pending alpha * s of the paged tiles, from their private s boxes
        if (inFlight) mbarWait(&sm.full[pb], (pb ? use1 : use0) & 1u);
        if (tid == 0) bulkWaitAll();
        __syncthreads();
        asm volatile("fence.proxy.async;" ::: "memory");
        if (executed > 0)
            for (int k = 0; k < nPaged; k++)
            {
                const long long a = static_cast<long long>(blockIdx.x) + static_cast<long long>(RES + k) * P;
                const int tile = g.a.activeTiles[a];
                const int ti = tile / g.a.tilesJ, tj = tile - ti * g.a.tilesJ;
#pragma unroll
                for (int c = 0; c < RES_CELLS; c++)
                {
                    const long long gi = static_cast<long long>(ti) * TR + lr + 4 * c, gj = static_cast<long long>(tj) * TC + lc;
                    if (gi < I && gj < J)
                    {
                        const double sv = __ldcg(boxS + a * PTILE_PAD + (lr + 4 * c + 1) * PSW + lc + 2);
                        g.x[gi * J + gj] = __dadd_rn(__ldcg(g.x + gi * J + gj), __dmul_rn(sv, alpha));
                    }
                }
            }
    }
    // ---- x = accumulated steps + the pending alpha * s of the last executed iteration (linearsolver.cpp:51)
#pragma unroll
    for (int t = 0; t < RES; t++)
        if (t < myTiles)
        {
            ResTile &rt = sm.t[t];
#pragma unroll
            for (int k = 0; k < RES_CELLS; k++)
            {
                const int p = q0 + k * 4 * PSW;
                if ((validBits[t] >> k) & 1u)
                {
                    double xv = xacc[t][k];
                    if (executed > 0) xv = __dadd_rn(xv, __dmul_rn(rt.S[p], alpha));
                    g.x[cell0[t] + k * rowStep] = xv;
                }
            }
        }
    if (scribe)
    {
        sc->alpha = alpha;
        sc->beta = beta;
        sc->sigma = sigma;
        sc->gamma = gamma;
        sc->err = err;
        sc->iter = executed;
        sc->result = result;
        sc->done = 1;
        sc->phaseNs[0] += tA;
        sc->phaseNs[1] += tB;
        sc->phaseLaunches += static_cast<unsigned int>(executed);
    }
}

// Reference-compatible convergence value (vmath.cpp:100-136): for each ThreadPool range
// the |r| of its LAST non-zero element (DBL_MIN when the range is all zero), then the max
// over ranges. One CTA per range scans backwards and stops at the first chunk with a hit.
__global__ void __launch_bounds__(NT) pcgRangeErrKernel(const double *r, long long N, int T, double *rangeVals,
                                                        PcgScalars *sc, double tol, double *trace, int traceCapacity)
{
    __shared__ long long best;
    __shared__ int isLast;
    if (sc->done) return;
    const int t = blockIdx.x;
    // splitRange (threadpool.cpp:41-76) for N >= T: partSize = N/T, first N%T ranges one longer
    const long long part = N / T, rem = N - part * T;
    const long long start = t * part + (t < rem ? t : rem);
    const long long end = start + part + (t < rem ? 1 : 0);
    if (threadIdx.x == 0) best = -1;
    __syncthreads();
    for (long long hi = end; hi > start; hi -= 4 * NT)
    {
        long long local = -1;
#pragma unroll
        for (int u = 0; u < 4; u++)
        {
            const long long idx = hi - 1 - (threadIdx.x + static_cast<long long>(u) * NT);
            if (idx >= start && r[idx] != 0.0 && idx > local) local = idx;
        }
        for (int o = 16; o > 0; o >>= 1)
        {
            long long other = __shfl_down_sync(0xffffffffu, local, o);
            local = other > local ? other : local;
        }
        if ((threadIdx.x & 31) == 0 && local >= 0) atomicMax(&best, local);
        __syncthreads();
        if (best >= 0) break;
        __syncthreads();
    }
    if (threadIdx.x == 0)
    {
        rangeVals[t] = best >= 0 ? fabs(r[best]) : 2.2250738585072014e-308;
        __threadfence();
        isLast = (atomicAdd(&sc->ticketC, 1u) == static_cast<unsigned int>(T - 1));
    }
    __syncthreads();
    if (!isLast || threadIdx.x != 0) return;
    __threadfence();
    double err = 0.0;
    for (int k = 0; k < T; k++)
    {
        const double v = __ldcg(rangeVals + k);
        if (v > err) err = v;
    }
    sc->ticketC = 0;
    const int it = sc->iter;
    const double sigmaNew = sc->gamma;
    double beta = 0.0;
    const double trueMax = sc->err;
    sc->err = err;
    if (err <= tol)
    {
        sc->done = 1;
        sc->result = it;
    }
    else
    {
        beta = sigmaNew / sc->sigma;
        sc->beta = beta;
        sc->sigma = sigmaNew;
    }
    if (trace && it < traceCapacity)
    {
        trace[4 * it + 0] = sc->alpha;
        trace[4 * it + 1] = beta;
        trace[4 * it + 2] = sigmaNew;
        trace[4 * it + 3] = err;
    }
    (void)trueMax;
    sc->iter = it + 1;
}

// ------------------------------------------------------------------ active tiles
// A tile whose cells have no matrix row and a zero right-hand side stays identically zero in every PCG
// vector for the whole solve (identity rows: q = s, z = r; r0 = rhs = 0), and contributes nothing to the
// dot products or the max. Such tiles are skipped: the iteration kernels walk an ordered list of the
// other ("active") tiles. Halo rows read from a skipped tile are the zeros pcgInitKernel put there.
// In a dam-break scene ~9 % of the cells are fluid, so this removes ~90 % of the PCG traffic; the
// iterates are the same numbers (only the grouping of the dot-product partials over CTAs changes).
__global__ void __launch_bounds__(NT) pcgTileFlagKernel(const uint8_t *__restrict__ rowInfo, const double *__restrict__ rhs, int I, int J,
                                                        int tilesJ, int tileBase, int *__restrict__ flags)
{
    __shared__ int any;
    if (threadIdx.x == 0) any = 0;
    __syncthreads();
    const int tile = tileBase + blockIdx.x;
    const int ti = tile / tilesJ, tj = tile - ti * tilesJ;
    const int i0 = ti * TR, j0 = tj * TC;
    bool hit = false;
    for (int e = threadIdx.x; e < TR * TC; e += NT)
    {
        const int i = i0 + e / TC, j = j0 + e % TC;
        if (i < I && j < J)
        {
            const long long n = static_cast<long long>(i) * J + j;
            if ((rowInfo[n] & FS2D_ROW_UNIT) || rhs[n] != 0.0) hit = true;
        }
    }
    if (hit) any = 1;
    __syncthreads();
    if (threadIdx.x == 0) flags[blockIdx.x] = any;
}

// Ordered compaction of the flagged tiles (a single CTA: there are only a few thousand tiles).
__global__ void __launch_bounds__(1024) pcgTileCompactKernel(const int *__restrict__ flags, int tiles, int tileBase,
                                                             int *__restrict__ list, int *__restrict__ count)
{
    __shared__ int warpSums[32];
    __shared__ int carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < tiles; base += 1024)
    {
        const int idx = base + threadIdx.x;
        const int v = idx < tiles ? flags[idx] : 0;
        int incl = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1)
        {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) warpSums[warp] = incl;
        __syncthreads();
        if (warp == 0)
        {
            int w = warpSums[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1)
            {
                const int t = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += t;
            }
            warpSums[lane] = w;
        }
        __syncthreads();
        const int excl = incl - v + (warp > 0 ? warpSums[warp - 1] : 0) + carry;
        if (v) list[excl] = tileBase + idx;
        __syncthreads();
        if (threadIdx.x == 1023) carry = excl + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = carry;
}

// result = 0; residual = aux = rhs; search = aux after the first K1 (linearsolver.cpp:32-46);
// sigma = rhs.rhs; zero test of :33-35.
// Slab mode: the walk covers the owned rows plus one halo row each side [nLo, nHi) (the halo rows of r0 = z = rhs
// come from the locally computed halo of the right-hand side); only owned cells [oLo, oHi) enter the sums, and
// the last CTA publishes them as phase 0 instead of writing the scalars.
// residentLimit > 0: the tile list was built BEFORE this kernel and a list of at most that many tiles will be taken by
// pcgResidentKernel, which never reads r0 / s0 / r1 / s1 (its vectors live in shared memory, the paged boxes are
// initialised by the kernel itself): those four passes (537 MB at 4096^2, 0.11 ms) are skipped.
template <bool MG>
__global__ void __launch_bounds__(NT) pcgInitKernel(const double *rhs, double *x, double *r0, double *z, double *s0, double *r1,
                                                    double *s1, double *q, long long N, double *partials, PcgScalars *sc,
                                                    long long nLo, long long oLo, long long oHi, MgArgs mg,
                                                    const int *activeCount, int residentLimit)
{
    __shared__ double red[8];
    __shared__ int isLast;
    double acc = 0.0, amax = 0.0;
    const bool light = residentLimit > 0 && activeCount && *activeCount > 0 && *activeCount <= residentLimit;
    for (long long n = nLo + blockIdx.x * static_cast<long long>(NT) + threadIdx.x; n < N; n += static_cast<long long>(gridDim.x) * NT)
    {
        const double v = rhs[n];
        x[n] = 0.0;
        z[n] = v;
        if (!light)
        {
            r0[n] = v;
            s0[n] = 0.0;
        }
        if (r1)  // active-tile mode: skipped tiles are never written again and must read as zero
        {
            if (!light)
            {
                r1[n] = 0.0;
                s1[n] = 0.0;
            }
            q[n] = 0.0;
        }
        if (MG && (n < oLo || n >= oHi)) continue;
        acc += v * v;
        amax = fmax(amax, fabs(v));
    }
    const int nb = gridDim.x;
    double bs = blockReduce<false>(acc, red);
    double bm = blockReduce<true>(amax, red);
    if (threadIdx.x == 0)
    {
        partials[blockIdx.x] = bs;
        partials[nb + blockIdx.x] = bm;
        __threadfence();
        isLast = (atomicAdd(&sc->ticketA, 1u) == static_cast<unsigned int>(nb - 1));
    }
    __syncthreads();
    if (!isLast) return;
    __threadfence();
    double total = finalReduce<false>(partials, nb, red);
    double emax = finalReduce<true>(partials + nb, nb, red);
    if (MG)
    {
        __shared__ double pub[2];
        if (threadIdx.x == 0)
        {
            pub[0] = total;
            pub[1] = emax;
            sc->done = 0;
            sc->iter = 0;
            sc->result = 0;
            sc->ticketA = 0;
            sc->ticketB = 0;
            sc->ticketC = 0;
        }
        __syncthreads();
        mgPublish(mg, pub[0], pub[1]);
        return;
    }
    if (threadIdx.x == 0)
    {
        sc->sigma = total;
        sc->alpha = 0.0;
        sc->beta = 0.0;
        sc->gamma = 0.0;
        sc->err = 0.0;
        sc->iter = 0;
        sc->result = 0;
        sc->done = (emax > 1.0e-15) ? 0 : 1;  // VOps::isZero (vmath.cpp:47-58)
        sc->ticketA = 0;
        sc->ticketB = 0;
        sc->ticketC = 0;
    }
}

// What a light pcgInitKernel left out, for the (rare) case that the resident kernel could not be launched after all.
__global__ void __launch_bounds__(NT) pcgInitRestKernel(const double *rhs, double *r0, double *s0, double *r1, double *s1, long long nLo, long long N)
{
    for (long long n = nLo + blockIdx.x * static_cast<long long>(NT) + threadIdx.x; n < N; n += static_cast<long long>(gridDim.x) * NT)
    {
        r0[n] = rhs[n];
        s0[n] = 0.0;
        r1[n] = 0.0;
        s1[n] = 0.0;
    }
}

// Pending x += alpha*s of the last executed iteration, and the return value.
__global__ void __launch_bounds__(NT) pcgFinalizeKernel(double *x, const double *sEven, const double *sOdd, long long nLo, long long N,
                                                        PcgScalars *sc, int iterLimit)
{
    const int iters = sc->iter;
    if (iters > 0 && !sc->pad)  // pad: the resident kernel wrote the finished x itself
    {
        const double alpha = sc->alpha;
        const double *s = (iters & 1) ? sOdd : sEven;
        for (long long n = nLo + blockIdx.x * static_cast<long long>(NT) + threadIdx.x; n < N;
             n += static_cast<long long>(gridDim.x) * NT)
            x[n] = __dadd_rn(x[n], __dmul_rn(s[n], alpha));
    }
}

__global__ void pcgResultKernel(PcgScalars *sc, int iterLimit)
{
    if (!sc->done) sc->result = iterLimit;
    sc->done = 1;
}

// Slab mode, after the last K2: wait for its partials (one thread, so that no grid-sized kernel ever spins), close
// the bookkeeping of the last iteration and leave alpha / iter for the pending x update of pcgFinalizeKernel.
__global__ void pcgMgCloseKernel(PcgArgs a, MgArgs mg)
{
    PcgScalars *sc = a.sc;
    const int L = mg.iterLimit - 1;
    if (!sc->done && mg.iterLimit > 0)
    {
        double sA = 0.0, mA = 0.0, sB = 0.0, sC = 0.0, dummy = 0.0;
        mgCollect(mg, 2 * L + 2, true, &sA, &mA);
        mgCollect(mg, 2 * L + 1, false, &sB, &dummy);
        mgCollect(mg, 2 * L, false, &sC, &dummy);
        const double alpha = sC / (sB + 1e-8);
        double beta = 0.0;
        sc->alpha = alpha;
        sc->gamma = sA;
        sc->err = mA;
        if (mA <= a.tol)
            sc->result = L;
        else
        {
            sc->result = mg.iterLimit;
            beta = sA / sC;
            sc->beta = beta;
            sc->sigma = sA;
        }
        if (a.trace && L < a.traceCapacity)
        {
            a.trace[4 * L + 0] = alpha;
            a.trace[4 * L + 1] = beta;
            a.trace[4 * L + 2] = sA;
            a.trace[4 * L + 3] = mA;
        }
        sc->iter = mg.iterLimit;
    }
    sc->done = 1;
}

PcgArgs baseArgs(Ctx *ctx)
{
    PcgArgs a;
    a.I = ctx->I;
    a.J = ctx->J;
    a.N = ctx->N;
    a.tilesJ = divUp(ctx->J, TC);
    a.scale = ctx->matrixScale;
    for (int k = 0; k < 8; k++) a.pre[k] = 1.0 / (static_cast<double>(k) * ctx->matrixScale);
    a.rowInfo = ctx->rowInfo;
    a.preInfo = ctx->preInfo;
    a.in0 = a.in1 = nullptr;
    a.out0 = a.out1 = a.x = nullptr;
    a.partials = ctx->partials;
    a.sc = ctx->scalars;
    a.trace = ctx->trace;
    a.traceCapacity = ctx->traceCapacity;
    a.tol = 0.0;
    a.compat = ctx->p.convergence_threads > 0 ? 1 : 0;
    a.activeTiles = nullptr;
    a.activeCount = nullptr;
    return a;
}
}  // namespace

// Force the module loader to resolve the kernels that only run in slab mode (see capi.cu on lazy loading).
void pcgPreloadSlabKernels()
{
    cudaFuncAttributes at;
    cudaFuncGetAttributes(&at, pcgInitKernel<true>);
    cudaFuncGetAttributes(&at, pcgPipeKernel<MODE_K1, true>);
    cudaFuncGetAttributes(&at, pcgPipeKernel<MODE_K2, true>);
    cudaFuncGetAttributes(&at, pcgMgCloseKernel);
    cudaFuncGetAttributes(&at, pcgSolveKernel<true>);
    cudaFuncGetAttributes(&at, pcgSolveKernel<false>);
    cudaFuncGetAttributes(&at, pcgResidentKernel<true, false>);
    cudaFuncGetAttributes(&at, pcgResidentKernel<false, false>);
    cudaFuncGetAttributes(&at, pcgResidentKernel<false, true>);
    cudaFuncGetAttributes(&at, pcgFinalizeKernel);
    cudaFuncGetAttributes(&at, pcgTileFlagKernel);
    cudaFuncGetAttributes(&at, pcgTileCompactKernel);
    cudaGetLastError();
}

// Tensor maps for the whole-solve kernel, encoded once per handle (the vectors never move). The driver entry point is
// fetched through the runtime, so the library still links against cudart only.
static bool buildSolveMaps(Ctx *ctx, SolveMaps *out)
{
    typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                 const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn ||
        qres != cudaDriverEntryPointSuccess)
    {
        cudaGetLastError();
        return false;
    }
    EncodeFn encode = reinterpret_cast<EncodeFn>(fn);
    double *vec[TM_COUNT] = {ctx->z, ctx->q, ctx->s[0], ctx->s[1], ctx->r[0], ctx->r[1], ctx->x};
    const cuuint64_t dims[2] = {static_cast<cuuint64_t>(ctx->J), static_cast<cuuint64_t>(ctx->I)};
    const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ctx->J) * sizeof(double)};
    const cuuint32_t elem[2] = {1, 1};
    const int promo = std::getenv("FS2D_PCG_L2PROMO") ? std::atoi(std::getenv("FS2D_PCG_L2PROMO")) : 256;
    const CUtensorMapL2promotion l2promo = promo >= 256 ? CU_TENSOR_MAP_L2_PROMOTION_L2_256B
                                           : promo >= 128 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B
                                           : promo >= 64  ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
                                                          : CU_TENSOR_MAP_L2_PROMOTION_NONE;
    for (int k = 0; k < TM_COUNT; k++)
    {
        const cuuint32_t box[2] = {static_cast<cuuint32_t>(k == TM_X ? TC : PSW), static_cast<cuuint32_t>(k == TM_X ? TR : PROWS)};
        if (encode(&out->m[k], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, vec[k], dims, strides, box, elem, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, l2promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return false;
    }
    return true;
}

// Dynamic shared-memory ceilings of the PCG kernels: set ONCE per device to the largest request any launch makes and never
// lowered. (Setting them per launch raced between the host threads of ranks that share a process in the tests: one
// thread lowered the ceiling between another thread's set and its cooperative launch -> "too many blocks in
// cooperative launch".)
constexpr size_t PCG_EXCLUSIVE_SMEM = 160 * 1024;
static void pcgKernelAttributes(int device)
{
    static std::mutex lock;
    static unsigned long long done = 0;
    std::lock_guard<std::mutex> guard(lock);
    if (device < 0 || device >= 64 || (done >> device) & 1ull) return;
    const int ceiling = static_cast<int>(std::max({sizeof(SolveSmem), sizeof(PipeSmem<MODE_K1>), sizeof(PipeSmem<MODE_K2>), PCG_EXCLUSIVE_SMEM}));
    cudaFuncSetAttribute(pcgPipeKernel<MODE_K1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ceiling);
    cudaFuncSetAttribute(pcgPipeKernel<MODE_K2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ceiling);
    cudaFuncSetAttribute(pcgPipeKernel<MODE_K1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ceiling);
    cudaFuncSetAttribute(pcgPipeKernel<MODE_K2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ceiling);
    cudaFuncSetAttribute(pcgSolveKernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, ceiling);
    cudaFuncSetAttribute(pcgSolveKernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, ceiling);
    cudaFuncSetAttribute(pcgResidentKernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(ResSmem)));
    cudaFuncSetAttribute(pcgResidentKernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(ResSmem)));
    cudaFuncSetAttribute(pcgResidentKernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(sizeof(ResSmem)));
    cudaGetLastError();
    done |= 1ull << device;
}

int pcgTileBlocks(const Ctx *ctx) { return divUp(ctx->I, TR) * divUp(ctx->J, TC); }

static bool pipeUsable(const Ctx *ctx) { return (ctx->J % 2) == 0 && !ctx->forceTileKernels; }

int pcgSolveDevice(Ctx *ctx, int iterLimit, double tol)
{
    // FS2D_MG_FORCE (measurement aid): a single rank in slab mode runs the <MG = true> kernels, without a neighbour
    static const bool mgForce = std::getenv("FS2D_MG_FORCE") != nullptr;
    const bool mgOn = ctx->slab.enabled && (ctx->slab.world > 1 || mgForce);
    const int tilesJ = divUp(ctx->J, TC);
    int blocks = pcgTileBlocks(ctx);  // tiles this rank walks
    int tileBase = 0;
    if (mgOn)
    {
        const int t0 = ctx->slab.rowBegin / TR, t1 = divUp(ctx->slab.rowEnd, TR);
        tileBase = t0 * tilesJ;
        blocks = (t1 - t0) * tilesJ;
    }
    const bool pipe = pipeUsable(ctx);
    if (mgOn && (!pipe || ctx->p.convergence_threads > 0))
    {
        ctx->lastError = "pcg: slab mode needs the pipelined kernels (even gridSizeJ) and convergence_threads = 0";
        return FS2D_ERR_STATE;
    }
    const int share = mgOn ? ctx->slab.share : 1;
    // CTAs of the persistent grids: what the device can hold co-resident (a cooperative launch refuses more; queried
    // once, for the kernel with the larger footprint), split between the ranks that share the GPU, capped by the
    // number of tiles and by the test knob.
    pcgKernelAttributes(ctx->device);
    if (pipe && ctx->pcgOccupancy == 0)
    {
        int perSm = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, pcgSolveKernel<true>, NT, sizeof(SolveSmem)) != cudaSuccess || perSm < 1)
        {
            cudaGetLastError();
            perSm = 1;
        }
        ctx->pcgOccupancy = std::min(perSm, 2);
    }
    // Several ranks on ONE GPU (tests): the kernels of different ranks spin on each other, so all of them have to be
    // co-resident whatever mix of streaming (109 KB, 256 threads) and resident (218 KB, 512 x 128 registers: a whole SM)
    // CTAs the ranks launch. Two streaming CTAs of one rank on an SM could keep a resident CTA of another rank out for
    // good, so with a shared GPU every CTA of these kernels takes an SM of its own (one per SM and rank share, a shared
    // memory request no two of them fit side by side).
    const int ctasPerSm = share > 1 ? 1 : std::max(1, ctx->pcgOccupancy);
    const size_t exclusiveSmem = PCG_EXCLUSIVE_SMEM;
    int pipeBlocks = std::max(1, std::min(blocks, ctasPerSm * ctx->smCount / share));
    if (ctx->pcgGridLimit > 0) pipeBlocks = std::min(pipeBlocks, ctx->pcgGridLimit);
    const int flat = std::max(1, ctx->smCount * 8 / share);
    if (pcgTileBlocks(ctx) > ctx->maxBlocks || flat > ctx->maxBlocks)
    {
        ctx->lastError = "pcg: partials buffer too small";
        return FS2D_ERR_STATE;
    }
    PcgArgs a = baseArgs(ctx);
    a.tol = tol;
    cudaStream_t st = ctx->stream;
    const int T = ctx->p.convergence_threads;
    if (T > 0 && ctx->N < T)
    {
        ctx->lastError = "pcg: convergence_threads larger than the cell count";
        return FS2D_ERR_ARG;
    }

    // The whole-solve kernel (one cooperative launch for all iterations) is the default; the per-iteration kernels
    // remain for odd gridSizeJ, the reference-compatible convergence test, and as the A/B baseline (FS2D_PCG_STEPWISE).
    const bool whole = pipe && T == 0 && !ctx->stepwisePcg;
    MgArgs mg;
    memset(&mg, 0, sizeof(mg));
    long long nLo = 0, nHi = ctx->N, oLo = 0, oHi = ctx->N;
    if (whole && !mgOn)
    {
        // one rank: the barrier of the whole-solve kernel still goes through the (local) mail ring
        SlabState &sl = ctx->slab;
        sl.solveSeq++;
        mg.rank = 0;
        mg.world = 1;
        mg.rowBegin = 0;
        mg.rowEnd = ctx->I;
        mg.ringBase = static_cast<int>(sl.solveSeq & 1ull) * 8;
        mg.solveTag = sl.solveSeq << 20;
        mg.mail = ctx->mail;
        mg.peerMail[0] = ctx->mail;
        mg.iterLimit = iterLimit;
        static const int dbg1 = std::getenv("FS2D_MG_DEBUG") ? std::atoi(std::getenv("FS2D_MG_DEBUG")) : 0;
        if (dbg1 & 8)
        {
            static unsigned long long *tl1 = nullptr;
            if (!tl1) cudaMalloc(reinterpret_cast<void **>(&tl1), 1024 * 8 * sizeof(unsigned long long));
            mg.timeline = tl1;
            ctx->mgTimeline = tl1;
        }
    }
    auto peerOf = [&](int r, double *p) -> double * {
        if (r < 0 || r >= ctx->slab.world) return nullptr;
        return reinterpret_cast<double *>(ctx->slab.peerHeap[r] + (reinterpret_cast<unsigned char *>(p) - ctx->heap));
    };
    if (mgOn)
    {
        SlabState &sl = ctx->slab;
        if (sl.connected != sl.world - 1)
        {
            ctx->lastError = "pcg: slab peers are not connected";
            return FS2D_ERR_COMM;
        }
        sl.solveSeq++;
        mg.rank = sl.rank;
        mg.world = sl.world;
        mg.rowBegin = sl.rowBegin;
        mg.rowEnd = sl.rowEnd;
        mg.tileBase = tileBase;
        mg.ringBase = static_cast<int>(sl.solveSeq & 1ull) * 8;
        mg.solveTag = sl.solveSeq << 20;
        mg.mail = ctx->mail;
        for (int r = 0; r < sl.world; r++)
            mg.peerMail[r] = reinterpret_cast<SlabMail *>(sl.peerHeap[r] + (reinterpret_cast<unsigned char *>(ctx->mail) - ctx->heap));
        mg.iterLimit = iterLimit;
        mg.haloLL = ctx->haloLL;
        mg.loHaloLL = sl.rank > 0 ? reinterpret_cast<unsigned long long *>(sl.peerHeap[sl.rank - 1] + (reinterpret_cast<unsigned char *>(ctx->haloLL) - ctx->heap)) : nullptr;
        mg.hiHaloLL = sl.rank + 1 < sl.world ? reinterpret_cast<unsigned long long *>(sl.peerHeap[sl.rank + 1] + (reinterpret_cast<unsigned char *>(ctx->haloLL) - ctx->heap)) : nullptr;
        static const int mgDebug = std::getenv("FS2D_MG_DEBUG") ? std::atoi(std::getenv("FS2D_MG_DEBUG")) : 0;
        mg.debug = mgDebug;
        if (mgDebug & 8)
        {
            // timeline dump: 1024 phases x 8 stamps, allocated once per process, printed by tools/mg_probe.py
            static unsigned long long *tl = nullptr;
            if (!tl) cudaMalloc(reinterpret_cast<void **>(&tl), 1024 * 8 * sizeof(unsigned long long));
            mg.timeline = tl;
            ctx->mgTimeline = tl;
        }
        const SlabRows ext = slabExt(ctx, 1);
        nLo = static_cast<long long>(ext.lo) * ctx->J;
        nHi = static_cast<long long>(ext.hi) * ctx->J;
        oLo = static_cast<long long>(sl.rowBegin) * ctx->J;
        oHi = static_cast<long long>(sl.rowEnd) * ctx->J;
    }

    const bool prof = ctx->profilePcg;
    if (prof)
        while (static_cast<int>(ctx->profEvents.size()) < 2 * iterLimit + 1)
        {
            cudaEvent_t e;
            FS2D_CUDA(cudaEventCreate(&e));
            ctx->profEvents.push_back(e);
        }
    const bool active = pipe && !ctx->densePcg;
    FS2D_CUDA(cudaMemsetAsync(&ctx->scalars->pad, 0, sizeof(int), st));  // set by pcgResidentKernel when it takes the solve
    if (active)
    {
        pcgTileFlagKernel<<<blocks, NT, 0, st>>>(ctx->rowInfo, ctx->rhs, ctx->I, ctx->J, a.tilesJ, tileBase, ctx->tileFlags);
        pcgTileCompactKernel<<<1, 1024, 0, st>>>(ctx->tileFlags, blocks, tileBase, ctx->activeTiles, ctx->activeCount);
        ctx->launches += 2;
        a.activeTiles = ctx->activeTiles;
        a.activeCount = ctx->activeCount;
    }
    // the longest tile list pcgResidentKernel takes (the same tests as in the kernel): such a solve needs no r / s arrays
    int resBlocks = std::max(1, ctx->smCount / share);
    if (ctx->pcgGridLimit > 0) resBlocks = std::min(resBlocks, ctx->pcgGridLimit);
    int residentLimit = 0;
    if (whole && active && ctx->residentPcg)
    {
        residentLimit = RES_TPC * resBlocks;
        if (!mgOn && ctx->pagedPcg && ctx->N < (1ll << 31))
            residentLimit = static_cast<int>(std::max<long long>(residentLimit, std::min<long long>(static_cast<long long>(RES_PAGED_TPC) * resBlocks, ctx->N / PTILE_PAD)));
    }
    bool initLight = residentLimit > 0;
    if (mgOn || whole)
    {
        mg.phase = 0;
        pcgInitKernel<true><<<flat, NT, 0, st>>>(ctx->rhs, ctx->x, ctx->r[0], ctx->z, ctx->s[0], active ? ctx->r[1] : nullptr, ctx->s[1], ctx->q,
                                                 nHi, ctx->partials, ctx->scalars, nLo, oLo, oHi, mg, ctx->activeCount, residentLimit);
    }
    else
    {
        pcgInitKernel<false><<<flat, NT, 0, st>>>(ctx->rhs, ctx->x, ctx->r[0], ctx->z, ctx->s[0], active ? ctx->r[1] : nullptr, ctx->s[1],
                                                  ctx->q, ctx->N, ctx->partials, ctx->scalars, 0, 0, ctx->N, mg, nullptr, 0);
    }
    ctx->launches++;
    if (whole)
    {
        SolveArgs g;
        g.a = a;
        g.z = ctx->z;
        g.q = ctx->q;
        g.x = ctx->x;
        g.s[0] = ctx->s[0];
        g.s[1] = ctx->s[1];
        g.r[0] = ctx->r[0];
        g.r[1] = ctx->r[1];
        g.loQ = mgOn ? peerOf(mg.rank - 1, ctx->q) : nullptr;
        g.hiQ = mgOn ? peerOf(mg.rank + 1, ctx->q) : nullptr;
        g.loZ = mgOn ? peerOf(mg.rank - 1, ctx->z) : nullptr;
        g.hiZ = mgOn ? peerOf(mg.rank + 1, ctx->z) : nullptr;
        g.numTiles = blocks;
        g.tileFlags = active ? ctx->tileFlags : nullptr;
        g.iterLimit = iterLimit;
        g.ticket = &ctx->scalars->ticketS;
        FS2D_CUDA(cudaMemsetAsync(g.ticket, 0, sizeof(unsigned int), st));
        if (prof)
        {
            FS2D_CUDA(cudaMemsetAsync(ctx->scalars->phaseNs, 0, 2 * sizeof(unsigned long long) + sizeof(unsigned int), st));
            cudaEventRecord(ctx->profEvents[0], st);
        }
        if (!ctx->solveMapsTried)
        {
            ctx->solveMapsTried = true;
            ctx->solveMaps = std::malloc(sizeof(SolveMaps));
            ctx->solveMapsOk = ctx->solveMaps && !ctx->rowCopyPcg && buildSolveMaps(ctx, static_cast<SolveMaps *>(ctx->solveMaps));
        }
        SolveMaps maps;
        memset(&maps, 0, sizeof(maps));
        int useTensor = ctx->solveMapsOk ? 1 : 0;
        if (useTensor) maps = *static_cast<SolveMaps *>(ctx->solveMaps);
        // L2 policies only make sense when the walked set is near the L2 size: the active-tile walk (see WalkHints)
        static const int hintEnv = std::getenv("FS2D_PCG_HINTS") ? std::atoi(std::getenv("FS2D_PCG_HINTS")) : -1;
        const int hintMask = hintEnv >= 0 ? hintEnv : 7;  // measured at 4096^2: 19.7 -> 18.2 us per phase with x, s, r evict-first and q, z evict-last
        if (useTensor && active) useTensor |= hintMask << 8;
        if (active && ctx->residentPcg)
        {
            // small active sets stay in shared memory for the whole solve (pcgResidentKernel decides on the device,
            // from the tile count, whether it can hold them; if not it returns at once and the streaming kernel runs)
            void *rargs[] = {&g, &mg};
            const size_t rsmem = sizeof(ResSmem);
            cudaError_t re;
            if (mgOn)
            {
                re = cudaLaunchCooperativeKernel(reinterpret_cast<void *>(pcgResidentKernel<true, false>), dim3(resBlocks), dim3(RNT), rargs, rsmem, st);
            }
            else
            {
                re = cudaLaunchCooperativeKernel(reinterpret_cast<void *>(pcgResidentKernel<false, false>), dim3(resBlocks), dim3(RNT), rargs, rsmem, st);
            }
            bool launched = true;
            if (re == cudaErrorCooperativeLaunchTooLarge)
            {
                cudaGetLastError();  // the device is shared: the streaming kernel (or its stepwise fallback) takes the solve
                launched = false;
            }
            else
            {
                FS2D_CUDA(re);
                ctx->launches++;
                if (!mgOn && ctx->pagedPcg)
                {
                    // more tiles than the SMs hold, but not many more (4096^2 dam break): resident + paged tiles; returns
                    // at once when the resident kernel took the solve or when the list is too long
                    re = cudaLaunchCooperativeKernel(reinterpret_cast<void *>(pcgResidentKernel<false, true>), dim3(resBlocks), dim3(RNT), rargs, rsmem, st);
                    if (re == cudaErrorCooperativeLaunchTooLarge)
                    {
                        cudaGetLastError();
                        launched = false;
                    }
                    else
                    {
                        FS2D_CUDA(re);
                        ctx->launches++;
                    }
                }
            }
            if (!launched && initLight)
            {
                // a kernel the light initialisation counted on did not start: the streaming kernel needs the r / s arrays
                pcgInitRestKernel<<<flat, NT, 0, st>>>(ctx->rhs, ctx->r[0], ctx->s[0], ctx->r[1], ctx->s[1], nLo, nHi);
                ctx->launches++;
                initLight = false;
            }
        }
        void *args[] = {&g, &mg, &maps, &useTensor};
        const size_t smem = (mgOn && share > 1) ? std::max(sizeof(SolveSmem), exclusiveSmem) : sizeof(SolveSmem);
        if (mgOn)
        {
            FS2D_CUDA(cudaLaunchCooperativeKernel(reinterpret_cast<void *>(pcgSolveKernel<true>), dim3(pipeBlocks), dim3(NT), args, smem, st));
        }
        else
        {
            const cudaError_t le = cudaLaunchCooperativeKernel(reinterpret_cast<void *>(pcgSolveKernel<false>), dim3(pipeBlocks), dim3(NT), args, smem, st);
            if (le == cudaErrorCooperativeLaunchTooLarge)
            {
                // the device cannot hold the grid co-resident right now (other resident work, MPS): the stepwise kernels
                // need no co-residency and produce the same iterates
                cudaGetLastError();
                ctx->stepwisePcg = true;
                return pcgSolveDevice(ctx, iterLimit, tol);
            }
            FS2D_CUDA(le);
        }
        ctx->launches++;
        if (prof) cudaEventRecord(ctx->profEvents[1], st);
        pcgFinalizeKernel<<<flat, NT, 0, st>>>(ctx->x, ctx->s[0], ctx->s[1], oLo, oHi, ctx->scalars, iterLimit);
        ctx->launches++;
        FS2D_CUDA(cudaGetLastError());
        if (prof)
        {
            PcgScalars sc;
            FS2D_CUDA(fs2dCopyToHost(ctx, &sc, ctx->scalars, sizeof(sc)));
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ctx->profEvents[0], ctx->profEvents[1]);
            // K1 / K2 split of the launch by the device-side phase clocks (CTA 0, globaltimer); the launch total is
            // what the CUDA events saw
            const double pa = static_cast<double>(sc.phaseNs[0]), pb = static_cast<double>(sc.phaseNs[1]);
            const double tot = pa + pb > 0.0 ? pa + pb : 1.0;
            ctx->profMs[0] += ms * pa / tot;
            ctx->profMs[1] += ms * pb / tot;
            ctx->profLaunches[0] += sc.iter;
            ctx->profLaunches[1] += sc.iter;
            ctx->profSolveMs += ms;
            ctx->profSolves++;
        }
        return FS2D_OK;
    }
    for (int i = 0; i < iterLimit; i++)
    {
        if (prof) cudaEventRecord(ctx->profEvents[2 * i], st);
        PcgArgs k1 = a;
        k1.in0 = ctx->z;
        k1.in1 = ctx->s[i & 1];
        k1.out0 = ctx->s[(i + 1) & 1];
        k1.out1 = ctx->q;
        k1.x = ctx->x;
        if (mgOn)
        {
            mg.iter = i;
            mg.phase = 2 * i + 1;
            mg.loOut1 = peerOf(mg.rank - 1, ctx->q);
            mg.hiOut1 = peerOf(mg.rank + 1, ctx->q);
            pcgPipeKernel<MODE_K1, true><<<pipeBlocks, NT, share > 1 ? exclusiveSmem : sizeof(PipeSmem<MODE_K1>), st>>>(k1, blocks, mg);
        }
        else if (pipe)
            pcgPipeKernel<MODE_K1, false><<<pipeBlocks, NT, sizeof(PipeSmem<MODE_K1>), st>>>(k1, blocks, mg);
        else
            pcgTileKernel<MODE_K1><<<blocks, NT, 0, st>>>(k1);
        if (prof) cudaEventRecord(ctx->profEvents[2 * i + 1], st);
        PcgArgs k2 = a;
        k2.in0 = ctx->r[i & 1];
        k2.in1 = ctx->q;
        k2.out0 = ctx->r[(i + 1) & 1];
        k2.out1 = ctx->z;
        if (mgOn)
        {
            mg.phase = 2 * i + 2;
            mg.loOut1 = peerOf(mg.rank - 1, ctx->z);
            mg.hiOut1 = peerOf(mg.rank + 1, ctx->z);
            pcgPipeKernel<MODE_K2, true><<<pipeBlocks, NT, share > 1 ? exclusiveSmem : sizeof(PipeSmem<MODE_K2>), st>>>(k2, blocks, mg);
        }
        else if (pipe)
            pcgPipeKernel<MODE_K2, false><<<pipeBlocks, NT, sizeof(PipeSmem<MODE_K2>), st>>>(k2, blocks, mg);
        else
            pcgTileKernel<MODE_K2><<<blocks, NT, 0, st>>>(k2);
        ctx->launches += 2;
        if (T > 0)
        {
            pcgRangeErrKernel<<<T, NT, 0, st>>>(ctx->r[(i + 1) & 1], ctx->N, T, ctx->partials + 2 * ctx->maxBlocks,
                                               ctx->scalars, tol, ctx->trace, ctx->traceCapacity);
            ctx->launches++;
        }
    }
    if (prof && iterLimit > 0) cudaEventRecord(ctx->profEvents[2 * iterLimit], st);
    if (mgOn)
    {
        pcgMgCloseKernel<<<1, 1, 0, st>>>(a, mg);
        pcgFinalizeKernel<<<flat, NT, 0, st>>>(ctx->x, ctx->s[0], ctx->s[1], oLo, oHi, ctx->scalars, iterLimit);
    }
    else
    {
        pcgFinalizeKernel<<<flat, NT, 0, st>>>(ctx->x, ctx->s[0], ctx->s[1], 0, ctx->N, ctx->scalars, iterLimit);
        pcgResultKernel<<<1, 1, 0, st>>>(ctx->scalars, iterLimit);
    }
    ctx->launches += 2;
    FS2D_CUDA(cudaGetLastError());
    if (prof && iterLimit > 0)
    {
        // only iterations that did work count (after convergence the kernels return at once)
        PcgScalars sc;
        FS2D_CUDA(fs2dCopyToHost(ctx, &sc, ctx->scalars, sizeof(sc)));
        FS2D_CUDA(cudaStreamSynchronize(st));
        const int executed = sc.iter < iterLimit ? sc.iter : iterLimit;
        for (int i = 0; i < executed; i++)
        {
            float a1 = 0.f, a2 = 0.f;
            cudaEventElapsedTime(&a1, ctx->profEvents[2 * i], ctx->profEvents[2 * i + 1]);
            cudaEventElapsedTime(&a2, ctx->profEvents[2 * i + 1], ctx->profEvents[2 * i + 2]);
            ctx->profMs[0] += a1;
            ctx->profMs[1] += a2;
        }
        ctx->profLaunches[0] += executed;
        ctx->profLaunches[1] += executed;
    }
    return FS2D_OK;
}

int pcgSpmvHost(Ctx *ctx, const double *in, double *out, bool precond)
{
    const size_t bytes = static_cast<size_t>(ctx->N) * sizeof(double);
    FS2D_CUDA(cudaMemcpyAsync(ctx->z, in, bytes, cudaMemcpyHostToDevice, ctx->stream));
    PcgArgs a = baseArgs(ctx);
    a.in0 = ctx->z;
    a.out1 = ctx->q;
    const int blocks = pcgTileBlocks(ctx);
    if (precond)
        pcgTileKernel<MODE_APPLY_M><<<blocks, NT, 0, ctx->stream>>>(a);
    else
        pcgTileKernel<MODE_APPLY_A><<<blocks, NT, 0, ctx->stream>>>(a);
    ctx->launches++;
    FS2D_CUDA(cudaGetLastError());
    FS2D_CUDA(fs2dCopyToHost(ctx, out, ctx->q, bytes));
    FS2D_CUDA(cudaStreamSynchronize(ctx->stream));
    return FS2D_OK;
}
