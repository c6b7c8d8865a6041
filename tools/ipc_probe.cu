// Probe: can two processes (one per GPU) exchange CUDA IPC handles and signal each other through
// peer stores + local spinning? Measures the one-way flag latency that sizes the fused
// reduction/halo exchange of the slab PCG (multi-GPU). Not part of the product library.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>

extern "C" {

static void *g_local = nullptr;
static void *g_peer = nullptr;

int probe_init(int device, size_t bytes, void *handleOut /* 64 bytes */)
{
    if (cudaSetDevice(device) != cudaSuccess) return -1;
    if (cudaMalloc(&g_local, bytes) != cudaSuccess) return -2;
    cudaMemset(g_local, 0, bytes);
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, g_local) != cudaSuccess) return -3;
    memcpy(handleOut, &h, sizeof(h));
    return 0;
}

int probe_open(const void *handleIn)
{
    cudaIpcMemHandle_t h;
    memcpy(&h, handleIn, sizeof(h));
    cudaError_t e = cudaIpcOpenMemHandle(&g_peer, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess)
    {
        fprintf(stderr, "cudaIpcOpenMemHandle: %s\n", cudaGetErrorString(e));
        return -1;
    }
    return 0;
}

// ping-pong: rank 0 writes seq to the peer flag, waits for the peer's echo in its own flag.
__global__ void pingpong(volatile unsigned long long *localFlag, unsigned long long *peerFlag, int rank, int rounds,
                         long long *cyclesOut, int *timeouts)
{
    const long long t0 = clock64();
    for (int k = 1; k <= rounds; k++)
    {
        if (rank == 0)
        {
            *peerFlag = k;
            __threadfence_system();
        }
        long long spins = 0;
        while (*localFlag < static_cast<unsigned long long>(k))
        {
            if (++spins > 20000000)
            {
                *timeouts = k;
                return;
            }
        }
        if (rank == 1)
        {
            *peerFlag = k;
            __threadfence_system();
        }
    }
    *cyclesOut = clock64() - t0;
}

// rows push: every CTA writes 'n' doubles to the peer then the last CTA raises the flag.
__global__ void pushRows(const double *src, double *peerDst, long long n, unsigned int *ticket, unsigned long long *peerFlag,
                         unsigned long long seq)
{
    for (long long i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += static_cast<long long>(gridDim.x) * blockDim.x)
        peerDst[i] = src[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0)
    {
        if (atomicAdd(ticket, 1u) == gridDim.x - 1)
        {
            *ticket = 0;
            __threadfence_system();
            *peerFlag = seq;
        }
    }
}

__global__ void waitFlag(volatile unsigned long long *localFlag, unsigned long long seq, int *timeouts)
{
    long long spins = 0;
    while (*localFlag < seq)
        if (++spins > 20000000)
        {
            *timeouts = 1;
            return;
        }
}

// returns one-way latency in us (ping-pong round trip / 2) or negative on failure
double probe_pingpong(int rank, int rounds)
{
    long long *cyc;
    int *to;
    cudaMalloc(&cyc, 8);
    cudaMalloc(&to, 4);
    cudaMemset(cyc, 0, 8);
    cudaMemset(to, 0, 4);
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    pingpong<<<1, 1>>>(reinterpret_cast<volatile unsigned long long *>(g_local), reinterpret_cast<unsigned long long *>(g_peer), rank,
                       rounds, cyc, to);
    cudaEventRecord(b);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1.0;
    int hto = 0;
    cudaMemcpy(&hto, to, 4, cudaMemcpyDeviceToHost);
    if (hto) return -2.0 - hto;
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return 1000.0 * ms / rounds / 2.0;
}

// returns us per (push n doubles + flag + peer wait) exchange, both directions at once
double probe_exchange(long long nDoubles, int rounds, unsigned long long seqBase)
{
    double *src;
    unsigned int *ticket;
    int *to;
    cudaMalloc(&src, nDoubles * 8);
    cudaMalloc(&ticket, 4);
    cudaMalloc(&to, 4);
    cudaMemset(ticket, 0, 4);
    cudaMemset(to, 0, 4);
    unsigned long long *localFlag = reinterpret_cast<unsigned long long *>(g_local) + 16;
    unsigned long long *peerFlag = reinterpret_cast<unsigned long long *>(g_peer) + 16;
    double *peerDst = reinterpret_cast<double *>(g_peer) + 1024;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a);
    for (int k = 1; k <= rounds; k++)
    {
        pushRows<<<8, 256>>>(src, peerDst, nDoubles, ticket, peerFlag, seqBase + k);
        waitFlag<<<1, 1>>>(localFlag, seqBase + k, to);
    }
    cudaEventRecord(b);
    if (cudaDeviceSynchronize() != cudaSuccess) return -1.0;
    int hto = 0;
    cudaMemcpy(&hto, to, 4, cudaMemcpyDeviceToHost);
    if (hto) return -2.0;
    float ms = 0;
    cudaEventElapsedTime(&ms, a, b);
    return 1000.0 * ms / rounds;
}
}
