// Micro-benchmarks that size the persistent round kernel: grid.sync() latency vs grid shape, and
// global / shared atomicMin(u64) throughput vs contention.  Build: nvcc -arch=sm_100a -O3 -o bench_sync bench_sync.cu
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;
typedef unsigned long long u64;

__global__ void k_sync(int n, unsigned *sink) {
    cg::grid_group g = cg::this_grid();
    unsigned acc = 0;
    for (int i = 0; i < n; ++i) { g.sync(); acc += i; }
    if (threadIdx.x == 0 && blockIdx.x == 0) *sink = acc;
}

// every thread does `per` atomicMin on one of `naddr` addresses
__global__ void k_atom(u64 *tab, unsigned naddr, int per, int stride_mode) {
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = 0; i < per; ++i) {
        unsigned a = stride_mode ? (t * 2654435761u + i * 40503u) % naddr : ((t + i * 977u) / 4) % naddr;
        atomicMin(tab + a, ((u64)(t ^ (i * 7919u)) << 20) | i);
    }
}
__global__ void k_atom32(unsigned *tab, unsigned naddr, int per) {
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = 0; i < per; ++i) {
        unsigned a = (t * 2654435761u + i * 40503u) % naddr;
        atomicAdd(tab + a, 1u);
    }
}
__global__ void k_satom(u64 *out, unsigned naddr, int per) {
    __shared__ u64 tab[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) tab[i] = ~0ull;
    __syncthreads();
    unsigned t = blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = 0; i < per; ++i) {
        unsigned a = (t * 2654435761u + i * 40503u) % naddr;
        atomicMin(tab + a, ((u64)(t ^ (i * 7919u)) << 20) | i);
    }
    __syncthreads();
    if (threadIdx.x < naddr) out[blockIdx.x * 1024 + threadIdx.x] = tab[threadIdx.x];
}
// dependent-load chain latency (L2-resident table)
__global__ void k_chase(const unsigned *tab, int n, unsigned *sink) {
    unsigned p = threadIdx.x;
    for (int i = 0; i < n; ++i) p = __ldcg(tab + p);
    if (p == 0xFFFFFFFF) *sink = p;
}

int main() {
    int nsm = 0;
    cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
    unsigned *sink; cudaMalloc(&sink, 4);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    int n = 200;
    for (int bs : {256, 512, 1024})
        for (int per : {1, 2, 3, 4}) {
            if (bs * per > 2048) continue;
            int grid = nsm * per;
            void *args[] = {&n, &sink};
            cudaLaunchCooperativeKernel((void *)k_sync, dim3(grid), dim3(bs), args, 0, 0);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            cudaError_t e = cudaLaunchCooperativeKernel((void *)k_sync, dim3(grid), dim3(bs), args, 0, 0);
            cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("grid.sync  block=%4d  blocks/SM=%d  grid=%4d : %.2f us per sync (%s)\n", bs, per, grid, ms * 1e3 / n, cudaGetErrorString(e));
        }
    u64 *tab; cudaMalloc(&tab, (size_t)(1 << 22) * 8);
    for (int mode : {1, 0})
        for (unsigned naddr : {1u, 16u, 128u, 1024u, 65536u, 1u << 20, 1u << 22}) {
            cudaMemset(tab, 0xFF, (size_t)(1 << 22) * 8);
            int per = 8, grid = nsm * 8, bs = 256;
            k_atom<<<grid, bs>>>(tab, naddr, per, mode);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            k_atom<<<grid, bs>>>(tab, naddr, per, mode);
            cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double ops = (double)grid * bs * per;
            printf("global atomicMin u64 %s naddr=%8u : %8.1f us for %.2fM ops = %7.2f Gops/s\n", mode ? "scattered" : "runs-of-4", naddr, ms * 1e3, ops / 1e6, ops / ms / 1e6);
        }
    for (unsigned naddr : {1u, 1024u, 1u << 20}) {
        int per = 8, grid = nsm * 8, bs = 256;
        k_atom32<<<grid, bs>>>((unsigned *)tab, naddr, per);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k_atom32<<<grid, bs>>>((unsigned *)tab, naddr, per);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double ops = (double)grid * bs * per;
        printf("global atomicAdd u32 scattered naddr=%8u : %8.1f us = %7.2f Gops/s\n", naddr, ms * 1e3, ops / ms / 1e6);
    }
    for (unsigned naddr : {1u, 16u, 90u, 1024u}) {
        int per = 16, grid = nsm, bs = 256;
        k_satom<<<grid, bs>>>(tab, naddr, per);
        cudaDeviceSynchronize();
        cudaEventRecord(e0);
        k_satom<<<grid, bs>>>(tab, naddr, per);
        cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("shared atomicMin u64 naddr=%5u : %8.1f us for %d ops per block (%.1f ns/op/block)\n", naddr, ms * 1e3, bs * per, ms * 1e6 / (bs * per));
    }
    unsigned *ct; cudaMalloc(&ct, 1 << 20);
    unsigned *h = new unsigned[1 << 18];
    for (int i = 0; i < (1 << 18); ++i) h[i] = (unsigned)((i * 40503u + 12345u) & ((1 << 18) - 1));
    cudaMemcpy(ct, h, 1 << 20, cudaMemcpyHostToDevice);
    int nn = 1000;
    k_chase<<<1, 32>>>(ct, nn, sink); cudaDeviceSynchronize();
    cudaEventRecord(e0); k_chase<<<1, 32>>>(ct, nn, sink); cudaEventRecord(e1); cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("dependent L2 load chain: %.1f ns per hop\n", ms * 1e6 / nn);
    cudaEventRecord(e0); k_chase<<<1, 32>>>(ct, 1, sink); cudaEventRecord(e1); cudaDeviceSynchronize();
    cudaEventElapsedTime(&ms, e0, e1);
    printf("near-empty kernel (event to event): %.2f us\n", ms * 1e3);
    return 0;
}
