// micro-benchmark: latency of dependent shared-memory loads / atomics issued by ONE warp in a cluster kernel,
// through (a) plain shared addressing, (b) the pointer cluster.map_shared_rank(p, own rank) returns,
// (c) a neighbour CTA's shared memory; plus S2R SR_CgaCtaId and __syncthreads.
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdint>
namespace cg = cooperative_groups;

__global__ void __cluster_dims__(2, 1, 1) k(long long *out) {
    extern __shared__ __align__(16) unsigned char dyn[];
    uint32_t *a = reinterpret_cast<uint32_t *>(dyn);
    cg::cluster_group cl = cg::this_cluster();
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) a[i] = (i * 17 + 1) & 1023;
    cl.sync();
    uint32_t *own = cl.map_shared_rank(a, cl.block_rank());
    uint32_t *other = cl.map_shared_rank(a, cl.block_rank() ^ 1);
    if (threadIdx.x < 32 && cl.block_rank() == 0) {
        long long t0, t1; uint32_t x = threadIdx.x;
        t0 = clock64(); for (int i = 0; i < 256; i++) x = a[x]; t1 = clock64();
        if (threadIdx.x == 0) out[0] = (t1 - t0) / 256 + (x & 0);
        t0 = clock64(); for (int i = 0; i < 256; i++) x = own[x]; t1 = clock64();
        if (threadIdx.x == 0) out[1] = (t1 - t0) / 256 + (x & 0);
        t0 = clock64(); for (int i = 0; i < 256; i++) x = other[x]; t1 = clock64();
        if (threadIdx.x == 0) out[2] = (t1 - t0) / 256 + (x & 0);
        t0 = clock64(); for (int i = 0; i < 256; i++) x = atomicMin(&a[x & 1023], 0xFFFFFFFFu) & 1023; t1 = clock64();
        if (threadIdx.x == 0) out[3] = (t1 - t0) / 256 + (x & 0);
        t0 = clock64(); for (int i = 0; i < 256; i++) x = atomicMin(&own[x & 1023], 0xFFFFFFFFu) & 1023; t1 = clock64();
        if (threadIdx.x == 0) out[4] = (t1 - t0) / 256 + (x & 0);
        t0 = clock64(); for (int i = 0; i < 256; i++) x = atomicCAS(&a[x & 1023], 0xFFFFFFFEu, 0u) & 1023; t1 = clock64();
        if (threadIdx.x == 0) out[5] = (t1 - t0) / 256 + (x & 0);
        uint32_t acc = 0;
        t0 = clock64(); for (int i = 0; i < 64; i++) { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); acc += r; } t1 = clock64();
        if (threadIdx.x == 0) out[6] = (t1 - t0) / 64 + (acc & 0);
        t0 = clock64(); for (int i = 0; i < 64; i++) acc += __match_any_sync(0xFFFFFFFFu, x + i + threadIdx.x); t1 = clock64();
        if (threadIdx.x == 0) out[7] = (t1 - t0) / 64 + (acc & 0);
        t0 = clock64(); for (int i = 0; i < 64; i++) acc += __shfl_sync(0xFFFFFFFFu, acc, i & 31); t1 = clock64();
        if (threadIdx.x == 0) out[8] = (t1 - t0) / 64 + (acc & 0);
    }
    __syncthreads();
    if (cl.block_rank() == 0) {
        long long t0 = clock64();
        for (int i = 0; i < 64; i++) __syncthreads();
        long long t1 = clock64();
        if (threadIdx.x == 0) out[9] = (t1 - t0) / 64;
        t0 = clock64();
        if (threadIdx.x < 128) for (int i = 0; i < 64; i++) asm volatile("bar.sync 1, 128;" ::: "memory");
        t1 = clock64();
        if (threadIdx.x == 0) out[10] = (t1 - t0) / 64;
    }
    long long t0 = clock64();
    for (int i = 0; i < 16; i++) cl.sync();
    long long t1 = clock64();
    if (threadIdx.x == 0 && cl.block_rank() == 0) out[11] = (t1 - t0) / 16;
}

int main() {
    long long *d, h[12];
    cudaMalloc(&d, sizeof h);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    k<<<2, 512, 100 * 1024>>>(d);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h, d, sizeof h, cudaMemcpyDeviceToHost);
    const char *nm[12] = {"lds plain", "lds mapped-own", "lds remote", "atoms.min plain", "atoms.min mapped-own", "atoms.cas plain", "cluster_ctarank read",
                          "match_any (32 distinct)", "shfl", "__syncthreads (512 thr)", "bar.sync 1,128", "cluster.sync (2 CTAs)"};
    printf("%s\n", cudaGetErrorString(e));
    for (int i = 0; i < 12; i++) printf("%-26s %lld cycles\n", nm[i], h[i]);
    return 0;
}
