// peerbw -- how fast can a kernel on GPU 0 read GPU 1's memory over NVLink (and both ways at once), by access shape.
//   mode 0: 4-byte loads, 16 independent loads per thread (what the merge kernels issue)
//   mode 1: 16-byte loads, 4 independent loads per thread
//   mode 2: cp.async.bulk (TMA) 16 KiB global -> shared chunks, 2 in flight per CTA, then shared -> local global
// One process, two devices, peer access enabled; "both" launches the same kernel on both GPUs concurrently.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/peerbw tools/peerbw.cu && tools/peerbw
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

__global__ void __launch_bounds__(256) copy32(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n) {
    const size_t tile = 4096;
    for (size_t base = blockIdx.x * tile; base < n; base += gridDim.x * tile) {
        uint32_t r[16];
#pragma unroll
        for (int k = 0; k < 16; ++k) r[k] = src[base + threadIdx.x + k * 256];
#pragma unroll
        for (int k = 0; k < 16; ++k) dst[base + threadIdx.x + k * 256] = r[k];
    }
}
__global__ void __launch_bounds__(256) copy128(const uint4* __restrict__ src, uint4* __restrict__ dst, size_t n4) {
    const size_t tile = 1024;
    for (size_t base = blockIdx.x * tile; base < n4; base += gridDim.x * tile) {
        uint4 r[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) r[k] = src[base + threadIdx.x + k * 256];
#pragma unroll
        for (int k = 0; k < 4; ++k) dst[base + threadIdx.x + k * 256] = r[k];
    }
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(256) copy_tma(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, size_t n) {
    extern __shared__ __align__(128) uint32_t sm[];   // 2 x 4096 words
    __shared__ __align__(8) unsigned long long bar[2];
    const size_t tile = 4096;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar[i])));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](size_t base, int buf) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar[buf])), "r"(16384) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm + buf * 4096)),
                     "l"(src + base), "r"(16384), "r"(smem_u32(&bar[buf])) : "memory");
    };
    size_t base = blockIdx.x * tile;
    const size_t step = gridDim.x * tile;
    if (threadIdx.x == 0 && base < n) issue(base, 0);
    int it = 0;
    for (; base < n; base += step, ++it) {
        const int buf = it & 1;
        if (threadIdx.x == 0 && base + step < n) issue(base + step, buf ^ 1);
        const uint32_t parity = (it >> 1) & 1;
        uint32_t done = 0;
        while (!done)
            asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(done) : "r"(smem_u32(&bar[buf])), "r"(parity) : "memory");
#pragma unroll
        for (int k = 0; k < 16; ++k) dst[base + threadIdx.x + k * 256] = sm[buf * 4096 + threadIdx.x + k * 256];
        __syncthreads();
    }
}

int main() {
    int nd = 0;
    CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("need 2 GPUs\n"); return 0; }
    const size_t n = 1ull << 30;  // 4 GiB of words per buffer
    uint32_t *remote[2], *local[2];
    cudaStream_t st[2];
    cudaEvent_t e0[2], e1[2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&remote[d], n * 4));   // read by the OTHER device
        CK(cudaMalloc(&local[d], n * 4));
        CK(cudaMemset(remote[d], d + 1, n * 4));
        CK(cudaStreamCreate(&st[d]));
        CK(cudaEventCreate(&e0[d])); CK(cudaEventCreate(&e1[d]));
        CK(cudaFuncSetAttribute(copy_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768));
        CK(cudaDeviceSynchronize());
    }
    const char* names[3] = {"ld.32 x16", "ld.128 x4", "TMA bulk 16K x2"};
    for (int both = 0; both < 2; ++both)
      for (int skew : {0, 1, 8})      // source misaligned by `skew` words (mode 0 only: the merge reads runs at arbitrary offsets)
        for (int mode = 0; mode < 3; ++mode)
            for (int ctas : {4}) {
                if (skew && mode) continue;
                float best[2] = {1e9f, 1e9f};
                for (int rep = 0; rep < 3; ++rep) {
                    for (int d = 0; d <= both; ++d) {
                        CK(cudaSetDevice(d));
                        CK(cudaEventRecord(e0[d], st[d]));
                        const uint32_t* src = remote[1 - d] + skew;
                        if (mode == 0) copy32<<<148 * ctas, 256, 0, st[d]>>>(src, local[d], n - 4096);
                        if (mode == 1) copy128<<<148 * ctas, 256, 0, st[d]>>>((const uint4*)src, (uint4*)local[d], n / 4);
                        if (mode == 2) copy_tma<<<148 * ctas, 256, 32768, st[d]>>>(src, local[d], n);
                        CK(cudaEventRecord(e1[d], st[d]));
                    }
                    for (int d = 0; d <= both; ++d) {
                        CK(cudaSetDevice(d));
                        CK(cudaEventSynchronize(e1[d]));
                        float ms; CK(cudaEventElapsedTime(&ms, e0[d], e1[d]));
                        if (rep && ms < best[d]) best[d] = ms;
                    }
                }
                printf("%s  skew %d  %-16s %d CTAs/SM: GPU0 %.0f GB/s", both ? "both directions" : "one direction ", skew, names[mode], ctas, n * 4 / best[0] * 1e-6);
                if (both) printf(", GPU1 %.0f GB/s", n * 4 / best[1] * 1e-6);
                printf("\n");
            }
    uint32_t h = 0;
    CK(cudaSetDevice(0));
    CK(cudaMemcpy(&h, local[0] + 12345, 4, cudaMemcpyDeviceToHost));
    printf("check: %08x (expect 02020202)\n", h);
    return 0;
}
