// bulkstore -- micro-benchmark: how fast can one SM push 256 short digit runs per tile to HBM
//   mode 0: warp-coalesced st.global of the sorted tile with a per-key base lookup (what rs_onesweep_kernel does today)
//   mode 1: one cp.async.bulk (shared -> global) per run for its 16-byte aligned middle + scalar head/tail stores
// The runs of a tile go to 256 different regions of the output, like one radix pass over uniform keys.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/bulkstore tools/bulkstore.cu && tools/bulkstore
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int THREADS = 512, RADIX = 256, SLOT = 48;  // a run owns SLOT words of smem (44 + alignment slack)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int MODE>
__global__ void __launch_bounds__(THREADS, 2) k(uint32_t* out, size_t n_words, int n_tiles, int run) {
    extern __shared__ __align__(128) uint32_t sm[];
    uint32_t* data = sm;                       // RADIX * SLOT words
    uint32_t* gofs = sm + RADIX * SLOT;        // per-digit global word offset of this tile's run (low 32 bits)
    const size_t region = n_words / RADIX;
    for (int i = threadIdx.x; i < RADIX * SLOT; i += THREADS) data[i] = i;
    __syncthreads();
    for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
        if (threadIdx.x < RADIX) gofs[threadIdx.x] = (uint32_t)(t * run + (threadIdx.x * 7 & 3));  // arbitrary word alignment
        __syncthreads();
        if (MODE == 0) {
            for (int i = threadIdx.x; i < RADIX * run; i += THREADS) {
                const int d = i / run, j = i - d * run;
                out[d * region + gofs[d] + j] = data[d * SLOT + j];
            }
        } else {
            if (threadIdx.x < RADIX) {
                const int d = threadIdx.x;
                uint32_t* dst = out + d * region + gofs[d];
                const int skew = (int)(((uintptr_t)dst >> 2) & 3);        // smem copy of the run starts at the same 16B phase
                const uint32_t* src = data + d * SLOT + skew;
                const int head = (4 - skew) & 3;
                const int mid = (run - head) & ~3;
                for (int j = 0; j < head; ++j) dst[j] = src[j];
                for (int j = head + mid; j < run; ++j) dst[j] = src[j];
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + head),
                             "r"(smem_u32(src + head)), "r"(mid * 4)
                             : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
        }
        __syncthreads();
    }
}

int main() {
    const size_t n_words = 1ull << 28;
    uint32_t* out;
    cudaMalloc(&out, n_words * 4 + 4096);
    cudaMemset(out, 0, n_words * 4);
    const int smem = (RADIX * SLOT + RADIX) * 4;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    for (int run : {44, 40, 24, 12}) {
        const int n_tiles = (int)(n_words / RADIX / run);
        for (int mode = 0; mode < 2; ++mode) {
            float best = 1e9f;
            for (int rep = 0; rep < 4; ++rep) {
                cudaEventRecord(e0);
                if (mode == 0) k<0><<<296, THREADS, smem>>>(out, n_words, n_tiles, run);
                else k<1><<<296, THREADS, smem>>>(out, n_words, n_tiles, run);
                cudaEventRecord(e1);
                cudaEventSynchronize(e1);
                float ms;
                cudaEventElapsedTime(&ms, e0, e1);
                if (rep && ms < best) best = ms;
            }
            const double bytes = (double)n_tiles * RADIX * run * 4;
            printf("run=%2d words  mode=%d (%s): %.3f ms  %.0f GB/s written  err=%s\n", run, mode,
                   mode ? "bulk S2G per run" : "warp st.global", best, bytes / best * 1e-6, cudaGetErrorString(cudaGetLastError()));
        }
    }
    // spot check of mode 1 placement
    uint32_t h[64];
    cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
    printf("out[0..7] = %u %u %u %u %u %u %u %u\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6], h[7]);
    return 0;
}
