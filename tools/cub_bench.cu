// Comparator only (dev tool, NOT part of the product or the bench contract): CUB DeviceRadixSort on the
// same workload, to know what the toolkit's own onesweep reaches on this B200.
#include <cub/cub.cuh>
#include <cstdio>
#include <cstdint>
__global__ void fill(uint32_t* p, size_t n, int kind) {
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint64_t z = 0x5EEDull + (i + 1) * 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; z ^= z >> 31;
        p[i] = kind == 0 ? (uint32_t)z : kind == 1 ? (uint32_t)i : 0u;
    }
}
int main() {
    const size_t n = 1ull << 28;
    uint32_t *a, *b, *va, *vb; cudaMalloc(&a, n * 4); cudaMalloc(&b, n * 4); cudaMalloc(&va, n * 4); cudaMalloc(&vb, n * 4);
    void* tmp = nullptr; size_t tb = 0, tb2 = 0;
    cub::DeviceRadixSort::SortKeys(nullptr, tb, a, b, n);
    cub::DeviceRadixSort::SortPairs(nullptr, tb2, a, b, va, vb, n);
    if (tb2 > tb) tb = tb2;
    cudaMalloc(&tmp, tb);
    const char* names[] = {"uniform", "range", "zeros"};
    for (int pairs = 0; pairs < 2; ++pairs)
        for (int kind = 0; kind < 3; ++kind) {
            fill<<<1184, 256>>>(a, n, kind);
            cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
            for (int it = 0; it < 2; ++it) { if (pairs) cub::DeviceRadixSort::SortPairs(tmp, tb, a, b, va, vb, n); else cub::DeviceRadixSort::SortKeys(tmp, tb, a, b, n); }
            cudaEventRecord(e0);
            const int iters = 5;
            for (int it = 0; it < iters; ++it) { if (pairs) cub::DeviceRadixSort::SortPairs(tmp, tb, a, b, va, vb, n); else cub::DeviceRadixSort::SortKeys(tmp, tb, a, b, n); }
            cudaEventRecord(e1); cudaDeviceSynchronize();
            float ms; cudaEventElapsedTime(&ms, e0, e1); ms /= iters;
            printf("CUB %s u32 %-8s n=2^28: %.3f ms  %.2f Gkeys/s  (%s)\n", pairs ? "pairs" : "keys ", names[kind], ms, n / ms / 1e6, cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
