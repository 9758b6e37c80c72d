// Experiment (dev tool): in what order does the hardware serialise same-address shared-memory atomics of one warp
// instruction?  For every row each lane does old = atomicAdd(&cnt[d], 1); lane-ascending order would mean
// old == count_before + popc(lanes below with the same d).  Counts violations over many random patterns.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__global__ void __launch_bounds__(512) k(unsigned long long* bad, unsigned long long* rows, int distinct, int stride, int iters, int partial) {
    __shared__ uint32_t cnt[16][512];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = lane; i < 512; i += 32) cnt[warp][i] = 0;
    __syncwarp();
    uint32_t v = hash(tid * 7919u + blockIdx.x * 104729u + distinct);
    unsigned long long mybad = 0;
    const uint32_t lt = (1u << lane) - 1u;
    for (int it = 0; it < iters; ++it) {
        v = v * 1664525u + 1013904223u;
        const uint32_t d = ((v >> 9) % distinct) * stride & 511;   // stride 32 => every address in one bank
        const bool active = !partial || ((hash(v ^ it) >> 5) & 3) != 0; // optionally a random subset of lanes
        const uint32_t amask = __ballot_sync(0xffffffffu, active);
        const uint32_t before = cnt[warp][d];
        __syncwarp();
        uint32_t old = 0;
        if (active) old = atomicAdd(&cnt[warp][d], 1u);
        __syncwarp();
        const uint32_t m = __match_any_sync(0xffffffffu, d) & amask;
        if (active && old != before + __popc(m & lt)) ++mybad;
        __syncwarp();
    }
    atomicAdd(bad, mybad);
    if (tid == 0) atomicAdd(rows, (unsigned long long)iters * 16);
}
int main() {
    unsigned long long *bad, *rows; cudaMalloc(&bad, 8); cudaMalloc(&rows, 8);
    for (int partial = 0; partial < 2; ++partial)
    for (int stride : {1, 2, 32, 33})
    for (int distinct : {1, 2, 3, 5, 8, 16, 31, 256}) {
        cudaMemset(bad, 0, 8); cudaMemset(rows, 0, 8);
        k<<<148 * 2, 512>>>(bad, rows, distinct, stride, 20000, partial);
        unsigned long long hb, hr; cudaMemcpy(&hb, bad, 8, cudaMemcpyDeviceToHost); cudaMemcpy(&hr, rows, 8, cudaMemcpyDeviceToHost);
        printf("partial=%d stride=%2d distinct=%3d: rows=%llu  lanes out of lane order=%llu  (%s)\n", partial, stride, distinct, hr, hb, cudaGetErrorString(cudaGetLastError()));
    }
    return 0;
}
