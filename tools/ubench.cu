// Microbenchmarks for the warp primitives the ranking loop can be built from (dev tool, not product).
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 2048
__device__ __forceinline__ uint32_t hash(uint32_t x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }

template <int MODE>
__global__ void __launch_bounds__(512, 2) k(uint32_t* out, long long* cyc, int distinct) {
    __shared__ uint32_t sh[16 * 256];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 16 * 256; i += 512) sh[i] = 0;
    __syncthreads();
    uint32_t acc = 0;
    uint32_t v = hash(tid * 7919u + blockIdx.x);
    long long t0 = clock64();
#pragma unroll 4
    for (int it = 0; it < ITERS; ++it) {
        v = v * 1664525u + 1013904223u;
        uint32_t d = (v >> 11) % distinct;   // digit with `distinct` possible values
        if (MODE == 0) {            // MATCH.ANY
            acc += __match_any_sync(0xffffffffu, d);
        } else if (MODE == 1) {     // 8 ballots
            uint32_t m = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; ++b) { const bool p = (d >> b) & 1u; const uint32_t bal = __ballot_sync(0xffffffffu, p); m &= p ? bal : ~bal; }
            acc += m;
        } else if (MODE == 2) {     // SHFL
            acc += __shfl_sync(0xffffffffu, d, (d >> 3) & 31);
        } else if (MODE == 3) {     // ATOMS all lanes, warp-private table
            acc += atomicAdd(&sh[warp * 256 + (d & 255)], 1u);
        } else if (MODE == 4) {     // LDS random within warp table
            acc += sh[warp * 256 + (d & 255)];
        } else if (MODE == 5) {     // STS random (scatter)
            sh[(d * 97u + lane) & 4095] = d;
        } else if (MODE == 6) {     // REDUX or
            acc += __reduce_or_sync(0xffffffffu, d);
        } else if (MODE == 7) {     // baseline loop (digit generation only)
            acc += d;
        } else if (MODE == 8) {     // match.any on two half warps with separate masks
            acc += __match_any_sync(lane < 16 ? 0x0000ffffu : 0xffff0000u, d);
        } else if (MODE == 9) {     // atomicOr bitmap + read back + clear  (WARP_MATCH_ATOMIC_OR style)
            uint32_t* bm = &sh[warp * 256 + (d & 255)];
            atomicOr(bm, 1u << lane);
            __syncwarp();
            const uint32_t m = *bm;
            __syncwarp();
            *bm = 0;
            __syncwarp();
            acc += m;
        }
    }
    long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    out[blockIdx.x * 512 + tid] = acc;
}

template <int MODE>
void run(const char* name, int distinct, uint32_t* out, long long* cyc) {
    const int grid = 148 * 2;
    k<MODE><<<grid, 512>>>(out, cyc, distinct);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    cudaEventRecord(a);
    k<MODE><<<grid, 512>>>(out, cyc, distinct);
    cudaEventRecord(b);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, a, b);
    long long h[296]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < grid; ++i) avg += h[i]; avg /= grid;
    // per SM: 2 blocks x 16 warps x ITERS warp-instructions in `avg` cycles
    printf("%-28s distinct=%3d  %8.1f cycles/warp-op/SM   (%.3f ms, %.0f cyc/block)\n", name, distinct, avg / (2.0 * 16 * ITERS), ms, avg);
}

int main() {
    uint32_t* out; long long* cyc;
    cudaMalloc(&out, 296 * 512 * 4); cudaMalloc(&cyc, 296 * 8);
    for (int distinct : {1, 2, 16, 256}) {
        run<7>("baseline", distinct, out, cyc);
        run<0>("match.any", distinct, out, cyc);
        run<8>("match.any half-warp masks", distinct, out, cyc);
        run<1>("8x ballot", distinct, out, cyc);
        run<9>("atomicOr bitmap", distinct, out, cyc);
        run<2>("shfl", distinct, out, cyc);
        run<3>("atoms.add all lanes", distinct, out, cyc);
        run<4>("lds random", distinct, out, cyc);
        run<5>("sts scatter", distinct, out, cyc);
        run<6>("redux.or", distinct, out, cyc);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
