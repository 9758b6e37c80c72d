// rs_merge.cuh -- pairwise merge of sorted key runs (multi-GPU final step).
//
// After the exchange every GPU owns P sorted runs (one per source GPU).  SURVEY.md 8(e).5 allows the final
// local step to be "one more local onesweep ... or a P-way merge"; merging is ~4x less shared-memory work than
// four onesweep passes for P = 2 and still cheaper for P = 8 (log2 P levels).  The first level reads its runs
// straight from PEER memory (CUDA IPC over NVLink): exchange and merge are one kernel.
//
//   rs_merge_partition_kernel : one thread per output-tile boundary, merge-path binary search (global loads)
//   rs_merge_tiles_kernel     : one CTA per output tile of MERGE_TILE keys: load the two input pieces into
//                               shared memory, per-thread merge-path split, MERGE_ITEMS-way serial merge in
//                               registers, padded shared-memory transpose, coalesced store
// Keys only (the distributed configuration is keys-only); order is the key type's (sign bit flipped for signed).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace rsb {

constexpr int MERGE_THREADS  = 256;
constexpr int MERGE_ITEMS    = 16;
constexpr int MERGE_TILE     = MERGE_THREADS * MERGE_ITEMS;  // 4096 outputs per CTA
constexpr int MERGE_MAX_JOBS = 8;

template <typename KeyT>
struct MergeBatch {
    int                num_jobs;
    const KeyT*        a[MERGE_MAX_JOBS];
    const KeyT*        b[MERGE_MAX_JOBS];
    KeyT*              out[MERGE_MAX_JOBS];
    unsigned long long na[MERGE_MAX_JOBS];
    unsigned long long nb[MERGE_MAX_JOBS];
    unsigned long long tile_start[MERGE_MAX_JOBS + 1];  // first tile of every job; [num_jobs] = total tiles
    KeyT               sign;                            // xor mask that maps keys to unsigned order
};

template <typename KeyT>
__device__ __forceinline__ void merge_job_of(const MergeBatch<KeyT>& mb, unsigned long long tile, int& job,
                                             const KeyT*& a, const KeyT*& b, KeyT*& out, unsigned long long& na,
                                             unsigned long long& nb, unsigned long long& first_tile) {
    job = 0;
#pragma unroll
    for (int j = 1; j < MERGE_MAX_JOBS; ++j)
        if (j < mb.num_jobs && tile >= mb.tile_start[j]) job = j;
    a = nullptr; b = nullptr; out = nullptr; na = 0; nb = 0; first_tile = 0;
#pragma unroll
    for (int j = 0; j < MERGE_MAX_JOBS; ++j)
        if (j == job) { a = mb.a[j]; b = mb.b[j]; out = mb.out[j]; na = mb.na[j]; nb = mb.nb[j]; first_tile = mb.tile_start[j]; }
}

// number of elements taken from A among the first `diag` outputs of merge(A, B); ties take A first
template <typename KeyT, typename IndexT>
__device__ __forceinline__ IndexT merge_path(const KeyT* a, IndexT na, const KeyT* b, IndexT nb, IndexT diag, KeyT sign) {
    IndexT lo = diag > nb ? diag - nb : 0;
    IndexT hi = diag < na ? diag : na;
    while (lo < hi) {
        const IndexT mid = lo + ((hi - lo) >> 1);
        const KeyT   ka  = a[mid] ^ sign;
        const KeyT   kb  = b[diag - 1 - mid] ^ sign;
        if (ka <= kb) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// splits[job_first_boundary + i] = A-count of the first i*MERGE_TILE outputs of the job, i = 0 .. tiles_of_job
template <typename KeyT>
__global__ void rs_merge_partition_kernel(const MergeBatch<KeyT> mb, unsigned long long* __restrict__ splits) {
    const unsigned long long total_boundaries = mb.tile_start[mb.num_jobs] + mb.num_jobs;
    for (unsigned long long g = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; g < total_boundaries;
         g += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        // boundary g of the batch: job j owns boundaries [tile_start[j] + j, tile_start[j+1] + j + 1)
        int job = 0;
#pragma unroll
        for (int j = 1; j < MERGE_MAX_JOBS; ++j)
            if (j < mb.num_jobs && g >= mb.tile_start[j] + j) job = j;
        const KeyT* a = nullptr; const KeyT* b = nullptr; unsigned long long na = 0, nb = 0, first = 0;
#pragma unroll
        for (int j = 0; j < MERGE_MAX_JOBS; ++j)
            if (j == job) { a = mb.a[j]; b = mb.b[j]; na = mb.na[j]; nb = mb.nb[j]; first = mb.tile_start[j] + j; }
        unsigned long long diag = (g - first) * MERGE_TILE;
        if (diag > na + nb) diag = na + nb;
        splits[g] = merge_path<KeyT, unsigned long long>(a, na, b, nb, diag, mb.sign);
    }
}

__device__ __forceinline__ uint32_t merge_pad(uint32_t i) { return i + (i >> 5); }  // conflict-free 16-stride transpose

template <typename KeyT>
__global__ void __launch_bounds__(MERGE_THREADS)
rs_merge_tiles_kernel(const MergeBatch<KeyT> mb, const unsigned long long* __restrict__ splits) {
    extern __shared__ __align__(16) unsigned char merge_smem[];
    KeyT* const sk = reinterpret_cast<KeyT*>(merge_smem);  // [MERGE_TILE + MERGE_TILE/32]
    const int                tid  = threadIdx.x;
    const unsigned long long tile = blockIdx.x;
    int job; const KeyT* a; const KeyT* b; KeyT* out; unsigned long long na, nb, first_tile;
    merge_job_of(mb, tile, job, a, b, out, na, nb, first_tile);
    const unsigned long long lt    = tile - first_tile;                 // tile index within the job
    const unsigned long long total = na + nb;
    const unsigned long long d0    = lt * MERGE_TILE;
    const unsigned long long d1    = (d0 + MERGE_TILE < total) ? d0 + MERGE_TILE : total;
    const unsigned long long a0    = splits[tile + job], a1 = splits[tile + job + 1];
    const unsigned long long b0    = d0 - a0, b1 = d1 - a1;
    const uint32_t           la    = static_cast<uint32_t>(a1 - a0), lb = static_cast<uint32_t>(b1 - b0);
    const KeyT               sign  = mb.sign;

    // the two input pieces, back to back, in shared memory (coalesced; A and B may be peer memory)
    for (uint32_t i = tid; i < la; i += MERGE_THREADS) sk[i] = a[a0 + i];
    for (uint32_t i = tid; i < lb; i += MERGE_THREADS) sk[la + i] = b[b0 + i];
    __syncthreads();

    // this thread's MERGE_ITEMS outputs start at local diagonal `diag`
    const uint32_t n_out = la + lb;
    uint32_t       diag  = tid * MERGE_ITEMS;
    if (diag > n_out) diag = n_out;
    uint32_t ai = merge_path<KeyT, uint32_t>(sk, la, sk + la, lb, diag, sign);
    uint32_t bi = diag - ai;
    KeyT     r[MERGE_ITEMS];
    KeyT     ka = (ai < la) ? sk[ai] : static_cast<KeyT>(0);
    KeyT     kb = (bi < lb) ? sk[la + bi] : static_cast<KeyT>(0);
#pragma unroll
    for (int k = 0; k < MERGE_ITEMS; ++k) {
        const bool take_a = (bi >= lb) || (ai < la && (ka ^ sign) <= (kb ^ sign));
        r[k] = take_a ? ka : kb;
        if (take_a) { ++ai; ka = (ai < la) ? sk[ai] : static_cast<KeyT>(0); }
        else        { ++bi; kb = (bi < lb) ? sk[la + bi] : static_cast<KeyT>(0); }
    }
    __syncthreads();  // every thread is done reading the inputs
#pragma unroll
    for (int k = 0; k < MERGE_ITEMS; ++k) sk[merge_pad(tid * MERGE_ITEMS + k)] = r[k];
    __syncthreads();
    KeyT* const dst = out + d0;
    for (uint32_t i = tid; i < n_out; i += MERGE_THREADS) dst[i] = sk[merge_pad(i)];
}

}  // namespace rsb
