// rs_dist.cuh -- device side of the splitter planning of the multi-GPU sort (see rs_dist_plan.h for the algorithm).
// All kernels are tiny (a few thousand binary searches, one thread per splitter for the selection): what matters is
// that the refinement state never leaves the device -- a round is three small launches and one ncclAllReduce on the
// same stream, no host round trip (the first generation paid ~10 D2H/H2D round trips per sort, 1.9 ms at 8 GPUs).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "rs_dist_plan.h"

namespace rsb {

// Refinement state of one rank, resident on the device (MAX_WORLD - 1 splitters).
constexpr int DIST_MAX_WORLD = 16;
struct DistPlanState {
    unsigned long long prefix[DIST_MAX_WORLD];   // per splitter: ordered-unsigned key bits resolved so far
    unsigned long long below[DIST_MAX_WORLD];    // per splitter: number of keys globally below `prefix`
    unsigned long long targets[DIST_MAX_WORLD];  // per splitter: global rank of the first key of rank j + 1
    unsigned long long total;                    // N (known after round 0)
};

// pos[b][i] = lower bound of boundary i (0 .. fan) of splitter-row b's current sub-ranges in the local sorted shard.
// rows = 1 in round 0 (one histogram of the whole key space serves every splitter), world - 1 afterwards.
template <typename U>
__global__ void rs_split_positions_kernel(const U* __restrict__ keys, unsigned long long n, const DistPlanState* __restrict__ state,
                                          int rows, int use_prefix, int fan, int shift, int bits, U sign,
                                          unsigned long long* __restrict__ pos) {
    const unsigned long long total = static_cast<unsigned long long>(rows) * (fan + 1);
    for (unsigned long long g = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; g < total;
         g += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        const int b = static_cast<int>(g / (fan + 1));
        const int i = static_cast<int>(g % (fan + 1));
        bool      is_top;
        const unsigned long long q = distplan::round_query(use_prefix ? state->prefix[b] : 0ull, i, fan, shift, bits, &is_top);
        unsigned long long lo = 0, hi = n;
        if (is_top) {
            lo = n;
        } else {
            const U x = static_cast<U>(q);
            while (lo < hi) {
                const unsigned long long mid = lo + ((hi - lo) >> 1);
                if (static_cast<U>(keys[mid] ^ sign) < x) lo = mid + 1; else hi = mid;
            }
        }
        pos[g] = lo;
    }
}

// counts[b][i] = pos[b][i + 1] - pos[b][i]
__global__ void rs_split_counts_kernel(const unsigned long long* __restrict__ pos, int rows, int fan,
                                       unsigned long long* __restrict__ counts) {
    const int total = rows * fan;
    for (int g = blockIdx.x * blockDim.x + threadIdx.x; g < total; g += gridDim.x * blockDim.x) {
        const int b = g / fan, i = g % fan;
        counts[g] = pos[b * (fan + 1) + i + 1] - pos[b * (fan + 1) + i];
    }
}

// After the all-reduce: every splitter descends into the sub-range that holds its target rank.
__global__ void rs_split_select_kernel(const unsigned long long* __restrict__ counts, int level, int fan, int shift, int world,
                                       DistPlanState* __restrict__ state) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= world - 1) return;
    const unsigned long long* row = counts + (level == 0 ? 0 : static_cast<size_t>(j) * fan);
    if (level == 0) {
        unsigned long long total = 0;
        for (int i = 0; i < fan; ++i) total += row[i];
        if (j == 0) state->total = total;
        state->targets[j] = (static_cast<unsigned long long>(j + 1) * total) / static_cast<unsigned long long>(world);
        state->prefix[j]  = 0;
        state->below[j]   = 0;
    }
    uint64_t  nb = 0;
    const int b  = distplan::select_bucket(row, fan, state->below[j], state->targets[j], &nb);
    state->prefix[j] += static_cast<unsigned long long>(b) << shift;
    state->below[j] = nb;
}

// eq[j] = copies of the exact splitter value v_j in the local shard (pos holds lower bounds of v_j and v_j + 1)
__global__ void rs_split_equal_kernel(const unsigned long long* __restrict__ pos, int world, unsigned long long* __restrict__ eq) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j < world - 1) eq[j] = pos[2 * j + 1] - pos[2 * j];
}

// mine[0 .. world] = cut positions of this rank's shard, mine[world + 1] = which buffer of the local pair holds it
__global__ void rs_split_cuts_kernel(const unsigned long long* __restrict__ pos, const unsigned long long* __restrict__ eq_all,
                                     const DistPlanState* __restrict__ state, int world, int rank, unsigned long long n_local,
                                     unsigned long long sel, unsigned long long* __restrict__ mine) {
    if (blockIdx.x != 0 || threadIdx.x != 0) return;
    unsigned long long lb[DIST_MAX_WORLD];
    for (int j = 0; j < world - 1; ++j) lb[j] = pos[2 * j];
    distplan::compute_cuts(world, rank, n_local, lb, eq_all, state->targets, state->below, mine);
    mine[world + 1] = sel;
}

}  // namespace rsb
