// rs_dist_plan.h -- splitter planning of the multi-GPU sort, as plain functions shared by the device kernels
// (rs_dist.cuh) and the host (CPU tests: tests/cpp/dist_plan_check.cpp emulates P ranks with them and compares the
// cuts with radix-sort_b200/dist.py:plan_exchange, which the gloo tests cover).
//
// Path (BASELINE.json north_star; the reference has no counterpart -- one OpenCL device,
// Common/ComputeState.cpp:10-12): every rank holds a SORTED shard.  The key space is refined LEVEL_BITS bits per round:
// every rank counts its keys in the 2^LEVEL_BITS sub-ranges of the current prefix of every splitter (binary searches in
// its sorted shard), the counts are all-reduced, and every rank picks -- identically -- the sub-range that holds the
// global rank j*N/P of splitter j.  After ceil(bits / LEVEL_BITS) rounds every splitter is an exact key value v_j; the
// copies of v_j are split by count in rank order (equal keys are indistinguishable), so rank j ends with exactly
// floor((j+1)N/P) - floor(jN/P) keys whatever the distribution.
// All key values here are "ordered unsigned": the key bits with the sign bit flipped for signed types.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define RS_PLAN_HD __host__ __device__
#else
#define RS_PLAN_HD
#endif

namespace rsb {
namespace distplan {

constexpr int LEVEL_BITS = 11;  // fan-out 2048 per round: 32-bit keys take 11 + 11 + 10 bits

RS_PLAN_HD inline uint64_t key_mask(int bits) { return bits >= 64 ? ~0ull : ((1ull << bits) - 1ull); }

RS_PLAN_HD inline int num_levels(int bits) { return (bits + LEVEL_BITS - 1) / LEVEL_BITS; }
// geometry of refinement round `level` (0-based): `width` bits are resolved, sub-range i starts at prefix + (i << shift)
RS_PLAN_HD inline void level_geometry(int bits, int level, int* width, int* shift) {
    const int resolved = level * LEVEL_BITS;
    const int w        = bits - resolved < LEVEL_BITS ? bits - resolved : LEVEL_BITS;
    *width             = w;
    *shift             = bits - resolved - w;
}

// Boundary i (0 .. fan) of the sub-ranges of `prefix`; *is_top: the boundary is 2^bits itself (one past the largest
// key), whose lower bound in any shard is the shard's length.
RS_PLAN_HD inline uint64_t round_query(uint64_t prefix, int i, int fan, int shift, int bits, bool* is_top) {
    const uint64_t q = prefix + (static_cast<uint64_t>(i) << shift);  // wraps to 0 for 64-bit keys exactly when it is 2^64
    *is_top          = (i == fan) && ((q & key_mask(bits)) == 0);
    return q & key_mask(bits);
}

// The sub-range b with cum[b] <= target < cum[b + 1], cum = below + prefix sums of the global counts; *new_below = cum[b].
RS_PLAN_HD inline int select_bucket(const unsigned long long* counts, int fan, uint64_t below, uint64_t target, uint64_t* new_below) {
    uint64_t cum   = below;
    int      b     = fan - 1;
    uint64_t at    = 0;
    bool     found = false;
    for (int i = 0; i < fan; ++i) {
        const uint64_t next = cum + counts[i];
        if (!found && target < next) { b = i; at = cum; found = true; }
        cum = next;
    }
    if (!found) at = cum - counts[fan - 1];  // only for an empty input: the last sub-range
    *new_below = at;
    return b;
}

// Cut positions of one rank's sorted shard: rank d receives [cuts[d], cuts[d+1]).
// lb[j] = first local index with key >= v_j, eq_all[r][j] = copies of v_j on rank r, targets[j] = global rank of the first
// key of rank j+1, below[j] = keys globally below v_j.
RS_PLAN_HD inline void compute_cuts(int world, int rank, uint64_t n_local, const unsigned long long* lb, const unsigned long long* eq_all,
                               const unsigned long long* targets, const unsigned long long* below, unsigned long long* cuts) {
    cuts[0] = 0;
    for (int j = 0; j < world - 1; ++j) {
        const uint64_t need_left = targets[j] - below[j];  // copies of v_j that belong to ranks <= j
        uint64_t       before_me = 0;
        for (int r = 0; r < rank; ++r) before_me += eq_all[r * (world - 1) + j];
        const uint64_t mine  = eq_all[rank * (world - 1) + j];
        uint64_t       share = need_left > before_me ? need_left - before_me : 0;
        if (share > mine) share = mine;
        cuts[j + 1] = lb[j] + share;
    }
    cuts[world] = n_local;
    for (int j = 1; j <= world; ++j)  // equal splitters (heavy duplicates) keep the cuts monotone
        if (cuts[j] < cuts[j - 1]) cuts[j] = cuts[j - 1];
}

}  // namespace distplan
}  // namespace rsb
