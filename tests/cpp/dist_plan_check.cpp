// dist_plan_check -- CPU emulation of the device-side splitter planner of rs_dist_sort (radix-sort_b200/csrc/
// rs_dist_api.inl): the SAME planning functions (rs_dist_plan.h: round_query, select_bucket, compute_cuts) driven in the
// same order, with std::lower_bound for the binary-search kernel and a sum over the emulated ranks for ncclAllReduce /
// ncclAllGather.  Built as a small shared library; tests/test_dist_plan_cpu.py checks that the cuts it returns make
// rank r's slice exactly its share of the global order, for every distribution the gloo tests use.
#include "../../radix-sort_b200/csrc/rs_dist_plan.h"

#include <algorithm>
#include <cstdint>
#include <vector>

using namespace rsb::distplan;

template <typename U>
static uint64_t lower_bound_pos(const U* keys, uint64_t n, uint64_t q, bool is_top, U sign) {
    if (is_top) return n;
    const U x = static_cast<U>(q);
    uint64_t lo = 0, hi = n;
    while (lo < hi) {
        const uint64_t mid = lo + ((hi - lo) >> 1);
        if (static_cast<U>(keys[mid] ^ sign) < x) lo = mid + 1; else hi = mid;
    }
    return lo;
}

template <typename U>
static void emulate(int world, bool is_signed, const void* const* shards, const uint64_t* n, unsigned long long* cuts_out) {
    const int bits = static_cast<int>(sizeof(U)) * 8;
    const U   sign = is_signed ? static_cast<U>(static_cast<U>(1) << (bits - 1)) : static_cast<U>(0);
    std::vector<unsigned long long> prefix(world, 0), below(world, 0), targets(world, 0);
    const int levels = num_levels(bits);
    for (int level = 0; level < levels; ++level) {
        int width = 0, shift = 0;
        level_geometry(bits, level, &width, &shift);
        const int fan  = 1 << width;
        const int rows = level == 0 ? 1 : world - 1;
        std::vector<unsigned long long> counts(static_cast<size_t>(rows) * fan, 0);   // the all-reduced table
        for (int r = 0; r < world; ++r) {
            const U* keys = static_cast<const U*>(shards[r]);
            for (int b = 0; b < rows; ++b) {
                std::vector<uint64_t> pos(fan + 1);
                for (int i = 0; i <= fan; ++i) {
                    bool top;
                    const uint64_t q = round_query(level ? prefix[b] : 0ull, i, fan, shift, bits, &top);
                    pos[i] = lower_bound_pos<U>(keys, n[r], q, top, sign);
                }
                for (int i = 0; i < fan; ++i) counts[static_cast<size_t>(b) * fan + i] += pos[i + 1] - pos[i];
            }
        }
        for (int j = 0; j < world - 1; ++j) {   // rs_split_select_kernel
            const unsigned long long* row = counts.data() + (level == 0 ? 0 : static_cast<size_t>(j) * fan);
            if (level == 0) {
                unsigned long long total = 0;
                for (int i = 0; i < fan; ++i) total += row[i];
                targets[j] = (static_cast<unsigned long long>(j + 1) * total) / static_cast<unsigned long long>(world);
                prefix[j] = 0; below[j] = 0;
            }
            uint64_t nb = 0;
            const int b = select_bucket(row, fan, below[j], targets[j], &nb);
            prefix[j] += static_cast<unsigned long long>(b) << shift;
            below[j] = nb;
        }
    }
    // copies of every exact splitter value
    std::vector<unsigned long long> lb(static_cast<size_t>(world) * world, 0), eq_all(static_cast<size_t>(world) * world, 0);
    for (int r = 0; r < world; ++r) {
        const U* keys = static_cast<const U*>(shards[r]);
        for (int j = 0; j < world - 1; ++j) {
            bool t0, t1;
            const uint64_t q0 = round_query(prefix[j], 0, 1, 0, bits, &t0);
            const uint64_t q1 = round_query(prefix[j], 1, 1, 0, bits, &t1);
            const uint64_t p0 = lower_bound_pos<U>(keys, n[r], q0, t0, sign), p1 = lower_bound_pos<U>(keys, n[r], q1, t1, sign);
            lb[static_cast<size_t>(r) * world + j]           = p0;
            eq_all[static_cast<size_t>(r) * (world - 1) + j] = p1 - p0;
        }
    }
    for (int r = 0; r < world; ++r)
        compute_cuts(world, r, n[r], lb.data() + static_cast<size_t>(r) * world, eq_all.data(), targets.data(), below.data(),
                     cuts_out + static_cast<size_t>(r) * (world + 1));
}

extern "C" int distplan_emulate(int key_bytes, int is_signed, int world, const void* const* sorted_shards, const uint64_t* n,
                                unsigned long long* cuts_out /* [world][world + 1] */) {
    if (world < 1 || world > 16) return 1;
    if (world == 1) { cuts_out[0] = 0; cuts_out[1] = n[0]; return 0; }
    if (key_bytes == 4) emulate<uint32_t>(world, is_signed != 0, sorted_shards, n, cuts_out);
    else if (key_bytes == 8) emulate<uint64_t>(world, is_signed != 0, sorted_shards, n, cuts_out);
    else return 2;
    return 0;
}
