// rs_merge_order.h -- the order in which the persistent merge CTAs walk the tiles of a batch (plain C++, also
// compiled on the host by tests/cpp/merge_order_check.cpp).
//
// Tiles of a batch are numbered job-major: [tile_start[j], tile_start[j+1]) belong to job j.  CTA c handles the
// slots c, c + G, c + 2G, ...; slot p names tile
//       t(p) = (p mod R) * C + p / R,        C = ceil(T / R),        valid iff t < T
// i.e. consecutive slots fall into R different regions of the tile space.  That gives both properties the level-1
// merge over peer memory needs, whatever the relative job sizes and whatever G mod (number of jobs):
//   * at any time the G running CTAs work on all R regions, so every job -- every pair of source GPUs -- is read at
//     a rate proportional to its size (no source's NVLink egress is the bottleneck while others idle);
//   * every CTA gets the same number of slots, and at most R - 1 + (R * C - T) < 2R slots of the whole walk are empty.
// (The first version used slot -> job p mod J: with G a multiple of J a CTA never left its job, so a batch whose keys
// all sit in one job -- already range-partitioned inputs -- kept only G / J CTAs busy.)
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define RS_HD __host__ __device__ __forceinline__
#else
#define RS_HD inline
#endif

namespace rsb {

constexpr uint32_t MERGE_REGIONS = 64;

struct MergeOrder {
    uint32_t tiles;      // T
    uint32_t per_region; // C
    uint32_t slots;      // R * C
};

RS_HD MergeOrder merge_order_of(uint32_t tiles) {
    MergeOrder o;
    o.tiles      = tiles;
    o.per_region = (tiles + MERGE_REGIONS - 1) / MERGE_REGIONS;
    o.slots      = o.per_region * MERGE_REGIONS;   // <= tiles + 63 * ... fits: tiles <= 2^31 - 1 is checked by the host
    return o;
}

// tile of slot p (p < o.slots); >= o.tiles means "empty slot"
RS_HD uint32_t merge_tile_of_slot(const MergeOrder& o, uint32_t p) {
    return (p % MERGE_REGIONS) * o.per_region + p / MERGE_REGIONS;
}

}  // namespace rsb
