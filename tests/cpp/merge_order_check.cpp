// Host check of radix-sort_b200/csrc/rs_merge_order.h: every tile is named by exactly one slot, empty slots are few,
// and a strided walk (CTA c: slots c, c+G, ...) touches all regions early.  Exit code 0 = ok.
#include "rs_merge_order.h"

#include <cstdio>
#include <vector>

int main() {
    using namespace rsb;
    const uint32_t sizes[] = {1, 2, 63, 64, 65, 127, 128, 129, 1000, 4095, 4096, 4097, 65536, 262145, 1048576, 1234567};
    for (uint32_t tiles : sizes) {
        const MergeOrder o = merge_order_of(tiles);
        if (o.slots < tiles || o.slots - tiles >= MERGE_REGIONS) { std::printf("bad slot count: T=%u S=%u\n", tiles, o.slots); return 1; }
        std::vector<uint8_t> seen(tiles, 0);
        uint32_t empty = 0;
        for (uint32_t p = 0; p < o.slots; ++p) {
            const uint32_t t = merge_tile_of_slot(o, p);
            if (t >= tiles) { ++empty; continue; }
            if (seen[t]) { std::printf("tile %u named twice (T=%u)\n", t, tiles); return 1; }
            seen[t] = 1;
        }
        for (uint32_t t = 0; t < tiles; ++t)
            if (!seen[t]) { std::printf("tile %u never named (T=%u)\n", t, tiles); return 1; }
        if (empty != o.slots - tiles) { std::printf("empty slot count mismatch (T=%u)\n", tiles); return 1; }
        // the first 592 slots (one per resident CTA on a B200) cover every region that has tiles
        if (tiles >= MERGE_REGIONS * 2) {
            std::vector<uint8_t> region(MERGE_REGIONS, 0);
            for (uint32_t p = 0; p < 592 && p < o.slots; ++p) {
                const uint32_t t = merge_tile_of_slot(o, p);
                if (t < tiles) region[t / o.per_region] = 1;
            }
            for (uint32_t r = 0; r < MERGE_REGIONS && r * o.per_region < tiles; ++r)
                if (!region[r]) { std::printf("region %u idle at the start (T=%u)\n", r, tiles); return 1; }
        }
    }
    std::printf("merge order ok\n");
    return 0;
}
