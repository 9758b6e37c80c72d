// rs_sweep.cuh -- the onesweep pass kernel of the default (atomic ranking) mode, second generation.
//
// Same algorithm and same global data structures as rs_onesweep_kernel in rs_kernels.cuh (which stays as the
// lane-bitmap fallback and as the A/B baseline, RSB200_SWEEP_V1=1): persistent CTAs, dynamic tile ids, TMA 1-D
// prefetch of the next tile, count atomics -> table scan -> position atomics while scattering into the stage ->
// decoupled look-back -> coalesced per-digit runs.  Replaces the reference's `reorder` kernel
// (src/kernels/RadixSort.cl:74-119) plus the per-tile part of scanhistograms / pastehistograms (:125-197).
//
// What changed, and why (ncu of the first generation, profiles/r1d_*: LSU data pipe 84 % busy with shared-memory
// wavefronts, 22 % of the stall samples at barriers, 9 % in look-back loads; 17.1 shared + 2.7 global wavefronts
// per 32 keys):
//   * the per-(warp, digit) counters are 16 bits wide, two warps per 32-bit word ({warp 2j : low, warp 2j+1 : high}):
//     the table scan and its re-zeroing touch half the words (1.3 -> 0.65 wavefronts per 32 keys), the table is
//     8 KB instead of 16 KB;
//   * the per-digit run pointers are 32-bit element offsets (a second word with the high half only when the pass
//     moves more than 2^32 keys): the data-dependent lookup in the output loop is a 32-bit load;
//   * roles: threads [0,256) scan the tables of ALL warps for one digit each (no partial totals, no extra barrier)
//     while threads [256,512) walk the look-back chain of the same digits AT THE SAME TIME -- the walk only needs
//     the predecessors' words, not this tile's counts.  The two latency-bound sections of a tile (table scan and
//     look-back round trips, about 1 us each) overlap each other instead of adding up, and four CTA-wide barriers
//     are left per tile (A: counts final, C: positions ready, D: stage ranked, E: stage free) instead of six.
//
// Stability rests on the same hardware property as before: lanes of ONE warp instruction that hit the same shared
// word are served in ascending lane order (self-tested per context, see rs_api.cu); two warps that share a word
// update disjoint 16-bit halves, and a half can never carry into its neighbour (every value is a tile position
// <= TILE <= 65535).
// Algorithmic traffic: sizeof(KeyT) read + sizeof(KeyT) written per key (+4 +4 with the index payload).
#pragma once
#include "rs_kernels.cuh"

namespace rsb {

// When the look-back warps walk the chain: 0 = right after (A), concurrently with the table scan; 1 = after (C), before
// their own share of the scatter (the other warps scatter meanwhile); 2 = after their scatter (the first generation's timing).
#ifndef RS_LB_MODE
#define RS_LB_MODE 1
#endif
constexpr int LB_MODE = RS_LB_MODE;
// Table layout: 1 = 16-bit counters, two warps per word (half the words to scan and re-zero, but the count atomic of
// the odd warps adds 65536 and cannot be the lane-merging ATOMS.POPC.INC); 0 = one 32-bit word per (warp, digit).
#ifndef RS_SWEEP_PACK16
#define RS_SWEEP_PACK16 1
#endif
constexpr bool SWEEP_PACK16 = RS_SWEEP_PACK16 != 0;

template <typename KeyT, bool HAS_VALS, int THREADS, int KPT>
struct SweepSmem {
    static constexpr int    TILE   = THREADS * KPT;
    static constexpr int    NWARPS = THREADS / 32;
    static constexpr int    GW     = SWEEP_PACK16 ? NWARPS / 2 : NWARPS;  // table words per digit
    static constexpr bool   SPLIT  = THREADS >= 2 * RADIX;   // scan and look-back run on different threads
    static constexpr size_t stage_bytes = sizeof(KeyT) * TILE;
    static constexpr size_t off_stage0  = 0;
    static constexpr size_t off_stage1  = off_stage0 + stage_bytes;
    static constexpr size_t off_vstage  = off_stage1 + stage_bytes;
    static constexpr size_t off_tab     = off_vstage + (HAS_VALS ? sizeof(uint32_t) * TILE : 0);
    static constexpr size_t off_koff    = off_tab + sizeof(uint32_t) * GW * RADIX;
    static constexpr size_t off_khi     = off_koff + sizeof(uint32_t) * RADIX;
    static constexpr size_t off_meta    = off_khi + sizeof(uint32_t) * RADIX;
    static constexpr size_t off_wsum    = off_meta + sizeof(uint32_t) * RADIX;
    static constexpr size_t off_mbar    = off_wsum + sizeof(uint32_t) * 16;
    static constexpr size_t off_tile    = off_mbar + sizeof(uint64_t) * 2;
    static constexpr size_t total       = off_tile + sizeof(uint32_t) * 4;
};

// constant digit: the pass is the identity -- nothing to do, or a plain copy to keep the ping-pong parity.
// Returns true when the calling CTA is done.
template <typename KeyT, bool HAS_VALS, int THREADS>
__device__ __forceinline__ bool sweep_identity_pass(const OnesweepArgs<KeyT>& a) {
    const uint32_t action = a.plan->action[a.pass];
    if (action == PASS_NOTHING) return true;
    if (action == PASS_COPY) {
        const uint64_t gtid = blockIdx.x * static_cast<uint64_t>(THREADS) + threadIdx.x;
        const uint64_t gstr = static_cast<uint64_t>(gridDim.x) * THREADS;
        constexpr int  KV   = 16 / sizeof(KeyT);
        const uint64_t nvec = a.use_tma ? a.n / KV : 0;  // use_tma == every key pointer is 16-byte aligned
        for (uint64_t i = gtid; i < nvec; i += gstr)
            reinterpret_cast<uint4*>(a.keys_out)[i] = __ldg(reinterpret_cast<const uint4*>(a.keys_in) + i);
        for (uint64_t i = nvec * KV + gtid; i < a.n; i += gstr) a.keys_out[i] = a.keys_in[i];
        if constexpr (HAS_VALS)
            for (uint64_t i = gtid; i < a.n; i += gstr) a.vals_out[i] = a.vals_in[i];
        return true;
    }
    return false;
}

template <typename KeyT, bool HAS_VALS, int THREADS, int KPT, int MIN_BLOCKS, bool FK>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
rs_sweep_kernel(const OnesweepArgs<KeyT> a) {
    using L = SweepSmem<KeyT, HAS_VALS, THREADS, KPT>;
    constexpr int  TILE  = L::TILE;
    constexpr int  GW    = L::GW;
    constexpr bool SPLIT = L::SPLIT;
    static_assert(THREADS % RADIX == 0, "THREADS must be a multiple of 256");
    static_assert((sizeof(KeyT) * TILE) % 128 == 0, "stage must keep 128B alignment");
    static_assert(TILE <= 0xffff, "16-bit count / position fields");
    static_assert(KPT <= 32, "uniform_rows is a 32-bit mask");

    if (sweep_identity_pass<KeyT, HAS_VALS, THREADS>(a)) return;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    KeyT* const        stage0  = reinterpret_cast<KeyT*>(smem_raw + L::off_stage0);  // stage s = stage0 + s*TILE
    uint32_t* const    vstage  = reinterpret_cast<uint32_t*>(smem_raw + L::off_vstage);
    uint32_t* const    tab     = reinterpret_cast<uint32_t*>(smem_raw + L::off_tab);   // [warp pair][digit], 2 x 16 bit
    uint32_t* const    koff    = reinterpret_cast<uint32_t*>(smem_raw + L::off_koff);  // [digit] run start - tile pos (low)
    uint32_t* const    khi     = reinterpret_cast<uint32_t*>(smem_raw + L::off_khi);   // high half, only if a.n > 2^32
    uint32_t* const    meta    = reinterpret_cast<uint32_t*>(smem_raw + L::off_meta);  // [digit] count | local_base << 16
    uint32_t* const    wsum    = reinterpret_cast<uint32_t*>(smem_raw + L::off_wsum);
    uint64_t* const    mbar    = reinterpret_cast<uint64_t*>(smem_raw + L::off_mbar);
    volatile uint32_t* sh_tile = reinterpret_cast<volatile uint32_t*>(smem_raw + L::off_tile);

    const int       tid    = threadIdx.x;
    const int       lane   = tid & 31;
    const int       warp   = tid >> 5;
    constexpr bool  PACK   = SWEEP_PACK16;
    const uint32_t  hshift = PACK ? (warp & 1) * 16 : 0;      // this warp's half of its table words
    const uint32_t  one    = 1u << hshift;
    const uint32_t  hsel   = hshift ? 0x4432u : 0x4410u;      // PRMT selector that extracts the half
    const DigitSpec ds     = a.digit;
    const uint32_t  full_tiles = static_cast<uint32_t>(a.n / TILE);  // tiles [0, full_tiles) are complete
    uint32_t* const gt     = tab + (PACK ? (warp >> 1) : warp) * RADIX;
    const int       sd     = tid & (RADIX - 1);               // digit of this thread's scan / look-back role
    const int       role   = tid / RADIX;                     // 0: table scan, 1: look-back (SPLIT), else: none
    const bool      wide   = a.n > (1ull << 32);

    auto issue_prefetch = [&](uint32_t t, int s) {
        // thread 0 only: bulk copy of a complete tile into stage s
        if (a.use_tma && t < full_tiles) {
            fence_proxy_async_smem();  // earlier generic-proxy writes to this stage are ordered before the copy
            mbar_expect_tx(&mbar[s], static_cast<uint32_t>(L::stage_bytes));
            bulk_load_g2s(stage0 + s * TILE, a.keys_in + static_cast<uint64_t>(t) * TILE,
                          static_cast<uint32_t>(L::stage_bytes), &mbar[s]);
        }
    };

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_fence_init();
        const uint32_t t0 = atomicAdd(a.counter, 1u);
        sh_tile[0]        = t0;
        if (t0 < a.num_tiles) issue_prefetch(t0, 0);
    }
    // the tables start clean; every tile re-zeroes them itself after its scatter
    for (int i = tid; i < GW * RADIX; i += THREADS) tab[i] = 0u;
    __syncthreads();

    uint32_t phase_bits = 0;  // bit s = parity to wait for on mbar[s]
    int      s          = 0;

    auto process_tile = [&](auto full_c, const uint32_t t) {
        constexpr bool FULL      = decltype(full_c)::value;
        const bool     via_tma   = FULL && a.use_tma;
        const uint64_t tile_base = static_cast<uint64_t>(t) * TILE;
        const uint32_t valid     = FULL ? TILE : static_cast<uint32_t>(a.n - tile_base);
        const uint32_t my_first  = warp * (32 * KPT) + lane;  // warp-striped: item r sits at my_first + 32*r
        KeyT* const    stg       = stage0 + s * TILE;

        // ---- keys (and values) into registers -------------------------------------------------
        KeyT key[KPT];
        if (via_tma) {
            mbar_wait(&mbar[s], (phase_bits >> s) & 1u);
            phase_bits ^= (1u << s);
#pragma unroll
            for (int r = 0; r < KPT; ++r) key[r] = stg[my_first + 32 * r];
        } else {
#pragma unroll
            for (int r = 0; r < KPT; ++r) {
                const uint32_t i = my_first + 32 * r;
                key[r] = (FULL || i < valid) ? a.keys_in[tile_base + i] : static_cast<KeyT>(~static_cast<KeyT>(0));
            }
        }
        uint32_t val[HAS_VALS ? KPT : 1];
        if constexpr (HAS_VALS) {
#pragma unroll
            for (int r = 0; r < KPT; ++r) {
                const uint32_t i = my_first + 32 * r;
                val[r]           = (FULL || i < valid) ? a.vals_in[tile_base + i] : 0u;
            }
        }

        // ---- count: one shared atomic per key, no return value ------------------------------------
        uint32_t uniform_rows = 0;  // bit r = all 32 digits of row r are equal
#pragma unroll
        for (int r = 0; r < KPT; ++r) {
            uint32_t d = digit_of<FK>(key[r], ds);
            if (!FULL && (my_first + 32 * r) >= valid) d = RADIX - 1;  // padding ranks last
            int all_same;
            __match_all_sync(0xffffffffu, d, &all_same);
            if (all_same) {  // 32 equal digits would serialise on one word: one lane counts them all
                uniform_rows |= 1u << r;
                if (lane == 0) atomicAdd(&gt[d], 32u << hshift);
            } else if constexpr (PACK) {
                atomicAdd(&gt[d], one);
            } else {
                atomicAdd(&gt[d], 1u);  // literal 1: ATOMS.POPC.INC, lanes on the same word are merged
            }
        }
        __syncthreads();  // (A) all counts final; every key of this stage is in registers

        const uint32_t chunk          = t / a.tiles_per_chunk;
        const uint32_t in_chunk       = t % a.tiles_per_chunk;   // predecessors of this tile in its chunk
        const bool     first_in_chunk = in_chunk == 0;
        uint32_t* const           my_status = a.status + static_cast<size_t>(t) * RADIX + sd;
        unsigned long long* const cb_ptr    = a.base + static_cast<size_t>(chunk) * RADIX + sd;
        uint32_t           count = 0, local_base = 0;  // scan role
        uint32_t           excl  = 0;                  // look-back role: sum of the predecessors' counts in this chunk
        unsigned long long cb    = 0;                  //                 global offset of the chunk's first key with digit sd

        // decoupled look-back over the predecessors of this chunk, LB_WINDOW status words per round trip
        auto look_back = [&]() {
            if (!first_in_chunk) {
                int32_t         left = static_cast<int32_t>(in_chunk);
                const uint32_t* p    = my_status - RADIX;
                bool            done = false;
                while (!done) {
                    uint32_t v[LB_WINDOW];
#pragma unroll
                    for (int j = 0; j < LB_WINDOW; ++j)
                        v[j] = (j < left) ? ld_status(p - j * RADIX) : STATUS_INCLUSIVE;  // before the chunk: prefix 0
                    bool unready = false;
#pragma unroll
                    for (int j = 0; j < LB_WINDOW; ++j) {
                        if (done || unready) break;
                        const uint32_t f = v[j] >> 30;
                        if (f == 0) { unready = true; break; }  // not published yet: resume from here
                        excl += v[j] & STATUS_VALUE_MASK;
                        p -= RADIX;
                        --left;
                        if (f == 2) done = true;
                    }
                    // wait for that one word (one load per poll, not a window of them), then read the next window
                    if (unready) while ((ld_status(p) >> 30) == 0) { }
                }
            }
            // global offset of this chunk's first key with this digit (ready at once for chunk 0)
            do { cb = ld_base(cb_ptr); } while (!(cb & BASE_READY));
            cb &= ~BASE_READY;
        };
        // publish the inclusive prefix, hand the next chunk its base, set the run offset of digit sd
        auto finish_look_back = [&](uint32_t cnt, uint32_t lbase) {
            if (!first_in_chunk) st_status(my_status, STATUS_INCLUSIVE | (excl + cnt));
            if (in_chunk == a.tiles_per_chunk - 1 && chunk + 1 < a.num_chunks)
                st_base(cb_ptr + RADIX, BASE_READY | (cb + excl + cnt));
            // run start minus the run's tile-local start: the element at tile position p goes to out[run0 + p]
            const unsigned long long run0 = cb + excl - lbase;  // may wrap below zero; run0 + p never does
            koff[sd] = static_cast<uint32_t>(run0);
            if (wide) khi[sd] = static_cast<uint32_t>(run0 >> 32);
        };

        if (role == 0) {
            // ---- scan: digit sd of every warp's counters ------------------------------------------------
            constexpr int NPRE = L::NWARPS / 2;
            uint32_t pre[NPRE];  // exclusive prefixes over the warps, two 16-bit values per register
            uint32_t run = 0;
#pragma unroll
            for (int j = 0; j < NPRE; ++j) {
                uint32_t lo, hi;
                if constexpr (PACK) {
                    const uint32_t w = tab[j * RADIX + sd];
                    lo = w & 0xffffu; hi = w >> 16;
                } else {
                    lo = tab[(2 * j) * RADIX + sd]; hi = tab[(2 * j + 1) * RADIX + sd];
                }
                pre[j] = run | ((run + lo) << 16);
                run += lo + hi;
            }
            count = run;
            if (!FULL && sd == RADIX - 1) count -= (TILE - valid);
            st_status(my_status, (first_in_chunk ? STATUS_INCLUSIVE : STATUS_LOCAL) | count);
            uint32_t incl = count;  // inclusive scan over the 256 digits
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += x;
            }
            if (lane == 31) wsum[warp] = incl;
            if constexpr (SPLIT) asm volatile("bar.sync 1, 256;" ::: "memory");  // the 8 scan warps only
            else                 __syncthreads();
            uint32_t woff = 0;
#pragma unroll
            for (int w = 0; w < RADIX / 32; ++w) woff += (w < warp) ? wsum[w] : 0u;
            local_base = woff + incl - count;  // tile-local start of digit sd
#pragma unroll
            for (int j = 0; j < NPRE; ++j) {
                if constexpr (PACK) {
                    tab[j * RADIX + sd] = pre[j] + local_base * 0x10001u;
                } else {
                    tab[(2 * j) * RADIX + sd]     = (pre[j] & 0xffffu) + local_base;
                    tab[(2 * j + 1) * RADIX + sd] = (pre[j] >> 16) + local_base;
                }
            }
            if constexpr (SPLIT) meta[sd] = count | (local_base << 16);
        } else if (SPLIT && role == 1 && LB_MODE == 0) {
            look_back();
        }
        __syncthreads();  // (C) a table half = tile-local position of that warp's first key with that digit
        if (SPLIT && role == 1 && LB_MODE <= 1) {
            if (LB_MODE == 1) look_back();
            const uint32_t m = meta[sd];
            finish_look_back(m & 0xffffu, m >> 16);
        }

        // ---- scatter into shared memory in ranked order -----------------------------------------
#pragma unroll
        for (int r = 0; r < KPT; ++r) {
            uint32_t d = digit_of<FK>(key[r], ds);
            if (!FULL && (my_first + 32 * r) >= valid) d = RADIX - 1;
            // the table half is the running tile position of (this warp, digit d).  Lanes that hit the same word are
            // served in ascending lane order (self-tested), so the returned value is the stable position of this key.
            uint32_t pos;
            if (uniform_rows & (1u << r)) {
                uint32_t c = 0;
                if (lane == 0) c = atoms_add_return(&gt[d], 32u << hshift);
                c   = __shfl_sync(0xffffffffu, c, 0);
                pos = (PACK ? __byte_perm(c, 0u, hsel) : c) + lane;
            } else {
                const uint32_t w = atoms_add_return(&gt[d], one);
                pos = PACK ? __byte_perm(w, 0u, hsel) : w;
            }
            stg[swz(pos)] = key[r];
            if constexpr (HAS_VALS) vstage[swz(pos)] = val[r];
        }
        if constexpr (!SPLIT) {
            look_back();
            finish_look_back(count, local_base);
        } else if (LB_MODE == 2 && role == 1) {
            look_back();
            const uint32_t m = meta[sd];
            finish_look_back(m & 0xffffu, m >> 16);
        }
        __syncthreads();  // (D) staged tile + run offsets visible; nobody reads the tables any more

        // ---- re-zero the tables; coalesced per-digit runs out to global memory -------------------
        for (int i = tid; i < GW * RADIX; i += THREADS) tab[i] = 0u;
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const uint32_t p = i * THREADS + tid;
            if (FULL || p < valid) {
                const KeyT     k = stg[swz(p)];
                const uint32_t d = digit_of<FK>(k, ds);
                uint64_t       o = static_cast<uint32_t>(koff[d] + p);
                if (wide) o = ((static_cast<uint64_t>(khi[d]) << 32) | koff[d]) + p;
                a.keys_out[o] = k;
                if constexpr (HAS_VALS) a.vals_out[o] = vstage[swz(p)];
            }
        }
        __syncthreads();  // (E) stage free again; sh_tile[s^1] published
    };

    while (true) {
        const uint32_t t = sh_tile[s];
        if (t >= a.num_tiles) break;
        if (tid == 0) {
            const uint32_t tn = atomicAdd(a.counter, 1u);
            sh_tile[s ^ 1]    = tn;
            if (tn < a.num_tiles) issue_prefetch(tn, s ^ 1);
        }
        if (t < full_tiles)
            process_tile(std::true_type{}, t);
        else
            process_tile(std::false_type{}, t);
        s ^= 1;
    }
}

}  // namespace rsb
