// rs_sweep.cuh -- the onesweep pass kernel of the default (atomic ranking) mode, second generation.
//
// Same algorithm and same global data structures as rs_onesweep_kernel in rs_kernels.cuh (which stays as the
// lane-bitmap fallback and as the A/B baseline, RSB200_SWEEP_V1=1): persistent CTAs, dynamic tile ids, TMA 1-D
// prefetch of coming tiles, count atomics -> table scan -> position atomics while scattering into the stage ->
// decoupled look-back -> coalesced per-digit runs.  Replaces the reference's `reorder` kernel
// (src/kernels/RadixSort.cl:74-119) plus the per-tile part of scanhistograms / pastehistograms (:125-197).
//
// What changed, and why.  Measured on the first generation (profiles/r1d_*, profiles/r2_*): the LSU data pipe is
// 84 % busy with shared-memory wavefronts, 20 % of the stall samples sit at barriers -- and every attempt to start
// the look-back EARLIER in a tile made the pass slower (0.70 -> 0.78 -> 0.85 -> 1.03 ms with the look-back after
// the scatter / before it / right after the count): a tile cannot finish before its predecessors have published
// their digit counts, tile ids are handed out a tile ahead of their processing, so any CTA that runs late stalls
// every successor on the whole GPU (a convoy).  What helps is SLACK between the publication of a tile's counts and
// their first use by anyone else.  So the tile loop is software-pipelined:
//
//     look-back(t) | scatter(t) | load + count(t+1) | (A) | scan(t+1) by threads [0,256), publishes counts(t+1) |
//                                                             output(t) by everybody (the scan warps join late)   | (E)
//
//   * the counts of tile t+1 are published most of a tile period before any successor's look-back needs them;
//   * the look-back is warp-cooperative: LANE j reads predecessor j, 32 predecessors per L2 round trip instead of 8
//     (at ~30 tiles per microsecond the nearest inclusive prefix is ~35 tiles back: one round trip instead of 4-5 on
//     the critical path of every tile);
//   * the table scan runs while the other warps already write the tile out; two CTA-wide barriers are left per tile
//     instead of six;
//   * the keys of tile t+1 wait in registers while tile t is written out (the output loop needs few registers);
//   * two count/position tables (tile parity), 16-bit counters, two warps per 32-bit word ({warp 2j : low,
//     warp 2j+1 : high}): half the words to scan and re-zero, 2 x 8 KB; the price is that the count atomic of the
//     odd warps adds 65536 and cannot be the lane-merging ATOMS.POPC.INC (+0.4 wavefronts per 32 keys);
//
// Stability rests on the same hardware property as before: lanes of ONE warp instruction that hit the same shared
// word are served in ascending lane order (self-tested per context, see rs_api.cu); two warps that share a word
// update disjoint 16-bit halves, and a half can never carry into its neighbour (every value is a tile position
// <= TILE <= 65535).
// Algorithmic traffic: sizeof(KeyT) read + sizeof(KeyT) written per key (+4 +4 with the index payload).
#pragma once
#include "rs_kernels.cuh"

namespace rsb {

// How the output loop finds a digit's run: 2 = 64-bit run pointers (keys and values), 0 = 32-bit element offsets plus a
// high word for passes over more than 2^32 keys.
#ifndef RS_SWEEP_RUNPTR
#define RS_SWEEP_RUNPTR 2
#endif
constexpr int SWEEP_RUNPTR = RS_SWEEP_RUNPTR;

template <typename KeyT, bool HAS_VALS, int THREADS, int KPT>
struct SweepSmem {
    static constexpr int    TILE   = THREADS * KPT;
    static constexpr int    NWARPS = THREADS / 32;
    static constexpr int    GW     = NWARPS / 2;             // packed table words per digit
    static constexpr bool   SPLIT  = THREADS >= 2 * RADIX;   // scan and look-back run on different threads
    static constexpr size_t stage_bytes = sizeof(KeyT) * TILE;
    static constexpr size_t off_stage0  = 0;
    static constexpr size_t off_stage1  = off_stage0 + stage_bytes;
    static constexpr size_t off_vstage  = off_stage1 + stage_bytes;
    static constexpr size_t off_tab     = off_vstage + (HAS_VALS ? sizeof(uint32_t) * TILE : 0);   // [2][GW][256]
    static constexpr size_t off_koff    = off_tab + sizeof(uint32_t) * 2 * GW * RADIX;
    static constexpr size_t off_khi     = off_koff + sizeof(uint32_t) * RADIX;
    static constexpr size_t off_vptr    = off_khi + sizeof(uint32_t) * RADIX;                       // value run pointers (RUNPTR 2)
    static constexpr size_t off_meta    = off_vptr + ((SWEEP_RUNPTR == 2 && HAS_VALS) ? sizeof(void*) * RADIX : 0);   // [2][256]
    static constexpr size_t off_wsum    = off_meta + sizeof(uint32_t) * 2 * RADIX;
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
    uint32_t* const    tab     = reinterpret_cast<uint32_t*>(smem_raw + L::off_tab);   // [slot][warp pair][digit], 2 x 16 bit
    uint32_t* const    koff    = reinterpret_cast<uint32_t*>(smem_raw + L::off_koff);  // [digit] run start - tile pos (low)
    uint32_t* const    khi     = reinterpret_cast<uint32_t*>(smem_raw + L::off_khi);   // high half, only if a.n > 2^32
    KeyT** const       kptr    = reinterpret_cast<KeyT**>(smem_raw + L::off_koff);     // RUNPTR 2: [digit] run pointer - tile pos
    uint32_t** const   vptr    = reinterpret_cast<uint32_t**>(smem_raw + L::off_vptr);
    uint32_t* const    meta    = reinterpret_cast<uint32_t*>(smem_raw + L::off_meta);  // [slot][digit] count | local_base << 16
    uint32_t* const    wsum    = reinterpret_cast<uint32_t*>(smem_raw + L::off_wsum);
    uint64_t* const    mbar    = reinterpret_cast<uint64_t*>(smem_raw + L::off_mbar);
    volatile uint32_t* sh_tile = reinterpret_cast<volatile uint32_t*>(smem_raw + L::off_tile);

    const int       tid    = threadIdx.x;
    const int       lane   = tid & 31;
    const int       warp   = tid >> 5;
    const uint32_t  hshift = (warp & 1) * 16;                 // this warp's half of its table words
    const uint32_t  one    = 1u << hshift;
    const uint32_t  hsel   = hshift ? 0x4432u : 0x4410u;      // PRMT selector that extracts the half
    const DigitSpec ds     = a.digit;
    const uint32_t  full_tiles = static_cast<uint32_t>(a.n / TILE);  // tiles [0, full_tiles) are complete
    const int       sd     = tid & (RADIX - 1);               // digit of this thread's scan / look-back role
    const int       role   = tid / RADIX;                     // 0: table scan, 1: look-back (SPLIT), else: none
    const bool      wide   = a.n > (1ull << 32);
    const uint32_t  my_first = warp * (32 * KPT) + lane;      // warp-striped: item r sits at my_first + 32*r

    auto issue_prefetch = [&](uint32_t t, int s) {
        // one thread only: bulk copy of a complete tile into stage s
        if (a.use_tma && t < full_tiles) {
            fence_proxy_async_smem();  // earlier generic-proxy accesses to this stage are ordered before the copy
            mbar_expect_tx(&mbar[s], static_cast<uint32_t>(L::stage_bytes));
            bulk_load_g2s(stage0 + s * TILE, a.keys_in + static_cast<uint64_t>(t) * TILE,
                          static_cast<uint32_t>(L::stage_bytes), &mbar[s]);
        }
    };

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        mbar_fence_init();
        // both stages are free: the first two tiles of this CTA are requested at once
        const uint32_t t0 = atomicAdd(a.counter, 1u);
        sh_tile[0]        = t0;
        if (t0 < a.num_tiles) issue_prefetch(t0, 0);
        const uint32_t t1 = t0 < a.num_tiles ? atomicAdd(a.counter, 1u) : t0;
        sh_tile[1]        = t1;
        if (t1 < a.num_tiles) issue_prefetch(t1, 1);
    }
    for (int i = tid; i < 2 * GW * RADIX; i += THREADS) tab[i] = 0u;  // every tile re-zeroes its table after its scatter
    __syncthreads();

    uint32_t phase_bits = 0;  // bit s = parity to wait for on mbar[s]

    // registers that carry a tile from its front half (load + count) to its scatter one iteration later
    KeyT     key[KPT];
    uint32_t val[HAS_VALS ? KPT : 1];
    uint32_t uniform_rows = 0;  // bit r = all 32 digits of row r are equal

    auto valid_of = [&](const uint32_t t) {
        const uint64_t rest = a.n - static_cast<uint64_t>(t) * TILE;
        return rest < static_cast<uint64_t>(TILE) ? static_cast<uint32_t>(rest) : static_cast<uint32_t>(TILE);
    };

    // ---- front half of a tile: keys (and values) into registers, one counting atomic per key ------------------
    auto front = [&](auto full_c, const uint32_t t, const int s) {
        constexpr bool FULL      = decltype(full_c)::value;
        const bool     via_tma   = t < full_tiles && a.use_tma;
        const uint64_t tile_base = static_cast<uint64_t>(t) * TILE;
        const uint32_t valid     = FULL ? TILE : valid_of(t);
        KeyT* const     stg      = stage0 + s * TILE;
        uint32_t* const gt       = tab + (s * GW + (warp >> 1)) * RADIX;
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
        if constexpr (HAS_VALS) {
#pragma unroll
            for (int r = 0; r < KPT; ++r) {
                const uint32_t i = my_first + 32 * r;
                val[r]           = (FULL || i < valid) ? a.vals_in[tile_base + i] : 0u;
            }
        }
        uniform_rows = 0;
#pragma unroll
        for (int r = 0; r < KPT; ++r) {
            uint32_t d = digit_of<FK>(key[r], ds);
            if (!FULL && (my_first + 32 * r) >= valid) d = RADIX - 1;  // padding ranks last
            int all_same;
            __match_all_sync(0xffffffffu, d, &all_same);
            if (all_same) {  // 32 equal digits would serialise on one word: one lane counts them all
                uniform_rows |= 1u << r;
                if (lane == 0) atomicAdd(&gt[d], 32u << hshift);
            } else {
                atomicAdd(&gt[d], one);
            }
        }
    };

    // ---- table scan of a tile (threads [0,256), digit sd): publish its digit counts, turn counts into positions ----
    auto scan = [&](auto full_c, const uint32_t t, const int s) {
        constexpr bool FULL  = decltype(full_c)::value;
        const uint32_t valid = FULL ? TILE : valid_of(t);
        uint32_t* const tb   = tab + s * GW * RADIX;
        uint32_t pre[GW];  // exclusive prefixes over the warps, packed like the table words
        uint32_t run = 0;
#pragma unroll
        for (int j = 0; j < GW; ++j) {
            const uint32_t w  = tb[j * RADIX + sd];
            const uint32_t lo = w & 0xffffu, hi = w >> 16;
            pre[j] = run | ((run + lo) << 16);
            run += lo + hi;
        }
        uint32_t count = run;
        if (!FULL && sd == RADIX - 1) count -= (TILE - valid);
        const bool first_in_chunk = (t % a.tiles_per_chunk) == 0;
        st_status(a.status + static_cast<size_t>(t) * RADIX + sd, (first_in_chunk ? STATUS_INCLUSIVE : STATUS_LOCAL) | count);
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
        const uint32_t local_base = woff + incl - count;  // tile-local start of digit sd
#pragma unroll
        for (int j = 0; j < GW; ++j) tb[j * RADIX + sd] = pre[j] + local_base * 0x10001u;
        meta[s * RADIX + sd] = count | (local_base << 16);
    };

    // ---- decoupled look-back of a tile (threads [256,512), or [0,256) when the CTA has only 256): digit sd ------
    auto look_back = [&](const uint32_t t, const int s) {
        const uint32_t chunk    = t / a.tiles_per_chunk;
        const uint32_t in_chunk = t % a.tiles_per_chunk;   // predecessors of this tile in its chunk
        uint32_t* const           my_status = a.status + static_cast<size_t>(t) * RADIX + sd;
        unsigned long long* const cb_ptr    = a.base + static_cast<size_t>(chunk) * RADIX + sd;
        uint32_t excl = 0;  // sum of the predecessors' counts in this chunk
        if (in_chunk != 0) {
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
                if (unready) while ((ld_status(p) >> 30) == 0) { __nanosleep(64); }
            }
        }
        // global offset of this chunk's first key with this digit (ready at once for chunk 0)
        unsigned long long cb;
        do { cb = ld_base(cb_ptr); } while (!(cb & BASE_READY));
        cb &= ~BASE_READY;
        const uint32_t m = meta[s * RADIX + sd];
        const uint32_t cnt = m & 0xffffu, lbase = m >> 16;
        // publish the inclusive prefix, hand the next chunk its base, set the run offset of digit sd
        RS_DCHECK(static_cast<unsigned long long>(excl) + cnt <= CHUNK_MAX_KEYS);
        if (in_chunk != 0) st_status(my_status, STATUS_INCLUSIVE | (excl + cnt));
        if (in_chunk == a.tiles_per_chunk - 1 && chunk + 1 < a.num_chunks)
            st_base(cb_ptr + RADIX, BASE_READY | (cb + excl + cnt));
        // run start minus the run's tile-local start: the element at tile position p goes to out[run0 + p]
        const unsigned long long run0 = cb + excl - lbase;  // may wrap below zero; run0 + p never does
        if constexpr (SWEEP_RUNPTR == 2) {
            kptr[sd] = a.keys_out + static_cast<long long>(run0);
            if constexpr (HAS_VALS) vptr[sd] = a.vals_out + static_cast<long long>(run0);
        } else {
            koff[sd] = static_cast<uint32_t>(run0);
            if (SWEEP_RUNPTR == 0 && wide) khi[sd] = static_cast<uint32_t>(run0 >> 32);
        }
    };

    // ---- scatter of the tile in registers into its stage, in ranked order -------------------------------------
    auto scatter = [&](auto full_c, const uint32_t t, const int s) {
        constexpr bool FULL  = decltype(full_c)::value;
        const uint32_t valid = FULL ? TILE : valid_of(t);
        KeyT* const     stg  = stage0 + s * TILE;
        uint32_t* const gt   = tab + (s * GW + (warp >> 1)) * RADIX;
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
                pos = __byte_perm(__shfl_sync(0xffffffffu, c, 0), 0u, hsel) + lane;
            } else {
                pos = __byte_perm(atoms_add_return(&gt[d], one), 0u, hsel);
            }
            RS_DCHECK(pos < static_cast<uint32_t>(TILE));
            stg[swz(pos)] = key[r];
            if constexpr (HAS_VALS) vstage[swz(pos)] = val[r];
        }
    };

    // ---- output of a ranked stage: coalesced per-digit runs; its table is re-zeroed on the way ------------------
    auto output = [&](auto full_c, const uint32_t t, const int s) {
        constexpr bool FULL    = decltype(full_c)::value;
        const uint32_t valid   = FULL ? TILE : valid_of(t);
        const KeyT* const stg  = stage0 + s * TILE;
        uint32_t* const   tb   = tab + s * GW * RADIX;
        for (int i = tid; i < GW * RADIX; i += THREADS) tb[i] = 0u;
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const uint32_t p = i * THREADS + tid;
            if (FULL || p < valid) {
                const KeyT     k = stg[swz(p)];
                const uint32_t d = digit_of<FK>(k, ds);
                if constexpr (SWEEP_RUNPTR == 2) {
                    RS_DCHECK(kptr[d] + p >= a.keys_out && kptr[d] + p < a.keys_out + a.n);
                    kptr[d][p] = k;
                    if constexpr (HAS_VALS) vptr[d][p] = vstage[swz(p)];
                } else {
                    uint64_t o = static_cast<uint32_t>(koff[d] + p);
                    if (SWEEP_RUNPTR == 0 && wide) o = ((static_cast<uint64_t>(khi[d]) << 32) | koff[d]) + p;
                    a.keys_out[o] = k;
                    if constexpr (HAS_VALS) a.vals_out[o] = vstage[swz(p)];
                }
            }
        }
    };

    // ---- prologue: front half and scan of this CTA's first tile ------------------------------------------------
    uint32_t t_cur = sh_tile[0];
    if (t_cur >= a.num_tiles) return;
    int s = 0;
    if (t_cur < full_tiles) front(std::true_type{}, t_cur, 0); else front(std::false_type{}, t_cur, 0);
    __syncthreads();
    if (role == 0) { if (t_cur < full_tiles) scan(std::true_type{}, t_cur, 0); else scan(std::false_type{}, t_cur, 0); }
    __syncthreads();

    // ---- steady state: keys(t_cur) in registers, table[s] holds its positions, its counts are published ---------
    auto iteration = [&](auto full_c, const uint32_t t_next) {
        const bool more = t_next < a.num_tiles;
        // the look-back warps walk the chain of t_cur first (L2 round trips) while the others already scatter
        if (SPLIT && role == 1) look_back(t_cur, s);
        scatter(full_c, t_cur, s);
        if (more) front(full_c, t_next, s ^ 1);   // the registers now carry tile t_next
        if (tid == THREADS - 1) {  // request the tile after next: (A)/(E) publish its id, its copy starts after (E)
            const uint32_t tn = more ? atomicAdd(a.counter, 1u) : t_next;
            sh_tile[s] = tn;
        }
        __syncthreads();  // (A) stage[s] ranked, run offsets of t_cur set; counts of t_next final
        if (role == 0) {   // the scan warps publish the counts of t_next and turn them into positions; the others write out
            if (more) scan(full_c, t_next, s ^ 1);
            if constexpr (!SPLIT) look_back(t_cur, s);
        }
        if constexpr (!SPLIT) __syncthreads();
        output(full_c, t_cur, s);
        __syncthreads();  // (E) stage[s] and table[s] free again; positions of t_next ready
        if (tid == THREADS - 1) {
            const uint32_t tn = sh_tile[s];
            if (tn < a.num_tiles) issue_prefetch(tn, s);
        }
    };

    while (true) {
        const uint32_t t_next = sh_tile[s ^ 1];
        const bool all_full = t_cur < full_tiles && (t_next < full_tiles || t_next >= a.num_tiles);
        if (all_full) iteration(std::true_type{}, t_next); else iteration(std::false_type{}, t_next);
        if (t_next >= a.num_tiles) break;
        t_cur = t_next;
        s ^= 1;
    }
}

}  // namespace rsb
