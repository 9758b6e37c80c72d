// rs_pass3.cuh -- the onesweep pass kernel of the default (atomic ranking) mode, third generation.
//
// Same algorithm, same global data structures and the same tile shapes as rs_onesweep_kernel in rs_kernels.cuh
// (which stays as the lane-bitmap fallback and as the A/B baseline, RSB200_PASS=1): persistent CTAs, dynamic tile
// ids, count atomics -> table scan -> position atomics while scattering into the stage -> decoupled look-back ->
// coalesced per-digit runs.  Replaces the reference's `reorder` kernel (src/kernels/RadixSort.cl:74-119) plus the
// per-tile part of scanhistograms / pastehistograms (:125-197).
//
// Why a third generation.  Rounds 1 and 2 measured (profiles/README.md) that the pass is bound by the NUMBER of
// wavefronts through the SM's load/store data pipe (84 % busy) -- about 19.3 per 32 keys: 15.1 from shared-memory
// instructions, 3.2 global, and 1.0 that no instruction shows: the TMA bulk copy of the next tile WRITES the stage
// through the same 32 banks.  Re-scheduling the same accesses (five variants, rs_sweep.cuh) changed nothing, so
// this version removes accesses:
//   * no TMA stage: the keys (and the index payload) of the NEXT tile are loaded straight into the key registers
//     with coalesced 32-lane loads right after the scatter, when those registers are dead, and arrive while the
//     current tile is written out (the output loop needs ten registers).  One wavefront per 32 keys (the load)
//     instead of two (bulk write + LDS); a ragged tile or an unaligned base needs no special path; shared memory
//     per CTA drops from 110 to 64 KB;
//   * the table scan reads AND clears a count with one ATOMS.EXCH and writes positions into a second table, so the
//     separate re-zero pass over the tables is gone (3 -> 2 accesses per table word);
//   * run starts are 32-bit element indices (one conflict-free LDS.32 per output row, shared by keys and payload)
//     whenever the pass has at most 2^32 keys; 64-bit indices otherwise (warp-uniform branch per tile).
// Stability rests on the same hardware property as before: lanes of ONE warp instruction that hit the same shared
// word are served in ascending lane order (self-tested per context, see rs_api.cu).
// Algorithmic traffic: sizeof(KeyT) read + sizeof(KeyT) written per key (+4 +4 with the index payload).
//
// RESULT (profiles/README.md, profiles/r2_pass_v1_vs_v3_ncu.txt, round 2): 14 % fewer shared-memory wavefronts, yet
// 0.729 ms per pass against 0.678 of the first generation (0.79 with the batched atomics that are in this file now):
// the loads cost 1.8 wavefronts per 32 keys in the global half of the same pipe, plus issue slots and MIO-queue stalls.
// Kept as the record of that experiment, compiled with -DRS_EXPERIMENTAL_PASSES only.
#pragma once
#include "rs_kernels.cuh"

namespace rsb {

#ifndef RS_P4_SB
#define RS_P4_SB 4
#endif
constexpr int P3_SB = RS_P4_SB;   // rows per batch of position atomics

template <typename KeyT, bool HAS_VALS, int THREADS, int KPT>
struct Pass3Smem {
    static constexpr int    TILE   = THREADS * KPT;
    static constexpr int    NWARPS = THREADS / 32;
    static constexpr int    PARTS  = THREADS / RADIX;                 // threads per digit in the scan phase
    static constexpr size_t off_stage  = 0;
    static constexpr size_t off_vstage = off_stage + sizeof(KeyT) * TILE;
    static constexpr size_t off_cnt    = off_vstage + (HAS_VALS ? sizeof(uint32_t) * TILE : 0);   // [warp][digit] counts
    static constexpr size_t off_pos    = off_cnt + sizeof(uint32_t) * NWARPS * RADIX;             // [warp][digit] positions
    static constexpr size_t off_run    = off_pos + sizeof(uint32_t) * NWARPS * RADIX;             // [digit] run start - local start
    static constexpr size_t off_ptot   = off_run + sizeof(unsigned long long) * RADIX;
    static constexpr size_t off_wsum   = off_ptot + sizeof(uint32_t) * PARTS * RADIX;
    static constexpr size_t off_tile   = off_wsum + sizeof(uint32_t) * 16;
    static constexpr size_t total      = off_tile + sizeof(uint32_t) * 4;
};

// Hides a key's value from common-subexpression elimination: the digit (and its table address) is recomputed in the
// scatter phase -- two ALU instructions -- instead of being kept in a register (or spilled) across the scan.
template <typename KeyT>
__device__ __forceinline__ void forget_derived(KeyT& k) {
    if constexpr (sizeof(KeyT) == 8) {
        unsigned long long w = static_cast<unsigned long long>(k);
        asm volatile("" : "+l"(w));
        k = static_cast<KeyT>(w);
    } else {
        uint32_t w = static_cast<uint32_t>(k);
        asm volatile("" : "+r"(w));
        k = static_cast<KeyT>(w);
    }
}

__device__ __forceinline__ uint32_t atoms_exch(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.exch.b32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}

template <typename KeyT, bool HAS_VALS, int THREADS, int KPT, int MIN_BLOCKS, bool FK>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
rs_pass3_kernel(const OnesweepArgs<KeyT> a) {
    using L = Pass3Smem<KeyT, HAS_VALS, THREADS, KPT>;
    constexpr int TILE   = L::TILE;
    constexpr int NWARPS = L::NWARPS;
    constexpr int PARTS  = L::PARTS;
    constexpr int GPP    = NWARPS / PARTS;              // warp tables per scan thread
    static_assert(THREADS % RADIX == 0 && NWARPS % PARTS == 0, "THREADS must be a multiple of 256");
    static_assert(TILE <= 0xffff, "16-bit prefix fields");
    static_assert(KPT <= 32, "uniform_rows is a 32-bit mask");

    {   // constant digit: the pass is the identity -- nothing to do, or a plain copy to keep the ping-pong parity
        const uint32_t action = a.plan->action[a.pass];
        if (action == PASS_NOTHING) return;
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
            return;
        }
    }

    extern __shared__ __align__(128) unsigned char smem_raw[];
    KeyT* const        stg     = reinterpret_cast<KeyT*>(smem_raw + L::off_stage);
    uint32_t* const    vstage  = reinterpret_cast<uint32_t*>(smem_raw + L::off_vstage);
    uint32_t* const    cnt     = reinterpret_cast<uint32_t*>(smem_raw + L::off_cnt);
    uint32_t* const    pos     = reinterpret_cast<uint32_t*>(smem_raw + L::off_pos);
    uint32_t* const    run32   = reinterpret_cast<uint32_t*>(smem_raw + L::off_run);
    unsigned long long* const run64 = reinterpret_cast<unsigned long long*>(smem_raw + L::off_run);
    uint32_t* const    ptot    = reinterpret_cast<uint32_t*>(smem_raw + L::off_ptot);
    uint32_t* const    wsum    = reinterpret_cast<uint32_t*>(smem_raw + L::off_wsum);
    volatile uint32_t* sh_tile = reinterpret_cast<volatile uint32_t*>(smem_raw + L::off_tile);

    const int       tid   = threadIdx.x;
    const int       lane  = tid & 31;
    const int       wrp   = tid >> 5;
    const DigitSpec ds    = a.digit;
    const uint32_t  full_tiles = static_cast<uint32_t>(a.n / TILE);  // tiles [0, full_tiles) are complete
    const bool      big   = a.n > 0xffffffffull;                      // 64-bit run starts
    uint32_t* const wc    = cnt + wrp * RADIX;
    uint32_t* const wp    = pos + wrp * RADIX;
    const uint32_t  my_first = wrp * (32 * KPT) + lane;  // warp-striped: item r sits at my_first + 32*r
    // scan-phase role: digit `sd`, warp tables [sp*GPP, sp*GPP + GPP)
    const int       sd    = tid & (RADIX - 1);
    const int       sp    = tid / RADIX;

    KeyT     key[KPT];
    uint32_t val[HAS_VALS ? KPT : 1];

    // keys (and payload) of tile t into the registers; slots past the end of the input hold the largest key
    auto load_tile = [&](const uint32_t t) {
        const uint64_t tile_base = static_cast<uint64_t>(t) * TILE;
        const KeyT* const src = a.keys_in + tile_base + my_first;
        if (t < full_tiles) {
#pragma unroll
            for (int r = 0; r < KPT; ++r) key[r] = __ldcs(src + 32 * r);
            if constexpr (HAS_VALS) {
                const uint32_t* const vsrc = a.vals_in + tile_base + my_first;
#pragma unroll
                for (int r = 0; r < KPT; ++r) val[r] = __ldcs(vsrc + 32 * r);
            }
        } else if (t < a.num_tiles) {
            const uint32_t valid = static_cast<uint32_t>(a.n - tile_base);
#pragma unroll
            for (int r = 0; r < KPT; ++r)
                key[r] = (my_first + 32 * r < valid) ? src[32 * r] : static_cast<KeyT>(~static_cast<KeyT>(0));
            if constexpr (HAS_VALS) {
                const uint32_t* const vsrc = a.vals_in + tile_base + my_first;
#pragma unroll
                for (int r = 0; r < KPT; ++r) val[r] = (my_first + 32 * r < valid) ? vsrc[32 * r] : 0u;
            }
        }
    };

    if (tid == 0) sh_tile[0] = atomicAdd(a.counter, 1u);
#pragma unroll
    for (int j = 0; j < GPP; ++j) cnt[(sp * GPP + j) * RADIX + sd] = 0u;
    __syncthreads();
    uint32_t t = sh_tile[0];
    int      s = 0;
    load_tile(t);

    // one tile; returns the id of the next one (whose keys are already on their way into the registers)
    auto process_tile = [&](auto full_c) -> uint32_t {
        constexpr bool full = decltype(full_c)::value;
        const uint64_t tile_base = static_cast<uint64_t>(t) * TILE;
        const uint32_t valid = full ? TILE : static_cast<uint32_t>(a.n - tile_base);

        // digit of row r (padding of a ragged tile ranks last)
        auto digit_row = [&](int r) -> uint32_t {
            uint32_t d = digit_of<FK>(key[r], ds);
            if (!full && (my_first + 32 * r) >= valid) d = RADIX - 1;
            return d;
        };
        // ---- count (so the tile's digit counts can be published early) -------------------------------------------
        // all row-uniformity tests first (MATCH.ALL in flight together), then the count atomics back to back
        uint32_t uniform_rows = 0;                       // bit r = all 32 digits of row r are equal
#pragma unroll
        for (int r = 0; r < KPT; ++r) {
            int all_same;
            __match_all_sync(0xffffffffu, digit_row(r), &all_same);
            uniform_rows |= all_same ? (1u << r) : 0u;
        }
        if (uniform_rows == 0) {
#pragma unroll
            for (int r = 0; r < KPT; ++r) atomicAdd(&wc[digit_row(r)], 1u);
        } else {
#pragma unroll
            for (int r = 0; r < KPT; ++r) {
                const uint32_t d = digit_row(r);
                if (uniform_rows & (1u << r)) {
                    if (lane == 0) atomicAdd(&wc[d], 32u);   // 32 equal digits would serialise on one word
                } else {
                    atomicAdd(&wc[d], 1u);
                }
            }
        }
        __syncthreads();  // (A) all warp counts final
#pragma unroll
        for (int r = 0; r < KPT; ++r) forget_derived(key[r]);   // recompute digits in the scatter phase, do not keep 22 of them
        const uint32_t tn = sh_tile[s ^ 1];

        // ---- scan: every thread owns digit sd of GPP warp tables; a count is read and cleared in one access ---------
        const uint32_t chunk          = t / a.tiles_per_chunk;
        const bool     first_in_chunk = (t % a.tiles_per_chunk) == 0;
        uint32_t       pre[(GPP + 1) / 2];  // exclusive prefixes over this thread's tables, two 16-bit values per register
        uint32_t       part_sum = 0;
#pragma unroll
        for (int j = 0; j < GPP; ++j) {
            if (j & 1) pre[j >> 1] |= part_sum << 16; else pre[j >> 1] = part_sum;
            part_sum += atoms_exch(&cnt[(sp * GPP + j) * RADIX + sd], 0u);
        }
        uint32_t count = part_sum, before_part = 0;
        if constexpr (PARTS > 1) {
            ptot[sp * RADIX + sd] = part_sum;
            __syncthreads();  // (B1)
            count = 0;
#pragma unroll
            for (int q = 0; q < PARTS; ++q) {
                const uint32_t v = ptot[q * RADIX + sd];
                count += v;
                before_part += (q < sp) ? v : 0u;
            }
        }
        if (!full && sd == RADIX - 1) count -= (TILE - valid);
        uint32_t* my_status = a.status + static_cast<size_t>(t) * RADIX + sd;
        if (sp == 0) st_status(my_status, (first_in_chunk ? STATUS_INCLUSIVE : STATUS_LOCAL) | count);
        uint32_t incl = count;  // inclusive scan over the 256 digits (each part computes it redundantly)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += x;
        }
        if (sp == 0 && lane == 31) wsum[sd >> 5] = incl;
        __syncthreads();  // (B2)
        uint32_t woff = 0;
#pragma unroll
        for (int w = 0; w < RADIX / 32; ++w) woff += (w < (sd >> 5)) ? wsum[w] : 0u;
        const uint32_t local_base = woff + incl - count;  // tile-local start of digit sd
#pragma unroll
        for (int j = 0; j < GPP; ++j)
            pos[(sp * GPP + j) * RADIX + sd] =
                local_base + before_part + ((j & 1) ? (pre[j >> 1] >> 16) : (pre[j >> 1] & 0xffffu));
        __syncthreads();  // (C) pos[w][d] = tile-local position of warp w's first key with digit d

        // ---- scatter into shared memory in ranked order -----------------------------------------
        // wp[d] is the running tile position of (this warp, digit d).  Lanes that hit the same word are served in
        // ascending lane order (self-tested), so the returned value is the stable position of this key.  Position
        // atomics go out in batches of P3_SB rows ahead of their stage stores (no warp waits on its own atomic).
        if (uniform_rows == 0) {
#pragma unroll
            for (int r0 = 0; r0 < KPT; r0 += P3_SB) {
                uint32_t p[P3_SB];
#pragma unroll
                for (int j = 0; j < P3_SB; ++j)
                    if (r0 + j < KPT) p[j] = atoms_add_return(&wp[digit_row(r0 + j)], 1u);
#pragma unroll
                for (int j = 0; j < P3_SB; ++j)
                    if (r0 + j < KPT) {
                        RS_DCHECK(p[j] < static_cast<uint32_t>(TILE));
                        stg[swz(p[j])] = key[r0 + j];
                        if constexpr (HAS_VALS) vstage[swz(p[j])] = val[r0 + j];
                    }
            }
        } else {
#pragma unroll
            for (int r = 0; r < KPT; ++r) {
                const uint32_t d = digit_row(r);
                uint32_t p;
                if (uniform_rows & (1u << r)) {
                    uint32_t c = 0;
                    if (lane == 0) c = atoms_add_return(&wp[d], 32u);
                    p = __shfl_sync(0xffffffffu, c, 0) + lane;
                } else {
                    p = atoms_add_return(&wp[d], 1u);
                }
                RS_DCHECK(p < static_cast<uint32_t>(TILE));
                stg[swz(p)] = key[r];
                if constexpr (HAS_VALS) vstage[swz(p)] = val[r];
            }
        }

        // ---- decoupled look-back: exclusive prefix of this tile's digit counts in its chunk ----
        if (sp == 0) {
            unsigned long long excl = 0;
            if (!first_in_chunk) {
                // walk back over the predecessors of this chunk, LB_WINDOW status words per round trip
                int32_t        left = static_cast<int32_t>(t % a.tiles_per_chunk);  // predecessors in this chunk
                const uint32_t* p   = my_status - RADIX;
                bool           done = false;
                while (!done) {
                    uint32_t v[LB_WINDOW];
#pragma unroll
                    for (int j = 0; j < LB_WINDOW; ++j)
                        v[j] = (j < left) ? ld_status(p - j * RADIX) : STATUS_INCLUSIVE;  // before the chunk: prefix 0
#pragma unroll
                    for (int j = 0; j < LB_WINDOW; ++j) {
                        if (done) break;
                        const uint32_t f = v[j] >> 30;
                        if (f == 0) break;  // not published yet: re-read from here
                        excl += v[j] & STATUS_VALUE_MASK;
                        p -= RADIX;
                        --left;
                        if (f == 2) done = true;
                    }
                }
                RS_DCHECK(excl + count <= CHUNK_MAX_KEYS);
                st_status(my_status, STATUS_INCLUSIVE | static_cast<uint32_t>(excl + count));
            }
            // global offset of this chunk's first key with this digit (ready at once for chunk 0)
            unsigned long long* cb_ptr = a.base + static_cast<size_t>(chunk) * RADIX + sd;
            unsigned long long  cb;
            do { cb = ld_base(cb_ptr); } while (!(cb & BASE_READY));
            cb &= ~BASE_READY;
            // the last tile of a chunk hands the next chunk its base (its inclusive prefix = the chunk's counts)
            if ((t % a.tiles_per_chunk) == a.tiles_per_chunk - 1 && chunk + 1 < a.num_chunks)
                st_base(cb_ptr + RADIX, BASE_READY | (cb + excl + count));
            // run start minus the run's tile-local start: the element at tile position p goes to out[run[d] + p]
            const unsigned long long run0 = cb + excl - local_base;   // modulo 2^64 (2^32): the sum with p is in range
            if (big) run64[sd] = run0; else run32[sd] = static_cast<uint32_t>(run0);
        }
        __syncthreads();  // (D) staged tile + run starts visible; nobody reads the tables any more

        // ---- next tile's keys into the (now dead) key registers; they arrive while this tile is written out ---------
        load_tile(tn);

        // ---- coalesced per-digit runs out to global memory -----------------------------------------
        if (!big) {
#pragma unroll
            for (int i = 0; i < KPT; ++i) {
                const uint32_t p = i * THREADS + tid;
                if (full || p < valid) {
                    const KeyT     k = stg[swz(p)];
                    const uint32_t g = run32[digit_of<FK>(k, ds)] + p;
                    RS_DCHECK(g < a.n);
                    a.keys_out[g] = k;
                    if constexpr (HAS_VALS) a.vals_out[g] = vstage[swz(p)];
                }
            }
        } else {
#pragma unroll
            for (int i = 0; i < KPT; ++i) {
                const uint32_t p = i * THREADS + tid;
                if (full || p < valid) {
                    const KeyT               k = stg[swz(p)];
                    const unsigned long long g = run64[digit_of<FK>(k, ds)] + p;
                    RS_DCHECK(g < a.n);
                    a.keys_out[g] = k;
                    if constexpr (HAS_VALS) a.vals_out[g] = vstage[swz(p)];
                }
            }
        }
        __syncthreads();  // (E) stage free again
        return tn;
    };

    while (t < a.num_tiles) {
        if (tid == 0) sh_tile[s ^ 1] = atomicAdd(a.counter, 1u);   // read after barrier (A)
        t = (t < full_tiles) ? process_tile(std::true_type{}) : process_tile(std::false_type{});
        s ^= 1;
    }
}

}  // namespace rsb
