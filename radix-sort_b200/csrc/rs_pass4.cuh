// rs_pass4.cuh -- the onesweep pass kernel of the default (atomic ranking) mode, fourth generation.
//
// Same algorithm, global data structures, tile shapes and TMA tile prefetch as rs_onesweep_kernel in rs_kernels.cuh
// (which stays as the lane-bitmap fallback and as the A/B baseline, RSB200_PASS=1).  Replaces the reference's
// `reorder` kernel (src/kernels/RadixSort.cl:74-119) plus the per-tile part of scanhistograms / pastehistograms
// (:125-197).
//
// Why.  ncu of the first generation (profiles/r2j_*): the LSU data pipe is 84 % busy, and 36 % of the stall samples
// are warps waiting for THEIR OWN previous shared-memory access -- the SASS shows why: in the scatter loop every row is
// ATOMS (position) -> wait -> STS, in the output loop LDS (key) -> wait -> LDS.64 (run pointer) -> wait -> ST (generic),
// 22 rows one after the other with no overlap inside a warp, because the position atomic is an asm with a memory
// clobber and the store through a pointer that was loaded from shared memory may alias anything.  A version that
// removed 14 % of the wavefronts but not this serialisation (rs_pass3.cuh) was SLOWER.  So this version keeps the
// accesses and un-serialises them:
//   * count: the 22 row-uniformity tests (MATCH.ALL) are issued first, then 22 count atomics back to back;
//   * scatter: position atomics in batches of RS_P4_SB rows, then their stage stores (tiles without an all-equal row
//     -- every tile of a random input -- take a branch-free loop);
//   * output: run starts are element indices (32-bit when the pass has at most 2^32 keys) added to the kernel's own
//     output pointer, so the stores are STG and the compiler batches the loads of several rows ahead of them;
//   * look-back: the chunk base of a digit is cached in shared memory (one L2 round trip less on the critical path
//     of every tile but the first of a chunk).
// Stability rests on the same hardware property as before: lanes of ONE warp instruction that hit the same shared
// word are served in ascending lane order (self-tested per context, see rs_api.cu).
// Algorithmic traffic: sizeof(KeyT) read + sizeof(KeyT) written per key (+4 +4 with the index payload).
//
// RESULT (profiles/README.md, round 2): 0.741-0.744 ms per pass against 0.678 of the first generation, whatever the batch
// size -- kept as the record of that experiment, compiled with -DRS_EXPERIMENTAL_PASSES only.  It was measured with
// 512 x 22-key tiles (build with -DRS_U32A_KPT=22 to repeat it): its 2 KB larger shared-memory layout leaves room for
// only one CTA per SM with the 512 x 23 tiles the default kernel uses since.
#pragma once
#include "rs_kernels.cuh"
#include "rs_pass3.cuh"   // forget_derived

namespace rsb {

#ifndef RS_P4_SB
#define RS_P4_SB 4
#endif
constexpr int P4_SB = RS_P4_SB;   // rows per batch of position atomics

// 0: rows of 32 equal digits are counted by one lane (ATOMS.ADD 32); 1: every lane counts itself -- ATOMS.POPC.INC
// merges the lanes of one instruction that hit the same word, so the shortcut is not needed for the count
#ifndef RS_P4_PLAIN_COUNT
#define RS_P4_PLAIN_COUNT 1
#endif

// atom.shared.add under a predicate (no branch around the asm): returns the old value, or `otherwise` when off
__device__ __forceinline__ uint32_t atoms_add_return_if(bool on, uint32_t* p, uint32_t v, uint32_t otherwise) {
    uint32_t old = otherwise;
    asm volatile(
        "{\n"
        ".reg .pred q;\n"
        "setp.ne.u32 q, %3, 0;\n"
        "@q atom.shared.add.u32 %0, [%1], %2;\n"
        "}\n"
        : "+r"(old)
        : "r"(smem_u32(p)), "r"(v), "r"(static_cast<uint32_t>(on))
        : "memory");
    return old;
}

template <typename KeyT, bool HAS_VALS, int THREADS, int KPT>
struct Pass4Smem {
    static constexpr int    TILE   = THREADS * KPT;
    static constexpr int    NWARPS = THREADS / 32;
    static constexpr int    PARTS  = THREADS / RADIX;                 // threads per digit in the scan phase
    static constexpr size_t stage_bytes = sizeof(KeyT) * TILE;
    static constexpr size_t off_stage0  = 0;
    static constexpr size_t off_stage1  = off_stage0 + stage_bytes;
    static constexpr size_t off_vstage  = off_stage1 + stage_bytes;
    static constexpr size_t off_tab     = off_vstage + (HAS_VALS ? sizeof(uint32_t) * TILE : 0);   // [warp][digit]
    static constexpr size_t off_run     = off_tab + sizeof(uint32_t) * NWARPS * RADIX;             // [digit] run start - local start
    static constexpr size_t off_cbase   = off_run + sizeof(unsigned long long) * RADIX;            // [digit] cached chunk base
    static constexpr size_t off_ptot    = off_cbase + sizeof(unsigned long long) * RADIX;
    static constexpr size_t off_wsum    = off_ptot + sizeof(uint32_t) * PARTS * RADIX;
    static constexpr size_t off_mbar    = off_wsum + sizeof(uint32_t) * 16;
    static constexpr size_t off_tile    = off_mbar + sizeof(uint64_t) * 2;
    static constexpr size_t total       = off_tile + sizeof(uint32_t) * 4;
};

template <typename KeyT, bool HAS_VALS, int THREADS, int KPT, int MIN_BLOCKS, bool FK>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
rs_pass4_kernel(const OnesweepArgs<KeyT> a) {
    using L = Pass4Smem<KeyT, HAS_VALS, THREADS, KPT>;
    constexpr int TILE   = L::TILE;
    constexpr int NWARPS = L::NWARPS;
    constexpr int PARTS  = L::PARTS;
    constexpr int GPP    = NWARPS / PARTS;              // warp tables per scan thread
    static_assert(THREADS % RADIX == 0 && NWARPS % PARTS == 0, "THREADS must be a multiple of 256");
    static_assert((sizeof(KeyT) * TILE) % 128 == 0, "stage must keep 128B alignment");
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
    KeyT* const        stage0  = reinterpret_cast<KeyT*>(smem_raw + L::off_stage0);  // stage s = stage0 + s*TILE
    uint32_t* const    vstage  = reinterpret_cast<uint32_t*>(smem_raw + L::off_vstage);
    uint32_t* const    tab     = reinterpret_cast<uint32_t*>(smem_raw + L::off_tab);   // [warp][digit]
    uint32_t* const    run32   = reinterpret_cast<uint32_t*>(smem_raw + L::off_run);
    unsigned long long* const run64 = reinterpret_cast<unsigned long long*>(smem_raw + L::off_run);
    // [digit] (chunk + 1) << 48 | global offset of the chunk's first key with this digit; read and written by the one
    // thread that owns the digit in the look-back
    unsigned long long* const cbase = reinterpret_cast<unsigned long long*>(smem_raw + L::off_cbase);
    uint32_t* const    ptot    = reinterpret_cast<uint32_t*>(smem_raw + L::off_ptot);  // [part][digit] partial digit totals
    uint32_t* const    wsum    = reinterpret_cast<uint32_t*>(smem_raw + L::off_wsum);
    uint64_t* const    mbar    = reinterpret_cast<uint64_t*>(smem_raw + L::off_mbar);
    volatile uint32_t* sh_tile = reinterpret_cast<volatile uint32_t*>(smem_raw + L::off_tile);

    const int       tid   = threadIdx.x;
    const int       lane  = tid & 31;
    const int       wrp   = tid >> 5;
    const DigitSpec ds    = a.digit;
    const uint32_t  full_tiles = static_cast<uint32_t>(a.n / TILE);  // tiles [0, full_tiles) are complete
    const bool      big   = a.n > 0xffffffffull;                      // 64-bit run starts
    uint32_t* const gt    = tab + wrp * RADIX;
    // scan-phase role: digit `sd`, warp tables [sp*GPP, sp*GPP + GPP)
    const int       sd    = tid & (RADIX - 1);
    const int       sp    = tid / RADIX;

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
#pragma unroll
    for (int j = 0; j < GPP; ++j) tab[(sp * GPP + j) * RADIX + sd] = 0u;
    if (sp == 0) cbase[sd] = 0ull;   // tag 0 = nothing cached
    __syncthreads();

    uint32_t phase_bits = 0;  // bit s = parity to wait for on mbar[s]
    int      s          = 0;

    auto process_tile = [&](auto full_c, const uint32_t t) {
        constexpr bool FULL     = decltype(full_c)::value;
        const bool     via_tma  = FULL && a.use_tma;
        const uint64_t tile_base = static_cast<uint64_t>(t) * TILE;
        const uint32_t valid    = FULL ? TILE : static_cast<uint32_t>(a.n - tile_base);
        const uint32_t my_first = wrp * (32 * KPT) + lane;  // warp-striped: item r sits at my_first + 32*r
        KeyT* const    stg      = stage0 + s * TILE;

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
        // digit of row r (padding of a ragged tile ranks last)
        auto digit_row = [&](int r) -> uint32_t {
            uint32_t d = digit_of<FK>(key[r], ds);
            if (!FULL && (my_first + 32 * r) >= valid) d = RADIX - 1;
            return d;
        };

        // ---- count (so the tile's digit counts can be published early); rank while scattering -----------------------
        // Rows whose 32 digits are all equal would serialise the position atomic on one word: they are found first
        // (all the MATCH.ALL of the tile in flight together) and take a one-lane shortcut.
        uint32_t uniform_rows = 0;                       // bit r = all 32 digits of row r are equal
#pragma unroll
        for (int r = 0; r < KPT; ++r) {
            int all_same;
            __match_all_sync(0xffffffffu, digit_row(r), &all_same);
            uniform_rows |= all_same ? (1u << r) : 0u;
        }
        // one straight-line sequence of predicated atomics (no second code path: a duplicated loop for tiles with
        // uniform rows cost 50 spilled registers)
#pragma unroll
        for (int r = 0; r < KPT; ++r) {
            const uint32_t d   = digit_row(r);
            if constexpr (RS_P4_PLAIN_COUNT) {
                atomicAdd(&gt[d], 1u);
            } else {
                const bool uni = (uniform_rows >> r) & 1u;
                if (!uni) atomicAdd(&gt[d], 1u);
                if (uni && lane == 0) atomicAdd(&gt[d], 32u);
            }
        }
        __syncthreads();  // (A) all warp counts final; every key of this stage is in registers
#pragma unroll
        for (int r = 0; r < KPT; ++r) forget_derived(key[r]);   // recompute digits in the scatter phase, do not keep 22 of them

        // ---- scan: every thread owns digit sd of GPP warp tables ----------------------------------
        const uint32_t chunk          = t / a.tiles_per_chunk;
        const bool     first_in_chunk = (t % a.tiles_per_chunk) == 0;
        uint32_t       pre[(GPP + 1) / 2];  // exclusive prefixes over this thread's tables, two 16-bit values per register
        uint32_t       part_sum = 0;
#pragma unroll
        for (int j = 0; j < GPP; ++j) {
            if (j & 1) pre[j >> 1] |= part_sum << 16; else pre[j >> 1] = part_sum;
            part_sum += tab[(sp * GPP + j) * RADIX + sd];
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
        if (!FULL && sd == RADIX - 1) count -= (TILE - valid);
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
            tab[(sp * GPP + j) * RADIX + sd] =
                local_base + before_part + ((j & 1) ? (pre[j >> 1] >> 16) : (pre[j >> 1] & 0xffffu));
        __syncthreads();  // (C) tab[w][d] = tile-local position of warp w's first key with digit d

        // ---- scatter into shared memory in ranked order -----------------------------------------
        // gt[d] is the running tile position of (this warp, digit d).  Lanes that hit the same word are served in
        // ascending lane order (self-tested, see the header comment), so the returned value is the stable position.
#pragma unroll
        for (int r0 = 0; r0 < KPT; r0 += P4_SB) {
            uint32_t pos[P4_SB];
#pragma unroll
            for (int j = 0; j < P4_SB; ++j)
                if (r0 + j < KPT) {
                    const bool uni = (uniform_rows >> (r0 + j)) & 1u;
                    pos[j] = atoms_add_return_if(!uni || lane == 0, &gt[digit_row(r0 + j)], uni ? 32u : 1u, 0u);
                }
            constexpr uint32_t batch_mask = (P4_SB >= 32 ? 0xffffffffu : ((1u << P4_SB) - 1u));
            if (uniform_rows & (batch_mask << r0)) {   // rare: a row of 32 equal digits took one atomic for all its lanes
#pragma unroll
                for (int j = 0; j < P4_SB; ++j)
                    if (r0 + j < KPT && ((uniform_rows >> (r0 + j)) & 1u)) pos[j] = __shfl_sync(0xffffffffu, pos[j], 0) + lane;
            }
#pragma unroll
            for (int j = 0; j < P4_SB; ++j)
                if (r0 + j < KPT) {
                    RS_DCHECK(pos[j] < static_cast<uint32_t>(TILE));
                    stg[swz(pos[j])] = key[r0 + j];
                    if constexpr (HAS_VALS) vstage[swz(pos[j])] = val[r0 + j];
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
            // global offset of this chunk's first key with this digit: cached per digit, polled once per chunk
            unsigned long long* cb_ptr = a.base + static_cast<size_t>(chunk) * RADIX + sd;
            unsigned long long  cb     = cbase[sd];
            if ((cb >> 48) == chunk + 1) {
                cb &= (1ull << 48) - 1;
            } else {
                do { cb = ld_base(cb_ptr); } while (!(cb & BASE_READY));
                cb &= ~BASE_READY;
                cbase[sd] = cb | (static_cast<unsigned long long>(chunk + 1) << 48);
            }
            // the last tile of a chunk hands the next chunk its base (its inclusive prefix = the chunk's counts)
            if ((t % a.tiles_per_chunk) == a.tiles_per_chunk - 1 && chunk + 1 < a.num_chunks)
                st_base(cb_ptr + RADIX, BASE_READY | (cb + excl + count));
            // run start minus the run's tile-local start: the element at tile position p goes to out[run[d] + p]
            const unsigned long long run0 = cb + excl - local_base;   // modulo 2^64 (2^32): the sum with p is in range
            if (big) run64[sd] = run0; else run32[sd] = static_cast<uint32_t>(run0);
        }
        __syncthreads();  // (D) staged tile + run starts visible; nobody reads the tables any more

        // ---- re-zero the tables; coalesced per-digit runs out to global memory -------------------
#pragma unroll
        for (int j = 0; j < GPP; ++j) tab[(sp * GPP + j) * RADIX + sd] = 0u;
        if (!big) {
#pragma unroll
            for (int i = 0; i < KPT; ++i) {
                const uint32_t p = i * THREADS + tid;
                if (FULL || p < valid) {
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
                if (FULL || p < valid) {
                    const KeyT               k = stg[swz(p)];
                    const unsigned long long g = run64[digit_of<FK>(k, ds)] + p;
                    RS_DCHECK(g < a.n);
                    a.keys_out[g] = k;
                    if constexpr (HAS_VALS) a.vals_out[g] = vstage[swz(p)];
                }
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
