// rs_merge.cuh -- pairwise merge of sorted key runs (multi-GPU final step).
//
// After the exchange every GPU owns P sorted runs (one per source GPU).  SURVEY.md 8(e).5 allows the final
// local step to be "one more local onesweep ... or a P-way merge"; merging is ~4x less shared-memory work than
// four onesweep passes for P = 2 and still cheaper for P = 8 (log2 P levels).  The first level reads its runs
// straight from PEER memory (CUDA IPC over NVLink): exchange and merge are one kernel.
//
//   rs_merge_partition_kernel       : one thread per output-tile boundary, merge-path binary search (global loads),
//                                     in two phases (coarse boundaries, then between neighbours)
//   rs_merge_tiles_pipelined_kernel : persistent CTAs, one tile of MERGE_TILE outputs at a time: line-aligned 16-byte
//                                     loads of the next tile in flight while the current one is merged out of shared
//                                     memory (per-thread merge-path split, MERGE_ITEMS-way serial merge, transpose,
//                                     coalesced store); jobs of a batch interleaved
//   rs_merge_tiles_kernel           : the first version (one CTA per tile, 4-byte loads, nothing in flight), kept for
//                                     A/B runs (RSB200_MERGE_UNPIPELINED=1)
// Keys only (the distributed configuration is keys-only); order is the key type's (sign bit flipped for signed).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "rs_merge_order.h"

namespace rsb {

// Tile shape of the merge kernels; overridable at compile time for A/B builds (tools/build_variants.sh, EXTRA=-D...).
#ifndef RS_MERGE_THREADS
#define RS_MERGE_THREADS 256
#endif
#ifndef RS_MERGE_ITEMS
#define RS_MERGE_ITEMS 16
#endif
#ifndef RS_MERGE_MINB32          // resident CTAs per SM the tile kernel is compiled for: 4-byte keys, local runs
#define RS_MERGE_MINB32 4
#endif
#ifndef RS_MERGE_MINB32_DENSE    // 4-byte keys, runs in peer memory
#define RS_MERGE_MINB32_DENSE 6
#endif
#ifndef RS_MERGE_MINB64          // 8-byte keys
#define RS_MERGE_MINB64 3
#endif
constexpr int MERGE_THREADS  = RS_MERGE_THREADS;
constexpr int MERGE_ITEMS    = RS_MERGE_ITEMS;
constexpr int MERGE_TILE     = MERGE_THREADS * MERGE_ITEMS;  // 4096 outputs per CTA
constexpr int MERGE_MAX_JOBS = 8;
static_assert(MERGE_THREADS % 32 == 0 && (MERGE_ITEMS == 8 || MERGE_ITEMS == 16 || MERGE_ITEMS == 32),
              "the padded transpose is conflict-free for 8, 16 or 32 items per thread");

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

// number of elements taken from A among the first `diag` outputs of merge(A, B); ties take A first.
// [lo_hint, hi_hint] brackets the answer (0 .. na when nothing is known).
template <typename KeyT, typename IndexT>
__device__ __forceinline__ IndexT merge_path(const KeyT* a, IndexT na, const KeyT* b, IndexT nb, IndexT diag, KeyT sign,
                                             IndexT lo_hint, IndexT hi_hint) {
    IndexT lo = diag > nb ? diag - nb : 0;
    IndexT hi = diag < na ? diag : na;
    if (lo < lo_hint) lo = lo_hint;
    if (hi > hi_hint) hi = hi_hint;
    while (lo < hi) {
        const IndexT mid = lo + ((hi - lo) >> 1);
        const KeyT   ka  = a[mid] ^ sign;
        const KeyT   kb  = b[diag - 1 - mid] ^ sign;
        if (ka <= kb) lo = mid + 1; else hi = mid;
    }
    return lo;
}
template <typename KeyT, typename IndexT>
__device__ __forceinline__ IndexT merge_path(const KeyT* a, IndexT na, const KeyT* b, IndexT nb, IndexT diag, KeyT sign) {
    return merge_path<KeyT, IndexT>(a, na, b, nb, diag, sign, 0, na);
}

// splits[job_first_boundary + i] = A-count of the first i*MERGE_TILE outputs of the job, i = 0 .. tiles_of_job.
// Every probe of a run in peer memory is a 32-byte NVLink read, so the search runs in two launches: `phase` 1 finds
// every MERGE_COARSE-th boundary (and the last one of each job) over the whole runs, `phase` 2 the others between
// their two coarse neighbours (the path is monotone) -- 18 instead of up to 32 steps for 2^31-key runs.  `phase` 0
// = everything in one launch over the whole runs.
constexpr unsigned long long MERGE_COARSE = 64;

template <typename KeyT>
__global__ void rs_merge_partition_kernel(const MergeBatch<KeyT> mb, unsigned long long* __restrict__ splits, const int phase) {
    const unsigned long long total_boundaries = mb.tile_start[mb.num_jobs] + mb.num_jobs;
    for (unsigned long long g = blockIdx.x * static_cast<unsigned long long>(blockDim.x) + threadIdx.x; g < total_boundaries;
         g += static_cast<unsigned long long>(gridDim.x) * blockDim.x) {
        // boundary g of the batch: job j owns boundaries [tile_start[j] + j, tile_start[j+1] + j + 1)
        int job = 0;
#pragma unroll
        for (int j = 1; j < MERGE_MAX_JOBS; ++j)
            if (j < mb.num_jobs && g >= mb.tile_start[j] + j) job = j;
        const KeyT* a = nullptr; const KeyT* b = nullptr; unsigned long long na = 0, nb = 0, first = 0, tiles = 0;
#pragma unroll
        for (int j = 0; j < MERGE_MAX_JOBS; ++j)
            if (j == job) { a = mb.a[j]; b = mb.b[j]; na = mb.na[j]; nb = mb.nb[j]; first = mb.tile_start[j] + j; tiles = mb.tile_start[j + 1] - mb.tile_start[j]; }
        const unsigned long long i = g - first;                       // boundary index within the job, 0 .. tiles
        const bool coarse = (i % MERGE_COARSE) == 0 || i == tiles;
        if (phase != 0 && coarse != (phase == 1)) continue;
        unsigned long long diag = i * MERGE_TILE;
        if (diag > na + nb) diag = na + nb;
        unsigned long long lo_hint = 0, hi_hint = na;
        if (phase == 2) {
            const unsigned long long prev = i - i % MERGE_COARSE;
            const unsigned long long next = prev + MERGE_COARSE < tiles ? prev + MERGE_COARSE : tiles;
            lo_hint = splits[first + prev];
            hi_hint = splits[first + next];
        }
        splits[g] = merge_path<KeyT, unsigned long long>(a, na, b, nb, diag, mb.sign, lo_hint, hi_hint);
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


// ---- persistent, software-pipelined variant (default) ------------------------------------------------------
// One CTA walks a strided sequence of tiles.  While it merges tile i out of shared memory, the loads of tile i+1
// are already in flight into registers and the two split words of tile i+2 are on their way, so a peer load's
// ~1-2 us NVLink round trip is hidden behind a whole merge phase per CTA.  The sequence interleaves the jobs of the
// batch (slot s -> job s % J, tile s / J): at level 1 of the tree every job reads two different source GPUs, so all
// P sources are read at the same rate all the time instead of every GPU pulling from GPU 0 and 1 first.
//
// Peer loads must be line-aligned per warp: a warp request that straddles 128-byte lines runs at ~430 GB/s over
// NVLink instead of ~790 (tools/peerbw.cu).  The pieces of a tile start at arbitrary elements, so the tile is loaded
// WHOLE LINES at a time with 16-byte loads into a "virtual" layout that is kept in shared memory as it is:
//     [ off_a slack | la elements of A | slack up to span_a ]  [ off_b slack | lb elements of B | slack up to span_b ]
// span_a / span_b are multiples of the line, so both pieces start line-aligned in global AND in shared memory; the
// merge simply reads A at sk + off_a and B at sk + span_a + off_b.  Only lines that hold at least one element of
// the piece are touched, so no load leaves the caller's allocation granule.
constexpr int MERGE_LINE_BYTES = 128;

template <typename KeyT>
struct MergeSlot {          // per thread
    const uint4*       pa;     // row-0 vector of this thread in A's lines
    const uint4*       pb;     // row-0 vector of this thread in B's lines (biased by -span_a)
    KeyT*              out;    // first output element of the tile
    uint32_t           la, lb;
    uint32_t           off_a;  // A starts at sk[off_a]
    uint32_t           span_a; // elements of shared memory taken by A's lines
    uint32_t           org_b;  // B starts at sk[org_b] (= span_a + off_b)
    uint32_t           total;  // span_a + span_b
};

// DENSE = more resident CTAs with fewer registers each (4-byte keys: 6 x 38 registers instead of 4 x 64).  Slower
// when both runs are local (2.56 vs 2.26 ms per 2^30 keys) but it holds 1.5x the bytes in flight, which is what bounds
// a level whose runs live in peer memory; the host picks it when an input pointer belongs to another device.
// REGIONS selects the walk order: false = slot s -> job s % J, tile s / J (the order every number in profiles/ was
// measured with; with the grid a multiple of J a CTA never leaves its job, so a batch whose tiles all sit in one job
// keeps only 1/J of the CTAs busy); true = the region-strided order of rs_merge_order.h, which balances any batch.
template <typename KeyT, bool DENSE, bool REGIONS>
__global__ void __launch_bounds__(MERGE_THREADS, sizeof(KeyT) == 4 ? (DENSE ? RS_MERGE_MINB32_DENSE : RS_MERGE_MINB32) : RS_MERGE_MINB64)
rs_merge_tiles_pipelined_kernel(const MergeBatch<KeyT> mb_param, const unsigned long long* __restrict__ splits,
                                const uint32_t max_tiles_per_job) {
    constexpr uint32_t LINE = MERGE_LINE_BYTES / sizeof(KeyT);          // elements per line
    constexpr uint32_t VEC  = 16 / sizeof(KeyT);                        // elements per 16-byte load
    constexpr uint32_t SK   = MERGE_TILE + 4 * LINE;                    // staged inputs incl. alignment slack
    constexpr uint32_t SO_STRIDE = MERGE_ITEMS + VEC;                   // outputs of one thread + one vector of padding
    constexpr int      ROWS = (MERGE_TILE + 3 * LINE + VEC * MERGE_THREADS - 1) / (VEC * MERGE_THREADS);
    extern __shared__ __align__(16) unsigned char merge_smem[];
    __shared__ MergeBatch<KeyT> mb;
    __shared__ MergeOrder       order;                      // rs_merge_order.h                        (REGIONS only)
    __shared__ uint32_t         first_tile[MERGE_MAX_JOBS]; // 32-bit copy of mb.tile_start, < 2^31    (REGIONS only)
    KeyT* const sk = reinterpret_cast<KeyT*>(merge_smem);   // [SK]                          virtual layout, see above
    KeyT* const so = sk + SK;                               // [MERGE_THREADS * SO_STRIDE]   outputs (layout by key size, below)
    const int tid = threadIdx.x;
    if (tid == 0) mb = mb_param;
    if constexpr (REGIONS) {
        if (tid == 0) order = merge_order_of(static_cast<uint32_t>(mb_param.tile_start[mb_param.num_jobs]));
        if (tid < MERGE_MAX_JOBS)   // jobs past num_jobs start at the total, i.e. never own a tile
            first_tile[tid] = static_cast<uint32_t>(mb_param.tile_start[tid < mb_param.num_jobs ? tid : mb_param.num_jobs]);
    }
    __syncthreads();
    const uint32_t J = static_cast<uint32_t>(mb.num_jobs);
    const uint32_t S = REGIONS ? order.slots : J * max_tiles_per_job;   // sequence slots (host guarantees < 2^32)
    const KeyT     sign = mb.sign;

    // slot -> (job, tile within the job)
    auto locate = [&](uint32_t s, uint32_t& job, uint32_t& lt) {
        if constexpr (REGIONS) {
            const uint32_t t = merge_tile_of_slot(order, s);
            job = 0;
#pragma unroll
            for (int j = 1; j < MERGE_MAX_JOBS; ++j)
                if (t >= first_tile[j]) job = j;     // the owner is the last job that starts at or before t
            lt = t - first_tile[job];
        } else {
            job = s % J;
            lt  = s / J;
        }
    };
    // does slot s name an existing tile?
    auto exists = [&](uint32_t s) {
        if constexpr (REGIONS) {
            return merge_tile_of_slot(order, s) < order.tiles;
        } else {
            const uint32_t job = s % J, lt = s / J;
            return lt < static_cast<uint32_t>(mb.tile_start[job + 1] - mb.tile_start[job]);
        }
    };
    // next slot >= s (stepping by the grid) that names an existing tile; S if none
    auto seek = [&](uint32_t s) {
        while (s < S) {
            if (exists(s)) break;
            if (S - s <= gridDim.x) { s = S; break; }
            s += gridDim.x;
        }
        return s;
    };
    auto advance = [&](uint32_t s) { return (S - s <= gridDim.x) ? S : seek(s + gridDim.x); };
    // the two split words of slot s
    auto load_splits = [&](uint32_t s, unsigned long long& s0, unsigned long long& s1) {
        uint32_t job, lt;
        locate(s, job, lt);
        const unsigned long long g = mb.tile_start[job] + lt + job;
        s0 = splits[g]; s1 = splits[g + 1];
    };
    auto make_slot = [&](uint32_t s, unsigned long long a0, unsigned long long a1) {
        uint32_t job, lt;
        locate(s, job, lt);
        const unsigned long long total = mb.na[job] + mb.nb[job];
        const unsigned long long d0 = static_cast<unsigned long long>(lt) * MERGE_TILE;
        const unsigned long long d1 = (d0 + MERGE_TILE < total) ? d0 + MERGE_TILE : total;
        MergeSlot<KeyT> t;
        const KeyT* const a = mb.a[job] + a0;
        const KeyT* const b = mb.b[job] + (d0 - a0);
        t.out = mb.out[job] + d0;
        t.la = static_cast<uint32_t>(a1 - a0); t.lb = static_cast<uint32_t>((d1 - a1) - (d0 - a0));
        t.off_a = t.la ? static_cast<uint32_t>(reinterpret_cast<uintptr_t>(a) / sizeof(KeyT)) & (LINE - 1) : 0;
        const uint32_t off_b = t.lb ? static_cast<uint32_t>(reinterpret_cast<uintptr_t>(b) / sizeof(KeyT)) & (LINE - 1) : 0;
        t.span_a = (t.off_a + t.la + LINE - 1) & ~(LINE - 1);
        t.org_b  = t.span_a + off_b;
        t.total  = t.span_a + ((off_b + t.lb + LINE - 1) & ~(LINE - 1));
        t.pa = reinterpret_cast<const uint4*>(a - t.off_a) + tid;
        t.pb = reinterpret_cast<const uint4*>(b - off_b - t.span_a) + tid;   // never dereferenced below span_a
        return t;
    };
    uint4 nx[ROWS];
    auto issue_loads = [&](const MergeSlot<KeyT>& t) {
#pragma unroll
        for (int k = 0; k < ROWS; ++k) {
            const uint32_t v = (tid + k * MERGE_THREADS) * VEC;     // virtual element index of this vector
            const uint4*   p = (v < t.span_a ? t.pa : t.pb) + k * MERGE_THREADS;
            if (v < t.total) nx[k] = *p;
        }
    };

    uint32_t s_cur = seek(blockIdx.x);
    if (s_cur >= S) return;
    unsigned long long p0, p1;
    load_splits(s_cur, p0, p1);
    MergeSlot<KeyT> cur = make_slot(s_cur, p0, p1);
    issue_loads(cur);
    uint32_t s_nxt = advance(s_cur);
    if (s_nxt < S) load_splits(s_nxt, p0, p1);
    KeyT sign4[VEC];
#pragma unroll
    for (uint32_t i = 0; i < VEC; ++i) sign4[i] = sign;
    const uint4 signv = *reinterpret_cast<const uint4*>(sign4);

    while (true) {
        // stage tile `cur` (its loads were issued one iteration ago), keys mapped to unsigned order
        const uint32_t la = cur.la, lb = cur.lb, n_out = la + lb;
        const KeyT* const sa = sk + cur.off_a;
        const KeyT* const sb = sk + cur.org_b;
#pragma unroll
        for (int k = 0; k < ROWS; ++k) {
            const uint32_t c = tid + k * MERGE_THREADS;
            if (c * VEC < cur.total)
                reinterpret_cast<uint4*>(sk)[c] = make_uint4(nx[k].x ^ signv.x, nx[k].y ^ signv.y, nx[k].z ^ signv.z, nx[k].w ^ signv.w);
        }
        KeyT* const dst = cur.out;
        __syncthreads();
        // put tile i+1 in flight, ask for the splits of tile i+2
        const bool more = s_nxt < S;
        if (more) {
            cur = make_slot(s_nxt, p0, p1);
            issue_loads(cur);
            s_nxt = advance(s_nxt);
            if (s_nxt < S) load_splits(s_nxt, p0, p1);
        }
        // merge tile i out of shared memory; reads past the end of a piece hit the slack (or the other piece) and
        // are never selected -- the index tests in take_a decide
        uint32_t diag = tid * MERGE_ITEMS;
        if (diag > n_out) diag = n_out;
        uint32_t ai;
        {
            uint32_t lo = diag > lb ? diag - lb : 0, hi = diag < la ? diag : la;
            while (lo < hi) {
                const uint32_t mid = (lo + hi) >> 1;
                if (sa[mid] <= sb[diag - 1 - mid]) lo = mid + 1; else hi = mid;
            }
            ai = lo;
        }
        uint32_t bi = diag - ai;
        KeyT     ka = sa[ai], kb = sb[bi];
        if constexpr (sizeof(KeyT) == 4) {
            // 4-byte keys: scalar stores into the padded transpose as the merge goes (measured faster than vectors here)
            KeyT* const wo = so + tid * MERGE_ITEMS + ((tid * MERGE_ITEMS) >> 5);   // == merge_pad(tid * MERGE_ITEMS + k) - k
#pragma unroll
            for (int k = 0; k < MERGE_ITEMS; ++k) {
                const bool take_a = (bi >= lb) || (ai < la && ka <= kb);
                wo[k] = take_a ? ka : kb;
                if (take_a) ka = sa[++ai]; else kb = sb[++bi];
            }
            __syncthreads();  // outputs complete, and nobody reads sk any more
            const KeyT* const ro = so + tid + (tid >> 5);              // == merge_pad(tid + k * MERGE_THREADS) - k * 264
            if (n_out == MERGE_TILE) {
#pragma unroll
                for (int k = 0; k < MERGE_ITEMS; ++k) dst[tid + k * MERGE_THREADS] = ro[k * (MERGE_THREADS + MERGE_THREADS / 32)] ^ sign;
            } else {
#pragma unroll
                for (int k = 0; k < MERGE_ITEMS; ++k)
                    if (tid + k * MERGE_THREADS < n_out) dst[tid + k * MERGE_THREADS] = ro[k * (MERGE_THREADS + MERGE_THREADS / 32)] ^ sign;
            }
        } else {
            // 8-byte keys: thread t's MERGE_ITEMS outputs go to so[t * SO_STRIDE ...] with 16-byte stores; the one vector
            // of padding per thread keeps these stores and the coalesced 16-byte reads below (almost) conflict-free
            KeyT r[MERGE_ITEMS];
#pragma unroll
            for (int k = 0; k < MERGE_ITEMS; ++k) {
                const bool take_a = (bi >= lb) || (ai < la && ka <= kb);
                r[k] = (take_a ? ka : kb) ^ sign;
                if (take_a) ka = sa[++ai]; else kb = sb[++bi];
            }
            uint4* const wo = reinterpret_cast<uint4*>(so + tid * SO_STRIDE);
#pragma unroll
            for (int q = 0; q < MERGE_ITEMS / static_cast<int>(VEC); ++q) {
                KeyT tmp[VEC];
#pragma unroll
                for (uint32_t e = 0; e < VEC; ++e) tmp[e] = r[q * VEC + e];
                wo[q] = *reinterpret_cast<const uint4*>(tmp);
            }
            __syncthreads();  // outputs complete, and nobody reads sk any more
            if (n_out == MERGE_TILE && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
#pragma unroll
                for (int k = 0; k < MERGE_TILE / static_cast<int>(VEC) / MERGE_THREADS; ++k) {
                    const uint32_t i = (tid + k * MERGE_THREADS) * VEC;     // first output element of this vector
                    reinterpret_cast<uint4*>(dst)[tid + k * MERGE_THREADS] =
                        *reinterpret_cast<const uint4*>(so + (i / MERGE_ITEMS) * SO_STRIDE + (i % MERGE_ITEMS));
                }
            } else {   // ragged last tile of a job, or an output run that does not start on a 16-byte boundary
#pragma unroll
                for (int k = 0; k < MERGE_ITEMS; ++k) {
                    const uint32_t i = tid + k * MERGE_THREADS;
                    if (i < n_out) dst[i] = so[(i / MERGE_ITEMS) * SO_STRIDE + (i % MERGE_ITEMS)];
                }
            }
        }
        if (!more) break;
    }
}

}  // namespace rsb
