// rs_api.cu -- the C ABI of include/rsb200.h over the CUDA runtime.
//
// Host side of the hot path: what RadixSortGPU<T>::{initialize, uploadData, calculate, downloadData,
// release} (reference src/RadixSortGPU.cpp:451-543, 299-357, 444-449) and ComputeDeviceData<T>
// (src/ComputeDeviceData.cpp:10-79) do over OpenCL is done here over cudaMalloc'd ping-pong buffers
// and stream-ordered launches.  No host synchronisation inside a sort: one memset, the histogram
// kernel, the scan kernel and one onesweep kernel per 8-bit digit are enqueued back to back.
#include "../../include/rsb200.h"
#include "rs_kernels.cuh"
#include "rs_sweep.cuh"
#ifdef RS_EXPERIMENTAL_PASSES   // round-2 experiments kept for the record (profiles/README.md); not part of the default build
#include "rs_pass3.cuh"
#include "rs_pass4.cuh"
#endif
#include "rs_merge.cuh"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <mutex>
#include <new>
#include <string>
#include <vector>

using namespace rsb;

// ------------------------------------------------------------------------------------------------
// error plumbing
// ------------------------------------------------------------------------------------------------
static thread_local std::string g_last_error;

// cudaFuncSetAttribute is a per-device setting: launch helpers remember, per call site and thread, on which devices
// their kernel has been configured (a process may hold contexts on several GPUs).
static uint64_t current_device_bit() {
    int dev = 0;
    cudaGetDevice(&dev);
    return 1ull << (dev & 63);
}

static int fail(int status, const char* what, cudaError_t e = cudaSuccess) {
    char buf[512];
    if (e != cudaSuccess)
        snprintf(buf, sizeof buf, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
    else
        snprintf(buf, sizeof buf, "%s", what);
    g_last_error = buf;
    return status;
}
#define RS_CUDA(expr, status)                                   \
    do {                                                        \
        cudaError_t _e = (expr);                                \
        if (_e != cudaSuccess) return fail((status), #expr, _e); \
    } while (0)

// Key-type table (what TypeNameString<T> + the generated `#define DataType` preamble are in the reference,
// Common/CLTypeInformation.h:8-46, src/RadixSortGPU.cpp:431-442).
struct KeyInfo {
    size_t             bytes;
    bool               is_signed;  // two's-complement integers: the sign bit is flipped in the digit that holds it
    unsigned long long finf;       // floating point: bit pattern of +infinity (see float_order_bits), else 0
};
static bool key_info(int type, KeyInfo* out) {
    switch (type) {
        case RS_KEY_U32:  *out = {4, false, 0}; return true;
        case RS_KEY_I32:  *out = {4, true, 0}; return true;
        case RS_KEY_U64:  *out = {8, false, 0}; return true;
        case RS_KEY_I64:  *out = {8, true, 0}; return true;
        case RS_KEY_U16:  *out = {2, false, 0}; return true;
        case RS_KEY_I16:  *out = {2, true, 0}; return true;
        case RS_KEY_F32:  *out = {4, false, 0x7f800000ull}; return true;
        case RS_KEY_F64:  *out = {8, false, 0x7ff0000000000000ull}; return true;
        case RS_KEY_F16:  *out = {2, false, 0x7c00ull}; return true;
        case RS_KEY_BF16: *out = {2, false, 0x7f80ull}; return true;
    }
    return false;
}

// Every entry runs on its context's device and puts the calling thread's current device back afterwards (a torch
// caller may be working on another GPU).
struct DeviceGuard {
    int  prev = -1;
    bool switched = false;
    cudaError_t status = cudaSuccess;
    explicit DeviceGuard(int device) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != device) { status = cudaSetDevice(device); switched = status == cudaSuccess; }
    }
    ~DeviceGuard() { if (switched && prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};

struct rs_context_s {
    int device;
    int sm_count;
    int max_smem_optin;
    int atomic_rank_ok;  // the device passed rs_atomic_order_selftest_kernel (same-word shared atomics in lane order)
    unsigned long long* d_result;  // 8 words for the blocking *_device checks (no cudaMalloc per call)
    std::mutex          result_mutex;
};
struct rs_stream_s {
    rs_context   ctx;
    cudaStream_t st;
    bool         owned;
};

struct EventQuad {
    cudaEvent_t e[4];  // start, after histogram, after scan, after last pass
};

struct rs_sorter_s {
    rs_context  ctx;
    rs_key_type type;
    uint32_t    flags;
    uint64_t    max_keys;
    size_t      key_bytes;
    bool        is_signed;
    unsigned long long finf;  // floating-point keys: +infinity bit pattern (0 for integers)
    bool        has_vals;
    bool        atomic_rank;  // ranking mode of the onesweep kernel (see rs_kernels.cuh)
    // device memory
    void*     keys[2];
    uint32_t* vals[2];
    unsigned char* scratch;
    size_t    scratch_bytes;
    // result of the last rs_sorter_sort
    int       result_index;
    uint64_t  last_n;
    int       last_passes;
    int       last_chunks;
    size_t    last_hist_offset;
    size_t    last_plan_offset;
    // profiling
    std::vector<EventQuad> free_events;
    std::vector<EventQuad> pending;
    rs_runtimes            rt;
    uint64_t               launches;
};

// Statistics::update of the reference (src/Statistics.h:21-31), including the else-if.
static void stat_update(rs_statistics& s, double v) {
    s.n++;
    s.sum += v;
    s.avg = s.sum / static_cast<double>(s.n);
    if (v > s.max)
        s.max = v;
    else if (v < s.min)
        s.min = v;
}
static void stat_reset(rs_statistics& s) {
    s.min = std::numeric_limits<double>::infinity();
    s.max = -std::numeric_limits<double>::infinity();
    s.avg = 0.0;
    s.sum = 0.0;
    s.n   = 0;
}

// ------------------------------------------------------------------------------------------------
// tile configurations (per key width / payload) and launch plumbing
// ------------------------------------------------------------------------------------------------
// THREADS x KPT keys per tile, MINB CTAs per SM.  Measured on B200 with tools/build_variants.sh.
template <typename KeyT, bool HAS_VALS, bool ATOMIC>
struct TileCfg;
#ifndef RS_U32_THREADS
#define RS_U32_THREADS 256
#define RS_U32_KPT 24
#define RS_U32_MINB 3
#endif
// atomic ranking mode (default on B200): larger tiles pay (fewer look-backs and table scans per key)
#ifndef RS_U32A_THREADS
#define RS_U32A_THREADS 512
#define RS_U32A_KPT 23
#define RS_U32A_MINB 2
#endif
#ifndef RS_U32VA_THREADS
#define RS_U32VA_THREADS 512
#define RS_U32VA_KPT 15
#define RS_U32VA_MINB 2
#endif
#ifndef RS_U64A_THREADS
#define RS_U64A_THREADS 512
#define RS_U64A_KPT 11
#define RS_U64A_MINB 2
#endif
#ifndef RS_U64VA_THREADS
#define RS_U64VA_THREADS 512
#define RS_U64VA_KPT 9
#define RS_U64VA_MINB 2
#endif
template <> struct TileCfg<uint32_t, false, false> { static constexpr int THREADS = RS_U32_THREADS, KPT = RS_U32_KPT, MINB = RS_U32_MINB; };
template <> struct TileCfg<uint32_t, true,  false> { static constexpr int THREADS = 512, KPT = 12, MINB = 2; };
template <> struct TileCfg<uint64_t, false, false> { static constexpr int THREADS = 512, KPT = 8,  MINB = 2; };
template <> struct TileCfg<uint64_t, true,  false> { static constexpr int THREADS = 512, KPT = 7,  MINB = 2; };  // 8 keys left room for ONE CTA per SM (found by the static_assert in launch_pass_mode)
template <> struct TileCfg<uint32_t, false, true>  { static constexpr int THREADS = RS_U32A_THREADS, KPT = RS_U32A_KPT, MINB = RS_U32A_MINB; };
template <> struct TileCfg<uint32_t, true,  true>  { static constexpr int THREADS = RS_U32VA_THREADS, KPT = RS_U32VA_KPT, MINB = RS_U32VA_MINB; };
template <> struct TileCfg<uint64_t, false, true>  { static constexpr int THREADS = RS_U64A_THREADS, KPT = RS_U64A_KPT, MINB = RS_U64A_MINB; };
template <> struct TileCfg<uint64_t, true,  true>  { static constexpr int THREADS = RS_U64VA_THREADS, KPT = RS_U64VA_KPT, MINB = RS_U64VA_MINB; };

// 16-bit keys (two passes): same shapes as 32-bit keys -- a key still takes one register
template <bool V, bool A> struct TileCfg<uint16_t, V, A> : TileCfg<uint32_t, V, A> {};

template <typename KeyT, bool HAS_VALS, bool ATOMIC>
static constexpr int tile_of() { return TileCfg<KeyT, HAS_VALS, ATOMIC>::THREADS * TileCfg<KeyT, HAS_VALS, ATOMIC>::KPT; }

static int tile_keys(size_t key_bytes, bool has_vals, bool atomic) {
    if (key_bytes <= 4) {
        if (atomic) return has_vals ? tile_of<uint32_t, true, true>() : tile_of<uint32_t, false, true>();
        return has_vals ? tile_of<uint32_t, true, false>() : tile_of<uint32_t, false, false>();
    }
    if (atomic) return has_vals ? tile_of<uint64_t, true, true>() : tile_of<uint64_t, false, true>();
    return has_vals ? tile_of<uint64_t, true, false>() : tile_of<uint64_t, false, false>();
}

struct Plan {
    int      tile;
    uint64_t num_tiles;
    uint32_t tiles_per_chunk;
    uint64_t chunk_keys;
    int      num_chunks;
    int      passes;
    // scratch layout (bytes)
    size_t off_hist, off_ctr, off_plan, off_status, zero_bytes, off_base, total;
};

static Plan make_plan(uint64_t n, size_t key_bytes, bool has_vals, bool atomic, int passes) {
    Plan p{};
    p.tile            = tile_keys(key_bytes, has_vals, atomic);
    p.num_tiles       = (n + p.tile - 1) / p.tile;
    p.tiles_per_chunk = static_cast<uint32_t>(CHUNK_MAX_KEYS / p.tile);
    // test hook: shrink the look-back chunk so the multi-chunk path can be checked exactly at small n
    static const long debug_tpc = [] { const char* e = getenv("RSB200_DEBUG_TILES_PER_CHUNK"); return e ? atol(e) : 0L; }();
    if (debug_tpc > 0 && static_cast<uint64_t>(debug_tpc) < p.tiles_per_chunk) p.tiles_per_chunk = static_cast<uint32_t>(debug_tpc);
    p.chunk_keys      = static_cast<uint64_t>(p.tiles_per_chunk) * p.tile;
    p.num_chunks      = static_cast<int>((n + p.chunk_keys - 1) / p.chunk_keys);
    if (p.num_chunks < 1) p.num_chunks = 1;
    p.passes          = passes;
    auto align = [](size_t x) { return (x + 255) & ~static_cast<size_t>(255); };
    size_t o     = 0;
    p.off_hist   = o; o = align(o + sizeof(uint64_t) * passes * RADIX);
    p.off_ctr    = o; o = align(o + sizeof(uint32_t) * MAX_PASSES);
    p.off_plan   = o; o = align(o + sizeof(SortPlanDev));
    p.off_base   = o; o = align(o + sizeof(uint64_t) * passes * p.num_chunks * RADIX);
    p.off_status = o; o = align(o + sizeof(uint32_t) * RADIX * p.num_tiles * passes);
    p.zero_bytes = o;
    p.total      = o;
    return p;
}

template <typename KeyT, int NPASS, bool FK>
static cudaError_t launch_hist_n(int grid, cudaStream_t st, const KeyT* keys, uint64_t n, unsigned long long* hist,
                                 const PassTable& pt) {
    if constexpr (NPASS > static_cast<int>(sizeof(KeyT))) {
        return cudaErrorInvalidValue;  // more 8-bit digits than the key has
    } else {
        auto kern = rs_histogram_kernel<KeyT, NPASS, FK>;
        static thread_local uint64_t configured = 0;
        if (const uint64_t bit = current_device_bit(); !(configured & bit)) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(HistCfg<NPASS>::bytes));
            if (e != cudaSuccess) return e;
            configured |= bit;
        }
        // whole bytes from bit 0 up: the kernel's fast digit extraction applies (see rs_histogram_kernel)
        bool bytes = true;
        KeyT key_flip = 0;
        for (int p = 0; p < NPASS; ++p) {
            bytes = bytes && pt.d[p].shift == 8 * p && pt.d[p].mask == 0xffu;
            key_flip = static_cast<KeyT>(key_flip | (static_cast<KeyT>(pt.d[p].flip) << pt.d[p].shift));
        }
        kern<<<grid, HIST_THREADS, HistCfg<NPASS>::bytes, st>>>(keys, n, hist, pt, bytes ? 1 : 0, key_flip);
        return cudaGetLastError();
    }
}
template <typename KeyT, bool FK>
static cudaError_t launch_hist(int sm_count, cudaStream_t st, const KeyT* keys, uint64_t n, unsigned long long* hist,
                               const PassTable& pt) {
    const uint64_t vec_per_block = static_cast<uint64_t>(HIST_THREADS) * HIST_VECS * (16 / sizeof(KeyT));
    uint64_t want = (n + vec_per_block - 1) / vec_per_block;
    // one block per SM (the per-lane tables take most of the shared memory); 32-bit counters: < 2^31 keys per block
    const uint64_t min_grid = (n >> 31) + 1;
    int grid = static_cast<int>(std::max<uint64_t>(
        min_grid, std::min<uint64_t>(std::max<uint64_t>(want, 1), static_cast<uint64_t>(sm_count))));
    switch (pt.num_passes) {
        case 1: return launch_hist_n<KeyT, 1, FK>(grid, st, keys, n, hist, pt);
        case 2: return launch_hist_n<KeyT, 2, FK>(grid, st, keys, n, hist, pt);
        case 3: return launch_hist_n<KeyT, 3, FK>(grid, st, keys, n, hist, pt);
        case 4: return launch_hist_n<KeyT, 4, FK>(grid, st, keys, n, hist, pt);
        case 5: return launch_hist_n<KeyT, 5, FK>(grid, st, keys, n, hist, pt);
        case 6: return launch_hist_n<KeyT, 6, FK>(grid, st, keys, n, hist, pt);
        case 7: return launch_hist_n<KeyT, 7, FK>(grid, st, keys, n, hist, pt);
        case 8: return launch_hist_n<KeyT, 8, FK>(grid, st, keys, n, hist, pt);
    }
    return cudaErrorInvalidValue;
}

template <typename KeyT, int NPASS>
static cudaError_t launch_gather_hist_n(int grid, cudaStream_t st, const SegmentTable<KeyT>& seg, KeyT* gathered,
                                        unsigned long long* hist, const PassTable& pt) {
    if constexpr (NPASS > static_cast<int>(sizeof(KeyT))) {
        return cudaErrorInvalidValue;
    } else {
        auto kern = rs_gather_histogram_kernel<KeyT, NPASS>;
        static thread_local uint64_t configured = 0;
        if (const uint64_t bit = current_device_bit(); !(configured & bit)) {
            cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                 static_cast<int>(HistCfg<NPASS>::bytes));
            if (e != cudaSuccess) return e;
            configured |= bit;
        }
        kern<<<grid, HIST_THREADS, HistCfg<NPASS>::bytes, st>>>(seg, gathered, hist, pt);
        return cudaGetLastError();
    }
}
template <typename KeyT>
static cudaError_t launch_gather_hist(int sm_count, cudaStream_t st, const SegmentTable<KeyT>& seg, KeyT* gathered,
                                      uint64_t n, unsigned long long* hist, const PassTable& pt) {
    const uint64_t min_grid = (n >> 31) + 1;  // 32-bit shared counters: < 2^31 keys per block
    const int grid = static_cast<int>(std::max<uint64_t>(min_grid, static_cast<uint64_t>(sm_count)));
    switch (pt.num_passes) {
        case 1: return launch_gather_hist_n<KeyT, 1>(grid, st, seg, gathered, hist, pt);
        case 2: return launch_gather_hist_n<KeyT, 2>(grid, st, seg, gathered, hist, pt);
        case 3: return launch_gather_hist_n<KeyT, 3>(grid, st, seg, gathered, hist, pt);
        case 4: return launch_gather_hist_n<KeyT, 4>(grid, st, seg, gathered, hist, pt);
        case 5: return launch_gather_hist_n<KeyT, 5>(grid, st, seg, gathered, hist, pt);
        case 6: return launch_gather_hist_n<KeyT, 6>(grid, st, seg, gathered, hist, pt);
        case 7: return launch_gather_hist_n<KeyT, 7>(grid, st, seg, gathered, hist, pt);
        case 8: return launch_gather_hist_n<KeyT, 8>(grid, st, seg, gathered, hist, pt);
    }
    return cudaErrorInvalidValue;
}

// Per-kernel launch state: dynamic shared memory attribute and resident CTAs per SM, per device (a process may hold
// contexts on several GPUs and alternate between them).
struct PassLaunchState {
    int blocks_per_sm[64] = {};
};
template <typename Kern>
static cudaError_t configure_pass(Kern kern, int device, int threads, size_t smem, PassLaunchState& st, int* blocks_per_sm) {
    int& b = st.blocks_per_sm[device & 63];
    if (!b) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
        if (e != cudaSuccess) return e;
        int occ = 0;
        e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem);
        if (e != cudaSuccess) return e;
        if (occ < 1) return cudaErrorLaunchOutOfResources;
        b = occ;
    }
    *blocks_per_sm = b;
    return cudaSuccess;
}

#ifndef RS_DEFAULT_PASS_GEN
#define RS_DEFAULT_PASS_GEN 1
#endif
template <typename KeyT, bool HAS_VALS, bool ATOMIC, bool FK>
static cudaError_t launch_pass_mode(rs_context ctx, cudaStream_t st, const OnesweepArgs<KeyT>& args) {
    using C = TileCfg<KeyT, HAS_VALS, ATOMIC>;
    int blocks_per_sm = 0;
    if constexpr (ATOMIC) {
        // RSB200_SWEEP_V2=1: the software-pipelined variant of the pass (rs_sweep.cuh; same tile shape and scratch layout).
        // Measured equal on 4-byte keys and slower on 8-byte keys and with the index payload (profiles/README.md), so
        // the first-generation kernel stays the default; the variant is kept runnable for that record.
        // RSB200_PASS=3: the third generation (rs_pass3.cuh: register prefetch instead of the TMA stage, exchange-and-clear
        // table scan, 32-bit run starts).
        static const int pass_gen = [] { const char* e = std::getenv("RSB200_PASS"); return e ? std::atoi(e) : RS_DEFAULT_PASS_GEN; }();
#ifdef RS_EXPERIMENTAL_PASSES
        if (pass_gen == 3) {
            using L = Pass3Smem<KeyT, HAS_VALS, C::THREADS, C::KPT>;
            auto kern = rs_pass3_kernel<KeyT, HAS_VALS, C::THREADS, C::KPT, C::MINB, FK>;
            static thread_local PassLaunchState state;
            cudaError_t e = configure_pass(kern, ctx->device, C::THREADS, L::total, state, &blocks_per_sm);
            if (e != cudaSuccess) return e;
            const uint64_t resident = static_cast<uint64_t>(blocks_per_sm) * ctx->sm_count;
            kern<<<static_cast<int>(std::min<uint64_t>(args.num_tiles, resident)), C::THREADS, L::total, st>>>(args);
            return cudaGetLastError();
        }
        if (pass_gen == 4) {   // rs_pass4.cuh: the first generation with its shared-memory accesses un-serialised
            using L = Pass4Smem<KeyT, HAS_VALS, C::THREADS, C::KPT>;
            auto kern = rs_pass4_kernel<KeyT, HAS_VALS, C::THREADS, C::KPT, C::MINB, FK>;
            static thread_local PassLaunchState state;
            cudaError_t e = configure_pass(kern, ctx->device, C::THREADS, L::total, state, &blocks_per_sm);
            if (e != cudaSuccess) return e;
            const uint64_t resident = static_cast<uint64_t>(blocks_per_sm) * ctx->sm_count;
            kern<<<static_cast<int>(std::min<uint64_t>(args.num_tiles, resident)), C::THREADS, L::total, st>>>(args);
            return cudaGetLastError();
        }
#endif
        static const bool sweep_v2 = std::getenv("RSB200_SWEEP_V2") != nullptr || pass_gen == 2;
        if (sweep_v2) {
            using L = SweepSmem<KeyT, HAS_VALS, C::THREADS, C::KPT>;
            auto kern = rs_sweep_kernel<KeyT, HAS_VALS, C::THREADS, C::KPT, C::MINB, FK>;
            static thread_local PassLaunchState state;
            cudaError_t e = configure_pass(kern, ctx->device, C::THREADS, L::total, state, &blocks_per_sm);
            if (e != cudaSuccess) return e;
            const uint64_t resident = static_cast<uint64_t>(blocks_per_sm) * ctx->sm_count;
            kern<<<static_cast<int>(std::min<uint64_t>(args.num_tiles, resident)), C::THREADS, L::total, st>>>(args);
            return cudaGetLastError();
        }
    }
    {
        using L = OnesweepSmem<KeyT, HAS_VALS, C::THREADS, C::KPT, ATOMIC>;
        // sm_100: 228 KiB of shared memory per SM, 1 KiB of it reserved per resident CTA -- a tile shape that does not
        // leave room for C::MINB CTAs would silently halve the occupancy the kernel was tuned for
        static_assert(L::total <= (228u * 1024u - 1024u * C::MINB) / C::MINB, "tile shape does not fit MINB CTAs per SM");
        auto kern = rs_onesweep_kernel<KeyT, HAS_VALS, C::THREADS, C::KPT, C::MINB, ATOMIC, FK>;
        static thread_local PassLaunchState state;
        cudaError_t e = configure_pass(kern, ctx->device, C::THREADS, L::total, state, &blocks_per_sm);
        if (e != cudaSuccess) return e;
        const uint64_t resident = static_cast<uint64_t>(blocks_per_sm) * ctx->sm_count;
        kern<<<static_cast<int>(std::min<uint64_t>(args.num_tiles, resident)), C::THREADS, L::total, st>>>(args);
        return cudaGetLastError();
    }
}
template <typename KeyT, bool HAS_VALS, bool FK>
static cudaError_t launch_pass(rs_context ctx, cudaStream_t st, const OnesweepArgs<KeyT>& args, bool atomic) {
    return atomic ? launch_pass_mode<KeyT, HAS_VALS, true, FK>(ctx, st, args) : launch_pass_mode<KeyT, HAS_VALS, false, FK>(ctx, st, args);
}

static PassTable make_pass_table(size_t key_bytes, bool is_signed, unsigned long long finf, int begin_bit, int end_bit) {
    PassTable pt{};
    const int key_bits = static_cast<int>(key_bytes) * 8;
    int p = 0;
    for (int b = begin_bit; b < end_bit && p < MAX_PASSES; b += RADIX_BITS, ++p) {
        const int bits = std::min(RADIX_BITS, end_bit - b);
        pt.d[p].shift  = b;
        pt.d[p].mask   = (1u << bits) - 1u;
        pt.d[p].flip   = 0;
        pt.d[p].finf   = finf;
        if (is_signed && key_bits - 1 >= b && key_bits - 1 < b + bits) pt.d[p].flip = 1u << (key_bits - 1 - b);
    }
    pt.num_passes = p;
    return pt;
}

template <typename KeyT, bool FK>
static int sort_typed_fk(rs_sorter s, cudaStream_t st, const KeyT* in, KeyT* buf0, KeyT* buf1, const uint32_t* vin,
                      uint32_t* vbuf0, uint32_t* vbuf1, uint64_t n, int begin_bit, int end_bit, bool profile,
                      int* result_index, const SegmentTable<KeyT>* segments) {
    const bool has_vals = vin != nullptr;
    const PassTable pt = make_pass_table(sizeof(KeyT), s->is_signed, s->finf, begin_bit, end_bit);
    if (pt.num_passes < 1) return fail(RS_CALCULATION_FAILED, "empty bit range");
    const Plan pl = make_plan(n, sizeof(KeyT), has_vals, s->atomic_rank, pt.num_passes);
    if (pl.total > s->scratch_bytes) return fail(RS_CALCULATION_FAILED, "n exceeds the sorter's max_keys (scratch too small)");
    if (pl.num_tiles >= (1ull << 32)) return fail(RS_CALCULATION_FAILED, "too many tiles");

    // ---- small inputs: one launch of one CTA does every pass in shared memory ---------------------------------
    if (n <= static_cast<uint64_t>(SMALL_CAP) && s->atomic_rank && !segments && !(s->flags & RS_NO_SMALL_PATH)) {
        EventQuad sq{};
        if (profile) {
            if (!s->free_events.empty()) { sq = s->free_events.back(); s->free_events.pop_back(); }
            else for (auto& e : sq.e) RS_CUDA(cudaEventCreate(&e), RS_CALCULATION_FAILED);
            for (int i = 0; i < 3; ++i) RS_CUDA(cudaEventRecord(sq.e[i], st), RS_CALCULATION_FAILED);
        }
        KeyT*     kout = (((pt.num_passes - 1) & 1) == 0) ? buf0 : buf1;
        uint32_t* vout = (((pt.num_passes - 1) & 1) == 0) ? vbuf0 : vbuf1;
        auto* hist_s = reinterpret_cast<unsigned long long*>(s->scratch + pl.off_hist);
        auto* plan_s = reinterpret_cast<SortPlanDev*>(s->scratch + pl.off_plan);
        const int allow = (s->flags & RS_NO_PASS_SKIP) ? 0 : 1;
        cudaError_t e;
        if (has_vals) {
            auto kern = rs_small_sort_kernel<KeyT, true, FK>;
            static thread_local uint64_t cfg = 0;
            if (const uint64_t bit = current_device_bit(); !(cfg & bit)) { RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SmallSmem<KeyT, true>::total)), RS_CALCULATION_FAILED); cfg |= bit; }
            kern<<<1, SMALL_THREADS, SmallSmem<KeyT, true>::total, st>>>(in, kout, vin, vout, static_cast<uint32_t>(n), pt, hist_s, plan_s, allow);
            e = cudaGetLastError();
        } else {
            auto kern = rs_small_sort_kernel<KeyT, false, FK>;
            static thread_local uint64_t cfg = 0;
            if (const uint64_t bit = current_device_bit(); !(cfg & bit)) { RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(SmallSmem<KeyT, false>::total)), RS_CALCULATION_FAILED); cfg |= bit; }
            kern<<<1, SMALL_THREADS, SmallSmem<KeyT, false>::total, st>>>(in, kout, nullptr, nullptr, static_cast<uint32_t>(n), pt, hist_s, plan_s, allow);
            e = cudaGetLastError();
        }
        if (e != cudaSuccess) return fail(RS_CALCULATION_FAILED, "small sort launch", e);
        s->launches++;
        if (profile) {
            RS_CUDA(cudaEventRecord(sq.e[3], st), RS_CALCULATION_FAILED);
            s->pending.push_back(sq);
        }
        s->last_passes      = pt.num_passes;
        s->last_chunks      = 1;
        s->last_hist_offset = pl.off_hist;
        s->last_plan_offset = pl.off_plan;
        if (result_index) *result_index = (pt.num_passes - 1) & 1;
        return RS_OK;
    }

    EventQuad q{};
    if (profile) {
        if (!s->free_events.empty()) {
            q = s->free_events.back();
            s->free_events.pop_back();
        } else {
            for (auto& e : q.e) RS_CUDA(cudaEventCreate(&e), RS_CALCULATION_FAILED);
        }
        RS_CUDA(cudaEventRecord(q.e[0], st), RS_CALCULATION_FAILED);
    }
    auto* hist   = reinterpret_cast<unsigned long long*>(s->scratch + pl.off_hist);
    auto* ctr    = reinterpret_cast<uint32_t*>(s->scratch + pl.off_ctr);
    auto* status = reinterpret_cast<uint32_t*>(s->scratch + pl.off_status);
    auto* base   = reinterpret_cast<unsigned long long*>(s->scratch + pl.off_base);

    RS_CUDA(cudaMemsetAsync(s->scratch, 0, pl.zero_bytes, st), RS_CALCULATION_FAILED);
    if (segments) {  // multi-GPU: pull the peer segments into `in` and count them in the same kernel (integer keys of 4 / 8 bytes)
        if constexpr (!FK && sizeof(KeyT) >= 4)
            RS_CUDA(launch_gather_hist<KeyT>(s->ctx->sm_count, st, *segments, const_cast<KeyT*>(in), n, hist, pt), RS_CALCULATION_FAILED);
        else
            return fail(RS_CALCULATION_FAILED, "the fused gather + sort entry serves 4- and 8-byte integer keys");
    } else {
        RS_CUDA((launch_hist<KeyT, FK>(s->ctx->sm_count, st, in, n, hist, pt)), RS_CALCULATION_FAILED);
    }
    s->launches++;
    if (profile) RS_CUDA(cudaEventRecord(q.e[1], st), RS_CALCULATION_FAILED);
    auto* plan = reinterpret_cast<SortPlanDev*>(s->scratch + pl.off_plan);
    const int allow_skip = (s->flags & RS_NO_PASS_SKIP) ? 0 : 1;
    rs_scan_kernel<<<1, RADIX, 0, st>>>(hist, base, pl.num_chunks, pt.num_passes, n, allow_skip,
                                       static_cast<const void*>(in) == static_cast<const void*>(buf1) ? 1 : 0, plan);
    RS_CUDA(cudaGetLastError(), RS_CALCULATION_FAILED);
    s->launches++;
    if (profile) RS_CUDA(cudaEventRecord(q.e[2], st), RS_CALCULATION_FAILED);

    const bool all_aligned = ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(buf0) |
                               reinterpret_cast<uintptr_t>(buf1)) & 15u) == 0;
    KeyT*     kb[2] = {buf0, buf1};
    uint32_t* vb[2] = {vbuf0, vbuf1};
    for (int p = 0; p < pt.num_passes; ++p) {
        OnesweepArgs<KeyT> a{};
        a.keys_in         = (p == 0) ? in : kb[(p - 1) & 1];
        a.keys_out        = kb[p & 1];
        a.vals_in         = has_vals ? ((p == 0) ? vin : vb[(p - 1) & 1]) : nullptr;
        a.vals_out        = has_vals ? vb[p & 1] : nullptr;
        a.plan            = plan;
        a.pass            = p;
        a.n               = n;
        a.base            = base + static_cast<size_t>(p) * pl.num_chunks * RADIX;
        a.status          = status + static_cast<size_t>(p) * pl.num_tiles * RADIX;
        a.counter         = ctr + p;
        a.num_tiles       = static_cast<uint32_t>(pl.num_tiles);
        a.tiles_per_chunk = pl.tiles_per_chunk;
        a.num_chunks      = static_cast<uint32_t>(pl.num_chunks);
        a.digit           = pt.d[p];
        a.use_tma         = ((s->flags & RS_NO_TMA) == 0) && all_aligned;
        cudaError_t e = has_vals ? launch_pass<KeyT, true, FK>(s->ctx, st, a, s->atomic_rank)
                                 : launch_pass<KeyT, false, FK>(s->ctx, st, a, s->atomic_rank);
        if (e != cudaSuccess) return fail(RS_CALCULATION_FAILED, "onesweep launch", e);
        s->launches++;
    }
    if (profile) {
        RS_CUDA(cudaEventRecord(q.e[3], st), RS_CALCULATION_FAILED);
        s->pending.push_back(q);
    }
    s->last_passes      = pt.num_passes;
    s->last_chunks      = pl.num_chunks;
    s->last_hist_offset = pl.off_hist;
    s->last_plan_offset = pl.off_plan;
    if (result_index) *result_index = (pt.num_passes - 1) & 1;
    return RS_OK;
}

template <typename KeyT>
static int sort_typed(rs_sorter s, cudaStream_t st, const KeyT* in, KeyT* buf0, KeyT* buf1, const uint32_t* vin,
                      uint32_t* vbuf0, uint32_t* vbuf1, uint64_t n, int begin_bit, int end_bit, bool profile,
                      int* result_index, const SegmentTable<KeyT>* segments = nullptr) {
    return s->finf ? sort_typed_fk<KeyT, true>(s, st, in, buf0, buf1, vin, vbuf0, vbuf1, n, begin_bit, end_bit, profile, result_index, segments)
                   : sort_typed_fk<KeyT, false>(s, st, in, buf0, buf1, vin, vbuf0, vbuf1, n, begin_bit, end_bit, profile, result_index, segments);
}

static int resolve_pending(rs_sorter s, bool block) {
    size_t kept = 0;
    for (size_t i = 0; i < s->pending.size(); ++i) {
        EventQuad q = s->pending[i];
        cudaError_t e = block ? cudaEventSynchronize(q.e[3]) : cudaEventQuery(q.e[3]);
        if (e == cudaErrorNotReady) {
            s->pending[kept++] = q;
            continue;
        }
        if (e != cudaSuccess) return fail(RS_CALCULATION_FAILED, "event wait", e);
        float h = 0, sc = 0, re = 0, to = 0;
        cudaEventElapsedTime(&h, q.e[0], q.e[1]);
        cudaEventElapsedTime(&sc, q.e[1], q.e[2]);
        cudaEventElapsedTime(&re, q.e[2], q.e[3]);
        cudaEventElapsedTime(&to, q.e[0], q.e[3]);
        stat_update(s->rt.time_histo, h);
        stat_update(s->rt.time_scan, sc);
        stat_update(s->rt.time_reorder, re);
        stat_update(s->rt.time_paste, 0.0);
        stat_update(s->rt.time_total, to);
        s->free_events.push_back(q);
    }
    s->pending.resize(kept);
    return RS_OK;
}

// ------------------------------------------------------------------------------------------------
// utility kernels (dataset fill, checks, lower bound, iota, pad)
// ------------------------------------------------------------------------------------------------
__host__ __device__ static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__host__ __device__ static inline uint64_t counter_rand(uint64_t seed, uint64_t i) {
    return mix64(seed + (i + 1) * 0x9E3779B97F4A7C15ull);
}

template <typename U>
__global__ void rs_fill_kernel(U* out, uint64_t n, int kind, bool is_signed, uint64_t seed, uint64_t off, uint64_t total) {
    const U sign = is_signed ? static_cast<U>(static_cast<U>(1) << (sizeof(U) * 8 - 1)) : static_cast<U>(0);
    const U tmin = sign;                                  // bit pattern of numeric_limits<T>::min()
    const U tmax = static_cast<U>(~static_cast<U>(0)) ^ sign;  // bit pattern of numeric_limits<T>::max()
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint64_t g = off + i;
        U v = 0;
        switch (kind) {
            case RS_DATASET_ZEROS: v = 0; break;
            case RS_DATASET_RANGE: v = static_cast<U>(tmin + static_cast<U>(g)); break;
            case RS_DATASET_INVERTED_RANGE: v = static_cast<U>(tmin + static_cast<U>(total - 1 - g)); break;
            case RS_DATASET_UNIFORM:
                v = static_cast<U>(counter_rand(seed, g));
                if (g == 0) v = tmax;
                if (g == total - 1) v = tmin;
                break;
            case RS_DATASET_RANDOM: v = static_cast<U>(static_cast<uint32_t>(counter_rand(seed ^ 0x52616e646f6dull, g))); break;
        }
        out[i] = v;
    }
}

__global__ void rs_iota_kernel(uint32_t* out, uint64_t n) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        out[i] = static_cast<uint32_t>(i);
}

template <typename U>
__global__ void rs_fill_value_kernel(U* out, uint64_t first, uint64_t last, U v) {
    for (uint64_t i = first + blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < last;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x)
        out[i] = v;
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ unsigned long long warp_xor(unsigned long long v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v ^= __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// key -> unsigned value in the key type's order: sign bit flipped for signed integers, float_order_bits for floats
template <typename U>
__device__ __forceinline__ U ordered_key(U k, U sign, unsigned long long finf) {
    return finf ? float_order_bits(k, finf) : static_cast<U>(k ^ sign);
}

template <typename U>
__global__ void rs_check_kernel(const U* keys, uint64_t n, U sign, unsigned long long finf, unsigned long long* result) {
    unsigned long long bad = 0, sum = 0, hsum = 0, hxor = 0;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const U k = keys[i];
        if (i + 1 < n) {
            const U nx = keys[i + 1];
            if (ordered_key(nx, sign, finf) < ordered_key(k, sign, finf)) bad++;
        }
        const unsigned long long h = mix64(static_cast<uint64_t>(k) + 0x9E3779B97F4A7C15ull);
        sum += static_cast<uint64_t>(k);
        hsum += h;
        hxor ^= h;
    }
    bad = warp_sum(bad); sum = warp_sum(sum); hsum = warp_sum(hsum); hxor = warp_xor(hxor);
    if ((threadIdx.x & 31) == 0) {
        if (bad) atomicAdd(&result[0], bad);
        atomicAdd(&result[1], sum);
        atomicAdd(&result[2], hsum);
        atomicXor(&result[3], hxor);
    }
}

template <typename U>
__global__ void rs_check_perm_kernel(const U* kin, const U* kout, const uint32_t* perm, uint64_t n,
                                     unsigned long long* result) {
    unsigned long long mism = 0, unstable = 0;
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const uint32_t p = perm[i];
        if (p >= n || kin[p] != kout[i]) mism++;
        if (i + 1 < n && kout[i] == kout[i + 1] && p >= perm[i + 1]) unstable++;
    }
    mism = warp_sum(mism); unstable = warp_sum(unstable);
    if ((threadIdx.x & 31) == 0) {
        if (mism) atomicAdd(&result[0], mism);
        if (unstable) atomicAdd(&result[1], unstable);
    }
}

template <typename U>
__global__ void rs_lower_bound_kernel(const U* keys, uint64_t n, const U* q, uint64_t nq, U sign, unsigned long long* out) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < nq;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        const U x = static_cast<U>(q[i] ^ sign);
        uint64_t lo = 0, hi = n;
        while (lo < hi) {
            const uint64_t mid = lo + ((hi - lo) >> 1);
            if (static_cast<U>(keys[mid] ^ sign) < x) lo = mid + 1; else hi = mid;
        }
        out[i] = lo;
    }
}

static int grid_for(uint64_t n, int threads, int sm_count) {
    uint64_t g = (n + threads - 1) / threads;
    return static_cast<int>(std::max<uint64_t>(1, std::min<uint64_t>(g, static_cast<uint64_t>(sm_count) * 16)));
}

template <typename KeyT>
static int sort_segments_typed(rs_sorter s, rs_stream st, int num_segments, const void* const* ptrs, const uint64_t* counts,
                               int first_segment, void* gathered, void* buf0, void* buf1, int begin_bit, int end_bit,
                               bool profile, int* result_index) {
    SegmentTable<KeyT> seg{};
    seg.num_segments  = num_segments;
    seg.first_segment = first_segment;
    uint64_t n = 0;
    for (int i = 0; i < num_segments; ++i) {
        seg.ptr[i]   = static_cast<const KeyT*>(ptrs[i]);
        seg.start[i] = n;
        n += counts[i];
    }
    for (int i = num_segments; i <= MAX_SEGMENTS; ++i) seg.start[i] = n;
    if (n > s->max_keys) return fail(RS_CALCULATION_FAILED, "segments exceed the sorter's max_keys");
    if (n == 0) return RS_OK;
    return sort_typed<KeyT>(s, st->st, static_cast<const KeyT*>(gathered), static_cast<KeyT*>(buf0), static_cast<KeyT*>(buf1),
                            nullptr, nullptr, nullptr, n, begin_bit, end_bit, profile, result_index, &seg);
}

template <typename KeyT>
static int merge_runs_typed(rs_context ctx, cudaStream_t st, bool is_signed, int num_jobs, const void* const* a_ptrs,
                            const uint64_t* a_counts, const void* const* b_ptrs, const uint64_t* b_counts, void* const* out_ptrs,
                            uint64_t* d_splits, uint64_t splits_capacity) {
    MergeBatch<KeyT> mb{};
    mb.num_jobs = num_jobs;
    mb.sign     = is_signed ? static_cast<KeyT>(static_cast<KeyT>(1) << (sizeof(KeyT) * 8 - 1)) : static_cast<KeyT>(0);
    unsigned long long tiles = 0;
    for (int j = 0; j < num_jobs; ++j) {
        mb.a[j] = static_cast<const KeyT*>(a_ptrs[j]);
        mb.b[j] = static_cast<const KeyT*>(b_ptrs[j]);
        mb.out[j] = static_cast<KeyT*>(out_ptrs[j]);
        mb.na[j] = a_counts[j];
        mb.nb[j] = b_counts[j];
        mb.tile_start[j] = tiles;
        tiles += (a_counts[j] + b_counts[j] + MERGE_TILE - 1) / MERGE_TILE;
    }
    for (int j = num_jobs; j <= MERGE_MAX_JOBS; ++j) mb.tile_start[j] = tiles;
    if (tiles == 0) return RS_OK;
    if (tiles + num_jobs > splits_capacity) return fail(RS_CALCULATION_FAILED, "split scratch too small for this merge level");
    if (tiles > 0x7fffffffull) return fail(RS_CALCULATION_FAILED, "too many merge tiles");
    const unsigned long long boundaries = tiles + num_jobs;
    const int pgrid = static_cast<int>(std::min<unsigned long long>((boundaries + 255) / 256, static_cast<unsigned long long>(ctx->sm_count) * 32));
    // developer switch (like RSB200_DIST_TIMING in dist.py): device time of the two kernels, printed to stderr; adds syncs
    static const bool merge_timing = std::getenv("RSB200_MERGE_TIMING") != nullptr;
    cudaEvent_t tev[3] = {nullptr, nullptr, nullptr};
    if (merge_timing) {
        for (auto& e : tev) cudaEventCreate(&e);
        cudaEventRecord(tev[0], st);
    }
    static const bool flat_partition = std::getenv("RSB200_MERGE_FLAT_PARTITION") != nullptr;  // A/B switch for profiling
    if (flat_partition || tiles <= 4 * MERGE_COARSE) {
        rs_merge_partition_kernel<KeyT><<<pgrid, 256, 0, st>>>(mb, reinterpret_cast<unsigned long long*>(d_splits), 0);
    } else {
        rs_merge_partition_kernel<KeyT><<<pgrid, 256, 0, st>>>(mb, reinterpret_cast<unsigned long long*>(d_splits), 1);
        rs_merge_partition_kernel<KeyT><<<pgrid, 256, 0, st>>>(mb, reinterpret_cast<unsigned long long*>(d_splits), 2);
    }
    RS_CUDA(cudaGetLastError(), RS_CALCULATION_FAILED);
    if (merge_timing) cudaEventRecord(tev[1], st);
    struct TimingReport {
        cudaEvent_t* ev; cudaStream_t st; bool on; unsigned long long tiles;
        ~TimingReport() {
            if (!on) return;
            cudaEventRecord(ev[2], st);
            cudaEventSynchronize(ev[2]);
            float part = 0.f, tile = 0.f;
            cudaEventElapsedTime(&part, ev[0], ev[1]);
            cudaEventElapsedTime(&tile, ev[1], ev[2]);
            int dev = 0;
            cudaGetDevice(&dev);
            std::fprintf(stderr, "[rsb200 merge, device %d] %llu tiles: partition %.3f ms, tiles %.3f ms\n", dev, tiles, part, tile);
            for (int i = 0; i < 3; ++i) cudaEventDestroy(ev[i]);
        }
    } timing_report{tev, st, merge_timing, tiles};
    static const bool one_tile_per_cta = std::getenv("RSB200_MERGE_UNPIPELINED") != nullptr;  // A/B switch for profiling
    if (!one_tile_per_cta) {
        unsigned long long max_tiles = 0;
        for (int j = 0; j < num_jobs; ++j) max_tiles = std::max(max_tiles, mb.tile_start[j + 1] - mb.tile_start[j]);
        if (max_tiles * num_jobs > 0xffffffffull) return fail(RS_CALCULATION_FAILED, "too many merge tiles");
        const size_t out_elems = sizeof(KeyT) == 4 ? MERGE_TILE + MERGE_TILE / 32 : MERGE_THREADS * (MERGE_ITEMS + 16 / sizeof(KeyT));
        const size_t smem = sizeof(KeyT) * (MERGE_TILE + 4 * (MERGE_LINE_BYTES / sizeof(KeyT)) + out_elems);
        // runs in peer memory (multi-GPU level 1) want the variant with more CTAs, i.e. more loads in flight
        bool remote = false;
        int  here = 0;
        cudaGetDevice(&here);
        for (int j = 0; j < num_jobs && !remote; ++j)
            for (const void* p : {a_ptrs[j], b_ptrs[j]}) {
                cudaPointerAttributes at{};
                if (p && cudaPointerGetAttributes(&at, p) == cudaSuccess && at.type == cudaMemoryTypeDevice && at.device != here) remote = true;
            }
        cudaGetLastError();  // a pointer the runtime does not know is simply treated as local
        static const char* const force = std::getenv("RSB200_MERGE_DENSE");  // "0" / "1": A/B switch for profiling
        if (force) remote = force[0] == '1';
        // walk order of the persistent CTAs: the region-strided order (rs_merge_order.h) balances batches whose tiles sit
        // in few jobs; opt-in until it has been run on the GPU box (the host emulation of the walk is in the CPU tests)
        static const bool regions = std::getenv("RSB200_MERGE_ORDER_REGIONS") != nullptr;
        auto kern = regions ? (remote ? rs_merge_tiles_pipelined_kernel<KeyT, true, true> : rs_merge_tiles_pipelined_kernel<KeyT, false, true>)
                            : (remote ? rs_merge_tiles_pipelined_kernel<KeyT, true, false> : rs_merge_tiles_pipelined_kernel<KeyT, false, false>);
        static thread_local int occupancy[4][64] = {};   // per variant and device
        int& ctas_per_sm = occupancy[(remote ? 1 : 0) + (regions ? 2 : 0)][here & 63];
        if (!ctas_per_sm) {
            RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)), RS_CALCULATION_FAILED);
            RS_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&ctas_per_sm, kern, MERGE_THREADS, smem), RS_CALCULATION_FAILED);
            if (ctas_per_sm < 1) return fail(RS_CALCULATION_FAILED, "merge kernel does not fit on an SM");
        }
        const unsigned grid = static_cast<unsigned>(std::min<unsigned long long>(tiles, static_cast<unsigned long long>(ctx->sm_count) * ctas_per_sm));
        kern<<<grid, MERGE_THREADS, smem, st>>>(mb, reinterpret_cast<const unsigned long long*>(d_splits), static_cast<uint32_t>(max_tiles));
        RS_CUDA(cudaGetLastError(), RS_CALCULATION_FAILED);
        return RS_OK;
    }
    const size_t smem = sizeof(KeyT) * (MERGE_TILE + MERGE_TILE / 32);
    auto kern = rs_merge_tiles_kernel<KeyT>;
    static thread_local uint64_t configured = 0;
    if (const uint64_t bit = current_device_bit(); !(configured & bit)) {
        RS_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)), RS_CALCULATION_FAILED);
        configured |= bit;
    }
    kern<<<static_cast<unsigned>(tiles), MERGE_THREADS, smem, st>>>(mb, reinterpret_cast<const unsigned long long*>(d_splits));
    RS_CUDA(cudaGetLastError(), RS_CALCULATION_FAILED);
    return RS_OK;
}

// ------------------------------------------------------------------------------------------------
// In-situ self-test of the atomic ranking mode.  The mode is stable only if the lanes of one warp instruction that hit
// the same shared-memory word are served in ascending lane order -- not documented by NVIDIA.  rs_context_create
// therefore (1) runs rs_atomic_order_selftest_kernel (the property in isolation, 4096 random rows per warp) and
// (2) sorts 2^20 + 12345 keys WITH an index payload through the production kernels themselves -- production tile
// shapes, resident CTAs per SM, TMA traffic, 4- and 8-byte keys -- on inputs whose digits take only four values (about
// eight lanes of every row collide on the same word, in every pass) and checks sortedness, the gather property and
// stability of the permutation on the device.  Any violation leaves the context in the lane-bitmap mode, which relies
// on nothing beyond the PTX memory model.
// ------------------------------------------------------------------------------------------------
template <typename U>
__global__ void rs_selftest_fill_kernel(U* keys, uint32_t* idx, uint64_t n, uint64_t seed) {
    for (uint64_t i = blockIdx.x * static_cast<uint64_t>(blockDim.x) + threadIdx.x; i < n;
         i += static_cast<uint64_t>(gridDim.x) * blockDim.x) {
        // two bits per byte: every 8-bit digit takes one of four values; every 7th row of 32 keys is constant
        U k = static_cast<U>(counter_rand(seed, i) & 0x0303030303030303ull);
        if (((i >> 5) % 7) == 0) k = static_cast<U>(0x0101010101010101ull & static_cast<U>(~static_cast<U>(0)));
        keys[i] = k;
        idx[i]  = static_cast<uint32_t>(i);
    }
}

template <typename U>
static bool selftest_sort_once(rs_context c, rs_key_type type) {
    const uint64_t n  = (1ull << 20) + 12345;
    const size_t   kb = sizeof(U);
    unsigned char* pool = nullptr;
    const size_t   key_bytes = ((n * kb + 255) / 256) * 256, val_bytes = ((n * 4 + 255) / 256) * 256;
    if (cudaMalloc(reinterpret_cast<void**>(&pool), 3 * key_bytes + 3 * val_bytes) != cudaSuccess) { cudaGetLastError(); return false; }
    U* in = reinterpret_cast<U*>(pool); U* b0 = reinterpret_cast<U*>(pool + key_bytes); U* b1 = reinterpret_cast<U*>(pool + 2 * key_bytes);
    uint32_t* vin = reinterpret_cast<uint32_t*>(pool + 3 * key_bytes);
    uint32_t* v0  = reinterpret_cast<uint32_t*>(pool + 3 * key_bytes + val_bytes);
    uint32_t* v1  = reinterpret_cast<uint32_t*>(pool + 3 * key_bytes + 2 * val_bytes);
    rs_stream_s st{c, nullptr, false};  // the legacy default stream: the context is not visible to anybody yet
    rs_sorter   sorter = nullptr;
    bool        ok     = false;
    if (rs_sorter_create(c, type, n, RS_NO_KEY_BUFFERS | RS_NO_PROFILE, &sorter) == RS_OK) {
        rs_selftest_fill_kernel<U><<<c->sm_count * 4, 256>>>(in, vin, n, 0xA70D1C5ull + kb);
        int idx = 0;
        if (rs_sorter_sort_device(sorter, &st, in, b0, b1, vin, v0, v1, n, 0, static_cast<int>(kb) * 8, &idx) == RS_OK) {
            U*        kout = idx ? b1 : b0;
            uint32_t* vout = idx ? v1 : v0;
            cudaMemset(c->d_result, 0, 8 * sizeof(unsigned long long));
            rs_check_kernel<U><<<c->sm_count * 4, 256>>>(kout, n, static_cast<U>(0), 0ull, c->d_result);
            rs_check_perm_kernel<U><<<c->sm_count * 4, 256>>>(in, kout, vout, n, c->d_result + 4);
            unsigned long long h[8] = {1, 1, 1, 1, 1, 1, 1, 1};
            if (cudaMemcpy(h, c->d_result, sizeof h, cudaMemcpyDeviceToHost) == cudaSuccess)
                ok = h[0] == 0 && h[4] == 0 && h[5] == 0;  // sorted, gathers to the input, stable
        }
        rs_sorter_destroy(sorter);
    }
    cudaFree(pool);
    (void)cudaGetLastError();
    return ok;
}

// typed dispatch over the key width: f(U{}) with U = uint16_t / uint32_t / uint64_t
template <typename F>
static void for_key_width(size_t bytes, F&& f) {
    if (bytes == 2) f(uint16_t{});
    else if (bytes == 4) f(uint32_t{});
    else f(uint64_t{});
}

// ------------------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------------------
extern "C" {

int rs_version(void) { return RSB200_VERSION_MAJOR * 1000 + RSB200_VERSION_MINOR; }

const char* rs_status_string(int status) {
    static const char* names[] = {"OK", "HOST_BUFFERS_FAILED", "INITIALIZATION_FAILED", "DATA_UPLOAD_FAILED",
                                  "CALCULATION_FAILED", "DATA_DOWNLOAD_FAILED", "CLEANUP_FAILED", "RESIZE_FAILED",
                                  "KERNEL_CREATION_FAILED", "PROGRAM_CREATION_FAILED", "NO_SOURCE_FOUND",
                                  "LOADING_SOURCE_FAILED"};
    if (status < 0 || status > 11) return "UNKNOWN";
    return names[status];
}

const char* rs_last_error_string(void) { return g_last_error.c_str(); }

int rs_device_count(int* n) {
    if (!n) return fail(RS_INITIALIZATION_FAILED, "null argument");
    int c = 0;
    cudaError_t e = cudaGetDeviceCount(&c);
    if (e != cudaSuccess) {
        *n = 0;
        return fail(RS_INITIALIZATION_FAILED, "cudaGetDeviceCount", e);
    }
    *n = c;
    return RS_OK;
}

int rs_context_create(int device, rs_context* out) {
    if (!out) return fail(RS_INITIALIZATION_FAILED, "null argument");
    *out = nullptr;
    int count = 0;
    RS_CUDA(cudaGetDeviceCount(&count), RS_INITIALIZATION_FAILED);
    if (device < 0 || device >= count) return fail(RS_INITIALIZATION_FAILED, "no such CUDA device");
    DeviceGuard guard(device); if (guard.status != cudaSuccess) return fail(RS_INITIALIZATION_FAILED, "cudaSetDevice", guard.status);
    int major = 0, sms = 0, smem = 0;
    RS_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device), RS_INITIALIZATION_FAILED);
    RS_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device), RS_INITIALIZATION_FAILED);
    RS_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, device), RS_INITIALIZATION_FAILED);
    if (major != 10) return fail(RS_INITIALIZATION_FAILED, "librsb200 is built for sm_100a (B200) only");
    rs_context c = new (std::nothrow) rs_context_s{};
    if (!c) return fail(RS_INITIALIZATION_FAILED, "out of host memory");
    c->device = device; c->sm_count = sms; c->max_smem_optin = smem;
    if (cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&c->d_result), 8 * sizeof(unsigned long long)); e != cudaSuccess) {
        delete c;
        return fail(RS_INITIALIZATION_FAILED, "cudaMalloc", e);
    }
    {   // self-test behind the atomic ranking mode (a few hundred microseconds; see rs_kernels.cuh)
        unsigned long long* d_bad = nullptr;
        unsigned long long  h_bad = 1;
        const char*         off   = getenv("RSB200_NO_ATOMIC_RANK");
        if (!(off && off[0] == '1') && cudaMalloc(reinterpret_cast<void**>(&d_bad), sizeof h_bad) == cudaSuccess) {
            cudaMemset(d_bad, 0, sizeof h_bad);
            rs_atomic_order_selftest_kernel<<<sms, 256>>>(d_bad, 4096);
            if (cudaMemcpy(&h_bad, d_bad, sizeof h_bad, cudaMemcpyDeviceToHost) != cudaSuccess) h_bad = 1;
            cudaFree(d_bad);
            (void)cudaGetLastError();
        }
        c->atomic_rank_ok = (h_bad == 0) ? 1 : 0;
        if (c->atomic_rank_ok)   // the property held in isolation: now under production conditions (see above)
            c->atomic_rank_ok = (selftest_sort_once<uint32_t>(c, RS_KEY_U32) && selftest_sort_once<uint64_t>(c, RS_KEY_U64)) ? 1 : 0;
    }
    *out = c;
    return RS_OK;
}

int rs_context_destroy(rs_context ctx) {
    if (!ctx) return RS_OK;
    DeviceGuard guard(ctx->device);
    if (ctx->d_result) cudaFree(ctx->d_result);
    delete ctx;
    return RS_OK;
}

int rs_context_atomic_rank(rs_context ctx, int* enabled) {
    if (!ctx || !enabled) return fail(RS_INITIALIZATION_FAILED, "null argument");
    *enabled = ctx->atomic_rank_ok;
    return RS_OK;
}

int rs_context_sm_count(rs_context ctx, int* n) {
    if (!ctx || !n) return fail(RS_INITIALIZATION_FAILED, "null argument");
    *n = ctx->sm_count;
    return RS_OK;
}

int rs_stream_create(rs_context ctx, rs_stream* out) {
    if (!ctx || !out) return fail(RS_INITIALIZATION_FAILED, "null argument");
    DeviceGuard guard(ctx->device); if (guard.status != cudaSuccess) return fail(RS_INITIALIZATION_FAILED, "cudaSetDevice", guard.status);
    cudaStream_t st;
    RS_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking), RS_INITIALIZATION_FAILED);
    *out = new rs_stream_s{ctx, st, true};
    return RS_OK;
}

int rs_stream_wrap(rs_context ctx, void* cuda_stream, rs_stream* out) {
    if (!ctx || !out) return fail(RS_INITIALIZATION_FAILED, "null argument");
    *out = new rs_stream_s{ctx, static_cast<cudaStream_t>(cuda_stream), false};
    return RS_OK;
}

int rs_stream_sync(rs_stream s) {
    if (!s) return fail(RS_CALCULATION_FAILED, "null stream");
    DeviceGuard guard(s->ctx->device); if (guard.status != cudaSuccess) return fail(RS_CALCULATION_FAILED, "cudaSetDevice", guard.status);
    RS_CUDA(cudaStreamSynchronize(s->st), RS_CALCULATION_FAILED);
    return RS_OK;
}

int rs_stream_destroy(rs_stream s) {
    if (!s) return RS_OK;
    if (s->owned) {
        DeviceGuard guard(s->ctx->device);
        cudaStreamDestroy(s->st);
    }
    delete s;
    return RS_OK;
}

int rs_host_alloc(size_t bytes, void** out) {
    if (!out) return fail(RS_HOST_BUFFERS_FAILED, "null argument");
    *out = nullptr;
    RS_CUDA(cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocDefault), RS_HOST_BUFFERS_FAILED);
    return RS_OK;
}
int rs_host_free(void* p) {
    if (p) RS_CUDA(cudaFreeHost(p), RS_CLEANUP_FAILED);
    return RS_OK;
}

uint64_t rs_round_up_keys(uint64_t n) {
    const uint64_t items = 1024;  // _NUM_GROUPS * _NUM_ITEMS_PER_GROUP of the reference
    const uint64_t rest  = n % items;
    return rest ? n + (items - rest) : n;
}

int rs_sorter_create(rs_context ctx, rs_key_type type, uint64_t max_keys, uint32_t flags, rs_sorter* out) {
    if (!ctx || !out) return fail(RS_INITIALIZATION_FAILED, "null argument");
    *out = nullptr;
    KeyInfo ki{};
    if (!key_info(type, &ki)) return fail(RS_INITIALIZATION_FAILED, "unknown key type");
    DeviceGuard guard(ctx->device); if (guard.status != cudaSuccess) return fail(RS_INITIALIZATION_FAILED, "cudaSetDevice", guard.status);
    rs_sorter s = new (std::nothrow) rs_sorter_s{};
    if (!s) return fail(RS_INITIALIZATION_FAILED, "out of host memory");
    s->ctx       = ctx;
    s->type      = type;
    s->flags     = flags;
    s->max_keys  = max_keys;
    s->key_bytes = ki.bytes;
    s->is_signed = ki.is_signed;
    s->finf      = ki.finf;
    s->has_vals  = (flags & RS_WITH_PERMUTATION) != 0;
    s->atomic_rank = ctx->atomic_rank_ok && !(flags & RS_NO_ATOMIC_RANK);
    stat_reset(s->rt.time_histo); stat_reset(s->rt.time_scan); stat_reset(s->rt.time_reorder);
    stat_reset(s->rt.time_paste); stat_reset(s->rt.time_total);
    if (s->has_vals && max_keys > (1ull << 32)) {
        delete s;
        return fail(RS_INITIALIZATION_FAILED, "a uint32 permutation cannot index more than 2^32 keys");
    }
    const uint64_t alloc_keys = std::max<uint64_t>(max_keys, 1);
    // scratch sized for the worst plan: the keys-only tile may differ from the key+value tile
    const int full_passes = static_cast<int>(s->key_bytes);
    size_t need = std::max(make_plan(alloc_keys, s->key_bytes, false, s->atomic_rank, full_passes).total,
                           make_plan(alloc_keys, s->key_bytes, true, s->atomic_rank, full_passes).total);
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&s->scratch), need);
    if (e != cudaSuccess) { delete s; return fail(RS_INITIALIZATION_FAILED, "cudaMalloc scratch", e); }
    s->scratch_bytes = need;
    if (!(flags & RS_NO_KEY_BUFFERS)) {
        for (int i = 0; i < 2; ++i) {
            e = cudaMalloc(&s->keys[i], alloc_keys * s->key_bytes);
            if (e == cudaSuccess && s->has_vals) e = cudaMalloc(reinterpret_cast<void**>(&s->vals[i]), alloc_keys * sizeof(uint32_t));
            if (e != cudaSuccess) {
                rs_sorter_destroy(s);
                return fail(RS_INITIALIZATION_FAILED, "cudaMalloc key/permutation buffers", e);
            }
        }
    }
    *out = s;
    return RS_OK;
}

int rs_sorter_destroy(rs_sorter s) {
    if (!s) return RS_OK;
    DeviceGuard guard(s->ctx->device);
    int rc = RS_OK;
    for (auto& q : s->pending) for (auto& e : q.e) cudaEventDestroy(e);
    for (auto& q : s->free_events) for (auto& e : q.e) cudaEventDestroy(e);
    for (int i = 0; i < 2; ++i) {
        if (s->keys[i] && cudaFree(s->keys[i]) != cudaSuccess) rc = RS_CLEANUP_FAILED;
        if (s->vals[i] && cudaFree(s->vals[i]) != cudaSuccess) rc = RS_CLEANUP_FAILED;
    }
    if (s->scratch && cudaFree(s->scratch) != cudaSuccess) rc = RS_CLEANUP_FAILED;
    delete s;
    return rc;
}

int rs_sorter_pad(rs_sorter s, rs_stream st, uint64_t first_key, uint64_t rounded) {
    if (!s || !st) return fail(RS_CALCULATION_FAILED, "null argument");
    if (!s->keys[0]) return fail(RS_CALCULATION_FAILED, "sorter has no key buffers");
    if (rounded > s->max_keys || first_key > rounded) return fail(RS_RESIZE_FAILED, "pad range exceeds max_keys");
    if (first_key == rounded) return RS_OK;
    DeviceGuard guard(s->ctx->device); if (guard.status != cudaSuccess) return fail(RS_CALCULATION_FAILED, "cudaSetDevice", guard.status);
    const int g = grid_for(rounded - first_key, 256, s->ctx->sm_count);
    // pattern = numeric_limits<T>::max() - 1 (src/RadixSortGPU.cpp:274-276)
    // floating point: max() - 1 rounds to max(), i.e. the pattern just below +infinity
    const int      bits = static_cast<int>(s->key_bytes) * 8;
    const uint64_t all  = bits == 64 ? ~0ull : ((1ull << bits) - 1ull);
    const uint64_t v    = s->finf ? s->finf - 1ull : (s->is_signed ? (all >> 1) : all) - 1ull;
    if (s->key_bytes == 2)
        rs_fill_value_kernel<uint16_t><<<g, 256, 0, st->st>>>(static_cast<uint16_t*>(s->keys[0]), first_key, rounded, static_cast<uint16_t>(v));
    else if (s->key_bytes == 4)
        rs_fill_value_kernel<uint32_t><<<g, 256, 0, st->st>>>(static_cast<uint32_t*>(s->keys[0]), first_key, rounded, static_cast<uint32_t>(v));
    else
        rs_fill_value_kernel<uint64_t><<<g, 256, 0, st->st>>>(static_cast<uint64_t*>(s->keys[0]), first_key, rounded, v);
    RS_CUDA(cudaGetLastError(), RS_CALCULATION_FAILED);
    s->launches++;
    return RS_OK;
}

// Host <-> device copies go out in pieces of 128 MiB (RSB200_COPY_CHUNK_MIB overrides, 0 = one cudaMemcpyAsync).  With an
// upload and a download in flight on two streams one host of this pool moved 35 GB/s each way with whole-GiB copies and
// 44 GB/s with 128 MiB pieces; others move 44-48 GB/s either way (tools/pcie_probe.py, tools/e2e_probe.py,
// profiles/r2_pcie_probe.txt) -- pieces never measured slower.
static cudaError_t copy_in_pieces(void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t st) {
    const char* e = std::getenv("RSB200_COPY_CHUNK_MIB");
    const size_t piece = (e ? static_cast<size_t>(std::atol(e)) : size_t{128}) << 20;
    if (piece == 0 || bytes <= piece) return cudaMemcpyAsync(dst, src, bytes, kind, st);
    for (size_t off = 0; off < bytes; off += piece) {
        const size_t len = std::min(piece, bytes - off);
        cudaError_t err = cudaMemcpyAsync(static_cast<char*>(dst) + off, static_cast<const char*>(src) + off, len, kind, st);
        if (err != cudaSuccess) return err;
    }
    return cudaSuccess;
}

int rs_sorter_upload(rs_sorter s, rs_stream st, const void* h_keys, const uint32_t* h_perm, uint64_t n) {
    if (!s || !st) return fail(RS_DATA_UPLOAD_FAILED, "null argument");
    if (!s->keys[0]) return fail(RS_DATA_UPLOAD_FAILED, "sorter has no key buffers");
    if (n > s->max_keys) return fail(RS_DATA_UPLOAD_FAILED, "n exceeds max_keys");
    if (n && !h_keys) return fail(RS_HOST_BUFFERS_FAILED, "null host key buffer");
    DeviceGuard guard(s->ctx->device); if (guard.status != cudaSuccess) return fail(RS_DATA_UPLOAD_FAILED, "cudaSetDevice", guard.status);
    if (n == 0) return RS_OK;
    RS_CUDA(copy_in_pieces(s->keys[0], h_keys, n * s->key_bytes, cudaMemcpyHostToDevice, st->st), RS_DATA_UPLOAD_FAILED);
    if (s->has_vals) {
        if (h_perm) {
            RS_CUDA(copy_in_pieces(s->vals[0], h_perm, n * sizeof(uint32_t), cudaMemcpyHostToDevice, st->st),
                    RS_DATA_UPLOAD_FAILED);
        } else {
            rs_iota_kernel<<<grid_for(n, 256, s->ctx->sm_count), 256, 0, st->st>>>(s->vals[0], n);
            RS_CUDA(cudaGetLastError(), RS_DATA_UPLOAD_FAILED);
            s->launches++;
        }
    }
    return RS_OK;
}

int rs_sorter_sort_device(rs_sorter s, rs_stream st, const void* d_keys_in, void* d_keys_buf0, void* d_keys_buf1,
                          const uint32_t* d_vals_in, uint32_t* d_vals_buf0, uint32_t* d_vals_buf1, uint64_t n,
                          int begin_bit, int end_bit, int* result_index) {
    if (!s || !st) return fail(RS_CALCULATION_FAILED, "null argument");
    const int key_bits = static_cast<int>(s->key_bytes) * 8;
    if (begin_bit < 0 || end_bit > key_bits || begin_bit >= end_bit) return fail(RS_CALCULATION_FAILED, "bad bit range");
    const int passes = (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS;
    if (result_index) *result_index = (passes - 1) & 1;
    if (!d_keys_in || !d_keys_buf0 || (passes > 1 && !d_keys_buf1)) return fail(RS_CALCULATION_FAILED, "null key buffer");
    const bool has_vals = d_vals_in != nullptr;
    if (has_vals && (!d_vals_buf0 || (passes > 1 && !d_vals_buf1))) return fail(RS_CALCULATION_FAILED, "null value buffer");
    if (has_vals && n > (1ull << 32)) return fail(RS_CALCULATION_FAILED, "uint32 payload limited to 2^32 keys");
    // a pass scatters into tiles other CTAs may still be reading if its source and destination are the same buffer
    if (d_keys_in == d_keys_buf0 || d_keys_buf0 == d_keys_buf1)
        return fail(RS_CALCULATION_FAILED, "key buffers alias: only d_keys_in == d_keys_buf1 is allowed");
    if (has_vals && (d_vals_in == d_vals_buf0 || d_vals_buf0 == d_vals_buf1))
        return fail(RS_CALCULATION_FAILED, "value buffers alias: only d_vals_in == d_vals_buf1 is allowed");
    DeviceGuard guard(s->ctx->device); if (guard.status != cudaSuccess) return fail(RS_CALCULATION_FAILED, "cudaSetDevice", guard.status);
    const bool profile = (s->flags & RS_NO_PROFILE) == 0;
    if (profile) {
        int rc = resolve_pending(s, false);
        if (rc != RS_OK) return rc;
    }
    if (n == 0) return RS_OK;
    if (s->key_bytes == 2)
        return sort_typed<uint16_t>(s, st->st, static_cast<const uint16_t*>(d_keys_in), static_cast<uint16_t*>(d_keys_buf0),
                                    static_cast<uint16_t*>(d_keys_buf1), d_vals_in, d_vals_buf0, d_vals_buf1, n, begin_bit,
                                    end_bit, profile, result_index);
    if (s->key_bytes == 4)
        return sort_typed<uint32_t>(s, st->st, static_cast<const uint32_t*>(d_keys_in), static_cast<uint32_t*>(d_keys_buf0),
                                    static_cast<uint32_t*>(d_keys_buf1), d_vals_in, d_vals_buf0, d_vals_buf1, n, begin_bit,
                                    end_bit, profile, result_index);
    return sort_typed<uint64_t>(s, st->st, static_cast<const uint64_t*>(d_keys_in), static_cast<uint64_t*>(d_keys_buf0),
                                static_cast<uint64_t*>(d_keys_buf1), d_vals_in, d_vals_buf0, d_vals_buf1, n, begin_bit,
                                end_bit, profile, result_index);
}

int rs_sorter_sort_segments_device(rs_sorter s, rs_stream st, int num_segments, const void* const* d_seg_ptrs,
                                   const uint64_t* seg_counts, int first_segment, void* d_gathered, void* d_keys_buf0,
                                   void* d_keys_buf1, int begin_bit, int end_bit, int* result_index) {
    if (!s || !st || !d_seg_ptrs || !seg_counts) return fail(RS_CALCULATION_FAILED, "null argument");
    if (num_segments < 1 || num_segments > MAX_SEGMENTS) return fail(RS_CALCULATION_FAILED, "1..16 segments supported");
    if (first_segment < 0 || first_segment >= num_segments) return fail(RS_CALCULATION_FAILED, "bad first_segment");
    if (s->key_bytes < 4) return fail(RS_CALCULATION_FAILED, "the fused gather + sort entry serves 4- and 8-byte keys");
    const int key_bits = static_cast<int>(s->key_bytes) * 8;
    if (begin_bit < 0 || end_bit > key_bits || begin_bit >= end_bit) return fail(RS_CALCULATION_FAILED, "bad bit range");
    const int passes = (end_bit - begin_bit + RADIX_BITS - 1) / RADIX_BITS;
    if (result_index) *result_index = (passes - 1) & 1;
    if (!d_gathered || !d_keys_buf0 || (passes > 1 && !d_keys_buf1)) return fail(RS_CALCULATION_FAILED, "null key buffer");
    if (d_gathered == d_keys_buf0) return fail(RS_CALCULATION_FAILED, "gathered must not alias buf0");
    DeviceGuard guard(s->ctx->device); if (guard.status != cudaSuccess) return fail(RS_CALCULATION_FAILED, "cudaSetDevice", guard.status);
    const bool profile = (s->flags & RS_NO_PROFILE) == 0;
    if (profile) {
        int rc = resolve_pending(s, false);
        if (rc != RS_OK) return rc;
    }
    if (s->key_bytes == 4)
        return sort_segments_typed<uint32_t>(s, st, num_segments, d_seg_ptrs, seg_counts, first_segment, d_gathered,
                                             d_keys_buf0, d_keys_buf1, begin_bit, end_bit, profile, result_index);
    return sort_segments_typed<uint64_t>(s, st, num_segments, d_seg_ptrs, seg_counts, first_segment, d_gathered,
                                         d_keys_buf0, d_keys_buf1, begin_bit, end_bit, profile, result_index);
}

int rs_merge_runs_device(rs_context ctx, rs_stream st, rs_key_type type, int num_jobs, const void* const* d_a_ptrs,
                         const uint64_t* a_counts, const void* const* d_b_ptrs, const uint64_t* b_counts,
                         void* const* d_out_ptrs, uint64_t* d_splits, uint64_t splits_capacity) {
    if (!ctx || !st || !d_a_ptrs || !a_counts || !d_b_ptrs || !b_counts || !d_out_ptrs || !d_splits)
        return fail(RS_CALCULATION_FAILED, "null argument");
    if (num_jobs < 1 || num_jobs > MERGE_MAX_JOBS) return fail(RS_CALCULATION_FAILED, "1..8 merge jobs per call");
    if (type < RS_KEY_U32 || type > RS_KEY_I64) return fail(RS_CALCULATION_FAILED, "the merge serves 32- and 64-bit integer keys");
    DeviceGuard guard(ctx->device); if (guard.status != cudaSuccess) return fail(RS_CALCULATION_FAILED, "cudaSetDevice", guard.status);
    const bool sg = (type == RS_KEY_I32 || type == RS_KEY_I64);
    if (type == RS_KEY_U32 || type == RS_KEY_I32)
        return merge_runs_typed<uint32_t>(ctx, st->st, sg, num_jobs, d_a_ptrs, a_counts, d_b_ptrs, b_counts, d_out_ptrs, d_splits, splits_capacity);
    return merge_runs_typed<uint64_t>(ctx, st->st, sg, num_jobs, d_a_ptrs, a_counts, d_b_ptrs, b_counts, d_out_ptrs, d_splits, splits_capacity);
}

int rs_device_alloc(rs_context ctx, size_t bytes, void** out) {
    if (!ctx || !out) return fail(RS_INITIALIZATION_FAILED, "null argument");
    *out = nullptr;
    DeviceGuard guard(ctx->device); if (guard.status != cudaSuccess) return fail(RS_INITIALIZATION_FAILED, "cudaSetDevice", guard.status);
    RS_CUDA(cudaMalloc(out, bytes ? bytes : 1), RS_INITIALIZATION_FAILED);
    return RS_OK;
}
int rs_device_free(rs_context ctx, void* p) {
    if (!ctx) return fail(RS_CLEANUP_FAILED, "null argument");
    if (!p) return RS_OK;
    DeviceGuard guard(ctx->device); if (guard.status != cudaSuccess) return fail(RS_CLEANUP_FAILED, "cudaSetDevice", guard.status);
    RS_CUDA(cudaFree(p), RS_CLEANUP_FAILED);
    return RS_OK;
}
int rs_ipc_export(rs_context ctx, const void* d_ptr, unsigned char handle[RS_IPC_HANDLE_BYTES]) {
    if (!ctx || !d_ptr || !handle) return fail(RS_INITIALIZATION_FAILED, "null argument");
    static_assert(sizeof(cudaIpcMemHandle_t) == RS_IPC_HANDLE_BYTES, "IPC handle size");
    DeviceGuard guard(ctx->device); if (guard.status != cudaSuccess) return fail(RS_INITIALIZATION_FAILED, "cudaSetDevice", guard.status);
    cudaIpcMemHandle_t h;
    RS_CUDA(cudaIpcGetMemHandle(&h, const_cast<void*>(d_ptr)), RS_INITIALIZATION_FAILED);
    memcpy(handle, &h, sizeof h);
    return RS_OK;
}
int rs_ipc_open(rs_context ctx, const unsigned char handle[RS_IPC_HANDLE_BYTES], void** out) {
    if (!ctx || !handle || !out) return fail(RS_INITIALIZATION_FAILED, "null argument");
    *out = nullptr;
    DeviceGuard guard(ctx->device); if (guard.status != cudaSuccess) return fail(RS_INITIALIZATION_FAILED, "cudaSetDevice", guard.status);
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    RS_CUDA(cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess), RS_INITIALIZATION_FAILED);
    return RS_OK;
}
int rs_ipc_close(rs_context ctx, void* p) {
    if (!ctx) return fail(RS_CLEANUP_FAILED, "null argument");
    if (!p) return RS_OK;
    DeviceGuard guard(ctx->device); if (guard.status != cudaSuccess) return fail(RS_CLEANUP_FAILED, "cudaSetDevice", guard.status);
    RS_CUDA(cudaIpcCloseMemHandle(p), RS_CLEANUP_FAILED);
    return RS_OK;
}

int rs_sorter_sort(rs_sorter s, rs_stream st, uint64_t n) {
    if (!s || !st) return fail(RS_CALCULATION_FAILED, "null argument");
    if (!s->keys[0]) return fail(RS_CALCULATION_FAILED, "sorter has no key buffers");
    if (n > s->max_keys) return fail(RS_CALCULATION_FAILED, "n exceeds max_keys");
    s->last_n = n;
    // pass p: keys[p&1] -> keys[(p+1)&1]; uploaded data lives in keys[0]
    int idx = 0;
    int rc  = rs_sorter_sort_device(s, st, s->keys[0], s->keys[1], s->keys[0], s->has_vals ? s->vals[0] : nullptr,
                                    s->has_vals ? s->vals[1] : nullptr, s->has_vals ? s->vals[0] : nullptr, n, 0,
                                    static_cast<int>(s->key_bytes) * 8, &idx);
    if (rc != RS_OK) return rc;
    s->result_index = (idx == 0) ? 1 : 0;  // buf0 is keys[1], buf1 is keys[0]
    if (n == 0) s->result_index = 0;
    return RS_OK;
}

int rs_sorter_download(rs_sorter s, rs_stream st, void* h_keys_out, uint32_t* h_perm_out, uint64_t n) {
    if (!s || !st) return fail(RS_DATA_DOWNLOAD_FAILED, "null argument");
    if (!s->keys[0]) return fail(RS_DATA_DOWNLOAD_FAILED, "sorter has no key buffers");
    if (n > s->max_keys) return fail(RS_DATA_DOWNLOAD_FAILED, "n exceeds max_keys");
    DeviceGuard guard(s->ctx->device); if (guard.status != cudaSuccess) return fail(RS_DATA_DOWNLOAD_FAILED, "cudaSetDevice", guard.status);
    const int r = s->result_index;
    if (n && h_keys_out)
        RS_CUDA(cudaMemcpyAsync(h_keys_out, s->keys[r], n * s->key_bytes, cudaMemcpyDeviceToHost, st->st),
                RS_DATA_DOWNLOAD_FAILED);
    if (n && h_perm_out) {
        if (!s->has_vals) return fail(RS_DATA_DOWNLOAD_FAILED, "sorter was created without RS_WITH_PERMUTATION");
        RS_CUDA(cudaMemcpyAsync(h_perm_out, s->vals[r], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st->st),
                RS_DATA_DOWNLOAD_FAILED);
    }
    RS_CUDA(cudaStreamSynchronize(st->st), RS_DATA_DOWNLOAD_FAILED);
    return RS_OK;
}

int rs_sorter_download_async(rs_sorter s, rs_stream st, void* h_keys_out, uint32_t* h_perm_out, uint64_t n) {
    if (!s || !st) return fail(RS_DATA_DOWNLOAD_FAILED, "null argument");
    if (!s->keys[0]) return fail(RS_DATA_DOWNLOAD_FAILED, "sorter has no key buffers");
    if (n > s->max_keys) return fail(RS_DATA_DOWNLOAD_FAILED, "n exceeds max_keys");
    DeviceGuard guard(s->ctx->device); if (guard.status != cudaSuccess) return fail(RS_DATA_DOWNLOAD_FAILED, "cudaSetDevice", guard.status);
    const int r = s->result_index;
    if (n && h_keys_out)
        RS_CUDA(copy_in_pieces(h_keys_out, s->keys[r], n * s->key_bytes, cudaMemcpyDeviceToHost, st->st),
                RS_DATA_DOWNLOAD_FAILED);
    if (n && h_perm_out) {
        if (!s->has_vals) return fail(RS_DATA_DOWNLOAD_FAILED, "sorter was created without RS_WITH_PERMUTATION");
        RS_CUDA(copy_in_pieces(h_perm_out, s->vals[r], n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st->st),
                RS_DATA_DOWNLOAD_FAILED);
    }
    return RS_OK;
}

int rs_sorter_download_histograms(rs_sorter s, rs_stream st, uint64_t* h_hist, size_t capacity, int* passes) {
    if (!s || !st || !h_hist) return fail(RS_DATA_DOWNLOAD_FAILED, "null argument");
    DeviceGuard guard(s->ctx->device); if (guard.status != cudaSuccess) return fail(RS_DATA_DOWNLOAD_FAILED, "cudaSetDevice", guard.status);
    const int P = s->last_passes;
    if (passes) *passes = P;
    if (P == 0) return RS_OK;
    if (capacity < static_cast<size_t>(P) * RADIX) return fail(RS_HOST_BUFFERS_FAILED, "histogram buffer too small");
    RS_CUDA(cudaMemcpyAsync(h_hist, s->scratch + s->last_hist_offset, static_cast<size_t>(P) * RADIX * sizeof(uint64_t),
                            cudaMemcpyDeviceToHost, st->st), RS_DATA_DOWNLOAD_FAILED);
    RS_CUDA(cudaStreamSynchronize(st->st), RS_DATA_DOWNLOAD_FAILED);
    return RS_OK;
}

int rs_sorter_last_skipped_passes(rs_sorter s, rs_stream st, int* skipped) {
    if (!s || !st || !skipped) return fail(RS_DATA_DOWNLOAD_FAILED, "null argument");
    *skipped = 0;
    if (s->last_passes == 0) return RS_OK;
    DeviceGuard guard(s->ctx->device); if (guard.status != cudaSuccess) return fail(RS_DATA_DOWNLOAD_FAILED, "cudaSetDevice", guard.status);
    SortPlanDev plan{};
    RS_CUDA(cudaMemcpyAsync(&plan, s->scratch + s->last_plan_offset, sizeof plan, cudaMemcpyDeviceToHost, st->st),
            RS_DATA_DOWNLOAD_FAILED);
    RS_CUDA(cudaStreamSynchronize(st->st), RS_DATA_DOWNLOAD_FAILED);
    *skipped = static_cast<int>(plan.num_skipped);
    return RS_OK;
}

int rs_sorter_runtimes(rs_sorter s, rs_runtimes* out) {
    if (!s || !out) return fail(RS_CALCULATION_FAILED, "null argument");
    DeviceGuard guard(s->ctx->device); if (guard.status != cudaSuccess) return fail(RS_CALCULATION_FAILED, "cudaSetDevice", guard.status);
    int rc = resolve_pending(s, true);
    if (rc != RS_OK) return rc;
    *out = s->rt;
    return RS_OK;
}

int rs_sorter_reset_runtimes(rs_sorter s) {
    if (!s) return fail(RS_CALCULATION_FAILED, "null argument");
    DeviceGuard guard(s->ctx->device);
    resolve_pending(s, true);
    stat_reset(s->rt.time_histo); stat_reset(s->rt.time_scan); stat_reset(s->rt.time_reorder);
    stat_reset(s->rt.time_paste); stat_reset(s->rt.time_total);
    return RS_OK;
}

int rs_sorter_launch_count(rs_sorter s, uint64_t* out) {
    if (!s || !out) return fail(RS_CALCULATION_FAILED, "null argument");
    *out = s->launches;
    return RS_OK;
}

int rs_dataset_fill_device(rs_context ctx, rs_stream st, rs_key_type type, rs_dataset_kind kind, uint64_t seed,
                           uint64_t global_offset, uint64_t total_n, void* d_out, uint64_t n) {
    if (!ctx || !st || (!d_out && n)) return fail(RS_CALCULATION_FAILED, "null argument");
    if (kind < RS_DATASET_ZEROS || kind > RS_DATASET_RANDOM) return fail(RS_CALCULATION_FAILED, "unknown dataset kind");
    KeyInfo ki{};
    if (!key_info(type, &ki)) return fail(RS_CALCULATION_FAILED, "unknown key type");
    if (n == 0) return RS_OK;
    DeviceGuard guard(ctx->device);
    const int g = grid_for(n, 256, ctx->sm_count);
    for_key_width(ki.bytes, [&](auto u) {
        using U = decltype(u);
        rs_fill_kernel<U><<<g, 256, 0, st->st>>>(static_cast<U*>(d_out), n, kind, ki.is_signed, seed, global_offset, total_n);
    });
    RS_CUDA(cudaGetLastError(), RS_CALCULATION_FAILED);
    return RS_OK;
}

int rs_check_device(rs_context ctx, rs_stream st, rs_key_type type, const void* d_keys, uint64_t n, uint64_t result[4]) {
    if (!ctx || !st || !result || (!d_keys && n)) return fail(RS_CALCULATION_FAILED, "null argument");
    KeyInfo ki{};
    if (!key_info(type, &ki)) return fail(RS_CALCULATION_FAILED, "unknown key type");
    DeviceGuard guard(ctx->device);
    std::lock_guard<std::mutex> lock(ctx->result_mutex);
    unsigned long long* d_res = ctx->d_result;
    cudaMemsetAsync(d_res, 0, 4 * sizeof(unsigned long long), st->st);
    if (n) {
        const int g = grid_for(n, 256, ctx->sm_count);
        for_key_width(ki.bytes, [&](auto u) {
            using U = decltype(u);
            const U sign = ki.is_signed ? static_cast<U>(static_cast<U>(1) << (sizeof(U) * 8 - 1)) : static_cast<U>(0);
            rs_check_kernel<U><<<g, 256, 0, st->st>>>(static_cast<const U*>(d_keys), n, sign, ki.finf, d_res);
        });
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(result, d_res, 4 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st->st);
    if (e != cudaSuccess) return fail(RS_CALCULATION_FAILED, "rs_check_device", e);
    return RS_OK;
}

int rs_check_permutation_device(rs_context ctx, rs_stream st, rs_key_type type, const void* d_keys_in,
                                const void* d_keys_out, const uint32_t* d_perm, uint64_t n, uint64_t result[2]) {
    if (!ctx || !st || !result) return fail(RS_CALCULATION_FAILED, "null argument");
    KeyInfo ki{};
    if (!key_info(type, &ki)) return fail(RS_CALCULATION_FAILED, "unknown key type");
    DeviceGuard guard(ctx->device);
    std::lock_guard<std::mutex> lock(ctx->result_mutex);
    unsigned long long* d_res = ctx->d_result;
    cudaMemsetAsync(d_res, 0, 2 * sizeof(unsigned long long), st->st);
    if (n) {
        const int g = grid_for(n, 256, ctx->sm_count);
        for_key_width(ki.bytes, [&](auto u) {
            using U = decltype(u);
            rs_check_perm_kernel<U><<<g, 256, 0, st->st>>>(static_cast<const U*>(d_keys_in), static_cast<const U*>(d_keys_out),
                                                          d_perm, n, d_res);
        });
    }
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(result, d_res, 2 * sizeof(uint64_t), cudaMemcpyDeviceToHost, st->st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st->st);
    if (e != cudaSuccess) return fail(RS_CALCULATION_FAILED, "rs_check_permutation_device", e);
    return RS_OK;
}

int rs_lower_bound_device(rs_context ctx, rs_stream st, rs_key_type type, const void* d_keys, uint64_t n,
                          const void* d_queries, uint64_t num_queries, uint64_t* d_out) {
    if (!ctx || !st || (num_queries && (!d_queries || !d_out))) return fail(RS_CALCULATION_FAILED, "null argument");
    KeyInfo ki{};
    if (!key_info(type, &ki) || ki.finf) return fail(RS_CALCULATION_FAILED, "the lower bound serves integer keys");
    if (num_queries == 0) return RS_OK;
    DeviceGuard guard(ctx->device);
    const int g   = grid_for(num_queries, 128, ctx->sm_count);
    auto*     out = reinterpret_cast<unsigned long long*>(d_out);
    for_key_width(ki.bytes, [&](auto u) {
        using U = decltype(u);
        const U sign = ki.is_signed ? static_cast<U>(static_cast<U>(1) << (sizeof(U) * 8 - 1)) : static_cast<U>(0);
        rs_lower_bound_kernel<U><<<g, 128, 0, st->st>>>(static_cast<const U*>(d_keys), n, static_cast<const U*>(d_queries),
                                                       num_queries, sign, out);
    });
    RS_CUDA(cudaGetLastError(), RS_CALCULATION_FAILED);
    return RS_OK;
}

}  // extern "C"

#include "rs_dist_api.inl"
