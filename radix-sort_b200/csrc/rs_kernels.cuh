// rs_kernels.cuh -- sm_100a device code of the B200-native LSD radix sort.
//
// Replaces the four OpenCL kernels of the reference (src/kernels/RadixSort.cl:16-197):
//   histogram        (:16-71)   -> rs_histogram_kernel : ONE read of all keys builds the 256-bin
//                                  histograms of every 8-bit digit (all passes) at once
//   scanhistograms x2 (:125-181)
//   pastehistograms  (:185-197) -> rs_scan_kernel (tiny) + decoupled look-back inside the pass kernel
//   reorder          (:74-119)  -> rs_onesweep_kernel : one stable scatter pass per digit; a tile's
//                                  global offsets come from a chained scan over per-tile status words
// The key transform `key + OFFSET` (:51,109,114 with OFFSET from src/RadixSortGPU.cpp:436) is a
// sign-bit flip; it is folded into the digit extraction of the pass that holds the sign bit.
//
// Data layout in HBM (all owned by rs_sorter, see rs_api.cu):
//   keys   : T[n] ping-pong pair           values : uint32[n] ping-pong pair (optional)
//   hist   : uint64[passes][256]           digit counts of the whole input (order-independent)
//   base   : uint64[passes][chunks][256]   bit 63 = ready | global output offset of (chunk, digit);
//                                          chunk 0 from rs_scan_kernel, chunk c+1 from the last tile of chunk c
//   status : uint32[passes][tiles][256]    2 flag bits | 30-bit count  (look-back words)
//   ctr    : uint32[passes]                dynamic tile-id dispensers
#pragma once
#include <cstdint>
#include <type_traits>
#include <cuda_runtime.h>

// Bounds checks of our own (compute-sanitizer is not available on the GPU pool this was developed on): a build with
// -DRS_DEBUG_CHECKS traps on any out-of-range tile position, output index or look-back sum; tools/debug_checks.sh runs
// the parity workloads against that build (profiles/r2_debug_checks.txt).
#ifdef RS_DEBUG_CHECKS
#include <cstdio>
#define RS_DCHECK(cond)                                                                                   \
    do {                                                                                                  \
        if (!(cond)) {                                                                                    \
            printf("RS_DCHECK failed: %s  (%s:%d, block %d thread %d)\n", #cond, __FILE__, __LINE__,      \
                   static_cast<int>(blockIdx.x), static_cast<int>(threadIdx.x));                          \
            __trap();                                                                                     \
        }                                                                                                 \
    } while (0)
#else
#define RS_DCHECK(cond) ((void)0)
#endif

namespace rsb {

constexpr int      RADIX_BITS        = 8;
constexpr int      RADIX             = 1 << RADIX_BITS;
constexpr uint32_t STATUS_LOCAL      = 1u << 30;   // tile aggregate available
constexpr uint32_t STATUS_INCLUSIVE  = 2u << 30;   // inclusive prefix (within chunk) available
constexpr uint32_t STATUS_VALUE_MASK = (1u << 30) - 1u;
constexpr uint64_t CHUNK_MAX_KEYS    = (1ull << 30) - 1ull;  // a look-back chain never sums past this
constexpr int      MAX_PASSES        = 8;
#ifndef RS_LB_WINDOW
#define RS_LB_WINDOW 8
#endif
constexpr int      LB_WINDOW         = RS_LB_WINDOW;  // look-back status words in flight per thread
constexpr unsigned long long BASE_READY = 1ull << 63;      // ready flag of a chunk-base word

// Per-sort plan written by rs_scan_kernel: a pass whose digit is the same for every key is the identity
// (stable sort on a constant digit) and is skipped on the device, without a host round trip.  The
// ping-pong assignment of the buffers stays the static one (pass p: buf[(p-1)&1] -> buf[p&1], pass 0
// reads the caller's input), so kernels keep compile-time-constant pointers; a skipped pass either does
// nothing (skipped passes cancel in pairs) or copies its source to its destination to keep the parity.
enum : uint32_t { PASS_SORT = 0, PASS_COPY = 1, PASS_NOTHING = 2 };
struct SortPlanDev {
    uint32_t action[MAX_PASSES];
    uint32_t num_skipped;   // passes whose digit was constant
    uint32_t num_copies;    // of those, how many had to copy
};

struct DigitSpec {
    int      shift;  // first bit of the digit
    uint32_t mask;   // (1<<bits)-1
    uint32_t flip;   // xor applied to the digit (sign bit of signed integer keys), else 0
    unsigned long long finf;  // floating-point keys: bit pattern of +infinity in the key's format; 0 = integer keys
};

struct PassTable {
    int       num_passes;
    DigitSpec d[MAX_PASSES];
};

// Floating-point keys (additive, the reference excludes floats, doc/doc.tex:64): the usual
// order-preserving map of IEEE bit patterns to unsigned integers -- negative values have all bits flipped, the
// others get the sign bit set -- applied on the fly inside the digit extraction, so keys are stored and moved
// unmodified and float sorts cost no extra pass.  Every NaN (either sign, any payload) maps to the largest value:
// NaNs end up last, in input order among themselves, as numpy and torch place them.  -0.0 orders before +0.0.
template <typename KeyT>
__device__ __forceinline__ KeyT float_order_bits(KeyT k, unsigned long long finf) {
    using W = typename std::conditional<sizeof(KeyT) == 8, unsigned long long, uint32_t>::type;
    constexpr W ALL  = static_cast<W>(~static_cast<KeyT>(0));           // all key bits
    constexpr W SIGN = static_cast<W>(1) << (sizeof(KeyT) * 8 - 1);
    const W     w    = static_cast<W>(k);
    if ((w & (ALL ^ SIGN)) > static_cast<W>(finf)) return static_cast<KeyT>(ALL);
    return static_cast<KeyT>((w & SIGN) ? (w ^ ALL) : (w | SIGN));
}

// FK (float keys) is a template parameter of every kernel: integer sorts pay nothing for the transform.
template <bool FK, typename KeyT>
__device__ __forceinline__ uint32_t digit_of(KeyT k, const DigitSpec& s) {
    if constexpr (FK) k = float_order_bits(k, s.finf);
    return (static_cast<uint32_t>(k >> s.shift) & s.mask) ^ s.flip;
}

// ------------------------------------------------------------------------------------------------
// small PTX helpers (mbarrier + bulk async copy = the TMA 1-D path, SASS: UBLKCP / SYNCS)
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_load_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ uint32_t ld_status(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_status(uint32_t* p, uint32_t v) {
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_base(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_base(unsigned long long* p, unsigned long long v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// Staging-buffer swizzle: sorted inputs (Range, Inverted Range) scatter with a power-of-two stride
// (one run per digit, equal lengths), which would put all 32 lanes of a row on one bank.
__device__ __forceinline__ uint32_t swz(uint32_t pos) { return pos ^ ((pos >> 5) & 31u); }
__device__ __forceinline__ uint32_t lanemask_lt() {
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}

// ------------------------------------------------------------------------------------------------
// Upfront multi-digit histogram.  One coalesced 128-bit read of every key; all `NPASS` 256-bin
// histograms are accumulated in shared memory and flushed once per block.
//
// ncu on the first version (one table per block, `atomicAdd(&sh[pass][digit])`): LSU data pipe 96 %
// busy, 68 % of the shared wavefronts were bank conflicts of the random-digit atomics, DRAM at 30 %.
// So the table is replicated per LANE: sh[pass][digit][col], col = lane % COLS.  With COLS = 32 the
// bank of an access is the lane, i.e. every ATOMS is one wavefront whatever the digits are -- this
// also removes the same-address serialisation of degenerate inputs (Zeros, Range) without any
// warp-uniformity test.  32 KB per pass: one 1024-thread block per SM; for more than 4 passes
// (64-bit keys) COLS = 16 (two lanes share a column).
// Algorithmic traffic: sizeof(KeyT) bytes read per key.
// ------------------------------------------------------------------------------------------------
constexpr int HIST_THREADS = 1024;
constexpr int HIST_VECS    = 4;  // 16-byte vectors in flight per thread

template <int NPASS>
struct HistCfg {
    static constexpr int    COLS  = NPASS <= 4 ? 32 : 16;
    // more than four passes (8-byte keys): the split layout of the byte-digit path is [4][256][32] whatever NPASS is
    static constexpr size_t bytes = NPASS <= 4 ? sizeof(uint32_t) * NPASS * RADIX * COLS : sizeof(uint32_t) * 4 * RADIX * 32;
};

template <typename KeyT, int NPASS, bool FK>
__global__ void __launch_bounds__(HIST_THREADS, 1)
rs_histogram_kernel(const KeyT* __restrict__ keys, uint64_t n, unsigned long long* __restrict__ hist /* [NPASS][256] */,
                    PassTable pt, int byte_digits, KeyT key_flip) {
    constexpr int KV   = 16 / sizeof(KeyT);
    constexpr int COLS = HistCfg<NPASS>::COLS;
    extern __shared__ __align__(16) uint32_t hist_sh[];  // [NPASS][256][COLS]  (SPLIT: [4][256][32], see below)
    const int      tid     = threadIdx.x;
    const int      lane    = tid & 31;
    const uint32_t col     = lane & (COLS - 1);
    const bool     aligned = (reinterpret_cast<uintptr_t>(keys) & 15u) == 0;

    for (int i = tid; i < static_cast<int>(HistCfg<NPASS>::bytes / sizeof(uint32_t)); i += HIST_THREADS) hist_sh[i] = 0;
    __syncthreads();
    const uint64_t nvec = aligned ? n / KV : 0;
    const uint4*   cv   = reinterpret_cast<const uint4*>(keys);

    // SPLIT (8-byte keys, more than four byte digits): 32 columns per digit row do not fit for eight passes, and with 16
    // two lanes share a column (every atomic two wavefronts: 0.62 ms for 2 GiB, LSU-bound).  Instead the two half-warps
    // read the SAME keys; lanes 0-15 count bytes 0-3 (the low word), lanes 16-31 bytes 4-7 (the high word), into
    // [4][256][32] where columns 0-15 of row (q, d) belong to pass q and columns 16-31 to pass q + 4.  The bank of an
    // access is the lane again: one wavefront per atomic, half as many atomics per lane.
    constexpr bool CAN_SPLIT = sizeof(KeyT) == 8 && NPASS > 4;
    if (CAN_SPLIT && byte_digits) {
        if constexpr (CAN_SPLIT) {
            const int      half     = lane >> 4;
            const uint32_t col_bits = static_cast<uint32_t>(lane) << 2;
            const int      my_npass = half ? NPASS - 4 : 4;       // byte digits this lane counts
            auto count_word = [&](KeyT k) {
                if constexpr (FK) k = float_order_bits(k, pt.d[0].finf); else k ^= key_flip;
                const uint32_t w = half ? static_cast<uint32_t>(static_cast<unsigned long long>(k) >> 32) : static_cast<uint32_t>(k);
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q < my_npass) {
                        uint32_t off = (8 * q >= 7) ? (w >> (8 * q - 7)) : (w << (7 - 8 * q));
                        off = (off & (0xffu << 7)) | col_bits;
                        atomicAdd(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(hist_sh) + q * RADIX * 32 * sizeof(uint32_t) + off), 1u);
                    }
                }
            };
            // both half-warps walk the same vectors: 512 logical threads per block
            const uint32_t ltid   = static_cast<uint32_t>(tid >> 5) * 16 + (lane & 15);
            const uint64_t stride = static_cast<uint64_t>(gridDim.x) * (HIST_THREADS / 2);
            const uint64_t first  = static_cast<uint64_t>(blockIdx.x) * (HIST_THREADS / 2) + ltid;
            for (uint64_t base = first; base < nvec; base += stride * HIST_VECS) {
                uint4 v[HIST_VECS];
                bool  act[HIST_VECS];
#pragma unroll
                for (int u = 0; u < HIST_VECS; ++u) {
                    const uint64_t idx = base + u * stride;
                    act[u]             = idx < nvec;
                    v[u]               = act[u] ? __ldg(cv + idx) : make_uint4(0, 0, 0, 0);
                }
#pragma unroll
                for (int u = 0; u < HIST_VECS; ++u) {
                    if (!act[u]) continue;
                    count_word(static_cast<KeyT>((static_cast<uint64_t>(v[u].y) << 32) | v[u].x));
                    count_word(static_cast<KeyT>((static_cast<uint64_t>(v[u].w) << 32) | v[u].z));
                }
            }
            for (uint64_t i = nvec * KV + first; i < n; i += stride) count_word(keys[i]);
            __syncthreads();
            // fold: row (q, d) -> pass q from columns 0-15, pass q + 4 from columns 16-31 (rotated by the lane: 32 banks)
            for (int row = tid; row < 4 * RADIX; row += HIST_THREADS) {
                unsigned long long lo = 0, hi = 0;
#pragma unroll
                for (int j = 0; j < 16; ++j) {
                    const int c = (j + lane) & 15;
                    lo += hist_sh[row * 32 + c];
                    hi += hist_sh[row * 32 + 16 + c];
                }
                if (lo) atomicAdd(&hist[row], lo);
                if (hi && row + 4 * RADIX < NPASS * RADIX) atomicAdd(&hist[row + 4 * RADIX], hi);
            }
        }
        return;
    }

    // byte_digits (the usual whole-key sort: digit p is byte p of the key, key_flip = the sign bit of signed keys): the
    // key is put into unsigned order ONCE, and the byte offset of a counter -- digit * COLS * 4 + col * 4 -- costs one
    // shift and one LOP3 per digit (the generic path spends ~4.4 ALU instructions per atomic and was issue-bound: 71 % of
    // the issue slots, profiles/r1d_ncu_full_summary.txt).
    const uint32_t col_bits = col << 2;
    auto count_key = [&](KeyT k) {
        if (byte_digits) {
            constexpr int LG = COLS == 32 ? 7 : 6;   // log2 of the bytes per digit row
            if constexpr (FK) k = float_order_bits(k, pt.d[0].finf); else k ^= key_flip;
#pragma unroll
            for (int p = 0; p < NPASS; ++p) {
                uint32_t off;
                if (8 * p >= LG) off = static_cast<uint32_t>(k >> (8 * p - LG)); else off = static_cast<uint32_t>(k) << (LG - 8 * p);
                off = (off & (0xffu << LG)) | col_bits;
                atomicAdd(reinterpret_cast<uint32_t*>(reinterpret_cast<char*>(hist_sh) + p * RADIX * COLS * sizeof(uint32_t) + off), 1u);
            }
        } else {
#pragma unroll
            for (int p = 0; p < NPASS; ++p)
                atomicAdd(&hist_sh[(p * RADIX + digit_of<FK>(k, pt.d[p])) * COLS + col], 1u);
        }
    };

    // a block never counts more than 2^32 keys per column (the launcher keeps blocks below 2^31 keys)
    const uint64_t stride = static_cast<uint64_t>(gridDim.x) * HIST_THREADS;
    const uint64_t first  = static_cast<uint64_t>(blockIdx.x) * HIST_THREADS + tid;
    for (uint64_t base = first; base < nvec; base += stride * HIST_VECS) {
        uint4 v[HIST_VECS];
        bool  act[HIST_VECS];
#pragma unroll
        for (int u = 0; u < HIST_VECS; ++u) {
            const uint64_t idx = base + u * stride;
            act[u]             = idx < nvec;
            v[u]               = act[u] ? __ldg(cv + idx) : make_uint4(0, 0, 0, 0);
        }
#pragma unroll
        for (int u = 0; u < HIST_VECS; ++u) {
            if (!act[u]) continue;
            if constexpr (sizeof(KeyT) == 2) {
                const uint32_t w4[4] = {v[u].x, v[u].y, v[u].z, v[u].w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    count_key(static_cast<KeyT>(w4[j] & 0xffffu));
                    count_key(static_cast<KeyT>(w4[j] >> 16));
                }
            } else if constexpr (sizeof(KeyT) == 4) {
                count_key(static_cast<KeyT>(v[u].x));
                count_key(static_cast<KeyT>(v[u].y));
                count_key(static_cast<KeyT>(v[u].z));
                count_key(static_cast<KeyT>(v[u].w));
            } else {
                count_key(static_cast<KeyT>((static_cast<uint64_t>(v[u].y) << 32) | v[u].x));
                count_key(static_cast<KeyT>((static_cast<uint64_t>(v[u].w) << 32) | v[u].z));
            }
        }
    }
    // scalar tail (and everything when the base pointer is not 16-byte aligned)
    for (uint64_t i = nvec * KV + first; i < n; i += stride) count_key(keys[i]);
    __syncthreads();
    // fold the columns (rotated by the lane so a warp's 32 rows hit 32 banks) and flush
    for (int row = tid; row < NPASS * RADIX; row += HIST_THREADS) {
        unsigned long long sum = 0;
#pragma unroll
        for (int j = 0; j < COLS; ++j) sum += hist_sh[row * COLS + ((j + lane) & (COLS - 1))];
        if (sum) atomicAdd(&hist[row], sum);
    }
}

// ------------------------------------------------------------------------------------------------
// Fused exchange + histogram (multi-GPU): the input of a sort is the concatenation of up to
// MAX_SEGMENTS device segments that may live in PEER memory (other GPUs of the NVSwitch box, mapped
// through CUDA IPC).  One kernel pulls every segment over NVLink with plain coalesced loads, writes it
// into the local `gathered` buffer and accumulates the digit histograms of the coming sort on the fly:
// the all-to-all and the histogram pass of the final local sort are one kernel, the received keys are
// never re-read for counting.  Segments are visited starting at `first_segment` (= this rank), so at
// any time every source GPU is read by one reader.  Same per-lane shared-memory tables as above.
// ------------------------------------------------------------------------------------------------
constexpr int MAX_SEGMENTS = 16;
constexpr int GATHER_UNROLL = 8;

template <typename KeyT>
struct SegmentTable {
    int                num_segments;
    int                first_segment;
    const KeyT*        ptr[MAX_SEGMENTS];
    unsigned long long start[MAX_SEGMENTS + 1];  // prefix of the segment lengths = position in `gathered`
};

template <typename KeyT, int NPASS>
__global__ void __launch_bounds__(HIST_THREADS, 1)
rs_gather_histogram_kernel(const SegmentTable<KeyT> seg, KeyT* __restrict__ gathered,
                           unsigned long long* __restrict__ hist /* [NPASS][256] */, PassTable pt) {
    constexpr int COLS = HistCfg<NPASS>::COLS;
    extern __shared__ __align__(16) uint32_t hist_sh[];  // [NPASS][256][COLS]
    const int      tid  = threadIdx.x;
    const int      lane = tid & 31;
    const uint32_t col  = lane & (COLS - 1);
    for (int i = tid; i < NPASS * RADIX * COLS; i += HIST_THREADS) hist_sh[i] = 0;
    __syncthreads();

    const unsigned long long stride = static_cast<unsigned long long>(gridDim.x) * HIST_THREADS;
    const unsigned long long first  = static_cast<unsigned long long>(blockIdx.x) * HIST_THREADS + tid;
    for (int k = 0; k < seg.num_segments; ++k) {
        int sidx = seg.first_segment + k;
        if (sidx >= seg.num_segments) sidx -= seg.num_segments;
        // select without dynamic indexing of the by-value table (keeps it in the constant bank)
        const KeyT*        src = nullptr;
        unsigned long long beg = 0, end = 0;
#pragma unroll
        for (int j = 0; j < MAX_SEGMENTS; ++j)
            if (j == sidx) { src = seg.ptr[j]; beg = seg.start[j]; end = seg.start[j + 1]; }
        // A warp load that straddles 128-byte lines runs at about half speed over NVLink (tools/peerbw.cu): index
        // the segment from the start of its first line (`skew` elements early) so that lane 0 sits on a line boundary.
        const unsigned long long skew = (reinterpret_cast<uintptr_t>(src) / sizeof(KeyT)) & (128 / sizeof(KeyT) - 1);
        const unsigned long long cnt  = end - beg + skew;
        const KeyT* const        from = src - skew;
        KeyT* const              dst  = gathered + beg - skew;
        for (unsigned long long base = first; base < cnt; base += stride * GATHER_UNROLL) {
            KeyT v[GATHER_UNROLL];
#pragma unroll
            for (int u = 0; u < GATHER_UNROLL; ++u) {
                const unsigned long long i = base + u * stride;
                if (i >= skew && i < cnt) v[u] = from[i];  // NVLink (or local) read
            }
#pragma unroll
            for (int u = 0; u < GATHER_UNROLL; ++u) {
                const unsigned long long i = base + u * stride;
                if (i >= skew && i < cnt) {
                    dst[i] = v[u];
#pragma unroll
                    for (int p = 0; p < NPASS; ++p)
                        atomicAdd(&hist_sh[(p * RADIX + digit_of<false>(v[u], pt.d[p])) * COLS + col], 1u);
                }
            }
        }
    }
    __syncthreads();
    for (int row = tid; row < NPASS * RADIX; row += HIST_THREADS) {
        unsigned long long sum = 0;
#pragma unroll
        for (int j = 0; j < COLS; ++j) sum += hist_sh[row * COLS + ((j + lane) & (COLS - 1))];
        if (sum) atomicAdd(&hist[row], sum);
    }
}

// ------------------------------------------------------------------------------------------------
// Digit-offset scan + pass plan: ONE block.  For every pass p: base[p][0][d] = READY | #keys with
// digit < d (chunks > 0 get their base from the last tile of the previous chunk and stay 0 = not
// ready here), and skip[p] = "one digit holds all n keys".  Thread 0 then assigns the ping-pong
// buffers to the passes that do run.  Replaces scanhistograms x2 + pastehistograms of the reference
// for the GLOBAL part; the per-tile part is the look-back inside rs_onesweep_kernel.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(RADIX)
rs_scan_kernel(const unsigned long long* __restrict__ hist, unsigned long long* __restrict__ base, int num_chunks,
               int num_passes, unsigned long long n, int allow_skip, int input_is_buf1, SortPlanDev* __restrict__ plan) {
    __shared__ unsigned long long wsum[RADIX / 32];
    __shared__ uint32_t           skip_sh[MAX_PASSES];
    const int d    = threadIdx.x;
    const int lane = d & 31, warp = d >> 5;
    for (int p = 0; p < num_passes; ++p) {
        const unsigned long long total = hist[static_cast<size_t>(p) * RADIX + d];
        unsigned long long       incl  = total;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        if (lane == 31) wsum[warp] = incl;
        const int constant_digit = __syncthreads_or(total == n);  // also orders the wsum writes
        unsigned long long woff = 0;
        for (int w = 0; w < warp; ++w) woff += wsum[w];
        base[static_cast<size_t>(p) * num_chunks * RADIX + d] = BASE_READY | (woff + incl - total);
        if (d == 0) skip_sh[p] = (allow_skip && constant_digit) ? 1u : 0u;
        __syncthreads();
    }
    if (d == 0) {
        // Runs of skipped passes cancel in pairs: after (p, p+1) the data is where pass p+2 expects it.  Pass 0 is
        // special unless the caller's input IS buf1 (the double-buffer call): its source is a third buffer.
        uint32_t skipped = 0, copies = 0;
        int p = 0;
        while (p < num_passes) {
            if (!skip_sh[p]) { plan->action[p] = PASS_SORT; ++p; continue; }
            int run_end = p;
            while (run_end < num_passes && skip_sh[run_end]) ++run_end;
            skipped += static_cast<uint32_t>(run_end - p);
            if (p == 0 && !input_is_buf1) { plan->action[0] = PASS_COPY; ++copies; ++p; }
            const int len = run_end - p;
            for (int j = 0; j < len; ++j) plan->action[p + j] = PASS_NOTHING;
            if (len & 1) { plan->action[run_end - 1] = PASS_COPY; ++copies; }
            p = run_end;
        }
        plan->num_skipped = skipped;
        plan->num_copies  = copies;
    }
}

// ------------------------------------------------------------------------------------------------
// Onesweep pass: persistent CTAs take tile ids from an atomic dispenser, prefetch the NEXT tile's
// keys into shared memory with a 1-D bulk async copy (TMA) while ranking the current one, rank the
// digits group by group, resolve the tile's global digit offsets with a decoupled look-back over
// status words, stage the keys (and values) in shared memory in ranked order and write them out in
// coalesced per-digit runs.  Stable.
//
// The kernel is bound by shared-memory wavefronts (LSU data pipe ~89 % busy, ncu), so ranking is
// chosen by how many data-dependent shared accesses it needs per key.  Measured on B200
// (tools/ubench.cu, tools/atomorder.cu, ncu): MATCH.ANY costs ~1.4 cycles per DISTINCT value per SM
// (45 for random 8-bit digits), eight ballots cost 17 and make the kernel issue-bound.  Two modes:
//
//  ATOMIC_RANK = true  (default when the device passes the self-test of rs_context_create):
//      count first (atomicAdd(&count[warp][digit], 1), no return value), publish the tile's digit
//      counts, turn the counts into tile positions, then position = atomicAdd(&pos[warp][digit], 1)
//      while scattering.  Two accesses per key, no rank registers.  The position is stable only
//      because the hardware serves the lanes of one warp instruction that hit the same shared word
//      in ascending lane order -- not documented, so it is TESTED: tools/atomorder.cu saw 0
//      violations in 6e9 warp rows on B200 (all-equal .. all-distinct digits, single-bank and
//      conflict-free layouts, partial warps), and every context re-runs that test before it enables
//      the mode; the parity tests compare keys AND the stable permutation bit-exactly.
//  ATOMIC_RANK = false (fallback, RS_NO_ATOMIC_RANK): a shared-memory lane bitmap per
//      (half-warp, digit), {count:16 | bitmap:16} in one 32-bit word: every lane ORs its lane bit
//      in, reads the word back, the lowest lane of each match group clears the bitmap and bumps
//      the count.  Three accesses per key; relies on nothing beyond the PTX memory model.
//
// A group (warp / half-warp) owns ROW*KPT consecutive keys of the tile and walks them ROW at a time,
// so the order of processing is the index order.  Rows whose digits are all equal (Zeros, the high
// digits of Range) would serialise on one word and take a warp-uniform shortcut.
// Algorithmic traffic: sizeof(KeyT) read + sizeof(KeyT) written per key (+4+4 with values).
// ------------------------------------------------------------------------------------------------
template <typename KeyT>
struct OnesweepArgs {
    const KeyT*               keys_in;
    KeyT*                     keys_out;
    const uint32_t*           vals_in;
    uint32_t*                 vals_out;
    const SortPlanDev*        plan;
    int                       pass;
    uint64_t                  n;
    unsigned long long*       base;    // [chunks][256] for this pass: READY | offset (chunks > 0 zeroed)
    uint32_t*                 status;  // [tiles][256] for this pass, zeroed
    uint32_t*                 counter; // tile-id dispenser for this pass, zeroed
    uint32_t                  num_tiles;
    uint32_t                  tiles_per_chunk;
    uint32_t                  num_chunks;
    DigitSpec                 digit;
    int                       use_tma;
};

template <typename KeyT, bool HAS_VALS, int THREADS, int KPT, bool ATOMIC_RANK>
struct OnesweepSmem {
    static constexpr int    TILE   = THREADS * KPT;
    static constexpr int    NWARPS = THREADS / 32;
    static constexpr int    ROW    = ATOMIC_RANK ? 32 : 16;           // keys a ranking group handles per step
    static constexpr int    GROUPS = THREADS / ROW;                   // warps or half-warps
    static constexpr int    PARTS  = THREADS / RADIX;                 // threads per digit in the scan phase
    static constexpr size_t stage_bytes = sizeof(KeyT) * TILE;
    static constexpr size_t off_stage0  = 0;
    static constexpr size_t off_stage1  = off_stage0 + stage_bytes;
    static constexpr size_t off_vstage  = off_stage1 + stage_bytes;
    static constexpr size_t off_tab     = off_vstage + (HAS_VALS ? sizeof(uint32_t) * TILE : 0);
    static constexpr size_t off_kptr    = off_tab + sizeof(uint32_t) * GROUPS * RADIX;
    static constexpr size_t off_vptr    = off_kptr + sizeof(void*) * RADIX;
    // the partial digit totals of the scan (written after barrier (A), read before (B2)) share the bytes of the run
    // pointers (written after (C), read until (E)) whenever they fit
    static constexpr bool   PTOT_IN_KPTR = sizeof(uint32_t) * PARTS <= sizeof(void*);
    static constexpr size_t off_ptot_own = off_vptr + (HAS_VALS ? sizeof(void*) * RADIX : 0);
    static constexpr size_t off_ptot    = PTOT_IN_KPTR ? off_kptr : off_ptot_own;
    static constexpr size_t off_wsum    = off_ptot_own + (PTOT_IN_KPTR ? 0 : sizeof(uint32_t) * PARTS * RADIX);
    static constexpr size_t off_mbar    = off_wsum + sizeof(uint32_t) * 16;
    static constexpr size_t off_tile    = off_mbar + sizeof(uint64_t) * 2;
    static constexpr size_t off_cbase   = off_tile + sizeof(uint32_t) * 4;   // [digit] (chunk + 1) << 48 | cached chunk base
    static constexpr size_t total       = off_cbase + sizeof(unsigned long long) * RADIX;
};

__device__ __forceinline__ uint32_t atoms_add_return(uint32_t* p, uint32_t v) {
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(smem_u32(p)), "r"(v) : "memory");
    return old;
}

template <typename KeyT, bool HAS_VALS, int THREADS, int KPT, int MIN_BLOCKS, bool ATOMIC_RANK, bool FK>
__global__ void __launch_bounds__(THREADS, MIN_BLOCKS)
rs_onesweep_kernel(const OnesweepArgs<KeyT> a) {
    using L = OnesweepSmem<KeyT, HAS_VALS, THREADS, KPT, ATOMIC_RANK>;
    constexpr int TILE   = L::TILE;
    constexpr int ROW    = L::ROW;
    constexpr int GROUPS = L::GROUPS;
    constexpr int PARTS  = L::PARTS;
    constexpr int GPP    = GROUPS / PARTS;              // group tables per scan thread
    constexpr int CSHIFT = ATOMIC_RANK ? 0 : 16;        // where the count / position lives in a table word
    static_assert(THREADS % RADIX == 0 && GROUPS % PARTS == 0, "THREADS must be a multiple of 256");
    static_assert((sizeof(KeyT) * TILE) % 128 == 0, "stage must keep 128B alignment");
    static_assert(TILE <= 0xffff, "16-bit count / position fields");
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
    uint32_t* const    tab     = reinterpret_cast<uint32_t*>(smem_raw + L::off_tab);   // [group][digit]
    // [digit] -> run pointer minus the run's tile-local start.  (32-bit element offsets instead of pointers were tried in
    // round 2: the lookup is one wavefront cheaper, the address arithmetic and the >2^32-key branch cost more -- 0.679 ->
    // 0.724 ms per pass, profiles/README.md.)
    KeyT** const       kptr    = reinterpret_cast<KeyT**>(smem_raw + L::off_kptr);
    uint32_t** const   vptr    = reinterpret_cast<uint32_t**>(smem_raw + L::off_vptr);
    uint32_t* const    ptot    = reinterpret_cast<uint32_t*>(smem_raw + L::off_ptot);  // [part][digit] partial digit totals
    uint32_t* const    wsum    = reinterpret_cast<uint32_t*>(smem_raw + L::off_wsum);
    uint64_t* const    mbar    = reinterpret_cast<uint64_t*>(smem_raw + L::off_mbar);
    volatile uint32_t* sh_tile = reinterpret_cast<volatile uint32_t*>(smem_raw + L::off_tile);
    unsigned long long* const cbase = reinterpret_cast<unsigned long long*>(smem_raw + L::off_cbase);

    const int       tid   = threadIdx.x;
    const int       lane  = tid & 31;
    const int       lig   = tid & (ROW - 1);               // lane within its ranking group
    const int       grp   = tid / ROW;                     // this lane's ranking group
    const uint32_t  bitg  = 1u << lig;
    const uint32_t  ltg   = bitg - 1u;
    const DigitSpec ds    = a.digit;
    const uint32_t  full_tiles = static_cast<uint32_t>(a.n / TILE);  // tiles [0, full_tiles) are complete
    uint32_t* const gt    = tab + grp * RADIX;
    // scan-phase role: digit `sd`, group tables [sp*GPP, sp*GPP + GPP)
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
    if (sp == 0) cbase[sd] = 0ull;   // tag 0: nothing cached
    __syncthreads();

    uint32_t phase_bits = 0;  // bit s = parity to wait for on mbar[s]
    int      s          = 0;

    auto process_tile = [&](auto full_c, const uint32_t t) {
        constexpr bool FULL     = decltype(full_c)::value;
        const bool     via_tma  = FULL && a.use_tma;
        const uint64_t tile_base = static_cast<uint64_t>(t) * TILE;
        const uint32_t valid    = FULL ? TILE : static_cast<uint32_t>(a.n - tile_base);
        const uint32_t my_first = grp * (ROW * KPT) + lig;  // group-striped: item r sits at my_first + ROW*r
        KeyT* const    stg      = stage0 + s * TILE;

        // ---- keys (and values) into registers -------------------------------------------------
        KeyT key[KPT];
        if (via_tma) {
            mbar_wait(&mbar[s], (phase_bits >> s) & 1u);
            phase_bits ^= (1u << s);
#pragma unroll
            for (int r = 0; r < KPT; ++r) key[r] = stg[my_first + ROW * r];
        } else {
#pragma unroll
            for (int r = 0; r < KPT; ++r) {
                const uint32_t i = my_first + ROW * r;
                key[r] = (FULL || i < valid) ? a.keys_in[tile_base + i] : static_cast<KeyT>(~static_cast<KeyT>(0));
            }
        }
        uint32_t val[HAS_VALS ? KPT : 1];
        if constexpr (HAS_VALS) {
#pragma unroll
            for (int r = 0; r < KPT; ++r) {
                const uint32_t i = my_first + ROW * r;
                val[r]           = (FULL || i < valid) ? a.vals_in[tile_base + i] : 0u;
            }
        }

        // ---- ATOMIC_RANK: count now (so the tile's digit counts can be published early), rank while scattering;
        //      bitmap mode: rank now -------------------------------------------------------------------------------
        uint32_t rk[ATOMIC_RANK ? 1 : (KPT + 1) / 2];  // bitmap mode: two 16-bit ranks per register (rank < ROW*KPT)
        uint32_t uniform_rows = 0;                       // atomic mode: bit r = all 32 digits of row r are equal
#pragma unroll
        for (int r = 0; r < KPT; ++r) {
            uint32_t d = digit_of<FK>(key[r], ds);
            if (!FULL && (my_first + ROW * r) >= valid) d = RADIX - 1;  // padding ranks last
            if constexpr (ATOMIC_RANK) {
                int all_same;
                __match_all_sync(0xffffffffu, d, &all_same);
                if (all_same) {  // 32 equal digits would serialise on one word: one lane counts them all
                    uniform_rows |= 1u << r;
                    if (lane == 0) atomicAdd(&gt[d], 32u);
                } else {
                    atomicAdd(&gt[d], 1u);
                }
            } else {
                uint32_t my_rank;
                // warp-uniform test "both half-warps hold 16 equal digits" (all masks stay compile-time constants)
                const uint32_t d_first = __shfl_sync(0xffffffffu, d, lane & 16);
                if (__all_sync(0xffffffffu, d == d_first)) {
                    const uint32_t e = gt[d];
                    __syncwarp();
                    if (lig == 0) gt[d] = e + (16u << 16);
                    my_rank = (e >> 16) + lig;
                } else {
                    atomicOr(&gt[d], bitg);
                    __syncwarp();
                    const uint32_t e = gt[d];
                    __syncwarp();
                    const uint32_t below = __popc(e & ltg);
                    if (below == 0) gt[d] = (e & 0xffff0000u) + (static_cast<uint32_t>(__popc(e & 0xffffu)) << 16);
                    my_rank = (e >> 16) + below;
                }
                __syncwarp();
                if (r & 1) rk[r >> 1] |= my_rank << 16; else rk[r >> 1] = my_rank;
            }
        }
        __syncthreads();  // (A) all group counts final; every key of this stage is in registers

        // ---- scan: every thread owns digit sd of GPP group tables ----------------------------------
        const uint32_t chunk          = t / a.tiles_per_chunk;
        const bool     first_in_chunk = (t % a.tiles_per_chunk) == 0;
        uint32_t       pre[(GPP + 1) / 2];  // exclusive prefixes over this thread's tables, two 16-bit values per register
        uint32_t       part_sum = 0;
#pragma unroll
        for (int j = 0; j < GPP; ++j) {
            if (j & 1) pre[j >> 1] |= part_sum << 16; else pre[j >> 1] = part_sum;
            part_sum += tab[(sp * GPP + j) * RADIX + sd] >> CSHIFT;
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
                (local_base + before_part + ((j & 1) ? (pre[j >> 1] >> 16) : (pre[j >> 1] & 0xffffu))) << CSHIFT;
        __syncthreads();  // (C) tab[g][d] >> CSHIFT = tile-local position of group g's first key with digit d

        // ---- scatter into shared memory in ranked order -----------------------------------------
#pragma unroll
        for (int r = 0; r < KPT; ++r) {
            uint32_t d = digit_of<FK>(key[r], ds);
            if (!FULL && (my_first + ROW * r) >= valid) d = RADIX - 1;
            uint32_t pos;
            if constexpr (ATOMIC_RANK) {
                // gt[d] is the running tile position of (this warp, digit d).  Lanes that hit the same word are served
                // in ascending lane order (self-tested, see the header comment), so the returned value is the stable
                // position of this key.
                if (uniform_rows & (1u << r)) {
                    uint32_t c = 0;
                    if (lane == 0) c = atoms_add_return(&gt[d], 32u);
                    pos = __shfl_sync(0xffffffffu, c, 0) + lane;
                } else {
                    pos = atoms_add_return(&gt[d], 1u);
                }
            } else {
                pos = (gt[d] >> CSHIFT) + ((r & 1) ? (rk[r >> 1] >> 16) : (rk[r >> 1] & 0xffffu));
            }
            RS_DCHECK(pos < static_cast<uint32_t>(TILE));
            stg[swz(pos)] = key[r];
            if constexpr (HAS_VALS) vstage[swz(pos)] = val[r];
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
            // polled once per chunk and digit, then cached in shared memory (the thread that owns the digit is the only
            // one that touches its slot): one L2 round trip less between the look-back and barrier (D), -1 % per pass
            unsigned long long  cb = cbase[sd];
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
            // run pointer minus the run's tile-local start: element at tile position p goes to kptr[d][p]
            const long long run0 = static_cast<long long>(cb + excl) - static_cast<long long>(local_base);
            kptr[sd] = a.keys_out + run0;
            if constexpr (HAS_VALS) vptr[sd] = a.vals_out + run0;
        }
        __syncthreads();  // (D) staged tile + run pointers visible; nobody reads the tables any more

        // ---- re-zero the tables; coalesced per-digit runs out to global memory -------------------
#pragma unroll
        for (int j = 0; j < GPP; ++j) tab[(sp * GPP + j) * RADIX + sd] = 0u;
#pragma unroll
        for (int i = 0; i < KPT; ++i) {
            const uint32_t p = i * THREADS + tid;
            if (FULL || p < valid) {
                const KeyT     k = stg[swz(p)];
                const uint32_t d = digit_of<FK>(k, ds);
                RS_DCHECK(kptr[d] + p >= a.keys_out && kptr[d] + p < a.keys_out + a.n);
                kptr[d][p]       = k;
                if constexpr (HAS_VALS) vptr[d][p] = vstage[swz(p)];
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

// ------------------------------------------------------------------------------------------------
// Small inputs (n <= SMALL_CAP): the whole sort in ONE launch of ONE CTA, all passes in shared memory
// (the general path costs a memset and 2 + passes launches, ~45 us, whatever n is).  Same ranking as
// the atomic mode of rs_onesweep_kernel (count, scan, position atomic while scattering), same
// skipping of constant-digit passes, same digit-histogram and plan read-backs.  Only used when the
// context passed the atomic-order self-test.
// ------------------------------------------------------------------------------------------------
constexpr int SMALL_THREADS = 512;
constexpr int SMALL_KPT     = 8;
constexpr int SMALL_CAP     = SMALL_THREADS * SMALL_KPT;  // 4096 keys

template <typename KeyT, bool HAS_VALS>
struct SmallSmem {
    static constexpr size_t off_keys = 0;
    static constexpr size_t off_vals = off_keys + 2 * sizeof(KeyT) * SMALL_CAP;
    static constexpr size_t off_tab  = off_vals + (HAS_VALS ? 2 * sizeof(uint32_t) * SMALL_CAP : 0);
    static constexpr size_t off_wsum = off_tab + sizeof(uint32_t) * (SMALL_THREADS / 32) * RADIX;
    static constexpr size_t total    = off_wsum + sizeof(uint32_t) * 16;
};

template <typename KeyT, bool HAS_VALS, bool FK>
__global__ void __launch_bounds__(SMALL_THREADS, 1)
rs_small_sort_kernel(const KeyT* in, KeyT* out, const uint32_t* vin, uint32_t* vout,  // `out` may alias `in`
                     uint32_t n, PassTable pt, unsigned long long* __restrict__ hist,
                     SortPlanDev* __restrict__ plan, int allow_skip) {
    using L = SmallSmem<KeyT, HAS_VALS>;
    constexpr int NW = SMALL_THREADS / 32;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    KeyT* const     kbuf = reinterpret_cast<KeyT*>(smem_raw + L::off_keys);      // [2][SMALL_CAP]
    uint32_t* const vbuf = reinterpret_cast<uint32_t*>(smem_raw + L::off_vals);  // [2][SMALL_CAP]
    uint32_t* const tab  = reinterpret_cast<uint32_t*>(smem_raw + L::off_tab);   // [warp][digit]
    uint32_t* const wsum = reinterpret_cast<uint32_t*>(smem_raw + L::off_wsum);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    uint32_t* const gt = tab + warp * RADIX;

    for (uint32_t i = tid; i < SMALL_CAP; i += SMALL_THREADS) {
        kbuf[i] = (i < n) ? in[i] : static_cast<KeyT>(~static_cast<KeyT>(0));
        if constexpr (HAS_VALS) vbuf[i] = (i < n) ? vin[i] : 0u;
    }
    int      cur     = 0;
    uint32_t skipped = 0;
    const uint32_t my_first = warp * (32 * SMALL_KPT) + lane;
    for (int p = 0; p < pt.num_passes; ++p) {
        const DigitSpec ds = pt.d[p];
        if (tid < RADIX)
#pragma unroll
            for (int w = 0; w < NW; ++w) tab[w * RADIX + tid] = 0u;
        __syncthreads();  // table clean; previous scatter (or the load) visible
        KeyT     key[SMALL_KPT];
        uint32_t val[HAS_VALS ? SMALL_KPT : 1];
        uint32_t uniform_rows = 0;
#pragma unroll
        for (int r = 0; r < SMALL_KPT; ++r) {
            const uint32_t i = my_first + 32 * r;
            key[r]           = kbuf[cur * SMALL_CAP + i];
            if constexpr (HAS_VALS) val[r] = vbuf[cur * SMALL_CAP + i];
            const uint32_t d = (i < n) ? digit_of<FK>(key[r], ds) : RADIX - 1;  // padding sits past n and ranks last
            int all_same;
            __match_all_sync(0xffffffffu, d, &all_same);
            if (all_same) {
                uniform_rows |= 1u << r;
                if (lane == 0) atomicAdd(&gt[d], 32u);
            } else {
                atomicAdd(&gt[d], 1u);
            }
        }
        __syncthreads();
        uint32_t count = 0, incl = 0;
        uint32_t pre[NW / 2];
        if (tid < RADIX) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                if (w & 1) pre[w >> 1] |= count << 16; else pre[w >> 1] = count;
                count += tab[w * RADIX + tid];
            }
            if (tid == RADIX - 1) count -= (SMALL_CAP - n);
            hist[p * RADIX + tid] = count;
            incl = count;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const uint32_t x = __shfl_up_sync(0xffffffffu, incl, o);
                if (lane >= o) incl += x;
            }
            if (lane == 31) wsum[warp] = incl;
        }
        const int constant_digit = __syncthreads_or(tid < RADIX && count == n);  // also orders the wsum writes
        if (allow_skip && constant_digit) {  // the pass is the identity
            ++skipped;
            if (tid == 0) plan->action[p] = PASS_NOTHING;
            continue;
        }
        if (tid == 0) plan->action[p] = PASS_SORT;
        if (tid < RADIX) {
            uint32_t woff = 0;
#pragma unroll
            for (int w = 0; w < RADIX / 32; ++w) woff += (w < warp) ? wsum[w] : 0u;
            const uint32_t local_base = woff + incl - count;
#pragma unroll
            for (int w = 0; w < NW; ++w)
                tab[w * RADIX + tid] = local_base + ((w & 1) ? (pre[w >> 1] >> 16) : (pre[w >> 1] & 0xffffu));
        }
        __syncthreads();
#pragma unroll
        for (int r = 0; r < SMALL_KPT; ++r) {
            const uint32_t i = my_first + 32 * r;
            const uint32_t d = (i < n) ? digit_of<FK>(key[r], ds) : RADIX - 1;
            uint32_t pos;
            if (uniform_rows & (1u << r)) {
                uint32_t c = 0;
                if (lane == 0) c = atoms_add_return(&gt[d], 32u);
                pos = __shfl_sync(0xffffffffu, c, 0) + lane;
            } else {
                pos = atoms_add_return(&gt[d], 1u);  // same-word lanes are served in lane order (self-tested)
            }
            RS_DCHECK(pos < static_cast<uint32_t>(SMALL_CAP));
            kbuf[(cur ^ 1) * SMALL_CAP + pos] = key[r];
            if constexpr (HAS_VALS) vbuf[(cur ^ 1) * SMALL_CAP + pos] = val[r];
        }
        cur ^= 1;
        __syncthreads();  // every warp is done with its table before the next pass re-zeroes all of them
    }
    __syncthreads();
    for (uint32_t i = tid; i < n; i += SMALL_THREADS) {
        out[i] = kbuf[cur * SMALL_CAP + i];
        if constexpr (HAS_VALS) vout[i] = vbuf[cur * SMALL_CAP + i];
    }
    if (tid == 0) {
        plan->num_skipped = skipped;
        plan->num_copies  = 0;
    }
}

// ------------------------------------------------------------------------------------------------
// Self-test behind ATOMIC_RANK: for random digit patterns (1 .. 256 distinct values, single-bank and
// spread layouts, partial warps) every lane does old = atomicAdd(&cnt[d], 1) and checks
// old == count_before + (number of lower lanes with the same d).  Any violation disables the mode.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
rs_atomic_order_selftest_kernel(unsigned long long* violations, int iters) {
    __shared__ uint32_t cnt[8][512];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = lane; i < 512; i += 32) cnt[warp][i] = 0;
    __syncwarp();
    uint32_t           v   = (tid * 7919u + blockIdx.x * 104729u) * 2654435761u + 12345u;
    unsigned long long bad = 0;
    const uint32_t     lt  = (1u << lane) - 1u;
    const int          distincts[8] = {1, 2, 3, 5, 8, 16, 31, 256};
    for (int it = 0; it < iters; ++it) {
        v = v * 1664525u + 1013904223u;
        const int      distinct = distincts[(it >> 3) & 7];
        const int      stride   = (it & 1) ? 32 : ((it & 2) ? 33 : 1);
        const uint32_t d        = (((v >> 9) % distinct) * stride) & 511;
        const bool     active   = (it & 4) ? (((v >> 20) & 3) != 0) : true;
        const uint32_t amask    = __ballot_sync(0xffffffffu, active);
        const uint32_t before   = cnt[warp][d];
        __syncwarp();
        uint32_t old = 0;
        if (active) old = atoms_add_return(&cnt[warp][d], 1u);
        __syncwarp();
        const uint32_t m = __match_any_sync(0xffffffffu, d) & amask;
        if (active && old != before + __popc(m & lt)) ++bad;
        __syncwarp();
    }
    if (bad) atomicAdd(violations, bad);
}

}  // namespace rsb
