/*
 * radix_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the reference's sort path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library, and only as the checker or the timed CPU baseline.  The product
 * (librsb200.so and everything under radix-sort_b200/) never links, imports or calls it.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks every function below against
 *   (a) the reference's own code compiled from /root/reference (oracle/_ref/libref_radix.so,
 *       recipe oracle/Makefile) on seeded inputs, in the build container, and
 *   (b) golden vectors generated from that library and committed under tests/golden/
 *       (script tests/golden/make_golden.py), which travel to the GPU box.
 *
 * What is restated (paths relative to the reference checkout):
 *   oracle_radix_sort_cpu_*   RadixSortCPU<T>::sort + countSort        src/CRadixSortCPU.h:57-72, 81-122
 *                             (base NUM_BINS = bits/4 counting sort by division, digit count from
 *                             log(|max|)/log(NUM_BINS); input-dependent bugs are reproduced, not fixed)
 *   oracle_opencl_path_*      the OpenCL GPU path: pass loop           src/RadixSortGPU.cpp:310-346
 *                             histogram kernel                          src/kernels/RadixSort.cl:16-71
 *                             scanhistograms x2 + pastehistograms       src/kernels/RadixSort.cl:125-197
 *                             reorder kernel                            src/kernels/RadixSort.cl:74-119
 *                             OFFSET = -numeric_limits<T>::min()        src/RadixSortGPU.cpp:431-442
 *                             with the payload move the reference lost (//#define PERMUT,
 *                             src/RadixSortGPU.cpp:578) added as an option, so the stable
 *                             permutation of the same algorithm is defined.
 *   oracle_dataset_*          Zeros/Range/InvertedRange/RandomDistributed/Random  src/Dataset.h:84-137
 *                             incl. std::seed_seq, std::mt19937 and libstdc++'s
 *                             uniform_int_distribution for full-range parameters.
 */
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* std::seed_seq::generate + std::mt19937 (ISO C++ [rand.util.seedseq], [rand.eng.mers])        */
/* ------------------------------------------------------------------------------------------ */
typedef struct {
    uint32_t mt[624];
    int      idx;
} oracle_mt19937;

static void seed_seq_generate(const uint32_t* v, size_t s, uint32_t* out, size_t n) {
    if (n == 0) return;
    for (size_t i = 0; i < n; ++i) out[i] = 0x8b8b8b8bu;
    const size_t t = (n >= 623) ? 11 : (n >= 68) ? 7 : (n >= 39) ? 5 : (n >= 7) ? 3 : (n - 1) / 2;
    const size_t p = (n - t) / 2;
    const size_t q = p + t;
    const size_t m = (s + 1 > n) ? s + 1 : n;
    for (size_t k = 0; k < m; ++k) {
        uint32_t x  = out[k % n] ^ out[(k + p) % n] ^ out[(k + n - 1) % n];
        uint32_t r1 = 1664525u * (x ^ (x >> 27));
        uint32_t r2 = r1;
        if (k == 0)
            r2 += (uint32_t)s;
        else if (k <= s)
            r2 += (uint32_t)(k % n) + v[k - 1];
        else
            r2 += (uint32_t)(k % n);
        out[(k + p) % n] += r1;
        out[(k + q) % n] += r2;
        out[k % n] = r2;
    }
    for (size_t k = m; k < m + n; ++k) {
        uint32_t x  = out[k % n] + out[(k + p) % n] + out[(k + n - 1) % n];
        uint32_t r3 = 1566083941u * (x ^ (x >> 27));
        uint32_t r4 = r3 - (uint32_t)(k % n);
        out[(k + p) % n] ^= r3;
        out[(k + q) % n] ^= r4;
        out[k % n] = r4;
    }
}

static void mt_seed_seq(oracle_mt19937* g, const uint32_t* v, size_t s) {
    seed_seq_generate(v, s, g->mt, 624);
    int zero = (g->mt[0] & 0x80000000u) == 0;
    for (int i = 1; zero && i < 624; ++i)
        if (g->mt[i] != 0) zero = 0;
    if (zero) g->mt[0] = 0x80000000u;
    g->idx = 624;
}

static uint32_t mt_next(oracle_mt19937* g) {
    if (g->idx >= 624) {
        for (int i = 0; i < 624; ++i) {
            uint32_t y = (g->mt[i] & 0x80000000u) | (g->mt[(i + 1) % 624] & 0x7fffffffu);
            g->mt[i]   = g->mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
        }
        g->idx = 0;
    }
    uint32_t y = g->mt[g->idx++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

/* ------------------------------------------------------------------------------------------ */
/* per-type stamps                                                                              */
/* ------------------------------------------------------------------------------------------ */
enum { ORACLE_ZEROS = 0, ORACLE_RANGE = 1, ORACLE_INVERTED = 2, ORACLE_UNIFORM = 3, ORACLE_RANDOM = 4 };

#define VPROC_GROUPS 16u   /* _NUM_GROUPS           src/Parameters.h:20 */
#define VPROC_ITEMS 64u    /* _NUM_ITEMS_PER_GROUP  src/Parameters.h:18 */
#define CL_BITS 4u         /* _NUM_BITS_PER_RADIX   src/Parameters.h:26 */
#define CL_RADIX 16u
#define HISTOSPLIT 512u    /* _NUM_HISTOSPLIT       src/Parameters.h:24 */

/* exclusive scan of the 16384-entry table the way the three scan launches do it:
 * 512 blocks of 32 entries scanned locally (block total -> globsum), globsum scanned, then pasted back. */
static void scan_histograms_like_cl(uint32_t* histo, uint32_t* globsum) {
    const uint32_t total = CL_RADIX * VPROC_GROUPS * VPROC_ITEMS;
    const uint32_t block = total / HISTOSPLIT;
    for (uint32_t b = 0; b < HISTOSPLIT; ++b) {
        uint32_t run = 0;
        for (uint32_t i = 0; i < block; ++i) {
            uint32_t c           = histo[b * block + i];
            histo[b * block + i] = run;
            run += c;
        }
        globsum[b] = run;
    }
    uint32_t run = 0;
    for (uint32_t b = 0; b < HISTOSPLIT; ++b) {
        uint32_t c = globsum[b];
        globsum[b] = run;
        run += c;
    }
    for (uint32_t b = 0; b < HISTOSPLIT; ++b)
        for (uint32_t i = 0; i < block; ++i) histo[b * block + i] += globsum[b];
}

#define STAMP(SUF, T, U, TMIN_BITS, TMAX_BITS, IS_SIGNED)                                                        \
    /* RadixSortCPU<T>::sort, src/CRadixSortCPU.h:57-72 with countSort :81-122 */                                \
    static void count_sort_##SUF(T* arr, T* output, size_t n, uint64_t exp, unsigned num_bins) {                  \
        size_t count[16];                                                                                        \
        memset(count, 0, sizeof count);                                                                          \
        for (size_t i = 0; i < n; ++i) {                                                                         \
            const U ev = (U)((U)arr[i] - (U)(TMIN_BITS)); /* inputData[i] - offset, :96 */                        \
            count[(ev / exp) % num_bins]++;                                                                      \
        }                                                                                                        \
        for (unsigned i = 1; i < num_bins; ++i) count[i] += count[i - 1];                                        \
        for (int64_t i = (int64_t)n - 1; i >= 0; --i) { /* backward, stable, :108-113 */                         \
            const U      ev = (U)((U)arr[i] - (U)(TMIN_BITS));                                                   \
            const size_t c  = (size_t)((ev / exp) % num_bins);                                                   \
            output[count[c] - 1] = arr[i];                                                                       \
            count[c]--;                                                                                          \
        }                                                                                                        \
        memcpy(arr, output, n * sizeof(T));                                                                      \
    }                                                                                                            \
    int oracle_radix_sort_cpu_##SUF(T* arr, size_t n) {                                                          \
        if (n == 0) return 0; /* the reference dereferences max_element of an empty span; callers never do */    \
        const unsigned num_bins = (unsigned)(sizeof(T) * 8u / CL_BITS); /* NUM_BINS, :35 */                      \
        T max_elem = arr[0];                                                                                     \
        for (size_t i = 1; i < n; ++i)                                                                           \
            if (arr[i] > max_elem) max_elem = arr[i];                                                            \
        uint64_t num_digits = 1;                                                                                 \
        if (max_elem) {                                                                                          \
            /* abs(): unsigned magnitude (:20-25); negated in the unsigned domain so INT_MIN is defined */       \
            const U      mag = (IS_SIGNED && (max_elem < (T)0)) ? (U)(0u - (U)max_elem) : (U)max_elem;           \
            const double a   = (IS_SIGNED && sizeof(T) == 4) ? (double)(int32_t)mag : (double)mag;               \
            /* all keys == numeric_limits<T>::min() (signed): the reference negates INT_MIN and takes the */   \
            /* log of a negative number -- undefined behaviour there, refused here */                           \
            if (!(a > 0.0) || (IS_SIGNED && (U)max_elem == (U)(TMIN_BITS))) return -4;                           \
            num_digits       = (uint64_t)ceil(log(a) / log((double)num_bins));                                   \
        }                                                                                                        \
        T* output = (T*)malloc(n * sizeof(T));                                                                   \
        if (!output) return -1;                                                                                  \
        for (uint64_t e = 0; e < num_digits; ++e)                                                                \
            count_sort_##SUF(arr, output, n, (uint64_t)pow((double)num_bins, (double)e), num_bins);              \
        free(output);                                                                                            \
        return 0;                                                                                                \
    }                                                                                                            \
    /* The OpenCL path on the host: n must be a multiple of 1024 (RadixSortGPU::Resize). */                      \
    /* keys are sorted in place; perm (may be NULL) is moved with the keys; histo_out (16384) and */             \
    /* globsum_out (512) receive the LAST pass's tables, like the read-backs in CopyDataFromDevice. */           \
    int oracle_opencl_path_##SUF(T* keys, uint32_t* perm, size_t n, uint32_t* histo_out, uint32_t* globsum_out) { \
        const uint32_t nproc = VPROC_GROUPS * VPROC_ITEMS;                                                       \
        if (n % nproc) return -2;                                                                                \
        const unsigned passes = (unsigned)(sizeof(T) * 8u / CL_BITS); /* _NUM_PASSES, src/Parameters.h:48 */     \
        const U        OFFSET = (U)(0u - (U)(TMIN_BITS));             /* BuildPreamble, RadixSortGPU.cpp:436 */  \
        T*        out   = (T*)malloc((n ? n : 1) * sizeof(T));                                                   \
        uint32_t* pout  = perm ? (uint32_t*)malloc((n ? n : 1) * sizeof(uint32_t)) : NULL;                       \
        uint32_t* histo = (uint32_t*)malloc(CL_RADIX * nproc * sizeof(uint32_t));                                \
        uint32_t  globsum[HISTOSPLIT];                                                                           \
        if (!out || !histo || (perm && !pout)) { free(out); free(pout); free(histo); return -1; }                \
        T*        in  = keys;                                                                                    \
        uint32_t* pin = perm;                                                                                    \
        const size_t sub = n / nproc; /* sublist_size, RadixSort.cl:39 */                                        \
        for (unsigned pass = 0; pass < passes; ++pass) {                                                         \
            memset(histo, 0, CL_RADIX * nproc * sizeof(uint32_t));                                               \
            for (uint32_t ig = 0; ig < nproc; ++ig) { /* kernel histogram */                                     \
                const uint32_t gr = ig / VPROC_ITEMS, it = ig % VPROC_ITEMS;                                     \
                for (size_t j = 0; j < sub; ++j) {                                                               \
                    const U key      = (U)((U)in[ig * sub + j] + OFFSET);                                        \
                    const U shortkey = (U)((key >> (pass * CL_BITS)) & (CL_RADIX - 1));                          \
                    histo[VPROC_ITEMS * (shortkey * VPROC_GROUPS + gr) + it]++;                                  \
                }                                                                                                \
            }                                                                                                    \
            scan_histograms_like_cl(histo, globsum); /* scanhistograms x2 + pastehistograms */                   \
            if (pass == passes - 1) {                                                                            \
                if (histo_out) memcpy(histo_out, histo, CL_RADIX * nproc * sizeof(uint32_t));                    \
                if (globsum_out) memcpy(globsum_out, globsum, sizeof globsum);                                   \
            }                                                                                                    \
            for (uint32_t ig = 0; ig < nproc; ++ig) { /* kernel reorder */                                       \
                const uint32_t gr = ig / VPROC_ITEMS, it = ig % VPROC_ITEMS;                                     \
                for (size_t j = 0; j < sub; ++j) {                                                               \
                    const size_t k        = ig * sub + j;                                                        \
                    const U      key      = (U)((U)in[k] + OFFSET);                                              \
                    const U      shortkey = (U)((key >> (pass * CL_BITS)) & (CL_RADIX - 1));                     \
                    uint32_t*    slot     = &histo[VPROC_ITEMS * (shortkey * VPROC_GROUPS + gr) + it];           \
                    out[*slot]            = (T)(U)(key - OFFSET);                                                \
                    if (perm) pout[*slot] = pin[k]; /* the PERMUT move the reference dropped */                  \
                    (*slot)++;                                                                                   \
                }                                                                                                \
            }                                                                                                    \
            { T* t = in; in = out; out = t; } /* std::swap of buffer names, RadixSortGPU.cpp:263-266 */          \
            if (perm) { uint32_t* t = pin; pin = pout; pout = t; }                                               \
        }                                                                                                        \
        if (in != keys) memcpy(keys, in, n * sizeof(T));                                                         \
        if (perm && pin != perm) memcpy(perm, pin, n * sizeof(uint32_t));                                        \
        free(in == keys ? out : in);                                                                             \
        if (perm) free(pin == perm ? pout : pin);                                                                \
        free(histo);                                                                                             \
        return 0;                                                                                                \
    }                                                                                                            \
    /* Dataset generators, src/Dataset.h:84-137.  `seed` replaces the clock of RandomDistributed. */             \
    int oracle_dataset_##SUF(int kind, T* out, size_t n, uint64_t seed) {                                        \
        oracle_mt19937 g;                                                                                        \
        switch (kind) {                                                                                          \
            case ORACLE_ZEROS:                                                                                   \
                for (size_t i = 0; i < n; ++i) out[i] = 0;                                                       \
                return 0;                                                                                        \
            case ORACLE_RANGE: /* iota from numeric_limits<T>::min() */                                          \
                for (size_t i = 0; i < n; ++i) out[i] = (T)(U)((U)(TMIN_BITS) + (U)i);                           \
                return 0;                                                                                        \
            case ORACLE_INVERTED:                                                                                \
                for (size_t i = 0; i < n; ++i) out[i] = (T)(U)((U)(TMIN_BITS) + (U)(n - 1 - i));                 \
                return 0;                                                                                        \
            case ORACLE_RANDOM: { /* seed_seq over the characters of "Random Test Seed", raw outputs */          \
                const char* s = "Random Test Seed";                                                              \
                uint32_t    v[16];                                                                               \
                for (int i = 0; i < 16; ++i) v[i] = (uint32_t)(unsigned char)s[i];                               \
                mt_seed_seq(&g, v, 16);                                                                          \
                for (size_t i = 0; i < n; ++i) out[i] = (T)mt_next(&g); /* T(uint_fast32_t) conversion */        \
                return 0;                                                                                        \
            }                                                                                                    \
            case ORACLE_UNIFORM: { /* seed_seq{lo32, hi32}; libstdc++ uniform_int_distribution(min, max) */      \
                uint32_t v[2] = {(uint32_t)(seed & 0xffffffffu), (uint32_t)(seed >> 32)};                        \
                mt_seed_seq(&g, v, 2);                                                                           \
                for (size_t i = 0; i < n; ++i) {                                                                 \
                    U r;                                                                                         \
                    if (sizeof(T) == 4) {                                                                        \
                        r = (U)mt_next(&g);                                                                      \
                    } else {                                                                                     \
                        const uint64_t hi = mt_next(&g);                                                         \
                        const uint64_t lo = mt_next(&g);                                                         \
                        r                 = (U)((hi << 32) + lo);                                                \
                    }                                                                                            \
                    out[i] = (T)(U)(r + (U)(TMIN_BITS)); /* ret + param.a() */                                   \
                }                                                                                                \
                if (n) { /* "Ensure that min and max are in the input array", :105-106 */                        \
                    out[0]     = (T)(U)(TMAX_BITS);                                                              \
                    out[n - 1] = (T)(U)(TMIN_BITS);                                                              \
                }                                                                                                \
                return 0;                                                                                        \
            }                                                                                                    \
        }                                                                                                        \
        return -3;                                                                                               \
    }

STAMP(u32, uint32_t, uint32_t, 0u, 0xffffffffu, 0)
STAMP(i32, int32_t, uint32_t, 0x80000000u, 0x7fffffffu, 1)
STAMP(u64, uint64_t, uint64_t, 0ull, 0xffffffffffffffffull, 0)
STAMP(i64, int64_t, uint64_t, 0x8000000000000000ull, 0x7fffffffffffffffull, 1)

int oracle_abi_version(void) { return 1; }
