// ref_shim.cpp -- TEST INFRASTRUCTURE ONLY.  extern "C" doors onto the UNMODIFIED reference code.
//
// Compiled by oracle/Makefile into oracle/_ref/libref_radix.so with -I/root/reference/src
// -I/root/reference; the reference's header-only RadixSortCPU (src/CRadixSortCPU.h), its
// dataset generators (src/Dataset.h) and Parameters.h are used where they lie -- nothing is
// copied into this repository.  Used (1) to pin the C restatement in radix_oracle.c,
// (2) to generate tests/golden/, (3) as the timed CPU baseline of bench.py
// (cpu_baseline.kind == "reference").  The product never loads it.
#include "CRadixSortCPU.h"
#include "Dataset.h"

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <span>

namespace {

template <typename T>
int radix_sort_cpu(T* data, size_t n) {
    if (n == 0) return 0;
    std::span<T> s(data, n);
    RadixSortCPU<T>::sort(s);  // src/CRadixSortCPU.h:57
    return 0;
}

template <typename T>
int dataset(int kind, T* out, size_t n) {
    std::unique_ptr<Dataset<T>> ds;
    switch (kind) {
        case 0: ds = std::make_unique<Zeros<T>>(n); break;
        case 1: ds = std::make_unique<Range<T>>(n); break;
        case 2: ds = std::make_unique<InvertedRange<T>>(n); break;
        case 3: ds = std::make_unique<RandomDistributed<T>>(n); break;  // clock-seeded in the reference
        case 4: ds = std::make_unique<Random<T>>(n); break;
        default: return -3;
    }
    std::memcpy(out, ds->dataset.data(), n * sizeof(T));
    return 0;
}

// What RandomDistributed<T> (src/Dataset.h:92-107) computes when its clock reading is `seed`:
// the same standard-library objects, fed a caller-chosen seed so the result is reproducible.
template <typename T>
int uniform_seeded(T* out, size_t n, uint64_t seed) {
    std::seed_seq seq({static_cast<uint32_t>(seed & 0xFFFFFFFF), static_cast<uint32_t>(seed >> 32)});
    std::mt19937 gen(seq);
    std::uniform_int_distribution<T> dist(std::numeric_limits<T>::min(), std::numeric_limits<T>::max());
    for (size_t i = 0; i < n; ++i) out[i] = dist(gen);
    if (n) {
        out[0]     = std::numeric_limits<T>::max();
        out[n - 1] = std::numeric_limits<T>::min();
    }
    return 0;
}

template <typename T>
double time_radix_sort_cpu_ms(const T* data, size_t n, int iterations, T* scratch) {
    // like CRadixSortTask::ComputeCPU (src/CRadixSortTask.cpp:203-221): copy + sort, averaged
    const auto t0 = std::chrono::high_resolution_clock::now();
    for (int i = 0; i < iterations; ++i) {
        std::copy(data, data + n, scratch);
        std::span<T> s(scratch, n);
        RadixSortCPU<T>::sort(s);
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double, std::milli>(t1 - t0).count() / iterations;
}

template <typename T>
double time_std_sort_ms(const T* data, size_t n, int iterations, T* scratch) {
    const auto t0 = std::chrono::high_resolution_clock::now();
    for (int i = 0; i < iterations; ++i) {
        std::copy(data, data + n, scratch);
        std::sort(scratch, scratch + n);
    }
    const auto t1 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration<double, std::milli>(t1 - t0).count() / iterations;
}

}  // namespace

#define DOORS(SUF, T)                                                                                        \
    extern "C" int ref_radix_sort_cpu_##SUF(T* d, size_t n) { return radix_sort_cpu<T>(d, n); }               \
    extern "C" int ref_std_sort_##SUF(T* d, size_t n) { std::sort(d, d + n); return 0; }                      \
    extern "C" int ref_dataset_##SUF(int kind, T* out, size_t n) { return dataset<T>(kind, out, n); }         \
    extern "C" int ref_uniform_seeded_##SUF(T* out, size_t n, uint64_t seed) { return uniform_seeded<T>(out, n, seed); } \
    extern "C" double ref_time_radix_sort_cpu_ms_##SUF(const T* d, size_t n, int it, T* scratch) {            \
        return time_radix_sort_cpu_ms<T>(d, n, it, scratch);                                                  \
    }                                                                                                         \
    extern "C" double ref_time_std_sort_ms_##SUF(const T* d, size_t n, int it, T* scratch) {                  \
        return time_std_sort_ms<T>(d, n, it, scratch);                                                        \
    }

DOORS(u32, uint32_t)
DOORS(i32, int32_t)
DOORS(u64, uint64_t)
DOORS(i64, int64_t)

extern "C" unsigned ref_num_bins_32(void) { return RadixSortCPU<uint32_t>::NUM_BINS; }
extern "C" unsigned ref_num_bins_64(void) { return RadixSortCPU<uint64_t>::NUM_BINS; }
extern "C" unsigned ref_num_performance_iterations(void) {
    return AlgorithmParameters<uint32_t>::_NUM_PERFORMANCE_ITERATIONS;
}
