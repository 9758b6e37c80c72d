// Parameters.h -- compile-time parameters of the sort (reference src/Parameters.h:9-60), B200 values.
//
// Names are the reference's so code written against AlgorithmParameters<T> keeps compiling; values differ
// where the algorithm does: 8-bit digits (256 bins, 4 passes for 32-bit keys, 8 for 64-bit) instead of
// 4-bit, no 2^25 cap (64-bit counts; 2^32 keys per GPU were run), and the launch shape is a property of the
// kernels in librsb200.so, not of this table.  The reference's values are kept under Reference:: because the
// CPU golden sort derives its digit base from THEM (src/CRadixSortCPU.h:35) -- see SURVEY.md section 5.
#pragma once
#include <cstdint>
#include <limits>

template <typename _DataType>
struct AlgorithmParameters {
    using DataType = _DataType;

    // ---- configurable ---------------------------------------------------------------------------------
    /// threads of one onesweep CTA / tile items per thread are fixed in csrc/rs_api.cu (TileCfg);
    /// these two only keep Resize()'s rounding unit (= 1024) identical to the reference's
    inline static constexpr auto _NUM_ITEMS_PER_GROUP = 64U;
    inline static constexpr auto _NUM_GROUPS          = 16U;
    inline static constexpr auto _NUM_ITEMS           = _NUM_ITEMS_PER_GROUP * _NUM_GROUPS;
    /// size of the `globsum` read-back buffer callers still allocate (never asserted by the reference)
    inline static constexpr auto _NUM_HISTOSPLIT      = 512U;
    /// bits per digit on the GPU path
    inline static constexpr auto _NUM_BITS_PER_RADIX  = 8U;
    /// largest input a single sorter accepts by default (the kernels index with 64 bits; this bounds the
    /// device allocation of the test task, 2^28 = the BASELINE.json single-GPU workload)
    inline static constexpr auto _NUM_MAX_INPUT_ELEMS = (1U << 28U);

    // ---- derived ----------------------------------------------------------------------------------------
    inline static constexpr uint32_t _TOTALBITS = sizeof(DataType) << 3U;
    inline static constexpr DataType _MAXINT    = std::numeric_limits<DataType>::max();
    inline static constexpr auto     _RADIX      = (1U << _NUM_BITS_PER_RADIX);
    inline static constexpr auto     _NUM_PASSES = (_TOTALBITS / _NUM_BITS_PER_RADIX);
    /// entries of the `histograms` read-back: passes x 256 digit counts (reference: 16 x 1024 per-item counts)
    inline static constexpr auto     _HISTOSIZE  = _NUM_PASSES * _RADIX;
    inline static constexpr auto     _NUM_PERFORMANCE_ITERATIONS = 5U;

    /// what the reference compiled with; its CPU radix sort's base is _TOTALBITS / 4 (8 or 16)
    struct Reference {
        inline static constexpr auto _NUM_BITS_PER_RADIX  = 4U;
        inline static constexpr auto _RADIX               = 16U;
        inline static constexpr auto _NUM_PASSES          = (sizeof(DataType) << 3U) / 4U;
        inline static constexpr auto _NUM_MAX_INPUT_ELEMS = (1U << 25U);
        inline static constexpr auto _HISTOSIZE           = 64U * 16U * 16U;
    };

    static_assert(_TOTALBITS % _NUM_BITS_PER_RADIX == 0);
    static_assert(_NUM_MAX_INPUT_ELEMS % _NUM_ITEMS == 0);
};
