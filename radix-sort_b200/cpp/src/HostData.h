// HostData.h -- caller-owned host buffers handed to the sorter (reference src/HostData.h:20-64,
// src/HostData.cpp:9-28).  Same five fields; spans are non-owning and must outlive the sorter.
#pragma once
#include "Dataset.h"
#include "Parameters.h"

#include <algorithm>
#include <cstdint>
#include <memory>
#include <numeric>
#include <span>
#include <vector>

template <typename BufferData, typename BufferAux>
struct HostBuffers {
    BufferData m_hKeys;           ///< input keys
    BufferAux  m_hHistograms;     ///< read-back: digit histograms of the last sort (passes x 256, low 32 bits)
    BufferAux  m_hGlobsum;        ///< read-back kept for layout compatibility (the look-back needs no globsum); zero-filled
    BufferAux  h_Permut;          ///< in: index payload (iota), out: stable sorting permutation
    BufferData m_hResultFromGPU;  ///< sorted keys
};

template <typename T>
using HostData = HostBuffers<std::vector<T>, std::vector<uint32_t>>;
template <typename T>
using HostSpans = HostBuffers<std::span<T>, std::span<uint32_t>>;

/// Test-side bundle: host buffers plus the two CPU results they are compared with.
template <typename T>
struct HostDataWithReference {
    using DataType     = T;
    using Parameters   = AlgorithmParameters<T>;
    using ResultBuffer = std::vector<T>;

    HostDataWithReference() = delete;
    /// Buffers are sized for `capacity` keys (the reference pre-sizes to its 2^25 maximum; pass 0 to use the
    /// dataset size rounded up to 1024).
    explicit HostDataWithReference(std::shared_ptr<Dataset<T>> dataset, std::size_t capacity = 0) {
        const std::size_t n   = dataset->dataset.size();
        const std::size_t cap = std::max<std::size_t>(capacity, (n + 1023) / 1024 * 1024);
        m_resultSTLCPU.resize(cap);
        m_resultRadixSortCPU.resize(cap);
        mHostBuffers.m_hKeys.resize(cap);  // value-initialised: keys past n are 0, as in the reference flow
        mHostBuffers.m_hHistograms.resize(Parameters::_HISTOSIZE);
        mHostBuffers.m_hGlobsum.resize(Parameters::_NUM_HISTOSPLIT);
        mHostBuffers.h_Permut.resize(cap);
        std::iota(mHostBuffers.h_Permut.begin(), mHostBuffers.h_Permut.end(), 0U);
        std::copy(dataset->dataset.begin(), dataset->dataset.end(), mHostBuffers.m_hKeys.begin());
    }

    ResultBuffer m_resultSTLCPU;
    ResultBuffer m_resultRadixSortCPU;
    HostData<T>  mHostBuffers;
};
