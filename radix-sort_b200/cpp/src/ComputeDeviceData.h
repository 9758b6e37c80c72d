// ComputeDeviceData.h -- device resource bundle of one sorter (reference src/ComputeDeviceData.h:12-28,
// src/ComputeDeviceData.cpp:10-79).  The reference keeps a cl::Program, a name->cl::Kernel map and a
// name->cl::Buffer map with seven buffers; here the kernels are precompiled sm_100a code inside
// librsb200.so and the buffers live behind one rs_sorter handle.  The buffer names are kept as a
// name -> size table so callers that inspect the layout still can.
#pragma once
#include "Common/CLTypeInformation.h"
#include "Common/ComputeState.h"
#include "Parameters.h"
#include "rsb200.h"

#include <cstddef>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

template <typename _DataType>
struct ComputeDeviceData {
    using DataType   = _DataType;
    using Parameters = AlgorithmParameters<DataType>;

    /// Allocates the ping-pong key and permutation buffers plus scratch for `buffer_size` keys.
    /// Unlike the reference (allocation errors are only printed, src/ComputeDeviceData.cpp:35-39) a failed
    /// allocation throws; RadixSortGPU::initialize turns it into INITIALIZATION_FAILED.
    ComputeDeviceData(rs::Context Context, std::size_t buffer_size, bool withPermutation = true) {
        const uint32_t flags = withPermutation ? RS_WITH_PERMUTATION : 0U;
        const int rc = rs_sorter_create(Context.get(), TypeNameString<DataType>::key_type, buffer_size, flags, &m_sorter);
        if (rc != RS_OK) throw std::runtime_error(std::string("rs_sorter_create: ") + rs_last_error_string());
        kernelNames = {"rs_histogram_kernel", "rs_scan_kernel", "rs_onesweep_kernel"};
        m_dMemoryMap["inputKeys"]          = sizeof(DataType) * buffer_size;
        m_dMemoryMap["outputKeys"]         = sizeof(DataType) * buffer_size;
        m_dMemoryMap["inputPermutations"]  = withPermutation ? sizeof(uint32_t) * buffer_size : 0;
        m_dMemoryMap["outputPermutations"] = withPermutation ? sizeof(uint32_t) * buffer_size : 0;
        m_dMemoryMap["histograms"]         = sizeof(uint64_t) * Parameters::_HISTOSIZE;
        m_dMemoryMap["lookback"]           = 0;  // sized inside the library from the tile count
    }
    ~ComputeDeviceData() {
        if (m_sorter) rs_sorter_destroy(m_sorter);
    }
    ComputeDeviceData(const ComputeDeviceData&)            = delete;
    ComputeDeviceData& operator=(const ComputeDeviceData&) = delete;

    rs_sorter                          m_sorter{nullptr};
    std::vector<std::string>           kernelNames;
    std::map<std::string, std::size_t> m_dMemoryMap;  ///< buffer name -> bytes
};
