// RadixSortGPU.h -- the GPU sorter class of the reference (src/RadixSortGPU.h:38-124, .cpp:15-602) over the
// C ABI of librsb200.so.  Same public methods, same OperationStatus results, same call order
// (initialize -> [padGPUData] -> uploadData -> calculate -> downloadData -> release).  Differences a
// caller can see:
//   * cl::Device / cl::Context / cl::CommandQueue become rs::Device / rs::Context / rs::Stream,
//   * `nn` is 64-bit (the reference caps at 2^25 keys),
//   * uploadData / calculate are stream-ordered and return at once; downloadData blocks (the reference
//     finish()es after every command),
//   * h_Permut comes back as the STABLE SORTING PERMUTATION (the reference's kernels never touch it and
//     return the uploaded identity, src/kernels/RadixSort.cl:79-80),
//   * m_hHistograms receives passes x 256 digit counts; m_hGlobsum is zero-filled (no globsum exists).
#pragma once
#include "Common/ComputeState.h"
#include "ComputeDeviceData.h"
#include "HostData.h"
#include "OperationStatus.h"
#include "Statistics.h"
#include "rsb200.h"

#include <algorithm>
#include <cstdint>
#include <iostream>
#include <memory>
#include <vector>

/// Per-kernel-class runtimes (reference src/RadixSortGPU.h:18-24), CUDA-event timed, milliseconds.
struct RuntimesGPU {
    Statistics timeHisto{};
    Statistics timeScan{};
    Statistics timeReorder{};
    Statistics timePaste{};  ///< always 0: the paste step is subsumed by the decoupled look-back
    Statistics timeTotal{};
};

template <typename DataType>
class RadixSortGPU {
public:
    OperationStatus initialize(rs::Device Device, rs::Context Context, uint64_t nn, const HostSpans<DataType>& hostSpans) {
        (void)Device;
        mNumberKeys        = nn;
        mNumberKeysRounded = Resize(nn);
        mHostSpans         = hostSpans;
        if (mHostSpans.m_hKeys.size() < mNumberKeysRounded || mHostSpans.m_hResultFromGPU.size() < mNumberKeysRounded)
            return OperationStatus::HOST_BUFFERS_FAILED;
        try {
            mDeviceData = std::make_shared<ComputeDeviceData<DataType>>(Context, mNumberKeysRounded,
                                                                        mHostSpans.h_Permut.size() >= mNumberKeysRounded);
        } catch (const std::exception& e) {
            if (mOutStream) *mOutStream << e.what() << std::endl;
            return OperationStatus::INITIALIZATION_FAILED;
        }
        mWithPermutation = mHostSpans.h_Permut.size() >= mNumberKeysRounded;
        return OperationStatus::OK;
    }

    /// H2D of `rounded` keys and the index payload (reference CopyDataToDevice, src/RadixSortGPU.cpp:365-387).
    OperationStatus uploadData(rs::Stream CommandQueue) {
        if (!mDeviceData) return OperationStatus::DATA_UPLOAD_FAILED;
        const int rc = rs_sorter_upload(mDeviceData->m_sorter, CommandQueue.get(), mHostSpans.m_hKeys.data(),
                                        mWithPermutation ? mHostSpans.h_Permut.data() : nullptr, mNumberKeysRounded);
        return static_cast<OperationStatus>(rc);
    }

    /// The pass loop (reference src/RadixSortGPU.cpp:310-346): one histogram launch, one scan launch and
    /// one onesweep launch per 8-bit digit, enqueued back to back.
    OperationStatus calculate(rs::Stream CommandQueue) {
        if (!mDeviceData) return OperationStatus::CALCULATION_FAILED;
        if (mOutStream) *mOutStream << "Sorting " << mNumberKeysRounded << " keys: histogram, scan, "
                                    << AlgorithmParameters<DataType>::_NUM_PASSES << " onesweep passes" << std::endl;
        return static_cast<OperationStatus>(rs_sorter_sort(mDeviceData->m_sorter, CommandQueue.get(), mNumberKeysRounded));
    }

    /// D2H of keys, permutation and the digit histograms; blocks (reference CopyDataFromDevice, :389-429).
    OperationStatus downloadData(rs::Stream CommandQueue) {
        if (!mDeviceData) return OperationStatus::DATA_DOWNLOAD_FAILED;
        int rc = rs_sorter_download(mDeviceData->m_sorter, CommandQueue.get(), mHostSpans.m_hResultFromGPU.data(),
                                    mWithPermutation ? mHostSpans.h_Permut.data() : nullptr, mNumberKeysRounded);
        if (rc != RS_OK) return static_cast<OperationStatus>(rc);
        if (!mHostSpans.m_hHistograms.empty()) {
            std::vector<uint64_t> h(8 * 256);
            int passes = 0;
            rc = rs_sorter_download_histograms(mDeviceData->m_sorter, CommandQueue.get(), h.data(), h.size(), &passes);
            if (rc != RS_OK) return static_cast<OperationStatus>(rc);
            const std::size_t m = std::min<std::size_t>(mHostSpans.m_hHistograms.size(), static_cast<std::size_t>(passes) * 256);
            for (std::size_t i = 0; i < m; ++i) mHostSpans.m_hHistograms[i] = static_cast<uint32_t>(h[i]);
        }
        std::fill(mHostSpans.m_hGlobsum.begin(), mHostSpans.m_hGlobsum.end(), 0U);
        return OperationStatus::OK;
    }

    OperationStatus release() {
        mDeviceData = nullptr;
        return OperationStatus::OK;
    }

    void setLogStream(std::ostream* out) noexcept { mOutStream = out; }

    /// Next multiple of 1024 (reference src/RadixSortGPU.cpp:287-297).  The kernels handle ragged sizes; the
    /// rounding is kept because callers size their host buffers with it.
    uint64_t Resize(uint64_t nn) const noexcept { return rs_round_up_keys(nn); }

    /// Fill device keys from byte offset `paddingOffset` up to the rounded length with max-1
    /// (reference src/RadixSortGPU.cpp:269-285).  As in the reference, a later uploadData overwrites it.
    void padGPUData(rs::Stream CommandQueue, std::size_t paddingOffset) {
        if (!mDeviceData) return;
        rs_sorter_pad(mDeviceData->m_sorter, CommandQueue.get(), paddingOffset / sizeof(DataType), mNumberKeysRounded);
    }

    RuntimesGPU getRuntimes() const {
        RuntimesGPU out;
        if (!mDeviceData) return out;
        rs_runtimes rt{};
        if (rs_sorter_runtimes(mDeviceData->m_sorter, &rt) != RS_OK) return out;
        out.timeHisto   = Statistics::from(rt.time_histo);
        out.timeScan    = Statistics::from(rt.time_scan);
        out.timeReorder = Statistics::from(rt.time_reorder);
        out.timePaste   = Statistics::from(rt.time_paste);
        out.timeTotal   = Statistics::from(rt.time_total);
        return out;
    }

private:
    std::shared_ptr<ComputeDeviceData<DataType>> mDeviceData;
    HostSpans<DataType>                          mHostSpans;
    uint64_t                                     mNumberKeys{0};
    uint64_t                                     mNumberKeysRounded{0};
    bool                                         mWithPermutation{false};
    std::ostream*                                mOutStream{nullptr};
};
