// CRadixSortTask.h -- the task orchestrator behind IComputeTask (reference src/CRadixSortTask.h:22-92,
// .cpp:60-436): InitResources -> ComputeCPU (std::sort and the CPU radix sort, 5 timed iterations each) ->
// ComputeGPU (pad, one warm ExecuteTask, then 5 timed ExecuteTask = upload + calculate + download) ->
// ValidateResults (memcmp of both results against std::sort) -> ReleaseResources, with the reference's
// performance printing and CSV schema (src/CRadixSortTask.cpp:317-436) plus the roofline columns.
//
// The CPU radix sort (RadixSortCPU<T>, reference src/CRadixSortCPU.h) is test infrastructure, so it is
// injected: SetCpuRadixSort(fn).  Without it ComputeCPU times std::sort only and the "CPU RadixSort"
// validation line is skipped.
#pragma once
#include "Common/CLTypeInformation.h"
#include "Common/CTimer.h"
#include "Common/IComputeTask.h"
#include "Dataset.h"
#include "HostData.h"
#include "Parameters.h"
#include "RadixSortGPU.h"
#include "RadixSortOptions.h"
#include "Statistics.h"

#include <algorithm>
#include <cstring>
#include <ctime>
#include <fstream>
#include <functional>
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>

struct RuntimesCPU {
    Statistics timeRadix{};
    Statistics timeSTL{};
};

template <typename T>
class CRadixSortTask : public IComputeTask {
public:
    using DataType       = T;
    using Parameters     = AlgorithmParameters<T>;
    using CpuRadixSortFn = std::function<void(T*, std::size_t)>;

    CRadixSortTask(const RadixSortOptions& options, std::shared_ptr<Dataset<T>> dataset)
        : mNumberKeys(std::min<std::size_t>(options.num_elements, dataset->dataset.size())),
          mHostData(dataset),
          m_selectedDataset(dataset),
          mOptions(options) {}

    void SetCpuRadixSort(CpuRadixSortFn fn) { mCpuRadixSort = std::move(fn); }

    bool InitResources(rs::Device Device, rs::Context Context) override {
        mNumberKeysRounded = mRadixSortGPU.Resize(mNumberKeys);
        auto& hb           = mHostData.mHostBuffers;
        hb.m_hResultFromGPU.resize(mNumberKeysRounded);
        HostSpans<T> spans{{hb.m_hKeys.data(), hb.m_hKeys.size()},
                           {hb.m_hHistograms.data(), hb.m_hHistograms.size()},
                           {hb.m_hGlobsum.data(), hb.m_hGlobsum.size()},
                           {hb.h_Permut.data(), hb.h_Permut.size()},
                           {hb.m_hResultFromGPU.data(), hb.m_hResultFromGPU.size()}};
        const auto status = mRadixSortGPU.initialize(Device, Context, mNumberKeys, spans);
        if (status != OperationStatus::OK)
            std::cerr << "Failed to initialize Radix Sort on GPU: " << static_cast<int>(status) << " ("
                      << rs_status_string(static_cast<int>(status)) << ": " << rs_last_error_string() << ")\n";
        return status == OperationStatus::OK;
    }

    void ReleaseResources() override { mRadixSortGPU.release(); }

    void ComputeCPU() override {
        auto&       hb = mHostData.mHostBuffers;
        const auto  n  = static_cast<std::size_t>(mNumberKeysRounded);
        const auto  it = Parameters::_NUM_PERFORMANCE_ITERATIONS;
        CTimer      timer;
        timer.Start();
        for (unsigned j = 0; j < it; ++j) {
            std::copy(hb.m_hKeys.begin(), hb.m_hKeys.begin() + n, mHostData.m_resultSTLCPU.begin());
            std::sort(mHostData.m_resultSTLCPU.begin(), mHostData.m_resultSTLCPU.begin() + n);
        }
        timer.Stop();
        mRuntimesCPU.timeSTL.avg = timer.GetElapsedMilliseconds() / it;
        if (mCpuRadixSort) {
            timer.Start();
            for (unsigned j = 0; j < it; ++j) {
                std::copy(hb.m_hKeys.begin(), hb.m_hKeys.begin() + n, mHostData.m_resultRadixSortCPU.begin());
                mCpuRadixSort(mHostData.m_resultRadixSortCPU.data(), n);
            }
            timer.Stop();
            mRuntimesCPU.timeRadix.avg = timer.GetElapsedMilliseconds() / it;
        }
    }

    void ComputeGPU(rs::Context Context, rs::Stream CommandQueue, const LocalWorkSize& LocalWorkSize) override {
        if (mNumberKeys != mNumberKeysRounded) mRadixSortGPU.padGPUData(CommandQueue, sizeof(T) * mNumberKeys);
        ExecuteTask(Context, CommandQueue, LocalWorkSize);  // warm run
        CommandQueue.finish();
        if (mOptions.perf_to_stdout) {
            std::cout << " radixsort cpu avg time: " << mRuntimesCPU.timeRadix.avg << " ms, throughput: "
                      << 1.0e-6 * static_cast<double>(mNumberKeysRounded) / mRuntimesCPU.timeRadix.avg << " Gelem/s\n"
                      << " stl cpu avg time: " << mRuntimesCPU.timeSTL.avg << " ms, throughput: "
                      << 1.0e-6 * static_cast<double>(mNumberKeysRounded) / mRuntimesCPU.timeSTL.avg << " Gelem/s\n"
                      << "Testing performance of GPU task RadixSort" << std::endl;
        }
        // the reference's TestPerformance: N timed end-to-end runs (upload + sort + download), host wall clock
        CTimer     timer;
        const auto it = Parameters::_NUM_PERFORMANCE_ITERATIONS;
        timer.Start();
        for (unsigned j = 0; j < it; ++j) ExecuteTask(Context, CommandQueue, LocalWorkSize);
        CommandQueue.finish();
        timer.Stop();
        mAvgTotalGPU = timer.GetElapsedMilliseconds() / it;
        ReportPerformance();
    }

    bool ValidateResults() override {
        bool        success = true;
        const auto  bytes   = sizeof(T) * static_cast<std::size_t>(mNumberKeys);
        std::cout << "Data set: " << m_selectedDataset->name() << "\nData type: " << TypeNameString<T>::stdint_name << std::endl;
        if (mCpuRadixSort) {
            const bool ok = std::memcmp(mHostData.m_resultRadixSortCPU.data(), mHostData.m_resultSTLCPU.data(), bytes) == 0;
            std::cout << "Validation of CPU RadixSort has " << (ok ? "passed" : "FAILED") << std::endl;
            success = success && ok;
        }
        const bool okGPU = std::memcmp(mHostData.mHostBuffers.m_hResultFromGPU.data(), mHostData.m_resultSTLCPU.data(), bytes) == 0;
        std::cout << "Validation of GPU RadixSort has " << (okGPU ? "passed" : "FAILED") << std::endl;
        success = success && okGPU;
        // new capability: the read-back permutation is the stable argsort of the uploaded keys
        const auto& keys = mHostData.mHostBuffers.m_hKeys;
        const auto& perm = mHostData.mHostBuffers.h_Permut;
        const auto& out  = mHostData.mHostBuffers.m_hResultFromGPU;
        bool        okPerm = true;
        for (std::size_t i = 0; i < mNumberKeysRounded && okPerm; ++i) {
            okPerm = perm[i] < mNumberKeysRounded && keys[perm[i]] == out[i] &&
                     (i == 0 || out[i - 1] != out[i] || perm[i - 1] < perm[i]);
        }
        std::cout << "Validation of GPU permutation (stable argsort) has " << (okPerm ? "passed" : "FAILED") << std::endl;
        return success && okPerm;
    }

    const RuntimesCPU& runtimesCPU() const { return mRuntimesCPU; }
    RuntimesGPU        runtimesGPU() const { return mRadixSortGPU.getRuntimes(); }
    double             avgTotalGPU() const { return mAvgTotalGPU; }

    /// CSV header: the reference's ten columns (src/CRadixSortTask.cpp:327-329) + four appended ones.
    static std::string CsvHeader() {
        return "NumElements,Datatype,Dataset,avgHistogram,avgScan,avgPaste,avgReorder,avgTotalGPU,avgTotalSTLCPU,"
               "avgTotalRDXCPU,GkeysPerSecDevice,BytesPerKey,AchievedGBs,RooflineFraction";
    }
    std::string CsvRow(double hbm_peak_gbs = 6550.1) const {
        const RuntimesGPU  g        = mRadixSortGPU.getRuntimes();
        const double       n        = static_cast<double>(mNumberKeysRounded);
        const double       bytesKey = sizeof(T) + Parameters::_NUM_PASSES * 2.0 * (sizeof(T) + sizeof(uint32_t));
        const double       gbs      = g.timeTotal.avg > 0 ? bytesKey * n / (g.timeTotal.avg * 1e-3) / 1e9 : 0.0;
        std::ostringstream os;
        os << mNumberKeysRounded << ',' << TypeNameString<T>::stdint_name << ',' << m_selectedDataset->name() << ','
           << g.timeHisto.avg << ',' << g.timeScan.avg << ',' << g.timePaste.avg << ',' << g.timeReorder.avg << ','
           << mAvgTotalGPU << ',' << mRuntimesCPU.timeSTL.avg << ',' << mRuntimesCPU.timeRadix.avg << ','
           << (g.timeTotal.avg > 0 ? 1.0e-6 * n / g.timeTotal.avg : 0.0) << ',' << bytesKey << ',' << gbs << ','
           << gbs / hbm_peak_gbs;
        return os.str();
    }

protected:
    void ExecuteTask(rs::Context, rs::Stream CommandQueue, const LocalWorkSize&) {
        if (mOptions.verbose) std::cout << "Sorting " << mNumberKeys << " keys..." << std::endl;
        // the uploaded permutation is the identity again on every run (the previous run overwrote it)
        auto& perm = mHostData.mHostBuffers.h_Permut;
        std::iota(perm.begin(), perm.end(), 0U);
        OperationStatus s = mRadixSortGPU.uploadData(CommandQueue);
        if (s == OperationStatus::OK) s = mRadixSortGPU.calculate(CommandQueue);
        if (s == OperationStatus::OK) s = mRadixSortGPU.downloadData(CommandQueue);
        if (s != OperationStatus::OK)
            std::cerr << "ExecuteTask failed: " << rs_status_string(static_cast<int>(s)) << ": " << rs_last_error_string() << std::endl;
        if (mOptions.verbose) std::cout << "Finished sorting." << std::endl;
    }

    void ReportPerformance() const {
        const RuntimesGPU g = mRadixSortGPU.getRuntimes();
        if (mOptions.perf_to_stdout) {
            std::cout << " histogram avg " << g.timeHisto.avg << " ms | scan avg " << g.timeScan.avg << " ms | paste avg "
                      << g.timePaste.avg << " ms | reorder (all passes) avg " << g.timeReorder.avg << " ms\n"
                      << " device total avg " << g.timeTotal.avg << " ms, throughput: "
                      << 1.0e-6 * static_cast<double>(mNumberKeysRounded) / g.timeTotal.avg << " Gelem/s\n"
                      << " end-to-end (upload+sort+download) avg " << mAvgTotalGPU << " ms, throughput: "
                      << 1.0e-6 * static_cast<double>(mNumberKeysRounded) / mAvgTotalGPU << " Gelem/s" << std::endl;
        }
        if (mOptions.perf_csv_to_stdout) std::cout << CsvHeader() << '\n' << CsvRow() << std::endl;
        if (mOptions.perf_to_csv) {
            const std::time_t t = std::time(nullptr);
            char              name[64];
            std::strftime(name, sizeof name, "radix_%H-%M-%S.csv", std::localtime(&t));
            const bool   fresh = !std::ifstream(name).good();
            std::ofstream f(name, std::ios::app);
            if (fresh) f << CsvHeader() << '\n';
            f << CsvRow() << '\n';
        }
    }

    std::size_t                  mNumberKeys{0};
    uint64_t                     mNumberKeysRounded{0};
    HostDataWithReference<T>     mHostData;
    std::shared_ptr<Dataset<T>>  m_selectedDataset;
    RuntimesCPU                  mRuntimesCPU{};
    RadixSortGPU<T>              mRadixSortGPU;
    RadixSortOptions             mOptions;
    CpuRadixSortFn               mCpuRadixSort;
    double                       mAvgTotalGPU{0.0};
};
