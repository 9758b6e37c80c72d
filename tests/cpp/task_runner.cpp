// task_runner -- the reference's integration test (tests/tests.cpp:90-113, tests/CTestBase.cpp:20-67) against
// the B200 facade: for uint32/int32/uint64/int64 x {Zeros, Range, InvertedRange, RandomDistributed, Random}
// run IComputeTask: InitResources -> ComputeCPU -> ComputeGPU -> ValidateResults -> ReleaseResources.
// The CPU radix sort is the oracle's restatement of RadixSortCPU (oracle/liboracle.so) -- test
// infrastructure injected into the task, never linked into the product.
// Unlike the reference (RunComputeTask returns true even on "INVALID RESULTS!") a failed validation fails
// the run.   usage: task_runner [--num-elements N] [--perf-to-stdout] [--perf-csv-to-stdout] [-v]
#include "CRadixSortTask.h"
#include "Common/Util.hpp"

#include <algorithm>
#include <iostream>
#include <memory>
#include <string>
#include <vector>

extern "C" {
int oracle_radix_sort_cpu_u32(uint32_t*, size_t);
int oracle_radix_sort_cpu_i32(int32_t*, size_t);
int oracle_radix_sort_cpu_u64(uint64_t*, size_t);
int oracle_radix_sort_cpu_i64(int64_t*, size_t);
}
static void cpuRadix(uint32_t* d, size_t n) { oracle_radix_sort_cpu_u32(d, n); }
static void cpuRadix(int32_t* d, size_t n) { oracle_radix_sort_cpu_i32(d, n); }
static void cpuRadix(uint64_t* d, size_t n) { oracle_radix_sort_cpu_u64(d, n); }
static void cpuRadix(int64_t* d, size_t n) { oracle_radix_sort_cpu_i64(d, n); }
// 16-bit keys: the reference's RadixSortCPU is not instantiated for them (src/CRadixSortTask.cpp:440-443); the
// "CPU radix" leg of those tasks is a stable sort
static void cpuRadix(uint16_t* d, size_t n) { std::stable_sort(d, d + n); }
static void cpuRadix(int16_t* d, size_t n) { std::stable_sort(d, d + n); }

class CTestBase {
public:
    explicit CTestBase(std::vector<std::string> arguments) : m_arguments(std::move(arguments)) {}
    bool InitCLContext() { return m_computeState.init(); }

    bool RunComputeTask(IComputeTask& Task, const LocalWorkSize& lws) {
        if (m_computeState.m_CLContext() == nullptr) {
            std::cerr << "Error: RunComputeTask() cannot execute because the context is null.\n";
            return false;
        }
        if (!Task.InitResources(m_computeState.device(), m_computeState.m_CLContext)) {
            std::cerr << "Error during resource allocation. Aborting execution." << std::endl;
            Task.ReleaseResources();
            return false;
        }
        std::cout << "Computing CPU reference result...";
        Task.ComputeCPU();
        std::cout << "DONE\nComputing GPU result..." << std::endl;
        Task.ComputeGPU(m_computeState.m_CLContext, m_computeState.m_CLCommandQueue, lws);
        std::cout << "DONE" << std::endl;
        const bool valid = Task.ValidateResults();
        std::cout << (valid ? "GOLD TEST PASSED!\n" : "INVALID RESULTS!\n");
        Task.ReleaseResources();
        return valid;
    }

protected:
    std::vector<std::string> m_arguments;
    ComputeState             m_computeState;
};

class CRunner : public CTestBase {
public:
    using CTestBase::CTestBase;

    template <typename T>
    bool runTask(const RadixSortOptions& options, const LocalWorkSize& lws) {
        const auto n        = options.num_elements;
        const auto datasets = make_array<std::shared_ptr<Dataset<T>>>(
            std::make_shared<Zeros<T>>(n), std::make_shared<Range<T>>(n), std::make_shared<InvertedRange<T>>(n),
            std::make_shared<RandomDistributed<T>>(n), std::make_shared<Random<T>>(n));
        bool success = true;
        for (const auto& ds : datasets) {
            CRadixSortTask<T> task(options, ds);
            task.SetCpuRadixSort([](T* d, std::size_t m) { cpuRadix(d, m); });
            success = RunComputeTask(task, lws) && success;
        }
        return success;
    }

    bool DoCompute() {
        RadixSortOptions options(m_arguments);
        bool has_n = false;
        for (const auto& a : m_arguments) has_n = has_n || a == "--num-elements";
        if (!has_n) options.num_elements = 1U << 20;  // keep the default run short; pass --num-elements 33554432 for the reference's size
        const LocalWorkSize lws{1, 1, 1};
        bool ok = runTask<uint32_t>(options, lws);
        ok      = runTask<int32_t>(options, lws) && ok;
        ok      = runTask<uint64_t>(options, lws) && ok;
        ok      = runTask<int64_t>(options, lws) && ok;
        for (const auto& a : m_arguments)
            if (a == "--narrow-types") {  // the other two rows of the reference's type table (Common/CLTypeInformation.h:12-22)
                ok = runTask<uint16_t>(options, lws) && ok;
                ok = runTask<int16_t>(options, lws) && ok;
            }
        return ok;
    }
};

int main(int argc, char** argv) {
    CRunner runner(std::vector<std::string>(argv + 1, argv + argc));
    if (!runner.InitCLContext()) {
        std::cerr << "no CUDA device / context: there is no CPU fallback\n";
        return 2;
    }
    const bool ok = runner.DoCompute();
    std::cout << (ok ? "ALL TASKS PASSED\n" : "SOME TASKS FAILED\n");
    return ok ? 0 : 1;
}
