// basic_sort -- what a library user writes (counterpart of the reference's examples/basic_sort/basic_sort.cpp):
// size the host buffers with Resize(), wrap them in HostSpans, initialize -> padGPUData -> uploadData -> calculate ->
// downloadData, compare with std::sort, print getRuntimes(), release.
//   usage: basic_sort [num_elements=1048576]
#include "Common/ComputeState.h"
#include "Dataset.h"
#include "RadixSortGPU.h"

#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <numeric>
#include <vector>

namespace {

template <typename T>
bool run_one(ComputeState& gpu, uint64_t n) {
    using P = AlgorithmParameters<T>;
    const RandomDistributed<T> input(n);
    RadixSortGPU<T>            sorter;
    const uint64_t             padded = sorter.Resize(n);  // next multiple of 1024, as callers of the reference do

    // host side: keys (+ tail), result, the two auxiliary read-backs and the index payload
    std::vector<T>        keys(padded, std::numeric_limits<T>::max()), sorted(padded);
    std::vector<uint32_t> histograms(P::_HISTOSIZE), globsum(P::_NUM_HISTOSPLIT), permutation(padded);
    std::copy(input.dataset.begin(), input.dataset.end(), keys.begin());
    // The tail past n is max, not padGPUData's max-1: uploadData ships the host tail over the device-side fill (as in
    // the reference), and max-1 would sort in front of the genuine max key every RandomDistributed set contains.
    std::iota(permutation.begin(), permutation.end(), 0U);

    const HostSpans<T> spans{{keys.data(), keys.size()},
                             {histograms.data(), histograms.size()},
                             {globsum.data(), globsum.size()},
                             {permutation.data(), permutation.size()},
                             {sorted.data(), sorted.size()}};
    auto ok = [](OperationStatus s, const char* what) {
        if (s == OperationStatus::OK) return true;
        std::cerr << what << " failed: " << rs_status_string(static_cast<int>(s)) << " (" << rs_last_error_string() << ")\n";
        return false;
    };
    auto& queue = gpu.m_CLCommandQueue;
    if (!ok(sorter.initialize(gpu.device(), gpu.m_CLContext, n, spans), "initialize")) return false;
    sorter.setLogStream(&std::cout);
    if (padded != n) sorter.padGPUData(queue, sizeof(T) * n);
    if (!ok(sorter.uploadData(queue), "uploadData") || !ok(sorter.calculate(queue), "calculate") ||
        !ok(sorter.downloadData(queue), "downloadData"))
        return false;

    std::vector<T> expected(input.dataset);
    std::sort(expected.begin(), expected.end());
    const bool same = std::equal(expected.begin(), expected.end(), sorted.begin());

    const RuntimesGPU rt = sorter.getRuntimes();
    std::cout << "avg ms (CUDA events): histogram " << rt.timeHisto.avg << ", scan " << rt.timeScan.avg << ", reorder "
              << rt.timeReorder.avg << ", paste " << rt.timePaste.avg << ", total " << rt.timeTotal.avg << "\n";
    sorter.release();
    return same;
}

}  // namespace

int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : (1ULL << 20);
    ComputeState   gpu;
    if (!gpu.init()) return 2;
    const bool ok = run_one<uint32_t>(gpu, n) && run_one<int64_t>(gpu, n);
    std::cout << (ok ? "Sort verified: results match std::sort\n" : "MISMATCH against std::sort\n");
    return ok ? 0 : 1;
}
