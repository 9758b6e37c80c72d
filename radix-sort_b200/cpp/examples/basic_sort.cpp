// basic_sort -- the documented library-user sequence of the reference (examples/basic_sort/basic_sort.cpp:24-139)
// against the B200 facade: Resize -> HostSpans -> initialize -> padGPUData -> uploadData -> calculate ->
// downloadData -> compare with std::sort -> getRuntimes -> release.
//   usage: basic_sort [num_elements=1048576]
#include "Common/ComputeState.h"
#include "Dataset.h"
#include "RadixSortGPU.h"

#include <algorithm>
#include <cstdlib>
#include <iostream>
#include <numeric>
#include <vector>

template <typename T>
static bool sortAndVerify(ComputeState& compute, uint64_t numElements) {
    using P = AlgorithmParameters<T>;
    RandomDistributed<T> dataset(numElements);
    RadixSortGPU<T>      sorter;
    const uint64_t       rounded = sorter.Resize(numElements);

    std::vector<T>        hKeys(rounded), hResult(rounded);
    std::vector<uint32_t> hHistograms(P::_HISTOSIZE), hGlobsum(P::_NUM_HISTOSPLIT), hPermut(rounded);
    std::copy_n(dataset.dataset.begin(), numElements, hKeys.begin());
    // keys past numElements must sort to the end.  The host tail is what uploadData ships (it overwrites the
    // device-side padGPUData fill, as in the reference), so fill it with max: padGPUData's max-1 would sort
    // in front of a genuine max key, and RandomDistributed always contains one.
    std::fill(hKeys.begin() + numElements, hKeys.end(), std::numeric_limits<T>::max());
    std::iota(hPermut.begin(), hPermut.end(), 0U);
    HostSpans<T> spans{{hKeys.data(), hKeys.size()}, {hHistograms.data(), hHistograms.size()},
                       {hGlobsum.data(), hGlobsum.size()}, {hPermut.data(), hPermut.size()},
                       {hResult.data(), hResult.size()}};

    if (sorter.initialize(compute.device(), compute.m_CLContext, numElements, spans) != OperationStatus::OK) {
        std::cerr << "Failed to initialize RadixSortGPU: " << rs_last_error_string() << "\n";
        return false;
    }
    sorter.setLogStream(&std::cout);
    auto& queue = compute.m_CLCommandQueue;
    if (rounded != numElements) sorter.padGPUData(queue, sizeof(T) * numElements);
    if (sorter.uploadData(queue) != OperationStatus::OK) { std::cerr << "Upload failed\n"; return false; }
    if (sorter.calculate(queue) != OperationStatus::OK) { std::cerr << "GPU sort failed\n"; return false; }
    if (sorter.downloadData(queue) != OperationStatus::OK) { std::cerr << "Download failed\n"; return false; }

    std::vector<T> reference(dataset.dataset.begin(), dataset.dataset.begin() + numElements);
    std::sort(reference.begin(), reference.end());
    const bool correct = std::equal(reference.begin(), reference.end(), hResult.begin());

    const auto rt = sorter.getRuntimes();
    std::cout << "\n--- Timing (avg ms, CUDA events) ---\n"
              << "  Histogram : " << rt.timeHisto.avg << "\n  Scan      : " << rt.timeScan.avg
              << "\n  Reorder   : " << rt.timeReorder.avg << "\n  Paste     : " << rt.timePaste.avg
              << "\n  Total     : " << rt.timeTotal.avg << "\n";
    sorter.release();
    return correct;
}

int main(int argc, char** argv) {
    const uint64_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : (1ULL << 20);
    ComputeState   compute;
    if (!compute.init()) return 2;
    const bool ok = sortAndVerify<uint32_t>(compute, n) && sortAndVerify<int64_t>(compute, n);
    std::cout << (ok ? "Sort verified: results match std::sort\n" : "MISMATCH against std::sort\n");
    return ok ? 0 : 1;
}
