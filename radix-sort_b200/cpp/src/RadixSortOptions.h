// RadixSortOptions.h -- runtime flags of the task (reference src/RadixSortOptions.h:8-40): same flag
// spellings and defaults, parsed from a vector of strings.
#pragma once
#include "Parameters.h"

#include <cstddef>
#include <string>
#include <vector>

struct RadixSortOptions {
    std::size_t num_elements{AlgorithmParameters<float>::Reference::_NUM_MAX_INPUT_ELEMS};  // 2^25, as the reference
    bool        perf_to_stdout{false};
    bool        perf_to_csv{false};
    bool        perf_csv_to_stdout{false};
    bool        verbose{false};

    RadixSortOptions() = default;
    explicit RadixSortOptions(const std::vector<std::string>& args) {
        for (std::size_t i = 0; i < args.size(); ++i) {
            const std::string& a = args[i];
            if (a == "--num-elements" && i + 1 < args.size()) {
                num_elements = static_cast<std::size_t>(std::stoull(args[++i]));  // reference: std::stoi, N < 2^31
            } else if (a == "--perf-to-stdout") {
                perf_to_stdout = true;
            } else if (a == "--perf-to-csv") {
                perf_to_csv = true;
            } else if (a == "--perf-csv-to-stdout") {
                perf_csv_to_stdout = true;
            } else if (a == "-v" || a == "--verbose") {
                verbose = true;
            }
        }
    }
};
