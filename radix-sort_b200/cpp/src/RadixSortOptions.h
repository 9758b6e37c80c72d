// RadixSortOptions.h -- runtime switches of the task, same names and command-line spellings as the
// reference's RadixSortOptions (src/RadixSortOptions.h:8-40); parsing is table-driven here.
#pragma once
#include "Parameters.h"

#include <array>
#include <cstddef>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

struct RadixSortOptions {
    /// keys to sort; the reference's default is its 2^25 maximum
    std::size_t num_elements = AlgorithmParameters<float>::Reference::_NUM_MAX_INPUT_ELEMS;
    bool perf_to_stdout     = false;  ///< --perf-to-stdout      human-readable timing table
    bool perf_to_csv        = false;  ///< --perf-to-csv         append to radix_HH-MM-SS.csv
    bool perf_csv_to_stdout = false;  ///< --perf-csv-to-stdout  CSV rows on stdout
    bool verbose            = false;  ///< -v | --verbose

    RadixSortOptions() = default;

    explicit RadixSortOptions(const std::vector<std::string>& args) {
        const std::array<std::pair<std::string_view, bool RadixSortOptions::*>, 5> switches{{
            {"--perf-to-stdout", &RadixSortOptions::perf_to_stdout},
            {"--perf-to-csv", &RadixSortOptions::perf_to_csv},
            {"--perf-csv-to-stdout", &RadixSortOptions::perf_csv_to_stdout},
            {"--verbose", &RadixSortOptions::verbose},
            {"-v", &RadixSortOptions::verbose},
        }};
        for (auto it = args.begin(); it != args.end(); ++it) {
            if (*it == "--num-elements") {
                if (++it == args.end()) break;  // the reference reads past the end here; we stop
                num_elements = static_cast<std::size_t>(std::stoull(*it));  // 64-bit: no 2^31 limit of std::stoi
                continue;
            }
            for (const auto& [name, member] : switches)
                if (*it == name) this->*member = true;
        }
    }
};
