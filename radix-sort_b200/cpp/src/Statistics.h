// Statistics.h -- running min/max/avg/sum (reference src/Statistics.h:4-31).  The update rule keeps the
// reference's `if / else if`: the first sample only raises `max`; `min` moves once a later sample is not a
// new maximum.  rs_statistics of the C ABI follows the same rule, so values can be copied field by field.
#pragma once
#include "rsb200.h"

#include <cstddef>
#include <limits>

struct Statistics {
    double      min{std::numeric_limits<double>::infinity()};
    double      max{-std::numeric_limits<double>::infinity()};
    double      avg{0.0};
    double      sum{0.0};
    std::size_t n{0U};

    void update(double value) {
        ++n;
        sum += value;
        avg = sum / static_cast<double>(n);
        if (value > max) {
            max = value;
        } else if (value < min) {
            min = value;
        }
    }

    static Statistics from(const rs_statistics& s) {
        Statistics out;
        out.min = s.min; out.max = s.max; out.avg = s.avg; out.sum = s.sum; out.n = static_cast<std::size_t>(s.n);
        return out;
    }
};
