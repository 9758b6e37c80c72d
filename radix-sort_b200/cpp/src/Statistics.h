// Statistics.h -- running min / max / avg / sum of timing samples, field for field the reference's
// struct (src/Statistics.h:4-31) and rs_statistics of the C ABI.
//
// Quirk kept on purpose: a sample that raises `max` is not looked at for `min` (the reference's update is an
// `if ... else if`), so after the first sample `min` is still +inf and only moves once a later sample is not a new
// maximum.  librsb200 applies the same rule, which lets from() copy the fields verbatim.
#pragma once
#include "rsb200.h"

#include <cstddef>
#include <limits>

struct Statistics {
    double      min = std::numeric_limits<double>::infinity();
    double      max = -std::numeric_limits<double>::infinity();
    double      avg = 0.0;
    double      sum = 0.0;
    std::size_t n   = 0;

    void update(double sample) {
        sum += sample;
        avg = sum / static_cast<double>(++n);
        const bool new_max = sample > max;
        if (new_max) max = sample;
        if (!new_max && sample < min) min = sample;
    }

    static Statistics from(const rs_statistics& c) {
        Statistics s;
        s.min = c.min;
        s.max = c.max;
        s.avg = c.avg;
        s.sum = c.sum;
        s.n   = static_cast<std::size_t>(c.n);
        return s;
    }
};
