// Dataset.h -- the five input generators (reference src/Dataset.h:22-137): same class names, name()
// strings and value sequences (std::mt19937 + std::seed_seq / std::uniform_int_distribution are the very
// standard-library objects the reference uses).  One addition: RandomDistributed takes an explicit seed
// (default 0x5EED) where the reference reads the clock, so runs are reproducible.
#pragma once
#include "Parameters.h"

#include <algorithm>
#include <cstdint>
#include <limits>
#include <numeric>
#include <random>
#include <string>
#include <vector>

template <typename T>
struct Dataset {
    using DataType = T;
    explicit Dataset(std::size_t size = AlgorithmParameters<T>::Reference::_NUM_MAX_INPUT_ELEMS) : dataset(size) {}
    virtual ~Dataset() = default;
    virtual const char* name() const { return "UNKNOWN"; }
    std::vector<T> dataset;
};

template <typename T>
struct Zeros : Dataset<T> {
    explicit Zeros(std::size_t size = AlgorithmParameters<T>::Reference::_NUM_MAX_INPUT_ELEMS) : Dataset<T>(size) {
        std::fill(this->dataset.begin(), this->dataset.end(), T{0});
    }
    const char* name() const override { return "Zeros"; }
};

template <typename T>
struct Range : Dataset<T> {
    explicit Range(std::size_t size = AlgorithmParameters<T>::Reference::_NUM_MAX_INPUT_ELEMS) : Dataset<T>(size) {
        std::iota(this->dataset.begin(), this->dataset.end(), std::numeric_limits<T>::min());
    }
    const char* name() const override { return "Range"; }
};

template <typename T>
struct InvertedRange : Dataset<T> {
    explicit InvertedRange(std::size_t size = AlgorithmParameters<T>::Reference::_NUM_MAX_INPUT_ELEMS) : Dataset<T>(size) {
        std::iota(this->dataset.rbegin(), this->dataset.rend(), std::numeric_limits<T>::min());
    }
    const char* name() const override { return "Inverted Range"; }
};

/// "Random Uniform": uniform over the whole value range, first element forced to max, last to min.
template <typename T>
struct RandomDistributed : Dataset<T> {
    explicit RandomDistributed(std::size_t size = AlgorithmParameters<T>::Reference::_NUM_MAX_INPUT_ELEMS,
                               uint64_t seed = 0x5EEDULL)
        : Dataset<T>(size) {
        std::seed_seq seq({static_cast<uint32_t>(seed & 0xFFFFFFFFULL), static_cast<uint32_t>(seed >> 32)});
        std::mt19937 engine(seq);
        std::uniform_int_distribution<T> pick(std::numeric_limits<T>::min(), std::numeric_limits<T>::max());
        for (auto& v : this->dataset) v = pick(engine);
        if (!this->dataset.empty()) {
            this->dataset.front() = std::numeric_limits<T>::max();
            this->dataset.back()  = std::numeric_limits<T>::min();
        }
    }
    const char* name() const override { return "Random Uniform"; }
};

/// "Random Random": raw mt19937 outputs of the fixed seed string (32-bit values also for 64-bit T).
template <typename T>
struct Random : Dataset<T> {
    explicit Random(std::size_t size = AlgorithmParameters<T>::Reference::_NUM_MAX_INPUT_ELEMS) : Dataset<T>(size) {
        const std::string text("Random Test Seed");
        std::seed_seq seq(text.begin(), text.end());
        std::mt19937 engine(seq);
        for (auto& v : this->dataset) v = static_cast<T>(engine());
    }
    const char* name() const override { return "Random Random"; }
};
