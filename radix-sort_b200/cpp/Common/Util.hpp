// Util.hpp -- make_array helper used by the dataset list of the test driver (reference Common/Util.hpp:6-12).
#pragma once
#include <array>
#include <utility>

template <typename V, typename... Args>
constexpr std::array<V, sizeof...(Args)> make_array(Args&&... args) {
    return {{V(std::forward<Args>(args))...}};
}
