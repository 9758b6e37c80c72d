// CTimer.h -- host wall-clock timer (reference Common/CTimer.h:12-33); GPU phases are CUDA-event timed
// inside the library, this one brackets the CPU golden sorts and whole-task wall time.
#pragma once
#include <chrono>

class CTimer {
public:
    void Start() { t0_ = clock::now(); }
    void Stop() { t1_ = clock::now(); }
    double GetElapsedMilliseconds() const { return std::chrono::duration<double, std::milli>(t1_ - t0_).count(); }
private:
    using clock = std::chrono::high_resolution_clock;
    clock::time_point t0_{}, t1_{};
};
