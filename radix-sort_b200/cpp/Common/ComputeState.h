// ComputeState.h -- CUDA device / context / stream state behind the C ABI.
// Replaces the OpenCL bring-up of the reference (Common/ComputeState.cpp:14-104: first GPU platform
// device -> cl::Context -> in-order cl::CommandQueue).  rs::Device / rs::Context / rs::Stream stand where
// the reference passes cl::Device / cl::Context / cl::CommandQueue.
#pragma once
#include "rsb200.h"

#include <iostream>
#include <memory>

namespace rs {

struct Device {
    int ordinal{0};
};

/// Shared-ownership wrapper of rs_context (a cl::Context is a ref-counted handle too).
class Context {
public:
    Context() = default;
    explicit Context(rs_context c) : h_(c, [](rs_context x) { rs_context_destroy(x); }) {}
    rs_context get() const { return h_.get(); }
    rs_context operator()() const { return h_.get(); }  // mirrors `m_CLContext()` null checks
private:
    std::shared_ptr<rs_context_s> h_;
};

/// Shared-ownership wrapper of rs_stream; finish() == cl::CommandQueue::finish().
class Stream {
public:
    Stream() = default;
    explicit Stream(rs_stream s) : h_(s, [](rs_stream x) { rs_stream_destroy(x); }) {}
    rs_stream get() const { return h_.get(); }
    int finish() const { return rs_stream_sync(h_.get()); }
private:
    std::shared_ptr<rs_stream_s> h_;
};

}  // namespace rs

/// Same member names as the reference's ComputeState (Common/ComputeState.h) so drivers read alike.
class ComputeState {
public:
    /// Picks the first CUDA device (the reference takes the first OpenCL GPU device).
    bool init(int ordinal = 0) {
        int count = 0;
        if (rs_device_count(&count) != RS_OK || count <= ordinal) {
            std::cerr << "No CUDA device: " << rs_last_error_string() << "\n";
            return false;
        }
        rs_context c = nullptr;
        if (rs_context_create(ordinal, &c) != RS_OK) {
            std::cerr << "rs_context_create failed: " << rs_last_error_string() << "\n";
            return false;
        }
        m_CLContext = rs::Context(c);
        rs_stream s = nullptr;
        if (rs_stream_create(c, &s) != RS_OK) {
            std::cerr << "rs_stream_create failed: " << rs_last_error_string() << "\n";
            return false;
        }
        m_CLCommandQueue = rs::Stream(s);
        m_device.ordinal = ordinal;
        return true;
    }
    rs::Device device() const { return m_device; }

    rs::Context m_CLContext;
    rs::Stream  m_CLCommandQueue;

private:
    rs::Device m_device{};
};
