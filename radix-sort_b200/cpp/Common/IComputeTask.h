// IComputeTask.h -- the five-call task contract of the reference (Common/IComputeTask.h:12-35), with the
// OpenCL handle types replaced by the CUDA-side wrappers of ComputeState.h.
#pragma once
#include "CommonDefs.h"
#include "ComputeState.h"

class IComputeTask {
public:
    virtual ~IComputeTask() = default;
    /// Allocate everything the task needs on the device.
    virtual bool InitResources(rs::Device Device, rs::Context Context) = 0;
    /// Undo InitResources.
    virtual void ReleaseResources() = 0;
    /// Run (and time) the GPU path.
    virtual void ComputeGPU(rs::Context Context, rs::Stream CommandQueue, const LocalWorkSize& LocalWorkSize) = 0;
    /// Produce the golden result on the host.
    virtual void ComputeCPU() = 0;
    /// GPU result == golden result?
    virtual bool ValidateResults() = 0;
};
