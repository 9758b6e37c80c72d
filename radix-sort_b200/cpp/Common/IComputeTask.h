// IComputeTask.h -- the task contract the reference's test runner drives (Common/IComputeTask.h:12-35):
// the same five virtuals in the same order of use; cl::Device / cl::Context / cl::CommandQueue are
// rs::Device / rs::Context / rs::Stream (Common/ComputeState.h).
#pragma once
#include "CommonDefs.h"
#include "ComputeState.h"

class IComputeTask {
public:
    virtual ~IComputeTask() = default;

    /// 1. device-side allocation; false aborts the run
    virtual bool InitResources(rs::Device Device, rs::Context Context) = 0;
    /// 2. golden result on the host (std::sort, CPU radix sort)
    virtual void ComputeCPU() = 0;
    /// 3. the GPU path, warm run + timed runs; LocalWorkSize is accepted and ignored
    virtual void ComputeGPU(rs::Context Context, rs::Stream CommandQueue, const LocalWorkSize& LocalWorkSize) = 0;
    /// 4. GPU result against the golden result
    virtual bool ValidateResults() = 0;
    /// 5. undo InitResources
    virtual void ReleaseResources() = 0;
};
