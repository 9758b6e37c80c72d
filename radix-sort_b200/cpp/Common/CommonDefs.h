// CommonDefs.h -- shared small types of the facade (counterpart of reference Common/CommonDefs.h:22).
#pragma once
#include <array>
#include <cstddef>

/// Kept for signature compatibility with IComputeTask::ComputeGPU; launch shapes are fixed inside
/// librsb200.so (persistent grids sized from the SM count), so the value is ignored.
using LocalWorkSize = std::array<std::size_t, 3>;
