// OperationStatus.h -- status codes of the sorter (reference src/OperationStatus.h:4-17); the ordinals are
// the rs_status values of include/rsb200.h, so a C-ABI return code casts straight to this enum.
#pragma once
#include "rsb200.h"

enum class OperationStatus : int {
    OK                      = RS_OK,
    HOST_BUFFERS_FAILED     = RS_HOST_BUFFERS_FAILED,
    INITIALIZATION_FAILED   = RS_INITIALIZATION_FAILED,
    DATA_UPLOAD_FAILED      = RS_DATA_UPLOAD_FAILED,
    CALCULATION_FAILED      = RS_CALCULATION_FAILED,
    DATA_DOWNLOAD_FAILED    = RS_DATA_DOWNLOAD_FAILED,
    CLEANUP_FAILED          = RS_CLEANUP_FAILED,
    RESIZE_FAILED           = RS_RESIZE_FAILED,
    KERNEL_CREATION_FAILED  = RS_KERNEL_CREATION_FAILED,
    PROGRAM_CREATION_FAILED = RS_PROGRAM_CREATION_FAILED,
    NO_SOURCE_FOUND         = RS_NO_SOURCE_FOUND,
    LOADING_SOURCE_FAILED   = RS_LOADING_SOURCE_FAILED,
};
