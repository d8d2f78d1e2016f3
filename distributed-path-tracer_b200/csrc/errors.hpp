// errors.hpp — internal exception type; never crosses the C ABI (api.cu catches everything).
#pragma once

#include <cuda_runtime.h>
#include <stdexcept>
#include <string>

#include "ptb.h"

namespace ptb {

struct Error : std::runtime_error {
    ptb_status code;
    Error(ptb_status c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void cuda_check(cudaError_t e, const char* what, const char* file, int line) {
    if (e != cudaSuccess) {
        ptb_status code = (e == cudaErrorMemoryAllocation) ? PTB_E_OOM : PTB_E_CUDA;
        throw Error(code, std::string(what) + ": " + cudaGetErrorString(e) + " (" + file + ":" + std::to_string(line) + ")");
    }
}

#define PTB_CUDA(expr) ::ptb::cuda_check((expr), #expr, __FILE__, __LINE__)

} // namespace ptb
