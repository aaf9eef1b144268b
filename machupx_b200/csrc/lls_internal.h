// Internal helpers shared by the CUDA translation units of liblls (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include "../../include/lls.h"

#define LLS_CUDA_TRY(expr)                                                                    \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess) {                                                              \
            lls::set_last_error(__FILE__, __LINE__, cudaGetErrorString(_e));                  \
            return LLS_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

namespace lls {
void set_last_error(const char* file, int line, const char* msg);
}
