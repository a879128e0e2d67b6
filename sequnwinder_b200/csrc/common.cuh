// common.cuh — error plumbing shared by the kernels behind the C ABI (include/sequnwinder_b200.h).
#pragma once

#include <cuda_runtime.h>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/sequnwinder_b200.h"

namespace sqw {

// thread-local message returned by sqw_last_error()
char *last_error_buf();
int set_error(int code, const char *fmt, ...);

#define SQW_CUDA_CHECK(expr)                                                                  \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return sqw::set_error(SQW_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,               \
                                  cudaGetErrorString(_e), __FILE__, __LINE__);                \
    } while (0)

// SQW_OK when a CUDA device is usable, SQW_ERR_NO_DEVICE otherwise (no CPU fallback).
int require_device();

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

}  // namespace sqw
