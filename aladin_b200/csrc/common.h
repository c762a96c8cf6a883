// Error plumbing shared by the C-ABI translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>

#include "alad_b200.h"

namespace alad {

char* error_buffer();  // thread-local, 512 bytes (defined in cabi.cu)

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(error_buffer(), 512, fmt, ap);
  va_end(ap);
  return code;
}

#define ALAD_REQUIRE(cond, ...)                                  \
  do {                                                           \
    if (!(cond)) return ::alad::fail(ALAD_ERR_INVALID, __VA_ARGS__); \
  } while (0)

#define ALAD_CUDA(expr)                                                                          \
  do {                                                                                           \
    cudaError_t _e = (expr);                                                                     \
    if (_e != cudaSuccess)                                                                       \
      return ::alad::fail(ALAD_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                          __FILE__, __LINE__);                                                   \
  } while (0)

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

int sm_count();  // SMs of the current device (cached)

}  // namespace alad
