// Shared C-ABI plumbing: version, last-error text, device queries.
#include "common.h"

namespace alad {

char* error_buffer() {
  static thread_local char buf[512] = {0};
  return buf;
}

int sm_count() {
  static thread_local int cached_dev = -1;
  static thread_local int cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

}  // namespace alad

extern "C" int alad_abi_version(void) { return 1; }
extern "C" const char* alad_last_error(void) { return alad::error_buffer(); }

extern "C" int alad_h2d_2d(void* dst, int64_t dst_pitch, const void* src_host, int64_t src_pitch, int64_t width_bytes,
                           int64_t height, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(width_bytes >= 0 && height >= 0 && dst_pitch >= width_bytes && src_pitch >= width_bytes,
               "alad_h2d_2d: bad geometry");
  if (width_bytes == 0 || height == 0) return ALAD_OK;
  ALAD_REQUIRE(dst && src_host, "alad_h2d_2d: NULL pointer");
  ALAD_CUDA(cudaMemcpy2DAsync(dst, (size_t)dst_pitch, src_host, (size_t)src_pitch, (size_t)width_bytes, (size_t)height,
                              cudaMemcpyHostToDevice, as_stream(stream)));
  return ALAD_OK;
}
