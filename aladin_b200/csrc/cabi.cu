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
