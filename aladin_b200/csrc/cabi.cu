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

extern "C" int alad_abi_version(void) { return 5; }

extern "C" int64_t alad_host_atomic_add(int64_t* p, int64_t v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
extern "C" int32_t alad_host_atomic_cas(int64_t* p, int64_t expected, int64_t desired) {
  return __atomic_compare_exchange_n(p, &expected, desired, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST) ? 1 : 0;
}
extern "C" int64_t alad_host_atomic_load(const int64_t* p) { return __atomic_load_n(p, __ATOMIC_SEQ_CST); }
extern "C" void alad_host_atomic_store(int64_t* p, int64_t v) { __atomic_store_n(p, v, __ATOMIC_SEQ_CST); }
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

// Host-side greedy tiling of consecutive images into <= ALAD_TILE_N packed region rows and <= ALAD_MAX_SEG
// images per tile (the table alad_mrsw_scores_fwd consumes).  Images without valid regions own no column and
// close the current tile.  All pointers are HOST pointers.
extern "C" int alad_region_tiles(const int32_t* nr, const uint8_t* clamp, int32_t Ni, alad_ntile* tiles, int32_t capacity,
                                 int64_t* row_off) {
  using namespace alad;
  ALAD_REQUIRE(Ni >= 0 && (Ni == 0 || (nr && tiles)) && capacity >= 0, "alad_region_tiles: bad arguments");
  int t = -1, cols = 0, seg = 0;
  bool open_tile = false;
  int64_t row = 0;
  for (int i = 0; i < Ni; ++i) {
    const int n = nr[i];
    ALAD_REQUIRE(n >= 0 && n <= ALAD_TILE_N, "alad_region_tiles: image %d has %d scored regions; the kernel supports at most %d",
                 i, n, ALAD_TILE_N);
    if (row_off) row_off[i] = row;
    if (n == 0) {
      open_tile = false;
      continue;
    }
    if (!open_tile || seg >= ALAD_MAX_SEG || cols + n > ALAD_TILE_N) {
      ++t;
      ALAD_REQUIRE(t < capacity, "alad_region_tiles: table capacity %d too small", capacity);
      alad_ntile& nt = tiles[t];
      nt.row_start = (int32_t)row;
      nt.img0 = i;
      nt.nseg = 0;
      nt.clamp_bits = 0;
      for (int k = 0; k < ALAD_MAX_SEG; ++k) nt.seg[k] = 0;
      cols = seg = 0;
      open_tile = true;
    }
    alad_ntile& nt = tiles[t];
    nt.seg[seg] = (uint16_t)(cols | (n << 8));
    if (clamp && clamp[i]) nt.clamp_bits |= 1u << seg;
    nt.nseg = ++seg;
    cols += n;
    row += n;
  }
  return t + 1;
}
