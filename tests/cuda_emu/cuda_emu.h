// TEST INFRASTRUCTURE ONLY -- a minimal host-thread emulation of the CUDA execution model, so that the CUDA-core
// kernels of aladin_b200/csrc (the ones without tcgen05 / TMA) can be executed on the GPU-less authoring container:
// one std::thread per CUDA thread, CTAs run one after the other, __syncthreads / __syncwarp / warp shuffles are
// pthread barriers plus an exchange buffer.  tests/cuda_emu/build_emu.py rewrites the `<<< >>>` launches and the
// dynamic shared-memory declaration of a .cu file and compiles it with g++ against this header.  It checks the
// kernels' index arithmetic and data flow against the oracle; it says nothing about performance, and the product
// never loads an emulated library.
#pragma once
#define __shared__ static
#include <cuda_runtime.h>
#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <string.h>

#include <thread>
#include <vector>

namespace cuda_emu {

struct Cta {
  int nthreads = 0;
  pthread_barrier_t cta_bar;
  std::vector<pthread_barrier_t> warp_bar;
  std::vector<uint64_t> xchg;
  std::vector<int> vote;
  std::vector<unsigned char> dyn;
};
inline thread_local Cta* cta = nullptr;
inline thread_local int tid = 0;

inline void* dynamic_smem() { return cta->dyn.data(); }

template <class F>
void launch_cfg(dim3 grid, dim3 block, size_t smem_bytes, cudaStream_t, F body);

}  // namespace cuda_emu

inline thread_local uint3 threadIdx, blockIdx;
inline thread_local dim3 blockDim, gridDim;

namespace cuda_emu {

template <class F>
void launch_cfg(dim3 grid, dim3 block, size_t smem_bytes, cudaStream_t, F body) {
  const int n = (int)(block.x * block.y * block.z);
  if (n <= 0 || n % 32 != 0) abort();                       // the emulated kernels all use whole warps
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        Cta c;
        c.nthreads = n;
        pthread_barrier_init(&c.cta_bar, nullptr, n);
        c.warp_bar.resize(n / 32);
        for (auto& b : c.warp_bar) pthread_barrier_init(&b, nullptr, 32);
        c.xchg.assign(n, 0);
        c.vote.assign(n, 0);
        c.dyn.assign(smem_bytes + 16, 0xCD);               // poisoned: reads of unwritten shared memory show up
        std::vector<std::thread> th;
        th.reserve(n);
        for (int t = 0; t < n; ++t)
          th.emplace_back([&, t] {
            cta = &c;
            tid = t;
            threadIdx.x = t % block.x;
            threadIdx.y = (t / block.x) % block.y;
            threadIdx.z = t / (block.x * block.y);
            blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
            blockDim = block;
            gridDim = grid;
            body();
          });
        for (auto& t : th) t.join();
        pthread_barrier_destroy(&c.cta_bar);
        for (auto& b : c.warp_bar) pthread_barrier_destroy(&b);
      }
}

inline void warp_barrier() { pthread_barrier_wait(&cta->warp_bar[tid >> 5]); }

}  // namespace cuda_emu

inline void __syncthreads() { pthread_barrier_wait(&cuda_emu::cta->cta_bar); }
inline void __syncwarp(unsigned = 0xffffffffu) { cuda_emu::warp_barrier(); }
inline int __syncthreads_or(int p) {
  using namespace cuda_emu;
  cta->vote[tid] = p;
  __syncthreads();
  int r = 0;
  for (int t = 0; t < cta->nthreads; ++t) r |= cta->vote[t];
  __syncthreads();
  return r != 0;
}
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  using namespace cuda_emu;
  static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  cta->xchg[tid] = raw;
  warp_barrier();
  raw = cta->xchg[(tid & ~31) | ((tid & 31) ^ lane_mask)];
  warp_barrier();
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src_lane) {
  using namespace cuda_emu;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  cta->xchg[tid] = raw;
  warp_barrier();
  raw = cta->xchg[(tid & ~31) | (src_lane & 31)];
  warp_barrier();
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}
template <class T>
inline T __shfl_down_sync(unsigned, T v, unsigned delta) {
  using namespace cuda_emu;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  cta->xchg[tid] = raw;
  warp_barrier();
  const int src = (tid & 31) + (int)delta;
  if (src < 32) raw = cta->xchg[(tid & ~31) | src];
  warp_barrier();
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}
inline unsigned __ballot_sync(unsigned, int pred) {
  using namespace cuda_emu;
  cta->xchg[tid] = pred ? 1 : 0;
  warp_barrier();
  unsigned m = 0;
  for (int l = 0; l < 32; ++l) m |= (unsigned)(cta->xchg[(tid & ~31) | l] & 1) << l;
  warp_barrier();
  return m;
}
template <class T>
inline T __reduce_add_sync(unsigned, T v) {
  using namespace cuda_emu;
  cta->xchg[tid] = (uint64_t)(int64_t)v;
  warp_barrier();
  int64_t acc = 0;
  for (int l = 0; l < 32; ++l) acc += (int64_t)cta->xchg[(tid & ~31) | l];
  warp_barrier();
  return (T)acc;
}
template <class T>
inline T __reduce_max_sync(unsigned, T v) {
  using namespace cuda_emu;
  cta->xchg[tid] = (uint64_t)(int64_t)v;
  warp_barrier();
  T acc = (T)(int64_t)cta->xchg[tid & ~31];
  for (int l = 1; l < 32; ++l) {
    const T o = (T)(int64_t)cta->xchg[(tid & ~31) | l];
    acc = o > acc ? o : acc;
  }
  warp_barrier();
  return acc;
}
template <class T>
inline T __ldg(const T* p) { return *p; }
template <class T>
inline T __ldcg(const T* p) { return *p; }
inline float __int_as_float(int i) { float f; memcpy(&f, &i, 4); return f; }
inline int __float_as_int(float f) { int i; memcpy(&i, &f, 4); return i; }
inline float __uint_as_float(unsigned i) { float f; memcpy(&f, &i, 4); return f; }
inline unsigned __float_as_uint(float f) { unsigned i; memcpy(&i, &f, 4); return i; }
inline float __fdividef(float a, float b) { return a / b; }
inline float __expf(float x) { return expf(x); }
inline float __logf(float x) { return logf(x); }
inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }
inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
inline unsigned long long atomicMax(unsigned long long* p, unsigned long long v) {
  unsigned long long old = __atomic_load_n(p, __ATOMIC_RELAXED);
  while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {}
  return old;
}
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_RELAXED); }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned delta) {
  using namespace cuda_emu;
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  cta->xchg[tid] = raw;
  warp_barrier();
  const int src = (tid & 31) - (int)delta;
  if (src >= 0) raw = cta->xchg[(tid & ~31) | src];
  warp_barrier();
  T out;
  memcpy(&out, &raw, sizeof(T));
  return out;
}
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline float atomicAdd(float* addr, float v) {
  uint32_t* p = reinterpret_cast<uint32_t*>(addr);
  uint32_t old = __atomic_load_n(p, __ATOMIC_RELAXED), want;
  float f;
  do {
    memcpy(&f, &old, 4);
    const float s = f + v;
    memcpy(&want, &s, 4);
  } while (!__atomic_compare_exchange_n(p, &old, want, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return f;
}

// the slice of the runtime API the emulated translation units call; "device" pointers are host pointers here
template <class T>
inline cudaError_t cudaFuncSetAttribute(T*, cudaFuncAttribute, int) { return cudaSuccess; }
// (inline: the header may be part of several translation units of one emulated library)
extern "C" {
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return cudaSuccess; }
inline cudaError_t cudaDeviceGetAttribute(int* v, cudaDeviceAttr attr, int) {
  *v = attr == cudaDevAttrMaxSharedMemoryPerBlockOptin ? 232448      // 227 KB, as on sm_100
       : attr == cudaDevAttrMultiProcessorCount ? 2 : 0;             // persistent grids stay small under emulation
  return cudaSuccess;
}
inline cudaError_t cudaGetLastError(void) { return cudaSuccess; }
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessor(int* n, const void*, int, size_t) { *n = 2; return cudaSuccess; }
inline cudaError_t cudaOccupancyMaxActiveBlocksPerMultiprocessorWithFlags(int* n, const void*, int, size_t, unsigned) {
  *n = 2;
  return cudaSuccess;
}
inline cudaError_t cudaFuncSetAttribute(const void*, cudaFuncAttribute, int) { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "emulated"; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemset2DAsync(void* p, size_t pitch, int v, size_t width, size_t height, cudaStream_t) {
  for (size_t r = 0; r < height; ++r) memset(static_cast<char*>(p) + r * pitch, v, width);
  return cudaSuccess;
}
inline cudaError_t cudaMemcpyAsync(void* dst, const void* src, size_t n, cudaMemcpyKind, cudaStream_t) {
  memcpy(dst, src, n);
  return cudaSuccess;
}
// streams and events: the emulation is synchronous, so a side stream is the same queue and events are no-ops
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventCreateWithFlags(cudaEvent_t* e, unsigned) { *e = nullptr; return cudaSuccess; }
inline cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned) { return cudaSuccess; }
inline cudaError_t cudaMemcpy2DAsync(void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                                     cudaMemcpyKind, cudaStream_t) {
  for (size_t r = 0; r < height; ++r) memcpy(static_cast<char*>(dst) + r * dpitch, static_cast<const char*>(src) + r * spitch, width);
  return cudaSuccess;
}
}

#ifndef CUDA_EMU_FULL_LIBRARY          // single translation unit: the two helpers csrc/cabi.cu would provide
namespace alad {
char* error_buffer() {
  static thread_local char buf[512];
  return buf;
}
int sm_count() { return 2; }              // persistent grids stay small under emulation
}  // namespace alad
extern "C" const char* alad_last_error(void) { return alad::error_buffer(); }
#endif
