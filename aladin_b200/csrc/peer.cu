// Peer exchange over NVLink without SM-resident collectives (multi-GPU retrieval).
//
// The host-resident gallery path uploads + packs 1/world of the captions on every rank and replicates the packed bf16
// rows on all ranks while the persistent tcgen05 scoring kernel of the previous caption phase owns every SM.  An
// SM-resident collective (NCCL all-gather) cannot start next to that kernel and lands in the gap between two scoring
// launches (profiles/r02_e2e_timeline.md); copy engines can.  This file provides the plumbing: IPC-exportable device
// buffers, peer-to-peer copies on the copy engines, and stream-ordered flags (a one-warp kernel that stores a
// sequence number into every peer's flag slot / spins on the local slots until every peer has caught up).
// No reference counterpart: the reference (alad/evaluation.py) is single-process.
#include <stdint.h>
#include <string.h>

#include "common.h"

#ifndef ALAD_CPU_EMU

namespace alad {
namespace {

struct PeerPtrs {
  int32_t* p[ALAD_MAX_PEERS];
};

// lane q < n: *(flags[q]) = value, visible system-wide after everything this stream did before
__global__ void peer_signal_kernel(PeerPtrs dst, int n, int32_t value) {
  const int q = threadIdx.x;
  if (q < n && dst.p[q]) {
    __threadfence_system();
    asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(dst.p[q]), "r"(value) : "memory");
  }
}

// lane q < n spins until flags[q] >= value (bounded: after ~timeout_ns of wall clock *error = 1 + q and the kernel traps, so
// that a lost peer can neither hang the GPU nor let stale rows be scored)
__global__ void peer_wait_kernel(const int32_t* flags, int n, int32_t value, int skip, long long timeout_ns, int32_t* error) {
  const int q = threadIdx.x;
  if (q >= n || q == skip) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  unsigned spins = 0;
  for (;;) {
    int32_t v;
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(flags + q) : "memory");
    if (v - value >= 0) break;                       // wrap-safe comparison of sequence numbers
    if ((++spins & 1023u) == 0) {
      unsigned long long t1;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
      if ((long long)(t1 - t0) > timeout_ns) {
        // a peer never delivered: the rows this stream is about to score would be stale.  Record who, then abort the
        // context -- the failure surfaces at the caller's next synchronisation instead of as silently wrong scores.
        if (error) atomicExch(error, 1 + q);
        __threadfence_system();
        __trap();
      }
    }
    __nanosleep(200);
  }
}

}  // namespace
}  // namespace alad

extern "C" int alad_peer_alloc(void** ptr, int64_t bytes) {
  using namespace alad;
  ALAD_REQUIRE(ptr && bytes > 0, "alad_peer_alloc: bad arguments");
  void* p = nullptr;
  ALAD_CUDA(cudaMalloc(&p, (size_t)bytes));
  ALAD_CUDA(cudaMemset(p, 0, (size_t)bytes));
  *ptr = p;
  return ALAD_OK;
}

extern "C" int alad_peer_free(void* ptr) {
  using namespace alad;
  if (ptr) ALAD_CUDA(cudaFree(ptr));
  return ALAD_OK;
}

extern "C" int alad_peer_export(const void* ptr, void* handle64) {
  using namespace alad;
  static_assert(sizeof(cudaIpcMemHandle_t) == ALAD_PEER_HANDLE_BYTES, "handle size");
  ALAD_REQUIRE(ptr && handle64, "alad_peer_export: NULL pointer");
  ALAD_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(ptr)));
  return ALAD_OK;
}

extern "C" int alad_peer_open(const void* handle64, void** ptr) {
  using namespace alad;
  ALAD_REQUIRE(ptr && handle64, "alad_peer_open: NULL pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  ALAD_CUDA(cudaIpcOpenMemHandle(ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return ALAD_OK;
}

extern "C" int alad_peer_close(void* ptr) {
  using namespace alad;
  if (ptr) ALAD_CUDA(cudaIpcCloseMemHandle(ptr));
  return ALAD_OK;
}

extern "C" int alad_peer_copy(void* dst, const void* src, int64_t bytes, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(bytes >= 0, "alad_peer_copy: negative size");
  if (bytes == 0) return ALAD_OK;
  ALAD_REQUIRE(dst && src, "alad_peer_copy: NULL pointer");
  ALAD_CUDA(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToDevice, as_stream(stream)));
  return ALAD_OK;
}

extern "C" int alad_peer_signal(void* const* flag_ptrs, int32_t n, int32_t value, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(n >= 0 && n <= ALAD_MAX_PEERS && (n == 0 || flag_ptrs), "alad_peer_signal: bad arguments");
  if (n == 0) return ALAD_OK;
  PeerPtrs d = {};
  for (int q = 0; q < n; ++q) d.p[q] = reinterpret_cast<int32_t*>(flag_ptrs[q]);
  peer_signal_kernel<<<1, 32, 0, as_stream(stream)>>>(d, n, value);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_peer_wait(const int32_t* flags, int32_t n, int32_t value, int32_t skip, int64_t timeout_ms, int32_t* error,
                              void* stream) {
  using namespace alad;
  ALAD_REQUIRE(n >= 0 && n <= ALAD_MAX_PEERS && (n == 0 || flags) && timeout_ms > 0, "alad_peer_wait: bad arguments");
  if (n == 0) return ALAD_OK;
  peer_wait_kernel<<<1, 32, 0, as_stream(stream)>>>(flags, n, value, skip, (long long)timeout_ms * 1000000ll, error);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

#else  // ALAD_CPU_EMU: no peers on the virtual device

extern "C" int alad_peer_alloc(void**, int64_t) { return alad::fail(ALAD_ERR_UNSUPPORTED, "peer memory needs CUDA devices"); }
extern "C" int alad_peer_free(void*) { return ALAD_OK; }
extern "C" int alad_peer_export(const void*, void*) { return alad::fail(ALAD_ERR_UNSUPPORTED, "peer memory needs CUDA devices"); }
extern "C" int alad_peer_open(const void*, void**) { return alad::fail(ALAD_ERR_UNSUPPORTED, "peer memory needs CUDA devices"); }
extern "C" int alad_peer_close(void*) { return ALAD_OK; }
extern "C" int alad_peer_copy(void*, const void*, int64_t, void*) { return alad::fail(ALAD_ERR_UNSUPPORTED, "peer memory needs CUDA devices"); }
extern "C" int alad_peer_signal(void* const*, int32_t, int32_t, void*) { return alad::fail(ALAD_ERR_UNSUPPORTED, "peer memory needs CUDA devices"); }
extern "C" int alad_peer_wait(const int32_t*, int32_t, int32_t, int32_t, int64_t, int32_t*, void*) {
  return alad::fail(ALAD_ERR_UNSUPPORTED, "peer memory needs CUDA devices");
}

#endif
