// alad_h2d_2d_staged: pitched host->device upload of PAGEABLE sources through pinned staging buffers filled by several
// host threads.  The reference's encode_data (alad/evaluation.py:98-130) hands i2t / t2i ordinary (pageable) CPU tensors;
// a cudaMemcpy2DAsync from pageable memory is staged by the driver on ONE thread (~10 GB/s on the B200 boxes, which made
// the end-to-end COCO-5k evaluation upload-bound: 625 ms against 337 ms from pinned tensors).  Here n_threads workers
// gather the rows into a ring of pinned buffers while the copy engine drains the previous buffer.
// Replaces the per-query `.cuda()` of alad/evaluation.py:179,202,267,291 for pageable galleries.
#include <stdint.h>
#include <string.h>

#include <condition_variable>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#include "common.h"

namespace alad {
namespace {

// minimal persistent pool: parallel_for(n, fn) runs fn(i) for i < n on the workers (and the caller)
class HostPool {
 public:
  explicit HostPool(int n) : n_(n > 1 ? n - 1 : 0) {
    for (int i = 0; i < n_; ++i) workers_.emplace_back([this] { loop(); });
  }
  ~HostPool() {
    {
      std::lock_guard<std::mutex> l(m_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return n_ + 1; }
  void parallel_for(int n, const std::function<void(int)>& fn) {
    {
      std::lock_guard<std::mutex> l(m_);
      fn_ = &fn;
      next_ = 0;
      total_ = n;
      pending_ = n;
      ++gen_;
    }
    cv_.notify_all();
    run();                                         // the caller takes part
    std::unique_lock<std::mutex> l(m_);
    done_.wait(l, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void run() {
    for (;;) {
      int i;
      const std::function<void(int)>* fn;
      {
        std::lock_guard<std::mutex> l(m_);
        if (!fn_ || next_ >= total_) return;
        i = next_++;
        fn = fn_;
      }
      (*fn)(i);
      std::lock_guard<std::mutex> l(m_);
      if (--pending_ == 0) done_.notify_all();
    }
  }
  void loop() {
    unsigned long long seen = 0;
    for (;;) {
      {
        std::unique_lock<std::mutex> l(m_);
        cv_.wait(l, [&] { return stop_ || gen_ != seen; });
        if (stop_) return;
        seen = gen_;
      }
      run();
    }
  }
  int n_;
  std::vector<std::thread> workers_;
  std::mutex m_;
  std::condition_variable cv_, done_;
  const std::function<void(int)>* fn_ = nullptr;
  int next_ = 0, total_ = 0, pending_ = 0;
  unsigned long long gen_ = 0;
  bool stop_ = false;
};

#ifndef ALAD_CPU_EMU
constexpr int kBuffers = 3;
constexpr size_t kBufBytes = 48u << 20;

struct Staging {
  std::mutex m;
  void* buf[kBuffers] = {nullptr, nullptr, nullptr};
  cudaEvent_t ev[kBuffers];
  bool used[kBuffers] = {false, false, false};
  int next = 0;
  HostPool* pool = nullptr;
};
// one set of staging buffers and events per device (events belong to the device they were created on)
Staging& staging(int dev) {
  static Staging s[16];
  return s[dev & 15];
}
#endif

}  // namespace
}  // namespace alad

extern "C" int alad_h2d_2d_staged(void* dst, int64_t dst_pitch, const void* src_host, int64_t src_pitch, int64_t width_bytes,
                                  int64_t height, int32_t n_threads, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(width_bytes >= 0 && height >= 0 && dst_pitch >= width_bytes && src_pitch >= width_bytes && n_threads >= 1,
               "alad_h2d_2d_staged: bad geometry");
  if (width_bytes == 0 || height == 0) return ALAD_OK;
  ALAD_REQUIRE(dst && src_host, "alad_h2d_2d_staged: NULL pointer");
#ifdef ALAD_CPU_EMU
  for (int64_t r = 0; r < height; ++r)
    memcpy(static_cast<char*>(dst) + r * dst_pitch, static_cast<const char*>(src_host) + r * src_pitch, (size_t)width_bytes);
  return ALAD_OK;
#else
  if ((size_t)width_bytes > kBufBytes)             // rows larger than a staging buffer: let the driver stage them
    return alad_h2d_2d(dst, dst_pitch, src_host, src_pitch, width_bytes, height, stream);
  int dev = 0;
  ALAD_CUDA(cudaGetDevice(&dev));
  Staging& S = staging(dev);
  std::lock_guard<std::mutex> lock(S.m);
  if (!S.pool || S.pool->size() != n_threads) {
    delete S.pool;
    S.pool = new HostPool(n_threads);
  }
  for (int b = 0; b < kBuffers; ++b)
    if (!S.buf[b]) {
      ALAD_CUDA(cudaHostAlloc(&S.buf[b], kBufBytes, cudaHostAllocPortable));
      ALAD_CUDA(cudaEventCreateWithFlags(&S.ev[b], cudaEventDisableTiming));
    }
  cudaStream_t st = as_stream(stream);
  const int64_t rows_per_buf = (int64_t)(kBufBytes / (size_t)width_bytes);
  const char* src = static_cast<const char*>(src_host);
  char* out = static_cast<char*>(dst);
  for (int64_t r0 = 0; r0 < height; r0 += rows_per_buf) {
    const int64_t rows = height - r0 < rows_per_buf ? height - r0 : rows_per_buf;
    const int b = S.next;
    S.next = (S.next + 1) % kBuffers;
    if (S.used[b]) ALAD_CUDA(cudaEventSynchronize(S.ev[b]));      // the DMA that last read this buffer has finished
    char* stage = static_cast<char*>(S.buf[b]);
    // gather: split the rows into ~4 pieces per thread so that the threads stay balanced
    const int pieces = (int)(rows < 4 * n_threads ? rows : 4 * n_threads);
    S.pool->parallel_for(pieces, [&](int p) {
      const int64_t a = rows * p / pieces, e = rows * (p + 1) / pieces;
      if (src_pitch == width_bytes) {
        memcpy(stage + a * width_bytes, src + (r0 + a) * src_pitch, (size_t)((e - a) * width_bytes));
      } else {
        for (int64_t r = a; r < e; ++r) memcpy(stage + r * width_bytes, src + (r0 + r) * src_pitch, (size_t)width_bytes);
      }
    });
    ALAD_CUDA(cudaMemcpy2DAsync(out + r0 * dst_pitch, (size_t)dst_pitch, stage, (size_t)width_bytes, (size_t)width_bytes,
                                (size_t)rows, cudaMemcpyHostToDevice, st));
    ALAD_CUDA(cudaEventRecord(S.ev[b], st));
    S.used[b] = true;
  }
  return ALAD_OK;
#endif
}
