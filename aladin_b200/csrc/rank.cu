// Ranking kernels: replace numpy.argsort + numpy.where of alad/evaluation.py:213-223 and
// 303-308 (and alad/recall_auxiliary.py:34-56).  A full sort is never needed:
//   rank  = number of gallery items ordered ahead of the ground truth,
//   top-k = running k-best per query.
// Total order: score descending, index descending on exact ties (what a stable argsort
// followed by [::-1] yields).  All HBM-bound streaming reads of the score matrix.
#include <math.h>

#include "common.h"

namespace alad {

__device__ __forceinline__ bool ahead(float v, int vi, float w, int wi) {  // (v,vi) ordered before (w,wi)?
  return v > w || (v == w && vi > wi);
}

// ------------------------------------------------------------------ i2t: one CTA per image row
constexpr int RR_THREADS = 256;

__global__ void __launch_bounds__(RR_THREADS)
rank_rows_kernel(const float* __restrict__ S, long long ldS, int Nc, int group, int img_off, int* __restrict__ rank,
                 int* __restrict__ top1) {
  const int i = blockIdx.x;
  const float* row = S + (long long)i * ldS;
  // best ground-truth caption of this image under the total order
  float gs = -INFINITY;
  int gi = -1;
  const long long g0 = (long long)group * (img_off + i);
  for (int g = 0; g < group; ++g) {
    const long long c = g0 + g;
    if (c < Nc) {
      const float v = __ldg(row + c);
      if (gi < 0 || ahead(v, (int)c, gs, gi)) {
        gs = v;
        gi = (int)c;
      }
    }
  }
  int cnt = 0;
  float ts = -INFINITY;
  int ti = -1;
  auto visit = [&](float v, int c) {
    cnt += (gi >= 0 && ahead(v, c, gs, gi)) ? 1 : 0;
    if (ti < 0 || ahead(v, c, ts, ti)) {
      ts = v;
      ti = c;
    }
  };
  const bool vec = ((ldS & 3) == 0) && ((reinterpret_cast<uintptr_t>(S) & 15) == 0);
  if (vec) {
    const float4* row4 = reinterpret_cast<const float4*>(row);
    const int n4 = Nc >> 2;
    for (int q = threadIdx.x; q < n4; q += RR_THREADS) {
      const float4 v = __ldg(row4 + q);
      visit(v.x, 4 * q);
      visit(v.y, 4 * q + 1);
      visit(v.z, 4 * q + 2);
      visit(v.w, 4 * q + 3);
    }
    for (int c = (n4 << 2) + threadIdx.x; c < Nc; c += RR_THREADS) visit(__ldg(row + c), c);
  } else {
    for (int c = threadIdx.x; c < Nc; c += RR_THREADS) visit(__ldg(row + c), c);
  }
  // block reduction: sum of counts, arg-best of (ts, ti)
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    const float os = __shfl_xor_sync(0xffffffffu, ts, o);
    const int oi = __shfl_xor_sync(0xffffffffu, ti, o);
    if (oi >= 0 && (ti < 0 || ahead(os, oi, ts, ti))) {
      ts = os;
      ti = oi;
    }
  }
  __shared__ int s_cnt[RR_THREADS / 32];
  __shared__ float s_ts[RR_THREADS / 32];
  __shared__ int s_ti[RR_THREADS / 32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    s_cnt[warp] = cnt;
    s_ts[warp] = ts;
    s_ti[warp] = ti;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int c = 0;
    float bs = -INFINITY;
    int bi = -1;
    for (int w = 0; w < RR_THREADS / 32; ++w) {
      c += s_cnt[w];
      if (s_ti[w] >= 0 && (bi < 0 || ahead(s_ts[w], s_ti[w], bs, bi))) {
        bs = s_ts[w];
        bi = s_ti[w];
      }
    }
    rank[i] = (gi >= 0) ? c : Nc;
    top1[i] = bi;
  }
}

// ------------------------------------------------------------------ t2i step 1: ground-truth score per caption
__global__ void col_gt_kernel(const float* __restrict__ S, long long ldS, int Ni, int Nc, int group, int img_off,
                              float* __restrict__ gt) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Nc) return;
  const int img = c / group - img_off;
  if (img >= 0 && img < Ni) gt[c] = __ldg(S + (long long)img * ldS + c);
}

// ------------------------------------------------------------------ t2i step 2: images ahead of the ground truth
constexpr int CC_ROWS = 8;      // row slices per CTA (threadIdx.y)
constexpr int CC_CHUNK = 512;   // rows per CTA along grid.y

__global__ void __launch_bounds__(32 * CC_ROWS)
col_count_kernel(const float* __restrict__ S, long long ldS, int Ni, int Nc, int group, int img_off,
                 const float* __restrict__ gt, int* __restrict__ count) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int r_begin = blockIdx.y * CC_CHUNK;
  const int r_end = min(Ni, r_begin + CC_CHUNK);
  int cnt = 0;
  if (c < Nc) {
    const float g = __ldg(gt + c);
    const int gimg = c / group;                       // global index of the ground-truth image
    for (int r = r_begin + threadIdx.y; r < r_end; r += CC_ROWS) {
      const float v = __ldg(S + (long long)r * ldS + c);
      cnt += ahead(v, img_off + r, g, gimg) ? 1 : 0;
    }
  }
  __shared__ int part[CC_ROWS][32];
  part[threadIdx.y][threadIdx.x] = cnt;
  __syncthreads();
  if (threadIdx.y == 0 && c < Nc) {
    int tot = 0;
#pragma unroll
    for (int y = 0; y < CC_ROWS; ++y) tot += part[y][threadIdx.x];
    if (tot) atomicAdd(count + c, tot);
  }
}

// ------------------------------------------------------------------ t2i step 3: running top-k per caption
// One warp per (32 captions, row slice); thread <-> caption (coalesced 128 B row reads).
// The k best entries of every caption live in a binary MIN-heap (root = worst kept entry) in
// shared memory, laid out [k][32] so that lane == bank whatever node a lane touches.  An insertion
// costs <= log2(k) levels instead of a k-entry rescan; the final order comes from k heap pops.
__device__ __forceinline__ void heap_sift_down(float* bs, int* bi, int k, int lane, float v, int vi) {
  int j = 0;
  while (true) {
    const int l = 2 * j + 1;
    if (l >= k) break;
    float cs = bs[l * 32 + lane];
    int ci = bi[l * 32 + lane];
    int cj = l;
    if (l + 1 < k) {
      const float rs = bs[(l + 1) * 32 + lane];
      const int ri = bi[(l + 1) * 32 + lane];
      if (ahead(cs, ci, rs, ri)) {          // the right child is the worse one
        cs = rs;
        ci = ri;
        cj = l + 1;
      }
    }
    if (!ahead(v, vi, cs, ci)) break;       // (v, vi) is not better than the worse child: it stays above
    bs[j * 32 + lane] = cs;
    bi[j * 32 + lane] = ci;
    j = cj;
  }
  bs[j * 32 + lane] = v;
  bi[j * 32 + lane] = vi;
}

__global__ void __launch_bounds__(32)
col_topk_kernel(const float* __restrict__ S, long long ldS, int Ni, int Nc, int k, int img_off, int splits,
                float* __restrict__ cand_score, int* __restrict__ cand_idx) {
  extern __shared__ uint8_t topk_smem[];
  float* bs = reinterpret_cast<float*>(topk_smem);       // [k][32]
  int* bi = reinterpret_cast<int*>(bs + k * 32);           // [k][32]
  const int lane = threadIdx.x;
  const int c = blockIdx.x * 32 + lane;
  const int split = blockIdx.y;
  const int per = (Ni + splits - 1) / splits;
  const int r_begin = split * per;
  const int r_end = min(Ni, r_begin + per);
  for (int j = 0; j < k; ++j) {             // k "empty" entries: every real entry is ahead of (-inf, -1)
    bs[j * 32 + lane] = -INFINITY;
    bi[j * 32 + lane] = -1;
  }
  float ws = -INFINITY;   // root of the heap = worst entry currently kept
  int wi = -1;
  const bool col_ok = c < Nc;
  const float* col = S + (col_ok ? c : 0);
  for (int r = r_begin; r < r_end; r += 4) {
    float v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) v[u] = (col_ok && r + u < r_end) ? __ldg(col + (long long)(r + u) * ldS) : 0.f;
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int gi = img_off + r + u;
      if (col_ok && r + u < r_end && ahead(v[u], gi, ws, wi)) {
        heap_sift_down(bs, bi, k, lane, v[u], gi);          // replaces the root
        ws = bs[lane];
        wi = bi[lane];
      }
    }
  }
  if (!col_ok) return;
  // heap sort: pop the worst entry k times, filling the output from the back (best first)
  float* out_s = cand_score + ((long long)split * Nc + c) * k;
  int* out_i = cand_idx + ((long long)split * Nc + c) * k;
  for (int n = k; n > 0; --n) {
    out_s[n - 1] = bs[lane];
    out_i[n - 1] = bi[lane];
    const float ls = bs[(n - 1) * 32 + lane];
    const int li = bi[(n - 1) * 32 + lane];
    if (n > 1) heap_sift_down(bs, bi, n - 1, lane, ls, li);
  }
}

// ------------------------------------------------------------------ merge P sorted candidate lists per caption
constexpr int MERGE_MAX_P = 64;

__global__ void topk_merge_kernel(const float* __restrict__ cand_score, const int* __restrict__ cand_idx, int P, int Nc,
                                  int k, float* __restrict__ out_score, int* __restrict__ out_idx) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Nc) return;
  unsigned char head[MERGE_MAX_P];
  for (int p = 0; p < P; ++p) head[p] = 0;
  for (int o = 0; o < k; ++o) {
    float s0 = -INFINITY;
    int i0 = -2;
    int p0 = -1;
    for (int p = 0; p < P; ++p) {
      if (head[p] >= k) continue;
      const long long at = ((long long)p * Nc + c) * k + head[p];
      const float s = __ldg(cand_score + at);
      const int ii = __ldg(cand_idx + at);
      if (p0 < 0 || ahead(s, ii, s0, i0)) {
        s0 = s;
        i0 = ii;
        p0 = p;
      }
    }
    if (p0 >= 0) head[p0]++;
    out_score[(long long)c * k + o] = (p0 >= 0) ? s0 : -INFINITY;
    if (out_idx) out_idx[(long long)c * k + o] = (p0 >= 0) ? i0 : -1;
  }
}

// ------------------------------------------------------------------ two-stage retrieval: keep only shortlisted pairs
__global__ void fill_kernel(float* __restrict__ x, long long ld, int rows, int cols, float v) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j < cols) x[(long long)blockIdx.y * ld + j] = v;
}
// list q holds up to k gallery indices for query q; by_column: q = caption (column), entries = global image
// indices; otherwise q = local image (row), entries = caption indices.
__global__ void shortlist_scatter_kernel(const float* __restrict__ S, long long ldS, float* __restrict__ S2, long long ld2,
                                         int Ni, int Nc, const int* __restrict__ idx, int n_lists, int k, int by_column,
                                         int img_off) {
  const long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (e >= (long long)n_lists * k) return;
  const int q = (int)(e / k);
  const int g = idx[e];
  if (g < 0) return;
  int i, c;
  if (by_column) {
    i = g - img_off;
    c = q;
  } else {
    i = q;
    c = g;
  }
  if (i < 0 || i >= Ni || c < 0 || c >= Nc) return;
  S2[(long long)i * ld2 + c] = S[(long long)i * ldS + c];
}

}  // namespace alad

extern "C" int alad_shortlist_scatter(const float* S, int64_t ldS, float* S2, int64_t ld2, int32_t Ni, int32_t Nc,
                                      const int32_t* idx, int32_t n_lists, int32_t k, int32_t by_column, int32_t img_off,
                                      void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Ni >= 0 && Nc >= 0 && ldS >= Nc && ld2 >= Nc && k > 0 && n_lists >= 0 && Ni <= 65535 * 1,
               "alad_shortlist_scatter: bad shape");
  if (Ni == 0 || Nc == 0) return ALAD_OK;
  ALAD_REQUIRE(S && S2 && (idx || n_lists == 0), "alad_shortlist_scatter: NULL pointer");
  cudaStream_t st = as_stream(stream);
  dim3 grid((Nc + 255) / 256, Ni);
  fill_kernel<<<grid, 256, 0, st>>>(S2, ld2, Ni, Nc, -INFINITY);
  const long long n = (long long)n_lists * k;
  if (n) shortlist_scatter_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(S, ldS, S2, ld2, Ni, Nc, idx, n_lists, k,
                                                                                by_column, img_off);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_rank_rows(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, int32_t group, int32_t img_off,
                              int32_t* rank, int32_t* top1, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Ni >= 0 && Nc >= 0 && ldS >= Nc && group > 0, "alad_rank_rows: bad shape");
  if (Ni == 0) return ALAD_OK;
  ALAD_REQUIRE(S && rank && top1, "alad_rank_rows: NULL pointer");
  rank_rows_kernel<<<Ni, RR_THREADS, 0, as_stream(stream)>>>(S, ldS, Nc, group, img_off, rank, top1);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_col_gt(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, int32_t group, int32_t img_off,
                           float* gt, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Ni >= 0 && Nc >= 0 && ldS >= Nc && group > 0, "alad_col_gt: bad shape");
  if (Ni == 0 || Nc == 0) return ALAD_OK;
  ALAD_REQUIRE(S && gt, "alad_col_gt: NULL pointer");
  col_gt_kernel<<<(Nc + 255) / 256, 256, 0, as_stream(stream)>>>(S, ldS, Ni, Nc, group, img_off, gt);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_col_count(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, int32_t group, int32_t img_off,
                              const float* gt, int32_t* count, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Ni >= 0 && Nc >= 0 && ldS >= Nc && group > 0, "alad_col_count: bad shape");
  if (Nc == 0) return ALAD_OK;
  ALAD_REQUIRE(count, "alad_col_count: NULL pointer");
  cudaStream_t st = as_stream(stream);
  ALAD_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t) * (size_t)Nc, st));
  if (Ni == 0) return ALAD_OK;
  ALAD_REQUIRE(S && gt, "alad_col_count: NULL pointer");
  dim3 grid((Nc + 31) / 32, (Ni + CC_CHUNK - 1) / CC_CHUNK), block(32, CC_ROWS);
  col_count_kernel<<<grid, block, 0, st>>>(S, ldS, Ni, Nc, group, img_off, gt, count);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_col_topk(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, int32_t k, int32_t img_off,
                             int32_t splits, float* cand_score, int32_t* cand_idx, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Ni >= 0 && Nc >= 0 && ldS >= Nc, "alad_col_topk: bad shape");
  ALAD_REQUIRE(k > 0 && k <= 256 && splits > 0 && splits <= 65535, "alad_col_topk: k=%d splits=%d out of range", k, splits);
  if (Nc == 0) return ALAD_OK;
  ALAD_REQUIRE(cand_score && cand_idx && (S || Ni == 0), "alad_col_topk: NULL pointer");
  const size_t smem = (size_t)k * 32 * 8;
  static thread_local size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    ALAD_CUDA(cudaFuncSetAttribute(col_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  dim3 grid((Nc + 31) / 32, splits);
  col_topk_kernel<<<grid, 32, smem, as_stream(stream)>>>(S, ldS, Ni, Nc, k, img_off, splits, cand_score, cand_idx);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_topk_merge(const float* cand_score, const int32_t* cand_idx, int32_t P, int32_t Nc, int32_t k,
                               float* out_score, int32_t* out_idx, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(P > 0 && P <= MERGE_MAX_P && k > 0 && k <= 255 && Nc >= 0, "alad_topk_merge: P=%d k=%d out of range", P, k);
  if (Nc == 0) return ALAD_OK;
  ALAD_REQUIRE(cand_score && cand_idx && out_score, "alad_topk_merge: NULL pointer");
  topk_merge_kernel<<<(Nc + 127) / 128, 128, 0, as_stream(stream)>>>(cand_score, cand_idx, P, Nc, k, out_score, out_idx);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}
