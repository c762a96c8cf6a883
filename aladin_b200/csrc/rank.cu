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

// col_list / n_list (optional): the kernel visits only the listed columns (the overflow columns of the
// threshold-select path below); CTAs beyond the list exit at once.
__global__ void __launch_bounds__(32)
col_topk_kernel(const float* __restrict__ S, long long ldS, int Ni, int Nc, int k, int img_off, int splits,
                float* __restrict__ cand_score, int* __restrict__ cand_idx, const int* __restrict__ col_list,
                const int* __restrict__ n_list) {
  extern __shared__ uint8_t topk_smem[];
  float* bs = reinterpret_cast<float*>(topk_smem);       // [k][32]
  int* bi = reinterpret_cast<int*>(bs + k * 32);           // [k][32]
  const int lane = threadIdx.x;
  int c = blockIdx.x * 32 + lane;
  if (col_list) {
    const int n = *n_list;
    if ((int)blockIdx.x * 32 >= n) return;
    c = c < n ? col_list[c] : Nc;
  }
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

// ------------------------------------------------------------------ t2i step 3, threshold-select variant
// Exact top-k per caption in two HBM sweeps of S instead of k-entry heaps:
//   (1) col_groupmax: the rows are cut into G >= k groups; gmax[g][c] = max of column c over group g.
//   (2) col_threshold: tau[c] = k-th largest of the G group maxima (radix select on order-preserving keys).
//       At least k entries of the column are >= tau[c] (one per qualifying group), so every top-k entry is.
//   (3) col_collect: second sweep; entries >= tau[c] are appended to a per-caption candidate list
//       (expected ~1.3 k entries for G = 2.5 k groups, capacity `cap`).
//   (4) col_select: one warp per caption ranks its candidates under the total order (score desc, index
//       desc) and writes the k best in order.  Captions whose list overflowed (mass ties, e.g. an all-zero
//       column) are queued for the heap kernel above -- exact in every case.
constexpr int SEL_ROWS = 8;           // row slices per CTA (threadIdx.y)
constexpr int SEL_CHUNK = 512;        // rows per collect CTA
constexpr int SEL_MAX_G = 352;        // group-maxima tile [G][33] floats stays under 48 KB
constexpr int SEL_MAX_CAP = 512;

__device__ __forceinline__ uint32_t order_key(float v) {       // monotone float -> uint (-0 == +0)
  const uint32_t u = __float_as_uint(v + 0.f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float key_to_float(uint32_t key) {
  return __uint_as_float((key & 0x80000000u) ? (key & 0x7fffffffu) : ~key);
}

__global__ void __launch_bounds__(32 * SEL_ROWS)
col_groupmax_kernel(const float* __restrict__ S, long long ldS, int Ni, int Nc, int rows_per_group, int G,
                    float* __restrict__ gmax) {
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int g = blockIdx.y * SEL_ROWS + threadIdx.y;
  if (c >= Nc || g >= G) return;
  const int r0 = g * rows_per_group;
  const int r1 = min(Ni, r0 + rows_per_group);
  const float* col = S + c;
  float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
  int r = r0;
  for (; r + 4 <= r1; r += 4) {
    m0 = fmaxf(m0, __ldg(col + (long long)(r + 0) * ldS));
    m1 = fmaxf(m1, __ldg(col + (long long)(r + 1) * ldS));
    m2 = fmaxf(m2, __ldg(col + (long long)(r + 2) * ldS));
    m3 = fmaxf(m3, __ldg(col + (long long)(r + 3) * ldS));
  }
  for (; r < r1; ++r) m0 = fmaxf(m0, __ldg(col + (long long)r * ldS));
  gmax[(long long)g * Nc + c] = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
}

// ------------------------------------------------------------------ both directions in ONE sweep of S
// rank_sweep_kernel replaces rank_rows + col_count + col_groupmax (three reads of S) by one: a CTA owns one row
// group of the threshold-select plan x FS_COLS columns; every thread keeps 8 columns in registers (two float4 per
// row: a warp reads 512 contiguous bytes per load) and carries their running maximum and "ahead of the ground
// truth" count down the rows, while the row statistics (entries ahead of the row's best ground-truth caption, row
// arg-max) are reduced across the warp with redux.sync and collected per CTA in shared memory.  Cross-CTA merges
// are integer atomics (adds, and a 64-bit max of (order key, column) = the total order's arg-best), so the results
// do not depend on the order CTAs finish in.  Exact ties take a rare slow branch (index comparison).
constexpr int FS_THREADS = 256;
#ifndef ALAD_FS_VEC
#define ALAD_FS_VEC 2
#endif
constexpr int FS_VEC = ALAD_FS_VEC;                // float4 per thread and row
constexpr int FS_CTAS = FS_VEC <= 2 ? 4 : 2;       // resident CTAs per SM the register budget is held to
constexpr int FS_E = 4 * FS_VEC;                   // columns per thread
constexpr int FS_COLS = FS_THREADS * FS_E;         // columns per CTA
constexpr int FS_ROWS = 64;                        // rows per block of shared row statistics

// largest float below a finite g (g + 0.f: no -0): "v >= g" as the strict comparison "v > below(g)"
__device__ __forceinline__ float float_below(float g) {
  const uint32_t u = __float_as_uint(g);
  if (u == 0u) return __uint_as_float(0x80000001u);
  return __uint_as_float((u & 0x80000000u) ? u + 1u : u - 1u);
}
__device__ __forceinline__ bool is_finite(float x) { return fabsf(x) < INFINITY; }
// c += (v > t) as one compare and one predicated add (the compiler's own lowering of the C expression is a
// three-instruction add / select / move chain per element, which makes the sweep issue-bound)
__device__ __forceinline__ void count_gt(int& c, float v, float t) {
#ifdef ALAD_CPU_EMU
  c += v > t ? 1 : 0;
#else
  asm("{\n\t.reg .pred p;\n\tsetp.gt.f32 p, %1, %2;\n\t@p add.s32 %0, %0, 1;\n\t}" : "+r"(c) : "f"(v), "f"(t));
#endif
}

// The total order's tie rule (equal scores: the higher index is ahead) is folded into the threshold wherever a whole
// CTA / thread sits on one side of the ground truth: entries after it compare against below(g) (>= g), entries before
// it against g itself.  Only the threads whose columns (rows) straddle the ground truth, and non-finite ground-truth
// scores (-inf of masked matrices), take the exact comparison.
struct __align__(16) RowQuery {
  float gs;          // the row's best ground-truth score (strict threshold) ...
  float ge;          // ... and the value just below it (ties count)
  int gi;            // its caption index; -1 none
  int mode;          // 0 not a query row, 1 thresholds, 2 exact comparison (non-finite ground-truth score)
};

template <bool COUNT>
__global__ void __launch_bounds__(FS_THREADS, FS_CTAS)
rank_sweep_kernel(const float* __restrict__ S, long long ldS, int Ni, int Nc, int group, int img_off, int q_rows,
                  int q_cols, int rows_per_group, const float* __restrict__ gt, int* __restrict__ rank,
                  unsigned long long* __restrict__ best, int* __restrict__ count, float* __restrict__ gmax) {
  // per-warp slots instead of shared-memory atomics: every warp writes its share of every query row once
  __shared__ RowQuery s_row[FS_ROWS];
  __shared__ int s_cnt[FS_THREADS / 32][FS_ROWS];
  __shared__ unsigned long long s_best[FS_THREADS / 32][FS_ROWS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = blockIdx.y;
  const int r0 = g * rows_per_group, r1 = min(Ni, r0 + rows_per_group);
  const int cbase = blockIdx.x * FS_COLS + 4 * threadIdx.x;
  auto col = [&](int e) { return cbase + (e >> 2) * (4 * FS_THREADS) + (e & 3); };
  const int c_first = col(0), c_last = col(FS_E - 1);
  float cmax[FS_E], cthr[FS_E];
  int ccnt[FS_E];
  bool c_exact = false;
#pragma unroll
  for (int e = 0; e < FS_E; ++e) {
    const int c = col(e);
    cmax[e] = -INFINITY;
    ccnt[e] = 0;
    cthr[e] = INFINITY;                              // non-query columns: nothing is ahead
    if (COUNT && c < q_cols) {
      const float gv = __ldg(gt + c) + 0.f;
      const int gl = c / group - img_off;            // block-local row of the caption's image
      if (!is_finite(gv) || (gl >= r0 && gl < r1 - 1)) c_exact = true;
      else cthr[e] = gl < r0 ? float_below(gv) : gv;
    }
  }
  // 0: the float4 lies beyond the row, 1: inside, 2: across its end
  int piece[FS_VEC];
#pragma unroll
  for (int u = 0; u < FS_VEC; ++u) {
    const int c4 = cbase + u * (4 * FS_THREADS);
    piece[u] = c4 + 3 < Nc ? 1 : (c4 < Nc ? 2 : 0);
  }
  auto load_row = [&](const float* rowp, float* v) {  // rowp = this thread's first column of the row
#pragma unroll
    for (int u = 0; u < FS_VEC; ++u) {
      float4 x = make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      const float* q = rowp + u * (4 * FS_THREADS);
      if (piece[u] == 1) x = __ldg(reinterpret_cast<const float4*>(q));
      else if (piece[u] == 2) {
        const int c4 = cbase + u * (4 * FS_THREADS);
        x.x = __ldg(q);
        if (c4 + 1 < Nc) x.y = __ldg(q + 1);
        if (c4 + 2 < Nc) x.z = __ldg(q + 2);
      }
      v[4 * u] = x.x; v[4 * u + 1] = x.y; v[4 * u + 2] = x.z; v[4 * u + 3] = x.w;
    }
  };
  auto consume = [&](int r, int rr, const float* v) {
    // ---- columns: running maximum, images ahead of the caption's ground truth
#pragma unroll
    for (int e = 0; e < FS_E; ++e) cmax[e] = fmaxf(cmax[e], v[e]);
    if (COUNT) {
      if (!c_exact) {
#pragma unroll
        for (int e = 0; e < FS_E; ++e) count_gt(ccnt[e], v[e], cthr[e]);
      } else {
#pragma unroll
        for (int e = 0; e < FS_E; ++e) {
          const int c = col(e);
          if (c < q_cols) ccnt[e] += ahead(v[e], img_off + r, __ldg(gt + c), c / group) ? 1 : 0;
        }
      }
    }
    // ---- row: captions ahead of the image's best ground truth, arg-max caption
    const RowQuery q = s_row[rr];
    if (q.mode == 0) return;                         // uniform over the CTA
    int cnt = 0;
    float tmax = -INFINITY;
#pragma unroll
    for (int e = 0; e < FS_E; ++e) tmax = fmaxf(tmax, v[e]);
    if (q.mode == 1 && (q.gi < c_first || q.gi >= c_last)) {
      const float thr = q.gi < c_first ? q.ge : q.gs;
#pragma unroll
      for (int e = 0; e < FS_E; ++e) count_gt(cnt, v[e], thr);
    } else {
#pragma unroll
      for (int e = 0; e < FS_E; ++e) cnt += (col(e) < Nc && ahead(v[e], col(e), q.gs, q.gi)) ? 1 : 0;
    }
    cnt = __reduce_add_sync(0xffffffffu, cnt);
    const unsigned key = order_key(tmax);
    const unsigned mk = __reduce_max_sync(0xffffffffu, key);
    int idx = -1;
    if (key == mk) {                                 // normally one lane of the warp
      int at = -1;                                   // slots ascend with the column: the largest index stays
      if (tmax > -INFINITY) {                        // columns beyond the row hold -inf: no bounds check needed
#pragma unroll
        for (int e = 0; e < FS_E; ++e) at = v[e] == tmax ? e : at;
      } else {
#pragma unroll
        for (int e = 0; e < FS_E; ++e) at = (v[e] == tmax && col(e) < Nc) ? e : at;
      }
      if (at >= 0) idx = cbase + (at >> 2) * (4 * FS_THREADS) + (at & 3);
    }
    idx = __reduce_max_sync(0xffffffffu, idx);
    if (lane == 0) {
      s_cnt[warp][rr] = cnt;
      s_best[warp][rr] = idx >= 0 ? (((unsigned long long)mk << 32) | (unsigned)idx) : 0ull;
    }
  };
  for (int rb = r0; rb < r1; rb += FS_ROWS) {
    const int nrows = min(FS_ROWS, r1 - rb);
    if ((int)threadIdx.x < nrows) {
      const int r = rb + threadIdx.x;
      RowQuery q;
      q.gs = INFINITY;
      q.gi = -1;
      q.mode = 0;
      if (r < q_rows) {
        const long long g0 = (long long)group * (img_off + r);
        for (int j = 0; j < group; ++j) {
          const long long c = g0 + j;
          if (c < Nc) {
            const float v = __ldg(S + (long long)r * ldS + c);
            if (q.gi < 0 || ahead(v, (int)c, q.gs, q.gi)) {
              q.gs = v;
              q.gi = (int)c;
            }
          }
        }
        if (q.gi < 0) q.gs = INFINITY;               // no ground truth in range: nothing is ahead, rank = Nc
        q.gs += 0.f;
        q.mode = (q.gi < 0 || is_finite(q.gs)) ? 1 : 2;
      }
      q.ge = is_finite(q.gs) ? float_below(q.gs) : q.gs;
      s_row[threadIdx.x] = q;
    }
    __syncthreads();
    float va[FS_E], vb[FS_E];                        // two rows in flight
    const float* rowp = S + (long long)rb * ldS + cbase;
    load_row(rowp, va);
    for (int rr = 0; rr < nrows; rr += 2) {
      if (rr + 1 < nrows) load_row(rowp + ldS, vb);
      consume(rb + rr, rr, va);
      if (rr + 2 < nrows) load_row(rowp + 2 * ldS, va);
      if (rr + 1 < nrows) consume(rb + rr + 1, rr + 1, vb);
      rowp += 2 * ldS;
    }
    __syncthreads();
    if ((int)threadIdx.x < nrows && s_row[threadIdx.x].mode != 0) {
      const int r = rb + threadIdx.x;
      int add = 0;
      unsigned long long top = 0ull;
#pragma unroll
      for (int w = 0; w < FS_THREADS / 32; ++w) {
        add += s_cnt[w][threadIdx.x];
        const unsigned long long b = s_best[w][threadIdx.x];
        top = b > top ? b : top;
      }
      if (s_row[threadIdx.x].gi < 0) add = blockIdx.x == 0 ? Nc : 0;
      if (add) atomicAdd(rank + r, add);
      if (top) atomicMax(best + r, top);
    }
    __syncthreads();
  }
#pragma unroll
  for (int e = 0; e < FS_E; ++e) {
    const int c = col(e);
    if (c < q_cols) {
      gmax[(long long)g * q_cols + c] = cmax[e];
      if (COUNT && ccnt[e]) atomicAdd(count + c, ccnt[e]);
    }
  }
}

__global__ void rank_sweep_finish_kernel(const unsigned long long* __restrict__ best, int q_rows, int* __restrict__ top1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < q_rows) top1[i] = best[i] ? (int)(unsigned)(best[i] & 0xffffffffull) : -1;
}

// 8 warps per CTA, 32 captions per CTA (4 per warp); the [G][32] tile of group maxima goes through shared
// memory so that the global reads stay coalesced
template <int PER_LANE>                                        // keys per lane: ceil(G / 32) rounded up to 4, 8 or 11
__global__ void __launch_bounds__(256)
col_threshold_kernel(const float* __restrict__ gmax, int Nc, int G, int k, float* __restrict__ tau) {
  extern __shared__ float tile[];                              // [G][33]
  const int c0 = blockIdx.x * 32;
  for (int e = threadIdx.x; e < G * 32; e += 256) {
    const int g = e >> 5, x = e & 31;
    tile[g * 33 + x] = (c0 + x < Nc) ? __ldg(gmax + (long long)g * Nc + c0 + x) : 0.f;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int q = 0; q < 4; ++q) {
    const int x = warp * 4 + q;
    if (c0 + x >= Nc) break;                                   // warp-uniform
    uint32_t key[PER_LANE];
#pragma unroll
    for (int j = 0; j < PER_LANE; ++j) {
      const int g = lane + 32 * j;
      key[j] = g < G ? order_key(tile[g * 33 + x]) : 0u;       // 0 is below every real key
    }
    uint32_t T = 0;                                            // largest T with #{key >= T} >= k  ==  k-th largest key
    for (int bit = 31; bit >= 0; --bit) {
      const uint32_t cand = T | (1u << bit);
      int n = 0;
#pragma unroll
      for (int j = 0; j < PER_LANE; ++j) n += key[j] >= cand ? 1 : 0;
      n = __reduce_add_sync(0xffffffffu, n);
      if (n >= k) T = cand;
    }
    if (lane == 0) tau[c0 + x] = key_to_float(T);
  }
}

constexpr int SEL_LOCAL = 24;         // CTA-local list entries per column (expected ~6 per 512-row chunk)
__global__ void __launch_bounds__(32 * SEL_ROWS)
col_collect_kernel(const float* __restrict__ S, long long ldS, int Ni, int Nc, int img_off,
                   const float* __restrict__ tau, int cap, int* __restrict__ cnt, float* __restrict__ cand_s,
                   int* __restrict__ cand_i) {
  // Candidates are first appended to CTA-local lists (shared-memory atomics), then every column reserves its
  // range of the global list with ONE global atomic per CTA; entries beyond the local capacity go to the
  // global list directly.
  __shared__ float ls[SEL_LOCAL][32];
  __shared__ int li[SEL_LOCAL][32];
  __shared__ int lcnt[32];
  __shared__ int lbase[32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int c = blockIdx.x * 32 + tx;
  if (ty == 0) lcnt[tx] = 0;
  __syncthreads();
  const int r_begin = blockIdx.y * SEL_CHUNK;
  const int r_end = min(Ni, r_begin + SEL_CHUNK);
  const bool col_ok = c < Nc;
  const float t = (tau && col_ok) ? __ldg(tau + c) : -INFINITY;   // no threshold: every entry is a candidate
  const float* col = S + (col_ok ? c : 0);
  float* out_s = cand_s + (long long)c * cap;
  int* out_i = cand_i + (long long)c * cap;
  if (col_ok) {
    for (int r = r_begin + ty; r < r_end; r += 4 * SEL_ROWS) {
      float v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        v[u] = (r + u * SEL_ROWS < r_end) ? __ldg(col + (long long)(r + u * SEL_ROWS) * ldS) : -INFINITY;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (r + u * SEL_ROWS < r_end && v[u] >= t) {
          const int lp = atomicAdd(&lcnt[tx], 1);
          if (lp < SEL_LOCAL) {
            ls[lp][tx] = v[u];
            li[lp][tx] = img_off + r + u * SEL_ROWS;
          } else {
            const int pos = atomicAdd(cnt + c, 1);
            if (pos < cap) {
              out_s[pos] = v[u];
              out_i[pos] = img_off + r + u * SEL_ROWS;
            }
          }
        }
      }
    }
  }
  __syncthreads();
  const int n_loc = min(lcnt[tx], SEL_LOCAL);
  if (ty == 0 && col_ok && n_loc > 0) lbase[tx] = atomicAdd(cnt + c, n_loc);
  __syncthreads();
  if (col_ok) {
    const int base = lbase[tx];
    for (int e = ty; e < n_loc; e += SEL_ROWS) {
      if (base + e < cap) {
        out_s[base + e] = ls[e][tx];
        out_i[base + e] = li[e][tx];
      }
    }
  }
}

// one warp per caption: rank of every candidate = number of candidates ahead of it (strict total order,
// so the ranks are a permutation and the output does not depend on the order the atomics filled the list in)
constexpr int SELECT_WARPS = 8;
__global__ void __launch_bounds__(32 * SELECT_WARPS)
col_select_kernel(const int* __restrict__ cnt, const float* __restrict__ cand_s, const int* __restrict__ cand_i, int Nc,
                  int k, int cap, float* __restrict__ out_score, int* __restrict__ out_idx, int* __restrict__ ovf_list,
                  int* __restrict__ ovf_n) {
  extern __shared__ float2 sel_smem[];                         // [SELECT_WARPS][cap] (score, index bits)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = blockIdx.x * SELECT_WARPS + warp;
  if (c >= Nc) return;
  const int n = __ldg(cnt + c);
  if (n > cap) {                                               // list overflowed: exact heap path for this caption
    if (lane == 0) ovf_list[atomicAdd(ovf_n, 1)] = c;
    return;
  }
  float2* mine = sel_smem + warp * cap;
  const float* cs = cand_s + (long long)c * cap;
  const int* ci = cand_i + (long long)c * cap;
  for (int e = lane; e < n; e += 32) mine[e] = make_float2(__ldg(cs + e), __int_as_float(__ldg(ci + e)));
  __syncwarp();
  float* os = out_score + (long long)c * k;
  int* oi = out_idx + (long long)c * k;
  for (int e = lane; e < n; e += 32) {
    const float v = mine[e].x;
    const int vi = __float_as_int(mine[e].y);
    int rank = 0;
#pragma unroll 4
    for (int o = 0; o < n; ++o) {
      const float2 w = mine[o];                                // broadcast read
      rank += ahead(w.x, __float_as_int(w.y), v, vi) ? 1 : 0;
    }
    if (rank < k) {
      os[rank] = v;
      oi[rank] = vi;
    }
  }
  for (int e = n + lane; e < k; e += 32) {                     // fewer than k gallery images
    os[e] = -INFINITY;
    oi[e] = -1;
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
  col_topk_kernel<<<grid, 32, smem, as_stream(stream)>>>(S, ldS, Ni, Nc, k, img_off, splits, cand_score, cand_idx,
                                                         nullptr, nullptr);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

namespace alad {
struct SelectPlan {
  bool select;          // threshold-select path (else: heaps over all columns)
  bool threshold;       // group maxima + tau (else every entry is a candidate: Ni <= cap)
  int cap, rows_per_group, G;
  size_t off_tau, off_cnt, off_ovf, off_cs, off_ci, off_gmax, bytes;
};
static SelectPlan select_plan(int Ni, int Nc, int k) {
  SelectPlan p = {};
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  int cap = 2 * k < 256 ? 256 : (2 * k + 31) / 32 * 32;
  p.cap = cap;
  p.select = k <= SEL_MAX_G / 2 && cap <= SEL_MAX_CAP && Ni > 0;
  if (p.select && Ni > cap) {
    int g_target = (5 * k / 2 + 31) / 32 * 32;
    g_target = g_target < 128 ? 128 : (g_target > SEL_MAX_G ? SEL_MAX_G : g_target);
    p.rows_per_group = (Ni + g_target - 1) / g_target;
    p.G = (Ni + p.rows_per_group - 1) / p.rows_per_group;
    p.threshold = true;
    if (p.G < k || p.G > SEL_MAX_G) p.select = false;         // cannot happen for k <= SEL_MAX_G / 2; be safe
  }
  size_t o = 0;
  p.off_tau = o;  o += up(sizeof(float) * (size_t)Nc);
  p.off_cnt = o;  o += up(sizeof(int) * ((size_t)Nc + 1));    // cnt[Nc] + overflow counter: one memset
  p.off_ovf = o;  o += up(sizeof(int) * (size_t)Nc);
  if (p.select) {
    p.off_cs = o;   o += up(sizeof(float) * (size_t)Nc * cap);
    p.off_ci = o;   o += up(sizeof(int) * (size_t)Nc * cap);
    p.off_gmax = o; o += up(sizeof(float) * (size_t)Nc * (p.threshold ? p.G : 0));
  }
  p.bytes = o + 256;
  return p;
}
}  // namespace alad

extern "C" int64_t alad_col_topk_select_workspace_bytes(int32_t Ni, int32_t Nc, int32_t k) {
  if (Ni < 0 || Nc < 0 || k <= 0) return 0;
  return (int64_t)alad::select_plan(Ni, Nc, k).bytes;
}

extern "C" int alad_col_topk_select(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, int32_t k, int32_t img_off,
                                    float* out_score, int32_t* out_idx, void* workspace, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Ni >= 0 && Nc >= 0 && ldS >= Nc, "alad_col_topk_select: bad shape");
  ALAD_REQUIRE(k > 0 && k <= 256, "alad_col_topk_select: k=%d out of range", k);
  if (Nc == 0) return ALAD_OK;
  ALAD_REQUIRE(out_score && out_idx && (S || Ni == 0), "alad_col_topk_select: NULL pointer");
  const SelectPlan pl = select_plan(Ni, Nc, k);
  if (!pl.select)      // large k or an empty gallery: the heap kernel over all columns (one row slice)
    return alad_col_topk(S, ldS, Ni, Nc, k, img_off, 1, out_score, out_idx, stream);
  ALAD_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "alad_col_topk_select: bad workspace");
  cudaStream_t st = as_stream(stream);
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  float* tau = reinterpret_cast<float*>(w + pl.off_tau);
  int* cnt = reinterpret_cast<int*>(w + pl.off_cnt);
  int* ovf_n = cnt + Nc;
  int* ovf_list = reinterpret_cast<int*>(w + pl.off_ovf);
  float* cs = reinterpret_cast<float*>(w + pl.off_cs);
  int* ci = reinterpret_cast<int*>(w + pl.off_ci);
  float* gmax = reinterpret_cast<float*>(w + pl.off_gmax);
  ALAD_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)Nc + 1), st));
  const unsigned col_blocks = (unsigned)((Nc + 31) / 32);
  if (pl.threshold) {
    dim3 g1(col_blocks, (unsigned)((pl.G + SEL_ROWS - 1) / SEL_ROWS)), b1(32, SEL_ROWS);
    col_groupmax_kernel<<<g1, b1, 0, st>>>(S, ldS, Ni, Nc, pl.rows_per_group, pl.G, gmax);
    const size_t tile_bytes = (size_t)pl.G * 33 * sizeof(float);
    if (pl.G <= 128)      col_threshold_kernel<4><<<col_blocks, 256, tile_bytes, st>>>(gmax, Nc, pl.G, k, tau);
    else if (pl.G <= 256) col_threshold_kernel<8><<<col_blocks, 256, tile_bytes, st>>>(gmax, Nc, pl.G, k, tau);
    else                  col_threshold_kernel<SEL_MAX_G / 32><<<col_blocks, 256, tile_bytes, st>>>(gmax, Nc, pl.G, k, tau);
  }
  dim3 g3(col_blocks, (unsigned)((Ni + SEL_CHUNK - 1) / SEL_CHUNK)), b3(32, SEL_ROWS);
  col_collect_kernel<<<g3, b3, 0, st>>>(S, ldS, Ni, Nc, img_off, pl.threshold ? tau : nullptr, pl.cap, cnt, cs, ci);
  const size_t sel_smem = (size_t)SELECT_WARPS * pl.cap * sizeof(float2);
  col_select_kernel<<<(unsigned)((Nc + SELECT_WARPS - 1) / SELECT_WARPS), 32 * SELECT_WARPS, sel_smem, st>>>(
      cnt, cs, ci, Nc, k, pl.cap, out_score, out_idx, ovf_list, ovf_n);
  // overflowed captions (normally none): per-caption heaps; CTAs beyond the queue exit immediately
  const size_t heap_smem = (size_t)k * 32 * 8;
  static thread_local size_t heap_set = 0;
  if (heap_smem > 48 * 1024 && heap_smem > heap_set) {
    ALAD_CUDA(cudaFuncSetAttribute(col_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heap_smem));
    heap_set = heap_smem;
  }
  col_topk_kernel<<<dim3(col_blocks, 1), 32, heap_smem, st>>>(S, ldS, Ni, Nc, k, img_off, 1, out_score, out_idx, ovf_list,
                                                              ovf_n);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

namespace alad {
struct FusedPlan {
  SelectPlan sel;
  size_t off_best, off_gt, bytes;
};
static FusedPlan fused_plan(int Ni, int q_rows, int q_cols, int k) {
  FusedPlan f;
  f.sel = select_plan(Ni, q_cols, k);
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  size_t o = up(f.sel.bytes);
  f.off_best = o; o += up(sizeof(unsigned long long) * (size_t)(q_rows > 0 ? q_rows : 1));
  f.off_gt = o;   o += up(sizeof(float) * (size_t)(q_cols > 0 ? q_cols : 1));
  f.bytes = o + 256;
  return f;
}
}  // namespace alad

extern "C" int64_t alad_rank_fused_workspace_bytes(int32_t Ni, int32_t q_rows, int32_t q_cols, int32_t k) {
  if (Ni < 0 || q_rows < 0 || q_cols < 0 || k <= 0) return 0;
  return (int64_t)alad::fused_plan(Ni, q_rows, q_cols, k).bytes;
}

extern "C" int alad_rank_fused(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, int32_t group, int32_t img_off,
                               int32_t q_rows, int32_t q_cols, int32_t k, const float* gt, int32_t* rank, int32_t* top1,
                               int32_t* count, float* out_score, int32_t* out_idx, void* workspace, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Ni >= 0 && Nc >= 0 && ldS >= Nc && group > 0, "alad_rank_fused: bad shape");
  ALAD_REQUIRE(q_rows >= 0 && q_rows <= Ni && q_cols >= 0 && q_cols <= Nc, "alad_rank_fused: q_rows=%d q_cols=%d out of range",
               q_rows, q_cols);
  ALAD_REQUIRE(k > 0 && k <= 256, "alad_rank_fused: k=%d out of range", k);
  ALAD_REQUIRE((rank && top1) || q_rows == 0, "alad_rank_fused: NULL pointer");
  ALAD_REQUIRE((out_score && out_idx) || q_cols == 0, "alad_rank_fused: NULL pointer");
  ALAD_REQUIRE(S || Ni == 0 || Nc == 0, "alad_rank_fused: NULL pointer");
  ALAD_REQUIRE(workspace && (reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "alad_rank_fused: bad workspace");
  cudaStream_t st = as_stream(stream);
  const FusedPlan fp = fused_plan(Ni, q_rows, q_cols, k);
  const SelectPlan& pl = fp.sel;
  uint8_t* w = reinterpret_cast<uint8_t*>(workspace);
  // count != NULL without gt: this block holds every query caption's image, the ground-truth scores are its own entries
  const float* gt_use = gt;
  if (count && !gt && q_cols > 0) {
    float* own = reinterpret_cast<float*>(w + fp.off_gt);
    ALAD_CUDA(cudaMemsetAsync(own, 0, sizeof(float) * (size_t)q_cols, st));
    if (Ni > 0) col_gt_kernel<<<(q_cols + 255) / 256, 256, 0, st>>>(S, ldS, Ni, q_cols, group, img_off, own);
    gt_use = own;
  }
  const bool fused = pl.select && pl.threshold && Nc > 0 && q_cols > 0 && (ldS & 3) == 0 &&
                     (reinterpret_cast<uintptr_t>(S) & 15) == 0;
  if (!fused) {          // small or unaligned blocks: the one-purpose kernels
    int rc = q_rows > 0 ? alad_rank_rows(S, ldS, q_rows, Nc, group, img_off, rank, top1, stream) : ALAD_OK;
    if (rc == ALAD_OK && count && q_cols > 0) rc = alad_col_count(S, ldS, Ni, q_cols, group, img_off, gt_use, count, stream);
    if (rc == ALAD_OK && q_cols > 0) rc = alad_col_topk_select(S, ldS, Ni, q_cols, k, img_off, out_score, out_idx, workspace, stream);
    return rc;
  }
  float* tau = reinterpret_cast<float*>(w + pl.off_tau);
  int* cnt = reinterpret_cast<int*>(w + pl.off_cnt);
  int* ovf_n = cnt + q_cols;
  int* ovf_list = reinterpret_cast<int*>(w + pl.off_ovf);
  float* cs = reinterpret_cast<float*>(w + pl.off_cs);
  int* ci = reinterpret_cast<int*>(w + pl.off_ci);
  float* gmax = reinterpret_cast<float*>(w + pl.off_gmax);
  unsigned long long* best = reinterpret_cast<unsigned long long*>(w + fp.off_best);
  ALAD_CUDA(cudaMemsetAsync(cnt, 0, sizeof(int) * ((size_t)q_cols + 1), st));
  if (q_rows > 0) {
    ALAD_CUDA(cudaMemsetAsync(rank, 0, sizeof(int32_t) * (size_t)q_rows, st));
    ALAD_CUDA(cudaMemsetAsync(best, 0, sizeof(unsigned long long) * (size_t)q_rows, st));
  }
  if (count) ALAD_CUDA(cudaMemsetAsync(count, 0, sizeof(int32_t) * (size_t)q_cols, st));
  // rows beyond the query columns still count for the i2t direction: the sweep covers all Nc columns
  dim3 gs((unsigned)((Nc + FS_COLS - 1) / FS_COLS), (unsigned)pl.G);
  if (count)
    rank_sweep_kernel<true><<<gs, FS_THREADS, 0, st>>>(S, ldS, Ni, Nc, group, img_off, q_rows, q_cols, pl.rows_per_group, gt_use,
                                                       rank, best, count, gmax);
  else
    rank_sweep_kernel<false><<<gs, FS_THREADS, 0, st>>>(S, ldS, Ni, Nc, group, img_off, q_rows, q_cols, pl.rows_per_group, nullptr,
                                                        rank, best, nullptr, gmax);
  if (q_rows > 0) rank_sweep_finish_kernel<<<(q_rows + 255) / 256, 256, 0, st>>>(best, q_rows, top1);
  const unsigned col_blocks = (unsigned)((q_cols + 31) / 32);
  const size_t tile_bytes = (size_t)pl.G * 33 * sizeof(float);
  if (pl.G <= 128)      col_threshold_kernel<4><<<col_blocks, 256, tile_bytes, st>>>(gmax, q_cols, pl.G, k, tau);
  else if (pl.G <= 256) col_threshold_kernel<8><<<col_blocks, 256, tile_bytes, st>>>(gmax, q_cols, pl.G, k, tau);
  else                  col_threshold_kernel<SEL_MAX_G / 32><<<col_blocks, 256, tile_bytes, st>>>(gmax, q_cols, pl.G, k, tau);
  dim3 g3(col_blocks, (unsigned)((Ni + SEL_CHUNK - 1) / SEL_CHUNK)), b3(32, SEL_ROWS);
  col_collect_kernel<<<g3, b3, 0, st>>>(S, ldS, Ni, q_cols, img_off, tau, pl.cap, cnt, cs, ci);
  const size_t sel_smem = (size_t)SELECT_WARPS * pl.cap * sizeof(float2);
  col_select_kernel<<<(unsigned)((q_cols + SELECT_WARPS - 1) / SELECT_WARPS), 32 * SELECT_WARPS, sel_smem, st>>>(
      cnt, cs, ci, q_cols, k, pl.cap, out_score, out_idx, ovf_list, ovf_n);
  const size_t heap_smem = (size_t)k * 32 * 8;
  static thread_local size_t heap_set = 0;
  if (heap_smem > 48 * 1024 && heap_smem > heap_set) {
    ALAD_CUDA(cudaFuncSetAttribute(col_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heap_smem));
    heap_set = heap_smem;
  }
  col_topk_kernel<<<dim3(col_blocks, 1), 32, heap_smem, st>>>(S, ldS, Ni, q_cols, k, img_off, 1, out_score, out_idx, ovf_list,
                                                              ovf_n);
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
