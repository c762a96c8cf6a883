// B x B loss kernels, forward + gradient in one go (HBM / launch-latency bound):
//   alad_triplet_fwd_bwd  -- VSE++ hinge with hardest negatives, alad/loss.py:42-67
//   alad_listnet_fwd_bwd  -- ListNet distillation, alad/loss.py:427-445 (teacher detached :370)
// Row direction: one CTA per row, coalesced 128-bit-friendly sweeps, warp-shuffle reductions.
// Column direction: 32-column strips x 8 row slices per CTA (coalesced 128 B row segments), the rows
// split into chunks of chunk_rows_of(B) along the grid so that large B fills the machine; per-chunk partial
// results are combined in a fixed order.  The loss scalar is reduced in a fixed order by the last
// CTA to finish (deterministic).
#include <math.h>

#include "common.h"

namespace alad {

constexpr int LT = 256;          // threads per CTA
constexpr int CS = 8;            // row slices of a column-strip CTA
constexpr int CV = 4;            // 32-column groups per column-strip CTA (4 x 128 B contiguous per visited row)
constexpr int SW = 32 * CV;      // columns per strip
__host__ __device__ inline int n_strips_of(int B) { return (B + SW - 1) / SW; }
// rows per column-strip CTA: B/16 rounded up to a multiple of CS, within [32, 256] -- enough CTAs to fill the
// machine at B = 512 without long partial lists at B = 8192
__host__ __device__ inline int chunk_rows_of(int B) {
  const int r = ((B + 15) / 16 + CS - 1) / CS * CS;
  return r < 32 ? 32 : (r > 256 ? 256 : r);
}
__host__ __device__ inline int n_chunks_of(int B) { return (B + chunk_rows_of(B) - 1) / chunk_rows_of(B); }

__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max_f(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ int warp_sum_i(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
// (value, index) arg-max, ties -> lower index (torch CPU first-occurrence semantics)
__device__ __forceinline__ void argmax_combine(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) {
    v = ov;
    i = oi;
  }
}
__device__ __forceinline__ void warp_argmax(float& v, int& i) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, v, o);
    const int oi = __shfl_xor_sync(0xffffffffu, i, o);
    argmax_combine(v, i, ov, oi);
  }
}
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* scratch) {   // LT threads, result valid in thread 0
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  T tot = 0;
  if (threadIdx.x == 0)
    for (int w = 0; w < LT / 32; ++w) tot += scratch[w];
  __syncthreads();
  return tot;
}
// fixed-order sum of n floats by one CTA (deterministic)
__device__ float cta_ordered_sum(const float* x, int n, float* scratch) {
  float s = 0.f;
  for (int i = threadIdx.x; i < n; i += LT) s += x[i];
  return block_sum<float>(s, scratch);
}
__device__ __forceinline__ bool last_cta_done(unsigned int* counter) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// true in every thread of the LAST of `expected` CTAs to arrive on `counter` (which is re-armed to 0)
__device__ __forceinline__ bool last_of_group(unsigned int* counter, unsigned int expected) {
  __shared__ bool group_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    group_last = (atomicAdd(counter, 1u) == expected - 1);
    if (group_last) *counter = 0;
  }
  __syncthreads();
  if (group_last) __threadfence();
  return group_last;
}

__global__ void diag_kernel(const float* __restrict__ S, long long ld, int B, float* __restrict__ diag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) diag[i] = S[(long long)i * ld + i];
}

// ------------------------------------------------------------------------------------ triplet
// workspace layout (floats/ints, B each): diag | rowval | colval | rowarg | colarg | counter
struct TripletParams {
  const float* S;
  long long ld;
  int B;
  float margin;
  int max_violation;
  float* loss;
  float* G;            // optional dense dL/dS, [B, ldG]
  long long ldG;
  float* diag;
  float* rowval;
  float* colval;
  int* rowarg;         // max_violation: hardest negative (-1 = none); sum mode: #violations
  int* colarg;
  float* cpart_val;    // [n_chunks][B] per-chunk column partials (best value / hinge sum)
  int* cpart_arg;      // [n_chunks][B]                           (arg-max row / violation count)
  unsigned int* strip_cnt;   // [n_strips] arrivals of the chunk CTAs of a 32-column strip
  unsigned int* counter;
};

__global__ void __launch_bounds__(LT) triplet_kernel(const TripletParams p) {
  __shared__ float sf[LT / 32];
  __shared__ int si[LT / 32];
  __shared__ float cv[CS][SW];
  __shared__ int ci[CS][SW];
  const int B = p.B;
  if ((int)blockIdx.x < B) {
    // ---------------- row direction: cost_s[i, j] = [margin + S_ij - S_ii]_+ (caption retrieval)
    const int i = blockIdx.x;
    const float* row = p.S + (long long)i * p.ld;
    const float dii = p.diag[i];
    float best = 0.f, sum = 0.f;
    int barg = B, cnt = 0;
    for (int j = threadIdx.x; j < B; j += LT) {
      const float s = __ldg(row + j);
      const float c_s = (j == i) ? 0.f : fmaxf(p.margin + s - dii, 0.f);
      if (p.max_violation) {
        argmax_combine(best, barg, c_s, j);
        if (p.G) p.G[(long long)i * p.ldG + j] = 0.f;
      } else {
        const float c_i = (j == i) ? 0.f : fmaxf(p.margin + s - __ldg(p.diag + j), 0.f);
        sum += c_s;
        cnt += c_s > 0.f;
        if (p.G) p.G[(long long)i * p.ldG + j] = (c_s > 0.f ? 1.f : 0.f) + (c_i > 0.f ? 1.f : 0.f);
      }
    }
    if (p.max_violation) {
      warp_argmax(best, barg);
      const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
      if (lane == 0) {
        sf[warp] = best;
        si[warp] = barg;
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        for (int w = 1; w < LT / 32; ++w) argmax_combine(best, barg, sf[w], si[w]);
        p.rowval[i] = best;
        p.rowarg[i] = best > 0.f ? barg : -1;
      }
    } else {
      const float tot = block_sum<float>(sum, sf);
      const int c = block_sum<int>(cnt, si);
      if (threadIdx.x == 0) {
        p.rowval[i] = tot;
        p.rowarg[i] = c;
      }
    }
  } else {
    // ---------------- column direction: cost_im[i, j] = [margin + S_ij - S_jj]_+ (image retrieval)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n_strips = n_strips_of(B);
    const int cta = blockIdx.x - B;
    const int strip = cta % n_strips, chunk = cta / n_strips;
    const int j0 = strip * SW + tx;                   // this thread's columns: j0 + 32 * u
    const int chunk_rows = chunk_rows_of(B);
    const int r_end = min(B, (chunk + 1) * chunk_rows);
    float best[CV], sum[CV], djj[CV];
    int barg[CV], cnt[CV];
#pragma unroll
    for (int u = 0; u < CV; ++u) {
      best[u] = 0.f; sum[u] = 0.f; barg[u] = B; cnt[u] = 0;
      djj[u] = (j0 + 32 * u < B) ? p.diag[j0 + 32 * u] : 0.f;
    }
    for (int i = chunk * chunk_rows + ty; i < r_end; i += CS) {
      const float* row = p.S + (long long)i * p.ld;
      float v[CV];
#pragma unroll
      for (int u = 0; u < CV; ++u) v[u] = (j0 + 32 * u < B) ? __ldg(row + j0 + 32 * u) : 0.f;
#pragma unroll
      for (int u = 0; u < CV; ++u) {
        const int j = j0 + 32 * u;
        const float c_i = (i == j || j >= B) ? 0.f : fmaxf(p.margin + v[u] - djj[u], 0.f);
        if (p.max_violation) {
          argmax_combine(best[u], barg[u], c_i, i);
        } else {
          sum[u] += c_i;
          cnt[u] += c_i > 0.f;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < CV; ++u) {
      cv[ty][tx + 32 * u] = p.max_violation ? best[u] : sum[u];
      ci[ty][tx + 32 * u] = p.max_violation ? barg[u] : cnt[u];
    }
    __syncthreads();
    if (ty < CV) {                                    // warp ty finishes column group ty
      const int c = tx + 32 * ty, j = strip * SW + c;
      if (j < B) {
        float v = cv[0][c];
        int a = ci[0][c];
        for (int y = 1; y < CS; ++y) {
          if (p.max_violation) {
            argmax_combine(v, a, cv[y][c], ci[y][c]);
          } else {
            v += cv[y][c];
            a += ci[y][c];
          }
        }
        p.cpart_val[(long long)chunk * B + j] = v;
        p.cpart_arg[(long long)chunk * B + j] = a;
      }
    }
    // the last chunk CTA of this strip combines the partials in chunk order (ties: lower row index wins)
    const int nch = n_chunks_of(B);
    if (last_of_group(p.strip_cnt + strip, nch) && ty < CV) {
      const int j = strip * SW + tx + 32 * ty;
      if (j < B) {
        float v = __ldcg(p.cpart_val + j);
        int a = __ldcg(p.cpart_arg + j);
        for (int ch = 1; ch < nch; ++ch) {
          const float ov = __ldcg(p.cpart_val + (long long)ch * B + j);
          const int oa = __ldcg(p.cpart_arg + (long long)ch * B + j);
          if (p.max_violation) {
            argmax_combine(v, a, ov, oa);
          } else {
            v += ov;
            a += oa;
          }
        }
        p.colval[j] = v;
        p.colarg[j] = p.max_violation ? (v > 0.f ? a : -1) : a;
      }
    }
  }
  // ---------------- last CTA: loss in a fixed order, sparse / diagonal part of the gradient
  if (!last_cta_done(p.counter)) return;
  const float lr = cta_ordered_sum(p.rowval, B, sf);
  const float lc = cta_ordered_sum(p.colval, B, sf);
  if (threadIdx.x == 0) {
    *p.loss = lr + lc;
    *p.counter = 0;                                   // re-arm for the next call
  }
  if (p.G) {
    for (int i = threadIdx.x; i < B; i += LT) {
      if (p.max_violation) {
        // small integers: float atomics are exact, order does not matter
        const int jr = p.rowarg[i];
        if (jr >= 0) {
          atomicAdd(p.G + (long long)i * p.ldG + jr, 1.f);
          atomicAdd(p.G + (long long)i * p.ldG + i, -1.f);
        }
        const int ic = p.colarg[i];                   // column i
        if (ic >= 0) {
          atomicAdd(p.G + (long long)ic * p.ldG + i, 1.f);
          atomicAdd(p.G + (long long)i * p.ldG + i, -1.f);
        }
      } else {
        p.G[(long long)i * p.ldG + i] = -static_cast<float>(p.rowarg[i] + p.colarg[i]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------ listnet
// per-row / per-column softmax statistics of tau*M (student) and T (teacher)
struct SoftStats {
  float m_s, z_s, m_t, z_t;   // max and sum(exp(x - max))
};
__device__ __forceinline__ void online_update(float& m, float& z, float x) {
  if (x > m) {
    z = z * expf(m - x) + 1.f;
    m = x;
  } else {
    z += expf(x - m);
  }
}
__device__ __forceinline__ void online_merge(float& m, float& z, float om, float oz) {
  const float nm = fmaxf(m, om);
  if (nm == -INFINITY) return;
  z = z * expf(m - nm) + oz * expf(om - nm);
  m = nm;
}

struct ListnetParams {
  const float* T;
  long long ldT;
  const float* M;
  long long ldM;
  int B;
  float tau, eps;
  float* loss;
  float* dM;           // optional [B, ldG]
  long long ldG;
  float* rstat;        // [B][5]: m_s, z_s, m_t, z_t, A   (row direction, dim=1)
  float* cstat;        // [B][5]                          (column direction, dim=0)
  float* rcost;        // [B]
  float* ccost;        // [B]
  float* cpart;        // [n_chunks][B][4] per-chunk column partials: softmax stats, then (cost, A)
  unsigned int* strip_cnt;   // [n_strips]
  unsigned int* counter;
};

// cost and A = sum_j t*p/(p+eps) contributions of one element given final statistics
__device__ __forceinline__ void listnet_elem(float xm, float xt, const SoftStats& st, float tau, float eps,
                                             float& cost, float& a, float& pout) {
  const float pr = expf(tau * xm - st.m_s) / st.z_s;
  const float t = expf(xt - st.m_t) / st.z_t;
  cost = -t * logf(pr + eps);
  a = t * (pr / (pr + eps));
  pout = pr;
}

__global__ void __launch_bounds__(LT) listnet_stats_kernel(const ListnetParams p) {
  __shared__ float sf[LT / 32];
  __shared__ float sm[4][LT / 32];
  __shared__ float cs[CS][SW][4];
  const int B = p.B;
  if ((int)blockIdx.x < B) {
    // ---------------- row i: softmax over j (sentence retrieval, dim=1)
    const int i = blockIdx.x;
    const float* rm = p.M + (long long)i * p.ldM;
    const float* rt = p.T + (long long)i * p.ldT;
    float ms = -INFINITY, zs = 0.f, mt = -INFINITY, zt = 0.f;
    for (int j = threadIdx.x; j < B; j += LT) {
      online_update(ms, zs, p.tau * __ldg(rm + j));
      online_update(mt, zt, __ldg(rt + j));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      online_merge(ms, zs, __shfl_xor_sync(0xffffffffu, ms, o), __shfl_xor_sync(0xffffffffu, zs, o));
      online_merge(mt, zt, __shfl_xor_sync(0xffffffffu, mt, o), __shfl_xor_sync(0xffffffffu, zt, o));
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
      sm[0][warp] = ms; sm[1][warp] = zs; sm[2][warp] = mt; sm[3][warp] = zt;
    }
    __syncthreads();
    SoftStats st;
    {
      float a = sm[0][0], b = sm[1][0], c = sm[2][0], d = sm[3][0];
      for (int w = 1; w < LT / 32; ++w) {
        online_merge(a, b, sm[0][w], sm[1][w]);
        online_merge(c, d, sm[2][w], sm[3][w]);
      }
      st.m_s = a; st.z_s = b; st.m_t = c; st.z_t = d;
    }
    __syncthreads();
    float cost = 0.f, A = 0.f;
    for (int j = threadIdx.x; j < B; j += LT) {
      float c, a, pr;
      listnet_elem(__ldg(rm + j), __ldg(rt + j), st, p.tau, p.eps, c, a, pr);
      cost += c;
      A += a;
    }
    const float ctot = block_sum<float>(cost, sf);
    const float atot = block_sum<float>(A, sf);
    if (threadIdx.x == 0) {
      float* o = p.rstat + 5 * (long long)i;
      o[0] = st.m_s; o[1] = st.z_s; o[2] = st.m_t; o[3] = st.z_t; o[4] = atot;
      p.rcost[i] = ctot;
    }
  } else {
    // ---------------- SW columns x one row chunk: partial softmax statistics over i (image retrieval, dim=0)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int n_strips = n_strips_of(B);
    const int cta = blockIdx.x - B;
    const int strip = cta % n_strips, chunk = cta / n_strips;
    const int j0 = strip * SW + tx;
    const int chunk_rows = chunk_rows_of(B);
    const int r_end = min(B, (chunk + 1) * chunk_rows);
    float ms[CV], zs[CV], mt[CV], zt[CV];
#pragma unroll
    for (int u = 0; u < CV; ++u) {
      ms[u] = -INFINITY; zs[u] = 0.f; mt[u] = -INFINITY; zt[u] = 0.f;
    }
    for (int i = chunk * chunk_rows + ty; i < r_end; i += CS) {
      const float* rm = p.M + (long long)i * p.ldM;
      const float* rt = p.T + (long long)i * p.ldT;
      float vm[CV], vt[CV];
#pragma unroll
      for (int u = 0; u < CV; ++u) {
        const bool ok = j0 + 32 * u < B;
        vm[u] = ok ? __ldg(rm + j0 + 32 * u) : 0.f;
        vt[u] = ok ? __ldg(rt + j0 + 32 * u) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < CV; ++u) {
        online_update(ms[u], zs[u], p.tau * vm[u]);
        online_update(mt[u], zt[u], vt[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < CV; ++u) {
      float* c = cs[ty][tx + 32 * u];
      c[0] = ms[u]; c[1] = zs[u]; c[2] = mt[u]; c[3] = zt[u];
    }
    __syncthreads();
    if (ty < CV) {
      const int c = tx + 32 * ty, j = strip * SW + c;
      if (j < B) {
        float a0 = cs[0][c][0], a1 = cs[0][c][1], a2 = cs[0][c][2], a3 = cs[0][c][3];
        for (int y = 1; y < CS; ++y) {
          online_merge(a0, a1, cs[y][c][0], cs[y][c][1]);
          online_merge(a2, a3, cs[y][c][2], cs[y][c][3]);
        }
        float* o = p.cpart + 4 * ((long long)chunk * B + j);
        o[0] = a0; o[1] = a1; o[2] = a2; o[3] = a3;
      }
    }
    // the last chunk CTA of this strip merges the partial statistics in chunk order
    const int nch = n_chunks_of(B);
    if (last_of_group(p.strip_cnt + strip, nch) && ty < CV) {
      const int j = strip * SW + tx + 32 * ty;
      if (j < B) {
        float a = -INFINITY, b = 0.f, c = -INFINITY, d = 0.f;
        for (int ch = 0; ch < nch; ++ch) {
          const float* o = p.cpart + 4 * ((long long)ch * B + j);
          online_merge(a, b, __ldcg(o), __ldcg(o + 1));
          online_merge(c, d, __ldcg(o + 2), __ldcg(o + 3));
        }
        float* o = p.cstat + 5 * (long long)j;
        o[0] = a; o[1] = b; o[2] = c; o[3] = d;
      }
    }
  }
}

// second column pass: final statistics of every column (merge of the chunk partials, fixed order), then the
// chunk's share of cost and A = sum t*p/(p+eps); the last CTA adds the shares up and writes the loss
__global__ void __launch_bounds__(LT) listnet_colcost_kernel(const ListnetParams p) {
  __shared__ float sf[LT / 32];
  __shared__ float cs[CS][SW][2];
  __shared__ SoftStats fin[SW];
  const int B = p.B;
  const int nch = n_chunks_of(B);
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n_strips = n_strips_of(B);
  const int strip = blockIdx.x % n_strips, chunk = blockIdx.x / n_strips;
  const int j0 = strip * SW + tx;
  const int chunk_rows = chunk_rows_of(B);
  const int r_end = min(B, (chunk + 1) * chunk_rows);
  if (ty < CV) {
    const int c = tx + 32 * ty, j = strip * SW + c;
    const float* o = p.cstat + 5 * (long long)(j < B ? j : 0);
    fin[c].m_s = o[0]; fin[c].z_s = o[1]; fin[c].m_t = o[2]; fin[c].z_t = o[3];
  }
  __syncthreads();
  SoftStats st[CV];
  float cost[CV], A[CV];
#pragma unroll
  for (int u = 0; u < CV; ++u) {
    st[u] = fin[tx + 32 * u];
    cost[u] = 0.f;
    A[u] = 0.f;
  }
  for (int i = chunk * chunk_rows + ty; i < r_end; i += CS) {
    const float* rm = p.M + (long long)i * p.ldM;
    const float* rt = p.T + (long long)i * p.ldT;
    float vm[CV], vt[CV];
#pragma unroll
    for (int u = 0; u < CV; ++u) {
      const bool ok = j0 + 32 * u < B;
      vm[u] = ok ? __ldg(rm + j0 + 32 * u) : 0.f;
      vt[u] = ok ? __ldg(rt + j0 + 32 * u) : 0.f;
    }
#pragma unroll
    for (int u = 0; u < CV; ++u) {
      if (j0 + 32 * u < B) {
        float c, a, pr;
        listnet_elem(vm[u], vt[u], st[u], p.tau, p.eps, c, a, pr);
        cost[u] += c;
        A[u] += a;
      }
    }
  }
#pragma unroll
  for (int u = 0; u < CV; ++u) {
    cs[ty][tx + 32 * u][0] = cost[u];
    cs[ty][tx + 32 * u][1] = A[u];
  }
  __syncthreads();
  float* share = p.cpart;      // the statistic partials are dead by now: reuse the buffer as [n_chunks][B][2]
  if (ty < CV) {
    const int c = tx + 32 * ty, j = strip * SW + c;
    if (j < B) {
      float ct = cs[0][c][0], at = cs[0][c][1];
      for (int y = 1; y < CS; ++y) {
        ct += cs[y][c][0];
        at += cs[y][c][1];
      }
      share[2 * ((long long)chunk * B + j)] = ct;
      share[2 * ((long long)chunk * B + j) + 1] = at;
    }
  }
  if (last_of_group(p.strip_cnt + strip, nch) && ty < CV) {
    const int j = strip * SW + tx + 32 * ty;
    if (j < B) {
      float ct = 0.f, at = 0.f;
      for (int ch = 0; ch < nch; ++ch) {
        ct += __ldcg(share + 2 * ((long long)ch * B + j));
        at += __ldcg(share + 2 * ((long long)ch * B + j) + 1);
      }
      p.cstat[5 * (long long)j + 4] = at;
      p.ccost[j] = ct;
    }
  }
  if (!last_cta_done(p.counter)) return;
  const float lr = cta_ordered_sum(p.rcost, B, sf);
  const float lc = cta_ordered_sum(p.ccost, B, sf);
  if (threadIdx.x == 0) {
    *p.loss = lc / B + lr / B;        // im_cost + s_cost (loss.py:445)
    *p.counter = 0;
  }
}

// dL/dM[i,j] = (tau/B) * [ (p_r*A_r(i) - a_r) + (p_c*A_c(j) - a_c) ]   
__global__ void __launch_bounds__(LT) listnet_grad_kernel(const ListnetParams p) {
  const int B = p.B;
  const int i = blockIdx.y;
  const int j = blockIdx.x * LT + threadIdx.x;
  if (j >= B) return;
  const float* r = p.rstat + 5 * (long long)i;
  const float* c = p.cstat + 5 * (long long)j;
  SoftStats sr{r[0], r[1], r[2], r[3]}, sc{c[0], c[1], c[2], c[3]};
  const float xm = __ldg(p.M + (long long)i * p.ldM + j), xt = __ldg(p.T + (long long)i * p.ldT + j);
  float cost, a_r, p_r, a_c, p_c;
  listnet_elem(xm, xt, sr, p.tau, p.eps, cost, a_r, p_r);
  listnet_elem(xm, xt, sc, p.tau, p.eps, cost, a_c, p_c);
  p.dM[(long long)i * p.ldG + j] = (p.tau / B) * ((p_r * r[4] - a_r) + (p_c * c[4] - a_c));
}

}  // namespace alad

extern "C" int64_t alad_loss_workspace_bytes(int32_t B) {
  // shared by both losses: 12 arrays of B 4-byte words + counter + padding, plus the per-chunk column
  // partials (listnet: 4 + 2 floats, triplet: 2 words per chunk and column)
  const int64_t b = B > 0 ? B : 1;
  return 12 * 4 * b + 256 + 6 * 4 * b * alad::n_chunks_of((int)b) + 4 * ((b + 31) / 32) + 64;
}

extern "C" int alad_triplet_fwd_bwd(const float* S, int64_t ldS, int32_t B, float margin, int32_t max_violation,
                                    float* loss, float* G, int64_t ldG, int32_t* row_arg, int32_t* col_arg,
                                    void* workspace, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(B >= 0 && ldS >= B && (G == nullptr || ldG >= B), "alad_triplet_fwd_bwd: bad shape");
  ALAD_REQUIRE(loss && workspace, "alad_triplet_fwd_bwd: NULL pointer");
  cudaStream_t st = as_stream(stream);
  if (B == 0) {
    ALAD_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    return ALAD_OK;
  }
  ALAD_REQUIRE(S && row_arg && col_arg, "alad_triplet_fwd_bwd: NULL pointer");
  float* w = reinterpret_cast<float*>(workspace);
  TripletParams p;
  p.S = S; p.ld = ldS; p.B = B; p.margin = margin; p.max_violation = max_violation; p.loss = loss; p.G = G; p.ldG = ldG;
  p.diag = w; p.rowval = w + B; p.colval = w + 2 * (size_t)B;
  p.rowarg = row_arg; p.colarg = col_arg;
  p.counter = reinterpret_cast<unsigned int*>(w + 3 * (size_t)B);
  const int nch = n_chunks_of(B);
  p.cpart_val = w + 12 * (size_t)B + 64;
  p.cpart_arg = reinterpret_cast<int*>(p.cpart_val + (size_t)nch * B);
  p.strip_cnt = reinterpret_cast<unsigned int*>(p.cpart_val + 6 * (size_t)nch * B);
  ALAD_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), st));
  ALAD_CUDA(cudaMemsetAsync(p.strip_cnt, 0, sizeof(unsigned int) * (size_t)((B + 31) / 32), st));
  diag_kernel<<<(B + 255) / 256, 256, 0, st>>>(S, ldS, B, p.diag);
  triplet_kernel<<<B + n_strips_of(B) * nch, LT, 0, st>>>(p);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_listnet_fwd_bwd(const float* teacher, int64_t ldT, const float* student, int64_t ldM, int32_t B,
                                    float temperature, float eps, float* loss, float* dM, int64_t ldG,
                                    void* workspace, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(B >= 0 && ldT >= B && ldM >= B && (dM == nullptr || ldG >= B), "alad_listnet_fwd_bwd: bad shape");
  ALAD_REQUIRE(loss && workspace, "alad_listnet_fwd_bwd: NULL pointer");
  cudaStream_t st = as_stream(stream);
  if (B == 0) {
    ALAD_CUDA(cudaMemsetAsync(loss, 0, sizeof(float), st));
    return ALAD_OK;
  }
  ALAD_REQUIRE(teacher && student, "alad_listnet_fwd_bwd: NULL pointer");
  float* w = reinterpret_cast<float*>(workspace);
  ListnetParams p;
  p.T = teacher; p.ldT = ldT; p.M = student; p.ldM = ldM; p.B = B; p.tau = temperature; p.eps = eps;
  p.loss = loss; p.dM = dM; p.ldG = ldG;
  p.rstat = w; p.cstat = w + 5 * (size_t)B; p.rcost = w + 10 * (size_t)B; p.ccost = w + 11 * (size_t)B;
  p.counter = reinterpret_cast<unsigned int*>(w + 12 * (size_t)B);
  const int nch = n_chunks_of(B);
  p.cpart = w + 12 * (size_t)B + 64;
  p.strip_cnt = reinterpret_cast<unsigned int*>(p.cpart + 6 * (size_t)nch * B);
  ALAD_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), st));
  ALAD_CUDA(cudaMemsetAsync(p.strip_cnt, 0, sizeof(unsigned int) * (size_t)((B + 31) / 32), st));
  listnet_stats_kernel<<<B + n_strips_of(B) * nch, LT, 0, st>>>(p);
  listnet_colcost_kernel<<<n_strips_of(B) * nch, LT, 0, st>>>(p);
  if (dM) {
    dim3 grid((B + LT - 1) / LT, B);
    listnet_grad_kernel<<<grid, LT, 0, st>>>(p);
  }
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}
