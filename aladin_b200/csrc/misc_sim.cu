// Small companions of the scoring path that are not GEMM-shaped:
//   alad_order_scores / alad_order_scores_bwd -- order-embedding similarity, alad/loss.py:20-26
//   alad_normalize_bwd   -- Jacobian of x / max(||x||, eps) applied in place (gradient of
//                           cosine_sim, alad/loss.py:13-18, after the two GEMMs of dot_sim)
//   alad_pool_tokens_bwd -- gradient of the pooled-token sums behind the 'sum' / 'mean'
//                           aggregations (alad/loss.py:120-123)
// CUDA-core kernels: coalesced feature-dim sweeps, warp-shuffle reductions, no atomics.
#include <math.h>

#include "common.h"

namespace alad {

// ------------------------------------------------------------------------------------ order_sim
// score[i, j] = -|| max(0, s_j - im_i) ||_2.  64 x 64 output tile per CTA (256 threads, 4 x 4
// outputs each), feature chunks of 16 staged in shared memory.
constexpr int OT = 64, OK = 16;

__global__ void __launch_bounds__(256) order_fwd_kernel(const float* __restrict__ im, long long ld_im,
                                                        const float* __restrict__ s, long long ld_s, int Ni, int Nc,
                                                        int d, float* __restrict__ out, long long ld_out) {
  __shared__ float a[OK][OT + 1];   // im tile, feature-major
  __shared__ float b[OK][OT + 1];   // s tile
  const int i0 = blockIdx.y * OT, j0 = blockIdx.x * OT;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4] = {};
  for (int k0 = 0; k0 < d; k0 += OK) {
    for (int e = threadIdx.x; e < OT * OK; e += 256) {
      const int r = e / OK, k = e % OK;
      a[k][r] = (i0 + r < Ni && k0 + k < d) ? __ldg(im + (long long)(i0 + r) * ld_im + k0 + k) : 0.f;
      b[k][r] = (j0 + r < Nc && k0 + k < d) ? __ldg(s + (long long)(j0 + r) * ld_s + k0 + k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < OK; ++k) {
      float av[4], bv[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        av[u] = a[k][ty * 4 + u];
        bv[u] = b[k][tx * 4 + u];
      }
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int v = 0; v < 4; ++v) {
          const float y = fmaxf(bv[v] - av[u], 0.f);
          acc[u][v] = fmaf(y, y, acc[u][v]);
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int u = 0; u < 4; ++u)
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      const int i = i0 + ty * 4 + u, j = j0 + tx * 4 + v;
      if (i < Ni && j < Nc) out[(long long)i * ld_out + j] = -sqrtf(acc[u][v]);
    }
}

// d_s[j, k] = sum_i w_ij * max(0, s_jk - im_ik),  d_im[i, k] = -sum_j w_ij * max(0, s_jk - im_ik),
// w_ij = G_ij / score_ij (0 where the score is 0).  One CTA per output row, threads over k.
__global__ void __launch_bounds__(256) order_bwd_kernel(const float* __restrict__ im, long long ld_im,
                                                        const float* __restrict__ s, long long ld_s, int Ni, int Nc,
                                                        int d, const float* __restrict__ score, long long ld_sc,
                                                        const float* __restrict__ G, long long ld_g,
                                                        float* __restrict__ d_im, float* __restrict__ d_s) {
  extern __shared__ float wrow[];
  const bool for_s = (int)blockIdx.x < Nc;
  const int row = for_s ? blockIdx.x : blockIdx.x - Nc;
  const int n_other = for_s ? Ni : Nc;
  for (int o = threadIdx.x; o < n_other; o += 256) {
    const long long i = for_s ? o : row, j = for_s ? row : o;
    const float sc = score[i * ld_sc + j];
    wrow[o] = sc != 0.f ? G[i * ld_g + j] / sc : 0.f;
  }
  __syncthreads();
  for (int k = threadIdx.x; k < d; k += 256) {
    float acc = 0.f;
    if (for_s) {
      const float sv = __ldg(s + (long long)row * ld_s + k);
      for (int i = 0; i < Ni; ++i) acc = fmaf(wrow[i], fmaxf(sv - __ldg(im + (long long)i * ld_im + k), 0.f), acc);
      d_s[(long long)row * d + k] = acc;
    } else {
      const float iv = __ldg(im + (long long)row * ld_im + k);
      for (int j = 0; j < Nc; ++j) acc = fmaf(wrow[j], fmaxf(__ldg(s + (long long)j * ld_s + k) - iv, 0.f), acc);
      d_im[(long long)row * d + k] = -acc;
    }
  }
}

// ------------------------------------------------------------------------------------ normalize bwd
__device__ __forceinline__ float m_wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per row: dx <- J(x) dx with J the Jacobian of x / max(||x||, eps)
__global__ void normalize_bwd_kernel(const float* __restrict__ x, long long ld_x, long long rows, int d, float eps,
                                     float* __restrict__ dx, long long ld_dx) {
  const long long row = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + row * ld_x;
  float* g = dx + row * ld_dx;
  float nn = 0.f, dot = 0.f;
  for (int e = lane; e < d; e += 32) {
    const float v = __ldg(xr + e);
    nn = fmaf(v, v, nn);
    dot = fmaf(v, g[e], dot);
  }
  nn = m_wsum(nn);
  dot = m_wsum(dot);
  const float nrm = sqrtf(nn);
  const bool clamped = nrm <= eps && eps > 0.f;
  const float iv = 1.f / (clamped ? eps : nrm);
  const float c = clamped ? 0.f : dot * iv * iv * iv;       // <xhat, g> / ||x||, spread along xhat
  for (int e = lane; e < d; e += 32) g[e] = g[e] * iv - __ldg(xr + e) * c;
}

// one warp per token: dx[b, slot, :] = J(x[b, slot]) d_pool[b, :] for the valid slots, 0 elsewhere
__global__ void pool_bwd_kernel(const float* __restrict__ x, long long sb, long long ss, int B, int S, int d, int slot0,
                                const int* __restrict__ count, float eps, const float* __restrict__ d_pool,
                                float* __restrict__ dx) {
  const long long tok = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= (long long)B * S) return;
  const int b = (int)(tok / S), slot = (int)(tok % S);
  float* g = dx + tok * d;
  if (slot < slot0 || slot >= slot0 + count[b]) {
    for (int e = lane; e < d; e += 32) g[e] = 0.f;
    return;
  }
  const float* xr = x + (long long)b * sb + (long long)slot * ss;
  const float* up = d_pool + (long long)b * d;
  float nn = 0.f, dot = 0.f;
  for (int e = lane; e < d; e += 32) {
    const float v = __ldg(xr + e);
    nn = fmaf(v, v, nn);
    dot = fmaf(v, __ldg(up + e), dot);
  }
  nn = m_wsum(nn);
  dot = m_wsum(dot);
  const float nrm = sqrtf(nn);
  const bool clamped = nrm <= eps;
  const float iv = 1.f / (clamped ? eps : nrm);
  const float c = clamped ? 0.f : dot * iv * iv * iv;
  for (int e = lane; e < d; e += 32) g[e] = __ldg(up + e) * iv - __ldg(xr + e) * c;
}

}  // namespace alad

extern "C" int alad_order_scores(const float* im, int64_t ld_im, const float* s, int64_t ld_s, int32_t Ni, int32_t Nc,
                                 int32_t d, float* scores, int64_t ldS, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Ni >= 0 && Nc >= 0 && d > 0 && ld_im >= d && ld_s >= d && ldS >= Nc, "alad_order_scores: bad shape");
  if (Ni == 0 || Nc == 0) return ALAD_OK;
  ALAD_REQUIRE(im && s && scores, "alad_order_scores: NULL pointer");
  dim3 grid((Nc + OT - 1) / OT, (Ni + OT - 1) / OT);
  order_fwd_kernel<<<grid, 256, 0, as_stream(stream)>>>(im, ld_im, s, ld_s, Ni, Nc, d, scores, ldS);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_order_scores_bwd(const float* im, int64_t ld_im, const float* s, int64_t ld_s, int32_t Ni,
                                     int32_t Nc, int32_t d, const float* scores, int64_t ldS, const float* G,
                                     int64_t ldG, float* d_im, float* d_s, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Ni >= 0 && Nc >= 0 && d > 0 && ld_im >= d && ld_s >= d && ldS >= Nc && ldG >= Nc,
               "alad_order_scores_bwd: bad shape");
  if (Ni == 0 || Nc == 0) return ALAD_OK;
  ALAD_REQUIRE(im && s && scores && G && d_im && d_s, "alad_order_scores_bwd: NULL pointer");
  const size_t smem = sizeof(float) * (size_t)(Ni > Nc ? Ni : Nc);
  ALAD_REQUIRE(smem <= 200 * 1024, "alad_order_scores_bwd: batch too large for the shared-memory weight row");
  ALAD_CUDA(cudaFuncSetAttribute(order_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  order_bwd_kernel<<<Ni + Nc, 256, smem, as_stream(stream)>>>(im, ld_im, s, ld_s, Ni, Nc, d, scores, ldS, G, ldG, d_im, d_s);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_normalize_bwd(const float* x, int64_t ld_x, int64_t rows, int32_t d, float eps, float* dx,
                                  int64_t ld_dx, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(rows >= 0 && d > 0 && ld_x >= d && ld_dx >= d, "alad_normalize_bwd: bad shape");
  if (rows == 0) return ALAD_OK;
  ALAD_REQUIRE(x && dx, "alad_normalize_bwd: NULL pointer");
  normalize_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, as_stream(stream)>>>(x, ld_x, rows, d, eps, dx, ld_dx);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_pool_tokens_bwd(const float* src, int64_t stride_b, int64_t stride_s, int32_t B, int32_t S, int32_t d,
                                    int32_t slot0, const int32_t* count, float eps, const float* d_pool, float* d_src,
                                    void* stream) {
  using namespace alad;
  ALAD_REQUIRE(B >= 0 && S >= 0 && d > 0 && slot0 >= 0, "alad_pool_tokens_bwd: bad shape");
  const long long n = (long long)B * S;
  if (n == 0) return ALAD_OK;
  ALAD_REQUIRE(src && count && d_pool && d_src, "alad_pool_tokens_bwd: NULL pointer");
  pool_bwd_kernel<<<(unsigned)((n + 7) / 8), 256, 0, as_stream(stream)>>>(src, stride_b, stride_s, B, S, d, slot0, count, eps,
                                                                          d_pool, d_src);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}
