// Aggregation 'scan-sentences' of AlignmentContrastiveLoss.forward (alad/loss.py:136-149): every region attends over
// the words of the caption and is compared with its attended word vector.
//
// The reference materialises B x B x R x W x d tensors (loss.py:143-146).  Here the attended vector never exists:
// with unit word rows Y_j and region x word cosines C = X_i Y_j^T,
//     <x_r, att_r> = sum_w alpha[r,w] C[r,w]          ||att_r||^2 = alpha_r^T K_j alpha_r,   K_j = Y_j Y_j^T
// so a pair needs its R x W cosine block (produced for ALL pairs by the tcgen05 GEMM of mrsw_fwd.cu with the plain
// epilogue) and the W x W Gram matrix of the caption.  This file holds the CUDA-core parts:
//   alad_scan_gram      K_j for every caption                                       (d-sweep, shared-memory tiles)
//   alad_scan_pool_fwd  relu -> L2 norm over regions -> softmax over words -> cosine -> sum over regions
//   alad_scan_pool_bwd  dL/dC (dense, consumed by two more GEMMs) and dL/dK from dL/dS
//   alad_scan_gram_bwd  d Y += 2 dK Y
// One warp per (image, caption) pair, the pair's C / alpha blocks in shared memory; one CTA serves one caption and a
// group of images so that K_j is staged once.
#include <math.h>
#include <stdlib.h>

#include "common.h"

namespace alad {

namespace {

constexpr int SCAN_MAX_EXTENT = 128;       // regions / words per item supported by these kernels
constexpr int SCAN_WARPS = 4;              // warps (pairs in flight) per CTA, reduced when shared memory is short
constexpr int SCAN_IMAGES_PER_CTA = 16;
constexpr float SCAN_NORM_EPS = 1e-12f;    // F.normalize (loss.py:138)
constexpr float SCAN_COS_EPS = 1e-8f;      // F.cosine_similarity (loss.py:146)

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmax(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// ------------------------------------------------------------------------------------------------ Gram matrices
constexpr int GRAM_K = 32;
constexpr int GRAM_ACC = 8;                // entries per thread and pass (256 threads -> 2048 entries per pass)

__global__ void __launch_bounds__(256) scan_gram_kernel(const float* __restrict__ yh, int W, int d,
                                                        const int* __restrict__ nw, float* __restrict__ K) {
  __shared__ float ys[SCAN_MAX_EXTENT][GRAM_K + 1];
  const int j = blockIdx.x;
  const int n = min(max(nw[j], 0), W);
  const float* y = yh + (long long)j * W * d;
  float* Kj = K + (long long)j * W * W;
  for (int e = threadIdx.x; e < W * W; e += 256) {                       // entries outside the valid block
    const int w = e / W, w2 = e % W;
    if (w >= n || w2 >= n) Kj[e] = 0.f;
  }
  const int total = n * n;
  for (int e0 = 0; e0 < total; e0 += 256 * GRAM_ACC) {
    float acc[GRAM_ACC];
    int wa[GRAM_ACC], wb[GRAM_ACC];
#pragma unroll
    for (int a = 0; a < GRAM_ACC; ++a) {
      const int e = e0 + a * 256 + threadIdx.x;
      acc[a] = 0.f;
      wa[a] = e < total ? e / n : 0;
      wb[a] = e < total ? e % n : 0;
    }
    for (int k0 = 0; k0 < d; k0 += GRAM_K) {
      __syncthreads();
      for (int e = threadIdx.x; e < n * GRAM_K; e += 256) {
        const int w = e / GRAM_K, k = e % GRAM_K;
        ys[w][k] = (k0 + k < d) ? __ldg(y + (long long)w * d + k0 + k) : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int a = 0; a < GRAM_ACC; ++a) {
        float s = acc[a];
#pragma unroll
        for (int k = 0; k < GRAM_K; ++k) s = fmaf(ys[wa[a]][k], ys[wb[a]][k], s);
        acc[a] = s;
      }
    }
#pragma unroll
    for (int a = 0; a < GRAM_ACC; ++a) {
      const int e = e0 + a * 256 + threadIdx.x;
      if (e < total) Kj[wa[a] * W + wb[a]] = acc[a];
    }
  }
}

// d_yh[j, w, :] += 2 * sum_w' dK[j, w, w'] * yh[j, w', :]   (dK is symmetric by construction: sum of c * alpha alpha^T)
__global__ void __launch_bounds__(128) scan_gram_bwd_kernel(const float* __restrict__ yh, int W, int d,
                                                            const int* __restrict__ nw, const float* __restrict__ dK,
                                                            float* __restrict__ d_yh) {
  const int j = blockIdx.x;
  const int k = blockIdx.y * 128 + threadIdx.x;
  const int n = min(max(nw[j], 0), W);
  if (k >= d) return;
  const float* y = yh + (long long)j * W * d + k;
  const float* g = dK + (long long)j * W * W;
  float* out = d_yh + (long long)j * W * d + k;
  for (int w = 0; w < n; ++w) {
    float acc = 0.f;
    for (int w2 = 0; w2 < n; ++w2) acc = fmaf(__ldg(g + w * W + w2), __ldg(y + (long long)w2 * d), acc);
    out[(long long)w * d] += 2.f * acc;
  }
}

// ------------------------------------------------------------------------------------------------ pair kernels
struct ScanPairArgs {
  const float* C;  long long ldC;          // [Bi*R, ldC] cosines, column j*W + w
  int Bi, R, Bc, W;
  int Rcap, Wcap;                          // upper bounds of nr / nw (shared-memory extents)
  const int* nr;  const int* nw;
  const float* K;                          // [Bc, W, W]
  float* S;  long long ldS;                // forward output
  const float* G;  long long ldG;          // backward: dL/dS
  float* dC;  long long lddC;              // backward: [Bi*R, lddC], zeroed by the entry point
  float* dK;                               // backward: [Bc, W, W], accumulated (the caller zeroes it)
  int warps;                               // warps per CTA
};

// Shared memory (floats): Ks[Wcap][Wp] | (bwd) dKs[Wcap][Wp] | per warp: Cs[Rcap][Wp], As[Rcap][Wp], (bwd) Ts[Rcap][Wp],
// inv[Wp], (bwd) cu[Rcap], cv[Rcap]
__host__ __device__ inline int scan_wp(int Wcap) { return Wcap | 1; }
inline size_t scan_smem_floats(int Rcap, int Wcap, int warps, bool bwd) {
  const size_t Wp = scan_wp(Wcap);
  const size_t per_warp = (size_t)Rcap * Wp * (bwd ? 3 : 2) + Wp + (bwd ? 2 * (size_t)Rcap : 0);
  return (size_t)Wcap * Wp * (bwd ? 2 : 1) + per_warp * warps;
}

// Steps shared by both directions: stage the pair's cosines, column norms over the regions (relu'd), then per
// region the softmax over the words.  Returns nothing; As holds alpha, inv holds 1 / max(||P[:, w]||, eps).
__device__ __forceinline__ void scan_stage_pair(const ScanPairArgs& a, int i, int j, int nri, int nwj, int Wp, float* Cs,
                                                float* inv, int lane) {
  const float* src = a.C + (long long)i * a.R * a.ldC + (long long)j * a.W;
  for (int r = 0; r < nri; ++r)
    for (int w = lane; w < nwj; w += 32) Cs[r * Wp + w] = __ldg(src + (long long)r * a.ldC + w);
  __syncwarp();
  for (int w = lane; w < nwj; w += 32) {
    float acc = 0.f;
    for (int r = 0; r < nri; ++r) {
      const float p = fmaxf(Cs[r * Wp + w], 0.f);
      acc = fmaf(p, p, acc);
    }
    inv[w] = 1.f / fmaxf(sqrtf(acc), SCAN_NORM_EPS);
  }
  __syncwarp();
}

// softmax over the words of region r -> As[r]; returns u_r = sum_w alpha C (all lanes)
__device__ __forceinline__ float scan_softmax_row(int r, int nwj, int Wp, const float* Cs, float* As, const float* inv,
                                                  int lane) {
  float m = -INFINITY;
  for (int w = lane; w < nwj; w += 32) {
    const float q = fmaxf(Cs[r * Wp + w], 0.f) * inv[w];
    As[r * Wp + w] = q;
    m = fmaxf(m, q);
  }
  m = wmax(m);
  float sum = 0.f;
  for (int w = lane; w < nwj; w += 32) {
    const float e = expf(As[r * Wp + w] - m);
    As[r * Wp + w] = e;
    sum += e;
  }
  sum = wsum(sum);
  const float rs = 1.f / sum;
  float u = 0.f;
  for (int w = lane; w < nwj; w += 32) {
    const float al = As[r * Wp + w] * rs;
    As[r * Wp + w] = al;
    u = fmaf(al, Cs[r * Wp + w], u);
  }
  __syncwarp();
  return wsum(u);
}

template <bool kBwd>
__global__ void __launch_bounds__(SCAN_WARPS * 32) scan_pair_kernel(const ScanPairArgs a) {
  extern __shared__ float smem[];
  const int j = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int Wp = scan_wp(a.Wcap);
  const int nwj = min(max(a.nw[j], 0), a.Wcap);
  float* Ks = smem;
  float* dKs = Ks + a.Wcap * Wp;
  float* wbase = (kBwd ? dKs + a.Wcap * Wp : dKs) +
                 (size_t)warp * ((size_t)a.Rcap * Wp * (kBwd ? 3 : 2) + Wp + (kBwd ? 2 * a.Rcap : 0));
  float* Cs = wbase;
  float* As = Cs + a.Rcap * Wp;
  float* Ts = As + a.Rcap * Wp;                       // backward only
  float* inv = kBwd ? Ts + a.Rcap * Wp : Ts;
  float* cus = inv + Wp;                              // backward only
  float* cvs = cus + a.Rcap;

  const float* Kj = a.K + (long long)j * a.W * a.W;
  for (int e = threadIdx.x; e < nwj * nwj; e += blockDim.x) {
    const int w = e / nwj, w2 = e % nwj;
    Ks[w * Wp + w2] = __ldg(Kj + w * a.W + w2);
    if (kBwd) dKs[w * Wp + w2] = 0.f;
  }
  __syncthreads();

  const int i_begin = blockIdx.y * SCAN_IMAGES_PER_CTA;
  const int i_end = min(i_begin + SCAN_IMAGES_PER_CTA, a.Bi);
  bool touched = false;
  if (warp < a.warps) {
    for (int i = i_begin + warp; i < i_end; i += a.warps) {
      const int nri = min(max(a.nr[i], 0), a.Rcap);
      if (!kBwd) {
        if (nri == 0 || nwj == 0) {
          // no valid region: every term is masked (loss.py:147) -> 0; no valid word: softmax over an all -inf
          // row (loss.py:139-140) -> NaN for every unmasked region
          if (lane == 0) a.S[(long long)i * a.ldS + j] = nri == 0 ? 0.f : __int_as_float(0x7fc00000);
          continue;
        }
      } else if (nri == 0 || nwj == 0) {
        continue;
      }
      float g = 0.f;
      if (kBwd) {
        g = __ldg(a.G + (long long)i * a.ldG + j);
        if (g == 0.f) continue;
      }
      scan_stage_pair(a, i, j, nri, nwj, Wp, Cs, inv, lane);
      float total = 0.f;
      for (int r = 0; r < nri; ++r) {
        const float u = scan_softmax_row(r, nwj, Wp, Cs, As, inv, lane);
        // t_w = (K alpha_r)[w], v = alpha_r^T K alpha_r
        float v = 0.f;
        for (int w = lane; w < nwj; w += 32) {
          float t = 0.f;
          for (int w2 = 0; w2 < nwj; ++w2) t = fmaf(Ks[w * Wp + w2], As[r * Wp + w2], t);
          if (kBwd) Ts[r * Wp + w] = t;
          v = fmaf(As[r * Wp + w], t, v);
        }
        v = wsum(v);
        const float b = sqrtf(fmaxf(v, 0.f));
        if (!kBwd) {
          total += u / fmaxf(b, SCAN_COS_EPS);
        } else {
          const bool live = b > SCAN_COS_EPS;                           // clamp of cosine_similarity inactive
          const float cu = live ? g / b : g / SCAN_COS_EPS;             // g * d new / d u
          const float cv2 = live ? -g * u / (b * b * b) : 0.f;          // g * 2 d new / d v
          float dot = 0.f;
          for (int w = lane; w < nwj; w += 32) {
            const float da = fmaf(cu, Cs[r * Wp + w], cv2 * Ts[r * Wp + w]);
            Ts[r * Wp + w] = da;
            dot = fmaf(As[r * Wp + w], da, dot);
          }
          dot = wsum(dot);
          for (int w = lane; w < nwj; w += 32) Ts[r * Wp + w] = As[r * Wp + w] * (Ts[r * Wp + w] - dot);   // dL/dQ
          if (lane == 0) {
            cus[r] = cu;
            cvs[r] = 0.5f * cv2;
          }
        }
      }
      if (!kBwd) {
        if (lane == 0) a.S[(long long)i * a.ldS + j] = total;
        __syncwarp();
        continue;
      }
      __syncwarp();
      // column pass: Jacobian of the L2 normalisation over the regions, relu mask, direct path through u
      float* dst = a.dC + (long long)i * a.R * a.lddC + (long long)j * a.W;
      for (int w = lane; w < nwj; w += 32) {
        const float iw = inv[w];
        const bool clamped = iw >= 1.f / SCAN_NORM_EPS;                 // ||P[:, w]|| <= eps: norm treated as constant
        float dot = 0.f;
        for (int r = 0; r < nri; ++r) dot = fmaf(fmaxf(Cs[r * Wp + w], 0.f) * iw, Ts[r * Wp + w], dot);
        if (clamped) dot = 0.f;
        for (int r = 0; r < nri; ++r) {
          const float c = Cs[r * Wp + w];
          const float q = fmaxf(c, 0.f) * iw;
          const float dP = (Ts[r * Wp + w] - q * dot) * iw;
          const float al = As[r * Wp + w];
          dst[(long long)r * a.lddC + w] = (c > 0.f ? dP : 0.f) + cus[r] * al;
          Ts[r * Wp + w] = cvs[r] * al;                                 // reused below: c_r * alpha[r, w]
        }
      }
      __syncwarp();
      // dK[w, w'] += sum_r c_r alpha[r, w] alpha[r, w']
      for (int w = lane; w < nwj; w += 32)
        for (int w2 = 0; w2 < nwj; ++w2) {
          float acc = 0.f;
          for (int r = 0; r < nri; ++r) acc = fmaf(Ts[r * Wp + w], As[r * Wp + w2], acc);
          atomicAdd(&dKs[w * Wp + w2], acc);
        }
      touched = true;
      __syncwarp();
    }
  }
  if (kBwd) {
    const int any = __syncthreads_or(touched ? 1 : 0);
    if (any) {
      float* out = a.dK + (long long)j * a.W * a.W;
      for (int e = threadIdx.x; e < nwj * nwj; e += blockDim.x) {
        const int w = e / nwj, w2 = e % nwj;
        atomicAdd(out + w * a.W + w2, dKs[w * Wp + w2]);
      }
    }
  }
}

// ------------------------------------------------------------------------------------------------ forward, <= 64 words
// Forward pooling with the caption's Gram rows in REGISTERS (lane l owns words l and l + 32, so 2 x 64 registers hold
// its two rows of K_j for all the images of the CTA) and no per-pair shared-memory block: the lane reads its two
// columns of the pair's cosine block straight from global memory (twice: column norms, then row by row; the second
// read hits L1), only the current row of exp(q) goes through a 256-byte double buffer for the K e product, read
// back as LDS.128 broadcasts.  q = relu(C) / ||relu(C)[:, w]|| lies in [0, 1], so the softmax needs no max shift, and
// with e = exp(q), s = sum e:  <x, att> = (e . c) / s,  ||att|| = sqrt(e' K e) / s  -- the three sums share one
// interleaved butterfly.  About 180 instructions per region row instead of ~450 and 4 dependent reductions.
constexpr int SCAN_REG_W = 64;

__global__ void __launch_bounds__(SCAN_WARPS * 32, 3) scan_pair_fwd_reg_kernel(const ScanPairArgs a) {
  __shared__ __align__(16) float erow[SCAN_WARPS][2][SCAN_REG_W];
  const int j = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nwj = min(max(a.nw[j], 0), SCAN_REG_W);
  const int w0 = lane, w1 = lane + 32;
  const bool has0 = w0 < nwj, has1 = w1 < nwj;
  float k0[SCAN_REG_W], k1[SCAN_REG_W];
  {
    const float* Kj = a.K + (long long)j * a.W * a.W;
#pragma unroll
    for (int w2 = 0; w2 < SCAN_REG_W; ++w2) {
      k0[w2] = (has0 && w2 < nwj) ? __ldg(Kj + w0 * a.W + w2) : 0.f;
      k1[w2] = (has1 && w2 < nwj) ? __ldg(Kj + w1 * a.W + w2) : 0.f;
    }
  }
  const int i_begin = blockIdx.y * SCAN_IMAGES_PER_CTA;
  const int i_end = min(i_begin + SCAN_IMAGES_PER_CTA, a.Bi);
  for (int i = i_begin + warp; i < i_end; i += SCAN_WARPS) {
    const int nri = min(max(a.nr[i], 0), a.R);
    if (nri == 0 || nwj == 0) {
      if (lane == 0) a.S[(long long)i * a.ldS + j] = nri == 0 ? 0.f : __int_as_float(0x7fc00000);
      continue;
    }
    const float* src = a.C + (long long)i * a.R * a.ldC + (long long)j * a.W;
    float n0 = 0.f, n1 = 0.f;
    for (int r = 0; r < nri; ++r) {
      const float p0 = has0 ? fmaxf(__ldg(src + (long long)r * a.ldC + w0), 0.f) : 0.f;
      const float p1 = has1 ? fmaxf(__ldg(src + (long long)r * a.ldC + w1), 0.f) : 0.f;
      n0 = fmaf(p0, p0, n0);
      n1 = fmaf(p1, p1, n1);
    }
    const float inv0 = 1.f / fmaxf(sqrtf(n0), SCAN_NORM_EPS), inv1 = 1.f / fmaxf(sqrtf(n1), SCAN_NORM_EPS);
    float total = 0.f;
    float c0 = has0 ? __ldg(src + w0) : 0.f, c1 = has1 ? __ldg(src + w1) : 0.f;
    for (int r = 0; r < nri; ++r) {
      float nc0 = 0.f, nc1 = 0.f;
      if (r + 1 < nri) {                                            // next row's cosines while this one is processed
        nc0 = has0 ? __ldg(src + (long long)(r + 1) * a.ldC + w0) : 0.f;
        nc1 = has1 ? __ldg(src + (long long)(r + 1) * a.ldC + w1) : 0.f;
      }
      float* buf = erow[warp][r & 1];
      const float e0 = has0 ? expf(fmaxf(c0, 0.f) * inv0) : 0.f;
      const float e1 = has1 ? expf(fmaxf(c1, 0.f) * inv1) : 0.f;
      buf[w0] = e0;
      buf[w1] = e1;
      __syncwarp();
      float t0 = 0.f, t1 = 0.f;
#pragma unroll
      for (int g = 0; g < SCAN_REG_W / 4; ++g) {
        if (4 * g < nwj) {
          const float4 e = *reinterpret_cast<const float4*>(buf + 4 * g);
          t0 = fmaf(k0[4 * g + 0], e.x, t0); t1 = fmaf(k1[4 * g + 0], e.x, t1);
          t0 = fmaf(k0[4 * g + 1], e.y, t0); t1 = fmaf(k1[4 * g + 1], e.y, t1);
          t0 = fmaf(k0[4 * g + 2], e.z, t0); t1 = fmaf(k1[4 * g + 2], e.z, t1);
          t0 = fmaf(k0[4 * g + 3], e.w, t0); t1 = fmaf(k1[4 * g + 3], e.w, t1);
        }
      }
      float s = e0 + e1;
      float u = fmaf(e0, c0, e1 * c1);
      float v = fmaf(e0, t0, e1 * t1);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, o);
        u += __shfl_xor_sync(0xffffffffu, u, o);
        v += __shfl_xor_sync(0xffffffffu, v, o);
      }
      const float rs = 1.f / s;
      total += (u * rs) / fmaxf(sqrtf(fmaxf(v, 0.f)) * rs, SCAN_COS_EPS);
      c0 = nc0;
      c1 = nc1;
      // the other half of the double buffer is written next; this half again two rows on, after every lane has
      // passed the shuffles of the row in between
    }
    if (lane == 0) a.S[(long long)i * a.ldS + j] = total;
  }
}

// ------------------------------------------------------------------------------------------------ sparse backward
// d xhat[i, r, :] += sum_w dC[i, r, j, w] yhat[j, w, :]  and  d yhat[j, w, :] += sum_r dC[i, r, j, w] xhat[i, r, :]
// for a LIST of (i, j) pairs: with the hardest-negative hinge dL/dS has <= 3B non-zero entries, so the two dense
// GEMMs over all B x B blocks of dC would spend > 99 % of their work (and of the pack / transpose passes that feed
// them) on zeros.  One CTA per pair and 256 feature columns: the pair's dC block in shared memory, one thread per
// column, results added with float atomics (several pairs share an image or a caption).
__global__ void __launch_bounds__(256) scan_apply_pairs_kernel(const float* __restrict__ dC, long long lddC,
                                                               const float* __restrict__ xh, const float* __restrict__ yh,
                                                               const int* __restrict__ pairs, int R, int W, int d,
                                                               const int* __restrict__ nr, const int* __restrict__ nw,
                                                               int Rcap, int Wcap, float* __restrict__ d_xh,
                                                               float* __restrict__ d_yh) {
  extern __shared__ float blk[];                       // [Rcap][Wp]
  const int Wp = scan_wp(Wcap);
  const int i = pairs[2 * blockIdx.x], j = pairs[2 * blockIdx.x + 1];
  const int nri = min(max(nr[i], 0), Rcap), nwj = min(max(nw[j], 0), Wcap);
  const float* src = dC + (long long)i * R * lddC + (long long)j * W;
  for (int e = threadIdx.x; e < nri * nwj; e += blockDim.x) {
    const int r = e / nwj, w = e % nwj;
    blk[r * Wp + w] = __ldg(src + (long long)r * lddC + w);
  }
  __syncthreads();
  const int k = blockIdx.y * blockDim.x + threadIdx.x;
  if (k >= d) return;
  const float* x = xh + (long long)i * R * d + k;
  const float* y = yh + (long long)j * W * d + k;
  for (int r = 0; r < nri; ++r) {
    float acc = 0.f;
    for (int w = 0; w < nwj; ++w) acc = fmaf(blk[r * Wp + w], __ldg(y + (long long)w * d), acc);
    atomicAdd(d_xh + ((long long)i * R + r) * d + k, acc);
  }
  for (int w = 0; w < nwj; ++w) {
    float acc = 0.f;
    for (int r = 0; r < nri; ++r) acc = fmaf(blk[r * Wp + w], __ldg(x + (long long)r * d), acc);
    atomicAdd(d_yh + ((long long)j * W + w) * d + k, acc);
  }
}

int scan_check(const char* what, int32_t Bi, int32_t R, int32_t Bc, int32_t W, int32_t max_nr, int32_t max_nw, int64_t ldC) {
  ALAD_REQUIRE(Bi >= 0 && Bc >= 0 && R >= 0 && W >= 0, "%s: bad shape", what);
  ALAD_REQUIRE(max_nr >= 0 && max_nr <= R && max_nw >= 0 && max_nw <= W, "%s: max_nr / max_nw outside the extents", what);
  ALAD_REQUIRE(ldC >= (int64_t)Bc * W, "%s: ldC too small", what);
  if (max_nr > SCAN_MAX_EXTENT || max_nw > SCAN_MAX_EXTENT)
    return fail(ALAD_ERR_UNSUPPORTED, "%s: at most %d scored regions / words per item (got %d / %d)", what, SCAN_MAX_EXTENT,
                max_nr, max_nw);
  ALAD_REQUIRE((Bi + SCAN_IMAGES_PER_CTA - 1) / SCAN_IMAGES_PER_CTA <= 65535, "%s: too many images per call (%d)", what, Bi);
  return ALAD_OK;
}

template <bool kBwd>
int scan_launch(const char* what, ScanPairArgs& p, cudaStream_t st) {
  int dev = 0, max_smem = 0;
  ALAD_CUDA(cudaGetDevice(&dev));
  ALAD_CUDA(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int warps = SCAN_WARPS;
  while (warps > 1 && scan_smem_floats(p.Rcap, p.Wcap, warps, kBwd) * 4 > (size_t)max_smem) --warps;
  const size_t bytes = scan_smem_floats(p.Rcap, p.Wcap, warps, kBwd) * 4;
  if (bytes > (size_t)max_smem)
    return fail(ALAD_ERR_UNSUPPORTED, "%s: %d regions x %d words need %zu bytes of shared memory (%d available)", what,
                p.Rcap, p.Wcap, bytes, max_smem);
  p.warps = warps;
  ALAD_CUDA(cudaFuncSetAttribute(scan_pair_kernel<kBwd>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  dim3 grid(p.Bc, (p.Bi + SCAN_IMAGES_PER_CTA - 1) / SCAN_IMAGES_PER_CTA);
  scan_pair_kernel<kBwd><<<grid, SCAN_WARPS * 32, bytes, st>>>(p);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

}  // namespace
}  // namespace alad

extern "C" int alad_scan_gram(const float* yh, int32_t Bc, int32_t W, int32_t d, const int32_t* nw, float* K, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Bc >= 0 && W >= 0 && d > 0, "alad_scan_gram: bad shape");
  if (W > SCAN_MAX_EXTENT) return fail(ALAD_ERR_UNSUPPORTED, "alad_scan_gram: at most %d words per caption (got %d)", SCAN_MAX_EXTENT, W);
  if (Bc == 0 || W == 0) return ALAD_OK;
  ALAD_REQUIRE(yh && nw && K, "alad_scan_gram: NULL pointer");
  scan_gram_kernel<<<Bc, 256, 0, as_stream(stream)>>>(yh, W, d, nw, K);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_scan_gram_bwd(const float* yh, int32_t Bc, int32_t W, int32_t d, const int32_t* nw, const float* dK,
                                  float* d_yh, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Bc >= 0 && W >= 0 && d > 0, "alad_scan_gram_bwd: bad shape");
  if (Bc == 0 || W == 0) return ALAD_OK;
  ALAD_REQUIRE(yh && nw && dK && d_yh, "alad_scan_gram_bwd: NULL pointer");
  ALAD_REQUIRE((d + 127) / 128 <= 65535, "alad_scan_gram_bwd: d too large");
  dim3 grid(Bc, (d + 127) / 128);
  scan_gram_bwd_kernel<<<grid, 128, 0, as_stream(stream)>>>(yh, W, d, nw, dK, d_yh);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_scan_pool_fwd(const float* C, int64_t ldC, int32_t Bi, int32_t R, int32_t Bc, int32_t W,
                                  const int32_t* nr, const int32_t* nw, int32_t max_nr, int32_t max_nw, const float* K,
                                  float* S, int64_t ldS, void* stream) {
  using namespace alad;
  const int rc = scan_check("alad_scan_pool_fwd", Bi, R, Bc, W, max_nr, max_nw, ldC);
  if (rc) return rc;
  ALAD_REQUIRE(ldS >= Bc, "alad_scan_pool_fwd: ldS too small");
  if (Bi == 0 || Bc == 0) return ALAD_OK;
  ALAD_REQUIRE(nr && nw && S && (C || max_nr == 0 || max_nw == 0) && (K || max_nw == 0), "alad_scan_pool_fwd: NULL pointer");
  ScanPairArgs p = {};
  p.C = C; p.ldC = ldC; p.Bi = Bi; p.R = R; p.Bc = Bc; p.W = W;
  p.Rcap = max_nr > 0 ? max_nr : 1; p.Wcap = max_nw > 0 ? max_nw : 1;
  p.nr = nr; p.nw = nw; p.K = K; p.S = S; p.ldS = ldS;
  if (max_nw <= SCAN_REG_W && getenv("ALAD_SCAN_SMEM_FWD") == nullptr) {      // env: A/B switch to the shared-memory kernel
    dim3 grid(Bc, (Bi + SCAN_IMAGES_PER_CTA - 1) / SCAN_IMAGES_PER_CTA);
    scan_pair_fwd_reg_kernel<<<grid, SCAN_WARPS * 32, 0, as_stream(stream)>>>(p);
    ALAD_CUDA(cudaGetLastError());
    return ALAD_OK;
  }
  return scan_launch<false>("alad_scan_pool_fwd", p, as_stream(stream));
}

extern "C" int alad_scan_pool_bwd(const float* C, int64_t ldC, int32_t Bi, int32_t R, int32_t Bc, int32_t W,
                                  const int32_t* nr, const int32_t* nw, int32_t max_nr, int32_t max_nw, const float* K,
                                  const float* G, int64_t ldG, float* dC, int64_t lddC, float* dK, void* stream) {
  using namespace alad;
  const int rc = scan_check("alad_scan_pool_bwd", Bi, R, Bc, W, max_nr, max_nw, ldC);
  if (rc) return rc;
  ALAD_REQUIRE(ldG >= Bc && lddC >= (int64_t)Bc * W, "alad_scan_pool_bwd: leading dimension too small");
  if (Bi == 0 || Bc == 0 || R == 0 || W == 0) return ALAD_OK;
  ALAD_REQUIRE(C && nr && nw && K && G && dC && dK, "alad_scan_pool_bwd: NULL pointer");
  cudaStream_t st = as_stream(stream);
  ALAD_CUDA(cudaMemset2DAsync(dC, (size_t)lddC * 4, 0, (size_t)Bc * W * 4, (size_t)Bi * R, st));
  if (max_nr == 0 || max_nw == 0) return ALAD_OK;
  ScanPairArgs p = {};
  p.C = C; p.ldC = ldC; p.Bi = Bi; p.R = R; p.Bc = Bc; p.W = W;
  p.Rcap = max_nr; p.Wcap = max_nw;
  p.nr = nr; p.nw = nw; p.K = K; p.G = G; p.ldG = ldG; p.dC = dC; p.lddC = lddC; p.dK = dK;
  return scan_launch<true>("alad_scan_pool_bwd", p, st);
}

extern "C" int alad_scan_apply_pairs(const float* dC, int64_t lddC, const float* xh, const float* yh, const int32_t* pairs,
                                     int32_t n_pairs, int32_t Bi, int32_t R, int32_t Bc, int32_t W, int32_t d,
                                     const int32_t* nr, const int32_t* nw, int32_t max_nr, int32_t max_nw, float* d_xh,
                                     float* d_yh, void* stream) {
  using namespace alad;
  const int rc = scan_check("alad_scan_apply_pairs", Bi, R, Bc, W, max_nr, max_nw, lddC);
  if (rc) return rc;
  ALAD_REQUIRE(n_pairs >= 0 && d > 0, "alad_scan_apply_pairs: bad shape");
  if (n_pairs == 0 || max_nr == 0 || max_nw == 0) return ALAD_OK;
  ALAD_REQUIRE(dC && xh && yh && pairs && nr && nw && d_xh && d_yh, "alad_scan_apply_pairs: NULL pointer");
  ALAD_REQUIRE((d + 255) / 256 <= 65535, "alad_scan_apply_pairs: d too large");
  const size_t bytes = (size_t)max_nr * scan_wp(max_nw) * 4;
  ALAD_CUDA(cudaFuncSetAttribute(scan_apply_pairs_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
  dim3 grid(n_pairs, (d + 255) / 256);
  scan_apply_pairs_kernel<<<grid, 256, bytes, as_stream(stream)>>>(dC, lddC, xh, yh, pairs, R, W, d, nr, nw, max_nr, max_nw,
                                                                   d_xh, d_yh);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}
