// alad_mrsw_scores_bwd: gradient of sum(G * S) w.r.t. the raw (un-normalised) token tensors,
// S = MrSw alignment scores (alad/loss.py:79-125).  Follows autograd through
// masked_fill_(0) / max(regions) / sum(words) / F.normalize:
//   1. compact the non-zero entries of G = g0*G0 + G1 into a pair list (with hardest-negative
//      mining G has <= 3B non-zeros; a dense G is handled by the same path),
//   2. per pair: recompute the (regions x words) cosine tile in fp32, arg-max over regions
//      (first occurrence, masked regions count as 0), scatter-add g*word into the winning
//      region's gradient and g*region into the word's gradient,
//   3. apply the F.normalize Jacobian per token in place.
#include <math.h>

#include <algorithm>

#include "common.h"

namespace alad {

constexpr int BW_THREADS = 256;
constexpr int BW_WARPS = BW_THREADS / 32;

struct Pair {
  int i, j;
  float g;
};

struct BwdParams {
  const float* im;
  long long im_sb, im_ss;
  const float* s;
  long long s_sb, s_ss;
  int Bi, S_im, Bc, S_s, d, R;
  const int* nr;
  const int* nw;
  const float* G0;
  long long ldG0;
  const float* g0_scale;
  const float* G1;
  long long ldG1;
  float* d_im;      // [Bi, S_im, d] contiguous
  float* d_s;       // [Bc, S_s, d] contiguous
  float* inv_im;    // [Bi * S_im]  1 / max(||x||, eps)
  float* inv_s;     // [Bc * S_s]
  Pair* pairs;
  unsigned int* n_pairs;
  unsigned int max_pairs;
  float eps;
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// one warp per token: inverse norms of every slot
__global__ void inv_norm_kernel(const float* __restrict__ x, long long sb, long long ss, int B, int S, int d, float eps,
                                float* __restrict__ inv) {
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= B * S) return;
  const float* row = x + (long long)(tok / S) * sb + (long long)(tok % S) * ss;
  float acc = 0.f;
  for (int e = lane; e < d; e += 32) {
    const float v = __ldg(row + e);
    acc += v * v;
  }
  acc = wsum(acc);
  if (lane == 0) inv[tok] = 1.f / fmaxf(sqrtf(acc), eps);
}

__global__ void compact_kernel(const BwdParams p) {
  const long long n = (long long)p.Bi * p.Bc;
  const float scale = p.g0_scale ? *p.g0_scale : 1.f;
  for (long long e = blockIdx.x * (long long)blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(e / p.Bc), j = (int)(e % p.Bc);
    float g = 0.f;
    if (p.G0) g += scale * p.G0[(long long)i * p.ldG0 + j];
    if (p.G1) g += p.G1[(long long)i * p.ldG1 + j];
    if (g != 0.f && p.nr[i] > 0 && p.nw[j] > 0) {
      const unsigned int at = atomicAdd(p.n_pairs, 1u);
      if (at < p.max_pairs) p.pairs[at] = Pair{i, j, g};
    }
  }
}

// persistent CTAs loop over the pair list; one warp per word of the caption
__global__ void __launch_bounds__(BW_THREADS) pair_kernel(const BwdParams p) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned int n_pairs = min(*p.n_pairs, p.max_pairs);
  const int d = p.d;
  for (unsigned int q = blockIdx.x; q < n_pairs; q += gridDim.x) {
    const Pair pr = p.pairs[q];
    const int nr = p.nr[pr.i], nw = p.nw[pr.j];
    const bool clamp = nr < p.R;
    const float* im_i = p.im + (long long)pr.i * p.im_sb + p.im_ss;        // slot 1
    const float* s_j = p.s + (long long)pr.j * p.s_sb + p.s_ss;
    const float* inv_i = p.inv_im + (long long)pr.i * p.S_im + 1;
    const float* inv_j = p.inv_s + (long long)pr.j * p.S_s + 1;
    for (int w = warp; w < nw; w += BW_WARPS) {
      const float* xw = s_j + (long long)w * p.s_ss;
      const float iw = inv_j[w];
      float best = -INFINITY;
      int rbest = -1;
      for (int r = 0; r < nr; ++r) {
        const float* xr = im_i + (long long)r * p.im_ss;
        float acc = 0.f;
        for (int e = lane; e < d; e += 32) acc += __ldg(xr + e) * __ldg(xw + e);
        acc = wsum(acc) * inv_i[r] * iw;
        if (acc > best) {            // strict: first occurrence wins, like torch.max on CPU
          best = acc;
          rbest = r;
        }
      }
      // masked regions take part in the max with value 0 and sit after the valid ones
      if (rbest < 0 || (clamp && best < 0.f)) continue;
      const float* xr = im_i + (long long)rbest * p.im_ss;
      const float ir = inv_i[rbest];
      float* gi = p.d_im + ((long long)pr.i * p.S_im + 1 + rbest) * d;
      float* gs = p.d_s + ((long long)pr.j * p.S_s + 1 + w) * d;
      for (int e = lane; e < d; e += 32) {
        atomicAdd(gi + e, pr.g * (__ldg(xw + e) * iw));
        atomicAdd(gs + e, pr.g * (__ldg(xr + e) * ir));
      }
    }
  }
}

// in place: dx = (dxhat - xhat * <xhat, dxhat>) / max(||x||, eps); below eps the norm is the constant eps
__global__ void norm_bwd_kernel(const float* __restrict__ x, long long sb, long long ss, int B, int S, int d, float eps,
                                const float* __restrict__ inv, float* __restrict__ dx) {
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= B * S) return;
  const float* row = x + (long long)(tok / S) * sb + (long long)(tok % S) * ss;
  float* g = dx + (long long)tok * d;
  const float iv = inv[tok];
  float dot = 0.f;
  for (int e = lane; e < d; e += 32) dot += (__ldg(row + e) * iv) * g[e];
  dot = wsum(dot);
  const bool clamped = iv >= 1.f / eps;      // ||x|| <= eps: x / eps, no projection term
  for (int e = lane; e < d; e += 32) {
    const float xh = __ldg(row + e) * iv;
    g[e] = clamped ? g[e] * iv : (g[e] - xh * dot) * iv;
  }
}

}  // namespace alad

extern "C" int64_t alad_mrsw_bwd_workspace_bytes(int32_t Bi, int32_t S_im, int32_t Bc, int32_t S_s, int64_t max_pairs) {
  return 4ll * ((int64_t)Bi * S_im + (int64_t)Bc * S_s) + 12ll * max_pairs + 256;
}

extern "C" int alad_mrsw_scores_bwd(const alad_mrsw_bwd_args* a, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(a != nullptr, "alad_mrsw_scores_bwd: NULL args");
  ALAD_REQUIRE(a->Bi >= 0 && a->Bc >= 0 && a->d > 0 && a->S_im >= 0 && a->S_s >= 0, "alad_mrsw_scores_bwd: bad shape");
  ALAD_REQUIRE(a->d_im && a->d_s && a->workspace, "alad_mrsw_scores_bwd: NULL output/workspace");
  ALAD_REQUIRE(a->max_pairs > 0 && a->workspace_bytes >= alad_mrsw_bwd_workspace_bytes(a->Bi, a->S_im, a->Bc, a->S_s, a->max_pairs),
               "alad_mrsw_scores_bwd: workspace too small");
  cudaStream_t st = as_stream(stream);
  const size_t n_im = (size_t)a->Bi * a->S_im, n_s = (size_t)a->Bc * a->S_s;
  if (n_im) ALAD_CUDA(cudaMemsetAsync(a->d_im, 0, n_im * a->d * sizeof(float), st));
  if (n_s) ALAD_CUDA(cudaMemsetAsync(a->d_s, 0, n_s * a->d * sizeof(float), st));
  if (a->Bi == 0 || a->Bc == 0 || a->S_im < 2 || a->S_s < 2 || (!a->G0 && !a->G1)) return ALAD_OK;
  ALAD_REQUIRE(a->im && a->s && a->nr && a->nw, "alad_mrsw_scores_bwd: NULL input");
  BwdParams p;
  p.im = a->im; p.im_sb = a->im_stride_b; p.im_ss = a->im_stride_s;
  p.s = a->s; p.s_sb = a->s_stride_b; p.s_ss = a->s_stride_s;
  p.Bi = a->Bi; p.S_im = a->S_im; p.Bc = a->Bc; p.S_s = a->S_s; p.d = a->d; p.R = a->region_extent > 0 ? a->region_extent : a->S_im - 1;
  p.nr = a->nr; p.nw = a->nw;
  p.G0 = a->G0; p.ldG0 = a->ldG0; p.g0_scale = a->g0_scale; p.G1 = a->G1; p.ldG1 = a->ldG1;
  p.d_im = a->d_im; p.d_s = a->d_s;
  float* w = reinterpret_cast<float*>(a->workspace);
  p.inv_im = w;
  p.inv_s = w + n_im;
  p.n_pairs = reinterpret_cast<unsigned int*>(w + n_im + n_s);
  p.pairs = reinterpret_cast<Pair*>(w + n_im + n_s + 4);
  p.max_pairs = (unsigned int)a->max_pairs;
  p.eps = a->eps;
  ALAD_CUDA(cudaMemsetAsync(p.n_pairs, 0, sizeof(unsigned int), st));
  inv_norm_kernel<<<(unsigned)((n_im + 7) / 8), 256, 0, st>>>(p.im, p.im_sb, p.im_ss, p.Bi, p.S_im, p.d, p.eps, p.inv_im);
  inv_norm_kernel<<<(unsigned)((n_s + 7) / 8), 256, 0, st>>>(p.s, p.s_sb, p.s_ss, p.Bc, p.S_s, p.d, p.eps, p.inv_s);
  const long long n = (long long)p.Bi * p.Bc;
  compact_kernel<<<(unsigned)std::min<long long>((n + 255) / 256, 4096), 256, 0, st>>>(p);
  pair_kernel<<<sm_count() * 4, BW_THREADS, 0, st>>>(p);
  norm_bwd_kernel<<<(unsigned)((n_im + 7) / 8), 256, 0, st>>>(p.im, p.im_sb, p.im_ss, p.Bi, p.S_im, p.d, p.eps, p.inv_im, p.d_im);
  norm_bwd_kernel<<<(unsigned)((n_s + 7) / 8), 256, 0, st>>>(p.s, p.s_sb, p.s_ss, p.Bc, p.S_s, p.d, p.eps, p.inv_s, p.d_s);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}
