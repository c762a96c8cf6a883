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
#include <stdint.h>
#include <stdlib.h>

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
  float* d_im;      // [Bi, S_im, d] with strides (dim_sb, dim_ss, 1)
  float* d_s;       // [Bc, S_s, d] with strides (ds_sb, ds_ss, 1)
  long long dim_sb, dim_ss, ds_sb, ds_ss;
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

// Generic fallback (any d / container extents): persistent CTAs loop over the pair list; one warp per word
// of the caption, warp-wide dot products straight from global memory.
__global__ void __launch_bounds__(BW_THREADS) pair_generic_kernel(const BwdParams p) {
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
      float* gi = p.d_im + (long long)pr.i * p.dim_sb + (long long)(1 + rbest) * p.dim_ss;
      float* gs = p.d_s + (long long)pr.j * p.ds_sb + (long long)(1 + w) * p.ds_ss;
      for (int e = lane; e < d; e += 32) {
        atomicAdd(gi + e, pr.g * (__ldg(xw + e) * iw));
        atomicAdd(gs + e, pr.g * (__ldg(xr + e) * ir));
      }
    }
  }
}


// ------------------------------------------------------------------------------------------------
// Tiled pair kernel (the training path: d % 32 == 0, 16-byte aligned rows, <= RA*8 regions and <= WB*32
// words per item).  One CTA per pair and iteration:
//   1. C[r, w] = <x_r, y_w> as an fp32 register-tiled GEMM: warp `wid` owns regions wid + 8a (a < RA),
//      lane l owns words l + 32b (b < WB); 32-float K chunks of both operands are staged in shared memory
//      with cp.async (double buffered).  Region values are read as warp-wide broadcasts, word values as
//      conflict-free LDS.128 (row pitch 36 floats).
//   2. arg-max over the regions of every word (first occurrence wins, masked regions = 0) through shared memory.
//   3. scatter: one warp per word adds the contribution of (word, winning region) to both gradient rows with
//      16-byte vector atomics (red.global.add.v4.f32), F.normalize Jacobian included (no norm_bwd pass).
// ------------------------------------------------------------------------------------------------
constexpr int PT_KC = 32;             // floats per K chunk
constexpr int PT_PITCH = PT_KC + 4;   // shared-memory row pitch (floats): 144 B, keeps 16 B alignment

#ifndef ALAD_CPU_EMU
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned int d = static_cast<unsigned int>(__cvta_generic_to_shared(smem_dst));
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
#else   // tests/cuda_emu (host threads): the copies complete at issue, the vector reduction is four scalar atomics
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) { memcpy(smem_dst, gsrc, 16); }
__device__ __forceinline__ void cp_async_commit() {}
template <int N>
__device__ __forceinline__ void cp_async_wait() {}
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  atomicAdd(addr, a); atomicAdd(addr + 1, b); atomicAdd(addr + 2, c); atomicAdd(addr + 3, d);
}
#endif

template <int RA, int WB>
struct PairTile {
  static constexpr int XR = RA * BW_WARPS;                 // region rows staged per chunk
  static constexpr int YR = WB * 32;                       // word rows staged per chunk
  static constexpr int STAGE_FLOATS = (XR + YR) * PT_PITCH;
  static constexpr int SMEM_BYTES = 2 * STAGE_FLOATS * 4 + BW_WARPS * YR * 8 + YR * 8;
};

template <int RA, int WB>
__global__ void __launch_bounds__(BW_THREADS, 2) pair_tile_kernel(const BwdParams p) {
  using T = PairTile<RA, WB>;
  extern __shared__ __align__(16) float pt_smem[];
  float* stage0 = pt_smem;
  float* bestv_s = pt_smem + 2 * T::STAGE_FLOATS;                        // [8 warps][YR]
  int* besti_s = reinterpret_cast<int*>(bestv_s + BW_WARPS * T::YR);     // [8 warps][YR]
  float* wbest_s = reinterpret_cast<float*>(besti_s + BW_WARPS * T::YR); // [YR] winning cosine
  int* wreg_s = reinterpret_cast<int*>(wbest_s + T::YR);                 // [YR] winning region (-1: none)

  const int tid = threadIdx.x, wid = tid >> 5, lane = tid & 31;
  const unsigned int n_pairs = min(*p.n_pairs, p.max_pairs);
  const int d = p.d, n_chunks = d / PT_KC;
  const int ld_row = tid >> 3, ld_seg = (tid & 7) * 4;                   // 32 rows x 8 x 16 B per pass

  for (unsigned int q = blockIdx.x; q < n_pairs; q += gridDim.x) {
    const Pair pr = p.pairs[q];
    const int nr = min(p.nr[pr.i], T::XR), nw = min(p.nw[pr.j], T::YR);
    const float* im_i = p.im + (long long)pr.i * p.im_sb + p.im_ss;        // slot 1
    const float* s_j = p.s + (long long)pr.j * p.s_sb + p.s_ss;
    const float* inv_i = p.inv_im + (long long)pr.i * p.S_im + 1;
    const float* inv_j = p.inv_s + (long long)pr.j * p.S_s + 1;

    auto load_chunk = [&](int c, float* st) {
      float* Xs = st;
      float* Ys = st + T::XR * PT_PITCH;
      const int k0 = c * PT_KC + ld_seg;
      for (int r = ld_row; r < nr; r += 32) cp_async16(Xs + r * PT_PITCH + ld_seg, im_i + (long long)r * p.im_ss + k0);
      for (int w = ld_row; w < nw; w += 32) cp_async16(Ys + w * PT_PITCH + ld_seg, s_j + (long long)w * p.s_ss + k0);
      cp_async_commit();
    };

    float acc[RA][WB];
#pragma unroll
    for (int a = 0; a < RA; ++a)
#pragma unroll
      for (int b = 0; b < WB; ++b) acc[a][b] = 0.f;

    load_chunk(0, stage0);
    for (int c = 0; c < n_chunks; ++c) {
      float* st = stage0 + (c & 1) * T::STAGE_FLOATS;
      if (c + 1 < n_chunks) {
        load_chunk(c + 1, stage0 + ((c + 1) & 1) * T::STAGE_FLOATS);
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
      __syncthreads();
      const float* Xs = st;
      const float* Ys = st + T::XR * PT_PITCH;
#pragma unroll
      for (int k = 0; k < PT_KC; k += 4) {
        float4 yv[WB];
#pragma unroll
        for (int b = 0; b < WB; ++b) yv[b] = *reinterpret_cast<const float4*>(Ys + (lane + 32 * b) * PT_PITCH + k);
#pragma unroll
        for (int a = 0; a < RA; ++a) {
          if (wid + BW_WARPS * a < nr) {                                  // warp-uniform
            const float4 xv = *reinterpret_cast<const float4*>(Xs + (wid + BW_WARPS * a) * PT_PITCH + k);
#pragma unroll
            for (int b = 0; b < WB; ++b) {
              acc[a][b] = fmaf(xv.x, yv[b].x, acc[a][b]);
              acc[a][b] = fmaf(xv.y, yv[b].y, acc[a][b]);
              acc[a][b] = fmaf(xv.z, yv[b].z, acc[a][b]);
              acc[a][b] = fmaf(xv.w, yv[b].w, acc[a][b]);
            }
          }
        }
      }
      __syncthreads();            // the other stage is refilled by the next iteration's cp.async
    }

    // ---- arg-max over this warp's regions (ascending index, strict >: first occurrence wins)
#pragma unroll
    for (int b = 0; b < WB; ++b) {
      const int w = lane + 32 * b;
      float best = -INFINITY;
      int rbest = -1;
      if (w < nw) {
        const float iw = inv_j[w];
#pragma unroll
        for (int a = 0; a < RA; ++a) {
          const int r = wid + BW_WARPS * a;
          if (r < nr) {
            const float v = acc[a][b] * inv_i[r] * iw;
            if (v > best) {
              best = v;
              rbest = r;
            }
          }
        }
      }
      bestv_s[wid * T::YR + w] = best;
      besti_s[wid * T::YR + w] = rbest;
    }
    __syncthreads();
    const bool clamp = nr < p.R;
    for (int w = tid; w < nw; w += BW_THREADS) {
      float best = -INFINITY;
      int rbest = -1;
#pragma unroll
      for (int g = 0; g < BW_WARPS; ++g) {
        const float v = bestv_s[g * T::YR + w];
        const int r = besti_s[g * T::YR + w];
        if (r >= 0 && (v > best || (v == best && r < rbest))) {
          best = v;
          rbest = r;
        }
      }
      // masked regions take part in the max with value 0 and sit after the valid ones
      if (clamp && best < 0.f) rbest = -1;
      wbest_s[w] = best;
      wreg_s[w] = rbest;
    }
    __syncthreads();

    // ---- scatter (one warp per word) with the F.normalize Jacobian applied per contribution: the gradient
    // w.r.t. the unit vectors is g * yhat_w (region side) and g * xhat_r (word side), and J(x) v =
    // (v - xhat <xhat, v>) / ||x|| is linear, with <xhat_r, yhat_w> = c already known -- no second pass over
    // the gradient tensors.  Below eps the norm is the constant eps (no projection term).
    const float inv_clamped = 1.f / p.eps;        // +inf for eps = 0: never clamped
    for (int w = wid; w < nw; w += BW_WARPS) {
      const int rbest = wreg_s[w];
      if (rbest < 0) continue;
      const float* xw = s_j + (long long)w * p.s_ss;
      const float* xr = im_i + (long long)rbest * p.im_ss;
      const float iw = inv_j[w], ir = inv_i[rbest], c = wbest_s[w];
      const float a_y = pr.g * ir * iw;                                   // coefficient of y_w in d x_r
      const float a_x = (ir >= inv_clamped) ? 0.f : -pr.g * c * ir * ir;  // coefficient of x_r in d x_r
      const float b_y = (iw >= inv_clamped) ? 0.f : -pr.g * c * iw * iw;  // coefficient of y_w in d y_w
      float* gi = p.d_im + (long long)pr.i * p.dim_sb + (long long)(1 + rbest) * p.dim_ss;
      float* gs = p.d_s + (long long)pr.j * p.ds_sb + (long long)(1 + w) * p.ds_ss;
      for (int e = lane * 4; e < d; e += 128) {
        const float4 yw = __ldg(reinterpret_cast<const float4*>(xw + e));
        const float4 xx = __ldg(reinterpret_cast<const float4*>(xr + e));
        red_add_v4(gi + e, fmaf(a_y, yw.x, a_x * xx.x), fmaf(a_y, yw.y, a_x * xx.y), fmaf(a_y, yw.z, a_x * xx.z),
                   fmaf(a_y, yw.w, a_x * xx.w));
        red_add_v4(gs + e, fmaf(a_y, xx.x, b_y * yw.x), fmaf(a_y, xx.y, b_y * yw.y), fmaf(a_y, xx.z, b_y * yw.z),
                   fmaf(a_y, xx.w, b_y * yw.w));
      }
    }
    __syncthreads();              // wreg_s / staging buffers are reused by the next pair
  }
}

template <int RA, int WB>
static cudaError_t launch_pair_tile(const BwdParams& p, cudaStream_t st) {
  using T = PairTile<RA, WB>;
  static int per_sm = 0;        // resident CTAs per SM (same for every sm_100a device)
  if (per_sm == 0) {
    cudaError_t e = cudaFuncSetAttribute(pair_tile_kernel<RA, WB>, cudaFuncAttributeMaxDynamicSharedMemorySize, T::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    int n = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, pair_tile_kernel<RA, WB>, BW_THREADS, T::SMEM_BYTES);
    if (e != cudaSuccess) return e;
    per_sm = n > 0 ? n : 1;
  }
  pair_tile_kernel<RA, WB><<<sm_count() * per_sm, BW_THREADS, T::SMEM_BYTES, st>>>(p);
  return cudaGetLastError();
}

// in place: dx = (dxhat - xhat * <xhat, dxhat>) / max(||x||, eps); below eps the norm is the constant eps
__global__ void norm_bwd_kernel(const float* __restrict__ x, long long sb, long long ss, int B, int S, int d, float eps,
                                const float* __restrict__ inv, float* __restrict__ dx, long long dsb, long long dss) {
  const int tok = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (tok >= B * S) return;
  const float* row = x + (long long)(tok / S) * sb + (long long)(tok % S) * ss;
  float* g = dx + (long long)(tok / S) * dsb + (long long)(tok % S) * dss;
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
  // gradient layouts: contiguous [B, S, d] unless strides are given (any dense permutation of the item / slot
  // dimensions, e.g. the [S, B, d] layout ALADModel.forward_loss permutes from, alad_model.py:377-378)
  p.dim_sb = a->d_im_stride_b > 0 ? a->d_im_stride_b : (long long)a->S_im * a->d;
  p.dim_ss = a->d_im_stride_s > 0 ? a->d_im_stride_s : a->d;
  p.ds_sb = a->d_s_stride_b > 0 ? a->d_s_stride_b : (long long)a->S_s * a->d;
  p.ds_ss = a->d_s_stride_s > 0 ? a->d_s_stride_s : a->d;
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
  {
    auto al16 = [](const void* q) { return (reinterpret_cast<uintptr_t>(q) & 15) == 0; };
    const bool vec_ok = p.d % PT_KC == 0 && al16(p.im) && al16(p.s) && al16(p.d_im) && al16(p.d_s) &&
                        ((p.im_sb | p.im_ss | p.s_sb | p.s_ss | p.dim_sb | p.dim_ss | p.ds_sb | p.ds_ss) & 3) == 0;
    const int r_ext = p.S_im - 1, w_ext = p.S_s - 1;      // valid counts never exceed the slots after slot 0
    const bool force_generic = getenv("ALAD_BWD_GENERIC") != nullptr;   // A/B switch (tests, experiments)
    bool fused_jacobian = true;
    if (vec_ok && !force_generic && r_ext <= 40 && w_ext <= 64) {
      ALAD_CUDA((launch_pair_tile<5, 2>(p, st)));
    } else if (vec_ok && !force_generic && r_ext <= 72 && w_ext <= 96) {
      ALAD_CUDA((launch_pair_tile<9, 3>(p, st)));
    } else if (vec_ok && !force_generic && r_ext <= 96 && w_ext <= 64) {   // roles swapped ('MwSr'): words on the max side
      ALAD_CUDA((launch_pair_tile<12, 2>(p, st)));
    } else {
      pair_generic_kernel<<<sm_count() * 4, BW_THREADS, 0, st>>>(p);
      fused_jacobian = false;
    }
    if (fused_jacobian) {
      ALAD_CUDA(cudaGetLastError());
      return ALAD_OK;
    }
  }
  norm_bwd_kernel<<<(unsigned)((n_im + 7) / 8), 256, 0, st>>>(p.im, p.im_sb, p.im_ss, p.Bi, p.S_im, p.d, p.eps, p.inv_im, p.d_im,
                                                             p.dim_sb, p.dim_ss);
  norm_bwd_kernel<<<(unsigned)((n_s + 7) / 8), 256, 0, st>>>(p.s, p.s_sb, p.s_ss, p.Bc, p.S_s, p.d, p.eps, p.inv_s, p.d_s,
                                                            p.ds_sb, p.ds_ss);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}
