// alad_pack_tokens: L2-normalise (F.normalize semantics, alad/loss.py:80-81), drop the
// unscored slots (loss.py:87-90), compact the valid tokens of every item into dense
// K-major bf16 rows (the layout the TMA descriptors of the scoring kernel read) and
// optionally split fp32 into bf16 hi/lo parts for the fp32-grade mode.
// HBM-bound: reads 4 B, writes 2 B (6 B split) per element; one warp per token row,
// 128-bit loads, 64-bit stores.
#include <cuda_bf16.h>

#include "common.h"

namespace alad {

constexpr int PACK_WARPS = 8;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ uint2 pack_bf16x4(float a, float b, float c, float d) {
  __nv_bfloat162 lo = __floats2bfloat162_rn(a, b);
  __nv_bfloat162 hi = __floats2bfloat162_rn(c, d);
  uint2 r;
  r.x = *reinterpret_cast<uint32_t*>(&lo);
  r.y = *reinterpret_cast<uint32_t*>(&hi);
  return r;
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// fp32 -> TF32 (10-bit mantissa), round to nearest, ties away from zero
__device__ __forceinline__ float round_tf32(float v) {
#ifndef ALAD_CPU_EMU
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
#else
  uint32_t u = __float_as_uint(v);
  u = (u + 0x1000u) & 0xffffe000u;
  return __uint_as_float(u);
#endif
}

template <bool kVec>
__global__ void __launch_bounds__(PACK_WARPS * 32) pack_tokens_kernel(const alad_pack_args a) {
  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  // containers of one slot per item (the plain-GEMM operands of dot_sim / compute_recall / the backward GEMMs): one
  // item per WARP, so that all eight warps of the CTA work; otherwise one item per CTA, its tokens over the warps
  const bool flat = a.S == 1;
  const int b = flat ? blockIdx.x * PACK_WARPS + warp : blockIdx.x;
  if (b >= a.B) return;
  const int cnt = a.count[b];
  const long long row0 = a.row_off[b];
  const int d = a.d;
  const int Kp = a.Kp;
  __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(a.dst);
  // offsets of the three output segments inside a packed row
  const int off_hi2 = (a.mode == 1) ? d : 2 * d;   // second copy of hi
  const int off_lo = (a.mode == 1) ? 2 * d : d;    // lo part
  const int used = (a.mode == 0) ? d : 3 * d;

  for (int t = flat ? 0 : warp; t < cnt; t += flat ? 1 : PACK_WARPS) {
    const float* x = a.src + (long long)b * a.stride_b + (long long)(a.slot0 + t) * a.stride_s;
    __nv_bfloat16* y = dst + (row0 + t) * (long long)Kp;
    if (a.mode == 3) {
      // TF32 operands: the normalised fp32 values rounded to TF32 (the tensor core reads the upper 19 bits); a row is
      // Kp / 2 floats = Kp two-byte units, so the tile geometry in BYTES is that of the bf16 rows
      float* yf = reinterpret_cast<float*>(y);
      float ssq = 0.f;
      if (a.normalize) {
        for (int i = lane; i < d; i += 32) {
          const float v = __ldg(x + i);
          ssq += v * v;
        }
        ssq = warp_sum(ssq);
      }
      const float den = a.normalize ? fmaxf(sqrtf(ssq), a.eps) : 1.f;
      // rounded to TF32 HERE (round to nearest): the tensor core would truncate the low 13 mantissa bits, a systematic
      // -5e-4 per operand that adds up over the d products instead of averaging out
      for (int i = lane; i < d; i += 32) {
        const float v = __ldg(x + i) / den;
        yf[i] = round_tf32(v);
      }
      for (int i = d + lane; i < Kp / 2; i += 32) yf[i] = 0.f;
      if (a.row_item != nullptr && lane == 0) a.row_item[row0 + t] = a.item_base + b;
      continue;
    }
    float ss = 0.f;
    if (a.normalize) {
      if (kVec) {
        const float4* x4 = reinterpret_cast<const float4*>(x);
        for (int i = lane; i < d / 4; i += 32) {
          const float4 v = __ldg(x4 + i);
          ss += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
        }
      } else {
        for (int i = lane; i < d; i += 32) {
          const float v = __ldg(x + i);
          ss += v * v;
        }
      }
      ss = warp_sum(ss);
    }
    const float denom = a.normalize ? fmaxf(sqrtf(ss), a.eps) : 1.f;
    if (kVec) {
      const float4* x4 = reinterpret_cast<const float4*>(x);
      for (int i = lane; i < d / 4; i += 32) {
        float4 v = __ldg(x4 + i);
        v.x /= denom; v.y /= denom; v.z /= denom; v.w /= denom;
        const uint2 hi = pack_bf16x4(v.x, v.y, v.z, v.w);
        *reinterpret_cast<uint2*>(y + 4 * i) = hi;
        if (a.mode != 0) {
          *reinterpret_cast<uint2*>(y + off_hi2 + 4 * i) = hi;
          const uint2 lo = pack_bf16x4(v.x - bf16_round(v.x), v.y - bf16_round(v.y), v.z - bf16_round(v.z),
                                       v.w - bf16_round(v.w));
          *reinterpret_cast<uint2*>(y + off_lo + 4 * i) = lo;
        }
      }
    } else {
      for (int i = lane; i < d; i += 32) {
        const float v = __ldg(x + i) / denom;
        const __nv_bfloat16 hi = __float2bfloat16_rn(v);
        y[i] = hi;
        if (a.mode != 0) {
          y[off_hi2 + i] = hi;
          y[off_lo + i] = __float2bfloat16_rn(v - __bfloat162float(hi));
        }
      }
    }
    for (int i = used + lane; i < Kp; i += 32) y[i] = __float2bfloat16_rn(0.f);   // K padding
    if (a.row_item != nullptr && lane == 0) a.row_item[row0 + t] = a.item_base + b;
  }
}

// ---------------------------------------------------------------------------------------------
// pooled tokens: out[b, :] = sum of the normalised valid tokens of item b (fp32)
constexpr int POOL_MAX_TOKENS = 256;

__global__ void __launch_bounds__(PACK_WARPS * 32)
pool_tokens_kernel(const float* __restrict__ src, long long sb, long long ss, int S, int d, int slot0,
                   const int* __restrict__ count, float eps, float* __restrict__ out) {
  __shared__ float inv[POOL_MAX_TOKENS];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int cnt = min(count[b], POOL_MAX_TOKENS);
  const float* base = src + (long long)b * sb + (long long)slot0 * ss;
  for (int t = warp; t < cnt; t += PACK_WARPS) {
    const float* x = base + (long long)t * ss;
    float acc = 0.f;
    for (int e = lane; e < d; e += 32) {
      const float v = __ldg(x + e);
      acc += v * v;
    }
    acc = warp_sum(acc);
    if (lane == 0) inv[t] = fmaxf(sqrtf(acc), eps);
  }
  __syncthreads();
  for (int e = threadIdx.x; e < d; e += PACK_WARPS * 32) {
    float acc = 0.f;
    for (int t = 0; t < cnt; ++t) acc += __ldg(base + (long long)t * ss + e) / inv[t];
    out[(long long)b * d + e] = acc;
  }
}

__global__ void scale_scores_kernel(float* __restrict__ S, long long ld, int Ni, int Nc, const float* __restrict__ col_div,
                                    float mul) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int i = blockIdx.y;
  if (j >= Nc) return;
  float v = S[(long long)i * ld + j] * mul;
  if (col_div) v = v / col_div[j];
  S[(long long)i * ld + j] = v;
}

}  // namespace alad

extern "C" int alad_pool_tokens(const float* src, int64_t stride_b, int64_t stride_s, int32_t B, int32_t S, int32_t d,
                                int32_t slot0, const int32_t* count, float eps, float* out, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(B >= 0 && S >= 0 && d > 0 && slot0 >= 0, "alad_pool_tokens: bad shape");
  ALAD_REQUIRE(S - slot0 <= POOL_MAX_TOKENS, "alad_pool_tokens: at most %d scored slots per item", POOL_MAX_TOKENS);
  if (B == 0) return ALAD_OK;
  ALAD_REQUIRE(src && count && out, "alad_pool_tokens: NULL pointer");
  pool_tokens_kernel<<<B, PACK_WARPS * 32, 0, as_stream(stream)>>>(src, stride_b, stride_s, S, d, slot0, count, eps, out);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_scale_scores(float* S, int64_t ldS, int32_t Ni, int32_t Nc, const float* col_div, float mul,
                                 void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Ni >= 0 && Nc >= 0 && ldS >= Nc && Ni <= 65535, "alad_scale_scores: bad shape");
  if (Ni == 0 || Nc == 0) return ALAD_OK;
  ALAD_REQUIRE(S, "alad_scale_scores: NULL pointer");
  dim3 grid((Nc + 255) / 256, Ni);
  scale_scores_kernel<<<grid, 256, 0, as_stream(stream)>>>(S, ldS, Ni, Nc, col_div, mul);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_pack_tokens(const alad_pack_args* a, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(a != nullptr, "alad_pack_tokens: NULL args");
  ALAD_REQUIRE(a->B >= 0 && a->S >= 0 && a->d > 0, "alad_pack_tokens: bad shape");
  ALAD_REQUIRE(a->mode >= 0 && a->mode <= 3, "alad_pack_tokens: unknown mode %d", a->mode);
  ALAD_REQUIRE(a->Kp % ALAD_TILE_K == 0 && a->Kp >= (a->mode == 0 ? a->d : a->mode == 3 ? 2 * a->d : 3 * a->d),
               "alad_pack_tokens: Kp=%d too small or not a multiple of %d", a->Kp, ALAD_TILE_K);
  if (a->B == 0) return ALAD_OK;
  ALAD_REQUIRE(a->src && a->dst && a->count && a->row_off, "alad_pack_tokens: NULL pointer");
  const bool vec = (a->d % 4 == 0) && (a->stride_b % 4 == 0) && (a->stride_s % 4 == 0) &&
                   ((reinterpret_cast<uintptr_t>(a->src) & 15) == 0) && ((reinterpret_cast<uintptr_t>(a->dst) & 7) == 0);
  cudaStream_t st = as_stream(stream);
  const unsigned grid = a->S == 1 ? (unsigned)((a->B + PACK_WARPS - 1) / PACK_WARPS) : (unsigned)a->B;
  if (vec)
    pack_tokens_kernel<true><<<grid, PACK_WARPS * 32, 0, st>>>(*a);
  else
    pack_tokens_kernel<false><<<grid, PACK_WARPS * 32, 0, st>>>(*a);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}
