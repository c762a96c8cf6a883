// The remaining DistillationLoss modes of alad/loss.py:371-425, forward + gradient in one call:
//   alad_distill_mse_fwd_bwd          -- mode 'mse'         (loss.py:371-373, owns the wb parameter)
//   alad_distill_contrastive_fwd_bwd  -- mode 'contrastive' (loss.py:397-418)
//   alad_distill_ordinal_fwd_bwd      -- mode 'ordinal'     (loss.py:374-396)
// All three are B x B sweeps (HBM / launch-latency bound): coalesced row sweeps, warp-shuffle
// reductions, per-CTA partial results reduced in a fixed order by the last CTA to finish, so the
// loss is bit-reproducible.  Gradient entries are small integers times one scale: no float atomics.
#include <math.h>

#include "common.h"

namespace alad {

constexpr int DT = 256;   // threads per CTA

__device__ __forceinline__ float d_warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int d_warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
template <typename T>
__device__ __forceinline__ T d_block_sum(T v, T* scratch) {   // DT threads, result valid in thread 0
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  v = d_warp_sum(v);
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  T tot = 0;
  if (threadIdx.x == 0)
    for (int w = 0; w < DT / 32; ++w) tot += scratch[w];
  __syncthreads();
  return tot;
}
__device__ __forceinline__ bool d_last_cta(unsigned int* counter) {
  __shared__ bool is_last;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) is_last = (atomicAdd(counter, 1u) == gridDim.x - 1);
  __syncthreads();
  if (is_last) __threadfence();
  return is_last;
}

// ------------------------------------------------------------------------------------ mse
// loss = mean((M*w0 + w1 - T)^2);  dM = 2*w0/n * diff;  dwb = (2/n) * (sum diff*M, sum diff)
struct MseParams {
  const float* T;
  long long ldT;
  const float* M;
  long long ldM;
  int B;
  const float* wb;       // device [2]
  float* loss;
  float* dM;             // optional [B, ldG]
  long long ldG;
  float* dwb;            // optional device [2]
  float* part;           // [grid][3]
  unsigned int* counter;
};

__global__ void __launch_bounds__(DT) distill_mse_kernel(const MseParams p) {
  __shared__ float sf[DT / 32];
  const int B = p.B;
  const float w0 = p.wb[0], w1 = p.wb[1];
  const float inv_n = 1.f / (static_cast<float>(B) * static_cast<float>(B));
  float s2 = 0.f, sm = 0.f, s1 = 0.f;
  for (int i = blockIdx.x; i < B; i += gridDim.x) {          // one row at a time: coalesced
    const float* rt = p.T + (long long)i * p.ldT;
    const float* rm = p.M + (long long)i * p.ldM;
    for (int j = threadIdx.x; j < B; j += DT) {
      const float m = __ldg(rm + j);
      const float diff = fmaf(m, w0, w1) - __ldg(rt + j);
      s2 = fmaf(diff, diff, s2);
      sm = fmaf(diff, m, sm);
      s1 += diff;
      if (p.dM) p.dM[(long long)i * p.ldG + j] = 2.f * w0 * inv_n * diff;
    }
  }
  const float a = d_block_sum<float>(s2, sf), b = d_block_sum<float>(sm, sf), c = d_block_sum<float>(s1, sf);
  if (threadIdx.x == 0) {
    float* o = p.part + 3 * (long long)blockIdx.x;
    o[0] = a; o[1] = b; o[2] = c;
  }
  if (!d_last_cta(p.counter)) return;
  if (threadIdx.x == 0) {                                    // fixed order over the CTAs
    float t2 = 0.f, tm = 0.f, t1 = 0.f;
    for (unsigned int g = 0; g < gridDim.x; ++g) {
      t2 += p.part[3 * g]; tm += p.part[3 * g + 1]; t1 += p.part[3 * g + 2];
    }
    *p.loss = t2 * inv_n;
    if (p.dwb) {
      p.dwb[0] = 2.f * inv_n * tm;
      p.dwb[1] = 2.f * inv_n * t1;
    }
    *p.counter = 0;
  }
}

// ------------------------------------------------------------------------------------ contrastive
// ns[k] = argmax_j Tnd[k, j], ni[k] = argmax_i Tnd[i, k] (Tnd = teacher with a zero diagonal, first
// occurrence); the reference index_selects whole columns / rows of the hinge matrices, so with
// cs[j] = #{k : ns[k] = j}, ci[i] = #{k : ni[k] = i}:
//   loss = sum_ij cs[j] * [m + M_ij - M_ii]_+  +  sum_ij ci[i] * [m + M_ij - M_jj]_+
struct ContrParams {
  float* T;              // teacher; its diagonal is zeroed in place when zero_diag (loss.py:400)
  long long ldT;
  const float* M;
  long long ldM;
  int B;
  float margin;
  int zero_diag;
  float* loss;
  float* dM;
  long long ldG;
  int* cs;               // [B] multiplicity of column j among the row-wise hard negatives
  int* ci;               // [B] multiplicity of row i among the column-wise hard negatives
  int* rowcorr;          // [B] sum_j cs[j] * active_s(i, j)
  int* colcorr;          // [B] sum_i ci[i] * active_im(i, j)
  float* part;           // [B] per-row loss partials
  float* diag;           // [B] diagonal of M
  unsigned int* counter;
};

__device__ __forceinline__ void amax_combine(float& v, int& i, float ov, int oi) {
  if (ov > v || (ov == v && oi < i)) {
    v = ov;
    i = oi;
  }
}

// grid = B row CTAs + ceil(B/32) column-strip CTAs
__global__ void __launch_bounds__(DT) distill_contr_select_kernel(const ContrParams p) {
  __shared__ float sv[DT / 32];
  __shared__ int si[DT / 32];
  __shared__ float cv[DT / 32][32];
  __shared__ int cx[DT / 32][32];
  const int B = p.B;
  if ((int)blockIdx.x < B) {
    const int k = blockIdx.x;
    const float* row = p.T + (long long)k * p.ldT;
    float best = -INFINITY;
    int arg = B;
    for (int j = threadIdx.x; j < B; j += DT) amax_combine(best, arg, j == k ? 0.f : row[j], j);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1)
      amax_combine(best, arg, __shfl_xor_sync(0xffffffffu, best, o), __shfl_xor_sync(0xffffffffu, arg, o));
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
      sv[warp] = best;
      si[warp] = arg;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int w = 1; w < DT / 32; ++w) amax_combine(best, arg, sv[w], si[w]);
      atomicAdd(p.cs + arg, 1);
      p.diag[k] = p.M[(long long)k * p.ldM + k];
    }
  } else {
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    const int k = (blockIdx.x - B) * 32 + tx;
    float best = -INFINITY;
    int arg = B;
    if (k < B)
      for (int i = ty; i < B; i += DT / 32) amax_combine(best, arg, i == k ? 0.f : p.T[(long long)i * p.ldT + k], i);
    cv[ty][tx] = best;
    cx[ty][tx] = arg;
    __syncthreads();
    if (ty == 0 && k < B) {
      for (int y = 1; y < DT / 32; ++y) amax_combine(best, arg, cv[y][tx], cx[y][tx]);
      atomicAdd(p.ci + arg, 1);
    }
  }
}

__global__ void __launch_bounds__(DT) distill_contr_loss_kernel(const ContrParams p) {
  __shared__ float sf[DT / 32];
  __shared__ int sn[DT / 32];
  const int B = p.B;
  const int i = blockIdx.x;
  const float* rm = p.M + (long long)i * p.ldM;
  const float dii = p.diag[i];
  const int ci_i = p.ci[i];
  float acc = 0.f;
  int rc = 0;
  for (int j = threadIdx.x; j < B; j += DT) {
    const float m = __ldg(rm + j);
    const float hs = p.margin + m - dii;                 // caption retrieval hinge (diagonal NOT cleared)
    const float hi = p.margin + m - __ldg(p.diag + j);   // image retrieval hinge
    const int cs_j = p.cs[j];
    const int as = hs > 0.f ? cs_j : 0, ai = hi > 0.f ? ci_i : 0;
    acc += static_cast<float>(cs_j) * fmaxf(hs, 0.f) + static_cast<float>(ci_i) * fmaxf(hi, 0.f);
    rc += as;
    if (ai) atomicAdd(p.colcorr + j, ai);                // integers: exact in any order
    if (p.dM) p.dM[(long long)i * p.ldG + j] = static_cast<float>(as + ai);
  }
  const float tot = d_block_sum<float>(acc, sf);
  const int rtot = d_block_sum<int>(rc, sn);
  if (threadIdx.x == 0) {
    p.part[i] = tot;
    p.rowcorr[i] = rtot;
  }
  if (p.zero_diag && threadIdx.x == 0) p.T[(long long)i * p.ldT + i] = 0.f;
  if (!d_last_cta(p.counter)) return;
  float s = 0.f;
  for (int r = threadIdx.x; r < B; r += DT) s += p.part[r];
  const float l = d_block_sum<float>(s, sf);
  if (threadIdx.x == 0) {
    *p.loss = l;
    *p.counter = 0;
  }
  if (p.dM)
    for (int r = threadIdx.x; r < B; r += DT)
      p.dM[(long long)r * p.ldG + r] -= static_cast<float>(p.rowcorr[r] + __ldcg(p.colcorr + r));
}

// ------------------------------------------------------------------------------------ ordinal
// One CTA per row (blockIdx < B) or column (blockIdx >= B): stable ascending sort of the teacher
// values (bitonic on (key, index) pairs in shared memory), then for sorted positions q:
//   valid(q) = Ts[q + stride] >= threshold,  h(q) = margin + Ms[q] - Ms[q + stride]
//   direction loss = sum_valid relu(h) / #valid;  d/dM: +1/#valid at q, -1/#valid at q + stride.
struct OrdParams {
  const float* T;
  long long ldT;
  const float* M;
  long long ldM;
  int B, P;              // P = next power of two >= B
  float margin, threshold;
  int stride;
  float* loss;
  float* dM;             // optional [B, ldG]; holds the row-direction integers until the scale pass
  long long ldG;
  float* colg;           // [B, B] column-direction integers
  float* psum;           // [2B] per-line hinge sums
  int* pcnt;             // [2B] per-line valid counts
  float* scale;          // [2] 1/#valid per direction (0 when empty)
  unsigned int* counter;
};

__global__ void __launch_bounds__(DT) distill_ordinal_kernel(const OrdParams p) {
  extern __shared__ unsigned char ord_smem[];
  float* key = reinterpret_cast<float*>(ord_smem);            // [P]
  int* idx = reinterpret_cast<int*>(key + p.P);               // [P]
  float* ms = reinterpret_cast<float*>(idx + p.P);            // [P] student in sorted order
  __shared__ float sf[DT / 32];
  __shared__ int sn[DT / 32];
  const int B = p.B, P = p.P;
  const bool is_row = (int)blockIdx.x < B;
  const int line = is_row ? blockIdx.x : blockIdx.x - B;
  const long long t_step = is_row ? 1 : p.ldT, m_step = is_row ? 1 : p.ldM;
  const float* tl = p.T + (is_row ? (long long)line * p.ldT : line);
  const float* ml = p.M + (is_row ? (long long)line * p.ldM : line);
  for (int e = threadIdx.x; e < P; e += DT) {
    key[e] = e < B ? tl[(long long)e * t_step] : INFINITY;
    idx[e] = e;                                               // padding sorts last (index >= B on ties)
  }
  __syncthreads();
  for (int k = 2; k <= P; k <<= 1) {
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int e = threadIdx.x; e < P; e += DT) {
        const int x = e ^ j;
        if (x > e) {
          const bool up = (e & k) == 0;
          const float ka = key[e], kb = key[x];
          const int ia = idx[e], ib = idx[x];
          const bool a_after_b = ka > kb || (ka == kb && ia > ib);
          if (a_after_b == up) {
            key[e] = kb; key[x] = ka;
            idx[e] = ib; idx[x] = ia;
          }
        }
      }
      __syncthreads();
    }
  }
  for (int e = threadIdx.x; e < B; e += DT) ms[e] = ml[(long long)idx[e] * m_step];
  __syncthreads();
  const int st = p.stride;
  float acc = 0.f;
  int cnt = 0;
  float* gline = p.dM ? (is_row ? p.dM + (long long)line * p.ldG : p.colg + line) : nullptr;
  const long long g_step = is_row ? 1 : B;
  for (int q = threadIdx.x; q < B; q += DT) {
    int g = 0;
    if (q + st < B && key[q + st] >= p.threshold) {           // q is the earlier element of a pair
      const float h = p.margin + ms[q] - ms[q + st];
      acc += fmaxf(h, 0.f);
      cnt += 1;
      g += h > 0.f;
    }
    if (q >= st && key[q] >= p.threshold) {                   // q is the later element of a pair
      const float h = p.margin + ms[q - st] - ms[q];
      g -= h > 0.f;
    }
    if (gline) gline[(long long)idx[q] * g_step] = static_cast<float>(g);
  }
  const float tot = d_block_sum<float>(acc, sf);
  const int ctot = d_block_sum<int>(cnt, sn);
  if (threadIdx.x == 0) {
    p.psum[blockIdx.x] = tot;
    p.pcnt[blockIdx.x] = ctot;
  }
  if (!d_last_cta(p.counter)) return;
  float total = 0.f;
  for (int dir = 0; dir < 2; ++dir) {
    float s = 0.f;
    int c = 0;
    for (int r = threadIdx.x; r < B; r += DT) {
      s += p.psum[dir * B + r];
      c += p.pcnt[dir * B + r];
    }
    const float ss = d_block_sum<float>(s, sf);
    const int cc = d_block_sum<int>(c, sn);
    if (threadIdx.x == 0) {
      total += ss / static_cast<float>(cc);                   // 0/0 = NaN, like torch's mean of an empty tensor
      p.scale[dir] = cc > 0 ? 1.f / static_cast<float>(cc) : 0.f;
    }
  }
  if (threadIdx.x == 0) {
    *p.loss = total;
    *p.counter = 0;
  }
}

__global__ void __launch_bounds__(DT) distill_ordinal_scale_kernel(const OrdParams p) {
  const int i = blockIdx.y, j = blockIdx.x * DT + threadIdx.x;
  if (j >= p.B) return;
  float* g = p.dM + (long long)i * p.ldG + j;
  *g = *g * p.scale[0] + p.colg[(long long)i * p.B + j] * p.scale[1];
}

}  // namespace alad

extern "C" int64_t alad_distill_workspace_bytes(int32_t B, int32_t mode) {
  const int64_t b = B > 0 ? B : 1;
  switch (mode) {
    case 0: return 3 * 4 * (b < 1184 ? b : 1184) + 256;               // mse: per-CTA partials
    case 1: return 6 * 4 * b + 256;                                    // contrastive
    case 2: return 4 * b * b + 4 * 4 * b + 256;                        // ordinal: column integers + partials
    default: return -1;
  }
}

extern "C" int alad_distill_mse_fwd_bwd(const float* teacher, int64_t ldT, const float* student, int64_t ldM, int32_t B,
                                        const float* wb, float* loss, float* dM, int64_t ldG, float* dwb,
                                        void* workspace, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(B > 0 && ldT >= B && ldM >= B && (dM == nullptr || ldG >= B), "alad_distill_mse_fwd_bwd: bad shape");
  ALAD_REQUIRE(teacher && student && wb && loss && workspace, "alad_distill_mse_fwd_bwd: NULL pointer");
  cudaStream_t st = as_stream(stream);
  MseParams p;
  p.T = teacher; p.ldT = ldT; p.M = student; p.ldM = ldM; p.B = B; p.wb = wb; p.loss = loss; p.dM = dM; p.ldG = ldG;
  p.dwb = dwb;
  int grid = sm_count() * 8;
  if (grid > B) grid = B;
  if (grid > 1184) grid = 1184;
  p.part = reinterpret_cast<float*>(workspace);
  p.counter = reinterpret_cast<unsigned int*>(p.part + 3 * (size_t)grid);
  ALAD_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), st));
  distill_mse_kernel<<<grid, DT, 0, st>>>(p);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_distill_contrastive_fwd_bwd(float* teacher, int64_t ldT, const float* student, int64_t ldM,
                                                int32_t B, float margin, int32_t zero_teacher_diag, float* loss,
                                                float* dM, int64_t ldG, void* workspace, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(B > 0 && ldT >= B && ldM >= B && (dM == nullptr || ldG >= B), "alad_distill_contrastive_fwd_bwd: bad shape");
  ALAD_REQUIRE(teacher && student && loss && workspace, "alad_distill_contrastive_fwd_bwd: NULL pointer");
  cudaStream_t st = as_stream(stream);
  ContrParams p;
  p.T = teacher; p.ldT = ldT; p.M = student; p.ldM = ldM; p.B = B; p.margin = margin; p.zero_diag = zero_teacher_diag;
  p.loss = loss; p.dM = dM; p.ldG = ldG;
  int* w = reinterpret_cast<int*>(workspace);
  p.cs = w; p.ci = w + B; p.rowcorr = w + 2 * (size_t)B; p.colcorr = w + 3 * (size_t)B;
  p.part = reinterpret_cast<float*>(w + 4 * (size_t)B);
  p.diag = reinterpret_cast<float*>(w + 5 * (size_t)B);
  p.counter = reinterpret_cast<unsigned int*>(w + 6 * (size_t)B);
  ALAD_CUDA(cudaMemsetAsync(workspace, 0, 4 * sizeof(int) * (size_t)B, st));
  ALAD_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), st));
  distill_contr_select_kernel<<<B + (B + 31) / 32, DT, 0, st>>>(p);
  distill_contr_loss_kernel<<<B, DT, 0, st>>>(p);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_distill_ordinal_fwd_bwd(const float* teacher, int64_t ldT, const float* student, int64_t ldM,
                                            int32_t B, float margin, float threshold, int32_t stride, float* loss,
                                            float* dM, int64_t ldG, void* workspace, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(B > 0 && ldT >= B && ldM >= B && (dM == nullptr || ldG >= B), "alad_distill_ordinal_fwd_bwd: bad shape");
  ALAD_REQUIRE(stride > 0, "alad_distill_ordinal_fwd_bwd: stride must be positive");
  ALAD_REQUIRE(teacher && student && loss && workspace, "alad_distill_ordinal_fwd_bwd: NULL pointer");
  int P = 1;
  while (P < B) P <<= 1;
  if (P > 16384)
    return fail(ALAD_ERR_UNSUPPORTED, "alad_distill_ordinal_fwd_bwd: B=%d exceeds the in-shared-memory sort (16384)", B);
  cudaStream_t st = as_stream(stream);
  OrdParams p;
  p.T = teacher; p.ldT = ldT; p.M = student; p.ldM = ldM; p.B = B; p.P = P; p.margin = margin; p.threshold = threshold;
  p.stride = stride; p.loss = loss; p.dM = dM; p.ldG = ldG;
  float* w = reinterpret_cast<float*>(workspace);
  p.colg = w;
  p.psum = w + (size_t)B * B;
  p.pcnt = reinterpret_cast<int*>(p.psum + 2 * (size_t)B);
  p.scale = reinterpret_cast<float*>(p.pcnt + 2 * (size_t)B);
  p.counter = reinterpret_cast<unsigned int*>(p.scale + 2);
  ALAD_CUDA(cudaMemsetAsync(p.counter, 0, sizeof(unsigned int), st));
  const size_t smem = 12 * (size_t)P;
  ALAD_CUDA(cudaFuncSetAttribute(distill_ordinal_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  distill_ordinal_kernel<<<2 * B, DT, smem, st>>>(p);
  if (dM) {
    dim3 grid((B + DT - 1) / DT, B);
    distill_ordinal_scale_kernel<<<grid, DT, 0, st>>>(p);
  }
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}
