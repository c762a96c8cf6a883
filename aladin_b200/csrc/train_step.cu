// alad_train_losses_fwd / alad_train_losses_bwd: the loss stack of ALADModel.forward_loss
// (alad/alad_model.py:371-428 -- matching_criterion :380, alignment_criterion :386, distillation_loss :405)
// in ONE native call each way.  Nothing new is computed here: the entry points compose alad_scores_fused,
// alad_triplet_fwd_bwd, alad_listnet_fwd_bwd and alad_mrsw_scores_bwd on one stream with one workspace, plus two
// small helper kernels (scaled sum + transpose) for the matching head's backward GEMM operands.  Why it exists: at
// B <= 512 the step launches ~35 kernels of a few microseconds each and is bound by the HOST cost of a dozen
// Python-level calls (0.8 ms enqueue for 0.3-0.6 ms of device work); done natively the enqueue takes ~0.1 ms.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.h"

namespace alad {

static inline int64_t up256(int64_t x) { return (x + 255) / 256 * 256; }

// out[r, c] = sa * A[r, c] + sb * Bm[r, c] and/or outT[c, r] = the same value.  sa / sb are DEVICE scalars (NULL = 1);
// a term whose scale is exactly 0 is dropped (0 * inf would poison the sum); Bm may be NULL.
__global__ void __launch_bounds__(256) axpby_t_kernel(const float* __restrict__ A, long long ldA, const float* __restrict__ sa,
                                                      const float* __restrict__ Bm, long long ldB, const float* __restrict__ sb,
                                                      int rows, int cols, float* __restrict__ out, long long ldo,
                                                      float* __restrict__ outT, long long ldt) {
  __shared__ float tile[32][33];
  const float fa = sa ? __ldg(sa) : 1.f;
  const float fb = (Bm && sb) ? __ldg(sb) : (Bm ? 1.f : 0.f);
  const int c = blockIdx.x * 32 + threadIdx.x;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int r = blockIdx.y * 32 + threadIdx.y + 8 * k;
    float v = 0.f;
    if (r < rows && c < cols) {
      if (A && fa != 0.f) v = fa * __ldg(A + (long long)r * ldA + c);
      if (fb != 0.f) v += fb * __ldg(Bm + (long long)r * ldB + c);
      if (out) out[(long long)r * ldo + c] = v;
    }
    tile[threadIdx.y + 8 * k][threadIdx.x] = v;
  }
  if (!outT) return;
  __syncthreads();
  const int rT = blockIdx.y * 32 + threadIdx.x;      // row of the source = column of outT
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int cT = blockIdx.x * 32 + threadIdx.y + 8 * k;
    if (rT < rows && cT < cols) outT[(long long)cT * ldt + rT] = tile[threadIdx.x][threadIdx.y + 8 * k];
  }
}

static int axpby_t(const float* A, int64_t ldA, const float* sa, const float* Bm, int64_t ldB, const float* sb, int rows,
                   int cols, float* out, int64_t ldo, float* outT, int64_t ldt, cudaStream_t st) {
  if (rows <= 0 || cols <= 0) return ALAD_OK;
  dim3 grid((cols + 31) / 32, (rows + 31) / 32), block(32, 8);
  axpby_t_kernel<<<grid, block, 0, st>>>(A, ldA, sa, Bm, ldB, sb, rows, cols, out, ldo, outT, ldt);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

// Fork / join inside one call: the matching head's chain (two packs, a small GEMM, the hinge kernels: ~50 us of launches that
// leave most SMs idle) runs on a side stream next to the alignment head's packs, and in the backward next to the sparse
// MrSw backward.  One side stream and two events per host thread and device; the pattern is capture-safe (event record / wait).
struct SideLane {
  int dev = -1;
  cudaStream_t side = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
// bit 0 = fork in the forward call, bit 1 = in the backward call.  Measured on one box (tools/train_step_quick.py, B = 512):
// forward 0.366 -> 0.313 ms with bit 0; the backward fork made forward + backward SLOWER or bimodal (0.66 -> 0.72 ms, 1.17 ms
// on another box: the matching head's CTA-pair GEMM competes with the 444 CTAs of the sparse MrSw backward for whole SMs),
// and at B = 128 the step is host-bound, where the four extra stream / event calls only cost.  Hence: forward only, B >= 256.
// A -DALAD_TUNING_ENV build reads ALAD_TRAIN_FORK instead.
static int fork_mask(int B) {
#ifdef ALAD_TUNING_ENV
  static const int m = [] {
    const char* e = getenv("ALAD_TRAIN_FORK");
    return e ? atoi(e) : -1;
  }();
  if (m >= 0) return m;
#endif
  return B >= 256 ? 1 : 0;
}
static int side_lane(SideLane** out) {
  static thread_local SideLane lanes[16];
  int dev = 0;
  ALAD_CUDA(cudaGetDevice(&dev));
  SideLane& l = lanes[dev & 15];
  if (l.dev != dev) {
    ALAD_CUDA(cudaStreamCreateWithFlags(&l.side, cudaStreamNonBlocking));
    ALAD_CUDA(cudaEventCreateWithFlags(&l.fork, cudaEventDisableTiming));
    ALAD_CUDA(cudaEventCreateWithFlags(&l.join, cudaEventDisableTiming));
    l.dev = dev;
  }
  *out = &l;
  return ALAD_OK;
}

struct TrainLayout {
  // forward (the matching chain has its own regions: it runs concurrently with the alignment chain)
  int64_t f_scores, f_scores_bytes, f_loss, f_args, f_scores_m, f_scores_m_bytes, f_loss_m, f_args_m, f_total;
  // backward
  int64_t b_gt, b_gtT, b_imT, b_sT, b_cnt, b_scores, b_scores_bytes, b_mrsw, b_mrsw_bytes, b_total;
};

static TrainLayout train_layout(int32_t B, int32_t S_im, int32_t S_s, int32_t d, int32_t precision, int32_t precision_m) {
  TrainLayout L;
  const int64_t bb = (int64_t)B * B * 4, bd = (int64_t)B * d * 4;
  const int64_t ws_match = alad_scores_fused_workspace_bytes(B, 1, 0, B, 1, 0, d, precision_m);
  const int64_t ws_align = alad_scores_fused_workspace_bytes(B, S_im, 1, B, S_s, 1, d, precision);
  int64_t o = 0;
  L.f_scores = o;  L.f_scores_bytes = ws_match > ws_align ? ws_match : ws_align;  o = up256(o + L.f_scores_bytes);
  L.f_loss = o;    o = up256(o + alad_loss_workspace_bytes(B));
  L.f_args = o;    o = up256(o + 8ll * (B > 0 ? B : 1));            // row_arg | col_arg of the triplet kernels
  L.f_scores_m = o;  L.f_scores_m_bytes = ws_match;  o = up256(o + ws_match);
  L.f_loss_m = o;  o = up256(o + alad_loss_workspace_bytes(B));
  L.f_args_m = o;  o = up256(o + 8ll * (B > 0 ? B : 1));
  L.f_total = o;
  o = 0;
  L.b_gt = o;   o = up256(o + bb);
  L.b_gtT = o;  o = up256(o + bb);
  L.b_imT = o;  o = up256(o + bd);
  L.b_sT = o;   o = up256(o + bd);
  L.b_cnt = o;  o = up256(o + 8ll * (B > 0 ? B : 1));               // nr | nw on the device
  L.b_scores = o;
  L.b_scores_bytes = alad_scores_fused_workspace_bytes(B, 1, 0, d, 1, 0, B > 0 ? B : 1, 1);
  o = up256(o + L.b_scores_bytes);
  L.b_mrsw = o;
  L.b_mrsw_bytes = alad_mrsw_bwd_workspace_bytes(B, S_im, B, S_s, (int64_t)B * B > 0 ? (int64_t)B * B : 1);
  o = up256(o + L.b_mrsw_bytes);
  L.b_total = o;
  return L;
}

static int check_common(const alad_train_losses_args* a, const char* who) {
  ALAD_REQUIRE(a != nullptr, "%s: NULL args", who);
  ALAD_REQUIRE(a->B >= 0 && a->S_im >= 0 && a->S_s >= 0 && a->d > 0, "%s: bad shape", who);
  ALAD_REQUIRE((a->precision == 0 || a->precision == 1) && (a->precision_m == 0 || a->precision_m == 1),
               "%s: unknown precision %d / %d", who, a->precision, a->precision_m);
  ALAD_REQUIRE(a->workspace && (reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, "%s: workspace must be 256-byte aligned", who);
  ALAD_REQUIRE(a->B == 0 || (a->nr && a->nw), "%s: NULL count arrays", who);
  return ALAD_OK;
}

}  // namespace alad

extern "C" int64_t alad_train_losses_workspace_bytes(int32_t B, int32_t S_im, int32_t S_s, int32_t d, int32_t precision,
                                                     int32_t backward) {
  if (B < 0 || S_im < 0 || S_s < 0 || d <= 0) return -1;
  const alad::TrainLayout L = alad::train_layout(B, S_im, S_s, d, precision, 1);   // sized for either matching precision
  return backward ? L.b_total : L.f_total;
}

extern "C" int alad_train_losses_fwd(const alad_train_losses_args* a, void* stream) {
  using namespace alad;
  int rc = check_common(a, "alad_train_losses_fwd");
  if (rc) return rc;
  const TrainLayout L = train_layout(a->B, a->S_im, a->S_s, a->d, a->precision, 1);
  ALAD_REQUIRE(a->workspace_bytes >= L.f_total, "alad_train_losses_fwd: workspace too small (%lld < %lld)",
               (long long)a->workspace_bytes, (long long)L.f_total);
  ALAD_REQUIRE(a->losses, "alad_train_losses_fwd: NULL output");
  cudaStream_t st = as_stream(stream);
  const int32_t B = a->B;
  if (B == 0) {
    ALAD_CUDA(cudaMemsetAsync(a->losses, 0, 3 * sizeof(float), st));
    return ALAD_OK;
  }
  ALAD_REQUIRE(a->M && a->S, "alad_train_losses_fwd: NULL output");
  ALAD_REQUIRE(!a->want_grad || (a->G_m && a->G_a && (a->dM || !a->with_distill)), "alad_train_losses_fwd: NULL gradient buffer");
  ALAD_REQUIRE(a->im_cls && a->s_cls && a->im_set && a->s_seq, "alad_train_losses_fwd: NULL input");
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  int32_t* row_arg = reinterpret_cast<int32_t*>(ws + L.f_args);
  int32_t* col_arg = row_arg + B;
  static thread_local std::vector<int32_t> ones;
  if ((int)ones.size() < B) ones.assign((size_t)B, 1);

  SideLane* lane = nullptr;
  if ((rc = side_lane(&lane))) return rc;
  const bool fork_f = (fork_mask(B) & 1) != 0;
  if (fork_f) {
    ALAD_CUDA(cudaEventRecord(lane->fork, st));
    ALAD_CUDA(cudaStreamWaitEvent(lane->side, lane->fork, 0));
  }
  void* side = fork_f ? (void*)lane->side : stream;
  int32_t* row_arg_m = reinterpret_cast<int32_t*>(ws + L.f_args_m);
  int32_t* col_arg_m = row_arg_m + B;

  // ---- matching head (side stream): M = im_cls @ s_cls.T (dot_sim, alad/loss.py:8-11) + hinge (loss.py:42-67)
  alad_scores_fused_args f;
  memset(&f, 0, sizeof(f));
  f.max_x = a->im_cls; f.max_stride_b = a->ld_im_cls; f.max_stride_s = a->ld_im_cls;
  f.sum_x = a->s_cls;  f.sum_stride_b = a->ld_s_cls;  f.sum_stride_s = a->ld_s_cls;
  f.n_max = B; f.S_max = 1; f.slot0_max = 0;
  f.n_sum = B; f.S_sum = 1; f.slot0_sum = 0;
  f.d = a->d;
  f.max_count = ones.data(); f.sum_count = ones.data(); f.max_clamp = nullptr;
  f.precision = a->precision_m; f.epilogue = 1; f.normalize = 0; f.eps = 0.f;
  f.S = a->M; f.ldS = B; f.transpose_out = 0;
  f.workspace = ws + L.f_scores_m; f.workspace_bytes = L.f_scores_m_bytes;
  if ((rc = alad_scores_fused(&f, side))) return rc;
  if ((rc = alad_triplet_fwd_bwd(a->M, B, B, a->margin_m, a->max_violation_m, a->losses + 0, a->want_grad ? a->G_m : nullptr, B,
                                 row_arg_m, col_arg_m, ws + L.f_loss_m, side)))
    return rc;
  if (fork_f) ALAD_CUDA(cudaEventRecord(lane->join, lane->side));

  // ---- alignment head: S = MrSw(im_set, s_seq) (loss.py:80-125) + hinge
  f.max_x = a->im_set; f.max_stride_b = a->im_stride_b; f.max_stride_s = a->im_stride_s;
  f.sum_x = a->s_seq;  f.sum_stride_b = a->s_stride_b;  f.sum_stride_s = a->s_stride_s;
  f.S_max = a->S_im; f.slot0_max = 1;
  f.S_sum = a->S_s;  f.slot0_sum = 1;
  f.max_count = a->nr; f.sum_count = a->nw; f.max_clamp = a->clamp;
  f.precision = a->precision; f.epilogue = 0; f.normalize = 1; f.eps = 1e-12f;
  f.S = a->S;
  f.workspace = ws + L.f_scores; f.workspace_bytes = L.f_scores_bytes;
  if ((rc = alad_scores_fused(&f, stream))) return rc;
  if ((rc = alad_triplet_fwd_bwd(a->S, B, B, a->margin_a, a->max_violation_a, a->losses + 1, a->want_grad ? a->G_a : nullptr, B,
                                 row_arg, col_arg, ws + L.f_loss, stream)))
    return rc;

  // ---- join: M and the matching loss are complete from here on
  if (fork_f) ALAD_CUDA(cudaStreamWaitEvent(st, lane->join, 0));
  // ---- distillation: ListNet(teacher = S detached, student = M) (loss.py:370,427-445)
  if (a->with_distill) {
    if ((rc = alad_listnet_fwd_bwd(a->S, B, a->M, B, B, a->temperature, a->listnet_eps, a->losses + 2,
                                   a->want_grad ? a->dM : nullptr, B, ws + L.f_loss, stream)))
      return rc;
  } else {
    ALAD_CUDA(cudaMemsetAsync(a->losses + 2, 0, sizeof(float), st));
  }
  return ALAD_OK;
}

extern "C" int alad_train_losses_bwd(const alad_train_losses_args* a, void* stream) {
  using namespace alad;
  int rc = check_common(a, "alad_train_losses_bwd");
  if (rc) return rc;
  const TrainLayout L = train_layout(a->B, a->S_im, a->S_s, a->d, a->precision, 1);
  ALAD_REQUIRE(a->workspace_bytes >= L.b_total, "alad_train_losses_bwd: workspace too small (%lld < %lld)",
               (long long)a->workspace_bytes, (long long)L.b_total);
  ALAD_REQUIRE(a->g, "alad_train_losses_bwd: NULL upstream gradients");
  cudaStream_t st = as_stream(stream);
  const int32_t B = a->B, d = a->d;
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  const bool use_m = a->has_g_m && a->G_m, use_d = a->has_g_d && a->with_distill && a->dM, use_a = a->has_g_a && a->G_a;
  SideLane* lane = nullptr;
  bool forked = false;

  // ---- matching head: dL/dM = g[0] * G_m + g[2] * dM;  d im_cls = (dL/dM) @ s_cls,  d s_cls = (dL/dM).T @ im_cls
  if (a->d_im_cls || a->d_s_cls) {
    if (B == 0 || !(use_m || use_d)) {
      if (a->d_im_cls && B) ALAD_CUDA(cudaMemsetAsync(a->d_im_cls, 0, (size_t)B * d * sizeof(float), st));
      if (a->d_s_cls && B) ALAD_CUDA(cudaMemsetAsync(a->d_s_cls, 0, (size_t)B * d * sizeof(float), st));
    } else {
      ALAD_REQUIRE(a->im_cls && a->s_cls, "alad_train_losses_bwd: NULL matching inputs");
      // the matching head's backward (transposes + two GEMMs) runs on the side stream next to the sparse MrSw backward
      if ((rc = side_lane(&lane))) return rc;
      forked = (fork_mask(B) & 2) != 0;
      if (forked) {
        ALAD_CUDA(cudaEventRecord(lane->fork, st));
        ALAD_CUDA(cudaStreamWaitEvent(lane->side, lane->fork, 0));
      }
      cudaStream_t ss = forked ? lane->side : st;
      void* side = forked ? (void*)lane->side : stream;
      float* Gt = reinterpret_cast<float*>(ws + L.b_gt);
      float* GtT = reinterpret_cast<float*>(ws + L.b_gtT);
      float* imT = reinterpret_cast<float*>(ws + L.b_imT);
      float* sT = reinterpret_cast<float*>(ws + L.b_sT);
      const float* A = use_m ? a->G_m : nullptr;
      const float* Bm = use_d ? a->dM : nullptr;
      // with only one of the two terms present it takes the "A" slot
      if (!A) {
        if ((rc = axpby_t(Bm, B, a->g + 2, nullptr, 0, nullptr, B, B, a->d_im_cls ? Gt : nullptr, B, a->d_s_cls ? GtT : nullptr, B, ss)))
          return rc;
      } else if ((rc = axpby_t(A, B, a->g + 0, Bm, B, a->g + 2, B, B, a->d_im_cls ? Gt : nullptr, B, a->d_s_cls ? GtT : nullptr, B, ss))) {
        return rc;
      }
      static thread_local std::vector<int32_t> ones;
      const int n1 = B > d ? B : d;
      if ((int)ones.size() < n1) ones.assign((size_t)n1, 1);
      alad_scores_fused_args f;
      memset(&f, 0, sizeof(f));
      f.n_max = B; f.S_max = 1; f.slot0_max = 0;
      f.n_sum = d; f.S_sum = 1; f.slot0_sum = 0;
      f.d = B;                                          // the contraction runs over the other batch index
      f.max_stride_b = B; f.max_stride_s = B; f.sum_stride_b = B; f.sum_stride_s = B;
      f.max_count = ones.data(); f.sum_count = ones.data();
      f.precision = 1; f.epilogue = 1; f.normalize = 0; f.eps = 0.f;   // split precision: fp32-grade gradients
      f.ldS = d; f.transpose_out = 0;
      f.workspace = ws + L.b_scores; f.workspace_bytes = L.b_scores_bytes;
      if (a->d_im_cls) {
        if ((rc = axpby_t(a->s_cls, a->ld_s_cls, nullptr, nullptr, 0, nullptr, B, d, nullptr, 0, sT, B, ss))) return rc;
        f.max_x = Gt; f.sum_x = sT; f.S = a->d_im_cls;
        if ((rc = alad_scores_fused(&f, side))) return rc;
      }
      if (a->d_s_cls) {
        if ((rc = axpby_t(a->im_cls, a->ld_im_cls, nullptr, nullptr, 0, nullptr, B, d, nullptr, 0, imT, B, ss))) return rc;
        f.max_x = GtT; f.sum_x = imT; f.S = a->d_s_cls;
        if ((rc = alad_scores_fused(&f, side))) return rc;
      }
      if (forked) ALAD_CUDA(cudaEventRecord(lane->join, lane->side));
    }
  }

  // ---- alignment head: dL/dS = g[1] * G_a (<= 3B non-zeros) -> sparse MrSw backward
  if (a->d_im_set || a->d_s_seq) {
    ALAD_REQUIRE(a->d_im_set && a->d_s_seq, "alad_train_losses_bwd: d_im_set and d_s_seq go together");
    if (B == 0) return ALAD_OK;                        // (B == 0 never forks)
    ALAD_REQUIRE(a->im_set && a->s_seq, "alad_train_losses_bwd: NULL alignment inputs");
    int32_t* cnt = reinterpret_cast<int32_t*>(ws + L.b_cnt);
    static thread_local std::vector<int32_t> stage;
    stage.resize((size_t)2 * B);
    memcpy(stage.data(), a->nr, sizeof(int32_t) * B);
    memcpy(stage.data() + B, a->nw, sizeof(int32_t) * B);
    ALAD_CUDA(cudaMemcpyAsync(cnt, stage.data(), sizeof(int32_t) * 2 * B, cudaMemcpyHostToDevice, st));
    alad_mrsw_bwd_args m;
    memset(&m, 0, sizeof(m));
    m.im = a->im_set; m.im_stride_b = a->im_stride_b; m.im_stride_s = a->im_stride_s;
    m.s = a->s_seq;   m.s_stride_b = a->s_stride_b;   m.s_stride_s = a->s_stride_s;
    m.Bi = B; m.S_im = a->S_im; m.Bc = B; m.S_s = a->S_s; m.d = d;
    m.nr = cnt; m.nw = cnt + B;
    m.G0 = use_a ? a->G_a : nullptr; m.ldG0 = B; m.g0_scale = a->g + 1;
    m.G1 = nullptr; m.ldG1 = 0;
    m.d_im = a->d_im_set; m.d_s = a->d_s_seq;
    m.eps = 1e-12f; m.region_extent = 0;
    m.max_pairs = (int64_t)B * B;
    m.workspace = ws + L.b_mrsw; m.workspace_bytes = L.b_mrsw_bytes;
    m.d_im_stride_b = a->d_im_stride_b; m.d_im_stride_s = a->d_im_stride_s;
    m.d_s_stride_b = a->d_s_stride_b;   m.d_s_stride_s = a->d_s_stride_s;
    if ((rc = alad_mrsw_scores_bwd(&m, stream))) return rc;
  }
  if (forked) ALAD_CUDA(cudaStreamWaitEvent(st, lane->join, 0));
  return ALAD_OK;
}
