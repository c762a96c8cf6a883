// alad_scores_fused: one entry point for "pack both operands + score" -- the whole body of
// AlignmentContrastiveLoss.forward up to the pooled similarity matrix (alad/loss.py:80-125) or of dot_sim /
// cosine_sim (alad/loss.py:8-18) on RAW fp32 tokens and HOST length arrays.  It composes the entry points of
// pack.cu and mrsw_fwd.cu; what it adds is the host side of a call done natively: valid-row offsets, the greedy
// region tile table, ONE metadata upload, and four launches, so that a training step pays one FFI call per
// criterion instead of a dozen Python-level tensor operations (the step is launch-latency bound at B <= 512).
#include <stdint.h>
#include <string.h>

#include <vector>

#include "common.h"

namespace alad {

static inline int64_t align_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

struct FusedLayout {
  int64_t Kp, rows_sum_ub, rows_max_ub, row_item_len;
  int64_t off_words, off_regions, off_row_item, off_meta;
  int64_t meta_sum_off, meta_max_off, meta_sum_cnt, meta_max_cnt, meta_tiles, meta_bytes;
  int64_t total;
};

static FusedLayout fused_layout(int32_t n_max, int32_t S_max, int32_t slot0_max, int32_t n_sum, int32_t S_sum,
                                int32_t slot0_sum, int32_t d, int32_t precision) {
  FusedLayout L;
  L.Kp = align_up((int64_t)d * (precision ? 3 : 1), ALAD_TILE_K);
  L.rows_sum_ub = (int64_t)n_sum * (S_sum > slot0_sum ? S_sum - slot0_sum : 0);
  L.rows_max_ub = (int64_t)n_max * (S_max > slot0_max ? S_max - slot0_max : 0);
  L.row_item_len = align_up(L.rows_sum_ub > 0 ? L.rows_sum_ub : 1, 2 * ALAD_TILE_M);
  int64_t o = 0;
  L.off_words = o;     o = align_up(o + (L.rows_sum_ub > 0 ? L.rows_sum_ub : 1) * L.Kp * 2, 256);
  L.off_regions = o;   o = align_up(o + (L.rows_max_ub > 0 ? L.rows_max_ub : 1) * L.Kp * 2, 256);
  L.off_row_item = o;  o = align_up(o + L.row_item_len * 4, 256);
  L.off_meta = o;
  int64_t m = 0;
  L.meta_sum_off = m;  m = align_up(m + 8ll * n_sum, 16);
  L.meta_max_off = m;  m = align_up(m + 8ll * n_max, 16);
  L.meta_sum_cnt = m;  m = align_up(m + 4ll * n_sum, 16);
  L.meta_max_cnt = m;  m = align_up(m + 4ll * n_max, 16);
  L.meta_tiles = m;    m = align_up(m + (int64_t)sizeof(alad_ntile) * (n_max > 0 ? n_max : 1), 16);
  L.meta_bytes = m;
  L.total = align_up(o + m, 256);
  return L;
}

}  // namespace alad

extern "C" int64_t alad_scores_fused_workspace_bytes(int32_t n_max, int32_t S_max, int32_t slot0_max, int32_t n_sum,
                                                     int32_t S_sum, int32_t slot0_sum, int32_t d, int32_t precision) {
  if (n_max < 0 || n_sum < 0 || S_max < 0 || S_sum < 0 || d <= 0) return -1;
  return alad::fused_layout(n_max, S_max, slot0_max, n_sum, S_sum, slot0_sum, d, precision).total;
}

extern "C" int alad_scores_fused(const alad_scores_fused_args* a, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(a != nullptr, "alad_scores_fused: NULL args");
  ALAD_REQUIRE(a->n_max >= 0 && a->n_sum >= 0 && a->S_max >= 0 && a->S_sum >= 0 && a->d > 0 && a->slot0_max >= 0 &&
                   a->slot0_sum >= 0,
               "alad_scores_fused: bad shape");
  ALAD_REQUIRE(a->epilogue == 0 || a->epilogue == 1, "alad_scores_fused: unknown epilogue %d", a->epilogue);
  ALAD_REQUIRE(a->precision == 0 || a->precision == 1, "alad_scores_fused: unknown precision %d", a->precision);
  const FusedLayout L = fused_layout(a->n_max, a->S_max, a->slot0_max, a->n_sum, a->S_sum, a->slot0_sum, a->d, a->precision);
  ALAD_REQUIRE(a->workspace && a->workspace_bytes >= L.total, "alad_scores_fused: workspace too small (%lld < %lld)",
               (long long)a->workspace_bytes, (long long)L.total);
  ALAD_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, "alad_scores_fused: workspace must be 256-byte aligned");
  ALAD_REQUIRE(a->n_max == 0 || a->max_count, "alad_scores_fused: NULL max_count");
  ALAD_REQUIRE(a->n_sum == 0 || a->sum_count, "alad_scores_fused: NULL sum_count");
  cudaStream_t st = as_stream(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);

  // ---- host bookkeeping into one staging blob (thread-local; a pageable cudaMemcpyAsync returns once the bytes
  //      are staged, so the buffer can be reused by the next call)
  static thread_local std::vector<uint8_t> blob;
  blob.assign((size_t)L.meta_bytes, 0);
  int64_t* sum_off = reinterpret_cast<int64_t*>(blob.data() + L.meta_sum_off);
  int64_t* max_off = reinterpret_cast<int64_t*>(blob.data() + L.meta_max_off);
  int32_t* sum_cnt = reinterpret_cast<int32_t*>(blob.data() + L.meta_sum_cnt);
  int32_t* max_cnt = reinterpret_cast<int32_t*>(blob.data() + L.meta_max_cnt);
  alad_ntile* tiles = reinterpret_cast<alad_ntile*>(blob.data() + L.meta_tiles);
  const int cap_sum = a->S_sum > a->slot0_sum ? a->S_sum - a->slot0_sum : 0;
  const int cap_max = a->S_max > a->slot0_max ? a->S_max - a->slot0_max : 0;
  int64_t rows_sum = 0, rows_max = 0;
  for (int j = 0; j < a->n_sum; ++j) {
    const int c = a->sum_count[j];
    ALAD_REQUIRE(c >= 0 && c <= cap_sum, "alad_scores_fused: sum_count[%d] = %d exceeds the container (%d)", j, c, cap_sum);
    sum_off[j] = rows_sum;
    sum_cnt[j] = c;
    rows_sum += c;
  }
  int n_tiles = 0;
  if (a->epilogue == 0) {
    for (int i = 0; i < a->n_max; ++i) {
      const int c = a->max_count[i];
      ALAD_REQUIRE(c >= 0 && c <= cap_max, "alad_scores_fused: max_count[%d] = %d exceeds the container (%d)", i, c, cap_max);
      max_cnt[i] = c;
    }
    n_tiles = alad_region_tiles(max_cnt, a->max_clamp, a->n_max, tiles, a->n_max > 0 ? a->n_max : 1, max_off);
    if (n_tiles < 0) return n_tiles;
    rows_max = a->n_max ? max_off[a->n_max - 1] + max_cnt[a->n_max - 1] : 0;
  } else {
    // plain GEMM: every item contributes max_count rows (1 for global vectors); tiles are blocks of 240 rows
    for (int i = 0; i < a->n_max; ++i) {
      const int c = a->max_count[i];
      ALAD_REQUIRE(c >= 0 && c <= cap_max, "alad_scores_fused: max_count[%d] = %d exceeds the container (%d)", i, c, cap_max);
      max_off[i] = rows_max;
      max_cnt[i] = c;
      rows_max += c;
    }
    n_tiles = (int)((rows_max + ALAD_TILE_N - 1) / ALAD_TILE_N);
    ALAD_REQUIRE(n_tiles <= (a->n_max > 0 ? a->n_max : 1), "alad_scores_fused: GEMM epilogue expects <= 1 row per item");
    for (int t = 0; t < n_tiles; ++t) {
      tiles[t].row_start = t * ALAD_TILE_N;
      tiles[t].img0 = t * ALAD_TILE_N;
      tiles[t].nseg = 1;
    }
  }
  if (L.meta_bytes > 0)
    ALAD_CUDA(cudaMemcpyAsync(ws + L.off_meta, blob.data(), (size_t)L.meta_bytes, cudaMemcpyHostToDevice, st));

  // ---- pack both operands
  int32_t* row_item = reinterpret_cast<int32_t*>(ws + L.off_row_item);
  if (a->epilogue == 0) ALAD_CUDA(cudaMemsetAsync(row_item, 0xFF, (size_t)L.row_item_len * 4, st));   // -1 = padding row
  alad_pack_args pk;
  memset(&pk, 0, sizeof(pk));
  pk.d = a->d;
  pk.Kp = (int32_t)L.Kp;
  pk.normalize = a->normalize;
  pk.eps = a->eps;
  if (rows_sum > 0) {
    pk.src = a->sum_x; pk.stride_b = a->sum_stride_b; pk.stride_s = a->sum_stride_s;
    pk.B = a->n_sum; pk.S = a->S_sum; pk.slot0 = a->slot0_sum;
    pk.count = reinterpret_cast<const int32_t*>(ws + L.off_meta + L.meta_sum_cnt);
    pk.row_off = reinterpret_cast<const int64_t*>(ws + L.off_meta + L.meta_sum_off);
    pk.dst = ws + L.off_words;
    pk.mode = a->precision ? 1 : 0;
    pk.row_item = a->epilogue == 0 ? row_item : nullptr;
    pk.item_base = 0;
    const int rc = alad_pack_tokens(&pk, stream);
    if (rc) return rc;
  }
  if (rows_max > 0) {
    pk.src = a->max_x; pk.stride_b = a->max_stride_b; pk.stride_s = a->max_stride_s;
    pk.B = a->n_max; pk.S = a->S_max; pk.slot0 = a->slot0_max;
    pk.count = reinterpret_cast<const int32_t*>(ws + L.off_meta + L.meta_max_cnt);
    pk.row_off = reinterpret_cast<const int64_t*>(ws + L.off_meta + L.meta_max_off);
    pk.dst = ws + L.off_regions;
    pk.mode = a->precision ? 2 : 0;
    pk.row_item = nullptr;
    const int rc = alad_pack_tokens(&pk, stream);
    if (rc) return rc;
  }

  // ---- score
  alad_mrsw_fwd_args f;
  memset(&f, 0, sizeof(f));
  f.words = ws + L.off_words;
  f.n_word_rows = rows_sum;
  f.regions = ws + L.off_regions;
  f.n_region_rows = rows_max;
  f.Kp = (int32_t)L.Kp;
  f.row_cap = a->epilogue == 0 ? row_item : nullptr;
  f.ntiles = reinterpret_cast<const alad_ntile*>(ws + L.off_meta + L.meta_tiles);
  f.n_ntiles = n_tiles;
  f.S = a->S;
  f.ldS = a->ldS;
  f.Ni = a->n_max;
  f.Nc = a->n_sum;
  f.epilogue = a->epilogue;
  f.num_ctas = 0;
  f.cta_group = 0;
  f.transpose_out = a->transpose_out;
  return alad_mrsw_scores_fwd(&f, stream);
}
