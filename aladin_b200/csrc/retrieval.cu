// alad_mrsw_retrieval: both retrieval directions of alad/evaluation.py:158-327 (i2t :213-223, t2i :303-308) from PACKED
// operands in one native call, the score matrix optional.  Nothing new is computed here: the entry point composes
// alad_region_tiles, alad_mrsw_scores_fwd, alad_col_gt, alad_rank_fused and alad_topk_merge on one stream with one
// workspace.  With S == NULL the gallery images are scored block by block into one reusable [block_images, Nc] buffer:
//   pass 1  every block against ITS OWN captions (the block diagonal, 1 / n_blocks of the work), launched on the same
//           256-row word units as the full pass, so the ground-truth scores are bit-identical to the entries of the
//           full matrix they stand for;
//   pass 2  every block against all captions: the i2t ranks of its rows are final, the t2i "images ahead" counts add
//           up over the blocks, the per-caption top-k lists are merged as they come (exactly what the shards of the
//           multi-GPU path do across ranks).
// Results equal the ranking of the dense matrix (tests/test_gpu_retrieval.py).
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.h"

namespace alad {

static inline int64_t r_up(int64_t x, int64_t a) { return (x + a - 1) / a * a; }

__global__ void add_i32_kernel(int* __restrict__ acc, const int* __restrict__ x, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) acc[i] += x[i];
}

struct RetrievalLayout {
  int64_t rows, ldb;        // block buffer geometry
  int64_t off_S, off_gt, off_cnt, off_tiles, off_ss, off_si, off_rank_ws, total;
};

static RetrievalLayout retrieval_layout(int32_t Ni, int32_t Nc, int32_t k, int32_t block_images, int32_t with_S) {
  RetrievalLayout L;
  const int64_t B = with_S ? Ni : std::min<int64_t>(Ni, std::max<int64_t>(std::max(block_images, k), 1));
  L.rows = B > 0 ? B : 1;
  L.ldb = r_up(Nc > 0 ? Nc : 1, 4);                  // 16-byte aligned rows: the fused ranking sweep reads float4
  const int64_t nc = Nc > 0 ? Nc : 1;
  int64_t o = 0;
  L.off_S = o;       o = r_up(o + (with_S ? 0 : L.rows * L.ldb * 4), 256);
  L.off_gt = o;      o = r_up(o + nc * 4, 256);
  L.off_cnt = o;     o = r_up(o + nc * 4, 256);
  L.off_tiles = o;   o = r_up(o + L.rows * (int64_t)sizeof(alad_ntile), 256);
  L.off_ss = o;      o = r_up(o + 2 * nc * k * 4, 256);
  L.off_si = o;      o = r_up(o + 2 * nc * k * 4, 256);
  // the ranking workspace depends on the block height (row groups of the top-k select): full blocks and the last one
  const int32_t tail = (int32_t)(Ni % L.rows);
  int64_t rank_ws = alad_rank_fused_workspace_bytes((int32_t)L.rows, (int32_t)L.rows, Nc, k);
  if (tail) rank_ws = std::max(rank_ws, alad_rank_fused_workspace_bytes(tail, tail, Nc, k));
  L.off_rank_ws = o; o = r_up(o + rank_ws, 256);
  L.total = o;
  return L;
}

}  // namespace alad

extern "C" int64_t alad_mrsw_retrieval_workspace_bytes(int32_t Ni, int32_t Nc, int32_t k, int32_t block_images, int32_t with_S) {
  if (Ni < 0 || Nc < 0 || k <= 0 || k > 256) return -1;
  return alad::retrieval_layout(Ni, Nc, k, block_images, with_S).total;
}

extern "C" int alad_mrsw_retrieval(const alad_mrsw_retrieval_args* a, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(a != nullptr, "alad_mrsw_retrieval: NULL args");
  ALAD_REQUIRE(a->Ni >= 0 && a->Nc >= 0 && a->group > 0 && a->Kp > 0 && a->n_word_rows >= 0 && a->n_region_rows >= 0,
               "alad_mrsw_retrieval: bad shape");
  ALAD_REQUIRE(a->k > 0 && a->k <= 256, "alad_mrsw_retrieval: k=%d out of range", a->k);
  ALAD_REQUIRE(a->S == nullptr || a->ldS >= a->Nc, "alad_mrsw_retrieval: ldS < Nc");
  const int Ni = a->Ni, Nc = a->Nc, k = a->k;
  if (Ni == 0 && Nc == 0) return ALAD_OK;
  const RetrievalLayout L = retrieval_layout(Ni, Nc, k, a->block_images, a->S != nullptr);
  ALAD_REQUIRE(a->workspace && a->workspace_bytes >= L.total, "alad_mrsw_retrieval: workspace too small (%lld < %lld)",
               (long long)a->workspace_bytes, (long long)L.total);
  ALAD_REQUIRE((reinterpret_cast<uintptr_t>(a->workspace) & 255) == 0, "alad_mrsw_retrieval: workspace must be 256-byte aligned");
  ALAD_REQUIRE((a->rank_i2t && a->top1) || Ni == 0, "alad_mrsw_retrieval: NULL i2t outputs");
  ALAD_REQUIRE((a->rank_t2i && a->topk_score && a->topk_idx) || Nc == 0, "alad_mrsw_retrieval: NULL t2i outputs");
  ALAD_REQUIRE(Ni == 0 || a->nr, "alad_mrsw_retrieval: NULL nr");
  ALAD_REQUIRE(Nc == 0 || a->cap_row, "alad_mrsw_retrieval: NULL cap_row");
  const int64_t n_rows = a->n_word_rows;
  ALAD_REQUIRE(Nc == 0 || (a->cap_row[0] == 0 && a->cap_row[Nc] == n_rows), "alad_mrsw_retrieval: cap_row does not span the word rows");
  const bool have_ops = n_rows > 0 && a->n_region_rows > 0 && Ni > 0 && Nc > 0;
  ALAD_REQUIRE(!have_ops || (a->words && a->regions && a->row_cap), "alad_mrsw_retrieval: NULL operands");
  cudaStream_t st = as_stream(stream);
  uint8_t* ws = reinterpret_cast<uint8_t*>(a->workspace);
  float* gt = reinterpret_cast<float*>(ws + L.off_gt);
  int* cnt_blk = reinterpret_cast<int*>(ws + L.off_cnt);
  alad_ntile* tiles_dev = reinterpret_cast<alad_ntile*>(ws + L.off_tiles);
  float* stack_s = reinterpret_cast<float*>(ws + L.off_ss);
  int* stack_i = reinterpret_cast<int*>(ws + L.off_si);
  void* rank_ws = ws + L.off_rank_ws;
  const int64_t row_bytes = (int64_t)a->Kp * (a->operand_format == 1 ? 4 : 2);
  const int64_t unit = 2 * ALAD_TILE_M;
  const int64_t n_units = (n_rows + unit - 1) / unit;

  // first packed region row of every image
  std::vector<int64_t> roff((size_t)Ni + 1, 0);
  for (int i = 0; i < Ni; ++i) {
    ALAD_REQUIRE(a->nr[i] >= 0, "alad_mrsw_retrieval: nr[%d] < 0", i);
    roff[i + 1] = roff[i] + a->nr[i];
  }
  ALAD_REQUIRE(roff[Ni] == a->n_region_rows || !have_ops, "alad_mrsw_retrieval: nr does not add up to n_region_rows");
  static thread_local std::vector<alad_ntile> tiles;

  // out[0 : hi-lo, :] = scores of images [lo, hi) against the captions whose rows lie in word rows [r0, r1)
  auto score = [&](int lo, int hi, int64_t r0, int64_t r1, int64_t u1, float* out, int64_t ld) -> int {
    const int n = hi - lo;
    int n_tiles = 0;
    if (have_ops && r1 > r0) {
      tiles.assign((size_t)n, alad_ntile{});
      n_tiles = alad_region_tiles(a->nr + lo, a->clamp ? a->clamp + lo : nullptr, n, tiles.data(), n, nullptr);
      if (n_tiles < 0) return n_tiles;
    }
    if (n_tiles == 0) {
      if (Nc > 0 && n > 0) ALAD_CUDA(cudaMemset2DAsync(out, (size_t)ld * 4, 0, (size_t)Nc * 4, (size_t)n, st));
      return ALAD_OK;
    }
    ALAD_CUDA(cudaMemcpyAsync(tiles_dev, tiles.data(), sizeof(alad_ntile) * (size_t)n_tiles, cudaMemcpyHostToDevice, st));
    alad_mrsw_fwd_args f;
    memset(&f, 0, sizeof(f));
    f.words = static_cast<const uint8_t*>(a->words) + r0 * row_bytes;
    f.n_word_rows = r1 - r0;
    f.regions = static_cast<const uint8_t*>(a->regions) + roff[lo] * row_bytes;
    f.n_region_rows = roff[hi] - roff[lo];
    f.Kp = a->Kp;
    f.row_cap = a->row_cap + r0;                     // readable up to u1 * 256 rows (padding rows hold -1)
    (void)u1;
    f.ntiles = tiles_dev;
    f.n_ntiles = n_tiles;
    f.S = out;
    f.ldS = ld;
    f.Ni = n;
    f.Nc = Nc;
    f.epilogue = 0;
    f.operand_format = a->operand_format;
    return alad_mrsw_scores_fwd(&f, stream);
  };

  if (a->S) {          // the whole matrix is wanted: one pass, one ranking
    if (Ni > 0 && Nc > 0) {
      const int rc = score(0, Ni, 0, n_rows, n_units, a->S, a->ldS);
      if (rc) return rc;
    }
    return alad_rank_fused(a->S, a->ldS, Ni, Nc, a->group, 0, Ni, Nc, k, nullptr, a->rank_i2t, a->top1, a->rank_t2i,
                           a->topk_score, a->topk_idx, rank_ws, stream);
  }

  float* Sb = reinterpret_cast<float*>(ws + L.off_S);
  const int B = (int)L.rows;
  if (Nc > 0) {
    ALAD_CUDA(cudaMemsetAsync(gt, 0, sizeof(float) * (size_t)Nc, st));
    ALAD_CUDA(cudaMemsetAsync(a->rank_t2i, 0, sizeof(int32_t) * (size_t)Nc, st));
  }
  if (Ni == 0) {       // an empty gallery: no image is ahead of anything, the lists are empty
    return alad_rank_fused(Sb, L.ldb, 0, Nc, a->group, 0, 0, Nc, k, gt, nullptr, nullptr, a->rank_t2i, a->topk_score, a->topk_idx,
                           rank_ws, stream);
  }
  // ---- pass 1: the block diagonal -> ground-truth scores
  if (have_ops) {
    for (int lo = 0; lo < Ni; lo += B) {
      const int hi = std::min(Ni, lo + B);
      const int64_t c0 = std::min<int64_t>(Nc, (int64_t)a->group * lo), c1 = std::min<int64_t>(Nc, (int64_t)a->group * hi);
      if (c1 <= c0) continue;
      const int64_t u0 = a->cap_row[c0] / unit, u1 = std::min(n_units, (a->cap_row[c1] + unit - 1) / unit);
      int rc = score(lo, hi, u0 * unit, std::min(u1 * unit, n_rows), u1, Sb, L.ldb);
      if (rc) return rc;
      rc = alad_col_gt(Sb, L.ldb, hi - lo, Nc, a->group, lo, gt, stream);
      if (rc) return rc;
    }
  }
  // ---- pass 2: every block against all captions
  const size_t list_bytes = sizeof(float) * (size_t)Nc * k;
  bool first = true;
  for (int lo = 0; lo < Ni; lo += B) {
    const int hi = std::min(Ni, lo + B), n = hi - lo;
    int rc = score(lo, hi, 0, n_rows, n_units, Sb, L.ldb);
    if (rc) return rc;
    float* ls = stack_s + (first ? 0 : (size_t)Nc * k);
    int* li = stack_i + (first ? 0 : (size_t)Nc * k);
    rc = alad_rank_fused(Sb, L.ldb, n, Nc, a->group, lo, n, Nc, k, gt, a->rank_i2t + lo, a->top1 + lo, cnt_blk, ls, li, rank_ws,
                         stream);
    if (rc) return rc;
    if (Nc > 0) {
      add_i32_kernel<<<(Nc + 255) / 256, 256, 0, st>>>(a->rank_t2i, cnt_blk, Nc);
      if (!first) {      // merge the running lists (slot 0) with this block's (slot 1), keep the result as the running lists
        rc = alad_topk_merge(stack_s, stack_i, 2, Nc, k, a->topk_score, a->topk_idx, stream);
        if (rc) return rc;
        ALAD_CUDA(cudaMemcpyAsync(stack_s, a->topk_score, list_bytes, cudaMemcpyDeviceToDevice, st));
        ALAD_CUDA(cudaMemcpyAsync(stack_i, a->topk_idx, list_bytes, cudaMemcpyDeviceToDevice, st));
      }
    }
    first = false;
  }
  if (Nc > 0 && Ni <= B) {       // a single block: its lists are the result
    ALAD_CUDA(cudaMemcpyAsync(a->topk_score, stack_s, list_bytes, cudaMemcpyDeviceToDevice, st));
    ALAD_CUDA(cudaMemcpyAsync(a->topk_idx, stack_i, list_bytes, cudaMemcpyDeviceToDevice, st));
  }
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}
