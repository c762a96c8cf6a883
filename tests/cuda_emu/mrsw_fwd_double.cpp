// TEST INFRASTRUCTURE ONLY -- CPU double of alad_mrsw_scores_fwd (csrc/mrsw_fwd.cu, the tcgen05 / TMA scoring kernel,
// which cannot run under the host-thread emulator).  It implements the entry point's CONTRACT as include/alad_b200.h
// states it, on host pointers: packed bf16 rows, the region tile table, the caption of every word row, S zeroed and
// accumulated.  With it the emulated library is complete, so the native compositions and the Python host layer can be
// exercised on CPU.  It never ships and nothing under aladin_b200/ loads it.
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <vector>

#include "common.h"

namespace {

inline float bf16_to_float(uint16_t v) {
  const uint32_t u = static_cast<uint32_t>(v) << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

// fp32 accumulation over the bf16 products (exact in fp32 each), like the tensor pipe up to summation order
inline float dot_rows(const uint16_t* a, const uint16_t* b, int Kp) {
  double acc = 0.0;
  for (int k = 0; k < Kp; ++k) acc += static_cast<double>(bf16_to_float(a[k])) * bf16_to_float(b[k]);
  return static_cast<float>(acc);
}

// operand_format 1: the rows are fp32 (Kp / 2 floats) and the tensor core reads them as TF32 (10-bit mantissa, truncated)
inline float tf32_of(float v) {
  uint32_t u;
  memcpy(&u, &v, 4);
  u &= 0xffffe000u;
  memcpy(&v, &u, 4);
  return v;
}
inline float dot_rows_tf32(const uint16_t* a, const uint16_t* b, int Kp) {
  const float* fa = reinterpret_cast<const float*>(a);
  const float* fb = reinterpret_cast<const float*>(b);
  double acc = 0.0;
  for (int k = 0; k < Kp / 2; ++k) acc += static_cast<double>(tf32_of(fa[k])) * tf32_of(fb[k]);
  return static_cast<float>(acc);
}

}  // namespace

extern "C" int alad_mrsw_scores_fwd(const alad_mrsw_fwd_args* a, void*) {
  using namespace alad;
  ALAD_REQUIRE(a != nullptr, "alad_mrsw_scores_fwd: NULL args");
  const int out_rows = a->transpose_out ? a->Nc : a->Ni, out_cols = a->transpose_out ? a->Ni : a->Nc;
  ALAD_REQUIRE(a->S != nullptr && a->Ni >= 0 && a->Nc >= 0 && a->ldS >= out_cols, "alad_mrsw_scores_fwd: bad output");
  ALAD_REQUIRE(a->Kp > 0 && a->Kp % ALAD_TILE_K == 0, "alad_mrsw_scores_fwd: Kp=%d must be a positive multiple of %d", a->Kp,
               ALAD_TILE_K);
  ALAD_REQUIRE(a->epilogue == 0 || a->epilogue == 1, "alad_mrsw_scores_fwd: unknown epilogue %d", a->epilogue);
  if (!a->accumulate)
    for (int r = 0; r < out_rows; ++r) memset(a->S + (long long)r * a->ldS, 0, sizeof(float) * (size_t)out_cols);
  if (a->n_word_rows == 0 || a->n_region_rows == 0 || a->n_ntiles == 0 || a->Ni == 0 || a->Nc == 0) return ALAD_OK;
  ALAD_REQUIRE(a->words && a->regions && a->ntiles, "alad_mrsw_scores_fwd: NULL operand");
  ALAD_REQUIRE(a->epilogue == 1 || a->row_cap, "alad_mrsw_scores_fwd: NULL row_cap");
  const uint16_t* words = static_cast<const uint16_t*>(a->words);
  const uint16_t* regions = static_cast<const uint16_t*>(a->regions);
  const long long ld_seg = a->transpose_out ? 1 : a->ldS, ld_row = a->transpose_out ? a->ldS : 1;
  auto dot_rows = [&](const uint16_t* x, const uint16_t* y, int Kp) {
    return a->operand_format == 1 ? dot_rows_tf32(x, y, Kp) : ::dot_rows(x, y, Kp);
  };
  for (int t = 0; t < a->n_ntiles; ++t) {
    const alad_ntile& nt = a->ntiles[t];
    if (a->epilogue == 1) {
      // plain GEMM: S[region row, word row]
      for (long long n = nt.row_start; n < nt.row_start + ALAD_TILE_N && n < a->n_region_rows; ++n)
        for (long long m = 0; m < a->n_word_rows; ++m)
          a->S[n * ld_seg + m * ld_row] = dot_rows(regions + n * a->Kp, words + m * a->Kp, a->Kp);
      continue;
    }
    ALAD_REQUIRE(nt.nseg >= 0 && nt.nseg <= ALAD_MAX_SEG, "alad_mrsw_scores_fwd: bad tile table");
    for (int s = 0; s < nt.nseg; ++s) {
      const int col0 = nt.seg[s] & 0xff, width = nt.seg[s] >> 8;
      if (width == 0) continue;
      const bool clamp = (nt.clamp_bits >> s) & 1u;
      float* Sout = a->S + (long long)(nt.img0 + s) * ld_seg;
      for (long long m = 0; m < a->n_word_rows; ++m) {
        const int cap = a->row_cap[m];
        if (cap < 0) continue;
        float best = clamp ? 0.f : -INFINITY;
        for (int c = 0; c < width; ++c)
          best = fmaxf(best, dot_rows(regions + (long long)(nt.row_start + col0 + c) * a->Kp, words + m * a->Kp, a->Kp));
        Sout[cap * ld_row] += best;
      }
    }
  }
  return ALAD_OK;
}

// CPU double of alad_mrsw_scores_pairs: the contract of the pair-list entry point (include/alad_b200.h) -- every tile
// scores captions cap_lo .. cap_hi-1 (rows from m_row0, 128 at most) against its image slots and stores
// S[slot_img, caption]; rows of other captions and unlisted pairs are left untouched.
extern "C" int alad_mrsw_scores_pairs(const alad_mrsw_pairs_args* a, void*) {
  using namespace alad;
  ALAD_REQUIRE(a != nullptr, "alad_mrsw_scores_pairs: NULL args");
  ALAD_REQUIRE(a->slot_rows >= ALAD_TILE_N / ALAD_PTILE_SLOTS && a->slot_rows <= ALAD_TILE_N, "alad_mrsw_scores_pairs: bad slot_rows");
  if (a->max_ptiles == 0 || a->n_word_rows == 0 || a->n_region_rows == 0 || a->Ni == 0 || a->Nc == 0) return ALAD_OK;
  ALAD_REQUIRE(a->words && a->regions && a->row_cap && a->ptiles && a->n_ptiles && a->S, "alad_mrsw_scores_pairs: NULL pointer");
  const uint16_t* words = static_cast<const uint16_t*>(a->words);
  const uint16_t* regions = static_cast<const uint16_t*>(a->regions);
  const long long ld_seg = a->transpose_out ? 1 : a->ldS, ld_row = a->transpose_out ? a->ldS : 1;
  const int n_t = *a->n_ptiles < a->max_ptiles ? *a->n_ptiles : a->max_ptiles;
  for (int t = 0; t < n_t; ++t) {
    const alad_ptile& pt = a->ptiles[t];
    ALAD_REQUIRE(pt.nseg >= 0 && pt.nseg <= ALAD_TILE_N / a->slot_rows, "alad_mrsw_scores_pairs: bad tile table");
    for (int s = 0; s < pt.nseg; ++s) {
      const int width = pt.slot_w[s];
      if (width == 0) continue;
      const bool clamp = (pt.clamp_bits >> s) & 1u;
      float* Sout = a->S + (long long)pt.slot_img[s] * ld_seg;
      int cur = -1;
      float sum = 0.f;
      for (int r = 0; r <= ALAD_TILE_M; ++r) {
        const long long m = (long long)pt.m_row0 + r;
        const int cap = (r < ALAD_TILE_M && m < a->n_word_rows) ? a->row_cap[m] : -2;
        if (cap != cur) {
          if (cur >= pt.cap_lo && cur < pt.cap_hi) Sout[cur * ld_row] = sum;
          cur = cap;
          sum = 0.f;
        }
        if (cap < 0) continue;
        float best = clamp ? 0.f : -INFINITY;
        for (int c = 0; c < width; ++c)
          best = fmaxf(best, dot_rows(regions + (long long)(pt.slot_row[s] + c) * a->Kp, words + m * a->Kp, a->Kp));
        sum += best;
      }
    }
  }
  return ALAD_OK;
}
