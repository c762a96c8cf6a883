// alad_mrsw_scores_fwd: all-pairs region x word cosine GEMM on tcgen05/TMEM with the
// TERAN-style MrSw pooling (max over regions, sum over words) fused in the epilogue.
// Replaces alad/loss.py:97-125 of the reference (expand + batched matmul + masked_fill +
// max(2)[0].sum(2)); the B x B x R x W tensor never exists.
//
// Mapping
//   M (TMEM lanes, 128 / tile)   = densely packed valid WORD rows of all captions
//   N (TMEM columns, 240 / tile) = densely packed valid REGION rows, whole images per tile
//   K                            = feature dim (bf16, or 3x for the split-precision mode)
// Persistent kernel, 6 warps per CTA; default variant = CTA PAIRS (tcgen05 cta_group::2, one 256 x 240 tile per pair
// and iteration, 74 pairs on 148 SMs), single-CTA variant (128 x 240) kept for A/B:
//   warps 0-3  epilogue: tcgen05.ld windows -> FMNMX3 max over each image's columns -> smem -> ordered
//              per-caption row sums -> <= 2 atomic addends per S entry (bit-reproducible)
//   warp 4     TMA producer (4-stage smem ring, 128B swizzle, L2 prefetch of the next word rows)
//   warp 5     tcgen05.mma issuer (one lane of the leader CTA), 2 accumulator stages in TMEM
#include <math.h>
#include <stdlib.h>

#include "common.h"
#include "sm100_ptx.cuh"

// This file is compiled twice: as is (bf16 operands, tcgen05 kind::f16) and from mrsw_fwd_tf32.cu with ALAD_MMA_TF32 defined
// (fp32 operands read as TF32, kind::tf32; dense kernel only).  Two translation units rather than one more template
// parameter, so that the bf16 kernel is generated from exactly the source it was tuned with (its code generation is
// sensitive: profiles/r01_tile_order_sweep.md, "build reproducibility").  A pipeline stage is 128-byte swizzle rows in both
// (64 bf16 or 32 floats), and one MMA consumes 32 bytes of a row (K = 16 or 8), so nothing else differs.
#ifdef ALAD_MMA_TF32
#define mrsw_fwd_kernel mrsw_fwd_tf32_kernel
#define ALAD_MAKE_IDESC make_idesc_tf32_f32
#define ALAD_UMMA_CG1 umma_tf32
#define ALAD_UMMA_CG2 umma_tf32_cg2
#else
#define ALAD_MAKE_IDESC make_idesc_bf16_f32
#define ALAD_UMMA_CG1 umma_bf16
#define ALAD_UMMA_CG2 umma_bf16_cg2
#endif

namespace alad {

constexpr int BM = ALAD_TILE_M;
constexpr int BN = ALAD_TILE_N;
constexpr int BK = ALAD_TILE_K;
constexpr int UMMA_K = 16;
constexpr int A_BYTES = BM * BK * 2;          // 16 KiB: 128 word rows x 64 bf16
constexpr int V_STRIDE = ALAD_MAX_SEG + 1;    // 33 floats: conflict-free row-major scratch
constexpr int V_BYTES = BM * V_STRIDE * 4;
constexpr int EPI_WARPS = 4;
constexpr int EPI_THREADS = EPI_WARPS * 32;
constexpr int THREADS = EPI_THREADS + 64;
constexpr int ACC_STAGES = 2;
constexpr int ACC_COLS = 256;                 // column stride between accumulator stages
constexpr int TMEM_COLS = 512;
constexpr int NTILE_WORDS = sizeof(alad_ntile) / 4;   // 20
static_assert(NTILE_WORDS <= 32, "one lane per table word");
constexpr int PTILE_WORDS = sizeof(alad_ptile) / 4;   // 24
static_assert(PTILE_WORDS <= 32 && ALAD_PTILE_SLOTS == 8, "one lane per table word; slot fields are unrolled by 8");
// word offsets inside alad_ptile
constexpr int PT_ROW0 = 0, PT_CAP_LO = 1, PT_CAP_HI = 2, PT_NSEG = 3, PT_CLAMP = 4, PT_SLOT_ROW = 5, PT_SLOT_IMG = 13,
              PT_SLOT_W = 21;

// Per-variant geometry.  CG = 1: one CTA per 128 x 240 tile.  CG = 2: a CTA pair (cta_group::2)
// per 256 x 240 tile -- each CTA stages its own 128 word rows and HALF of the region rows, which
// cuts the shared-memory fill per SM from 46 KB to 31 KB per K block.  Ring depth: 4 stages cover the L2 round
// trip; deeper rings only lower the clock the power cap allows (4 stages 322 ms, 6 stages 341 ms per COCO-5k launch
// on the same box, profiles/r01_tile_order_sweep.md section 10).
template <int CG>
struct Cfg {
#ifndef ALAD_STAGES_CG2
#define ALAD_STAGES_CG2 4
#endif
#ifndef ALAD_RES_KB
#define ALAD_RES_KB 0
#endif
  static constexpr int STAGES = CG == 2 ? ALAD_STAGES_CG2 : 4;
  // K blocks of the region tile kept RESIDENT in shared memory across consecutive tiles (CTA pairs in the
  // lock-step tile order meet the same region tile for a whole sweep): those B loads are skipped
  static constexpr int RES = CG == 2 ? ALAD_RES_KB : 0;
  static constexpr int B_ROWS = BN / CG;                       // region rows staged by this CTA
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int OFF_A = 0;
  static constexpr int OFF_B = OFF_A + STAGES * A_BYTES;
  static constexpr int OFF_RES = OFF_B + STAGES * B_BYTES;
  static constexpr int OFF_V = OFF_RES + RES * B_BYTES;
  static constexpr int OFF_CAP = OFF_V + 2 * V_BYTES;          // int capS[2][128]
  static constexpr int OFF_TAB = OFF_CAP + 2 * BM * 4;         // uint32 tab[4 warps][32]
  static constexpr int OFF_RUN = OFF_TAB + EPI_WARPS * 32 * 4; // uint32 run_start_mask[2][4]
  static constexpr int OFF_BAR = OFF_RUN + 2 * EPI_WARPS * 4;  // mbarriers
  static constexpr int NUM_BARS = 2 * STAGES + 2 * ACC_STAGES + 2;   // + resident full / empty
  static constexpr int OFF_TMEMPTR = OFF_BAR + NUM_BARS * 8;
  static constexpr int SMEM_USED = OFF_TMEMPTR + 16;
  static constexpr int SMEM_BYTES = SMEM_USED + 1024;          // slack for manual 1024 B alignment
  static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KiB per-CTA shared memory limit");
  static_assert(A_BYTES % 1024 == 0 && B_BYTES % 1024 == 0, "SWIZZLE_128B tiles must stay 1024 B aligned");
};

struct MrswParams {
  const int32_t* row_cap;
  const alad_ntile* ntiles;
  float* S;
  long long ld_seg;   // address = S + seg_item * ld_seg + row_item * ld_row  (image-major: ldS, 1; transposed: 1, ldS)
  long long ld_row;
  long long n_word_rows;
  long long n_region_rows;
  int n_mtiles;
  int n_ntiles;
  int num_kb;
  int epilogue;
  int n_block;      // N tiles swept per pass over the M tiles (their region rows stay hot in L2)
  int l2_hints;     // TMA L2 policies: bit 0 = words evict_first, bit 1 = region block evict_last
  int l2_prefetch;  // 1: the CTAs cooperatively prefetch the next M unit's word rows into L2
  int b_resident;   // 1: keep the first Cfg::RES K blocks of the unit's region tile resident in shared memory
  // pair-list mode (LIST = true, alad_mrsw_scores_pairs): tiles come from a device table instead of the dense
  // (word tile, region tile) enumeration; every tile names its own word rows and up to 8 gathered image slots
  const alad_ptile* ptiles;
  const int32_t* n_ptiles;   // device scalar
  int max_ptiles;
  int slot_rows;
  int a_bytes;               // bytes of word rows a pair-list tile loads per K block (word_box_rows x 128)
};

// full = the tile lies in a complete block of n_block region tiles (the last block may be shorter)
__device__ __forceinline__ void tile_coord(int t, int n_mtiles, int n_ntiles, int n_block, int& mt, int& nt, bool& full) {
  const int per_block = n_mtiles * n_block;
  const int nb = t / per_block;
  const int rem = t - nb * per_block;
  const int nb_size = min(n_block, n_ntiles - nb * n_block);
  mt = rem / nb_size;
  nt = nb * n_block + (rem - mt * nb_size);
  full = nb_size == n_block;
}
__device__ __forceinline__ void tile_coord(int t, int n_mtiles, int n_ntiles, int n_block, int& mt, int& nt) {
  bool full;
  tile_coord(t, n_mtiles, n_ntiles, n_block, mt, nt, full);
}

template <int CG, bool LIST>
__global__ void __launch_bounds__(THREADS, 1)
mrsw_fwd_kernel(const __grid_constant__ CUtensorMap map_words, const __grid_constant__ CUtensorMap map_regions,
                const MrswParams p) {
  static_assert(!LIST || CG == 1, "the pair-list mode runs single-CTA tiles (one or two captions per 128 word rows)");
  using C = Cfg<CG>;
  constexpr int STAGES = C::STAGES;
  extern __shared__ uint8_t smem_raw[];
  // manual 1024 B alignment by OFFSET (keeps the pointer in the shared address space -> LDS/STS)
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);

  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + C::OFF_BAR);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tfull_bar = bars + 2 * STAGES;
  uint64_t* tempty_bar = bars + 2 * STAGES + ACC_STAGES;
  uint64_t* res_full = bars + 2 * STAGES + 2 * ACC_STAGES;
  uint64_t* res_empty = res_full + 1;
  const bool use_res = C::RES > 0 && p.b_resident != 0;
  const int res_kb = min(C::RES, p.num_kb);
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + C::OFF_TMEMPTR);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int cta_rank = (CG == 2) ? static_cast<int>(cluster_ctarank()) : 0;   // rank inside the CTA pair
  const bool leader = cta_rank == 0;
  const int unit = blockIdx.x / CG;                  // work unit: a CTA (CG=1) or a CTA pair (CG=2)
  const int n_units = gridDim.x / CG;
  const int n_munits = (p.n_mtiles + CG - 1) / CG;   // M tiles are consumed CG at a time
  int total_tiles = n_munits * p.n_ntiles;
  if constexpr (LIST) total_tiles = min(__ldg(p.n_ptiles), p.max_ptiles);

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&map_words);
    tma_prefetch_desc(&map_regions);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < ACC_STAGES; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], EPI_THREADS * CG);   // the leader's copy collects both CTAs' epilogues
    }
    mbar_init(res_full, 1);
    mbar_init(res_empty, 1);
    fence_mbar_init();
  }
  if (warp == 5) {
    if (CG == 2) {
      tmem_alloc_cg2(tmem_ptr_smem, TMEM_COLS);
      tmem_relinquish_cg2();
    } else {
      tmem_alloc(tmem_ptr_smem, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 4) {
    // ============================== TMA producer (every CTA stages its own operands) ==========
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // the word rows stream past once per region block; the region block is re-read by every M unit
      const bool hints = p.l2_hints != 0;
      int res_nt = -1;
      uint32_t res_loads = 0;
      const uint64_t pol_words = (p.l2_hints & 1) ? l2_policy_evict_first() : l2_policy_evict_normal();
      const uint64_t pol_regions = (p.l2_hints & 2) ? l2_policy_evict_last() : l2_policy_evict_normal();
      // pair-list mode: the record of the NEXT tile is fetched while the current one streams (a dependent global
      // load at the head of every tile would drain the 4-stage ring)
      int nx_row0 = 0, nx_nseg = 0, nx_srow[ALAD_PTILE_SLOTS] = {};
      auto fetch_rec = [&](int t) {
        const int32_t* rec = reinterpret_cast<const int32_t*>(&p.ptiles[t]);
        nx_row0 = __ldg(rec + PT_ROW0);
        nx_nseg = __ldg(rec + PT_NSEG);
#pragma unroll
        for (int s_i = 0; s_i < ALAD_PTILE_SLOTS; ++s_i) nx_srow[s_i] = __ldg(rec + PT_SLOT_ROW + s_i);
      };
      if constexpr (LIST) {
        if (unit < total_tiles) fetch_rec(unit);
      }
      for (int t = unit; t < total_tiles; t += n_units) {
        if constexpr (LIST) {
          // pair-list tile: 128 word rows from m_row0, one TMA box of slot_rows region rows per image slot
          const int m_row0 = nx_row0;
          const int nseg = nx_nseg;
          int srow[ALAD_PTILE_SLOTS];
#pragma unroll
          for (int s_i = 0; s_i < ALAD_PTILE_SLOTS; ++s_i) srow[s_i] = nx_srow[s_i];
          if (t + n_units < total_tiles) fetch_rec(t + n_units);
          const uint32_t slot_bytes = static_cast<uint32_t>(p.slot_rows) * (BK * 2);
          for (int kb = 0; kb < p.num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1u);
            uint8_t* sa = smem + C::OFF_A + stage * A_BYTES;
            uint8_t* sb = smem + C::OFF_B + stage * C::B_BYTES;
            mbar_expect_tx(&full_bar[stage], static_cast<uint32_t>(p.a_bytes) + static_cast<uint32_t>(nseg) * slot_bytes);
            tma_load_2d(sa, &map_words, &full_bar[stage], kb * BK, m_row0);
#pragma unroll
            for (int s_i = 0; s_i < ALAD_PTILE_SLOTS; ++s_i)
              if (s_i < nseg) tma_load_2d(sb + s_i * slot_bytes, &map_regions, &full_bar[stage], kb * BK, srow[s_i]);
            if (++stage == STAGES) {
              stage = 0;
              phase ^= 1u;
            }
          }
          continue;
        }
        int mu, nt;
        bool full_block;
        tile_coord(t, n_munits, p.n_ntiles, p.n_block, mu, nt, full_block);
        const bool res_t = CG == 2 && use_res && full_block;    // resident rows only where the unit keeps its tile
        const int n_row0 = __ldg(&p.ntiles[nt].row_start) + cta_rank * C::B_ROWS;
        const int m_row0 = (mu * CG + cta_rank) * BM;
        if (res_t && nt != res_nt) {
          // new region tile for this unit: wait until the MMAs of the previous one have read the resident
          // rows (in both CTAs), then load the first res_kb K blocks once
          if (res_loads > 0) mbar_wait(res_empty, (res_loads - 1) & 1u);
          if (leader) mbar_expect_tx(res_full, 2 * res_kb * C::B_BYTES);
          for (int kb = 0; kb < res_kb; ++kb)
            tma_load_2d_cg2(smem + C::OFF_RES + kb * C::B_BYTES, &map_regions, res_full, kb * BK, n_row0);
          res_nt = nt;
          ++res_loads;
        }
        if (p.l2_prefetch && t + n_units < total_tiles) {
          // All units sweep the M units in near lock step, so the word rows of the next M unit are
          // about to be requested by every SM at once; concurrent first-touch misses on the same
          // lines are not merged into one DRAM read.  Pull them into L2 one tile ahead instead:
          // box b of the next unit (BM*CG/128 row groups x num_kb K blocks) is prefetched by the
          // CTAs whose index is congruent to b (one per die).
          int mu2, nt2;
          tile_coord(t + n_units, n_munits, p.n_ntiles, p.n_block, mu2, nt2);
          if (mu2 != mu) {
            const int n_boxes = CG * p.num_kb;
            const int half = static_cast<int>(gridDim.x) / 2 > 0 ? static_cast<int>(gridDim.x) / 2 : 1;
            for (int b = static_cast<int>(blockIdx.x) % half; b < n_boxes; b += half)
              tma_prefetch_l2_2d(&map_words, (b % p.num_kb) * BK, (mu2 * CG + b / p.num_kb) * BM);
          }
        }
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          uint8_t* sa = smem + C::OFF_A + stage * A_BYTES;
          uint8_t* sb = smem + C::OFF_B + stage * C::B_BYTES;
          if (CG == 2) {
            // completion bytes of BOTH CTAs are credited to the leader's barrier
            const bool b_ring = !(res_t && kb < res_kb);       // resident K blocks need no B load
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * (A_BYTES + (b_ring ? C::B_BYTES : 0)));
            if (hints) {
              tma_load_2d_cg2_hint(sa, &map_words, &full_bar[stage], kb * BK, m_row0, pol_words);
              if (b_ring) tma_load_2d_cg2_hint(sb, &map_regions, &full_bar[stage], kb * BK, n_row0, pol_regions);
            } else {
              tma_load_2d_cg2(sa, &map_words, &full_bar[stage], kb * BK, m_row0);
              if (b_ring) tma_load_2d_cg2(sb, &map_regions, &full_bar[stage], kb * BK, n_row0);
            }
          } else {
            mbar_expect_tx(&full_bar[stage], C::STAGE_BYTES);
            if (hints) {
              tma_load_2d_hint(sa, &map_words, &full_bar[stage], kb * BK, m_row0, pol_words);
              tma_load_2d_hint(sb, &map_regions, &full_bar[stage], kb * BK, n_row0, pol_regions);
            } else {
              tma_load_2d(sa, &map_words, &full_bar[stage], kb * BK, m_row0);
              tma_load_2d(sb, &map_regions, &full_bar[stage], kb * BK, n_row0);
            }
          }
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ============================== MMA issuer (leader CTA only) ==============================
    if (lane == 0 && leader) {
      constexpr uint32_t idesc = ALAD_MAKE_IDESC(BM * CG, BN);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      int res_nt = -1;
      uint32_t res_loads = 0;
      for (int t = unit; t < total_tiles; t += n_units, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1u;
        int nt = 0, nt_next = -1;
        bool res_t = false;
        if (CG == 2 && use_res) {
          int mu;
          bool full_block, full_next = false;
          tile_coord(t, n_munits, p.n_ntiles, p.n_block, mu, nt, full_block);
          if (t + n_units < total_tiles) tile_coord(t + n_units, n_munits, p.n_ntiles, p.n_block, mu, nt_next, full_next);
          if (!full_next) nt_next = -1;                // the next tile does not use the resident rows
          res_t = full_block;
          if (res_t && nt != res_nt) {                 // first tile on a new region tile: wait for its resident rows
            mbar_wait(res_full, res_loads & 1u);
            res_nt = nt;
            ++res_loads;
          }
        }
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
        for (int kb = 0; kb < p.num_kb; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint64_t a_desc = make_sw128_kmajor_desc(smem_u32(smem + C::OFF_A + stage * A_BYTES));
          const uint64_t b_desc = (res_t && kb < res_kb)
                                      ? make_sw128_kmajor_desc(smem_u32(smem + C::OFF_RES + kb * C::B_BYTES))
                                      : make_sw128_kmajor_desc(smem_u32(smem + C::OFF_B + stage * C::B_BYTES));
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // +32 B per K step inside the 128 B swizzle row: start-address field is in 16 B units
            if (CG == 2) ALAD_UMMA_CG2(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
            else         ALAD_UMMA_CG1(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
          if (CG == 2) umma_commit_cg2_both(&empty_bar[stage]); else umma_commit(&empty_bar[stage]);
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        // accumulator complete -> epilogue warps (of both CTAs)
        if (CG == 2) umma_commit_cg2_both(&tfull_bar[acc]); else umma_commit(&tfull_bar[acc]);
        // last tile on this region tile: the resident rows may be overwritten once these MMAs are done
        if (res_t && nt_next != nt) umma_commit_cg2_both(res_empty);
      }
    }
    __syncwarp();
  } else {
    // ============================== epilogue (warps 0-3) ======================
    const int row = warp * 32 + lane;                       // TMEM lane == row of the M tile
    float* V = reinterpret_cast<float*>(smem + C::OFF_V);
    int* capS = reinterpret_cast<int*>(smem + C::OFF_CAP);
    uint32_t* tab = reinterpret_cast<uint32_t*>(smem + C::OFF_TAB) + warp * 32;
    uint32_t* runS = reinterpret_cast<uint32_t*>(smem + C::OFF_RUN);
    const bool mrsw = p.epilogue == 0;

    // metadata of a tile: N-tile record word (lanes 0..11), caption of my row and of the row above
    uint32_t nx_tab = 0;
    int nx_cap = 0, nx_above = 0;
    auto fetch_meta = [&](int t) {
      if constexpr (LIST) {
        if (lane < PTILE_WORDS) nx_tab = __ldg(reinterpret_cast<const uint32_t*>(&p.ptiles[t]) + lane);
        const long long mrow = static_cast<long long>(__shfl_sync(0xffffffffu, static_cast<int>(nx_tab), PT_ROW0)) + row;
        nx_cap = mrow < p.n_word_rows ? __ldg(&p.row_cap[mrow]) : -1;
        nx_above = (row > 0 && lane == 0 && mrow - 1 < p.n_word_rows) ? __ldg(&p.row_cap[mrow - 1]) : -1;
        return;
      }
      int mu, nt;
      tile_coord(t, n_munits, p.n_ntiles, p.n_block, mu, nt);
      const int mt = mu * CG + cta_rank;
      if (lane < NTILE_WORDS) nx_tab = __ldg(reinterpret_cast<const uint32_t*>(&p.ntiles[nt]) + lane);
      if (mrsw) {
        const long long mrow = static_cast<long long>(mt) * BM + row;
        nx_cap = __ldg(&p.row_cap[mrow]);
        nx_above = (row > 0 && lane == 0) ? __ldg(&p.row_cap[mrow - 1]) : 0;
      }
    };
    if (unit < total_tiles) fetch_meta(unit);

    int it = 0;
    for (int t = unit; t < total_tiles; t += n_units, ++it) {
      int mu = 0, nt = 0;
      if constexpr (!LIST) tile_coord(t, n_munits, p.n_ntiles, p.n_block, mu, nt);
      const int mt = mu * CG + cta_rank;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1u;
      const int buf = it & 1;
      const long long mrow = static_cast<long long>(mt) * BM + row;
      const int mycap = nx_cap;
      if (lane < (LIST ? PTILE_WORDS : NTILE_WORDS)) tab[lane] = nx_tab;
      if (mrsw) {
        // caption runs of this M tile: a row starts a run when its caption differs from the row above
        int above = __shfl_up_sync(0xffffffffu, mycap, 1);
        if (lane == 0) above = nx_above;
        const bool is_start = (row == 0) || (mycap != above);
        const uint32_t smask = __ballot_sync(0xffffffffu, is_start);
        if (lane == 0) runS[buf * EPI_WARPS + warp] = smask;
        capS[buf * BM + row] = mycap;
      }
      __syncwarp();
      const int n_row0 = static_cast<int>(tab[0]);
      const int img0 = static_cast<int>(tab[1]);
      // shfl => provably warp-uniform
      const int nseg = __shfl_sync(0xffffffffu, static_cast<int>(tab[LIST ? PT_NSEG : 2]), 0);
      const uint32_t clamp = tab[LIST ? PT_CLAMP : 3];
      uint32_t mydesc;                                                                // descriptor of image `lane`
      int cap_lo = 0, cap_hi = 0x7fffffff, my_img = img0 + lane;
      if constexpr (LIST) {
        const int sl = lane & (ALAD_PTILE_SLOTS - 1);
        const uint32_t w = (tab[PT_SLOT_W + (sl >> 2)] >> ((sl & 3) * 8)) & 0xffu;
        mydesc = static_cast<uint32_t>(sl * p.slot_rows) | (w << 8);
        cap_lo = static_cast<int>(tab[PT_CAP_LO]);
        cap_hi = static_cast<int>(tab[PT_CAP_HI]);
        my_img = static_cast<int>(tab[PT_SLOT_IMG + sl]);
      } else {
        mydesc = (tab[4 + (lane >> 1)] >> ((lane & 1) * 16)) & 0xffffu;
      }
      // prefetch the next tile's metadata while this one is processed
      if (t + n_units < total_tiles) fetch_meta(t + n_units);

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + acc * ACC_COLS;

      if (mrsw) {
        // ---- phase 1: per-row max over each image's columns -> V[row][seg].
        // Every image is covered by windows of P consecutive columns that lie entirely inside
        // it (P = 32, or the largest power of two <= width; the last window overlaps the
        // previous one), so no masking is needed and each window is a FMNMX3 tree.
        float* Vrow = V + buf * (BM * V_STRIDE) + row * V_STRIDE;
        for (int s_i = 0; s_i < nseg; ++s_i) {
          const uint32_t dsc = __shfl_sync(0xffffffffu, mydesc, s_i);   // warp-uniform (start | width << 8)
          const uint32_t c0 = taddr + (dsc & 0xffu);
          const int w = static_cast<int>(dsc >> 8);
          float m;
          if (w >= 32) {
            uint32_t va[32], vb[32];
            tmem_ld_32x32(c0, va);
            tmem_ld_32x32(c0 + (w > 64 ? 32 : w - 32), vb);
            tmem_ld_wait();
            m = fmaxf(tree_max_bits(va), tree_max_bits(vb));
            for (int off = 64; off < w; off += 32) {                    // images wider than 64 columns
              tmem_ld_32x32(c0 + min(off, w - 32), va);
              tmem_ld_wait();
              m = fmaxf(m, tree_max_bits(va));
            }
          } else if (w >= 16) {
            uint32_t va[16], vb[16];
            tmem_ld_32x16(c0, va);
            tmem_ld_32x16(c0 + w - 16, vb);
            tmem_ld_wait();
            m = fmaxf(tree_max_bits(va), tree_max_bits(vb));
          } else if (w >= 8) {
            uint32_t va[8], vb[8];
            tmem_ld_32x8(c0, va);
            tmem_ld_32x8(c0 + w - 8, vb);
            tmem_ld_wait();
            m = fmaxf(tree_max_bits(va), tree_max_bits(vb));
          } else if (w >= 4) {
            uint32_t va[4], vb[4];
            tmem_ld_32x4(c0, va);
            tmem_ld_32x4(c0 + w - 4, vb);
            tmem_ld_wait();
            m = fmaxf(tree_max_bits(va), tree_max_bits(vb));
          } else if (w >= 2) {
            uint32_t va[2], vb[2];
            tmem_ld_32x2(c0, va);
            tmem_ld_32x2(c0 + w - 2, vb);
            tmem_ld_wait();
            m = fmaxf(tree_max_bits(va), tree_max_bits(vb));
          } else {
            uint32_t va[1];
            tmem_ld_32x1(c0, va);
            tmem_ld_wait();
            m = __uint_as_float(va[0]);
          }
          Vrow[s_i] = m;
        }
        tmem_ld_wait();
        tc_fence_before();
        // stage may be overwritten.  Every thread arrives for itself: one arrival per warp behind a __syncwarp measured
        // 1.3-1.9 % SLOWER on the same box (325-327 ms against 321 ms per COCO-5k launch, profiles/r02_scoring_kernel.md)
        if (CG == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]);
        asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory");
        // ---- phase 2: per-caption row sums in a fixed order; warp w owns caption runs w, w+4, ...
        const int* caps = capS + buf * BM;
        const float* Vcol = V + buf * (BM * V_STRIDE) + lane;
        const bool active = lane < nseg;
        const float floor_v = ((clamp >> lane) & 1u) ? 0.f : -INFINITY;   // masked slots take part in the max as 0
        float* Sout = p.S + static_cast<long long>(my_img) * p.ld_seg;
        unsigned long long lo = runS[buf * EPI_WARPS + 0] | (static_cast<unsigned long long>(runS[buf * EPI_WARPS + 1]) << 32);
        unsigned long long hi = runS[buf * EPI_WARPS + 2] | (static_cast<unsigned long long>(runS[buf * EPI_WARPS + 3]) << 32);
        auto pop_start = [&]() -> int {                    // next run start (128 when exhausted)
          if (lo) {
            const int b = __ffsll(static_cast<long long>(lo)) - 1;
            lo &= lo - 1;
            return b;
          }
          if (hi) {
            const int b = __ffsll(static_cast<long long>(hi)) - 1;
            hi &= hi - 1;
            return 64 + b;
          }
          return BM;
        };
        int start = pop_start();                           // row 0 always starts a run
        int ord = 0;
        while (start < BM) {
          const int end = pop_start();
          if ((ord & 3) == warp) {
            const int cap = caps[start];
            if (cap >= cap_lo && cap < cap_hi && active) {
              float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
              int r = start;
              for (; r + 4 <= end; r += 4) {
                s0 += fmaxf(Vcol[(r + 0) * V_STRIDE], floor_v);
                s1 += fmaxf(Vcol[(r + 1) * V_STRIDE], floor_v);
                s2 += fmaxf(Vcol[(r + 2) * V_STRIDE], floor_v);
                s3 += fmaxf(Vcol[(r + 3) * V_STRIDE], floor_v);
              }
              for (; r < end; ++r) s0 += fmaxf(Vcol[r * V_STRIDE], floor_v);
              if constexpr (LIST) Sout[cap * p.ld_row] = (s0 + s1) + (s2 + s3);   // every listed pair lives in ONE tile
              else atomicAdd(Sout + cap * p.ld_row, (s0 + s1) + (s2 + s3));
            }
          }
          start = end;
          ++ord;
        }
      } else {
        // ---- plain GEMM epilogue: S[region row, word row] = accumulator (coalesced over lanes)
        const int ncols = static_cast<int>(min(static_cast<long long>(BN), p.n_region_rows - n_row0));
        const bool row_ok = mrow < p.n_word_rows;
#pragma unroll 1
        for (int chunk = 0; chunk < 8; ++chunk) {
          uint32_t v[32];
          tmem_ld_32x32(taddr + chunk * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int col = chunk * 32 + c;
            if (row_ok && col < ncols) p.S[static_cast<long long>(n_row0 + col) * p.ld_seg + mrow * p.ld_row] = __uint_as_float(v[c]);
          }
        }
        tc_fence_before();
        if (CG == 2) mbar_arrive_leader(&tempty_bar[acc]); else mbar_arrive(&tempty_bar[acc]);
      }
    }
  }

  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();   // the pair stays alive until both CTAs are done
  if (warp == 5) {
    tc_fence_after();
    if (CG == 2) tmem_dealloc_cg2(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(sym);
  }
  return fn;
}

// Experiment knobs: compile-time defaults; only a -DALAD_TUNING_ENV build looks at the environment.
static inline long long tuning_env(const char* name, long long dflt) {
#ifdef ALAD_TUNING_ENV
  const char* e = getenv(name);
  return e ? atoll(e) : dflt;
#else
  (void)name;
  return dflt;
#endif
}

// [rows, Kp] bf16 row-major; box = 64 elements (128 B, one swizzle row) x box_rows.
static int make_map(CUtensorMap* m, const void* ptr, long long rows, int Kp, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(ALAD_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(Kp), static_cast<cuuint64_t>(rows)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(Kp) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(BK), static_cast<cuuint32_t>(box_rows)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ALAD_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

// the dense kernel of THIS translation unit's operand format (cfg: grid, block, stream and cluster attribute already set)
#ifdef ALAD_MMA_TF32
int launch_mrsw_dense_tf32(int cg, cudaLaunchConfig_t* cfg, const CUtensorMap* map_w, const CUtensorMap* map_r, const MrswParams* p) {
#else
int launch_mrsw_dense_tf32(int cg, cudaLaunchConfig_t* cfg, const CUtensorMap* map_w, const CUtensorMap* map_r, const MrswParams* p);
int launch_mrsw_dense_bf16(int cg, cudaLaunchConfig_t* cfg, const CUtensorMap* map_w, const CUtensorMap* map_r, const MrswParams* p) {
#endif
  if (cg == 2) {
    ALAD_CUDA(cudaFuncSetAttribute(mrsw_fwd_kernel<2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<2>::SMEM_BYTES));
    cfg->dynamicSmemBytes = Cfg<2>::SMEM_BYTES;
    ALAD_CUDA(cudaLaunchKernelEx(cfg, mrsw_fwd_kernel<2, false>, *map_w, *map_r, *p));
  } else {
    ALAD_CUDA(cudaFuncSetAttribute(mrsw_fwd_kernel<1, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM_BYTES));
    cfg->dynamicSmemBytes = Cfg<1>::SMEM_BYTES;
    ALAD_CUDA(cudaLaunchKernelEx(cfg, mrsw_fwd_kernel<1, false>, *map_w, *map_r, *p));
  }
  return ALAD_OK;
}

}  // namespace alad

#ifndef ALAD_MMA_TF32
extern "C" int alad_mrsw_scores_fwd(const alad_mrsw_fwd_args* a, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(a != nullptr, "alad_mrsw_scores_fwd: NULL args");
  // S is [Ni, ldS] with Ni = #segment items (images) and Nc = #row items (captions), or its
  // transpose [Nc, ldS] when transpose_out is set; it is zeroed here
  const int out_rows = a->transpose_out ? a->Nc : a->Ni, out_cols = a->transpose_out ? a->Ni : a->Nc;
  ALAD_REQUIRE(a->S != nullptr && a->Ni >= 0 && a->Nc >= 0 && a->ldS >= out_cols, "alad_mrsw_scores_fwd: bad output");
  ALAD_REQUIRE(a->Kp > 0 && a->Kp % BK == 0, "alad_mrsw_scores_fwd: Kp=%d must be a positive multiple of %d", a->Kp, BK);
  ALAD_REQUIRE(a->n_word_rows >= 0 && a->n_region_rows >= 0 && a->n_word_rows < (1ll << 31) &&
                   a->n_region_rows < (1ll << 31),
               "alad_mrsw_scores_fwd: bad row counts");
  ALAD_REQUIRE(a->epilogue == 0 || a->epilogue == 1, "alad_mrsw_scores_fwd: unknown epilogue %d", a->epilogue);
  ALAD_REQUIRE(a->accumulate == 0 || (a->accumulate == 1 && a->epilogue == 0), "alad_mrsw_scores_fwd: accumulate needs epilogue 0");
  ALAD_REQUIRE(a->operand_format == 0 || a->operand_format == 1, "alad_mrsw_scores_fwd: unknown operand_format %d", a->operand_format);
  cudaStream_t st = as_stream(stream);
  if (a->Ni > 0 && a->Nc > 0 && !a->accumulate) {
    if (a->ldS == out_cols) {
      ALAD_CUDA(cudaMemsetAsync(a->S, 0, sizeof(float) * (size_t)out_rows * (size_t)out_cols, st));
    } else {
      ALAD_CUDA(cudaMemset2DAsync(a->S, sizeof(float) * (size_t)a->ldS, 0, sizeof(float) * (size_t)out_cols,
                                  (size_t)out_rows, st));
    }
  }
  if (a->n_word_rows == 0 || a->n_region_rows == 0 || a->n_ntiles == 0 || a->Ni == 0 || a->Nc == 0) return ALAD_OK;
  ALAD_REQUIRE(a->words && a->regions && a->ntiles, "alad_mrsw_scores_fwd: NULL operand");
  ALAD_REQUIRE(a->epilogue == 1 || a->row_cap, "alad_mrsw_scores_fwd: NULL row_cap");
  ALAD_REQUIRE((reinterpret_cast<uintptr_t>(a->words) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->regions) & 15) == 0,
               "alad_mrsw_scores_fwd: operands must be 16-byte aligned");

  MrswParams p = {};
  p.row_cap = a->row_cap;
  p.ntiles = a->ntiles;
  p.S = a->S;
  p.ld_seg = a->transpose_out ? 1 : a->ldS;
  p.ld_row = a->transpose_out ? a->ldS : 1;
  p.n_word_rows = a->n_word_rows;
  p.n_region_rows = a->n_region_rows;
  p.n_mtiles = (int)((a->n_word_rows + BM - 1) / BM);
  p.n_ntiles = a->n_ntiles;
  p.num_kb = a->Kp / BK;
  p.epilogue = a->epilogue;
  const long long total = (long long)p.n_mtiles * p.n_ntiles;
  ALAD_REQUIRE(total < (1ll << 31), "alad_mrsw_scores_fwd: too many tiles (%lld)", total);

  int cg = a->cta_group;
  if (cg == 0) cg = static_cast<int>(tuning_env("ALAD_CTA_GROUP", 2)) == 1 ? 1 : 2;
  ALAD_REQUIRE(cg == 1 || cg == 2, "alad_mrsw_scores_fwd: cta_group must be 0, 1 or 2");
  if (sm_count() < 2) cg = 1;

  CUtensorMap map_w, map_r;
  int rc = make_map(&map_w, a->words, a->n_word_rows, a->Kp, BM);
  if (rc) return rc;
  rc = make_map(&map_r, a->regions, a->n_region_rows, a->Kp, BN / cg);
  if (rc) return rc;

  const long long units = (long long)((p.n_mtiles + cg - 1) / cg) * p.n_ntiles;
  int ctas = a->num_ctas > 0 ? a->num_ctas : sm_count();
  if ((long long)ctas > units * cg) ctas = (int)(units * cg);
  ctas = (ctas / cg) * cg;
  if (ctas < cg) ctas = cg;
  {
    // Tile order.  The region tiles are swept in blocks of n_block tiles while all word tiles stream
    // past.  With n_block a divisor of the number of work units (CTAs or CTA pairs), unit u meets the
    // SAME region tile(s) on every pass (t = u + k * units), all units read the same word rows at the
    // same time, and the block (n_block x 240 rows x Kp bf16) stays L2-resident: measured on B200 at
    // COCO-5k shape, n_block = 74 runs ~8 % faster than a 30 MB block of 61 tiles and 25 % faster than
    // 16 tiles (profiles/r01_tile_order_sweep.md).  Pick the largest divisor whose block fits the budget
    // (64 MB by default: beyond ~100 MB the block thrashes the 126 MB L2).
    // The knobs below are fixed at their measured defaults; a build with -DALAD_TUNING_ENV reads them from the
    // environment per call (tools/sweep_tile_order.py), the product build never calls getenv.
    const int units_in_flight = ctas / cg;
    const long long ev = tuning_env("ALAD_L2_BLOCK_MB", 0);
    const long long budget = (ev > 0 ? ev : 64) << 20;
    const long long tile_bytes = (long long)BN * a->Kp * 2;
    long long nb = 0;
    for (int j = 1; j <= units_in_flight && nb == 0; ++j)
      if (units_in_flight % j == 0 && (units_in_flight / j) * tile_bytes <= budget) nb = units_in_flight / j;
    if (nb < 8 || ev > 0) nb = budget / tile_bytes;    // tiny grids / explicit budget: plain budget rule
    nb = nb < 8 ? 8 : nb;
    const long long eb = tuning_env("ALAD_N_BLOCK", 0);   // block size in tiles (experiments)
    if (eb > 0) nb = eb;
    p.n_block = (int)(nb > p.n_ntiles ? p.n_ntiles : nb);
    if ((long long)p.n_mtiles * p.n_block >= (1ll << 31)) p.n_block = 8;
    p.l2_hints = (int)tuning_env("ALAD_L2_HINTS", 0);
    // the word-row prefetch is OFF: with the 6-stage ring it measured +0.6 % at 74 tiles / bf16 (-8 % at 148 tiles,
    // -19 % at 37 tiles / 3x split), but with the 4-stage ring the same-box A/B gives 322-324 ms without it against
    // 337 ms with it (profiles/r01_tile_order_sweep.md section 10)
    p.l2_prefetch = (int)tuning_env("ALAD_L2_PREFETCH", 0);
    // resident region K blocks (compile-time Cfg<2>::RES > 0): only where a unit keeps its region tile
    p.b_resident = (int)tuning_env("ALAD_B_RESIDENT", p.n_block > 0 && units_in_flight % p.n_block == 0 ? 1 : 0);
  }

  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(THREADS);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (a->operand_format == 1) {
    if ((rc = launch_mrsw_dense_tf32(cg, &cfg, &map_w, &map_r, &p))) return rc;
  } else if ((rc = launch_mrsw_dense_bf16(cg, &cfg, &map_w, &map_r, &p))) {
    return rc;
  }
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

// Pair-list scoring (two-stage retrieval, stage 2): same mainloop and epilogue, tiles from a device table.
extern "C" int alad_mrsw_scores_pairs(const alad_mrsw_pairs_args* a, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(a != nullptr, "alad_mrsw_scores_pairs: NULL args");
  const int out_cols = a->transpose_out ? a->Ni : a->Nc;
  ALAD_REQUIRE(a->Ni >= 0 && a->Nc >= 0 && a->ldS >= out_cols, "alad_mrsw_scores_pairs: bad output");
  ALAD_REQUIRE(a->Kp > 0 && a->Kp % BK == 0, "alad_mrsw_scores_pairs: Kp=%d must be a positive multiple of %d", a->Kp, BK);
  ALAD_REQUIRE(a->n_word_rows >= 0 && a->n_region_rows >= 0 && a->n_word_rows < (1ll << 31) && a->n_region_rows < (1ll << 31),
               "alad_mrsw_scores_pairs: bad row counts");
  ALAD_REQUIRE(a->slot_rows >= BN / ALAD_PTILE_SLOTS && a->slot_rows <= BN,
               "alad_mrsw_scores_pairs: slot_rows=%d outside [%d, %d]", a->slot_rows, BN / ALAD_PTILE_SLOTS, BN);
  ALAD_REQUIRE(a->max_ptiles >= 0, "alad_mrsw_scores_pairs: bad max_ptiles");
  if (a->max_ptiles == 0 || a->n_word_rows == 0 || a->n_region_rows == 0 || a->Ni == 0 || a->Nc == 0) return ALAD_OK;
  ALAD_REQUIRE(a->words && a->regions && a->row_cap && a->ptiles && a->n_ptiles && a->S, "alad_mrsw_scores_pairs: NULL pointer");
  ALAD_REQUIRE((reinterpret_cast<uintptr_t>(a->words) & 15) == 0 && (reinterpret_cast<uintptr_t>(a->regions) & 15) == 0,
               "alad_mrsw_scores_pairs: operands must be 16-byte aligned");
  MrswParams p = {};
  p.row_cap = a->row_cap;
  p.S = a->S;
  p.ld_seg = a->transpose_out ? 1 : a->ldS;
  p.ld_row = a->transpose_out ? a->ldS : 1;
  p.n_word_rows = a->n_word_rows;
  p.n_region_rows = a->n_region_rows;
  p.n_mtiles = 1;
  p.n_ntiles = 1;
  p.n_block = 1;
  p.num_kb = a->Kp / BK;
  p.epilogue = 0;
  p.ptiles = a->ptiles;
  p.n_ptiles = a->n_ptiles;
  p.max_ptiles = a->max_ptiles;
  p.slot_rows = a->slot_rows;
  const int box_rows = a->word_box_rows > 0 ? a->word_box_rows : BM;
  ALAD_REQUIRE(box_rows % 8 == 0 && box_rows >= 8 && box_rows <= BM, "alad_mrsw_scores_pairs: word_box_rows=%d must be a multiple of 8 in [8, %d]",
               box_rows, BM);
  p.a_bytes = box_rows * BK * 2;
  CUtensorMap map_w, map_r;
  int rc = make_map(&map_w, a->words, a->n_word_rows, a->Kp, box_rows);
  if (rc) return rc;
  rc = make_map(&map_r, a->regions, a->n_region_rows, a->Kp, a->slot_rows);
  if (rc) return rc;
  int ctas = a->num_ctas > 0 ? a->num_ctas : sm_count();
  if (ctas > a->max_ptiles) ctas = a->max_ptiles;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(ctas);
  cfg.blockDim = dim3(THREADS);
  cfg.stream = as_stream(stream);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  ALAD_CUDA(cudaFuncSetAttribute(mrsw_fwd_kernel<1, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg<1>::SMEM_BYTES));
  cfg.dynamicSmemBytes = Cfg<1>::SMEM_BYTES;
  ALAD_CUDA(cudaLaunchKernelEx(&cfg, mrsw_fwd_kernel<1, true>, map_w, map_r, p));
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}
#endif  // !ALAD_MMA_TF32
