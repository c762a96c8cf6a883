// Two-stage retrieval, stage 2 bookkeeping (BASELINE config 5): turns the stage-1 shortlists into the tile table
// of alad_mrsw_scores_pairs, and re-ranks the shortlists from the stage-2 scores.  The reference pieces composed
// are alad/recall_auxiliary.py:30 (matching-head scores -> shortlist) and alad/loss.py:97-125 (alignment scores);
// the re-ranking replaces a numpy argsort per query over the shortlisted entries.  Integer / index work on CUDA
// cores: a few MB of lists, bitmaps and tables per call.
#include <math.h>

#include "common.h"

namespace alad {
namespace {

constexpr int PT_THREADS = 256;

__device__ __forceinline__ bool ahead(float v, int vi, float w, int wi) {  // (v,vi) ordered before (w,wi)?
  return v > w || (v == w && vi > wi);
}

// ---------------------------------------------------------------------------------- union bitmaps
// bitmap[g][l >> 5] bit (l & 31): local image l is needed by caption group g
__global__ void __launch_bounds__(PT_THREADS)
pairtile_mark_t2i_kernel(const int32_t* __restrict__ lists, long long n_entries, int k, const int32_t* __restrict__ cap_group,
                         int img_off, int n_loc, const int32_t* __restrict__ nr, int bw, uint32_t* __restrict__ bitmap) {
  const long long e = blockIdx.x * (long long)PT_THREADS + threadIdx.x;
  if (e >= n_entries) return;
  const int c = (int)(e / k);
  const int g = __ldg(cap_group + c);
  const int l = __ldg(lists + e) - img_off;
  if (g < 0 || l < 0 || l >= n_loc || __ldg(nr + l) <= 0) return;
  atomicOr(bitmap + (long long)g * bw + (l >> 5), 1u << (l & 31));
}

__global__ void __launch_bounds__(PT_THREADS)
pairtile_mark_i2t_kernel(const int32_t* __restrict__ lists, long long n_entries, int k, const int32_t* __restrict__ cap_group,
                         int Nc, const int32_t* __restrict__ nr, int bw, uint32_t* __restrict__ bitmap) {
  const long long e = blockIdx.x * (long long)PT_THREADS + threadIdx.x;
  if (e >= n_entries) return;
  const int l = (int)(e / k);
  const int c = __ldg(lists + e);
  if (c < 0 || c >= Nc || __ldg(nr + l) <= 0) return;
  const int g = __ldg(cap_group + c);
  if (g < 0) return;
  atomicOr(bitmap + (long long)g * bw + (l >> 5), 1u << (l & 31));
}

// Image blocks: the local images are cut into blocks of block_words * 32 images whose packed region rows fit the L2
// (a gathered slot is re-used by ~Nc*K/Ni caption groups; with the whole gallery in play every slot load would come
// from HBM: 200 GB per call at COCO-5k).  Tiles are emitted block-major, so one launch sweeps block after block.
// A "virtual group" v = block * n_groups + group owns the images of its group's union that fall into its block.
// one warp per virtual group: images in the union -> tiles of `slots` images
__global__ void __launch_bounds__(PT_THREADS)
pairtile_count_kernel(const uint32_t* __restrict__ bitmap, int bw, int n_groups, int n_blocks, int block_words, int slots,
                      int32_t* __restrict__ tiles_of) {
  const int v = blockIdx.x * (PT_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (v >= n_groups * n_blocks) return;
  const int b = v / n_groups, g = v - b * n_groups;
  const int w_lo = b * block_words, w_hi = min(bw, w_lo + block_words);
  int n = 0;
  for (int w = w_lo + lane; w < w_hi; w += 32) n += __popc(__ldg(bitmap + (long long)g * bw + w));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if (lane == 0) tiles_of[v] = (n + slots - 1) / slots;
}

// single CTA: exclusive prefix sum of tiles_of -> tile_off[0 .. n_groups], n_ptiles = min(total, capacity)
__global__ void __launch_bounds__(1024)
pairtile_scan_kernel(const int32_t* __restrict__ tiles_of, int n_groups, int capacity, int32_t* __restrict__ tile_off,
                     int32_t* __restrict__ n_ptiles) {
  __shared__ long long part[1024];
  const int t = threadIdx.x;
  const int chunk = (n_groups + 1023) / 1024;
  const int g0 = min(n_groups, t * chunk), g1 = min(n_groups, g0 + chunk);
  long long s = 0;
  for (int g = g0; g < g1; ++g) s += tiles_of[g];
  part[t] = s;
  __syncthreads();
  if (t == 0) {
    long long run = 0;
    for (int i = 0; i < 1024; ++i) {
      const long long v = part[i];
      part[i] = run;
      run += v;
    }
    *n_ptiles = (int32_t)(run < capacity ? run : capacity);
    tile_off[n_groups] = (int32_t)(run < 0x7fffffffll ? run : 0x7fffffffll);
  }
  __syncthreads();
  long long run = part[t];
  for (int g = g0; g < g1; ++g) {
    tile_off[g] = (int32_t)(run < 0x7fffffffll ? run : 0x7fffffffll);
    run += tiles_of[g];
  }
}

// one warp per virtual group: compact its part of the bitmap in ascending image order into its tile records
__global__ void __launch_bounds__(PT_THREADS)
pairtile_emit_kernel(const uint32_t* __restrict__ bitmap, int bw, int n_groups, int n_blocks, int block_words, int slots,
                     int slot_rows, const int32_t* __restrict__ tile_off, const int32_t* __restrict__ group_row0,
                     const int32_t* __restrict__ group_cap_lo, const int32_t* __restrict__ region_row,
                     const int32_t* __restrict__ nr, const uint8_t* __restrict__ clamp, int capacity,
                     alad_ptile* __restrict__ ptiles) {
  const int v = blockIdx.x * (PT_THREADS / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (v >= n_groups * n_blocks) return;
  const int b = v / n_groups, g = v - b * n_groups;
  const int t0 = tile_off[v];
  const int n_t = min(tile_off[v + 1], capacity) - t0;
  if (n_t <= 0) return;
  // headers (and zeroed slot fields) first; the slot writers below fill them in
  const int row0 = __ldg(group_row0 + g), cap_lo = __ldg(group_cap_lo + g), cap_hi = __ldg(group_cap_lo + g + 1);
  for (int t = lane; t < n_t; t += 32) {
    alad_ptile* rec = ptiles + t0 + t;
    rec->m_row0 = row0;
    rec->cap_lo = cap_lo;
    rec->cap_hi = cap_hi;
    rec->nseg = 0;
    rec->clamp_bits = 0u;
    for (int s = 0; s < ALAD_PTILE_SLOTS; ++s) {
      rec->slot_row[s] = 0;
      rec->slot_img[s] = 0;
      rec->slot_w[s] = 0;
    }
    rec->reserved = 0;
  }
  __syncwarp();
  const int w_lo = b * block_words, w_hi = min(bw, w_lo + block_words);
  int base = 0;
  for (int w0 = w_lo; w0 < w_hi; w0 += 32) {
    const int w = w0 + lane;
    uint32_t bits = w < w_hi ? __ldg(bitmap + (long long)g * bw + w) : 0u;
    const int n = __popc(bits);
    int incl = n;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    int idx = base + incl - n;
    while (bits) {
      const int bit = __ffs((int)bits) - 1;
      bits &= bits - 1;
      const int l = w * 32 + bit;
      const int t = idx / slots, s_i = idx - t * slots;
      if (t < n_t) {
        alad_ptile* rec = ptiles + t0 + t;
        rec->slot_row[s_i] = __ldg(region_row + l);
        rec->slot_img[s_i] = l;
        rec->slot_w[s_i] = (uint8_t)min(__ldg(nr + l), slot_rows);
        if (clamp && __ldg(clamp + l)) atomicOr(&rec->clamp_bits, 1u << s_i);
        atomicMax(&rec->nseg, s_i + 1);
      }
      ++idx;
    }
    base += __shfl_sync(0xffffffffu, incl, 31);
  }
}

// ---------------------------------------------------------------------------------- consumers
__global__ void __launch_bounds__(PT_THREADS)
gather_list_scores_kernel(const float* __restrict__ S, long long ldS, int Ni, int Nc, const int32_t* __restrict__ ids,
                          long long n_entries, int k, int by_column, int img_off, const int32_t* __restrict__ nr,
                          const int32_t* __restrict__ nw, float* __restrict__ out) {
  const long long e = blockIdx.x * (long long)PT_THREADS + threadIdx.x;
  if (e >= n_entries) return;
  const int q = (int)(e / k);
  const int id = __ldg(ids + e);
  int i, c;
  if (by_column) {
    i = id - img_off;
    c = q;
  } else {
    i = q;
    c = id;
  }
  float v = 0.f;
  if (id >= 0 && i >= 0 && i < Ni && c >= 0 && c < Nc && __ldg(nr + i) > 0 && __ldg(nw + c) > 0)
    v = __ldg(S + (long long)i * ldS + c);
  out[e] = v;
}

constexpr int RR_WARPS = 8;
__global__ void __launch_bounds__(32 * RR_WARPS)
list_rerank_kernel(const float* __restrict__ scores, const int32_t* __restrict__ ids, int Q, int k, int q_off, int gt_mul,
                   int gt_div, int gt_n, const int32_t* __restrict__ fallback, int32_t* __restrict__ rank,
                   int32_t* __restrict__ order) {
  extern __shared__ float2 rr_smem[];                          // [RR_WARPS][k] (score, id bits)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int q = blockIdx.x * RR_WARPS + warp;
  if (q >= Q) return;
  float2* mine = rr_smem + warp * k;
  for (int e = lane; e < k; e += 32) mine[e] = make_float2(__ldg(scores + (long long)q * k + e), __int_as_float(__ldg(ids + (long long)q * k + e)));
  __syncwarp();
  // Order inside the list: score descending, LIST POSITION descending on exact ties -- what numpy.argsort(...)[::-1]
  // of the shortlisted scores gives (alad/evaluation.py:213,305 applied to the shortlist).
  // best ground-truth candidate under that order
  const long long gt_lo = ((long long)q + q_off) * gt_mul / gt_div;
  float gs = -INFINITY;
  int gp = -1;
  for (int e = lane; e < k; e += 32) {
    const int id = __float_as_int(mine[e].y);
    if (id >= gt_lo && id < gt_lo + gt_n && (gp < 0 || ahead(mine[e].x, e, gs, gp))) {
      gs = mine[e].x;
      gp = e;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float os = __shfl_xor_sync(0xffffffffu, gs, o);
    const int op = __shfl_xor_sync(0xffffffffu, gp, o);
    if (op >= 0 && (gp < 0 || ahead(os, op, gs, gp))) {
      gs = os;
      gp = op;
    }
  }
  int cnt = 0, n_valid = 0;
  for (int e = lane; e < k; e += 32) {
    const float v = mine[e].x;
    const int vi = __float_as_int(mine[e].y);
    if (vi < 0) continue;
    ++n_valid;
    if (gp >= 0 && ahead(v, e, gs, gp)) ++cnt;
    if (order) {
      int r = 0;
      for (int o = 0; o < k; ++o) {
        const float2 w = mine[o];                              // broadcast read
        r += (__float_as_int(w.y) >= 0 && ahead(w.x, o, v, e)) ? 1 : 0;
      }
      order[(long long)q * k + r] = vi;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    n_valid += __shfl_xor_sync(0xffffffffu, n_valid, o);
  }
  if (order)
    for (int e = n_valid + lane; e < k; e += 32) order[(long long)q * k + e] = -1;
  if (lane == 0 && rank) rank[q] = gp >= 0 ? cnt : (fallback ? __ldg(fallback + q) : -1);
}

struct PairPlan {
  int bw;
  size_t off_bitmap, off_tiles_of, off_tile_off, bytes;
};
int pair_blocks(int n_loc, int block_images) {
  if (block_images <= 0 || block_images >= n_loc) return 1;
  return (n_loc + block_images - 1) / block_images;
}
PairPlan pair_plan(int n_groups, int n_loc, int block_images) {
  PairPlan p = {};
  auto up = [](size_t x) { return (x + 255) / 256 * 256; };
  p.bw = (n_loc + 31) / 32;
  const size_t n_v = (size_t)n_groups * pair_blocks(n_loc, block_images);
  size_t o = 0;
  p.off_bitmap = o;   o += up(sizeof(uint32_t) * (size_t)n_groups * (size_t)(p.bw > 0 ? p.bw : 1));
  p.off_tiles_of = o; o += up(sizeof(int32_t) * (n_v + 1));
  p.off_tile_off = o; o += up(sizeof(int32_t) * (n_v + 1));
  p.bytes = o + 256;
  return p;
}

}  // namespace
}  // namespace alad

extern "C" int64_t alad_pairtile_workspace_bytes(int32_t n_groups, int32_t n_loc, int32_t block_images) {
  if (n_groups < 0 || n_loc < 0 || block_images < 0 || block_images % 32 != 0) return 0;
  return (int64_t)alad::pair_plan(n_groups, n_loc, block_images).bytes;
}

/* Host helper (HOST pointers, no CUDA work): consecutive captions are grouped greedily into M tiles of <= 128 packed
 * word rows; captions without scored words belong to no group (cap_group = -1).  group_cap_lo gets n_groups + 1
 * entries.  Returns the number of groups or a negative alad_status (a caption with more than 128 scored words). */
extern "C" int alad_caption_groups(const int32_t* nw, int32_t Nc, int32_t* group_row0, int32_t* group_cap_lo,
                                   int32_t* cap_group) {
  using namespace alad;
  ALAD_REQUIRE(Nc >= 0 && (Nc == 0 || (nw && group_row0 && group_cap_lo && cap_group)), "alad_caption_groups: bad arguments");
  int n_g = 0;
  long long row = 0;
  int rows_in = 0;
  bool open = false;
  for (int c = 0; c < Nc; ++c) {
    const int n = nw[c];
    ALAD_REQUIRE(n >= 0, "alad_caption_groups: negative count");
    if (n > ALAD_TILE_M)
      return fail(ALAD_ERR_UNSUPPORTED, "alad_caption_groups: caption %d has %d scored words; the pair-list kernel supports at most %d",
                  c, n, ALAD_TILE_M);
    if (n == 0) {
      cap_group[c] = -1;
      continue;
    }
    if (!open || rows_in + n > ALAD_TILE_M) {
      ALAD_REQUIRE(row < (1ll << 31), "alad_caption_groups: too many word rows");
      group_row0[n_g] = (int32_t)row;
      group_cap_lo[n_g] = c;
      ++n_g;
      rows_in = 0;
      open = true;
    }
    cap_group[c] = n_g - 1;
    rows_in += n;
    row += n;
  }
  if (Nc) group_cap_lo[n_g] = Nc;
  return n_g;
}

extern "C" int alad_pairtile_build(const alad_pairtile_args* a, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(a != nullptr, "alad_pairtile_build: NULL args");
  ALAD_REQUIRE(a->n_groups >= 0 && a->Nc >= 0 && a->n_loc >= 0 && a->k_t2i >= 0 && a->k_i2t >= 0 && a->capacity >= 0,
               "alad_pairtile_build: bad shape");
  ALAD_REQUIRE(a->slot_rows >= ALAD_TILE_N / ALAD_PTILE_SLOTS && a->slot_rows <= ALAD_TILE_N,
               "alad_pairtile_build: slot_rows=%d outside [%d, %d]", a->slot_rows, ALAD_TILE_N / ALAD_PTILE_SLOTS, ALAD_TILE_N);
  ALAD_REQUIRE(a->n_ptiles, "alad_pairtile_build: NULL n_ptiles");
  cudaStream_t st = as_stream(stream);
  if (a->n_groups == 0 || a->n_loc == 0 || a->capacity == 0) {
    ALAD_CUDA(cudaMemsetAsync(a->n_ptiles, 0, sizeof(int32_t), st));
    return ALAD_OK;
  }
  ALAD_REQUIRE(a->block_images >= 0 && a->block_images % 32 == 0, "alad_pairtile_build: block_images must be a multiple of 32");
  const PairPlan pl = pair_plan(a->n_groups, a->n_loc, a->block_images);
  const int n_blocks = pair_blocks(a->n_loc, a->block_images);
  const int block_words = n_blocks == 1 ? pl.bw : a->block_images / 32;
  ALAD_REQUIRE((long long)a->n_groups * n_blocks < (1ll << 31), "alad_pairtile_build: too many (group, block) pairs");
  const int n_v = a->n_groups * n_blocks;
  ALAD_REQUIRE(a->workspace && a->workspace_bytes >= (int64_t)pl.bytes && (reinterpret_cast<uintptr_t>(a->workspace) & 15) == 0,
               "alad_pairtile_build: workspace too small or misaligned");
  ALAD_REQUIRE(a->group_row0 && a->group_cap_lo && a->cap_group && a->region_row && a->nr && a->ptiles,
               "alad_pairtile_build: NULL pointer");
  ALAD_REQUIRE((a->lists_t2i || a->k_t2i == 0) && (a->lists_i2t || a->k_i2t == 0), "alad_pairtile_build: NULL list");
  uint8_t* w = reinterpret_cast<uint8_t*>(a->workspace);
  uint32_t* bitmap = reinterpret_cast<uint32_t*>(w + pl.off_bitmap);
  int32_t* tiles_of = reinterpret_cast<int32_t*>(w + pl.off_tiles_of);
  int32_t* tile_off = reinterpret_cast<int32_t*>(w + pl.off_tile_off);
  const int slots = ALAD_TILE_N / a->slot_rows;
  ALAD_CUDA(cudaMemsetAsync(bitmap, 0, sizeof(uint32_t) * (size_t)a->n_groups * pl.bw, st));
  const long long n1 = (long long)a->Nc * a->k_t2i, n2 = (long long)a->n_loc * a->k_i2t;
  ALAD_REQUIRE((n1 + PT_THREADS - 1) / PT_THREADS < (1ll << 31) && (n2 + PT_THREADS - 1) / PT_THREADS < (1ll << 31),
               "alad_pairtile_build: lists too long");
  if (n1)
    pairtile_mark_t2i_kernel<<<(unsigned)((n1 + PT_THREADS - 1) / PT_THREADS), PT_THREADS, 0, st>>>(
        a->lists_t2i, n1, a->k_t2i, a->cap_group, a->img_off, a->n_loc, a->nr, pl.bw, bitmap);
  if (n2)
    pairtile_mark_i2t_kernel<<<(unsigned)((n2 + PT_THREADS - 1) / PT_THREADS), PT_THREADS, 0, st>>>(
        a->lists_i2t, n2, a->k_i2t, a->cap_group, a->Nc, a->nr, pl.bw, bitmap);
  const unsigned gblocks = (unsigned)((n_v + PT_THREADS / 32 - 1) / (PT_THREADS / 32));
  pairtile_count_kernel<<<gblocks, PT_THREADS, 0, st>>>(bitmap, pl.bw, a->n_groups, n_blocks, block_words, slots, tiles_of);
  pairtile_scan_kernel<<<1, 1024, 0, st>>>(tiles_of, n_v, a->capacity, tile_off, a->n_ptiles);
  pairtile_emit_kernel<<<gblocks, PT_THREADS, 0, st>>>(bitmap, pl.bw, a->n_groups, n_blocks, block_words, slots, a->slot_rows,
                                                      tile_off, a->group_row0, a->group_cap_lo, a->region_row, a->nr, a->clamp,
                                                      a->capacity, a->ptiles);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_gather_list_scores(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, const int32_t* ids, int32_t Q,
                                       int32_t k, int32_t by_column, int32_t img_off, const int32_t* nr, const int32_t* nw,
                                       float* out, void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Ni >= 0 && Nc >= 0 && Q >= 0 && k >= 0 && ldS >= Nc, "alad_gather_list_scores: bad shape");
  const long long n = (long long)Q * k;
  if (n == 0) return ALAD_OK;
  ALAD_REQUIRE(ids && out && nr && nw && (S || Ni == 0 || Nc == 0), "alad_gather_list_scores: NULL pointer");
  ALAD_REQUIRE((n + PT_THREADS - 1) / PT_THREADS < (1ll << 31), "alad_gather_list_scores: lists too long");
  gather_list_scores_kernel<<<(unsigned)((n + PT_THREADS - 1) / PT_THREADS), PT_THREADS, 0, as_stream(stream)>>>(
      S, ldS, Ni, Nc, ids, n, k, by_column, img_off, nr, nw, out);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}

extern "C" int alad_list_rerank(const float* scores, const int32_t* ids, int32_t Q, int32_t k, int32_t q_off, int32_t gt_mul,
                                int32_t gt_div, int32_t gt_n, const int32_t* fallback, int32_t* rank, int32_t* order,
                                void* stream) {
  using namespace alad;
  ALAD_REQUIRE(Q >= 0 && k > 0 && k <= 1024 && gt_mul > 0 && gt_div > 0 && gt_n > 0, "alad_list_rerank: bad arguments");
  if (Q == 0) return ALAD_OK;
  ALAD_REQUIRE(scores && ids && (rank || order), "alad_list_rerank: NULL pointer");
  const size_t smem = (size_t)RR_WARPS * k * sizeof(float2);
  static thread_local size_t smem_set = 0;
  if (smem > 48 * 1024 && smem > smem_set) {
    ALAD_CUDA(cudaFuncSetAttribute(list_rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  list_rerank_kernel<<<(unsigned)((Q + RR_WARPS - 1) / RR_WARPS), 32 * RR_WARPS, smem, as_stream(stream)>>>(
      scores, ids, Q, k, q_off, gt_mul, gt_div, gt_n, fallback, rank, order);
  ALAD_CUDA(cudaGetLastError());
  return ALAD_OK;
}
