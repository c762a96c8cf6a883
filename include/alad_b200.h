/* alad_b200 -- C ABI of the B200 (sm_100a) implementation of ALADIN's all-pairs
 * cross-modal scoring path.
 *
 * The reference (mesnico/ALADIN) is pure Python/PyTorch and has no FFI; its "plugin
 * boundary" for this path is the Python call sites listed next to each entry point.
 * The host side (aladin_b200/*.py) mirrors those call sites and reaches the kernels
 * only through the functions below (ctypes).
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the comment says "host";
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - no allocation inside: the caller owns inputs, outputs and workspaces;
 *   - return 0 on success, negative alad_status otherwise; alad_last_error() gives the
 *     text of the last failure on the calling thread;
 *   - re-entrant across streams; the device is the current CUDA device of the caller.
 */
#ifndef ALAD_B200_H_
#define ALAD_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum alad_status {
  ALAD_OK = 0,
  ALAD_ERR_INVALID = -1,   /* bad argument (shape, alignment, NULL) */
  ALAD_ERR_CUDA = -2,      /* a CUDA runtime / driver call failed    */
  ALAD_ERR_UNSUPPORTED = -3
} alad_status;

/* geometry of the scoring kernel, exposed so the host can build tile tables */
#define ALAD_TILE_M 128      /* packed word rows per tile (TMEM lanes)            */
#define ALAD_TILE_N 240      /* packed region rows per tile (TMEM columns)        */
#define ALAD_TILE_K 64       /* bf16 elements per pipeline stage (128 B rows)     */
#define ALAD_MAX_SEG 32      /* images per N tile                                 */

int alad_abi_version(void);
const char* alad_last_error(void);

/* ---------------------------------------------------------------------------------
 * alad_h2d_2d -- pitched host->device upload (cudaMemcpy2DAsync): only the scored slots of
 * the [N, 71, d] evaluation containers (alad/evaluation.py:98-99) cross PCIe, and the 5x
 * duplicated image rows are read with a 5-row pitch.  Replaces the per-query `.cuda()` of
 * the whole gallery (evaluation.py:179,202,267,291).  src_host may be pinned or pageable.
 * ------------------------------------------------------------------------------- */
int alad_h2d_2d(void* dst, int64_t dst_pitch, const void* src_host, int64_t src_pitch, int64_t width_bytes,
                int64_t height, void* stream);
/* alad_h2d_2d_staged -- the same upload for PAGEABLE sources (what the reference's encode_data returns,
 * alad/evaluation.py:98-130): n_threads host threads gather the rows into a ring of pinned staging buffers while the
 * copy engine drains the previous one (a pageable cudaMemcpy2DAsync is staged by the driver on one thread, ~10 GB/s).
 * Returns once the source has been read; the last DMA may still be in flight on `stream`.  Process-wide staging
 * buffers (3 x 48 MB of pinned memory, allocated on first use). */
int alad_h2d_2d_staged(void* dst, int64_t dst_pitch, const void* src_host, int64_t src_pitch, int64_t width_bytes,
                       int64_t height, int32_t n_threads, void* stream);

/* ---------------------------------------------------------------------------------
 * alad_pack_tokens  -- replaces F.normalize + slot slicing at alad/loss.py:80-90 and the
 * per-query `.cuda()` uploads of alad/evaluation.py:179,202,267,291.
 * For item b and scored token t < count[b] reads src[b, slot0 + t, :], scales it by
 * 1/max(||x||_2, eps) when `normalize`, converts to bf16 and writes packed row
 * row_off[b] + t of dst ([rows, Kp] bf16, K-major, Kp % 64 == 0, zero padded).
 *   mode 0: plain bf16 (Kp >= d)
 *   mode 1: split-precision "word side"   [hi | hi | lo] (Kp >= 3d)
 *   mode 2: split-precision "region side" [hi | lo | hi] (Kp >= 3d)
 *   mode 3: TF32 operands: the normalised fp32 values, Kp / 2 floats per row (Kp >= 2d counts 2-byte units, so
 *           the row pitch is 2 * Kp bytes in every mode)
 * so that <mode1 row, mode2 row> = hi*hi + hi*lo + lo*hi (fp32-grade dot product on the
 * bf16 tensor pipe).  row_item (optional) receives b for every packed row.
 * ------------------------------------------------------------------------------- */
typedef struct alad_pack_args {
  const float* src;          /* [B, S, d] fp32, innermost stride 1                   */
  int64_t stride_b;          /* element strides of src                               */
  int64_t stride_s;
  int32_t B, S, d;
  int32_t slot0;             /* first scored slot (1: loss.py:87-88)                 */
  const int32_t* count;      /* [B] scored tokens per item (<= S - slot0)            */
  const int64_t* row_off;    /* [B] first packed row of item b                       */
  void* dst;                 /* [rows, Kp] bf16                                      */
  int32_t Kp;
  int32_t mode;
  int32_t normalize;         /* 1: F.normalize semantics; 0: copy                    */
  float eps;                 /* 1e-12 (F.normalize) or 0 (alad.utils.l2norm)         */
  int32_t* row_item;         /* optional [rows]: receives item_base + b              */
  int32_t item_base;         /* global index of item 0 (sharded packing)             */
} alad_pack_args;
int alad_pack_tokens(const alad_pack_args* a, void* stream);

/* ---------------------------------------------------------------------------------
 * alad_mrsw_scores_fwd -- replaces alad/loss.py:97-125 (expand + batched matmul + masks +
 * `alignments.max(2)[0].sum(2)`), i.e. the body of AlignmentContrastiveLoss.forward for
 * aggregation 'MrSw'.  S[i, j] = sum_w max_r <region(i,r), word(j,w)> over packed valid
 * tokens; the max starts from 0 when the image has masked slots inside its container
 * (clamp bit) and from -inf otherwise.  TMA + tcgen05/TMEM; the 4-D tensor is never
 * written.  S is zeroed by the call and accumulated with <= 2 addends per entry.
 * ------------------------------------------------------------------------------- */
typedef struct alad_ntile {      /* one 240-column tile of packed region rows (device table) */
  int32_t row_start;             /* first packed region row of the tile                      */
  int32_t img0;                  /* first image (row of S) of the tile; images are consecutive */
  int32_t nseg;                  /* images in the tile (<= ALAD_MAX_SEG)                     */
  uint32_t clamp_bits;           /* bit s: image img0+s has masked slots -> max starts at 0  */
  uint16_t seg[ALAD_MAX_SEG];    /* image s: low byte = first column, high byte = #columns   */
} alad_ntile;

/* Host-side helper (HOST pointers, no CUDA work): greedy grouping of consecutive images into tiles of
 * <= ALAD_TILE_N packed region rows and <= ALAD_MAX_SEG images; images with nr == 0 own no column.  Writes
 * the table to tiles[0 .. return value) (capacity Ni always suffices) and, if row_off != NULL, the first packed
 * region row of every image.  Returns the number of tiles or a negative alad_status.  Replaces the Python
 * mask loops of alad/loss.py:105-106 as the per-call host bookkeeping. */
int alad_region_tiles(const int32_t* nr, const uint8_t* clamp, int32_t Ni, alad_ntile* tiles, int32_t capacity,
                      int64_t* row_off);

typedef struct alad_mrsw_fwd_args {
  const void* words;             /* [n_word_rows, Kp] bf16 packed (mode 0 or 1)              */
  int64_t n_word_rows;
  const void* regions;           /* [n_region_rows, Kp] bf16 packed (mode 0 or 2)            */
  int64_t n_region_rows;
  int32_t Kp;
  const int32_t* row_cap;        /* [ceil(n_word_rows/256)*256] caption of each row, -1 pad  */
  const alad_ntile* ntiles;
  int32_t n_ntiles;
  float* S;                      /* [Ni, ldS] fp32                                           */
  int64_t ldS;
  int32_t Ni, Nc;
  int32_t epilogue;              /* 0 = MrSw pooling; 1 = plain GEMM: S[n, m] = <region n, word m> */
  int32_t num_ctas;              /* 0 = one persistent CTA per SM                            */
  int32_t cta_group;             /* 0 = default (2, or $ALAD_CTA_GROUP); 1 = one CTA per 128x240 tile;
                                    2 = CTA pair (tcgen05 cta_group::2) per 256x240 tile       */
  int32_t transpose_out;         /* 1: write S[word item, region item] (S is [Nc, ldS]); used for the
                                    'MwSr' pooling, which is MrSw with the two token sets swapped */
  int32_t accumulate;            /* epilogue 0 only.  0: S is zeroed here first; 1: the launch ADDS its partial sums onto S
                                    (the caller zeroed it): several launches -- also from other GPUs, through a peer
                                    mapping of S -- each score a range of word rows of the same block (ABI >= 4) */
  int32_t operand_format;        /* 0 = bf16 rows (pack modes 0-2), 1 = fp32 rows read as TF32 (pack mode 3): tcgen05.mma
                                    kind::tf32, half the tensor rate of bf16 on twice the bytes: an intermediate precision (worst
                                    score entry 3e-4 .. 6e-4 relative; the 1e-4 parity mode is the 3-term split-precision
                                    bf16 one) at half that mode's time (ABI >= 5; append-only: older callers pass 0)    */
} alad_mrsw_fwd_args;
int alad_mrsw_scores_fwd(const alad_mrsw_fwd_args* a, void* stream);

/* ---------------------------------------------------------------------------------
 * alad_scores_fused -- "pack both operands + score" in ONE call on RAW fp32 tokens and HOST count arrays:
 * the body of AlignmentContrastiveLoss.forward up to `aggr_similarity` (alad/loss.py:80-125, epilogue 0) or of
 * dot_sim / cosine_sim (alad/loss.py:8-18, epilogue 1: one row per item, slot0 = 0).  The "max" items
 * become tile columns (images for 'MrSw'), the "sum" items packed rows (captions).  Host work done inside:
 * row offsets, the greedy region tile table (alad_region_tiles), one metadata upload; device work: the two
 * alad_pack_tokens launches and alad_mrsw_scores_fwd.  The packed operands live in `workspace`
 * (alad_scores_fused_workspace_bytes: an upper bound from the container extents; 256-byte aligned).
 * ------------------------------------------------------------------------------- */
typedef struct alad_scores_fused_args {
  const float* max_x;            /* [n_max, S_max, d] fp32 device, innermost stride 1         */
  int64_t max_stride_b, max_stride_s;
  const float* sum_x;            /* [n_sum, S_sum, d]                                         */
  int64_t sum_stride_b, sum_stride_s;
  int32_t n_max, S_max, slot0_max;
  int32_t n_sum, S_sum, slot0_sum;
  int32_t d;
  const int32_t* max_count;      /* HOST [n_max] scored tokens per item (from slot0 on)       */
  const int32_t* sum_count;      /* HOST [n_sum]                                              */
  const uint8_t* max_clamp;      /* HOST [n_max] or NULL: item has masked slots -> max starts at 0 */
  int32_t precision;             /* 0 = bf16 operands, 1 = split hi/lo (fp32-grade)           */
  int32_t epilogue;              /* 0 = max/sum pooling, 1 = plain GEMM                       */
  int32_t normalize;             /* 1: F.normalize the tokens while packing                   */
  float eps;
  float* S;                      /* [n_max, ldS] (or [n_sum, ldS] when transpose_out)         */
  int64_t ldS;
  int32_t transpose_out;
  void* workspace;
  int64_t workspace_bytes;
} alad_scores_fused_args;
int64_t alad_scores_fused_workspace_bytes(int32_t n_max, int32_t S_max, int32_t slot0_max, int32_t n_sum, int32_t S_sum,
                                          int32_t slot0_sum, int32_t d, int32_t precision);
int alad_scores_fused(const alad_scores_fused_args* a, void* stream);

/* ---------------------------------------------------------------------------------
 * alad_pool_tokens -- sum of the L2-normalised valid tokens of every item: out[b, :] =
 * sum_{t < count[b]} normalize(src[b, slot0 + t, :]).  With it the 'sum' / 'mean' pooling modes of
 * alad/loss.py:120-123 collapse to one GEMM: sum_{r,w} <r, w> = <sum_r r, sum_w w>.
 * alad_scale_scores -- S[i, j] = S[i, j] * mul / col_div[j] (col_div optional): the division by
 * the caption lengths of 'MrAVGw' (loss.py:126-129) and the 1/(R*W) of 'mean'.
 * ------------------------------------------------------------------------------- */
int alad_pool_tokens(const float* src, int64_t stride_b, int64_t stride_s, int32_t B, int32_t S, int32_t d,
                     int32_t slot0, const int32_t* count, float eps, float* out, void* stream);
int alad_scale_scores(float* S, int64_t ldS, int32_t Ni, int32_t Nc, const float* col_div, float mul, void* stream);

/* ---------------------------------------------------------------------------------
 * alad_mrsw_scores_bwd -- autograd of alad/loss.py:80-125 for aggregation 'MrSw': given
 * dL/dS = g0_scale * G0 + G1 (either may be NULL; g0_scale is a DEVICE scalar or NULL = 1)
 * writes dL/d im_set [Bi,S_im,d] and dL/d s_seq [Bc,S_s,d] (contiguous fp32, zeroed inside).
 * Non-zero entries of dL/dS are compacted into a pair list and only those (image, caption)
 * tiles are recomputed (<= 3B pairs with hardest-negative mining).
 * ------------------------------------------------------------------------------- */
typedef struct alad_mrsw_bwd_args {
  const float* im;               /* raw image tokens [Bi, S_im, d], innermost stride 1       */
  int64_t im_stride_b, im_stride_s;
  const float* s;                /* raw caption tokens [Bc, S_s, d]                          */
  int64_t s_stride_b, s_stride_s;
  int32_t Bi, S_im, Bc, S_s, d;
  const int32_t* nr;             /* [Bi] valid scored regions (slots 1 .. nr)                */
  const int32_t* nw;             /* [Bc] valid scored words                                  */
  const float* G0;               /* optional [Bi, ldG0]                                      */
  int64_t ldG0;
  const float* g0_scale;         /* optional device scalar multiplying G0                    */
  const float* G1;               /* optional [Bi, ldG1]                                      */
  int64_t ldG1;
  float* d_im;                   /* [Bi, S_im, d] gradient, zeroed here (all Bi*S_im*d floats from d_im on:    */
  float* d_s;                    /* contiguous, or a dense permutation of the item / slot dims -- strides below) */
  float eps;                     /* F.normalize eps (1e-12)                                  */
  int32_t region_extent;         /* container extent R of the max side (clamp iff nr < R); 0 = S_im - 1 */
  int64_t max_pairs;             /* capacity of the pair list (Bi*Bc is always enough)       */
  void* workspace;
  int64_t workspace_bytes;       /* >= alad_mrsw_bwd_workspace_bytes(...)                    */
  int64_t d_im_stride_b, d_im_stride_s;   /* element strides of d_im / d_s over items and slots; 0 = contiguous.   */
  int64_t d_s_stride_b, d_s_stride_s;     /* e.g. (d, Bi*d) for the [S, B, d] layout of alad_model.py:377-378      */
} alad_mrsw_bwd_args;
int64_t alad_mrsw_bwd_workspace_bytes(int32_t Bi, int32_t S_im, int32_t Bc, int32_t S_s, int64_t max_pairs);
int alad_mrsw_scores_bwd(const alad_mrsw_bwd_args* a, void* stream);

/* ---------------------------------------------------------------------------------
 * B x B losses, forward + gradient in one call (workspace: alad_loss_workspace_bytes(B)).
 * alad_triplet_fwd_bwd -- Contrastive.compute_contrastive_loss, alad/loss.py:42-67:
 *   loss = sum_i max_j [m + S_ij - S_ii]_+ + sum_j max_i [m + S_ij - S_jj]_+   (max_violation)
 *   or the plain sums; G (optional) = dloss/dS dense; row_arg/col_arg = hardest negative of
 *   every row / column (-1 when nothing violates; violation counts in sum mode).
 * alad_listnet_fwd_bwd -- DistillationLoss(mode='listnet'), alad/loss.py:427-445: teacher is
 *   detached (loss.py:370); dM (optional) = dloss/dstudent.
 * ------------------------------------------------------------------------------- */
int64_t alad_loss_workspace_bytes(int32_t B);
int alad_triplet_fwd_bwd(const float* S, int64_t ldS, int32_t B, float margin, int32_t max_violation, float* loss,
                         float* G, int64_t ldG, int32_t* row_arg, int32_t* col_arg, void* workspace, void* stream);
int alad_listnet_fwd_bwd(const float* teacher, int64_t ldT, const float* student, int64_t ldM, int32_t B,
                         float temperature, float eps, float* loss, float* dM, int64_t ldG, void* workspace,
                         void* stream);

/* ---------------------------------------------------------------------------------
 * alad_train_losses_fwd / alad_train_losses_bwd -- the loss stack of ALADModel.forward_loss
 * (alad/alad_model.py:371-428: matching_criterion :380, alignment_criterion :386, distillation_loss :405) in ONE
 * native call per direction, for the configuration every shipped config selects (measure 'dot', aggregation
 * 'MrSw', distillation 'listnet').  Pure composition of alad_scores_fused, alad_triplet_fwd_bwd,
 * alad_listnet_fwd_bwd and alad_mrsw_scores_bwd on the caller's stream (plus a scaled-sum/transpose helper
 * kernel for the matching head's backward GEMM operands): at B <= 512 the step is bound by the host cost of
 * its ~35 launches, not by the device.
 *   fwd: M = im_cls @ s_cls.T, losses[0] = hinge(M); S = MrSw(im_set, s_seq), losses[1] = hinge(S);
 *        losses[2] = ListNet(teacher = S, student = M) (0 when !with_distill); with want_grad the dense
 *        gradients G_m = dlosses[0]/dM, G_a = dlosses[1]/dS, dM = dlosses[2]/dM are kept for the backward call.
 *   bwd: g = DEVICE [3] upstream gradients of the three losses (has_g_* = 0: that loss got no gradient);
 *        d im_cls = (g0 G_m + g2 dM) @ s_cls, d s_cls = (g0 G_m + g2 dM).T @ im_cls (split-precision GEMMs),
 *        (d im_set, d s_seq) = sparse MrSw backward of g1 G_a.  Output pointers may be NULL (not needed).
 * nr / nw / clamp are HOST arrays (valid scored counts per image / caption, "has masked slots" flags);
 * M, S, G_m, G_a, dM are contiguous [B, B]; d im_cls / d s_cls contiguous [B, d]; workspace 256-byte aligned,
 * alad_train_losses_workspace_bytes(..., backward = 0 | 1).
 * ------------------------------------------------------------------------------- */
typedef struct alad_train_losses_args {
  const float* im_cls;  int64_t ld_im_cls;       /* [B, d] matching-head vectors (normalised by the model) */
  const float* s_cls;   int64_t ld_s_cls;
  const float* im_set;  int64_t im_stride_b, im_stride_s;   /* [B, S_im, d] raw tokens, innermost stride 1 */
  const float* s_seq;   int64_t s_stride_b, s_stride_s;     /* [B, S_s, d]                                 */
  int32_t B, S_im, S_s, d;
  const int32_t* nr;             /* HOST [B] */
  const int32_t* nw;             /* HOST [B] */
  const uint8_t* clamp;          /* HOST [B] or NULL */
  int32_t precision;             /* alignment score kernel: 0 = bf16 operands, 1 = split hi/lo */
  int32_t precision_m;           /* matching GEMM (forward) likewise */
  float margin_m;  int32_t max_violation_m;      /* matching_criterion  */
  float margin_a;  int32_t max_violation_a;      /* alignment_criterion */
  int32_t with_distill;
  float temperature, listnet_eps;                /* 6.0, 1e-10 (alad/loss.py:430-443) */
  int32_t want_grad;
  float* losses;                 /* [3] matching, alignment, distillation */
  float* M;  float* S;           /* [B, B] score matrices */
  float* G_m;  float* G_a;  float* dM;
  const float* g;                /* backward: [3] upstream gradients */
  int32_t has_g_m, has_g_a, has_g_d;
  float* d_im_cls;  float* d_s_cls;
  float* d_im_set;  int64_t d_im_stride_b, d_im_stride_s;   /* element strides, 0 = contiguous [B, S, d] */
  float* d_s_seq;   int64_t d_s_stride_b, d_s_stride_s;
  void* workspace;
  int64_t workspace_bytes;
} alad_train_losses_args;
int64_t alad_train_losses_workspace_bytes(int32_t B, int32_t S_im, int32_t S_s, int32_t d, int32_t precision,
                                          int32_t backward);
int alad_train_losses_fwd(const alad_train_losses_args* a, void* stream);
int alad_train_losses_bwd(const alad_train_losses_args* a, void* stream);

/* ---------------------------------------------------------------------------------
 * The remaining DistillationLoss modes (alad/loss.py:371-425), forward + gradient in one call;
 * workspace: alad_distill_workspace_bytes(B, mode) with mode 0 = mse, 1 = contrastive, 2 = ordinal.
 * alad_distill_mse_fwd_bwd -- 'mse' (loss.py:371-373): loss = mean((student*wb[0] + wb[1] - teacher)^2);
 *   wb is the module's DEVICE parameter [2]; dM (optional) = dloss/dstudent, dwb (optional) = dloss/dwb.
 * alad_distill_contrastive_fwd_bwd -- 'contrastive' (loss.py:397-418): hinge on the student with the
 *   hard negatives picked by the teacher (arg-max per row / column with the diagonal zeroed, first
 *   occurrence); like the reference, WHOLE columns / rows of the un-cleared hinge matrices are
 *   selected and summed.  zero_teacher_diag = 1 reproduces the reference's in-place
 *   `teacher.masked_fill_(eye, 0)` side effect on the caller's matrix (loss.py:400).
 * alad_distill_ordinal_fwd_bwd -- 'ordinal' (loss.py:374-396): per row and per column, the student
 *   reordered by the ascending (stable) teacher order must keep that order by `margin` between
 *   entries `stride` apart, counted where the later teacher value >= threshold; each direction is a
 *   mean over its selected pairs (NaN when none, like torch).  B <= 16384.
 * ------------------------------------------------------------------------------- */
int64_t alad_distill_workspace_bytes(int32_t B, int32_t mode);
int alad_distill_mse_fwd_bwd(const float* teacher, int64_t ldT, const float* student, int64_t ldM, int32_t B,
                             const float* wb, float* loss, float* dM, int64_t ldG, float* dwb, void* workspace,
                             void* stream);
int alad_distill_contrastive_fwd_bwd(float* teacher, int64_t ldT, const float* student, int64_t ldM, int32_t B,
                                     float margin, int32_t zero_teacher_diag, float* loss, float* dM, int64_t ldG,
                                     void* workspace, void* stream);
int alad_distill_ordinal_fwd_bwd(const float* teacher, int64_t ldT, const float* student, int64_t ldM, int32_t B,
                                 float margin, float threshold, int32_t stride, float* loss, float* dM, int64_t ldG,
                                 void* workspace, void* stream);

/* ---------------------------------------------------------------------------------
 * Similarity measures other than the dot product, and gradient helpers.
 * alad_order_scores(_bwd) -- order_sim, alad/loss.py:20-26: scores[i, j] = -||max(0, s_j - im_i)||_2
 *   (d_im [Ni, d], d_s [Nc, d] contiguous; entries with a zero score get no gradient).
 * alad_normalize_bwd -- in place dx <- J(x) dx, J = Jacobian of x / max(||x||, eps) per row: with the two
 *   GEMMs of dot_sim it is the gradient of cosine_sim (loss.py:13-18; eps = 0 like alad.utils.l2norm).
 * alad_pool_tokens_bwd -- gradient of alad_pool_tokens: d_src[b, slot, :] = J(src[b, slot]) d_pool[b, :]
 *   for slot0 <= slot < slot0 + count[b], 0 elsewhere (d_src contiguous [B, S, d]); serves the
 *   'sum' / 'mean' aggregations of loss.py:120-123.
 * ------------------------------------------------------------------------------- */
int alad_order_scores(const float* im, int64_t ld_im, const float* s, int64_t ld_s, int32_t Ni, int32_t Nc, int32_t d,
                      float* scores, int64_t ldS, void* stream);
int alad_order_scores_bwd(const float* im, int64_t ld_im, const float* s, int64_t ld_s, int32_t Ni, int32_t Nc,
                          int32_t d, const float* scores, int64_t ldS, const float* G, int64_t ldG, float* d_im,
                          float* d_s, void* stream);
int alad_normalize_bwd(const float* x, int64_t ld_x, int64_t rows, int32_t d, float eps, float* dx, int64_t ld_dx,
                       void* stream);
int alad_pool_tokens_bwd(const float* src, int64_t stride_b, int64_t stride_s, int32_t B, int32_t S, int32_t d,
                         int32_t slot0, const int32_t* count, float eps, const float* d_pool, float* d_src,
                         void* stream);

/* ---------------------------------------------------------------------------------
 * Aggregation 'scan-sentences' of AlignmentContrastiveLoss.forward (alad/loss.py:136-149): relu of the masked
 * region x word cosines, L2 normalisation over the regions, softmax over the unmasked words, attended word
 * vector per region, cosine of region and attended vector, sum over the unmasked regions.  The reference
 * materialises B x B x R x W x d tensors (loss.py:143-146); here <x_r, att_r> = sum_w alpha[r,w] C[r,w] and
 * ||att_r||^2 = alpha_r' K_j alpha_r, so a pair needs only its cosine block C (a plain GEMM of the unit token rows
 * on the tcgen05 kernel: alad_scores_fused, epilogue 1) and the Gram matrix K_j of the caption's unit word rows.
 *   yh  [Bc*W, d]   unit word rows, caption j = rows j*W .. j*W+W-1 (slots 1 .. W of the caption container)
 *   C   [Bi*R, ldC] cosines: row i*R + r, column j*W + w
 *   nr / nw         DEVICE valid scored counts per image / caption; max_nr / max_nw = host upper bounds of them
 *                   (shared-memory extents; at most 128 each)
 * alad_scan_gram      K [Bc, W, W] (entries outside the valid nw x nw block are 0)
 * alad_scan_pool_fwd  S[i, j]; 0 for an image without valid regions, NaN for a caption without valid words
 *                     (softmax over an all -inf row, loss.py:139-140), like the reference
 * alad_scan_pool_bwd  dC [Bi*R, lddC] = dL/dC for dL/dS = G (zeroed inside, dense: feeds the two backward GEMMs)
 *                     and dK [Bc, W, W] += dL/dK (the CALLER zeroes dK before the first image chunk); pairs with
 *                     G = 0 are skipped.  Masked regions get a zero gradient (the reference yields NaN there).
 * alad_scan_gram_bwd  d_yh [Bc*W, d] += 2 dK yh (dK is symmetric by construction)
 * alad_scan_apply_pairs  for a DEVICE list of n_pairs (image, caption) index pairs (int32 [n_pairs, 2], each pair once):
 *                     d_xh [Bi*R, d] += dC_ij yh_j and d_yh [Bc*W, d] += dC_ij' xh_i (float atomics; the caller zeroes
 *                     both) -- the sparse form of the two backward GEMMs when dL/dS has few non-zero entries
 *                     (<= 3B with the hardest-negative hinge)
 * ------------------------------------------------------------------------------- */
int alad_scan_gram(const float* yh, int32_t Bc, int32_t W, int32_t d, const int32_t* nw, float* K, void* stream);
int alad_scan_gram_bwd(const float* yh, int32_t Bc, int32_t W, int32_t d, const int32_t* nw, const float* dK,
                       float* d_yh, void* stream);
int alad_scan_pool_fwd(const float* C, int64_t ldC, int32_t Bi, int32_t R, int32_t Bc, int32_t W, const int32_t* nr,
                       const int32_t* nw, int32_t max_nr, int32_t max_nw, const float* K, float* S, int64_t ldS,
                       void* stream);
int alad_scan_apply_pairs(const float* dC, int64_t lddC, const float* xh, const float* yh, const int32_t* pairs,
                          int32_t n_pairs, int32_t Bi, int32_t R, int32_t Bc, int32_t W, int32_t d, const int32_t* nr,
                          const int32_t* nw, int32_t max_nr, int32_t max_nw, float* d_xh, float* d_yh, void* stream);
int alad_scan_pool_bwd(const float* C, int64_t ldC, int32_t Bi, int32_t R, int32_t Bc, int32_t W, const int32_t* nr,
                       const int32_t* nw, int32_t max_nr, int32_t max_nw, const float* K, const float* G, int64_t ldG,
                       float* dC, int64_t lddC, float* dK, void* stream);

/* ---------------------------------------------------------------------------------
 * Ranking -- replaces numpy.argsort + numpy.where of alad/evaluation.py:213-223,303-308
 * and alad/recall_auxiliary.py:34-56.  Order = score descending, index descending on
 * exact ties (= stable argsort reversed).  S is image-major [Ni, ldS]; ground truth of
 * image i (global index img_off + i) are captions 5*(img_off+i) .. +4 (group = 5).
 * ------------------------------------------------------------------------------- */
/* i2t: rank[i] = position of the best ground-truth caption; top1[i] = arg-max caption. */
int alad_rank_rows(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, int32_t group, int32_t img_off,
                   int32_t* rank, int32_t* top1, void* stream);
/* t2i step 1: gt[c] = S[c/group - img_off, c] where this shard owns that image, else untouched. */
int alad_col_gt(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, int32_t group, int32_t img_off,
                float* gt, void* stream);
/* t2i step 2: count[c] = #{ local images ahead of the ground truth of caption c }. */
int alad_col_count(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, int32_t group, int32_t img_off,
                   const float* gt, int32_t* count, void* stream);
/* t2i step 3: per-caption top-k candidates of this shard, `splits` row slices each:
 * cand_score/cand_idx are [splits, Nc, k], sorted, idx = global image index (-1 = empty). */
int alad_col_topk(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, int32_t k, int32_t img_off, int32_t splits,
                  float* cand_score, int32_t* cand_idx, void* stream);
/* t2i step 3, threshold-select variant (the default of the drop-ins): exact, sorted top-k of every caption
 * over this shard's images in two sweeps of S -- group maxima -> per-caption threshold = k-th largest group
 * maximum -> candidates >= threshold -> rank among the candidates; captions whose candidate list overflows
 * (mass ties) fall back to the heaps of alad_col_topk.  out_score/out_idx are [Nc, k], idx = global image
 * index (-1 = fewer than k images).  Same order as numpy.argsort(...)[::-1][:k] on a stable sort
 * (alad/evaluation.py:303-308).  Workspace: alad_col_topk_select_workspace_bytes, 16-byte aligned. */
int64_t alad_col_topk_select_workspace_bytes(int32_t Ni, int32_t Nc, int32_t k);
int alad_col_topk_select(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, int32_t k, int32_t img_off,
                         float* out_score, int32_t* out_idx, void* workspace, void* stream);
/* Both directions of one score block in TWO sweeps of S instead of four (replaces alad_rank_rows + alad_col_count +
 * alad_col_topk_select on the same block; alad/evaluation.py:213-223 and :303-308 in one pass): the first sweep carries
 * the i2t counts / arg-max of rows [0, q_rows), the per-caption group maxima of the threshold select and -- when
 * `count` is given -- the t2i "images ahead of the ground truth" counts of captions [0, q_cols); the second sweep
 * collects the top-k candidates.  rank/top1 are [q_rows] (i2t gallery = all Nc captions), count [q_cols] or NULL,
 * out_score/out_idx [q_cols, k] (t2i gallery = all Ni rows of the block).  gt [q_cols]: the captions' ground-truth
 * scores (after the exchange between shards); gt == NULL with count != NULL takes them from the block itself (single
 * block).  Small or unaligned blocks run the one-purpose kernels.  Results are identical to the separate entry points.
 * Workspace: alad_rank_fused_workspace_bytes, 16-byte aligned. */
int64_t alad_rank_fused_workspace_bytes(int32_t Ni, int32_t q_rows, int32_t q_cols, int32_t k);
int alad_rank_fused(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, int32_t group, int32_t img_off, int32_t q_rows,
                    int32_t q_cols, int32_t k, const float* gt, int32_t* rank, int32_t* top1, int32_t* count,
                    float* out_score, int32_t* out_idx, void* workspace, void* stream);
/* ---------------------------------------------------------------------------------
 * alad_mrsw_retrieval -- i2t + t2i of alad/evaluation.py:158-327 from PACKED operands in one call, the score matrix
 * OPTIONAL: a composition of alad_region_tiles, alad_mrsw_scores_fwd, alad_col_gt, alad_rank_fused and
 * alad_topk_merge.  S != NULL: the whole [Ni, ldS] matrix is written and ranked.  S == NULL: the images are scored in blocks
 * of max(block_images, k) rows into one reusable buffer of the workspace -- a first pass over the block diagonal gives the
 * captions' ground-truth scores (same 256-row word units as the full pass: bit-identical entries), a second pass ranks every
 * block against all captions, sums the t2i counts and merges the per-caption top-k lists.  Both give the ranking of the
 * dense matrix exactly.  Image i owns captions group*i .. group*i + group-1.
 * ------------------------------------------------------------------------------- */
typedef struct alad_mrsw_retrieval_args {
  const void* words;             /* packed caption rows [n_word_rows, Kp] (alad_pack_tokens), device          */
  int64_t n_word_rows;
  const int32_t* row_cap;        /* device [ceil(n_word_rows/256)*256]: caption of every packed row, -1 padding */
  const void* regions;           /* packed region rows [n_region_rows, Kp], device                            */
  int64_t n_region_rows;
  int32_t Kp;
  int32_t operand_format;        /* as alad_mrsw_fwd_args: 0 = bf16 rows, 1 = fp32 rows read as TF32           */
  const int32_t* nr;             /* HOST [Ni]: packed region rows of every image                              */
  const uint8_t* clamp;          /* HOST [Ni] or NULL: image has masked region slots                          */
  const int64_t* cap_row;        /* HOST [Nc + 1]: first packed word row of every caption, cap_row[Nc] = n_word_rows */
  int32_t Ni, Nc, group, k, block_images;
  float* S;                      /* optional [Ni, ldS]                                                        */
  int64_t ldS;
  int32_t* rank_i2t;             /* [Ni]  position of the image's best ground-truth caption                   */
  int32_t* top1;                 /* [Ni]  arg-max caption                                                     */
  int32_t* rank_t2i;             /* [Nc]  images ahead of the caption's ground-truth image                    */
  float* topk_score;             /* [Nc, k]                                                                   */
  int32_t* topk_idx;             /* [Nc, k] image indices, -1 = fewer than k images                           */
  void* workspace;               /* alad_mrsw_retrieval_workspace_bytes, 256-byte aligned                     */
  int64_t workspace_bytes;
} alad_mrsw_retrieval_args;
int64_t alad_mrsw_retrieval_workspace_bytes(int32_t Ni, int32_t Nc, int32_t k, int32_t block_images, int32_t with_S);
int alad_mrsw_retrieval(const alad_mrsw_retrieval_args* a, void* stream);

/* merge P sorted candidate lists per caption (after the all-gather across shards). */
int alad_topk_merge(const float* cand_score, const int32_t* cand_idx, int32_t P, int32_t Nc, int32_t k,
                    float* out_score, int32_t* out_idx, void* stream);

/* ---------------------------------------------------------------------------------
 * Pair-list scoring (BASELINE config 5, stage 2 of two-stage retrieval): MrSw scores of a SPARSE set of
 * (image, caption) pairs on the same tcgen05 mainloop as alad_mrsw_scores_fwd.  The reference pieces composed are
 * the matching-head shortlist of alad/recall_auxiliary.py:30 and the alignment scores of alad/loss.py:97-125.
 *
 * A "pair tile" is 128 packed word rows starting at the first row of caption cap_lo (captions cap_lo .. cap_hi-1
 * lie completely inside it; rows of later captions are ignored) times up to ALAD_PTILE_SLOTS image SLOTS of
 * slot_rows packed region rows each: slot s is loaded by its own TMA box from packed region row slot_row[s], so the
 * images of a tile need not be neighbours in the gallery.  Every listed (image, caption) pair is produced by exactly
 * one tile and written with a plain store to S[slot_img, caption]; S is NOT zeroed and entries of unlisted pairs are
 * left untouched.
 *
 * alad_pairtile_build (device-side bookkeeping, no host sync): the image set of a tile group = union over its
 * captions of their t2i shortlist (lists_t2i[c, :], GLOBAL image ids, -1 = empty; ids outside
 * [img_off, img_off + n_loc) are another shard's) and of the local images that shortlist the caption
 * (lists_i2t[i, :], caption ids).  The union is a bitmap per group (workspace), compacted in ascending image order
 * into tiles of 240 / slot_rows slots; n_ptiles (DEVICE int32) receives the tile count, clipped to `capacity`
 * (capacity >= (entries of both lists) / slots + n_groups * n_blocks always suffices).
 * ------------------------------------------------------------------------------- */
#define ALAD_PTILE_SLOTS 8
typedef struct alad_ptile {      /* device table entry, 96 bytes                                             */
  int32_t m_row0;                /* first packed word row of the tile                                        */
  int32_t cap_lo, cap_hi;        /* captions (row_cap values) whose scores this tile emits                   */
  int32_t nseg;                  /* image slots in use                                                       */
  uint32_t clamp_bits;           /* bit s: slot s's image has masked region slots -> max starts at 0         */
  int32_t slot_row[ALAD_PTILE_SLOTS];   /* first packed region row of slot s                                 */
  int32_t slot_img[ALAD_PTILE_SLOTS];   /* LOCAL image index (row of S) of slot s                            */
  uint8_t slot_w[ALAD_PTILE_SLOTS];     /* scored regions of slot s (<= slot_rows)                           */
  int32_t reserved;
} alad_ptile;

typedef struct alad_pairtile_args {
  int32_t n_groups;              /* caption groups = M tiles (host-built: consecutive captions, <= 128 word rows) */
  const int32_t* group_row0;     /* [n_groups] first packed word row                                          */
  const int32_t* group_cap_lo;   /* [n_groups + 1] first caption of every group (last entry = Nc)             */
  const int32_t* cap_group;      /* [Nc] group of every caption                                               */
  int32_t Nc;
  const int32_t* lists_t2i;      /* [Nc, k_t2i] or NULL                                                       */
  int32_t k_t2i;
  const int32_t* lists_i2t;      /* [n_loc, k_i2t] or NULL                                                    */
  int32_t k_i2t;
  int32_t img_off, n_loc;        /* this shard's image block                                                  */
  const int32_t* region_row;     /* [n_loc] first packed region row of every local image                      */
  const int32_t* nr;             /* [n_loc] scored regions                                                    */
  const uint8_t* clamp;          /* [n_loc] or NULL                                                           */
  int32_t slot_rows;             /* >= max nr, <= 240                                                         */
  int32_t block_images;          /* 0 = one block; else a multiple of 32: the local images are processed in blocks
                                    of this many (tiles are emitted block-major) so that a block's packed region
                                    rows stay L2-resident while the caption groups stream past                */
  alad_ptile* ptiles;            /* [capacity] out                                                            */
  int32_t capacity;
  int32_t* n_ptiles;             /* DEVICE scalar out                                                         */
  void* workspace;
  int64_t workspace_bytes;       /* >= alad_pairtile_workspace_bytes(n_groups, n_loc, block_images)           */
} alad_pairtile_args;
/* Host helper (HOST pointers, no CUDA work): greedy grouping of consecutive captions into M tiles of <= ALAD_TILE_M
 * packed word rows.  group_row0 [>= Nc], group_cap_lo [>= Nc + 1] (n_groups + 1 entries are written), cap_group [Nc]
 * (-1 for a caption without scored words).  Returns n_groups or a negative alad_status. */
int alad_caption_groups(const int32_t* nw, int32_t Nc, int32_t* group_row0, int32_t* group_cap_lo, int32_t* cap_group);
int64_t alad_pairtile_workspace_bytes(int32_t n_groups, int32_t n_loc, int32_t block_images);
int alad_pairtile_build(const alad_pairtile_args* a, void* stream);

typedef struct alad_mrsw_pairs_args {
  const void* words;             /* [n_word_rows, Kp] bf16 packed (mode 0 or 1)                              */
  int64_t n_word_rows;
  const void* regions;           /* [n_region_rows, Kp] bf16 packed (mode 0 or 2)                            */
  int64_t n_region_rows;
  int32_t Kp;
  const int32_t* row_cap;        /* [n_word_rows] caption of each packed word row                            */
  const alad_ptile* ptiles;      /* device table                                                             */
  const int32_t* n_ptiles;       /* DEVICE scalar: tiles in the table                                        */
  int32_t max_ptiles;            /* host upper bound (sizes the grid)                                        */
  int32_t slot_rows;
  float* S;                      /* [Ni, ldS] (or [Nc, ldS] when transpose_out); not zeroed                   */
  int64_t ldS;
  int32_t Ni, Nc;
  int32_t transpose_out;
  int32_t num_ctas;              /* 0 = one persistent CTA per SM                                            */
  int32_t word_box_rows;         /* word rows loaded per tile: 0 = 128, else a multiple of 8 >= the longest caption
                                    group (e.g. 104 for two 50-word captions): the rows a tile never scores stay out of
                                    the L2 -> shared-memory traffic that bounds this kernel                     */
} alad_mrsw_pairs_args;
int alad_mrsw_scores_pairs(const alad_mrsw_pairs_args* a, void* stream);

/* Re-ranking of per-query candidate lists (stage 2 consumers).
 * alad_gather_list_scores: out[q, k] = S entry of (query q, candidate ids[q, k]); by_column = 1: q is a caption
 *   (column of S), ids are GLOBAL image ids -- candidates outside this shard's block [img_off, img_off + Ni) or with
 *   no scored token on either side give 0, so the per-shard outputs add up (all-reduce) to the full list;
 *   by_column = 0: q is a LOCAL image (row of S), ids are caption ids.
 * alad_list_rerank: order[q, :] = candidate ids by (score desc, list position desc on exact ties, i.e.
 *   numpy.argsort(scores)[::-1] of the shortlisted scores); rank[q] = position of the best
 *   ground-truth candidate (ids gt_lo .. gt_lo + gt_n - 1 with gt_lo = (q + q_off) * gt_mul / gt_div) inside that
 *   order, or fallback[q] when no ground-truth id is in the list (it keeps its stage-1 rank). */
int alad_gather_list_scores(const float* S, int64_t ldS, int32_t Ni, int32_t Nc, const int32_t* ids, int32_t Q, int32_t k,
                            int32_t by_column, int32_t img_off, const int32_t* nr, const int32_t* nw, float* out,
                            void* stream);
int alad_list_rerank(const float* scores, const int32_t* ids, int32_t Q, int32_t k, int32_t q_off, int32_t gt_mul,
                     int32_t gt_div, int32_t gt_n, const int32_t* fallback, int32_t* rank, int32_t* order, void* stream);

/* Two-stage retrieval (BASELINE config 5; not in the reference): S2 = -inf everywhere except the
 * shortlisted (image, caption) pairs, which keep their alignment score.  idx is [n_lists, k];
 * by_column = 1: list q belongs to caption q and holds GLOBAL image indices (t2i shortlist),
 * by_column = 0: list q belongs to local image q and holds caption indices (i2t shortlist). */
int alad_shortlist_scatter(const float* S, int64_t ldS, float* S2, int64_t ld2, int32_t Ni, int32_t Nc,
                           const int32_t* idx, int32_t n_lists, int32_t k, int32_t by_column, int32_t img_off,
                           void* stream);

/* ---------------------------------------------------------------------------------
 * Peer exchange over NVLink (multi-GPU gallery; no reference counterpart: alad/evaluation.py is
 * single-process).  The packed caption rows a rank prepares are replicated on every rank by COPY ENGINES while the
 * persistent scoring kernel owns the SMs (an SM-resident collective cannot start next to it).
 *   alad_peer_alloc / _free     zeroed cudaMalloc buffer (the one exception to "no allocation inside": the memory
 *                               has to be exportable); *ptr is a device pointer
 *   alad_peer_export / _open / _close   cudaIpc handle (ALAD_PEER_HANDLE_BYTES, HOST memory) of such a buffer / mapping
 *                               of another rank's buffer into this process (peer access is enabled on first use)
 *   alad_peer_copy              device-to-device copy (local or peer) on the stream's copy engine
 *   alad_peer_signal            after everything enqueued on `stream` so far: *flag_ptrs[q] = value for q < n
 *                               (flag_ptrs: HOST array of device pointers, normally one slot in each peer's buffer)
 *   alad_peer_wait              blocks `stream` until flags[q] - value >= 0 for every q < n, q != skip (sequence
 *                               numbers; wrap-safe); after timeout_ms it stores 1 + q into *error (device) and traps: the
 *                               failure surfaces at the caller's next synchronisation, stale rows are never scored
 * ------------------------------------------------------------------------------- */
#define ALAD_MAX_PEERS 32
#define ALAD_PEER_HANDLE_BYTES 64
int alad_peer_alloc(void** ptr, int64_t bytes);
int alad_peer_free(void* ptr);
int alad_peer_export(const void* ptr, void* handle64);
int alad_peer_open(const void* handle64, void** ptr);
int alad_peer_close(void* ptr);
int alad_peer_copy(void* dst, const void* src, int64_t bytes, void* stream);
int alad_peer_signal(void* const* flag_ptrs, int32_t n, int32_t value, void* stream);
int alad_peer_wait(const int32_t* flags, int32_t n, int32_t value, int32_t skip, int64_t timeout_ms, int32_t* error,
                   void* stream);

/* Host-side atomics on a word of HOST memory shared by the ranks of one box (a /dev/shm mapping): the claim / done
 * counters of the cross-GPU work pool (aladin_b200/steal.py).  No CUDA work.
 *   alad_host_atomic_add   returns the value before the addition
 *   alad_host_atomic_cas   returns 1 when *p was `expected` and has been replaced by `desired`
 *   alad_host_atomic_load / _store   sequentially consistent */
int64_t alad_host_atomic_add(int64_t* p, int64_t v);
int32_t alad_host_atomic_cas(int64_t* p, int64_t expected, int64_t desired);
int64_t alad_host_atomic_load(const int64_t* p);
void alad_host_atomic_store(int64_t* p, int64_t v);

#ifdef __cplusplus
}
#endif
#endif /* ALAD_B200_H_ */
