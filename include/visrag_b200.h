/*
 * visrag_b200 — C ABI of the B200-native VisRAG-Ret embedding + retrieval hot path.
 *
 * The reference (OpenBMB/VisRAG) is pure Python: it has no FFI for this path. The
 * drop-in boundary is the three Python call signatures of SURVEY.md §8(b); this C ABI
 * is what sits underneath the Python mirror of those signatures
 * (visrag_b200/{modeling,encoder,retriever}.py), and each entry point below names the
 * reference code whose GPU work it replaces (paths relative to the reference repo).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer unless the name ends in _host; the caller owns
 *    all buffers (inputs, outputs, workspaces); the library allocates nothing;
 *  - all work is enqueued on the caller's cudaStream_t (passed as void*); no hidden
 *    synchronisation, so every call is CUDA-graph capturable;
 *  - return value: 0 on success, non-zero on error; vr_last_error() returns a
 *    thread-local message; nothing throws or exits across the ABI;
 *  - row-major matrices; "ld*" = leading dimension in ELEMENTS.
 */
#ifndef VISRAG_B200_H
#define VISRAG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VR_ABI_VERSION 2

typedef enum { VR_BF16 = 0, VR_F16 = 1, VR_F32 = 2 } vr_dtype;

const char* vr_last_error(void);
int vr_abi_version(void);

/* ------------------------------------------------------------------------------------
 * Dense contraction  C[M,N] = A[M,K] * B[N,K]^T  on tcgen05 tensor cores (TMA-staged
 * 128B-swizzled tiles -> UMMA -> fp32 accumulators in TMEM -> fused epilogue).
 * Replaces every nn.Linear / Conv2d-as-GEMM the reference dispatches to cuBLAS/cuDNN:
 *   timm/layers/patch_embed.py:87 (patch conv), timm/models/vision_transformer.py:88,105
 *   (qkv, proj), timm/layers/mlp.py:41-49 (fc1+GELU, fc2), resampler.py:154,159-167,
 *   modeling_minicpm.py:850-852,908 (q,k,v,o), modeling_minicpm.py:333 (SwiGLU MLP).
 * A and B are bf16 (VR_BF16) or fp16 (VR_F16); K*2 bytes and lda/ldb*2 bytes must be
 * multiples of 16 (TMA); N must be a multiple of 8.
 * ---------------------------------------------------------------------------------- */
typedef enum {
    VR_EPI_LINEAR = 0, /* out = [resid +] scale*(gelu?(acc + bias)) [+ rowadd[row % period]] */
    VR_EPI_ROPE = 1,   /* MiniCPM q|k|v: rotate-half RoPE on 64-wide heads in columns < rope_cols; bf16 out */
    VR_EPI_SWIGLU = 2  /* B rows interleaved [32 gate | 32 up] per 64: out[:, j] = silu(g_j)*u_j; bf16, N/2 cols */
} vr_epi_mode;

typedef struct {
    int32_t mode;          /* vr_epi_mode */
    int32_t out_dtype;     /* VR_BF16 or VR_F32 (LINEAR); others write bf16 */
    int32_t act_gelu;      /* LINEAR: exact erf-GELU applied to (acc + bias) */
    float scale;           /* LINEAR: multiplies the activation before the residual add */
    const float* bias;     /* [N] fp32 or NULL */
    const float* resid;    /* [M, ldo] fp32 or NULL; may alias out (in-place residual stream) */
    const float* rowadd;   /* [period, N] fp32 or NULL (ViT position embedding) */
    int32_t rowadd_period; /* rows of rowadd; row index used is (row % period) */
    const int32_t* positions; /* ROPE: [M] position of each packed token inside its sequence */
    const float* rope_cos; /* ROPE: [max_pos, 32] fp32 */
    const float* rope_sin; /* ROPE: [max_pos, 32] fp32 */
    int32_t rope_cols;     /* ROPE: columns [0, rope_cols) are rotated (q and k); the rest (v) pass through */
    void* out;             /* [M, ldo] */
    int64_t ldo;
} vr_gemm_epilogue;

int vr_gemm(const void* A, int64_t lda, const void* B, int64_t ldb, int32_t ab_dtype, int32_t M, int32_t N, int32_t K,
            const vr_gemm_epilogue* epi, void* stream);

/* Same operation with the kernel variant chosen by the caller (benchmarks, parity tests of every variant):
 *   block_n = 0    what vr_gemm picks: the CTA-pair kernel when M > 128 and N >= 256, 128 x 64 tiles when M <= 128
 *                  (weight-streaming bound: more, narrower tiles), else 128-row tiles
 *   block_n = 2    CTA-pair kernel (tcgen05 cta_group::2, one 256x256 tile per pair of SMs, each CTA stages half of B);
 *                  LINEAR epilogues with N % 192 == 0, N % 256 != 0 and K <= 2304 (proj: N = K = 1152) use 256x192 tiles
 *   block_n = 4    CTA-pair kernel with 256x192 tiles forced (LINEAR epilogues only)
 *   block_n = 256 / 128 / 64   single-CTA kernel, token-major accumulator, 128 tokens x block_n features per tile
 *   block_n = 3    single-CTA kernel, feature-major accumulator (the weight tile is the MMA's M operand), LINEAR
 *                  epilogues only; its epilogue needs no shared-memory transpose */
int vr_gemm_tuned(const void* A, int64_t lda, const void* B, int64_t ldb, int32_t ab_dtype, int32_t M, int32_t N,
                  int32_t K, const vr_gemm_epilogue* epi, int32_t block_n, void* stream);


/* ------------------------------------------------------------------------------------
 * Device image front-end: Pillow-bit-compatible 8-bit bicubic resampling + grid crop.
 * Replaces the host-side `image.resize(size, Image.BICUBIC)` calls of slice_image
 * (modeling_minicpmv/modeling_minicpmv.py:509,519,531) and split_to_patches (:571-592); the arithmetic is
 * Pillow's ImagingResample (src/libImaging/Resample.c): horizontal pass over source rows
 * [row_first, row_first+row_count) into an 8-bit intermediate, then the vertical pass, 22-bit fixed-point
 * coefficients, 32-bit sums, clip8. Results are bit-identical to PIL.
 *   src        [n, in_h, in_w, src_pixel_bytes] uint8 (n pages of one size); src_pixel_bytes = 3 (packed RGB) or
 *              4 (RGBX: Pillow's native row layout for mode "RGB", so a page can travel without any repacking on
 *              the host; the 4th byte is ignored)
 *   bounds_*   [out, 2] int32 (first source index, tap count); coefficients int32 fixed point as Pillow's
 *              precompute_coeffs + normalize_coeffs_8bpc produce them (frontend.resample_coeffs): coeffs_v
 *              [out_h, ksize_v] (row per output row), coeffs_h TAP-MAJOR [ksize_h, out_w] (coalesced across the
 *              threads of the horizontal pass); NULL for an axis whose size does not change (Pillow skips that
 *              pass). in_w <= 12288.
 *   tmp        workspace of n * row_count * ((out_w*3 + 3) & ~3) bytes (4-byte row pitch), needed when both passes run
 *   out        slice buffer [*, cell_h, cell_w, 3] uint8: the out_h x out_w result of page i is cut into
 *              (out_h/cell_h) x (out_w/cell_w) cells, row-major, stored as slices first_cell[i], first_cell[i]+1, ...
 *              (cell = whole image for the thumbnail). first_cell: [n] int32, device.
 * ---------------------------------------------------------------------------------- */
int vr_resample_u8(const uint8_t* src, int32_t src_pixel_bytes, int32_t n, int32_t in_h, int32_t in_w, const int32_t* bounds_h,
                   const int32_t* coeffs_h, int32_t ksize_h, const int32_t* bounds_v, const int32_t* coeffs_v,
                   int32_t ksize_v, int32_t row_first, int32_t row_count, int32_t out_h, int32_t out_w, uint8_t* tmp,
                   uint8_t* out, const int32_t* first_cell, int32_t cell_h, int32_t cell_w, void* stream);

/* ------------------------------------------------------------------------------------
 * Fused softmax(Q K^T * scale) V on tcgen05. Sequences longer than 128 queries: two query tiles per CTA, S, P
 * (bf16, TS-MMA operand) and the O accumulator all live in TMEM, lazy rescaling. Up to 128 queries per sequence:
 * single-tile kernel, P re-staged through 128B-swizzled shared memory, running max/sum/O in registers.
 * Replaces F.scaled_dot_product_attention in timm/models/vision_transformer.py:92-96 (ViT,
 * 16 heads x 72, no mask), modeling_minicpm.py:895-903 (MiniCPM, causal + right padding ->
 * here: packed var-len sequences, no padding rows at all) and nn.MultiheadAttention in
 * resampler.py:159-163 (64 learned queries x N keys, 18 heads x 128).
 * q/k/v are bf16 row-major token matrices; head h starts at column *_col0 + h*head_stride
 * (head_stride = head_dim rounded up to a multiple of 16; pad columns must hold zeros).
 * ---------------------------------------------------------------------------------- */
typedef struct {
    const void* q;  int64_t ldq;  int64_t q_rows;   /* q_rows: rows in the q buffer (TMA bound) */
    const void* k;  int64_t ldk;
    const void* v;  int64_t ldv;  int64_t kv_rows;  /* rows in the k and v buffers */
    int32_t q_col0, k_col0, v_col0;
    int32_t head_stride;      /* 64, 80 or 128 */
    int32_t head_dim;         /* 64, 72 or 128: output columns per head */
    int32_t heads, batch;
    const int32_t* cu_q;      /* [batch+1] packed query offsets, or NULL: every item uses q rows [0, max_q) */
    const int32_t* cu_k;      /* [batch+1] packed key offsets */
    int32_t max_q, max_k;     /* longest query / key sequence (grid sizing) */
    int32_t causal;
    float scale;
    void* out; int64_t ldo;   /* bf16; row = cu_q ? cu_q[b]+i : b*max_q+i ; head h at column h*head_dim */
    int32_t flags;            /* VR_ATTN_* */
} vr_attn_params;

/* The caller guarantees V[:, head_dim] == 1 for every head (head_dim == head_stride - 8; e.g. a bias of 1 in the zero
 * padding of the QKV projection). Kernels that can use it take the softmax denominator out of the P.V MMA (column
 * head_dim of the accumulator) instead of summing P in registers; the others ignore the column. Results are the same
 * up to fp32 summation order. */
#define VR_ATTN_V_ONES_COLUMN 1

int vr_attention(const vr_attn_params* p, void* stream);
/* Dispatch: non-causal sequences longer than 128 queries (the ViT) run the persistent kernel of attention4.cuh (one CTA per
 * SM loops over (query-tile pair, head, sequence) items; P in its own TMEM buffer so Q.K^T of the next key block is issued
 * while the exps of the current one run); causal long sequences the two-tile kernel of attention2.cuh; everything else
 * the single-tile kernel.
 * test / benchmark hook (process-wide): 0 = default dispatch, 1 = always the one-tile-per-CTA kernel, 2 = attention2 wherever
 * attention4 is the default, 5 = two-tile kernel with Q in tensor memory as well (measured slower than the default; kept
 * as a tested alternative) */
void vr_attention_force_v1(int32_t variant);


/* ------------------------------------------------------------------------------------
 * HBM-bound standalone kernels (128-bit vectorised, one pass over the activation).
 * ---------------------------------------------------------------------------------- */

/* uint8 HWC slices -> normalised bf16 patch matrix (ToTensor + Normalize(0.5,0.5) + the unfold of the
 * 14x14/stride-14 patch conv): modeling_minicpmv.py:84-92 + timm/layers/patch_embed.py:87.
 * pixels: [n_slices, h, w, 3] uint8 (h, w multiples of `patch`); out: [n_slices*(h/patch)*(w/patch), ldo] bf16,
 * column c*patch*patch + ky*patch + kx (the Conv2d weight's flattening); columns [3*patch^2, ldo) are zeroed.
 * bf16(fma(u, 2/255, -1)) - bit-identical to bf16((u/255 - 0.5)/0.5) for all 256 byte values. patch <= 85; a strip of `patch`
 * pixel rows must fit 200 KB of shared memory (w up to ~4800 for patch 14). */
int vr_im2col_norm(const uint8_t* pixels, int32_t n_slices, int32_t h, int32_t w, int32_t patch, void* out, int64_t ldo,
                   void* stream);

/* LayerNorm over the last dim (timm vision_transformer.py:142,155,525; resampler.py:155,166): fp32 in -> bf16 out.
 * If out2 != NULL also writes out2 = LN(x) + add[row % add_period] (bf16) — the resampler's K input (kv + pos). */
int vr_layernorm(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int32_t rows, int32_t dim,
                 void* out, int64_t ldo, void* out2, const float* add, int32_t add_period, void* stream);

/* RMSNorm (modeling_minicpm.py:119-123): fp32 in -> bf16 out. */
int vr_rmsnorm(const float* x, int64_t ldx, const float* gamma, float eps, int32_t rows, int32_t dim, void* out, int64_t ldo,
               void* stream);

/* LM input assembly (modeling_minicpmv.py:139-166): for packed token t,
 *   src[t] >= 0 : h[t] = vision[src[t]]            (resampler output row, fp32)
 *   src[t] <  0 : h[t] = embed[-(src[t]+1)] * scale_emb   (bf16 table)
 * h: [tokens, dim] fp32. */
int vr_build_lm_input(const int32_t* src, int32_t tokens, int32_t dim, const void* embed_bf16, float scale_emb,
                      const float* vision, int64_t ldv, float* h, int64_t ldh, void* stream);

/* Final RMSNorm + pooling + L2 normalise (modeling_minicpm.py:1280; dense_retrieval_model.py:170-223):
 * per packed sequence b (rows cu[b]..cu[b+1]) of h [tokens, dim] fp32 -> reps [batch, dim] fp32.
 * pooling: 0 wmean (w_t = t+1), 1 mean, 2 lasttoken, 3 cls. normalise: x / max(||x||, 1e-12).
 * One thread-block cluster of 8 (or 4) CTAs per sequence, every row read once; h, gamma and reps 16-byte aligned; any
 * sequence length (no per-length shared memory), dim <= 4096. */
int vr_pool_norm(const float* h, int64_t ldh, const float* gamma, float eps, const int32_t* cu, int32_t batch, int32_t dim,
                 int32_t pooling, int32_t normalize, float* reps, void* stream);


/* ------------------------------------------------------------------------------------
 * Similarity + top-k: replaces `torch.matmul(Q, D^T)` + `torch.topk` of
 * retriever/dense_retriever.py:25-30 (fp32 scores) and the Python merge loop `:88-92`.
 * The [nq, nd] score matrix is never materialised:
 *   vr_score_filter  : tcgen05 fp16 GEMM on CTA pairs (256 queries x 256 docs per MMA tile) with a fused per-row running
 *                      top-16 per (query, doc range); writes `lists = 2*ranges` sorted 16-entry candidate lists per query into
 *                      cand_scores / cand_ids [nq, lists*16] (unused lists: score -inf, id -1; the first score of the LAST
 *                      list slot is scratch - the query's running threshold - and carries id -1);
 *   vr_score_rescore : keeps the max(32, 2k) best candidates by approximate score, rescoring them exactly in fp32, top-k by
 *                      (score desc, id asc), and a proof that nothing dropped (by a list or by the pruning) could belong to
 *                      the top-k (flags[q] = 1 when the proof fails; the caller then reruns that query through
 *                      vr_score_exact + vr_topk_rows).
 * Results are therefore exactly the fp32 top-k. Doc ids returned are `local index + id_offset`.
 * ---------------------------------------------------------------------------------- */
int vr_score_ranges(int32_t nq, int64_t nd);   /* sizes the candidate buffers: [nq, ranges*2*16]; host arithmetic only (needs no
                                                * GPU); pass the value on to vr_score_filter / vr_score_rescore unchanged */
int vr_score_list_len(void);                   /* 16 */
/* The filter's work decomposition for (nq, nd), host arithmetic only: out6 = {doc tiles of 256, doc ranges R, 256-query blocks,
 * work items = R * blocks (item i = range i / blocks of query block i % blocks, tiles [T*r/R, T*(r+1)/R)), CTA pairs launched,
 * candidate lists per query (= 2 * vr_score_ranges)}. For tests and capacity planning. */
int vr_score_plan(int32_t nq, int64_t nd, int32_t* out6);
int vr_f32_to_f16_rows(const float* src, int64_t rows, int32_t dim, void* dst_f16, float* norms, float* max_norm,
                       void* stream);          /* norms / max_norm optional; *max_norm must be pre-zeroed */
int vr_score_filter(const void* q_f16, int32_t nq, const void* d_f16, int64_t nd, int32_t dim, int32_t ranges,
                    float* cand_scores, int32_t* cand_ids, void* stream);
int vr_score_rescore(const float* q_f32, int32_t nq, const float* d_f32, int64_t nd, int32_t dim, int32_t ranges,
                     const float* cand_scores, const int32_t* cand_ids, const float* max_doc_norm, int32_t k,
                     int64_t id_offset, float* out_scores, int64_t* out_ids, int32_t* flags, void* stream);
/* plain fp32 scan (small problems, flagged queries): scores [nq, nd] */
int vr_score_exact(const float* q_f32, int32_t nq, const float* d_f32, int64_t nd, int32_t dim, float* scores, void* stream);
/* top-k of every row of scores [rows, cols]; ids == NULL -> column index (+ id_offset), else ids[row, col]
 * (negative ids are skipped): also the k-way merge of per-shard / per-rank partial top-k lists. */
int vr_topk_rows(const float* scores, const int64_t* ids, int32_t rows, int64_t cols, int32_t k, int64_t id_offset,
                 float* out_scores, int64_t* out_ids, void* stream);
/* Same result for few rows x many columns (single-query retrieval over a large index): `chunks` blocks per row each
 * reduce a column range to a top-k list (pass 1, workspace ws_scores / ws_ids [rows, chunks, k]), then the lists are
 * merged (pass 2). One block per row would scan a 1 M-column row k times on a single SM. */
int vr_topk_rows_chunked(const float* scores, int32_t rows, int64_t cols, int32_t k, int64_t id_offset, int32_t chunks,
                         float* ws_scores, int64_t* ws_ids, float* out_scores, int64_t* out_ids, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VISRAG_B200_H */
