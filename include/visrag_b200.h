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

#define VR_ABI_VERSION 1

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

#ifdef __cplusplus
}
#endif
#endif /* VISRAG_B200_H */
