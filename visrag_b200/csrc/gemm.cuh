// Persistent warp-specialised tcgen05 GEMM:  C[M,N] = A[M,K] * B[N,K]^T  (both K-major).
//
//   warp 0 (one lane)  TMA producer : A/B 128B-swizzled tiles -> 4-stage smem ring
//   warp 1 (one lane)  MMA issuer   : tcgen05.mma 128 x BN x 16, fp32 accumulators in TMEM,
//                                     two accumulator stages (2 x BN columns) so the epilogue
//                                     of tile i overlaps the main loop of tile i+1
//   warp 2             TMEM allocator
//   warps 4..11        epilogue     : tcgen05.ld (thread = one output row, 32 columns per
//                                     chunk) -> fused bias / GELU / residual / RoPE / SwiGLU
//                                     -> 16-byte global stores
//
// Three mbarrier pipelines: smem full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
// Tiles are scheduled round-robin over a grid of min(#tiles, #SMs) CTAs, n-block fastest so
// that CTAs running together share the same A rows in L2.
#pragma once
#include "ptx.cuh"
#include "../../include/visrag_b200.h"

namespace vr {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;  // 64 bf16 = 128 B = one swizzle-128B row
constexpr int GEMM_EPI_WARPS = 8;
constexpr int GEMM_THREADS = (4 + GEMM_EPI_WARPS) * 32;

template <int BN>
struct GemmCfg {
    static constexpr int STAGES = (BN == 256) ? 4 : 6;
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = 2 * BN;  // 256 or 512: power of two >= 32
};

struct GemmArgs {
    int M, N, K;
    vr_gemm_epilogue epi;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }

// ---------------------------------------------------------------------------------------
// Epilogue bodies. Each thread owns output row `row`; `v` holds 32 consecutive accumulator
// columns starting at global column `col0`.
// ---------------------------------------------------------------------------------------
template <bool OUT_F32, bool GELU>
__device__ __forceinline__ void epi_linear(const GemmArgs& g, int row, int col0, float (&v)[32]) {
    const vr_gemm_epilogue& e = g.epi;
    if (row >= g.M) return;
    const int N = g.N;
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
        const int c = col0 + j8 * 8;
        if (c >= N) break;  // N % 8 == 0: a group of 8 is entirely in or out
        float x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = v[j8 * 8 + j];
        if (e.bias) {
            const float4 b0 = *reinterpret_cast<const float4*>(e.bias + c);
            const float4 b1 = *reinterpret_cast<const float4*>(e.bias + c + 4);
            x[0] += b0.x; x[1] += b0.y; x[2] += b0.z; x[3] += b0.w;
            x[4] += b1.x; x[5] += b1.y; x[6] += b1.z; x[7] += b1.w;
        }
        if (GELU) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] = gelu_erf(x[j]);
        }
        if (e.scale != 1.0f) {
#pragma unroll
            for (int j = 0; j < 8; ++j) x[j] *= e.scale;
        }
        if (e.rowadd) {
            const float* p = e.rowadd + static_cast<int64_t>(row % e.rowadd_period) * N + c;
            const float4 a0 = *reinterpret_cast<const float4*>(p);
            const float4 a1 = *reinterpret_cast<const float4*>(p + 4);
            x[0] += a0.x; x[1] += a0.y; x[2] += a0.z; x[3] += a0.w;
            x[4] += a1.x; x[5] += a1.y; x[6] += a1.z; x[7] += a1.w;
        }
        if (e.resid) {
            const float* p = e.resid + static_cast<int64_t>(row) * e.ldo + c;
            const float4 a0 = *reinterpret_cast<const float4*>(p);
            const float4 a1 = *reinterpret_cast<const float4*>(p + 4);
            x[0] += a0.x; x[1] += a0.y; x[2] += a0.z; x[3] += a0.w;
            x[4] += a1.x; x[5] += a1.y; x[6] += a1.z; x[7] += a1.w;
        }
        if (OUT_F32) {
            float* o = reinterpret_cast<float*>(e.out) + static_cast<int64_t>(row) * e.ldo + c;
            *reinterpret_cast<float4*>(o) = make_float4(x[0], x[1], x[2], x[3]);
            *reinterpret_cast<float4*>(o + 4) = make_float4(x[4], x[5], x[6], x[7]);
        } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e.out) + static_cast<int64_t>(row) * e.ldo + c;
            uint4 pk;
            pk.x = pack_bf16x2(x[0], x[1]);
            pk.y = pack_bf16x2(x[2], x[3]);
            pk.z = pack_bf16x2(x[4], x[5]);
            pk.w = pack_bf16x2(x[6], x[7]);
            *reinterpret_cast<uint4*>(o) = pk;
        }
    }
}

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* o, const float (&x)[32]) {
#pragma unroll
    for (int j8 = 0; j8 < 4; ++j8) {
        uint4 pk;
        pk.x = pack_bf16x2(x[j8 * 8 + 0], x[j8 * 8 + 1]);
        pk.y = pack_bf16x2(x[j8 * 8 + 2], x[j8 * 8 + 3]);
        pk.z = pack_bf16x2(x[j8 * 8 + 4], x[j8 * 8 + 5]);
        pk.w = pack_bf16x2(x[j8 * 8 + 6], x[j8 * 8 + 7]);
        *reinterpret_cast<uint4*>(o + j8 * 8) = pk;
    }
}

// RoPE (modeling_minicpm.py:259-290): a head is 64 columns [lo(32) | hi(32)];
//   lo' = lo*cos - hi*sin ; hi' = hi*cos + lo*sin   with cos/sin[pos, 0..31].
__device__ __forceinline__ void epi_rope(const GemmArgs& g, int row, int col0, float (&lo)[32], float (&hi)[32]) {
    const vr_gemm_epilogue& e = g.epi;
    if (row >= g.M) return;
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e.out) + static_cast<int64_t>(row) * e.ldo + col0;
    if (col0 < e.rope_cols) {
        const int pos = e.positions[row];
        const float* cs = e.rope_cos + static_cast<int64_t>(pos) * 32;
        const float* sn = e.rope_sin + static_cast<int64_t>(pos) * 32;
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 c4 = *reinterpret_cast<const float4*>(cs + j4 * 4);
            const float4 s4 = *reinterpret_cast<const float4*>(sn + j4 * 4);
            const float c[4] = {c4.x, c4.y, c4.z, c4.w};
            const float s[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float a = lo[j4 * 4 + j], b = hi[j4 * 4 + j];
                lo[j4 * 4 + j] = a * c[j] - b * s[j];
                hi[j4 * 4 + j] = b * c[j] + a * s[j];
            }
        }
    }
    store_bf16x32(o, lo);
    store_bf16x32(o + 32, hi);
}

// SwiGLU (modeling_minicpm.py:333): accumulator columns [gate(32) | up(32)] -> 32 outputs.
__device__ __forceinline__ void epi_swiglu(const GemmArgs& g, int row, int col0, float (&gt)[32], float (&up)[32]) {
    const vr_gemm_epilogue& e = g.epi;
    if (row >= g.M) return;
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(e.out) + static_cast<int64_t>(row) * e.ldo + (col0 >> 1);
#pragma unroll
    for (int j = 0; j < 32; ++j) gt[j] = silu(gt[j]) * up[j];
    store_bf16x32(o, gt);
}

// ---------------------------------------------------------------------------------------
template <int BN, int MODE, bool OUT_F32, bool GELU, int AB_FMT /*0 f16, 1 bf16*/>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                    const GemmArgs g) {
    using Cfg = GemmCfg<BN>;
    constexpr int STAGES = Cfg::STAGES;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES;
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    const int tiles_m = (g.M + GEMM_BM - 1) / GEMM_BM;
    const int tiles_n = (g.N + BN - 1) / BN;
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = (g.K + GEMM_BK - 1) / GEMM_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], GEMM_EPI_WARPS);
        }
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const int m0 = (t / tiles_n) * GEMM_BM;
                const int n0 = (t % tiles_n) * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(&tmap_a, &full_bar[stage], smem_a + stage * Cfg::A_BYTES, kb * GEMM_BK, m0);
                    if (BN == 256) {
                        // a TMA box is at most 256 rows; keep boxes at 128 rows for both operands
                        tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * Cfg::B_BYTES, kb * GEMM_BK, n0);
                        tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * Cfg::B_BYTES + Cfg::B_BYTES / 2,
                                    kb * GEMM_BK, n0 + 128);
                    } else {
                        tma_load_2d(&tmap_b, &full_bar[stage], smem_b + stage * Cfg::B_BYTES, kb * GEMM_BK, n0);
                    }
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(GEMM_BM, BN, AB_FMT, 0, 0);
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t a_addr = smem_u32(smem_a + stage * Cfg::A_BYTES);
                    const uint32_t b_addr = smem_u32(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < GEMM_BK / 16; ++k) {
                        // K-major SW128: 8-row groups are 1024 B apart; +32 B per 16-element K step
                        const uint64_t ad = make_smem_desc(a_addr + k * 32, 16, 1024, kLayoutSW128);
                        const uint64_t bd = make_smem_desc(b_addr + k * 32, 16, 1024, kLayoutSW128);
                        umma_f16_ss(d_tmem, ad, bd, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    umma_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
                    if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue
        const int ew = warp - 4;
        const int quarter = warp & 3;          // TMEM lane quarter this warp may touch
        const int half = ew >> 2;              // which half of the BN columns
        constexpr int COLS_PER_WARP = BN / 2;  // 128 or 64
        int it = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int m0 = (t / tiles_n) * GEMM_BM;
            const int n0 = (t % tiles_n) * BN;
            const int row = m0 + quarter * 32 + lane;
            mbar_wait(&tfull_bar[acc], acc_phase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * BN + half * COLS_PER_WARP;
            if (MODE == VR_EPI_LINEAR) {
#pragma unroll 1
                for (int c = 0; c < COLS_PER_WARP / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld_32x32(taddr + c * 32, r);
                    tmem_ld_wait();
                    if (c == COLS_PER_WARP / 32 - 1) {
                        // all of this warp's TMEM reads for the tile are done: hand the stage back
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    epi_linear<OUT_F32, GELU>(g, row, n0 + half * COLS_PER_WARP + c * 32, v);
                }
            } else {
#pragma unroll 1
                for (int c = 0; c < COLS_PER_WARP / 64; ++c) {
                    uint32_t r0[32], r1[32];
                    tmem_ld_32x32(taddr + c * 64, r0);
                    tmem_ld_32x32(taddr + c * 64 + 32, r1);
                    tmem_ld_wait();
                    if (c == COLS_PER_WARP / 64 - 1) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                    }
                    float a[32], b[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        a[j] = __uint_as_float(r0[j]);
                        b[j] = __uint_as_float(r1[j]);
                    }
                    const int col0 = n0 + half * COLS_PER_WARP + c * 64;
                    if (col0 < g.N) {
                        if (MODE == VR_EPI_ROPE) epi_rope(g, row, col0, a, b);
                        else epi_swiglu(g, row, col0, a, b);
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
    }
}

}  // namespace vr
