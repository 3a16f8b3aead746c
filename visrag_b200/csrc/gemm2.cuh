// CTA-pair variant of the tcgen05 GEMM (tcgen05 cta_group::2): a cluster of two CTAs on one TPC computes a 256 x 256
// output tile. Each CTA stages ITS 128 rows of A and HALF of the B tile (128 of the 256 N rows) — so shared-memory
// traffic per MMA halves and six 32 KB stages fit where the single-CTA kernel has four 48 KB ones — and holds its 128
// accumulator rows in its own TMEM. Only the leader CTA (rank 0) issues tcgen05.mma (M = 256); both CTAs run a TMA
// producer and the epilogue warps.
//   full[stage]   lives in the LEADER: both producers arrive.expect_tx on it and their TMA bytes complete on it
//   empty[stage]  per CTA: released by the leader's tcgen05.commit multicast to both CTAs
//   tfull[acc]    per CTA: same multicast commit after a tile's last k-block
//   tempty[acc]   in the LEADER: 16 arrivals (8 epilogue warps of each CTA)
#pragma once
#include "gemm.cuh"

namespace vr {

// BN_ = 256, or 192 for N = 1152 (proj, fc2: 1152 = 6 x 192 tiles exactly, where 256-wide tiles compute 4.5 -> 5 tiles, 10 %
// of the MMAs on zero padding). 192-wide tiles support the LINEAR epilogues only (RoPE / SwiGLU work on 64-column pairs).
template <int BN_>
struct Gemm2CfgT {
    static constexpr int BN = BN_;
    static constexpr int STAGES = 6;
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;      // this CTA's 128 rows
    static constexpr int B_BYTES = (BN_ / 2) * GEMM_BK * 2;    // this CTA's half of the BN N rows
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;      // 32 KB (28 KB for BN = 192)
    static constexpr int EPI_STAGE_BYTES = 4096;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + GEMM_EPI_WARPS * EPI_STAGE_BYTES + 1024 + 256;
    static constexpr int TMEM_COLS = 512;
};
using Gemm2Cfg = Gemm2CfgT<256>;

template <int MODE, bool OUT_F32, bool GELU, int BN_ = 256>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
gemm2_tcgen05_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                     const GemmArgs g) {
    using Cfg = Gemm2CfgT<BN_>;
    static_assert(BN_ == 256 || (BN_ == 192 && MODE == VR_EPI_LINEAR), "192-wide tiles: LINEAR epilogues only");
    constexpr int STAGES = Cfg::STAGES;
    constexpr int BN = Cfg::BN;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
    uint8_t* smem_stage = smem + STAGES * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_stage + GEMM_EPI_WARPS * Cfg::EPI_STAGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES;
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
    volatile uint32_t* progress = tmem_slot + 1;  // tiles whose main loop has started; the leader writes BOTH CTAs' copy

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int cluster_id = blockIdx.x >> 1;
    const int num_clusters = gridDim.x >> 1;

    const int tiles_m = (g.M + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
    const int tiles_n = (g.N + BN - 1) / BN;
    const int num_tiles = tiles_m * tiles_n;
    const int num_kb = (g.K + GEMM_BK - 1) / GEMM_BK;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_a);
        tma_prefetch_desc(&tmap_b);
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < STAGES; ++i) {
            mbar_init(&full_bar[i], 2);   // one arrive.expect_tx per CTA (used in the leader only)
            mbar_init(&empty_bar[i], 1);
        }
        for (int i = 0; i < 2; ++i) {
            mbar_init(&tfull_bar[i], 1);
            mbar_init(&tempty_bar[i], 2 * GEMM_EPI_WARPS);  // used in the leader only
        }
        *progress = 0;
        fence_mbar_init();
    }
    cluster_sync_all();  // both CTAs resident and their barriers initialised before the pair-wide TMEM allocation
    if (warp == 2) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ------------------------------------------------------------ TMA producer (both CTAs)
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                const int m0 = (t / tiles_n) * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM;
                const int n0 = (t % tiles_n) * BN + static_cast<int>(rank) * (BN / 2);
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    const uint32_t lfull = mapa_u32(smem_u32(&full_bar[stage]), 0);
                    mbar_expect_tx_cluster(lfull, Cfg::STAGE_BYTES);
                    tma_load_2d_2sm(&tmap_a, lfull, smem_a + stage * Cfg::A_BYTES, kb * GEMM_BK, m0);
                    tma_load_2d_2sm(&tmap_b, lfull, smem_b + stage * Cfg::B_BYTES, kb * GEMM_BK, n0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------ MMA issuer (leader CTA only)
        if (rank == 0) {
            constexpr uint32_t idesc = make_idesc_f16(2 * GEMM_BM, BN, 1, 0, 0);
            const uint64_t desc_hi = make_smem_desc(0, 16, 1024, kLayoutSW128);
            const uint32_t a_lo0 = smem_u32(smem_a) >> 4, b_lo0 = smem_u32(smem_b) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
                tc_fence_after();
                if (lane == 0) {  // main loop of tile `it` starts: pace the residual prefetch warps of both CTAs
                    *progress = static_cast<uint32_t>(it + 1);
                    st_shared_cluster_u32(mapa_u32(smem_u32(const_cast<uint32_t*>(progress)), 1), static_cast<uint32_t>(it + 1));
                }
                const uint32_t d_tmem = tmem_base + acc * BN;
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = desc_hi | static_cast<uint64_t>(a_lo0 + stage * (Cfg::A_BYTES >> 4));
                        const uint64_t bd = desc_hi | static_cast<uint64_t>(b_lo0 + stage * (Cfg::B_BYTES >> 4));
                        umma_f16_ss_2sm(d_tmem, ad, bd, idesc, kb != 0 ? 1u : 0u);
                        umma_f16_ss_2sm(d_tmem, ad + 2, bd + 2, idesc, 1u);
                        umma_f16_ss_2sm(d_tmem, ad + 4, bd + 4, idesc, 1u);
                        umma_f16_ss_2sm(d_tmem, ad + 6, bd + 6, idesc, 1u);
                        umma_commit_2sm(&empty_bar[stage], 3);  // frees the slot in BOTH CTAs
                        if (kb == num_kb - 1) umma_commit_2sm(&tfull_bar[acc], 3);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 3) {
        // ------------------------------------------------------------ residual prefetcher (both CTAs, own 128 rows);
        // see gemm.cuh: pulls tile i's fp32 residual rows into L2 while tile i's main loop runs
        if (MODE == VR_EPI_LINEAR && OUT_F32 && g.prefetch_resid) {
            int it = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
                uint32_t started;
                while ((started = *progress) < static_cast<uint32_t>(it + 1)) __nanosleep(256);
                if (started > static_cast<uint32_t>(it + 1)) continue;  // too late to be useful
                const int m0 = (t / tiles_n) * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM;
                const int n0 = (t % tiles_n) * BN;
                const int cols = min(BN, g.N - n0) & ~3;
                if (cols <= 0) continue;
#pragma unroll
                for (int r = lane; r < GEMM_BM; r += 32) {
                    if (m0 + r < g.M)
                        l2_prefetch_bulk(g.epi.resid + static_cast<int64_t>(m0 + r) * g.epi.ldo + n0, cols * 4);
                }
            }
        }
    } else if (warp >= 4) {
        // ------------------------------------------------------------ epilogue (both CTAs, own 128 rows)
        const int ew = warp - 4;
        const int quarter = warp & 3;
        const int half = ew >> 2;
        constexpr int COLS_PER_WARP = BN / 2;
        uint8_t* st = smem_stage + ew * Cfg::EPI_STAGE_BYTES;
        int it = 0;
        for (int t = cluster_id; t < num_tiles; t += num_clusters, ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int m0 = (t / tiles_n) * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM;
            const int n0 = (t % tiles_n) * BN;
            const int row0 = m0 + quarter * 32;
            const uint32_t ltempty = mapa_u32(smem_u32(&tempty_bar[acc]), 0);
            float4 bias4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (MODE == VR_EPI_LINEAR) bias4 = epi_bias_prefetch(g, lane, n0 + half * COLS_PER_WARP, COLS_PER_WARP);
            // fp32 residual: the first chunk's loads fly while this tile's main loop finishes, every later chunk's while the
            // chunk before it is stored (ncu on proj, K = 1152: the epilogue warps sat on these loads, long_scoreboard 3.6,
            // tensor pipe 55 %)
            constexpr bool PRE = MODE == VR_EPI_LINEAR && OUT_F32;
            const bool pre = PRE && g.epi.resid != nullptr;
            uint4 rq[8];
            if (pre) {
                const int c0 = n0 + half * COLS_PER_WARP;
                resid_issue(rq, lane, g.epi.resid, g.epi.ldo, row0, g.M - row0, c0, g.N - c0);
            }
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
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(ltempty);
                    }
                    float v[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
                    const int col = n0 + half * COLS_PER_WARP + c * 32;
                    if (pre)
                        epi_linear<OUT_F32, GELU, PRE>(g, st, lane, row0, col, v, bias4, c, &rq,
                                                       c + 1 < COLS_PER_WARP / 32 ? col + 32 : -1);
                    else
                        epi_linear<OUT_F32, GELU>(g, st, lane, row0, col, v, bias4, c);
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
                        if (lane == 0) mbar_arrive_cluster(ltempty);
                    }
                    float a[32], b[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        a[j] = __uint_as_float(r0[j]);
                        b[j] = __uint_as_float(r1[j]);
                    }
                    const int col0 = n0 + half * COLS_PER_WARP + c * 64;
                    if (col0 < g.N) {
                        if (MODE == VR_EPI_ROPE) epi_rope(g, st, lane, row0, col0, a, b);
                        else epi_swiglu(g, st, lane, row0, col0, a, b);
                    }
                }
            }
        }
    }

    tc_fence_before();
    cluster_sync_all();  // the peer may still be reading this CTA's smem / arriving on its barriers until here
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
    }
}

}  // namespace vr
