// Attention v4 (ViT shape: many 1024-token sequences, 16 heads x 72): persistent two-tile kernel with P outside S.
//
// What the engines cost per key block of 128 keys and CTA (two 128-query tiles), measured with tools/ubench/ub_tc.cu:
//   MUFU   32768 ex2 at 16/clk/SM                                                   2048 cycles
//   tensor 2 x (5 SS-MMA N=128 at 98 + 8 x (TS N=64 at 41 + TS N=16 at 17))          1900 cycles
//   TMEM   128 KB of S at ~280 B/clk/SM (64 B/clk per sub-partition)                 460 cycles
//   hand-offs: mbarrier arrive -> waiter runs 170 cycles, tcgen05.commit -> waiter 190 cycles
// The v2 kernel (attention2.cuh) needs ~4400 cycles per key block: P overlays S, so each tile runs
// QK(j) -> ld S -> max -> exp -> st P -> PV(j) -> QK(j+1) as ONE dependency chain, and every CTA (8 key blocks) pays its
// own prologue (barrier init, TMEM alloc, first TMA round trip) and epilogue. v4 changes two things:
//   * P does not overlay S. It lives in its own 64-column TMEM buffer that the two tiles share in strict alternation
//     (A(j), B(j), A(j+1), ...): 2 x 128 (S) + 2 x HS (O) + 64 (P) = 480 of the 512 columns. S_x is free again as soon as
//     the softmax warps have pulled it into registers, so QK_x(j+1) is issued while the exps of block j are still running.
//     The packed P row (64 words) waits in registers until the preceding use's P.V has retired.
//   * The CTA is persistent: one CTA per SM loops over (query-tile pair, head, sequence) work items; barriers, TMEM and
//     the K/V ring live across items, the producer runs ahead into the next item (double-buffered Q), the MMA issuer is
//     event driven (polls S_x consumed / P_x written per tile with mbarrier.test_wait).
//   * ONES: when the caller guarantees V[:, head_dim] == 1 (a bias in the QKV GEMM's zero padding), column head_dim of O
//     is the softmax denominator, accumulated by the tensor core from exactly the bf16 P values the numerator uses; the
//     128 FADDs per row and key block disappear (0.99 instead of 1.06 ms).
// Measured at 128 x 16 x 1024 x 72 (profiles/r02_attention4_experiments.md): this kernel 0.985 ms, v2 1.118 ms. Variants that
// were built, measured and dropped (kept in the git history): a private 32-column P buffer per tile reused for both halves
// of a block (1.05 ms), the same with the second half's store deferred behind the next S load (1.27 ms), the second half
// of P through swizzled shared memory + SS-MMA so that every tile owns a whole block of P (1.08 ms, with one blocking
// issuer warp per tile, prefetched work-item descriptors and the O read-out deferred into the next item: 1.07 ms), tile B
// started 800-2400 cycles late (no change). Ablations of that last variant: no exp2 0.86 ms, no P stores/exps 0.68 ms,
// no MMAs 1.00 ms, no K/V loads 1.07 ms, barrier protocol alone 0.36 ms - the phases of a key block add up almost
// serially: with two softmax warps per sub-partition there is too little parallelism to hide the LDTM / hand-off /
// store-completion latencies behind the other warp's exps.
// Warp roles: 0-3 softmax tile A, 4-7 softmax tile B (thread = query row), 8 MMA issuer, 9 TMA producer, 10-11 idle.
#pragma once
#ifndef VR_A_SLEEP
#define VR_A_SLEEP 32
#endif
#include "attention2.cuh"

namespace vr {

#ifndef VR_A_TRACE
#define VR_A_TRACE 0
#endif
#if VR_A_TRACE
// Debug build only (tools/trace_attention.py): clock64 stamps of CTA 0's softmax warps 0 (tile A) and 4 (tile B), and of
// the MMA issuer, for the first 96 key blocks.
__device__ long long g_att_trace[3][96][8];
#define VR_TR(x, blk, slot) do { if (blockIdx.x == 0 && lane == 0 && (warp & 3) == 0 && (blk) < 96) g_att_trace[x][blk][slot] = clock64(); } while (0)
#else
#define VR_TR(x, blk, slot) do { } while (0)
#endif

constexpr int ATT4_THREADS = 384;

template <int HS>
struct Att4Cfg {
    using C1 = AttCfg<HS>;
    static_assert(HS == 64 || HS == 80, "v4 is built for head stride 64 / 80");
    static constexpr int TILE = C1::TILE_BYTES;
    static constexpr int KS = 2;                       // K and V ring depth (3 before the output staging tiles took the room)
    static constexpr int OFF_Q = 0;                    // [2 buffers][2 tiles]
    static constexpr int OFF_K = 4 * TILE;
    static constexpr int OFF_V = OFF_K + KS * TILE;
    static constexpr int OFF_OST = OFF_V + KS * TILE;   // [2 tiles] output staging for the bulk tensor store: 128 rows x head_dim bf16
    static constexpr int OFF_BAR = OFF_OST + 2 * TILE;
    static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
    static constexpr uint32_t T_S = 0, T_O = 256, T_P = 256 + 2 * HS;  // TMEM columns
    static_assert(T_P + 64 <= 512, "TMEM budget");
};

__device__ __forceinline__ bool mbar_test(uint64_t* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, P1;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}

struct Att4Item {
    int head, b, k_begin, len_k, q_begin, len_q, q0, nkt;
    bool valid, b_active;
};

__device__ __forceinline__ Att4Item att4_item(const AttArgs& a, int w, int nqp) {
    Att4Item it;
    const int qp = w % nqp;
    const int t = w / nqp;
    it.head = t % a.heads;
    it.b = t / a.heads;
    it.k_begin = a.cu_k[it.b];
    it.len_k = a.cu_k[it.b + 1] - it.k_begin;
    it.q_begin = a.cu_q ? a.cu_q[it.b] : 0;
    it.len_q = a.cu_q ? a.cu_q[it.b + 1] - it.q_begin : a.max_q;
    it.q0 = qp * 2 * ATT_BM;
    it.valid = it.q0 < it.len_q && it.len_k > 0;
    it.b_active = it.q0 + ATT_BM < it.len_q;
    it.nkt = (it.len_k + ATT_BN - 1) / ATT_BN;
    return it;
}

template <int HS, bool ONES>
__global__ void __launch_bounds__(ATT4_THREADS, 1)
attention4_tcgen05_kernel(const __grid_constant__ AttMaps maps, const AttArgs a, const int total_items, const int nqp) {
    using Cfg = Att4Cfg<HS>;
    using C1 = AttCfg<HS>;
    constexpr int KS = Cfg::KS;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* q_full = bars + 0;          // [2] TMA bytes of a Q buffer (both tiles)
    uint64_t* q_empty = bars + 2;         // [2] commit: every QK of the item retired
    uint64_t* k_full = bars + 4;          // [KS]
    uint64_t* k_empty = bars + 4 + KS;    // [KS] commit
    uint64_t* v_full = bars + 4 + 2 * KS; // [KS]
    uint64_t* v_empty = bars + 4 + 3 * KS;// [KS] commit
    uint64_t* s_full = bars + 4 + 4 * KS; // [2] commit: S_x holds Q K^T of the next block
    uint64_t* s_free = s_full + 2;        // [2] 4 warps: S_x is in registers
    uint64_t* p_full = s_full + 4;        // [2] 4 warps: P_x is in tensor memory
    uint64_t* pv_done = s_full + 6;       // [2] commit: P.V of tile x retired (P buffer free, O_x updated)
    uint64_t* o_free = s_full + 8;        // [2] 4 warps: O_x of the finished item is in registers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 1);
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 4);
            mbar_init(&p_full[i], 4);
            mbar_init(&pv_done[i], 1);
            mbar_init(&o_free[i], 4);
        }
        for (int i = 0; i < KS; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 80;");
        if (warp == 9) {
            if (lane == 0) {
                // ------------------------------------------------------------ TMA producer (runs ahead across items)
                auto load_tile = [&](const CUtensorMap* m64, const CUtensorMap* m16, uint64_t* bar, uint8_t* dst, int col, int row) {
#pragma unroll
                    for (int c = 0; c < C1::NCH; ++c) tma_load_2d(m64, bar, dst + c * 16384, col + c * 64, row);
                    if (C1::HAS16) tma_load_2d(m16, bar, dst + C1::NCH * 16384, col + C1::NCH * 64, row);
                };
                uint32_t n_items = 0, n_kv = 0;
                for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
                    const Att4Item it = att4_item(a, w, nqp);
                    if (!it.valid) continue;
                    const int qcol = a.q_col0 + it.head * HS, kcol = a.k_col0 + it.head * HS, vcol = a.v_col0 + it.head * HS;
                    const uint32_t qb = n_items & 1;
                    mbar_wait(&q_empty[qb], ((n_items >> 1) & 1) ^ 1);
                    mbar_expect_tx(&q_full[qb], Cfg::TILE * (it.b_active ? 2 : 1));
                    uint8_t* qdst = smem + Cfg::OFF_Q + qb * 2 * Cfg::TILE;
                    load_tile(&maps.q64, &maps.q16, &q_full[qb], qdst, qcol, it.q_begin + it.q0);
                    if (it.b_active) load_tile(&maps.q64, &maps.q16, &q_full[qb], qdst + Cfg::TILE, qcol, it.q_begin + it.q0 + ATT_BM);
                    for (int j = 0; j < it.nkt; ++j, ++n_kv) {
                        const uint32_t st = n_kv % KS, par = (n_kv / KS) & 1;
                        mbar_wait(&k_empty[st], par ^ 1);
                        mbar_expect_tx(&k_full[st], Cfg::TILE);
                        load_tile(&maps.k64, &maps.k16, &k_full[st], smem + Cfg::OFF_K + st * Cfg::TILE, kcol, it.k_begin + j * ATT_BN);
                        mbar_wait(&v_empty[st], par ^ 1);
                        mbar_expect_tx(&v_full[st], Cfg::TILE);
                        load_tile(&maps.v64, &maps.v16, &v_full[st], smem + Cfg::OFF_V + st * Cfg::TILE, vcol, it.k_begin + j * ATT_BN);
                    }
                    ++n_items;
                }
            }
        } else if (warp == 8) {
            // ------------------------------------------------------------ MMA issuer: warp-uniform control flow, one elected
            // lane issues; event driven over {S_x consumed -> QK_x(next), P_x written -> PV_x}
            constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, 1, 0, 0);
            constexpr uint32_t idesc_pv64 = make_idesc_f16(128, 64, 1, 0, 1);
            constexpr uint32_t idesc_pv16 = make_idesc_f16(128, 16, 1, 0, 1);
            const uint64_t hi128 = make_smem_desc(0, 16, 1024, kLayoutSW128);
            const uint64_t hi32 = make_smem_desc(0, 16, 256, kLayoutSW32);
            const uint32_t qbase = smem_u32(smem + Cfg::OFF_Q) >> 4;
            const uint32_t kbase = smem_u32(smem + Cfg::OFF_K) >> 4, vbase = smem_u32(smem + Cfg::OFF_V) >> 4;
            constexpr uint32_t TILE16 = Cfg::TILE >> 4;
            const uint32_t tP = tmem_base + Cfg::T_P;
            auto issue_qk = [&](uint32_t q16, uint32_t k16, uint32_t d_tmem) {
                uint32_t acc = 0;
#pragma unroll
                for (int c = 0; c < C1::NCH; ++c)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        umma_f16_ss(d_tmem, hi128 | (q16 + c * 1024 + kk * 2), hi128 | (k16 + c * 1024 + kk * 2), idesc_qk, acc);
                        acc = 1;
                    }
                if (C1::HAS16) umma_f16_ss(d_tmem, hi32 | (q16 + C1::NCH * 1024), hi32 | (k16 + C1::NCH * 1024), idesc_qk, acc);
            };
            auto issue_pv = [&](uint32_t v16, uint32_t d_tmem, uint32_t first_block) {
#pragma unroll
                for (int kk = 0; kk < ATT_BN / 16; ++kk) {
                    const uint32_t pa_t = tP + kk * 8;  // 16 keys = 8 packed 32-bit columns per k-step
                    const uint32_t accum = (kk != 0 || !first_block) ? 1u : 0u;
#pragma unroll
                    for (int c = 0; c < C1::NCH; ++c)
                        umma_f16_ts(d_tmem + c * 64, pa_t, hi128 | (v16 + c * 1024 + kk * 128), idesc_pv64, accum);
                    if (C1::HAS16)
                        umma_f16_ts(d_tmem + C1::NCH * 64, pa_t, hi32 | (v16 + C1::NCH * 1024 + kk * 32), idesc_pv16, accum);
                }
            };
            uint32_t n_items = 0, n_kv = 0;        // n_kv: ring uses before this item
            uint32_t cs[2] = {0, 0};               // QKs issued per tile (all items)
            uint32_t cp[2] = {0, 0};               // PVs issued per tile (all items)
            uint32_t ci[2] = {0, 0};               // finished items per tile (o_free phases)
            for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
                const Att4Item it = att4_item(a, w, nqp);
                if (!it.valid) continue;
                const uint32_t qb = n_items & 1;
                const int nt = it.b_active ? 2 : 1;
                const int nkt = it.nkt;
                int jq[2] = {0, 0}, jp[2] = {0, 0};
                if (!it.b_active) { jq[1] = nkt; jp[1] = nkt; }
                mbar_wait(&q_full[qb], (n_items >> 1) & 1);
                int idle = 0;
                while (jp[0] < nkt || jp[1] < nkt) {
                    bool progress = false;
#pragma unroll
                    for (int x = 0; x < 2; ++x) {
                        if (x >= nt) continue;
                        // ---- S_x = Q_x K_j^T as soon as S_x(j-1) sits in the softmax warps' registers
                        if (jq[x] < nkt) {
                            const uint32_t kv = n_kv + jq[x], st = kv % KS;
                            const bool ready = (cs[x] == 0 || mbar_test(&s_free[x], (cs[x] - 1) & 1)) &&
                                               mbar_test(&k_full[st], (kv / KS) & 1);
                            if (ready) {
                                tc_fence_after();
                                const int j = jq[x];
                                const bool release_k = jq[x ^ 1] > j;  // the other tile already used (or never uses) this K stage
                                if (elect_one()) {
#if VR_A_TRACE
                                    if (blockIdx.x == 0 && cs[x] < 96) g_att_trace[2][cs[x]][x * 2] = clock64();
#endif
                                    issue_qk(qbase + (qb * 2 + x) * TILE16, kbase + st * TILE16, tmem_base + Cfg::T_S + x * 128);
                                    umma_commit(&s_full[x]);
                                    if (release_k) umma_commit(&k_empty[st]);
                                    if (release_k && j == nkt - 1) umma_commit(&q_empty[qb]);
                                }
                                __syncwarp();
                                ++cs[x];
                                ++jq[x];
                                progress = true;
                            }
                        }
                        // ---- O_x (+)= P_x V_j once P_x(j) is in tensor memory
                        if (jp[x] < jq[x]) {
                            const uint32_t kv = n_kv + jp[x], st = kv % KS;
                            const bool ready = mbar_test(&p_full[x], cp[x] & 1) && mbar_test(&v_full[st], (kv / KS) & 1) &&
                                               (jp[x] > 0 || ci[x] == 0 || mbar_test(&o_free[x], (ci[x] - 1) & 1));
                            if (ready) {
                                tc_fence_after();
                                const int j = jp[x];
                                const bool release_v = jp[x ^ 1] > j;
                                if (elect_one()) {
#if VR_A_TRACE
                                    if (blockIdx.x == 0 && cp[x] < 96) g_att_trace[2][cp[x]][x * 2 + 1] = clock64();
#endif
                                    issue_pv(vbase + st * TILE16, tmem_base + Cfg::T_O + x * HS, j == 0);
                                    umma_commit(&pv_done[x]);
                                    if (release_v) umma_commit(&v_empty[st]);
                                }
                                __syncwarp();
                                ++cp[x];
                                ++jp[x];
                                progress = true;
                            }
                        }
                    }
                    if (!progress) {
                        if (++idle > 4) __nanosleep(VR_A_SLEEP);
                        if (idle > (1 << 24)) __trap();  // a pipeline bug must not hang the GPU box
                    } else {
                        idle = 0;
                    }
                }
                n_kv += nkt;
                ++n_items;
                ++ci[0];
                if (it.b_active) ++ci[1];
            }
        }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
        // ---------------------------------------------------------------- softmax warpgroups (thread = query row)
        const int x = warp >> 2;  // 0 = tile A, 1 = tile B
        const int r = threadIdx.x & 127;
        const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t tmem_s = tmem_base + Cfg::T_S + x * 128 + lane_off;
        const uint32_t tmem_o = tmem_base + Cfg::T_O + x * HS + lane_off;
        const uint32_t tmem_p = tmem_base + Cfg::T_P + lane_off;
        const float sl2 = a.scale_log2;
        constexpr float RESCALE_LOG2 = 8.0f;
        uint32_t base_a = 0, base_b = 0; // P.V MMAs of tile A / B before the current item
        uint32_t n_mine = 0;             // S blocks this tile consumed so far (s_full phases)
        int prev_tile = -1;              // tile of the last use of the shared P buffer before the current item
        for (int w = blockIdx.x; w < total_items; w += gridDim.x) {
            const Att4Item it = att4_item(a, w, nqp);
            if (!it.valid) continue;
            const int nkt = it.nkt;
            if (x == 0 || it.b_active) {
                const int q_idx = it.q0 + x * ATT_BM + r;
                float m_ref = -INFINITY, l_run = 0.f;
                for (int kt = 0; kt < nkt; ++kt, ++n_mine) {
                    const int limit = it.len_k - kt * ATT_BN;
                    const bool full = limit >= ATT_BN;
                    VR_TR(x, n_mine, 0);
                    mbar_wait(&s_full[x], n_mine & 1);
                    tc_fence_after();
                    VR_TR(x, n_mine, 1);
                    uint32_t sv[4][32];
                    tmem_ld_32x32(tmem_s, sv[0]);
                    tmem_ld_32x32(tmem_s + 32, sv[1]);
                    tmem_ld_wait();
                    tmem_ld_32x32(tmem_s + 64, sv[2]);  // second half in flight while the first half's maximum is taken
                    tmem_ld_32x32(tmem_s + 96, sv[3]);
                    float m_tile;
                    {
                        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
                        if (full) {
#pragma unroll
                            for (int c = 0; c < 2; ++c)
#pragma unroll
                                for (int j = 0; j < 32; j += 8) {
                                    m0 = max3(m0, __uint_as_float(sv[c][j]), __uint_as_float(sv[c][j + 1]));
                                    m1 = max3(m1, __uint_as_float(sv[c][j + 2]), __uint_as_float(sv[c][j + 3]));
                                    m2 = max3(m2, __uint_as_float(sv[c][j + 4]), __uint_as_float(sv[c][j + 5]));
                                    m3 = max3(m3, __uint_as_float(sv[c][j + 6]), __uint_as_float(sv[c][j + 7]));
                                }
                        }
                        tmem_ld_wait();
                        VR_TR(x, n_mine, 2);
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive(&s_free[x]);  // QK of the next key block may overwrite S_x now
                        if (full) {
#pragma unroll
                            for (int c = 2; c < 4; ++c)
#pragma unroll
                                for (int j = 0; j < 32; j += 8) {
                                    m0 = max3(m0, __uint_as_float(sv[c][j]), __uint_as_float(sv[c][j + 1]));
                                    m1 = max3(m1, __uint_as_float(sv[c][j + 2]), __uint_as_float(sv[c][j + 3]));
                                    m2 = max3(m2, __uint_as_float(sv[c][j + 4]), __uint_as_float(sv[c][j + 5]));
                                    m3 = max3(m3, __uint_as_float(sv[c][j + 6]), __uint_as_float(sv[c][j + 7]));
                                }
                        } else {
#pragma unroll
                            for (int c = 0; c < 4; ++c)
#pragma unroll
                                for (int j = 0; j < 32; ++j)
                                    if (c * 32 + j < limit) m0 = fmaxf(m0, __uint_as_float(sv[c][j]));
                        }
                        m_tile = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                    }
                    // The use of the shared P buffer that precedes this one: its P.V must have retired before P is
                    // overwritten. That MMA was issued after this tile's own previous P.V, so the same wait also makes
                    // O_x safe to rescale.
                    int pred_tile;
                    uint32_t pred_idx;
                    if (x == 1) {
                        pred_tile = 0; pred_idx = base_a + kt;
                    } else if (kt > 0) {
                        pred_tile = it.b_active ? 1 : 0; pred_idx = (it.b_active ? base_b : base_a) + kt - 1;
                    } else {
                        pred_tile = prev_tile; pred_idx = (prev_tile == 1 ? base_b : base_a) - 1;
                    }
                    bool waited = false;
                    const bool grow = (m_tile - m_ref) * sl2 > RESCALE_LOG2;
                    if (kt == 0) {
                        m_ref = (m_tile == -INFINITY) ? 0.f : m_tile;
                    } else if (__any_sync(0xffffffffu, grow)) {
                        mbar_wait(&pv_done[pred_tile], pred_idx & 1);
                        tc_fence_after();
                        waited = true;
                        const float alpha = grow ? ex2_approx((m_ref - m_tile) * sl2) : 1.0f;
                        if (grow) {
                            m_ref = m_tile;
                            l_run *= alpha;
                        }
#pragma unroll
                        for (int c = 0; c < HS / 32; ++c) {
                            uint32_t v[32];
                            tmem_ld_32x32(tmem_o + c * 32, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * alpha);
                            tmem_st_32x32(tmem_o + c * 32, v);
                        }
                        if (HS % 32 == 16) {
                            uint32_t v[16];
                            tmem_ld_32x16(tmem_o + (HS / 32) * 32, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * alpha);
                            tmem_st_32x16(tmem_o + (HS / 32) * 32, v);
                        }
                        tmem_st_wait();
                    }
                    const float neg_ms = -m_ref * sl2;
                    float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
                    uint32_t pk[4][16];
                    VR_TR(x, n_mine, 3);
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float p[32];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            float t0, t1;
                            fma2(t0, t1, __uint_as_float(sv[c][j]), __uint_as_float(sv[c][j + 1]), sl2, sl2, neg_ms, neg_ms);
                            p[j] = ex2_approx(t0);
                            p[j + 1] = ex2_approx(t1);
                        }
                        if (!full) {
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c * 32 + j >= limit) p[j] = 0.f;
                        }
                        if (!ONES) {
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                l0 += p[j];
                                l1 += p[j + 1];
                                l2 += p[j + 2];
                                l3 += p[j + 3];
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[c][j] = pack_bf16x2(p[2 * j], p[2 * j + 1]);
                    }
                    VR_TR(x, n_mine, 4);
                    if (!waited && pred_tile >= 0) {
                        mbar_wait(&pv_done[pred_tile], pred_idx & 1);
                        tc_fence_after();
                    }
                    VR_TR(x, n_mine, 5);
#pragma unroll
                    for (int c = 0; c < 4; ++c) tmem_st_32x16(tmem_p + c * 16, pk[c]);
                    if (!ONES) l_run += (l0 + l1) + (l2 + l3);
                    tmem_st_wait();
                    VR_TR(x, n_mine, 6);
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_full[x]);
                }
                // ---- O / l : read the accumulator once, after the item's last P.V
                mbar_wait(&pv_done[x], ((x ? base_b : base_a) + nkt - 1) & 1);
                tc_fence_after();
                {
                    uint32_t o[HS];
#pragma unroll
                    for (int c = 0; c < HS / 32; ++c) tmem_ld_32x32(tmem_o + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&o[c * 32]));
                    if (HS % 32 == 16) tmem_ld_32x16(tmem_o + (HS / 32) * 32, *reinterpret_cast<uint32_t(*)[16]>(&o[(HS / 32) * 32]));
                    tmem_ld_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&o_free[x]);  // the next item's first P.V may overwrite O_x
                    const float l = ONES ? __uint_as_float(o[HS - 8]) : l_run;  // ONES: head_dim == HS - 8 (checked by the launcher)
                    const float inv = 1.0f / l;
                    const long long row = a.cu_q ? (long long)(it.q_begin + q_idx) : (long long)it.b * a.max_q + q_idx;
                    const bool tile_full = it.q0 + x * ATT_BM + ATT_BM <= it.len_q;  // uniform over the tile's four warps
                    if (a.tma_out && tile_full) {
                        // Output through shared memory + ONE bulk tensor store per tile: per-thread 16-byte stores put 32
                        // different rows into every warp instruction (2304 scattered L2 requests per work item; the ablation
                        // without any output stores ran 0.868 instead of 0.986 ms).
                        uint8_t* ost = smem + Cfg::OFF_OST + x * Cfg::TILE;
                        if (r == 0) bulk_wait_read0();  // the previous item's store has finished reading this buffer
                        asm volatile("bar.sync %0, 128;" ::"r"(1 + x) : "memory");
                        uint4* dsts = reinterpret_cast<uint4*>(ost + r * (a.head_dim * 2));
#pragma unroll
                        for (int j8 = 0; j8 < HS / 8; ++j8) {
                            if (j8 * 8 < a.head_dim) {
                                uint4 v4;
                                v4.x = pack_bf16x2(__uint_as_float(o[j8 * 8 + 0]) * inv, __uint_as_float(o[j8 * 8 + 1]) * inv);
                                v4.y = pack_bf16x2(__uint_as_float(o[j8 * 8 + 2]) * inv, __uint_as_float(o[j8 * 8 + 3]) * inv);
                                v4.z = pack_bf16x2(__uint_as_float(o[j8 * 8 + 4]) * inv, __uint_as_float(o[j8 * 8 + 5]) * inv);
                                v4.w = pack_bf16x2(__uint_as_float(o[j8 * 8 + 6]) * inv, __uint_as_float(o[j8 * 8 + 7]) * inv);
                                dsts[j8] = v4;
                            }
                        }
                        fence_proxy_async_smem();
                        asm volatile("bar.sync %0, 128;" ::"r"(1 + x) : "memory");
                        if (r == 0) tma_store_2d(&maps.o, ost, it.head * a.head_dim, static_cast<int>(row));  // row of r == 0
                    } else {
                        __nv_bfloat16* dst = a.out + row * a.ldo + it.head * a.head_dim;
                        if (q_idx < it.len_q) {
#pragma unroll
                            for (int j8 = 0; j8 < HS / 8; ++j8) {
                                if (j8 * 8 < a.head_dim) {
                                    uint4 v4;
                                    v4.x = pack_bf16x2(__uint_as_float(o[j8 * 8 + 0]) * inv, __uint_as_float(o[j8 * 8 + 1]) * inv);
                                    v4.y = pack_bf16x2(__uint_as_float(o[j8 * 8 + 2]) * inv, __uint_as_float(o[j8 * 8 + 3]) * inv);
                                    v4.z = pack_bf16x2(__uint_as_float(o[j8 * 8 + 4]) * inv, __uint_as_float(o[j8 * 8 + 5]) * inv);
                                    v4.w = pack_bf16x2(__uint_as_float(o[j8 * 8 + 6]) * inv, __uint_as_float(o[j8 * 8 + 7]) * inv);
                                    *reinterpret_cast<uint4*>(dst + j8 * 8) = v4;
                                }
                            }
                        }
                    }
                }
            }
            base_a += nkt;
            if (it.b_active) base_b += nkt;
            prev_tile = it.b_active ? 1 : 0;
        }
        if (r == 0) bulk_wait_read0();  // shared memory must outlive the last bulk store's reads
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace vr
