// Attention v2: two 128-query tiles per CTA in ping-pong (the hot ViT shape: 1024 tokens, 16 heads x 72).
//
// The softmax of a 128x128 score tile needs 16384 exp2 on the SFU (MUFU: 16/clk/SM => 1024 cycles), more than the two
// MMAs of the tile (640 cycles at head stride 80), so the design goal is to keep the SFUs busy all the time:
//   warps 0-3  : softmax of query tile A (thread = one query row, S read from TMEM with tcgen05.ld)
//   warps 4-7  : softmax of query tile B
//   warp  8    : MMA issuer (one lane): S_x = Q_x K_j^T into TMEM, PV_x = P_x V_j into TMEM, x in {A, B}.
//                S_A(j+1) is issued as soon as tile A's softmax(j) has consumed S_A(j), i.e. while tile B is still in
//                its softmax: tensor core and SFU work overlap inside one CTA.
//   warp  9    : TMA producer (one lane): Q_A, Q_B once; K_j / V_j through two-stage full/empty mbarrier rings.
// The O accumulator stays in TMEM for the whole key loop (P.V MMAs accumulate into it); the softmax keeps a lazy
// reference maximum and rescales O in place (tcgen05.ld/st) only when a row maximum grows by more than 2^8, so the
// whole S row fits in registers and is read from TMEM exactly once per key block.
// Per-element ALU cost is cut with the sm_100 packed/3-input forms: FMNMX3 (max), FFMA2 (scale), FADD2 (sum), ex2.approx.
//
// TMEM (512 columns): S_A [0,128)  S_B [128,256)  PV_A [256,256+HS)  Q_A [256+HS, 256+1.5HS)  PV_B [384,384+HS)  Q_B behind it.
#pragma once
#include "attention.cuh"

namespace vr {

#ifndef VR_ATT_ABL
#define VR_ATT_ABL 0  // timing ablations (debug builds only): 1 no exp2, 2 no P stores, 3 no TMEM S loads, 4 no row max
#endif
#ifndef VR_ATT2_SETMAXNREG
#define VR_ATT2_SETMAXNREG 1
#endif
constexpr int ATT2_THREADS = 384;  // 2 softmax warpgroups + 1 control warpgroup (issuer, producer, 2 idle warps)

template <int HS>
struct Att2Cfg {
    using C1 = AttCfg<HS>;
    static_assert(HS == 64 || HS == 80, "v2 is built for head stride 64 / 80");
    static constexpr int TILE = C1::TILE_BYTES;
    static constexpr int OFF_QA = 0;
    static constexpr int OFF_QB = TILE;
    static constexpr int OFF_K = 2 * TILE;   // 2 stages
    static constexpr int OFF_V = 4 * TILE;   // 2 stages
    static constexpr int OFF_BAR = 6 * TILE;
    static constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
};

__device__ __forceinline__ float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float max3(float a, float b, float c) {
    float y;
    asm("max.f32 %0, %1, %2, %3;" : "=f"(y) : "f"(a), "f"(b), "f"(c));
    return y;
}
// (a0,a1) * (b0,b1) + (c0,c1) in one FFMA2
__device__ __forceinline__ void fma2(float& d0, float& d1, float a0, float a1, float b0, float b1, float c0, float c1) {
    unsigned long long A, B, Cc, D;
    asm("mov.b64 %0, {%1, %2};" : "=l"(A) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(B) : "f"(b0), "f"(b1));
    asm("mov.b64 %0, {%1, %2};" : "=l"(Cc) : "f"(c0), "f"(c1));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(D) : "l"(A), "l"(B), "l"(Cc));
    asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(D));
}

// QT: the two Q tiles live in TENSOR MEMORY (written once by the softmax threads, global -> registers -> tcgen05.st,
// packed bf16x2 in the HS/2 columns behind each tile's O accumulator) and S = Q K^T is a TS-MMA. With Q in shared
// memory every Q.K MMA pays the ~128-cycle shared-memory A-operand read (640 cycles per tile and key block). MEASURED:
// QT is slower in situ (1.20 vs 1.11 ms per ViT layer) - the prologue and the extra TMEM traffic outweigh the cheaper
// operand - so the dispatcher uses QT = false; QT = true stays selectable (vr_attention_force_v1(5)) and tested.
template <int HS, bool CAUSAL, bool QT>
__global__ void __launch_bounds__(ATT2_THREADS, 1)
attention2_tcgen05_kernel(const __grid_constant__ AttMaps maps, const AttArgs a) {
    using Cfg = Att2Cfg<HS>;
    using C1 = AttCfg<HS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* q_bar = bars + 0;
    uint64_t* k_full = bars + 1;    // [2]
    uint64_t* k_empty = bars + 3;   // [2]
    uint64_t* v_full = bars + 5;    // [2]
    uint64_t* v_empty = bars + 7;   // [2]
    uint64_t* s_bar = bars + 9;     // [2] per query tile
    uint64_t* p_bar = bars + 11;    // [2]
    uint64_t* o_bar = bars + 13;    // [2]
    uint64_t* q_ready = bars + 15;  // [2] QT only: the 128 rows of a Q tile are in TMEM
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qp = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int k_begin = a.cu_k[b];
    const int len_k = a.cu_k[b + 1] - k_begin;
    const int q_begin = a.cu_q ? a.cu_q[b] : 0;
    const int len_q = a.cu_q ? a.cu_q[b + 1] - q_begin : a.max_q;
    const int q0 = qp * 2 * ATT_BM;
    if (q0 >= len_q || len_k <= 0) return;
    const bool b_active = q0 + ATT_BM < len_q;
    int nkt = (len_k + ATT_BN - 1) / ATT_BN;
    if (CAUSAL) {
        const int last_q = min(q0 + 2 * ATT_BM, len_q) - 1;
        nkt = min(nkt, (last_q + (len_k - len_q)) / ATT_BN + 1);
    }

    if (threadIdx.x == 0) {
        mbar_init(q_bar, 1);
        for (int i = 0; i < 2; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
            mbar_init(&s_bar[i], 1);
            mbar_init(&p_bar[i], 128);
            mbar_init(&o_bar[i], 1);
            mbar_init(&q_ready[i], 128);
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    // Register rebalancing (per-SMSP register files: 12 warps x 168 regs at launch): the control warpgroup gives
    // registers away (168 -> 64), the softmax warpgroups grow (168 -> 208; 64*32 + 2*208*32 = 15360 <= 16384 per SMSP) so that the O row (80 fp32) and two S chunks fit without spills.
    // Each setmaxnreg sits at the top of its own role branch (the allocator applies the limit to the code it dominates).
    if (warp >= 8) {
      // control warpgroup: ONE setmaxnreg executed by all four warps together (it is .sync.aligned per warpgroup)
#if VR_ATT2_SETMAXNREG
      asm volatile("setmaxnreg.dec.sync.aligned.u32 64;");
#endif
      if (warp == 9) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer
            auto load_tile = [&](const CUtensorMap* m64, const CUtensorMap* m16, uint64_t* bar, uint8_t* dst, int col, int row) {
#pragma unroll
                for (int c = 0; c < C1::NCH; ++c) tma_load_2d(m64, bar, dst + c * 16384, col + c * 64, row);
                if (C1::HAS16) tma_load_2d(m16, bar, dst + C1::NCH * 16384, col + C1::NCH * 64, row);
            };
            const int qcol = a.q_col0 + head * HS, kcol = a.k_col0 + head * HS, vcol = a.v_col0 + head * HS;
            if (!QT) {
                mbar_expect_tx(q_bar, Cfg::TILE * (b_active ? 2 : 1));
                load_tile(&maps.q64, &maps.q16, q_bar, smem + Cfg::OFF_QA, qcol, q_begin + q0);
                if (b_active) load_tile(&maps.q64, &maps.q16, q_bar, smem + Cfg::OFF_QB, qcol, q_begin + q0 + ATT_BM);
            }
            for (int j = 0; j < nkt; ++j) {
                const int st = j & 1;
                const uint32_t use_parity = (j >> 1) & 1;
                mbar_wait(&k_empty[st], use_parity ^ 1);
                mbar_expect_tx(&k_full[st], Cfg::TILE);
                load_tile(&maps.k64, &maps.k16, &k_full[st], smem + Cfg::OFF_K + st * Cfg::TILE, kcol, k_begin + j * ATT_BN);
                mbar_wait(&v_empty[st], use_parity ^ 1);
                mbar_expect_tx(&v_full[st], Cfg::TILE);
                load_tile(&maps.v64, &maps.v16, &v_full[st], smem + Cfg::OFF_V + st * Cfg::TILE, vcol, k_begin + j * ATT_BN);
            }
        }
      } else if (warp == 8) {
        // ------------------------------------------------------------ MMA issuer: the whole warp runs the (warp-uniform)
        // control flow, one elected lane issues each block of tcgen05.mma (keeps descriptors in uniform registers)
        constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, 1, 0, 0);
        constexpr uint32_t idesc_pv64 = make_idesc_f16(128, 64, 1, 0, 1);
        constexpr uint32_t idesc_pv16 = make_idesc_f16(128, 16, 1, 0, 1);
        const uint64_t hi128 = make_smem_desc(0, 16, 1024, kLayoutSW128);
        const uint64_t hi32 = make_smem_desc(0, 16, 256, kLayoutSW32);
        const uint32_t qa = smem_u32(smem + Cfg::OFF_QA) >> 4, qb = smem_u32(smem + Cfg::OFF_QB) >> 4;
        const uint32_t kbase = smem_u32(smem + Cfg::OFF_K) >> 4, vbase = smem_u32(smem + Cfg::OFF_V) >> 4;
        constexpr uint32_t TILE16 = Cfg::TILE >> 4;
        // all addresses below are in 16-byte units (the descriptor's address field)
        // q16: Q tile in shared memory (16-byte units) when !QT; with QT the A operand is the tile's TMEM copy
        // (8 packed 32-bit columns per 16-element k-step) at d_tmem's tile: columns 256 + x*128 + HS
        auto issue_qk = [&](uint32_t q16, uint32_t k16, uint32_t d_tmem) {
            const uint32_t qt = d_tmem + 256 + HS;  // d_tmem = tmem_base + x*128
            uint32_t acc = 0;
#pragma unroll
            for (int c = 0; c < C1::NCH; ++c)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    if (QT) umma_f16_ts(d_tmem, qt + (c * 4 + kk) * 8, hi128 | (k16 + c * 1024 + kk * 2), idesc_qk, acc);
                    else umma_f16_ss(d_tmem, hi128 | (q16 + c * 1024 + kk * 2), hi128 | (k16 + c * 1024 + kk * 2), idesc_qk, acc);
                    acc = 1;
                }
            if (C1::HAS16) {
                if (QT) umma_f16_ts(d_tmem, qt + C1::NCH * 32, hi32 | (k16 + C1::NCH * 1024), idesc_qk, acc);
                else umma_f16_ss(d_tmem, hi32 | (q16 + C1::NCH * 1024), hi32 | (k16 + C1::NCH * 1024), idesc_qk, acc);
            }
        };
        // O (+)= P V : O lives in TMEM for the whole key loop; the first key block overwrites, later ones accumulate.
        // P is the A operand and is read from TENSOR MEMORY (it was written there by the softmax warps, packed bf16x2,
        // over the first 64 columns of the tile's own S region): an A operand in shared memory costs 128 rows x 32 B per
        // MMA regardless of N (~128 cycles), which made the 16 small-N P.V MMAs of a key block 4x more expensive than
        // their math (measured: every smem-A variant of this kernel sat at ~1.27 ms per ViT layer).
        auto issue_pv = [&](uint32_t p_tmem, uint32_t v16, uint32_t d_tmem, uint32_t first_block) {
#pragma unroll
            for (int kk = 0; kk < ATT_BN / 16; ++kk) {
                const uint32_t pa_t = p_tmem + kk * 8;  // 16 keys = 8 packed 32-bit columns per k-step
                const uint32_t accum = (kk != 0 || !first_block) ? 1u : 0u;
#pragma unroll
                for (int c = 0; c < C1::NCH; ++c)
                    umma_f16_ts(d_tmem + c * 64, pa_t, hi128 | (v16 + c * 1024 + kk * 128), idesc_pv64, accum);
                if (C1::HAS16)
                    umma_f16_ts(d_tmem + C1::NCH * 64, pa_t, hi32 | (v16 + C1::NCH * 1024 + kk * 32), idesc_pv16, accum);
            }
        };
        const uint32_t tS0 = tmem_base, tS1 = tmem_base + 128, tO0 = tmem_base + 256, tO1 = tmem_base + 384;
        if (QT) {
            mbar_wait(&q_ready[0], 0);
            if (b_active) mbar_wait(&q_ready[1], 0);
        } else {
            mbar_wait(q_bar, 0);
        }
        mbar_wait(&k_full[0], 0);
        tc_fence_after();
        if (elect_one()) {
            issue_qk(qa, kbase, tS0);
            umma_commit(&s_bar[0]);
            if (b_active) {
                issue_qk(qb, kbase, tS1);
                umma_commit(&s_bar[1]);
            }
            umma_commit(&k_empty[0]);
        }
        __syncwarp();
        for (int j = 0; j < nkt; ++j) {
            const int st = j & 1, nst = st ^ 1;
            const uint32_t ph = j & 1, use_parity = (j >> 1) & 1, nuse_parity = ((j + 1) >> 1) & 1;
            const bool more = j + 1 < nkt;
            mbar_wait(&v_full[st], use_parity);
            mbar_wait(&p_bar[0], ph);
            if (more) mbar_wait(&k_full[nst], nuse_parity);
            tc_fence_after();
            if (elect_one()) {
                issue_pv(tS0, vbase + st * TILE16, tO0, j == 0);
                umma_commit(&o_bar[0]);
                if (more) {
                    issue_qk(qa, kbase + nst * TILE16, tS0);
                    umma_commit(&s_bar[0]);
                }
            }
            __syncwarp();
            if (b_active) {
                mbar_wait(&p_bar[1], ph);
                tc_fence_after();
            }
            if (elect_one()) {
                if (b_active) {
                    issue_pv(tS1, vbase + st * TILE16, tO1, j == 0);
                    umma_commit(&o_bar[1]);
                }
                umma_commit(&v_empty[st]);
                if (more) {
                    if (b_active) {
                        issue_qk(qb, kbase + nst * TILE16, tS1);
                        umma_commit(&s_bar[1]);
                    }
                    umma_commit(&k_empty[nst]);
                }
            }
            __syncwarp();
        }
      }
    } else {
#if VR_ATT2_SETMAXNREG
        asm volatile("setmaxnreg.inc.sync.aligned.u32 208;");
#endif
        // ---------------------------------------------------------------- softmax warpgroups
        const int x = warp >> 2;  // 0 = tile A, 1 = tile B
        if (x == 0 || b_active) {
            const int r = threadIdx.x & 127;
            const int q_idx = q0 + x * ATT_BM + r;
            const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
            const uint32_t tmem_s = tmem_base + x * 128 + lane_off;
            const uint32_t tmem_o = tmem_base + 256 + x * 128 + lane_off;
            const int causal_shift = len_k - len_q;
            const float sl2 = a.scale_log2;
            // Online softmax with a LAZY reference maximum (the FlashAttention-4 trick): O accumulates in TMEM and is only
            // rescaled when the row maximum grows by more than 2^RESCALE_LOG2 relative to the reference; otherwise the
            // stale reference is kept (p <= 2^8 stays harmless in bf16/fp32, and O/l is invariant to the reference).
            constexpr float RESCALE_LOG2 = 8.0f;
            float m_ref = -INFINITY, l_run = 0.f;
            if (QT) {  // this row of Q: global -> registers -> TMEM (packed bf16x2); rows past the sequence are zero
                const uint32_t tmem_q = tmem_o + HS;
                const uint4* src = reinterpret_cast<const uint4*>(a.q + static_cast<long long>(q_begin + q_idx) * a.ldq +
                                                                  a.q_col0 + head * HS);
                const bool valid = q_idx < len_q;
                uint32_t w[32];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const uint4 t = valid ? src[i] : make_uint4(0, 0, 0, 0);
                    w[4 * i] = t.x; w[4 * i + 1] = t.y; w[4 * i + 2] = t.z; w[4 * i + 3] = t.w;
                }
                tmem_st_32x32(tmem_q, w);
                if (HS == 80) {
                    uint32_t w2[8];
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const uint4 t = valid ? src[8 + i] : make_uint4(0, 0, 0, 0);
                        w2[4 * i] = t.x; w2[4 * i + 1] = t.y; w2[4 * i + 2] = t.z; w2[4 * i + 3] = t.w;
                    }
                    tmem_st_32x8(tmem_q + 32, w2);
                }
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&q_ready[x]);
            }

            for (int kt = 0; kt < nkt; ++kt) {
                const uint32_t ph = kt & 1;
                const int key0 = kt * ATT_BN;
                int limit = len_k - key0;
                if (CAUSAL) limit = min(limit, q_idx + causal_shift - key0 + 1);
                const bool full = limit >= ATT_BN;
                mbar_wait(&s_bar[x], ph);
                tc_fence_after();
                // the whole S row (128 fp32) comes to registers in one go: O is in TMEM, so there is room
                uint32_t sv[4][32];
#if VR_ATT_ABL == 3
#pragma unroll
                for (int c = 0; c < 4; ++c)
#pragma unroll
                    for (int j = 0; j < 32; ++j) sv[c][j] = __float_as_uint(0.01f * (c * 32 + j + kt));
#else
                tmem_ld_32x32(tmem_s, sv[0]);
                tmem_ld_32x32(tmem_s + 32, sv[1]);
                tmem_ld_32x32(tmem_s + 64, sv[2]);
                tmem_ld_32x32(tmem_s + 96, sv[3]);
                tmem_ld_wait();
#endif
                float m_tile;
                {
                    float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
                    if (full) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
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
                // previous P.V must have retired before P (TMEM, over this tile's S columns) is overwritten and before O is rescaled
                if (kt > 0) {
                    mbar_wait(&o_bar[x], ph ^ 1);
                    tc_fence_after();
                }
                // lazy rescale: only when some row of this warp outgrew its reference by more than 2^8
                const bool grow = (m_tile - m_ref) * sl2 > RESCALE_LOG2;  // false for m_tile = -inf; true for m_ref = -inf
                if (kt == 0) {
                    m_ref = (m_tile == -INFINITY) ? 0.f : m_tile;
                } else if (__any_sync(0xffffffffu, grow)) {
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
                // p = 2^(s*scale*log2e - m_ref*scale*log2e), packed bf16x2 into tensor memory (the A operand of the P.V TS-MMA)
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    float p[32];
#pragma unroll
                    for (int j = 0; j < 32; j += 2) {
                        float t0, t1;
                        fma2(t0, t1, __uint_as_float(sv[c][j]), __uint_as_float(sv[c][j + 1]), sl2, sl2, neg_ms, neg_ms);
#if VR_ATT_ABL == 1
                        p[j] = t0 * 0.001f;
                        p[j + 1] = t1 * 0.001f;
#else
                        p[j] = ex2_approx(t0);
                        p[j + 1] = ex2_approx(t1);
#endif
                    }
                    if (!full) {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            if (c * 32 + j >= limit) p[j] = 0.f;
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        l0 += p[j];
                        l1 += p[j + 1];
                        l2 += p[j + 2];
                        l3 += p[j + 3];
                    }
                    // packed bf16x2 -> TMEM columns [16c, 16c+16) of this tile's S region (the row's S values are already
                    // in registers, and the tensor core only overwrites S again after P.V has consumed P)
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(p[2 * j], p[2 * j + 1]);
                    tmem_st_32x16(tmem_s + c * 16, pk);
                }
                l_run += (l0 + l1) + (l2 + l3);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_bar[x]);
            }
            // ---- O / l : read the accumulator once, after the last P.V
            mbar_wait(&o_bar[x], (nkt - 1) & 1);
            tc_fence_after();
            {  // TMEM loads are warp collective: every lane reads, only valid rows store
                const float inv = 1.0f / l_run;
                const long long row = a.cu_q ? (long long)(q_begin + q_idx) : (long long)b * a.max_q + q_idx;
                __nv_bfloat16* dst = a.out + row * a.ldo + head * a.head_dim;
                const bool valid = q_idx < len_q;
#pragma unroll
                for (int c = 0; c < HS / 32; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(tmem_o + c * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j8 = 0; j8 < 4; ++j8) {
                        if (valid && c * 32 + j8 * 8 < a.head_dim) {
                            uint4 pk;
                            pk.x = pack_bf16x2(__uint_as_float(v[j8 * 8 + 0]) * inv, __uint_as_float(v[j8 * 8 + 1]) * inv);
                            pk.y = pack_bf16x2(__uint_as_float(v[j8 * 8 + 2]) * inv, __uint_as_float(v[j8 * 8 + 3]) * inv);
                            pk.z = pack_bf16x2(__uint_as_float(v[j8 * 8 + 4]) * inv, __uint_as_float(v[j8 * 8 + 5]) * inv);
                            pk.w = pack_bf16x2(__uint_as_float(v[j8 * 8 + 6]) * inv, __uint_as_float(v[j8 * 8 + 7]) * inv);
                            *reinterpret_cast<uint4*>(dst + c * 32 + j8 * 8) = pk;
                        }
                    }
                }
                if (HS % 32 == 16) {
                    uint32_t v[16];
                    tmem_ld_32x16(tmem_o + (HS / 32) * 32, v);
                    tmem_ld_wait();
#pragma unroll
                    for (int j8 = 0; j8 < 2; ++j8) {
                        if (valid && (HS / 32) * 32 + j8 * 8 < a.head_dim) {
                            uint4 pk;
                            pk.x = pack_bf16x2(__uint_as_float(v[j8 * 8 + 0]) * inv, __uint_as_float(v[j8 * 8 + 1]) * inv);
                            pk.y = pack_bf16x2(__uint_as_float(v[j8 * 8 + 2]) * inv, __uint_as_float(v[j8 * 8 + 3]) * inv);
                            pk.z = pack_bf16x2(__uint_as_float(v[j8 * 8 + 4]) * inv, __uint_as_float(v[j8 * 8 + 5]) * inv);
                            pk.w = pack_bf16x2(__uint_as_float(v[j8 * 8 + 6]) * inv, __uint_as_float(v[j8 * 8 + 7]) * inv);
                            *reinterpret_cast<uint4*>(dst + (HS / 32) * 32 + j8 * 8) = pk;
                        }
                    }
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace vr
