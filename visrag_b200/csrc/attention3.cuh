// Attention for sequences longer than one query tile (the hot ViT shape: 1024 tokens, 16 heads x 72): fully pipelined.
//
// Measurements on B200 that shaped this kernel (profiles/r01_ncu_attention*.txt, ablation runs): with 128-key blocks
// and ONE S buffer per query tile, a tile's softmax cannot start on block j+1 before its own P.V(j) + Q.K(j+1) MMAs
// have retired, and the two query tiles of a CTA drift into lock-step (both in softmax, then both waiting for the
// tensor core): 4900 cycles per 2x128x128 block although the SFU needs 2048 and the tensor core 1664. Here the key
// block is 64 wide, which lets every query tile own TWO S buffers and TWO P buffers in the same TMEM / shared memory
// budget: Q.K(j+1) runs while softmax(j) is still working, P.V(j) runs behind softmax(j+1), and the softmax warps
// never wait for the tensor core in steady state.
//
// Both MMA A operands live in TENSOR MEMORY: an A operand in shared memory costs 128 rows x 32 B of smem reads per MMA
// whatever N is (~128 cycles), which dominated the small-N MMAs of attention; from TMEM the cost follows N.
//   * Q_x (128 x HS bf16 = HS/2 packed columns) is written once by the softmax warps (global -> registers -> tcgen05.st);
//   * P_x(j) overwrites the first 32 columns of its own S buffer (the S row is already in registers by then).
//
//   warps 0-3 / 4-7 : softmax of query tile A / B (thread = one query row; S row of 64 fp32 read once from TMEM;
//                     lazy-maximum online softmax; O accumulates in TMEM and is rescaled in place only when a row
//                     maximum grows by more than 2^8)
//   warp 8          : MMA issuer (warp-uniform control flow, one elected lane issues)
//   warp 9          : TMA producer: K_j / V_j (64 keys) through 4-stage full/empty mbarrier rings
//
// TMEM (496 of 512 columns): S_x[b] at (2x+b)*64 | O_A 256 | O_B 336 | Q_A 416 | Q_B 456.
// smem: K ring x4 | V ring x4 | barriers.
#pragma once
#include "attention2.cuh"

namespace vr {

constexpr int ATT3_THREADS = 320;
constexpr int ATT3_BN = 64;
constexpr int ATT3_STAGES = 4;

template <int HS>
struct Att3Cfg {
    using C1 = AttCfg<HS>;
    static_assert(HS == 64 || HS == 80, "built for head stride 64 / 80");
    static constexpr int KTILE = C1::NCH * 8192 + (C1::HAS16 ? 2048 : 0);          // 64 rows
    static constexpr int OFF_K = 0;
    static constexpr int OFF_V = OFF_K + ATT3_STAGES * KTILE;
    static constexpr int OFF_BAR = OFF_V + ATT3_STAGES * KTILE;
    static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
    static constexpr int TM_O = 256, TM_Q = 256 + 2 * HS;                          // TMEM columns
    static_assert(TM_Q + HS <= 512, "TMEM budget");
    static_assert(OFF_K % 1024 == 0 && OFF_V % 1024 == 0 && KTILE % 1024 == 0, "swizzle alignment");
};

struct AttMaps3 {
    CUtensorMap k64, k16;      // 64-row boxes
    CUtensorMap v64, v16;      // 64-row boxes
};

template <int HS, bool CAUSAL>
__global__ void __launch_bounds__(ATT3_THREADS, 1)
attention3_tcgen05_kernel(const __grid_constant__ AttMaps3 maps, const AttArgs a) {
    using Cfg = Att3Cfg<HS>;
    using C1 = AttCfg<HS>;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* q_ready = bars + 0;                   // [x]     Q_x written to TMEM (128 arrivals)
    uint64_t* k_full = bars + 2;                    // [4]
    uint64_t* k_empty = bars + 6;                   // [4]
    uint64_t* v_full = bars + 10;                   // [4]
    uint64_t* v_empty = bars + 14;                  // [4]
    uint64_t* s_full = bars + 18;                   // [x][b]  S_x[b] holds Q_x K_j^T
    uint64_t* p_full = bars + 22;                   // [x][b]  P_x(j) written over S_x[b] (128 arrivals)
    uint64_t* pv_done = bars + 26;                  // [x][b]  P.V that read P_x[b] retired (O up to date, S_x[b] reusable)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 30);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int qp = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const int k_begin = a.cu_k[b];
    const int len_k = a.cu_k[b + 1] - k_begin;
    const int q_begin = a.cu_q ? a.cu_q[b] : 0;
    const int len_q = a.cu_q ? a.cu_q[b + 1] - q_begin : a.max_q;
    const int q0 = qp * 2 * ATT_BM;
    if (q0 >= len_q || len_k <= 0) return;
    const bool b_active = q0 + ATT_BM < len_q;
    int nkt = (len_k + ATT3_BN - 1) / ATT3_BN;
    if (CAUSAL) {
        const int last_q = min(q0 + 2 * ATT_BM, len_q) - 1;
        nkt = min(nkt, (last_q + (len_k - len_q)) / ATT3_BN + 1);
    }

    if (threadIdx.x == 0) {
        mbar_init(&q_ready[0], 128);
        mbar_init(&q_ready[1], 128);
        for (int i = 0; i < ATT3_STAGES; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 1);
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 1);
        }
        for (int i = 0; i < 4; ++i) {
            mbar_init(&s_full[i], 1);
            mbar_init(&p_full[i], 128);
            mbar_init(&pv_done[i], 1);
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc<512>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 9) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer
            auto load_tile = [&](const CUtensorMap* m64, const CUtensorMap* m16, uint64_t* bar, uint8_t* dst, int chunk_bytes,
                                 int col, int row) {
#pragma unroll
                for (int c = 0; c < C1::NCH; ++c) tma_load_2d(m64, bar, dst + c * chunk_bytes, col + c * 64, row);
                if (C1::HAS16) tma_load_2d(m16, bar, dst + C1::NCH * chunk_bytes, col + C1::NCH * 64, row);
            };
            const int kcol = a.k_col0 + head * HS, vcol = a.v_col0 + head * HS;
            for (int j = 0; j < nkt; ++j) {
                const int st = j % ATT3_STAGES;
                const uint32_t use_parity = (j / ATT3_STAGES) & 1;
                mbar_wait(&k_empty[st], use_parity ^ 1);
                mbar_expect_tx(&k_full[st], Cfg::KTILE);
                load_tile(&maps.k64, &maps.k16, &k_full[st], smem + Cfg::OFF_K + st * Cfg::KTILE, 8192, kcol, k_begin + j * ATT3_BN);
                mbar_wait(&v_empty[st], use_parity ^ 1);
                mbar_expect_tx(&v_full[st], Cfg::KTILE);
                load_tile(&maps.v64, &maps.v16, &v_full[st], smem + Cfg::OFF_V + st * Cfg::KTILE, 8192, vcol, k_begin + j * ATT3_BN);
            }
        }
    } else if (warp == 8) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc_qk = make_idesc_f16(128, 64, 1, 0, 0);
        constexpr uint32_t idesc_pv64 = make_idesc_f16(128, 64, 1, 0, 1);
        constexpr uint32_t idesc_pv16 = make_idesc_f16(128, 16, 1, 0, 1);
        const uint64_t hi128 = make_smem_desc(0, 16, 1024, kLayoutSW128);
        const uint64_t hi32 = make_smem_desc(0, 16, 256, kLayoutSW32);
        const uint32_t kbase = smem_u32(smem + Cfg::OFF_K) >> 4, vbase = smem_u32(smem + Cfg::OFF_V) >> 4;
        constexpr uint32_t KT16 = Cfg::KTILE >> 4;
        // S_x[buf] = Q_x K_j^T : A = Q_x from TMEM (8 packed columns per 16-element k-step), B = K chunk (64 keys, K-major)
        auto issue_qk = [&](int x, int j) {
            const uint32_t qt = tmem_base + Cfg::TM_Q + x * (HS / 2), ka = kbase + (j % ATT3_STAGES) * KT16;
            const uint32_t d_tmem = tmem_base + (2 * x + (j & 1)) * 64;
            uint32_t acc = 0;
#pragma unroll
            for (int c = 0; c < C1::NCH; ++c)
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    umma_f16_ts(d_tmem, qt + (c * 4 + kk) * 8, hi128 | (ka + c * 512 + kk * 2), idesc_qk, acc);
                    acc = 1;
                }
            if (C1::HAS16) umma_f16_ts(d_tmem, qt + C1::NCH * 32, hi32 | (ka + C1::NCH * 512), idesc_qk, acc);
            umma_commit(&s_full[2 * x + (j & 1)]);
        };
        // O_x (+)= P_x(j) V_j : A = P from TMEM (first 32 columns of S_x[j&1]); V chunk: 64 keys x 64 dims, MN-major
        auto issue_pv = [&](int x, int j) {
            const uint32_t pt = tmem_base + (2 * x + (j & 1)) * 64, va = vbase + (j % ATT3_STAGES) * KT16;
            const uint32_t d_tmem = tmem_base + Cfg::TM_O + x * HS;
#pragma unroll
            for (int kk = 0; kk < ATT3_BN / 16; ++kk) {
                const uint32_t accum = (kk != 0 || j != 0) ? 1u : 0u;
#pragma unroll
                for (int c = 0; c < C1::NCH; ++c)
                    umma_f16_ts(d_tmem + c * 64, pt + kk * 8, hi128 | (va + c * 512 + kk * 128), idesc_pv64, accum);
                if (C1::HAS16) umma_f16_ts(d_tmem + C1::NCH * 64, pt + kk * 8, hi32 | (va + C1::NCH * 512 + kk * 32), idesc_pv16, accum);
            }
            umma_commit(&pv_done[2 * x + (j & 1)]);
        };
        const int ntile = b_active ? 2 : 1;
        for (int x = 0; x < ntile; ++x) mbar_wait(&q_ready[x], 0);
        // prologue: S(0) and S(1) of both tiles
        for (int j = 0; j < 2 && j < nkt; ++j) {
            mbar_wait(&k_full[j % ATT3_STAGES], 0);
            tc_fence_after();
            if (elect_one()) {
                for (int x = 0; x < ntile; ++x) issue_qk(x, j);
                umma_commit(&k_empty[j % ATT3_STAGES]);
            }
            __syncwarp();
        }
        for (int j = 0; j < nkt; ++j) {
            const int st = j % ATT3_STAGES;
            const uint32_t buf_parity = (j >> 1) & 1;
            const bool more = j + 2 < nkt;
            mbar_wait(&v_full[st], (j / ATT3_STAGES) & 1);
            if (more) mbar_wait(&k_full[(j + 2) % ATT3_STAGES], ((j + 2) / ATT3_STAGES) & 1);
            for (int x = 0; x < ntile; ++x) {
                mbar_wait(&p_full[2 * x + (j & 1)], buf_parity);  // P_x(j) ready; S_x[j&1] consumed
                tc_fence_after();
                if (elect_one()) {
                    issue_pv(x, j);
                    if (x == ntile - 1) umma_commit(&v_empty[st]);
                    if (more) {
                        issue_qk(x, j + 2);  // into the S buffer softmax(j) has just released
                        if (x == ntile - 1) umma_commit(&k_empty[(j + 2) % ATT3_STAGES]);
                    }
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------------------------------------------------------- softmax warpgroups
        const int x = warp >> 2;
        if (x == 0 || b_active) {
            const int r = threadIdx.x & 127;
            const int q_idx = q0 + x * ATT_BM + r;
            const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
            const uint32_t tmem_o = tmem_base + Cfg::TM_O + x * HS + lane_off;
            const int causal_shift = len_k - len_q;
            const float sl2 = a.scale_log2;
            constexpr float RESCALE_LOG2 = 8.0f;
            float m_ref = -INFINITY, l_run = 0.f;
            {   // this row of Q: global -> registers -> TMEM (packed bf16x2, HS/2 columns); rows past the sequence are zero
                const uint32_t tmem_q = tmem_base + Cfg::TM_Q + x * (HS / 2) + lane_off;
                const uint4* src = reinterpret_cast<const uint4*>(a.q + static_cast<long long>(q_begin + q_idx) * a.ldq + a.q_col0 + head * HS);
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
                const int buf = kt & 1;
                const uint32_t buf_parity = (kt >> 1) & 1;
                const int key0 = kt * ATT3_BN;
                int limit = len_k - key0;
                if (CAUSAL) limit = min(limit, q_idx + causal_shift - key0 + 1);
                const bool full = limit >= ATT3_BN;
                mbar_wait(&s_full[2 * x + buf], buf_parity);
                tc_fence_after();
                uint32_t sv[2][32];
                const uint32_t tmem_s = tmem_base + (2 * x + buf) * 64 + lane_off;
                tmem_ld_32x32(tmem_s, sv[0]);
                tmem_ld_32x32(tmem_s + 32, sv[1]);
                tmem_ld_wait();
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
                    } else {
#pragma unroll
                        for (int c = 0; c < 2; ++c)
#pragma unroll
                            for (int j = 0; j < 32; ++j)
                                if (c * 32 + j < limit) m0 = fmaxf(m0, __uint_as_float(sv[c][j]));
                    }
                    m_tile = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
                }
                // lazy rescale of O (rare): every earlier P.V must have retired before O is read-modified-written
                const bool grow = (m_tile - m_ref) * sl2 > RESCALE_LOG2;
                if (kt == 0) {
                    m_ref = (m_tile == -INFINITY) ? 0.f : m_tile;
                } else if (__any_sync(0xffffffffu, grow)) {
                    mbar_wait(&pv_done[2 * x + (buf ^ 1)], ((kt - 1) >> 1) & 1);  // P.V(kt-1): MMAs retire in order
                    tc_fence_after();
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
                // P overwrites the first 32 columns of this S buffer: the tensor core last touched it with Q.K(kt) itself.
                // (The P.V of two blocks ago on this buffer has necessarily retired - Q.K(kt) was issued behind it - but its
                // mbarrier phase is still consumed here so that every phase of pv_done is observed in order.)
                if (kt >= 2) mbar_wait(&pv_done[2 * x + buf], buf_parity ^ 1);
                const float neg_ms = -m_ref * sl2;
                float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
#pragma unroll
                for (int c = 0; c < 2; ++c) {
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
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        l0 += p[j];
                        l1 += p[j + 1];
                        l2 += p[j + 2];
                        l3 += p[j + 3];
                    }
                    uint32_t pk[16];
#pragma unroll
                    for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(p[2 * j], p[2 * j + 1]);
                    tmem_st_32x16(tmem_s + c * 16, pk);
                }
                l_run += (l0 + l1) + (l2 + l3);
                tmem_st_wait();
                tc_fence_before();
                mbar_arrive(&p_full[2 * x + buf]);
            }
            // ---- O / l after the last P.V
            mbar_wait(&pv_done[2 * x + ((nkt - 1) & 1)], ((nkt - 1) >> 1) & 1);
            tc_fence_after();
            {
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
