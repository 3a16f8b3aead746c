// Attention v4 (ViT shape: many 1024-token sequences, 16 heads x 72): persistent, decoupled two-tile kernel.
//
// What the engines cost per key block of 128 keys and CTA (two 128-query tiles), measured with tools/ubench/ub_tc.cu:
//   MUFU   32768 ex2 at 16/clk/SM                               2048 cycles   <- the real bound
//   tensor 2 x (5 SS-MMA N=128 at 98 + 8 x (TS N=64 at 41 + TS N=16 at 17))  1900 cycles
//   TMEM   128 KB of S at ~280 B/clk/SM (64 B/clk per sub-partition)          460 cycles
// The v2 kernel (attention2.cuh) needs 3900 cycles per key block: P overlays S, so each tile runs
// QK(j) -> ld S -> max -> exp -> st P -> PV(j) -> QK(j+1) as ONE dependency chain, and every CTA (8 key blocks) pays its
// own prologue (barrier init, TMEM alloc, first TMA round trip) and epilogue. v4 removes both:
//   * P does not overlay S, and every tile owns private room for a WHOLE key block of P, so a P store only ever waits
//     for the P.V of the PREVIOUS key block (issued a full exp phase earlier): TMEM has 512 columns = 2 x 128 (S) +
//     2 x HS (O) + 96, i.e. room for only half a block of P per tile (32 columns = 64 keys, bf16x2); the second half of
//     every block goes through 128B-swizzled SHARED memory and that half of P.V is an SS-MMA (A operand from shared
//     memory: 70 instead of 41 cycles per N=64 MMA - the tensor pipe has the headroom, the SFU does not).
//     S_x is free again as soon as the softmax warps have pulled it into registers, so QK_x(j+1) is issued while the exps
//     of block j are still running: the MMAs leave the softmax critical path and the two tiles never wait for each other.
//     (Measured alternatives, 128 x 16 x 1024 x 72: ONE 64-column TMEM buffer shared by the tiles in strict alternation
//     0.99 ms; a private 32-column buffer per tile reused for both halves of a block 1.05 ms - the store of the second
//     half then waits for the first half's P.V round trip every block; v2 kernel 1.12 ms.)
//   * The CTA is persistent: one CTA per SM loops over (query-tile pair, head, sequence) work items; barriers, TMEM and
//     the K/V ring live across items, the producer runs ahead into the next item (double-buffered Q), and the O read-out
//     of item i overlaps the first QK of item i+1.
//   * One MMA issuer warp per tile, each with its own cursor over the tile's key blocks: the events of one tile arrive
//     in a fixed order, so the issuer follows it with blocking mbarrier waits (no polling); tile A may already be in the
//     next work item while tile B finishes the current one (ring stages and Q buffers are released by two commits, one per
//     tile).
//   * ONES: when the caller guarantees V[:, head_dim] == 1 (a bias in the QKV GEMM's zero padding), column head_dim of O
//     is the softmax denominator, accumulated by the tensor core from exactly the bf16 P values the numerator uses; the
//     128 FADDs per row and key block disappear.
// Warp roles: 0-3 softmax tile A, 4-7 softmax tile B (thread = query row), 8 / 10 MMA issuers of tile A / B, 9 TMA producer.
#pragma once
#include "attention2.cuh"

namespace vr {

constexpr int ATT4_THREADS = 384;

template <int HS>
struct Att4Cfg {
    using C1 = AttCfg<HS>;
    static_assert(HS == 64 || HS == 80, "v4 is built for head stride 64 / 80");
    static constexpr int TILE = C1::TILE_BYTES;
    static constexpr int KSK = 3, KSV = 2;             // K / V ring depths
    static constexpr int OFF_Q = 0;                    // [2 buffers][2 tiles]
    static constexpr int OFF_K = 4 * TILE;
    static constexpr int OFF_V = OFF_K + KSK * TILE;
    static constexpr int OFF_P = OFF_V + KSV * TILE;   // [2 tiles] bf16 P(keys 64-127): 128 rows x 128 B, 128B-swizzled (K-major A operand)
    static constexpr int OFF_BAR = OFF_P + 2 * 16384;
    static constexpr int SMEM_BYTES = OFF_BAR + 512 + 1024;
    static_assert(OFF_P % 1024 == 0, "swizzled tiles must be 1024-byte aligned");
    static constexpr uint32_t T_S = 0, T_O = 256, T_P = 256 + 2 * HS;  // TMEM columns; P_x(keys 0-63) at T_P + 32 x
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

// debug timeline (only in builds with -DVR_ATT4_TRACE=1: the two extra live registers cost spills in the softmax loop):
// lane 0 of every warp of CTA 0 appends (event << 48 | clock)
#ifndef VR_ATT4_TRACE
#define VR_ATT4_TRACE 0
#endif
#ifndef VR_ATT4_ABL
#define VR_ATT4_ABL 0  // timing ablations (experiment builds only, results are wrong): 1 no exp2, 2 no P stores, 4 no S loads, 8 no row max, 16 no MMAs, 32 no smem half of P, 64 no TMEM half of P, 128 no K/V loads after the first three
#endif
#ifndef VR_ATT4_BDELAY
#define VR_ATT4_BDELAY 0  // cycles tile B's softmax warps wait before their first key block
#endif
struct Att4Trace {
#if VR_ATT4_TRACE
    unsigned long long* p;
    int left;
    __device__ __forceinline__ void init(const AttArgs& a, int warp, int lane) {
        const bool on = a.trace != nullptr && blockIdx.x == 0 && lane == 0;
        p = on ? a.trace + static_cast<long long>(warp) * a.trace_cap : nullptr;
        left = on ? a.trace_cap : 0;
    }
    __device__ __forceinline__ void ev(unsigned ev_id) {
        if (left > 0) {
            *p++ = (static_cast<unsigned long long>(ev_id) << 48) | (static_cast<unsigned long long>(clock64()) & 0xFFFFFFFFFFFFull);
            --left;
        }
    }
#else
    __device__ __forceinline__ void init(const AttArgs&, int, int) {}
    __device__ __forceinline__ void ev(unsigned) {}
#endif
};

struct Att4Item {
    int w, head, b, k_begin, len_k, q_begin, len_q, q0, nkt;
    bool valid, b_active;
};

// Descriptor of work item w. Branch free (the loads are clamped in bounds and nothing branches on them), so a caller can
// issue it a whole item ahead and the global-load latency hides under the current item's math. valid == false: the item is
// empty (ragged batch) or w is past the end.
__device__ __forceinline__ Att4Item att4_load(const AttArgs& a, int w, int total, int nqp) {
    Att4Item it;
    const int wc = w < total ? w : total - 1;
    const int qp = wc % nqp;
    const int t = wc / nqp;
    it.w = w;
    it.head = t % a.heads;
    it.b = t / a.heads;
    it.k_begin = a.cu_k[it.b];
    it.len_k = a.cu_k[it.b + 1] - it.k_begin;
    it.q_begin = a.cu_q ? a.cu_q[it.b] : 0;
    it.len_q = a.cu_q ? a.cu_q[it.b + 1] - it.q_begin : a.max_q;
    it.q0 = qp * 2 * ATT_BM;
    it.b_active = it.q0 + ATT_BM < it.len_q;
    it.nkt = (it.len_k + ATT_BN - 1) / ATT_BN;
    it.valid = w < total && it.q0 < it.len_q && it.len_k > 0;
    return it;
}
// first non-empty work item at or after `it` (stride = gridDim.x); valid == false: no more work
__device__ __forceinline__ Att4Item att4_settle(const AttArgs& a, Att4Item it, int total, int nqp, int stride) {
    while (!it.valid && it.w < total) it = att4_load(a, it.w + stride, total, nqp);
    return it;
}
__device__ __forceinline__ Att4Item att4_next(const AttArgs& a, int w, int total, int nqp, int stride) {
    return att4_settle(a, att4_load(a, w, total, nqp), total, nqp, stride);
}

template <int HS, bool ONES>
__global__ void __launch_bounds__(ATT4_THREADS, 1)
attention4_tcgen05_kernel(const __grid_constant__ AttMaps maps, const AttArgs a, const int total_items, const int nqp) {
    using Cfg = Att4Cfg<HS>;
    using C1 = AttCfg<HS>;
    constexpr int KSK = Cfg::KSK, KSV = Cfg::KSV;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BAR);
    uint64_t* q_full = bars + 0;           // [2] TMA bytes of a Q buffer (both tiles)
    uint64_t* q_empty = bars + 2;          // [2] commit: every QK of the item retired
    uint64_t* k_full = bars + 4;                     // [KSK]
    uint64_t* k_empty = bars + 4 + KSK;              // [KSK] commit
    uint64_t* v_full = bars + 4 + 2 * KSK;           // [KSV]
    uint64_t* v_empty = bars + 4 + 2 * KSK + KSV;    // [KSV] commit
    uint64_t* s_full = bars + 4 + 2 * KSK + 2 * KSV; // [2] commit: S_x holds Q K^T of the next block
    uint64_t* s_free = s_full + 2;         // [2] 4 warps: S_x is in registers
    uint64_t* p_full = s_full + 4;         // [2] 4 warps: P_x is written (keys 0-63 in tensor memory, keys 64-127 in shared memory)
    uint64_t* pv_done = s_full + 6;        // [2] commit: P.V of the block retired (both P buffers free, O_x complete for this block)
    uint64_t* o_free = s_full + 8;         // [2] 4 warps: O_x of the finished item is in registers
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(s_full + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stride = gridDim.x;
    Att4Trace tr;
    tr.init(a, warp, lane);

    if (threadIdx.x == 0) {
        for (int i = 0; i < 2; ++i) {
            mbar_init(&q_full[i], 1);
            mbar_init(&q_empty[i], 2);  // one commit per tile (a tile that is alone in an item commits twice)
            mbar_init(&s_full[i], 1);
            mbar_init(&s_free[i], 4);
            mbar_init(&p_full[i], 4);
            mbar_init(&pv_done[i], 1);
            mbar_init(&o_free[i], 4);
        }
        for (int i = 0; i < KSK; ++i) {
            mbar_init(&k_full[i], 1);
            mbar_init(&k_empty[i], 2);
        }
        for (int i = 0; i < KSV; ++i) {
            mbar_init(&v_full[i], 1);
            mbar_init(&v_empty[i], 2);
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
                Att4Item it = att4_next(a, blockIdx.x, total_items, nqp, stride);
                while (it.valid) {
                    const Att4Item nx = att4_load(a, it.w + stride, total_items, nqp);  // descriptor loads overlap the TMA issue
                    const int qcol = a.q_col0 + it.head * HS, kcol = a.k_col0 + it.head * HS, vcol = a.v_col0 + it.head * HS;
                    const uint32_t qb = n_items & 1;
                    mbar_wait(&q_empty[qb], ((n_items >> 1) & 1) ^ 1);
                    mbar_expect_tx(&q_full[qb], Cfg::TILE * (it.b_active ? 2 : 1));
                    uint8_t* qdst = smem + Cfg::OFF_Q + qb * 2 * Cfg::TILE;
                    load_tile(&maps.q64, &maps.q16, &q_full[qb], qdst, qcol, it.q_begin + it.q0);
                    if (it.b_active) load_tile(&maps.q64, &maps.q16, &q_full[qb], qdst + Cfg::TILE, qcol, it.q_begin + it.q0 + ATT_BM);
                    for (int j = 0; j < it.nkt; ++j, ++n_kv) {
                        const uint32_t sk = n_kv % KSK, sv = n_kv % KSV;
                        mbar_wait(&k_empty[sk], ((n_kv / KSK) & 1) ^ 1);
                        if ((VR_ATT4_ABL & 128) && n_kv >= 3) {
                            mbar_arrive(&k_full[sk]);
                        } else {
                            mbar_expect_tx(&k_full[sk], Cfg::TILE);
                            load_tile(&maps.k64, &maps.k16, &k_full[sk], smem + Cfg::OFF_K + sk * Cfg::TILE, kcol, it.k_begin + j * ATT_BN);
                        }
                        mbar_wait(&v_empty[sv], ((n_kv / KSV) & 1) ^ 1);
                        if ((VR_ATT4_ABL & 128) && n_kv >= 3) {
                            mbar_arrive(&v_full[sv]);
                        } else {
                            mbar_expect_tx(&v_full[sv], Cfg::TILE);
                            load_tile(&maps.v64, &maps.v16, &v_full[sv], smem + Cfg::OFF_V + sv * Cfg::TILE, vcol, it.k_begin + j * ATT_BN);
                        }
                    }
                    ++n_items;
                    it = att4_settle(a, nx, total_items, nqp, stride);
                }
            }
        } else if (warp == 8 || warp == 10) {
            // ------------------------------------------------------------ MMA issuers: warp 8 serves tile A, warp 10 tile B.
            // Warp-uniform control flow, one elected lane issues. Every wait is a blocking mbarrier wait (the warp sleeps in
            // hardware). The events of ONE tile arrive in a fixed order - S(j) consumed, P(j) stored - so each issuer simply
            // follows that order:  QK(j+1) | PV(j)
            const int x = warp == 8 ? 0 : 1;
            constexpr uint32_t idesc_qk = make_idesc_f16(128, 128, 1, 0, 0);
            constexpr uint32_t idesc_pv64 = make_idesc_f16(128, 64, 1, 0, 1);
            constexpr uint32_t idesc_pv16 = make_idesc_f16(128, 16, 1, 0, 1);
            const uint64_t hi128 = make_smem_desc(0, 16, 1024, kLayoutSW128);
            const uint64_t hi32 = make_smem_desc(0, 16, 256, kLayoutSW32);
            const uint32_t qbase = smem_u32(smem + Cfg::OFF_Q) >> 4;
            const uint32_t kbase = smem_u32(smem + Cfg::OFF_K) >> 4, vbase = smem_u32(smem + Cfg::OFF_V) >> 4;
            constexpr uint32_t TILE16 = Cfg::TILE >> 4;
            const uint32_t tS = tmem_base + Cfg::T_S + x * 128, tO = tmem_base + Cfg::T_O + x * HS, tP = tmem_base + Cfg::T_P + x * 32;
            auto issue_qk = [&](uint32_t q16, uint32_t k16) {
                uint32_t acc = 0;
#pragma unroll
                for (int c = 0; c < C1::NCH; ++c)
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        umma_f16_ss(tS, hi128 | (q16 + c * 1024 + kk * 2), hi128 | (k16 + c * 1024 + kk * 2), idesc_qk, acc);
                        acc = 1;
                    }
                if (C1::HAS16) umma_f16_ss(tS, hi32 | (q16 + C1::NCH * 1024), hi32 | (k16 + C1::NCH * 1024), idesc_qk, acc);
            };
            // O (+)= P[:, 64*half .. 64*half+63] V[64*half .. , :]  (4 k-steps of 16 keys). First half: P from tensor memory
            // (packed bf16x2 = 8 columns per k-step, TS-MMA); second half: P from 128B-swizzled shared memory (SS-MMA).
            const uint32_t p16 = smem_u32(smem + Cfg::OFF_P + x * 16384) >> 4;
            auto issue_pv_half = [&](uint32_t v16, uint32_t half, uint32_t overwrite) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk) {
                    const uint32_t ks = half * 4 + kk;
                    const uint32_t accum = (kk != 0 || !overwrite) ? 1u : 0u;
                    if (half == 0) {
                        const uint32_t pa_t = tP + kk * 8;
#pragma unroll
                        for (int c = 0; c < C1::NCH; ++c)
                            umma_f16_ts(tO + c * 64, pa_t, hi128 | (v16 + c * 1024 + ks * 128), idesc_pv64, accum);
                        if (C1::HAS16)
                            umma_f16_ts(tO + C1::NCH * 64, pa_t, hi32 | (v16 + C1::NCH * 1024 + ks * 32), idesc_pv16, accum);
                    } else {
                        const uint64_t pd = hi128 | (p16 + kk * 2);
#pragma unroll
                        for (int c = 0; c < C1::NCH; ++c)
                            umma_f16_ss(tO + c * 64, pd, hi128 | (v16 + c * 1024 + ks * 128), idesc_pv64, accum);
                        if (C1::HAS16)
                            umma_f16_ss(tO + C1::NCH * 64, pd, hi32 | (v16 + C1::NCH * 1024 + ks * 32), idesc_pv16, accum);
                    }
                }
            };
            // cursor over the key blocks this tile processes, in order, across work items
            struct Cur {
                Att4Item it;
                int j;          // key block inside the item
                uint32_t n_it;  // index of the item among the CTA's valid items (Q buffer + parity)
                uint32_t kvb;   // ring index of the item's key block 0
            };
            auto settle = [&](Cur& c) {  // skip items in which this tile has no rows (tile B of an odd tile count)
                while (c.it.valid && x == 1 && !c.it.b_active) {
                    c.kvb += c.it.nkt;
                    ++c.n_it;
                    c.it = att4_next(a, c.it.w + stride, total_items, nqp, stride);
                }
            };
            auto advance = [&](Cur& c) {
                if (++c.j == c.it.nkt) {
                    c.j = 0;
                    c.kvb += c.it.nkt;
                    ++c.n_it;
                    c.it = att4_next(a, c.it.w + stride, total_items, nqp, stride);
                    settle(c);
                }
            };
            uint32_t cs = 0, cp = 0, ci = 0;  // QKs / P.Vs issued, items finished (barrier phases)
            auto do_qk = [&](const Cur& c) {
                const uint32_t g = c.kvb + c.j, st = g % KSK, qb = c.n_it & 1;
                if (c.j == 0) mbar_wait(&q_full[qb], (c.n_it >> 1) & 1);
                if (cs > 0) mbar_wait(&s_free[x], (cs - 1) & 1);  // S_x(previous block) sits in the softmax warps' registers
                mbar_wait(&k_full[st], (g / KSK) & 1);
                tc_fence_after();
                const bool solo = !c.it.b_active;  // the other tile does not exist in this item: release on its behalf
                const bool last = c.j == c.it.nkt - 1;
                if (elect_one()) {
                    if (!(VR_ATT4_ABL & 16)) issue_qk(qbase + (qb * 2 + x) * TILE16, kbase + st * TILE16);
                    umma_commit(&s_full[x]);
                    umma_commit(&k_empty[st]);
                    if (solo) umma_commit(&k_empty[st]);
                    if (last) {
                        umma_commit(&q_empty[qb]);
                        if (solo) umma_commit(&q_empty[qb]);
                    }
                }
                __syncwarp();
                tr.ev(0x10 + x);
                ++cs;
            };
            auto do_pv = [&](const Cur& c) {
                const uint32_t g = c.kvb + c.j, st = g % KSV;
                mbar_wait(&p_full[x], cp & 1);
                mbar_wait(&v_full[st], (g / KSV) & 1);
                if (c.j == 0 && ci > 0) mbar_wait(&o_free[x], (ci - 1) & 1);  // O_x of the previous item has been read out
                tc_fence_after();
                const bool solo = !c.it.b_active;
                if (elect_one()) {
                    if (!(VR_ATT4_ABL & 16)) {
                        issue_pv_half(vbase + st * TILE16, 0, c.j == 0 ? 1u : 0u);
                        issue_pv_half(vbase + st * TILE16, 1, 0u);
                    }
                    umma_commit(&pv_done[x]);
                    umma_commit(&v_empty[st]);
                    if (solo) umma_commit(&v_empty[st]);
                }
                __syncwarp();
                tr.ev(0x20 + x * 2);
                ++cp;
                if (c.j == c.it.nkt - 1) ++ci;
            };
            Cur qc;
            qc.it = att4_next(a, blockIdx.x, total_items, nqp, stride);
            qc.j = 0;
            qc.n_it = 0;
            qc.kvb = 0;
            settle(qc);
            Cur pc = qc;
            if (qc.it.valid) {
                do_qk(qc);
                advance(qc);
            }
            while (pc.it.valid) {
                if (qc.it.valid) {  // S_x(j) was consumed right after it arrived: Q.K^T of block j+1 runs under the exps of block j
                    do_qk(qc);
                    advance(qc);
                }
                do_pv(pc);
                advance(pc);
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
        const uint32_t tmem_p = tmem_base + Cfg::T_P + x * 32 + lane_off;
        uint8_t* p_row = smem + Cfg::OFF_P + x * 16384 + (r >> 3) * 1024 + (r & 7) * 128;  // this row of the smem half of P
        const float sl2 = a.scale_log2;
        constexpr float RESCALE_LOG2 = 8.0f;
        uint32_t n_blk = 0;  // key blocks this tile has processed over all items (phase of every per-block barrier)
        // O / l of a finished item: read the accumulator once, scale, store. It runs DEFERRED, inside the first key block of
        // the tile's next item (after that block's exps, before its P is published): the last P.V of the finished item has
        // long retired by then, so the softmax warps never sit out its round trip (~2000 cycles per item before).
        bool pend = false;             // an item's O_x still sits in tensor memory
        __nv_bfloat16* pend_dst = nullptr;
        bool pend_store = false;
        float pend_l = 1.f;
        auto epilogue = [&]() {
            mbar_wait(&pv_done[x], (n_blk - 1) & 1);  // n_blk - 1 = the finished item's last key block
            tc_fence_after();
            tr.ev(11);
            uint32_t o[HS];
#pragma unroll
            for (int c = 0; c < HS / 32; ++c) tmem_ld_32x32(tmem_o + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&o[c * 32]));
            if (HS % 32 == 16) tmem_ld_32x16(tmem_o + (HS / 32) * 32, *reinterpret_cast<uint32_t(*)[16]>(&o[(HS / 32) * 32]));
            tmem_ld_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&o_free[x]);  // the next item's first P.V may overwrite O_x
            const float l = ONES ? __uint_as_float(o[HS - 8]) : pend_l;  // ONES: head_dim == HS - 8 (checked by the launcher)
            const float inv = 1.0f / l;
            if (pend_store) {
#pragma unroll
                for (int j8 = 0; j8 < HS / 8; ++j8) {
                    if (j8 * 8 < a.head_dim) {
                        uint4 v4;
                        v4.x = pack_bf16x2(__uint_as_float(o[j8 * 8 + 0]) * inv, __uint_as_float(o[j8 * 8 + 1]) * inv);
                        v4.y = pack_bf16x2(__uint_as_float(o[j8 * 8 + 2]) * inv, __uint_as_float(o[j8 * 8 + 3]) * inv);
                        v4.z = pack_bf16x2(__uint_as_float(o[j8 * 8 + 4]) * inv, __uint_as_float(o[j8 * 8 + 5]) * inv);
                        v4.w = pack_bf16x2(__uint_as_float(o[j8 * 8 + 6]) * inv, __uint_as_float(o[j8 * 8 + 7]) * inv);
                        *reinterpret_cast<uint4*>(pend_dst + j8 * 8) = v4;
                    }
                }
            }
            pend = false;
            tr.ev(12);
        };
        Att4Item it = att4_next(a, blockIdx.x, total_items, nqp, stride);
#if VR_ATT4_BDELAY > 0
        // De-phase the two tiles once: the two softmax warps of a sub-partition share its SFU. Started together they stay
        // in phase for the whole kernel (both in their exp phase at half speed, then both in their load/max/store phase with
        // the SFU idle); half a block apart, one computes exps while the other fetches S and stores P.
        if (x == 1) {
            const long long t0 = clock64();
            while (clock64() - t0 < VR_ATT4_BDELAY) {}
        }
#endif
        while (it.valid) {
            const Att4Item nx = att4_load(a, it.w + stride, total_items, nqp);  // its loads fly under this item's math
            const int nkt = it.nkt;
            if (x == 0 || it.b_active) {
                const int q_idx = it.q0 + x * ATT_BM + r;
                float m_ref = -INFINITY, l_run = 0.f;
                for (int kt = 0; kt < nkt; ++kt, ++n_blk) {
                    const int limit = it.len_k - kt * ATT_BN;
                    const bool full = limit >= ATT_BN;
                    tr.ev(1);  // block start (waiting for S)
                    mbar_wait(&s_full[x], n_blk & 1);
                    tc_fence_after();
                    tr.ev(2);  // S ready
                    uint32_t sv[4][32];
                    if (VR_ATT4_ABL & 4) {
#pragma unroll
                        for (int c = 0; c < 4; ++c)
#pragma unroll
                            for (int j = 0; j < 32; ++j) sv[c][j] = __float_as_uint(0.01f * (c * 32 + j + kt + r));
                    } else {
                        tmem_ld_32x32(tmem_s, sv[0]);
                        tmem_ld_32x32(tmem_s + 32, sv[1]);
                        tmem_ld_32x32(tmem_s + 64, sv[2]);
                        tmem_ld_32x32(tmem_s + 96, sv[3]);
                        tmem_ld_wait();
                    }
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&s_free[x]);  // QK of the next key block may overwrite S_x now
                    tr.ev(3);  // S in registers
                    float m_tile;
                    {
                        float m0 = -INFINITY, m1 = -INFINITY, m2 = -INFINITY, m3 = -INFINITY;
                        if (VR_ATT4_ABL & 8) {
                            m0 = __uint_as_float(sv[0][0]);
                        } else if (full) {
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
                    // Both P buffers of this tile were last read by the P.V of the PREVIOUS key block (issued one exp phase
                    // ago): the waits below are normally no-ops. The first block of an item needs none (the epilogue of
                    // the previous item waited for everything).
                    bool prev_done = (kt == 0);
                    const bool grow = (m_tile - m_ref) * sl2 > RESCALE_LOG2;
                    if (kt == 0) {
                        m_ref = (m_tile == -INFINITY) ? 0.f : m_tile;
                    } else if (__any_sync(0xffffffffu, grow)) {
                        mbar_wait(&pv_done[x], (n_blk - 1) & 1);  // O_x complete up to the previous block
                        tc_fence_after();
                        prev_done = true;
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
                    tr.ev(5);  // max done, exps start
#pragma unroll
                    for (int c = 0; c < 4; ++c) {
                        float p[32];
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            float t0, t1;
                            fma2(t0, t1, __uint_as_float(sv[c][j]), __uint_as_float(sv[c][j + 1]), sl2, sl2, neg_ms, neg_ms);
                            if (VR_ATT4_ABL & 1) {
                                p[j] = t0 * 0.001f;
                                p[j + 1] = t1 * 0.001f;
                            } else {
                                p[j] = ex2_approx(t0);
                                p[j + 1] = ex2_approx(t1);
                            }
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
                        uint32_t pk[16];
#pragma unroll
                        for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(p[2 * j], p[2 * j + 1]);
                        // Each 32-key chunk of P leaves for its buffer as soon as it is packed, so the store latency (TMEM
                        // write + completion, shared-memory write + proxy fence: 0.4 ms of the kernel when they were all
                        // issued at the end, ablation in profiles/r02_attention4_ablations.txt) hides under the remaining exps.
                        // The buffers were last read by this tile's previous P.V, issued a whole exp phase ago.
                        if (c == 0 && !prev_done) {
                            mbar_wait(&pv_done[x], (n_blk - 1) & 1);
                            tc_fence_after();
                        }
                        if (VR_ATT4_ABL & 2) {
                            uint32_t acc = 0;
#pragma unroll
                            for (int j = 0; j < 16; ++j) acc ^= pk[j];
                            if (acc == 0x12345678u) l_run += 1.f;  // keep the values alive
                        } else if (c < 2) {
                            if (!(VR_ATT4_ABL & 64)) tmem_st_32x16(tmem_p + c * 16, pk);  // keys 0-63 -> tensor memory (A operand of the TS-MMAs)
                        } else if (!(VR_ATT4_ABL & 32)) {
                            // keys 64-127 -> shared memory, 128 B per row, 16-byte pieces XOR-swizzled by the row (SW128 K-major)
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint4 v4 = make_uint4(pk[i * 4], pk[i * 4 + 1], pk[i * 4 + 2], pk[i * 4 + 3]);
                                *reinterpret_cast<uint4*>(p_row + ((((c - 2) * 4 + i) ^ (r & 7)) << 4)) = v4;
                            }
                        }
                    }
                    if (kt == 0 && pend) epilogue();  // the previous item's O_x (its last P.V retired long ago)
                    tr.ev(7);
                    if (!(VR_ATT4_ABL & (2 | 32))) fence_proxy_async_smem();  // generic-proxy stores -> visible to the tensor core's async proxy
                    if (!(VR_ATT4_ABL & (2 | 64))) tmem_st_wait();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(&p_full[x]);
                    tr.ev(8);  // P stored
                    if (!ONES) l_run += (l0 + l1) + (l2 + l3);
                }
                // the item's O_x stays in tensor memory; it is read out inside the next item's first key block (or after the loop)
                {
                    const long long row = a.cu_q ? (long long)(it.q_begin + q_idx) : (long long)it.b * a.max_q + q_idx;
                    pend_dst = a.out + row * a.ldo + it.head * a.head_dim;
                    pend_store = q_idx < it.len_q;
                    pend_l = l_run;
                    pend = true;
                }
            }
            it = att4_settle(a, nx, total_items, nqp, stride);
        }
        if (pend) epilogue();
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        tmem_dealloc<512>(tmem_base);
    }
}

}  // namespace vr
