// Dense query x corpus similarity + top-k (reference: torch.matmul + torch.topk, fp32, in
// retriever/dense_retriever.py:25-30), built so that the [nq, nd] score matrix never touches HBM and the
// result is the EXACT fp32 top-k:
//
//  1. score_filter_kernel : tcgen05 GEMM on fp16 copies of Q and D (fp32 accumulate in TMEM). Each CTA owns a
//     (128-query block, contiguous range of 256-doc tiles); its epilogue threads (one query row each) keep the 16
//     best approximate scores of their row in registers across all tiles of the range. Output: per query
//     `lists = ranges*2` sorted candidate lists of 16 (score, doc) pairs.
//  2. rescore_topk_kernel : one CTA per query recomputes every candidate's score in fp32 on CUDA cores
//     (q . d, 2304 FMAs each), selects the top-k by (score desc, doc id asc), and PROVES the selection: every
//     doc that was dropped by a list has approximate score <= that list's 16th entry, hence exact score
//     <= tail + eps with eps = fp16 rounding bound * |q| * max|d|. If max(tail) + eps < k-th exact score the
//     result equals the full fp32 scan; otherwise the query is flagged and the caller reruns it through
//  3. exact_scores_kernel + topk_rows_kernel : plain fp32 scan (also the path for tiny problems).
//
// Tie rule everywhere: higher score first, then lower doc id.
#include "common.h"
#include "gemm2.cuh"
#include <math.h>

namespace vr {

constexpr int SC_KT = 16;    // candidates kept per list
constexpr int SC_BN = 256;   // docs per tile
constexpr int SC_MAX_RANGES = 64;  // doc ranges (= candidate lists per query) the filter may use

// Work decomposition of the filter: the unit of work is one 256-query x 256-doc MMA tile (one CTA pair). The doc axis
// is cut into R equal ranges; item i = r*QB + b (QB = 256-query blocks) is the sweep of query block b over doc range r,
// and pair p runs items p, p+P, p+2P, ... Items have the same length and start together, so all pairs of a wave walk
// their doc range in lockstep: at any moment the whole GPU reads at most ceil(P/QB)+1 distinct doc tiles and every doc
// tile is fetched from HBM once and then served to the other query blocks from L2. (A contiguous "stream-K" split of
// the b-major tile list balances perfectly but de-phases the pairs: measured 23 GB of DRAM reads instead of 0.6 GB
// and 6.0 ms instead of the ~4 ms this layout takes at 10 k x 125 k.) R is chosen by the host to fill whole waves
// (score_plan); every item emits ONE 16-entry candidate list per query, so a query has R lists. Items of later waves
// start from the threshold the finished items of the same query published (tau, see the epilogue).
struct ScoreArgs {
    int nq;
    long long nd;
    int dim;
    int lists;         // candidate lists per query (>= R; the ones beyond R are written empty)
    int T;             // doc tiles
    int R;             // doc ranges
    int QB;            // 256-query blocks
    int items;         // R * QB
    float* cand_scores;  // [nq, lists*SC_KT]
    int* cand_ids;
};

struct Score2Cfg {
    static constexpr int STAGES = 6;
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;   // this CTA's 128 query rows
    static constexpr int B_BYTES = 128 * GEMM_BK * 2;       // this CTA's half of the 256 docs
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 32 KB
    static constexpr int MERGE_BYTES = 4 * SC_KT * 32 * 8;  // per quarter: 16 x 32 (score, id) pairs
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + MERGE_BYTES + 1024 + 256;
    static constexpr int TMEM_COLS = 512;
};

__device__ __forceinline__ void topk_insert(float (&sc)[SC_KT], int (&id)[SC_KT], float v, int i) {
    // precondition: v > sc[SC_KT-1]; lists are sorted descending, ties keep the earlier entry first
    sc[SC_KT - 1] = v;
    id[SC_KT - 1] = i;
#pragma unroll
    for (int j = SC_KT - 1; j > 0; --j) {
        const bool sw = sc[j] > sc[j - 1];
        const float a = sc[j], b = sc[j - 1];
        const int ia = id[j], ib = id[j - 1];
        sc[j - 1] = sw ? a : b;
        sc[j] = sw ? b : a;
        id[j - 1] = sw ? ia : ib;
        id[j] = sw ? ib : ia;
    }
}

__device__ __forceinline__ void named_bar_sync(int id, int threads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// CTA pair (tcgen05 cta_group::2): 256 queries x 256 docs per MMA tile, fp16 operands, fp32 accumulators in the TMEM of
// both CTAs (two stages). Roles as in gemm2.cuh; the epilogue threads (one query row each, the two column halves of a
// row in two warps) keep a sorted top-16 in registers over all tiles of a piece, merge the halves through shared memory
// (top-16 of the union: its tail bounds everything either half dropped) and write one list per (query, piece).
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(GEMM_THREADS, 1)
score_filter_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                    const ScoreArgs g) {
    using Cfg = Score2Cfg;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem;
    uint8_t* smem_b = smem + STAGES * Cfg::A_BYTES;
    float* merge_s = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES);  // [4][SC_KT][32]
    int* merge_i = reinterpret_cast<int*>(merge_s + 4 * SC_KT * 32);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::MERGE_BYTES);
    uint64_t* full_bar = bars;
    uint64_t* empty_bar = bars + STAGES;
    uint64_t* tfull_bar = bars + 2 * STAGES;
    uint64_t* tempty_bar = bars + 2 * STAGES + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const int pair = blockIdx.x >> 1;
    const int num_pairs = gridDim.x >> 1;
    const int num_kb = (g.dim + GEMM_BK - 1) / GEMM_BK;
    // item -> (query block b, doc tiles [t0, t1))
    auto item_b = [&](int item) { return item % g.QB; };
    auto item_t0 = [&](int item) { return static_cast<int>(static_cast<long long>(g.T) * (item / g.QB) / g.R); };
    auto item_t1 = [&](int item) { return static_cast<int>(static_cast<long long>(g.T) * (item / g.QB + 1) / g.R); };

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_d);
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
        fence_mbar_init();
    }
    cluster_sync_all();
    if (warp == 2) tmem_alloc_2sm<Cfg::TMEM_COLS>(tmem_slot);
    tc_fence_before();
    cluster_sync_all();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int item = pair; item < g.items; item += num_pairs) {
                const int m0 = item_b(item) * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM;
                const int t1 = item_t1(item);
                for (int t = item_t0(item); t < t1; ++t) {
                    const int n0 = t * SC_BN + static_cast<int>(rank) * 128;
                    for (int kb = 0; kb < num_kb; ++kb) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        const uint32_t lfull = mapa_u32(smem_u32(&full_bar[stage]), 0);
                        mbar_expect_tx_cluster(lfull, Cfg::STAGE_BYTES);
                        tma_load_2d_2sm(&tmap_q, lfull, smem_a + stage * Cfg::A_BYTES, kb * GEMM_BK, m0);
                        tma_load_2d_2sm(&tmap_d, lfull, smem_b + stage * Cfg::B_BYTES, kb * GEMM_BK, n0);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {
            // whole warp, warp-uniform control flow, one elected lane issues (see gemm.cuh)
            constexpr uint32_t idesc = make_idesc_f16(2 * GEMM_BM, SC_BN, 0 /*fp16*/, 0, 0);
            const uint64_t desc_hi = make_smem_desc(0, 16, 1024, kLayoutSW128);
            const uint32_t a_lo0 = smem_u32(smem_a) >> 4, b_lo0 = smem_u32(smem_b) >> 4;
            int stage = 0;
            uint32_t phase = 0;
            int units = 0;  // tiles this pair computes
            for (int item = pair; item < g.items; item += num_pairs) units += item_t1(item) - item_t0(item);
            for (int it = 0; it < units; ++it) {
                const int acc = it & 1;
                mbar_wait(&tempty_bar[acc], ((it >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * SC_BN;
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
    } else if (warp >= 4) {
        const int quarter = warp & 3, half = (warp - 4) >> 2;
        float sc[SC_KT];
        int id[SC_KT];
#pragma unroll
        for (int j = 0; j < SC_KT; ++j) { sc[j] = -INFINITY; id[j] = -1; }
        float* ms = merge_s + quarter * SC_KT * 32;
        int* mi = merge_i + quarter * SC_KT * 32;
        const uint32_t ltempty0 = mapa_u32(smem_u32(&tempty_bar[0]), 0);
        const uint32_t ltempty1 = mapa_u32(smem_u32(&tempty_bar[1]), 0);
        int it = 0;
        for (int item = pair; item < g.items; item += num_pairs) {
            const int b = item_b(item), r = item / g.QB;
            const int row = b * 2 * GEMM_BM + static_cast<int>(rank) * GEMM_BM + quarter * 32 + lane;
            // tau: a lower bound on what this query can still use, published by the items of this query that already
            // finished (the 16th best score of their doc range). Dropping everything <= tau is covered by the proof: the
            // rescoring kernel's bound is the maximum over all list tails, and tau is one of them.
            float* tau_ptr = g.cand_scores + (static_cast<long long>(min(row, g.nq - 1)) * g.lists + g.lists - 1) * SC_KT;
            const float tau = __ldcg(tau_ptr);
            float thr = tau;  // max(tau, sc[SC_KT-1])
            const int t1 = item_t1(item);
            for (int t = item_t0(item); t < t1; ++t, ++it) {
                const int acc = it & 1;
                mbar_wait(&tfull_bar[acc], (it >> 1) & 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + acc * SC_BN + half * 128;
                const long long col_base = static_cast<long long>(t) * SC_BN + half * 128;
#pragma unroll 1
                for (int c = 0; c < 4; ++c) {
                    uint32_t v[32];
                    tmem_ld_32x32(taddr + c * 32, v);
                    tmem_ld_wait();
                    if (c == 3) {
                        tc_fence_before();
                        __syncwarp();
                        if (lane == 0) mbar_arrive_cluster(acc ? ltempty1 : ltempty0);
                    }
                    // fast path: nothing of the 32 scores beats the threshold (the common case after the first tiles)
                    float mx = fmaxf(__uint_as_float(v[0]), __uint_as_float(v[1]));
#pragma unroll
                    for (int j = 2; j < 32; j += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
                    if (mx > thr) {
                        // slow path, ONE compact instance of the insertion code (the fully unrolled form - 32 inlined
                        // insertions per chunk - thrashed the instruction cache: ncu top stall "no_instruction")
                        uint32_t mask = 0;
#pragma unroll
                        for (int j = 0; j < 32; ++j) mask |= (__uint_as_float(v[j]) > thr ? 1u : 0u) << j;
                        const long long c0 = col_base + c * 32;
#pragma unroll 1
                        while (mask) {
                            const int j = __ffs(mask) - 1;
                            mask &= mask - 1;
                            // v[j] with a run-time j: binary select tree over the register array
                            uint32_t s16[16], s8[8], s4[4], s2[2];
#pragma unroll
                            for (int i = 0; i < 16; ++i) s16[i] = (j & 1) ? v[2 * i + 1] : v[2 * i];
#pragma unroll
                            for (int i = 0; i < 8; ++i) s8[i] = (j & 2) ? s16[2 * i + 1] : s16[2 * i];
#pragma unroll
                            for (int i = 0; i < 4; ++i) s4[i] = (j & 4) ? s8[2 * i + 1] : s8[2 * i];
#pragma unroll
                            for (int i = 0; i < 2; ++i) s2[i] = (j & 8) ? s4[2 * i + 1] : s4[2 * i];
                            const float sv = __uint_as_float((j & 16) ? s2[1] : s2[0]);
                            if (sv > thr && c0 + j < g.nd) {
                                topk_insert(sc, id, sv, static_cast<int>(c0 + j));
                                thr = fmaxf(tau, sc[SC_KT - 1]);
                            }
                        }
                    }
                }
            }
            // one candidate list per (query row, doc range): merge the two column halves (top-16 of the union: its tail
            // bounds everything either half dropped), write it, publish its tail as the query's new tau
            if (half == 1) {
#pragma unroll
                for (int j = 0; j < SC_KT; ++j) { ms[j * 32 + lane] = sc[j]; mi[j * 32 + lane] = id[j]; }
            }
            named_bar_sync(1 + quarter, 64);
            if (half == 0) {
#pragma unroll 1
                for (int j = 0; j < SC_KT; ++j) {
                    const float v2 = ms[j * 32 + lane];
                    if (v2 > sc[SC_KT - 1]) topk_insert(sc, id, v2, mi[j * 32 + lane]);
                }
            }
            named_bar_sync(1 + quarter, 64);
            if (half == 0 && row < g.nq) {
                const long long base = (static_cast<long long>(row) * g.lists + r) * SC_KT;
#pragma unroll
                for (int j4 = 0; j4 < SC_KT / 4; ++j4) {
                    *reinterpret_cast<float4*>(g.cand_scores + base + j4 * 4) =
                        make_float4(sc[j4 * 4], sc[j4 * 4 + 1], sc[j4 * 4 + 2], sc[j4 * 4 + 3]);
                    *reinterpret_cast<int4*>(g.cand_ids + base + j4 * 4) =
                        make_int4(id[j4 * 4], id[j4 * 4 + 1], id[j4 * 4 + 2], id[j4 * 4 + 3]);
                }
                const float tail = sc[SC_KT - 1];
                if (tail > tau) {  // float max through the integer atomics (tail may be negative)
                    if (tail >= 0.f) atomicMax(reinterpret_cast<int*>(tau_ptr), __float_as_int(tail));
                    else atomicMin(reinterpret_cast<unsigned int*>(tau_ptr), __float_as_uint(tail));
                }
            }
#pragma unroll
            for (int j = 0; j < SC_KT; ++j) { sc[j] = -INFINITY; id[j] = -1; }
        }
    }
    tc_fence_before();
    cluster_sync_all();  // the peer may still be arriving on this CTA's barriers until here
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc_2sm<Cfg::TMEM_COLS>(tmem_base);
    }
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum_f(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// (score, id) ordering: a before b  <=>  a.s > b.s || (a.s == b.s && a.id < b.id)
__device__ __forceinline__ bool before(float sa, long long ia, float sb, long long ib) {
    return sa > sb || (sa == sb && ia < ib);
}

constexpr int RS_THREADS = 128;
constexpr int RS_MAX_KEEP = 256;

// One CTA per query. The filter hands over `lists` sorted 16-entry candidate lists (approximate scores). Step 1 (warp 0):
// multi-way merge of the list heads keeps the `keep` best candidates by approximate score; everything else - docs a list
// dropped (<= that list's 16th entry) and candidates pruned here (<= the best remaining head) - is bounded by `bound`.
// Step 2: exact fp32 rescoring of the kept candidates, one warp per candidate. Step 3: top-k by (score desc, id asc)
// and the proof  bound + eps < k-th exact score  (else the query is flagged for the fp32 scan).
__global__ void __launch_bounds__(RS_THREADS)
rescore_topk_kernel(const float* __restrict__ Q, const float* __restrict__ D, long long nd, int dim, int lists, int keep,
                    const float* __restrict__ cand_scores, const int* __restrict__ cand_ids,
                    const float* __restrict__ max_doc_norm, int k, long long id_offset, float* __restrict__ out_scores,
                    long long* __restrict__ out_ids, int* __restrict__ flags) {
    extern __shared__ float sm[];
    float* qs = sm;                                   // [dim]
    float* ex = sm + dim;                             // [keep] exact scores
    int* sel = reinterpret_cast<int*>(ex + keep);     // [keep] doc ids of the kept candidates (-1 = none)
    __shared__ float red_s[RS_THREADS / 32];
    __shared__ long long red_i[RS_THREADS / 32];
    __shared__ float sh_bound, sh_qnorm;
    const int q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* qrow = Q + static_cast<long long>(q) * dim;
    const float* cs = cand_scores + static_cast<long long>(q) * lists * SC_KT;
    const int* ci = cand_ids + static_cast<long long>(q) * lists * SC_KT;
    if (warp == 0) {
        // lane l owns lists l, l+32, ...: head position per owned list (<= 2 lists per lane for lists <= 64; general loop)
        float tail = -INFINITY;
        for (int l = lane; l < lists; l += 32) tail = fmaxf(tail, cs[l * SC_KT + SC_KT - 1]);
        int head[(SC_MAX_RANGES + 31) / 32 + 1];
#pragma unroll
        for (int j = 0; j < (SC_MAX_RANGES + 31) / 32 + 1; ++j) head[j] = 0;
        for (int m = 0; m < keep; ++m) {
            // best head of this lane
            float bs = -INFINITY;
            int bj = -1;
#pragma unroll
            for (int j = 0; j < (SC_MAX_RANGES + 31) / 32 + 1; ++j) {
                const int l = lane + j * 32;
                if (l < lists && head[j] < SC_KT) {
                    const float v = cs[l * SC_KT + head[j]];
                    if (ci[l * SC_KT + head[j]] >= 0 && (bj < 0 || v > bs)) { bs = v; bj = j; }
                }
            }
            // warp arg-max (ties: lower lane)
            float ws = bs;
            int wl = bj >= 0 ? lane : 64;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                const float os = __shfl_xor_sync(0xffffffffu, ws, o);
                const int ol = __shfl_xor_sync(0xffffffffu, wl, o);
                if (ol < 64 && (wl >= 64 || os > ws || (os == ws && ol < wl))) { ws = os; wl = ol; }
            }
            if (wl >= 64) {  // every list exhausted
                for (int r = m + lane; r < keep; r += 32) sel[r] = -1;
                break;
            }
            if (lane == wl) {
#pragma unroll
                for (int j = 0; j < (SC_MAX_RANGES + 31) / 32 + 1; ++j)
                    if (j == bj) {
                        sel[m] = ci[(lane + j * 32) * SC_KT + head[j]];
                        ++head[j];
                    }
            }
        }
        // what is left in the lists was pruned: bounded by the best remaining head
        float rem = -INFINITY;
#pragma unroll
        for (int j = 0; j < (SC_MAX_RANGES + 31) / 32 + 1; ++j) {
            const int l = lane + j * 32;
            if (l < lists && head[j] < SC_KT && ci[l * SC_KT + head[j]] >= 0) rem = fmaxf(rem, cs[l * SC_KT + head[j]]);
        }
        float bnd = fmaxf(tail, rem);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) bnd = fmaxf(bnd, __shfl_xor_sync(0xffffffffu, bnd, o));
        if (lane == 0) sh_bound = bnd;
    }
    float qq = 0.f;
    for (int i = threadIdx.x; i < dim; i += RS_THREADS) {
        const float v = qrow[i];
        qs[i] = v;
        qq += v * v;
    }
    qq = warp_sum_f(qq);
    if (lane == 0) red_s[warp] = qq;
    __syncthreads();  // qs, sel, sh_bound, red_s visible
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < RS_THREADS / 32; ++i) s += red_s[i];
        sh_qnorm = sqrtf(s);
    }
    // exact fp32 rescoring: one warp per kept candidate
    for (int c = warp; c < keep; c += RS_THREADS / 32) {
        const int id = sel[c];
        float s = -INFINITY;
        if (id >= 0) {
            const float4* drow = reinterpret_cast<const float4*>(D + static_cast<long long>(id) * dim);
            const float4* q4 = reinterpret_cast<const float4*>(qs);
            float a = 0.f;
            for (int i = lane; i < (dim >> 2); i += 32) {
                const float4 x = drow[i], y = q4[i];
                a = fmaf(x.x, y.x, a);
                a = fmaf(x.y, y.y, a);
                a = fmaf(x.z, y.z, a);
                a = fmaf(x.w, y.w, a);
            }
            s = warp_sum_f(a);
        }
        if (lane == 0) ex[c] = s;
    }
    __syncthreads();  // ex[] complete, red_s reusable
    // k rounds of block arg-max in (score desc, id asc) order, strictly after the previous winner
    float last_s = INFINITY;
    long long last_i = -1;
    float kth = -INFINITY;
    for (int round = 0; round < k; ++round) {
        float bs = -INFINITY;
        long long bi = 0x7fffffffffffffffll;
        for (int c = threadIdx.x; c < keep; c += RS_THREADS) {
            const int id = sel[c];
            if (id < 0) continue;
            const float s = ex[c];
            if (!before(last_s, last_i, s, id)) continue;  // already emitted (or equal to the previous winner)
            if (before(s, id, bs, bi)) { bs = s; bi = id; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (before(os, oi, bs, bi)) { bs = os; bi = oi; }
        }
        if (lane == 0) { red_s[warp] = bs; red_i[warp] = bi; }
        __syncthreads();
        bs = red_s[0]; bi = red_i[0];
        for (int i = 1; i < RS_THREADS / 32; ++i)
            if (before(red_s[i], red_i[i], bs, bi)) { bs = red_s[i]; bi = red_i[i]; }
        __syncthreads();
        const bool valid = bi != 0x7fffffffffffffffll;
        if (threadIdx.x == 0) {
            out_scores[static_cast<long long>(q) * k + round] = valid ? bs : -INFINITY;
            out_ids[static_cast<long long>(q) * k + round] = valid ? bi + id_offset : -1;
        }
        if (!valid) {
            for (int r2 = round + 1 + threadIdx.x; r2 < k; r2 += RS_THREADS) {
                out_scores[static_cast<long long>(q) * k + r2] = -INFINITY;
                out_ids[static_cast<long long>(q) * k + r2] = -1;
            }
            kth = -INFINITY;
            break;
        }
        last_s = bs; last_i = bi; kth = bs;
    }
    if (threadIdx.x == 0) {
        // fp16 operand rounding (2^-11 each) + fp32 accumulation slack, times |q| * max|d|, plus an absolute
        // term for fp16 subnormals (elements below 6.1e-5 carry an absolute error up to 2^-25)
        const float dn = *max_doc_norm;
        const float eps = (9.765625e-4f + static_cast<float>(dim) * 1.1920929e-7f) * sh_qnorm * dn +
                          sqrtf(static_cast<float>(dim)) * 5.9604645e-8f * (sh_qnorm + dn) + 1e-6f;
        int flag = 0;
        if (sh_bound > -INFINITY && !(sh_bound + eps < kth)) flag = 1;  // something was dropped that might belong
        // the bound assumes finite fp16 copies of both operands: a row norm >= 65504 (or inf / NaN, for which the
        // comparison is false as well) means some |x| may have overflowed fp16 -> rerun this query through the fp32 scan
        if (!(sh_qnorm < 65504.f) || !(dn < 65504.f)) flag = 1;
        flags[q] = flag;
    }
}

// ---------------------------------------------------------------------------------------------
// Plain fp32 scan: scores[q, doc] = Q[q] . D[doc].  Block = 8 warps; each warp owns docs, up to 8 queries per pass.
// ---------------------------------------------------------------------------------------------
constexpr int EX_QB = 8;   // queries per pass (their rows sit in shared memory)
constexpr int EX_DW = 4;   // docs per warp per pass: each query float4 read from shared memory feeds 4 doc rows

template <int NQ>  // queries per pass this instantiation is compiled for (1, 2, 4 or EX_QB): fewer accumulators, deeper unroll
__global__ void __launch_bounds__(256)
exact_scores_kernel(const float* __restrict__ Q, int nq, const float* __restrict__ D, long long nd, int dim,
                    float* __restrict__ scores) {
    extern __shared__ float qsm[];  // [NQ, dim]
    const int q0 = blockIdx.y * NQ;
    const int nqb = min(NQ, nq - q0);
    for (int i = threadIdx.x; i < nqb * dim; i += blockDim.x) qsm[i] = Q[static_cast<long long>(q0) * dim + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = dim >> 2;
    const long long groups = (nd + EX_DW - 1) / EX_DW;
    for (long long gi = static_cast<long long>(blockIdx.x) * 8 + warp; gi < groups; gi += static_cast<long long>(gridDim.x) * 8) {
        const long long doc0 = gi * EX_DW;
        const float4* drow[EX_DW];
#pragma unroll
        for (int d = 0; d < EX_DW; ++d) drow[d] = reinterpret_cast<const float4*>(D + min(doc0 + d, nd - 1) * dim);
        float acc[EX_DW][NQ];
#pragma unroll
        for (int d = 0; d < EX_DW; ++d)
#pragma unroll
            for (int j = 0; j < NQ; ++j) acc[d][j] = 0.f;
#pragma unroll(NQ <= 2 ? 3 : 1)
        for (int i = lane; i < nv; i += 32) {
            float4 x[EX_DW];
#pragma unroll
            for (int d = 0; d < EX_DW; ++d) x[d] = __ldcs(drow[d] + i);  // streamed once: do not keep in L1
#pragma unroll
            for (int j = 0; j < NQ; ++j) {
                if (j < nqb) {
                    const float4 y = reinterpret_cast<const float4*>(qsm + j * dim)[i];
#pragma unroll
                    for (int d = 0; d < EX_DW; ++d) {
                        acc[d][j] = fmaf(x[d].x, y.x, acc[d][j]);
                        acc[d][j] = fmaf(x[d].y, y.y, acc[d][j]);
                        acc[d][j] = fmaf(x[d].z, y.z, acc[d][j]);
                        acc[d][j] = fmaf(x[d].w, y.w, acc[d][j]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NQ; ++j) {
            if (j < nqb) {
#pragma unroll
                for (int d = 0; d < EX_DW; ++d) {
                    const float s = warp_sum_f(acc[d][j]);
                    if (lane == d && doc0 + d < nd) scores[static_cast<long long>(q0 + j) * nd + doc0 + d] = s;
                }
            }
        }
    }
}

// top-k of each row of a dense [rows, cols] fp32 matrix (optionally with explicit ids per entry).
__global__ void __launch_bounds__(256)
topk_rows_kernel(const float* __restrict__ scores, const long long* __restrict__ ids, long long cols, int k,
                 long long id_offset, long long chunk_cols, float* __restrict__ out_scores,
                 long long* __restrict__ out_ids) {
    // block (row, chunk): top-k of columns [chunk*chunk_cols, ...) of one row, written as list `row*gridDim.y + chunk`
    __shared__ float red_s[8];
    __shared__ long long red_i[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const long long c_lo = static_cast<long long>(blockIdx.y) * chunk_cols;
    const long long c_hi = min(cols, c_lo + chunk_cols);
    const float* srow = scores + static_cast<long long>(blockIdx.x) * cols;
    const long long* irow = ids ? ids + static_cast<long long>(blockIdx.x) * cols : nullptr;
    const long long row = static_cast<long long>(blockIdx.x) * gridDim.y + blockIdx.y;  // output list
    float last_s = INFINITY;
    long long last_i = -1;
    for (int round = 0; round < k; ++round) {
        float bs = -INFINITY;
        long long bi = 0x7fffffffffffffffll;
        for (long long c = c_lo + threadIdx.x; c < c_hi; c += 256) {
            const long long id = irow ? irow[c] : c;
            if (id < 0) continue;
            const float s = srow[c];
            if (!before(last_s, last_i, s, id)) continue;
            if (before(s, id, bs, bi)) { bs = s; bi = id; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (before(os, oi, bs, bi)) { bs = os; bi = oi; }
        }
        if (lane == 0) { red_s[warp] = bs; red_i[warp] = bi; }
        __syncthreads();
        bs = red_s[0]; bi = red_i[0];
        for (int i = 1; i < 8; ++i)
            if (before(red_s[i], red_i[i], bs, bi)) { bs = red_s[i]; bi = red_i[i]; }
        __syncthreads();
        const bool valid = bi != 0x7fffffffffffffffll;
        if (threadIdx.x == 0) {
            out_scores[row * k + round] = valid ? bs : -INFINITY;
            out_ids[row * k + round] = valid ? bi + id_offset : -1;
        }
        if (!valid) {
            for (int r2 = round + 1 + threadIdx.x; r2 < k; r2 += 256) {
                out_scores[row * k + r2] = -INFINITY;
                out_ids[row * k + r2] = -1;
            }
            break;
        }
        last_s = bs; last_i = bi;
    }
}

// Short rows (cols <= 32*NPL): one WARP per row, the row lives in registers, k rounds of shuffle arg-max - no block barriers.
// This is the merge of per-rank / per-shard partial top-k lists ([nq, world*k]) and the second pass of the chunked top-k.
template <int NPL>
__global__ void __launch_bounds__(256)
topk_rows_warp_kernel(const float* __restrict__ scores, const long long* __restrict__ ids, int rows, int cols, int k,
                      long long id_offset, float* __restrict__ out_scores, long long* __restrict__ out_ids) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const float* srow = scores + static_cast<long long>(row) * cols;
    const long long* irow = ids ? ids + static_cast<long long>(row) * cols : nullptr;
    float s[NPL];
    long long id[NPL];
#pragma unroll
    for (int j = 0; j < NPL; ++j) {
        const int c = lane + j * 32;
        const bool in = c < cols;
        id[j] = in ? (irow ? irow[c] : static_cast<long long>(c)) : -1;
        s[j] = (in && id[j] >= 0) ? srow[c] : -INFINITY;
    }
    float last_s = INFINITY;
    long long last_i = -1;
    for (int round = 0; round < k; ++round) {
        float bs = -INFINITY;
        long long bi = 0x7fffffffffffffffll;
#pragma unroll
        for (int j = 0; j < NPL; ++j) {
            if (id[j] < 0) continue;
            if (!before(last_s, last_i, s[j], id[j])) continue;
            if (before(s[j], id[j], bs, bi)) { bs = s[j]; bi = id[j]; }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const float os = __shfl_xor_sync(0xffffffffu, bs, o);
            const long long oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (before(os, oi, bs, bi)) { bs = os; bi = oi; }
        }
        const bool valid = bi != 0x7fffffffffffffffll;
        if (lane == 0) {
            out_scores[static_cast<long long>(row) * k + round] = valid ? bs : -INFINITY;
            out_ids[static_cast<long long>(row) * k + round] = valid ? bi + id_offset : -1;
        }
        if (!valid) {
            for (int r2 = round + 1 + lane; r2 < k; r2 += 32) {
                out_scores[static_cast<long long>(row) * k + r2] = -INFINITY;
                out_ids[static_cast<long long>(row) * k + r2] = -1;
            }
            break;
        }
        last_s = bs; last_i = bi;
    }
}

static int launch_topk_rows(const float* scores, const long long* ids, int rows, long long cols, int k, long long id_offset,
                            float* out_scores, long long* out_ids, cudaStream_t s) {
    if (cols <= 128)
        topk_rows_warp_kernel<4><<<(rows + 7) / 8, 256, 0, s>>>(scores, ids, rows, static_cast<int>(cols), k, id_offset, out_scores, out_ids);
    else if (cols <= 512)
        topk_rows_warp_kernel<16><<<(rows + 7) / 8, 256, 0, s>>>(scores, ids, rows, static_cast<int>(cols), k, id_offset, out_scores, out_ids);
    else
        topk_rows_kernel<<<rows, 256, 0, s>>>(scores, ids, cols, k, id_offset, cols, out_scores, out_ids);
    return 0;
}

// fp32 -> fp16 rows, with the row L2 norms and their maximum (norms are >= 0, so the int view orders them).
__global__ void f32_to_f16_rows_kernel(const float* __restrict__ src, long long rows, int dim, __half* __restrict__ dst,
                                       float* __restrict__ norms, float* __restrict__ max_norm) {
    constexpr int CH = 6;  // float4 per lane in flight (dim 2304 = 3 chunks of 6 x 32 float4)
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    const int nv = dim >> 2;
    for (long long r = static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5); r < rows;
         r += static_cast<long long>(gridDim.x) * warps) {
        const float4* s4 = reinterpret_cast<const float4*>(src + r * dim);
        uint2* d2 = reinterpret_cast<uint2*>(dst + r * dim);
        float ss = 0.f;
        for (int i0 = 0; i0 < nv; i0 += CH * 32) {
            float4 v[CH];
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int i = i0 + j * 32 + lane;
                v[j] = i < nv ? __ldcs(s4 + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int j = 0; j < CH; ++j) {
                const int i = i0 + j * 32 + lane;
                ss += (v[j].x * v[j].x + v[j].y * v[j].y) + (v[j].z * v[j].z + v[j].w * v[j].w);
                if (i < nv) {
                    uint2 pk;
                    pk.x = pack_f16x2(v[j].x, v[j].y);
                    pk.y = pack_f16x2(v[j].z, v[j].w);
                    d2[i] = pk;
                }
            }
        }
        ss = warp_sum_f(ss);
        if (lane == 0) {
            const float n = sqrtf(ss);
            if (norms) norms[r] = n;
            if (max_norm) atomicMax(reinterpret_cast<int*>(max_norm), __float_as_int(n));
        }
    }
}

// Before the filter: every list slot beyond the R real ones becomes an empty list (-inf, -1). The LAST slot doubles as
// the query's running threshold tau (its first score; id -1 keeps the rescoring kernel from reading it as a candidate).
__global__ void score_init_lists_kernel(float* __restrict__ cand_scores, int* __restrict__ cand_ids, int nq, int lists, int R) {
    const int per_row = (lists - R) * SC_KT;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < static_cast<long long>(nq) * per_row;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const long long row = i / per_row;
        const long long o = (row * lists + R) * SC_KT + (i - row * per_row);
        cand_scores[o] = -INFINITY;
        cand_ids[o] = -1;
    }
}

// Co-resident CTA pairs of the filter kernel on the current device (GPCs with an odd number of usable SMs cannot pair
// all of them; a persistent kernel must not launch more clusters than fit at once).
static int score_pairs() {
    static int cached[64] = {};
    const int dev = current_device();
    const int slot = (dev >= 0 && dev < 64) ? dev : 0;
    if (cached[slot] == 0) {
        int n = 0;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(num_sms() / 2 * 2);
        cfg.blockDim = dim3(GEMM_THREADS);
        cfg.dynamicSmemBytes = Score2Cfg::SMEM_BYTES;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = 2; attr.val.clusterDim.y = 1; attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        if (dev < 0 ||
            cudaFuncSetAttribute(score_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Score2Cfg::SMEM_BYTES) != cudaSuccess ||
            cudaOccupancyMaxActiveClusters(&n, score_filter_kernel, &cfg) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            n = num_sms() / 2;
        }
        cached[slot] = n < num_sms() / 2 ? n : num_sms() / 2;
    }
    return cached[slot];
}

struct ScorePlan {
    int T, R, QB, items, lists, pairs;
};

static ScorePlan score_plan(int nq, long long nd) {
    ScorePlan p;
    p.QB = (nq + 2 * GEMM_BM - 1) / (2 * GEMM_BM);
    p.T = static_cast<int>((nd + SC_BN - 1) / SC_BN);
    const int P = score_pairs();
    // time ~ waves * tiles per item, plus a quarter tile per item for the pipeline fill and the list write-out
    double best = 1e30;
    p.R = 1;
    for (int R = 1; R <= SC_MAX_RANGES && R <= p.T; ++R) {
        const long long waves = (static_cast<long long>(p.QB) * R + P - 1) / P;
        const double cost = static_cast<double>(waves) * ((p.T + R - 1) / R + 0.25);
        if (cost < best * 0.98) { best = cost; p.R = R; }  // a larger R must win by 2 %: fewer lists to rescore
    }
    p.items = p.QB * p.R;
    p.pairs = p.items < P ? p.items : P;
    p.lists = 2 * ((p.R + 2) / 2);  // >= R + 1: the last slot carries the per-query threshold (score_init_lists_kernel)
    return p;
}

static int score_ranges_for(int nq, long long nd) { return score_plan(nq, nd).lists / 2; }

}  // namespace vr

using namespace vr;

extern "C" int vr_score_ranges(int32_t nq, int64_t nd) { return score_ranges_for(nq, nd); }
extern "C" int vr_score_plan(int32_t nq, int64_t nd, int32_t* out6) {
    VR_REQUIRE(out6 && nq > 0 && nd > 0, "vr_score_plan: bad arguments");
    const ScorePlan p = score_plan(nq, nd);
    out6[0] = p.T; out6[1] = p.R; out6[2] = p.QB; out6[3] = p.items; out6[4] = p.pairs; out6[5] = p.lists;
    return 0;
}
extern "C" int vr_score_list_len(void) { return SC_KT; }

extern "C" int vr_f32_to_f16_rows(const float* src, int64_t rows, int32_t dim, void* dst_f16, float* norms, float* max_norm,
                                  void* stream) {
    VR_REQUIRE(src && dst_f16, "vr_f32_to_f16_rows: null pointer");
    VR_REQUIRE(rows > 0 && dim > 0 && dim % 4 == 0, "vr_f32_to_f16_rows: bad shape rows=%lld dim=%d", (long long)rows, dim);
    long long blocks = (rows + 7) / 8;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    f32_to_f16_rows_kernel<<<static_cast<int>(blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        src, rows, dim, reinterpret_cast<__half*>(dst_f16), norms, max_norm);
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int vr_score_filter(const void* q_f16, int32_t nq, const void* d_f16, int64_t nd, int32_t dim, int32_t ranges,
                               float* cand_scores, int32_t* cand_ids, void* stream) {
    VR_REQUIRE(q_f16 && d_f16 && cand_scores && cand_ids, "vr_score_filter: null pointer");
    VR_REQUIRE(nq > 0 && nd > 0 && nd < 2147483647ll && dim % 8 == 0, "vr_score_filter: bad shape nq=%d nd=%lld dim=%d", nq,
               (long long)nd, dim);
    VR_REQUIRE(nq < (1 << 30), "vr_score_filter: too many queries");
    const ScorePlan plan = score_plan(nq, nd);
    VR_REQUIRE(ranges * 2 == plan.lists, "vr_score_filter: ranges must come from vr_score_ranges()");
    using Cfg = Score2Cfg;
    CUtensorMap tq, td;
    if (int rc = make_tmap_2d(&tq, q_f16, nq, dim, dim, GEMM_BM, GEMM_BK, 128, false)) return rc;
    if (int rc = make_tmap_2d(&td, d_f16, nd, dim, dim, 128, GEMM_BK, 128, false)) return rc;
    static unsigned long long attr_set = 0;
    if (first_use_on_device(&attr_set))
        VR_CHECK_CUDA(cudaFuncSetAttribute(score_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    ScoreArgs g;
    g.nq = nq; g.nd = nd; g.dim = dim; g.lists = plan.lists; g.T = plan.T; g.R = plan.R; g.QB = plan.QB; g.items = plan.items;
    g.cand_scores = cand_scores; g.cand_ids = cand_ids;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    {
        const long long n = static_cast<long long>(nq) * (plan.lists - plan.R) * SC_KT;
        long long blocks = (n + 255) / 256;
        if (blocks > num_sms() * 8) blocks = num_sms() * 8;
        score_init_lists_kernel<<<static_cast<int>(blocks), 256, 0, st>>>(cand_scores, cand_ids, nq, plan.lists, plan.R);
        VR_CHECK_CUDA(cudaGetLastError());
    }
    score_filter_kernel<<<2 * plan.pairs, GEMM_THREADS, Cfg::SMEM_BYTES, st>>>(tq, td, g);
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int vr_score_rescore(const float* q_f32, int32_t nq, const float* d_f32, int64_t nd, int32_t dim, int32_t ranges,
                                const float* cand_scores, const int32_t* cand_ids, const float* max_doc_norm, int32_t k,
                                int64_t id_offset, float* out_scores, int64_t* out_ids, int32_t* flags, void* stream) {
    VR_REQUIRE(q_f32 && d_f32 && cand_scores && cand_ids && max_doc_norm && out_scores && out_ids && flags,
               "vr_score_rescore: null pointer");
    VR_REQUIRE(nq > 0 && k > 0 && dim % 4 == 0, "vr_score_rescore: bad shape");
    const int lists = ranges * 2;
    VR_REQUIRE(ranges > 0 && lists <= SC_MAX_RANGES + 2, "vr_score_rescore: ranges must come from vr_score_ranges()");
    int keep = 2 * k > 32 ? 2 * k : 32;  // candidates rescored per query (the best by approximate score)
    if (keep > lists * SC_KT) keep = lists * SC_KT;
    if (keep > RS_MAX_KEEP) keep = RS_MAX_KEEP;
    const size_t smem = (static_cast<size_t>(dim) + 2 * static_cast<size_t>(keep)) * sizeof(float);
    VR_REQUIRE(smem <= 200 * 1024, "vr_score_rescore: dim too large for shared memory (%zu bytes)", smem);
    static unsigned long long attr_set = 0;
    if (smem > 48 * 1024 && first_use_on_device(&attr_set))
        VR_CHECK_CUDA(cudaFuncSetAttribute(rescore_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    rescore_topk_kernel<<<nq, RS_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        q_f32, d_f32, nd, dim, lists, keep, cand_scores, cand_ids, max_doc_norm, k, id_offset, out_scores,
        reinterpret_cast<long long*>(out_ids), flags);
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int vr_score_exact(const float* q_f32, int32_t nq, const float* d_f32, int64_t nd, int32_t dim, float* scores,
                              void* stream) {
    VR_REQUIRE(q_f32 && d_f32 && scores, "vr_score_exact: null pointer");
    VR_REQUIRE(nq > 0 && nd > 0 && dim % 4 == 0, "vr_score_exact: bad shape");
    const int NQ = nq >= EX_QB ? EX_QB : (nq >= 4 ? 4 : (nq >= 2 ? 2 : 1));
    const size_t smem = static_cast<size_t>(NQ) * dim * sizeof(float);
    VR_REQUIRE(smem <= 200 * 1024, "vr_score_exact: dim too large");
    long long bx = (nd + 8 * EX_DW - 1) / (8 * EX_DW);
    const long long cap = static_cast<long long>(num_sms()) * 3;
    if (bx > cap) bx = cap;
    dim3 grid(static_cast<unsigned>(bx), (nq + NQ - 1) / NQ);
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
#define VR_EXACT_LAUNCH(N)                                                                                               \
    do {                                                                                                                 \
        static unsigned long long attr_set = 0;                                                                          \
        if (smem > 48 * 1024 && first_use_on_device(&attr_set))                                                          \
            VR_CHECK_CUDA(cudaFuncSetAttribute(exact_scores_kernel<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024)); \
        exact_scores_kernel<N><<<grid, 256, smem, st>>>(q_f32, nq, d_f32, nd, dim, scores);                              \
    } while (0)
    if (NQ == EX_QB) VR_EXACT_LAUNCH(EX_QB);
    else if (NQ == 4) VR_EXACT_LAUNCH(4);
    else if (NQ == 2) VR_EXACT_LAUNCH(2);
    else VR_EXACT_LAUNCH(1);
#undef VR_EXACT_LAUNCH
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int vr_topk_rows(const float* scores, const int64_t* ids, int32_t rows, int64_t cols, int32_t k, int64_t id_offset,
                            float* out_scores, int64_t* out_ids, void* stream) {
    VR_REQUIRE(scores && out_scores && out_ids, "vr_topk_rows: null pointer");
    VR_REQUIRE(rows > 0 && cols > 0 && k > 0, "vr_topk_rows: bad shape");
    launch_topk_rows(scores, reinterpret_cast<const long long*>(ids), rows, cols, k, id_offset, out_scores,
                     reinterpret_cast<long long*>(out_ids), reinterpret_cast<cudaStream_t>(stream));
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int vr_topk_rows_chunked(const float* scores, int32_t rows, int64_t cols, int32_t k, int64_t id_offset,
                                    int32_t chunks, float* ws_scores, int64_t* ws_ids, float* out_scores,
                                    int64_t* out_ids, void* stream) {
    VR_REQUIRE(scores && ws_scores && ws_ids && out_scores && out_ids, "vr_topk_rows_chunked: null pointer");
    VR_REQUIRE(rows > 0 && cols > 0 && k > 0 && chunks > 0 && chunks <= 65535, "vr_topk_rows_chunked: bad shape");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    const long long chunk_cols = (cols + chunks - 1) / chunks;
    // pass 1: every (row, chunk) block reduces its column range to a sorted top-k list (ids = column + id_offset)
    topk_rows_kernel<<<dim3(rows, chunks), 256, 0, s>>>(scores, nullptr, cols, k, id_offset, chunk_cols, ws_scores,
                                                        reinterpret_cast<long long*>(ws_ids));
    VR_CHECK_CUDA(cudaGetLastError());
    // pass 2: merge the `chunks` lists of each row (explicit ids; exhausted lists carry id -1 and are skipped)
    launch_topk_rows(ws_scores, reinterpret_cast<const long long*>(ws_ids), rows, static_cast<long long>(chunks) * k, k, 0,
                     out_scores, reinterpret_cast<long long*>(out_ids), s);
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
