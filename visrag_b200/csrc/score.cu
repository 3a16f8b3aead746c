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
#include "gemm.cuh"
#include <math.h>

namespace vr {

constexpr int SC_KT = 16;    // candidates kept per list
constexpr int SC_BN = 256;   // docs per tile

struct ScoreArgs {
    int nq;
    long long nd;
    int dim;
    int ranges;
    float* cand_scores;  // [nq, ranges*2*SC_KT]
    int* cand_ids;
};

__device__ __forceinline__ void topk_insert(float (&sc)[SC_KT], int (&id)[SC_KT], float v, int i) {
    // precondition: v > sc[SC_KT-1]; lists are sorted descending, ties keep the earlier (lower) doc id first
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

__global__ void __launch_bounds__(GEMM_THREADS, 1)
score_filter_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_d,
                    const ScoreArgs g) {
    using Cfg = GemmCfg<SC_BN>;
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

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r = blockIdx.x, qb = blockIdx.y;
    const long long doc_tiles = (g.nd + SC_BN - 1) / SC_BN;
    const int t_begin = static_cast<int>(doc_tiles * r / g.ranges);
    const int t_end = static_cast<int>(doc_tiles * (r + 1) / g.ranges);
    const int num_kb = (g.dim + GEMM_BK - 1) / GEMM_BK;
    const int m0 = qb * GEMM_BM;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmap_q);
        tma_prefetch_desc(&tmap_d);
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
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = t_begin; t < t_end; ++t) {
                for (int kb = 0; kb < num_kb; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    tma_load_2d(&tmap_q, &full_bar[stage], smem_a + stage * Cfg::A_BYTES, kb * GEMM_BK, m0);
                    tma_load_2d(&tmap_d, &full_bar[stage], smem_b + stage * Cfg::B_BYTES, kb * GEMM_BK, t * SC_BN);
                    tma_load_2d(&tmap_d, &full_bar[stage], smem_b + stage * Cfg::B_BYTES + Cfg::B_BYTES / 2, kb * GEMM_BK,
                                t * SC_BN + 128);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // whole warp, warp-uniform control flow, one elected lane issues (see gemm.cuh)
        constexpr uint32_t idesc = make_idesc_f16(GEMM_BM, SC_BN, 0 /*fp16*/, 0, 0);
        const uint64_t desc_hi = make_smem_desc(0, 16, 1024, kLayoutSW128);
        const uint32_t a_lo0 = smem_u32(smem_a) >> 4, b_lo0 = smem_u32(smem_b) >> 4;
        int stage = 0;
        uint32_t phase = 0;
        int it = 0;
        for (int t = t_begin; t < t_end; ++t, ++it) {
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
                    umma_f16_ss(d_tmem, ad, bd, idesc, kb != 0 ? 1u : 0u);
                    umma_f16_ss(d_tmem, ad + 2, bd + 2, idesc, 1u);
                    umma_f16_ss(d_tmem, ad + 4, bd + 4, idesc, 1u);
                    umma_f16_ss(d_tmem, ad + 6, bd + 6, idesc, 1u);
                    umma_commit(&empty_bar[stage]);
                    if (kb == num_kb - 1) umma_commit(&tfull_bar[acc]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        const int quarter = warp & 3, half = (warp - 4) >> 2;
        const int row = m0 + quarter * 32 + lane;
        float sc[SC_KT];
        int id[SC_KT];
#pragma unroll
        for (int j = 0; j < SC_KT; ++j) { sc[j] = -INFINITY; id[j] = -1; }
        int it = 0;
        for (int t = t_begin; t < t_end; ++t, ++it) {
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
                    if (lane == 0) mbar_arrive(&tempty_bar[acc]);
                }
                const long long c0 = col_base + c * 32;
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const float s = __uint_as_float(v[j]);
                    if (s > sc[SC_KT - 1] && c0 + j < g.nd) topk_insert(sc, id, s, static_cast<int>(c0 + j));
                }
            }
        }
        if (row < g.nq) {
            const long long base = (static_cast<long long>(row) * g.ranges * 2 + r * 2 + half) * SC_KT;
#pragma unroll
            for (int j4 = 0; j4 < SC_KT / 4; ++j4) {
                *reinterpret_cast<float4*>(g.cand_scores + base + j4 * 4) =
                    make_float4(sc[j4 * 4], sc[j4 * 4 + 1], sc[j4 * 4 + 2], sc[j4 * 4 + 3]);
                *reinterpret_cast<int4*>(g.cand_ids + base + j4 * 4) =
                    make_int4(id[j4 * 4], id[j4 * 4 + 1], id[j4 * 4 + 2], id[j4 * 4 + 3]);
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

__global__ void __launch_bounds__(RS_THREADS)
rescore_topk_kernel(const float* __restrict__ Q, const float* __restrict__ D, long long nd, int dim, int lists,
                    const float* __restrict__ cand_scores, const int* __restrict__ cand_ids,
                    const float* __restrict__ max_doc_norm, int k, long long id_offset, float* __restrict__ out_scores,
                    long long* __restrict__ out_ids, int* __restrict__ flags) {
    extern __shared__ float sm[];
    const int C = lists * SC_KT;
    float* qs = sm;               // [dim]
    float* ex = sm + dim;         // [C] exact scores
    __shared__ float red_s[RS_THREADS / 32];
    __shared__ long long red_i[RS_THREADS / 32];
    __shared__ float sh_bound, sh_qnorm;
    const int q = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float* qrow = Q + static_cast<long long>(q) * dim;
    float qq = 0.f;
    for (int i = threadIdx.x; i < dim; i += RS_THREADS) {
        const float v = qrow[i];
        qs[i] = v;
        qq += v * v;
    }
    qq = warp_sum_f(qq);
    if (lane == 0) red_s[warp] = qq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < RS_THREADS / 32; ++i) s += red_s[i];
        sh_qnorm = sqrtf(s);
    }
    const float* cs = cand_scores + static_cast<long long>(q) * C;
    const int* ci = cand_ids + static_cast<long long>(q) * C;
    // exact fp32 rescoring: one warp per candidate
    for (int c = warp; c < C; c += RS_THREADS / 32) {
        const int id = ci[c];
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
    // bound on everything the filter dropped: max over lists of the list tail (approximate score)
    float tail = -INFINITY;
    for (int l = threadIdx.x; l < lists; l += RS_THREADS) tail = fmaxf(tail, cs[l * SC_KT + SC_KT - 1]);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tail = fmaxf(tail, __shfl_xor_sync(0xffffffffu, tail, o));
    __syncthreads();  // ex[] complete, red_s reusable
    if (lane == 0) red_s[warp] = tail;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = red_s[0];
        for (int i = 1; i < RS_THREADS / 32; ++i) s = fmaxf(s, red_s[i]);
        sh_bound = s;
    }
    __syncthreads();
    // k rounds of block arg-max in (score desc, id asc) order, strictly after the previous winner
    float last_s = INFINITY;
    long long last_i = -1;
    float kth = -INFINITY;
    for (int round = 0; round < k; ++round) {
        float bs = -INFINITY;
        long long bi = 0x7fffffffffffffffll;
        for (int c = threadIdx.x; c < C; c += RS_THREADS) {
            const int id = ci[c];
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
        if (!valid) { kth = -INFINITY; break; }
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
constexpr int EX_QB = 8;

__global__ void __launch_bounds__(256)
exact_scores_kernel(const float* __restrict__ Q, int nq, const float* __restrict__ D, long long nd, int dim,
                    float* __restrict__ scores) {
    extern __shared__ float qsm[];  // [EX_QB, dim]
    const int q0 = blockIdx.y * EX_QB;
    const int nqb = min(EX_QB, nq - q0);
    for (int i = threadIdx.x; i < nqb * dim; i += blockDim.x) qsm[i] = Q[static_cast<long long>(q0) * dim + i];
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nv = dim >> 2;
    for (long long doc = static_cast<long long>(blockIdx.x) * 8 + warp; doc < nd; doc += static_cast<long long>(gridDim.x) * 8) {
        const float4* drow = reinterpret_cast<const float4*>(D + doc * dim);
        float acc[EX_QB];
#pragma unroll
        for (int j = 0; j < EX_QB; ++j) acc[j] = 0.f;
        for (int i = lane; i < nv; i += 32) {
            const float4 x = drow[i];
#pragma unroll
            for (int j = 0; j < EX_QB; ++j) {
                if (j < nqb) {
                    const float4 y = reinterpret_cast<const float4*>(qsm + j * dim)[i];
                    acc[j] = fmaf(x.x, y.x, acc[j]);
                    acc[j] = fmaf(x.y, y.y, acc[j]);
                    acc[j] = fmaf(x.z, y.z, acc[j]);
                    acc[j] = fmaf(x.w, y.w, acc[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < EX_QB; ++j) {
            if (j < nqb) {
                const float s = warp_sum_f(acc[j]);
                if (lane == 0) scores[static_cast<long long>(q0 + j) * nd + doc] = s;
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

// fp32 -> fp16 rows, with the row L2 norms and their maximum (norms are >= 0, so the int view orders them).
__global__ void f32_to_f16_rows_kernel(const float* __restrict__ src, long long rows, int dim, __half* __restrict__ dst,
                                       float* __restrict__ norms, float* __restrict__ max_norm) {
    const int warps = blockDim.x >> 5, lane = threadIdx.x & 31;
    for (long long r = static_cast<long long>(blockIdx.x) * warps + (threadIdx.x >> 5); r < rows;
         r += static_cast<long long>(gridDim.x) * warps) {
        const float4* s4 = reinterpret_cast<const float4*>(src + r * dim);
        uint2* d2 = reinterpret_cast<uint2*>(dst + r * dim);
        float ss = 0.f;
        for (int i = lane; i < (dim >> 2); i += 32) {
            const float4 v = s4[i];
            ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
            uint2 pk;
            pk.x = pack_f16x2(v.x, v.y);
            pk.y = pack_f16x2(v.z, v.w);
            d2[i] = pk;
        }
        ss = warp_sum_f(ss);
        if (lane == 0) {
            const float n = sqrtf(ss);
            if (norms) norms[r] = n;
            if (max_norm) atomicMax(reinterpret_cast<int*>(max_norm), __float_as_int(n));
        }
    }
}

static int score_ranges_for(int nq, long long nd) {
    const int qb = (nq + GEMM_BM - 1) / GEMM_BM;
    long long tiles = (nd + SC_BN - 1) / SC_BN;
    long long r = num_sms() / qb;
    if (r < 1) r = 1;
    if (r > tiles) r = tiles;
    if (r < 1) r = 1;
    return static_cast<int>(r);
}

}  // namespace vr

using namespace vr;

extern "C" int vr_score_ranges(int32_t nq, int64_t nd) { return score_ranges_for(nq, nd); }
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
    VR_REQUIRE(ranges == score_ranges_for(nq, nd), "vr_score_filter: ranges must come from vr_score_ranges()");
    using Cfg = GemmCfg<SC_BN>;
    CUtensorMap tq, td;
    if (int rc = make_tmap_2d(&tq, q_f16, nq, dim, dim, GEMM_BM, GEMM_BK, 128, false)) return rc;
    if (int rc = make_tmap_2d(&td, d_f16, nd, dim, dim, 128, GEMM_BK, 128, false)) return rc;
    static unsigned long long attr_set = 0;
    if (first_use_on_device(&attr_set))
        VR_CHECK_CUDA(cudaFuncSetAttribute(score_filter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    ScoreArgs g;
    g.nq = nq; g.nd = nd; g.dim = dim; g.ranges = ranges; g.cand_scores = cand_scores; g.cand_ids = cand_ids;
    dim3 grid(ranges, (nq + GEMM_BM - 1) / GEMM_BM);
    score_filter_kernel<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, reinterpret_cast<cudaStream_t>(stream)>>>(tq, td, g);
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
    const size_t smem = (static_cast<size_t>(dim) + static_cast<size_t>(lists) * SC_KT) * sizeof(float);
    VR_REQUIRE(smem <= 200 * 1024, "vr_score_rescore: candidate set too large for shared memory (%zu bytes)", smem);
    static unsigned long long attr_set = 0;
    if (smem > 48 * 1024 && first_use_on_device(&attr_set))
        VR_CHECK_CUDA(cudaFuncSetAttribute(rescore_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    rescore_topk_kernel<<<nq, RS_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(
        q_f32, d_f32, nd, dim, lists, cand_scores, cand_ids, max_doc_norm, k, id_offset, out_scores,
        reinterpret_cast<long long*>(out_ids), flags);
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int vr_score_exact(const float* q_f32, int32_t nq, const float* d_f32, int64_t nd, int32_t dim, float* scores,
                              void* stream) {
    VR_REQUIRE(q_f32 && d_f32 && scores, "vr_score_exact: null pointer");
    VR_REQUIRE(nq > 0 && nd > 0 && dim % 4 == 0, "vr_score_exact: bad shape");
    const size_t smem = static_cast<size_t>(EX_QB) * dim * sizeof(float);
    VR_REQUIRE(smem <= 200 * 1024, "vr_score_exact: dim too large");
    static unsigned long long attr_set = 0;
    if (smem > 48 * 1024 && first_use_on_device(&attr_set))
        VR_CHECK_CUDA(cudaFuncSetAttribute(exact_scores_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    long long bx = (nd + 7) / 8;
    const long long cap = static_cast<long long>(num_sms()) * 4;
    if (bx > cap) bx = cap;
    dim3 grid(static_cast<unsigned>(bx), (nq + EX_QB - 1) / EX_QB);
    exact_scores_kernel<<<grid, 256, smem, reinterpret_cast<cudaStream_t>(stream)>>>(q_f32, nq, d_f32, nd, dim, scores);
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int vr_topk_rows(const float* scores, const int64_t* ids, int32_t rows, int64_t cols, int32_t k, int64_t id_offset,
                            float* out_scores, int64_t* out_ids, void* stream) {
    VR_REQUIRE(scores && out_scores && out_ids, "vr_topk_rows: null pointer");
    VR_REQUIRE(rows > 0 && cols > 0 && k > 0, "vr_topk_rows: bad shape");
    topk_rows_kernel<<<rows, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        scores, reinterpret_cast<const long long*>(ids), cols, k, id_offset, cols, out_scores,
        reinterpret_cast<long long*>(out_ids));
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
    topk_rows_kernel<<<rows, 256, 0, s>>>(ws_scores, reinterpret_cast<const long long*>(ws_ids),
                                          static_cast<long long>(chunks) * k, k, 0, static_cast<long long>(chunks) * k,
                                          out_scores, reinterpret_cast<long long*>(out_ids));
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
