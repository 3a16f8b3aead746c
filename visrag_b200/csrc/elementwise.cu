// HBM-bound kernels of the encode path: patch unfold + normalise, LayerNorm, RMSNorm, LM input
// assembly, final norm + pooling + L2 normalise. All of them stream the activation once with 16-byte
// accesses; row statistics are computed in fp32 with a warp per row (rows are <= 9 KB, so the second
// and third sweep of a row hit L1).
#include "common.h"
#include "ptx.cuh"
#include "../../include/visrag_b200.h"

namespace vr {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// ---------------------------------------------------------------------------------------------
// im2col + normalise. One thread produces 8 consecutive output columns (one 16-byte store).
// ---------------------------------------------------------------------------------------------
__global__ void im2col_norm_kernel(const uint8_t* __restrict__ px, int n_slices, int h, int w, int patch,
                                   __nv_bfloat16* __restrict__ out, long long ldo) {
    const int gh = h / patch, gw = w / patch;
    const int groups = static_cast<int>(ldo / 8);
    const long long total = static_cast<long long>(n_slices) * gh * gw * groups;
    const int pp = patch * patch, kvalid = 3 * pp;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < total;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
        const int g = static_cast<int>(i % groups);
        const long long prow = i / groups;
        const int pxi = static_cast<int>(prow % gw);
        const int pyi = static_cast<int>((prow / gw) % gh);
        const int s = static_cast<int>(prow / (static_cast<long long>(gw) * gh));
        const uint8_t* img = px + static_cast<long long>(s) * h * w * 3;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int col = g * 8 + j;
            if (col < kvalid) {
                const int c = col / pp, rem = col - c * pp;
                const int ky = rem / patch, kx = rem - ky * patch;
                const int y = pyi * patch + ky, x = pxi * patch + kx;
                const float u = static_cast<float>(img[(static_cast<long long>(y) * w + x) * 3 + c]);
                v[j] = (u / 255.0f - 0.5f) / 0.5f;  // ToTensor then Normalize(0.5, 0.5), fp32 like torchvision
            } else {
                v[j] = 0.f;
            }
        }
        uint4 pk;
        pk.x = pack_bf16x2(v[0], v[1]);
        pk.y = pack_bf16x2(v[2], v[3]);
        pk.z = pack_bf16x2(v[4], v[5]);
        pk.w = pack_bf16x2(v[6], v[7]);
        *reinterpret_cast<uint4*>(out + prow * ldo + g * 8) = pk;
    }
}

// ---------------------------------------------------------------------------------------------
// LayerNorm / RMSNorm: one warp per row, dim % 4 == 0.
// ---------------------------------------------------------------------------------------------
template <bool RMS>
__global__ void norm_kernel(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma,
                            const float* __restrict__ beta, float eps, int rows, int dim,
                            __nv_bfloat16* __restrict__ out, long long ldo, __nv_bfloat16* __restrict__ out2,
                            const float* __restrict__ add, int add_period) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    const int nvec = dim >> 2;
    for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
        const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * ldx);
        float mean = 0.f;
        if (!RMS) {
            float s = 0.f;
            for (int i = lane; i < nvec; i += 32) {
                const float4 v = xr[i];
                s += (v.x + v.y) + (v.z + v.w);
            }
            mean = warp_sum(s) / static_cast<float>(dim);
        }
        float ss = 0.f;
        for (int i = lane; i < nvec; i += 32) {
            const float4 v = xr[i];
            const float a = v.x - mean, b = v.y - mean, c = v.z - mean, d = v.w - mean;
            ss += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(warp_sum(ss) / static_cast<float>(dim) + eps);
        __nv_bfloat16* orow = out + static_cast<long long>(row) * ldo;
        __nv_bfloat16* orow2 = out2 ? out2 + static_cast<long long>(row) * ldo : nullptr;
        const float* arow = add ? add + static_cast<long long>(row % add_period) * dim : nullptr;
        for (int i = lane; i < nvec; i += 32) {
            const float4 v = xr[i];
            const float4 g = reinterpret_cast<const float4*>(gamma)[i];
            float y0 = (v.x - mean) * rstd * g.x, y1 = (v.y - mean) * rstd * g.y;
            float y2 = (v.z - mean) * rstd * g.z, y3 = (v.w - mean) * rstd * g.w;
            if (!RMS) {
                const float4 bb = reinterpret_cast<const float4*>(beta)[i];
                y0 += bb.x; y1 += bb.y; y2 += bb.z; y3 += bb.w;
            }
            uint2 pk;
            pk.x = pack_bf16x2(y0, y1);
            pk.y = pack_bf16x2(y2, y3);
            reinterpret_cast<uint2*>(orow)[i] = pk;
            if (orow2) {
                const float4 aa = reinterpret_cast<const float4*>(arow)[i];
                uint2 pk2;
                pk2.x = pack_bf16x2(y0 + aa.x, y1 + aa.y);
                pk2.y = pack_bf16x2(y2 + aa.z, y3 + aa.w);
                reinterpret_cast<uint2*>(orow2)[i] = pk2;
            }
        }
    }
}

// Register-resident variant for dim == VPL * 128 (1152 -> VPL 9, 2304 -> VPL 18): the row is read from HBM exactly once.
template <bool RMS, int VPL>
__global__ void __launch_bounds__(256)
norm_kernel_reg(const float* __restrict__ x, long long ldx, const float* __restrict__ gamma, const float* __restrict__ beta,
                float eps, int rows, __nv_bfloat16* __restrict__ out, long long ldo, __nv_bfloat16* __restrict__ out2,
                const float* __restrict__ add, int add_period) {
    constexpr int DIM = VPL * 128;
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int row = blockIdx.x * warps_per_block + (threadIdx.x >> 5); row < rows; row += gridDim.x * warps_per_block) {
        const float4* xr = reinterpret_cast<const float4*>(x + static_cast<long long>(row) * ldx);
        float4 v[VPL];
#pragma unroll
        for (int i = 0; i < VPL; ++i) v[i] = xr[lane + i * 32];
        float mean = 0.f;
        if (!RMS) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < VPL; ++i) s += (v[i].x + v[i].y) + (v[i].z + v[i].w);
            mean = warp_sum(s) * (1.0f / DIM);
        }
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const float a = v[i].x - mean, b = v[i].y - mean, c = v[i].z - mean, d = v[i].w - mean;
            ss += (a * a + b * b) + (c * c + d * d);
        }
        const float rstd = rsqrtf(warp_sum(ss) * (1.0f / DIM) + eps);
        uint2* orow = reinterpret_cast<uint2*>(out + static_cast<long long>(row) * ldo);
        uint2* orow2 = out2 ? reinterpret_cast<uint2*>(out2 + static_cast<long long>(row) * ldo) : nullptr;
        const float4* arow = add ? reinterpret_cast<const float4*>(add + static_cast<long long>(row % add_period) * DIM) : nullptr;
#pragma unroll
        for (int i = 0; i < VPL; ++i) {
            const int c = lane + i * 32;
            const float4 g = reinterpret_cast<const float4*>(gamma)[c];
            float y0 = (v[i].x - mean) * rstd * g.x, y1 = (v[i].y - mean) * rstd * g.y;
            float y2 = (v[i].z - mean) * rstd * g.z, y3 = (v[i].w - mean) * rstd * g.w;
            if (!RMS) {
                const float4 bb = reinterpret_cast<const float4*>(beta)[c];
                y0 += bb.x; y1 += bb.y; y2 += bb.z; y3 += bb.w;
            }
            uint2 pk;
            pk.x = pack_bf16x2(y0, y1);
            pk.y = pack_bf16x2(y2, y3);
            orow[c] = pk;
            if (orow2) {
                const float4 aa = arow[c];
                uint2 pk2;
                pk2.x = pack_bf16x2(y0 + aa.x, y1 + aa.y);
                pk2.y = pack_bf16x2(y2 + aa.z, y3 + aa.w);
                orow2[c] = pk2;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// LM input assembly: one warp per token row.
// ---------------------------------------------------------------------------------------------
__global__ void build_lm_input_kernel(const int* __restrict__ src, int tokens, int dim,
                                      const __nv_bfloat16* __restrict__ embed, float scale_emb,
                                      const float* __restrict__ vision, long long ldv, float* __restrict__ h,
                                      long long ldh) {
    const int warps_per_block = blockDim.x >> 5;
    const int lane = threadIdx.x & 31;
    for (int t = blockIdx.x * warps_per_block + (threadIdx.x >> 5); t < tokens; t += gridDim.x * warps_per_block) {
        const int s = src[t];
        float4* dst = reinterpret_cast<float4*>(h + static_cast<long long>(t) * ldh);
        if (s >= 0) {
            const float4* v = reinterpret_cast<const float4*>(vision + static_cast<long long>(s) * ldv);
            for (int i = lane; i < (dim >> 2); i += 32) dst[i] = v[i];
        } else {
            const uint2* e = reinterpret_cast<const uint2*>(embed + static_cast<long long>(-(s + 1)) * dim);
            for (int i = lane; i < (dim >> 2); i += 32) {
                const uint2 raw = e[i];
                const __nv_bfloat162 a = *reinterpret_cast<const __nv_bfloat162*>(&raw.x);
                const __nv_bfloat162 b = *reinterpret_cast<const __nv_bfloat162*>(&raw.y);
                dst[i] = make_float4(__low2float(a) * scale_emb, __high2float(a) * scale_emb,
                                     __low2float(b) * scale_emb, __high2float(b) * scale_emb);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Final RMSNorm + pooling + L2 normalise: one CTA per sequence.
//   phase 1: warps compute 1/rms of every row of the sequence into shared memory;
//   phase 2: thread-per-column weighted accumulation over the rows (coalesced across threads);
//   phase 3: block reduction of the squared norm, normalise, write.
// ---------------------------------------------------------------------------------------------
constexpr int POOL_THREADS = 256;
constexpr int POOL_MAX_COLS_PER_THREAD = 16;  // dim <= 4096

__global__ void __launch_bounds__(POOL_THREADS)
pool_norm_kernel(const float* __restrict__ h, long long ldh, const float* __restrict__ gamma, float eps,
                 const int* __restrict__ cu, int dim, int pooling, int normalize, float* __restrict__ reps) {
    extern __shared__ float inv_rms[];  // [len]
    __shared__ float red[POOL_THREADS / 32];
    __shared__ float total;
    const int b = blockIdx.x;
    const int begin = cu[b], len = cu[b + 1] - begin;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float* out = reps + static_cast<long long>(b) * dim;
    if (len <= 0) {
        for (int c = threadIdx.x; c < dim; c += POOL_THREADS) out[c] = 0.f;
        return;
    }
    int t_lo = 0, t_hi = len;  // rows that carry weight
    if (pooling == 2) t_lo = len - 1;
    if (pooling == 3) t_hi = 1;
    for (int t = t_lo + warp; t < t_hi; t += POOL_THREADS / 32) {
        const float4* xr = reinterpret_cast<const float4*>(h + static_cast<long long>(begin + t) * ldh);
        float ss = 0.f;
        for (int i = lane; i < (dim >> 2); i += 32) {
            const float4 v = xr[i];
            ss += (v.x * v.x + v.y * v.y) + (v.z * v.z + v.w * v.w);
        }
        ss = warp_sum(ss);
        if (lane == 0) inv_rms[t] = rsqrtf(ss / static_cast<float>(dim) + eps);
    }
    __syncthreads();
    float acc[POOL_MAX_COLS_PER_THREAD];
#pragma unroll
    for (int j = 0; j < POOL_MAX_COLS_PER_THREAD; ++j) acc[j] = 0.f;
    float wsum = 0.f;
    for (int t = t_lo; t < t_hi; ++t) {
        const float w = (pooling == 0) ? static_cast<float>(t + 1) : 1.0f;
        wsum += w;
        const float s = w * inv_rms[t];
        const float* xr = h + static_cast<long long>(begin + t) * ldh;
#pragma unroll
        for (int j = 0; j < POOL_MAX_COLS_PER_THREAD; ++j) {
            const int c = threadIdx.x + j * POOL_THREADS;
            if (c < dim) acc[j] += s * xr[c];
        }
    }
    float sq = 0.f;
#pragma unroll
    for (int j = 0; j < POOL_MAX_COLS_PER_THREAD; ++j) {
        const int c = threadIdx.x + j * POOL_THREADS;
        if (c < dim) {
            acc[j] = acc[j] * gamma[c] / wsum;
            sq += acc[j] * acc[j];
        }
    }
    sq = warp_sum(sq);
    if (lane == 0) red[warp] = sq;
    __syncthreads();
    if (threadIdx.x == 0) {
        float s = 0.f;
        for (int i = 0; i < POOL_THREADS / 32; ++i) s += red[i];
        total = s;
    }
    __syncthreads();
    const float inv = normalize ? 1.0f / fmaxf(sqrtf(total), 1e-12f) : 1.0f;
#pragma unroll
    for (int j = 0; j < POOL_MAX_COLS_PER_THREAD; ++j) {
        const int c = threadIdx.x + j * POOL_THREADS;
        if (c < dim) out[c] = acc[j] * inv;
    }
}

static int grid_for(long long work_items, int per_block) {
    long long blocks = (work_items + per_block - 1) / per_block;
    const long long cap = static_cast<long long>(num_sms()) * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return static_cast<int>(blocks);
}

}  // namespace vr

using namespace vr;

extern "C" int vr_im2col_norm(const uint8_t* pixels, int32_t n_slices, int32_t h, int32_t w, int32_t patch, void* out,
                              int64_t ldo, void* stream) {
    VR_REQUIRE(pixels && out, "vr_im2col_norm: null pointer");
    VR_REQUIRE(n_slices > 0 && h > 0 && w > 0 && patch > 0 && h % patch == 0 && w % patch == 0,
               "vr_im2col_norm: bad geometry n=%d h=%d w=%d patch=%d", n_slices, h, w, patch);
    VR_REQUIRE(ldo % 8 == 0 && ldo >= 3 * patch * patch, "vr_im2col_norm: ldo=%lld must be a multiple of 8 and >= %d",
               (long long)ldo, 3 * patch * patch);
    const long long total = static_cast<long long>(n_slices) * (h / patch) * (w / patch) * (ldo / 8);
    im2col_norm_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        pixels, n_slices, h, w, patch, reinterpret_cast<__nv_bfloat16*>(out), ldo);
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int vr_layernorm(const float* x, int64_t ldx, const float* gamma, const float* beta, float eps, int32_t rows,
                            int32_t dim, void* out, int64_t ldo, void* out2, const float* add, int32_t add_period,
                            void* stream) {
    VR_REQUIRE(x && gamma && beta && out, "vr_layernorm: null pointer");
    VR_REQUIRE(rows > 0 && dim > 0 && dim % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "vr_layernorm: bad shape rows=%d dim=%d",
               rows, dim);
    VR_REQUIRE(!out2 || (add && add_period > 0), "vr_layernorm: out2 needs add/add_period");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    __nv_bfloat16* o2 = reinterpret_cast<__nv_bfloat16*>(out2);
    if (dim == 1152)
        norm_kernel_reg<false, 9><<<grid_for(rows, 8), 256, 0, s>>>(x, ldx, gamma, beta, eps, rows, o, ldo, o2, add, add_period);
    else if (dim == 2304)
        norm_kernel_reg<false, 18><<<grid_for(rows, 8), 256, 0, s>>>(x, ldx, gamma, beta, eps, rows, o, ldo, o2, add, add_period);
    else
        norm_kernel<false><<<grid_for(rows, 8), 256, 0, s>>>(x, ldx, gamma, beta, eps, rows, dim, o, ldo, o2, add, add_period);
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int vr_rmsnorm(const float* x, int64_t ldx, const float* gamma, float eps, int32_t rows, int32_t dim, void* out,
                          int64_t ldo, void* stream) {
    VR_REQUIRE(x && gamma && out, "vr_rmsnorm: null pointer");
    VR_REQUIRE(rows > 0 && dim > 0 && dim % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0, "vr_rmsnorm: bad shape rows=%d dim=%d",
               rows, dim);
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(out);
    if (dim == 2304)
        norm_kernel_reg<true, 18><<<grid_for(rows, 8), 256, 0, s>>>(x, ldx, gamma, nullptr, eps, rows, o, ldo, nullptr, nullptr, 1);
    else
        norm_kernel<true><<<grid_for(rows, 8), 256, 0, s>>>(x, ldx, gamma, nullptr, eps, rows, dim, o, ldo, nullptr, nullptr, 1);
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int vr_build_lm_input(const int32_t* src, int32_t tokens, int32_t dim, const void* embed_bf16, float scale_emb,
                                 const float* vision, int64_t ldv, float* h, int64_t ldh, void* stream) {
    VR_REQUIRE(src && embed_bf16 && h, "vr_build_lm_input: null pointer");
    VR_REQUIRE(tokens > 0 && dim % 4 == 0 && ldh % 4 == 0 && (vision == nullptr || ldv % 4 == 0),
               "vr_build_lm_input: bad shape tokens=%d dim=%d", tokens, dim);
    build_lm_input_kernel<<<grid_for(tokens, 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
        src, tokens, dim, reinterpret_cast<const __nv_bfloat16*>(embed_bf16), scale_emb, vision, ldv, h, ldh);
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}

extern "C" int vr_pool_norm(const float* h, int64_t ldh, const float* gamma, float eps, const int32_t* cu, int32_t batch,
                            int32_t dim, int32_t pooling, int32_t normalize, float* reps, void* stream) {
    VR_REQUIRE(h && gamma && cu && reps, "vr_pool_norm: null pointer");
    VR_REQUIRE(batch > 0 && dim > 0 && dim % 4 == 0 && dim <= POOL_THREADS * POOL_MAX_COLS_PER_THREAD && ldh % 4 == 0,
               "vr_pool_norm: bad shape batch=%d dim=%d", batch, dim);
    VR_REQUIRE(pooling >= 0 && pooling <= 3, "vr_pool_norm: pooling must be 0..3");
    const int smem = 2048 * sizeof(float) * 4;  // sequences up to 8192 tokens
    pool_norm_kernel<<<batch, POOL_THREADS, smem, reinterpret_cast<cudaStream_t>(stream)>>>(h, ldh, gamma, eps, cu, dim,
                                                                                          pooling, normalize, reps);
    VR_CHECK_CUDA(cudaGetLastError());
    return 0;
}
