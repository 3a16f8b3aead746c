// Device image front-end: Pillow-bit-compatible 8-bit separable resampling (the arithmetic of PIL.Image.resize(...,
// BICUBIC), i.e. Pillow src/libImaging/Resample.c ImagingResampleHorizontal_8bpc / ImagingResampleVertical_8bpc) with the
// grid crop of split_to_patches (modeling_minicpmv.py:571-592) folded into the output addressing.
//
// Both passes are byte/integer work bounded by memory traffic, not arithmetic: the horizontal pass reads every needed
// source row once (coalesced, staged in shared memory because each output pixel taps a window of ~4*scale pixels)
// and writes an 8-bit intermediate `out_w` wide; the vertical pass reads that intermediate (L2 resident: out_w x rows
// x 3 bytes per page) and writes the final pixels. Coefficients are 22-bit fixed point computed on the host exactly
// as Pillow does (visrag_b200/frontend.py); sums are 32-bit like Pillow's.
#include "common.h"

namespace vr {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

__device__ __forceinline__ uint8_t rs_clip8(int v) {
    v >>= RS_PRECISION_BITS;  // arithmetic shift, like Pillow's table lookup on (in >> PRECISION_BITS)
    return static_cast<uint8_t>(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// Where output pixel (y, x) of image `img` goes. Plain: [n, rows, w, 3]. Cells: the image is a grid of cell_h x cell_w
// crops, crop (cy, cx) of image img is slice  first[img] + cy*gx + cx  of a [*, cell_h, cell_w, 3] buffer.
struct RsDst {
    uint8_t* base;
    const int32_t* first;  // NULL -> plain layout
    int rows, pitch;       // plain layout: rows per image, bytes per row (a multiple of 4)
    int cell_h, cell_w, gx;
};
__device__ __forceinline__ uint8_t* rs_dst(const RsDst& d, int img, int y, int x) {
    if (d.first == nullptr) return d.base + (static_cast<int64_t>(img) * d.rows + y) * d.pitch + x * 3;
    const int cy = y / d.cell_h, ly = y - cy * d.cell_h;
    const int cx = x / d.cell_w, lx = x - cx * d.cell_w;
    const int64_t cell = static_cast<int64_t>(d.first[img]) + cy * d.gx + cx;
    return d.base + ((cell * d.cell_h + ly) * d.cell_w + lx) * 3;
}

// ---- horizontal pass: one block per (RS_ROWS source rows, image). bounds: [out_w][2]; kk: [ksize][out_w] (TAP-major, so
// that the threads of a warp - consecutive output pixels - read consecutive coefficients). Every tap coefficient is
// loaded once and applied to RS_ROWS staged rows.
constexpr int RS_ROWS = 4;

__device__ __forceinline__ void rs_stage_row(uint8_t* dst, const uint8_t* __restrict__ line, int nbytes) {
    // the row start is only byte aligned in general (3 bytes per pixel): peel to a 4-byte boundary, then 32-bit loads
    const int head = min(nbytes, static_cast<int>((4 - (reinterpret_cast<uintptr_t>(line) & 3)) & 3));
    for (int i = threadIdx.x; i < head; i += blockDim.x) dst[i] = line[i];
    const int words = (nbytes - head) >> 2;
    const uint32_t* lw = reinterpret_cast<const uint32_t*>(line + head);
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
        const uint32_t v = lw[i];
        uint8_t* p = dst + head + i * 4;
        p[0] = v & 255; p[1] = (v >> 8) & 255; p[2] = (v >> 16) & 255; p[3] = v >> 24;
    }
    for (int i = head + words * 4 + threadIdx.x; i < nbytes; i += blockDim.x) dst[i] = line[i];
}

__global__ void __launch_bounds__(256)
resample_h_kernel(const uint8_t* __restrict__ src, int in_h, int in_w, int ps, int row0, int rows,
                  const int32_t* __restrict__ bounds, const int32_t* __restrict__ kk, int out_w, RsDst dst) {
    extern __shared__ uint8_t rowbuf[];
    const int y0 = blockIdx.x * RS_ROWS, img = blockIdx.y;
    const int nbytes = in_w * ps;  // ps = source bytes per pixel: 3 (packed RGB) or 4 (Pillow's native RGBX rows)
    const int pitch = (nbytes + 3) & ~3;
    const int nrows = min(RS_ROWS, rows - y0);
    for (int r = 0; r < nrows; ++r)
        rs_stage_row(rowbuf + r * pitch, src + (static_cast<int64_t>(img) * in_h + row0 + y0 + r) * nbytes, nbytes);
    __syncthreads();
    for (int xx = threadIdx.x; xx < out_w; xx += blockDim.x) {
        const int xmin = bounds[2 * xx], n = bounds[2 * xx + 1];
        int acc[RS_ROWS][3];
#pragma unroll
        for (int r = 0; r < RS_ROWS; ++r) acc[r][0] = acc[r][1] = acc[r][2] = 1 << (RS_PRECISION_BITS - 1);
        const uint8_t* p = rowbuf + xmin * ps;
        for (int j = 0; j < n; ++j) {
            const int c = kk[static_cast<int64_t>(j) * out_w + xx];
#pragma unroll
            for (int r = 0; r < RS_ROWS; ++r) {
                const uint8_t* q = p + r * pitch + ps * j;  // rows past nrows read stale shared memory; never stored
                acc[r][0] += q[0] * c;
                acc[r][1] += q[1] * c;
                acc[r][2] += q[2] * c;
            }
        }
#pragma unroll
        for (int r = 0; r < RS_ROWS; ++r) {
            if (r < nrows) {
                uint8_t* o = rs_dst(dst, img, y0 + r, xx);
                o[0] = rs_clip8(acc[r][0]); o[1] = rs_clip8(acc[r][1]); o[2] = rs_clip8(acc[r][2]);
            }
        }
    }
}

// ---- vertical pass: one block per (output row, image); threads run over the bytes of the row. `pitch` = bytes between
// source rows. WORDS: the source is the 4-byte-pitched intermediate of the horizontal pass -> 32-bit loads, 4 bytes per thread.
template <bool WORDS>
__global__ void __launch_bounds__(256)
resample_v_kernel(const uint8_t* __restrict__ src, int in_rows, int w, int pitch, int ps, const int32_t* __restrict__ bounds,
                  const int32_t* __restrict__ kk, int ksize, int shift, RsDst dst) {
    const int yy = blockIdx.x, img = blockIdx.y;
    const int ymin = bounds[2 * yy] - shift, n = bounds[2 * yy + 1];
    const int32_t* k = kk + static_cast<int64_t>(yy) * ksize;
    const int nbytes = w * 3;
    const uint8_t* col = src + (static_cast<int64_t>(img) * in_rows + ymin) * pitch;
    if (WORDS) {
        for (int wi = threadIdx.x; wi * 4 < nbytes; wi += blockDim.x) {
            int s0 = 1 << (RS_PRECISION_BITS - 1), s1 = s0, s2 = s0, s3 = s0;
            for (int j = 0; j < n; ++j) {
                const uint32_t v = *reinterpret_cast<const uint32_t*>(col + static_cast<int64_t>(j) * pitch + wi * 4);
                const int c = k[j];
                s0 += static_cast<int>(v & 255) * c;
                s1 += static_cast<int>((v >> 8) & 255) * c;
                s2 += static_cast<int>((v >> 16) & 255) * c;
                s3 += static_cast<int>(v >> 24) * c;
            }
            const int sv[4] = {s0, s1, s2, s3};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
                const int b = wi * 4 + t;
                if (b < nbytes) {
                    const int x = b / 3;
                    rs_dst(dst, img, yy, x)[b - x * 3] = rs_clip8(sv[t]);
                }
            }
        }
    } else {
        for (int b = threadIdx.x; b < nbytes; b += blockDim.x) {
            const int x = b / 3, ch = b - x * 3;
            const int off = x * ps + ch;  // byte of (pixel x, channel ch) inside a source row
            int s = 1 << (RS_PRECISION_BITS - 1);
            for (int j = 0; j < n; ++j) s += col[static_cast<int64_t>(j) * pitch + off] * k[j];
            rs_dst(dst, img, yy, x)[ch] = rs_clip8(s);
        }
    }
}

// ---- no resampling at all (Image.resize to the same size is a copy): scatter rows into the cell layout
__global__ void __launch_bounds__(256)
resample_copy_kernel(const uint8_t* __restrict__ src, int h, int w, int ps, RsDst dst) {
    const int y = blockIdx.x, img = blockIdx.y;
    const uint8_t* line = src + (static_cast<int64_t>(img) * h + y) * w * ps;
    if (ps == 4) {  // RGBX source: one 32-bit load per pixel, three byte stores
        const uint32_t* lw = reinterpret_cast<const uint32_t*>(line);
        for (int x = threadIdx.x; x < w; x += blockDim.x) {
            const uint32_t v = lw[x];
            uint8_t* o = rs_dst(dst, img, y, x);
            o[0] = v & 255; o[1] = (v >> 8) & 255; o[2] = (v >> 16) & 255;
        }
    } else {
        for (int b = threadIdx.x; b < w * 3; b += blockDim.x) {
            const int x = b / 3;
            rs_dst(dst, img, y, x)[b - x * 3] = line[b];
        }
    }
}

}  // namespace vr

extern "C" int vr_resample_u8(const uint8_t* src, int32_t src_pixel_bytes, int32_t n, int32_t in_h, int32_t in_w,
                              const int32_t* bounds_h,
                              const int32_t* coeffs_h, int32_t ksize_h, const int32_t* bounds_v, const int32_t* coeffs_v,
                              int32_t ksize_v, int32_t row_first, int32_t row_count, int32_t out_h, int32_t out_w,
                              uint8_t* tmp, uint8_t* out, const int32_t* first_cell, int32_t cell_h, int32_t cell_w,
                              void* stream) {
    using namespace vr;
    VR_REQUIRE(src && out && first_cell, "vr_resample_u8: null pointer argument");
    VR_REQUIRE(n > 0 && in_h > 0 && in_w > 0 && out_h > 0 && out_w > 0, "vr_resample_u8: empty problem");
    VR_REQUIRE(cell_h > 0 && cell_w > 0 && out_h % cell_h == 0 && out_w % cell_w == 0,
               "vr_resample_u8: the output %dx%d is not a whole grid of %dx%d cells", out_w, out_h, cell_w, cell_h);
    const bool need_h = bounds_h != nullptr, need_v = bounds_v != nullptr;
    VR_REQUIRE(need_h || in_w == out_w, "vr_resample_u8: width changes but no horizontal coefficients were given");
    VR_REQUIRE(need_v || in_h == out_h, "vr_resample_u8: height changes but no vertical coefficients were given");
    VR_REQUIRE(!need_h || (coeffs_h && ksize_h > 0), "vr_resample_u8: horizontal pass needs coefficients");
    VR_REQUIRE(!need_v || (coeffs_v && ksize_v > 0), "vr_resample_u8: vertical pass needs coefficients");
    VR_REQUIRE(!(need_h && need_v) || tmp, "vr_resample_u8: two passes need the intermediate buffer");
    VR_REQUIRE(in_w <= 12288, "vr_resample_u8: rows wider than 12288 pixels are not supported");
    const int ps = src_pixel_bytes;
    VR_REQUIRE(ps == 3 || ps == 4, "vr_resample_u8: src_pixel_bytes must be 3 (RGB) or 4 (RGBX), got %d", ps);
    VR_REQUIRE(ps == 3 || (reinterpret_cast<uintptr_t>(src) & 3) == 0, "vr_resample_u8: RGBX sources must be 4-byte aligned");
    cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
    RsDst cells{out, first_cell, 0, 0, cell_h, cell_w, out_w / cell_w};
    const int tmp_pitch = (out_w * 3 + 3) & ~3;  // rows of the intermediate start on 4-byte boundaries
    if (!need_h && !need_v) {
        resample_copy_kernel<<<dim3(in_h, n), 256, 0, s>>>(src, in_h, in_w, ps, cells);
        VR_CHECK_CUDA(cudaGetLastError());
        return 0;
    }
    const uint8_t* vsrc = src;
    int vrows = in_h, shift = 0, vpitch = in_w * ps, vps = ps;
    if (need_h) {
        // Pillow runs the horizontal pass over source rows [row_first, row_first + row_count) only (the rows the
        // vertical pass reads); with no vertical pass that is every row
        const int r0 = need_v ? row_first : 0, rc = need_v ? row_count : in_h;
        VR_REQUIRE(r0 >= 0 && rc > 0 && r0 + rc <= in_h, "vr_resample_u8: bad source row range [%d, %d)", r0, r0 + rc);
        RsDst hd = need_v ? RsDst{tmp, nullptr, rc, tmp_pitch, 0, 0, 0} : cells;
        const size_t smem = static_cast<size_t>(RS_ROWS) * ((static_cast<size_t>(in_w) * ps + 3) & ~size_t(3));
        static unsigned long long smem_set = 0;
        if (smem > 48 * 1024 && first_use_on_device(&smem_set))
            VR_CHECK_CUDA(cudaFuncSetAttribute(resample_h_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        resample_h_kernel<<<dim3((rc + RS_ROWS - 1) / RS_ROWS, n), 256, smem, s>>>(src, in_h, in_w, ps, r0, rc, bounds_h,
                                                                                 coeffs_h, out_w, hd);
        VR_CHECK_CUDA(cudaGetLastError());
        vsrc = tmp; vrows = rc; shift = r0; vpitch = tmp_pitch; vps = 3;
    }
    if (need_v) {
        if (need_h && (reinterpret_cast<uintptr_t>(tmp) & 3) == 0)
            resample_v_kernel<true><<<dim3(out_h, n), 256, 0, s>>>(vsrc, vrows, out_w, vpitch, vps, bounds_v, coeffs_v, ksize_v, shift, cells);
        else
            resample_v_kernel<false><<<dim3(out_h, n), 256, 0, s>>>(vsrc, vrows, out_w, vpitch, vps, bounds_v, coeffs_v, ksize_v, shift, cells);
        VR_CHECK_CUDA(cudaGetLastError());
    }
    return 0;
}
