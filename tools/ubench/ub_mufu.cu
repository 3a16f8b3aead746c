// MUFU.EX2 issue cadence per sub-partition as a function of the number of warps issuing it (sm_100a).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ub_mufu ub_mufu.cu && ./ub_mufu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void k_ex2(float* out, long long* cyc, int iters) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    const long long t1 = clock64();
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 1 << 22);
    cudaMalloc(&cyc, 1 << 16);
    const int iters = 256;
    for (int threads : {32, 128, 256, 384, 512, 1024}) {
        k_ex2<<<148, threads>>>(out, cyc, iters);
        k_ex2<<<148, threads>>>(out, cyc, iters);
        cudaDeviceSynchronize();
        long long h[148];
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < 148; ++i) avg += h[i];
        avg /= 148;
        const int warps = threads / 32, per_smsp = (warps + 3) / 4;
        const double instr_per_warp = 16.0 * iters;
        printf("%4d threads/CTA (%d warp(s) per sub-partition): %.1f cycles per MUFU.EX2 warp-instruction per warp, %.2f per sub-partition\n",
               threads, per_smsp, avg / instr_per_warp, avg / (instr_per_warp * per_smsp));
    }
    return 0;
}
