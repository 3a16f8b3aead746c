// Microbenchmarks that size the attention kernel's engines on sm_100a (build: make -C tools/ubench; run on the GPU box).
//   1. tcgen05.mma SS (A,B in smem) cycles per K=16 instruction for N = 64..256
//   2. tcgen05.mma TS (A in TMEM)   cycles per K=16 instruction for N = 16..128
//   3. tcgen05.ld 32x32b.x32 throughput with 1 / 2 warps per SM sub-partition
//   4. ex2.approx.ftz.f32 vs ex2.approx.f16x2 issue rate
//   5. the same loads while the tensor core runs SS / TS MMAs (contention)
//   6. kind::f16 with A = f16 (TMEM) and B = bf16 (smem): is the mix legal and correct?
#include <cstdio>
#include <cstdlib>
#include "../../visrag_b200/csrc/ptx.cuh"
using namespace vr;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Res { long long mma_cycles; long long ld_cycles; long long aux; float val; };

// mode 0: SS  mode 1: TS ; n = MMA N ; reps MMAs ; ld_warps: number of warps (0, 4 or 8) running LDTM loops of ld_reps x 4 loads
__global__ void __launch_bounds__(384, 1) k_mma_ld(int mode, int n, int reps, int ld_warps, int ld_reps, int swz32, Res* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 8) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    if (warp == 8) {
        if (reps > 0) {
            const uint32_t idesc = make_idesc_f16(128, n, 1, 0, 0);
            const uint64_t hi = swz32 ? make_smem_desc(0, 16, 256, kLayoutSW32) : make_smem_desc(0, 16, 1024, kLayoutSW128);
            const uint32_t a16 = smem_u32(smem) >> 4, b16 = smem_u32(smem + 32768) >> 4;
            long long t0 = clock64();
            if (elect_one()) {
                for (int r = 0; r < reps; ++r) {
                    const uint32_t ko = swz32 ? 0 : (r & 3) * 2;
                    if (mode == 0) umma_f16_ss(tm, hi | (a16 + ko), hi | (b16 + ko), idesc, r != 0);
                    else umma_f16_ts(tm, tm + 256 + (r & 3) * 8, hi | (b16 + ko), idesc, r != 0);
                }
                umma_commit(&bar);
            }
            __syncwarp();
            mbar_wait(&bar, 0);
            long long t1 = clock64();
            if ((threadIdx.x & 31) == 0) out->mma_cycles = t1 - t0;
        }
    } else if (warp < ld_warps) {
        const uint32_t lane_off = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t base = tm + lane_off + (warp >= 4 ? 128 : 0) + 256;  // read columns away from the accumulator
        uint32_t acc = 0;
        long long t0 = clock64();
        for (int r = 0; r < ld_reps; ++r) {
            uint32_t v0[32], v1[32], v2[32], v3[32];
            tmem_ld_32x32(base, v0);
            tmem_ld_32x32(base + 32, v1);
            tmem_ld_32x32(base + 64, v2);
            tmem_ld_32x32(base + 96, v3);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc += v0[j] ^ v1[j] ^ v2[j] ^ v3[j];
        }
        long long t1 = clock64();
        if (threadIdx.x == 0) { out->ld_cycles = t1 - t0; out->aux = acc; }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

// MUFU issue rate: warps_per_smsp x 4 warps, each thread runs `reps` x 8 independent ex2
__global__ void k_ex2(int use_f16x2, int reps, Res* out) {
    float x[8];
    uint32_t h[8];
    for (int i = 0; i < 8; ++i) { x[i] = -0.001f * (threadIdx.x + i); h[i] = 0xb800b400u + i + threadIdx.x; }
    __syncthreads();
    long long t0 = clock64();
    if (use_f16x2) {
        for (int r = 0; r < reps; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(h[i]));
    } else {
        for (int r = 0; r < reps; ++r)
#pragma unroll
            for (int i = 0; i < 8; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
    }
    long long t1 = clock64();
    float s = 0; uint32_t hs = 0;
    for (int i = 0; i < 8; ++i) { s += x[i]; hs ^= h[i]; }
    if (threadIdx.x == 0) { out->mma_cycles = t1 - t0; out->val = s + hs; }
}

// A = f16 ones in TMEM (packed pairs), B = bf16 3.0 in smem, K = 16 -> 48 if the f16 x bf16 mix is honoured
__global__ void __launch_bounds__(128, 1) k_mix(int a_fmt, int b_fmt, uint32_t a_bits, uint32_t b_bits, Res* out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = b_bits;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<512>(&slot);
    fence_proxy_async_smem();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = slot;
    const uint32_t lane_off = static_cast<uint32_t>(warp * 32) << 16;
    uint32_t a[8];
    for (int i = 0; i < 8; ++i) a[i] = a_bits;
    tmem_st_32x8(tm + lane_off + 256, a);
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 0) {
        const uint32_t idesc = (1u << 4) | (uint32_t(a_fmt) << 7) | (uint32_t(b_fmt) << 10) | ((64u >> 3) << 17) | ((128u >> 4) << 24);
        const uint64_t hi = make_smem_desc(0, 16, 1024, kLayoutSW128);
        if (elect_one()) {
            umma_f16_ts(tm, tm + 256, hi | (smem_u32(smem) >> 4), idesc, 0);
            umma_commit(&bar);
        }
        __syncwarp();
    }
    mbar_wait(&bar, 0);
    tc_fence_after();
    uint32_t v[16];
    tmem_ld_32x16(tm + lane_off, v);
    tmem_ld_wait();
    if (threadIdx.x == 37) out->val = __uint_as_float(v[5]);
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { tc_fence_after(); tmem_dealloc<512>(tm); }
}

// 7. hand-off latencies: (a) tcgen05.commit with nothing pending -> own mbarrier wait; (b) warp-to-warp ping-pong through
//    mbarrier.arrive / try_wait; (c) the same with tcgen05.commit on one side (the issuer's side of the attention pipeline)
__global__ void __launch_bounds__(128, 1) k_handoff(int mode, int reps, Res* out) {
    __shared__ uint64_t bar_a, bar_b;
    __shared__ uint32_t slot;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) { mbar_init(&bar_a, 1); mbar_init(&bar_b, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<32>(&slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    long long t0 = clock64();
    if (mode == 0) {
        if (warp == 0) {
            for (int r = 0; r < reps; ++r) {
                if (elect_one()) umma_commit(&bar_a);
                __syncwarp();
                mbar_wait(&bar_a, r & 1);
            }
        }
    } else {
        // warp 0: arrive A (plain or commit), wait B.   warp 1: wait A, arrive B
        for (int r = 0; r < reps; ++r) {
            if (warp == 0) {
                if (mode == 1) { if (lane == 0) mbar_arrive(&bar_a); }
                else { if (elect_one()) umma_commit(&bar_a); }
                __syncwarp();
                mbar_wait(&bar_b, r & 1);
            } else if (warp == 1) {
                mbar_wait(&bar_a, r & 1);
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_b);
            }
        }
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) out->mma_cycles = t1 - t0;
    __syncthreads();
    if (warp == 0) tmem_dealloc<32>(slot);
}

int main() {
    Res* d; Res h;
    CK(cudaMalloc(&d, sizeof(Res)));
    const int SM = 100 * 1024;
    CK(cudaFuncSetAttribute(k_mma_ld, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
    CK(cudaFuncSetAttribute(k_mix, cudaFuncAttributeMaxDynamicSharedMemorySize, SM));
    auto run = [&](int mode, int n, int reps, int ldw, int ldr, int sw32) {
        CK(cudaMemset(d, 0, sizeof(Res)));
        k_mma_ld<<<1, 384, SM>>>(mode, n, reps, ldw, ldr, sw32, d);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost));
    };
    run(0, 128, 64, 0, 0, 0);  // warm-up
    printf("== 1. SS MMA (M=128, K=16): cycles per instruction\n");
    for (int sw = 0; sw < 2; ++sw)
        for (int n : {64, 80, 96, 112, 128, 160, 256}) { run(0, n, 256, 0, 0, sw); printf("  %s N=%3d : %.1f\n", sw ? "SW32 " : "SW128", n, h.mma_cycles / 256.0); }
    printf("== 2. TS MMA (A in TMEM)\n");
    for (int n : {16, 64, 80, 96, 128, 256}) { run(1, n, 256, 0, 0, 0); printf("  N=%3d : %.1f\n", n, h.mma_cycles / 256.0); }
    printf("== 3. LDTM 32x32b.x32, 128 fp32 columns per warp-row-block (16 KB per warp iteration)\n");
    for (int w : {1, 4, 8}) { run(0, 128, 0, w, 64, 0); printf("  %d warps: %.0f cycles per 128-column row block (per warp), %.1f B/clk/SM\n", w, h.ld_cycles / 64.0, w * 16384.0 * 64 / h.ld_cycles); }
    printf("== 5. contention: LDTM (8 warps) while MMAs run\n");
    for (int mode = 0; mode < 2; ++mode)
        for (int n : {128, 80}) {
            run(mode, n, 1024, 8, 64, 0);
            printf("  %s N=%3d: mma %.1f cyc/instr, ld %.0f cycles per row block\n", mode ? "TS" : "SS", n, h.mma_cycles / 1024.0, h.ld_cycles / 64.0);
        }
    printf("== 4. ex2 issue rate (cycles per warp instruction per SM sub-partition)\n");
    for (int f16 = 0; f16 < 2; ++f16)
        for (int warps : {4, 8}) {
            k_ex2<<<1, warps * 32>>>(f16, 512, d);
            CK(cudaDeviceSynchronize());
            CK(cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost));
            printf("  %s %d warps: %.2f cycles per warp-instruction on its SMSP\n", f16 ? "f16x2" : "f32  ", warps, h.mma_cycles / (512.0 * 8) / (warps / 4));
        }
    printf("== 7. hand-off latency (cycles per round)\n");
    for (int mode = 0; mode < 3; ++mode) {
        CK(cudaMemset(d, 0, sizeof(Res)));
        k_handoff<<<1, 128>>>(mode, 1000, d);
        CK(cudaDeviceSynchronize());
        CK(cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost));
        const char* names[3] = {"tcgen05.commit (nothing pending) -> own wait", "ping-pong arrive/wait between two warps (round trip)",
                                "ping-pong with tcgen05.commit on one side (round trip)"};
        printf("  %s: %.0f\n", names[mode], h.mma_cycles / 1000.0);
    }
    printf("== 6. kind::f16 operand type mix (A in TMEM x B in smem), expect 16*a*b\n");
    struct { int af, bf; uint32_t ab, bb; const char* name; float expect; } cases[] = {
        {1, 1, 0x3f803f80u, 0x40404040u, "A bf16 1.0 x B bf16 3.0", 48.f},
        {0, 0, 0x3c003c00u, 0x42004200u, "A f16 1.0  x B f16 3.0 ", 48.f},
        {0, 1, 0x3c003c00u, 0x40404040u, "A f16 1.0  x B bf16 3.0", 48.f},
    };
    for (auto& c : cases) {
        CK(cudaMemset(d, 0, sizeof(Res)));
        k_mix<<<1, 128, SM>>>(c.af, c.bf, c.ab, c.bb, d);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("  %s: launch failed: %s\n", c.name, cudaGetErrorString(e)); return 0; }
        CK(cudaMemcpy(&h, d, sizeof(Res), cudaMemcpyDeviceToHost));
        printf("  %s -> %.3f (expect %.1f)\n", c.name, h.val, c.expect);
    }
    return 0;
}
