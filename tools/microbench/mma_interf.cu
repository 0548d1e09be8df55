// mma_interf.cu -- does other warps' work slow tcgen05.mma down?  One thread issues f16 TS-mode
// M128 N128 K16 MMAs (alternating two accumulators); the other warps run an interference loop:
//   0 none | 1 LDS.128 reads | 2 STS.128 writes | 4 tcgen05.ld of the idle accumulator | 8 FFMA ALU
//   16 cp.async (LDGSTS) global->shared | 32 fence.proxy.async after STS
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I papc_b200/csrc -o mma_interf tools/microbench/mma_interf.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace papc::umma;

__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_of(int fmt, int n) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__global__ void __launch_bounds__(544, 1) k(int reps, int mode, const float *g, long long *out, float *sink) {
    long long ld_cyc = 0, ld_n = 0;
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int done;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < 196608 / 4; i += 544) ((uint32_t *)smem)[i] = 0;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); done = 0; }
    if (warp == 16) tmem_alloc<512>(&slot);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = slot;
    if (warp == 16) {
        if (lane == 0) {
            constexpr uint32_t idesc = idesc_of(0, 128);
            const uint32_t sb = smem_u32(smem);
            long long t0 = clock64();
            for (int r = 0; r < reps; ++r) {
                const uint32_t d = tb + ((r / 12) & 1) * 128;
                mma_f16_ts(d, tb + 256 + (r & 15) * 8, make_desc_sw128(sb + (r & 3) * 32 + ((r >> 2) & 1) * 16384), idesc, 1);
            }
            mma_commit(&bar);
            mbar_wait(&bar, 0);
            long long t1 = clock64();
            if (blockIdx.x == 0) out[0] = t1 - t0;
            done = 1;
        }
    } else if (mode != 0) {
        // warps 0-7 "epilogue", 8-15 "producers"
        const uint32_t sm = smem_u32(smem);
        float acc = 0.f;
        uint32_t it = 0;
        while (!done) {
            if ((mode & 1) && warp >= 8) {
#pragma unroll
                for (int q = 0; q < 8; ++q) { float4 v = lds128f(sm + 65536 + ((q * 256 + (tid - 256)) & 4095) * 16); acc += v.x; }
            }
            if ((mode & 2) && warp >= 8) {
#pragma unroll
                for (int q = 0; q < 8; ++q) sts128f(sm + 32768 + ((q * 256 + (tid - 256)) & 2047) * 16, make_float4(acc, 1.f, 2.f, 3.f));
            }
            if ((mode & 32) && warp >= 8) fence_proxy_async();
            if ((mode & 16) && warp >= 8) {
#pragma unroll
                for (int q = 0; q < 8; ++q)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sm + 131072 + ((q * 256 + (tid - 256)) & 4095) * 16),
                                 "l"(g + ((size_t)blockIdx.x * 65536 + ((it * 2048 + q * 256 + (tid - 256)) & 65535)) * 4) : "memory");
                asm volatile("cp.async.commit_group;" ::: "memory");
                asm volatile("cp.async.wait_group 2;" ::: "memory");
            }
            if ((mode & 4) && warp < 8) {
                uint32_t r[32];
                const long long c0 = clock64();
                tmem_ld32_nowait(tb + ((uint32_t)((warp & 3) * 32) << 16) + 384 + (it & 1) * 32, r);
                tmem_wait_ld();
                acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
                ld_cyc += clock64() - c0;
                ++ld_n;
                for (int q = 0; q < 100; ++q) acc = fmaf(acc, 1.0001f, 0.5f);  // ~0.2 us between loads
            }
            if ((mode & 8)) {
#pragma unroll
                for (int q = 0; q < 64; ++q) acc = fmaf(acc, 1.0001f, 0.5f);
            }
            ++it;
        }
        if (acc == 12345.f) sink[tid] = acc;
        if (blockIdx.x == 0 && tid == 0 && ld_n > 0) { out[1] = ld_cyc; out[2] = ld_n; }
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 16) tmem_dealloc<512>(tb);
}

int main() {
    long long *d_out; float *sink, *g;
    cudaMalloc(&d_out, 24); cudaMemset(d_out, 0, 24); cudaMalloc(&sink, 4096); cudaMalloc(&g, (size_t)148 * 65536 * 16);
    cudaMemset(g, 0, (size_t)148 * 65536 * 16);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200000);
    const int modes[] = {4, 4 + 1, 4 + 2, 4 + 2 + 32, 4 + 16, 4 + 1 + 2 + 16, 4 + 1 + 2 + 16 + 32, 4 + 8};
    for (int m : modes) {
        const int reps = 12 * 512;
        k<<<148, 544, 200000>>>(reps, m, g, d_out, sink);
        k<<<148, 544, 200000>>>(reps, m, g, d_out, sink);
        cudaError_t e = cudaDeviceSynchronize();
        long long h[3] = {0, 0, 0};
        cudaMemcpy(h, d_out, 24, cudaMemcpyDeviceToHost);
        printf("mode %2d (%s%s%s%s%s%s): %8.1f cycles/MMA   tmem ld+wait %7.1f cycles (%lld)  (%s)\n", m, m & 1 ? "LDS " : "",
               m & 2 ? "STS " : "", m & 32 ? "fence " : "", m & 16 ? "cp.async " : "", m & 4 ? "tmem.ld " : "", m & 8 ? "FFMA " : "",
               (double)h[0] / reps, h[2] ? (double)h[1] / h[2] : 0.0, h[2], cudaGetErrorString(e));
        cudaMemset(d_out, 0, 24);
    }
    return 0;
}
