// layout_probe.cu -- correctness probes for the two layouts the fused (chained) MLP kernel relies on:
//   (1) an MN-major, SWIZZLE_128B B operand for tcgen05.mma kind::f16 (A = weights in tensor memory):
//       which descriptor field carries the stride between 64-element N groups / between 8-row K groups;
//   (2) cp.async.bulk.tensor.2d ... tile::gather4: where the four gathered rows land in shared memory
//       under SWIZZLE_128B, and which box shape the tensor map wants;
//   (3) cycles per MMA with the MN-major B operand.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I papc_b200/csrc -o layout_probe tools/microbench/layout_probe.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace papc::umma;

__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// D fp32, A/B f16, A K-major (tensor memory), B MN-major (bit 16), M = 128, N = n
__host__ __device__ constexpr uint32_t idesc_mn(int n) {
    return (1u << 4) | (1u << 16) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
// MN-major SWIZZLE_128B descriptor: lbo / sbo in bytes
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

constexpr int M = 128, N = 128, K = 32;
constexpr uint32_t kNG = 1024;   // bytes between the two 64-row N groups of one 8-row K group
constexpr uint32_t kKG = 2048;   // bytes between consecutive 8-row K groups

__host__ __device__ inline int a_val(int c, int k) { return ((c * 7 + k * 3) % 13) - 6; }
__host__ __device__ inline int b_val(int k, int n) { return ((k * 5 + n) % 11) - 5; }

// variant 0: lbo field = N-group stride, sbo field = K-group stride (CUTLASS canonical reading)
// variant 1: fields swapped
__global__ void __launch_bounds__(128, 1) mn_probe(int variant, float *out, long long *cycles) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // B[k][n] -> MN-major SW128: atom = 8 k-rows of 128 B (64 n); chunk index xor (k & 7)
    for (int e = tid; e < K * N; e += 128) {
        const int k = e / N, n = e % N;
        const uint32_t off = (k >> 3) * kKG + (n >> 6) * kNG + (k & 7) * 128 + ((((n & 63) >> 3) ^ (k & 7)) << 4) + (n & 7) * 2;
        *reinterpret_cast<__half *>(smem + off) = __float2half((float)b_val(k, n));
    }
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<256>(&slot);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = slot;
    // A[c][k] into tensor memory: lane = c, column j = (k = 2j, 2j+1) packed f16x2, columns 128..143
    {
        uint32_t r[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) {
            const __half2 h = __floats2half2_rn(j < K / 2 ? (float)a_val(tid, 2 * j) : 0.f, j < K / 2 ? (float)a_val(tid, 2 * j + 1) : 0.f);
            r[j] = *reinterpret_cast<const uint32_t *>(&h);
        }
        tmem_st32(tb + ((uint32_t)(warp * 32) << 16) + 128, r);
        tmem_wait_st();
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (tid == 0) {
        const uint32_t sb = smem_u32(smem);
        const uint32_t lbo = variant == 0 ? kNG : kKG, sbo = variant == 0 ? kKG : kNG;
        for (int ks = 0; ks < K / 16; ++ks)
            mma_f16_ts(tb, tb + 128 + ks * 8, make_desc_mn(sb + ks * 2 * kKG, lbo, sbo), idesc_mn(N), ks > 0);
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        // timing: 512 more MMAs into columns 0..127 of a scratch accumulator (results unused)
        long long t0 = clock64();
        for (int r = 0; r < 512; ++r)
            mma_f16_ts(tb + 0, tb + 128 + (r & 1) * 8, make_desc_mn(sb + (r & 1) * 2 * kKG, lbo, sbo), idesc_mn(N), 1);
        mma_commit(&bar);
        mbar_wait(&bar, 1);
        long long t1 = clock64();
        if (cycles) cycles[0] = t1 - t0;
    }
    // NOTE: the timing loop above accumulates into the same columns, so read the result BEFORE it:
    // (re-run the two real MMAs with accumulate = 0 after the timing loop)
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (tid == 0) {
        const uint32_t sb = smem_u32(smem);
        const uint32_t lbo = variant == 0 ? kNG : kKG, sbo = variant == 0 ? kKG : kNG;
        for (int ks = 0; ks < K / 16; ++ks)
            mma_f16_ts(tb, tb + 128 + ks * 8, make_desc_mn(sb + ks * 2 * kKG, lbo, sbo), idesc_mn(N), ks > 0);
        mma_commit(&bar);
        mbar_wait(&bar, 0);
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    for (int blk = 0; blk < 4; ++blk) {
        uint32_t r[32];
        tmem_ld32_nowait(tb + ((uint32_t)(warp * 32) << 16) + blk * 32, r);
        tmem_wait_ld();
#pragma unroll
        for (int i = 0; i < 32; ++i) out[(size_t)tid * N + blk * 32 + i] = __uint_as_float(r[i]);
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (warp == 0) tmem_dealloc<256>(tb);
}

// gather4: rows r0..r3 of a [R][C] f16 matrix, 64 columns starting at col0, SWIZZLE_128B
__global__ void __launch_bounds__(32, 1) gather4_probe(const __grid_constant__ CUtensorMap map, int col0, int r0, int r1, int r2,
                                                       int r3, uint32_t dst_off, uint8_t *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    for (int i = tid; i < 4096; i += 32) smem[i] = 0xEE;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    fence_proxy_async();
    __syncthreads();
    if (tid == 0) {
        mbar_expect_tx(&bar, 4 * 128);
        asm volatile(
            "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
            ::"r"(smem_u32(smem + dst_off)), "l"(reinterpret_cast<uint64_t>(&map)), "r"(col0), "r"(r0), "r"(r1), "r"(r2), "r"(r3),
            "r"(smem_u32(&bar))
            : "memory");
    }
    mbar_wait(&bar, 0);
    __syncthreads();
    for (int i = tid; i < 4096; i += 32) out[i] = smem[i];
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                             const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
    // ---------------- (1) MN-major B operand
    float *d_out;
    long long *d_cyc;
    cudaMalloc(&d_out, sizeof(float) * M * N);
    cudaMalloc(&d_cyc, 8);
    cudaFuncSetAttribute(mn_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 40000);
    std::vector<float> h(M * N);
    for (int variant = 0; variant < 2; ++variant) {
        cudaMemset(d_out, 0, sizeof(float) * M * N);
        mn_probe<<<1, 128, 40000>>>(variant, d_out, d_cyc);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h.data(), d_out, sizeof(float) * M * N, cudaMemcpyDeviceToHost);
        long long cyc = 0;
        cudaMemcpy(&cyc, d_cyc, 8, cudaMemcpyDeviceToHost);
        int bad = 0;
        double maxerr = 0;
        for (int c = 0; c < M; ++c)
            for (int n = 0; n < N; ++n) {
                double ref = 0;
                for (int k = 0; k < K; ++k) ref += (double)a_val(c, k) * b_val(k, n);
                const double err = fabs(ref - h[c * N + n]);
                if (err > 1e-3) ++bad;
                if (err > maxerr) maxerr = err;
            }
        printf("[mn-major] variant %d (%s): %d / %d wrong, max err %.3g, %.1f cycles/MMA (%s)\n", variant,
               variant == 0 ? "lbo = N-group stride, sbo = K-group stride" : "swapped", bad, M * N, maxerr, (double)cyc / 512,
               cudaGetErrorString(e));
        if (e != cudaSuccess) { printf("sticky error, stopping\n"); return 1; }
    }
    // ---------------- (2) gather4
    const int R = 512, C = 144;   // pitch 288 B: columns 128..143 valid, 144..191 out of bounds (zero fill)
    std::vector<__half> src((size_t)R * C);
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < C; ++c) src[(size_t)r * C + c] = __float2half((float)((r % 32) * 8 + ((c / 8) % 8)));
    // value encodes the row (mod 32) and the 16-byte source chunk (c / 8) mod 8
    __half *d_src;
    uint8_t *d_dump;
    cudaMalloc(&d_src, src.size() * 2);
    cudaMalloc(&d_dump, 4096);
    cudaMemcpy(d_src, src.data(), src.size() * 2, cudaMemcpyHostToDevice);
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        printf("[gather4] no cuTensorMapEncodeTiled\n");
        return 0;
    }
    EncodeFn encode = (EncodeFn)fn;
    cudaFuncSetAttribute(gather4_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 8192);
    for (int boxrows = 1; boxrows <= 4; boxrows += 3) {
        CUtensorMap map;
        const cuuint64_t gdim[2] = {(cuuint64_t)C, (cuuint64_t)R};
        const cuuint64_t gstride[1] = {(cuuint64_t)C * 2};
        const cuuint32_t box[2] = {64, (cuuint32_t)boxrows};
        const cuuint32_t estr[2] = {1, 1};
        CUresult cr = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, d_src, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        printf("[gather4] box {64,%d}: encode -> %d\n", boxrows, (int)cr);
        if (cr != CUDA_SUCCESS) continue;
        for (int t = 0; t < 3; ++t) {
            const int col0 = t == 2 ? 128 : t * 64;           // last: 16 valid columns, the rest out of bounds
            const uint32_t dst_off = t == 1 ? 512 : 0;         // second: rows 4..7 of the swizzle atom
            const int rows[4] = {5, 300, 17, 129};
            cudaMemset(d_dump, 0, 4096);
            gather4_probe<<<1, 32, 8192>>>(map, col0, rows[0], rows[1], rows[2], rows[3], dst_off, d_dump);
            cudaError_t e = cudaDeviceSynchronize();
            printf("[gather4]  col0 %d dst +%u: %s\n", col0, dst_off, cudaGetErrorString(e));
            if (e != cudaSuccess) { printf("sticky error, stopping\n"); return 1; }
            std::vector<uint8_t> dump(4096);
            cudaMemcpy(dump.data(), d_dump, 4096, cudaMemcpyDeviceToHost);
            // for each 16-byte chunk of the first 2 KB: which (row, source chunk) does it hold?
            for (int line = 0; line < 16; ++line) {
                printf("[gather4]   smem line %2d:", line);
                for (int ch = 0; ch < 8; ++ch) {
                    const __half *p = reinterpret_cast<const __half *>(dump.data() + line * 128 + ch * 16);
                    const uint8_t *b = dump.data() + line * 128 + ch * 16;
                    if (b[0] == 0xEE && b[1] == 0xEE) { printf("  ----  "); continue; }
                    const int v0 = (int)__half2float(p[0]), v7 = (int)__half2float(p[7]);
                    if (v0 == v7) printf(" r%02dc%d ", v0 / 8, v0 % 8);
                    else printf(" ?%3d/%3d", v0, v7);
                }
                printf("\n");
            }
        }
    }
    return 0;
}
