// mma_rate.cu -- cycles per tcgen05.mma (cta_group::1, M=128) for kind::tf32 / kind::f16, A from shared or
// tensor memory, N = 64/128/256, same or alternating accumulator.  One CTA per SM, one issuing thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I papc_b200/csrc -o mma_rate tools/microbench/mma_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "umma.cuh"
using namespace papc::umma;

__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
// idesc: D fp32 (1<<4); A/B format at bits 7 / 10: tf32 = 2, f16 = 0, bf16 = 1
__host__ __device__ constexpr uint32_t idesc_of(int fmt, int n) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

template <int KIND /*0 tf32, 1 f16*/, int ASRC /*0 smem, 1 tmem*/, int N, int NACC>
__global__ void __launch_bounds__(128, 1) k(int reps, long long *out) {
    extern __shared__ uint8_t raw[];
    uint8_t *smem = (uint8_t *)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    for (int i = threadIdx.x; i < 65536 / 4; i += 128) ((uint32_t *)smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    if (threadIdx.x < 32) tmem_alloc<512>(&slot);
    fence_proxy_async();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = slot;
    if (threadIdx.x == 0) {
        constexpr uint32_t idesc = idesc_of(KIND == 0 ? 2 : 0, N);
        const uint32_t sa = smem_u32(smem), sb = smem_u32(smem + 32768);
        long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = tb + (NACC > 1 ? (r % NACC) * N : 0);
            const uint64_t db = make_desc_sw128(sb + (r & 3) * 32);
            if (KIND == 0) {
                if (ASRC == 0) mma_tf32_ss(d, make_desc_sw128(sa + (r & 3) * 32), db, idesc, 1);
                else mma_tf32_ts(d, tb + 256 + (r & 15) * 8, db, idesc, 1);
            } else {
                if (ASRC == 0) mma_f16_ss(d, make_desc_sw128(sa + (r & 3) * 32), db, idesc, 1);
                else mma_f16_ts(d, tb + 256 + (r & 15) * 8, db, idesc, 1);
            }
        }
        mma_commit(&bar);
        mbar_wait(&bar, 0);
        long long t1 = clock64();
        if (blockIdx.x == 0) out[0] = t1 - t0;
    }
    tc_fence_before(); __syncthreads(); tc_fence_after();
    if (threadIdx.x < 32) tmem_dealloc<512>(tb);
}

template <int KIND, int ASRC, int N, int NACC>
void run(const char *name, long long *d_out) {
    auto kern = k<KIND, ASRC, N, NACC>;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 70000);
    const int reps = 4096;
    kern<<<148, 128, 70000>>>(reps, d_out);
    kern<<<148, 128, 70000>>>(reps, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    long long h = 0;
    cudaMemcpy(&h, d_out, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %8.1f cycles/MMA  (%s)\n", name, (double)h / reps, cudaGetErrorString(e));
}

int main() {
    long long *d_out;
    cudaMalloc(&d_out, 8);
    run<0, 0, 128, 1>("tf32 SS  M128 N128 K8  same D", d_out);
    run<0, 1, 128, 1>("tf32 TS  M128 N128 K8  same D", d_out);
    run<0, 1, 128, 2>("tf32 TS  M128 N128 K8  2 accumulators", d_out);
    run<0, 1, 256, 1>("tf32 TS  M128 N256 K8  same D", d_out);
    run<0, 1, 64, 1>("tf32 TS  M128 N64  K8  same D", d_out);
    run<0, 0, 256, 1>("tf32 SS  M128 N256 K8  same D", d_out);
    run<1, 0, 128, 1>("f16  SS  M128 N128 K16 same D", d_out);
    run<1, 1, 128, 1>("f16  TS  M128 N128 K16 same D", d_out);
    run<1, 1, 256, 1>("f16  TS  M128 N256 K16 same D", d_out);
    run<1, 0, 256, 1>("f16  SS  M128 N256 K16 same D", d_out);
    run<1, 1, 64, 1>("f16  TS  M128 N64  K16 same D", d_out);
    return 0;
}
