// common.cuh -- shared helpers for the papc_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "papc_b200.h"

namespace papc {

extern thread_local int g_last_cuda_error;
extern thread_local unsigned long long g_launch_count;  // kernels launched by this host thread

inline int cuda_fail(cudaError_t e) {
    g_last_cuda_error = (int)e;
    return PAPC_ECUDA;
}

#define PAPC_CUDA_TRY(expr)                                   \
    do {                                                      \
        cudaError_t _e = (expr);                              \
        if (_e != cudaSuccess) return ::papc::cuda_fail(_e);  \
    } while (0)

// Launch errors (bad configuration etc.) are synchronous; cudaPeekAtLastError does not
// synchronise the stream.
#define PAPC_LAUNCH_CHECK()                     \
    do {                                        \
        ++::papc::g_launch_count;               \
        PAPC_CUDA_TRY(cudaPeekAtLastError());   \
    } while (0)

inline cudaStream_t as_stream(papc_stream_t s) { return reinterpret_cast<cudaStream_t>(s); }

// ---- launch profiler (papc_prof_* in the C ABI): when enabled on the calling host thread, the
// instrumented kernel launches are bracketed by CUDA events recorded on the launching stream.
struct ProfRec {
    char name[56];
    long long M;
    int cin, cout;
    double flops, bytes;  // algorithmic work of the launch (0 = not stated)
    cudaEvent_t e0, e1;
};
struct Prof {
    bool on = false;
    int n = 0, cap = 0;
    ProfRec *rec = nullptr;
};
extern thread_local Prof g_prof;
struct ProfScope {
    ProfRec *r = nullptr;
    cudaStream_t st;
    ProfScope(cudaStream_t stream, const char *name, long long M, int cin, int cout, double flops,
              double bytes);
    ~ProfScope();
    ProfScope(const ProfScope &) = delete;
    ProfScope &operator=(const ProfScope &) = delete;
};

template <typename T>
__host__ __device__ constexpr T ceil_div(T a, T b) {
    return (a + b - 1) / b;
}

inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// ---- pinned fp32 arithmetic (never contracted into FMAs by the compiler) ----------------
// |p|^2 as numpy/paddle sum(p**2, -1) evaluates it: (x*x + y*y) + z*z.
__device__ __forceinline__ float sq3(float x, float y, float z) {
    return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z));
}
// square_distance element, layers.py:36-38: (-2*dot + |q|^2) + |p|^2, dot = BLAS-style FMA chain.
__device__ __forceinline__ float sqdist_expanded(float qx, float qy, float qz, float qn,
                                                 float px, float py, float pz, float pn) {
    float dot = __fmul_rn(qx, px);
    dot = __fmaf_rn(qy, py, dot);
    dot = __fmaf_rn(qz, pz, dot);
    return __fadd_rn(__fadd_rn(__fmul_rn(-2.0f, dot), qn), pn);
}

// ---- float atomic max / min through the integer atomics ---------------------------------
__device__ __forceinline__ void atomic_max_f32(float *addr, float v) {
    v += 0.0f;  // -0 -> +0 so the sign test below matches the bit pattern
    if (v >= 0.0f)
        atomicMax(reinterpret_cast<int *>(addr), __float_as_int(v));
    else
        atomicMin(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_min_f32(float *addr, float v) {
    v += 0.0f;
    if (v >= 0.0f)
        atomicMin(reinterpret_cast<int *>(addr), __float_as_int(v));
    else
        atomicMax(reinterpret_cast<unsigned int *>(addr), __float_as_uint(v));
}

}  // namespace papc
