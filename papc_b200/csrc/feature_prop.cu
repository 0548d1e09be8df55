// feature_prop.cu -- PointNetFeaturePropagation interpolation (reference: layers.py:306-329, SURVEY.md
// 8f row N1) for sm_100a.
//
// For every point of the dense cloud: the three smallest squared distances to the S sampled points
// (square_distance's expansion arithmetic, pinned as in ball_query.cu), inverse-distance weights,
// and the weighted sum of three rows of points2 -- fused with the concat [points1 | interpolated]
// into one channels-last row, which is what the pointwise MLP that follows reads.
//
// Reference quirk reproduced on purpose: layers.py sorts `dists` (:317) and then takes the argsort
// of the SORTED array (:318), i.e. the identity -- the weights come from the three nearest sampled
// points but multiply the features of sampled points 0, 1, 2.  The [B,N,S] matrix and both sorts
// are never materialised: one warp per point, lanes split the S candidates (register top-3 each),
// a three-round warp merge, then the lanes split the channels (coalesced 128-byte row writes).
#include "common.cuh"

namespace papc {

constexpr int kFpWarps = 8;

__device__ __forceinline__ void top3_insert(float d, float &a0, float &a1, float &a2) {
    if (d < a2) {
        if (d < a1) {
            a2 = a1;
            if (d < a0) { a1 = a0; a0 = d; } else { a1 = d; }
        } else {
            a2 = d;
        }
    }
}

__global__ void __launch_bounds__(kFpWarps * 32)
fp_interp_kernel(const float *__restrict__ xyz1, const float *__restrict__ xyz2,
                 const float *__restrict__ points1, const float *__restrict__ points2, int N, int S,
                 int D1, int D2, int ld, float *__restrict__ out) {
    extern __shared__ float4 s_q[];  // [S] sampled points (x, y, z, |p|^2)
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *c2 = xyz2 + (size_t)b * S * 3;
    for (int j = tid; j < S; j += kFpWarps * 32) {
        const float x = c2[j * 3 + 0], y = c2[j * 3 + 1], z = c2[j * 3 + 2];
        s_q[j] = make_float4(x, y, z, sq3(x, y, z));
    }
    __syncthreads();
    const float *p2 = points2 + (size_t)b * S * D2;
    const int K3 = S < 3 ? S : 3;
    for (int n = blockIdx.x * kFpWarps + warp; n < N; n += gridDim.x * kFpWarps) {
        float w0 = 1.f, w1 = 0.f, w2 = 0.f;
        if (S > 1) {
            const float *q = xyz1 + ((size_t)b * N + n) * 3;
            const float qx = q[0], qy = q[1], qz = q[2];
            const float qn = sq3(qx, qy, qz);
            float a0 = INFINITY, a1 = INFINITY, a2 = INFINITY;
            for (int j = lane; j < S; j += 32) {
                const float4 p = s_q[j];
                top3_insert(sqdist_expanded(qx, qy, qz, qn, p.x, p.y, p.z, p.w), a0, a1, a2);
            }
            // merge the 32 sorted triples: three rounds of warp-min + pop at the lowest holder
            float m[3];
#pragma unroll
            for (int r = 0; r < 3; ++r) {
                float v = a0;
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, o));
                m[r] = v;
                const unsigned holders = __ballot_sync(0xffffffffu, a0 == v);
                if (lane == __ffs(holders) - 1) { a0 = a1; a1 = a2; a2 = INFINITY; }
            }
            // layers.py:321-323 (fp32, separately rounded)
            const float r0 = __fdiv_rn(1.0f, __fadd_rn(m[0], 1e-8f));
            const float r1 = K3 > 1 ? __fdiv_rn(1.0f, __fadd_rn(m[1], 1e-8f)) : 0.f;
            const float r2 = K3 > 2 ? __fdiv_rn(1.0f, __fadd_rn(m[2], 1e-8f)) : 0.f;
            float norm = r0;
            if (K3 > 1) norm = __fadd_rn(norm, r1);
            if (K3 > 2) norm = __fadd_rn(norm, r2);
            w0 = __fdiv_rn(r0, norm);
            w1 = __fdiv_rn(r1, norm);
            w2 = __fdiv_rn(r2, norm);
        }
        float *o = out + ((size_t)b * N + n) * ld;
        if (points1 != nullptr) {
            const float *p1 = points1 + ((size_t)b * N + n) * D1;
            for (int c = lane; c < D1; c += 32) o[c] = p1[c];
        }
        for (int c = lane; c < D2; c += 32) {
            float v;
            if (S == 1) {
                v = p2[c];                                             // :314 tile
            } else {
                v = __fmul_rn(p2[c], w0);                              // :324, rows 0, 1, 2 (the quirk)
                if (K3 > 1) v = __fadd_rn(v, __fmul_rn(p2[D2 + c], w1));
                if (K3 > 2) v = __fadd_rn(v, __fmul_rn(p2[2 * D2 + c], w2));
            }
            o[D1 + c] = v;
        }
        for (int c = D1 + D2 + lane; c < ld; c += 32) o[c] = 0.f;     // row padding
    }
}

}  // namespace papc

using namespace papc;

extern "C" int papc_fp_interpolate_f32(const float *xyz1, const float *xyz2, const float *points1,
                                       const float *points2, int B, int N, int S, int D1, int D2,
                                       int ld_out, float *out, papc_stream_t stream) {
    if (B < 0 || N < 0 || S <= 0 || D1 < 0 || D2 <= 0 || ld_out < D1 + D2) return PAPC_EINVAL;
    if (B == 0 || N == 0) return PAPC_OK;
    if (!xyz1 || !xyz2 || !points2 || !out || (D1 > 0 && !points1)) return PAPC_EINVAL;
    if (B > 65535) return PAPC_EUNSUPPORTED;
    const size_t smem = (size_t)S * sizeof(float4);
    if (smem > 200 * 1024) return PAPC_EUNSUPPORTED;
    if (smem > 48 * 1024)
        PAPC_CUDA_TRY(cudaFuncSetAttribute(fp_interp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int gx = ceil_div(N, kFpWarps);
    const int cap = ceil_div(kNumSMs * 8, B > 0 ? B : 1);
    if (gx > cap) gx = cap > 0 ? cap : 1;
    cudaStream_t st = as_stream(stream);
    ProfScope prof(st, "fp_interpolate", (long long)B * N, S, D1 + D2, 0.0,
                   12.0 * B * (N + S) + 4.0 * B * S * D2 + 4.0 * B * N * (D1 + (double)ld_out));
    fp_interp_kernel<<<dim3(gx, B), kFpWarps * 32, smem, st>>>(xyz1, xyz2, D1 > 0 ? points1 : nullptr, points2, N, S,
                                                               D1, D2, ld_out, out);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}
