// fps.cu -- farthest point sampling (reference: layers.py:65-95) for sm_100a.
//
// One persistent CTA per cloud runs all `npoint` dependent iterations: the cloud's xyz and the
// running min-distance live in REGISTERS (PPT points per thread, blocked so that lower threads
// own lower indices), a shared-memory AoS copy of xyz serves the winner-coordinate broadcast,
// and the per-iteration argmax is a redux.sync (integer max on the fp32 bit pattern -- all
// distances are >= 0) + ballot inside each warp, then one double-buffered shared-memory
// exchange across warps: ONE __syncthreads per iteration.
//
// Bit-exactness vs the oracle: d = (dx*dx + dy*dy) + dz*dz with separately rounded ops
// (__fmul_rn/__fadd_rn are never contracted), running = min(running, d) (== the reference's
// masked assignment for non-NaN data), argmax = lowest index among maxima (paddle.argmax).
#include "common.cuh"

namespace papc {

constexpr int kFpsRegMaxN = 8192;

template <int THREADS, int PPT>
__global__ void __launch_bounds__(THREADS)
fps_reg_kernel(const float *__restrict__ xyz, int N, int npoint,
               const int64_t *__restrict__ start_idx, float init_dist,
               int64_t *__restrict__ out_idx, float *__restrict__ out_new_xyz) {
    extern __shared__ float s_xyz[];  // [N*3] AoS
    constexpr int NW = THREADS / 32;
    __shared__ unsigned long long s_red[2][32];

    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const float *cloud = xyz + (size_t)b * N * 3;

    for (int i = tid; i < N * 3; i += THREADS) s_xyz[i] = cloud[i];
    __syncthreads();

    float px[PPT], py[PPT], pz[PPT], pd[PPT];
#pragma unroll
    for (int p = 0; p < PPT; ++p) {
        const int j = tid * PPT + p;
        if (j < N) {
            px[p] = s_xyz[j * 3 + 0];
            py[p] = s_xyz[j * 3 + 1];
            pz[p] = s_xyz[j * 3 + 2];
            pd[p] = init_dist;
        } else {
            px[p] = py[p] = pz[p] = 0.0f;
            pd[p] = -1.0f;  // negative bit pattern: never the (signed) maximum
        }
    }

    int far = (int)start_idx[b];
    far = min(max(far, 0), N - 1);
    int64_t *out = out_idx + (size_t)b * npoint;

    for (int it = 0; it < npoint; ++it) {
        const float cx = s_xyz[far * 3 + 0];
        const float cy = s_xyz[far * 3 + 1];
        const float cz = s_xyz[far * 3 + 2];
        if (tid == 0) {
            out[it] = far;
            if (out_new_xyz != nullptr) {
                float *o = out_new_xyz + ((size_t)b * npoint + it) * 3;
                o[0] = cx;
                o[1] = cy;
                o[2] = cz;
            }
        }
        int best = -2147483647 - 1;
        int besti = 0;
#pragma unroll
        for (int p = 0; p < PPT; ++p) {
            const float dx = __fsub_rn(px[p], cx);
            const float dy = __fsub_rn(py[p], cy);
            const float dz = __fsub_rn(pz[p], cz);
            const float d = sq3(dx, dy, dz);
            // tid*PPT+p >= N keeps pd = -1 (d >= 0 is never smaller)
            pd[p] = (d < pd[p]) ? d : pd[p];
            const int bits = __float_as_int(pd[p]);
            if (bits > best) {  // strict: first maximum wins inside the thread
                best = bits;
                besti = tid * PPT + p;
            }
        }
        // warp argmax: integer max, then the lowest lane holding it (lanes own ascending indices)
        const int wmax = __reduce_max_sync(0xffffffffu, best);
        const unsigned ball = __ballot_sync(0xffffffffu, best == wmax);
        const int src = __ffs(ball) - 1;
        const int widx = __shfl_sync(0xffffffffu, besti, src);
        if (NW == 1) {
            far = widx;
        } else {
            if (lane == 0)
                s_red[it & 1][warp] =
                    ((unsigned long long)(unsigned)wmax << 32) | (unsigned)widx;
            __syncthreads();
            unsigned long long k = s_red[it & 1][lane < NW ? lane : 0];
            const int kmax = (int)(unsigned)(k >> 32);
            const int m2 = __reduce_max_sync(0xffffffffu, kmax);
            const unsigned ball2 = __ballot_sync(0xffffffffu, (kmax == m2) && (lane < NW));
            const int src2 = __ffs(ball2) - 1;  // lowest warp == lowest index
            far = (int)__shfl_sync(0xffffffffu, (unsigned)(k & 0xffffffffu), src2);
        }
    }
}

// Generic path for N > 8192: running distances in global memory (L2-resident), xyz read through
// L1/L2 every iteration.  Correct for any N; not the tuned path.
template <int THREADS>
__global__ void __launch_bounds__(THREADS)
fps_global_kernel(const float *__restrict__ xyz, int N, int npoint,
                  const int64_t *__restrict__ start_idx, float init_dist,
                  int64_t *__restrict__ out_idx, float *__restrict__ out_new_xyz,
                  float *__restrict__ dist_ws) {
    constexpr int NW = THREADS / 32;
    __shared__ unsigned long long s_red[2][32];
    const int b = blockIdx.x;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const float *cloud = xyz + (size_t)b * N * 3;
    float *dist = dist_ws + (size_t)b * N;
    // blocked ownership: thread t owns [t*chunk, (t+1)*chunk)
    const int chunk = (N + THREADS - 1) / THREADS;
    const int j0 = tid * chunk;
    const int j1 = min(N, j0 + chunk);
    for (int j = j0; j < j1; ++j) dist[j] = init_dist;

    int far = (int)start_idx[b];
    far = min(max(far, 0), N - 1);
    int64_t *out = out_idx + (size_t)b * npoint;
    for (int it = 0; it < npoint; ++it) {
        const float cx = cloud[far * 3 + 0];
        const float cy = cloud[far * 3 + 1];
        const float cz = cloud[far * 3 + 2];
        if (tid == 0) {
            out[it] = far;
            if (out_new_xyz != nullptr) {
                float *o = out_new_xyz + ((size_t)b * npoint + it) * 3;
                o[0] = cx;
                o[1] = cy;
                o[2] = cz;
            }
        }
        int best = -2147483647 - 1;
        int besti = 0;
        for (int j = j0; j < j1; ++j) {
            const float dx = __fsub_rn(cloud[j * 3 + 0], cx);
            const float dy = __fsub_rn(cloud[j * 3 + 1], cy);
            const float dz = __fsub_rn(cloud[j * 3 + 2], cz);
            const float d = sq3(dx, dy, dz);
            float r = dist[j];
            r = (d < r) ? d : r;
            dist[j] = r;
            const int bits = __float_as_int(r);
            if (bits > best) {
                best = bits;
                besti = j;
            }
        }
        const int wmax = __reduce_max_sync(0xffffffffu, best);
        const unsigned ball = __ballot_sync(0xffffffffu, best == wmax);
        const int src = __ffs(ball) - 1;
        const int widx = __shfl_sync(0xffffffffu, besti, src);
        if (lane == 0)
            s_red[it & 1][warp] = ((unsigned long long)(unsigned)wmax << 32) | (unsigned)widx;
        __syncthreads();
        unsigned long long k = s_red[it & 1][lane < NW ? lane : 0];
        const int kmax = (int)(unsigned)(k >> 32);
        const int m2 = __reduce_max_sync(0xffffffffu, kmax);
        const unsigned ball2 = __ballot_sync(0xffffffffu, (kmax == m2) && (lane < NW));
        const int src2 = __ffs(ball2) - 1;
        far = (int)__shfl_sync(0xffffffffu, (unsigned)(k & 0xffffffffu), src2);
    }
}

template <int THREADS, int PPT>
static int launch_fps_reg(const float *xyz, int B, int N, int npoint, const int64_t *start,
                          float init_dist, int64_t *out_idx, float *out_new_xyz,
                          cudaStream_t st) {
    const size_t smem = (size_t)N * 3 * sizeof(float);
    auto k = fps_reg_kernel<THREADS, PPT>;
    if (smem > 48 * 1024)
        PAPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
    k<<<B, THREADS, smem, st>>>(xyz, N, npoint, start, init_dist, out_idx, out_new_xyz);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

}  // namespace papc

extern "C" size_t papc_fps_workspace_bytes(int B, int N) {
    if (N <= papc::kFpsRegMaxN || B <= 0) return 0;
    return (size_t)B * (size_t)N * sizeof(float);
}

extern "C" int papc_fps_f32(const float *xyz, int B, int N, int npoint,
                            const int64_t *start_idx, float init_dist, int64_t *out_idx,
                            float *out_new_xyz, void *workspace, size_t workspace_bytes,
                            papc_stream_t stream) {
    using namespace papc;
    if (B < 0 || N <= 0 || npoint < 0 || !(init_dist >= 0.0f)) return PAPC_EINVAL;
    if (B == 0 || npoint == 0) return PAPC_OK;
    if (!xyz || !start_idx || !out_idx) return PAPC_EINVAL;
    cudaStream_t st = as_stream(stream);
    if (N <= 32) return launch_fps_reg<32, 1>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_new_xyz, st);
    if (N <= 64) return launch_fps_reg<64, 1>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_new_xyz, st);
    if (N <= 128) return launch_fps_reg<128, 1>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_new_xyz, st);
    if (N <= 256) return launch_fps_reg<256, 1>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_new_xyz, st);
    if (N <= 512) return launch_fps_reg<256, 2>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_new_xyz, st);
    if (N <= 1024) return launch_fps_reg<256, 4>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_new_xyz, st);
    if (N <= 2048) return launch_fps_reg<512, 4>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_new_xyz, st);
    if (N <= 4096) return launch_fps_reg<1024, 4>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_new_xyz, st);
    if (N <= kFpsRegMaxN) return launch_fps_reg<1024, 8>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_new_xyz, st);
    const size_t need = papc_fps_workspace_bytes(B, N);
    if (!workspace || workspace_bytes < need) return PAPC_EWORKSPACE;
    fps_global_kernel<1024><<<B, 1024, 0, st>>>(xyz, N, npoint, start_idx, init_dist, out_idx,
                                                 out_new_xyz, reinterpret_cast<float *>(workspace));
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}
