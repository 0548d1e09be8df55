// ball_query.cu -- square_distance, index_points, query_ball_point and the group gather
// (reference: layers.py:26-62, 98-126, 146-151, 263-267) for sm_100a.
//
// query_ball_point never materialises the reference's [B,S,N] distance matrix nor sorts it:
// "the nsample lowest in-radius indices, ascending" is an in-order scan with a warp ballot +
// prefix popcount compaction.  One warp per query point (kBqQPW queries in turn);
// all warps of a CTA share one cloud
// whose xyz block is staged into shared memory by a TMA bulk copy (cp.async.bulk, completion
// on an mbarrier) -- chunked, so any N works with a fixed 28 KB of shared memory.
#include "common.cuh"

namespace papc {

// ------------------------------------------------------------------ mbarrier / bulk copy PTX
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t phase) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(phase)
        : "memory");
    return ok != 0;
}
// Bounded wait: a lost transaction traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, phase); ++spin)
        if (spin > (1u << 22)) __trap();
}
// 1-D TMA bulk copy global -> shared; size and both addresses must be multiples of 16 bytes.
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------ square_distance
__global__ void __launch_bounds__(256)
square_distance_kernel(const float *__restrict__ src, const float *__restrict__ dst, int N, int M,
                       float *__restrict__ out, size_t total) {
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; e < total; e += stride) {
        const size_t j = e % M;
        const size_t bi = e / M;  // b*N + i
        const size_t b = bi / N;
        const float *q = src + bi * 3;
        const float *p = dst + (b * M + j) * 3;
        const float qx = q[0], qy = q[1], qz = q[2];
        const float px = p[0], py = p[1], pz = p[2];
        out[e] = sqdist_expanded(qx, qy, qz, sq3(qx, qy, qz), px, py, pz, sq3(px, py, pz));
    }
}

// ------------------------------------------------------------------ index_points
template <typename VecT>
__global__ void __launch_bounds__(256)
gather_rows_kernel(const float *__restrict__ points, const int64_t *__restrict__ idx, int N,
                   int Cv /* row length in VecT units */, int M, size_t total,
                   float *__restrict__ out) {
    const VecT *pin = reinterpret_cast<const VecT *>(points);
    VecT *pout = reinterpret_cast<VecT *>(out);
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; e < total; e += stride) {
        const size_t c = e % Cv;
        const size_t bm = e / Cv;  // b*M + m
        const size_t b = bm / M;
        long long n = idx[bm];
        n = n < 0 ? 0 : (n >= N ? N - 1 : n);
        pout[e] = pin[(b * N + (size_t)n) * Cv + c];
    }
}

// ------------------------------------------------------------------ query_ball_point
constexpr int kBqWarps = 8;
constexpr int kBqQPW = 1;       // queries per warp (4 measured slower: fewer, longer warps; scanning, not staging, is the cost)
constexpr int kBqChunk = 2048;  // points per shared-memory stage: 32 KB (x,y,z,|p|^2) + 24 KB AoS landing zone

template <typename IdxT>
__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz, int N, int S,
                  float radius2, int K, IdxT *__restrict__ out_idx,
                  int32_t *__restrict__ empty_count) {
    // dynamic: chunk = min(N, kBqChunk) points -- [chunk*3] AoS xyz of the current chunk, then
    // [chunk] |p|^2 (small clouds take a few KB, so the kernel co-resides with the MLP kernels)
    extern __shared__ __align__(128) float s_dyn[];
    const int chunk = N < kBqChunk ? ((N + 3) & ~3) : kBqChunk;
    float4 *s_q = reinterpret_cast<float4 *>(s_dyn);  // [chunk] (x, y, z, |p|^2): one LDS.128 per test
    float *s_p = s_dyn + chunk * 4;                    // [chunk*3] AoS landing zone of the bulk copy
    __shared__ __align__(8) uint64_t s_bar;

    const int b = blockIdx.y;
    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const int s0 = (blockIdx.x * kBqWarps + warp) * kBqQPW;  // this warp's queries s0 .. s0+QPW-1
    const float *cloud = xyz + (size_t)b * N * 3;

    float qx[kBqQPW], qy[kBqQPW], qz[kBqQPW], qn[kBqQPW];
    int cnt[kBqQPW];    // in-radius points found so far (may exceed K)
    int first[kBqQPW];  // lowest in-radius index
#pragma unroll
    for (int i = 0; i < kBqQPW; ++i) {
        qx[i] = qy[i] = qz[i] = qn[i] = 0.f;
        cnt[i] = 0;
        first[i] = -1;
        if (s0 + i < S) {
            const float *q = new_xyz + ((size_t)b * S + s0 + i) * 3;
            qx[i] = q[0];
            qy[i] = q[1];
            qz[i] = q[2];
            qn[i] = sq3(qx[i], qy[i], qz[i]);
        }
    }

    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();

    uint32_t phase = 0;
    for (int base = 0; base < N; base += chunk) {
        const int n = min(chunk, N - base);
        const float *gsrc = cloud + (size_t)base * 3;
        const uint32_t bytes = (uint32_t)n * 12u;
        // TMA bulk copy needs 16-byte aligned source and size; otherwise plain coalesced loads.
        const bool tma_ok = ((reinterpret_cast<uintptr_t>(gsrc) & 15u) == 0) && ((bytes & 15u) == 0);
        if (tma_ok) {
            if (tid == 0) {
                mbar_expect_tx(&s_bar, bytes);
                bulk_g2s(s_p, gsrc, bytes, &s_bar);
            }
            mbar_wait(&s_bar, phase);
            phase ^= 1;
        } else {
            for (int i = tid; i < n * 3; i += kBqWarps * 32) s_p[i] = gsrc[i];
            __syncthreads();
        }
        for (int j = tid; j < n; j += kBqWarps * 32) {
            const float x = s_p[j * 3 + 0], y = s_p[j * 3 + 1], z = s_p[j * 3 + 2];
            s_q[j] = make_float4(x, y, z, sq3(x, y, z));
        }
        __syncthreads();

#pragma unroll
        for (int i = 0; i < kBqQPW; ++i) {
            if (s0 + i >= S || cnt[i] >= K) continue;
            IdxT *out = out_idx + ((size_t)b * S + s0 + i) * K;
            for (int j0 = 0; j0 < n; j0 += 32) {
                const int j = j0 + lane;
                bool in = false;
                if (j < n) {
                    const float4 p = s_q[j];
                    const float d = sqdist_expanded(qx[i], qy[i], qz[i], qn[i], p.x, p.y, p.z, p.w);
                    in = !(d > radius2);  // layers.py:112 masks "> r^2" OUT
                }
                const unsigned m = __ballot_sync(0xffffffffu, in);
                if (m) {
                    if (first[i] < 0) first[i] = base + j0 + __ffs(m) - 1;
                    const int pos = cnt[i] + __popc(m & ((1u << lane) - 1u));
                    if (in && pos < K) out[pos] = (IdxT)(base + j);
                    cnt[i] += __popc(m);
                    if (cnt[i] >= K) break;
                }
            }
        }
        __syncthreads();  // everyone done with s_q before the next stage overwrites it
    }
#pragma unroll
    for (int i = 0; i < kBqQPW; ++i) {
        if (s0 + i >= S) continue;
        IdxT *out = out_idx + ((size_t)b * S + s0 + i) * K;
        const IdxT pad = (IdxT)(cnt[i] > 0 ? first[i] : N);  // empty ball: N, as the sort leaves it
        for (int k = min(cnt[i], K) + lane; k < K; k += 32) out[k] = pad;
        if (cnt[i] == 0 && lane == 0 && empty_count != nullptr) atomicAdd(empty_count, 1);
    }
}


// ------------------------------------------------------------------ query_ball_point, several radii at once
// PointNetSetAbstractionMsg (layers.py:258-267) runs query_ball_point once per radius over the SAME centroids:
// three (two) full distance passes per layer.  Here one pass evaluates every (query, point) distance once and
// compacts into up to kBqMaxR index lists (one ballot per radius on the same distance register); the cloud is
// staged once.  Results are identical to R separate papc_ball_query_f32 calls.
constexpr int kBqMaxR = 4;
struct BqMultiArgs {
    float r2[kBqMaxR];
    int K[kBqMaxR];
    void *out[kBqMaxR];
    int32_t *empty;   // [R] or null
    int R;
};

template <typename IdxT>
__global__ void __launch_bounds__(kBqWarps * 32)
ball_query_multi_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz, int N, int S,
                        const BqMultiArgs a) {
    extern __shared__ __align__(128) float s_dyn[];
    const int chunk = N < kBqChunk ? ((N + 3) & ~3) : kBqChunk;
    float4 *s_q = reinterpret_cast<float4 *>(s_dyn);
    float *s_p = s_dyn + chunk * 4;
    __shared__ __align__(8) uint64_t s_bar;
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = blockIdx.x * kBqWarps + warp;
    const float *cloud = xyz + (size_t)b * N * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f, qn = 0.f;
    if (s < S) {
        const float *q = new_xyz + ((size_t)b * S + s) * 3;
        qx = q[0]; qy = q[1]; qz = q[2];
        qn = sq3(qx, qy, qz);
    }
    int cnt[kBqMaxR], first[kBqMaxR];
#pragma unroll
    for (int r = 0; r < kBqMaxR; ++r) { cnt[r] = 0; first[r] = -1; }
    if (tid == 0) {
        mbar_init(&s_bar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t phase = 0;
    for (int base = 0; base < N; base += chunk) {
        const int n = min(chunk, N - base);
        const float *gsrc = cloud + (size_t)base * 3;
        const uint32_t bytes = (uint32_t)n * 12u;
        const bool tma_ok = ((reinterpret_cast<uintptr_t>(gsrc) & 15u) == 0) && ((bytes & 15u) == 0);
        if (tma_ok) {
            if (tid == 0) {
                mbar_expect_tx(&s_bar, bytes);
                bulk_g2s(s_p, gsrc, bytes, &s_bar);
            }
            mbar_wait(&s_bar, phase);
            phase ^= 1;
        } else {
            for (int i = tid; i < n * 3; i += kBqWarps * 32) s_p[i] = gsrc[i];
            __syncthreads();
        }
        for (int j = tid; j < n; j += kBqWarps * 32) {
            const float x = s_p[j * 3 + 0], y = s_p[j * 3 + 1], z = s_p[j * 3 + 2];
            s_q[j] = make_float4(x, y, z, sq3(x, y, z));
        }
        __syncthreads();
        if (s < S) {
            for (int j0 = 0; j0 < n; j0 += 32) {
                bool open = false;   // any list still short?
#pragma unroll
                for (int r = 0; r < kBqMaxR; ++r) open = open || (r < a.R && cnt[r] < a.K[r]);
                if (!open) break;
                const int j = j0 + lane;
                float d = INFINITY;
                if (j < n) {
                    const float4 p = s_q[j];
                    d = sqdist_expanded(qx, qy, qz, qn, p.x, p.y, p.z, p.w);
                }
#pragma unroll
                for (int r = 0; r < kBqMaxR; ++r) {
                    if (r >= a.R || cnt[r] >= a.K[r]) continue;
                    const bool in = (j < n) && !(d > a.r2[r]);   // layers.py:112 masks "> r^2" OUT
                    const unsigned m = __ballot_sync(0xffffffffu, in);
                    if (m) {
                        if (first[r] < 0) first[r] = base + j0 + __ffs(m) - 1;
                        const int pos = cnt[r] + __popc(m & ((1u << lane) - 1u));
                        IdxT *out = reinterpret_cast<IdxT *>(a.out[r]) + ((size_t)b * S + s) * a.K[r];
                        if (in && pos < a.K[r]) out[pos] = (IdxT)(base + j);
                        cnt[r] += __popc(m);
                    }
                }
            }
        }
        __syncthreads();
    }
    if (s >= S) return;
#pragma unroll
    for (int r = 0; r < kBqMaxR; ++r) {
        if (r >= a.R) continue;
        IdxT *out = reinterpret_cast<IdxT *>(a.out[r]) + ((size_t)b * S + s) * a.K[r];
        const IdxT pad = (IdxT)(cnt[r] > 0 ? first[r] : N);
        for (int k = min(cnt[r], a.K[r]) + lane; k < a.K[r]; k += 32) out[k] = pad;
        if (cnt[r] == 0 && lane == 0 && a.empty != nullptr) atomicAdd(a.empty + r, 1);
    }
}

// ------------------------------------------------------------------ k nearest neighbours
// The k (<= 32) nearest points of every query under square_distance (expansion form, as A1), ascending
// distance, ties -> the lower index (a stable argsort of the distance row, what layers.py:316-318 takes its
// first three columns from).  One warp per query; lane i holds the i-th best (distance, index) so far, a
// candidate that beats the k-th is inserted with one ballot + two shuffles.
template <typename IdxT>
__global__ void __launch_bounds__(kBqWarps * 32)
knn_kernel(const float *__restrict__ xyz, const float *__restrict__ query, int N, int S, int k,
           IdxT *__restrict__ out_idx, float *__restrict__ out_dist) {
    extern __shared__ __align__(128) float s_dyn[];
    const int chunk = N < kBqChunk ? ((N + 3) & ~3) : kBqChunk;
    float4 *s_q = reinterpret_cast<float4 *>(s_dyn);
    const int b = blockIdx.y;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = blockIdx.x * kBqWarps + warp;
    const float *cloud = xyz + (size_t)b * N * 3;
    float qx = 0.f, qy = 0.f, qz = 0.f, qn = 0.f;
    if (s < S) {
        const float *q = query + ((size_t)b * S + s) * 3;
        qx = q[0]; qy = q[1]; qz = q[2];
        qn = sq3(qx, qy, qz);
    }
    float bd = INFINITY;      // lane i: i-th best distance (lanes >= k stay at +inf and never matter)
    int bi = 0x7fffffff;
    for (int base = 0; base < N; base += chunk) {
        const int n = min(chunk, N - base);
        __syncthreads();
        for (int j = tid; j < n; j += kBqWarps * 32) {
            const float *g = cloud + (size_t)(base + j) * 3;
            const float x = g[0], y = g[1], z = g[2];
            s_q[j] = make_float4(x, y, z, sq3(x, y, z));
        }
        __syncthreads();
        if (s >= S) continue;
        for (int j0 = 0; j0 < n; j0 += 32) {
            const int j = j0 + lane;
            float d = INFINITY;
            if (j < n) {
                const float4 p = s_q[j];
                d = sqdist_expanded(qx, qy, qz, qn, p.x, p.y, p.z, p.w);
            }
            const int gj = base + j;
            float thr_d = __shfl_sync(0xffffffffu, bd, k - 1);
            int thr_i = __shfl_sync(0xffffffffu, bi, k - 1);
            unsigned m = __ballot_sync(0xffffffffu, (j < n) && (d < thr_d || (d == thr_d && gj < thr_i)));
            while (m) {
                const int src = __ffs(m) - 1;
                m &= m - 1;
                const float cd = __shfl_sync(0xffffffffu, d, src);
                const int ci = __shfl_sync(0xffffffffu, gj, src);
                if (!(cd < thr_d || (cd == thr_d && ci < thr_i))) continue;   // the threshold moved meanwhile
                // position = entries that sort before the candidate
                const bool before = bd < cd || (bd == cd && bi < ci);
                const int pos = __popc(__ballot_sync(0xffffffffu, before));
                const float ud = __shfl_up_sync(0xffffffffu, bd, 1);
                const int ui = __shfl_up_sync(0xffffffffu, bi, 1);
                if (lane == pos) { bd = cd; bi = ci; }
                else if (lane > pos) { bd = ud; bi = ui; }
                thr_d = __shfl_sync(0xffffffffu, bd, k - 1);
                thr_i = __shfl_sync(0xffffffffu, bi, k - 1);
            }
        }
    }
    if (s < S && lane < k) {
        const size_t o = ((size_t)b * S + s) * k + lane;
        out_idx[o] = (IdxT)(bi == 0x7fffffff ? N : bi);   // fewer than k points: N, like an exhausted sort
        if (out_dist != nullptr) out_dist[o] = bd;
    }
}

// ------------------------------------------------------------------ group gather (A5)
__global__ void __launch_bounds__(256)
group_gather_kernel(const float *__restrict__ xyz, const float *__restrict__ new_xyz,
                    const float *__restrict__ feats, const int64_t *__restrict__ idx, int N, int S,
                    int K, int D, int order, float *__restrict__ out, size_t total) {
    const int C = 3 + D;
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; e < total; e += stride) {
        const int c = (int)(e % C);
        const size_t row = e / C;      // (b*S + s)*K + k
        const size_t g = row / K;      // b*S + s
        const size_t b = g / S;
        long long n = idx[row];
        n = n < 0 ? 0 : (n >= N ? N - 1 : n);
        const int cx = (order == PAPC_XYZ_FIRST) ? c : c - D;  // xyz channel if in [0,3)
        float v;
        if (cx >= 0 && cx < 3) {
            v = __fsub_rn(xyz[(b * N + (size_t)n) * 3 + cx], new_xyz[g * 3 + cx]);
        } else {
            const int cf = (order == PAPC_XYZ_FIRST) ? c - 3 : c;
            v = feats[(b * N + (size_t)n) * D + cf];
        }
        out[e] = v;
    }
}

static int grid_for(size_t total, int threads) {
    size_t blocks = (total + threads - 1) / threads;
    const size_t cap = (size_t)kNumSMs * 16;
    return (int)(blocks < cap ? (blocks ? blocks : 1) : cap);
}

}  // namespace papc

using namespace papc;

extern "C" int papc_square_distance_f32(const float *src, const float *dst, int B, int N, int M,
                                        float *out, papc_stream_t stream) {
    if (B < 0 || N < 0 || M < 0) return PAPC_EINVAL;
    const size_t total = (size_t)B * N * M;
    if (total == 0) return PAPC_OK;
    if (!src || !dst || !out) return PAPC_EINVAL;
    square_distance_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(src, dst, N, M, out,
                                                                               total);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

extern "C" int papc_gather_f32(const float *points, const int64_t *idx, int B, int N, int C, int M,
                               float *out, papc_stream_t stream) {
    if (B < 0 || N <= 0 || C < 0 || M < 0) return PAPC_EINVAL;
    if ((size_t)B * M * C == 0) return PAPC_OK;
    if (!points || !idx || !out) return PAPC_EINVAL;
    cudaStream_t st = as_stream(stream);
    const bool vec4 = (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15u) == 0) &&
                      ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    if (vec4) {
        const size_t total = (size_t)B * M * (C / 4);
        gather_rows_kernel<float4><<<grid_for(total, 256), 256, 0, st>>>(points, idx, N, C / 4, M,
                                                                         total, out);
    } else {
        const size_t total = (size_t)B * M * C;
        gather_rows_kernel<float><<<grid_for(total, 256), 256, 0, st>>>(points, idx, N, C, M, total,
                                                                        out);
    }
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

extern "C" int papc_ball_query_f32(const float *xyz, const float *new_xyz, int B, int N, int S,
                                   float radius2, int nsample, void *out_idx, int idx_bits,
                                   int32_t *empty_count, papc_stream_t stream) {
    if (B < 0 || N <= 0 || S < 0 || nsample <= 0) return PAPC_EINVAL;
    if (idx_bits != 32 && idx_bits != 64) return PAPC_EINVAL;
    if (nsample > N) return PAPC_EINVAL;  // the reference fails with a shape mismatch
    if (B == 0 || S == 0) return PAPC_OK;
    if (!xyz || !new_xyz || !out_idx) return PAPC_EINVAL;
    if (B > 65535) return PAPC_EUNSUPPORTED;
    dim3 grid(ceil_div(S, kBqWarps * kBqQPW), B);
    ProfScope prof(as_stream(stream), "ball_query", (long long)B * S, N, nsample, 0.0,
                   12.0 * B * (N + S) + (idx_bits / 8.0) * B * S * nsample);
    const int chunk = N < kBqChunk ? ((N + 3) & ~3) : kBqChunk;
    const size_t smem = (size_t)chunk * 7 * sizeof(float);
    if (smem > 48 * 1024) {  // only the largest chunk needs the opt-in
        PAPC_CUDA_TRY(cudaFuncSetAttribute(ball_query_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PAPC_CUDA_TRY(cudaFuncSetAttribute(ball_query_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (idx_bits == 64)
        ball_query_kernel<int64_t><<<grid, kBqWarps * 32, smem, as_stream(stream)>>>(
            xyz, new_xyz, N, S, radius2, nsample, reinterpret_cast<int64_t *>(out_idx), empty_count);
    else
        ball_query_kernel<int32_t><<<grid, kBqWarps * 32, smem, as_stream(stream)>>>(
            xyz, new_xyz, N, S, radius2, nsample, reinterpret_cast<int32_t *>(out_idx), empty_count);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}


extern "C" int papc_ball_query_multi_f32(const float *xyz, const float *new_xyz, int B, int N, int S, int R,
                                         const float *radius2_host, const int32_t *nsample_host,
                                         void *const *out_idx_host, int idx_bits, int32_t *empty_count,
                                         papc_stream_t stream) {
    if (B < 0 || N <= 0 || S < 0 || R < 1 || R > kBqMaxR || !radius2_host || !nsample_host || !out_idx_host)
        return PAPC_EINVAL;
    if (idx_bits != 32 && idx_bits != 64) return PAPC_EINVAL;
    BqMultiArgs a{};
    a.R = R;
    a.empty = empty_count;
    double out_bytes = 0.0;
    for (int r = 0; r < R; ++r) {
        if (nsample_host[r] <= 0 || nsample_host[r] > N) return PAPC_EINVAL;
        if (!out_idx_host[r] && B * S > 0) return PAPC_EINVAL;
        a.r2[r] = radius2_host[r];
        a.K[r] = nsample_host[r];
        a.out[r] = out_idx_host[r];
        out_bytes += (idx_bits / 8.0) * B * S * nsample_host[r];
    }
    if (B == 0 || S == 0) return PAPC_OK;
    if (!xyz || !new_xyz) return PAPC_EINVAL;
    if (B > 65535) return PAPC_EUNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    dim3 grid(ceil_div(S, kBqWarps), B);
    ProfScope prof(st, "ball_query_multi", (long long)B * S, N, R, 0.0, 12.0 * B * (N + S) + out_bytes);
    const int chunk = N < kBqChunk ? ((N + 3) & ~3) : kBqChunk;
    const size_t smem = (size_t)chunk * 7 * sizeof(float);
    if (smem > 48 * 1024) {
        PAPC_CUDA_TRY(cudaFuncSetAttribute(ball_query_multi_kernel<int64_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        PAPC_CUDA_TRY(cudaFuncSetAttribute(ball_query_multi_kernel<int32_t>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    }
    if (idx_bits == 64) ball_query_multi_kernel<int64_t><<<grid, kBqWarps * 32, smem, st>>>(xyz, new_xyz, N, S, a);
    else ball_query_multi_kernel<int32_t><<<grid, kBqWarps * 32, smem, st>>>(xyz, new_xyz, N, S, a);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

extern "C" int papc_knn_f32(const float *xyz, const float *query, int B, int N, int S, int k, void *out_idx,
                            int idx_bits, float *out_dist, papc_stream_t stream) {
    if (B < 0 || N <= 0 || S < 0 || k < 1 || k > 32) return PAPC_EINVAL;
    if (idx_bits != 32 && idx_bits != 64) return PAPC_EINVAL;
    if (B == 0 || S == 0) return PAPC_OK;
    if (!xyz || !query || !out_idx) return PAPC_EINVAL;
    if (B > 65535) return PAPC_EUNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    dim3 grid(ceil_div(S, kBqWarps), B);
    ProfScope prof(st, "knn", (long long)B * S, N, k, 0.0, 12.0 * B * (N + S) + (idx_bits / 8.0 + 4.0) * B * S * k);
    const int chunk = N < kBqChunk ? ((N + 3) & ~3) : kBqChunk;
    const size_t smem = (size_t)chunk * 4 * sizeof(float);
    if (idx_bits == 64)
        knn_kernel<int64_t><<<grid, kBqWarps * 32, smem, st>>>(xyz, query, N, S, k, reinterpret_cast<int64_t *>(out_idx), out_dist);
    else
        knn_kernel<int32_t><<<grid, kBqWarps * 32, smem, st>>>(xyz, query, N, S, k, reinterpret_cast<int32_t *>(out_idx), out_dist);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

extern "C" int papc_group_gather_f32(const float *xyz, const float *new_xyz, const float *feats,
                                     const int64_t *idx, int B, int N, int S, int K, int D,
                                     int order, float *out, papc_stream_t stream) {
    if (B < 0 || N <= 0 || S < 0 || K < 0 || D < 0) return PAPC_EINVAL;
    if (order != PAPC_XYZ_FIRST && order != PAPC_FEATS_FIRST) return PAPC_EINVAL;
    const size_t total = (size_t)B * S * K * (3 + D);
    if (total == 0) return PAPC_OK;
    if (!xyz || !new_xyz || !idx || !out || (D > 0 && !feats)) return PAPC_EINVAL;
    group_gather_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(
        xyz, new_xyz, feats, idx, N, S, K, D, order, out, total);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}
