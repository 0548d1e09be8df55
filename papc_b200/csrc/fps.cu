// fps.cu -- farthest point sampling (reference: layers.py:65-95) for sm_100a.
//
// One persistent CTA per cloud runs all `npoint` dependent iterations; the kernel is LATENCY bound
// (npoint serial argmax steps over a 12 KB cloud), so everything is arranged to shorten the
// per-iteration dependency chain:
//   * the cloud's xyz and the running min-distance live in REGISTERS (PPT points per thread, blocked
//     so that lower threads own lower indices), two points per packed f32x2 instruction;
//   * the per-thread / per-warp argmax works on the fp32 bit pattern as a signed integer (all
//     distances are >= 0): integer max in the thread, redux.sync in the warp, ballot for the lowest
//     lane holding the maximum;
//   * that lane publishes (distance bits, index) of its candidate, so after the ONE __syncthreads
//     of the iteration every thread reads all warps' candidates (broadcast LDS.128, two candidates
//     each) and reduces them itself in a fixed tree (lower warp wins ties == lowest index): no
//     second redux / ballot / shuffle round;
//   * the cloud is also kept in shared memory as float4 so the next centroid is ONE LDS.128.
//
// Bit-exactness vs the oracle: d = (dx*dx + dy*dy) + dz*dz with separately rounded ops (packed
// subtractions and squares, scalar add.rn sums -- checked in SASS: no FFMA in the loop;
// x - c == x + (-c) exactly), running = min(running, d)
// (== the reference's masked assignment for non-NaN data), argmax = lowest index among maxima
// (paddle.argmax).
#include "common.cuh"

#include <stdlib.h>

namespace papc {

constexpr int kFpsRegMaxN = 8192;
constexpr int kFpsOutCap = 4096;  // picks parked in shared memory between flushes to global

__device__ __forceinline__ uint64_t f2pack(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void f2unpack(uint64_t v, float &lo, float &hi) {
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ uint64_t f2add(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
// explicit shared-window accesses on 32-bit addresses computed once: generic accesses to __shared__
// objects make ptxas re-derive the window base (S2R SR_CgaCtaId, ~30 cycles) inside the hot loop
__device__ __forceinline__ uint32_t fps_smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ float4 fps_lds128f(uint32_t saddr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(saddr)
                 : "memory");
    return v;
}
__device__ __forceinline__ int4 fps_lds128i(uint32_t saddr) {
    int4 v;
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(saddr)
                 : "memory");
    return v;
}
__device__ __forceinline__ int2 fps_lds64i(uint32_t saddr) {
    int2 v;
    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void fps_sts64i(uint32_t saddr, int a, int b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(saddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ uint64_t f2mul(uint64_t a, uint64_t b) {
    uint64_t r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}

// DBG (timing experiments only, results are wrong): 1 = no cross-warp exchange, 2 = no warp redux /
// ballot either, 4 = no centroid LDS dependency
// PUB: every pick is also stored to pub[it] (global, relaxed) the moment it is known, so that the consumer
// CTAs of sample_group_kernel can run the ball query of centroid `it` while the recurrence continues.
// CTA-wide barrier of the FPS role: the whole CTA, or -- inside sample_group_kernel, whose CTAs carry more
// warps than the FPS role uses -- named barrier 1 over the role's THREADS threads.
template <int THREADS, bool NAMED>
__device__ __forceinline__ void fps_sync() {
    if (NAMED) asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
    else __syncthreads();
}

template <int THREADS, int PPT, int DBG, bool PUB>
__device__ __forceinline__ void
fps_reg_body(const int b, const float *__restrict__ xyz, int N, int npoint,
             const int64_t *__restrict__ start_idx, float init_dist,
             int64_t *__restrict__ out_idx, float *__restrict__ out_new_xyz, int32_t *__restrict__ pub) {
    static_assert(PPT % 2 == 0, "points are processed in packed pairs");
    extern __shared__ float4 s_xyz4[];  // [N] (x, y, z, 0): one LDS.128 fetches a centroid
    int *s_out = reinterpret_cast<int *>(s_xyz4 + N);  // [min(npoint, kFpsOutCap)] picks awaiting the flush
    constexpr int NW = THREADS / 32;
    constexpr int PP = PPT / 2;
    constexpr int NWP = NW > 1 ? NW : 2;
    __shared__ __align__(16) int2 s_red[2][NWP];  // per warp: (distance bits, point index)

    const int tid = threadIdx.x;
    const int lane = tid & 31;
    const int warp = tid >> 5;
    const float *cloud = xyz + (size_t)b * N * 3;

    for (int i = tid; i < N; i += THREADS)
        s_xyz4[i] = make_float4(cloud[i * 3 + 0], cloud[i * 3 + 1], cloud[i * 3 + 2], 0.f);
    fps_sync<THREADS, PUB>();

    uint64_t px[PP], py[PP], pz[PP];
    float pd[PPT];
#pragma unroll
    for (int q = 0; q < PP; ++q) {
        float4 v[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int j = tid * PPT + 2 * q + h;
            if (j < N) {
                v[h] = s_xyz4[j];
                pd[2 * q + h] = init_dist;
            } else {
                v[h] = make_float4(0.f, 0.f, 0.f, 0.f);
                pd[2 * q + h] = -1.0f;  // negative bit pattern: never the (signed) maximum
            }
        }
        px[q] = f2pack(v[0].x, v[1].x);
        py[q] = f2pack(v[0].y, v[1].y);
        pz[q] = f2pack(v[0].z, v[1].z);
    }

    int far = (int)start_idx[b];
    far = min(max(far, 0), N - 1);
    unsigned lt_mask;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(lt_mask));
    // window addresses laundered through a volatile shared-memory round trip: ptxas otherwise
    // rematerialises them (an S2UR of SR_CgaCtaId at the head of the dependency chain) in every
    // iteration instead of keeping two registers
    __shared__ uint32_t s_addr[3];
    if (tid == 0) {
        s_addr[0] = fps_smem_u32(s_xyz4);
        s_addr[1] = fps_smem_u32(&s_red[0][0]);
        s_addr[2] = fps_smem_u32(s_out);
    }
    fps_sync<THREADS, PUB>();
    uint32_t xyz_sa, red_sa, out_sa;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(out_sa) : "r"(fps_smem_u32(&s_addr[2])) : "memory");
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(xyz_sa) : "r"(fps_smem_u32(&s_addr[0])) : "memory");
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(red_sa) : "r"(fps_smem_u32(&s_addr[1])) : "memory");

    int64_t *out = out_idx + (size_t)b * npoint;
    float *oxyz = out_new_xyz != nullptr ? out_new_xyz + (size_t)b * npoint * 3 : nullptr;
    auto flush = [&](int base, int n) {  // picks [base, base+n) -> global (all threads)
        fps_sync<THREADS, PUB>();
        for (int i = tid; i < n; i += THREADS) out[base + i] = s_out[i];
        if (oxyz != nullptr)
            for (int i = tid; i < n * 3; i += THREADS) {
                const int pt = i / 3, d = i - pt * 3;
                const float4 v = s_xyz4[s_out[pt]];
                oxyz[(size_t)base * 3 + i] = d == 0 ? v.x : (d == 1 ? v.y : v.z);
            }
        fps_sync<THREADS, PUB>();
    };

    int32_t *pub_ptr = pub;
    const uint32_t is_pub = (PUB && tid == THREADS - 1) ? 1u : 0u;
    for (int it = 0; it < npoint; ++it) {
        const float4 c = fps_lds128f(xyz_sa + 16u * (uint32_t)((DBG & 4) ? (it & 1023) : far));
        const uint32_t red_it = red_sa + (uint32_t)(it & 1) * (NWP * 8);
        // the pick is parked in shared memory (one predicated STS): global stores with their 64-bit
        // address arithmetic would sit on warp 0's path to the barrier in every iteration
        if (tid == 0)
            asm volatile("st.shared.u32 [%0], %1;" ::"r"(out_sa + 4u * (uint32_t)(it & (kFpsOutCap - 1))), "r"(far)
                         : "memory");
        // the last warp's spare lane publishes the pick: ONE predicated store through a per-thread running pointer
        // (an `if` here became a divergent branch with a reconvergence barrier at the head of every iteration and
        // cost 20 % of the recurrence).  A weak store: the consumers only ever need the 4-byte value itself.
        if (PUB) {
            asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.global.s32 [%0], %1;\n\t}"
                         ::"l"(pub_ptr), "r"(far), "r"(is_pub) : "memory");
            ++pub_ptr;
        }
        if (it == npoint - 1) break;  // the reference's last argmax is discarded (:93 after :80)
        if ((it & (kFpsOutCap - 1)) == kFpsOutCap - 1) flush(it + 1 - kFpsOutCap, kFpsOutCap);  // rare
        const uint64_t ncx = f2pack(-c.x, -c.x), ncy = f2pack(-c.y, -c.y), ncz = f2pack(-c.z, -c.z);
        int bits[PPT];
#pragma unroll
        for (int q = 0; q < PP; ++q) {
            const uint64_t dx = f2add(px[q], ncx);
            const uint64_t dy = f2add(py[q], ncy);
            const uint64_t dz = f2add(pz[q], ncz);
            // squares packed, sums as scalar add.rn.f32: ptxas contracts mul.rn.f32x2 + add.rn.f32x2
            // into FFMA2 (even under -fmad=false) but never touches an explicit scalar add.rn
            float xx0, xx1, yy0, yy1, zz0, zz1;
            f2unpack(f2mul(dx, dx), xx0, xx1);
            f2unpack(f2mul(dy, dy), yy0, yy1);
            f2unpack(f2mul(dz, dz), zz0, zz1);
            const float d0 = __fadd_rn(__fadd_rn(xx0, yy0), zz0);
            const float d1 = __fadd_rn(__fadd_rn(xx1, yy1), zz1);
            // tid*PPT+p >= N keeps pd = -1 (d >= 0 is never smaller)
            pd[2 * q] = fminf(pd[2 * q], d0);
            pd[2 * q + 1] = fminf(pd[2 * q + 1], d1);
            bits[2 * q] = __float_as_int(pd[2 * q]);
            bits[2 * q + 1] = __float_as_int(pd[2 * q + 1]);
        }
        // thread maximum as a balanced tree (short dependency chain)
        int tmax[PPT];
#pragma unroll
        for (int p = 0; p < PPT; ++p) tmax[p] = bits[p];
#pragma unroll
        for (int step = 1; step < PPT; step <<= 1)
#pragma unroll
            for (int p = 0; p + step < PPT; p += 2 * step) tmax[p] = max(tmax[p], tmax[p + step]);
        const int best = tmax[0];
        const int wmax = (DBG & 2) ? best : __reduce_max_sync(0xffffffffu, best);
        // first point of this thread holding its maximum (independent of the redux result)
        unsigned eq = 0u;
#pragma unroll
        for (int p = 0; p < PPT; ++p) eq |= (bits[p] == best ? 1u : 0u) << p;
        const int bi = __ffs(eq) - 1;
        const bool mine = best == wmax;
        const unsigned ball = (DBG & 2) ? 1u : __ballot_sync(0xffffffffu, mine);
        if (DBG & 1) {
            far = (tid * PPT + bi + it) & 1023;
        } else if (NW == 1) {
            const int src = __ffs(ball) - 1;  // lowest lane == lowest index
            far = __shfl_sync(0xffffffffu, tid * PPT + bi, src);
        } else {
            {   // lowest lane holding the warp maximum publishes it: a PREDICATED store, not a branch (a divergent
                // branch + reconvergence barrier in front of the bar.sync sits on every iteration's critical path)
                const uint32_t wr = (mine && (ball & lt_mask) == 0u) ? 1u : 0u;
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %3, 0;\n\t@p st.shared.v2.b32 [%0], {%1, %2};\n\t}"
                             ::"r"(red_it + 8u * warp), "r"(wmax), "r"(tid * PPT + bi), "r"(wr) : "memory");
            }
            fps_sync<THREADS, PUB>();
            if (NW <= 8) {
                // every thread reduces all warps' candidates itself: broadcast LDS.128 (two
                // candidates each), fixed tree, the lower warp (= lower indices) wins ties
                int kb[NWP], ki[NWP];
#pragma unroll
                for (int w = 0; w < NWP; w += 2) {
                    const int4 e = fps_lds128i(red_it + 8u * w);
                    kb[w] = e.x; ki[w] = e.y; kb[w + 1] = e.z; ki[w + 1] = e.w;
                }
#pragma unroll
                for (int step = 1; step < NW; step <<= 1)
#pragma unroll
                    for (int w = 0; w + step < NW; w += 2 * step) {
                        const bool tb = kb[w + step] > kb[w];
                        kb[w] = tb ? kb[w + step] : kb[w];
                        ki[w] = tb ? ki[w + step] : ki[w];
                    }
                far = ki[0];
            } else {
                // many warps: lane w holds warp w's candidate, second redux round + shuffle
                const int2 e = fps_lds64i(red_it + 8u * (lane < NW ? lane : 0));
                const int kb = lane < NW ? e.x : (-2147483647 - 1);
                const int m2 = __reduce_max_sync(0xffffffffu, kb);
                const int src = __ffs(__ballot_sync(0xffffffffu, kb == m2)) - 1;  // lowest warp
                far = __shfl_sync(0xffffffffu, e.y, src);
            }
        }
    }
    flush((npoint - 1) & ~(kFpsOutCap - 1), ((npoint - 1) & (kFpsOutCap - 1)) + 1);
}

template <int THREADS, int PPT, int DBG = 0>
__global__ void __launch_bounds__(THREADS)
fps_reg_kernel(const float *__restrict__ xyz, int N, int npoint,
               const int64_t *__restrict__ start_idx, float init_dist,
               int64_t *__restrict__ out_idx, float *__restrict__ out_new_xyz) {
    fps_reg_body<THREADS, PPT, DBG, false>(blockIdx.x, xyz, N, npoint, start_idx, init_dist, out_idx, out_new_xyz, nullptr);
}

// ------------------------------------------------------------------ FPS + ball query + moments in ONE launch
// sample_and_group's three dependent stages (layers.py:143-146) for the SetAbstraction layers: CTAs [0, B) run
// the FPS recurrence (one per cloud, exactly fps_reg_kernel) and publish every pick as it is made; CTAs
// [B, B + B*P) are consumers -- P per cloud, the cloud staged once in shared memory as (x, y, z, |p|^2) -- whose
// warps take the centroids of their cloud round-robin, wait for the pick, and run query_ball_point for it
// (the same in-order ballot / popcount scan as ball_query_kernel) while the recurrence is still going.  The
// in-radius points are in registers at that moment, so the nine second-moment sums of the centred neighbours
// that the folded first MLP layer needs (MomentArgs, sa_mlp_tt.cuh) are accumulated on the way: the separate
// 15 us gather pass and all but the last few microseconds of the 30 us ball query leave the critical path.
// All CTAs are co-resident (host: B + B*P <= number of SMs) and the FPS CTAs have the lowest block indices, so
// the consumers' spin-waits cannot starve a producer.
constexpr int kSgThreads = 512;   // consumer CTAs: 16 warps hide the scan's shared-memory latency
// the FPS role uses the first FT / 32 warps of its CTA (the rest exit at once)
template <int FT, int PPT>
__global__ void __launch_bounds__(kSgThreads)
sample_group_kernel(const float *__restrict__ xyz, int B, int N, int npoint, const int64_t *__restrict__ start_idx,
                    float init_dist, int64_t *__restrict__ out_idx, float *__restrict__ out_new_xyz,
                    int32_t *__restrict__ pub, int P, float radius2, int K, int32_t *__restrict__ grp_idx,
                    int32_t *__restrict__ empty_count, double *__restrict__ mom_partial, int dbg) {
    if ((int)blockIdx.x < B) {
        if (threadIdx.x >= FT) return;
        fps_reg_body<FT, PPT, 0, true>(blockIdx.x, xyz, N, npoint, start_idx, init_dist, out_idx, out_new_xyz,
                                               pub + (size_t)blockIdx.x * npoint);
        return;
    }
    extern __shared__ float4 s_q[];   // [N] (x, y, z, |p|^2)
    __shared__ double s_red[kSgThreads / 32][9];
    const int c = (int)blockIdx.x - B;
    const int b = c / P, part = c - b * P;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float *cloud = xyz + (size_t)b * N * 3;
    for (int j = tid; j < N; j += kSgThreads) {
        const float x = cloud[j * 3 + 0], y = cloud[j * 3 + 1], z = cloud[j * 3 + 2];
        s_q[j] = make_float4(x, y, z, sq3(x, y, z));
    }
    __syncthreads();
    const int32_t *picks = pub + (size_t)b * npoint;
    constexpr int NW = kSgThreads / 32;
#ifdef PAPC_TRIAGE
    if (dbg & 1) return;   // timing experiment: producers only
#endif
    double dacc[9];
#pragma unroll
    for (int i = 0; i < 9; ++i) dacc[i] = 0.0;
    for (int s = part * NW + warp; s < npoint; s += NW * P) {
        int pick = 0;
        if (lane == 0) {
            for (uint32_t spin = 0;; ++spin) {
                asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(pick) : "l"(picks + s) : "memory");
                if (pick >= 0) break;
                __nanosleep(100);
                if (spin > (1u << 24)) __trap();   // a lost producer traps instead of hanging the GPU
            }
        }
        pick = __shfl_sync(0xffffffffu, pick, 0);
#ifdef PAPC_TRIAGE
        if (dbg & 2) continue;   // timing experiment: wait for the picks, no scan
#endif
        const float4 q = s_q[pick];
        int32_t *out = grp_idx + ((size_t)b * npoint + s) * K;
        int cnt = 0, first = -1;
        // in-order scan, 64 points per trip: both loads and both distances are in flight before the first ballot
        for (int j0 = 0; j0 < N && cnt < K; j0 += 64) {
            const int ja = j0 + lane, jb = j0 + 32 + lane;
            const float4 pa = s_q[ja < N ? ja : 0], pb = s_q[jb < N ? jb : 0];
            const float da = sqdist_expanded(q.x, q.y, q.z, q.w, pa.x, pa.y, pa.z, pa.w);
            const float db = sqdist_expanded(q.x, q.y, q.z, q.w, pb.x, pb.y, pb.z, pb.w);
            const bool ina = ja < N && !(da > radius2);   // layers.py:112 masks "> r^2" OUT
            const bool inb = jb < N && !(db > radius2);
            const unsigned ma = __ballot_sync(0xffffffffu, ina);
            const unsigned mb = __ballot_sync(0xffffffffu, inb);
            if (ma | mb) {
                if (first < 0) first = ma ? j0 + __ffs(ma) - 1 : j0 + 32 + __ffs(mb) - 1;
                const unsigned lt = (1u << lane) - 1u;
                const int posa = cnt + __popc(ma & lt);
                if (ina && posa < K) out[posa] = ja;
                cnt += __popc(ma);
                const int posb = cnt + __popc(mb & lt);
                if (inb && posb < K) out[posb] = jb;
                cnt += __popc(mb);
            }
        }
        const int have = min(cnt, K);
        const int pad = cnt > 0 ? first : N;   // empty ball: N in every slot, as the sort leaves it
        for (int k = have + lane; k < K; k += 32) out[k] = pad;
        if (lane == 0 && cnt == 0 && empty_count != nullptr) atomicAdd(empty_count, 1);
        __syncwarp();   // the warp's own index writes are visible to all of its lanes
        // second moments of the K centred neighbours (padding rows included; an empty ball's N clamps to N-1
        // exactly as the MLP's gather does): every lane takes rows lane, lane+32, ...
        float f[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) f[i] = 0.f;
        if (mom_partial != nullptr) {
            for (int k = lane; k < K; k += 32) {
                const int j = min(out[k], N - 1);
                const float4 p = s_q[j];
                const float x = __fsub_rn(p.x, q.x), y = __fsub_rn(p.y, q.y), z = __fsub_rn(p.z, q.z);
                f[0] += x; f[1] += y; f[2] += z;
                f[3] = fmaf(x, x, f[3]); f[4] = fmaf(x, y, f[4]); f[5] = fmaf(x, z, f[5]);
                f[6] = fmaf(y, y, f[6]); f[7] = fmaf(y, z, f[7]); f[8] = fmaf(z, z, f[8]);
            }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) dacc[i] += (double)f[i];
    }
    if (mom_partial != nullptr) {
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            double v = dacc[i];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) s_red[warp][i] = v;
        }
        __syncthreads();
        if (tid < 9) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < NW; ++w) v += s_red[w][tid];
            mom_partial[(size_t)c * 9 + tid] = v;
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

template <int THREADS, int PPT, int DBG = 0>
static int launch_fps_reg(const float *xyz, int B, int N, int npoint, const int64_t *start,
                          float init_dist, int64_t *out_idx, float *out_new_xyz,
                          cudaStream_t st) {
    const size_t smem = (size_t)N * sizeof(float4) + (size_t)(npoint < kFpsOutCap ? npoint : kFpsOutCap) * sizeof(int);
    auto k = fps_reg_kernel<THREADS, PPT, DBG>;
    if (smem > 48 * 1024)
        PAPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
    ProfScope prof(st, "fps_reg", (long long)B * N, npoint, 0, 0.0,
                   12.0 * B * N + 8.0 * B * npoint + (out_new_xyz ? 12.0 * B * npoint : 0.0));
    k<<<B, THREADS, smem, st>>>(xyz, N, npoint, start, init_dist, out_idx, out_new_xyz);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

}  // namespace papc

extern "C" int papc_sample_group_parts(int B, int N, int npoint, int nsample) {
    // consumer CTAs per cloud, 0 = this shape does not run fused (use papc_fps_f32 + papc_ball_query_f32)
    if (B < 1 || N <= 256 || N > 1024 || npoint < 1 || nsample < 1 || nsample > N) return 0;
    int P = (papc::kNumSMs - B) / B;
    if (P > 4) P = 4;
    return P >= 1 ? P : 0;
}
extern "C" size_t papc_sample_group_workspace_bytes(int B, int npoint) {
    return B > 0 && npoint > 0 ? papc::align_up((size_t)B * npoint * sizeof(int32_t), 256) : 0;
}
extern "C" int papc_sample_group_f32(const float *xyz, int B, int N, int npoint, const int64_t *start_idx,
                                     float init_dist, float radius2, int nsample, int64_t *out_fps_idx,
                                     float *out_new_xyz, int32_t *out_group_idx, int32_t *empty_count,
                                     double *moments_partial, void *workspace, size_t workspace_bytes,
                                     papc_stream_t stream) {
    using namespace papc;
    const int P = papc_sample_group_parts(B, N, npoint, nsample);
    if (P == 0) return PAPC_EUNSUPPORTED;
    if (!(init_dist >= 0.0f) || !xyz || !start_idx || !out_fps_idx || !out_new_xyz || !out_group_idx) return PAPC_EINVAL;
    const size_t need = papc_sample_group_workspace_bytes(B, npoint);
    if (!workspace || workspace_bytes < need) return PAPC_EWORKSPACE;
    cudaStream_t st = as_stream(stream);
    int32_t *pub = reinterpret_cast<int32_t *>(workspace);
    PAPC_CUDA_TRY(cudaMemsetAsync(pub, 0xFF, (size_t)B * npoint * sizeof(int32_t), st));   // -1 = not picked yet
    const size_t smem = (size_t)N * sizeof(float4) + (size_t)(npoint < kFpsOutCap ? npoint : kFpsOutCap) * sizeof(int);
    ProfScope prof(st, "sample_group", (long long)B * N, npoint, nsample, 0.0,
                   12.0 * B * N + 20.0 * B * npoint + 4.0 * B * npoint * nsample);
    const int grid = B + B * P;
    int dbg = 0;
#ifdef PAPC_TRIAGE
    { const char *e = getenv("PAPC_SG_DBG"); dbg = e ? atoi(e) : 0; }
#endif
    // FPS role shape as papc_fps_f32 picks it (PAPC_FPS_WIDE=1: twice the warps, half the points per thread)
    static const bool wide = [] { const char *e = getenv("PAPC_FPS_WIDE"); return e && atoi(e) == 1; }();
#define SG_GO(FT, PPT)                                                                                              \
    do {                                                                                                            \
        auto k = sample_group_kernel<FT, PPT>;                                                                      \
        if (smem > 48 * 1024)                                                                                       \
            PAPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));        \
        k<<<grid, kSgThreads, smem, st>>>(xyz, B, N, npoint, start_idx, init_dist, out_fps_idx, out_new_xyz, pub, P, \
                                           radius2, nsample, out_group_idx, empty_count, moments_partial, dbg);     \
    } while (0)
    if (N <= 512) { if (wide) SG_GO(256, 2); else SG_GO(128, 4); }
    else { if (wide) SG_GO(256, 4); else SG_GO(128, 8); }
#undef SG_GO
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

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
    // threads x points-per-thread; PAPC_FPS_WIDE=1 selects the wider (more warps, fewer points per
    // thread) variant for the mid sizes -- A/B switch for tuning, same results either way
    static const int shape = [] { const char *e = getenv("PAPC_FPS_WIDE"); return e ? atoi(e) : 0; }();
    const bool wide = shape == 1;
#define FPS_GO(T, P) return launch_fps_reg<T, P>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_new_xyz, st)
    if (N <= 64) FPS_GO(32, 2);
    if (N <= 128) FPS_GO(32, 4);
    if (N <= 256) { if (wide) FPS_GO(128, 2); FPS_GO(64, 4); }
    if (N <= 512) { if (wide) FPS_GO(256, 2); if (shape == 2) FPS_GO(64, 8); if (shape == 3) FPS_GO(32, 16); FPS_GO(128, 4); }
    if (N <= 1024) {
#ifdef PAPC_TRIAGE   // timing experiments with stages of the reduction removed: WRONG RESULTS, triage builds only
        static const int dbg = [] { const char *e = getenv("PAPC_FPS_DBG"); return e ? atoi(e) : 0; }();
#define FPS_DBG(D) if (dbg == D) return launch_fps_reg<128, 8, D>(xyz, B, N, npoint, start_idx, init_dist, out_idx, out_new_xyz, st)
        FPS_DBG(1); FPS_DBG(2); FPS_DBG(3); FPS_DBG(4); FPS_DBG(7);
#undef FPS_DBG
#endif
        if (wide) FPS_GO(256, 4); if (shape == 2) FPS_GO(64, 16); if (shape == 3) FPS_GO(32, 32); FPS_GO(128, 8); }
    if (N <= 2048) { if (wide) FPS_GO(512, 4); FPS_GO(256, 8); }
    if (N <= 4096) { if (wide) FPS_GO(1024, 4); FPS_GO(512, 8); }
    if (N <= kFpsRegMaxN) FPS_GO(1024, 8);
#undef FPS_GO
    const size_t need = papc_fps_workspace_bytes(B, N);
    if (!workspace || workspace_bytes < need) return PAPC_EWORKSPACE;
    fps_global_kernel<1024><<<B, 1024, 0, st>>>(xyz, N, npoint, start_idx, init_dist, out_idx,
                                                 out_new_xyz, reinterpret_cast<float *>(workspace));
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}
