// sa_chain.cu -- chained grouped-MLP layers on the sm_100a tensor cores (layers.py:214-219, 271-276).
//
// The train-mode BatchNorm between two SetAbstraction MLP layers is a grid-wide dependency, so the
// layer-at-a-time kernels (sa_mlp_tt.cu) write every hidden layer's pre-BN output to HBM and read it
// back (134 MB each way at BASELINE config 2).  Here the statistics of a layer come from a
// compute-only pass and the layer is then RECOMPUTED inside the launch that consumes it:
//
//     pass "stats":   in -> [A] -> sum / sum^2 of y_A                                      (nl = 1)
//     pass "chain":   in -> [A] -> BN_A + ReLU -> [B] -> stats (+ max/min pool) of y_B     (nl = 2)
//
// where the activation between A and B never leaves the SM: the accumulator of A comes back from
// tensor memory with lane = channel, column = row (transposed formulation: W is the tcgen05 A operand
// and lives in tensor memory for the whole kernel), each epilogue thread applies its channel's
// BatchNorm scale / shift + ReLU, splits the value into fp16 hi + lo and writes 32 consecutive rows
// of its channel as 64 contiguous bytes of an MN-major SWIZZLE_128B operand -- exactly the B operand
// layer B's tcgen05.mma reads (descriptor: leading offset = stride between 64-row groups, stride
// offset = stride between 8-channel groups; tools/microbench/layout_probe.cu pins both).
//
// Inputs of layer A arrive without conversion work wherever possible:
//   IN_GATHER  : rows of a pre-split fp16 hi/lo image of [feats | xyz] (built once per SOURCE point, 16x
//                fewer conversions than per gathered row) copied by 16-byte cp.async straight into the
//                K-major SWIZZLE_128B operand ring; completion is tracked by the stage's mbarrier
//                (cp.async.mbarrier.arrive.noinc), no wait_group / fence on the loader's path.  The centring
//                xyz[idx] - new_xyz[g] is linear: W_xyz * xyz[idx] goes through the tensor core, the
//                per-group constant -W_xyz * new_xyz[g] (+ bias) is added by the epilogue thread.
//                (TMA tile::gather4 was tried first: issued per lane it serialises through the uniform
//                datapath, ~90 cycles per 512-byte instruction -- 8.8 us per 128-row tile.)
//   IN_TILE    : pre-split activation tiles stored by an earlier chain launch (store_mid), one linear
//                bulk copy per tile, already in the MN-major operand layout;
//   IN_POINTMLP: the cin = 3 first layer folded with its (analytic) BatchNorm, recomputed per row.
// Weights reach tensor memory from a pre-split image (prep_weights: one tiny launch per MLP) with
// independent 16-byte loads -- per-CTA conversion of the fp32 weights cost 20-60 us per launch.
//
// fp32 parity (<= 1e-5) as in sa_mlp_tt.cu: two-term fp16 operand split, three MMAs per product,
// fp32 accumulation in tensor memory, exact power-of-two column scales.  Bias / group constants are
// not added per element: statistics and extrema are taken on the raw accumulator and corrected per
// 32-row block in fp64 (sum (v+c) = S + 32 c, sum (v+c)^2 = Q + c (2 S + 32 c); max (v+c) = max v + c).
//
// Warp roles (one persistent CTA per SM): warps 0-7 epilogue (weight staging, conversion, statistics /
// pooling), then the loader / producer warps, the LAST warp issues the MMAs (highest warp id wins issue
// arbitration) and owns the tensor-memory allocation.  Schedules:
//   nl = 1           two accumulators: layer A of tile i+1 runs while the statistics of tile i are taken;
//   nl = 2, two_acc  accumulators A and B plus a double-buffered B operand: the epilogue converts tile
//                    i+1 while the tensor core runs layer B of tile i, then takes the statistics of tile i
//                    while layer A of tile i+2 runs (period = max(MMA, SIMT));
//   nl = 2, shared   (tensor memory full, e.g. 128 -> 128 -> 256): one accumulator, the layers of a tile
//                    run back to back.
#include "common.cuh"
#include "sa_chain.cuh"
#include "sa_mlp_tt.cuh"
#include "umma.cuh"

#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>

namespace papc {
namespace chain {

using namespace umma;
using tt::f16_colscale_sq;

constexpr int kEpiWarps = 8;
constexpr int kProdWarps = 8;                  // IN_POINTMLP producers
constexpr int kLoadWarps = 4;                  // IN_GATHER cp.async loaders (32 rows each)
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kFirstAuxWarp = kEpiWarps;
template <int IN> struct Roles {
    static constexpr int aux = IN == IN_POINTMLP ? kProdWarps + 1 : IN == IN_GATHER ? kLoadWarps : 1;
    static constexpr int mma_warp = kEpiWarps + aux;
    static constexpr int threads = (mma_warp + 1) * 32;   // 576 / 416 / 320
};

#ifdef PAPC_CHAIN_TRIAGE
// clock64 stamps of CTA 0: per local tile 4..10 at clk[(tl - 4) * 32 + e], kernel phases at clk[224 + e]
#define CH_CLK(a, tl, e)                                                                             \
    do {                                                                                             \
        if ((a).clk != nullptr && blockIdx.x == 0 && (tl) >= 4u && (tl) < 11u)                       \
            (a).clk[((tl) - 4u) * 32 + (e)] = (unsigned long long)clock64();                         \
    } while (0)
#define CH_PH(a, e)                                                                                  \
    do {                                                                                             \
        if ((a).clk != nullptr && blockIdx.x == 0) (a).clk[224 + (e)] = (unsigned long long)clock64(); \
    } while (0)
#else
#define CH_CLK(a, tl, e) do { } while (0)
#define CH_PH(a, e) do { } while (0)
#endif

__device__ __forceinline__ void cp_async16(uint32_t dst_saddr, const void *src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst_saddr), "l"(src) : "memory");
}
// the mbarrier receives this thread's arrival once all its prior cp.async have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t *bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
constexpr int kMaxStages = 6;
constexpr int kChunkBytes = 2 * kTile * 128;   // one K-major chunk stage: [128 rows][128 B] hi + lo
constexpr int kHalfChunk = kTile * 128;
constexpr uint32_t kNG = 1024;                 // MN-major: bytes between the two 64-row groups
constexpr uint32_t kKG = 2048;                 // MN-major: bytes between 8-channel groups
constexpr int kRPT = 4, kRowStride = 32;       // IN_POINTMLP producer mapping (rows rb + 32 j)
constexpr int kMaxKA = 192;
constexpr int kTmemCols = 512;
constexpr int kWChunkWords = 2 * 128 * 32;     // one 64-k chunk of a weight image (hi + lo)

// instruction descriptors: D fp32, A/B f16, A K-major, M = 128, N = 128; B K-major or MN-major
constexpr uint32_t kIdescK = (1u << 4) | ((uint32_t)(kTile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
constexpr uint32_t kIdescMN = kIdescK | (1u << 16);

__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(ta), "l"(db),
                 "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ uint64_t make_desc_mn(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)(kNG >> 4) << 16;   // leading byte offset: next 64-row (N) group
    d |= (uint64_t)(kKG >> 4) << 32;   // stride byte offset: next 8-channel (K) group
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;            // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}
__device__ __forceinline__ uint32_t fastdiv(uint32_t x, uint32_t mul, uint32_t shr) {
    return mul == 0u ? x : (__umulhi(x, mul) >> shr);
}

constexpr int kPtSlots = 4;                    // IN_POINTMLP: tiles of gathered points in flight
struct SmemLayout {
    static constexpr uint32_t xyz = 0;                        // [kPtSlots][2][128] float4: point, centroid (IN_POINTMLP)
    static constexpr uint32_t fold = xyz + kPtSlots * 2 * kTile * 16;    // [128] float4
    static constexpr uint32_t xpool = fold + 128 * 16;        // [2][128] float2 (K = 128 hand-over)
    static constexpr uint32_t xstat = xpool + 2 * kTile * 8;  // [2][128] double2
    static constexpr uint32_t bars = xstat + 2 * kTile * 16;
    static constexpr uint32_t nbars = 2 * kMaxStages + 9 + 2 * kPtSlots;
    static constexpr uint32_t misc = bars + nbars * 8;
    static constexpr uint32_t ring = (misc + 16 + 1023) / 1024 * 1024;
};
constexpr uint32_t kSmemBudget = 227 * 1024 - 1024;  // dynamic shared memory requested (+ alignment slack)

struct Geo {
    int KC;          // K-major chunks per tile (IN_POINTMLP / IN_GATHER)
    int nst;         // ring stages
    uint32_t stage_bytes, abuf_off, abuf_bytes, wa_cols, wb_cols, colW, colB;
};
__host__ __device__ inline Geo make_geo(const ChainArgs &a) {
    Geo g;
    g.KC = (a.ka + 63) / 64;
    g.stage_bytes = a.in_mode == IN_TILE ? (uint32_t)a.ka * 512u : (uint32_t)kChunkBytes;
    g.abuf_bytes = a.nl == 2 ? (uint32_t)a.ca * 512u : 0u;
    const uint32_t nab = a.nl == 2 ? (a.two_acc ? 2u : 1u) : 0u;
    const uint32_t avail = kSmemBudget - SmemLayout::ring - nab * g.abuf_bytes;
    int nst = (int)(avail / g.stage_bytes);
    g.nst = nst > kMaxStages ? kMaxStages : nst;
    g.abuf_off = SmemLayout::ring + (uint32_t)g.nst * g.stage_bytes;
    g.wa_cols = (uint32_t)((a.ka + 63) / 64) * 32u;
    g.wb_cols = (uint32_t)((a.ca + 63) / 64) * 32u;
    g.colW = (a.nl == 1 || a.two_acc) ? 256u : 128u;   // accumulators first
    g.colB = (a.nl == 2 && a.two_acc) ? 128u : 0u;     // accumulator of layer B
    return g;
}
__host__ __device__ inline uint32_t tmem_cols_needed(const ChainArgs &a, const Geo &g) {
    return g.colW + 2 * g.wa_cols + (a.nl == 2 ? (uint32_t)a.nt * 2 * g.wb_cols : 0u);
}

// ---- last-CTA BatchNorm finalisation (fixed-point sums first, partial rows as the fallback)
__device__ void bn_finalize(const ChainArgs &a, int cout, int gm, uint8_t *scratch, int tid, int nthreads) {
    double *all = reinterpret_cast<double *>(scratch);   // [2*cout]
    bool fixed = false;
    if (a.fix_acc != nullptr) {
        fixed = __ldcg(a.fix_acc + (size_t)4 * cout) == 0ull;
        __syncthreads();
        for (int i = tid; i < 2 * cout; i += nthreads) {
            const int which = i / cout, ch = i - which * cout;
            unsigned long long *pi = a.fix_acc + (size_t)(2 * which) * cout + ch;
            unsigned long long *pf = a.fix_acc + (size_t)(2 * which + 1) * cout + ch;
            const long long vi = (long long)__ldcg(pi), vf = (long long)__ldcg(pf);
            if (fixed) all[i] = (double)vi + (double)vf * 0x1p-54;
            *pi = 0ull;
            *pf = 0ull;
        }
        if (tid == 0) a.fix_acc[(size_t)4 * cout] = 0ull;
        __syncthreads();
    }
    if (!fixed) {   // fixed-order reduction of the per-CTA rows [gm][2][cout]
        for (int i = tid; i < 2 * cout; i += nthreads) {
            double s = 0.0;
            for (int r = 0; r < gm; ++r) s += __ldcg(a.stats_partial + (size_t)r * 2 * cout + i);
            all[i] = s;
        }
        __syncthreads();
    }
    for (int ch = tid; ch < cout; ch += nthreads) {
        const double mean = all[ch] * a.inv_count;
        double var = all[cout + ch] * a.inv_count - mean * mean;   // biased, as Paddle's training BN
        var = var > 0.0 ? var : 0.0;
        const double g = a.gamma ? (double)a.gamma[ch] : 1.0;
        const double b = a.beta ? (double)a.beta[ch] : 0.0;
        const double ve = var + (double)a.eps;
        double rs = (double)rsqrtf((float)ve);
        rs = rs * (1.5 - 0.5 * ve * rs * rs);
        rs = rs * (1.5 - 0.5 * ve * rs * rs);
        const double sc = g * rs;
        float cs = 1.f;
        if (a.out_colscale != nullptr) {
            cs = f16_colscale_sq(a.gamma ? a.gamma[ch] : 1.f, a.beta ? a.beta[ch] : 0.f, a.sqrt_count);
            a.out_colscale[ch] = cs;
        }
        a.scale[ch] = (float)sc / cs;
        a.shift[ch] = (float)(b - mean * sc) / cs;
        if (a.mean_out) a.mean_out[ch] = (float)mean;
        if (a.var_out) a.var_out[ch] = (float)var;
    }
    if (tid == 0) *a.counter = 0u;
}

template <int IN, int NL, bool POOL>
__global__ void __launch_bounds__(Roles<IN>::threads, 1)
chain_kernel(const __grid_constant__ ChainArgs a) {
    constexpr int kThreads = Roles<IN>::threads;
    constexpr int kMmaWarp = Roles<IN>::mma_warp;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const uint32_t sm = smem_u32(smem);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SmemLayout::bars);
    uint64_t *in_full = bars;
    uint64_t *in_empty = bars + kMaxStages;
    uint64_t *accA_full = bars + 2 * kMaxStages;   // [2]
    uint64_t *accA_free = accA_full + 2;           // [2]
    uint64_t *accB_full = accA_full + 4;
    uint64_t *accB_free = accA_full + 5;
    uint64_t *abuf_full = accA_full + 6;           // [2]
    uint64_t *w_ready = accA_full + 8;
    uint64_t *pts_full = accA_full + 9;            // [kPtSlots] IN_POINTMLP
    uint64_t *pts_empty = pts_full + kPtSlots;     // [kPtSlots]
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SmemLayout::misc);
    uint32_t *s_last = tmem_slot + 1;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const Geo geo = make_geo(a);
    const int gm = gridDim.x;
    const int mi = blockIdx.x;
    const long long tiles_m = ceil_div<long long>(a.M, kTile);
    const int cout_last = NL == 2 ? a.cb : a.ca;
    const int NT = NL == 2 ? a.nt : 1;
    const bool two_acc = NL == 2 && a.two_acc != 0;
    if (tid == 0) CH_PH(a, 0);

    if (tid == 0) {
        for (int s = 0; s < kMaxStages; ++s) {
            mbar_init(in_full + s, IN == IN_POINTMLP ? kProdWarps : IN == IN_GATHER ? kLoadWarps * 32 : 1);
            mbar_init(in_empty + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(accA_full + b, 1);
            mbar_init(accA_free + b, kEpiWarps);
            mbar_init(abuf_full + b, kEpiWarps);
        }
        mbar_init(accB_full, 1);
        mbar_init(accB_free, kEpiWarps);
        mbar_init(w_ready, kEpiWarps);
        for (int b = 0; b < kPtSlots; ++b) {
            mbar_init(pts_full + b, 32);
            mbar_init(pts_empty + b, kProdWarps);
        }
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) CH_PH(a, 1);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    auto pdl_wait = [&]() {
        if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    };
    auto tile_of = [&](long long t) -> long long { return a.reverse ? tiles_m - 1 - t : t; };
    const long long ntiles_local = mi < tiles_m ? (tiles_m - mi + gm - 1) / gm : 0;

    if (warp < kEpiWarps) {
        // ==================================== epilogue warps ====================================
        const int quad = warp & 3;                // tensor-memory lane quadrant of this warp
        const int half = warp >> 2;               // accumulator columns [64*half, 64*half + 64)
        const int c = quad * 32 + lane;           // channel inside a 128-channel tile == TMEM lane
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        // ---- weights: pre-split image -> tensor memory (the two warps of a quadrant alternate chunks)
        {
            uint32_t cnt = 0;
            auto stage = [&](const uint32_t *img, int chunks, uint32_t col_hi, uint32_t col_lo) {
                for (int kc = 0; kc < chunks; ++kc, ++cnt) {
                    if ((int)(cnt & 1u) != half) continue;
                    const uint4 *ph = reinterpret_cast<const uint4 *>(img + (size_t)kc * kWChunkWords + (size_t)c * 32);
                    const uint4 *pl = ph + 128 * 32 / 4;
                    uint32_t hi[32], lo[32];
#pragma unroll
                    for (int q = 0; q < 8; ++q) {
                        const uint4 h = __ldg(ph + q), l = __ldg(pl + q);
                        hi[4 * q] = h.x; hi[4 * q + 1] = h.y; hi[4 * q + 2] = h.z; hi[4 * q + 3] = h.w;
                        lo[4 * q] = l.x; lo[4 * q + 1] = l.y; lo[4 * q + 2] = l.z; lo[4 * q + 3] = l.w;
                    }
                    tmem_st32(lane_base + col_hi + kc * 32, hi);
                    tmem_st32(lane_base + col_lo + kc * 32, lo);
                }
            };
            const int KCA = (a.ka + 63) / 64, KCB = (a.ca + 63) / 64;
            stage(a.wimgA, KCA, geo.colW, geo.colW + geo.wa_cols);
            if (NL == 2)
                for (int nt = 0; nt < NT; ++nt)
                    stage(a.wimgB + (size_t)nt * KCB * kWChunkWords, KCB, geo.colW + 2 * geo.wa_cols + nt * 2 * geo.wb_cols,
                          geo.colW + 2 * geo.wa_cols + nt * 2 * geo.wb_cols + geo.wb_cols);
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(w_ready);
        }
        if (tid == 0) CH_PH(a, 2);
        pdl_wait();
        if (tid == 0) CH_PH(a, 3);

        // ---- per-thread constants
        const bool cvalidA = c < a.ca;
        const float biasA = (a.biasA != nullptr && cvalidA) ? __ldg(a.biasA + c) : 0.f;
        float scA = 0.f, shA = 0.f;
        if (NL == 2 && cvalidA) {
            scA = __ldg(a.scaleA + c);
            shA = __ldg(a.shiftA + c);
        }
        float wx = 0.f, wy = 0.f, wz = 0.f;   // IN_GATHER: the group constant is biasA - W_xyz . new_xyz[g]
        const bool has_gc = IN == IN_GATHER && a.new_xyz != nullptr && a.wxyz != nullptr;
        if (has_gc && cvalidA) {
            const float *wp = a.wxyz + (size_t)c * a.wxyz_ld;
            wx = __ldg(wp); wy = __ldg(wp + 1); wz = __ldg(wp + 2);
        }
        float biasL0 = 0.f, biasL1 = 0.f;     // bias of the last layer per n-tile
        bool cvalidL0 = c < cout_last, cvalidL1 = NT > 1 && 128 + c < cout_last;
        if (NL == 2) {
            if (a.biasB != nullptr && cvalidL0) biasL0 = __ldg(a.biasB + c);
            if (a.biasB != nullptr && cvalidL1) biasL1 = __ldg(a.biasB + 128 + c);
        } else {
            biasL0 = biasA;
        }
        const int kshift = a.K == 32 ? 5 : a.K == 64 ? 6 : 7;
        float2 *s_xpool = reinterpret_cast<float2 *>(smem + SmemLayout::xpool);
        double acc_s0 = 0.0, acc_q0 = 0.0, acc_s1 = 0.0, acc_q1 = 0.0;
        uint32_t npool = 0;  // K = 128 hand-overs
        const bool stamp = warp == 0 && lane == 0;

        // constants of layer A for the two 32-row blocks of this warp in the tile starting at row m0
        auto group_consts = [&](long long m0, float &g0, float &g1) {
            g0 = g1 = biasA;
            if (!has_gc) return;
            const long long r0 = m0 + half * 64;
            const long long ga = r0 >> kshift, gb = (r0 + 32) >> kshift;
            const long long gmax = (a.M >> kshift) - 1;
            const float *pa = a.new_xyz + (ga < gmax ? ga : gmax) * 3, *pb = a.new_xyz + (gb < gmax ? gb : gmax) * 3;
            const float ax = __ldg(pa), ay = __ldg(pa + 1), az = __ldg(pa + 2);
            const float bx = __ldg(pb), by = __ldg(pb + 1), bz = __ldg(pb + 2);
            g0 = biasA - fmaf(wz, az, fmaf(wy, ay, wx * ax));
            g1 = biasA - fmaf(wz, bz, fmaf(wy, by, wx * bx));
        };
        auto load_acc = [&](uint32_t col, uint32_t (&r0)[32], uint32_t (&r1)[32]) {
            tmem_ld32_nowait(lane_base + col + (uint32_t)(half * 64), r0);
            tmem_ld32_nowait(lane_base + col + (uint32_t)(half * 64 + 32), r1);
            tmem_wait_ld();
            tc_fence_before();
        };

        // statistics (+ pooling) of a completed accumulator; c0 / c1 = the additive constants of its two blocks
        auto stats_pool = [&](const uint32_t (&r0)[32], const uint32_t (&r1)[32], float c0, float c1, int nt, long long m0,
                              int nrows, bool valid, double &acc_s, double &acc_q) {
            float mx = -INFINITY, mn = INFINITY;
            long long pend_g[2] = {-1, -1};
            float pend_mx[2] = {0.f, 0.f}, pend_mn[2] = {0.f, 0.f};
            auto body = [&](const uint32_t (&r)[32], int bi, float cst) {
                const int col0 = half * 64 + bi * 32;
                const int nr = nrows - col0;           // a multiple of 32: the block is whole or absent
                if (nr > 0) {
                    uint64_t s2 = 0ull, q2 = 0ull;
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const uint64_t v2 = pack2u(r[i], r[i + 1]);
                        s2 = add2(s2, v2);
                        q2 = fma2(v2, v2, q2);
                        if (POOL) {
                            const float va = __uint_as_float(r[i]), vb = __uint_as_float(r[i + 1]);
                            mx = fmaxf(fmaxf(mx, va), vb);
                            mn = fminf(fminf(mn, va), vb);
                        }
                    }
                    float sa, sb, qa, qb;
                    unpack2(s2, sa, sb);
                    unpack2(q2, qa, qb);
                    // sum (v + c) = S + 32 c ;  sum (v + c)^2 = Q + c (2 S + 32 c), in fp64
                    const double S = (double)(sa + sb), Q = (double)(qa + qb), cd = (double)cst;
                    acc_s += S + 32.0 * cd;
                    acc_q += Q + cd * (2.0 * S + 32.0 * cd);
                }
                if (POOL && kshift != 7 && (((col0 + 32) & (a.K - 1)) == 0)) {
                    pend_g[bi] = (valid && nr > 0) ? (long long)((m0 + col0) >> kshift) : -1;
                    pend_mx[bi] = mx + cst;            // max (v + c) = max v + c (rounding is monotone)
                    pend_mn[bi] = mn + cst;
                    mx = -INFINITY;
                    mn = INFINITY;
                }
            };
            body(r0, 0, c0);
            body(r1, 1, c1);
            const int cg = nt * 128 + c;
            if (POOL && kshift != 7) {
#pragma unroll
                for (int bi = 0; bi < 2; ++bi)
                    if (pend_g[bi] >= 0) {
                        a.pool_max[pend_g[bi] * cout_last + cg] = pend_mx[bi];
                        a.pool_min[pend_g[bi] * cout_last + cg] = pend_mn[bi];
                    }
            }
            if (POOL && kshift == 7) {   // the tile is one group: the upper half hands over to the lower half
                const uint32_t slot = npool & 1;
                ++npool;
                if (half == 1) s_xpool[slot * kTile + c] = make_float2(mx, mn);
                named_bar_sync(2 + quad, 64);
                if (half == 0) {
                    const float2 o = s_xpool[slot * kTile + c];
                    mx = fmaxf(mx, o.x);
                    mn = fminf(mn, o.y);
                    if (valid && nrows > 0) {
                        const long long g = m0 >> 7;
                        a.pool_max[g * cout_last + cg] = mx + c0;
                        a.pool_min[g * cout_last + cg] = mn + c0;
                    }
                }
            }
        };

        // BatchNorm_A + ReLU + fp16 split of layer A's accumulator -> MN-major operand of layer B (slot of abuf)
        auto convert = [&](const uint32_t (&r0)[32], const uint32_t (&r1)[32], float c0, float c1, uint32_t slot) {
            const uint32_t row_addr = sm + geo.abuf_off + slot * geo.abuf_bytes + (uint32_t)(c >> 3) * kKG + (uint32_t)half * kNG +
                                      (uint32_t)(c & 7) * 128u;
            const uint32_t lo_off = (uint32_t)a.ca * 256u;
            auto body = [&](const uint32_t (&r)[32], int bi, float cst) {
                const float sh = fmaf(cst, scA, shA);
#pragma unroll
                for (int qq = 0; qq < 4; ++qq) {
                    uint32_t hi[4], lo[4];
#pragma unroll
                    for (int p = 0; p < 4; ++p) {
                        const float v0 = fmaxf(fmaf(__uint_as_float(r[qq * 8 + 2 * p]), scA, sh), 0.f);
                        const float v1 = fmaxf(fmaf(__uint_as_float(r[qq * 8 + 2 * p + 1]), scA, sh), 0.f);
                        split_f16x2(v0, v1, hi[p], lo[p]);
                    }
                    const uint32_t chunk = (uint32_t)((bi * 4 + qq) ^ (c & 7)) << 4;
                    if (cvalidA) {
                        sts128(row_addr + chunk, make_uint4(hi[0], hi[1], hi[2], hi[3]));
                        sts128(row_addr + lo_off + chunk, make_uint4(lo[0], lo[1], lo[2], lo[3]));
                    }
                }
            };
            body(r0, 0, c0);
            body(r1, 1, c1);
        };

        if (NL == 1) {
            // ---- statistics of layer A, two accumulators
            uint32_t li = 0;
            for (long long t = mi; t < tiles_m; t += gm, ++li) {
                const long long m0 = tile_of(t) * kTile;
                const int nrows = (int)((a.M - m0) < kTile ? (a.M - m0) : kTile);
                const uint32_t buf = li & 1;
                float g0, g1;
                group_consts(m0, g0, g1);
                mbar_wait(accA_full + buf, (li >> 1) & 1);
                tc_fence_after();
                if (stamp) CH_CLK(a, li, 12);
                uint32_t r0[32], r1[32];
                load_acc(buf * 128, r0, r1);
                __syncwarp();
                if (lane == 0) mbar_arrive(accA_free + buf);
                if (stamp) CH_CLK(a, li, 13);
                if (quad * 32 < cout_last) stats_pool(r0, r1, g0, g1, 0, m0, nrows, cvalidL0, acc_s0, acc_q0);
                if (stamp) CH_CLK(a, li, 14);
            }
        } else {
            uint32_t nconv = 0, nepi = 0;
            auto conv_tile = [&](long long t) {
                const long long m0 = tile_of(t) * kTile;
                float g0, g1;
                group_consts(m0, g0, g1);
                mbar_wait(accA_full, nconv & 1);
                tc_fence_after();
                if (stamp) CH_CLK(a, nconv, 8);
                uint32_t r0[32], r1[32];
                load_acc(0, r0, r1);
                __syncwarp();
                if (lane == 0) mbar_arrive(accA_free);
                if (stamp) CH_CLK(a, nconv, 9);
                const uint32_t slot = two_acc ? (nconv & 1u) : 0u;
                if (quad * 32 < a.ca) convert(r0, r1, g0, g1, slot);   // (warp-uniform) channels of this warp exist
                if (stamp) CH_CLK(a, nconv, 10);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(abuf_full + slot);
                if (stamp) CH_CLK(a, nconv, 11);
                ++nconv;
            };
            auto epi_tile = [&](long long t, uint32_t li) {
                const long long m0 = tile_of(t) * kTile;
                const int nrows = (int)((a.M - m0) < kTile ? (a.M - m0) : kTile);
                for (int nt = 0; nt < NT; ++nt) {
                    mbar_wait(accB_full, nepi & 1);
                    ++nepi;
                    tc_fence_after();
                    if (stamp) CH_CLK(a, li, 12 + 3 * nt);
                    uint32_t r0[32], r1[32];
                    load_acc(geo.colB, r0, r1);
                    __syncwarp();
                    if (lane == 0) mbar_arrive(accB_free);
                    if (stamp) CH_CLK(a, li, 13 + 3 * nt);
                    if (nt * 128 + quad * 32 < cout_last) {
                        if (nt == 0) stats_pool(r0, r1, biasL0, biasL0, 0, m0, nrows, cvalidL0, acc_s0, acc_q0);
                        else stats_pool(r0, r1, biasL1, biasL1, 1, m0, nrows, cvalidL1, acc_s1, acc_q1);
                    }
                    if (stamp) CH_CLK(a, li, 14 + 3 * nt);
                }
            };
            uint32_t li = 0;
            if (two_acc) {
                if (mi < tiles_m) conv_tile(mi);
                for (long long t = mi; t < tiles_m; t += gm, ++li) {
                    if (t + gm < tiles_m) conv_tile(t + gm);
                    epi_tile(t, li);
                }
            } else {
                for (long long t = mi; t < tiles_m; t += gm, ++li) {
                    conv_tile(t);
                    epi_tile(t, li);
                }
            }
        }
        if (tid == 0) CH_PH(a, 5);
        // ---- fold the two halves' statistics, publish the per-CTA sums
        double2 *s_xstat = reinterpret_cast<double2 *>(smem + SmemLayout::xstat);
        if (half == 1) {
            s_xstat[c] = make_double2(acc_s0, acc_q0);
            s_xstat[kTile + c] = make_double2(acc_s1, acc_q1);
        }
        named_bar_sync(6 + quad, 64);
        if (half == 0 && a.stats_partial != nullptr) {
            for (int nt = 0; nt < NT; ++nt) {
                if (!(nt == 0 ? cvalidL0 : cvalidL1)) continue;
                const int cg = nt * 128 + c;
                const double2 o = s_xstat[nt * kTile + c];
                const double vs = (nt == 0 ? acc_s0 : acc_s1) + o.x, vq = (nt == 0 ? acc_q0 : acc_q1) + o.y;
                a.stats_partial[((long long)mi * 2 + 0) * cout_last + cg] = vs;
                a.stats_partial[((long long)mi * 2 + 1) * cout_last + cg] = vq;
                for (long long rr = mi + gm; rr < a.partial_rows; rr += gm) {
                    a.stats_partial[(rr * 2 + 0) * cout_last + cg] = 0.0;
                    a.stats_partial[(rr * 2 + 1) * cout_last + cg] = 0.0;
                }
                if (a.fix_acc != nullptr && a.counter != nullptr) {
                    const double v2[2] = {vs, vq};
#pragma unroll
                    for (int i = 0; i < 2; ++i) {
                        const double p = v2[i];
                        if (fabs(p) < 0x1p53) {
                            const double ip = trunc(p);
                            atomicAdd(a.fix_acc + (size_t)(2 * i) * cout_last + cg, (unsigned long long)(long long)ip);
                            atomicAdd(a.fix_acc + (size_t)(2 * i + 1) * cout_last + cg,
                                      (unsigned long long)__double2ll_rn((p - ip) * 0x1p54));
                        } else {
                            atomicOr(a.fix_acc + (size_t)4 * cout_last, 1ull);
                        }
                    }
                }
            }
            __threadfence();
        }
    } else if (warp == kMmaWarp) {
        // ==================================== MMA issuer ========================================
        const bool leader = lane == 0;   // one thread issues every tcgen05 / bulk-group operation of this warp
        pdl_wait();
        mbar_wait(w_ready, 0);
        tc_fence_after();
        const uint32_t ring = sm + SmemLayout::ring;
        const uint32_t colWA = geo.colW, colWB = geo.colW + 2 * geo.wa_cols;
        uint32_t it = 0;            // ring items consumed (chunks, or tiles for IN_TILE)
        bool store_pending = false;
        // layer A of one tile into accumulator column `dcol`
        auto issue_A = [&](uint32_t dcol, uint32_t li) {
            const uint32_t d_tmem = tmem_base + dcol;
            if (IN == IN_TILE) {
                const uint32_t s = it % geo.nst;
                mbar_wait(in_full + s, (it / geo.nst) & 1);
                tc_fence_after();
                if (leader) CH_CLK(a, li, 0);
                const uint32_t hi = ring + s * geo.stage_bytes, lo = hi + (uint32_t)a.ka * 256u;
                if (leader) {
                    const int nks = a.ka / 16;
                    uint64_t dh = make_desc_mn(hi), dl = make_desc_mn(lo);
                    uint32_t w_hi = tmem_base + colWA;
                    for (int ks = 0; ks < nks; ++ks, dh += (2 * kKG) >> 4, dl += (2 * kKG) >> 4, w_hi += 8) {
                        mma_ts(d_tmem, w_hi + geo.wa_cols, dh, kIdescMN, ks != 0);
                        mma_ts(d_tmem, w_hi, dl, kIdescMN, 1);
                        mma_ts(d_tmem, w_hi, dh, kIdescMN, 1);
                    }
                    mma_commit(in_empty + s);
                }
                __syncwarp();
                ++it;
            } else {
                for (int cc = 0; cc < geo.KC; ++cc, ++it) {
                    const uint32_t s = it % geo.nst;
                    mbar_wait(in_full + s, (it / geo.nst) & 1);
                    tc_fence_after();
                    if (leader && cc == 0) CH_CLK(a, li, 0);
                    const uint32_t hi = ring + s * kChunkBytes, lo = hi + kHalfChunk;
                    if (leader) {
                        const int left = a.ka - cc * 64;
                        const int nks = (left < 64 ? left : 64) / 16;
                        const uint64_t dh = make_desc_sw128(hi), dl = make_desc_sw128(lo);
                        for (int ks = 0; ks < nks; ++ks) {
                            const uint32_t w_hi = tmem_base + colWA + cc * 32 + ks * 8, w_lo = w_hi + geo.wa_cols;
                            mma_ts(d_tmem, w_lo, dh + ks * 2, kIdescK, (cc | ks) != 0);
                            mma_ts(d_tmem, w_hi, dl + ks * 2, kIdescK, 1);
                            mma_ts(d_tmem, w_hi, dh + ks * 2, kIdescK, 1);
                        }
                        mma_commit(in_empty + s);
                    }
                    __syncwarp();
                }
            }
        };
        // layer B of one tile (n-tile nt) from operand slot `slot` into the B accumulator
        auto issue_B = [&](int nt, uint32_t slot) {
            if (leader) {
                const uint32_t abuf = sm + geo.abuf_off + slot * geo.abuf_bytes;
                const int nks = a.ca / 16;
                const uint32_t lo_off = (uint32_t)a.ca * 256u;
                const uint32_t wb = tmem_base + colWB + nt * 2 * geo.wb_cols;
                const uint32_t d_tmem = tmem_base + geo.colB;
                uint64_t dh = make_desc_mn(abuf), dl = make_desc_mn(abuf + lo_off);
                uint32_t w_hi = wb;
                for (int ks = 0; ks < nks; ++ks, dh += (2 * kKG) >> 4, dl += (2 * kKG) >> 4, w_hi += 8) {
                    mma_ts(d_tmem, w_hi + geo.wb_cols, dh, kIdescMN, ks != 0);
                    mma_ts(d_tmem, w_hi, dl, kIdescMN, 1);
                    mma_ts(d_tmem, w_hi, dh, kIdescMN, 1);
                }
                mma_commit(accB_full);
            }
            __syncwarp();
        };
        auto store_mid = [&](long long t, uint32_t slot) {
            if (leader) {
                uint8_t *dst = a.mid_out + (size_t)tile_of(t) * ((size_t)a.ca * 512u);
                const uint32_t abuf = sm + geo.abuf_off + slot * geo.abuf_bytes;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(abuf),
                             "r"(geo.abuf_bytes) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            __syncwarp();
            store_pending = true;
        };
        auto commit_A = [&](uint64_t *bar, uint32_t li) {   // layer A of a tile is complete once this fires
            if (leader) {
                // the previous stores have finished READING their operand slot before the conversion that this
                // commit releases may overwrite it
                if (store_pending) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                mma_commit(bar);
                CH_CLK(a, li, 1);
            }
            __syncwarp();
            store_pending = false;
        };

        if (NL == 1) {
            uint32_t li = 0;
            for (long long t = mi; t < tiles_m; t += gm, ++li) {
                const uint32_t buf = li & 1;
                mbar_wait(accA_free + buf, ((li >> 1) & 1) ^ 1);
                tc_fence_after();
                issue_A(buf * 128, li);
                commit_A(accA_full + buf, li);
            }
        } else if (two_acc) {
            uint32_t li = 0, nB = 0;
            if (mi < tiles_m) {
                issue_A(0, 0);
                commit_A(accA_full, 0);
            }
            for (long long t = mi; t < tiles_m; t += gm, ++li) {
                // layer A of the next tile as soon as the conversion of this one has taken the accumulator
                mbar_wait(accA_free, li & 1);
                tc_fence_after();
                if (t + gm < tiles_m) {
                    issue_A(0, li + 1);
                    commit_A(accA_full, li + 1);
                }
                const uint32_t slot = li & 1;
                mbar_wait(abuf_full + slot, (li >> 1) & 1);
                tc_fence_after();
                if (leader) CH_CLK(a, li, 2);
                if (a.store_mid) store_mid(t, slot);
                for (int nt = 0; nt < NT; ++nt, ++nB) {
                    mbar_wait(accB_free, (nB & 1) ^ 1);   // the previous statistics step has loaded accumulator B
                    tc_fence_after();
                    if (leader) CH_CLK(a, li, 4 + 2 * nt);
                    issue_B(nt, slot);
                    if (leader) CH_CLK(a, li, 3 + 2 * nt);
                }
            }
        } else {
            uint32_t li = 0, nB = 0;
            for (long long t = mi; t < tiles_m; t += gm, ++li) {
                issue_A(0, li);
                commit_A(accA_full, li);
                mbar_wait(abuf_full, li & 1);
                tc_fence_after();
                if (leader) CH_CLK(a, li, 2);
                if (a.store_mid) store_mid(t, 0);
                for (int nt = 0; nt < NT; ++nt, ++nB) {
                    issue_B(nt, 0);
                    if (leader) CH_CLK(a, li, 3 + 2 * nt);
                    mbar_wait(accB_free, nB & 1);         // one accumulator: wait until it is in registers
                    tc_fence_after();
                    if (leader) CH_CLK(a, li, 4 + 2 * nt);
                }
            }
        }
        if (store_pending && leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        __syncwarp();
    } else if (IN == IN_TILE) {
        // ==================================== tile loader =======================================
        pdl_wait();
        uint32_t it = 0;
        const uint32_t half_bytes = (uint32_t)a.ka * 256u;
        for (long long t = mi; t < tiles_m; t += gm, ++it) {
            const uint32_t s = it % geo.nst;
            mbar_wait(in_empty + s, ((it / geo.nst) & 1) ^ 1);
            if (lane == 0) {
                CH_CLK(a, it, 20);
                mbar_expect_tx(in_full + s, 2 * half_bytes);
                const uint8_t *src = a.mid_in + (size_t)tile_of(t) * ((size_t)a.ka * 512u);
                uint8_t *dst = smem + SmemLayout::ring + s * geo.stage_bytes;
                bulk_g2s(dst, src, half_bytes, in_full + s);
                bulk_g2s(dst + half_bytes, src + half_bytes, half_bytes, in_full + s);
            }
            __syncwarp();
        }
    } else if (IN == IN_GATHER) {
        // ==================================== gather loaders ====================================
        // Loader warp lw copies rows [32 lw, 32 lw + 32) of every chunk stage: 16-byte cp.async straight into
        // the K-major SWIZZLE_128B operand position (the image is already fp16 hi / lo), lane = (row r4 of 4,
        // 16-byte unit q of 8).  Every thread then hands its arrival to the stage's mbarrier
        // (cp.async.mbarrier.arrive.noinc fires when its copies have landed): nothing on this path waits, so
        // the depth of the ring is the number of stages in flight.
        const int lw = warp - kFirstAuxWarp;
        const int q = lane & 7, r4 = lane >> 3;
        pdl_wait();
        const size_t pitch = (size_t)a.img_ld * 2;
        const size_t lo_off = (size_t)a.img_rows * pitch;
        long long srow[8];
        auto fetch_rows = [&](long long tile) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                long long row = tile * kTile + 32 * lw + 4 * j + r4;
                row = row < a.M ? row : a.M - 1;
                const uint32_t rw = (uint32_t)row;
                const uint32_t gg = fastdiv(rw, a.kmul, a.kshr);
                const uint32_t b = fastdiv(gg, a.smul, a.sshr);
                int n = a.idx != nullptr ? __ldg(a.idx + row) : (int)(rw - gg * (uint32_t)a.K);
                n = min(max(n, 0), a.N - 1);
                srow[j] = ((long long)b * a.N + n) * (long long)pitch;
            }
        };
        uint32_t it = 0, ltl = 0;
        if (mi < tiles_m) fetch_rows(tile_of(mi));
        for (long long t = mi; t < tiles_m; t += gm, ++ltl) {
            long long cur[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) cur[j] = srow[j];
            if (t + gm < tiles_m) fetch_rows(tile_of(t + gm));
            for (int cc = 0; cc < geo.KC; ++cc, ++it) {
                const uint32_t s = it % geo.nst;
                mbar_wait(in_empty + s, ((it / geo.nst) & 1) ^ 1);
                if (lw == 0 && lane == 0 && cc == 0) CH_CLK(a, ltl, 20);
                const int units = min(8, (a.img_ld - cc * 64) / 8);   // valid 16-byte units of this chunk
                if (q < units) {
                    const uint32_t stage = sm + SmemLayout::ring + s * kChunkBytes;
                    const uint8_t *src0 = a.image + (size_t)cc * 128 + (size_t)q * 16;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int rl = 32 * lw + 4 * j + r4;
                        const uint32_t dst = stage + (uint32_t)rl * 128u + (uint32_t)((q ^ (rl & 7)) << 4);
                        cp_async16(dst, src0 + cur[j]);
                        cp_async16(dst + kHalfChunk, src0 + cur[j] + lo_off);
                    }
                }
                cp_async_arrive_noinc(in_full + s);
                if (lw == 0 && lane == 0 && cc == geo.KC - 1) CH_CLK(a, ltl, 21);
            }
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (IN == IN_POINTMLP && warp == kFirstAuxWarp + kProdWarps) {
        // ==================================== point loader ======================================
        // One warp gathers the tile's 128 points and centroids (lane = 4 consecutive rows) with 4-byte cp.async
        // straight into a shared-memory table, kPtSlots tiles ahead; the stage's mbarrier collects the copies
        // (cp.async.mbarrier.arrive.noinc).  The producer warps therefore never have a global load in flight:
        // with the gathers issued from the producers themselves, the shared-memory loads that followed waited
        // for the same scoreboard slots and every tile cost a full global round trip (2 800 of 3 100 cycles).
        pdl_wait();
        uint32_t ptl = 0;
        for (long long t = mi; t < tiles_m; t += gm, ++ptl) {
            const uint32_t slot = ptl % kPtSlots;
            mbar_wait(pts_empty + slot, ((ptl / kPtSlots) & 1) ^ 1);
            const long long m0 = tile_of(t) * kTile + 4 * lane;
            int nidx[4];
            if (a.idx != nullptr) {
                const long long r0 = m0 < a.M ? m0 : a.M - 4;     // M is a multiple of 32
                const int4 v = __ldg(reinterpret_cast<const int4 *>(a.idx + r0));
                nidx[0] = v.x; nidx[1] = v.y; nidx[2] = v.z; nidx[3] = v.w;
            }
            const uint32_t tab = sm + SmemLayout::xyz + slot * (2 * kTile * 16);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                long long row = m0 + j;
                row = row < a.M ? row : a.M - 1;
                const uint32_t rw = (uint32_t)row;
                const uint32_t gg = fastdiv(rw, a.kmul, a.kshr);
                const long long bN = (long long)fastdiv(gg, a.smul, a.sshr) * a.N;
                int n = a.idx != nullptr ? nidx[j] : (int)(rw - gg * (uint32_t)a.K);
                n = min(max(n, 0), a.N - 1);
                const float *qp = a.xyz + (bN + n) * 3;
                const float *cp = a.new_xyz != nullptr ? a.new_xyz + (long long)gg * 3 : qp;
                const uint32_t dp = tab + 16u * (uint32_t)(4 * lane + j), dc = dp + kTile * 16;
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dp + 4 * d), "l"(qp + d) : "memory");
                    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dc + 4 * d), "l"(cp + d) : "memory");
                }
            }
            cp_async_arrive_noinc(pts_full + slot);
        }
        asm volatile("cp.async.wait_all;" ::: "memory");
    } else if (IN == IN_POINTMLP) {
        // ==================================== producers =========================================
        // relu(bn(W0 p + b0)) for the tile's 128 rows from the folded first layer: thread (rb, u) computes the
        // 8 channels of 16-byte unit u for rows rb + 32 j from the point table the loader warp fills.
        const int ptid = tid - kFirstAuxWarp * 32;   // 0..255
        const int u = ptid & 7;
        const int rb = ptid >> 3;
        const uint32_t swz = (uint32_t)((u ^ (rb & 7)) << 4);
        float4 *s_fold = reinterpret_cast<float4 *>(smem + SmemLayout::fold);
        pdl_wait();
        if (ptid < 128) s_fold[ptid] = ptid < a.ka ? reinterpret_cast<const float4 *>(a.l0_fold)[ptid] : make_float4(0.f, 0.f, 0.f, 0.f);
        named_bar_sync(1, kProdThreads);
        const bool centred = a.new_xyz != nullptr;
        uint32_t it = 0, ptl = 0;
        for (long long t = mi; t < tiles_m; t += gm, ++ptl) {
            const uint32_t slot = ptl % kPtSlots;
            const uint32_t tab = sm + SmemLayout::xyz + slot * (2 * kTile * 16);
            if (ptid == 0) CH_CLK(a, ptl, 22);
            mbar_wait(pts_full + slot, (ptl / kPtSlots) & 1);
            if (ptid == 0) CH_CLK(a, ptl, 23);
            float4 p_cur[kRPT];
#pragma unroll
            for (int j = 0; j < kRPT; ++j) {
                const float4 pp = lds128f(tab + 16u * (uint32_t)(rb + kRowStride * j));
                const float4 pc = lds128f(tab + kTile * 16 + 16u * (uint32_t)(rb + kRowStride * j));
                p_cur[j] = centred ? make_float4(__fsub_rn(pp.x, pc.x), __fsub_rn(pp.y, pc.y), __fsub_rn(pp.z, pc.z), 0.f) : pp;
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(pts_empty + slot);
            for (int cc = 0; cc < geo.KC; ++cc, ++it) {
                const uint32_t s = it % geo.nst;
                const int k0 = cc * 64 + u * 8;
                float o[kRPT][8];
                float4 fo[8];   // all eight loads in flight before the first FMA: one exposed shared-memory latency, not eight
#pragma unroll
                for (int qq = 0; qq < 8; ++qq) fo[qq] = lds128f(sm + SmemLayout::fold + 16 * ((k0 + qq) & 127));
#pragma unroll
                for (int qq = 0; qq < 8; ++qq) {
                    const float4 f = fo[qq];
#pragma unroll
                    for (int j = 0; j < kRPT; ++j) {
                        const float4 p = p_cur[j];
                        o[j][qq] = fmaxf(fmaf(f.x, p.x, fmaf(f.y, p.y, fmaf(f.z, p.z, f.w))), 0.f);
                    }
                }
                if (ptid == 0 && cc == 0) CH_CLK(a, ptl, 24);
                mbar_wait(in_empty + s, ((it / geo.nst) & 1) ^ 1);
                if (ptid == 0 && cc == 0) CH_CLK(a, ptl, 20);
                const uint32_t stage = sm + SmemLayout::ring + s * kChunkBytes;
#pragma unroll
                for (int j = 0; j < kRPT; ++j) {
                    uint4 hi, lo;
                    split_f16x2(o[j][0], o[j][1], hi.x, lo.x);
                    split_f16x2(o[j][2], o[j][3], hi.y, lo.y);
                    split_f16x2(o[j][4], o[j][5], hi.z, lo.z);
                    split_f16x2(o[j][6], o[j][7], hi.w, lo.w);
                    const uint32_t off = (uint32_t)(rb + kRowStride * j) * 128u + swz;
                    sts128(stage + off, hi);
                    sts128(stage + kHalfChunk + off, lo);
                }
                if (ptid == 0 && cc == geo.KC - 1) CH_CLK(a, ptl, 25);
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_arrive(in_full + s);
                if (ptid == 0 && cc == geo.KC - 1) CH_CLK(a, ptl, 21);
            }
        }
    }
    (void)ntiles_local;

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) CH_PH(a, 6);
    if (warp == kMmaWarp) tmem_dealloc<kTmemCols>(tmem_base);
    if (a.counter != nullptr) {
        if (tid == 0) {
            __threadfence();
            *s_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
        }
        __syncthreads();
        if (*s_last != 0u) {
            __threadfence();
            bn_finalize(a, cout_last, gm, smem + SmemLayout::ring, tid, kThreads);
        }
    }
    if (tid == 0) CH_PH(a, 7);
}

// =============================================================================== weight images
struct WPrepArgs {
    WeightSpec spec[4];
    int first_block[5];   // spec i owns blocks [first_block[i], first_block[i+1])
    int n;
};
__global__ void __launch_bounds__(1024)
weight_image_kernel(const WPrepArgs p) {
    int si = 0;
    while (si + 1 < p.n && (int)blockIdx.x >= p.first_block[si + 1]) ++si;
    const WeightSpec &w = p.spec[si];
    const int KC = (w.K + 63) / 64;
    const int blk = blockIdx.x - p.first_block[si];   // = nt * KC + kc
    const int nt = blk / KC, kc = blk - nt * KC;
    // thread = (word group g of 8, channel c): consecutive threads read consecutive k of one row (coalesced)
    const int kk = threadIdx.x & 63;                  // k inside the chunk
    const int cb = threadIdx.x >> 6;                  // channels cb, cb + 16, ...
    __shared__ float s_v[128][65];
    for (int c = cb; c < 128; c += 16) {
        const int cg = nt * 128 + c;
        const int k = kc * 64 + kk;
        float x = 0.f;
        if (cg < w.rows && k < w.K) {
            int col = -1;
            if (k < w.nk) col = w.k0 + k;
            else if (w.xyz >= 0 && k < w.nk + 6) col = w.xyz + (k - w.nk) % 3;
            if (col >= 0) {
                float cs = 1.f;
                if (w.cs_on) cs = f16_colscale_sq(w.cs_gamma ? w.cs_gamma[k] : 1.f, w.cs_beta ? w.cs_beta[k] : 0.f, w.cs_sqrt_count);
                else if (w.colscale != nullptr) cs = w.colscale[k];
                x = w.W[(size_t)cg * w.ld + col] * cs;
            }
        }
        s_v[c][kk] = x;
    }
    __syncthreads();
    // word j of channel c = fp16 pair (k = 2j, 2j+1): thread -> (c = tid / 8, words 4 (tid % 8) .. +3), 16-byte stores
    const int c = threadIdx.x >> 3, j0 = (threadIdx.x & 7) * 4;
    uint32_t hi[4], lo[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) split_f16x2(s_v[c][2 * (j0 + q)], s_v[c][2 * (j0 + q) + 1], hi[q], lo[q]);
    uint32_t *out_hi = w.image + (size_t)blk * kWChunkWords + (size_t)c * 32 + j0;
    *reinterpret_cast<uint4 *>(out_hi) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
    *reinterpret_cast<uint4 *>(out_hi + 128 * 32) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
}

size_t weight_image_bytes(int rows, int K) {
    return (size_t)((rows + 127) / 128) * ((K + 63) / 64) * kWChunkWords * sizeof(uint32_t);
}

int prep_weights(const WeightSpec *specs, int n, cudaStream_t st) {
    if (n < 1 || n > 4) return PAPC_EINVAL;
    WPrepArgs p{};
    p.n = n;
    int blocks = 0;
    for (int i = 0; i < n; ++i) {
        if (!specs[i].W || !specs[i].image || specs[i].rows < 1 || specs[i].K < 16 || specs[i].K % 16 != 0) return PAPC_EINVAL;
        p.spec[i] = specs[i];
        p.first_block[i] = blocks;
        blocks += ((specs[i].rows + 127) / 128) * ((specs[i].K + 63) / 64);
    }
    p.first_block[n] = blocks;
    ProfScope prof(st, "chain_weights", blocks, 0, 0, 0.0, (double)blocks * kWChunkWords * 4.0);
    weight_image_kernel<<<blocks, 1024, 0, st>>>(p);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

// =============================================================================== source image
constexpr int kImgThreads = 256;

__global__ void __launch_bounds__(kImgThreads)
image_absmax_kernel(const float *__restrict__ feats, const float *__restrict__ xyz, long long R, int D,
                    unsigned int *__restrict__ absmax) {
    // column-wise max |value| over a slice of rows; columns [0,D) feats, [D,D+3) xyz
    const int C = D + 3;
    for (int col = threadIdx.x; col < C; col += kImgThreads) {
        float m = 0.f;
        for (long long r = blockIdx.x; r < R; r += gridDim.x) {
            const float v = col < D ? feats[r * D + col] : xyz[r * 3 + (col - D)];
            m = fmaxf(m, fabsf(v));
        }
        atomicMax(absmax + col, __float_as_uint(m));   // non-negative floats order as their bit patterns
    }
}

__device__ __forceinline__ float pow2_scale(float absmax) {
    float s = 1.f;
    while (s * 32000.f < absmax && s < 1e30f) s *= 2.f;
    return s;
}

__global__ void __launch_bounds__(kImgThreads)
image_build_kernel(const float *__restrict__ feats, const float *__restrict__ xyz, long long R, int D, int ld,
                   const unsigned int *__restrict__ absmax, __half *__restrict__ img, float *__restrict__ colscale) {
    const long long total = R * ld;
    __half *hi = img, *lo = img + total;
    if (blockIdx.x == 0)
        for (int k = threadIdx.x; k < D + 6; k += kImgThreads) {
            const int src = k < D + 3 ? k : k - 3;
            colscale[k] = pow2_scale(__uint_as_float(absmax[src]));
        }
    for (long long e = (long long)blockIdx.x * kImgThreads + threadIdx.x; e < total; e += (long long)gridDim.x * kImgThreads) {
        const long long r = e / ld;
        const int k = (int)(e - r * ld);
        float h = 0.f, l = 0.f;
        if (k < D) {
            const float v = feats[r * D + k] / pow2_scale(__uint_as_float(absmax[k]));
            h = __half2float(__float2half_rn(v));
            l = v - h;
        } else if (k < D + 6) {
            const int d = (k - D) % 3;
            const float v = xyz[r * 3 + d] / pow2_scale(__uint_as_float(absmax[D + d]));
            const float t0 = __half2float(__float2half_rn(v));
            const float r1 = v - t0;
            const float t1 = __half2float(__float2half_rn(r1));
            if (k < D + 3) { h = t0; l = t1; }
            else { h = r1 - t1; l = 0.f; }
        }
        hi[e] = __float2half_rn(h);
        lo[e] = __float2half_rn(l);
    }
}

int image_ld(int D) { return (D + 6 + 15) / 16 * 16; }
size_t image_bytes(long long R, int D) { return (size_t)2 * R * image_ld(D) * sizeof(__half); }

int build_image(const float *feats, const float *xyz, long long R, int D, uint8_t *image, float *colscale,
                unsigned int *absmax, cudaStream_t st) {
    if (R <= 0 || D < 0 || !xyz || !image || !colscale || !absmax || (D > 0 && !feats)) return PAPC_EINVAL;
    PAPC_CUDA_TRY(cudaMemsetAsync(absmax, 0, sizeof(unsigned int) * (D + 3), st));
    long long rows_blocks = R < 4LL * kNumSMs ? R : 4LL * kNumSMs;
    image_absmax_kernel<<<(unsigned)rows_blocks, kImgThreads, 0, st>>>(feats, xyz, R, D, absmax);
    PAPC_LAUNCH_CHECK();
    const int ld = image_ld(D);
    long long blocks = (R * ld + kImgThreads - 1) / kImgThreads;
    if (blocks > 8LL * kNumSMs) blocks = 8LL * kNumSMs;
    ProfScope prof(st, "chain_image", R, D, ld, 0.0, (double)R * (4.0 * (D + 3) + 4.0 * ld));
    image_build_kernel<<<(unsigned)blocks, kImgThreads, 0, st>>>(feats, xyz, R, D, ld, absmax, reinterpret_cast<__half *>(image),
                                                                colscale);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

// ==================================================================================== host side
static bool shape_ok(const ChainArgs &a) {
    if (a.nl != 1 && a.nl != 2) return false;
    if (a.M <= 0 || a.M >= (1LL << 31)) return false;
    if (!(a.K == 32 || a.K == 64 || a.K == 128) || a.M % a.K != 0) return false;
    if (a.ka < 16 || a.ka % 16 != 0 || a.ka > kMaxKA) return false;
    if (a.ca < 1 || a.ca > 128) return false;
    if (a.in_mode == IN_POINTMLP && a.ka > 128) return false;
    if (a.in_mode == IN_TILE && (a.ka > 128 || a.nl != 2)) return false;
    if (a.nl == 2) {
        if (a.ca % 16 != 0) return false;
        if (a.nt < 1 || a.nt > 2 || a.cb < 1 || a.cb > 128 * a.nt) return false;
    }
    return true;
}
static bool fits(const ChainArgs &a) {
    const Geo g = make_geo(a);
    const int min_stages = a.in_mode == IN_TILE ? 2 : 3;
    return g.nst >= min_stages && tmem_cols_needed(a, g) <= (uint32_t)kTmemCols;
}
// chooses two_acc; false if the shape does not fit at all
static bool plan(ChainArgs &a) {
    if (!shape_ok(a)) return false;
    if (a.nl == 2) {
        a.two_acc = 1;
        if (fits(a)) return true;
    }
    a.two_acc = 0;
    return fits(a);
}
bool eligible(const ChainArgs &a_in) {
    ChainArgs a = a_in;
    return plan(a);
}

template <int IN, int NL, bool POOL>
static int launch_inst(const ChainArgs &a, int grid, cudaStream_t st) {
    auto k = chain_kernel<IN, NL, POOL>;
    constexpr int threads = Roles<IN>::threads;
    const int smem = (int)kSmemBudget + 1024;
    static bool configured = false;
    if (!configured) {
        PAPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        configured = true;
    }
    char name[56];
    snprintf(name, sizeof(name), "mlp_chain<%s,%s%s%s%s>", IN == IN_POINTMLP ? "pointmlp" : IN == IN_GATHER ? "gather" : "tile",
             NL == 2 ? "A+B" : "A", POOL ? ",pool" : ",stats", a.store_mid ? ",store" : "", (NL == 2 && !a.two_acc) ? ",1acc" : "");
    ProfScope prof(st, name, a.M, a.prof_cin, a.prof_cout, a.prof_flops, a.prof_bytes);
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3(threads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = a.pdl ? 1 : 0;
    ++g_launch_count;
#ifdef PAPC_CHAIN_TRIAGE
    static unsigned long long *d_clk = nullptr;
    ChainArgs b = a;
    if (getenv("PAPC_CHAIN_CLK") != nullptr) {
        if (d_clk == nullptr) cudaMalloc(&d_clk, 256 * sizeof(unsigned long long));
        cudaMemsetAsync(d_clk, 0, 256 * sizeof(unsigned long long), st);
        b.clk = d_clk;
        b.pdl = 0;
        cfg.numAttrs = 0;
    }
    PAPC_CUDA_TRY(cudaLaunchKernelEx(&cfg, k, b));
    if (b.clk != nullptr) {
        unsigned long long h[256];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
        const unsigned long long *ph = h + 224;
        auto pu = [&](int e) { return ph[e] ? (double)(ph[e] - ph[0]) / 1965.0 : -1.0; };
        fprintf(stderr, "[chain clk] %s M %lld ka %d ca %d cb %d nt %d two_acc %d | CTA0 us: setup %.2f Wstaged %.2f pdl %.2f tiles_done %.2f "
                "teardown %.2f exit %.2f\n", name, a.M, a.ka, a.ca, a.cb, a.nt, a.two_acc, pu(1), pu(2), pu(3), pu(5), pu(6), pu(7));
        fprintf(stderr, "[chain clk] tile | mma: LAdone abuf LB0 free0 LB1 free1 | epi: cvAcc cvLd cvSts cvArr e0Acc e0Ld e0End e1Acc e1Ld e1End | load: slot issued | next LA\n");
        for (int t = 0; t < 6; ++t) {
            const unsigned long long *q = h + t * 32;
            if (q[0] == 0) continue;
            auto d = [&](int e) { return q[e] ? (long long)(q[e] - q[0]) : -1LL; };
            fprintf(stderr, "[chain clk] %2d | %6lld %6lld %6lld %6lld %6lld %6lld | %6lld %6lld %6lld %6lld %6lld %6lld %6lld %6lld %6lld %6lld | %6lld %6lld | %6lld\n",
                    t + 4, d(1), d(2), d(3), d(4), d(5), d(6), d(8), d(9), d(10), d(11), d(12), d(13), d(14), d(15), d(16), d(17), d(20),
                    d(21), h[(t + 1) * 32] ? (long long)(h[(t + 1) * 32] - q[0]) : -1LL);
            if (q[22]) fprintf(stderr, "[chain clk]      producer (rel. its tile start): bar %lld computed %lld slot %lld sts %lld published %lld | next tile %lld\n",
                    (long long)(q[23] - q[22]), (long long)(q[24] - q[22]), (long long)(q[20] - q[22]), (long long)(q[25] - q[22]),
                    (long long)(q[21] - q[22]), h[(t + 1) * 32 + 22] ? (long long)(h[(t + 1) * 32 + 22] - q[22]) : -1LL);
        }
    }
    return PAPC_OK;
#else
    PAPC_CUDA_TRY(cudaLaunchKernelEx(&cfg, k, a));
    return PAPC_OK;
#endif
}

int launch(const ChainArgs &a_in, cudaStream_t st) {
    ChainArgs a = a_in;
    if (!plan(a)) return PAPC_EUNSUPPORTED;
    if (!a.wimgA || (a.nl == 2 && !a.wimgB)) return PAPC_EINVAL;
    {
        const char *e = getenv("PAPC_CHAIN_2ACC");   // A/B switch: PAPC_CHAIN_2ACC=0 forces the shared accumulator
        if (e && e[0] == '0') a.two_acc = 0;
    }
    tt::make_fastdiv((uint32_t)a.K, &a.kmul, &a.kshr);
    tt::make_fastdiv((uint32_t)(a.S >= 1 ? a.S : 1), &a.smul, &a.sshr);
    a.inv_count = a.count > 0.0 ? 1.0 / a.count : 0.0;
    {
        const char *e = getenv("PAPC_TT_PDL");
        if (e && e[0] == '0') a.pdl = 0;
    }
    const long long tiles_m = ceil_div<long long>(a.M, kTile);
    long long gm = kNumSMs;
    if (gm > tiles_m) gm = tiles_m;
    if (a.stats_partial != nullptr && gm > a.partial_rows) gm = a.partial_rows;
    const int grid = (int)gm;
    const bool pool = a.pool != 0;
    if (pool && (!a.pool_max || !a.pool_min)) return PAPC_EINVAL;
    switch (a.in_mode) {
        case IN_POINTMLP:
            if (a.nl == 1) return pool ? launch_inst<IN_POINTMLP, 1, true>(a, grid, st) : launch_inst<IN_POINTMLP, 1, false>(a, grid, st);
            return pool ? launch_inst<IN_POINTMLP, 2, true>(a, grid, st) : launch_inst<IN_POINTMLP, 2, false>(a, grid, st);
        case IN_GATHER:
            if (a.nl == 1) return pool ? launch_inst<IN_GATHER, 1, true>(a, grid, st) : launch_inst<IN_GATHER, 1, false>(a, grid, st);
            return pool ? launch_inst<IN_GATHER, 2, true>(a, grid, st) : launch_inst<IN_GATHER, 2, false>(a, grid, st);
        case IN_TILE:
            return pool ? launch_inst<IN_TILE, 2, true>(a, grid, st) : launch_inst<IN_TILE, 2, false>(a, grid, st);
    }
    return PAPC_EINVAL;
}

}  // namespace chain
}  // namespace papc
