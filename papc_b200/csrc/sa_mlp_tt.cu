// sa_mlp_tt.cu -- grouped shared-MLP layer on the sm_100a tensor cores, transposed formulation:
//
//      Y^T[cout, rows] = W[cout, cin] * act(X)^T[cin, rows]        (one SetAbstraction MLP layer,
//                                                                   layers.py:214-219 / 271-276)
//
// The M = B*S*K grouped rows are the MMA *N* dimension and the output channels the MMA *M*
// dimension, so that
//   * W (hi/lo split) is the A operand and -- whenever it fits -- lives in TENSOR MEMORY for the
//     whole kernel: the tensor core then reads only the activation tile from shared memory (half
//     the operand traffic of a shared/shared UMMA, which at M=N=128 saturates the 128 B/clk
//     shared-memory port).  For longer reductions W chunks are streamed through the same ring by
//     TMA from a pre-split, pre-swizzled image (shared/shared UMMA);
//   * the accumulator comes back with lane = channel, column = row: every epilogue thread owns one
//     channel, so bias, BatchNorm sum / sum^2 and the per-group max/min pooling are plain
//     per-thread running reductions (no shuffles), and a warp stores 32 consecutive channels of one
//     row (128-byte coalesced) straight from the tcgen05.ld registers.
//
// fp32 in / fp32 out inside the 1e-5 parity budget through a two-term operand split and three MMAs
// per product (lo*hi + hi*lo + hi*hi, fp32 accumulation in tensor memory):
//   PREC_TF32: hi/lo are TF32, K = 8 per MMA -- any finite input;
//   PREC_F16 : hi/lo are fp16, K = 16 per MMA (half the MMAs, half the shared-memory bytes) --
//              only for inputs that are relu(batch-norm(.)), whose magnitude is bounded by
//              |gamma| sqrt(count) + |beta|; an exact power-of-two column scale keeps them < 2^15.
//
// Persistent, one CTA per SM, 21 warps:
//   warps 0-3   epilogue  (TMEM lane quadrant = warp).  They first stage W into tensor memory.
//   warp  20    MMA issuer (one elected lane; highest warp id = issue priority) + TMEM allocation.
//   warps 4-19  producers: build the 128-row activation tile in shared memory (UMMA K-major
//               SWIZZLE_128B, hi and lo halves, ring of 128-byte-wide chunks), from one of
//                 SRC_PLAIN    : x [M,cin] with the previous layer's BatchNorm+ReLU applied on load,
//                 SRC_GATHER   : feats[b, idx[row]] rows (the grouped tensor is never materialised);
//                                the 3 centred xyz channels are added in the epilogue in fp32,
//                 SRC_POINTMLP : relu(bn(W0 p + b0)) recomputed from the centred point p with the
//                                folded first layer -- the first layer's output never exists.
//   The last CTA to finish reduces the per-CTA statistic partials in fixed order and writes the
//   BatchNorm scale / shift of this layer (no separate kernel, deterministic).
#include "common.cuh"
#include "sa_mlp_tt.cuh"
#include "umma.cuh"

#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>

#include <type_traits>

namespace papc {
namespace tt {

using namespace umma;

constexpr int kTile = 128;                    // rows per tile == UMMA N
constexpr int kHalfBytes = kTile * 128;       // one [128 rows][128 B] SWIZZLE_128B block
constexpr int kXBytes = 2 * kHalfBytes;       // X chunk: hi block + lo block
#ifndef PAPC_TT_EPI_WARPS
#define PAPC_TT_EPI_WARPS 8
#endif
// Epilogue warps: warp w reads TMEM lane quadrant w % 4; with 8 warps the two warps of a quadrant
// (same SM sub-partition) split the 128 accumulator columns (= rows of the tile) in halves, so the
// latency-bound tcgen05.ld -> math -> store chains of two warps interleave on one scheduler.
constexpr int kEpiWarps = PAPC_TT_EPI_WARPS;
constexpr int kEpiHalves = kEpiWarps / 4;
constexpr int kBlkPerHalf = 4 / kEpiHalves;   // 32-column accumulator blocks per epilogue warp per tile
static_assert(kEpiWarps == 4 || kEpiWarps == 8, "epilogue warps: one or two per TMEM lane quadrant");
#ifdef PAPC_TT_TRIAGE
#define TT_DBG(a, bit) ((a).dbg & (bit))
// phase stamps of CTA 0 / 1 (slot = blockIdx) for the timeline printed by launch() under PAPC_TT_CLK=1
#define TT_CLK(a, e)                                                                       \
    do {                                                                                   \
        if ((a).clk != nullptr && blockIdx.x < 2) (a).clk[blockIdx.x * 16 + (e)] = clock64(); \
    } while (0)
// per-tile stamps of CTA 0 (local tiles 8..23): [48 + (tl-8)*8 + e]
#define TT_TCLK(a, tl, e)                                                                           \
    do {                                                                                            \
        if ((a).clk != nullptr && blockIdx.x == 0 && (tl) >= 8u && (tl) < 24u)                      \
            (a).clk[48 + ((tl) - 8u) * 8 + (e)] = clock64();                                        \
    } while (0)
// epilogue detail stamps of CTA 0 warp 0 (local tiles 8..23): [176 + (tl-8)*8 + e]
#define TT_ECLK(a, tl, e)                                                                           \
    do {                                                                                            \
        if ((a).clk != nullptr && blockIdx.x == 0 && threadIdx.x == 0 && (tl) >= 8u && (tl) < 24u)  \
            (a).clk[176 + ((tl) - 8u) * 8 + (e)] = clock64();                                       \
    } while (0)
// per-CTA %globaltimer stamps of a launch (PAPC_TT_GCLK=1; no host synchronisation, so programmatic
// dependent launch and the real kernel-to-kernel boundaries are observed): [blockIdx][8]
#define TT_GCLK(a, e)                                                                      \
    do {                                                                                   \
        if ((a).gclk != nullptr) {                                                         \
            unsigned long long t_;                                                         \
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_));                         \
            (a).gclk[blockIdx.x * 16 + (e)] = t_;                                           \
        }                                                                                  \
    } while (0)
#else
#define TT_DBG(a, bit) 0
#define TT_GCLK(a, e) do { } while (0)
#define TT_CLK(a, e) do { } while (0)
#define TT_TCLK(a, tl, e) do { } while (0)
#define TT_ECLK(a, tl, e) do { } while (0)
#endif
#ifndef PAPC_TT_PROD_WARPS
#define PAPC_TT_PROD_WARPS 8
#endif
constexpr int kProdWarps = PAPC_TT_PROD_WARPS;  // each thread: kRPT rows x one 16-byte unit per chunk
constexpr int kRPT = 128 * 8 / (kProdWarps * 32);  // rows per producer thread per chunk
constexpr int kRowStride = kTile / kRPT;       // rows rb + kRowStride * j
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kMmaWarp = kEpiWarps + kProdWarps;  // highest warp id: wins issue arbitration
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;  // 544 -> at most 120 registers / thread
constexpr int kMaxAct = 512;                  // reduction length with per-column scale / shift in smem
constexpr int kMaxFold = 128;                 // SRC_POINTMLP: channels of the folded first layer
static_assert(kRPT == 2 || kRPT == 4, "producer row mapping");
// tensor-memory columns: two accumulators, then W hi, then W lo (128 columns each)
constexpr uint32_t kColAcc = 0, kColWHi = 2 * kTile, kColWLo = 2 * kTile + 128;
constexpr int kTmemCols = 512;

template <int PREC> struct Prec;
template <> struct Prec<PREC_TF32> {
    static constexpr int kEPC = 32;    // K elements per 128-byte chunk row
    static constexpr int kEPU = 4;     // K elements per 16-byte unit
    static constexpr int kMmaK = 8;
    static constexpr int kTmemK = 128; // longest reduction whose W fits in tensor memory
    static constexpr uint32_t kFmt = 2;
};
template <> struct Prec<PREC_F16> {
    static constexpr int kEPC = 64;
    static constexpr int kEPU = 8;
    static constexpr int kMmaK = 16;
    static constexpr int kTmemK = 256;
    static constexpr uint32_t kFmt = 0;
};
// instruction descriptor: D fp32, A/B format fmt, both K-major, M = 128, N = 128
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt) {
    return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(kTile >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d), "l"(da), "l"(db),
                 "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t ta, uint64_t db, uint32_t idesc, uint32_t acc) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(ta), "l"(db),
                 "r"(idesc), "r"(acc) : "memory");
}
// two fp32 -> packed (fp16 hi pair, fp16 lo pair): hi = rn_f16(v), lo = rn_f16(v - hi)
__device__ __forceinline__ void split_f16x2(float a, float b, uint32_t &hi, uint32_t &lo) {
    const __half2 h = __floats2half2_rn(a, b);
    const float2 f = __half22float2(h);
    const __half2 l = __floats2half2_rn(a - f.x, b - f.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// The same split with the ReLU folded into the two conversions (every fp16-split operand is a post-ReLU
// activation): hi = relu(rz_f16(v)) -- rounded TOWARD ZERO so that the residual of a positive v is >= 0 --
// and lo = relu(rn_f16(v - hi)).  v < 0: hi = 0, residual = v < 0 -> lo = 0.  v >= 0: v - hi is exact and
// non-negative, the second ReLU does nothing.  |v - (hi + lo)| <= 2^-22 |v|, as for the round-to-nearest
// pair; saves the FMNMX per element in the producer warps, which bound the layer kernels.
__device__ __forceinline__ void split_relu_f16x2(float a, float b, uint32_t &hi, uint32_t &lo) {
    asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));   // first source -> upper half
    const float2 f = __half22float2(*reinterpret_cast<const __half2 *>(&hi));
    // residual as ONE packed FMA: (a, b) + (-1, -1) * (f.x, f.y), exact
    float ra, rb;
    unpack2(fma2(pack2(f.x, f.y), pack2(-1.f, -1.f), pack2(a, b)), ra, rb);
    asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(rb), "f"(ra));
}

// Shared memory: the small fixed tables first, then the operand ring (read by the tensor core) and
// the raw ring (fp32 activations landed by cp.async, thread-private slots), sized per instantiation.
struct SmemLayout {
    static constexpr uint32_t xyz = 0;                               // [4][128] float4 (SRC_GATHER)
    static constexpr uint32_t scale = xyz + 4 * kTile * 16;          // [512] float
    static constexpr uint32_t shift = scale + kMaxAct * 4;           // [512] float
    static constexpr uint32_t fold = shift + kMaxAct * 4;            // [128] float4 (SRC_POINTMLP)
    // hand-over slots between the two epilogue warps of a quadrant
    static constexpr uint32_t xpool = fold + kMaxFold * 16;          // [2][128] float2: K = 128 max/min
    static constexpr uint32_t xstat = xpool + 2 * kTile * 8;         // [128] double2: statistic sums
    static constexpr uint32_t cs = xstat + kTile * 16;               // [512] float: fp16 column scale of W
    static constexpr uint32_t bars = cs + kMaxAct * 4;
    static constexpr uint32_t nbars = 2 * 6 + 2 + 2 + 1 + 4 + 2 * 4;  // ... + raw_full[4], raw_empty[4]
    static constexpr uint32_t misc = bars + nbars * 8;               // tmem slot, last-CTA flag
    static constexpr uint32_t ring = (misc + 16 + 1023) / 1024 * 1024;  // operand ring (1024-aligned)
};
#ifndef PAPC_TT_PM_STAGES
#define PAPC_TT_PM_STAGES 6
#endif
#ifndef PAPC_TT_PAIR_STAGES
#define PAPC_TT_PAIR_STAGES 3
#define PAPC_TT_PAIR_RAW 3
#endif
template <int MODE, int PREC, int WMODE, bool PAIR = false>
struct Cfg {
    static constexpr int kNV = Prec<PREC>::kEPU / 4;                 // 16-byte loads per row per chunk
    static constexpr int kStageBytes = WMODE ? 2 * kXBytes : kXBytes;  // [X hi][X lo]([W hi][W lo])
    static constexpr int kRawStageBytes =
        MODE == SRC_POINTMLP ? 0 : kProdThreads * 16 * (kRPT * kNV) + (MODE == SRC_GATHER ? kTile * 4 * 6 : 0);
    // operand stages / raw stages (cp.async groups in flight per thread = kRawStages - 1)
    static constexpr int kStages = PAIR ? PAPC_TT_PAIR_STAGES
                                   : MODE == SRC_POINTMLP ? PAPC_TT_PM_STAGES : WMODE ? 2 : (PREC == PREC_F16 || MODE == SRC_GATHER) ? 3 : 4;
    static constexpr int kRawStages =
        PAIR ? PAPC_TT_PAIR_RAW
             : MODE == SRC_POINTMLP ? 0 : (PREC == PREC_F16 ? (WMODE ? 2 : 3) : (MODE == SRC_GATHER && WMODE) ? 3 : 4);
    static constexpr uint32_t raw = SmemLayout::ring + kStages * kStageBytes;
    static constexpr uint32_t total = raw + kRawStages * kRawStageBytes;
    static constexpr uint32_t bytes = total + 1024;  // + alignment slack
    static_assert(bytes <= 227 * 1024, "shared memory budget");
    static_assert(kStages <= 6, "barrier array");
};

__device__ __forceinline__ void cp_async16(uint32_t dst_saddr, const void *src, bool valid) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst_saddr), "l"(src),
                 "r"(valid ? 16 : 0) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst_saddr, const void *src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst_saddr), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

struct RowGeom {
    long long bN;  // b * N
    int g;         // group index b*S + s
    int k;         // position inside the group
};
__device__ __forceinline__ uint32_t fastdiv(uint32_t x, uint32_t mul, uint32_t shr) {
    return mul == 0u ? x : (__umulhi(x, mul) >> shr);
}
template <typename A>
__device__ __forceinline__ RowGeom row_geom(long long row, const A &a) {
    RowGeom r;
    if (a.fastgeom) {
        const uint32_t rw = (uint32_t)row;
        const uint32_t gg = fastdiv(rw, a.kmul, a.kshr);
        r.k = (int)(rw - gg * (uint32_t)a.K);
        r.g = (int)gg;
        r.bN = (long long)fastdiv(gg, a.smul, a.sshr) * a.N;
    } else {
        const long long gg = row / a.K;
        r.k = (int)(row - gg * a.K);
        r.g = (int)gg;
        r.bN = (gg / a.S) * (long long)a.N;
    }
    return r;
}

// ------------------------------------------------------------------ streamed-W image (WMODE 1)
// img: [nt][KC][hi | lo][128 rows][128 B] in the SWIZZLE_128B layout, so that a CTA stages the
// (tile_n, chunk) operand with ONE linear TMA bulk copy of 32 KiB.
template <int PREC>
__global__ void __launch_bounds__(256)
prep_wimg_kernel(const float *__restrict__ W, int wld, int wk0, int cin, int cout,
                 const float *__restrict__ colscale, int cs_on, const float *__restrict__ cs_gamma,
                 const float *__restrict__ cs_beta, float cs_sqrt_count, int KC,
                 uint8_t *__restrict__ img) {
    using P = Prec<PREC>;
    // the layer kernel that follows is launched as a programmatic dependent of THIS kernel: let its CTAs start
    // their prologue (barriers, tensor memory, weight staging) now; it reads the image only after griddepcontrol.wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    const int nt = ceil_div(cout, kTile);
    const long long total = (long long)nt * KC * kTile * 8;  // one 16-byte unit per thread step
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int u = (int)(e & 7);
        const int r = (int)((e >> 3) & (kTile - 1));
        const int c = (int)((e >> 10) % KC);
        const int tn = (int)((e >> 10) / KC);
        const int row = tn * kTile + r;
        float v[P::kEPU];
#pragma unroll
        for (int q = 0; q < P::kEPU; ++q) {
            const int k = c * P::kEPC + u * P::kEPU + q;
            float w = 0.f;
            if (row < cout && k < cin) {
                w = W[(size_t)row * wld + wk0 + k];
                if (cs_on)
                    w *= f16_colscale_sq(cs_gamma ? cs_gamma[k] : 1.f, cs_beta ? cs_beta[k] : 0.f, cs_sqrt_count);
                else if (colscale != nullptr) w *= colscale[k];
            }
            v[q] = w;
        }
        uint32_t hi[4], lo[4];
        if (PREC == PREC_TF32) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                float h, l;
                split_tf32(v[q], h, l);
                hi[q] = __float_as_uint(h);
                lo[q] = __float_as_uint(l);
            }
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) split_f16x2(v[(2 * q) % P::kEPU], v[(2 * q + 1) % P::kEPU], hi[q], lo[q]);
        }
        uint8_t *blk = img + ((size_t)(tn * KC + c) * 2) * kHalfBytes + (size_t)r * 128 + ((u ^ (r & 7)) << 4);
        *reinterpret_cast<uint4 *>(blk) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(blk + kHalfBytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// ------------------------------------------------------------------ activation image (SRC_PLAIN, small M)
// img: [tiles_m][KC][hi | lo][128 rows][128 B], the layout the producers write into an operand stage, holding
// relu(in_scale[k] * x[row][k] + in_shift[k]) split into fp16 hi / lo.  Rows >= M and columns >= cin are zero.
// d.fix != nullptr: the BatchNorm of x was deferred to this kernel (TtArgs::in_fix) -- every block derives the
// scale / shift table into shared memory first (cin <= kMaxAct), block 0 records the batch mean / variance.
__global__ void __launch_bounds__(256)
prep_ximg_kernel(const float *__restrict__ x, long long M, int cin, const float *__restrict__ in_scale,
                 const float *__restrict__ in_shift, int KC, uint8_t *__restrict__ img, const DeferredIn d) {
    using P = Prec<PREC_F16>;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // see prep_wimg_kernel
    __shared__ __align__(16) float s_sc[kMaxAct], s_sh[kMaxAct];
    if (d.fix != nullptr) {
        deferred_table<kMaxAct / 256>(d, cin, threadIdx.x, 256, blockIdx.x == 0, s_sc, s_sh, nullptr, nullptr);
        __syncthreads();
        in_scale = s_sc;
        in_shift = s_sh;
    }
    const long long tiles_m = ceil_div<long long>(M, kTile);
    const long long total = tiles_m * KC * kTile * 8;  // one 16-byte unit (8 halves) per thread step
    for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < total;
         e += (long long)gridDim.x * blockDim.x) {
        const int u = (int)(e & 7);
        const int r = (int)((e >> 3) & (kTile - 1));
        const int c = (int)((e >> 10) % KC);
        const long long tm = (e >> 10) / KC;
        const long long row = tm * kTile + r;
        const int k0 = c * P::kEPC + u * P::kEPU;
        float v[P::kEPU];
#pragma unroll
        for (int q = 0; q < P::kEPU; ++q) v[q] = 0.f;
        if (row < M && k0 < cin) {   // cin % 8 == 0: whole units are valid or not
            const float4 a0 = *reinterpret_cast<const float4 *>(x + row * cin + k0);
            const float4 a1 = *reinterpret_cast<const float4 *>(x + row * cin + k0 + 4);
            const float4 s0 = *reinterpret_cast<const float4 *>(in_scale + k0), s1 = *reinterpret_cast<const float4 *>(in_scale + k0 + 4);
            const float4 h0 = *reinterpret_cast<const float4 *>(in_shift + k0), h1 = *reinterpret_cast<const float4 *>(in_shift + k0 + 4);
            v[0] = fmaxf(fmaf(a0.x, s0.x, h0.x), 0.f); v[1] = fmaxf(fmaf(a0.y, s0.y, h0.y), 0.f);
            v[2] = fmaxf(fmaf(a0.z, s0.z, h0.z), 0.f); v[3] = fmaxf(fmaf(a0.w, s0.w, h0.w), 0.f);
            v[4] = fmaxf(fmaf(a1.x, s1.x, h1.x), 0.f); v[5] = fmaxf(fmaf(a1.y, s1.y, h1.y), 0.f);
            v[6] = fmaxf(fmaf(a1.z, s1.z, h1.z), 0.f); v[7] = fmaxf(fmaf(a1.w, s1.w, h1.w), 0.f);
        }
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) split_f16x2(v[2 * q], v[2 * q + 1], hi[q], lo[q]);
        uint8_t *blk = img + ((size_t)(tm * KC + c) * 2) * kHalfBytes + (size_t)r * 128 + ((u ^ (r & 7)) << 4);
        *reinterpret_cast<uint4 *>(blk) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
        *reinterpret_cast<uint4 *>(blk + kHalfBytes) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
    }
}

// --------------------------------------------------------------------------------- main kernel
// WMODE 0: W resident in tensor memory (TS UMMA).  WMODE 1: W chunks streamed through the ring.
// PAIR: one CTA owns TWO 128-channel tiles (cout up to 256 per CTA) of every row tile it visits: the activation
// chunk is converted ONCE and multiplied by both weight tiles (two passes of the MMA issuer over the same operand
// stages, accumulator h = channel tile h), instead of two CTAs converting the same rows.  Needs W of both tiles
// in tensor memory: 2 x (64 + 64) columns, i.e. cin <= kTmemK / 2 (SRC_PLAIN, W resident).
template <int MODE, int PREC, int WMODE, bool POOL, bool PAIR = false>
__global__ void __launch_bounds__(kThreads, 1)
mlp_layer_tt_kernel(const __grid_constant__ TtArgs a) {
    using P = Prec<PREC>;
    using CF = Cfg<MODE, PREC, WMODE, PAIR>;
    static_assert(!PAIR || (MODE == SRC_PLAIN && WMODE == 0), "paired channel tiles: plain source, resident W");
    constexpr int kChT = PAIR ? 2 : 1;                        // channel tiles per CTA
    // tensor-memory columns of channel tile h: W hi / lo (PAIR: 64 columns each)
    constexpr uint32_t kWHi0 = kColWHi, kWLo0 = PAIR ? kColWHi + 64 : kColWLo, kWStep = PAIR ? 128 : 0;
    constexpr int kStages = CF::kStages;
    constexpr int kStageBytes = CF::kStageBytes;
    constexpr int NV = CF::kNV;
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float *s_scale = reinterpret_cast<float *>(smem + SmemLayout::scale);
    float *s_shift = reinterpret_cast<float *>(smem + SmemLayout::shift);
    float4 *s_fold = reinterpret_cast<float4 *>(smem + SmemLayout::fold);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SmemLayout::bars);
    uint64_t *x_full = bars;
    uint64_t *x_empty = bars + 6;
    uint64_t *acc_full = bars + 12;
    uint64_t *acc_empty = acc_full + 2;
    uint64_t *w_ready = acc_empty + 2;
    uint64_t *xyz_full = w_ready + 1;
    uint64_t *raw_full = xyz_full + 4;    // SRC_PLAIN: the raw stage's 128 row copies (TMA) have landed
    uint64_t *raw_empty = raw_full + 4;   // ... every producer warp has read the raw stage
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SmemLayout::misc);
    uint32_t *s_last = tmem_slot + 1;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    if (tid == 0) TT_CLK(a, 0);
    if (tid == 0) TT_GCLK(a, 0);   // entry
#ifdef PAPC_TT_TRIAGE
    unsigned long long gt0 = 0;
    if (tid == 0 && a.clk != nullptr) {
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
        atomicMin(a.clk + 44, gt0);
    }
#endif
    const int nt = ceil_div(a.cout, kTile * kChT);
    const int tile_n = blockIdx.x % nt;
    const int mi = blockIdx.x / nt;
    const int gm = gridDim.x / nt;
    const int n0 = tile_n * kTile * kChT;
    const long long tiles_m = ceil_div<long long>(a.M, kTile);
    const int KC = ceil_div(a.cin, P::kEPC);
    const int kpad = ceil_div(a.cin, P::kMmaK) * P::kMmaK;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(x_full + s, kProdWarps + (WMODE ? 1 : 0));  // + the TMA issuer's expect_tx
            mbar_init(x_empty + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full + b, 1);
            mbar_init(acc_empty + b, kEpiWarps);
        }
        mbar_init(w_ready, kEpiWarps);
        for (int b = 0; b < 4; ++b) mbar_init(xyz_full + b, kProdWarps);
        for (int b = 0; b < 4; ++b) {
            mbar_init(raw_full + b, a.tma2d ? 1 : kTile);  // one arrive.expect_tx per issuing thread
            mbar_init(raw_empty + b, kProdWarps);
        }
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    if (tid == 0) TT_CLK(a, 1);
    if (tid == 0) TT_GCLK(a, 1);   // barriers initialised, tensor memory allocated
    // Programmatic dependent launch.  As the primary: let the next kernel's CTAs be scheduled on
    // SMs as ours retire (they only run their prologue until we are completely done).  As the
    // dependent (a.pdl): everything up to pdl_wait() reads launch arguments and the weights only.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    auto pdl_wait = [&]() {
        if (a.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    };

    if (warp < kEpiWarps) {
        // ================================ epilogue =========================================
        const int quad = warp & 3;        // TMEM lane quadrant this warp may access (warp id % 4)
        const int half = warp >> 2;       // which accumulator column blocks (rows of the tile) it owns
        const int c = quad * 32 + lane;   // channel inside a 128-channel tile == TMEM lane
        const int cg_base = n0 + c;       // channel tile h of this CTA: cg_base + h * kTile
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        // ---- stage W into tensor memory (hi / lo split).  Thread = channel reads ITS OWN row
        //      W[cg, kc*kEPC ...] straight into the registers tcgen05.st takes (16-byte loads when the
        //      row is 16-byte aligned): every load of a chunk is independent, so the whole chunk is in
        //      flight at once -- no shared-memory transposition, no per-row round trips.  The two
        //      epilogue warps of a quadrant take alternate chunks.
        if (WMODE == 0) {
            float *s_cs = reinterpret_cast<float *>(smem + SmemLayout::cs);
            const bool scaled = a.cs_on || a.w_colscale != nullptr;
            if (scaled) {  // the column scale of every k once per CTA (256 threads), then 16-byte reads
                for (int k = tid; k < kMaxAct; k += kEpiWarps * 32) {
                    float c1 = 1.f;
                    if (k < a.cin) {
                        if (a.cs_on)
                            c1 = f16_colscale_sq(a.cs_gamma ? __ldg(a.cs_gamma + k) : 1.f,
                                                 a.cs_beta ? __ldg(a.cs_beta + k) : 0.f, a.cs_sqrt_count);
                        else
                            c1 = __ldg(a.w_colscale + k);
                    }
                    s_cs[k] = c1;
                }
                named_bar_sync(10, kEpiWarps * 32);
            }
            const bool vec = ((a.wld | a.wk0) & 3) == 0 && (reinterpret_cast<uintptr_t>(a.W) & 15u) == 0 &&
                             (a.w_colscale == nullptr || (reinterpret_cast<uintptr_t>(a.w_colscale) & 15u) == 0);
            for (int h = 0; h < kChT; ++h) {
            const int cg = cg_base + h * kTile;
            const bool cvalid = cg < a.cout;
            const float *wrow = a.W + (size_t)(cvalid ? cg : 0) * a.wld + a.wk0;
            for (int kc = half; kc < KC; kc += kEpiHalves) {  // one 32-column TMEM chunk = kEPC elements
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int pass = 0; pass < P::kEPC / 32; ++pass) {
                    const int k0 = kc * P::kEPC + pass * 32;
                    float v[32];
                    if (vec && k0 + 32 <= a.cin) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 t = __ldg(reinterpret_cast<const float4 *>(wrow + k0) + q);
                            v[4 * q] = t.x; v[4 * q + 1] = t.y; v[4 * q + 2] = t.z; v[4 * q + 3] = t.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            const int k = k0 + i;
                            v[i] = k < a.cin ? __ldg(wrow + k) : 0.f;
                        }
                    }
                    if (scaled) {
#pragma unroll
                        for (int q = 0; q < 8; ++q) {
                            const float4 t = *reinterpret_cast<const float4 *>(s_cs + ((k0 + 4 * q) & (kMaxAct - 1)));
                            v[4 * q] *= t.x; v[4 * q + 1] *= t.y; v[4 * q + 2] *= t.z; v[4 * q + 3] *= t.w;
                        }
                    }
                    if (!cvalid) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) v[i] = 0.f;
                    }
                    if (PREC == PREC_TF32) {
#pragma unroll
                        for (int i = 0; i < 32; ++i) {
                            float h, l;
                            split_tf32(v[i], h, l);
                            hi[i] = __float_as_uint(h);
                            lo[i] = __float_as_uint(l);
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; i += 2)
                            split_f16x2(v[i], v[i + 1], hi[(pass * 16 + i / 2) & 31], lo[(pass * 16 + i / 2) & 31]);
                    }
                }
                tmem_st32(lane_base + kWHi0 + h * kWStep + kc * 32, hi);
                tmem_st32(lane_base + kWLo0 + h * kWStep + kc * 32, lo);
            }
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(w_ready);
        }
        if (tid == 0) TT_CLK(a, 2);
        if (tid == 0) TT_GCLK(a, 2);   // W staged
        pdl_wait();
        if (tid == 0) TT_GCLK(a, 3);   // the previous kernel has completed
        float bias_h[kChT];
        double acc_s_h[kChT], acc_q_h[kChT];
#pragma unroll
        for (int h = 0; h < kChT; ++h) {
            const int cgh = cg_base + h * kTile;
            bias_h[h] = (a.bias != nullptr && cgh < a.cout) ? a.bias[cgh] : 0.f;
            acc_s_h[h] = 0.0;
            acc_q_h[h] = 0.0;
        }
        float wx = 0.f, wy = 0.f, wz = 0.f;
        const bool has_xyz = (MODE == SRC_GATHER) && a.wxyz >= 0;
        if (has_xyz && cg_base < a.cout) {
            const float *wp = a.W + (size_t)cg_base * a.wld + a.wxyz;
            wx = wp[0]; wy = wp[1]; wz = wp[2];
        }
        const uint64_t wx2 = pack2(wx, wx), wy2 = pack2(wy, wy), wz2 = pack2(wz, wz);
        const int kshift = POOL ? (a.K == 32 ? 5 : a.K == 64 ? 6 : 7) : 0;  // K in {32, 64, 128}
        const size_t ystride = (size_t)a.cout;
        float2 *s_xpool = reinterpret_cast<float2 *>(smem + SmemLayout::xpool);
        // one accumulator of one row tile: channel tile H (PAIR: accumulator H, filled once per row tile;
        // otherwise the two accumulators alternate between row tiles)
        auto tile_body = [&](auto hc, long long tile, uint32_t tl) {
            constexpr int H = decltype(hc)::value;
            const int cg = cg_base + H * kTile;
            const bool cvalid = cg < a.cout;
            const bool do_y = (!POOL || a.y != nullptr) && cvalid;
            const bool do_pool = POOL && cvalid;
            const uint64_t bias2 = pack2(bias_h[H], bias_h[H]);
            double &acc_s = acc_s_h[H], &acc_q = acc_q_h[H];
            const uint32_t buf = PAIR ? (uint32_t)H : (tl & 1);
            const uint32_t par = PAIR ? (tl & 1) : ((tl >> 1) & 1);
            const long long m0 = tile * kTile;
            const int nrows = (int)((a.M - m0) < kTile ? (a.M - m0) : kTile);
            TT_ECLK(a, tl, 0);   // loop top
            if (MODE == SRC_GATHER) mbar_wait(xyz_full + (tl & 3), (tl >> 2) & 1);
            mbar_wait(acc_full + buf, par);
            TT_ECLK(a, tl, 1);   // accumulator barrier passed
            tc_fence_after();
            TT_ECLK(a, tl, 2);   // tcgen05 fence done
            if (tid == 0 && tl == 0) TT_CLK(a, 5);
            if (tid == 0 && tl == 0) TT_GCLK(a, 4);   // first accumulator ready
            if (tid == 0 && tl == 1) TT_CLK(a, 11);
            if (tid == 0) TT_TCLK(a, tl, 4);   // epilogue: accumulator ready
            uint64_t s2 = 0ull, q2 = 0ull;     // packed (even rows, odd rows) running sums
            float mx = -INFINITY, mn = INFINITY;
            long long pend_g[kBlkPerHalf];
            float pend_mx[kBlkPerHalf], pend_mn[kBlkPerHalf];
#pragma unroll
            for (int bi = 0; bi < kBlkPerHalf; ++bi) { pend_g[bi] = -1; pend_mx[bi] = 0.f; pend_mn[bi] = 0.f; }
#pragma unroll
            for (int bi = 0; bi < kBlkPerHalf; ++bi) {  // 32 accumulator columns (= rows) at a time
                const int blk = half * kBlkPerHalf + bi;
                uint32_t r[32];
                tmem_ld32_nowait(lane_base + kColAcc + buf * kTile + blk * 32, r);
                tmem_wait_ld();
                TT_ECLK(a, tl, 3 + 2 * bi);   // block bi loaded
                if (bi == kBlkPerHalf - 1) {
                    // this warp's share of the accumulator is in registers: hand the buffer back
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty + buf);
                    if (tid == 0) TT_TCLK(a, tl, 5);   // epilogue: buffer handed back
                }
                const uint32_t xs = smem_u32(smem) + SmemLayout::xyz + 16 * ((tl & 3) * kTile + blk * 32);
                float *yp = a.y + ((size_t)m0 + blk * 32) * ystride + cg;   // dereferenced only if do_y
                const int nr = nrows - blk * 32;  // valid rows of this block (may be <= 0)
                auto body = [&](auto full_tag, auto store_tag) {
                    constexpr bool FULL = decltype(full_tag)::value;
                    constexpr bool STORE = decltype(store_tag)::value;
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        uint64_t v2 = add2(pack2u(r[i], r[i + 1]), bias2);
                        if (has_xyz) {
                            const float4 p0 = lds128f(xs + 16 * i), p1 = lds128f(xs + 16 * (i + 1));
                            v2 = fma2(wx2, pack2(p0.x, p1.x), v2);
                            v2 = fma2(wy2, pack2(p0.y, p1.y), v2);
                            v2 = fma2(wz2, pack2(p0.z, p1.z), v2);
                        }
                        float va, vb;
                        unpack2(v2, va, vb);
                        if (FULL) {
                            s2 = add2(s2, v2);
                            q2 = fma2(v2, v2, q2);
                            if (STORE) {
                                yp[0] = va;
                                yp[ystride] = vb;
                            }
                            if (POOL) {
                                mx = fmaxf(fmaxf(mx, va), vb);
                                mn = fminf(fminf(mn, va), vb);
                            }
                        } else {
                            const bool rva = i < nr, rvb = i + 1 < nr;
                            const uint64_t m2 = pack2(rva ? va : 0.f, rvb ? vb : 0.f);
                            s2 = add2(s2, m2);
                            q2 = fma2(m2, m2, q2);
                            if (STORE && rva) yp[0] = va;
                            if (STORE && rvb) yp[ystride] = vb;
                            if (POOL) {
                                if (rva) { mx = fmaxf(mx, va); mn = fminf(mn, va); }
                                if (rvb) { mx = fmaxf(mx, vb); mn = fminf(mn, vb); }
                            }
                        }
                        if (STORE) yp += 2 * ystride;
                    }
                };
                if (TT_DBG(a, 4)) {
                    s2 = pack2u(r[0], r[31]);
                } else if (nr >= 32) {
                    if (do_y) body(std::true_type{}, std::true_type{});
                    else body(std::true_type{}, std::false_type{});
                } else {
                    if (do_y) body(std::false_type{}, std::true_type{});
                    else body(std::false_type{}, std::false_type{});
                }
                TT_ECLK(a, tl, 4 + 2 * bi);   // block bi arithmetic done
                // pooled groups of K rows end at multiples of K (K in {32, 64, 128})
                // (K is 32 / 64 / 128 here: a mask, not a runtime integer division per block)
                // The extrema are parked and stored after the LAST block of the tile: tcgen05.wait::ld
                // also drains the warp's outstanding global stores (measured: the wait after a block
                // with stores in flight took 1500 cycles instead of 300), so no store may sit between
                // two accumulator loads.
                if (POOL && kshift != 7 && ((((blk + 1) * 32) & (a.K - 1)) == 0)) {
                    pend_g[bi] = (do_pool && nr > 0) ? (long long)((m0 + blk * 32) >> kshift) : -1;
                    pend_mx[bi] = mx;
                    pend_mn[bi] = mn;
                    mx = -INFINITY;
                    mn = INFINITY;
                }
            }
            if (POOL && kshift != 7) {
#pragma unroll
                for (int bi = 0; bi < kBlkPerHalf; ++bi)
                    if (pend_g[bi] >= 0) {
                        a.pool_max[pend_g[bi] * a.cout + cg] = pend_mx[bi];
                        a.pool_min[pend_g[bi] * a.cout + cg] = pend_mn[bi];
                    }
            }
            if (POOL && kshift == 7) {  // K = 128: the tile is one group
                if (kEpiHalves == 2) {
                    // the upper half hands its extrema to the lower half (double buffered: the next
                    // hand-over into this slot happens two named barriers later)
                    if (half == 1) s_xpool[buf * kTile + c] = make_float2(mx, mn);
                    named_bar_sync(2 + quad, 64);
                    if (half == 0) {
                        const float2 o = s_xpool[buf * kTile + c];
                        mx = fmaxf(mx, o.x);
                        mn = fminf(mn, o.y);
                    }
                }
                if (do_pool && half == 0) {
                    const long long g = m0 >> 7;
                    a.pool_max[g * a.cout + cg] = mx;
                    a.pool_min[g * a.cout + cg] = mn;
                }
            }
            if (tid == 0) TT_TCLK(a, tl, 6);   // epilogue: tile done
            TT_ECLK(a, tl, 6);   // (overwrites the block-1 stamp: stores issued)
            float sa, sb, qa, qb;
            unpack2(s2, sa, sb);
            unpack2(q2, qa, qb);
            // one fp64 add per quantity and tile (sa + sb in fp32 costs one rounding at 2^-24
            // relative, far inside the statistic's own fp32 accumulation error).  Flushing only every
            // fourth tile was measured: no gain (the epilogue is starved, not fp64 bound).
            acc_s += (double)(sa + sb);
            acc_q += (double)(qa + qb);
            TT_ECLK(a, tl, 7);   // tile bookkeeping done
        };
        uint32_t tl = 0;
        for (long long tile = mi; tile < tiles_m; tile += gm, ++tl) {
            tile_body(std::integral_constant<int, 0>{}, tile, tl);
            if (PAIR) tile_body(std::integral_constant<int, kChT - 1>{}, tile, tl);
        }
        if (tid == 0) TT_CLK(a, 6);
#pragma unroll
        for (int h = 0; h < kChT; ++h) {
        const int cg = cg_base + h * kTile;
        const bool cvalid = cg < a.cout;
        double acc_s = acc_s_h[h], acc_q = acc_q_h[h];
        if (kEpiHalves == 2) {  // fold the two halves' statistics (fixed order -> deterministic)
            double2 *s_xstat = reinterpret_cast<double2 *>(smem + SmemLayout::xstat);
            if (h > 0) named_bar_sync(6 + quad, 64);   // the lower half has read the previous tile's slot
            if (half == 1) s_xstat[c] = make_double2(acc_s, acc_q);
            named_bar_sync(6 + quad, 64);
            if (half == 0) {
                const double2 o = s_xstat[c];
                acc_s += o.x;
                acc_q += o.y;
            }
        }
        if (a.stats_partial != nullptr && cvalid && half == 0) {
            // one partial row per CTA (fixed order across launches -> deterministic)
            a.stats_partial[((long long)mi * 2 + 0) * a.cout + cg] = acc_s;
            a.stats_partial[((long long)mi * 2 + 1) * a.cout + cg] = acc_q;
            // partial rows this launch does not own are zeroed (fixed row count per M)
            for (long long rr = mi + gm; rr < a.partial_rows; rr += gm) {
                a.stats_partial[(rr * 2 + 0) * a.cout + cg] = 0.0;
                a.stats_partial[(rr * 2 + 1) * a.cout + cg] = 0.0;
            }
            if (a.fix_acc != nullptr) {
                // the same sums once more, for the last CTA's fast path (TtArgs::fix_acc) -- or, deferred
                // finalisation (no counter), for the next layer's CTAs
                const double v2[2] = {acc_s, acc_q};
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double p = v2[i];
                    if (fabs(p) < 0x1p53) {   // <= 296 CTAs: the integer words cannot overflow
                        const double ip = trunc(p);
                        atomicAdd(a.fix_acc + (size_t)(2 * i) * a.cout + cg, (unsigned long long)(long long)ip);
                        atomicAdd(a.fix_acc + (size_t)(2 * i + 1) * a.cout + cg,
                                  (unsigned long long)__double2ll_rn((p - ip) * 0x1p54));
                    } else {
                        atomicOr(a.fix_acc + (size_t)4 * a.cout, 1ull);
                    }
                }
            }
            __threadfence();
        }
        }
    } else if (warp == kMmaWarp) {
        // ================================ MMA issuer =======================================
        // The whole warp runs the (uniform) control flow; one elected lane issues the tcgen05 ops.
        pdl_wait();
        if (WMODE == 0) {
            mbar_wait(w_ready, 0);
            tc_fence_after();
        }
        constexpr uint32_t idesc = make_idesc(P::kFmt);
        const uint32_t ring_base = smem_u32(smem + SmemLayout::ring);
        uint32_t it = 0, tl = 0;
        for (long long tile = mi; tile < tiles_m; tile += gm, ++tl) {
          // PAIR: two passes over the row tile's operand stages, pass h -> accumulator h with the weights of
          // channel tile h; the stages are awaited in pass 0 and handed back to the producers in pass 1
          const uint32_t it0 = it;
#pragma unroll
          for (int h = 0; h < kChT; ++h) {
            const uint32_t buf = PAIR ? (uint32_t)h : (tl & 1);
            mbar_wait(acc_empty + buf, (PAIR ? (tl & 1) : ((tl >> 1) & 1)) ^ 1);
            tc_fence_after();
            if (lane == 0) TT_TCLK(a, tl, 2);   // MMA: accumulator buffer free
            const uint32_t d_tmem = tmem_base + kColAcc + buf * kTile;
            it = it0;
            for (int c = 0; c < KC; ++c, ++it) {
                const uint32_t s = it % kStages;
                if (h == 0) mbar_wait(x_full + s, (it / kStages) & 1);
                tc_fence_after();
                const int nks = TT_DBG(a, 2) ? 0 : min(P::kEPC, kpad - c * P::kEPC) / P::kMmaK;  // 1..4 K steps
                const uint32_t x_hi = ring_base + s * kStageBytes;
                // descriptors of K step 0; one K step = 32 bytes of the swizzle row = +2 in the
                // (address >> 4) field, = 8 tensor-memory columns
                const uint64_t dxh = make_desc_sw128(x_hi), dxl = make_desc_sw128(x_hi + kHalfBytes);
                const uint64_t dwh = make_desc_sw128(x_hi + kXBytes), dwl = make_desc_sw128(x_hi + kXBytes + kHalfBytes);
                const uint32_t w_hi = tmem_base + kWHi0 + h * kWStep + c * 32, w_lo = tmem_base + kWLo0 + h * kWStep + c * 32;
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < 4; ++ks) {
                        if (ks < nks) {
                            const uint32_t acc0 = (c | ks) != 0;
                            // small terms first
                            if (WMODE == 0) {
                                if (PREC == PREC_TF32) {
                                    mma_tf32_ts(d_tmem, w_lo + ks * 8, dxh + ks * 2, idesc, acc0);
                                    mma_tf32_ts(d_tmem, w_hi + ks * 8, dxl + ks * 2, idesc, 1);
                                    mma_tf32_ts(d_tmem, w_hi + ks * 8, dxh + ks * 2, idesc, 1);
                                } else {
                                    mma_f16_ts(d_tmem, w_lo + ks * 8, dxh + ks * 2, idesc, acc0);
                                    mma_f16_ts(d_tmem, w_hi + ks * 8, dxl + ks * 2, idesc, 1);
                                    mma_f16_ts(d_tmem, w_hi + ks * 8, dxh + ks * 2, idesc, 1);
                                }
                            } else {
                                if (PREC == PREC_TF32) {
                                    mma_tf32_ss(d_tmem, dwl + ks * 2, dxh + ks * 2, idesc, acc0);
                                    mma_tf32_ss(d_tmem, dwh + ks * 2, dxl + ks * 2, idesc, 1);
                                    mma_tf32_ss(d_tmem, dwh + ks * 2, dxh + ks * 2, idesc, 1);
                                } else {
                                    mma_f16_ss(d_tmem, dwl + ks * 2, dxh + ks * 2, idesc, acc0);
                                    mma_f16_ss(d_tmem, dwh + ks * 2, dxl + ks * 2, idesc, 1);
                                    mma_f16_ss(d_tmem, dwh + ks * 2, dxh + ks * 2, idesc, 1);
                                }
                            }
                        }
                    }
                    if (h == kChT - 1) mma_commit(x_empty + s);    // chunk reusable once these MMAs have read it
                    if (c == KC - 1) mma_commit(acc_full + buf);   // accumulator complete
                    if (it == 0) TT_CLK(a, 4);
                    if (it == 0) TT_GCLK(a, 11);           // first MMAs issued
                    if (c == KC - 1) TT_TCLK(a, tl, 3);   // MMA: last chunk issued + committed
                }
                __syncwarp();
            }
            if (TT_DBG(a, 64)) {  // triage: the issuer itself waits for the tile's MMAs to complete
                mbar_wait(acc_full + buf, PAIR ? (tl & 1) : ((tl >> 1) & 1));
                if (lane == 0) TT_TCLK(a, tl, 7);
            }
          }
        }
    } else {
        // ================================ producers ========================================
        const int ptid = tid - kEpiWarps * 32;  // 0..255
        const int u = ptid & 7;       // 16-byte unit inside the 128-byte chunk row
        const int rb = ptid >> 3;     // rows rb + kRowStride*j, j = 0..kRPT-1
        const uint32_t swz = (uint32_t)((u ^ (rb & 7)) << 4);
        const bool has_act = (MODE == SRC_PLAIN && (a.in_scale != nullptr || a.in_fix != nullptr)) ||
                             (MODE == SRC_GATHER && a.x_colscale != nullptr);
        // the tensor-map descriptor lives in the kernel parameters: have the TMA unit fetch it while
        // this kernel still waits for the previous one (PAPC_TT_DBG=512 switches the prefetch off;
        // measured on the B200: no difference in the step time either way)
        if (MODE == SRC_PLAIN && a.tma2d && ptid == 0 && (a.dbg & 512) == 0)
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&a.xmap)) : "memory");
        pdl_wait();
        // SRC_PLAIN fills the scale / shift table after its first activation copies are in flight
        // (below): the two global round trips overlap instead of adding up at every kernel start.
        if (MODE == SRC_POINTMLP && ptid < kMaxFold)
            s_fold[ptid] = ptid < a.cin ? reinterpret_cast<const float4 *>(a.l0_fold)[ptid]
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        if (MODE == SRC_GATHER && a.x_colscale != nullptr) {
            // fp16-split gather: features are >= 0 (post-ReLU) and bounded; "activation" = division by the
            // power-of-two column scale (exact), through the same scale / shift table the PLAIN mode uses
            for (int k = ptid; k < kMaxAct; k += kProdThreads) {
                s_scale[k] = k < a.cin ? __frcp_rn(__ldg(a.x_colscale + k)) : 0.f;
                s_shift[k] = 0.f;
            }
        }
        if (MODE != SRC_PLAIN) named_bar_sync(1, kProdThreads);

        // ---- per-row geometry pipeline (SRC_GATHER / SRC_POINTMLP)
        long long src[kRPT] = {};          // b*N + n of the rows of the tile being issued
        long long srcD[kRPT] = {};         // ... times D: element offset of the feature row
        int grp[kRPT] = {};
        int idx_pf[kRPT] = {};             // prefetched neighbour indices of a later tile
        long long bN_pf[kRPT] = {};
        int grp_pf[kRPT] = {};
        float4 p_cur[kRPT];                // SRC_POINTMLP: centred points of the current tile
#pragma unroll
        for (int j = 0; j < kRPT; ++j) p_cur[j] = make_float4(0.f, 0.f, 0.f, 0.f);

        auto fetch_idx = [&](long long tile) {  // -> idx_pf / bN_pf / grp_pf for `tile`
#pragma unroll
            for (int j = 0; j < kRPT; ++j) {
                long long row = tile * kTile + rb + kRowStride * j;
                row = row < a.M ? row : a.M - 1;
                const RowGeom rg = row_geom(row, a);
                bN_pf[j] = rg.bN;
                grp_pf[j] = rg.g;
                idx_pf[j] = a.idx != nullptr ? __ldg(a.idx + row) : rg.k;
            }
        };
        // ---- SRC_POINTMLP: the tile's 128 centred points are fetched ONCE (producer thread r < 128
        //      owns row r: index two tiles ahead, point one tile ahead, both in flight across a whole
        //      tile) and handed to the 8 threads that expand that row through a double-buffered
        //      shared-memory table -- instead of every thread gathering its rows' points itself.
        int idx1 = 0, grp1 = 0;
        long long bN1 = 0;
        // raw loads of a later tile's point and centroid; they are combined only when that tile is published (any
        // arithmetic on them here would stall this in-order warp for the whole global round trip)
        float px1 = 0.f, py1 = 0.f, pz1 = 0.f, cx1 = 0.f, cy1 = 0.f, cz1 = 0.f;
        auto fetch_idx1 = [&](long long tile) {
            long long row = tile * kTile + (ptid & (kTile - 1));
            row = row < a.M ? row : a.M - 1;
            const RowGeom rg = row_geom(row, a);
            bN1 = rg.bN;
            grp1 = rg.g;
            idx1 = a.idx != nullptr ? __ldg(a.idx + row) : rg.k;
        };
        auto fetch_pt1 = [&]() {  // idx1 (arrived) -> raw point / centroid loads
            const int n = min(max(idx1, 0), a.N - 1);
            const float *q = a.xyz + (bN1 + n) * 3;
            px1 = __ldg(q); py1 = __ldg(q + 1); pz1 = __ldg(q + 2);
            if (a.new_xyz != nullptr) {
                const float *cc = a.new_xyz + (long long)grp1 * 3;
                cx1 = __ldg(cc); cy1 = __ldg(cc + 1); cz1 = __ldg(cc + 2);
            }
        };

        // ---- raw ring (SRC_PLAIN / SRC_GATHER): every thread lands ITS OWN 16-byte units of a chunk
        //      with cp.async into thread-private slots (slot q of thread t at (q*256 + t)*16: conflict
        //      free) and reads them back itself -- global latency is hidden by the depth of the ring,
        //      not by registers or warps.
        const uint32_t sm = smem_u32(smem);
        auto raw_slot = [&](int rs, int q) -> uint32_t {
            return sm + CF::raw + rs * CF::kRawStageBytes + (q * kProdThreads + ptid) * 16;
        };
        auto raw_xyz = [&](int rs, int q) -> uint32_t {  // SRC_GATHER: 6 floats per row (threads u < kRPT)
            return sm + CF::raw + rs * CF::kRawStageBytes + kProdThreads * 16 * (kRPT * NV) +
                   4 * (q * kTile + rb + kRowStride * (u & (kRPT - 1)));
        };
        uint32_t raw_issued = 0, raw_read = 0;  // SRC_PLAIN: chunks issued into / read from the raw ring
        auto issue = [&](long long tile, int c, int rs) {  // chunk (tile, c) -> raw stage rs
            if (TT_DBG(a, 16)) return;                          // triage: no global loads
            const long long m0 = tile * kTile;
            const int k0 = c * P::kEPC + u * P::kEPU;
            if (MODE == SRC_PLAIN) {
                // The raw fp32 tile arrives through the TMA unit, not through the LSU: the LSU / MIO path of
                // this SM is shared with the epilogue's TMEM loads and stores, and 2 048 16-byte cp.async per
                // chunk cost 25-40 % of the kernel at a third of the HBM rate.  Default: ONE 2-D tensor copy
                // per chunk (measured on B200: sa2.l3 141 -> 98 us, sa1.l3 87 -> 59 us, step 0.645 -> 0.601 ms).
                // Fallback (no descriptor: M >= 2^31 or the driver entry point is missing): one bulk copy per
                // row, thread r < 128 owns row r (correct but slower than cp.async: 128 small copies per chunk).
                if (a.tma2d) {
                    // one 2-D tensor copy per chunk: box = 128 rows x kEPC floats at (k = c*kEPC, row = m0);
                    // rows >= M and columns >= cin arrive as zeros (out-of-bounds fill)
                    if (ptid == 0) {
                        const uint32_t use = raw_issued / (uint32_t)(CF::kRawStages > 0 ? CF::kRawStages : 1);
                        mbar_wait(raw_empty + rs, (use & 1u) ^ 1u);
                        mbar_expect_tx(raw_full + rs, (uint32_t)(kTile * P::kEPC * 4));
                        asm volatile(
                            "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                            ::"r"(sm + CF::raw + rs * CF::kRawStageBytes), "l"(reinterpret_cast<uint64_t>(&a.xmap)),
                            "r"(c * P::kEPC), "r"((int)m0), "r"(smem_u32(raw_full + rs))
                            : "memory");
                    }
                } else if (ptid < kTile) {
                    long long row = m0 + ptid;
                    row = row < a.M ? row : a.M - 1;
                    if (TT_DBG(a, 128)) row = ptid;
                    const int kleft = a.cin - c * P::kEPC;                       // > 0
                    const uint32_t bytes = (uint32_t)(kleft < P::kEPC ? kleft : P::kEPC) * 4u;
                    const uint32_t use = raw_issued / (uint32_t)(CF::kRawStages > 0 ? CF::kRawStages : 1);  // previous uses of this stage
                    mbar_wait(raw_empty + rs, (use & 1u) ^ 1u);
                    mbar_expect_tx(raw_full + rs, bytes);
                    bulk_g2s(smem + CF::raw + rs * CF::kRawStageBytes + ptid * (P::kEPC * 4),
                             a.x + row * a.cin + c * P::kEPC, bytes, raw_full + rs);
                }
                ++raw_issued;
            } else if (MODE == SRC_GATHER) {
                if (c == 0) {
#pragma unroll
                    for (int j = 0; j < kRPT; ++j) {
                        src[j] = bN_pf[j] + min(max(idx_pf[j], 0), a.N - 1);
                        srcD[j] = src[j] * a.D;
                        grp[j] = grp_pf[j];
                    }
                    fetch_idx(tile + gm < tiles_m ? tile + gm : tile);
                    if (u < kRPT) {  // this thread also stages the centred xyz of row rb + 64*u
                        long long sj = src[0];
                        int gj = grp[0];
#pragma unroll
                        for (int j = 1; j < kRPT; ++j)
                            if (u == j) { sj = src[j]; gj = grp[j]; }
                        const float *q = a.xyz + sj * 3;
                        const float *cc = a.new_xyz != nullptr ? a.new_xyz + (long long)gj * 3 : q;
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            cp_async4(raw_xyz(rs, d), q + d);
                            cp_async4(raw_xyz(rs, 3 + d), cc + d);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < kRPT; ++j) {
                    const bool ok = k0 < a.D;
                    const float *p = ok ? a.feats + (srcD[j] + k0) : a.feats;
#pragma unroll
                    for (int h = 0; h < NV; ++h) cp_async16(raw_slot(rs, j * NV + h), p + (ok ? 4 * h : 0), ok);
                }
            }
        };

        uint32_t it = 0, ptl = 0;
        const uint8_t *wimg = reinterpret_cast<const uint8_t *>(a.wimg);
        auto process = [&](long long tile, int c, int rs) {
            const uint32_t s = it % kStages;
            const int k0 = c * P::kEPC + u * P::kEPU;
            float o[kRPT][P::kEPU];
            float e0 = 0.f, e1 = 0.f, e2 = 0.f;
            if (MODE == SRC_PLAIN || MODE == SRC_GATHER) {
                float4 v[kRPT][NV];
                if (MODE == SRC_PLAIN && !TT_DBG(a, 16)) {
                    mbar_wait(raw_full + rs, (raw_read / (uint32_t)(CF::kRawStages > 0 ? CF::kRawStages : 1)) & 1u);
                    if (ptid == 0 && raw_read == 0) TT_GCLK(a, 13);  // first raw chunk landed
                    const bool kok = k0 < a.cin;  // whole 16-byte units are valid or not (cin % kEPU == 0)
#pragma unroll
                    for (int j = 0; j < kRPT; ++j)
#pragma unroll
                        for (int h = 0; h < NV; ++h) {
                            const float4 t = lds128f(sm + CF::raw + rs * CF::kRawStageBytes +
                                                     (uint32_t)(rb + kRowStride * j) * (P::kEPC * 4) + (u * NV + h) * 16);
                            v[j][h] = kok ? t : make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(raw_empty + rs);  // the values are in registers
                    ++raw_read;
                } else {
#pragma unroll
                    for (int j = 0; j < kRPT; ++j)
#pragma unroll
                        for (int h = 0; h < NV; ++h)
                            v[j][h] = TT_DBG(a, 16) ? make_float4(1.f, 2.f, 3.f, (float)c)
                                                   : lds128f(raw_slot(rs, j * NV + h));
                }
                if (MODE == SRC_GATHER && c == 0 && u < kRPT && !TT_DBG(a, 16)) {
                    e0 = lds32f(raw_xyz(rs, 0)); e1 = lds32f(raw_xyz(rs, 1)); e2 = lds32f(raw_xyz(rs, 2));
                    if (a.new_xyz != nullptr) {
                        e0 = __fsub_rn(e0, lds32f(raw_xyz(rs, 3)));
                        e1 = __fsub_rn(e1, lds32f(raw_xyz(rs, 4)));
                        e2 = __fsub_rn(e2, lds32f(raw_xyz(rs, 5)));
                    }
                }
#pragma unroll
                for (int h = 0; h < NV; ++h) {
                    float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (has_act) {
                        sc = lds128f(sm + SmemLayout::scale + 4 * ((k0 & (kMaxAct - 1)) + 4 * h));
                        sh = lds128f(sm + SmemLayout::shift + 4 * ((k0 & (kMaxAct - 1)) + 4 * h));
                    }
#pragma unroll
                    for (int j = 0; j < kRPT; ++j) {
                        const float4 w = v[j][h];
                        if (has_act && PREC == PREC_F16) {
                            // BatchNorm as two packed FMAs; the ReLU happens inside the fp16 conversions below
                            unpack2(fma2(pack2(w.x, w.y), pack2(sc.x, sc.y), pack2(sh.x, sh.y)), o[j][4 * h + 0], o[j][4 * h + 1]);
                            unpack2(fma2(pack2(w.z, w.w), pack2(sc.z, sc.w), pack2(sh.z, sh.w)), o[j][4 * h + 2], o[j][4 * h + 3]);
                        } else if (has_act) {
                            o[j][4 * h + 0] = fmaxf(fmaf(w.x, sc.x, sh.x), 0.f);
                            o[j][4 * h + 1] = fmaxf(fmaf(w.y, sc.y, sh.y), 0.f);
                            o[j][4 * h + 2] = fmaxf(fmaf(w.z, sc.z, sh.z), 0.f);
                            o[j][4 * h + 3] = fmaxf(fmaf(w.w, sc.w, sh.w), 0.f);
                        } else {
                            o[j][4 * h + 0] = w.x; o[j][4 * h + 1] = w.y;
                            o[j][4 * h + 2] = w.z; o[j][4 * h + 3] = w.w;
                        }
                    }
                }
            } else {  // SRC_POINTMLP
                if (c == 0) {
                    // pt1 holds this tile's point of row ptid (issued one tile ago): publish it, pick up
                    // the rows this thread expands, then refill pt1 / idx1 for the next tiles
                    const uint32_t tab = sm + SmemLayout::xyz + 16u * (uint32_t)((ptl & 1) * kTile);
                    if (ptid < kTile)
                        sts128f(tab + 16u * (uint32_t)ptid,
                                make_float4(__fsub_rn(px1, cx1), __fsub_rn(py1, cy1), __fsub_rn(pz1, cz1), 0.f));
                    named_bar_sync(1, kProdThreads);
#pragma unroll
                    for (int j = 0; j < kRPT; ++j) p_cur[j] = lds128f(tab + 16u * (uint32_t)(rb + kRowStride * j));
                    if (ptid < kTile) {
                        fetch_pt1();
                        const long long t2 = tile + 2LL * gm;
                        fetch_idx1(t2 < tiles_m ? t2 : tile);
                    }
                }
                float4 fo[P::kEPU];   // all loads in flight before the first FMA: one exposed shared-memory latency
#pragma unroll
                for (int q = 0; q < P::kEPU; ++q) fo[q] = lds128f(sm + SmemLayout::fold + 16 * ((k0 + q) & (kMaxFold - 1)));
#pragma unroll
                for (int q = 0; q < P::kEPU; ++q) {
                    const float4 f = fo[q];
#pragma unroll
                    for (int j = 0; j < kRPT; ++j) {
                        const float4 p = p_cur[j];
                        const float t = fmaf(f.x, p.x, fmaf(f.y, p.y, fmaf(f.z, p.z, f.w)));
                        o[j][q] = PREC == PREC_F16 ? t : fmaxf(t, 0.f);   // fp16: ReLU inside the conversions
                    }
                }
            }
            if (ptid == 0 && c == 0) TT_TCLK(a, ptl, 0);   // producer: first chunk of the tile computed
            mbar_wait(x_empty + s, ((it / kStages) & 1) ^ 1);
            const uint32_t stage = sm + SmemLayout::ring + s * kStageBytes;
            if (WMODE == 1 && ptid == 0) {
                // stream this chunk of W (hi + lo blocks, 32 KiB contiguous in the image)
                mbar_expect_tx(x_full + s, kXBytes);
                bulk_g2s(smem + SmemLayout::ring + s * kStageBytes + kXBytes,
                         wimg + ((size_t)tile_n * KC + c) * kXBytes, kXBytes, x_full + s);
            }
            if (MODE == SRC_GATHER && c == 0) {
                if (u < kRPT)
                    sts128f(sm + SmemLayout::xyz + 16 * ((ptl & 3) * kTile + rb + kRowStride * u), make_float4(e0, e1, e2, 0.f));
                __syncwarp();
                if (lane == 0) mbar_arrive(xyz_full + (ptl & 3));
            }
#pragma unroll
            for (int j = 0; j < kRPT; ++j) {
                uint4 hi, lo;
                if (PREC == PREC_TF32) {
                    float h[4], l[4];
#pragma unroll
                    for (int q = 0; q < 4; ++q) split_tf32(o[j][q], h[q], l[q]);
                    hi = make_uint4(__float_as_uint(h[0]), __float_as_uint(h[1]), __float_as_uint(h[2]), __float_as_uint(h[3]));
                    lo = make_uint4(__float_as_uint(l[0]), __float_as_uint(l[1]), __float_as_uint(l[2]), __float_as_uint(l[3]));
                } else {
                    split_relu_f16x2(o[j][0], o[j][1], hi.x, lo.x);
                    split_relu_f16x2(o[j][2], o[j][3], hi.y, lo.y);
                    split_relu_f16x2(o[j][4 % P::kEPU], o[j][5 % P::kEPU], hi.z, lo.z);
                    split_relu_f16x2(o[j][6 % P::kEPU], o[j][7 % P::kEPU], hi.w, lo.w);
                }
                const uint32_t off = (uint32_t)(rb + kRowStride * j) * 128u + swz;
                sts128(stage + off, hi);
                sts128(stage + kHalfBytes + off, lo);
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(x_full + s);
            if (TT_DBG(a, 32)) __nanosleep(1500);  // triage: producers step aside after every chunk
            if (ptid == 0 && c == KC - 1) TT_TCLK(a, ptl, 1);  // producer: last chunk of the tile published
            if (ptid == 0 && it == 0) TT_CLK(a, 3);
            if (ptid == 0 && it == 0) TT_GCLK(a, 10);  // first operand stage published
            if (ptid == 0 && it == 1) TT_CLK(a, 12);
            if (ptid == 0) TT_CLK(a, 10);
            ++it;
            if (c == KC - 1) ++ptl;
        };

        auto advance = [&](long long &t, int &ci) {
            if (++ci == KC) { ci = 0; t += gm; }
        };
        const long long tile0 = mi;
        if (TT_DBG(a, 1)) {
            // triage: ring protocol only
            for (long long t = tile0; t < tiles_m; t += gm)
                for (int cq = 0; cq < KC; ++cq) {
                    const uint32_t s = it % kStages;
                    mbar_wait(x_empty + s, ((it / kStages) & 1) ^ 1);
                    if (WMODE == 1 && ptid == 0) mbar_arrive(x_full + s);
                    if (MODE == SRC_GATHER && cq == 0) { __syncwarp(); if (lane == 0) mbar_arrive(xyz_full + (ptl & 3)); }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(x_full + s);
                    ++it;
                    if (cq == KC - 1) ++ptl;
                }
        } else if (MODE == SRC_PLAIN && a.ximg_on) {
            // the activations were converted once by prep_ximg_kernel: every chunk is ONE bulk copy of the
            // pre-split [hi | lo] block (plus the W chunk when W is streamed); no producer arithmetic at all
            const uint8_t *ximg = reinterpret_cast<const uint8_t *>(a.ximg);
            for (long long t = tile0; t < tiles_m; t += gm)
                for (int cq = 0; cq < KC; ++cq) {
                    const uint32_t s = it % kStages;
                    mbar_wait(x_empty + s, ((it / kStages) & 1) ^ 1);
                    if (ptid == 0) {
                        uint8_t *stage_p = smem + SmemLayout::ring + s * kStageBytes;
                        if (WMODE == 1) {
                            mbar_expect_tx(x_full + s, 2 * kXBytes);   // the issuer's arrival is part of the count
                            bulk_g2s(stage_p + kXBytes, wimg + ((size_t)tile_n * KC + cq) * kXBytes, kXBytes, x_full + s);
                        } else {
                            mbar_expect_tx_only(x_full + s, kXBytes);
                        }
                        bulk_g2s(stage_p, ximg + ((size_t)t * KC + cq) * kXBytes, kXBytes, x_full + s);
                    }
                    __syncwarp();
                    if (lane == 0) mbar_arrive(x_full + s);
                    if (ptid == 0 && it == 0) TT_GCLK(a, 10);
                    ++it;
                    if (cq == KC - 1) ++ptl;
                }
        } else if (MODE == SRC_POINTMLP) {
            if (tile0 < tiles_m && ptid < kTile) {
                fetch_idx1(tile0);                      // idx of the first tile
                fetch_pt1();                            // -> pt1 = point of the first tile
                fetch_idx1(tile0 + gm < tiles_m ? tile0 + gm : tile0);
            }
            for (long long t = tile0; t < tiles_m; t += gm)
                for (int cq = 0; cq < KC; ++cq) process(t, cq, 0);
        } else {
            constexpr int R = CF::kRawStages > 0 ? CF::kRawStages : 1;
            if (MODE == SRC_GATHER && tile0 < tiles_m) fetch_idx(tile0);
            long long ti = tile0, tp = tile0;  // issue cursor, process cursor
            int ci = 0, cp = 0;
            for (int n = 0; n < R - 1; ++n) {
                if (ti < tiles_m) { issue(ti, ci, n); advance(ti, ci); }
                cp_async_commit();
            }
            if (MODE == SRC_PLAIN && a.in_fix != nullptr) {
                // deferred finalisation of the producing layer (TtArgs::in_fix): words -> scale / shift table
                const unsigned long long *fx = a.in_fix;
                const bool fixed = __ldcg(fx + (size_t)4 * a.cin) == 0ull;
                for (int k = ptid; k < kMaxAct; k += kProdThreads) {
                    float sc = 0.f, sh = 0.f;
                    if (k < a.cin) {
                        double sum, sq;
                        if (fixed) {
                            const long long si = (long long)__ldcg(fx + k), sf = (long long)__ldcg(fx + (size_t)a.cin + k);
                            const long long qi = (long long)__ldcg(fx + (size_t)2 * a.cin + k);
                            const long long qf = (long long)__ldcg(fx + (size_t)3 * a.cin + k);
                            sum = (double)si + (double)sf * 0x1p-54;
                            sq = (double)qi + (double)qf * 0x1p-54;
                        } else {   // a sum left the fixed-point range (or is not finite): the partial rows, in order
                            sum = 0.0; sq = 0.0;
                            for (long long r = 0; r < a.in_partial_rows; ++r) {
                                sum += __ldcg(a.in_partial + (r * 2 + 0) * a.cin + k);
                                sq += __ldcg(a.in_partial + (r * 2 + 1) * a.cin + k);
                            }
                        }
                        const float g = a.in_gamma ? __ldg(a.in_gamma + k) : 1.f, b = a.in_beta ? __ldg(a.in_beta + k) : 0.f;
                        const float cs = a.in_cs ? f16_colscale_sq(g, b, a.cs_sqrt_count) : 1.f;
                        float mean, var;
                        bn_from_sums(sum, sq, a.in_inv_count, g, b, a.in_eps, cs, sc, sh, mean, var);
                        if (blockIdx.x == 0) {
                            if (a.in_mean_out) a.in_mean_out[k] = mean;
                            if (a.in_var_out) a.in_var_out[k] = var;
                        }
                    }
                    s_scale[k] = sc;
                    s_shift[k] = sh;
                }
                named_bar_sync(1, kProdThreads);
            } else if (MODE == SRC_PLAIN) {
                for (int k = ptid; k < kMaxAct; k += kProdThreads) {
                    s_scale[k] = (has_act && k < a.cin) ? a.in_scale[k] : 0.f;
                    s_shift[k] = (has_act && k < a.cin) ? a.in_shift[k] : 0.f;
                }
                named_bar_sync(1, kProdThreads);
            }
            // The next chunk's copies are issued AFTER this chunk has been processed: process() ends in
            // fence.proxy.async (a CTA-scope membar underneath), which waits for the thread's outstanding
            // cp.async traffic -- with freshly issued copies in flight every fence would cost a full
            // global-memory round trip.  Issued here they have a whole chunk period to land.
            uint32_t i = 0;
            while (tp < tiles_m) {
                cp_async_wait<(R >= 2 ? R - 2 : 0)>();  // chunk i has landed (this thread's own copies)
                process(tp, cp, (int)(i % R));
                advance(tp, cp);
                if (ti < tiles_m) { issue(ti, ci, (int)((i + R - 1) % R)); advance(ti, ci); }
                cp_async_commit();              // (possibly empty) group: keeps the group count uniform
                ++i;
            }
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (tid == 0) TT_CLK(a, 7);
    if (tid == 0) TT_GCLK(a, 5);   // all tiles done
    if (warp == kMmaWarp) tmem_dealloc<kTmemCols>(tmem_base);

    // ---- fused BatchNorm finalisation: the last CTA reduces the partial rows in fixed order
#ifdef PAPC_TT_EXPERIMENT
    if (a.dbg & 1024) return;   // timing experiment (stale scale / shift from an earlier launch): no finalisation at all
#endif
    if (a.counter != nullptr) {
        if (tid == 0) {
            __threadfence();
            *s_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
        }
        __syncthreads();
        if (*s_last != 0u) {
            __threadfence();
#ifdef PAPC_TT_TRIAGE
            if (tid == 0 && a.clk != nullptr) a.clk[32 + 7] = clock64();
#endif
            if (tid == 0) TT_GCLK(a, 7);   // this CTA is the last one: finalisation starts
            // gm partial rows x 2*cout doubles (sum | sum^2 per channel), read as 16-byte column pairs:
            // up to 512 pairs per pass, the rows split over S = 512 / pairs thread slices so that a
            // thread's loads are all independent and few (fixed order -> deterministic); slices are
            // combined through shared memory.  This is a serial tail of the kernel (one CTA works,
            // the GPU waits), so it is arranged for the fewest dependent L2 round trips.
            // Measured (profiles/r01_layer_boundary_timeline.txt): 5-7 us per layer whatever the row count.
            // Two rewrites were timed on the B200 and dropped because the step did not move: a two-level
            // tree (last CTA of every 16-row group folds its group, the last of those folds the groups)
            // and TMA bulk staging of the rows into the idle operand ring.
            double2 *red = reinterpret_cast<double2 *>(smem + SmemLayout::ring);            // [512]
            double *all = reinterpret_cast<double *>(smem + SmemLayout::ring + kXBytes);    // [2*cout]
            const int npairs = a.cout;  // 2*cout doubles
            const int P = npairs < 512 ? npairs : 512;
            const int S = 512 / P;
            const int pair_l = tid % P, sl = tid / P;
            const double2 *part = reinterpret_cast<const double2 *>(a.stats_partial);
            // Fast path: the CTAs' sums were also added into fix_acc with integer atomics; 4*cout words to
            // read instead of gm rows of 2*cout doubles (303 KB through one SM's L2 port: 5 us).
            bool fixed = false;
            if (a.fix_acc != nullptr && (a.dbg & 256) == 0) {   // PAPC_TT_DBG=256: A/B switch, partial rows
                fixed = __ldcg(a.fix_acc + (size_t)4 * a.cout) == 0ull;
                __syncthreads();   // everyone has read the flag before it is cleaned
                for (int i = tid; i < 2 * a.cout; i += kThreads) {
                    const int which = i / a.cout, ch = i - which * a.cout;   // 0: sum, 1: sum of squares
                    unsigned long long *pi = a.fix_acc + (size_t)(2 * which) * a.cout + ch;
                    unsigned long long *pf = a.fix_acc + (size_t)(2 * which + 1) * a.cout + ch;
                    const long long vi = (long long)__ldcg(pi), vf = (long long)__ldcg(pf);
                    if (fixed) all[i] = (double)vi + (double)vf * 0x1p-54;
                    *pi = 0ull;   // self-cleaning for the next launch
                    *pf = 0ull;
                }
                if (tid == 0) a.fix_acc[(size_t)4 * a.cout] = 0ull;
                __syncthreads();
            }
            for (int base = fixed ? npairs : 0; base < npairs; base += P) {
                const int pr = base + pair_l;
                if (sl < S && pr < npairs) {
                    double ax = 0.0, ay = 0.0;
                    const double2 *p = part + pr;
                    // all loads of a thread are independent: issue them 16 at a time (each batch is one
                    // L2 round trip of this serial tail), two accumulator pairs to halve the add chain
                    double bx = 0.0, by = 0.0;
                    int r = sl;
                    for (; r + 15 * S < gm; r += 16 * S) {
                        double2 v[16];
#pragma unroll
                        for (int q = 0; q < 16; ++q) v[q] = __ldcg(p + (long long)(r + q * S) * npairs);
#pragma unroll
                        for (int q = 0; q < 16; q += 2) {
                            ax += v[q].x; ay += v[q].y;
                            bx += v[q + 1].x; by += v[q + 1].y;
                        }
                    }
#pragma unroll 8
                    for (; r < gm; r += S) {
                        const double2 v = __ldcg(p + (long long)r * npairs);
                        ax += v.x;
                        ay += v.y;
                    }
                    ax += bx;
                    ay += by;
                    red[sl * P + pair_l] = make_double2(ax, ay);
                }
                if (tid == 0 && base == 0) TT_GCLK(a, 12);  // thread 0: its loads and adds done
                __syncthreads();
                if (tid == 0 && base == 0) TT_GCLK(a, 14);  // every thread's loads and adds done
                if (sl == 0 && pr < npairs) {
                    double ax = 0.0, ay = 0.0;
                    for (int q = 0; q < S; ++q) {
                        const double2 v = red[q * P + pair_l];
                        ax += v.x;
                        ay += v.y;
                    }
                    all[2 * pr] = ax;
                    all[2 * pr + 1] = ay;
                }
                __syncthreads();
            }
#ifdef PAPC_TT_TRIAGE
            if (tid == 0 && a.clk != nullptr) a.clk[32 + 10] = clock64();
#endif
            if (tid == 0) TT_GCLK(a, 8);   // last CTA: partial rows reduced
            // fp64 division and square root are long software sequences on a slow pipe and this is the
            // kernel's serial tail: reciprocal of the count from the host, 1/sqrt by two Newton steps
            // from the fp32 estimate (full double accuracy)
            const double inv_count = a.inv_count;
            for (int ch = tid; ch < a.cout; ch += kThreads) {
                const float g = a.gamma ? a.gamma[ch] : 1.f, b = a.beta ? a.beta[ch] : 0.f;
                float cs = 1.f;
                if (a.out_colscale != nullptr) {
                    cs = f16_colscale_sq(g, b, a.sqrt_count);
                    a.out_colscale[ch] = cs;
                }
                float sc, sh, mean, var;
                bn_from_sums(all[ch], all[a.cout + ch], inv_count, g, b, a.eps, cs, sc, sh, mean, var);
                a.scale[ch] = sc;
                a.shift[ch] = sh;
                if (a.mean_out) a.mean_out[ch] = mean;
                if (a.var_out) a.var_out[ch] = var;
            }
            if (tid == 0) *a.counter = 0u;  // self-cleaning for the next launch
            if (tid == 0) TT_GCLK(a, 9);   // last CTA: scale / shift written (this thread's share)
#ifdef PAPC_TT_TRIAGE
            __syncthreads();
            if (tid == 0 && a.clk != nullptr) { a.clk[32 + 8] = clock64(); a.clk[32 + 9] = blockIdx.x; }
#endif
        }
        if (tid == 0) TT_CLK(a, 8);
    }
    if (tid == 0) TT_GCLK(a, 6);       // exit (after the finalisation in the last CTA)
#ifdef PAPC_TT_TRIAGE
    if (tid == 0 && a.clk != nullptr) {
        unsigned long long gt1;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
        atomicMax(a.clk + 45, gt1);
        atomicMax(a.clk + 46, gt1 - gt0);
        atomicAdd(a.clk + 47, gt1 - gt0);
    }
#endif
}

// ------------------------------------------------------------------------------ point moments
constexpr int kMomThreads = 256;

__global__ void __launch_bounds__(kMomThreads)
point_moments_kernel(const MomentArgs a) {
    __shared__ double s_red[kMomThreads / 32][9];
    __shared__ double s_tot[9];
    __shared__ uint32_t s_islast;
    const int tid = threadIdx.x;
    const bool batch = a.running_mean == nullptr;
    if (batch && a.pre_partial != nullptr) {
        // the nine sums were accumulated by the fused sampling kernel: fixed-order reduction of its rows
        __shared__ double s_sl[28][9];
        const int q = tid % 9, sl = tid / 9;
        if (sl < 28) {
            double v = 0.0;
            for (int b = sl; b < a.pre_rows; b += 28) v += __ldcg(a.pre_partial + (long long)b * 9 + q);
            s_sl[sl][q] = v;
        }
        __syncthreads();
        if (tid < 9) {
            double v = 0.0;
#pragma unroll
            for (int k = 0; k < 28; ++k) v += s_sl[k][tid];
            s_tot[tid] = v / (double)a.M;
        }
        __syncthreads();
    } else if (batch) {
        float acc[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] = 0.f;
        double dacc[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) dacc[i] = 0.0;
        // four rows per trip, their dependent load chains (index -> point, centroid) issued together;
        // the fp32 run length stays bounded (flush to fp64 every 4 trips = 16 rows)
        constexpr int R = 4;
        int trips = 0;
        auto accumulate = [&](float x, float y, float z) {
            acc[0] += x; acc[1] += y; acc[2] += z;
            acc[3] = fmaf(x, x, acc[3]); acc[4] = fmaf(x, y, acc[4]); acc[5] = fmaf(x, z, acc[5]);
            acc[6] = fmaf(y, y, acc[6]); acc[7] = fmaf(y, z, acc[7]); acc[8] = fmaf(z, z, acc[8]);
        };
        auto flush = [&]() {
            if (++trips == 4) {
#pragma unroll
                for (int i = 0; i < 9; ++i) { dacc[i] += (double)acc[i]; acc[i] = 0.f; }
                trips = 0;
            }
        };
        if (a.nslice > 0) {
            // Staged variant: a block works on ONE cloud.  Its points (and centroids) sit in shared
            // memory, so the random 12-byte gathers -- the cost of this kernel: 3.7 M scattered
            // loads through L1 -- become LDS; only the (coalesced) index reads go to global memory.
            extern __shared__ float s_pts[];               // [N*3] xyz, then [S*3] new_xyz
            const int b = blockIdx.x / a.nslice, sl = blockIdx.x - b * a.nslice;
            const float *cloud = a.xyz + (long long)b * a.N * 3;
            for (int i = tid; i < a.N * 3; i += kMomThreads) s_pts[i] = __ldg(cloud + i);
            float *s_ctr = s_pts + a.N * 3;
            if (a.new_xyz != nullptr) {
                const float *ctr = a.new_xyz + (long long)b * a.S * 3;
                for (int i = tid; i < a.S * 3; i += kMomThreads) s_ctr[i] = __ldg(ctr + i);
            }
            __syncthreads();
            const int rows = a.S * a.K;                    // rows of this cloud
            const int per = (rows + a.nslice - 1) / a.nslice;
            const int r_end = min(rows, (sl + 1) * per);
            const long long row_base = (long long)b * rows;
            for (int r0 = sl * per + tid; r0 < r_end; r0 += R * kMomThreads) {
                int n[R];
                int g[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int rl = r0 + r * kMomThreads;
                    const bool ok = rl < r_end;
                    g[r] = ok ? (int)fastdiv((uint32_t)rl, a.kmul, a.kshr) : -1;
                    n[r] = !ok ? 0 : (a.idx != nullptr ? __ldg(a.idx + row_base + rl) : rl - g[r] * a.K);
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (g[r] < 0) continue;
                    const int nn = min(max(n[r], 0), a.N - 1);
                    float x = s_pts[nn * 3], y = s_pts[nn * 3 + 1], z = s_pts[nn * 3 + 2];
                    if (a.new_xyz != nullptr) {
                        x = __fsub_rn(x, s_ctr[g[r] * 3]);
                        y = __fsub_rn(y, s_ctr[g[r] * 3 + 1]);
                        z = __fsub_rn(z, s_ctr[g[r] * 3 + 2]);
                    }
                    accumulate(x, y, z);
                }
                flush();
            }
        } else {
            const long long stride = (long long)gridDim.x * kMomThreads;
            for (long long row0 = (long long)blockIdx.x * kMomThreads + tid; row0 < a.M; row0 += R * stride) {
                int n[R];
                RowGeom rg[R];
                bool ok[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const long long row = row0 + r * stride;
                    ok[r] = row < a.M;
                    rg[r] = row_geom(ok[r] ? row : 0, a);
                    n[r] = (a.idx != nullptr && ok[r]) ? __ldg(a.idx + row) : rg[r].k;
                }
                float px[R], py[R], pz[R], cx[R], cy[R], cz[R];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const int nn = min(max(n[r], 0), a.N - 1);
                    const float *q = a.xyz + (rg[r].bN + nn) * 3;
                    px[r] = __ldg(q); py[r] = __ldg(q + 1); pz[r] = __ldg(q + 2);
                    cx[r] = cy[r] = cz[r] = 0.f;
                    if (a.new_xyz != nullptr) {
                        const float *cc = a.new_xyz + (long long)rg[r].g * 3;
                        cx[r] = __ldg(cc); cy[r] = __ldg(cc + 1); cz[r] = __ldg(cc + 2);
                    }
                }
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    if (!ok[r]) continue;
                    accumulate(a.new_xyz != nullptr ? __fsub_rn(px[r], cx[r]) : px[r],
                               a.new_xyz != nullptr ? __fsub_rn(py[r], cy[r]) : py[r],
                               a.new_xyz != nullptr ? __fsub_rn(pz[r], cz[r]) : pz[r]);
                }
                flush();
            }
        }
#pragma unroll
        for (int i = 0; i < 9; ++i) {
            double v = dacc[i] + (double)acc[i];
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((tid & 31) == 0) s_red[tid >> 5][i] = v;
        }
        __syncthreads();
        if (tid < 9) {
            double v = 0.0;
#pragma unroll
            for (int w = 0; w < kMomThreads / 32; ++w) v += s_red[w][tid];
            a.partial[(long long)blockIdx.x * 9 + tid] = v;
            __threadfence();
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            s_islast = (atomicAdd(a.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
        }
        __syncthreads();
        if (s_islast == 0u) return;
        __threadfence();
        {
            // 9 quantities x 28 slices of the block rows in parallel (fixed order -> deterministic):
            // nine threads walking ~300 rows one load at a time was most of this kernel's time
            __shared__ double s_sl[28][9];
            const int q = tid % 9, sl = tid / 9;
            if (sl < 28) {
                double v = 0.0;
                for (int b = sl; b < (int)gridDim.x; b += 28) v += __ldcg(a.partial + (long long)b * 9 + q);
                s_sl[sl][q] = v;
            }
            __syncthreads();
            if (tid < 9) {
                double v = 0.0;
#pragma unroll
                for (int k = 0; k < 28; ++k) v += s_sl[k][tid];
                s_tot[tid] = v / (double)a.M;
            }
        }
        if (tid == 0) *a.counter = 0u;
        __syncthreads();
    }
    // ---- statistics of y0 = W0 p + b0 per channel, scale / shift, folded first layer
    for (int c = tid; c < a.c0; c += kMomThreads) {
        const double w0 = a.W0[c * 3 + 0], w1 = a.W0[c * 3 + 1], w2 = a.W0[c * 3 + 2];
        const double b0 = a.b0 ? (double)a.b0[c] : 0.0;
        double mean, var;
        if (batch) {
            const double mx = s_tot[0], my = s_tot[1], mz = s_tot[2];
            const double cxx = s_tot[3] - mx * mx, cxy = s_tot[4] - mx * my, cxz = s_tot[5] - mx * mz;
            const double cyy = s_tot[6] - my * my, cyz = s_tot[7] - my * mz, czz = s_tot[8] - mz * mz;
            mean = w0 * mx + w1 * my + w2 * mz + b0;
            var = w0 * w0 * cxx + w1 * w1 * cyy + w2 * w2 * czz +
                  2.0 * (w0 * w1 * cxy + w0 * w2 * cxz + w1 * w2 * cyz);
            var = var > 0.0 ? var : 0.0;
            if (a.mean_out) a.mean_out[c] = (float)mean;
            if (a.var_out) a.var_out[c] = (float)var;
        } else {
            mean = (double)a.running_mean[c];
            var = (double)a.running_var[c];
        }
        const double g = a.gamma ? (double)a.gamma[c] : 1.0;
        const double be = a.beta ? (double)a.beta[c] : 0.0;
        const double sc = g / sqrt(var + (double)a.eps);
        const double sh = be - mean * sc;
        a.scale[c] = (float)sc;
        a.shift[c] = (float)sh;
        double cs = 1.0;  // power of two keeping relu(bn(y0)) / cs below 2^15 (fp16 operands)
        if (a.out_colscale != nullptr) {
            const float csf = f16_colscale_sq(a.gamma ? a.gamma[c] : 1.f, a.beta ? a.beta[c] : 0.f, a.sqrt_M);
            a.out_colscale[c] = csf;
            cs = (double)csf;
        }
        reinterpret_cast<float4 *>(a.l0_fold)[c] =
            make_float4((float)(sc * w0 / cs), (float)(sc * w1 / cs), (float)(sc * w2 / cs),
                        (float)((sc * b0 + sh) / cs));
    }
}

// ------------------------------------------------------------------------------------ host side
static int tmem_k(int prec) { return prec == PREC_F16 ? Prec<PREC_F16>::kTmemK : Prec<PREC_TF32>::kTmemK; }
static int epc(int prec) { return prec == PREC_F16 ? Prec<PREC_F16>::kEPC : Prec<PREC_TF32>::kEPC; }
static int epu(int prec) { return prec == PREC_F16 ? Prec<PREC_F16>::kEPU : Prec<PREC_TF32>::kEPU; }

void make_fastdiv(uint32_t d, uint32_t *mul, uint32_t *shr) {
    if (d <= 1u) { *mul = 0u; *shr = 0u; return; }
    uint32_t lg = 0;
    while ((1ull << lg) < d) ++lg;  // ceil(log2(d))
    const uint32_t p = 31u + lg;
    *mul = (uint32_t)(((1ull << p) + d - 1ull) / d);
    *shr = p - 32u;
}

size_t ximg_bytes(long long M, int cin, int cout) {
    // worth it when a row tile feeds several channel tiles and there are few row tiles (small-M layers: today
    // every channel tile's CTA converts the same activation rows again)
    const long long tiles_m = ceil_div<long long>(M, kTile);
    const int nt = ceil_div(cout, kTile);
    if (M <= 0 || cin < 8 || cin % 8 != 0 || nt < 2 || tiles_m > 128) return 0;
    return (size_t)tiles_m * ceil_div(cin, Prec<PREC_F16>::kEPC) * kXBytes;
}

bool eligible(const TtProblem &p) {
    if (p.prec != PREC_TF32 && p.prec != PREC_F16) return false;
    if (p.cout < 1 || p.cout > 8 * kTile) return false;
    if (p.cin < epu(p.prec) || p.cin % epu(p.prec) != 0) return false;
    if (p.mode == SRC_POINTMLP) {
        if (p.cin > kMaxFold) return false;
    } else if (p.cin > kMaxAct && !(p.mode == SRC_PLAIN && p.no_act)) {
        return false;  // the scale / shift / column-scale tables hold kMaxAct channels
    }
    if (p.mode == SRC_GATHER && p.D != p.cin) return false;
    if (p.pool && !(p.K == 32 || p.K == 64 || p.K == 128)) return false;
    return true;
}

size_t wimg_bytes(int prec, int cin, int cout) {
    if (cin <= tmem_k(prec)) return 0;
    return (size_t)ceil_div(cout, kTile) * ceil_div(cin, epc(prec)) * kXBytes;
}

template <int MODE, int PREC, int WMODE, bool POOL, bool PAIR = false>
static int launch_inst(const TtArgs &a, int grid, cudaStream_t st) {
    auto k = mlp_layer_tt_kernel<MODE, PREC, WMODE, POOL, PAIR>;
    static bool configured = false;
    if (!configured) {
        PAPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)Cfg<MODE, PREC, WMODE, PAIR>::bytes));
        configured = true;
    }
    char name[56];
    snprintf(name, sizeof(name), "mlp_tt<%s,%s,%s%s%s>", MODE == SRC_PLAIN ? "plain" : MODE == SRC_GATHER ? "gather" : "pointmlp",
             PREC == PREC_F16 ? "f16x3" : "tf32x3", WMODE ? "Wstream" : "Wtmem", POOL ? ",pool" : "", PAIR ? ",pair" : "");
    // algorithmic bytes: the activation rows read (gathered rows count once per row read), the
    // pre-BN output written (if any) and the pooled extrema
    const double in_b = MODE == SRC_POINTMLP ? 16.0 * a.M : 4.0 * (double)a.M * a.cin;
    const double out_b = (a.y ? 4.0 * (double)a.M * a.cout : 0.0) +
                         (POOL ? 8.0 * (double)(a.M / (a.K > 0 ? a.K : 1)) * a.cout : 0.0);
    ProfScope prof(st, name, a.M, a.cin, a.cout, 2.0 * (double)a.M * a.cin * a.cout, in_b + out_b);
    if (a.pdl) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(kThreads);
        cfg.dynamicSmemBytes = Cfg<MODE, PREC, WMODE, PAIR>::bytes;
        cfg.stream = st;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        ++g_launch_count;
        PAPC_CUDA_TRY(cudaLaunchKernelEx(&cfg, k, a));
        return PAPC_OK;
    }
    k<<<grid, kThreads, Cfg<MODE, PREC, WMODE, PAIR>::bytes, st>>>(a);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

template <int MODE, int PREC>
static int launch_mp(const TtArgs &a, bool streamed, bool pool, int grid, cudaStream_t st) {
    if (streamed) {
        if (MODE == SRC_POINTMLP) return PAPC_EUNSUPPORTED;
        constexpr int M1 = MODE == SRC_POINTMLP ? SRC_PLAIN : MODE;  // never instantiated for POINTMLP
        return pool ? launch_inst<M1, PREC, 1, true>(a, grid, st) : launch_inst<M1, PREC, 1, false>(a, grid, st);
    }
    return pool ? launch_inst<MODE, PREC, 0, true>(a, grid, st) : launch_inst<MODE, PREC, 0, false>(a, grid, st);
}

#ifdef PAPC_TT_TRIAGE
// PAPC_TT_GCLK=1: every launch gets a [152][16] slice of one device buffer (zeroed once, 64 launches);
// papc_tt_gclk_dump() writes them out after the run.
namespace {
constexpr int kGclkLaunches = 64, kGclkCtas = 152;
unsigned long long *g_gclk = nullptr;
int g_gclk_n = 0;
struct GclkMeta { int mode, prec, cin, cout; long long M; } g_gclk_meta[kGclkLaunches];
unsigned long long *gclk_slice(const TtArgs &a) {
    if (getenv("PAPC_TT_GCLK") == nullptr || g_gclk_n >= kGclkLaunches) return nullptr;
    if (g_gclk == nullptr) {
        cudaMalloc(&g_gclk, sizeof(unsigned long long) * kGclkLaunches * kGclkCtas * 16);
        cudaMemset(g_gclk, 0, sizeof(unsigned long long) * kGclkLaunches * kGclkCtas * 16);
    }
    g_gclk_meta[g_gclk_n] = GclkMeta{a.mode, a.prec, a.cin, a.cout, a.M};
    return g_gclk + (size_t)(g_gclk_n++) * kGclkCtas * 16;
}
}  // namespace
extern "C" int papc_tt_gclk_dump(const char *path) {
    if (g_gclk == nullptr) return -1;
    cudaDeviceSynchronize();
    const size_t n = (size_t)kGclkLaunches * kGclkCtas * 16;
    unsigned long long *h = new unsigned long long[n];
    cudaMemcpy(h, g_gclk, n * sizeof(unsigned long long), cudaMemcpyDeviceToHost);
    FILE *f = fopen(path, "w");
    if (f == nullptr) { delete[] h; return -2; }
    for (int l = 0; l < g_gclk_n; ++l) {
        const GclkMeta &m = g_gclk_meta[l];
        fprintf(f, "launch %d mode %d prec %d M %lld cin %d cout %d\n", l, m.mode, m.prec, m.M, m.cin, m.cout);
        for (int c = 0; c < kGclkCtas; ++c) {
            const unsigned long long *q = h + ((size_t)l * kGclkCtas + c) * 16;
            if (q[0] == 0) continue;
            fprintf(f, "cta %d", c);
            for (int e = 0; e < 16; ++e) fprintf(f, " %llu", q[e]);
            fprintf(f, "\n");
        }
    }
    fclose(f);
    delete[] h;
    return g_gclk_n;
}
#endif

int launch(const TtArgs &a_in, cudaStream_t st) {
    TtArgs a = a_in;
    a.dbg = 0;
#ifdef PAPC_TT_TRIAGE
    {
        const char *e = getenv("PAPC_TT_DBG");   // role masks: wrong results, triage builds only
        a.dbg = e ? atoi(e) : 0;
    }
#else
    {
        const char *e = getenv("PAPC_TT_DBG");   // release builds honour only the result-preserving A/B bits
#ifdef PAPC_TT_EXPERIMENT
        a.dbg = e ? (atoi(e) & (256 | 512 | 1024)) : 0;
#else
        a.dbg = e ? (atoi(e) & (256 | 512)) : 0;
#endif
    }
#endif
#ifdef PAPC_TT_TRIAGE
    static unsigned long long *d_clk = nullptr;
    const bool want_clk = getenv("PAPC_TT_CLK") != nullptr;
    if (want_clk) {
        if (d_clk == nullptr) cudaMalloc(&d_clk, 304 * sizeof(unsigned long long));
        cudaMemsetAsync(d_clk, 0, 304 * sizeof(unsigned long long), st);
        cudaMemsetAsync(d_clk + 44, 0xff, sizeof(unsigned long long), st);  // atomicMin target
        a.clk = d_clk;
    }
#endif
#ifdef PAPC_TT_TRIAGE
    a.gclk = gclk_slice(a_in);
#endif
    const bool pool = a.pool_max != nullptr;
    if (!pool && a.y == nullptr) return PAPC_EINVAL;
    a.fastgeom = a.M < (1LL << 31) && a.K >= 1 && a.S >= 1;
    make_fastdiv((uint32_t)(a.K >= 1 ? a.K : 1), &a.kmul, &a.kshr);
    make_fastdiv((uint32_t)(a.S >= 1 ? a.S : 1), &a.smul, &a.sshr);
    a.inv_count = a.count > 0.0 ? 1.0 / a.count : 0.0;
    const bool streamed = a.cin > tmem_k(a.prec);
    int nt = ceil_div(a.cout, kTile);
    {
        const char *e = getenv("PAPC_TT_PDL");  // A/B switch: PAPC_TT_PDL=0 disables dependent launch
        // a streamed-W / activation-image launch has its image kernel(s) as stream predecessor: still a dependent
        // launch -- the image kernels release their dependents at once, so this kernel's prologue overlaps them, and
        // griddepcontrol.wait returns only when they (hence everything before them) are complete
        if (streamed) a.pdl = 1;
        if (e && e[0] == '0') a.pdl = 0;
    }
    a.ximg_on = 0;
    {
        static const bool ximg_off = [] { const char *e = getenv("PAPC_TT_XIMG"); return e && e[0] == '0'; }();  // A/B switch
        const size_t xb = ximg_bytes(a.M, a.cin, a.cout);
        const bool act_tab = a.in_scale != nullptr && a.in_shift != nullptr &&
                             (reinterpret_cast<uintptr_t>(a.in_scale) & 15u) == 0 && (reinterpret_cast<uintptr_t>(a.in_shift) & 15u) == 0;
        if (!ximg_off && a.mode == SRC_PLAIN && a.prec == PREC_F16 && (act_tab || a.in_fix != nullptr) &&
            a.ximg != nullptr && xb > 0 && (reinterpret_cast<uintptr_t>(a.x) & 15u) == 0) {
            const int KCx = ceil_div(a.cin, epc(a.prec));
            long long blocks = (ceil_div<long long>(a.M, kTile) * KCx * kTile * 8 + 255) / 256;
            if (blocks > 8LL * kNumSMs) blocks = 8LL * kNumSMs;
            DeferredIn d{};
            if (a.in_fix != nullptr)
                d = DeferredIn{a.in_fix, a.in_partial, a.in_partial_rows, a.in_gamma, a.in_beta, a.in_eps, a.in_inv_count,
                               a.in_cs, a.cs_sqrt_count, a.in_mean_out, a.in_var_out};
            prep_ximg_kernel<<<(unsigned)blocks, 256, 0, st>>>(a.x, a.M, a.cin, a.in_scale, a.in_shift, KCx,
                                                               reinterpret_cast<uint8_t *>(a.ximg), d);
            PAPC_LAUNCH_CHECK();
            a.ximg_on = 1;
            if (a.pdl == 0) { const char *e = getenv("PAPC_TT_PDL"); a.pdl = (e && e[0] == '0') ? 0 : 1; }
        }
    }
    if (streamed) {
        if (a.wimg == nullptr) return PAPC_EWORKSPACE;
        const int KC = ceil_div(a.cin, epc(a.prec));
        long long blocks = ((long long)nt * KC * kTile * 8 + 255) / 256;
        if (blocks > 4LL * kNumSMs) blocks = 4LL * kNumSMs;
        if (a.prec == PREC_F16)
            prep_wimg_kernel<PREC_F16><<<(unsigned)blocks, 256, 0, st>>>(
                a.W, a.wld, a.wk0, a.cin, a.cout, a.w_colscale, a.cs_on, a.cs_gamma, a.cs_beta, a.cs_sqrt_count, KC,
                reinterpret_cast<uint8_t *>(a.wimg));
        else
            prep_wimg_kernel<PREC_TF32><<<(unsigned)blocks, 256, 0, st>>>(
                a.W, a.wld, a.wk0, a.cin, a.cout, a.w_colscale, a.cs_on, a.cs_gamma, a.cs_beta, a.cs_sqrt_count, KC,
                reinterpret_cast<uint8_t *>(a.wimg));
        PAPC_LAUNCH_CHECK();
    }
    const long long tiles_m = ceil_div<long long>(a.M, kTile);
    // paired channel tiles (one CTA converts a row tile once for 256 output channels): plain fp16-split layers
    // whose W pair fits in tensor memory, when the row tiles alone fill the GPU
    bool pair = false;
    {
        static const bool pair_off = [] { const char *e = getenv("PAPC_TT_PAIR"); return e && e[0] == '0'; }();  // A/B switch
        pair = !pair_off && a.mode == SRC_PLAIN && a.prec == PREC_F16 && !streamed && !a.ximg_on && nt >= 2 &&
               a.cin <= tmem_k(PREC_F16) / 2 && tiles_m * ((nt + 1) / 2) >= kNumSMs;
        if (pair) nt = (nt + 1) / 2;
    }
    long long gm = kNumSMs / nt;
    if (gm < 1) gm = 1;
    if (gm > tiles_m) gm = tiles_m;
    if (a.stats_partial != nullptr && gm > a.partial_rows) gm = a.partial_rows;
    const int grid = (int)(gm * nt);
    a.tma2d = 0;
    {
        const char *e = getenv("PAPC_TT_TMA2D");  // A/B switch: PAPC_TT_TMA2D=0 falls back to per-row bulk copies
        if (a.mode == SRC_PLAIN && !(e && e[0] == '0') && a.M < (1LL << 31)) {
            typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                         const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                         CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
            static EncodeFn encode = [] {
                void *fn = nullptr;
                cudaDriverEntryPointQueryResult qres;
                if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess ||
                    qres != cudaDriverEntryPointSuccess)
                    fn = nullptr;
                return reinterpret_cast<EncodeFn>(fn);
            }();
            if (encode != nullptr) {
                const cuuint64_t gdim[2] = {(cuuint64_t)a.cin, (cuuint64_t)a.M};
                const cuuint64_t gstride[1] = {(cuuint64_t)a.cin * 4};
                const cuuint32_t box[2] = {(cuuint32_t)epc(a.prec), (cuuint32_t)kTile};
                const cuuint32_t estr[2] = {1, 1};
                if (encode(&a.xmap, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(a.x), gdim, gstride, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS)
                    a.tma2d = 1;
            }
        }
    }
    int rc = PAPC_EINVAL;
    switch (a.mode) {
        case SRC_PLAIN:
            if (pair)
                rc = pool ? launch_inst<SRC_PLAIN, PREC_F16, 0, true, true>(a, grid, st)
                          : launch_inst<SRC_PLAIN, PREC_F16, 0, false, true>(a, grid, st);
            else
                rc = a.prec == PREC_F16 ? launch_mp<SRC_PLAIN, PREC_F16>(a, streamed, pool, grid, st)
                                        : launch_mp<SRC_PLAIN, PREC_TF32>(a, streamed, pool, grid, st);
            break;
        case SRC_GATHER:
            rc = a.prec == PREC_F16 ? launch_mp<SRC_GATHER, PREC_F16>(a, streamed, pool, grid, st)
                                    : launch_mp<SRC_GATHER, PREC_TF32>(a, streamed, pool, grid, st);
            break;
        case SRC_POINTMLP:
            rc = a.prec == PREC_F16 ? launch_mp<SRC_POINTMLP, PREC_F16>(a, streamed, pool, grid, st)
                                    : launch_mp<SRC_POINTMLP, PREC_TF32>(a, streamed, pool, grid, st);
            break;
    }
#ifdef PAPC_TT_TRIAGE
    if (want_clk && rc == PAPC_OK) {
        unsigned long long h[304];
        cudaStreamSynchronize(st);
        cudaMemcpy(h, d_clk, sizeof(h), cudaMemcpyDeviceToHost);
        auto us = [&](int slot, int e) { return h[slot * 16 + e] ? (double)(h[slot * 16 + e] - h[slot * 16]) / 1965.0 : -1.0; };
        fprintf(stderr, "[tt clk] mode %d prec %d M %lld cin %d cout %d grid %d | CTA0 us: setup %.2f Wstaged %.2f "
                "x_full0 %.2f x_full1 %.2f mma0 %.2f acc0 %.2f acc1 %.2f prod_end %.2f epi_end %.2f teardown %.2f "
                "counted %.2f | last CTA %llu: finalisation %.2f us (reduction part %.2f)\n",
                a.mode, a.prec, a.M, a.cin, a.cout, grid, us(0, 1), us(0, 2), us(0, 3), us(0, 12), us(0, 4), us(0, 5),
                us(0, 11), us(0, 10), us(0, 6), us(0, 7), us(0, 8), h[32 + 9],
                h[32 + 7] ? (double)(h[32 + 8] - h[32 + 7]) / 1965.0 : -1.0,
                h[32 + 10] ? (double)(h[32 + 10] - h[32 + 7]) / 1965.0 : -1.0);
        if (getenv("PAPC_TT_TILECLK") != nullptr) {
            fprintf(stderr, "[tt tile] us rel. CTA0 entry: prod_ready prod_pub | mma_accfree mma_commit | epi_accfull epi_release epi_done\n");
            for (int t = 0; t < 16; ++t) {
                const unsigned long long *q = h + 48 + t * 8;
                if (q[3] == 0) continue;
                auto u = [&](int e) { return q[e] ? (double)(q[e] - h[0]) / 1965.0 : -1.0; };
                fprintf(stderr, "[tt tile] %2d: %7.2f %7.2f | %7.2f %7.2f (done %7.2f) | %7.2f %7.2f %7.2f\n", t + 8, u(0), u(1), u(2),
                        u(3), u(7), u(4), u(5), u(6));
            }
        }
        if (getenv("PAPC_TT_EPICLK") != nullptr) {
            fprintf(stderr, "[tt epi] cycles per phase (warp 0): wait_acc fence ld0 math0 ld1 math1 bookkeeping | tile total\n");
            for (int t = 0; t < 16; ++t) {
                const unsigned long long *q = h + 176 + t * 8;
                if (q[7] == 0 || q[0] == 0) continue;
                fprintf(stderr, "[tt epi] %2d: %6llu %6llu %6llu %6llu %6llu %6llu %6llu | %6llu\n", t + 8, q[1] - q[0], q[2] - q[1],
                        q[3] - q[2], q[4] - q[3], q[5] - q[4], q[6] - q[5], q[7] - q[6], q[7] - q[0]);
            }
        }
        fprintf(stderr, "[tt clk]   globaltimer: first entry -> last exit %.2f us, longest CTA %.2f us, mean CTA %.2f us\n",
                (double)(h[45] - h[44]) / 1e3, (double)h[46] / 1e3, (double)h[47] / 1e3 / grid);
    }
#endif
    return rc;
}

int moment_blocks(long long M) {
    long long b = ceil_div<long long>(M, 1024);
    if (b < 1) b = 1;
    if (b > 2LL * kNumSMs) b = 2LL * kNumSMs;
    return (int)b;
}

int launch_moments(const MomentArgs &a_in, cudaStream_t st) {
    MomentArgs a = a_in;
    a.fastgeom = a.M < (1LL << 31) && a.K >= 1 && a.S >= 1;
    make_fastdiv((uint32_t)(a.K >= 1 ? a.K : 1), &a.kmul, &a.kshr);
    make_fastdiv((uint32_t)(a.S >= 1 ? a.S : 1), &a.smul, &a.sshr);
    int blocks = a.running_mean != nullptr ? 1 : moment_blocks(a.M);
    size_t smem = 0;
    a.nslice = 0;
    if (a.running_mean != nullptr) a.pre_partial = nullptr;
    if (a.pre_partial != nullptr) {
        if (a.pre_rows < 1) return PAPC_EINVAL;
        ProfScope prof(st, "point_moments_finish", a.M, 3, a.c0, 0.0, 72.0 * a.pre_rows);
        point_moments_kernel<<<1, kMomThreads, 0, st>>>(a);
        PAPC_LAUNCH_CHECK();
        return PAPC_OK;
    }
    // one cloud per block, staged in shared memory, when the batch provides enough blocks and the
    // cloud fits; B = M / (S * K)
    const long long rows_per_cloud = (long long)a.S * a.K;
    const long long B = rows_per_cloud > 0 ? a.M / rows_per_cloud : 0;
    const size_t need = ((size_t)a.N + (a.new_xyz ? (size_t)a.S : 0)) * 3 * sizeof(float);
    if (a.running_mean == nullptr && a.fastgeom && B >= 1 && B * rows_per_cloud == a.M && B <= 2 * kNumSMs &&
        need <= 96 * 1024 && rows_per_cloud < (1LL << 31)) {
        int ns = (int)((2 * kNumSMs) / B);
        const int max_ns = (int)ceil_div<long long>(rows_per_cloud, 4 * kMomThreads);
        if (ns > max_ns) ns = max_ns;
        if (ns < 1) ns = 1;
        a.nslice = ns;
        blocks = (int)(B * ns);
        smem = need;
        static bool configured = false;
        if (!configured && smem > 48 * 1024) {
            PAPC_CUDA_TRY(cudaFuncSetAttribute(point_moments_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
            configured = true;
        }
    }
    ProfScope prof(st, "point_moments", a.M, 3, a.c0, 0.0, 16.0 * (double)a.M);
    point_moments_kernel<<<blocks, kMomThreads, smem, st>>>(a);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

}  // namespace tt
}  // namespace papc
