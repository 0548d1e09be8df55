// sa_mlp_tc.cu -- tensor-core (tcgen05 / TMEM) path of the grouped shared MLP layer for sm_100a.
//
//   y[M,cout] = act(x)[M,cin] * W^T[cin,cout] + bias          (one SetAbstraction MLP layer)
//
// fp32 in, fp32 out, within the 1e-5 parity budget: every operand is split into two TF32 terms
// (hi = rna_tf32(v), lo = rna_tf32(v - hi)) and each product is issued as THREE kind::tf32 UMMAs
// (lo*hi + hi*lo + hi*hi) accumulating in fp32 in tensor memory ("3xTF32").
//
// One persistent CTA per SM, warp-specialised:
//   warps 0-3  producers : gather / load the 128-row activation tile from HBM, apply the previous
//                          layer's BatchNorm scale/shift + ReLU, split hi/lo and write both halves
//                          into shared memory in the UMMA K-major SWIZZLE_128B layout (ring of
//                          32-column stages, mbarrier full/empty).
//   warp  8    MMA issuer: one thread issues tcgen05.mma (M=128, N=BN, K=8 per instruction) from the
//                          shared-memory descriptors; W (hi and lo) is resident in shared memory for
//                          the whole kernel, staged once by TMA bulk copies from a pre-swizzled image.
//                          tcgen05.commit frees the activation stage / publishes the accumulator.
//   warps 4-7  epilogue  : tcgen05.ld the accumulator (double-buffered in TMEM so tile i+1's MMAs
//                          overlap tile i's epilogue), + bias, pre-BN store, per-channel sum / sum^2
//                          for the batch statistics (warp transpose-reduce, fp64 across tiles) and,
//                          on the last layer, the per-group max / min (the max-pool commutes with
//                          the monotone BN+ReLU).
#include "common.cuh"
#include "sa_mlp_tc.cuh"

namespace papc {
namespace tc {

constexpr int kBM = 128;           // rows per tile == UMMA M
constexpr int kChunkK = 32;        // fp32 per K chunk == one 128-byte swizzle row
constexpr int kAStageBytes = kBM * 128;  // one (hi or lo) activation chunk
constexpr int kProducerWarps = 4;
constexpr int kEpilogueWarps = 4;
constexpr int kThreads = (kProducerWarps + kEpilogueWarps + 1) * 32;  // 288
constexpr int kMaxStages = 4;

// ------------------------------------------------------------------------------- PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t *bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    return ok != 0;
}
// Bounded wait: a protocol bug traps (launch error) instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    for (uint32_t spin = 0; !mbar_try_wait(bar, parity); ++spin)
        if (spin > (1u << 24)) __trap();
}
__device__ __forceinline__ void bulk_g2s(void *dst_smem, const void *src_gmem, uint32_t bytes,
                                         uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_u32(dst_smem)),
        "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)),
                 "n"(COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(COLS)
                 : "memory");
}
// D[tmem] (+)= A[smem] * B[smem], kind::tf32, cta_group::1
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// mbarrier arrives once every tcgen05 op issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}
// 32 lanes x 32 columns of fp32: thread t of the warp gets row (lane base + t), columns c0..c0+31
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ float tf32_rna(float x) {
    uint32_t r;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
    return __uint_as_float(r);
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// UMMA shared-memory descriptor: K-major operand, SWIZZLE_128B, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);       // start address  [0,14)
    d |= (uint64_t)1 << 16;                       // leading byte offset (unused for SW128 K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // stride byte offset [32,46): 8 rows * 128 B
    d |= (uint64_t)1 << 46;                       // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                       // layout type: SWIZZLE_128B
    return d;
}
// instruction descriptor: D fp32, A/B tf32, both K-major, M = 128, N = BN
__host__ __device__ constexpr uint32_t make_idesc(int n) {
    return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

// physical float offset of element (row r, k in [0,32)) inside a [rows][32] SWIZZLE_128B block
__host__ __device__ __forceinline__ int sw128_off(int r, int k) {
    return r * 32 + ((((k >> 2) ^ (r & 7)) << 2) | (k & 3));
}

// warp transpose-reduce: v[j] of lane l = value (row l, column j); returns for lane l the
// reduction over the 32 rows of column l.  31 shuffles.
template <typename Op>
__device__ __forceinline__ float warp_col_reduce(float (&v)[32], int lane, Op op) {
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
        const bool up = (lane & s) != 0;
#pragma unroll
        for (int j = 0; j < s; ++j) {
            const float mine = up ? v[j + s] : v[j];
            const float other = up ? v[j] : v[j + s];
            v[j] = op(mine, __shfl_xor_sync(0xffffffffu, other, s));
        }
    }
    return v[0];
}

// --------------------------------------------------------------------------------- W image
// img: [nt][KC][2 (hi,lo)][BN][32] floats, each [BN][32] block in the SWIZZLE_128B layout, so a CTA
// stages its column tile with linear TMA bulk copies.  Internal K order of a gathered source is
// [feats 0..D) | xyz 0..3) | zero pad] (16-byte aligned feature loads); W columns are permuted here.
__global__ void __launch_bounds__(256)
prep_w_kernel(const float *__restrict__ W, int cin, int cout, int gather, int D, int order, int BN,
              int KC, float *__restrict__ img) {
    const int nt = ceil_div(cout, BN);
    const size_t total = (size_t)nt * KC * BN * 32;
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; e < total; e += stride) {
        const int kk = (int)(e & 31);
        const int rl = (int)((e >> 5) % BN);
        const int c = (int)((e / (32 * (size_t)BN)) % KC);
        const int tn = (int)(e / (32 * (size_t)BN * KC));
        const int n = tn * BN + rl;
        const int kin = c * 32 + kk;
        int col = -1;
        if (!gather) {
            col = kin < cin ? kin : -1;
        } else if (kin < D) {
            col = order == PAPC_XYZ_FIRST ? kin + 3 : kin;
        } else if (kin < D + 3) {
            col = order == PAPC_XYZ_FIRST ? kin - D : kin;
        }
        float v = 0.f;
        if (n < cout && col >= 0) v = W[(size_t)n * cin + col];
        const float hi = tf32_rna(v);
        const float lo = tf32_rna(v - hi);
        float *blk = img + ((size_t)(tn * KC + c) * 2) * BN * 32;
        blk[sw128_off(rl, kk)] = hi;
        blk[(size_t)BN * 32 + sw128_off(rl, kk)] = lo;
    }
}

// --------------------------------------------------------------------------------- main kernel
struct Smem {
    // offsets in bytes from the 1024-aligned base
    uint32_t w, a, bars, tmem_slot, pool, total;
};
__host__ __device__ inline Smem carve(int BN, int KC, int stages, bool pool) {
    Smem s;
    uint32_t off = 0;
    s.w = off;
    off += (uint32_t)KC * 2 * BN * 128;
    s.a = off;
    off += (uint32_t)stages * 2 * kAStageBytes;
    s.bars = off;
    off += 8 * (1 + 2 * kMaxStages + 4);
    s.tmem_slot = off;
    off += 16;
    s.pool = off;
    if (pool) off += 2 * kEpilogueWarps * BN * 4;
    s.total = off;
    return s;
}

template <int BN, bool GATHER, bool POOL>
__global__ void __launch_bounds__(kThreads, 1)
mlp_layer_tc_kernel(const TcArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const Smem L = carve(BN, a.KC, a.stages, POOL);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
    uint64_t *w_full = bars;
    uint64_t *a_full = bars + 1;
    uint64_t *a_empty = bars + 1 + kMaxStages;
    uint64_t *acc_full = bars + 1 + 2 * kMaxStages;
    uint64_t *acc_empty = acc_full + 2;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + L.tmem_slot);

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int nt = ceil_div(a.cout, BN);
    const int tile_n = blockIdx.x % nt;
    const int mi = blockIdx.x / nt;
    const int gm = gridDim.x / nt;
    const int n0 = tile_n * BN;
    const long long tiles_m = ceil_div<long long>(a.M, kBM);
    const int KC = a.KC;
    const int S = a.stages;
    constexpr int kTmemCols = 2 * BN;  // double-buffered accumulator (256 or 128 columns)

    if (tid == 0) {
        mbar_init(w_full, 1);
        for (int s = 0; s < kMaxStages; ++s) {
            mbar_init(a_full + s, kProducerWarps * 32);
            mbar_init(a_empty + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full + b, 1);
            mbar_init(acc_empty + b, kEpilogueWarps * 32);
        }
        fence_mbar_init();
    }
    if (warp == 8) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kProducerWarps) {
        // ============================ producers ============================================
        const int u = tid & 7;        // 16-byte unit inside the 128-byte chunk row
        const int rbase = tid >> 3;   // rows rbase + 16*j, j = 0..7
        uint32_t it = 0;
        for (long long tile = mi; tile < tiles_m; tile += gm) {
            const long long m0 = tile * kBM;
            // per-row source addresses for this tile
            long long src[8];
            int grp[8];
            bool valid[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const long long row = m0 + rbase + 16 * j;
                valid[j] = row < a.M;
                src[j] = row;
                grp[j] = 0;
                if (GATHER && valid[j]) {
                    const long long gg = row / a.K;
                    const int k = (int)(row - gg * a.K);
                    const long long b = gg / a.S;
                    int n = a.idx ? a.idx[row] : k;
                    n = min(max(n, 0), a.N - 1);
                    src[j] = b * a.N + n;
                    grp[j] = (int)gg;
                }
            }
            for (int c = 0; c < KC; ++c, ++it) {
                const int s = it % S;
                mbar_wait(a_empty + s, ((it / S) & 1) ^ 1);
                const int k0 = c * kChunkK + u * 4;  // first of this thread's 4 k values
                float4 v[8];
                if (!GATHER) {
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (valid[j] && k0 < a.cin)
                            v[j] = *reinterpret_cast<const float4 *>(a.x + src[j] * a.cin + k0);
                    }
                    if (a.in_scale != nullptr && k0 < a.cin) {
                        const float4 sc = *reinterpret_cast<const float4 *>(a.in_scale + k0);
                        const float4 sh = *reinterpret_cast<const float4 *>(a.in_shift + k0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            v[j].x = fmaxf(fmaf(v[j].x, sc.x, sh.x), 0.f);
                            v[j].y = fmaxf(fmaf(v[j].y, sc.y, sh.y), 0.f);
                            v[j].z = fmaxf(fmaf(v[j].z, sc.z, sh.z), 0.f);
                            v[j].w = fmaxf(fmaf(v[j].w, sc.w, sh.w), 0.f);
                            if (!valid[j]) v[j] = make_float4(0.f, 0.f, 0.f, 0.f);
                        }
                    }
                } else {
                    const int D = a.D;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        float t[4] = {0.f, 0.f, 0.f, 0.f};
                        if (valid[j]) {
                            if (k0 + 3 < D) {
                                const float4 f = *reinterpret_cast<const float4 *>(a.feats + src[j] * D + k0);
                                t[0] = f.x; t[1] = f.y; t[2] = f.z; t[3] = f.w;
                            } else if (k0 < D + 3) {
#pragma unroll
                                for (int q = 0; q < 4; ++q) {
                                    const int kk = k0 + q;
                                    if (kk < D) {
                                        t[q] = a.feats[src[j] * D + kk];
                                    } else if (kk < D + 3) {
                                        float pv = a.xyz[src[j] * 3 + (kk - D)];
                                        if (a.new_xyz != nullptr)
                                            pv = __fsub_rn(pv, a.new_xyz[(long long)grp[j] * 3 + (kk - D)]);
                                        t[q] = pv;
                                    }
                                }
                            }
                        }
                        v[j] = make_float4(t[0], t[1], t[2], t[3]);
                    }
                }
                // split hi / lo and store into the swizzled stage
                float *ahi = reinterpret_cast<float *>(smem + L.a + (uint32_t)s * 2 * kAStageBytes);
                float *alo = ahi + kAStageBytes / 4;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int r = rbase + 16 * j;
                    float4 hi, lo;
                    hi.x = tf32_rna(v[j].x); lo.x = tf32_rna(v[j].x - hi.x);
                    hi.y = tf32_rna(v[j].y); lo.y = tf32_rna(v[j].y - hi.y);
                    hi.z = tf32_rna(v[j].z); lo.z = tf32_rna(v[j].z - hi.z);
                    hi.w = tf32_rna(v[j].w); lo.w = tf32_rna(v[j].w - hi.w);
                    const int off = r * 32 + ((u ^ (r & 7)) << 2);
                    *reinterpret_cast<float4 *>(ahi + off) = hi;
                    *reinterpret_cast<float4 *>(alo + off) = lo;
                }
                fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
                mbar_arrive(a_full + s);
            }
        }
    } else if (warp == 8) {
        // ============================ MMA issuer ===========================================
        if (lane == 0) {
            // stage this CTA's W column tile (hi + lo, all K chunks) once
            const uint32_t wbytes = (uint32_t)KC * 2 * BN * 128;
            const uint8_t *wsrc = reinterpret_cast<const uint8_t *>(a.wimg) + (size_t)tile_n * wbytes;
            mbar_expect_tx(w_full, wbytes);
            for (uint32_t o = 0; o < wbytes; o += 16384) {
                const uint32_t n = min(16384u, wbytes - o);
                bulk_g2s(smem + L.w + o, wsrc + o, n, w_full);
            }
            mbar_wait(w_full, 0);
            constexpr uint32_t idesc = make_idesc(BN);
            const uint32_t w_base = smem_u32(smem + L.w);
            const uint32_t a_base = smem_u32(smem + L.a);
            uint32_t it = 0, tl = 0;
            for (long long tile = mi; tile < tiles_m; tile += gm, ++tl) {
                const uint32_t buf = tl & 1;
                mbar_wait(acc_empty + buf, ((tl >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * BN;
                for (int c = 0; c < KC; ++c, ++it) {
                    const int s = it % S;
                    mbar_wait(a_full + s, (it / S) & 1);
                    tc_fence_after();
                    const int kreal = min(kChunkK, a.kpad - c * kChunkK);  // multiple of 8
                    const uint32_t a_hi = a_base + (uint32_t)s * 2 * kAStageBytes;
                    const uint32_t a_lo = a_hi + kAStageBytes;
                    const uint32_t w_hi = w_base + (uint32_t)c * 2 * BN * 128;
                    const uint32_t w_lo = w_hi + BN * 128;
                    for (int ks = 0; ks * 8 < kreal; ++ks) {
                        const uint32_t ko = ks * 32;  // 8 tf32 = 32 bytes along K inside the swizzle row
                        const uint64_t dah = make_desc_sw128(a_hi + ko), dal = make_desc_sw128(a_lo + ko);
                        const uint64_t dwh = make_desc_sw128(w_hi + ko), dwl = make_desc_sw128(w_lo + ko);
                        umma_tf32(d_tmem, dal, dwh, idesc, (c | ks) != 0);  // small terms first
                        umma_tf32(d_tmem, dah, dwl, idesc, 1);
                        umma_tf32(d_tmem, dah, dwh, idesc, 1);
                    }
                    umma_commit(a_empty + s);  // stage reusable once these MMAs have read it
                }
                umma_commit(acc_full + buf);   // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // ============================ epilogue =============================================
        const int q = warp - kProducerWarps;  // == warp % 4: TMEM lane quadrant
        const int et = tid - kProducerWarps * 32;  // 0..127
        float *pool_s = reinterpret_cast<float *>(smem + L.pool);  // [2][4][BN]
        double acc_s[BN / 32], acc_q[BN / 32];
#pragma unroll
        for (int j = 0; j < BN / 32; ++j) acc_s[j] = acc_q[j] = 0.0;
        uint32_t tl = 0;
        for (long long tile = mi; tile < tiles_m; tile += gm, ++tl) {
            const uint32_t buf = tl & 1;
            const long long m0 = tile * kBM;
            const long long row = m0 + q * 32 + lane;
            const bool rvalid = row < a.M;
            mbar_wait(acc_full + buf, (tl >> 1) & 1);
            tc_fence_after();
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * BN + j * 32, v);
                const int c0 = n0 + j * 32;
                if (a.bias != nullptr) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (c0 + i < a.cout) v[i] += a.bias[c0 + i];
                }
                if (a.y != nullptr && rvalid) {
                    float *yr = a.y + row * a.cout + c0;
                    if (c0 + 31 < a.cout && a.vec_y) {
#pragma unroll
                        for (int i = 0; i < 8; ++i)
                            *reinterpret_cast<float4 *>(yr + 4 * i) =
                                make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 32; ++i)
                            if (c0 + i < a.cout) yr[i] = v[i];
                    }
                }
                if (POOL) {
                    float t[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) t[i] = rvalid ? v[i] : -INFINITY;
                    const float mx = warp_col_reduce(t, lane, [](float x, float y) { return fmaxf(x, y); });
#pragma unroll
                    for (int i = 0; i < 32; ++i) t[i] = rvalid ? v[i] : INFINITY;
                    const float mn = warp_col_reduce(t, lane, [](float x, float y) { return fminf(x, y); });
                    pool_s[(0 * 4 + q) * BN + j * 32 + lane] = mx;
                    pool_s[(1 * 4 + q) * BN + j * 32 + lane] = mn;
                }
                if (a.stats_partial != nullptr) {
                    float t[32];
#pragma unroll
                    for (int i = 0; i < 32; ++i) t[i] = rvalid ? v[i] : 0.f;
                    const float s1 = warp_col_reduce(t, lane, [](float x, float y) { return x + y; });
#pragma unroll
                    for (int i = 0; i < 32; ++i) t[i] = rvalid ? v[i] * v[i] : 0.f;
                    const float s2 = warp_col_reduce(t, lane, [](float x, float y) { return x + y; });
                    acc_s[j] += (double)s1;
                    acc_q[j] += (double)s2;
                }
            }
            // accumulator buffer drained: hand it back to the MMA issuer
            tc_fence_before();
            mbar_arrive(acc_empty + buf);
            if (POOL) {
                // combine the warps that share a group (K = 32: none, 64: pairs, 128: all four)
                named_bar_sync(1, kEpilogueWarps * 32);
                const int wpg = a.K / 32;  // warps per group: 1, 2 or 4
                for (int e = et; e < (4 / wpg) * BN; e += kEpilogueWarps * 32) {
                    const int gl = e / BN, cl = e - gl * BN;
                    const long long grow = m0 + (long long)gl * a.K;
                    if (grow < a.M && n0 + cl < a.cout) {
                        float mx = -INFINITY, mn = INFINITY;
                        for (int w = 0; w < wpg; ++w) {
                            mx = fmaxf(mx, pool_s[(0 * 4 + gl * wpg + w) * BN + cl]);
                            mn = fminf(mn, pool_s[(1 * 4 + gl * wpg + w) * BN + cl]);
                        }
                        const long long g = grow / a.K;
                        a.pool_max[g * a.cout + n0 + cl] = mx;
                        a.pool_min[g * a.cout + n0 + cl] = mn;
                    }
                }
                named_bar_sync(1, kEpilogueWarps * 32);  // pool_s free for the next tile
            }
        }
        // batch-statistics partial of this CTA: combine the four epilogue warps in fixed order
        if (a.stats_partial != nullptr) {
            // all MMAs of this CTA have completed (their accumulators were read above), so the
            // activation ring is dead: reuse it as fp64 scratch [4 warps][2][BN]
            named_bar_sync(1, kEpilogueWarps * 32);
            double *red = reinterpret_cast<double *>(smem + L.a);
#pragma unroll
            for (int j = 0; j < BN / 32; ++j) {
                red[(q * 2 + 0) * BN + j * 32 + lane] = acc_s[j];
                red[(q * 2 + 1) * BN + j * 32 + lane] = acc_q[j];
            }
            named_bar_sync(1, kEpilogueWarps * 32);
            for (int e = et; e < 2 * BN; e += kEpilogueWarps * 32) {
                const int which = e / BN, cl = e - which * BN;
                if (n0 + cl < a.cout) {
                    double tot = 0.0;
#pragma unroll
                    for (int w = 0; w < 4; ++w) tot += red[(w * 2 + which) * BN + cl];
                    a.stats_partial[((long long)mi * 2 + which) * a.cout + n0 + cl] = tot;
                    // partial rows this launch does not own are zeroed (fixed row count per M)
                    for (long long rr = mi + gm; rr < a.partial_rows; rr += gm)
                        a.stats_partial[(rr * 2 + which) * a.cout + n0 + cl] = 0.0;
                }
            }
        }
    }
    // ---- teardown
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == 8) tmem_dealloc<kTmemCols>(tmem_base);
}

// ------------------------------------------------------------------------------------ host side
static int pick_stages(int BN, int KC, bool pool, int *stages_out, uint32_t *smem_out) {
    for (int s = kMaxStages; s >= 2; --s) {
        const Smem L = carve(BN, KC, s, pool);
        const uint32_t need = L.total + 1024;  // alignment slack
        if (need <= 227u * 1024u) {
            *stages_out = s;
            *smem_out = need;
            return PAPC_OK;
        }
    }
    return PAPC_EUNSUPPORTED;
}

bool eligible(const TcProblem &p) {
    if (p.cout < 16 || p.cin < 1) return false;
    if (!p.gather && (p.cin % 4 != 0)) return false;           // float4 activation loads
    if (p.gather && p.D % 4 != 0) return false;
    if (p.pool && !(p.K == 32 || p.K == 64 || p.K == 128)) return false;
    const int BN = p.cout <= 64 ? 64 : 128;
    const int KC = ceil_div(p.cin, kChunkK);
    int st;
    uint32_t sm;
    return pick_stages(BN, KC, p.pool, &st, &sm) == PAPC_OK;
}

size_t wimg_bytes(int cin, int cout) {
    const int BN = cout <= 64 ? 64 : 128;
    const int KC = ceil_div(cin, kChunkK);
    return (size_t)ceil_div(cout, BN) * KC * 2 * BN * 128;
}

template <int BN>
static int launch_bn(const TcArgs &a, bool gather, bool pool, int grid, uint32_t smem, cudaStream_t st) {
#define PAPC_TC_LAUNCH(G, P)                                                                          \
    do {                                                                                              \
        auto k = mlp_layer_tc_kernel<BN, G, P>;                                                       \
        PAPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k<<<grid, kThreads, smem, st>>>(a);                                                           \
    } while (0)
    if (gather && pool) PAPC_TC_LAUNCH(true, true);
    else if (gather) PAPC_TC_LAUNCH(true, false);
    else if (pool) PAPC_TC_LAUNCH(false, true);
    else PAPC_TC_LAUNCH(false, false);
#undef PAPC_TC_LAUNCH
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

int launch(TcArgs a, const float *W, float *wimg, bool gather, int order, cudaStream_t st) {
    const int BN = a.cout <= 64 ? 64 : 128;
    const bool pool = a.pool_max != nullptr;
    a.KC = ceil_div(a.cin, kChunkK);
    a.kpad = ceil_div(a.cin, 8) * 8;
    uint32_t smem;
    int rc = pick_stages(BN, a.KC, pool, &a.stages, &smem);
    if (rc != PAPC_OK) return rc;
    const int nt = ceil_div(a.cout, BN);
    // W image (hi/lo split, swizzled, K order permuted for gathered sources)
    {
        const size_t total = (size_t)nt * a.KC * BN * 32;
        size_t blocks = (total + 255) / 256;
        if (blocks > (size_t)kNumSMs * 4) blocks = (size_t)kNumSMs * 4;
        prep_w_kernel<<<(unsigned)blocks, 256, 0, st>>>(W, a.cin, a.cout, gather ? 1 : 0, a.D, order, BN, a.KC,
                                                        wimg);
        PAPC_LAUNCH_CHECK();
    }
    a.wimg = wimg;
    const long long tiles_m = ceil_div<long long>(a.M, kBM);
    long long gm = kNumSMs / nt;
    if (gm < 1) gm = 1;
    if (gm > tiles_m) gm = tiles_m;
    if (gm > a.partial_rows) gm = a.partial_rows;
    const int grid = (int)(gm * nt);
    if (BN == 64) return launch_bn<64>(a, gather, pool, grid, smem, st);
    return launch_bn<128>(a, gather, pool, grid, smem, st);
}

}  // namespace tc
}  // namespace papc
