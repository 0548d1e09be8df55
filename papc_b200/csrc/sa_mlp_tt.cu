// sa_mlp_tt.cu -- grouped shared-MLP layer on the sm_100a tensor cores, transposed formulation:
//
//      Y^T[cout, rows] = W[cout, cin] * act(X)^T[cin, rows]        (one SetAbstraction MLP layer,
//                                                                   layers.py:214-219 / 271-276)
//
// The M = B*S*K grouped rows are the MMA *N* dimension and the output channels the MMA *M*
// dimension, so that
//   * W (hi/lo TF32 split) is the A operand and lives in TENSOR MEMORY for the whole kernel: the
//     tensor core reads only the activation tile from shared memory (half the operand traffic of a
//     shared/shared UMMA, which at M=N=128 saturates the 128 B/clk shared-memory port);
//   * the accumulator comes back with lane = channel, column = row: every epilogue thread owns one
//     channel, so bias, BatchNorm sum / sum^2, the per-group max/min pooling are plain per-thread
//     running reductions (no shuffles), and a warp stores 32 consecutive channels of one row
//     (128-byte coalesced) straight from the tcgen05.ld registers.
//
// fp32 in / fp32 out inside the 1e-5 parity budget: 3xTF32 (lo*hi + hi*lo + hi*hi, fp32
// accumulation in tensor memory).
//
// Persistent, one CTA per SM, 13 warps:
//   warps 0-3   epilogue  (TMEM lane quadrant = warp id).  They first stage W into tensor memory.
//   warp  4     MMA issuer (one thread) + TMEM allocation.
//   warps 5-12  producers: build the 128-row activation tile in shared memory (UMMA K-major
//               SWIZZLE_128B, hi and lo halves, ring of 32-column chunks), from one of
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

#include <stdlib.h>

#include <type_traits>

namespace papc {
namespace tt {

using namespace umma;

constexpr int kTile = 128;        // rows per tile == UMMA N
constexpr int kChunkK = 32;       // fp32 per K chunk == one 128-byte swizzle row
constexpr int kHalfBytes = kTile * 128;       // hi (or lo) half of one chunk stage
constexpr int kStageBytes = 2 * kHalfBytes;   // 32 KiB
constexpr int kStages = 6;
constexpr int kEpiWarps = 8;        // two per TMEM lane quadrant: rows 0-63 / 64-127 of each tile
constexpr int kProdWarps = 8;
constexpr int kProdThreads = kProdWarps * 32;
constexpr int kMmaWarp = kEpiWarps;
constexpr int kThreads = (kEpiWarps + 1 + kProdWarps) * 32;  // 544
constexpr int kMaxK = 128;        // reduction length held in tensor memory
// tensor-memory columns: two accumulators, then W hi, then W lo
constexpr uint32_t kColAcc = 0, kColWHi = 2 * kTile, kColWLo = 2 * kTile + kMaxK;
constexpr int kTmemCols = 512;

struct SmemLayout {
    static constexpr uint32_t ring = 0;
    static constexpr uint32_t xyz = ring + kStages * kStageBytes;   // [4][128] float4
    static constexpr uint32_t scale = xyz + 4 * kTile * 16;          // [128] float
    static constexpr uint32_t shift = scale + kMaxK * 4;             // [128] float
    static constexpr uint32_t fold = shift + kMaxK * 4;              // [128] float4
    static constexpr uint32_t wst = fold + kMaxK * 16;               // [4 warps][32][33] float
    static constexpr uint32_t bars = wst + 4 * 32 * 33 * 4;
    static constexpr uint32_t nbars = 2 * kStages + 2 + 2 + 1 + 4;
    static constexpr uint32_t misc = bars + nbars * 8;               // tmem slot, last-CTA flag
    static constexpr uint32_t total = misc + 16;
};
constexpr uint32_t kSmemBytes = SmemLayout::total + 1024;  // + alignment slack

struct RowGeom {
    long long bN;  // b * N
    int g;         // group index b*S + s
    int k;         // position inside the group
};
__device__ __forceinline__ RowGeom row_geom(long long row, int K, int S, int N) {
    RowGeom r;
    const long long gg = row / K;
    r.k = (int)(row - gg * K);
    r.g = (int)gg;
    r.bN = (gg / S) * (long long)N;
    return r;
}

struct Chunk {
    float4 v[4];
    float e0, e1, e2;  // SRC_GATHER: centred xyz of this thread's staging row
};

template <int MODE, bool POOL, bool HAS_Y>
__global__ void __launch_bounds__(kThreads, 1)
mlp_layer_tt_kernel(const TtArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    float4 *xyz_stage = reinterpret_cast<float4 *>(smem + SmemLayout::xyz);
    float *s_scale = reinterpret_cast<float *>(smem + SmemLayout::scale);
    float *s_shift = reinterpret_cast<float *>(smem + SmemLayout::shift);
    float4 *s_fold = reinterpret_cast<float4 *>(smem + SmemLayout::fold);
    float *s_wst = reinterpret_cast<float *>(smem + SmemLayout::wst);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + SmemLayout::bars);
    uint64_t *x_full = bars;
    uint64_t *x_empty = bars + kStages;
    uint64_t *acc_full = bars + 2 * kStages;
    uint64_t *acc_empty = acc_full + 2;
    uint64_t *w_ready = acc_empty + 2;
    uint64_t *xyz_full = w_ready + 1;
    uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(smem + SmemLayout::misc);
    uint32_t *s_last = tmem_slot + 1;

    const int tid = threadIdx.x;
    const int warp = tid >> 5;
    const int lane = tid & 31;
    const int nt = ceil_div(a.cout, kTile);
    const int tile_n = blockIdx.x % nt;
    const int mi = blockIdx.x / nt;
    const int gm = gridDim.x / nt;
    const int n0 = tile_n * kTile;
    const long long tiles_m = ceil_div<long long>(a.M, kTile);
    const int KC = ceil_div(a.cin, kChunkK);
    const int kpad = ceil_div(a.cin, 8) * 8;

    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(x_full + s, kProdWarps);
            mbar_init(x_empty + s, 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(acc_full + b, 1);
            mbar_init(acc_empty + b, kEpiWarps);
        }
        mbar_init(w_ready, 4);
        for (int b = 0; b < 4; ++b) mbar_init(xyz_full + b, kProdWarps);
        fence_mbar_init();
    }
    if (warp == kMmaWarp) tmem_alloc<kTmemCols>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kEpiWarps) {
        // ================================ epilogue =========================================
        const int quad = warp & 3;        // TMEM lane quadrant this warp may access
        const int half = warp >> 2;       // rows [64*half, 64*half + 64) of every tile
        const int c = quad * 32 + lane;   // channel inside this CTA's 128-channel tile == TMEM lane
        const int cg = n0 + c;
        const bool cvalid = cg < a.cout;
        const uint32_t lane_base = tmem_base + ((uint32_t)(quad * 32) << 16);
        // ---- warps 0-3 stage 32 W rows each into tensor memory (hi / lo split): coalesced row
        //      loads, transposed through shared memory so that thread = channel owns its row
        if (half == 0) {
            float *wst = s_wst + quad * 32 * 33;
            for (int kc = 0; kc < KC; ++kc) {
                const int k = kc * kChunkK + lane;
#pragma unroll
                for (int rr = 0; rr < 32; ++rr) {
                    const int row = n0 + quad * 32 + rr;
                    wst[rr * 33 + lane] = (row < a.cout && k < a.cin)
                                              ? __ldg(a.W + (size_t)row * a.wld + a.wk0 + k) : 0.f;
                }
                __syncwarp();
                uint32_t hi[32], lo[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float h, l;
                    split_tf32(wst[lane * 33 + i], h, l);
                    hi[i] = __float_as_uint(h);
                    lo[i] = __float_as_uint(l);
                }
                tmem_st32(lane_base + kColWHi + kc * kChunkK, hi);
                tmem_st32(lane_base + kColWLo + kc * kChunkK, lo);
                __syncwarp();
            }
            tmem_wait_st();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(w_ready);
        }
        const float bias = (a.bias != nullptr && cvalid) ? a.bias[cg] : 0.f;
        float wx = 0.f, wy = 0.f, wz = 0.f;
        const bool has_xyz = (MODE == SRC_GATHER) && a.wxyz >= 0;
        if (has_xyz && cvalid) {
            const float *wp = a.W + (size_t)cg * a.wld + a.wxyz;
            wx = wp[0]; wy = wp[1]; wz = wp[2];
        }
        const bool do_y = HAS_Y && cvalid;
        const bool do_pool = POOL && cvalid;
        const uint64_t bias2 = pack2(bias, bias);
        const uint64_t wx2 = pack2(wx, wx), wy2 = pack2(wy, wy), wz2 = pack2(wz, wz);
        const int kshift = POOL ? (a.K == 32 ? 5 : a.K == 64 ? 6 : 7) : 0;  // K in {32, 64, 128}
        // s_wst is dead once W is staged (all eight warps pass the named barrier below first):
        // reused to combine the two row-halves (pool extrema for K = 128, final statistics)
        float *s_comb = s_wst;  // [2 tile parities][2][128] floats
        named_bar_sync(2, kEpiWarps * 32);
        double acc_s = 0.0, acc_q = 0.0;
        uint32_t tl = 0;
        for (long long tile = mi; tile < tiles_m; tile += gm, ++tl) {
            const uint32_t buf = tl & 1;
            const long long m0 = tile * kTile;
            const int nrows = (int)((a.M - m0) < kTile ? (a.M - m0) : kTile);
            if (MODE == SRC_GATHER) mbar_wait(xyz_full + (tl & 3), (tl >> 2) & 1);
            mbar_wait(acc_full + buf, (tl >> 1) & 1);
            tc_fence_after();
            // this warp's 64 accumulator columns -> registers, then hand the buffer straight back
            uint32_t r[2][32];
            tmem_ld32_nowait(lane_base + kColAcc + buf * kTile + half * 64, r[0]);
            tmem_ld32_nowait(lane_base + kColAcc + buf * kTile + half * 64 + 32, r[1]);
            tmem_wait_ld();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty + buf);

            const float4 *xs = xyz_stage + (tl & 3) * kTile + half * 64;
            float *yrow = do_y ? a.y + ((size_t)m0 + half * 64) * a.cout + cg : nullptr;
            const int nr = nrows - half * 64;  // valid rows of this half (may be <= 0)
            uint64_t s2 = 0ull, q2 = 0ull;     // packed (even rows, odd rows) running sums
            float mx = -INFINITY, mn = INFINITY;
            auto body = [&](auto full_tag) {
                constexpr bool FULL = decltype(full_tag)::value;
#pragma unroll
                for (int j = 0; j < 2; ++j) {
#pragma unroll
                    for (int i = 0; i < 32; i += 2) {
                        const int rr = j * 32 + i;
                        uint64_t v2 = add2(pack2u(r[j][i], r[j][i + 1]), bias2);
                        if (has_xyz) {
                            const float4 p0 = xs[rr], p1 = xs[rr + 1];
                            v2 = fma2(wx2, pack2(p0.x, p1.x), v2);
                            v2 = fma2(wy2, pack2(p0.y, p1.y), v2);
                            v2 = fma2(wz2, pack2(p0.z, p1.z), v2);
                        }
                        float va, vb;
                        unpack2(v2, va, vb);
                        if (FULL) {
                            s2 = add2(s2, v2);
                            q2 = fma2(v2, v2, q2);
                            if (do_y) {
                                yrow[(size_t)rr * a.cout] = va;
                                yrow[(size_t)(rr + 1) * a.cout] = vb;
                            }
                            if (POOL) {
                                mx = fmaxf(fmaxf(mx, va), vb);
                                mn = fminf(fminf(mn, va), vb);
                            }
                        } else {
                            const bool rva = rr < nr, rvb = rr + 1 < nr;
                            const uint64_t m2 = pack2(rva ? va : 0.f, rvb ? vb : 0.f);
                            s2 = add2(s2, m2);
                            q2 = fma2(m2, m2, q2);
                            if (do_y && rva) yrow[(size_t)rr * a.cout] = va;
                            if (do_y && rvb) yrow[(size_t)(rr + 1) * a.cout] = vb;
                            if (POOL) {
                                if (rva) { mx = fmaxf(mx, va); mn = fminf(mn, va); }
                                if (rvb) { mx = fmaxf(mx, vb); mn = fminf(mn, vb); }
                            }
                        }
                    }
                    if (POOL && kshift == 5) {  // K = 32: one group per 32-row block
                        if (do_pool && (FULL || j * 32 < nr)) {
                            const long long g = (m0 >> 5) + half * 2 + j;
                            a.pool_max[g * a.cout + cg] = mx;
                            a.pool_min[g * a.cout + cg] = mn;
                        }
                        mx = -INFINITY;
                        mn = INFINITY;
                    }
                }
            };
            if (a.dbg & 4) {
                s2 = pack2u(r[0][0], r[1][31]);
            } else if (nr >= 64) body(std::true_type{});
            else body(std::false_type{});
            if (POOL && kshift == 6) {  // K = 64: this half is exactly one group
                if (do_pool && nr > 0) {
                    const long long g = (m0 >> 6) + half;
                    a.pool_max[g * a.cout + cg] = mx;
                    a.pool_min[g * a.cout + cg] = mn;
                }
            }
            if (POOL && kshift == 7) {  // K = 128: the group spans both halves -> combine
                float *cb = s_comb + (tl & 1) * 256;
                if (half == 1) {
                    cb[c] = mx;
                    cb[128 + c] = mn;
                }
                named_bar_sync(2, kEpiWarps * 32);
                if (half == 0 && do_pool) {
                    const long long g = m0 >> 7;
                    a.pool_max[g * a.cout + cg] = fmaxf(mx, cb[c]);
                    a.pool_min[g * a.cout + cg] = fminf(mn, cb[128 + c]);
                }
            }
            float sa, sb, qa, qb;
            unpack2(s2, sa, sb);
            unpack2(q2, qa, qb);
            acc_s += (double)sa + (double)sb;
            acc_q += (double)qa + (double)qb;
        }
        if (a.stats_partial != nullptr) {
            // combine the two row-halves, one partial row per CTA (fixed order -> deterministic)
            named_bar_sync(2, kEpiWarps * 32);  // every pool exchange through s_comb is finished
            double *cd = reinterpret_cast<double *>(s_wst);  // [2][128] doubles
            if (half == 1) {
                cd[c] = acc_s;
                cd[128 + c] = acc_q;
            }
            named_bar_sync(2, kEpiWarps * 32);
            if (half == 0 && cvalid) {
                a.stats_partial[((long long)mi * 2 + 0) * a.cout + cg] = acc_s + cd[c];
                a.stats_partial[((long long)mi * 2 + 1) * a.cout + cg] = acc_q + cd[128 + c];
                // partial rows this launch does not own are zeroed (fixed row count per M)
                for (long long rr = mi + gm; rr < a.partial_rows; rr += gm) {
                    a.stats_partial[(rr * 2 + 0) * a.cout + cg] = 0.0;
                    a.stats_partial[(rr * 2 + 1) * a.cout + cg] = 0.0;
                }
                __threadfence();
            }
        }
    } else if (warp == kMmaWarp) {
        // ================================ MMA issuer =======================================
        if (lane == 0) {
            mbar_wait(w_ready, 0);
            tc_fence_after();
            constexpr uint32_t idesc = make_idesc_tf32_m128(kTile);
            const uint32_t ring_base = smem_u32(smem + SmemLayout::ring);
            uint32_t it = 0, tl = 0;
            for (long long tile = mi; tile < tiles_m; tile += gm, ++tl) {
                const uint32_t buf = tl & 1;
                mbar_wait(acc_empty + buf, ((tl >> 1) & 1) ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + kColAcc + buf * kTile;
                for (int c = 0; c < KC; ++c, ++it) {
                    const uint32_t s = it % kStages;
                    mbar_wait(x_full + s, (it / kStages) & 1);
                    tc_fence_after();
                    const int kreal = min(kChunkK, kpad - c * kChunkK);  // multiple of 8
                    const uint32_t x_hi = ring_base + s * kStageBytes;
                    const uint32_t x_lo = x_hi + kHalfBytes;
                    const uint32_t w_hi = tmem_base + kColWHi + c * kChunkK;
                    const uint32_t w_lo = tmem_base + kColWLo + c * kChunkK;
                    for (int ks = 0; ks * 8 < kreal && !(a.dbg & 2); ++ks) {
                        const uint64_t dxh = make_desc_sw128(x_hi + ks * 32);
                        const uint64_t dxl = make_desc_sw128(x_lo + ks * 32);
                        mma_tf32_ts(d_tmem, w_lo + ks * 8, dxh, idesc, (c | ks) != 0);  // small terms first
                        mma_tf32_ts(d_tmem, w_hi + ks * 8, dxl, idesc, 1);
                        mma_tf32_ts(d_tmem, w_hi + ks * 8, dxh, idesc, 1);
                    }
                    mma_commit(x_empty + s);  // chunk reusable once these MMAs have read it
                }
                mma_commit(acc_full + buf);   // accumulator complete
            }
        }
        __syncwarp();
    } else {
        // ================================ producers ========================================
        const int ptid = tid - (kEpiWarps + 1) * 32;  // 0..255
        const int u = ptid & 7;       // 16-byte unit inside the 128-byte chunk row
        const int rb = ptid >> 3;     // rows rb + 32*j, j = 0..3
        const uint32_t swz = (uint32_t)((u ^ (rb & 7)) << 4);
        const bool has_act = (MODE == SRC_PLAIN) && a.in_scale != nullptr;
        if (MODE == SRC_PLAIN && ptid < kMaxK) {
            s_scale[ptid] = (has_act && ptid < a.cin) ? a.in_scale[ptid] : 0.f;
            s_shift[ptid] = (has_act && ptid < a.cin) ? a.in_shift[ptid] : 0.f;
        }
        if (MODE == SRC_POINTMLP && ptid < kMaxK)
            s_fold[ptid] = ptid < a.cin ? reinterpret_cast<const float4 *>(a.l0_fold)[ptid]
                                        : make_float4(0.f, 0.f, 0.f, 0.f);
        named_bar_sync(1, kProdThreads);

        // ---- per-row geometry pipeline (SRC_GATHER / SRC_POINTMLP)
        long long src[4] = {0, 0, 0, 0};   // b*N + n of the current tile's rows
        int grp[4] = {0, 0, 0, 0};
        int idx_pf[4] = {0, 0, 0, 0};      // prefetched neighbour indices of a later tile
        long long bN_pf[4] = {0, 0, 0, 0};
        int grp_pf[4] = {0, 0, 0, 0};
        float4 p_nx[4];                    // SRC_POINTMLP: centred points of the next tile
#pragma unroll
        for (int j = 0; j < 4; ++j) p_nx[j] = make_float4(0.f, 0.f, 0.f, 0.f);

        auto fetch_idx = [&](long long tile) {  // -> idx_pf / bN_pf / grp_pf for `tile`
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                long long row = tile * kTile + rb + 32 * j;
                row = row < a.M ? row : a.M - 1;
                const RowGeom rg = row_geom(row, a.K, a.S, a.N);
                bN_pf[j] = rg.bN;
                grp_pf[j] = rg.g;
                idx_pf[j] = a.idx != nullptr ? __ldg(a.idx + row) : rg.k;
            }
        };
        auto fetch_points = [&]() {  // idx_pf (arrived) -> p_nx
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = min(max(idx_pf[j], 0), a.N - 1);
                const float *q = a.xyz + (bN_pf[j] + n) * 3;
                float px = __ldg(q), py = __ldg(q + 1), pz = __ldg(q + 2);
                if (a.new_xyz != nullptr) {
                    const float *cc = a.new_xyz + (long long)grp_pf[j] * 3;
                    px = __fsub_rn(px, __ldg(cc));
                    py = __fsub_rn(py, __ldg(cc + 1));
                    pz = __fsub_rn(pz, __ldg(cc + 2));
                }
                p_nx[j] = make_float4(px, py, pz, 0.f);
            }
        };

        auto load = [&](long long tile, int c, Chunk &ch) {
            const long long m0 = tile * kTile;
            const int k0 = c * kChunkK + u * 4;
            if (MODE == SRC_PLAIN) {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    long long row = m0 + rb + 32 * j;
                    row = row < a.M ? row : a.M - 1;
                    ch.v[j] = k0 < a.cin ? __ldg(reinterpret_cast<const float4 *>(a.x + row * a.cin + k0))
                                         : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            } else if (MODE == SRC_GATHER) {
                if (c == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        src[j] = bN_pf[j] + min(max(idx_pf[j], 0), a.N - 1);
                        grp[j] = grp_pf[j];
                    }
                    fetch_idx(tile + gm < tiles_m ? tile + gm : tile);
                    if (u < 4) {  // this thread also stages the centred xyz of row rb + 32*u
                        const long long sj = u == 0 ? src[0] : u == 1 ? src[1] : u == 2 ? src[2] : src[3];
                        const int gj = u == 0 ? grp[0] : u == 1 ? grp[1] : u == 2 ? grp[2] : grp[3];
                        const float *q = a.xyz + sj * 3;
                        ch.e0 = __ldg(q); ch.e1 = __ldg(q + 1); ch.e2 = __ldg(q + 2);
                        if (a.new_xyz != nullptr) {
                            const float *cc = a.new_xyz + (long long)gj * 3;
                            ch.e0 = __fsub_rn(ch.e0, __ldg(cc));
                            ch.e1 = __fsub_rn(ch.e1, __ldg(cc + 1));
                            ch.e2 = __fsub_rn(ch.e2, __ldg(cc + 2));
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    ch.v[j] = k0 < a.D ? __ldg(reinterpret_cast<const float4 *>(a.feats + src[j] * a.D + k0))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {  // SRC_POINTMLP: no per-chunk loads, only the per-tile point pipeline
                if (c == 0) {
                    // p_nx holds this tile's points (issued one tile ago); refill it for the next
                    // tile from idx_pf (issued one tile ago), then prefetch idx two tiles ahead
#pragma unroll
                    for (int j = 0; j < 4; ++j) ch.v[j] = p_nx[j];
                    fetch_points();
                    const long long t2 = tile + 2LL * gm;
                    fetch_idx(t2 < tiles_m ? t2 : tile);
                }
            }
        };
        // SRC_POINTMLP keeps the current tile's points here (chunks c >= 1 reuse them)
        float4 p_cur[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) p_cur[j] = make_float4(0.f, 0.f, 0.f, 0.f);

        uint32_t it = 0, ptl = 0;
        auto process = [&](int c, Chunk &ch) {
            const uint32_t s = it % kStages;
            const int k0 = c * kChunkK + u * 4;
            float4 o[4];
            if (MODE == SRC_PLAIN) {
                if (has_act) {
                    const float4 sc = *reinterpret_cast<const float4 *>(s_scale + k0);
                    const float4 sh = *reinterpret_cast<const float4 *>(s_shift + k0);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        o[j].x = fmaxf(fmaf(ch.v[j].x, sc.x, sh.x), 0.f);
                        o[j].y = fmaxf(fmaf(ch.v[j].y, sc.y, sh.y), 0.f);
                        o[j].z = fmaxf(fmaf(ch.v[j].z, sc.z, sh.z), 0.f);
                        o[j].w = fmaxf(fmaf(ch.v[j].w, sc.w, sh.w), 0.f);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = ch.v[j];
                }
            } else if (MODE == SRC_GATHER) {
#pragma unroll
                for (int j = 0; j < 4; ++j) o[j] = ch.v[j];
            } else {
                if (c == 0) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) p_cur[j] = ch.v[j];
                }
                const float4 f0 = s_fold[k0], f1 = s_fold[k0 + 1], f2 = s_fold[k0 + 2], f3 = s_fold[k0 + 3];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float4 p = p_cur[j];
                    o[j].x = fmaxf(fmaf(f0.x, p.x, fmaf(f0.y, p.y, fmaf(f0.z, p.z, f0.w))), 0.f);
                    o[j].y = fmaxf(fmaf(f1.x, p.x, fmaf(f1.y, p.y, fmaf(f1.z, p.z, f1.w))), 0.f);
                    o[j].z = fmaxf(fmaf(f2.x, p.x, fmaf(f2.y, p.y, fmaf(f2.z, p.z, f2.w))), 0.f);
                    o[j].w = fmaxf(fmaf(f3.x, p.x, fmaf(f3.y, p.y, fmaf(f3.z, p.z, f3.w))), 0.f);
                }
            }
            mbar_wait(x_empty + s, ((it / kStages) & 1) ^ 1);
            if (MODE == SRC_GATHER && c == 0) {
                if (u < 4) xyz_stage[(ptl & 3) * kTile + rb + 32 * u] = make_float4(ch.e0, ch.e1, ch.e2, 0.f);
                __syncwarp();
                if (lane == 0) mbar_arrive(xyz_full + (ptl & 3));
            }
            uint8_t *hi_base = smem + SmemLayout::ring + s * kStageBytes;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float4 hi, lo;
                split_tf32(o[j].x, hi.x, lo.x);
                split_tf32(o[j].y, hi.y, lo.y);
                split_tf32(o[j].z, hi.z, lo.z);
                split_tf32(o[j].w, hi.w, lo.w);
                const uint32_t off = (uint32_t)(rb + 32 * j) * 128u + swz;
                *reinterpret_cast<float4 *>(hi_base + off) = hi;
                *reinterpret_cast<float4 *>(hi_base + kHalfBytes + off) = lo;
            }
            fence_proxy_async();  // generic-proxy writes -> visible to the tensor core (async proxy)
            __syncwarp();
            if (lane == 0) mbar_arrive(x_full + s);
            ++it;
            if (c == KC - 1) ++ptl;
        };

        const long long tile = mi;
        const bool have = tile < tiles_m;
        if (have && MODE != SRC_PLAIN) {
            fetch_idx(tile);                       // idx of the first tile
            if (MODE == SRC_POINTMLP) {
                fetch_points();                    // -> p_nx = points of the first tile
                fetch_idx(tile + gm < tiles_m ? tile + gm : tile);
            }
        }
        // three chunks of global loads in flight per thread (registers ca / cb / cc rotate)
        Chunk ca, cb, cc;
        ca.e0 = ca.e1 = ca.e2 = cb.e0 = cb.e1 = cb.e2 = cc.e0 = cc.e1 = cc.e2 = 0.f;
        auto advance = [&](long long &t, int &ci) {
            if (++ci == KC) { ci = 0; t += gm; }
        };
        long long t0 = tile, t1 = tile, t2;
        int c0 = 0, c1 = 0, c2;
        advance(t1, c1);
        bool have0 = have, have1 = have && t1 < tiles_m;
        if (a.dbg & 1) {
            // triage: ring protocol only
            for (long long t = tile; t < tiles_m; t += gm)
                for (int cq = 0; cq < KC; ++cq) {
                    const uint32_t s = it % kStages;
                    mbar_wait(x_empty + s, ((it / kStages) & 1) ^ 1);
                    if (MODE == SRC_GATHER && cq == 0) { __syncwarp(); if (lane == 0) mbar_arrive(xyz_full + (ptl & 3)); }
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(x_full + s);
                    ++it;
                    if (cq == KC - 1) ++ptl;
                }
            have0 = false;
        }
        if (have0) load(t0, c0, ca);
        if (have1) load(t1, c1, cb);
        while (have0) {
            t2 = t1; c2 = c1; advance(t2, c2);
            const bool have2 = have1 && t2 < tiles_m;
            if (have2) load(t2, c2, cc);
            process(c0, ca);
            if (!have1) break;
            long long t3 = t2; int c3 = c2; advance(t3, c3);
            const bool have3 = have2 && t3 < tiles_m;
            if (have3) load(t3, c3, ca);
            process(c1, cb);
            if (!have2) break;
            long long t4 = t3; int c4 = c3; advance(t4, c4);
            const bool have4 = have3 && t4 < tiles_m;
            if (have4) load(t4, c4, cb);
            process(c2, cc);
            t0 = t3; c0 = c3; have0 = have3;
            t1 = t4; c1 = c4; have1 = have4;
        }
    }

    // ---- teardown
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (warp == kMmaWarp) tmem_dealloc<kTmemCols>(tmem_base);

    // ---- fused BatchNorm finalisation: the last CTA reduces the partial rows in fixed order
    if (a.counter != nullptr) {
        if (tid == 0) {
            __threadfence();
            *s_last = (atomicAdd(a.counter, 1u) == gridDim.x - 1) ? 1u : 0u;
        }
        __syncthreads();
        if (*s_last != 0u) {
            __threadfence();
            // 2*cout (channel, sum | sum^2) columns x gm partial rows: 256 columns at a time, the
            // rows split over two thread slices (fixed order -> deterministic), combined in smem
            double *red = reinterpret_cast<double *>(smem + SmemLayout::ring);  // [2][256]
            const int col_l = tid & 255, sl = tid >> 8;  // threads >= 512 idle
            for (int base = 0; base < 2 * a.cout; base += 256) {
                const int col = base + col_l;  // = which * cout + ch
                double acc = 0.0;
                if (sl < 2 && col < 2 * a.cout) {
                    const int which = col / a.cout, ch = col - which * a.cout;
                    const double *p = a.stats_partial + (long long)which * a.cout + ch;
#pragma unroll 16
                    for (int r = sl; r < gm; r += 2) acc += __ldcg(p + (long long)r * 2 * a.cout);
                    red[sl * 256 + col_l] = acc;
                }
                __syncthreads();
                if (sl == 0 && col < 2 * a.cout) red[col_l] = red[col_l] + red[256 + col_l];
                __syncthreads();
                // channels whose sum AND sum^2 columns are both inside this 256-column window
                // (cout >= 128: the two columns of a channel are cout apart -> handled below)
                if (2 * a.cout <= 256) {
                    if (tid < a.cout) {
                        const int ch = tid;
                        const double mean = red[ch] / a.count;
                        double var = red[a.cout + ch] / a.count - mean * mean;  // biased (Paddle training BN)
                        var = var > 0.0 ? var : 0.0;
                        const double g = a.gamma ? (double)a.gamma[ch] : 1.0;
                        const double b = a.beta ? (double)a.beta[ch] : 0.0;
                        const double sc = g / sqrt(var + (double)a.eps);
                        a.scale[ch] = (float)sc;
                        a.shift[ch] = (float)(b - mean * sc);
                        if (a.mean_out) a.mean_out[ch] = (float)mean;
                        if (a.var_out) a.var_out[ch] = (float)var;
                    }
                } else {
                    // park the reduced columns in the (dead) second ring stage: [2*cout] doubles
                    double *all = reinterpret_cast<double *>(smem + SmemLayout::ring + kStageBytes);
                    if (sl == 0 && col < 2 * a.cout) all[col] = red[col_l];
                }
                __syncthreads();
            }
            if (2 * a.cout > 256) {
                const double *all = reinterpret_cast<const double *>(smem + SmemLayout::ring + kStageBytes);
                for (int ch = tid; ch < a.cout; ch += kThreads) {
                    const double mean = all[ch] / a.count;
                    double var = all[a.cout + ch] / a.count - mean * mean;
                    var = var > 0.0 ? var : 0.0;
                    const double g = a.gamma ? (double)a.gamma[ch] : 1.0;
                    const double b = a.beta ? (double)a.beta[ch] : 0.0;
                    const double sc = g / sqrt(var + (double)a.eps);
                    a.scale[ch] = (float)sc;
                    a.shift[ch] = (float)(b - mean * sc);
                    if (a.mean_out) a.mean_out[ch] = (float)mean;
                    if (a.var_out) a.var_out[ch] = (float)var;
                }
            }
            if (tid == 0) *a.counter = 0u;  // self-cleaning for the next launch
        }
    }
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
    if (batch) {
        float acc[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) acc[i] = 0.f;
        double dacc[9];
#pragma unroll
        for (int i = 0; i < 9; ++i) dacc[i] = 0.0;
        int n_in_acc = 0;
        for (long long row = (long long)blockIdx.x * kMomThreads + tid; row < a.M;
             row += (long long)gridDim.x * kMomThreads) {
            const RowGeom rg = row_geom(row, a.K, a.S, a.N);
            int n = a.idx != nullptr ? __ldg(a.idx + row) : rg.k;
            n = min(max(n, 0), a.N - 1);
            const float *q = a.xyz + (rg.bN + n) * 3;
            float px = __ldg(q), py = __ldg(q + 1), pz = __ldg(q + 2);
            if (a.new_xyz != nullptr) {
                const float *cc = a.new_xyz + (long long)rg.g * 3;
                px = __fsub_rn(px, __ldg(cc));
                py = __fsub_rn(py, __ldg(cc + 1));
                pz = __fsub_rn(pz, __ldg(cc + 2));
            }
            acc[0] += px; acc[1] += py; acc[2] += pz;
            acc[3] = fmaf(px, px, acc[3]); acc[4] = fmaf(px, py, acc[4]); acc[5] = fmaf(px, pz, acc[5]);
            acc[6] = fmaf(py, py, acc[6]); acc[7] = fmaf(py, pz, acc[7]); acc[8] = fmaf(pz, pz, acc[8]);
            if (++n_in_acc == 16) {  // bound the fp32 run length
#pragma unroll
                for (int i = 0; i < 9; ++i) { dacc[i] += (double)acc[i]; acc[i] = 0.f; }
                n_in_acc = 0;
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
        if (tid < 9) {
            double v = 0.0;
            for (int b = 0; b < (int)gridDim.x; ++b) v += __ldcg(a.partial + (long long)b * 9 + tid);
            s_tot[tid] = v / (double)a.M;
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
        reinterpret_cast<float4 *>(a.l0_fold)[c] =
            make_float4((float)(sc * w0), (float)(sc * w1), (float)(sc * w2), (float)(sc * b0 + sh));
    }
}

// ------------------------------------------------------------------------------------ host side
bool eligible(const TtProblem &p) {
    if (p.cout < 1 || p.cout > 4 * kTile) return false;
    if (p.cin < 4 || p.cin > kMaxK || p.cin % 4 != 0) return false;
    if (p.mode == SRC_GATHER && p.D != p.cin) return false;
    if (p.pool && !(p.K == 32 || p.K == 64 || p.K == 128)) return false;
    return true;
}

template <int MODE, bool POOL, bool HAS_Y>
static int launch_inst(const TtArgs &a, int grid, cudaStream_t st) {
    auto k = mlp_layer_tt_kernel<MODE, POOL, HAS_Y>;
    static bool configured = false;
    if (!configured) {
        PAPC_CUDA_TRY(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmemBytes));
        configured = true;
    }
    k<<<grid, kThreads, kSmemBytes, st>>>(a);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

int launch(const TtArgs &a_in, cudaStream_t st) {
    TtArgs a = a_in;
    {
        const char *e = getenv("PAPC_TT_DBG");
        a.dbg = e ? atoi(e) : 0;
    }
    const bool pool = a.pool_max != nullptr;
    const int nt = ceil_div(a.cout, kTile);
    const long long tiles_m = ceil_div<long long>(a.M, kTile);
    long long gm = kNumSMs / nt;
    if (gm < 1) gm = 1;
    if (gm > tiles_m) gm = tiles_m;
    if (a.stats_partial != nullptr && gm > a.partial_rows) gm = a.partial_rows;
    const int grid = (int)(gm * nt);
    const bool has_y = a.y != nullptr;
#define PAPC_TT_CASE(MODE_)                                                        \
    case MODE_:                                                                    \
        if (pool && has_y) return launch_inst<MODE_, true, true>(a, grid, st);     \
        if (pool) return launch_inst<MODE_, true, false>(a, grid, st);             \
        if (has_y) return launch_inst<MODE_, false, true>(a, grid, st);            \
        return launch_inst<MODE_, false, false>(a, grid, st);
    switch (a.mode) {
        PAPC_TT_CASE(SRC_PLAIN)
        PAPC_TT_CASE(SRC_GATHER)
        PAPC_TT_CASE(SRC_POINTMLP)
    }
#undef PAPC_TT_CASE
    return PAPC_EINVAL;
}

int moment_blocks(long long M) {
    long long b = ceil_div<long long>(M, 2048);
    if (b < 1) b = 1;
    if (b > 2LL * kNumSMs) b = 2LL * kNumSMs;
    return (int)b;
}

int launch_moments(const MomentArgs &a, cudaStream_t st) {
    const int blocks = a.running_mean != nullptr ? 1 : moment_blocks(a.M);
    point_moments_kernel<<<blocks, kMomThreads, 0, st>>>(a);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

}  // namespace tt
}  // namespace papc
