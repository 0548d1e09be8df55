// pillars.cu -- PointPillars pillar encode for sm_100a:
//   points_to_voxel (reference: point_cloud_ops.py:7-166), PillarFeatureNet with one PFNLayer
//   (pillars.py:9-108) and PointPillarsScatter (pillars.py:110-142).
//
// The reference voxeliser is a sequential loop ("need mutex if write in cuda", :65).  Here its
// order-dependent semantics are reproduced deterministically in parallel:
//   1. cell id per point; atomicMin gives each cell its FIRST point index.
//   2. a point is a "cell opener" iff it is its cell's first point; an in-order scan over the
//      opener flags numbers the voxels first-come.  The opener whose scan value equals
//      max_voxels is where the reference `break`s: every point from there on is dropped.
//   3. surviving points are counted and bucketed per voxel (CSR; bucket order arbitrary),
//   4. one warp per voxel selects the max_points smallest point indices of its bucket
//      (bisection on the index value when the bucket is larger), ranks them, and writes the
//      voxel's rows -- real rows then zero padding -- so every byte of `voxels` is written
//      exactly once with coalesced 16-byte stores; no separate 19 MB memset.
#include "common.cuh"

namespace papc {

struct VoxGeom {
    float lo[3], vs[3];
    int grid[3];   // x,y,z cell counts
    int shape[3];  // dense map shape (grid reversed when reverse_index)
    int reverse;
};

static bool make_geom(const float *vs, const float *cr, int reverse, VoxGeom *g) {
    for (int j = 0; j < 3; ++j) {
        if (!(vs[j] > 0.f)) return false;
        g->lo[j] = cr[j];
        g->vs[j] = vs[j];
        const float q = (cr[3 + j] - cr[j]) / vs[j];           // fp32, pc_ops.py:24
        g->grid[j] = (int)nearbyintf(q);                        // np.round: half-even
        if (g->grid[j] <= 0) return false;
    }
    for (int j = 0; j < 3; ++j) g->shape[j] = reverse ? g->grid[2 - j] : g->grid[j];
    g->reverse = reverse;
    return true;
}

struct VoxWs {
    int32_t *cell_first;  // [cells]  first point index of each cell (INT_MAX = empty)
    int32_t *pt_cell;     // [N]      cell id, later voxel id (or -1)
    int32_t *pt_vid;      // [N]      exclusive scan of opener flags
    int32_t *cnt;         // [max_voxels] points per voxel before the max_points cap
    int32_t *off;         // [max_voxels] bucket offsets
    int32_t *fillc;       // [max_voxels] bucket fill counters
    int32_t *list;        // [N]      bucketed point indices
    int32_t *meta;        // [0] cutoff point index i*  [1] total openers
    size_t total;
};

static void carve_vox_ws(char *base, int N, size_t cells, int max_voxels, VoxWs *w) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t o_cf = take(cells * 4), o_pc = take((size_t)N * 4), o_pv = take((size_t)N * 4);
    const size_t o_cnt = take((size_t)max_voxels * 4), o_off = take((size_t)max_voxels * 4);
    const size_t o_fill = take((size_t)max_voxels * 4), o_list = take((size_t)N * 4);
    const size_t o_meta = take(16);
    w->total = off;
    if (!base) return;
    w->cell_first = (int32_t *)(base + o_cf);
    w->pt_cell = (int32_t *)(base + o_pc);
    w->pt_vid = (int32_t *)(base + o_pv);
    w->cnt = (int32_t *)(base + o_cnt);
    w->off = (int32_t *)(base + o_off);
    w->fillc = (int32_t *)(base + o_fill);
    w->list = (int32_t *)(base + o_list);
    w->meta = (int32_t *)(base + o_meta);
}

__global__ void vox_init_kernel(int32_t *cell_first, size_t cells, int32_t *cnt, int32_t *fillc,
                                int max_voxels, int32_t *meta, int N) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (i == 0) {
        meta[0] = N;  // cutoff: no break
        meta[1] = 0;
    }
    for (size_t j = i; j < cells; j += stride) cell_first[j] = 0x7fffffff;
    for (size_t j = i; j < (size_t)max_voxels; j += stride) {
        cnt[j] = 0;
        fillc[j] = 0;
    }
}

__global__ void __launch_bounds__(256)
vox_cell_kernel(const float *__restrict__ points, int N, int F, const VoxGeom g,
                int32_t *__restrict__ pt_cell, int32_t *__restrict__ cell_first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    int coor[3];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        // floor((p - lo) / vs) with separately rounded fp32 ops (pc_ops.py:34)
        const float c = floorf(__fdiv_rn(__fsub_rn(points[(size_t)i * F + j], g.lo[j]), g.vs[j]));
        if (!(c >= 0.f) || !(c < (float)g.grid[j])) ok = false;  // also rejects NaN
        coor[g.reverse ? 2 - j : j] = (int)c;
    }
    int cell = -1;
    if (ok) {
        cell = (coor[0] * g.shape[1] + coor[1]) * g.shape[2] + coor[2];
        atomicMin(cell_first + cell, i);
    }
    pt_cell[i] = cell;
}

// Exclusive prefix of one value per thread over a 1024-thread CTA (32 warps): returns the prefix,
// *total receives the CTA sum.  Two __syncthreads; s_warp is a 32-int scratch.
__device__ __forceinline__ int block_excl_scan_1024(int v, int *s_warp, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();  // previous use of s_warp is over
    if (lane == 31) s_warp[warp] = incl;
    __syncthreads();
    const int wsum = s_warp[lane];
    int wincl = wsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, wincl, o);
        if (lane >= o) wincl += t;
    }
    *total = __shfl_sync(0xffffffffu, wincl, 31);
    const int warp_excl = __shfl_sync(0xffffffffu, wincl - wsum, warp);
    return warp_excl + incl - v;
}

constexpr int kScanEPT = 16;  // consecutive elements per thread and pass: all their loads in flight at once

// Single CTA: in-order exclusive scan of the opener flags; writes voxel ids of openers, the
// coordinates of the first max_voxels voxels, the break position and the opener total.  Each thread
// owns kScanEPT CONSECUTIVE points per pass (two batches of independent loads, a register prefix, one
// block scan of the thread totals), so 20 000 points are two passes instead of twenty dependent
// load -> scan -> barrier rounds.
__global__ void __launch_bounds__(1024)
vox_scan_kernel(const int32_t *__restrict__ pt_cell, const int32_t *__restrict__ cell_first, int N,
                const VoxGeom g, int max_voxels, int32_t *__restrict__ pt_vid,
                int32_t *__restrict__ coors, int32_t *__restrict__ meta) {
    __shared__ int s_warp[32];
    const int tid = threadIdx.x;
    int carry = 0;
    for (int base = 0; base < N; base += 1024 * kScanEPT) {
        const int i0 = base + tid * kScanEPT;
        int cell[kScanEPT], first[kScanEPT];
#pragma unroll
        for (int e = 0; e < kScanEPT; ++e) cell[e] = (i0 + e < N) ? pt_cell[i0 + e] : -1;
#pragma unroll
        for (int e = 0; e < kScanEPT; ++e) first[e] = cell[e] >= 0 ? cell_first[cell[e]] : -1;
        int local = 0;
        unsigned flags = 0u;
#pragma unroll
        for (int e = 0; e < kScanEPT; ++e) {
            const bool f = cell[e] >= 0 && first[e] == i0 + e;
            flags |= (f ? 1u : 0u) << e;
            local += f ? 1 : 0;
        }
        int total;
        int vid = carry + block_excl_scan_1024(local, s_warp, &total);
#pragma unroll
        for (int e = 0; e < kScanEPT; ++e) {
            const int i = i0 + e;
            if (i < N) {
                pt_vid[i] = vid;
                if ((flags >> e) & 1u) {
                    if (vid < max_voxels) {
                        const int c2 = cell[e] % g.shape[2];
                        const int c1 = (cell[e] / g.shape[2]) % g.shape[1];
                        const int c0 = cell[e] / (g.shape[2] * g.shape[1]);
                        coors[vid * 3 + 0] = c0;
                        coors[vid * 3 + 1] = c1;
                        coors[vid * 3 + 2] = c2;
                    } else if (vid == max_voxels) {
                        meta[0] = i;  // the reference breaks here (pc_ops.py:44-45)
                    }
                    ++vid;
                }
            }
        }
        carry += total;
    }
    if (tid == 0) meta[1] = carry;
}

__global__ void __launch_bounds__(256)
vox_count_kernel(int32_t *__restrict__ pt_cell, const int32_t *__restrict__ cell_first,
                 const int32_t *__restrict__ pt_vid, const int32_t *__restrict__ meta, int N,
                 int32_t *__restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int cell = pt_cell[i];
    int v = -1;
    if (cell >= 0 && i < meta[0]) {
        v = pt_vid[cell_first[cell]];  // voxel id of the cell's opener (opened before the break)
        atomicAdd(cnt + v, 1);
    }
    pt_cell[i] = v;  // from here on: voxel id of the point, -1 = dropped
}

// Single CTA: exclusive scan of cnt over the voxels -> bucket offsets; final num_points and the
// zero rows of coors / num_points beyond voxel_num.  Same consecutive-elements-per-thread scheme.
__global__ void __launch_bounds__(1024)
vox_offsets_kernel(const int32_t *__restrict__ cnt, const int32_t *__restrict__ meta,
                   int max_voxels, int max_points, int32_t *__restrict__ off,
                   int32_t *__restrict__ num_points, int32_t *__restrict__ coors,
                   int32_t *__restrict__ voxel_num_out) {
    __shared__ int s_warp[32];
    const int tid = threadIdx.x;
    const int vnum = min(meta[1], max_voxels);
    if (tid == 0) *voxel_num_out = vnum;
    int carry = 0;
    for (int base = 0; base < max_voxels; base += 1024 * kScanEPT) {
        const int v0 = base + tid * kScanEPT;
        int c[kScanEPT];
        int local = 0;
#pragma unroll
        for (int e = 0; e < kScanEPT; ++e) {
            c[e] = (v0 + e < vnum) ? cnt[v0 + e] : 0;
            local += c[e];
        }
        int total;
        int run = carry + block_excl_scan_1024(local, s_warp, &total);
#pragma unroll
        for (int e = 0; e < kScanEPT; ++e) {
            const int v = v0 + e;
            if (v < max_voxels) {
                off[v] = run;
                num_points[v] = min(c[e], max_points);
                if (v >= vnum) {
                    coors[v * 3 + 0] = 0;
                    coors[v * 3 + 1] = 0;
                    coors[v * 3 + 2] = 0;
                }
            }
            run += c[e];
        }
        carry += total;
    }
}

__global__ void __launch_bounds__(256)
vox_bucket_kernel(const int32_t *__restrict__ pt_vox, const int32_t *__restrict__ off, int N,
                  int32_t *__restrict__ fillc, int32_t *__restrict__ list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int v = pt_vox[i];
    if (v >= 0) list[off[v] + atomicAdd(fillc + v, 1)] = i;
}

// One warp per voxel slot.  Dynamic smem: [warps][max_points] selected + [warps][max_points] sorted.
constexpr int kVoxWarps = 8;
__global__ void __launch_bounds__(kVoxWarps * 32)
vox_write_kernel(const float *__restrict__ points, int N, int F, const int32_t *__restrict__ cnt,
                 const int32_t *__restrict__ off, const int32_t *__restrict__ list,
                 const int32_t *__restrict__ voxel_num, int max_voxels, int max_points,
                 float *__restrict__ voxels) {
    extern __shared__ int32_t s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t *s_sel = s_dyn + (size_t)warp * 2 * max_points;
    int32_t *s_sorted = s_sel + max_points;
    const int v = blockIdx.x * kVoxWarps + warp;
    if (v >= max_voxels) return;
    const int vnum = *voxel_num;
    int nsel = 0;
    if (v < vnum) {
        const int n = cnt[v];
        const int32_t *seg = list + off[v];
        int thresh = 0x7fffffff;
        if (n > max_points) {
            // smallest T with |{e in seg : e <= T}| >= max_points (indices are unique)
            int lo = 0, hi = N - 1;
            while (lo < hi) {
                const int mid = lo + ((hi - lo) >> 1);
                int c = 0;
                for (int e = lane; e < n; e += 32) c += (seg[e] <= mid);
                c = __reduce_add_sync(0xffffffffu, c);
                if (c >= max_points) hi = mid; else lo = mid + 1;
            }
            thresh = lo;
        }
        // compact the selected indices into shared memory (arbitrary order)
        for (int e0 = 0; e0 < n; e0 += 32) {
            const int e = e0 + lane;
            const int val = (e < n) ? seg[e] : 0x7fffffff;
            const bool take = (e < n) && (val <= thresh);
            const unsigned m = __ballot_sync(0xffffffffu, take);
            if (take) s_sel[nsel + __popc(m & ((1u << lane) - 1u))] = val;
            nsel += __popc(m);
        }
        __syncwarp();
        // rank by counting -> input order
        for (int e = lane; e < nsel; e += 32) {
            const int val = s_sel[e];
            int r = 0;
            for (int f = 0; f < nsel; ++f) r += (s_sel[f] < val);
            s_sorted[r] = val;
        }
        __syncwarp();
    }
    float *vout = voxels + (size_t)v * max_points * F;
    const bool vec = (F % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(voxels) & 15u) == 0);
    if (vec) {
        const int F4 = F / 4;
        const float4 *pin = reinterpret_cast<const float4 *>(points);
        float4 *po = reinterpret_cast<float4 *>(vout);
        for (int e = lane; e < max_points * F4; e += 32) {
            const int r = e / F4, c = e - r * F4;
            po[e] = (r < nsel) ? pin[(size_t)s_sorted[r] * F4 + c] : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else {
        for (int e = lane; e < max_points * F; e += 32) {
            const int r = e / F, c = e - r * F;
            vout[e] = (r < nsel) ? points[(size_t)s_sorted[r] * F + c] : 0.f;
        }
    }
}

// ====================================================================== PillarFeatureNet
constexpr int kPfnWarps = 8;
constexpr int kPfnMaxCin = 16;

// One warp per pillar (grid-stride).  Per pillar: sequential fp32 mean of xyz over the real
// points (same order as NumPy/Paddle sum over the T axis with zero padding), decoration,
// x = dec * W (+bias), running max/min over the T rows (padding rows contribute x_pad), and
// per-channel sum / sum^2 for the BatchNorm1D statistics (padding rows included).
template <int CPL>
__global__ void __launch_bounds__(kPfnWarps * 32)
pfn_main_kernel(const float *__restrict__ features, const int32_t *__restrict__ num_voxels,
                const int32_t *__restrict__ coors, int P, int T, int F, float vx, float vy,
                float x_off, float y_off, const float *__restrict__ weight,
                const float *__restrict__ bias, int cout, const int32_t *__restrict__ num_valid,
                float *__restrict__ pmax, float *__restrict__ pmin,
                double *__restrict__ partial /* [gridDim.x][2][cout] or null */) {
    extern __shared__ float s_dynf[];
    const int cin = F + 5;
    float *s_w = s_dynf;                                   // [cin][cout]
    float *s_pts = s_w + cin * cout;                       // [warps][T*F] raw rows of the pillar
    size_t red_off = (size_t)cin * cout + (size_t)kPfnWarps * T * F;
    red_off = (red_off + 1) & ~(size_t)1;  // 8-byte alignment for the fp64 scratch
    double *s_red = reinterpret_cast<double *>(s_dynf + red_off);                   // [warps][2][cout]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < cin * cout; i += blockDim.x) s_w[i] = weight[i];
    __syncthreads();
    const int Pv = num_valid ? min(*num_valid, P) : P;
    float *mypts = s_pts + (size_t)warp * T * F;

    float bch[CPL];
    double ssum[CPL], ssq[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const int c = lane + 32 * q;
        bch[q] = (bias && c < cout) ? bias[c] : 0.f;
        ssum[q] = 0.0;
        ssq[q] = 0.0;
    }

    // Software pipeline over the warp's pillars: the NEXT pillar's point count, cell and first 32 floats (8 points
    // at F = 4 -- most pillars hold fewer) are requested before this pillar is computed, so the three dependent
    // global round trips per pillar (count -> rows -> ...) overlap the arithmetic instead of serialising 48 000
    // times (the kernel was latency bound at ~5 us per pillar and warp).
    const int pstride = gridDim.x * kPfnWarps;
    const int TF = T * F;
    int n_nx = 0, cx_nx = 0, cy_nx = 0;
    float v_nx = 0.f;
    auto prefetch = [&](int pp) {
        if (pp < Pv) {
            n_nx = num_voxels[pp];
            cx_nx = coors[(size_t)pp * 4 + 3];
            cy_nx = coors[(size_t)pp * 4 + 2];
            v_nx = lane < TF ? features[(size_t)pp * TF + lane] : 0.f;
        }
    };
    prefetch(blockIdx.x * kPfnWarps + warp);
    for (int p = blockIdx.x * kPfnWarps + warp; p < Pv; p += pstride) {
        const int nraw = n_nx, cxi = cx_nx, cyi = cy_nx;
        const float v0 = v_nx;
        const int n = min(max(nraw, 0), T);
        prefetch(p + pstride);
        const float *src = features + (size_t)p * TF;
        if (lane < n * F) mypts[lane] = v0;
        for (int i = 32 + lane; i < n * F; i += 32) mypts[i] = src[i];
        __syncwarp();
        // sequential fp32 sums of x,y,z (lanes 0..2), divided by num_voxels as in pillars.py:82
        float mean = 0.f;
        if (lane < 3) {
            float s = 0.f;
            for (int t = 0; t < n; ++t) s = __fadd_rn(s, mypts[t * F + lane]);
            mean = __fdiv_rn(s, (float)nraw);
        }
        const float mx = __shfl_sync(0xffffffffu, mean, 0);
        const float my = __shfl_sync(0xffffffffu, mean, 1);
        const float mz = __shfl_sync(0xffffffffu, mean, 2);
        // pillar centre: coors[:,3]*vx + x_offset, coors[:,2]*vy + y_offset (pillars.py:87-88)
        const float ccx = __fadd_rn(__fmul_rn((float)cxi, vx), x_off);
        const float ccy = __fadd_rn(__fmul_rn((float)cyi, vy), y_off);

        float vmax[CPL], vmin[CPL], fs[CPL], fq[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            vmax[q] = -INFINITY;
            vmin[q] = INFINITY;
            fs[q] = 0.f;
            fq[q] = 0.f;
        }
        for (int t = 0; t < n; ++t) {
            const float *pt = mypts + t * F;
            float acc[CPL];
#pragma unroll
            for (int q = 0; q < CPL; ++q) acc[q] = bch[q];
            // decorated row = [features(F), f_cluster(3), f_center(2)]
            for (int k = 0; k < cin; ++k) {
                float d;
                if (k < F) d = pt[k];
                else if (k == F) d = __fsub_rn(pt[0], mx);
                else if (k == F + 1) d = __fsub_rn(pt[1], my);
                else if (k == F + 2) d = __fsub_rn(pt[2], mz);
                else if (k == F + 3) d = __fsub_rn(pt[0], ccx);
                else d = __fsub_rn(pt[1], ccy);
#pragma unroll
                for (int q = 0; q < CPL; ++q) {
                    const int c = lane + 32 * q;
                    if (c < cout) acc[q] = fmaf(d, s_w[k * cout + c], acc[q]);
                }
            }
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                vmax[q] = fmaxf(vmax[q], acc[q]);
                vmin[q] = fminf(vmin[q], acc[q]);
                fs[q] += acc[q];
                fq[q] = fmaf(acc[q], acc[q], fq[q]);
            }
        }
        const float npad = (float)(T - n);
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const int c = lane + 32 * q;
            if (c >= cout) continue;
            if (n < T) {  // zeroed padding rows go through the Linear too: x_pad = bias
                vmax[q] = fmaxf(vmax[q], bch[q]);
                vmin[q] = fminf(vmin[q], bch[q]);
            }
            pmax[(size_t)p * cout + c] = vmax[q];
            pmin[(size_t)p * cout + c] = vmin[q];
            ssum[q] += (double)fs[q] + (double)npad * (double)bch[q];
            ssq[q] += (double)fq[q] + (double)npad * (double)bch[q] * (double)bch[q];
        }
        __syncwarp();
    }
    if (partial != nullptr) {
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const int c = lane + 32 * q;
            if (c < cout) {
                s_red[(warp * 2 + 0) * cout + c] = ssum[q];
                s_red[(warp * 2 + 1) * cout + c] = ssq[q];
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * cout; i += blockDim.x) {
            double tot = 0.0;
            for (int w = 0; w < kPfnWarps; ++w) tot += s_red[w * 2 * cout + i];
            partial[(size_t)blockIdx.x * 2 * cout + i] = tot;
        }
    }
}

// Single CTA: fixed-order reduction of the block partials -> scale / shift (+ mean / var).  The
// 2*cout columns are split over all 1024 threads (slices of the block rows, independent loads),
// the slices are combined through shared memory in a fixed order -> deterministic.
constexpr int kPfnBnThreads = 1024;
__global__ void __launch_bounds__(kPfnBnThreads)
pfn_bn_kernel(const double *__restrict__ partial, int nblocks, int P, int T,
              const int32_t *__restrict__ num_valid, const float *__restrict__ gamma,
              const float *__restrict__ beta, const float *__restrict__ rm,
              const float *__restrict__ rv, int bn_mode, float eps, int cout,
              float *__restrict__ scale, float *__restrict__ shift, float *__restrict__ mean_out,
              float *__restrict__ var_out) {
    __shared__ double s_red[kPfnBnThreads];
    __shared__ double s_sum[2 * 512];  // S | Q per channel (cout <= 512)
    const int tid = threadIdx.x;
    const int Pv = num_valid ? min(*num_valid, P) : P;
    if (bn_mode == PAPC_BN_BATCH) {
        const int C2 = 2 * cout;
        for (int cbase = 0; cbase < C2; cbase += kPfnBnThreads) {
            const int ncol = min(C2 - cbase, kPfnBnThreads);
            int slices = kPfnBnThreads / ncol;          // >= 1
            const int col = tid % ncol, sl = tid / ncol;
            double acc = 0.0;
            if (sl < slices) {
                const double *p = partial + cbase + col;
#pragma unroll 8
                for (int b = sl; b < nblocks; b += slices) acc += p[(size_t)b * C2];
            }
            s_red[tid] = acc;
            __syncthreads();
            if (tid < ncol) {
                double t = 0.0;
                for (int q = 0; q < slices; ++q) t += s_red[q * ncol + tid];
                s_sum[cbase + tid] = t;
            }
            __syncthreads();
        }
    }
    for (int c = tid; c < cout; c += kPfnBnThreads) {
        double sc = 1.0, sh = 0.0;
        if (bn_mode == PAPC_BN_BATCH) {
            const double S = s_sum[c], Q = s_sum[cout + c];
            const double count = (double)Pv * (double)T;
            const double mean = count > 0 ? S / count : 0.0;
            double var = count > 0 ? Q / count - mean * mean : 0.0;
            var = var > 0.0 ? var : 0.0;
            const double g = gamma ? (double)gamma[c] : 1.0;
            const double b2 = beta ? (double)beta[c] : 0.0;
            sc = g / sqrt(var + (double)eps);
            sh = b2 - mean * sc;
            if (mean_out) mean_out[c] = (float)mean;
            if (var_out) var_out[c] = (float)var;
        } else if (bn_mode == PAPC_BN_RUNNING) {
            const double g = gamma ? (double)gamma[c] : 1.0;
            const double b2 = beta ? (double)beta[c] : 0.0;
            sc = g / sqrt((double)rv[c] + (double)eps);
            sh = b2 - (double)rm[c] * sc;
        }
        scale[c] = (float)sc;
        shift[c] = (float)sh;
    }
}

__global__ void __launch_bounds__(256)
pfn_apply_kernel(const float *__restrict__ pmax, const float *__restrict__ pmin,
                 const float *__restrict__ scale, const float *__restrict__ shift, int P, int cout,
                 const int32_t *__restrict__ num_valid, float *__restrict__ out) {
    const int Pv = num_valid ? min(*num_valid, P) : P;
    const size_t total = (size_t)P * cout;
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; e < total; e += stride) {
        const int c = (int)(e % cout);
        const size_t p = e / cout;
        float r = 0.f;
        if (p < (size_t)Pv) {
            const float sc = scale[c];
            const float v = sc >= 0.f ? pmax[e] : pmin[e];
            r = fmaxf(fmaf(v, sc, shift[c]), 0.f);
        }
        out[e] = r;
    }
}

// ====================================================================== PointPillarsScatter
__global__ void __launch_bounds__(256)
scatter_map_kernel(const int32_t *__restrict__ coords, int P, const int32_t *__restrict__ num_valid,
                   int batch, int ny, int nx, int32_t *__restrict__ map) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int Pv = num_valid ? min(*num_valid, P) : P;
    if (p >= Pv) return;
    const int b = coords[p * 4 + 0], y = coords[p * 4 + 2], x = coords[p * 4 + 3];
    if (b < 0 || b >= batch || y < 0 || y >= ny || x < 0 || x >= nx) return;
    atomicMax(map + ((size_t)b * ny + y) * nx + x, p);  // duplicates: the last pillar wins
}

constexpr int kScatCSplit = 16;  // channels handled per thread
__global__ void __launch_bounds__(256)
scatter_canvas_kernel(const float *__restrict__ feat, const int32_t *__restrict__ map, int C,
                      int batch, size_t cells /* ny*nx */, float *__restrict__ canvas) {
    // thread -> (b, channel block, 4 consecutive cells); requires cells % 4 == 0
    const size_t q4 = cells / 4;
    const int cblocks = ceil_div(C, kScatCSplit);
    const size_t total = (size_t)batch * cblocks * q4;
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; e < total; e += stride) {
        const size_t cell4 = e % q4;
        const size_t bc = e / q4;
        const int cb = (int)(bc % cblocks);
        const size_t b = bc / cblocks;
        const int4 pm = *reinterpret_cast<const int4 *>(map + b * cells + cell4 * 4);
        const int c1 = min(C, (cb + 1) * kScatCSplit);
        for (int c = cb * kScatCSplit; c < c1; ++c) {
            float4 v;
            v.x = pm.x >= 0 ? feat[(size_t)pm.x * C + c] : 0.f;
            v.y = pm.y >= 0 ? feat[(size_t)pm.y * C + c] : 0.f;
            v.z = pm.z >= 0 ? feat[(size_t)pm.z * C + c] : 0.f;
            v.w = pm.w >= 0 ? feat[(size_t)pm.w * C + c] : 0.f;
            *reinterpret_cast<float4 *>(canvas + (b * C + c) * cells + cell4 * 4) = v;
        }
    }
}

__global__ void __launch_bounds__(256)
scatter_canvas_scalar_kernel(const float *__restrict__ feat, const int32_t *__restrict__ map, int C,
                             int batch, size_t cells, float *__restrict__ canvas) {
    const size_t total = (size_t)batch * C * cells;
    size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (; e < total; e += stride) {
        const size_t cell = e % cells;
        const size_t bc = e / cells;
        const int c = (int)(bc % C);
        const size_t b = bc / C;
        const int p = map[b * cells + cell];
        canvas[e] = p >= 0 ? feat[(size_t)p * C + c] : 0.f;
    }
}

static unsigned blocks_for(size_t total, int threads, int per_sm) {
    size_t b = (total + threads - 1) / threads;
    const size_t cap = (size_t)kNumSMs * per_sm;
    if (b > cap) b = cap;
    return (unsigned)(b ? b : 1);
}

}  // namespace papc

using namespace papc;

extern "C" size_t papc_voxelize_workspace_bytes(int N, const float *voxel_size_host,
                                                const float *coors_range_host, int max_voxels) {
    VoxGeom g;
    if (N < 0 || max_voxels <= 0 || !voxel_size_host || !coors_range_host) return 0;
    if (!make_geom(voxel_size_host, coors_range_host, 1, &g)) return 0;
    VoxWs w;
    carve_vox_ws(nullptr, N > 0 ? N : 1, (size_t)g.grid[0] * g.grid[1] * g.grid[2], max_voxels, &w);
    return w.total;
}

extern "C" int papc_voxelize_f32(const float *points, int N, int F, const float *voxel_size_host,
                                 const float *coors_range_host, int max_points, int reverse_index,
                                 int max_voxels, float *voxels, int32_t *coors, int32_t *num_points,
                                 int32_t *voxel_num, void *workspace, size_t workspace_bytes,
                                 papc_stream_t stream) {
    if (N < 0 || F < 3 || max_points <= 0 || max_voxels <= 0) return PAPC_EINVAL;
    if (!voxel_size_host || !coors_range_host || !voxels || !coors || !num_points || !voxel_num)
        return PAPC_EINVAL;
    if (N > 0 && !points) return PAPC_EINVAL;
    VoxGeom g;
    if (!make_geom(voxel_size_host, coors_range_host, reverse_index ? 1 : 0, &g)) return PAPC_EINVAL;
    const size_t cells = (size_t)g.grid[0] * g.grid[1] * g.grid[2];
    if (cells > 0x7fffffffULL) return PAPC_EUNSUPPORTED;
    const size_t smem = (size_t)kVoxWarps * 2 * max_points * sizeof(int32_t);
    if (smem > 200 * 1024) return PAPC_EUNSUPPORTED;
    VoxWs w;
    carve_vox_ws(nullptr, N > 0 ? N : 1, cells, max_voxels, &w);
    if (!workspace || workspace_bytes < w.total) return PAPC_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PAPC_EINVAL;
    carve_vox_ws(reinterpret_cast<char *>(workspace), N > 0 ? N : 1, cells, max_voxels, &w);
    cudaStream_t st = as_stream(stream);

    const size_t init_n = cells > (size_t)max_voxels ? cells : (size_t)max_voxels;
    vox_init_kernel<<<blocks_for(init_n, 256, 8), 256, 0, st>>>(w.cell_first, cells, w.cnt, w.fillc,
                                                                max_voxels, w.meta, N);
    PAPC_LAUNCH_CHECK();
    if (N > 0) {
        vox_cell_kernel<<<ceil_div(N, 256), 256, 0, st>>>(points, N, F, g, w.pt_cell, w.cell_first);
        PAPC_LAUNCH_CHECK();
        vox_scan_kernel<<<1, 1024, 0, st>>>(w.pt_cell, w.cell_first, N, g, max_voxels, w.pt_vid, coors,
                                            w.meta);
        PAPC_LAUNCH_CHECK();
        vox_count_kernel<<<ceil_div(N, 256), 256, 0, st>>>(w.pt_cell, w.cell_first, w.pt_vid, w.meta, N,
                                                           w.cnt);
        PAPC_LAUNCH_CHECK();
    }
    vox_offsets_kernel<<<1, 1024, 0, st>>>(w.cnt, w.meta, max_voxels, max_points, w.off, num_points,
                                           coors, voxel_num);
    PAPC_LAUNCH_CHECK();
    if (N > 0) {
        vox_bucket_kernel<<<ceil_div(N, 256), 256, 0, st>>>(w.pt_cell, w.off, N, w.fillc, w.list);
        PAPC_LAUNCH_CHECK();
    }
    if (smem > 48 * 1024)
        PAPC_CUDA_TRY(cudaFuncSetAttribute(vox_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                           (int)smem));
    vox_write_kernel<<<ceil_div(max_voxels, kVoxWarps), kVoxWarps * 32, smem, st>>>(
        points, N, F, w.cnt, w.off, w.list, voxel_num, max_voxels, max_points, voxels);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

// ---------------------------------------------------------------------------------- PFN
namespace {

// ====================================================================== generic PFNLayer (any position)
// pillars.py:29-41 for a layer that is NOT the fused (decorate + single last layer) case above: a first layer
// with the distance channel, a non-last layer (its output is [x | max-repeat], :36-40) or a later layer reading
// a materialised [P,T,cin] tensor.  Two kernels around the shared BatchNorm reduction (pfn_bn_kernel):
//   pfn_gen_linear_kernel : one warp per pillar, rows of the pillar one at a time (lanes load the row, the
//                           reduction broadcasts it with shuffles), y = x W (+ bias) -> y [P,T,u] and the
//                           per-block sum / sum^2 over ALL P*T rows (padding rows included, as the reference);
//   pfn_gen_finish_kernel : relu(scale * y + shift), max over T, then either [P,u] (last layer) or
//                           [P,T,2u] = [x | repeat(max)].
constexpr int kPfnGenMaxCin = 128;
template <int CPL, bool DECORATE>
__global__ void __launch_bounds__(kPfnWarps * 32)
pfn_gen_linear_kernel(const float *__restrict__ x, const int32_t *__restrict__ num_voxels,
                      const int32_t *__restrict__ coors, int P, int T, int F, int with_distance, float vx, float vy,
                      float x_off, float y_off, const float *__restrict__ weight, const float *__restrict__ bias,
                      int cin, int u, const int32_t *__restrict__ num_valid, float *__restrict__ y,
                      double *__restrict__ partial) {
    extern __shared__ float s_dynf[];
    float *s_w = s_dynf;                                   // [cin][u]
    size_t red_off = ((size_t)cin * u + 1) & ~(size_t)1;
    double *s_red = reinterpret_cast<double *>(s_dynf + red_off);   // [warps][2][u]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < cin * u; i += blockDim.x) s_w[i] = weight[i];
    __syncthreads();
    const int Pv = num_valid ? min(*num_valid, P) : P;
    float bch[CPL];
    double ssum[CPL], ssq[CPL];
#pragma unroll
    for (int q = 0; q < CPL; ++q) {
        const int c = lane + 32 * q;
        bch[q] = (bias && c < u) ? bias[c] : 0.f;
        ssum[q] = 0.0;
        ssq[q] = 0.0;
    }
    constexpr int XR = kPfnGenMaxCin / 32;
    for (int p = blockIdx.x * kPfnWarps + warp; p < Pv; p += gridDim.x * kPfnWarps) {
        float mx = 0.f, my = 0.f, mz = 0.f, ccx = 0.f, ccy = 0.f;
        int n = T;
        if (DECORATE) {
            n = min(max(num_voxels[p], 0), T);
            const float *src = x + (size_t)p * T * F;
            float mean = 0.f;
            if (lane < 3) {   // sequential fp32 sum over the T axis (zero padding adds nothing), pillars.py:82
                float sacc = 0.f;
                for (int t = 0; t < n; ++t) sacc = __fadd_rn(sacc, src[t * F + lane]);
                mean = __fdiv_rn(sacc, (float)num_voxels[p]);
            }
            mx = __shfl_sync(0xffffffffu, mean, 0);
            my = __shfl_sync(0xffffffffu, mean, 1);
            mz = __shfl_sync(0xffffffffu, mean, 2);
            ccx = __fadd_rn(__fmul_rn((float)coors[p * 4 + 3], vx), x_off);
            ccy = __fadd_rn(__fmul_rn((float)coors[p * 4 + 2], vy), y_off);
        }
        float fs[CPL], fq[CPL];
#pragma unroll
        for (int q = 0; q < CPL; ++q) { fs[q] = 0.f; fq[q] = 0.f; }
        for (int t = 0; t < T; ++t) {
            // the row (cin values) spread over the lanes: lane holds elements lane, lane+32, ...
            float xr[XR];
#pragma unroll
            for (int j = 0; j < XR; ++j) {
                const int k = lane + 32 * j;
                float d = 0.f;
                if (k < cin) {
                    if (!DECORATE) {
                        d = x[((size_t)p * T + t) * cin + k];
                    } else if (t < n) {   // padding rows stay zero (features *= mask, pillars.py:102)
                        const float *pt = x + ((size_t)p * T + t) * F;
                        if (k < F) d = pt[k];
                        else if (k == F) d = __fsub_rn(pt[0], mx);
                        else if (k == F + 1) d = __fsub_rn(pt[1], my);
                        else if (k == F + 2) d = __fsub_rn(pt[2], mz);
                        else if (k == F + 3) d = __fsub_rn(pt[0], ccx);
                        else if (k == F + 4) d = __fsub_rn(pt[1], ccy);
                        else if (with_distance && k == F + 5)   // paddle.norm(features[:, :, :3], 2, 2), :93
                            d = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(pt[0], pt[0]), __fmul_rn(pt[1], pt[1])),
                                                     __fmul_rn(pt[2], pt[2])));
                    }
                }
                xr[j] = d;
            }
            float acc[CPL];
#pragma unroll
            for (int q = 0; q < CPL; ++q) acc[q] = bch[q];
#pragma unroll
            for (int j = 0; j < XR; ++j) {
                if (32 * j >= cin) break;
                const int kend = min(32, cin - 32 * j);
                for (int kk = 0; kk < kend; ++kk) {
                    const float d = __shfl_sync(0xffffffffu, xr[j], kk);
                    const float *wr = s_w + (size_t)(32 * j + kk) * u;
#pragma unroll
                    for (int q = 0; q < CPL; ++q) {
                        const int c = lane + 32 * q;
                        if (c < u) acc[q] = fmaf(d, wr[c], acc[q]);
                    }
                }
            }
#pragma unroll
            for (int q = 0; q < CPL; ++q) {
                const int c = lane + 32 * q;
                if (c < u) {
                    y[((size_t)p * T + t) * u + c] = acc[q];
                    fs[q] += acc[q];
                    fq[q] = fmaf(acc[q], acc[q], fq[q]);
                }
            }
        }
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            ssum[q] += (double)fs[q];
            ssq[q] += (double)fq[q];
        }
    }
    if (partial != nullptr) {
#pragma unroll
        for (int q = 0; q < CPL; ++q) {
            const int c = lane + 32 * q;
            if (c < u) {
                s_red[(warp * 2 + 0) * u + c] = ssum[q];
                s_red[(warp * 2 + 1) * u + c] = ssq[q];
            }
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 2 * u; i += blockDim.x) {
            double tot = 0.0;
            for (int w = 0; w < kPfnWarps; ++w) tot += s_red[w * 2 * u + i];
            partial[(size_t)blockIdx.x * 2 * u + i] = tot;
        }
    }
}

__global__ void __launch_bounds__(kPfnWarps * 32)
pfn_gen_finish_kernel(const float *__restrict__ y, const float *__restrict__ scale, const float *__restrict__ shift,
                      int P, int T, int u, int last_layer, const int32_t *__restrict__ num_valid,
                      float *__restrict__ out) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int Pv = num_valid ? min(*num_valid, P) : P;
    for (int p = blockIdx.x * kPfnWarps + warp; p < P; p += gridDim.x * kPfnWarps) {
        for (int c = lane; c < u; c += 32) {
            const float sc = scale[c], sh = shift[c];
            float m = -INFINITY;
            if (p < Pv) {
                for (int t = 0; t < T; ++t) {
                    const float v = fmaxf(fmaf(y[((size_t)p * T + t) * u + c], sc, sh), 0.f);
                    m = fmaxf(m, v);
                    if (!last_layer) out[((size_t)p * T + t) * 2 * u + c] = v;
                }
            } else {
                m = 0.f;   // rows beyond *num_valid are zero
                if (!last_layer)
                    for (int t = 0; t < T; ++t) out[((size_t)p * T + t) * 2 * u + c] = 0.f;
            }
            if (last_layer) out[(size_t)p * u + c] = m;
            else
                for (int t = 0; t < T; ++t) out[((size_t)p * T + t) * 2 * u + u + c] = m;
        }
    }
}

struct PfnGenWs {
    size_t y, partial, scale, shift, total;
    int nblocks;
};
static void plan_pfn_gen(int P, int T, int u, PfnGenWs *w) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    int nb = ceil_div(P > 0 ? P : 1, kPfnWarps);
    if (nb > kNumSMs * 4) nb = kNumSMs * 4;
    w->nblocks = nb;
    w->y = take((size_t)P * T * u * 4);
    w->partial = take((size_t)nb * 2 * u * 8);
    w->scale = take((size_t)u * 4);
    w->shift = take((size_t)u * 4);
    w->total = off;
}
}  // namespace

extern "C" size_t papc_pfn_layer_workspace_bytes(int P, int T, int units) {
    if (P < 0 || T <= 0 || units <= 0) return 0;
    PfnGenWs w;
    plan_pfn_gen(P, T, units, &w);
    return w.total;
}

extern "C" int papc_pfn_layer_f32(const float *x, int decorate, int with_distance, const int32_t *num_voxels,
                                  const int32_t *coors, int P, int T, int F, float vx, float vy, float x_offset,
                                  float y_offset, const float *weight, const float *bias, const float *gamma,
                                  const float *beta, const float *running_mean, const float *running_var,
                                  int bn_mode, float eps, int units, int last_layer, const int32_t *num_valid,
                                  float *out, float *batch_mean, float *batch_var, void *workspace,
                                  size_t workspace_bytes, papc_stream_t stream) {
    if (P < 0 || T <= 0 || F < 1 || units <= 0) return PAPC_EINVAL;
    if (bn_mode != PAPC_BN_BATCH && bn_mode != PAPC_BN_RUNNING && bn_mode != PAPC_BN_NONE) return PAPC_EINVAL;
    if (decorate && F < 3) return PAPC_EINVAL;
    const int cin = decorate ? F + 5 + (with_distance ? 1 : 0) : F;
    if (cin > kPfnGenMaxCin || units > 128) return PAPC_EUNSUPPORTED;
    if (P == 0) return PAPC_OK;
    if (!x || !weight || !out) return PAPC_EINVAL;
    if (decorate && (!num_voxels || !coors)) return PAPC_EINVAL;
    if (bn_mode == PAPC_BN_RUNNING && (!running_mean || !running_var)) return PAPC_EINVAL;
    PfnGenWs w;
    plan_pfn_gen(P, T, units, &w);
    if (!workspace || workspace_bytes < w.total) return PAPC_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PAPC_EINVAL;
    char *ws = reinterpret_cast<char *>(workspace);
    float *y = (float *)(ws + w.y);
    double *partial = (double *)(ws + w.partial);
    float *scale = (float *)(ws + w.scale), *shift = (float *)(ws + w.shift);
    cudaStream_t st = as_stream(stream);
    size_t smem = align_up((size_t)cin * units * 4, 8) + (size_t)kPfnWarps * 2 * units * 8;
    double *part_arg = (bn_mode == PAPC_BN_BATCH) ? partial : nullptr;
    const int cpl = ceil_div(units, 32);
#define PAPC_PFNG_LAUNCH(CPL, DEC)                                                                     \
    do {                                                                                              \
        auto kfn = pfn_gen_linear_kernel<CPL, DEC>;                                                   \
        if (smem > 48 * 1024)                                                                         \
            PAPC_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kfn<<<w.nblocks, kPfnWarps * 32, smem, st>>>(x, num_voxels, coors, P, T, F, with_distance, vx, vy, x_offset, \
                                                     y_offset, weight, bias, cin, units, num_valid, y, part_arg); \
    } while (0)
    if (decorate) {
        if (cpl <= 1) PAPC_PFNG_LAUNCH(1, true);
        else if (cpl <= 2) PAPC_PFNG_LAUNCH(2, true);
        else PAPC_PFNG_LAUNCH(4, true);
    } else {
        if (cpl <= 1) PAPC_PFNG_LAUNCH(1, false);
        else if (cpl <= 2) PAPC_PFNG_LAUNCH(2, false);
        else PAPC_PFNG_LAUNCH(4, false);
    }
#undef PAPC_PFNG_LAUNCH
    PAPC_LAUNCH_CHECK();
    pfn_bn_kernel<<<1, kPfnBnThreads, 0, st>>>(partial, w.nblocks, P, T, num_valid, gamma, beta, running_mean,
                                               running_var, bn_mode, eps, units, scale, shift, batch_mean, batch_var);
    PAPC_LAUNCH_CHECK();
    pfn_gen_finish_kernel<<<w.nblocks, kPfnWarps * 32, 0, st>>>(y, scale, shift, P, T, units, last_layer, num_valid, out);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

namespace {
struct PfnWs {
    size_t pmax, pmin, partial, scale, shift, total;
    int nblocks;
};
static void plan_pfn(int P, int cout, PfnWs *w) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    int nb = ceil_div(P > 0 ? P : 1, kPfnWarps);
    if (nb > kNumSMs * 4) nb = kNumSMs * 4;
    w->nblocks = nb;
    w->pmax = take((size_t)P * cout * 4);
    w->pmin = take((size_t)P * cout * 4);
    w->partial = take((size_t)nb * 2 * cout * 8);
    w->scale = take((size_t)cout * 4);
    w->shift = take((size_t)cout * 4);
    w->total = off;
}
}  // namespace

extern "C" size_t papc_pfn_workspace_bytes(int P, int cout) {
    if (P < 0 || cout <= 0) return 0;
    PfnWs w;
    plan_pfn(P, cout, &w);
    return w.total;
}

extern "C" int papc_pfn_f32(const float *features, const int32_t *num_voxels, const int32_t *coors,
                            int P, int T, int F, float vx, float vy, float x_offset, float y_offset,
                            const float *weight, const float *bias, const float *gamma,
                            const float *beta, const float *running_mean, const float *running_var,
                            int bn_mode, float eps, int cout, const int32_t *num_valid, float *out,
                            float *batch_mean, float *batch_var, void *workspace,
                            size_t workspace_bytes, papc_stream_t stream) {
    if (P < 0 || T <= 0 || F < 3 || cout <= 0) return PAPC_EINVAL;
    if (bn_mode != PAPC_BN_BATCH && bn_mode != PAPC_BN_RUNNING && bn_mode != PAPC_BN_NONE) return PAPC_EINVAL;
    if (F + 5 > kPfnMaxCin || cout > 256) return PAPC_EUNSUPPORTED;
    if (P == 0) return PAPC_OK;
    if (!features || !num_voxels || !coors || !weight || !out) return PAPC_EINVAL;
    if (bn_mode == PAPC_BN_RUNNING && (!running_mean || !running_var)) return PAPC_EINVAL;
    PfnWs w;
    plan_pfn(P, cout, &w);
    if (!workspace || workspace_bytes < w.total) return PAPC_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PAPC_EINVAL;
    char *ws = reinterpret_cast<char *>(workspace);
    float *pmax = (float *)(ws + w.pmax), *pmin = (float *)(ws + w.pmin);
    double *partial = (double *)(ws + w.partial);
    float *scale = (float *)(ws + w.scale), *shift = (float *)(ws + w.shift);
    cudaStream_t st = as_stream(stream);
    const int cin = F + 5;
    size_t smem = (size_t)cin * cout * 4 + (size_t)kPfnWarps * T * F * 4;
    smem = align_up(smem, 8) + (size_t)kPfnWarps * 2 * cout * 8;
    if (smem > 200 * 1024) return PAPC_EUNSUPPORTED;
    double *part_arg = (bn_mode == PAPC_BN_BATCH) ? partial : nullptr;
    const int cpl = ceil_div(cout, 32);
#define PAPC_PFN_LAUNCH(CPL)                                                                        \
    do {                                                                                            \
        auto kfn = pfn_main_kernel<CPL>;                                                            \
        if (smem > 48 * 1024)                                                                       \
            PAPC_CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize,    \
                                               (int)smem));                                         \
        kfn<<<w.nblocks, kPfnWarps * 32, smem, st>>>(features, num_voxels, coors, P, T, F, vx, vy,  \
                                                     x_offset, y_offset, weight, bias, cout,        \
                                                     num_valid, pmax, pmin, part_arg);              \
    } while (0)
    if (cpl <= 1) PAPC_PFN_LAUNCH(1);
    else if (cpl <= 2) PAPC_PFN_LAUNCH(2);
    else if (cpl <= 4) PAPC_PFN_LAUNCH(4);
    else PAPC_PFN_LAUNCH(8);
#undef PAPC_PFN_LAUNCH
    PAPC_LAUNCH_CHECK();
    pfn_bn_kernel<<<1, kPfnBnThreads, 0, st>>>(partial, w.nblocks, P, T, num_valid, gamma, beta, running_mean,
                                     running_var, bn_mode, eps, cout, scale, shift, batch_mean,
                                     batch_var);
    PAPC_LAUNCH_CHECK();
    pfn_apply_kernel<<<blocks_for((size_t)P * cout, 256, 8), 256, 0, st>>>(pmax, pmin, scale, shift, P,
                                                                           cout, num_valid, out);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

// ---------------------------------------------------------------------------------- scatter
extern "C" size_t papc_pillar_scatter_workspace_bytes(int batch, int ny, int nx) {
    if (batch <= 0 || ny <= 0 || nx <= 0) return 0;
    return align_up((size_t)batch * ny * nx * sizeof(int32_t), 256);
}

extern "C" int papc_pillar_scatter_f32(const float *voxel_features, const int32_t *coords, int P,
                                       int C, int batch, int ny, int nx, const int32_t *num_valid,
                                       float *canvas, void *workspace, size_t workspace_bytes,
                                       papc_stream_t stream) {
    if (P < 0 || C <= 0 || batch <= 0 || ny <= 0 || nx <= 0) return PAPC_EINVAL;
    if (!canvas || (P > 0 && (!voxel_features || !coords))) return PAPC_EINVAL;
    const size_t need = papc_pillar_scatter_workspace_bytes(batch, ny, nx);
    if (!workspace || workspace_bytes < need) return PAPC_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PAPC_EINVAL;
    cudaStream_t st = as_stream(stream);
    int32_t *map = reinterpret_cast<int32_t *>(workspace);
    const size_t cells = (size_t)ny * nx;
    PAPC_CUDA_TRY(cudaMemsetAsync(map, 0xff, (size_t)batch * cells * sizeof(int32_t), st));
    if (P > 0) {
        scatter_map_kernel<<<ceil_div(P, 256), 256, 0, st>>>(coords, P, num_valid, batch, ny, nx, map);
        PAPC_LAUNCH_CHECK();
    }
    const bool vec = (cells % 4 == 0) && ((reinterpret_cast<uintptr_t>(canvas) & 15u) == 0);
    if (vec) {
        const size_t total = (size_t)batch * ceil_div(C, kScatCSplit) * (cells / 4);
        scatter_canvas_kernel<<<blocks_for(total, 256, 8), 256, 0, st>>>(voxel_features, map, C, batch,
                                                                          cells, canvas);
    } else {
        const size_t total = (size_t)batch * C * cells;
        scatter_canvas_scalar_kernel<<<blocks_for(total, 256, 8), 256, 0, st>>>(voxel_features, map, C,
                                                                                 batch, cells, canvas);
    }
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}
