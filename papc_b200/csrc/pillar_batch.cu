// pillar_batch.cu -- the steps either side of the pillar encode:
//   * batched voxelisation with the merged batch layout of merge_second_batch (pp/data/preprocess.py:16-42):
//     a whole batch of frames in SEVEN launches (the single-frame path spends eight per frame), rows of all
//     frames compacted back to back, coordinates [P,4] = (b, z, y, x) -- the DataLoader stage the reference runs
//     on the CPU, on the device;
//   * sparse_sum_for_anchors_mask / fused_get_anchors_area (pp/libs/ops/box_np_ops.py:772-806): the pillar
//     occupancy map, its 2-D inclusive prefix sum (box_np_ops.py:748-768 / the caller's cumsum) and the
//     per-anchor area query;
//   * points_to_bev (pp/libs/ops/point_cloud/bev_ops.py:6-103).
// The voxeliser follows pillars.cu step by step (cell id -> atomicMin opener -> in-order opener scan with the
// max_voxels break -> CSR buckets -> one warp per output row); here every per-point / per-cell / per-voxel
// array carries a frame dimension and the writer resolves (frame, voxel) from the merged row index.
#include "common.cuh"

#include <limits.h>

namespace papc {

constexpr int kMaxFrames = 32;
constexpr int kMaxSlices = 64;
struct HeightLowers { float v[kMaxSlices]; };   // np.linspace(lo_z, hi_z, D, endpoint=False), computed by the host mirror
struct BatchGeom {
    float lo[3], vs[3];
    int grid[3], shape[3];
    int reverse;
    int B;
    int off[kMaxFrames + 1];   // point offsets of the frames in the concatenated cloud
    uint32_t mul2, shr2, mul1, shr1;   // division by shape[2] / shape[1] (cells < 2^31): q = umulhi(x, mul) >> shr
};
__device__ __forceinline__ uint32_t vb_div(uint32_t x, uint32_t mul, uint32_t shr) {
    return mul == 0u ? x : (__umulhi(x, mul) >> shr);
}
// cell -> (c0, c1, c2) of the [shape0, shape1, shape2] grid without integer divisions (12 000 openers x 5
// runtime div / mod were most of the opener scan's 49 us)
__device__ __forceinline__ void vb_cell_coords(const BatchGeom &g, int cell, int &c0, int &c1, int &c2) {
    const uint32_t q2 = vb_div((uint32_t)cell, g.mul2, g.shr2);
    c2 = cell - (int)q2 * g.shape[2];
    const uint32_t q1 = vb_div(q2, g.mul1, g.shr1);
    c1 = (int)q2 - (int)q1 * g.shape[1];
    c0 = (int)q1;
}
__device__ __forceinline__ int frame_of(const BatchGeom &g, int i) {
    int b = 0;
    while (b + 1 < g.B && i >= g.off[b + 1]) ++b;
    return b;
}

__global__ void vb_init_kernel(int32_t *cell_first, size_t cells_total, int32_t *cnt, int32_t *fillc, size_t vox_total,
                               int32_t *meta, const BatchGeom g) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    if (i < (size_t)g.B) {
        meta[3 * i] = g.off[i + 1] - g.off[i];   // cutoff (local point index): no break
        meta[3 * i + 1] = 0;                     // openers found by the scan
        meta[3 * i + 2] = 0;                     // segment cursor of the bucket list (vb_alloc_kernel)
    }
    for (size_t j = i; j < cells_total; j += stride) cell_first[j] = INT_MAX;
    for (size_t j = i; j < vox_total; j += stride) {
        cnt[j] = 0;
        fillc[j] = 0;
    }
}

__global__ void __launch_bounds__(256)
vb_cell_kernel(const float *__restrict__ points, int Ntot, int F, const BatchGeom g, size_t cells,
               int32_t *__restrict__ pt_cell, int32_t *__restrict__ cell_first) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ntot) return;
    const int b = frame_of(g, i);
    int coor[3];
    bool ok = true;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float c = floorf(__fdiv_rn(__fsub_rn(points[(size_t)i * F + j], g.lo[j]), g.vs[j]));
        if (!(c >= 0.f) || !(c < (float)g.grid[j])) ok = false;
        coor[g.reverse ? 2 - j : j] = (int)c;
    }
    int cell = -1;
    if (ok) {
        cell = (coor[0] * g.shape[1] + coor[1]) * g.shape[2] + coor[2];
        atomicMin(cell_first + (size_t)b * cells + cell, i - g.off[b]);
    }
    pt_cell[i] = cell;
}

__device__ __forceinline__ int vb_block_excl_scan_1024(int v, int *s_warp, int *total) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    __syncthreads();
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
constexpr int kVbEPT = 16;

// one CTA per frame: in-order scan of the opener flags (see vox_scan_kernel in pillars.cu)
__global__ void __launch_bounds__(1024)
vb_scan_kernel(const int32_t *__restrict__ pt_cell, const int32_t *__restrict__ cell_first, const BatchGeom g,
               size_t cells, int max_voxels, int32_t *__restrict__ pt_vid, int32_t *__restrict__ coors3,
               int32_t *__restrict__ meta) {
    __shared__ int s_warp[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int p0 = g.off[b], N = g.off[b + 1] - g.off[b];
    const int32_t *pc = pt_cell + p0;
    const int32_t *cf = cell_first + (size_t)b * cells;
    int32_t *pv = pt_vid + p0;
    int32_t *co = coors3 + (size_t)b * max_voxels * 3;
    int carry = 0;
    for (int base = 0; base < N; base += 1024 * kVbEPT) {
        const int i0 = base + tid * kVbEPT;
        int cell[kVbEPT], first[kVbEPT];
#pragma unroll
        for (int e = 0; e < kVbEPT; ++e) cell[e] = (i0 + e < N) ? pc[i0 + e] : -1;
#pragma unroll
        for (int e = 0; e < kVbEPT; ++e) first[e] = cell[e] >= 0 ? cf[cell[e]] : -1;
        int local = 0;
        unsigned flags = 0u;
#pragma unroll
        for (int e = 0; e < kVbEPT; ++e) {
            const bool f = cell[e] >= 0 && first[e] == i0 + e;
            flags |= (f ? 1u : 0u) << e;
            local += f ? 1 : 0;
        }
        int total;
        int vid = carry + vb_block_excl_scan_1024(local, s_warp, &total);
#pragma unroll
        for (int e = 0; e < kVbEPT; ++e) {
            const int i = i0 + e;
            if (i < N) {
                pv[i] = vid;
                if ((flags >> e) & 1u) {
                    if (vid < max_voxels) {
                        if (coors3 == nullptr) { ++vid; continue; }
                        int c0, c1, c2;
                        vb_cell_coords(g, cell[e], c0, c1, c2);
                        co[vid * 3 + 0] = c0;
                        co[vid * 3 + 1] = c1;
                        co[vid * 3 + 2] = c2;
                    } else if (vid == max_voxels) {
                        meta[3 * b] = i;   // the reference breaks here (point_cloud_ops.py:44-45)
                    }
                    ++vid;
                }
            }
        }
        carry += total;
    }
    if (tid == 0) meta[3 * b + 1] = carry;
}

__global__ void __launch_bounds__(256)
vb_count_kernel(int32_t *__restrict__ pt_cell, const int32_t *__restrict__ cell_first, const int32_t *__restrict__ pt_vid,
                const int32_t *__restrict__ meta, int Ntot, const BatchGeom g, size_t cells, int max_voxels,
                int32_t *__restrict__ cnt) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ntot) return;
    const int b = frame_of(g, i);
    const int li = i - g.off[b];
    const int cell = pt_cell[i];
    int v = -1;
    if (cell >= 0 && li < meta[3 * b]) {
        v = pt_vid[g.off[b] + cell_first[(size_t)b * cells + cell]];
        atomicAdd(cnt + (size_t)b * max_voxels + v, 1);
    }
    pt_cell[i] = v;   // from here on: voxel id (inside its frame) of the point, -1 = dropped
}

// bucket segments of the frame's point list: one atomic cursor per frame hands every voxel a contiguous
// segment of cnt[v] slots (the ORDER of the segments is irrelevant -- the writer sorts each voxel's points and
// addresses them through off[] -- so no scan over the 12 000 counts is needed); also the frame's voxel count
__global__ void __launch_bounds__(256)
vb_alloc_kernel(const int32_t *__restrict__ cnt, int32_t *__restrict__ meta, int batch, int max_voxels,
                int32_t *__restrict__ off, int32_t *__restrict__ frame_voxels) {
    // grid.y = frame; one atomicAdd per WARP (the warp's 32 counts are scanned with shuffles first): 12 000
    // single-thread atomics on one cursor serialise to ~35 us
    const int b = blockIdx.y;
    const int v = blockIdx.x * blockDim.x + threadIdx.x;
    const int lane = threadIdx.x & 31;
    const int vnum = min(meta[3 * b + 1], max_voxels);
    if (v == 0) frame_voxels[b] = vnum;
    const size_t i = (size_t)b * max_voxels + v;
    const int c = (v < vnum) ? cnt[i] : 0;
    int incl = c;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    const int total = __shfl_sync(0xffffffffu, incl, 31);
    int base = 0;
    if (lane == 0 && total > 0) base = atomicAdd(meta + 3 * b + 2, total);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (v < max_voxels) off[i] = base + incl - c;
}

__global__ void __launch_bounds__(256)
vb_bucket_kernel(const int32_t *__restrict__ pt_vox, const int32_t *__restrict__ off, int Ntot, const BatchGeom g,
                 int max_voxels, int32_t *__restrict__ fillc, int32_t *__restrict__ list) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Ntot) return;
    const int v = pt_vox[i];
    if (v < 0) return;
    const int b = frame_of(g, i);
    const size_t s = (size_t)b * max_voxels + v;
    list[g.off[b] + off[s] + atomicAdd(fillc + s, 1)] = i;   // GLOBAL point index
}

// one warp per MERGED output row: resolves (frame, voxel) from the per-frame voxel counts, writes the voxel's
// points in input order (zero padded), its (b, z, y, x) coordinate and point count; rows past the batch total
// are zero.  Block 0 also publishes the total.
constexpr int kVbWarps = 8;
__global__ void __launch_bounds__(kVbWarps * 32)
vb_write_kernel(const float *__restrict__ points, int F, const int32_t *__restrict__ cnt, const int32_t *__restrict__ off,
                const int32_t *__restrict__ list, const int32_t *__restrict__ coors3,
                const int32_t *__restrict__ frame_voxels, const BatchGeom g, int max_voxels, int max_points,
                float *__restrict__ voxels, int32_t *__restrict__ coors4, int32_t *__restrict__ num_points,
                int32_t *__restrict__ total_out) {
    extern __shared__ int32_t s_dyn[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    int32_t *s_sel = s_dyn + (size_t)warp * 2 * max_points;
    int32_t *s_sorted = s_sel + max_points;
    const long long row = (long long)blockIdx.x * kVbWarps + warp;
    const long long rows_total = (long long)g.B * max_voxels;
    if (row >= rows_total) return;
    int b = -1, v = 0;
    long long acc = 0;
    for (int f = 0; f < g.B; ++f) {
        const int n = frame_voxels[f];
        if (b < 0 && row < acc + n) { b = f; v = (int)(row - acc); }
        acc += n;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) *total_out = (int32_t)acc;
    int nsel = 0;
    if (b >= 0) {
        const size_t s = (size_t)b * max_voxels + v;
        const int n = cnt[s];
        const int32_t *seg = list + g.off[b] + off[s];
        int thresh = INT_MAX;
        if (n > max_points) {   // the max_points lowest point indices: binary search on the index value
            int lo = g.off[b], hi = g.off[b + 1] - 1;
            while (lo < hi) {
                const int mid = lo + ((hi - lo) >> 1);
                int c = 0;
                for (int e = lane; e < n; e += 32) c += (seg[e] <= mid);
                c = __reduce_add_sync(0xffffffffu, c);
                if (c >= max_points) hi = mid; else lo = mid + 1;
            }
            thresh = lo;
        }
        for (int e0 = 0; e0 < n; e0 += 32) {
            const int e = e0 + lane;
            const int val = (e < n) ? seg[e] : INT_MAX;
            const bool take = (e < n) && (val <= thresh);
            const unsigned m = __ballot_sync(0xffffffffu, take);
            if (take) s_sel[nsel + __popc(m & ((1u << lane) - 1u))] = val;
            nsel += __popc(m);
        }
        __syncwarp();
        for (int e = lane; e < nsel; e += 32) {   // rank by counting -> input order
            const int val = s_sel[e];
            int r = 0;
            for (int f = 0; f < nsel; ++f) r += (s_sel[f] < val);
            s_sorted[r] = val;
        }
        __syncwarp();
        if (lane == 0) {
            const int32_t *c3 = coors3 + s * 3;
            coors4[row * 4 + 0] = b;
            coors4[row * 4 + 1] = c3[0];
            coors4[row * 4 + 2] = c3[1];
            coors4[row * 4 + 3] = c3[2];
            num_points[row] = min(n, max_points);
        }
    } else {
        return;   // rows past the batch total stay zero (the launcher clears the three outputs first)
    }
    // only the real rows are written: the zero padding (98 % of the 19.2 MB per frame) comes from one memset at
    // full HBM write rate instead of from 48 000 latency-bound warps
    float *vout = voxels + (size_t)row * max_points * F;
    const bool vec = (F % 4 == 0) && ((reinterpret_cast<uintptr_t>(points) & 15u) == 0) &&
                     ((reinterpret_cast<uintptr_t>(voxels) & 15u) == 0);
    if (vec) {
        const int F4 = F / 4;
        const float4 *pin = reinterpret_cast<const float4 *>(points);
        float4 *po = reinterpret_cast<float4 *>(vout);
        for (int e = lane; e < nsel * F4; e += 32) {
            const int r = e / F4, c = e - r * F4;
            po[e] = pin[(size_t)s_sorted[r] * F4 + c];
        }
    } else {
        for (int e = lane; e < nsel * F; e += 32) {
            const int r = e / F, c = e - r * F;
            vout[e] = points[(size_t)s_sorted[r] * F + c];
        }
    }
}

// ------------------------------------------------------------------ anchors mask (box_np_ops.py:772-806)
__global__ void __launch_bounds__(256)
anchors_count_kernel(const int32_t *__restrict__ coors, int P, int stride, int c_y, int c_x, const int32_t *__restrict__ num_valid,
                     int ny, int nx, float *__restrict__ dense) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    const int Pv = num_valid ? min(*num_valid, P) : P;
    if (p >= Pv) return;
    const int y = coors[(size_t)p * stride + c_y], x = coors[(size_t)p * stride + c_x];
    if (y >= 0 && y < ny && x >= 0 && x < nx) atomicAdd(dense + (size_t)y * nx + x, 1.0f);   // small integers: exact
}
// inclusive prefix along one axis, one thread per line (the maps are a few hundred cells wide)
__global__ void __launch_bounds__(256)
cumsum_axis_kernel(float *__restrict__ m, int lines, int len, size_t line_stride, size_t elem_stride) {
    const int l = blockIdx.x * blockDim.x + threadIdx.x;
    if (l >= lines) return;
    float *p = m + (size_t)l * line_stride;
    float run = 0.f;
    for (int i = 0; i < len; ++i) {
        run += p[(size_t)i * elem_stride];
        p[(size_t)i * elem_stride] = run;
    }
}
__global__ void __launch_bounds__(256)
anchors_area_kernel(const float *__restrict__ dense, int ny, int nx, const float *__restrict__ anchors_bv, int N,
                    float sx, float sy, float ox, float oy, int gx, int gy, float *__restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const float *a = anchors_bv + (size_t)i * 4;
    int c0 = (int)floorf(__fdiv_rn(__fsub_rn(a[0], ox), sx));
    int c1 = (int)floorf(__fdiv_rn(__fsub_rn(a[1], oy), sy));
    int c2 = (int)floorf(__fdiv_rn(__fsub_rn(a[2], ox), sx));
    int c3 = (int)floorf(__fdiv_rn(__fsub_rn(a[3], oy), sy));
    c0 = max(c0, 0); c1 = max(c1, 0);
    c2 = min(c2, gx - 1); c3 = min(c3, gy - 1);
    // NumPy negative indices wrap (an anchor entirely left of / below the map)
    auto at = [&](int y, int x) -> float {
        if (y < 0) y += ny;
        if (x < 0) x += nx;
        y = min(max(y, 0), ny - 1);
        x = min(max(x, 0), nx - 1);
        return dense[(size_t)y * nx + x];
    };
    const float ID = at(c3, c2), IA = at(c1, c0), IB = at(c3, c0), IC = at(c1, c2);
    out[i] = __fadd_rn(__fsub_rn(__fsub_rn(ID, IB), IC), IA);
}

// ------------------------------------------------------------------ points_to_bev (bev_ops.py:6-103)
// Height slices keep the MAXIMUM normalised height of their cell, the last map counts points, the optional
// reflectivity map holds the intensity of the point that set the highest maximum of its (slice, y, x) cell.
// The sequential reference also numbers cells first-come and stops at the first NEW cell once max_voxels
// exist; that cut-off is found with the same opener scan as the voxeliser (single frame).
__global__ void __launch_bounds__(256)
bev_accumulate_kernel(const float *__restrict__ points, int N, int F, const BatchGeom g, const int32_t *__restrict__ pt_cell,
                      const int32_t *__restrict__ meta, const HeightLowers hl, float slice, int with_refl,
                      float *__restrict__ bev, unsigned long long *__restrict__ refl_key) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= N) return;
    const int cell = pt_cell[i];
    if (cell < 0 || i >= meta[0]) return;
    const int D = g.shape[0], H = g.shape[1], W = g.shape[2];
    int x, y, z;
    vb_cell_coords(g, cell, z, y, x);
    const size_t plane = (size_t)H * W;
    const int nmaps = D + 1 + (with_refl ? 1 : 0);
    atomicAdd(bev + (size_t)(nmaps - 1) * plane + (size_t)y * W + x, 1.0f);
    // incoming = (p_z - height_lowers[z]) / slice (:52-53), fp32 like the jitted loop
    const float h = __fdiv_rn(__fsub_rn(points[(size_t)i * F + 2], hl.v[z]), slice);
    if (h > 0.f) {   // the maps start at 0 and only a larger value replaces it (:54)
        atomic_max_f32(bev + (size_t)z * plane + (size_t)y * W + x, h);
        if (with_refl) {
            // intensity of the point with the largest height; equal heights: the EARLIER point (the later one
            // does not satisfy '>').  key = (height bits, ~index) so that atomicMax picks exactly that point
            const unsigned long long key = ((unsigned long long)__float_as_uint(h) << 32) | (unsigned)(~(unsigned)i);
            atomicMax(refl_key + (size_t)z * plane + (size_t)y * W + x, key);
        }
    }
}
__global__ void __launch_bounds__(256)
bev_refl_kernel(const float *__restrict__ points, int F, const unsigned long long *__restrict__ refl_key, int D, size_t plane,
                const float *__restrict__ bev, float *__restrict__ refl_map) {
    const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= plane) return;
    // the reference overwrites ONE reflectivity cell per (y, x) from every slice in point order (:54-57): the last
    // writer is the latest point, over all slices, that raised its slice's maximum.  Within a slice every
    // intermediate raise precedes the final one (the first point that reaches the slice's maximum, recorded in
    // refl_key), so the last writer is the largest point index among the per-slice record holders -- exact.
    unsigned best_idx = 0;
    bool any = false;
    for (int z = 0; z < D; ++z) {
        const unsigned long long key = refl_key[(size_t)z * plane + c];
        if (key == 0ull) continue;
        const unsigned idx = ~(unsigned)(key & 0xffffffffull);
        if (!any || idx > best_idx) { best_idx = idx; any = true; }
    }
    (void)bev;
    refl_map[c] = any ? points[(size_t)best_idx * F + 3] : 0.f;
}

static unsigned vb_blocks(size_t total, int threads, int per_sm) {
    size_t b = (total + threads - 1) / threads;
    const size_t cap = (size_t)kNumSMs * per_sm;
    if (b > cap) b = cap;
    return (unsigned)(b ? b : 1);
}

struct VbWs {
    size_t cell_first, pt_cell, pt_vid, cnt, off, fillc, list, coors3, meta, total;
};
static void plan_vb(int Ntot, size_t cells, int B, int max_voxels, VbWs *w) {
    size_t off = 0;
    auto take = [&](size_t bytes) { size_t o = off; off = align_up(off + bytes, 256); return o; };
    const size_t n = Ntot > 0 ? Ntot : 1, vt = (size_t)B * max_voxels;
    w->cell_first = take(cells * B * 4);
    w->pt_cell = take(n * 4);
    w->pt_vid = take(n * 4);
    w->cnt = take(vt * 4);
    w->off = take(vt * 4);
    w->fillc = take(vt * 4);
    w->list = take(n * 4);
    w->coors3 = take(vt * 3 * 4);
    w->meta = take((size_t)B * 12 + 16);
    w->total = off;
}
static bool make_batch_geom(const float *vs, const float *cr, int reverse, BatchGeom *g) {
    for (int j = 0; j < 3; ++j) {
        if (!(vs[j] > 0.f)) return false;
        g->lo[j] = cr[j];
        g->vs[j] = vs[j];
        const float q = (cr[3 + j] - cr[j]) / vs[j];
        g->grid[j] = (int)nearbyintf(q);
        if (g->grid[j] <= 0) return false;
    }
    for (int j = 0; j < 3; ++j) g->shape[j] = reverse ? g->grid[2 - j] : g->grid[j];
    g->reverse = reverse;
    auto magic = [](uint32_t d, uint32_t *mul, uint32_t *shr) {   // as tt::make_fastdiv: dividends < 2^31
        if (d <= 1u) { *mul = 0u; *shr = 0u; return; }
        uint32_t lg = 0;
        while ((1ull << lg) < d) ++lg;
        const uint32_t p = 31u + lg;
        *mul = (uint32_t)(((1ull << p) + d - 1ull) / d);
        *shr = p - 32u;
    };
    magic((uint32_t)g->shape[2], &g->mul2, &g->shr2);
    magic((uint32_t)g->shape[1], &g->mul1, &g->shr1);
    return true;
}

}  // namespace papc

using namespace papc;

extern "C" size_t papc_voxelize_batch_workspace_bytes(int total_points, int batch, const float *voxel_size_host,
                                                      const float *coors_range_host, int max_voxels) {
    BatchGeom g;
    if (total_points < 0 || batch < 1 || batch > kMaxFrames || max_voxels <= 0 || !voxel_size_host || !coors_range_host) return 0;
    if (!make_batch_geom(voxel_size_host, coors_range_host, 1, &g)) return 0;
    VbWs w;
    plan_vb(total_points, (size_t)g.grid[0] * g.grid[1] * g.grid[2], batch, max_voxels, &w);
    return w.total;
}

extern "C" int papc_voxelize_batch_f32(const float *points, const int32_t *frame_offsets_host, int batch, int F,
                                       const float *voxel_size_host, const float *coors_range_host, int max_points,
                                       int reverse_index, int max_voxels, float *voxels, int32_t *coors4,
                                       int32_t *num_points, int32_t *frame_voxels, int32_t *total_voxels,
                                       void *workspace, size_t workspace_bytes, papc_stream_t stream) {
    if (batch < 1 || batch > kMaxFrames || F < 3 || max_points <= 0 || max_voxels <= 0) return PAPC_EINVAL;
    if (!frame_offsets_host || !voxel_size_host || !coors_range_host || !voxels || !coors4 || !num_points ||
        !frame_voxels || !total_voxels)
        return PAPC_EINVAL;
    BatchGeom g;
    if (!make_batch_geom(voxel_size_host, coors_range_host, reverse_index ? 1 : 0, &g)) return PAPC_EINVAL;
    g.B = batch;
    if (frame_offsets_host[0] != 0) return PAPC_EINVAL;
    for (int b = 0; b <= batch; ++b) {
        if (b > 0 && frame_offsets_host[b] < frame_offsets_host[b - 1]) return PAPC_EINVAL;
        g.off[b] = frame_offsets_host[b];
    }
    const int Ntot = g.off[batch];
    if (Ntot > 0 && !points) return PAPC_EINVAL;
    const size_t cells = (size_t)g.grid[0] * g.grid[1] * g.grid[2];
    if (cells * batch > 0x7fffffffULL) return PAPC_EUNSUPPORTED;
    const size_t smem = (size_t)kVbWarps * 2 * max_points * sizeof(int32_t);
    if (smem > 200 * 1024) return PAPC_EUNSUPPORTED;
    VbWs w;
    plan_vb(Ntot, cells, batch, max_voxels, &w);
    if (!workspace || workspace_bytes < w.total) return PAPC_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PAPC_EINVAL;
    char *ws = reinterpret_cast<char *>(workspace);
    int32_t *cell_first = (int32_t *)(ws + w.cell_first), *pt_cell = (int32_t *)(ws + w.pt_cell);
    int32_t *pt_vid = (int32_t *)(ws + w.pt_vid), *cnt = (int32_t *)(ws + w.cnt), *off = (int32_t *)(ws + w.off);
    int32_t *fillc = (int32_t *)(ws + w.fillc), *list = (int32_t *)(ws + w.list), *coors3 = (int32_t *)(ws + w.coors3);
    int32_t *meta = (int32_t *)(ws + w.meta);
    cudaStream_t st = as_stream(stream);
    const size_t vt = (size_t)batch * max_voxels;
    ProfScope prof(st, "voxelize_batch", Ntot, F, batch, 0.0, 4.0 * Ntot * F + (double)vt * (4.0 * max_points * F + 20.0));
    const size_t init_n = cells * batch > vt ? cells * batch : vt;
    vb_init_kernel<<<vb_blocks(init_n, 256, 8), 256, 0, st>>>(cell_first, cells * batch, cnt, fillc, vt, meta, g);
    PAPC_LAUNCH_CHECK();
    if (Ntot > 0) {
        vb_cell_kernel<<<ceil_div(Ntot, 256), 256, 0, st>>>(points, Ntot, F, g, cells, pt_cell, cell_first);
        PAPC_LAUNCH_CHECK();
        vb_scan_kernel<<<batch, 1024, 0, st>>>(pt_cell, cell_first, g, cells, max_voxels, pt_vid, coors3, meta);
        PAPC_LAUNCH_CHECK();
        vb_count_kernel<<<ceil_div(Ntot, 256), 256, 0, st>>>(pt_cell, cell_first, pt_vid, meta, Ntot, g, cells, max_voxels, cnt);
        PAPC_LAUNCH_CHECK();
    }
    vb_alloc_kernel<<<dim3((unsigned)ceil_div(max_voxels, 256), (unsigned)batch), 256, 0, st>>>(cnt, meta, batch, max_voxels, off,
                                                                                                frame_voxels);
    PAPC_LAUNCH_CHECK();
    if (Ntot > 0) {
        vb_bucket_kernel<<<ceil_div(Ntot, 256), 256, 0, st>>>(pt_cell, off, Ntot, g, max_voxels, fillc, list);
        PAPC_LAUNCH_CHECK();
    }
    PAPC_CUDA_TRY(cudaMemsetAsync(voxels, 0, vt * (size_t)max_points * F * sizeof(float), st));
    PAPC_CUDA_TRY(cudaMemsetAsync(coors4, 0, vt * 4 * sizeof(int32_t), st));
    PAPC_CUDA_TRY(cudaMemsetAsync(num_points, 0, vt * sizeof(int32_t), st));
    if (smem > 48 * 1024)
        PAPC_CUDA_TRY(cudaFuncSetAttribute(vb_write_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    vb_write_kernel<<<(unsigned)ceil_div<long long>((long long)vt, kVbWarps), kVbWarps * 32, smem, st>>>(
        points, F, cnt, off, list, coors3, frame_voxels, g, max_voxels, max_points, voxels, coors4, num_points, total_voxels);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

extern "C" int papc_anchors_mask_f32(const int32_t *coors, int P, int coor_dim, const int32_t *num_valid, int ny, int nx,
                                     const float *anchors_bv, int N, const float *stride_host, const float *offset_host,
                                     const int32_t *grid_size_host, float *dense_map, float *anchors_area,
                                     papc_stream_t stream) {
    if (P < 0 || ny <= 0 || nx <= 0 || N < 0 || (coor_dim != 3 && coor_dim != 4) || !dense_map) return PAPC_EINVAL;
    if (P > 0 && !coors) return PAPC_EINVAL;
    cudaStream_t st = as_stream(stream);
    ProfScope prof(st, "anchors_mask", P, N, 0, 0.0, 4.0 * ny * nx * 3 + 16.0 * N + 4.0 * N);
    PAPC_CUDA_TRY(cudaMemsetAsync(dense_map, 0, sizeof(float) * (size_t)ny * nx, st));
    if (P > 0) {
        // sparse_sum_for_anchors_mask(coors, shape) adds at [coors[:,1], coors[:,2]] of a (z,y,x) row (:775)
        anchors_count_kernel<<<ceil_div(P, 256), 256, 0, st>>>(coors, P, coor_dim, coor_dim - 2, coor_dim - 1, num_valid, ny, nx, dense_map);
        PAPC_LAUNCH_CHECK();
    }
    if (N > 0) {
        if (!anchors_bv || !stride_host || !offset_host || !grid_size_host || !anchors_area) return PAPC_EINVAL;
        // dense_voxel_map.cumsum(0).cumsum(1) (the caller of fused_get_anchors_area, target_assigner / voxelnet)
        cumsum_axis_kernel<<<ceil_div(nx, 256), 256, 0, st>>>(dense_map, nx, ny, 1, (size_t)nx);
        PAPC_LAUNCH_CHECK();
        cumsum_axis_kernel<<<ceil_div(ny, 256), 256, 0, st>>>(dense_map, ny, nx, (size_t)nx, 1);
        PAPC_LAUNCH_CHECK();
        anchors_area_kernel<<<ceil_div(N, 256), 256, 0, st>>>(dense_map, ny, nx, anchors_bv, N, stride_host[0], stride_host[1],
                                                             offset_host[0], offset_host[1], grid_size_host[0], grid_size_host[1],
                                                             anchors_area);
        PAPC_LAUNCH_CHECK();
    }
    return PAPC_OK;
}

extern "C" size_t papc_points_to_bev_workspace_bytes(int N, const float *voxel_size_host, const float *coors_range_host) {
    BatchGeom g;
    if (N < 0 || !voxel_size_host || !coors_range_host || !make_batch_geom(voxel_size_host, coors_range_host, 1, &g)) return 0;
    const size_t cells = (size_t)g.grid[0] * g.grid[1] * g.grid[2];
    const size_t n = N > 0 ? N : 1;
    return align_up(cells * 4, 256) + 2 * align_up(n * 4, 256) + align_up(cells * 8, 256) + 256;
}

extern "C" int papc_points_to_bev_f32(const float *points, int N, int F, const float *voxel_size_host,
                                      const float *coors_range_host, const float *height_lowers_host,
                                      int with_reflectivity, int max_voxels, float *bev_map,
                                      void *workspace, size_t workspace_bytes, papc_stream_t stream) {
    if (N < 0 || F < 3 || (with_reflectivity && F < 4) || max_voxels <= 0 || !voxel_size_host || !coors_range_host ||
        !height_lowers_host || !bev_map)
        return PAPC_EINVAL;
    BatchGeom g;
    if (!make_batch_geom(voxel_size_host, coors_range_host, 1, &g)) return PAPC_EINVAL;
    g.B = 1;
    g.off[0] = 0;
    g.off[1] = N;
    const size_t cells = (size_t)g.grid[0] * g.grid[1] * g.grid[2];
    if (cells > 0x7fffffffULL) return PAPC_EUNSUPPORTED;
    if (!workspace || workspace_bytes < papc_points_to_bev_workspace_bytes(N, voxel_size_host, coors_range_host)) return PAPC_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PAPC_EINVAL;
    char *ws = reinterpret_cast<char *>(workspace);
    const size_t n = N > 0 ? N : 1;
    size_t o = 0;
    int32_t *cell_first = (int32_t *)(ws + o); o += align_up(cells * 4, 256);
    int32_t *pt_cell = (int32_t *)(ws + o); o += align_up(n * 4, 256);
    int32_t *pt_vid = (int32_t *)(ws + o); o += align_up(n * 4, 256);
    unsigned long long *refl_key = (unsigned long long *)(ws + o); o += align_up(cells * 8, 256);
    int32_t *meta = (int32_t *)(ws + o); o += 256;
    cudaStream_t st = as_stream(stream);
    const int D = g.shape[0];
    if (D > kMaxSlices) return PAPC_EUNSUPPORTED;
    HeightLowers hl;
    for (int z = 0; z < kMaxSlices; ++z) hl.v[z] = z < D ? height_lowers_host[z] : 0.f;
    const size_t plane = (size_t)g.shape[1] * g.shape[2];
    const int nmaps = D + 1 + (with_reflectivity ? 1 : 0);
    ProfScope prof(st, "points_to_bev", N, F, nmaps, 0.0, 4.0 * N * F + 4.0 * nmaps * plane);
    PAPC_CUDA_TRY(cudaMemsetAsync(bev_map, 0, sizeof(float) * nmaps * plane, st));
    if (N == 0) return PAPC_OK;
    if (with_reflectivity) PAPC_CUDA_TRY(cudaMemsetAsync(refl_key, 0, cells * 8, st));
    // cut-off of the sequential loop (the break at the first new cell once max_voxels cells exist, :44-46)
    int32_t *cnt_dummy = pt_vid;   // vb_init zeroes two per-voxel arrays: give it one harmless element
    vb_init_kernel<<<vb_blocks(cells, 256, 8), 256, 0, st>>>(cell_first, cells, cnt_dummy, cnt_dummy, 0, meta, g);
    PAPC_LAUNCH_CHECK();
    vb_cell_kernel<<<ceil_div(N, 256), 256, 0, st>>>(points, N, F, g, cells, pt_cell, cell_first);
    PAPC_LAUNCH_CHECK();
    // the opener scan gives the cut-off point; the openers' coordinates are not needed (coors3 = null)
    vb_scan_kernel<<<1, 1024, 0, st>>>(pt_cell, cell_first, g, cells, max_voxels, pt_vid, nullptr, meta);
    PAPC_LAUNCH_CHECK();
    bev_accumulate_kernel<<<ceil_div(N, 256), 256, 0, st>>>(points, N, F, g, pt_cell, meta, hl, g.vs[2], with_reflectivity,
                                                           bev_map, refl_key);
    PAPC_LAUNCH_CHECK();
    if (with_reflectivity) {
        bev_refl_kernel<<<(unsigned)ceil_div<size_t>(plane, 256), 256, 0, st>>>(points, F, refl_key, D, plane, bev_map,
                                                                               bev_map + (size_t)D * plane);
        PAPC_LAUNCH_CHECK();
    }
    return PAPC_OK;
}
