// nms.cu -- detector post-processing on sm_100a: axis-aligned / rotated non-maximum suppression and the rotated
// IoU matrix.  Reference: PAPC/models/detect/pointpillars/libs/ops/non_max_suppression/
// nms_gpu.py (numba.cuda): nms_gpu :133-164, rotate_nms_gpu :453-488, rotate_iou_gpu(_eval) :518-653, and the C++
// twin libs/ops/cc/nms/nms_kernel.cu.cc:39-157.
//
// What the reference does per call: argsort on the host, H2D of the boxes and of a ZERO-FILLED n x n/64 mask,
// mask kernel (every 64 x 64 tile, also the lower triangle nobody reads), D2H of the mask, a sequential
// suppress loop in numba on the host.  Here the whole call is device resident and allocation free:
//   nms_rank_kernel   score order by counting (descending score, ties to the higher index == numpy's stable
//                     argsort reversed) -- deterministic, no sort library;
//   nms_mask_kernel   upper-triangle 64 x 64 tiles only; the rotated variant converts every box to corners ONCE
//                     (the reference recomputes cos / sin for every pair);
//   nms_scan_kernel   the suppress scan on the device: one CTA walks the 64-box blocks, a single thread resolves
//                     the in-block dependencies on the diagonal mask words, all threads OR the kept rows into
//                     the removal bitmap (16 independent row loads per thread); the kept ORIGINAL indices and
//                     their count stay on the device.
// Arithmetic: IEEE fp32, every operation separately rounded (this file is compiled with -fmad=false), in the
// reference's order; cos / sin in double, rounded to float -- identical to oracle/nms_oracle.c, so keep lists
// and IoU values are compared bit for bit.
#include "common.cuh"

namespace papc {

constexpr int kNmsTile = 64;

__device__ __forceinline__ float iou_device(const float *a, const float *b) {
    const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    const float width = fmaxf(right - left + 1.0f, 0.0f);
    const float height = fmaxf(bottom - top + 1.0f, 0.0f);
    const float interS = width * height;
    const float Sa = (a[2] - a[0] + 1.0f) * (a[3] - a[1] + 1.0f);
    const float Sb = (b[2] - b[0] + 1.0f) * (b[3] - b[1] + 1.0f);
    return interS / (Sa + Sb - interS);
}

// ---- rotated IoU geometry (nms_gpu.py:178-412)
__device__ __forceinline__ float trangle_area(const float *a, const float *b, const float *c) {
    return ((a[0] - c[0]) * (b[1] - c[1]) - (a[1] - c[1]) * (b[0] - c[0])) / 2.0f;
}
__device__ __forceinline__ void rbbox_to_corners(float *corners, const float *rbbox) {
    const float angle = rbbox[4];
    const float a_cos = (float)cos((double)angle), a_sin = (float)sin((double)angle);
    const float center_x = rbbox[0], center_y = rbbox[1], x_d = rbbox[2], y_d = rbbox[3];
    const float cx[4] = {-x_d / 2.0f, -x_d / 2.0f, x_d / 2.0f, x_d / 2.0f};
    const float cy[4] = {-y_d / 2.0f, y_d / 2.0f, y_d / 2.0f, -y_d / 2.0f};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        corners[2 * i] = a_cos * cx[i] + a_sin * cy[i] + center_x;
        corners[2 * i + 1] = -a_sin * cx[i] + a_cos * cy[i] + center_y;
    }
}
__device__ __forceinline__ bool point_in_quadrilateral(float pt_x, float pt_y, const float *corners) {
    const float ab0 = corners[2] - corners[0], ab1 = corners[3] - corners[1];
    const float ad0 = corners[6] - corners[0], ad1 = corners[7] - corners[1];
    const float ap0 = pt_x - corners[0], ap1 = pt_y - corners[1];
    const float abab = ab0 * ab0 + ab1 * ab1;
    const float abap = ab0 * ap0 + ab1 * ap1;
    const float adad = ad0 * ad0 + ad1 * ad1;
    const float adap = ad0 * ap0 + ad1 * ap1;
    return abab >= abap && abap >= 0 && adad >= adap && adap >= 0;
}
__device__ __forceinline__ bool line_segment_intersection(const float *pts1, const float *pts2, int i, int j, float *temp_pts) {
    const float A0 = pts1[2 * i], A1 = pts1[2 * i + 1];
    const float B0 = pts1[2 * ((i + 1) & 3)], B1 = pts1[2 * ((i + 1) & 3) + 1];
    const float C0 = pts2[2 * j], C1 = pts2[2 * j + 1];
    const float D0 = pts2[2 * ((j + 1) & 3)], D1 = pts2[2 * ((j + 1) & 3) + 1];
    const float BA0 = B0 - A0, BA1 = B1 - A1, DA0 = D0 - A0, CA0 = C0 - A0, DA1 = D1 - A1, CA1 = C1 - A1;
    const bool acd = DA1 * CA0 > CA1 * DA0;
    const bool bcd = (D1 - B1) * (C0 - B0) > (C1 - B1) * (D0 - B0);
    if (acd != bcd) {
        const bool abc = CA1 * BA0 > BA1 * CA0;
        const bool abd = DA1 * BA0 > BA1 * DA0;
        if (abc != abd) {
            const float DC0 = D0 - C0, DC1 = D1 - C1;
            const float ABBA = A0 * B1 - B0 * A1;
            const float CDDC = C0 * D1 - D0 * C1;
            const float DH = BA1 * DC0 - BA0 * DC1;
            const float Dx = ABBA * DC0 - BA0 * CDDC;
            const float Dy = ABBA * DC1 - BA1 * CDDC;
            temp_pts[0] = Dx / DH;
            temp_pts[1] = Dy / DH;
            return true;
        }
    }
    return false;
}
// intersection area of two quadrilaterals given as corners (quadrilateral_intersection :341-362,
// sort_vertex_in_convex_polygon :194-232, area :184-191)
__device__ float inter_corners(const float *pts1, const float *pts2) {
    float ip[16];   // 8 points, as the reference's local array
    int n = 0;
    auto push = [&](float x, float y) {
        if (n < 8) { ip[2 * n] = x; ip[2 * n + 1] = y; }
        ++n;
    };
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        if (point_in_quadrilateral(pts1[2 * i], pts1[2 * i + 1], pts2)) push(pts1[2 * i], pts1[2 * i + 1]);
        if (point_in_quadrilateral(pts2[2 * i], pts2[2 * i + 1], pts1)) push(pts2[2 * i], pts2[2 * i + 1]);
    }
    float tp[2];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (line_segment_intersection(pts1, pts2, i, j, tp)) push(tp[0], tp[1]);
    if (n > 8) n = 8;
    if (n <= 0) return 0.0f;
    float cx = 0.0f, cy = 0.0f;
    for (int i = 0; i < n; ++i) { cx += ip[2 * i]; cy += ip[2 * i + 1]; }
    cx /= (float)n;
    cy /= (float)n;
    float vs[8];
    for (int i = 0; i < n; ++i) {
        float v0 = ip[2 * i] - cx, v1 = ip[2 * i + 1] - cy;
        const float d = sqrtf(v0 * v0 + v1 * v1);
        v0 = v0 / d;
        v1 = v1 / d;
        if (v1 < 0) v0 = -2.0f - v0;
        vs[i] = v0;
    }
    for (int i = 1; i < n; ++i) {
        if (vs[i - 1] > vs[i]) {
            const float temp = vs[i], tx = ip[2 * i], ty = ip[2 * i + 1];
            int j = i;
            while (j > 0 && vs[j - 1] > temp) {
                vs[j] = vs[j - 1];
                ip[j * 2] = ip[j * 2 - 2];
                ip[j * 2 + 1] = ip[j * 2 - 1];
                --j;
            }
            vs[j] = temp;
            ip[j * 2] = tx;
            ip[j * 2 + 1] = ty;
        }
    }
    float area_val = 0.0f;
    for (int i = 0; i < n - 2; ++i) area_val += fabsf(trangle_area(ip, ip + 2 * i + 2, ip + 2 * i + 4));
    return area_val;
}
// devRotateIoUEval :556-566 on precomputed corners; area1 / area2 = w * h of the two boxes
__device__ __forceinline__ float rotate_iou_corners(const float *c1, float area1, const float *c2, float area2, int criterion) {
    const float area_inter = inter_corners(c1, c2);
    if (criterion == -1) return area_inter / (area1 + area2 - area_inter);
    if (criterion == 0) return area_inter / area1;
    if (criterion == 1) return area_inter / area2;
    return area_inter;
}

// ------------------------------------------------------------------ score order
__global__ void __launch_bounds__(256)
nms_rank_kernel(const float *__restrict__ dets, int n, int stride, int32_t *__restrict__ order) {
    __shared__ float s_s[1024];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float si = i < n ? dets[(size_t)i * stride + stride - 1] : 0.f;
    int rank = 0;
    for (int base = 0; base < n; base += 1024) {
        const int m = min(1024, n - base);
        __syncthreads();
        for (int j = threadIdx.x; j < m; j += blockDim.x) s_s[j] = dets[(size_t)(base + j) * stride + stride - 1];
        __syncthreads();
        if (i < n)
            for (int j = 0; j < m; ++j) {
                const float sj = s_s[j];
                rank += (sj > si || (sj == si && base + j > i)) ? 1 : 0;   // descending, ties: higher index first
            }
    }
    if (i < n) order[rank] = i;
}

// ------------------------------------------------------------------ suppression mask (upper triangle)
// 256 threads per 64 x 64 tile: thread = (row, quarter of the columns), the four 16-bit parts of a row's mask word
// meet in shared memory -- four times the warps per tile of the one-thread-per-row form (the rotated IoU is a long
// serial polygon clip per pair; a 1000-box call has only 136 tiles for 148 SMs).
constexpr int kMaskThreads = 4 * kNmsTile;
template <bool ROT>
__global__ void __launch_bounds__(kMaskThreads)
nms_mask_kernel(const float *__restrict__ dets, const int32_t *__restrict__ order, int n, float thresh,
                unsigned long long *__restrict__ mask) {
    const int col_start = blockIdx.x, row_start = blockIdx.y;
    if (col_start < row_start) return;     // never read by the scan
    const int tx = threadIdx.x & (kNmsTile - 1), part = threadIdx.x >> 6;
    const int row_size = min(n - row_start * kNmsTile, kNmsTile);
    const int col_size = min(n - col_start * kNmsTile, kNmsTile);
    const int col_blocks = ceil_div(n, kNmsTile);
    constexpr int W = ROT ? 9 : 4;   // rotated: 8 corner floats + area; plain: the 4 box coordinates
    __shared__ float s_box[kNmsTile * W];
    __shared__ unsigned long long s_word[kNmsTile];
    if (part == 0) {
        s_word[tx] = 0ull;
        if (tx < col_size) {
            const float *b = dets + (size_t)order[col_start * kNmsTile + tx] * (ROT ? 6 : 5);
            if (ROT) {
                rbbox_to_corners(s_box + tx * W, b);
                s_box[tx * W + 8] = b[2] * b[3];
            } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) s_box[tx * W + q] = b[q];
            }
        }
    }
    __syncthreads();
    if (tx < row_size) {
        const int cur = row_start * kNmsTile + tx;
        const float *b = dets + (size_t)order[cur] * (ROT ? 6 : 5);
        float mine[ROT ? 8 : 4];
        float my_area = 0.f;
        if (ROT) {
            rbbox_to_corners(mine, b);
            my_area = b[2] * b[3];
        } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) mine[q] = b[q];
        }
        unsigned long long t = 0ull;
        const int start = max((row_start == col_start) ? tx + 1 : 0, part * 16);
        const int stop = min(col_size, part * 16 + 16);
        for (int i = start; i < stop; ++i) {
            const float v = ROT ? rotate_iou_corners(mine, my_area, s_box + i * W, s_box[i * W + 8], -1)
                                : iou_device(mine, s_box + i * W);
            if (v > thresh) t |= 1ull << i;
        }
        if (t != 0ull) atomicOr(s_word + tx, t);
    }
    __syncthreads();
    if (part == 0 && tx < row_size) mask[(size_t)(row_start * kNmsTile + tx) * col_blocks + col_start] = s_word[tx];
}

// ------------------------------------------------------------------ suppress scan (nms_postprocess :111-130)
constexpr int kScanThreads = 256;
__global__ void __launch_bounds__(kScanThreads)
nms_scan_kernel(const unsigned long long *__restrict__ mask, const int32_t *__restrict__ order, int n,
                int32_t *__restrict__ keep, int32_t *__restrict__ num_out) {
    extern __shared__ unsigned long long s_remv[];   // [col_blocks]
    __shared__ unsigned long long s_diag[kNmsTile];
    __shared__ unsigned long long s_kept;
    const int col_blocks = ceil_div(n, kNmsTile);
    const int tid = threadIdx.x;
    for (int j = tid; j < col_blocks; j += kScanThreads) s_remv[j] = 0ull;
    int base = 0;
    for (int b = 0; b < col_blocks; ++b) {
        const int rows = min(kNmsTile, n - b * kNmsTile);
        __syncthreads();
        if (tid < kNmsTile) s_diag[tid] = tid < rows ? mask[(size_t)(b * kNmsTile + tid) * col_blocks + b] : 0ull;
        __syncthreads();
        if (tid == 0) {
            unsigned long long rem = s_remv[b], kept = 0ull;
            for (int t = 0; t < rows; ++t)
                if (!((rem >> t) & 1ull)) {
                    kept |= 1ull << t;
                    rem |= s_diag[t];
                }
            s_kept = kept;
        }
        __syncthreads();
        const unsigned long long kept = s_kept;
        if (tid < rows && ((kept >> tid) & 1ull))
            keep[base + __popcll(kept & ((1ull << tid) - 1ull))] = order[b * kNmsTile + tid];
        // OR the mask rows of the kept boxes into the removal bitmap.  Thread = (word j, slice of 16 rows): the
        // rows of a slice are independent predicated loads, all in flight together (one L2 round trip per block
        // of 64 boxes instead of a chain of up to 64), the four slices meet in shared memory.
        for (int j0 = b + 1; j0 < col_blocks; j0 += kNmsTile) {
            const int j = j0 + (tid & (kNmsTile - 1));
            const int sl = tid >> 6;                       // kScanThreads / kNmsTile = 4 slices
            unsigned long long acc = 0ull;
            if (j < col_blocks) {
                const unsigned long long *mrow = mask + (size_t)(b * kNmsTile + sl * 16) * col_blocks + j;
                unsigned long long v[16];
#pragma unroll
                for (int t = 0; t < 16; ++t)
                    v[t] = ((kept >> (sl * 16 + t)) & 1ull) ? mrow[(size_t)t * col_blocks] : 0ull;
#pragma unroll
                for (int t = 0; t < 16; ++t) acc |= v[t];
                if (acc != 0ull) atomicOr(s_remv + j, acc);
            }
        }
        base += __popcll(kept);
    }
    __syncthreads();
    for (int i = base + tid; i < n; i += kScanThreads) keep[i] = -1;
    if (tid == 0) *num_out = base;
}

// ------------------------------------------------------------------ rotated IoU matrix (:491-653)
__global__ void __launch_bounds__(kNmsTile)
rotate_iou_kernel(const float *__restrict__ boxes, int N, const float *__restrict__ query, int K, int criterion,
                  float *__restrict__ out) {
    const int row_start = blockIdx.x, col_start = blockIdx.y;
    const int tx = threadIdx.x;
    const int row_size = min(N - row_start * kNmsTile, kNmsTile);
    const int col_size = min(K - col_start * kNmsTile, kNmsTile);
    __shared__ float s_q[kNmsTile * 9];
    if (tx < col_size) {
        const float *q = query + (size_t)(col_start * kNmsTile + tx) * 5;
        rbbox_to_corners(s_q + tx * 9, q);
        s_q[tx * 9 + 8] = q[2] * q[3];
    }
    __syncthreads();
    if (tx < row_size) {
        const float *b = boxes + (size_t)(row_start * kNmsTile + tx) * 5;
        float mine[8];
        rbbox_to_corners(mine, b);
        const float my_area = b[2] * b[3];
        float *o = out + (size_t)(row_start * kNmsTile + tx) * K + col_start * kNmsTile;
        // devRotateIoUEval(query_box, box): rbox1 = the query box (:597-599)
        for (int i = 0; i < col_size; ++i) o[i] = rotate_iou_corners(s_q + i * 9, s_q[i * 9 + 8], mine, my_area, criterion);
    }
}

}  // namespace papc

using namespace papc;

extern "C" size_t papc_nms_workspace_bytes(int n) {
    if (n <= 0) return 0;
    const size_t cb = (size_t)ceil_div(n, kNmsTile);
    return align_up((size_t)n * sizeof(int32_t), 256) + align_up((size_t)n * cb * sizeof(unsigned long long), 256);
}

extern "C" int papc_nms_f32(const float *dets, int n, int box_dim, float thresh, int32_t *keep_out, int32_t *num_out,
                            void *workspace, size_t workspace_bytes, papc_stream_t stream) {
    if (n < 0 || (box_dim != 5 && box_dim != 6) || !num_out) return PAPC_EINVAL;
    cudaStream_t st = as_stream(stream);
    if (n == 0) {
        PAPC_CUDA_TRY(cudaMemsetAsync(num_out, 0, sizeof(int32_t), st));
        return PAPC_OK;
    }
    if (!dets || !keep_out) return PAPC_EINVAL;
    if (n > 65536) return PAPC_EUNSUPPORTED;     // the removal bitmap lives in shared memory
    if (!workspace || workspace_bytes < papc_nms_workspace_bytes(n)) return PAPC_EWORKSPACE;
    if ((reinterpret_cast<uintptr_t>(workspace) & 255u) != 0) return PAPC_EINVAL;
    char *ws = reinterpret_cast<char *>(workspace);
    int32_t *order = reinterpret_cast<int32_t *>(ws);
    unsigned long long *mask = reinterpret_cast<unsigned long long *>(ws + align_up((size_t)n * sizeof(int32_t), 256));
    const int cb = ceil_div(n, kNmsTile);
    ProfScope prof(st, box_dim == 6 ? "rotate_nms" : "nms", n, box_dim, 0, 0.0, 4.0 * n * box_dim + 4.0 * n);
    nms_rank_kernel<<<ceil_div(n, 256), 256, 0, st>>>(dets, n, box_dim, order);
    PAPC_LAUNCH_CHECK();
    dim3 grid(cb, cb);
    if (box_dim == 6) nms_mask_kernel<true><<<grid, kMaskThreads, 0, st>>>(dets, order, n, thresh, mask);
    else nms_mask_kernel<false><<<grid, kMaskThreads, 0, st>>>(dets, order, n, thresh, mask);
    PAPC_LAUNCH_CHECK();
    nms_scan_kernel<<<1, kScanThreads, (size_t)cb * sizeof(unsigned long long), st>>>(mask, order, n, keep_out, num_out);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}

extern "C" int papc_rotate_iou_f32(const float *boxes, int N, const float *query_boxes, int K, int criterion,
                                   float *out, papc_stream_t stream) {
    if (N < 0 || K < 0 || criterion < -1 || criterion > 2) return PAPC_EINVAL;
    if (N == 0 || K == 0) return PAPC_OK;
    if (!boxes || !query_boxes || !out) return PAPC_EINVAL;
    cudaStream_t st = as_stream(stream);
    dim3 grid(ceil_div(N, kNmsTile), ceil_div(K, kNmsTile));
    if (grid.y > 65535) return PAPC_EUNSUPPORTED;
    ProfScope prof(st, "rotate_iou", (long long)N * K, 5, 5, 0.0, 20.0 * (N + K) + 4.0 * N * K);
    rotate_iou_kernel<<<grid, kNmsTile, 0, st>>>(boxes, N, query_boxes, K, criterion, out);
    PAPC_LAUNCH_CHECK();
    return PAPC_OK;
}
