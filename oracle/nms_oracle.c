/*
 * nms_oracle.c -- CPU restatement (TEST INFRASTRUCTURE, NOT PRODUCT) of the detector post-processing
 * kernels of PAPC/models/detect/pointpillars/libs/ops/non_max_suppression/nms_gpu.py:
 *   iou_device :22-33, nms_kernel :72-103, nms_postprocess :111-130, nms_gpu :133-164,
 *   trangle_area :178-181, area :184-191, sort_vertex_in_convex_polygon :194-232,
 *   line_segment_intersection :235-277, point_in_quadrilateral :323-338, quadrilateral_intersection :341-362,
 *   rbbox_to_corners :365-388, inter :391-404, devRotateIoU(Eval) :407-412 / :556-566,
 *   rotate_nms_kernel :415-449, rotate_nms_gpu :453-488, rotate_iou_gpu(_eval) :518-553 / :603-653.
 *
 * Arithmetic: every operation in IEEE fp32, separately rounded (-ffp-contract=off), in the reference's order.
 * cos / sin of the box angle are evaluated in double and rounded to float (the reference calls libdevice's
 * single-precision cos / sin, accurate to 2 ulp, and NVVM may contract its multiplies and adds): IoU values are
 * therefore "parity unpinned" at the 1e-6 level; the golden vectors (tests/golden/nms_ref.npz, the reference's own
 * kernels run under numba's CUDA simulator) pin the logic and the keep lists.
 * Score order: descending score, ties to the HIGHER index (= numpy's stable argsort reversed, :147).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float iou_device(const float *a, const float *b) {
    const float left = fmaxf(a[0], b[0]), right = fminf(a[2], b[2]);
    const float top = fmaxf(a[1], b[1]), bottom = fminf(a[3], b[3]);
    const float width = fmaxf(right - left + 1.0f, 0.0f);
    const float height = fmaxf(bottom - top + 1.0f, 0.0f);
    const float interS = width * height;
    const float Sa = (a[2] - a[0] + 1.0f) * (a[3] - a[1] + 1.0f);
    const float Sb = (b[2] - b[0] + 1.0f) * (b[3] - b[1] + 1.0f);
    return interS / (Sa + Sb - interS);
}

static float trangle_area(const float *a, const float *b, const float *c) {
    return ((a[0] - c[0]) * (b[1] - c[1]) - (a[1] - c[1]) * (b[0] - c[0])) / 2.0f;
}

static float poly_area(const float *int_pts, int num_of_inter) {
    float area_val = 0.0f;
    for (int i = 0; i < num_of_inter - 2; ++i)
        area_val += fabsf(trangle_area(int_pts, int_pts + 2 * i + 2, int_pts + 2 * i + 4));
    return area_val;
}

static void sort_vertex_in_convex_polygon(float *int_pts, int num_of_inter) {
    if (num_of_inter <= 0) return;
    float center[2] = {0.0f, 0.0f};
    for (int i = 0; i < num_of_inter; ++i) {
        center[0] += int_pts[2 * i];
        center[1] += int_pts[2 * i + 1];
    }
    center[0] /= (float)num_of_inter;
    center[1] /= (float)num_of_inter;
    float v[2], vs[16];
    for (int i = 0; i < num_of_inter; ++i) {
        v[0] = int_pts[2 * i] - center[0];
        v[1] = int_pts[2 * i + 1] - center[1];
        const float d = sqrtf(v[0] * v[0] + v[1] * v[1]);
        v[0] = v[0] / d;
        v[1] = v[1] / d;
        if (v[1] < 0) v[0] = -2.0f - v[0];
        vs[i] = v[0];
    }
    for (int i = 1; i < num_of_inter; ++i) {
        if (vs[i - 1] > vs[i]) {
            const float temp = vs[i], tx = int_pts[2 * i], ty = int_pts[2 * i + 1];
            int j = i;
            while (j > 0 && vs[j - 1] > temp) {
                vs[j] = vs[j - 1];
                int_pts[j * 2] = int_pts[j * 2 - 2];
                int_pts[j * 2 + 1] = int_pts[j * 2 - 1];
                --j;
            }
            vs[j] = temp;
            int_pts[j * 2] = tx;
            int_pts[j * 2 + 1] = ty;
        }
    }
}

static int line_segment_intersection(const float *pts1, const float *pts2, int i, int j, float *temp_pts) {
    const float A0 = pts1[2 * i], A1 = pts1[2 * i + 1];
    const float B0 = pts1[2 * ((i + 1) % 4)], B1 = pts1[2 * ((i + 1) % 4) + 1];
    const float C0 = pts2[2 * j], C1 = pts2[2 * j + 1];
    const float D0 = pts2[2 * ((j + 1) % 4)], D1 = pts2[2 * ((j + 1) % 4) + 1];
    const float BA0 = B0 - A0, BA1 = B1 - A1, DA0 = D0 - A0, CA0 = C0 - A0, DA1 = D1 - A1, CA1 = C1 - A1;
    const int acd = DA1 * CA0 > CA1 * DA0;
    const int bcd = (D1 - B1) * (C0 - B0) > (C1 - B1) * (D0 - B0);
    if (acd != bcd) {
        const int abc = CA1 * BA0 > BA1 * CA0;
        const int abd = DA1 * BA0 > BA1 * DA0;
        if (abc != abd) {
            const float DC0 = D0 - C0, DC1 = D1 - C1;
            const float ABBA = A0 * B1 - B0 * A1;
            const float CDDC = C0 * D1 - D0 * C1;
            const float DH = BA1 * DC0 - BA0 * DC1;
            const float Dx = ABBA * DC0 - BA0 * CDDC;
            const float Dy = ABBA * DC1 - BA1 * CDDC;
            temp_pts[0] = Dx / DH;
            temp_pts[1] = Dy / DH;
            return 1;
        }
    }
    return 0;
}

static int point_in_quadrilateral(float pt_x, float pt_y, const float *corners) {
    const float ab0 = corners[2] - corners[0], ab1 = corners[3] - corners[1];
    const float ad0 = corners[6] - corners[0], ad1 = corners[7] - corners[1];
    const float ap0 = pt_x - corners[0], ap1 = pt_y - corners[1];
    const float abab = ab0 * ab0 + ab1 * ab1;
    const float abap = ab0 * ap0 + ab1 * ap1;
    const float adad = ad0 * ad0 + ad1 * ad1;
    const float adap = ad0 * ap0 + ad1 * ap1;
    return abab >= abap && abap >= 0 && adad >= adap && adap >= 0;
}

static int quadrilateral_intersection(const float *pts1, const float *pts2, float *int_pts) {
    int num_of_inter = 0;
    for (int i = 0; i < 4; ++i) {
        if (point_in_quadrilateral(pts1[2 * i], pts1[2 * i + 1], pts2)) {
            int_pts[num_of_inter * 2] = pts1[2 * i];
            int_pts[num_of_inter * 2 + 1] = pts1[2 * i + 1];
            ++num_of_inter;
        }
        if (point_in_quadrilateral(pts2[2 * i], pts2[2 * i + 1], pts1)) {
            int_pts[num_of_inter * 2] = pts2[2 * i];
            int_pts[num_of_inter * 2 + 1] = pts2[2 * i + 1];
            ++num_of_inter;
        }
    }
    float temp_pts[2];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j)
            if (line_segment_intersection(pts1, pts2, i, j, temp_pts)) {
                int_pts[num_of_inter * 2] = temp_pts[0];
                int_pts[num_of_inter * 2 + 1] = temp_pts[1];
                ++num_of_inter;
            }
    return num_of_inter;
}

static void rbbox_to_corners(float *corners, const float *rbbox) {
    const float angle = rbbox[4];
    const float a_cos = (float)cos((double)angle), a_sin = (float)sin((double)angle);
    const float center_x = rbbox[0], center_y = rbbox[1], x_d = rbbox[2], y_d = rbbox[3];
    const float cx[4] = {-x_d / 2.0f, -x_d / 2.0f, x_d / 2.0f, x_d / 2.0f};
    const float cy[4] = {-y_d / 2.0f, y_d / 2.0f, y_d / 2.0f, -y_d / 2.0f};
    for (int i = 0; i < 4; ++i) {
        corners[2 * i] = a_cos * cx[i] + a_sin * cy[i] + center_x;
        corners[2 * i + 1] = -a_sin * cx[i] + a_cos * cy[i] + center_y;
    }
}

static float inter(const float *rbbox1, const float *rbbox2) {
    float corners1[8], corners2[8], ic[16 + 8];   /* up to 8 corner hits + 16 edge crossings can be appended */
    rbbox_to_corners(corners1, rbbox1);
    rbbox_to_corners(corners2, rbbox2);
    /* the reference's 16-float buffer holds 8 points; more can only occur for degenerate (coincident) boxes */
    float buf[64];
    int n = quadrilateral_intersection(corners1, corners2, buf);
    if (n > 8) n = 8;
    memcpy(ic, buf, sizeof(float) * 2 * (size_t)n);
    sort_vertex_in_convex_polygon(ic, n);
    return poly_area(ic, n);
}

static float rotate_iou_eval(const float *rbox1, const float *rbox2, int criterion) {
    const float area1 = rbox1[2] * rbox1[3], area2 = rbox2[2] * rbox2[3];
    const float area_inter = inter(rbox1, rbox2);
    if (criterion == -1) return area_inter / (area1 + area2 - area_inter);
    if (criterion == 0) return area_inter / area1;
    if (criterion == 1) return area_inter / area2;
    return area_inter;
}

/* descending score, ties to the higher index: order[] */
typedef struct { float s; int32_t i; } key_t_;
static int cmp_key(const void *pa, const void *pb) {
    const key_t_ *a = (const key_t_ *)pa, *b = (const key_t_ *)pb;
    if (a->s > b->s) return -1;
    if (a->s < b->s) return 1;
    return (a->i > b->i) ? -1 : (a->i < b->i);
}
static void score_order(const float *dets, int n, int stride, int score_col, int32_t *order) {
    key_t_ *k = (key_t_ *)malloc(sizeof(key_t_) * (size_t)(n > 0 ? n : 1));
    for (int i = 0; i < n; ++i) { k[i].s = dets[(size_t)i * stride + score_col]; k[i].i = i; }
    qsort(k, (size_t)n, sizeof(key_t_), cmp_key);
    for (int i = 0; i < n; ++i) order[i] = k[i].i;
    free(k);
}

/* nms_gpu / rotate_nms_gpu: mask matrix (:72-103 / :415-449) then the suppress scan (:111-130).
 * rotated != 0: dets [n,6] (cx, cy, w, h, angle, score); else [n,5] (x1, y1, x2, y2, score).
 * keep [n] receives ORIGINAL indices in kept order; returns their number. */
int oracle_nms_f32(const float *dets, int n, float thresh, int rotated, int32_t *keep) {
    if (n <= 0) return 0;
    const int stride = rotated ? 6 : 5;
    int32_t *order = (int32_t *)malloc(sizeof(int32_t) * (size_t)n);
    score_order(dets, n, stride, stride - 1, order);
    const int col_blocks = n / 64 + (n % 64 > 0);
    uint64_t *mask = (uint64_t *)calloc((size_t)n * col_blocks, sizeof(uint64_t));
    for (int i = 0; i < n; ++i) {
        const float *bi = dets + (size_t)order[i] * stride;
        for (int cb = 0; cb < col_blocks; ++cb) {
            uint64_t t = 0;
            const int col_size = (n - cb * 64 < 64) ? n - cb * 64 : 64;
            int start = 0;
            if (i / 64 == cb) start = i % 64 + 1;
            for (int j = start; j < col_size; ++j) {
                const float *bj = dets + (size_t)order[cb * 64 + j] * stride;
                const float v = rotated ? rotate_iou_eval(bi, bj, -1) : iou_device(bi, bj);
                if (v > thresh) t |= (uint64_t)1 << j;
            }
            mask[(size_t)i * col_blocks + cb] = t;
        }
    }
    uint64_t *remv = (uint64_t *)calloc((size_t)col_blocks, sizeof(uint64_t));
    int num = 0;
    for (int i = 0; i < n; ++i) {
        const int nblock = i / 64, inblock = i % 64;
        if (!(remv[nblock] & ((uint64_t)1 << inblock))) {
            keep[num++] = order[i];
            for (int j = nblock; j < col_blocks; ++j) remv[j] |= mask[(size_t)i * col_blocks + j];
        }
    }
    free(remv); free(mask); free(order);
    return num;
}

/* rotate_iou_gpu_eval (:603-653): out [N,K], out[i][k] = devRotateIoUEval(query[k], boxes[i], criterion)
 * (note the argument order at :597-599: area1 is the QUERY box's). */
void oracle_rotate_iou_f32(const float *boxes, int N, const float *query, int K, int criterion, float *out) {
    for (int i = 0; i < N; ++i)
        for (int k = 0; k < K; ++k)
            out[(size_t)i * K + k] = rotate_iou_eval(query + (size_t)k * 5, boxes + (size_t)i * 5, criterion);
}
