// sa_chain.cuh -- interface of the chained grouped-MLP kernel (sa_chain.cu): one or two consecutive
// SetAbstraction MLP layers per launch with the intermediate activation kept on the SM
// (tensor memory -> registers -> BatchNorm + ReLU + fp16 hi/lo split -> shared-memory operand of the
// next tcgen05.mma), fed by cp.async row gathers of a pre-split source image, by bulk copies of
// pre-split activation tiles, or by the folded first layer recomputed from the points.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace papc {
namespace chain {

enum { IN_POINTMLP = 0,  // relu(bn(W0 p + b0)) recomputed per row from the centred point (cin = 3 first layer)
       IN_GATHER = 1,    // rows of the pre-split source image [feats | xyz] gathered by 16-byte cp.async
       IN_TILE = 2 };    // pre-split activation tiles written by an earlier launch (store_mid)

constexpr int kTile = 128;                 // rows per tile == UMMA N

// One weight matrix in the layout the kernel's tensor-memory staging wants (prep_weights below):
// uint32 [n-tiles][chunks][2 (hi, lo)][128 channels][32 words], one chunk = 64 reduction elements, word j
// of a channel = fp16 pair (k = 2j, 2j + 1) of that channel's row, already multiplied by the column scale.
struct WeightSpec {
    const float *W;          // [rows][ld]
    int rows, ld;
    int K;                   // reduction length (multiple of 16)
    int k0, nk;              // k < nk: W[c*ld + k0 + k]
    int xyz;                 // >= 0: k in [nk, nk+3) and [nk+3, nk+6): W[c*ld + xyz + (k-nk)%3]; < 0: none
    const float *colscale;   // nullable device [K]: factor of column k
    int cs_on;               // or f16_colscale_sq(cs_gamma[k], cs_beta[k], cs_sqrt_count) (sa_mlp_tt.cuh)
    const float *cs_gamma, *cs_beta;
    float cs_sqrt_count;
    uint32_t *image;         // out
};
size_t weight_image_bytes(int rows, int K);
// One launch converts up to 4 matrices.
int prep_weights(const WeightSpec *specs, int n, cudaStream_t st);

struct ChainArgs {
    int in_mode;
    int nl;          // tensor-core layers in this launch: 1 (A) or 2 (A then B)
    int nt;          // 128-channel tiles of the LAST layer's output (1 or 2)
    int pool;        // last layer: per-group max / min of the pre-BN output
    int store_mid;   // nl == 2: the converted output of layer A is also stored to mid_out (tile images)
    int reverse;     // walk the tiles from the last to the first (the most recently written are still in L2)
    long long M;     // rows = G * K
    int K;           // rows per group (32, 64 or 128)
    int N, S;        // points per cloud, groups per cloud
    // ---- input of layer A
    const float *xyz, *new_xyz;   // IN_POINTMLP: the points; new_xyz also gives IN_GATHER's per-group constant
    const int32_t *idx;           // [M] neighbour index of every row (nullable = identity, row k of a group = point k)
    const float *l0_fold;         // IN_POINTMLP: [ka][4] folded first layer (scale*w_x, scale*w_y, scale*w_z, scale*b+shift)
    const uint8_t *image;         // IN_GATHER: f16 [2][img_rows][img_ld]: hi block, then lo block (build_image)
    int img_rows, img_ld;         // B * N, image_ld(D)
    const uint8_t *mid_in;        // IN_TILE: [tiles][ka * 512 bytes]: hi block then lo block, MN-major SWIZZLE_128B
    uint8_t *mid_out;             // store_mid: [tiles][ca * 512 bytes]
    // ---- layer A: y_A[c] = sum_k WA'[c][k] * in[k]  (+ biasA, + the group constant for IN_GATHER)
    int ka;                       // reduction length on the tensor core (multiple of 16, <= 192)
    int ca;                       // output channels (<= 128)
    const uint32_t *wimgA;        // prep_weights image of layer A's weights
    const float *wxyz;            // IN_GATHER: &WA[0][xyz column], row stride wxyz_ld (the per-group constant), or null
    int wxyz_ld;
    const float *biasA;
    const float *scaleA, *shiftA; // nl == 2: BatchNorm scale / shift of layer A (already divided by its column scale)
    // ---- layer B (nl == 2): input = relu(scaleA * y_A + shiftA)
    int cb;                       // output channels (<= 128 * nt)
    const uint32_t *wimgB;
    const float *biasB;
    // ---- outputs of the last layer
    float *pool_max, *pool_min;   // [M/K][cout] (pool)
    double *stats_partial;        // [partial_rows][2][cout]
    long long partial_rows;
    unsigned int *counter;        // last-CTA ticket (zeroed, self-cleaning)
    unsigned long long *fix_acc;  // [4][cout] + flag word (zeroed, self-cleaning): see sa_mlp_tt.cuh
    const float *gamma, *beta;
    float eps;
    double count, inv_count;
    float sqrt_count;
    float *scale, *shift, *mean_out, *var_out, *out_colscale;
    // ---- filled in by launch()
    uint32_t kmul, kshr, smul, sshr;
    int two_acc;                  // nl == 2: separate accumulators for A and B and a double-buffered operand of B
    int pdl;
    unsigned long long *clk;      // triage builds (-DPAPC_CHAIN_TRIAGE): per-tile clock64 stamps of CTA 0
    // profiler labelling: algorithmic FLOPs / bytes attributed to this launch
    double prof_flops, prof_bytes;
    int prof_cin, prof_cout;
};

bool eligible(const ChainArgs &a);
int launch(const ChainArgs &a, cudaStream_t st);

// ---- pre-split source image for IN_GATHER -------------------------------------------------------
// feats [R][D] (nullable, D = 0), xyz [R][3] -> image f16 [2][R][ld] (hi block, then lo block),
// ld = image_ld(D): columns [0,D) feats / colscale, [D,D+3) xyz / colscale (hi + mid terms),
// [D+3,D+6) xyz third term, zero padding.  colscale [D+6] (power of two per column, >= 1) is what the
// weight staging multiplies the matching weight column with.  absmax: [D+3] scratch (uint32 bit patterns).
int image_ld(int D);
size_t image_bytes(long long R, int D);
int build_image(const float *feats, const float *xyz, long long R, int D, uint8_t *image, float *colscale,
                unsigned int *absmax, cudaStream_t st);

}  // namespace chain
}  // namespace papc
