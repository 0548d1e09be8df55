// sa_mlp_tt.cuh -- interface of the transposed tcgen05 layer kernel and of the point-moment kernel
// that lets a cin<=8 first layer be recomputed on the fly.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace papc {
namespace tt {

enum { SRC_PLAIN = 0,     // x [M,cin] (+ optional relu(in_scale*x + in_shift))
       SRC_GATHER = 1,    // feats[b, idx] rows through the tensor core, centred xyz in the epilogue
       SRC_POINTMLP = 2 };// relu(bn(W0 * centred xyz + b0)) recomputed per row from the folded W0

enum { PREC_TF32 = 0,     // 3xTF32: hi/lo TF32 split of both operands, K = 8 per MMA
       PREC_F16 = 1 };    // 3xFP16: hi/lo fp16 split, K = 16 per MMA; needs |activation| < 2^15 --
                          // used only where a bound is known (inputs that are relu(batch-norm(.)))

struct TtArgs {
    int mode, prec;
    long long M;
    int K;          // rows per group (pooling)
    int cin;        // reduction length on the tensor core (PLAIN: columns of x, GATHER: D, POINTMLP: c0)
    int cout;
    // SRC_PLAIN
    const float *x, *in_scale, *in_shift;
    // SRC_GATHER / SRC_POINTMLP
    // SRC_GATHER with the fp16 split (prec == PREC_F16): x_colscale [D] are powers of two with
    // 0 <= feats[., k] / x_colscale[k] < 2^15 guaranteed by the caller (post-ReLU features of a BatchNorm layer,
    // bound |gamma| sqrt(count) + |beta|); the producers divide by it, w_colscale (= the same array) multiplies W.
    const float *x_colscale;
    const float *xyz, *new_xyz, *feats;
    const int32_t *idx;
    int N, S, D;
    const float *l0_fold;  // SRC_POINTMLP: [cin][4] = scale*w_x, scale*w_y, scale*w_z, scale*b + shift
    // weights of THIS layer: W[c*wld + wk0 + k] multiplies tensor column k; W[c*wld + wxyz + {0,1,2}]
    // the centred xyz (SRC_GATHER, -1 = none).  w_colscale[k] (nullable): power-of-two factor folded
    // into column k (the producer of the activations divided them by it).
    const float *W;
    int wld, wk0, wxyz;
    const float *w_colscale;
    // ... or, cs_on != 0, computed on the fly as f16_colscale_sq(cs_gamma[k], cs_beta[k], cs_sqrt_count)
    // from the BatchNorm parameters of the layer that produced the activations (nullable = 1 / 0):
    // nothing the weight staging reads then depends on the previous kernel.
    int cs_on;
    const float *cs_gamma, *cs_beta;
    float cs_sqrt_count;
    // programmatic dependent launch: this launch may start while the previous kernel on the stream
    // drains (its prologue -- barrier init, tensor-memory allocation, weight staging -- touches nothing
    // that kernel writes); griddepcontrol.wait precedes every dependent access
    int pdl;
    void *wimg;                  // streamed-W mode: workspace for the pre-split, pre-swizzled image
    // SRC_PLAIN, fp16 split, small M (few row tiles, several channel tiles): workspace (ximg_bytes(M, cin), nullable)
    // for a pre-split image of the ACTIVATIONS.  launch() then converts relu(bn(x)) once (prep_ximg_kernel) and
    // every CTA streams its operand chunks by TMA instead of converting the same rows once per channel tile.
    void *ximg;
    int ximg_on;                 // filled in by launch()
    const float *bias;
    float *y;                    // [M,cout] pre-BN output (nullable)
    float *pool_max, *pool_min;  // [M/K,cout] (nullable)
    double *stats_partial;       // [partial_rows][2][cout] (nullable)
    long long partial_rows;
    // fused BatchNorm finalisation by the last CTA to finish (counter nullable = off).
    // out_colscale (nullable): per-channel power of two 2^e such that relu(bn(y)) / 2^e < 2^15 for
    // every possible y (bound |gamma| sqrt(count) + |beta|); scale/shift are written divided by it.
    unsigned int *counter;
    // nullable: zeroed, self-cleaning unsigned long long [4][cout] (+1 flag word at [4*cout]): every CTA adds
    // its fp64 statistic sums as an exact (integer part, 2^-54 fraction) pair with 64-bit integer
    // atomics -- associative, so still deterministic -- and the last CTA reads 4*cout words instead
    // of reducing gridDim partial rows.  The flag is raised by a sum outside +-2^53 (or NaN): the last
    // CTA then falls back to the partial rows.
    unsigned long long *fix_acc;
    // Deferred finalisation.  PRODUCING side: counter == nullptr with fix_acc != nullptr -- the CTAs add their sums
    // into fix_acc (zeroed by the caller, NOT cleaned here) and nobody finalises: the kernel ends with its last
    // tile.  CONSUMING side (SRC_PLAIN, no activation image): in_fix != nullptr -- every CTA derives the scale /
    // shift table of its input from the producer's words (one L2 round trip that overlaps its first tile copy)
    // with the arithmetic of the last-CTA finalisation, bit for bit; CTA 0 also writes the batch mean / variance.
    // Removes ticket, reduction and scale / shift round trips (~5 us) from the layer-to-layer boundary.
    const unsigned long long *in_fix;  // [4][cin] + flag word at [4*cin]
    const double *in_partial;          // the producer's partial rows (used when the flag is raised)
    long long in_partial_rows;
    const float *in_gamma, *in_beta;   // BatchNorm parameters of the producing layer (nullable = 1 / 0)
    float in_eps;
    double in_inv_count;
    int in_cs;                         // scale / shift divided by f16_colscale_sq(in_gamma, in_beta, cs_sqrt_count)
    float *in_mean_out, *in_var_out;   // nullable
    const float *gamma, *beta;
    float eps;
    double count;
    double inv_count;  // 1 / count (filled in by launch())
    float sqrt_count;
    float *scale, *shift, *mean_out, *var_out, *out_colscale;
    // row -> (group, position, cloud) without 64-bit divisions: magic-number division by K and S,
    // filled in by launch(); valid while M < 2^31 (fastgeom = 0 falls back to long long division)
    uint32_t kmul, kshr, smul, sshr;
    int fastgeom;
    int dbg;  // PAPC_TT_DBG bit mask (performance triage only): 1 = producers skip loads+math,
              // 2 = no MMAs issued, 4 = epilogue skips its math / stores
    // SRC_PLAIN: 2-D TMA descriptor of x [M, cin] (box = 128 rows x one chunk), built by launch();
    // tma2d != 0: one cp.async.bulk.tensor per chunk replaces the per-row copies
    alignas(64) CUtensorMap xmap;
    int tma2d;
    unsigned long long *gclk; // triage builds (PAPC_TT_GCLK=1): [grid][8] %globaltimer stamps of this launch
    unsigned long long *clk;  // triage builds: [3][16] per-phase clock64() stamps (CTA 0, CTA 1, last CTA)
};

struct TtProblem {
    int mode, prec, cin, cout, K, D;
    bool pool;
    bool no_act = false;  // SRC_PLAIN without an input activation: no scale / shift table, any cin
};
bool eligible(const TtProblem &p);
// Division by a runtime constant d >= 1 for dividends < 2^31: q = d == 1 ? x : umulhi(x, mul) >> shr
void make_fastdiv(uint32_t d, uint32_t *mul, uint32_t *shr);
// Bytes of the streamed-W image the launch needs in TtArgs::wimg (0 = W fits in tensor memory).
size_t wimg_bytes(int prec, int cin, int cout);
// Bytes of the activation image for an [M, cin] fp16-split layer (0 = the layer does not use one).
size_t ximg_bytes(long long M, int cin, int cout);
int launch(const TtArgs &a, cudaStream_t st);

// 2^e (e >= 0) with relu(bn(y)) / 2^e < 2^15 for every possible y: |bn(y)| <= |gamma| sqrt(count) +
// |beta| (a sample is at most sqrt(count - 1) standard deviations from the batch mean).  Single
// precision on purpose (fp64 is slow on this part and the staging code evaluates it per weight
// column) with a 2 % margin for the rounding; every user -- the finalisation that divides scale /
// shift, the weight staging and the streamed-W image -- calls this one function with the same
// arguments, so producer and consumer always agree on the power of two.
__host__ __device__ inline float f16_colscale_sq(float gamma, float beta, float sqrt_count) {
    const float bound = (gamma < 0.f ? -gamma : gamma) * sqrt_count + (beta < 0.f ? -beta : beta);
    float s = 1.f;
    while (s * 32000.f < bound && s < 1e30f) s *= 2.f;
    return s;
}

// Consumer side of a deferred finalisation (TtArgs::in_fix) for kernels other than the layer kernel itself
// (pool_finish, the activation-image kernel): what they need to derive scale / shift of C channels.
struct DeferredIn {
    const unsigned long long *fix;   // [4][C] + flag word at [4*C]; nullptr = not deferred
    const double *partial;           // the producer's partial rows (flag raised)
    long long partial_rows;
    const float *gamma, *beta;       // nullable = 1 / 0
    float eps;
    double inv_count;
    int cs_on;                       // divide by f16_colscale_sq(gamma, beta, sqrt_count)
    float sqrt_count;
    float *mean_out, *var_out;       // nullable; written by block 0
};

#ifdef __CUDACC__
// BatchNorm scale / shift of one channel from its batch sums -- the one place this arithmetic lives (last-CTA
// finalisation and every consumer-side derivation must agree bit for bit).  fp64 division and square root are long
// software sequences on a slow pipe: reciprocal of the count from the host, 1/sqrt by two Newton steps from the
// fp32 estimate (full double accuracy).  cs: power-of-two column scale the results are divided by (1 = none).
__device__ __forceinline__ void bn_from_sums(double sum, double sumsq, double inv_count, float gamma, float beta,
                                             float eps, float cs, float &scale, float &shift, float &mean_f,
                                             float &var_f) {
    const double mean = sum * inv_count;
    double var = sumsq * inv_count - mean * mean;  // biased, as Paddle's training BN
    var = var > 0.0 ? var : 0.0;
    const double g = (double)gamma, b = (double)beta;
    const double ve = var + (double)eps;
    double rs = (double)rsqrtf((float)ve);
    rs = rs * (1.5 - 0.5 * ve * rs * rs);
    rs = rs * (1.5 - 0.5 * ve * rs * rs);
    const double sc = g * rs;
    scale = (float)sc / cs;
    shift = (float)(b - mean * sc) / cs;
    mean_f = (float)mean;
    var_f = (float)var;
}
// The scale / shift table of C <= NPT * nthreads channels into shared memory, one block: every global load of a
// thread's channels (statistic words, flag, gamma, beta) is issued before the first is used -- ONE L2 round trip
// for the whole table instead of one per channel and stage (this sits on the step's critical path: the kernel's
// input only exists once the previous kernel has ended).  scale_out / shift_out / mean / var: written when
// write_stats (block 0).
template <int NPT>
__device__ __forceinline__ void deferred_table(const DeferredIn &d, int C, int tid, int nthreads, bool write_stats,
                                               float *s_sc, float *s_sh, float *scale_out, float *shift_out) {
    unsigned long long w[NPT][4];
    float g[NPT], b[NPT];
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
        const int k = tid + j * nthreads;
        g[j] = 1.f; b[j] = 0.f;
#pragma unroll
        for (int q = 0; q < 4; ++q) w[j][q] = k < C ? __ldcg(d.fix + (size_t)q * C + k) : 0ull;
        if (k < C) {
            if (d.gamma) g[j] = __ldg(d.gamma + k);
            if (d.beta) b[j] = __ldg(d.beta + k);
        }
    }
    const bool fixed = __ldcg(d.fix + (size_t)4 * C) == 0ull;
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
        const int k = tid + j * nthreads;
        if (k >= C) continue;
        double sum, sq;
        if (fixed) {
            sum = (double)(long long)w[j][0] + (double)(long long)w[j][1] * 0x1p-54;
            sq = (double)(long long)w[j][2] + (double)(long long)w[j][3] * 0x1p-54;
        } else {   // a sum left the fixed-point range (or is not finite): the partial rows, in order
            sum = 0.0; sq = 0.0;
            for (long long r = 0; r < d.partial_rows; ++r) {
                sum += __ldcg(d.partial + (r * 2 + 0) * C + k);
                sq += __ldcg(d.partial + (r * 2 + 1) * C + k);
            }
        }
        const float cs = d.cs_on ? f16_colscale_sq(g[j], b[j], d.sqrt_count) : 1.f;
        float sc, sh, mean, var;
        bn_from_sums(sum, sq, d.inv_count, g[j], b[j], d.eps, cs, sc, sh, mean, var);
        s_sc[k] = sc;
        s_sh[k] = sh;
        if (write_stats) {
            if (scale_out) scale_out[k] = sc;
            if (shift_out) shift_out[k] = sh;
            if (d.mean_out) d.mean_out[k] = mean;
            if (d.var_out) d.var_out[k] = var;
        }
    }
}
#endif

// Moments of the centred grouped points p = xyz[b, idx] - new_xyz over all M rows, then -- in the
// last block -- BatchNorm statistics of y0 = W0 p + b0 derived analytically in fp64
// (mean = W0 mu + b0, var = W0 Cov W0^T), the resulting scale / shift and the folded first layer.
struct MomentArgs {
    const float *xyz, *new_xyz;
    const int32_t *idx;
    int N, S, K;
    long long M;
    const float *W0, *b0;  // [c0,3], [c0] (nullable)
    const float *gamma, *beta;
    const float *running_mean, *running_var;  // used instead of the batch moments when non-null
    float eps;
    int c0;
    float sqrt_M;
    uint32_t kmul, kshr, smul, sshr;  // as TtArgs (filled in by launch_moments)
    int fastgeom;
    int nslice;             // > 0: staged variant, nslice blocks per cloud (filled in by launch_moments)
    const double *pre_partial;  // non-null: [pre_rows][9] sums already accumulated (papc_sample_group_f32): one block
    int pre_rows;               //           reduces them in fixed order and finalises -- no pass over the rows
    double *partial;        // [blocks][9]
    unsigned int *counter;
    float *scale, *shift, *mean_out, *var_out;  // [c0] (mean/var nullable)
    float *l0_fold;         // [c0][4]
    float *out_colscale;    // [c0] (nullable): see TtArgs::out_colscale; folded into l0_fold
};
int moment_blocks(long long M);
int launch_moments(const MomentArgs &a, cudaStream_t st);

}  // namespace tt
}  // namespace papc
