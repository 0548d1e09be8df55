/*
 * papc_b200.h -- C ABI of the B200-native point-cloud primitive library.
 *
 * This is the drop-in boundary for ONE hot path of AgentMaker/PAPC: the PointNet++
 * SetAbstraction forward and the PointPillars pillar encode.  The reference has no FFI
 * layer -- its "ops" are plain Python callables -- so every entry point below cites the
 * reference callable (file:line under /root/reference) whose computation it replaces.
 * Short names used in the citations:
 *   layers.py  = PAPC/models/layers/pointnet2_basic_layers.py
 *   pc_ops.py  = PAPC/models/detect/pointpillars/libs/ops/point_cloud/point_cloud_ops.py
 *   pillars.py = PAPC/models/detect/pointpillars/models/bones/pillars.py
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no C++/torch types.
 *   - Every pointer is a DEVICE pointer on the current CUDA device unless the parameter
 *     name ends in _host.  The caller allocates every input, output and workspace buffer;
 *     the library never allocates, frees or retains memory.  Inputs are read-only.
 *   - Tensors are dense row-major in the stated shape.
 *   - `stream` is a cudaStream_t passed as void*.  Calls only enqueue work (asynchronous).
 *   - Return value: PAPC_OK (0) or a negative papc_status.  Invalid arguments are rejected
 *     before anything is launched.  papc_status_string() names a code.
 *   - Re-entrant: no global mutable state; concurrent calls on distinct streams/workspaces
 *     are safe.  A few environment variables are read as kernel-selection A/B switches
 *     (PAPC_MLP_TC, PAPC_CHAIN, PAPC_TT_PDL, PAPC_TT_TMA2D, PAPC_TT_XIMG, PAPC_TT_PAIR, PAPC_TT_DEFER,
 *     PAPC_FPS_WIDE): every setting gives
 *     the same results within the stated tolerances.  Switches that change results exist in
 *     triage builds only (-DPAPC_TRIAGE / -DPAPC_TT_TRIAGE).
 */
#ifndef PAPC_B200_H_
#define PAPC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PAPC_ABI_VERSION 1

typedef void *papc_stream_t; /* cudaStream_t */

typedef enum papc_status {
    PAPC_OK = 0,
    PAPC_EINVAL = -1,       /* bad shape / null pointer / unsupported flag value          */
    PAPC_EWORKSPACE = -2,   /* workspace too small (see the matching *_workspace_bytes)     */
    PAPC_ECUDA = -3,        /* a CUDA runtime call failed; papc_last_cuda_error() has it    */
    PAPC_EUNSUPPORTED = -4  /* valid request outside what this build implements             */
} papc_status;

const char *papc_status_string(int status);
int papc_abi_version(void);
/* cudaError_t of the most recent PAPC_ECUDA on this host thread (0 if none). */
int papc_last_cuda_error(void);
/* Number of CUDA kernels this library has launched from the calling host thread (monotonic;
 * diagnostics for bench.py's gpu_launches -- memsets are not counted). */
uint64_t papc_launch_count(void);

/* Launch profiler (diagnostics; this is how bench.py times "the dominant kernel" live).  While
 * enabled on the calling host thread, every kernel launch of papc_fps_f32, papc_ball_query_f32 and
 * the grouped-MLP entry points is bracketed by two CUDA events recorded on the launching stream.
 * papc_prof_get(i) synchronises record i's end event and returns the kernel's name, its problem
 * size, the algorithmic FLOPs / bytes of that launch and the elapsed milliseconds.  Records
 * accumulate (at most 8192) until papc_prof_reset().  Not for use inside a CUDA-graph capture. */
int papc_prof_enable(int on);
int papc_prof_reset(void);
int papc_prof_count(void);
int papc_prof_get(int i, char *name, int name_cap, int64_t *M, int32_t *cin, int32_t *cout,
                  double *flops, double *bytes, float *ms);

/* ----------------------------------------------------------------------------------------
 * A1  square_distance(src, dst)                                    layers.py:26-40
 *     src [B,N,3], dst [B,M,3] -> out [B,N,M];  out = (-2*dot + |src|^2) + |dst|^2 with
 *     dot = fma(z,z', fma(y,y', x*x')) and |.|^2 = (x*x + y*y) + z*z  (the pinned arithmetic
 *     of oracle/papc_oracle.c).  C must be 3.
 */
int papc_square_distance_f32(const float *src, const float *dst, int B, int N, int M,
                             float *out, papc_stream_t stream);

/* A2  index_points(points, idx)                                    layers.py:43-62
 *     points [B,N,C], idx [B,M] int64 (flatten trailing index dims into M) -> out [B,M,C].
 *     Indices outside [0,N) are a caller error (the reference raises IndexError); they are
 *     clamped so the kernel never reads out of bounds.
 */
int papc_gather_f32(const float *points, const int64_t *idx, int B, int N, int C, int M,
                    float *out, papc_stream_t stream);

/* A3  farthest_point_sample(xyz, npoint)                           layers.py:65-95
 *     xyz [B,N,3] -> out_idx [B,npoint] int64.  start_idx [B] replaces the reference's
 *     paddle.randint draw (:76); init_dist is the initial running distance (the reference
 *     uses 1.0, :75 -- NOT 1e10; must be >= 0).  out_new_xyz (nullable) [B,npoint,3] receives
 *     xyz[b, out_idx[b,i]] (the index_points call that always follows, :144 / :258).
 *     workspace: papc_fps_workspace_bytes(B,N) bytes (0 for N <= 8192), may be NULL then.
 */
size_t papc_fps_workspace_bytes(int B, int N);
int papc_fps_f32(const float *xyz, int B, int N, int npoint, const int64_t *start_idx,
                 float init_dist, int64_t *out_idx, float *out_new_xyz, void *workspace,
                 size_t workspace_bytes, papc_stream_t stream);

/* A4  query_ball_point(radius, nsample, xyz, new_xyz)              layers.py:98-126
 *     xyz [B,N,3], new_xyz [B,S,3] -> out_idx [B,S,nsample]: the nsample lowest indices j with
 *     NOT(square_distance(new_xyz, xyz)[j] > radius2), ascending, padded with the first one.
 *     radius2 must be float32(double(radius)**2) (:112).  An empty ball yields N in every slot
 *     and increments *empty_count (nullable int32 device counter, caller zeroes it); the
 *     reference fails with IndexError in that case.  idx_bits selects int64 (64, the
 *     reference dtype) or int32 (32, used by the fused SetAbstraction path) output.
 */
int papc_ball_query_f32(const float *xyz, const float *new_xyz, int B, int N, int S,
                        float radius2, int nsample, void *out_idx, int idx_bits,
                        int32_t *empty_count, papc_stream_t stream);

/* A3 + A4 + the grouping statistics in ONE launch      layers.py:143-146 (sample_and_group's first three calls)
 *     farthest_point_sample -> index_points -> query_ball_point for one radius, with the ball query of
 *     centroid i running (on other SMs) while the FPS recurrence is still producing centroid i+1...: results
 *     are bit-identical to papc_fps_f32 followed by papc_ball_query_f32 (idx_bits 32).  out_fps_idx [B,npoint]
 *     int64, out_new_xyz [B,npoint,3], out_group_idx [B,npoint,nsample] int32; moments_partial (nullable)
 *     [papc_sample_group_parts(...) * B][9] doubles for papc_group_source.xyz_moments.
 *     papc_sample_group_parts returns the number of consumer CTAs per cloud, or 0 when the shape does not run
 *     fused (256 < N <= 1024 and B + B*parts <= 148 SMs are required: every CTA must be co-resident) -- then
 *     papc_sample_group_f32 returns PAPC_EUNSUPPORTED and the caller uses the two separate entry points.
 *     workspace: papc_sample_group_workspace_bytes(B, npoint).
 */
int papc_sample_group_parts(int B, int N, int npoint, int nsample);
size_t papc_sample_group_workspace_bytes(int B, int npoint);
int papc_sample_group_f32(const float *xyz, int B, int N, int npoint, const int64_t *start_idx,
                          float init_dist, float radius2, int nsample, int64_t *out_fps_idx,
                          float *out_new_xyz, int32_t *out_group_idx, int32_t *empty_count,
                          double *moments_partial, void *workspace, size_t workspace_bytes,
                          papc_stream_t stream);

/* A4 (several radii)  the per-radius query_ball_point calls of PointNetSetAbstractionMsg   layers.py:258-267
 *     One pass over the cloud per centroid evaluates every distance once and fills R (<= 4) index lists:
 *     out_idx_host[r] -> [B,S,nsample_host[r]] with radius2_host[r]; results identical to R calls of
 *     papc_ball_query_f32.  radius2_host / nsample_host / out_idx_host are HOST arrays of length R (the
 *     entries of out_idx_host are device pointers); empty_count (nullable) is a device int32 [R].
 */
int papc_ball_query_multi_f32(const float *xyz, const float *new_xyz, int B, int N, int S, int R,
                              const float *radius2_host, const int32_t *nsample_host,
                              void *const *out_idx_host, int idx_bits, int32_t *empty_count,
                              papc_stream_t stream);

/* kNN  the k nearest points of every query under square_distance (A1's arithmetic): ascending distance,
 *     ties to the lower index -- a stable argsort of a square_distance row cut at k (what
 *     PointNetFeaturePropagation takes its three neighbours from, layers.py:316-318).  1 <= k <= 32.
 *     out_idx [B,S,k] int64 / int32 (idx_bits); out_dist [B,S,k] nullable.  Fewer than k points: index N,
 *     distance +inf in the unused slots.
 */
int papc_knn_f32(const float *xyz, const float *query, int B, int N, int S, int k, void *out_idx,
                 int idx_bits, float *out_dist, papc_stream_t stream);

/* A5  the gather+centre+concat of sample_and_group                 layers.py:146-151, 263-267
 *     out [B,S,K,3+D]: PAPC_XYZ_FIRST  -> [xyz[idx]-new_xyz, feats[idx]]  (SSG, :151)
 *                      PAPC_FEATS_FIRST-> [feats[idx], xyz[idx]-new_xyz]  (MSG, :267)
 *     feats may be NULL (D=0).  idx is int64 [B,S,K].
 */
enum { PAPC_XYZ_FIRST = 0, PAPC_FEATS_FIRST = 1 };
int papc_group_gather_f32(const float *xyz, const float *new_xyz, const float *feats,
                          const int64_t *idx, int B, int N, int S, int K, int D, int order,
                          float *out, papc_stream_t stream);

/* ----------------------------------------------------------------------------------------
 * A7/A8  the grouped shared MLP: (Conv2D 1x1 + bias -> BatchNorm2D -> ReLU) x L -> max over
 *        the K neighbours.                                          layers.py:214-219, 271-276
 *
 * Rows are the M = G*K grouped neighbours (G = B*S groups of K consecutive rows).  The input
 * is either an explicit grouped tensor or -- the fused path, which never materialises
 * [B,S,K,3+D] -- the (xyz, new_xyz, feats, idx) tuple it would be gathered from.
 */
#define PAPC_MAX_MLP_LAYERS 8
enum { PAPC_BN_BATCH = 0,    /* training-mode statistics over all M rows (biased variance);   */
                             /* what the reference's unregistered SA layers always do (A7)   */
       PAPC_BN_RUNNING = 1,  /* normalise with running_mean / running_var                    */
       PAPC_BN_NONE = 2 };   /* no normalisation (PFNLayer use_norm=False only, pillars.py:26-27) */
enum { PAPC_OUT_BSC = 0,     /* out [B,S,Cout]  (channels-last)                               */
       PAPC_OUT_BCS = 1 };   /* out [B,Cout,S]  (the reference's layout, layers.py:219)       */

typedef struct papc_group_source {
    const float *grouped; /* [G,K,cin] explicit rows, or NULL to gather:                     */
    const float *xyz;     /* [B,N,3]                                                         */
    const float *new_xyz; /* [B,S,3] subtracted from the xyz channels; NULL = no centring    */
    const float *feats;   /* [B,N,D] or NULL (D = 0)                                         */
    const int32_t *idx;   /* [B,S,K] int32, or NULL = identity (group_all: row k = point k)  */
    int32_t B, N, S, K, D;
    int32_t order;        /* PAPC_XYZ_FIRST / PAPC_FEATS_FIRST                               */
    const double *xyz_moments; /* nullable: [xyz_moment_rows][9] partial sums (x, y, z, xx, xy, xz, yy, yz, zz)  */
    int32_t xyz_moment_rows;   /* of the centred grouped points over all M rows, as papc_sample_group_f32      */
    int32_t reserved;          /* writes them: the folded first layer (D = 0) then skips its own gather pass    */
    const float *feats_colscale; /* nullable: [D] powers of two c_k with 0 <= feats[.,k] / c_k < 2^15 for every      */
                                 /* row -- true for post-ReLU outputs of a BatchNorm layer with c_k >= (|gamma_k|  */
                                 /* sqrt(count) + |beta_k|) / 32000.  Lets the first MLP layer run the fp16 operand  */
                                 /* split (half the tensor-core products of the 3xTF32 form); NULL = 3xTF32         */
} papc_group_source;

typedef struct papc_mlp_layer {
    const float *weight;       /* [cout,cin]  == Conv2D.weight [cout,cin,1,1]                */
    const float *bias;         /* [cout] or NULL                                             */
    const float *gamma;        /* [cout] BatchNorm weight                                    */
    const float *beta;         /* [cout] BatchNorm bias                                      */
    const float *running_mean; /* [cout], read when bn_mode == PAPC_BN_RUNNING               */
    const float *running_var;  /* [cout]                                                     */
    float *batch_mean;         /* [cout] out, nullable: batch mean (PAPC_BN_BATCH)           */
    float *batch_var;          /* [cout] out, nullable: biased batch variance                */
    int32_t cout;
    int32_t reserved;
} papc_mlp_layer;

typedef struct papc_mlp {
    int32_t num_layers; /* 1..PAPC_MAX_MLP_LAYERS */
    int32_t cin;        /* 3+D for a gathered source */
    int32_t bn_mode;
    float eps;          /* 1e-5 for BatchNorm2D (layers.py:190) */
    papc_mlp_layer layers[PAPC_MAX_MLP_LAYERS];
} papc_mlp;

size_t papc_sa_mlp_workspace_bytes(const papc_group_source *src, const papc_mlp *mlp);
/* out: [B,S,cout_last] or [B,cout_last,S] per out_layout. */
int papc_sa_mlp_f32(const papc_group_source *src, const papc_mlp *mlp, float *out,
                    int out_layout, void *workspace, size_t workspace_bytes,
                    papc_stream_t stream);

/* Step-wise form of the same computation, for batch-sharded multi-GPU runs where the
 * BatchNorm statistics of each layer are summed across ranks between the steps
 * (SURVEY.md 8e).  papc_sa_mlp_f32 == for each layer { layer_forward; stats_reduce;
 * bn_scale_shift } then pool_finish.
 *
 *   papc_mlp_layer_forward_f32: y = W * act(x) + bias for one layer.
 *     layer 0: src != NULL (x, in_scale, in_shift ignored).
 *     later  : x [M,cin] = previous layer's pre-BN output, act(v) = relu(in_scale*v+in_shift).
 *     y [M,cout] (nullable when pooling) receives the pre-BN output; pool_max / pool_min
 *     [M/K,cout] (nullable) receive the per-group extrema of y (valid because BN+ReLU is
 *     monotone per channel, so max_k relu(bn(y_k)) = relu(bn(max_k y_k or min_k y_k))).
 *     stats_partial: double [papc_mlp_stats_partial_rows(M), 2, cout] per-CTA sums of y, y^2.
 *     workspace: papc_mlp_layer_workspace_bytes(cin, cout) bytes (the hi/lo-split, swizzled weight
 *     image of the tcgen05 kernel); without it the layer runs on the fp32 SIMT kernel.
 *   papc_mlp_stats_reduce_f64: fixed-order reduction -> sums double [2,cout].
 *   papc_bn_scale_shift_f32: sums (already summed over ranks) + count -> scale, shift (and
 *     optionally mean / biased var).
 *   papc_sa_pool_finish_f32: out = relu(scale * (scale>=0 ? pool_max : pool_min) + shift).
 */
int64_t papc_mlp_stats_partial_rows(int64_t M);
size_t papc_mlp_layer_workspace_bytes(int32_t cin, int32_t cout);
int papc_mlp_layer_forward_f32(const papc_group_source *src, const float *x,
                               const float *in_scale, const float *in_shift, int64_t M,
                               int32_t cin, int32_t cout, int32_t K, const float *weight,
                               const float *bias, float *y, float *pool_max, float *pool_min,
                               double *stats_partial, void *workspace, size_t workspace_bytes,
                               papc_stream_t stream);
int papc_mlp_stats_reduce_f64(const double *stats_partial, int64_t partial_rows, int32_t cout,
                              double *sums, papc_stream_t stream);
int papc_bn_scale_shift_f32(const double *sums, double count, const float *gamma,
                            const float *beta, float eps, int32_t cout, float *scale,
                            float *shift, float *mean_out, float *var_out,
                            papc_stream_t stream);
int papc_bn_running_scale_shift_f32(const float *running_mean, const float *running_var,
                                    const float *gamma, const float *beta, float eps,
                                    int32_t cout, float *scale, float *shift,
                                    papc_stream_t stream);
int papc_sa_pool_finish_f32(const float *pool_max, const float *pool_min, const float *scale,
                            const float *shift, int32_t B, int32_t S, int32_t cout,
                            float *out, int out_layout, papc_stream_t stream);

/* ----------------------------------------------------------------------------------------
 * N1  PointNetFeaturePropagation (the first "next" row)          layers.py:284-335
 *
 *   papc_fp_interpolate_f32: lines :306-329.  xyz1 [B,N,3], xyz2 [B,S,3], points1 [B,N,D1] (nullable,
 *     D1 = 0), points2 [B,S,D2], all channels-last -> out [B*N, ld_out] rows
 *     [points1 | interpolated | zero padding], ld_out >= D1 + D2.  S == 1 tiles points2 (:314).
 *     Otherwise the three smallest square_distance values of each point weight -- through
 *     1/(d + 1e-8), normalised -- the features of sampled points 0, 1, 2: the reference takes the
 *     argsort of the already SORTED distances (:317-318), i.e. the identity, and that quirk is
 *     reproduced.
 *   papc_pointwise_mlp_f32: lines :332-335, (Conv1D k=1 + BatchNorm1D + ReLU) x L over the M rows
 *     of x [M, ld_x] (ld_x >= mlp->cin, padding columns must be zero) -> out [M, cout_last].
 *     bn_mode / eps / per-layer pointers as in papc_mlp.
 */
/*   papc_bn_relu_apply_f32: out = relu(scale * y + shift) over rows [M,cout] -- the last step of the step-wise
 *     (batch-sharded, SyncBN) form of papc_pointwise_mlp_f32: per layer papc_mlp_layer_forward_f32 (src = NULL,
 *     K = 1, no pooling) -> papc_mlp_stats_reduce_f64 -> all-reduce -> papc_bn_scale_shift_f32, then this.
 */
int papc_bn_relu_apply_f32(const float *y, const float *scale, const float *shift, int64_t M,
                           int32_t cout, float *out, papc_stream_t stream);
int papc_fp_interpolate_f32(const float *xyz1, const float *xyz2, const float *points1,
                            const float *points2, int B, int N, int S, int D1, int D2, int ld_out,
                            float *out, papc_stream_t stream);
size_t papc_pointwise_mlp_workspace_bytes(int64_t M, int32_t ld_x, const papc_mlp *mlp);
int papc_pointwise_mlp_f32(const float *x, int64_t M, int32_t ld_x, const papc_mlp *mlp, float *out,
                           void *workspace, size_t workspace_bytes, papc_stream_t stream);

/* ----------------------------------------------------------------------------------------
 * A9  points_to_voxel(points, voxel_size, coors_range, max_points, reverse_index, max_voxels)
 *                                                                   pc_ops.py:106-166
 *     points [N,F] (F >= 3) -> voxels [max_voxels,max_points,F] (zero padded; every element is
 *     written), coors [max_voxels,3] int32 ((z,y,x) when reverse_index), num_points
 *     [max_voxels] int32, *voxel_num int32 (device).  Rows >= *voxel_num are zero; the caller
 *     slices [:voxel_num] as pc_ops.py:161-163 does.  Reproduces the sequential semantics
 *     deterministically: first-come voxel numbering, per-voxel points in input order capped at
 *     max_points, and the `break` at the first NEW cell once max_voxels exist (which drops
 *     every later point).  voxel_size_host [3], coors_range_host [6] are HOST arrays.
 */
size_t papc_voxelize_workspace_bytes(int N, const float *voxel_size_host,
                                     const float *coors_range_host, int max_voxels);
int papc_voxelize_f32(const float *points, int N, int F, const float *voxel_size_host,
                      const float *coors_range_host, int max_points, int reverse_index,
                      int max_voxels, float *voxels, int32_t *coors, int32_t *num_points,
                      int32_t *voxel_num, void *workspace, size_t workspace_bytes,
                      papc_stream_t stream);

/* A10 PillarFeatureNet.forward with a single (last) PFNLayer          pillars.py:79-108, 29-41
 *     features [P,T,F] (F >= 3), num_voxels [P] int32, coors [P,4] int32 (b,z,y,x) -> out [P,cout].
 *     Decorations [features, f_cluster(3), f_center(2)] (:82-95), padding rows zeroed (:99-102),
 *     Linear(F+5 -> cout, weight [F+5,cout], bias nullable) -> BatchNorm1D over all P*T rows
 *     (padding rows included, as the reference does) -> ReLU -> max over T.
 *     bn_mode as above; eps 1e-3 in the reference (:24).  num_valid (nullable int32 device
 *     scalar): only the first *num_valid pillars are processed (device-side voxel_num).
 */
size_t papc_pfn_workspace_bytes(int P, int cout);
int papc_pfn_f32(const float *features, const int32_t *num_voxels, const int32_t *coors, int P,
                 int T, int F, float vx, float vy, float x_offset, float y_offset,
                 const float *weight, const float *bias, const float *gamma, const float *beta,
                 const float *running_mean, const float *running_var, int bn_mode, float eps,
                 int cout, const int32_t *num_valid, float *out, float *batch_mean,
                 float *batch_var, void *workspace, size_t workspace_bytes,
                 papc_stream_t stream);

/* A10 (general form) one PFNLayer at any position                      pillars.py:9-41, 79-108
 *     decorate != 0: x = features [P,T,F] (F >= 3); the layer input is the decorated tensor
 *       [features, f_cluster(3), f_center(2), (|xyz| when with_distance, :92-94)] with padding rows zeroed
 *       (num_voxels / coors as papc_pfn_f32).  decorate == 0: x [P,T,F] is the previous layer's output (F = cin).
 *     y = x W (+ bias), weight [cin,units]; BatchNorm1D over all P*T rows; ReLU; max over T.
 *     last_layer != 0: out [P,units] = the max (:34-36); else out [P,T,2*units] = [x | repeat(max)] (:38-40).
 *     The default constructor num_filters=(64,128) is papc_pfn_layer_f32(decorate, units 32, not last) followed by
 *     papc_pfn_layer_f32(plain, F = 64, units 128, last).  cin <= 128, units <= 128.
 */
size_t papc_pfn_layer_workspace_bytes(int P, int T, int units);
int papc_pfn_layer_f32(const float *x, int decorate, int with_distance, const int32_t *num_voxels,
                       const int32_t *coors, int P, int T, int F, float vx, float vy, float x_offset,
                       float y_offset, const float *weight, const float *bias, const float *gamma,
                       const float *beta, const float *running_mean, const float *running_var, int bn_mode,
                       float eps, int units, int last_layer, const int32_t *num_valid, float *out,
                       float *batch_mean, float *batch_var, void *workspace, size_t workspace_bytes,
                       papc_stream_t stream);

/* A11 PointPillarsScatter.forward                                     pillars.py:121-142
 *     voxel_features [P,C], coords [P,4] int32 (b,z,y,x) -> canvas [batch,C,ny,nx]; every canvas
 *     element is written exactly once (zero where no pillar).  Duplicate (b,y,x) keep the last
 *     pillar, as NumPy fancy assignment does.  num_valid as above.
 */
size_t papc_pillar_scatter_workspace_bytes(int batch, int ny, int nx);
int papc_pillar_scatter_f32(const float *voxel_features, const int32_t *coords, int P, int C,
                            int batch, int ny, int nx, const int32_t *num_valid, float *canvas,
                            void *workspace, size_t workspace_bytes, papc_stream_t stream);

/* ----------------------------------------------------------------------------------------
 * 8e  the batch-sharded step's one exchange, over peer memory instead of an NCCL all-gather
 *     local [n_per_rank] floats of this rank -> every peer's result buffer at offset rank * n_per_rank.
 *     peer_bufs_host [world] is a HOST array of device pointers to the ranks' result buffers (this rank's own
 *     included), all peer-accessible -- e.g. torch symmetric memory (papc_b200/dist.py: PeerAllGather, which also
 *     runs the signal-pad barriers that open and close the exchange).  world <= 16; 16-byte aligned buffers.
 */
int papc_p2p_allgather_f32(const float *local, int64_t n_per_rank, void *const *peer_bufs_host, int rank,
                           int world, papc_stream_t stream);

/* ----------------------------------------------------------------------------------------
 * N3  detector post-processing      pp/libs/ops/non_max_suppression/nms_gpu.py
 *
 *   papc_nms_f32: nms_gpu (:133-164, box_dim 5: x1, y1, x2, y2, score) and rotate_nms_gpu (:453-488, box_dim 6:
 *     cx, cy, w, h, angle, score).  Boxes are visited by descending score (ties: higher index first, numpy's
 *     stable argsort reversed, :147); a box is kept unless a kept box before it has IoU > thresh (:64, :98, :445).
 *     keep_out [n] int32 receives the ORIGINAL indices of the kept boxes in visiting order (what the reference
 *     returns as list(order[keep])), -1 beyond *num_out (device int32).  Everything stays on the device: no
 *     host sort, no mask round trip, no host suppress loop (cc/nms/nms_kernel.cu.cc:100-157).  n <= 65536.
 *   papc_rotate_iou_f32: rotate_iou_gpu (:518-553, criterion -1) / rotate_iou_gpu_eval (:603-653, criterion
 *     -1 IoU, 0 inter / area(query), 1 inter / area(box), 2 inter): boxes [N,5], query_boxes [K,5] (cx, cy, w, h,
 *     angle clockwise) -> out [N,K].
 *   Arithmetic: fp32 in the reference's operation order, no FMA contraction, cos / sin evaluated in double.
 *   Reference quirk, reproduced: two IDENTICAL rotated boxes do not score IoU 1 (all eight corners enter the
 *   clipped polygon, which then holds every vertex twice: IoU 1/3 or 0, see tests/test_nms_oracle.py).
 */
size_t papc_nms_workspace_bytes(int n);
int papc_nms_f32(const float *dets, int n, int box_dim, float thresh, int32_t *keep_out, int32_t *num_out,
                 void *workspace, size_t workspace_bytes, papc_stream_t stream);
int papc_rotate_iou_f32(const float *boxes, int N, const float *query_boxes, int K, int criterion, float *out,
                        papc_stream_t stream);

/* ----------------------------------------------------------------------------------------
 * N4  the steps either side of the pillar encode
 *
 *   papc_voxelize_batch_f32: points_to_voxel (pc_ops.py:106-166) over a whole batch of frames plus the merged
 *     layout of merge_second_batch (pp/data/preprocess.py:16-42).  points [total,F] = the frames back to back,
 *     frame_offsets_host [batch+1] (HOST, offsets[0] = 0); every frame is voxelised exactly as papc_voxelize_f32
 *     does (first-come numbering, max_points cap, the max_voxels break per frame) and the rows of all frames are
 *     written back to back: voxels [batch*max_voxels,max_points,F], coors4 [batch*max_voxels,4] int32
 *     (b, z, y, x) = np.pad(coors, ((0,0),(1,0)), constant_values=b) (:31-36), num_points [batch*max_voxels],
 *     frame_voxels [batch] int32 (the 'num_voxels' entry), *total_voxels int32 (device); rows >= *total_voxels
 *     are zero.  batch <= 32.
 *   papc_anchors_mask_f32: sparse_sum_for_anchors_mask (pp/libs/ops/box_np_ops.py:772-777) -> dense_map [ny,nx]
 *     (+1 at (coors[i,-2], coors[i,-1]) of every pillar; coor_dim 3 = (z,y,x), 4 = (b,z,y,x)); when n_anchors > 0
 *     the map is then turned into its inclusive 2-D prefix sum in place (the cumsum(0).cumsum(1) the caller of
 *     fused_get_anchors_area applies) and anchors_area [n_anchors] = fused_get_anchors_area(dense_map, anchors_bv
 *     [n_anchors,4] (x1,y1,x2,y2), stride, offset, grid_size) (:781-806).  stride/offset [2], grid_size [2] HOST.
 *   papc_points_to_bev_f32: points_to_bev (pp/libs/ops/point_cloud/bev_ops.py:6-103): bev_map
 *     [D + 1 (+1 with reflectivity), H, W]: per height slice the maximum normalised height, (the intensity of the
 *     last point that raised a maximum of its column), the point count.  height_lowers_host [D] = the reference's
 *     np.linspace(lo_z, hi_z, D, endpoint=False) (:89-90), computed by the caller.  The sequential loop's break at
 *     the first new cell once max_voxels cells exist (:44-46) is reproduced.
 */
size_t papc_voxelize_batch_workspace_bytes(int total_points, int batch, const float *voxel_size_host,
                                           const float *coors_range_host, int max_voxels);
int papc_voxelize_batch_f32(const float *points, const int32_t *frame_offsets_host, int batch, int F,
                            const float *voxel_size_host, const float *coors_range_host, int max_points,
                            int reverse_index, int max_voxels, float *voxels, int32_t *coors4,
                            int32_t *num_points, int32_t *frame_voxels, int32_t *total_voxels,
                            void *workspace, size_t workspace_bytes, papc_stream_t stream);
int papc_anchors_mask_f32(const int32_t *coors, int P, int coor_dim, const int32_t *num_valid, int ny, int nx,
                          const float *anchors_bv, int n_anchors, const float *stride_host, const float *offset_host,
                          const int32_t *grid_size_host, float *dense_map, float *anchors_area,
                          papc_stream_t stream);
size_t papc_points_to_bev_workspace_bytes(int N, const float *voxel_size_host, const float *coors_range_host);
int papc_points_to_bev_f32(const float *points, int N, int F, const float *voxel_size_host,
                           const float *coors_range_host, const float *height_lowers_host,
                           int with_reflectivity, int max_voxels, float *bev_map, void *workspace,
                           size_t workspace_bytes, papc_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PAPC_B200_H_ */
