"""NumPy restatement of PAPC's PointNet++ primitives -- TEST INFRASTRUCTURE, NOT PRODUCT.

Follows /root/reference/PAPC/models/layers/pointnet2_basic_layers.py statement by
statement (line numbers cited per function), replacing every ``paddle.X`` by the
NumPy call of the same meaning.  The reference's ``.numpy()`` / ``paddle.to_tensor``
host round trips (layers.py:57-60, 81-92, 113-116, 120-124) are identities here.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this module; the product never does.

PARITY UNPINNED: the reference executes these lines inside PaddlePaddle, which is not
installable in the build image, and the reference ships no tests or golden vectors.
The restatement is pinned by (i) vectors produced by EXECUTING the reference's own source
over a NumPy stand-in for its paddle calls (tests/golden/make_golden_layers.py ->
tests/golden/layers_ref.npz, checked bit-exactly in tests/test_oracle_vs_reference_source.py;
this pins the logic, not the arithmetic inside Paddle's kernels), (ii) the known-answer
vectors K1-K4 of SURVEY.md section 8(c), (iii) agreement with the arithmetic-pinned C
restatement in oracle/papc_oracle.c.

BatchNorm follows Paddle 2.x semantics: training mode normalises with the *biased*
batch variance over (B, H, W) per channel, eps 1e-5 (BatchNorm2D default), momentum 0.9
(``running = 0.9*running + 0.1*batch``).  Because the SA layers keep their conv/bn
objects in plain Python lists (layers.py:185-190, 230-241) Paddle never registers
them, so they stay in training mode even after ``model.eval()``; ``bn_mode='batch'``
is therefore the default and ``'running'`` is offered for folded inference.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32


# ----------------------------------------------------------------------------------
def pc_normalize(pc):
    """layers.py:17-23."""
    centroid = np.mean(pc, axis=0)
    pc = pc - centroid
    m = np.max(np.sqrt(np.sum(pc ** 2, axis=1)))
    pc = pc / m
    return pc


def square_distance(src, dst):
    """layers.py:26-40.  src [B,N,C], dst [B,M,C] -> [B,N,M] fp32."""
    B, N, _ = src.shape
    _, M, _ = dst.shape
    dist = F32(-2) * np.matmul(src, dst.transpose(0, 2, 1))          # :36
    dist += np.sum(src ** 2, axis=-1).reshape(B, N, 1)               # :37
    dist += np.sum(dst ** 2, axis=-1).reshape(B, 1, M)               # :38
    return dist


def index_points(points, idx):
    """layers.py:43-62.  idx may be float32 (FPS output); cast as :59 does."""
    B = points.shape[0]
    view_shape = list(idx.shape)
    view_shape[1:] = [1] * (len(view_shape) - 1)
    repeat_shape = list(idx.shape)
    repeat_shape[0] = 1
    batch_indices = np.tile(np.arange(B).reshape(view_shape), repeat_shape)  # :56
    idx_np = np.asarray(idx).astype("int64")                                  # :59
    return points[batch_indices.astype("int64"), idx_np, :]                   # :60


def farthest_point_sample(xyz, npoint, start_idx=None, rng=None, init_dist=1.0):
    """layers.py:65-95.  Returns float32-encoded indices [B,npoint] like :74.

    ``start_idx`` replaces ``paddle.randint(0, N, (B,))`` (:76) so runs are
    reproducible; when omitted it is drawn from ``rng``.
    """
    B, N, C = xyz.shape
    centroids = np.zeros([B, npoint], dtype=F32)                     # :74
    distance = np.full([B, N], init_dist, dtype=F32)                 # :75 (ones)
    if start_idx is None:
        rng = rng or np.random.default_rng()
        farthest = rng.integers(0, N, (B,))                          # :76
    else:
        farthest = np.asarray(start_idx).astype("int64")
    batch_indices = np.arange(B)                                     # :77
    for i in range(npoint):                                          # :79
        centroids[:, i] = farthest                                   # :80
        centroid = xyz[batch_indices, farthest, :][:, None, :]       # :84-85
        dist = np.sum((xyz - centroid) ** 2, -1)                     # :86
        mask = dist < distance                                       # :87
        distance[mask] = dist[mask]                                  # :88-92
        farthest = np.argmax(distance, -1)                           # :93 (first max)
    return centroids


def query_ball_point(radius, nsample, xyz, new_xyz):
    """layers.py:98-126.  -> int64 [B,S,nsample]."""
    B, N, C = xyz.shape
    _, S, _ = new_xyz.shape
    group_idx = np.tile(np.arange(N, dtype=np.int64).reshape(1, 1, N), [B, S, 1])  # :110
    sqrdists = square_distance(new_xyz, xyz)                         # :111
    mask = sqrdists > F32(radius ** 2)                               # :112 (python double r^2 -> fp32)
    group_idx[mask] = N                                              # :115
    group_idx = np.sort(group_idx, axis=-1)[:, :, :nsample]          # :117
    group_first = np.tile(group_idx[:, :, 0].reshape(B, S, 1), [1, 1, nsample])   # :118
    mask = group_idx == N                                            # :119
    group_idx[mask] = group_first[mask]                              # :123
    return group_idx


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False, start_idx=None):
    """layers.py:129-157.  xyz-first concat (:151)."""
    B, N, C = xyz.shape
    S = npoint
    fps_idx = farthest_point_sample(xyz, npoint, start_idx=start_idx)    # :143
    new_xyz = index_points(xyz, fps_idx)                                 # :144
    idx = query_ball_point(radius, nsample, xyz, new_xyz)                # :145
    grouped_xyz = index_points(xyz, idx)                                 # :146
    grouped_xyz_norm = grouped_xyz - new_xyz.reshape(B, S, 1, C)         # :147
    if points is not None:
        grouped_points = index_points(points, idx)                       # :150
        new_points = np.concatenate([grouped_xyz_norm, grouped_points], axis=-1)  # :151
    else:
        new_points = grouped_xyz_norm
    if returnfps:
        return new_xyz, new_points, grouped_xyz, fps_idx
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """layers.py:160-176.  No centring; new_xyz is all zeros."""
    B, N, C = xyz.shape
    new_xyz = np.zeros([B, 1, C], dtype=F32)
    grouped_xyz = xyz.reshape(B, 1, N, C)
    if points is not None:
        new_points = np.concatenate([grouped_xyz, points.reshape(B, 1, N, -1)], axis=-1)
    else:
        new_points = grouped_xyz
    return new_xyz, new_points


# ----------------------------------------------------------------------------------
class Conv2D1x1:
    """paddle.nn.Conv2D(cin, cout, 1) restated: weight [cout,cin,1,1], bias [cout]."""

    def __init__(self, cin, cout, rng=None):
        rng = rng or np.random.default_rng(0)
        # Paddle default init: Normal(0, sqrt(2/(k*k*cin))) weight, zero bias.  Parity tests
        # always overwrite these (bias U(-.1,.1)).
        self.weight = (rng.standard_normal((cout, cin, 1, 1)) * np.sqrt(2.0 / cin)).astype(F32)
        self.bias = np.zeros((cout,), dtype=F32)

    def __call__(self, x, acc=np.float64):
        # x [B,Cin,H,W] -> [B,Cout,H,W]; accumulate in float64 (or float32 for the timed port)
        w = self.weight.reshape(self.weight.shape[0], -1)
        B, Cin, H, W = x.shape
        # a 1x1 convolution is a [cout,cin] x [cin,H*W] matrix product per batch item (BLAS)
        y = np.matmul(w.astype(acc), np.ascontiguousarray(x).reshape(B, Cin, H * W).astype(acc))
        y = y.reshape(B, -1, H, W) + self.bias.astype(acc).reshape(1, -1, 1, 1)
        return y.astype(F32)


class BatchNorm2D:
    """paddle.nn.BatchNorm2D(c): eps 1e-5, momentum 0.9, biased batch variance."""

    def __init__(self, c, eps=1e-5, momentum=0.9):
        self.weight = np.ones((c,), dtype=F32)
        self.bias = np.zeros((c,), dtype=F32)
        self._mean = np.zeros((c,), dtype=F32)
        self._variance = np.ones((c,), dtype=F32)
        self.eps = eps
        self.momentum = momentum
        self.training = True

    def __call__(self, x, acc=np.float64):
        axes = tuple(i for i in range(x.ndim) if i != 1)
        shp = [1, -1] + [1] * (x.ndim - 2)
        if self.training:
            xa = x.astype(acc)
            mean = xa.mean(axis=axes)
            var = xa.var(axis=axes)  # biased
            self.last_mean, self.last_var = mean.astype(F32), var.astype(F32)
            self._mean = (self.momentum * self._mean + (1 - self.momentum) * mean).astype(F32)
            self._variance = (self.momentum * self._variance + (1 - self.momentum) * var).astype(F32)
        else:
            mean, var = self._mean.astype(acc), self._variance.astype(acc)
        y = (x.astype(acc) - mean.reshape(shp)) / np.sqrt(var.reshape(shp) + acc(self.eps))
        y = y * self.weight.astype(acc).reshape(shp) + self.bias.astype(acc).reshape(shp)
        return y.astype(F32)


def relu(x):
    return np.maximum(x, F32(0))


class PointNetSetAbstraction:
    """layers.py:179-221."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all, rng=None):
        self.npoint = npoint
        self.radius = radius
        self.nsample = nsample
        self.mlp_convs = []
        self.mlp_bns = []
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(Conv2D1x1(last_channel, out_channel, rng))
            self.mlp_bns.append(BatchNorm2D(out_channel))
            last_channel = out_channel
        self.group_all = group_all
        self.acc = np.float64

    def forward(self, xyz, points, start_idx=None):
        xyz = xyz.transpose(0, 2, 1)                                     # :203
        if points is not None:
            points = points.transpose(0, 2, 1)                           # :205
        if self.group_all:
            new_xyz, new_points = sample_and_group_all(xyz, points)      # :211
        else:
            new_xyz, new_points = sample_and_group(self.npoint, self.radius, self.nsample,
                                                   xyz, points, start_idx=start_idx)  # :213
        new_points = new_points.transpose(0, 3, 2, 1)                    # :214  [B,C,K,S]
        for i, conv in enumerate(self.mlp_convs):
            bn = self.mlp_bns[i]
            new_points = relu(bn(conv(new_points, self.acc), self.acc))  # :217
        new_points = np.max(new_points, 2)                               # :219
        new_xyz = new_xyz.transpose(0, 2, 1)                             # :220
        return new_xyz, new_points

    __call__ = forward


class PointNetSetAbstractionMsg:
    """layers.py:224-281.  features-first concat (:267)."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list, rng=None):
        self.npoint = npoint
        self.radius_list = radius_list
        self.nsample_list = nsample_list
        self.conv_blocks = []
        self.bn_blocks = []
        for i in range(len(mlp_list)):
            convs, bns = [], []
            last_channel = in_channel + 3
            for out_channel in mlp_list[i]:
                convs.append(Conv2D1x1(last_channel, out_channel, rng))
                bns.append(BatchNorm2D(out_channel))
                last_channel = out_channel
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)
        self.acc = np.float64

    def forward(self, xyz, points, start_idx=None):
        xyz = xyz.transpose(0, 2, 1)
        if points is not None:
            points = points.transpose(0, 2, 1)
        B, N, C = xyz.shape
        S = self.npoint
        new_xyz = index_points(xyz, farthest_point_sample(xyz, S, start_idx=start_idx))  # :258
        new_points_list = []
        for i, radius in enumerate(self.radius_list):
            K = self.nsample_list[i]
            group_idx = query_ball_point(radius, K, xyz, new_xyz)        # :262
            grouped_xyz = index_points(xyz, group_idx)                   # :263
            grouped_xyz = grouped_xyz - new_xyz.reshape(B, S, 1, C)      # :264
            if points is not None:
                grouped_points = index_points(points, group_idx)         # :266
                grouped_points = np.concatenate([grouped_points, grouped_xyz], axis=-1)  # :267
            else:
                grouped_points = grouped_xyz
            grouped_points = grouped_points.transpose(0, 3, 2, 1)        # :271
            for j in range(len(self.conv_blocks[i])):
                conv = self.conv_blocks[i][j]
                bn = self.bn_blocks[i][j]
                grouped_points = relu(bn(conv(grouped_points, self.acc), self.acc))  # :275
            new_points = np.max(grouped_points, 2)                       # :276
            new_points_list.append(new_points)
        new_xyz = new_xyz.transpose(0, 2, 1)
        new_points_concat = np.concatenate(new_points_list, axis=1)      # :280
        return new_xyz, new_points_concat

    __call__ = forward


# ----------------------------------------------------------------------------------
def grouped_mlp(new_points, weights, biases, gammas, betas, eps=1e-5, bn_mode="batch",
                running_mean=None, running_var=None, acc=np.float64):
    """(Conv1x1 -> BN -> ReLU) x L -> max over K on an explicit grouped tensor.

    new_points [B,S,K,Cin] (the layout sample_and_group returns) -> [B,Cout,S]; this is
    layers.py:214-219 with explicit parameter arrays, used to check the CUDA grouped MLP.
    Also returns the per-layer batch mean / biased variance.
    """
    x = new_points.transpose(0, 3, 2, 1)
    stats = []
    for l, w in enumerate(weights):
        conv = Conv2D1x1(w.shape[1], w.shape[0])
        conv.weight = w.reshape(w.shape[0], w.shape[1], 1, 1).astype(F32)
        conv.bias = (biases[l] if biases[l] is not None else np.zeros(w.shape[0])).astype(F32)
        bn = BatchNorm2D(w.shape[0], eps=eps)
        bn.weight, bn.bias = gammas[l].astype(F32), betas[l].astype(F32)
        if bn_mode == "running":
            bn.training = False
            bn._mean, bn._variance = running_mean[l].astype(F32), running_var[l].astype(F32)
        x = relu(bn(conv(x, acc), acc))
        if bn_mode == "batch":
            stats.append((bn.last_mean, bn.last_var))
    return np.max(x, 2), stats


# ----------------------------------------------------------------------------------
class PointNetFeaturePropagation:
    """layers.py:284-335, statement by statement.

    Conv1D(k=1) / BatchNorm1D are restated with the 2-D holders above on a [B,C,N,1] view (a
    1x1 convolution and a per-channel normalisation over (B, N) are the same arithmetic).

    Reference quirk kept on purpose: ``dists`` is SORTED first (:316) and ``idx`` is the argsort
    of the already sorted array (:317), i.e. the identity permutation -- so the three *smallest
    distances* weight the features of sampled points 0, 1, 2, not of the three nearest ones.
    (``np.argsort(kind='stable')`` makes the identity exact under ties; with ties Paddle may swap
    equal entries, which leaves the weighted sum unchanged unless the tie straddles positions 2|3.)
    """

    def __init__(self, in_channel, mlp, rng=None):
        self.mlp_convs = []
        self.mlp_bns = []
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(Conv2D1x1(last_channel, out_channel, rng))   # nn.Conv1D(.., 1)  :292
            self.mlp_bns.append(BatchNorm2D(out_channel))                      # nn.BatchNorm1D    :293
            last_channel = out_channel
        self.acc = np.float64

    def interpolate(self, xyz1, xyz2, points1, points2):
        """:306-329 -> new_points [B, N, D1+D2] (channels-last, before the transpose at :331)."""
        xyz1 = xyz1.transpose(0, 2, 1)                                   # :306
        xyz2 = xyz2.transpose(0, 2, 1)                                   # :307
        points2 = points2.transpose(0, 2, 1)                             # :309
        B, N, C = xyz1.shape
        _, S, _ = xyz2.shape
        if S == 1:
            interpolated_points = np.tile(points2, [1, N, 1])            # :314
        else:
            dists = square_distance(xyz1, xyz2)                          # :316
            dists = np.sort(dists, axis=-1)                              # :317
            idx = np.argsort(dists, axis=-1, kind="stable")              # :318 (of the SORTED array)
            dists, idx = dists[:, :, :3], idx[:, :, :3]                  # :319
            dist_recip = F32(1.0) / (dists + F32(1e-8))                  # :321
            norm = np.sum(dist_recip, axis=2, keepdims=True)             # :322
            weight = dist_recip / norm                                   # :323
            k3 = idx.shape[2]
            interpolated_points = np.sum(index_points(points2, idx) * weight.reshape(B, N, k3, 1), axis=2)  # :324
        if points1 is not None:
            points1 = points1.transpose(0, 2, 1)                         # :327
            new_points = np.concatenate([points1, interpolated_points], axis=-1)  # :328
        else:
            new_points = interpolated_points
        return new_points.astype(F32)

    def forward(self, xyz1, xyz2, points1, points2):
        new_points = self.interpolate(xyz1, xyz2, points1, points2)
        new_points = new_points.transpose(0, 2, 1)                       # :332  [B,C,N]
        x = new_points[:, :, :, None]
        for i, conv in enumerate(self.mlp_convs):
            bn = self.mlp_bns[i]
            x = relu(bn(conv(x, self.acc), self.acc))                    # :335
        return x[:, :, :, 0]

    __call__ = forward
