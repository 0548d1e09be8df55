"""Host-side mirror of PAPC/models/layers/pointnet2_basic_layers.py over the sm_100a kernels.

Same callables, argument order, shapes and dtypes as the reference (file:line cited per
function), with torch CUDA tensors where the reference has Paddle tensors -- PaddlePaddle is
not installable in this environment, see INTEGRATION.md for the Paddle-side stub.  Every op
launches hand-written CUDA through the C ABI (include/papc_b200.h); nothing here computes on
the CPU and nothing falls back to PyTorch ops for the hot path.

Differences a caller can see (all opt-in):
  * ``farthest_point_sample`` / ``sample_and_group`` / the layer ``forward`` accept an explicit
    ``start_idx`` so runs are reproducible; omitted, it is drawn with ``torch.randint`` exactly
    where the reference calls ``paddle.randint`` (layers.py:76).
  * the SetAbstraction layers never materialise the ``[B,S,K,3+D]`` grouped tensor and return
    ``[B,D',S]`` as a transposed *view* of channels-last storage (same shape and values).
  * ``bn_mode``: ``'batch'`` (default) reproduces what the reference's unregistered conv/bn lists
    always do -- training-mode batch statistics, even after ``model.eval()``;
    ``'running'`` normalises with the stored running statistics.
"""
from __future__ import annotations

import ctypes as C
import math
import os

import torch
import torch.distributed as dist

from . import _lib as L


# ----------------------------------------------------------------------------------------------
def _ws(nbytes, device):
    """Scratch buffer from torch's stream-aware caching allocator (256-byte aligned)."""
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)


def square_distance(src, dst):
    """layers.py:26-40.  src [B,N,3], dst [B,M,3] -> [B,N,M] fp32."""
    L.require_cuda(src, dst)
    src, dst = L.f32c(src), L.f32c(dst)
    B, N, Cc = src.shape
    _, M, _ = dst.shape
    if Cc != 3 or dst.shape[2] != 3 or dst.shape[0] != B:
        raise ValueError("square_distance expects src [B,N,3] and dst [B,M,3]")
    out = torch.empty((B, N, M), dtype=torch.float32, device=src.device)
    L.check(L.lib().papc_square_distance_f32(L.ptr(src), L.ptr(dst), B, N, M, L.ptr(out),
                                             L.stream_ptr(src.device)), "square_distance")
    return out


def index_points(points, idx, strict=False):
    """layers.py:43-62.  points [B,N,C]; idx [B,S] or [B,S,K] (float32 or integer, cast to int64
    as :59 does) -> [B,S,C] / [B,S,K,C].

    Out-of-range indices: the reference raises IndexError (NumPy fancy indexing, :58); the kernel
    clamps them to [0, N-1] so that it never reads out of bounds.  ``strict=True`` restores the
    reference behaviour at the price of a device synchronisation."""
    L.require_cuda(points, idx)
    points = L.f32c(points)
    B, N, Cc = points.shape
    idx64 = idx.to(torch.int64).contiguous()
    if idx64.shape[0] != B:
        raise ValueError("index_points: batch mismatch")
    if strict and idx64.numel() and (int(idx64.min()) < -N or int(idx64.max()) >= N):
        raise IndexError(f"index_points: index out of range for {N} points")
    M = idx64.numel() // max(B, 1)
    out = torch.empty(tuple(idx64.shape) + (Cc,), dtype=torch.float32, device=points.device)
    L.check(L.lib().papc_gather_f32(L.ptr(points), L.ptr(idx64), B, N, Cc, M, L.ptr(out),
                                    L.stream_ptr(points.device)), "index_points")
    return out


_START_QUEUE = []   # seeded draws for layers called WITHOUT start_idx (papc_b200.compat.queue_fps_starts)


def _draw_start(B, N, device, start_idx):
    if start_idx is None and _START_QUEUE:
        start_idx = _START_QUEUE.pop(0)
    if start_idx is None:
        return torch.randint(0, N, (B,), device=device, dtype=torch.int64)  # layers.py:76
    s = torch.as_tensor(start_idx, device=device).to(torch.int64).reshape(B).contiguous()
    return s


def farthest_point_sample_idx(xyz, npoint, start_idx=None, return_xyz=False, init_dist=1.0):
    """Native form of layers.py:65-95: int64 indices [B,npoint] (and the gathered new_xyz)."""
    L.require_cuda(xyz)
    xyz = L.f32c(xyz)
    B, N, Cc = xyz.shape
    if Cc != 3:
        raise ValueError("farthest_point_sample expects xyz [B,N,3]")
    if npoint < 0:
        raise ValueError("npoint must be >= 0")
    start = _draw_start(B, N, xyz.device, start_idx)
    out = torch.empty((B, npoint), dtype=torch.int64, device=xyz.device)
    new_xyz = torch.empty((B, npoint, 3), dtype=torch.float32, device=xyz.device) if return_xyz else None
    lib = L.lib()
    wsb = lib.papc_fps_workspace_bytes(B, N)
    ws = _ws(wsb, xyz.device) if wsb else None
    L.check(lib.papc_fps_f32(L.ptr(xyz), B, N, npoint, L.ptr(start), float(init_dist), L.ptr(out),
                             L.ptr(new_xyz), L.ptr(ws), wsb, L.stream_ptr(xyz.device)),
            "farthest_point_sample")
    return (out, new_xyz) if return_xyz else out


def farthest_point_sample(xyz, npoint, start_idx=None):
    """layers.py:65-95.  Returns float32-encoded indices [B,npoint] as the reference does
    (``centroids = paddle.zeros([B, npoint])``, :74); ``index_points`` casts them back."""
    return farthest_point_sample_idx(xyz, npoint, start_idx).to(torch.float32)


def radius2_f32(radius):
    """``radius ** 2`` is a Python double compared against an fp32 tensor (layers.py:112)."""
    return float(torch.tensor(float(radius) ** 2, dtype=torch.float32).item())


def _ball_query(radius, nsample, xyz, new_xyz, idx_dtype, check_empty=False):
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    if nsample > N:
        # the reference fails with a shape mismatch at layers.py:118-123
        raise ValueError(f"query_ball_point: nsample ({nsample}) > N ({N})")
    out = torch.empty((B, S, nsample), dtype=idx_dtype, device=xyz.device)
    empty = torch.zeros((1,), dtype=torch.int32, device=xyz.device) if check_empty else None
    L.check(L.lib().papc_ball_query_f32(L.ptr(xyz), L.ptr(new_xyz), B, N, S, radius2_f32(radius),
                                        nsample, L.ptr(out), 64 if idx_dtype == torch.int64 else 32,
                                        L.ptr(empty), L.stream_ptr(xyz.device)), "query_ball_point")
    if check_empty and int(empty.item()) > 0:
        # the reference raises IndexError in index_points (index N out of range)
        raise IndexError(f"query_ball_point: {int(empty.item())} query points have no neighbour "
                         f"within radius {radius}")
    return out


def query_ball_point(radius, nsample, xyz, new_xyz, check_empty=False):
    """layers.py:98-126.  xyz [B,N,3], new_xyz [B,S,3] -> int64 [B,S,nsample].

    A query point with no neighbour inside the radius yields index N in every slot (what the
    reference's sort leaves there, :115-117); the reference then fails with IndexError in
    ``index_points``.  ``check_empty=True`` reproduces that error (one device synchronisation);
    the default keeps the call asynchronous and downstream gathers clamp N to N-1."""
    L.require_cuda(xyz, new_xyz)
    return _ball_query(radius, nsample, L.f32c(xyz), L.f32c(new_xyz), torch.int64, check_empty)


def _ball_query_multi(queries, xyz, new_xyz, idx_dtype):
    """All (radius, nsample) pairs of an MSG layer in ONE pass over the distances (papc_ball_query_multi_f32);
    same results as one ``_ball_query`` per pair."""
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    R = len(queries)
    if R == 1 or R > 4:
        return [_ball_query(r, k, xyz, new_xyz, idx_dtype) for r, k in queries]
    for _, k in queries:
        if k > N:
            raise ValueError(f"query_ball_point: nsample ({k}) > N ({N})")
    outs = [torch.empty((B, S, k), dtype=idx_dtype, device=xyz.device) for _, k in queries]
    r2 = (C.c_float * R)(*[radius2_f32(r) for r, _ in queries])
    ks = (C.c_int32 * R)(*[int(k) for _, k in queries])
    ptrs = (C.c_void_p * R)(*[o.data_ptr() for o in outs])
    L.check(L.lib().papc_ball_query_multi_f32(L.ptr(xyz), L.ptr(new_xyz), B, N, S, R, r2, ks, ptrs,
                                              64 if idx_dtype == torch.int64 else 32, None,
                                              L.stream_ptr(xyz.device)), "query_ball_point (multi-radius)")
    return outs


def query_ball_point_multi(radius_list, nsample_list, xyz, new_xyz):
    """The per-radius ``query_ball_point`` calls of PointNetSetAbstractionMsg.forward (layers.py:258-267) as one
    kernel: -> list of int64 [B,S,nsample_i]."""
    L.require_cuda(xyz, new_xyz)
    return _ball_query_multi(list(zip(radius_list, nsample_list)), L.f32c(xyz), L.f32c(new_xyz), torch.int64)


def knn_points(k, xyz, new_xyz, return_dist=False):
    """k nearest neighbours of every ``new_xyz`` point in ``xyz`` under ``square_distance`` (stable order:
    ascending distance, ties to the lower index).  xyz [B,N,3], new_xyz [B,S,3] -> int64 [B,S,k]
    (+ fp32 distances).  The neighbour search inside PointNetFeaturePropagation (layers.py:316-318), standalone."""
    L.require_cuda(xyz, new_xyz)
    xyz, new_xyz = L.f32c(xyz), L.f32c(new_xyz)
    B, N, _ = xyz.shape
    S = new_xyz.shape[1]
    if not 1 <= k <= 32:
        raise ValueError("knn_points: 1 <= k <= 32")
    idx = torch.empty((B, S, k), dtype=torch.int64, device=xyz.device)
    dist_ = torch.empty((B, S, k), dtype=torch.float32, device=xyz.device) if return_dist else None
    L.check(L.lib().papc_knn_f32(L.ptr(xyz), L.ptr(new_xyz), B, N, S, k, L.ptr(idx), 64, L.ptr(dist_),
                                 L.stream_ptr(xyz.device)), "knn_points")
    return (idx, dist_) if return_dist else idx


def _group_gather(xyz, new_xyz, feats, idx64, order):
    B, N, _ = xyz.shape
    _, S, K = idx64.shape
    D = 0 if feats is None else feats.shape[2]
    out = torch.empty((B, S, K, 3 + D), dtype=torch.float32, device=xyz.device)
    L.check(L.lib().papc_group_gather_f32(L.ptr(xyz), L.ptr(new_xyz), L.ptr(feats), L.ptr(idx64), B, N,
                                          S, K, D, order, L.ptr(out), L.stream_ptr(xyz.device)),
            "group_gather")
    return out


def sample_and_group(npoint, radius, nsample, xyz, points, returnfps=False, start_idx=None):
    """layers.py:129-157.  -> new_xyz [B,S,3], new_points [B,S,K,3+D] (xyz-first concat, :151)."""
    L.require_cuda(xyz, points)
    xyz = L.f32c(xyz)
    points = L.f32c(points)
    fps_idx, new_xyz = farthest_point_sample_idx(xyz, npoint, start_idx, return_xyz=True)  # :143-144
    idx = _ball_query(radius, nsample, xyz, new_xyz, torch.int64)                           # :145
    new_points = _group_gather(xyz, new_xyz, points, idx, L.XYZ_FIRST)                      # :146-151
    if returnfps:
        grouped_xyz = index_points(xyz, idx)
        return new_xyz, new_points, grouped_xyz, fps_idx.to(torch.float32)
    return new_xyz, new_points


def sample_and_group_all(xyz, points):
    """layers.py:160-176.  new_xyz zeros [B,1,3]; new_points [B,1,N,3+D], no centring."""
    L.require_cuda(xyz, points)
    xyz = L.f32c(xyz)
    B, N, Cc = xyz.shape
    new_xyz = torch.zeros((B, 1, Cc), dtype=torch.float32, device=xyz.device)
    if points is None:
        return new_xyz, xyz.reshape(B, 1, N, Cc)
    points = L.f32c(points)
    idx = torch.arange(N, device=xyz.device, dtype=torch.int64).reshape(1, 1, N).expand(B, 1, N).contiguous()
    return new_xyz, _group_gather(xyz, new_xyz, points, idx, L.XYZ_FIRST)


# ---------------------------------------------------------------------------------------------
# Sampling overlap.  farthest_point_sample / query_ball_point of a SetAbstraction layer read only the
# xyz coordinates, i.e. for every layer but the first the *sampled centroids* of the previous layer
# -- available long before that layer's MLP has finished.  Each layer therefore tags the xyz tensor
# it returns with the CUDA event recorded right after its own FPS, and a layer whose input carries
# such a tag runs its FPS + ball query on a high-priority side stream that waits for that event
# only; the main stream joins before the MLP.  The sampling kernels are small (one CTA per cloud, a
# few KB of shared memory) and co-reside with the previous layer's MLP kernels.  Transparent to the
# caller (same call sequence, same results); PAPC_OVERLAP_SAMPLING=0 switches it off.
OVERLAP_SAMPLING = os.environ.get("PAPC_OVERLAP_SAMPLING", "1") != "0"
_SIDE_STREAMS = {}


def _side_stream(device):
    key = torch.device(device).index
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = torch.cuda.Stream(device=device, priority=-1)
    return _SIDE_STREAMS[key]


FUSED_SAMPLING = os.environ.get("PAPC_FUSED_SAMPLING", "1") != "0"


def _sample_group_fused(xyz, npoint, start_idx, radius, nsample, want_moments):
    """FPS + ball query (+ the second moments of the centred neighbours) in ONE launch
    (``papc_sample_group_f32``): the ball query of centroid i runs on other SMs while the FPS recurrence is
    still producing the next centroids.  Bit-identical to ``farthest_point_sample_idx`` + ``_ball_query``.
    Returns None when the shape does not run fused."""
    B, N, _ = xyz.shape
    lib = L.lib()
    P = lib.papc_sample_group_parts(B, N, npoint, nsample)
    if P == 0:
        return None
    dev = xyz.device
    start = _draw_start(B, N, dev, start_idx)
    fps = torch.empty((B, npoint), dtype=torch.int64, device=dev)
    new_xyz = torch.empty((B, npoint, 3), dtype=torch.float32, device=dev)
    idx = torch.empty((B, npoint, nsample), dtype=torch.int32, device=dev)
    mom = torch.empty((B * P, 9), dtype=torch.float64, device=dev) if want_moments else None
    wsb = lib.papc_sample_group_workspace_bytes(B, npoint)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    L.check(lib.papc_sample_group_f32(L.ptr(xyz), B, N, npoint, L.ptr(start), 1.0, radius2_f32(radius), nsample,
                                      L.ptr(fps), L.ptr(new_xyz), L.ptr(idx), None, L.ptr(mom), L.ptr(ws), wsb,
                                      L.stream_ptr(dev)), "sample_and_group (fused)")
    return new_xyz, idx, mom


_KEEPALIVE = {}    # device -> buffers of the most recent side-stream sampling (see _sample)


def precede(t):
    """Mark a CUDA tensor (e.g. a ``start_idx``) as complete at THIS point of the current stream: an event is
    recorded and attached, and an overlapped layer waits for exactly that event instead of for everything the
    main stream has been given so far (which would include the previous layer's MLP and undo the overlap)."""
    if isinstance(t, torch.Tensor) and t.is_cuda:
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(t.device))
        t._papc_ev = ev
    return t


def _sample(xyz, ready, npoint, start_idx, queries, want_moments=False):
    """FPS + one ball query per (radius, nsample) in ``queries`` -> (new_xyz, [idx int32...], event
    recorded right after the FPS[, moment partials when ``want_moments``: one query, fused launch, else None]).  Runs on the side stream when ``ready`` -- the producer layer's
    (FPS-done event, entry event) pair -- is set, so that the sampling of layer n+1 overlaps the MLP of layer n.

    Stream safety without giving the overlap up:
      * the side stream waits for the producer's FPS (``ready[0]``) and for the producer layer's ENTRY event
        (``ready[1]``, recorded on the main stream before the producer enqueued its MLP): everything the main
        stream did before the producer layer is complete, the producer's own MLP is not waited for;
      * a CUDA ``start_idx`` is waited for through the event ``precede()`` attached to it (the stacks / models
        attach one at their entry); an unmarked one falls back to waiting for the main stream as it is now;
      * buffers allocated here are read by main-stream kernels after this call returns: outside a capture
        they are ``record_stream``-ed to the main stream (the allocator defers their reuse); in every mode the
        most recent sampling's buffers are also kept alive until the NEXT overlapped sampling has made its
        allocations, and that one waits for this layer's entry event, i.e. for the MLP that read them."""
    dev = xyz.device
    main = torch.cuda.current_stream(dev)
    if ready is not None and OVERLAP_SAMPLING:
        side = _side_stream(dev)
        ready_ev, entry_ev = ready
        if entry_ev is not None:
            side.wait_event(entry_ev)
        if isinstance(start_idx, torch.Tensor) and start_idx.is_cuda:
            ev0 = getattr(start_idx, "_papc_ev", None)
            if ev0 is None:
                ev0 = torch.cuda.Event()
                ev0.record(main)           # conservative: everything enqueued on main so far
            side.wait_event(ev0)
        side.wait_event(ready_ev)
        with torch.cuda.stream(side):
            # separate small kernels here: the fused launch (512-thread CTAs) cannot co-reside with the
            # previous layer's MLP kernels, which is the whole point of this branch
            new_xyz, idxs, ev, mom = _sample_kernels(xyz, npoint, start_idx, queries, False, side, fused=False)
            done = torch.cuda.Event()
            done.record(side)
        main.wait_event(done)
        bufs = [new_xyz] + idxs + ([mom] if mom is not None else [])
        if not torch.cuda.is_current_stream_capturing():
            for t in bufs:
                t.record_stream(main)
        _KEEPALIVE[dev] = bufs
        return (new_xyz, idxs, ev, mom) if want_moments else (new_xyz, idxs, ev)
    new_xyz, idxs, ev, mom = _sample_kernels(xyz, npoint, start_idx, queries, want_moments, main)
    return (new_xyz, idxs, ev, mom) if want_moments else (new_xyz, idxs, ev)


def _sample_kernels(xyz, npoint, start_idx, queries, want_moments, stream, fused=True):
    """The sampling launches on the current stream: the fused kernel for a single-radius layer whose shape
    qualifies, else FPS followed by the (multi-radius) ball query."""
    if fused and FUSED_SAMPLING and len(queries) == 1:
        r = _sample_group_fused(xyz, npoint, start_idx, queries[0][0], queries[0][1], want_moments)
        if r is not None:
            new_xyz, idx, mom = r
            ev = torch.cuda.Event()
            ev.record(stream)
            return new_xyz, [idx], ev, mom
    _, new_xyz = farthest_point_sample_idx(xyz, npoint, start_idx, return_xyz=True)
    ev = torch.cuda.Event()
    ev.record(stream)
    idxs = _ball_query_multi(queries, xyz, new_xyz, torch.int32)
    return new_xyz, idxs, ev, None


def _entry_event(dev):
    """Recorded on the main stream when a sampled layer is entered (before its MLP is enqueued)."""
    ev = torch.cuda.Event()
    ev.record(torch.cuda.current_stream(dev))
    return ev


def _tag_ready(t, ev, entry_ev=None):
    """Attach the producer's events (and the tensor's version counter: an in-place modification by the
    caller after the layer returned invalidates the tag, see ``_ready_event``)."""
    t._papc_ready = (ev, entry_ev, t._version)
    return t


def _ready_event(t):
    tag = getattr(t, "_papc_ready", None)
    if tag is None:
        return None
    ev, entry_ev, version = tag
    return (ev, entry_ev) if t._version == version else None


DEFAULT_DEVICE = None   # device of freshly built parameter holders (None = CPU until .to(); compat.install sets cuda)


class Conv2D:
    """Parameter holder mirroring ``paddle.nn.Conv2D(cin, cout, 1)``: weight [cout,cin,1,1], bias
    [cout].  Default init as Paddle: Normal(0, sqrt(2/fan_in)) weight, zero bias."""

    def __init__(self, in_channels, out_channels, kernel_size=1, device=None, generator=None):
        assert kernel_size == 1
        device = DEFAULT_DEVICE if device is None else device
        std = math.sqrt(2.0 / in_channels)
        self.weight = (torch.randn((out_channels, in_channels, 1, 1), generator=generator) * std).to(device)
        self.bias = torch.zeros((out_channels,), device=device)

    def _apply(self, fn):
        self.weight, self.bias = fn(self.weight), fn(self.bias)


class BatchNorm2D:
    """Parameter holder mirroring ``paddle.nn.BatchNorm2D(c)`` (epsilon 1e-5, momentum 0.9):
    weight/bias and the running ``_mean`` / ``_variance`` under Paddle's attribute names."""

    def __init__(self, num_features, momentum=0.9, epsilon=1e-5, device=None):
        device = DEFAULT_DEVICE if device is None else device
        self.weight = torch.ones((num_features,), device=device)
        self.bias = torch.zeros((num_features,), device=device)
        self._mean = torch.zeros((num_features,), device=device)
        self._variance = torch.ones((num_features,), device=device)
        self._momentum = momentum
        self._epsilon = epsilon

    def _apply(self, fn):
        self.weight, self.bias = fn(self.weight), fn(self.bias)
        self._mean, self._variance = fn(self._mean), fn(self._variance)


class _MlpRunner:
    """Builds the C-ABI descriptors for one conv/bn stack and runs the grouped MLP."""

    def __init__(self, convs, bns):
        self.convs, self.bns = convs, bns
        self.last_batch_stats = None

    def _mlp_struct(self, cin, bn_mode, device, want_stats):
        mlp = L.Mlp()
        mlp.num_layers = len(self.convs)
        mlp.cin = cin
        mlp.bn_mode = L.BN_BATCH if bn_mode == "batch" else L.BN_RUNNING
        mlp.eps = float(self.bns[0]._epsilon)
        keep = []
        stats = []
        for l, (conv, bn) in enumerate(zip(self.convs, self.bns)):
            w = L.f32c(conv.weight.reshape(conv.weight.shape[0], -1))
            tensors = [w, L.f32c(conv.bias) if conv.bias is not None else None, L.f32c(bn.weight),
                       L.f32c(bn.bias), L.f32c(bn._mean), L.f32c(bn._variance)]
            keep.append(tensors)
            ly = mlp.layers[l]
            ly.weight, ly.bias, ly.gamma, ly.beta, ly.running_mean, ly.running_var = [
                (t.data_ptr() if t is not None else None) for t in tensors]
            ly.cout = w.shape[0]
            if want_stats and bn_mode == "batch":
                bm = torch.empty((w.shape[0],), dtype=torch.float32, device=device)
                bv = torch.empty((w.shape[0],), dtype=torch.float32, device=device)
                ly.batch_mean, ly.batch_var = bm.data_ptr(), bv.data_ptr()
                stats.append((bm, bv))
        return mlp, keep, stats

    def run(self, src, keep_src, cin, B, S, bn_mode, device, update_running=False, sync=(False, None)):
        """-> [B,S,cout] channels-last.  ``sync`` = (enabled, process group) from ``_SAMixin._sync_group``."""
        if bn_mode not in ("batch", "running"):
            raise ValueError("bn_mode must be 'batch' or 'running'")
        if any(c.weight.device != device for c in self.convs):
            raise L.PapcError("layer parameters are not on the input's device; call .to(device)")
        cout = self.convs[-1].weight.shape[0]
        out = torch.empty((B, S, cout), dtype=torch.float32, device=device)
        lib = L.lib()
        st = L.stream_ptr(device)
        if bn_mode == "batch" and sync[0]:
            self._run_stepwise_synced(src, cin, B, S, out, device, sync[1], update_running)
            return out
        mlp, keep, stats = self._mlp_struct(cin, bn_mode, device, update_running)
        wsb = lib.papc_sa_mlp_workspace_bytes(C.byref(src), C.byref(mlp))
        ws = _ws(wsb, device)
        L.check(lib.papc_sa_mlp_f32(C.byref(src), C.byref(mlp), L.ptr(out), L.OUT_BSC, L.ptr(ws), wsb, st),
                "sa_mlp")
        if stats:
            self.last_batch_stats = stats
            self._update_running(stats)
        del keep, keep_src
        return out

    def _update_running(self, stats):
        # Paddle: running = momentum*running + (1-momentum)*batch, biased batch variance
        for bn, (bm, bv) in zip(self.bns, stats):
            bn._mean.mul_(bn._momentum).add_(bm, alpha=1.0 - bn._momentum)
            bn._variance.mul_(bn._momentum).add_(bv, alpha=1.0 - bn._momentum)

    def _run_stepwise_synced(self, src, cin, B, S, out, device, group, update_running):
        """Batch-sharded SyncBN: per layer, all-reduce the [2,cout] fp64 sums over the ranks
         so sharded == unsharded results."""
        lib = L.lib()
        st = L.stream_ptr(device)
        K = src.K
        M = B * S * K
        total = float(M) * dist.get_world_size(group)  # equal shards (papc_b200.dist.shard_range)
        prows = lib.papc_mlp_stats_partial_rows(M)
        x = None
        scale = shift = None
        stats = []
        G = B * S
        for l, (conv, bn) in enumerate(zip(self.convs, self.bns)):
            w = L.f32c(conv.weight.reshape(conv.weight.shape[0], -1))
            cout = w.shape[0]
            last = l == len(self.convs) - 1
            y = None if last else torch.empty((M, cout), dtype=torch.float32, device=device)
            pmax = torch.empty((G, cout), dtype=torch.float32, device=device) if last else None
            pmin = torch.empty((G, cout), dtype=torch.float32, device=device) if last else None
            partial = torch.empty((prows, 2, cout), dtype=torch.float64, device=device)
            bias = L.f32c(conv.bias) if conv.bias is not None else None
            lwsb = lib.papc_mlp_layer_workspace_bytes(cin, cout)
            lws = _ws(lwsb, device)
            L.check(lib.papc_mlp_layer_forward_f32(C.byref(src) if l == 0 else None, L.ptr(x),
                                                   L.ptr(scale), L.ptr(shift), M, cin, cout, K, L.ptr(w),
                                                   L.ptr(bias), L.ptr(y), L.ptr(pmax), L.ptr(pmin),
                                                   L.ptr(partial), L.ptr(lws), lwsb, st),
                    "mlp_layer_forward")
            sums = torch.empty((2, cout), dtype=torch.float64, device=device)
            L.check(lib.papc_mlp_stats_reduce_f64(L.ptr(partial), prows, cout, L.ptr(sums), st),
                    "mlp_stats_reduce")
            dist.all_reduce(sums, group=group)
            scale = torch.empty((cout,), dtype=torch.float32, device=device)
            shift = torch.empty((cout,), dtype=torch.float32, device=device)
            bm = torch.empty((cout,), dtype=torch.float32, device=device)
            bv = torch.empty((cout,), dtype=torch.float32, device=device)
            L.check(lib.papc_bn_scale_shift_f32(L.ptr(sums), total, L.ptr(L.f32c(bn.weight)),
                                                L.ptr(L.f32c(bn.bias)), float(bn._epsilon), cout,
                                                L.ptr(scale), L.ptr(shift), L.ptr(bm), L.ptr(bv), st),
                    "bn_scale_shift")
            stats.append((bm, bv))
            x, cin = y, cout
        L.check(lib.papc_sa_pool_finish_f32(L.ptr(pmax), L.ptr(pmin), L.ptr(scale), L.ptr(shift), B, S,
                                            cout, L.ptr(out), L.OUT_BSC, st), "sa_pool_finish")
        self.last_batch_stats = stats
        if update_running:
            self._update_running(stats)


# fp16 operand split for the GATHERED first layer of a SetAbstraction module fed by post-ReLU features: built,
# parity-green, and measured SLOWER on B200 than 3xTF32 (sa2.l1 102 vs 74 us, sa3.l1 37 vs 25 us: the kernel is
# bound by its producer warps, and the fp16 split + the column-scale division cost more producer instructions per
# element than the TF32 split saves in tensor-core products) -- off unless PAPC_F16_GATHER=1.
F16_GATHER = os.environ.get("PAPC_F16_GATHER", "0") == "1"


def _feature_colscale(bn, count):
    """Powers of two c_k with 0 <= relu(bn(y))[., k] / c_k < 2^15 for ANY input of a train-mode BatchNorm over
    ``count`` rows: a sample is at most sqrt(count - 1) standard deviations from the batch mean, so
    |bn(y)| <= |gamma| sqrt(count) + |beta| (the same bound the layer kernels use between MLP layers,
    f16_colscale_sq in sa_mlp_tt.cuh).  Cached on the holder; recomputed when gamma / beta change."""
    key = (bn.weight.data_ptr(), bn.weight._version, bn.bias.data_ptr(), bn.bias._version, int(count))
    cache = getattr(bn, "_papc_cs", None)
    if cache is not None and cache[0] == key:
        return cache[1]
    bound = (bn.weight.double().abs() * math.sqrt(float(count)) + bn.bias.double().abs()) * 1.001
    cs = torch.exp2(torch.ceil(torch.log2(torch.clamp(bound / 32000.0, min=1.0)))).to(torch.float32).contiguous()
    bn._papc_cs = (key, cs)
    return cs


def _tag_colscale(t, cs):
    """Attach the bound of a layer's post-ReLU output features (consumed by the next layer's gather kernel)."""
    t._papc_colscale = (cs, t._version)
    return t


def _colscale_of(t, D):
    tag = getattr(t, "_papc_colscale", None) if t is not None else None
    if tag is None or not F16_GATHER:
        return None
    cs, version = tag
    return cs if (t._version == version and cs.numel() == D and cs.device == t.device) else None


def _make_src(xyz, new_xyz, feats, idx32, B, N, S, K, order, moments=None, feats_colscale=None):
    src = L.GroupSource()
    if feats_colscale is not None:   # [D] power-of-two bounds of the (post-ReLU) feature columns: fp16-split layer 0
        src.feats_colscale = feats_colscale.data_ptr()
    if moments is not None:   # [rows,9] fp64 partial sums from the fused sampling kernel (D = 0 layers)
        src.xyz_moments = moments.data_ptr()
        src.xyz_moment_rows = moments.shape[0]
    src.grouped = None
    src.xyz = xyz.data_ptr()
    src.new_xyz = new_xyz.data_ptr() if new_xyz is not None else None
    src.feats = feats.data_ptr() if feats is not None else None
    src.idx = idx32.data_ptr() if idx32 is not None else None
    src.B, src.N, src.S, src.K = B, N, S, K
    src.D = 0 if feats is None else feats.shape[2]
    src.order = order
    return src


def grouped_mlp(new_points, convs, bns, bn_mode="batch"):
    """The MLP + max-pool tail of a SetAbstraction layer on an explicit grouped tensor
    ``new_points [B,S,K,Cin]`` (what ``sample_and_group`` returns) -> ``[B,Cout,S]``;
    layers.py:214-219."""
    L.require_cuda(new_points)
    new_points = L.f32c(new_points)
    B, S, K, Cin = new_points.shape
    src = L.GroupSource()
    src.grouped = new_points.data_ptr()
    src.B, src.N, src.S, src.K, src.D, src.order = B, K, S, K, Cin - 3, L.XYZ_FIRST
    out = _MlpRunner(convs, bns).run(src, new_points, Cin, B, S, bn_mode, new_points.device)
    return out.transpose(1, 2)


class _SAMixin(torch.nn.Module):
    """Shared plumbing.  The conv/bn holders live in plain Python lists exactly like the reference
    (layers.py:185-190, 230-241), so they are invisible to ``parameters()`` / ``state_dict()``;
    ``_apply`` is overridden so ``.to()`` / ``.cuda()`` still move them."""

    bn_mode = "batch"
    update_running_stats = False
    # SyncBN over the batch shards: ``sync_bn = True`` all-reduces the per-layer statistics over
    # ``sync_bn_group`` (None = the default / world group).  papc_b200.dist.set_sync_bn sets both.
    sync_bn = False
    sync_bn_group = None

    def _sync_group(self):
        """(enabled, group) -- enabled only when torch.distributed runs with more than one rank."""
        if not (self.sync_bn or self.sync_bn_group is not None):
            return False, None
        if not dist.is_initialized():
            return False, None
        group = self.sync_bn_group if self.sync_bn_group is not None else dist.group.WORLD
        return dist.get_world_size(group) > 1, group

    def _holders(self):
        raise NotImplementedError

    def _apply(self, fn, *args, **kwargs):
        super()._apply(fn, *args, **kwargs)
        for h in self._holders():
            h._apply(fn)
        return self


class PointNetSetAbstraction(_SAMixin):
    """layers.py:179-221 (SSG / group_all)."""

    def __init__(self, npoint, radius, nsample, in_channel, mlp, group_all):
        super().__init__()
        self.npoint = npoint
        self.radius = radius
        self.nsample = nsample
        self.mlp_convs = []
        self.mlp_bns = []
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(Conv2D(last_channel, out_channel, 1))
            self.mlp_bns.append(BatchNorm2D(out_channel))
            last_channel = out_channel
        self.group_all = group_all
        self.in_channel = in_channel

    def _holders(self):
        return list(self.mlp_convs) + list(self.mlp_bns)

    def forward(self, xyz, points, start_idx=None):
        """xyz [B,3,N], points [B,D,N] | None -> (new_xyz [B,3,S], new_points [B,D',S])."""
        L.require_cuda(xyz, points)
        ready = _ready_event(xyz)
        entry_ev = _entry_event(xyz.device) if not self.group_all else None
        fcs = _colscale_of(points, points.shape[1]) if points is not None else None
        xyz = L.f32c(xyz.transpose(1, 2))                                  # :203
        feats = L.f32c(points.transpose(1, 2)) if points is not None else None  # :205
        B, N, Cc = xyz.shape
        if Cc != 3:
            raise ValueError("xyz must be [B,3,N]")
        D = 0 if feats is None else feats.shape[2]
        if 3 + D != self.in_channel:
            raise ValueError(f"in_channel={self.in_channel} but input has 3+{D} channels")
        dev = xyz.device
        if self.group_all:                                                 # :211
            new_xyz = torch.zeros((B, 1, 3), dtype=torch.float32, device=dev)
            S = 1
            src = _make_src(xyz, None, feats, None, B, N, 1, N, L.XYZ_FIRST, feats_colscale=fcs)
            keep = (xyz, feats, fcs)
        else:                                                              # :213
            S = self.npoint
            if self.nsample > N:
                raise ValueError(f"query_ball_point: nsample ({self.nsample}) > N ({N})")
            new_xyz, (idx,), ev, mom = _sample(xyz, ready, S, start_idx, [(self.radius, self.nsample)],
                                               want_moments=True)
            if feats is not None or self.bn_mode != "batch":
                mom = None        # only the folded first layer (features = centred xyz, batch statistics) reads them
            src = _make_src(xyz, new_xyz, feats, idx, B, N, S, self.nsample, L.XYZ_FIRST, moments=mom, feats_colscale=fcs)
            keep = (xyz, feats, new_xyz, idx, mom, fcs)
        out = _MlpRunner(self.mlp_convs, self.mlp_bns).run(                # :214-219
            src, keep, 3 + D, B, S, self.bn_mode, dev, self.update_running_stats, self._sync_group())
        out_xyz = new_xyz.transpose(1, 2)
        if not self.group_all:
            _tag_ready(out_xyz, ev, entry_ev)
        out_feats = out.transpose(1, 2)
        if self.bn_mode == "batch" and not self._sync_group()[0]:
            # post-ReLU outputs of a batch-statistics BatchNorm over B*S*K rows: bounded, and the next layer's
            # gather kernel can use the fp16 operand split on them
            _tag_colscale(out_feats, _feature_colscale(self.mlp_bns[-1], B * S * (N if self.group_all else self.nsample)))
        return out_xyz, out_feats                                          # :220-221


class PointNetSetAbstractionMsg(_SAMixin):
    """layers.py:224-281 (multi-scale grouping; features-first concat, :267)."""

    def __init__(self, npoint, radius_list, nsample_list, in_channel, mlp_list):
        super().__init__()
        self.npoint = npoint
        self.radius_list = radius_list
        self.nsample_list = nsample_list
        self.conv_blocks = []
        self.bn_blocks = []
        for i in range(len(mlp_list)):
            convs, bns = [], []
            last_channel = in_channel + 3
            for out_channel in mlp_list[i]:
                convs.append(Conv2D(last_channel, out_channel, 1))
                bns.append(BatchNorm2D(out_channel))
                last_channel = out_channel
            self.conv_blocks.append(convs)
            self.bn_blocks.append(bns)
        self.in_channel = in_channel

    def _holders(self):
        return [h for blk in self.conv_blocks for h in blk] + [h for blk in self.bn_blocks for h in blk]

    def forward(self, xyz, points, start_idx=None):
        L.require_cuda(xyz, points)
        ready = _ready_event(xyz)
        entry_ev = _entry_event(xyz.device)
        xyz = L.f32c(xyz.transpose(1, 2))
        feats = L.f32c(points.transpose(1, 2)) if points is not None else None
        B, N, Cc = xyz.shape
        D = 0 if feats is None else feats.shape[2]
        if D != self.in_channel:
            raise ValueError(f"in_channel={self.in_channel} but points has {D} channels")
        dev = xyz.device
        S = self.npoint
        if max(self.nsample_list) > N:
            raise ValueError(f"query_ball_point: nsample ({max(self.nsample_list)}) > N ({N})")
        new_xyz, idxs, ev = _sample(xyz, ready, S, start_idx,                        # :258, :262
                                    list(zip(self.radius_list, self.nsample_list)))
        outs = []
        for i, radius in enumerate(self.radius_list):
            K = self.nsample_list[i]
            idx = idxs[i]
            src = _make_src(xyz, new_xyz, feats, idx, B, N, S, K, L.FEATS_FIRST)     # :263-267
            outs.append(_MlpRunner(self.conv_blocks[i], self.bn_blocks[i]).run(     # :271-276
                src, (xyz, feats, new_xyz, idx), 3 + D, B, S, self.bn_mode, dev,
                self.update_running_stats, self._sync_group()))
        new_points_concat = torch.cat(outs, dim=2)                                   # :280 (channels-last)
        return _tag_ready(new_xyz.transpose(1, 2), ev, entry_ev), new_points_concat.transpose(1, 2)


# ---------------------------------------------------------------------------------------------
class Conv1D(Conv2D):
    """Parameter holder mirroring ``paddle.nn.Conv1D(cin, cout, 1)``: weight [cout,cin,1], bias [cout]."""

    def __init__(self, in_channels, out_channels, kernel_size=1, device=None, generator=None):
        super().__init__(in_channels, out_channels, kernel_size, device, generator)
        self.weight = self.weight.reshape(out_channels, in_channels, 1)


class BatchNorm1D(BatchNorm2D):
    """Parameter holder mirroring ``paddle.nn.BatchNorm1D(c)`` (epsilon 1e-5, momentum 0.9)."""


class Conv1DLayer(torch.nn.Module):
    """REGISTERED form of ``Conv1D`` (weight / bias are parameters, so they appear in ``parameters()`` and
    ``state_dict()`` under Paddle's names) for sublayers the reference assigns as attributes, e.g. the
    segmentation head's ``conv1`` / ``conv2`` (segment/pointnet2/pointnet2.py:21-24)."""

    def __init__(self, in_channels, out_channels, kernel_size=1):
        super().__init__()
        h = Conv1D(in_channels, out_channels, kernel_size)
        self.weight = torch.nn.Parameter(h.weight, requires_grad=False)
        self.bias = torch.nn.Parameter(h.bias, requires_grad=False)


class BatchNorm1DLayer(torch.nn.Module):
    """REGISTERED form of ``BatchNorm1D``: weight / bias parameters, ``_mean`` / ``_variance`` buffers
    (Paddle's state_dict keys), so running statistics can be saved, loaded and mapped from a reference
    checkpoint."""

    def __init__(self, num_features, momentum=0.9, epsilon=1e-5):
        super().__init__()
        self.weight = torch.nn.Parameter(torch.ones(num_features), requires_grad=False)
        self.bias = torch.nn.Parameter(torch.zeros(num_features), requires_grad=False)
        self.register_buffer("_mean", torch.zeros(num_features))
        self.register_buffer("_variance", torch.ones(num_features))
        self._momentum = momentum
        self._epsilon = epsilon


def feature_interpolate(xyz1, xyz2, points1, points2, pad_to=8):
    """layers.py:306-329 on channels-last inputs: xyz1 [B,N,3], xyz2 [B,S,3], points1 [B,N,D1] | None,
    points2 [B,S,D2] -> rows [B*N, ld] = [points1 | interpolated | 0-padding], ld = D1+D2 rounded up
    to ``pad_to`` (what the tensor-core MLP wants).  Returns (rows, D1 + D2)."""
    L.require_cuda(xyz1, xyz2, points1, points2)
    xyz1, xyz2, points2 = L.f32c(xyz1), L.f32c(xyz2), L.f32c(points2)
    points1 = L.f32c(points1) if points1 is not None else None
    B, N, _ = xyz1.shape
    S = xyz2.shape[1]
    D1 = 0 if points1 is None else points1.shape[2]
    D2 = points2.shape[2]
    if xyz2.shape[0] != B or points2.shape[:2] != (B, S) or (points1 is not None and points1.shape[:2] != (B, N)):
        raise ValueError("feature_interpolate: inconsistent shapes")
    ld = (D1 + D2 + pad_to - 1) // pad_to * pad_to
    out = torch.empty((B * N, ld), dtype=torch.float32, device=xyz1.device)
    L.check(L.lib().papc_fp_interpolate_f32(L.ptr(xyz1), L.ptr(xyz2), L.ptr(points1), L.ptr(points2), B, N, S,
                                            D1, D2, ld, L.ptr(out), L.stream_ptr(xyz1.device)), "fp_interpolate")
    return out, D1 + D2


def _pointwise_mlp_rows_synced(rows, cin, convs, bns, group, update_running):
    """Batch-sharded form of ``pointwise_mlp_rows`` with BatchNorm statistics summed over the ranks of
    ``group``: step-wise C ABI, one all-reduce of [2,cout] fp64 sums per layer."""
    lib = L.lib()
    dev = rows.device
    st = L.stream_ptr(dev)
    M, ld = rows.shape
    total = float(M) * dist.get_world_size(group)   # equal shards (papc_b200.dist.shard_range)
    prows = lib.papc_mlp_stats_partial_rows(M)
    x, c_in = rows, ld
    scale = shift = None
    stats = []
    for l, (conv, bn) in enumerate(zip(convs, bns)):
        w = L.f32c(conv.weight.reshape(conv.weight.shape[0], -1))
        if l == 0 and ld != cin:   # the rows are zero-padded to ld columns: pad the first weight alike
            w = torch.nn.functional.pad(w, (0, ld - cin)).contiguous()
        cout = w.shape[0]
        y = torch.empty((M, cout), dtype=torch.float32, device=dev)
        partial = torch.empty((prows, 2, cout), dtype=torch.float64, device=dev)
        bias = L.f32c(conv.bias) if conv.bias is not None else None
        lwsb = lib.papc_mlp_layer_workspace_bytes(c_in, cout)
        lws = _ws(lwsb, dev)
        L.check(lib.papc_mlp_layer_forward_f32(None, L.ptr(x), L.ptr(scale), L.ptr(shift), M, c_in, cout, 1,
                                               L.ptr(w), L.ptr(bias), L.ptr(y), None, None, L.ptr(partial),
                                               L.ptr(lws), lwsb, st), "mlp_layer_forward")
        sums = torch.empty((2, cout), dtype=torch.float64, device=dev)
        L.check(lib.papc_mlp_stats_reduce_f64(L.ptr(partial), prows, cout, L.ptr(sums), st), "mlp_stats_reduce")
        dist.all_reduce(sums, group=group)
        scale = torch.empty((cout,), dtype=torch.float32, device=dev)
        shift = torch.empty((cout,), dtype=torch.float32, device=dev)
        bm = torch.empty((cout,), dtype=torch.float32, device=dev)
        bv = torch.empty((cout,), dtype=torch.float32, device=dev)
        L.check(lib.papc_bn_scale_shift_f32(L.ptr(sums), total, L.ptr(L.f32c(bn.weight)), L.ptr(L.f32c(bn.bias)),
                                            float(bn._epsilon), cout, L.ptr(scale), L.ptr(shift), L.ptr(bm),
                                            L.ptr(bv), st), "bn_scale_shift")
        stats.append((bm, bv))
        x, c_in = y, cout
    out = torch.empty_like(x)
    L.check(lib.papc_bn_relu_apply_f32(L.ptr(x), L.ptr(scale), L.ptr(shift), M, c_in, L.ptr(out), st), "bn_relu_apply")
    if update_running:
        _MlpRunner(convs, bns)._update_running(stats)
    return out


def pointwise_mlp_rows(rows, cin, convs, bns, bn_mode="batch", update_running=False, sync=(False, None)):
    """(1x1 conv -> BatchNorm -> ReLU) x len(convs) over channels-last rows [M, ld] (ld >= cin, ld % 4 == 0,
    columns >= cin ignored) -> [M, cout], on the tcgen05 layer kernels (``papc_pointwise_mlp_f32``).
    ``bn_mode`` 'batch' normalises with the statistics of the M rows, 'running' with the holders'
    ``_mean`` / ``_variance``; ``update_running`` folds the batch statistics into them (Paddle momentum)."""
    L.require_cuda(rows)
    if bn_mode not in ("batch", "running"):
        raise ValueError("bn_mode must be 'batch' or 'running'")
    rows = L.f32c(rows)
    dev = rows.device
    if any(c.weight.device != dev for c in convs):
        raise L.PapcError("layer parameters are not on the input's device; call .to(device)")
    if bn_mode == "batch" and sync[0]:
        return _pointwise_mlp_rows_synced(rows, cin, convs, bns, sync[1], update_running)
    runner = _MlpRunner(convs, bns)
    mlp, keep, stats = runner._mlp_struct(cin, bn_mode, dev, update_running)
    lib = L.lib()
    M, ld = rows.shape
    cout = convs[-1].weight.shape[0]
    out = torch.empty((M, cout), dtype=torch.float32, device=dev)
    wsb = lib.papc_pointwise_mlp_workspace_bytes(M, ld, C.byref(mlp))
    ws = _ws(wsb, dev)
    L.check(lib.papc_pointwise_mlp_f32(L.ptr(rows), M, ld, C.byref(mlp), L.ptr(out), L.ptr(ws), wsb,
                                       L.stream_ptr(dev)), "pointwise_mlp")
    if stats:
        runner._update_running(stats)
    del keep
    return out


class PointNetFeaturePropagation(_SAMixin):
    """layers.py:284-335.  The conv/bn holders live in plain lists like the
    reference's (:287-294), so -- as for the SetAbstraction layers -- BatchNorm runs on batch
    statistics unless ``bn_mode='running'``.  The reference's interpolation quirk (weights of the
    three nearest sampled points applied to the features of sampled points 0, 1, 2, :317-318) is
    reproduced; see include/papc_b200.h."""

    def __init__(self, in_channel, mlp):
        super().__init__()
        self.mlp_convs = []
        self.mlp_bns = []
        last_channel = in_channel
        for out_channel in mlp:
            self.mlp_convs.append(Conv1D(last_channel, out_channel, 1))
            self.mlp_bns.append(BatchNorm1D(out_channel))
            last_channel = out_channel
        self.in_channel = in_channel

    def _holders(self):
        return list(self.mlp_convs) + list(self.mlp_bns)

    def forward(self, xyz1, xyz2, points1, points2):
        """xyz1 [B,C,N], xyz2 [B,C,S], points1 [B,D,N] | None, points2 [B,D,S] -> [B,D',N]."""
        L.require_cuda(xyz1, xyz2, points1, points2)
        x1 = L.f32c(xyz1.transpose(1, 2))                                   # :306
        x2 = L.f32c(xyz2.transpose(1, 2))                                   # :307
        p2 = L.f32c(points2.transpose(1, 2))                                # :309
        p1 = L.f32c(points1.transpose(1, 2)) if points1 is not None else None  # :327
        B, N, _ = x1.shape
        rows, cin = feature_interpolate(x1, x2, p1, p2)                     # :311-329
        if cin != self.in_channel:
            raise ValueError(f"in_channel={self.in_channel} but the concatenated input has {cin} channels")
        out = pointwise_mlp_rows(rows, cin, self.mlp_convs, self.mlp_bns, self.bn_mode,   # :332-335
                                 update_running=self.update_running_stats, sync=self._sync_group())
        return out.reshape(B, N, -1).transpose(1, 2)
