"""GPU parity tests of the SetAbstraction path: CUDA (through the C ABI) vs the CPU oracle on the
same seeded inputs.  Bit-exact for indices / gathers; fp32 features within 1e-5 (abs + rel)."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import capi, layers_np  # noqa: E402
from papc_b200 import layers, synth  # noqa: E402

DEV = "cuda:0"
TOL = dict(rtol=1e-5, atol=1e-5)  # the north-star tolerance for fp32 MLP / pool outputs


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _xyz(B, N, seed=0):
    return np.ascontiguousarray(synth.clouds(B, N, seed).transpose(0, 2, 1))


# ------------------------------------------------------------------ KATs through the C ABI
def _line(xs):
    return np.array([[[x, 0.0, 0.0] for x in xs]], np.float32)


def test_kat_k1_k2_fps():
    out = layers.farthest_point_sample(_cu(_line([0, .1, .5, .9])), 3, start_idx=[0])
    assert out.dtype == torch.float32 and out.cpu().tolist() == [[0.0, 3.0, 2.0]]
    out = layers.farthest_point_sample_idx(_cu(_line([0, 2, 3])), 3, start_idx=[0])
    assert out.cpu().tolist() == [[0, 1, 2]]  # clamp quirk (distance initialised to 1.0)


def test_kat_k3_k4_ball_query():
    xyz = _line([0, .1, .15, .3, .19, .7])
    q = xyz[:, :1, :]
    assert layers.query_ball_point(0.2, 3, _cu(xyz), _cu(q)).cpu().tolist() == [[[0, 1, 2]]]
    assert layers.query_ball_point(0.2, 6, _cu(xyz), _cu(q)).cpu().tolist() == [[[0, 1, 2, 4, 0, 0]]]
    xyz4 = _line([0, .1, .15, .3, .2, .7])
    assert layers.query_ball_point(0.2, 6, _cu(xyz4), _cu(q)).cpu().tolist() == [[[0, 1, 2, 0, 0, 0]]]


# ------------------------------------------------------------------ primitives vs oracle
@pytest.mark.parametrize("B,N,npoint", [(32, 1024, 512), (32, 512, 128), (16, 2048, 512), (3, 100, 100),
                                        (2, 33, 7), (2, 5000, 64), (1, 9000, 16), (5, 1, 1)])
def test_fps_bit_exact(B, N, npoint):
    xyz = _xyz(B, N, seed=B + N)
    for start in (synth.fps_start(B, N, seed=1), np.zeros(B, np.int64)):
        ref = capi.farthest_point_sample(xyz, npoint, start)
        idx, new_xyz = layers.farthest_point_sample_idx(_cu(xyz), npoint, _cu(start), return_xyz=True)
        np.testing.assert_array_equal(idx.cpu().numpy(), ref)
        np.testing.assert_array_equal(new_xyz.cpu().numpy(), layers_np.index_points(xyz, ref))


def test_fps_duplicates_and_ties():
    # duplicated points and an all-equal cloud: ties must resolve to the lowest index
    xyz = _xyz(2, 64, seed=9)
    xyz[:, 32:] = xyz[:, :32]
    start = np.array([5, 40], np.int64)
    np.testing.assert_array_equal(layers.farthest_point_sample_idx(_cu(xyz), 64, _cu(start)).cpu().numpy(),
                                  capi.farthest_point_sample(xyz, 64, start))
    same = np.zeros((1, 16, 3), np.float32)
    np.testing.assert_array_equal(layers.farthest_point_sample_idx(_cu(same), 4, [3]).cpu().numpy(),
                                  capi.farthest_point_sample(same, 4, [3]))


@pytest.mark.parametrize("B,N,S,r,K", [(32, 1024, 512, 0.2, 32), (32, 512, 128, 0.4, 64),
                                       (16, 2048, 512, 0.1, 32), (16, 2048, 512, 0.4, 128),
                                       (2, 5000, 77, 0.3, 48), (3, 37, 37, 0.5, 37), (2, 64, 9, 10.0, 64)])
def test_ball_query_bit_exact(B, N, S, r, K):
    xyz = _xyz(B, N, seed=N + S)
    fps = capi.farthest_point_sample(xyz, S, synth.fps_start(B, N))
    new_xyz = layers_np.index_points(xyz, fps)
    ref, empty = capi.query_ball_point(r, K, xyz, new_xyz)
    assert empty == 0
    out = layers.query_ball_point(r, K, _cu(xyz), _cu(new_xyz), check_empty=True)
    assert out.dtype == torch.int64
    np.testing.assert_array_equal(out.cpu().numpy(), ref)
    out32 = layers._ball_query(r, K, _cu(xyz), _cu(new_xyz), torch.int32)
    np.testing.assert_array_equal(out32.cpu().numpy().astype(np.int64), ref)


def test_ball_query_errors():
    xyz = _xyz(1, 16)
    with pytest.raises(ValueError):
        layers.query_ball_point(0.2, 17, _cu(xyz), _cu(xyz[:, :2]))
    far = np.full((1, 1, 3), 50.0, np.float32)
    with pytest.raises(IndexError):
        layers.query_ball_point(0.2, 4, _cu(xyz), _cu(far), check_empty=True)
    out = layers.query_ball_point(0.2, 4, _cu(xyz), _cu(far))
    assert (out.cpu().numpy() == 16).all()  # N in every slot, as the reference's sort leaves it


def test_square_distance_and_index_points():
    rng = np.random.default_rng(0)
    a = rng.uniform(-1, 1, (3, 50, 3)).astype(np.float32)
    b = rng.uniform(-1, 1, (3, 70, 3)).astype(np.float32)
    np.testing.assert_array_equal(layers.square_distance(_cu(a), _cu(b)).cpu().numpy(),
                                  capi.square_distance(a, b))
    pts = rng.standard_normal((3, 50, 7)).astype(np.float32)
    idx2 = rng.integers(0, 50, (3, 11))
    idx3 = rng.integers(0, 50, (3, 11, 5))
    np.testing.assert_array_equal(layers.index_points(_cu(pts), _cu(idx2)).cpu().numpy(),
                                  layers_np.index_points(pts, idx2))
    np.testing.assert_array_equal(layers.index_points(_cu(pts), _cu(idx3.astype(np.float32))).cpu().numpy(),
                                  layers_np.index_points(pts, idx3))
    pts8 = rng.standard_normal((3, 50, 8)).astype(np.float32)  # float4 path
    np.testing.assert_array_equal(layers.index_points(_cu(pts8), _cu(idx3)).cpu().numpy(),
                                  layers_np.index_points(pts8, idx3))


@pytest.mark.parametrize("D", [0, 5, 8])
def test_sample_and_group_bit_exact(D):
    B, N, S, K = 3, 300, 40, 12
    xyz = _xyz(B, N, seed=4)
    feats = np.random.default_rng(5).standard_normal((B, N, D)).astype(np.float32) if D else None
    start = synth.fps_start(B, N, seed=6)
    rx, rp, rg, rf = layers_np.sample_and_group(S, 0.35, K, xyz, feats, returnfps=True, start_idx=start)
    gx, gp, gg, gf = layers.sample_and_group(S, 0.35, K, _cu(xyz), _cu(feats) if D else None,
                                             returnfps=True, start_idx=_cu(start))
    np.testing.assert_array_equal(gx.cpu().numpy(), rx)
    np.testing.assert_array_equal(gf.cpu().numpy(), rf)
    np.testing.assert_array_equal(gg.cpu().numpy(), rg)
    np.testing.assert_array_equal(gp.cpu().numpy(), rp)
    ax, ap = layers.sample_and_group_all(_cu(xyz), _cu(feats) if D else None)
    ox, op = layers_np.sample_and_group_all(xyz, feats)
    np.testing.assert_array_equal(ax.cpu().numpy(), ox)
    np.testing.assert_array_equal(ap.cpu().numpy(), op)


# ------------------------------------------------------------------ grouped MLP vs oracle
def _set_params(gpu_convs, gpu_bns, ref_convs, ref_bns, params, rng):
    for l, p in enumerate(params):
        w4 = p["weight"].reshape(*p["weight"].shape, 1, 1)
        gamma = rng.uniform(0.5, 1.5, p["gamma"].shape).astype(np.float32)
        gamma[::7] *= -1  # negative BN scales exercise the min-pool branch
        beta = rng.uniform(-0.2, 0.2, p["beta"].shape).astype(np.float32)
        gpu_convs[l].weight, gpu_convs[l].bias = _cu(w4), _cu(p["bias"])
        gpu_bns[l].weight, gpu_bns[l].bias = _cu(gamma), _cu(beta)
        ref_convs[l].weight, ref_convs[l].bias = w4, p["bias"]
        ref_bns[l].weight, ref_bns[l].bias = gamma, beta


@pytest.mark.parametrize("B,S,K,cin,mlp", [(2, 16, 32, 3, [64, 64, 128]), (2, 8, 64, 131, [128, 128, 256]),
                                           # 20 480 rows: paired channel tiles with a ragged second tile (cout = 192)
                                           (4, 80, 64, 20, [64, 192]), (4, 80, 64, 20, [128, 256, 64]),
                                           (3, 1, 128, 259, [256, 512, 1024]), (2, 5, 12, 10, [20, 36]),
                                           (1, 3, 7, 6, [9]), (2, 4, 256, 19, [32, 48])])
def test_grouped_mlp_explicit_input(B, S, K, cin, mlp):
    rng = np.random.default_rng(cin)
    x = rng.standard_normal((B, S, K, cin)).astype(np.float32)
    params = synth.mlp_params(cin, mlp, seed=cin)
    convs = [layers.Conv2D(1, 1) for _ in mlp]
    bns = [layers.BatchNorm2D(c) for c in mlp]
    rconvs = [layers_np.Conv2D1x1(1, 1) for _ in mlp]
    rbns = [layers_np.BatchNorm2D(c) for c in mlp]
    _set_params(convs, bns, rconvs, rbns, params, rng)
    ref, stats = layers_np.grouped_mlp(x, [c.weight.reshape(c.weight.shape[0], -1) for c in rconvs],
                                       [c.bias for c in rconvs], [b.weight for b in rbns],
                                       [b.bias for b in rbns])
    out = layers.grouped_mlp(_cu(x), convs, bns)
    assert tuple(out.shape) == (B, mlp[-1], S)
    np.testing.assert_allclose(out.cpu().numpy(), ref, **TOL)


@pytest.mark.parametrize("cfg", ["ssg_sa1", "ssg_sa2", "group_all", "msg"])
@pytest.mark.parametrize("bn_mode", ["batch", "running"])
def test_set_abstraction_layer(cfg, bn_mode):
    rng = np.random.default_rng(11)
    B, N = 4, 512
    xyz = synth.clouds(B, N, seed=3)
    start = synth.fps_start(B, N, seed=4)
    if cfg == "ssg_sa1":
        D, args = 0, (128, 0.2, 32, 3, [64, 64, 128], False)
    elif cfg == "ssg_sa2":
        D, args = 128, (64, 0.4, 64, 131, [128, 128, 256], False)
    elif cfg == "group_all":
        D, args = 61, (None, None, None, 64, [64, 128, 256], True)
    else:
        D, args = 6, None
    feats = rng.standard_normal((B, D, N)).astype(np.float32) if D else None
    if cfg != "msg":
        gpu = layers.PointNetSetAbstraction(*args)
        ref = layers_np.PointNetSetAbstraction(*args)
        _set_params(gpu.mlp_convs, gpu.mlp_bns, ref.mlp_convs, ref.mlp_bns,
                    synth.mlp_params(args[3], args[4], seed=5), rng)
        blocks = [(gpu.mlp_bns, ref.mlp_bns)]
    else:
        margs = (64, [0.2, 0.4, 0.8], [16, 32, 64], D, [[32, 32, 64], [64, 64, 128], [64, 96, 128]])
        gpu = layers.PointNetSetAbstractionMsg(*margs)
        ref = layers_np.PointNetSetAbstractionMsg(*margs)
        for i, m in enumerate(margs[4]):
            _set_params(gpu.conv_blocks[i], gpu.bn_blocks[i], ref.conv_blocks[i], ref.bn_blocks[i],
                        synth.mlp_params(D + 3, m, seed=6 + i), rng)
        blocks = list(zip(gpu.bn_blocks, ref.bn_blocks))
    if bn_mode == "running":
        gpu.bn_mode = "running"
        for gb, rb in blocks:
            for g, r in zip(gb, rb):
                m = rng.uniform(-0.3, 0.3, r._mean.shape).astype(np.float32)
                v = rng.uniform(0.5, 2.0, r._variance.shape).astype(np.float32)
                g._mean, g._variance = _cu(m), _cu(v)
                r._mean, r._variance, r.training = m, v, False
    gpu.to(DEV)
    gx, gp = gpu(_cu(xyz), _cu(feats) if D else None, start_idx=_cu(start))
    rx, rp = ref(xyz, feats, start_idx=start)
    assert tuple(gx.shape) == rx.shape and tuple(gp.shape) == rp.shape
    np.testing.assert_array_equal(gx.cpu().numpy(), rx)
    np.testing.assert_allclose(gp.cpu().numpy(), rp, **TOL)


def test_fp16_split_column_scale_large_gamma():
    """BatchNorm weights large enough that relu(bn(.)) exceeds the fp16 range: the exact
    power-of-two column scale of the fp16-split layers must keep the result within tolerance."""
    rng = np.random.default_rng(5)
    B, N = 4, 512
    xyz = synth.clouds(B, N, seed=8)
    start = synth.fps_start(B, N, seed=9)
    args = (128, 0.2, 32, 3, [64, 64, 128], False)
    gpu, ref = layers.PointNetSetAbstraction(*args), layers_np.PointNetSetAbstraction(*args)
    _set_params(gpu.mlp_convs, gpu.mlp_bns, ref.mlp_convs, ref.mlp_bns, synth.mlp_params(3, args[4], seed=5), rng)
    for l in range(2):  # inputs of layers 1 and 2 reach ~ 5000 * sqrt(16384) = 6.4e5 >> 65504
        g = (ref.mlp_bns[l].weight * 5000.0).astype(np.float32)
        ref.mlp_bns[l].weight = g
        gpu.mlp_bns[l].weight = _cu(g)
    gpu.to(DEV)
    gx, gp = gpu(_cu(xyz), None, start_idx=_cu(start))
    rx, rp = ref(xyz, None, start_idx=start)
    assert np.isfinite(gp.cpu().numpy()).all()
    np.testing.assert_array_equal(gx.cpu().numpy(), rx)
    np.testing.assert_allclose(gp.cpu().numpy(), rp, **TOL)


@pytest.mark.parametrize("cfg", ["ssg_sa1", "ssg_sa2"])
def test_statistics_outside_the_fixed_point_range(cfg):
    """A convolution bias large enough that the per-channel sums of squares exceed 2^53 (32 768 rows x (6e5)^2):
    the layer kernels raise the flag of their exact fixed-point statistic words and every consumer of a deferred
    BatchNorm finalisation (the next layer kernel, pool_finish) must fall back to the fp64 partial rows.  The
    pre-BN values sit at 6e5 +- 1e4, where fp32 resolves 0.06: both sides (the oracle computes the layer in fp32 as
    the reference does) lose digits in y - mean over three layers (measured: 5e-4 for the sa2 shape, 8e-3 for
    sa1, whose first layers spread the least).  This is a test of the fallback PATH -- a broken one gives zeros,
    NaNs or O(1) errors -- with a gross-error bound of 3e-2; the 1e-5 parity bound is tested everywhere else."""
    rng = np.random.default_rng(21)
    B, N = 8, 512
    xyz = synth.clouds(B, N, seed=13)
    start = synth.fps_start(B, N, seed=14)
    if cfg == "ssg_sa1":
        D, args = 0, (128, 0.2, 32, 3, [64, 64, 128], False)
    else:
        D, args = 128, (64, 0.4, 64, 131, [128, 128, 256], False)
    feats = rng.standard_normal((B, D, N)).astype(np.float32) if D else None
    gpu, ref = layers.PointNetSetAbstraction(*args), layers_np.PointNetSetAbstraction(*args)
    params = synth.mlp_params(args[3], args[4], seed=15)
    for p in params:
        p["weight"] = (p["weight"] * np.float32(1e4)).astype(np.float32)
        p["bias"] = np.full_like(p["bias"], 6e5)
    _set_params(gpu.mlp_convs, gpu.mlp_bns, ref.mlp_convs, ref.mlp_bns, params, rng)
    gpu.to(DEV)
    gx, gp = gpu(_cu(xyz), _cu(feats) if D else None, start_idx=_cu(start))
    rx, rp = ref(xyz, feats, start_idx=start)
    assert np.isfinite(gp.cpu().numpy()).all()
    np.testing.assert_array_equal(gx.cpu().numpy(), rx)
    np.testing.assert_allclose(gp.cpu().numpy(), rp, rtol=3e-2, atol=3e-2)


@pytest.mark.parametrize("cfg", ["ssg_sa1", "ssg_sa2", "group_all"])
def test_running_statistics_of_every_layer(cfg):
    """update_running_stats: the batch mean / variance of every MLP layer.  With the deferred BatchNorm
    finalisation they are written by whoever finalises the layer -- the next layer kernel's CTA 0, block 0 of the
    activation-image kernel (group_all: small M), block 0 of pool_finish (last layer), the moments kernel (folded
    first layer) -- so each writer is checked against the oracle's running buffers."""
    rng = np.random.default_rng(31)
    B, N = 4, 512
    xyz = synth.clouds(B, N, seed=23)
    start = synth.fps_start(B, N, seed=24)
    if cfg == "ssg_sa1":
        D, args = 0, (128, 0.2, 32, 3, [64, 64, 128], False)
    elif cfg == "ssg_sa2":
        D, args = 128, (64, 0.4, 64, 131, [128, 128, 256], False)
    else:
        D, args = 256, (None, None, None, 259, [256, 512, 1024], True)
    feats = rng.standard_normal((B, D, N)).astype(np.float32) if D else None
    gpu, ref = layers.PointNetSetAbstraction(*args), layers_np.PointNetSetAbstraction(*args)
    _set_params(gpu.mlp_convs, gpu.mlp_bns, ref.mlp_convs, ref.mlp_bns, synth.mlp_params(args[3], args[4], seed=25), rng)
    gpu.update_running_stats = True
    gpu.to(DEV)
    gx, gp = gpu(_cu(xyz), _cu(feats) if D else None, start_idx=_cu(start))
    rx, rp = ref(xyz, feats, start_idx=start)
    np.testing.assert_allclose(gp.cpu().numpy(), rp, **TOL)
    for l, (g, r) in enumerate(zip(gpu.mlp_bns, ref.mlp_bns)):
        np.testing.assert_allclose(g._mean.cpu().numpy(), r._mean, rtol=1e-5, atol=1e-6, err_msg=f"running mean, layer {l}")
        np.testing.assert_allclose(g._variance.cpu().numpy(), r._variance, rtol=1e-5, atol=1e-6,
                                   err_msg=f"running variance, layer {l}")


def test_ssg_stack_c2_full_size_vs_oracle():
    """BASELINE config 2 at full size (B=32, N=1024, the three SSG SetAbstraction layers of
    PointNet2_SSG_Clas, classify/pointnet2/pointnet2.py:11-16) against the oracle."""
    B, N = 32, 1024
    rng = np.random.default_rng(21)
    xyz = synth.clouds(B, N, seed=0)
    start = synth.fps_start(B, N, seed=1)
    cfgs = [(512, 0.2, 32, 3, [64, 64, 128], False), (128, 0.4, 64, 131, [128, 128, 256], False),
            (None, None, None, 259, [256, 512, 1024], True)]
    gpu_layers, ref_layers = [], []
    for i, c in enumerate(cfgs):
        g, r = layers.PointNetSetAbstraction(*c), layers_np.PointNetSetAbstraction(*c)
        _set_params(g.mlp_convs, g.mlp_bns, r.mlp_convs, r.mlp_bns, synth.mlp_params(c[3], c[4], seed=2 + i), rng)
        gpu_layers.append(g.to(DEV))
        ref_layers.append(r)
    starts = [start, np.zeros(B, np.int64), None]
    gx, gp = _cu(xyz), None
    rx, rp = xyz, None
    for g, r, st in zip(gpu_layers, ref_layers, starts):
        gx, gp = g(gx, gp, start_idx=_cu(st) if st is not None else None)
        rx, rp = r(rx, rp, start_idx=st)
        assert tuple(gx.shape) == rx.shape and tuple(gp.shape) == rp.shape
        np.testing.assert_array_equal(gx.cpu().numpy(), rx)
        np.testing.assert_allclose(gp.cpu().numpy(), rp, **TOL)
        # feed the ORACLE's features forward on both sides so each layer is judged on equal inputs
        gp = _cu(rp)
    assert tuple(gp.shape) == (B, 1024, 1)
    # determinism: same inputs -> bit-identical outputs
    a = gpu_layers[0](_cu(xyz), None, start_idx=_cu(start))[1]
    b = gpu_layers[0](_cu(xyz), None, start_idx=_cu(start))[1]
    assert torch.equal(a, b)


def test_sampling_overlap_is_transparent(monkeypatch):
    """The side-stream sampling overlap (layers._sample) must not change any result: the chained
    SSG stack with the overlap on == off, bit for bit, and the tagged xyz tensor really carries the
    event the next layer waits for."""
    from papc_b200 import sa_stack
    B, N = 8, 1024
    xyz = _cu(synth.clouds(B, N, seed=3))
    st1 = _cu(synth.fps_start(B, N, seed=4))
    st2 = torch.zeros(B, dtype=torch.int64, device=DEV)
    model = sa_stack.SSGSetAbstractionStack().to(DEV)
    for i, sa in enumerate(model.layers_()):
        c = [(3, [64, 64, 128]), (131, [128, 128, 256]), (259, [256, 512, 1024])][i]
        sa_stack.load_conv_bn(sa.mlp_convs, sa.mlp_bns, synth.mlp_params(c[0], c[1], seed=10 + i))
    outs = {}
    for flag in (True, False):
        monkeypatch.setattr(layers, "OVERLAP_SAMPLING", flag)
        l1_xyz, _ = model.sa1(xyz, None, start_idx=st1)
        assert hasattr(l1_xyz, "_papc_ready") and layers._ready_event(l1_xyz) is not None
        l1_xyz.add_(0.0)   # an in-place edit by the caller invalidates the tag (no stale overlap)
        assert layers._ready_event(l1_xyz) is None
        for _ in range(3):  # repeated calls exercise the side stream's buffer reuse
            l3_xyz, l3 = model(xyz, None, start_idx=(st1, st2))
        torch.cuda.synchronize()
        outs[flag] = l3.clone()
    assert torch.equal(outs[True], outs[False])


def test_sampling_overlap_three_sampled_layers(monkeypatch):
    """Three consecutive SAMPLED layers (the shipped models have two) plus a caller-produced start_idx: the
    side stream's buffers must not be recycled while an earlier layer's MLP still reads them.
    Overlap on == off, bit for bit, over repeated calls."""
    B, N = 4, 1024
    xyz = _cu(synth.clouds(B, N, seed=11))
    cfgs = [(512, 0.2, 32, 3, [32, 32, 64]), (256, 0.3, 32, 64 + 3, [64, 64, 64]), (64, 0.5, 32, 64 + 3, [64, 64, 128])]
    sas = []
    for i, c in enumerate(cfgs):
        sa = layers.PointNetSetAbstraction(c[0], c[1], c[2], c[3], c[4], False).to(DEV)
        from papc_b200 import sa_stack
        sa_stack.load_conv_bn(sa.mlp_convs, sa.mlp_bns, synth.mlp_params(c[3], c[4], seed=20 + i))
        sas.append(sa)
    outs = {}
    for flag in (True, False):
        monkeypatch.setattr(layers, "OVERLAP_SAMPLING", flag)
        for rep in range(4):
            st = (torch.arange(B, device=DEV) * 7 + rep * 0) % N       # produced on the main stream right here
            x, p = xyz, None
            for sa in sas:
                x, p = sa(x, p, start_idx=st if sa is sas[0] else torch.zeros(B, dtype=torch.int64, device=DEV))
        torch.cuda.synchronize()
        outs[flag] = p.clone()
    assert torch.equal(outs[True], outs[False])


def test_launch_profiler_records_kernels():
    """papc_prof_*: every instrumented launch yields one record with a positive duration and the
    algorithmic work of the launch (what bench.py's roofline block is built from)."""
    from papc_b200 import _lib as L
    lib = L.lib()
    xyz = _cu(_xyz(4, 256))
    start = torch.zeros(4, dtype=torch.int64, device=DEV)
    L.check(lib.papc_prof_reset(), "reset")
    L.check(lib.papc_prof_enable(1), "enable")
    _, new_xyz = layers.farthest_point_sample_idx(xyz, 64, start, return_xyz=True)
    layers.query_ball_point(0.3, 16, xyz, new_xyz)
    torch.cuda.synchronize()
    L.check(lib.papc_prof_enable(0), "disable")
    recs = L.prof_records()
    L.check(lib.papc_prof_reset(), "reset")
    names = [r["name"] for r in recs]
    assert names == ["fps_reg", "ball_query"]
    assert all(r["ms"] > 0 and r["bytes"] > 0 for r in recs)
    assert recs[0]["M"] == 4 * 256 and recs[0]["cin"] == 64
    assert lib.papc_prof_count() == 0


def test_graphed_forward_matches_eager():
    """sa_stack.GraphedForward: the captured graph (side-stream sampling and dependent launches
    included) replays to exactly the eager result, also after the input buffer is refilled."""
    from papc_b200 import sa_stack
    B, N = 4, 1024
    model = sa_stack.SSGSetAbstractionStack().to(DEV)
    for i, sa in enumerate(model.layers_()):
        c = [(3, [64, 64, 128]), (131, [128, 128, 256]), (259, [256, 512, 1024])][i]
        sa_stack.load_conv_bn(sa.mlp_convs, sa.mlp_bns, synth.mlp_params(c[0], c[1], seed=20 + i))
    st1 = _cu(synth.fps_start(B, N, seed=5))
    st2 = torch.zeros(B, dtype=torch.int64, device=DEV)
    xa, xb = _cu(synth.clouds(B, N, seed=6)), _cu(synth.clouds(B, N, seed=7))
    fn = lambda x: model(x, None, start_idx=(st1, st2))  # noqa: E731
    g = sa_stack.GraphedForward(fn, xa)
    assert g.kernels_per_replay > 10
    for x in (xa, xb, xa):
        want = fn(x)[1].clone()
        got = g(x)[1]
        torch.cuda.synchronize()
        assert torch.equal(got, want)


def test_msg_sa1_c3_shapes_vs_oracle():
    """BASELINE config 3 (PointNet2_MSG_Seg sa1, segment/pointnet2/pointnet2.py:62: 2048 points,
    512 centroids, radii .1/.2/.4 with 32/64/128 samples, features = xyz) at the real layer sizes
    with a reduced batch, against the oracle: new_xyz bit-exact, [B,320,512] features within 1e-5."""
    rng = np.random.default_rng(31)
    B, N = 2, 2048
    xyz = synth.clouds(B, N, seed=12)
    start = synth.fps_start(B, N, seed=13)
    margs = (512, [0.1, 0.2, 0.4], [32, 64, 128], 3, [[32, 32, 64], [64, 64, 128], [64, 96, 128]])
    gpu = layers.PointNetSetAbstractionMsg(*margs)
    ref = layers_np.PointNetSetAbstractionMsg(*margs)
    for i, m in enumerate(margs[4]):
        _set_params(gpu.conv_blocks[i], gpu.bn_blocks[i], ref.conv_blocks[i], ref.bn_blocks[i],
                    synth.mlp_params(6, m, seed=14 + i), rng)
    gpu.to(DEV)
    gx, gp = gpu(_cu(xyz), _cu(xyz), start_idx=_cu(start))
    rx, rp = ref(xyz, xyz, start_idx=start)
    assert tuple(gp.shape) == rp.shape == (B, 320, 512)
    np.testing.assert_array_equal(gx.cpu().numpy(), rx)
    np.testing.assert_allclose(gp.cpu().numpy(), rp, **TOL)


# ------------------------------------------------------------------ multi-radius ball query, kNN
@pytest.mark.parametrize("B,N,S,radii,ks", [(16, 2048, 512, (0.1, 0.2, 0.4), (32, 64, 128)),   # BASELINE config 3, sa1
                                            (16, 512, 128, (0.4, 0.8), (64, 128)),               # config 3, sa2
                                            (3, 5000, 70, (0.05, 0.3, 0.5, 2.0), (8, 40, 17, 64)),
                                            (2, 33, 9, (0.2, 10.0), (33, 5))])
def test_ball_query_multi_equals_separate_queries(B, N, S, radii, ks):
    """One distance pass, R index lists (papc_ball_query_multi_f32) == R calls of query_ball_point, bit for bit,
    and both == the oracle (layers.py:258-267)."""
    xyz = _xyz(B, N, seed=N + S)
    fps = capi.farthest_point_sample(xyz, S, synth.fps_start(B, N))
    new_xyz = layers_np.index_points(xyz, fps)
    outs = layers.query_ball_point_multi(radii, ks, _cu(xyz), _cu(new_xyz))
    assert len(outs) == len(radii)
    for r, k, o in zip(radii, ks, outs):
        assert o.dtype == torch.int64 and tuple(o.shape) == (B, S, k)
        ref, _ = capi.query_ball_point(r, k, xyz, new_xyz)
        np.testing.assert_array_equal(o.cpu().numpy(), ref)
        np.testing.assert_array_equal(layers.query_ball_point(r, k, _cu(xyz), _cu(new_xyz)).cpu().numpy(), ref)


@pytest.mark.parametrize("B,N,S,k", [(16, 128, 512, 3), (4, 2048, 100, 16), (2, 5000, 33, 32), (3, 7, 5, 3), (2, 2, 4, 3)])
def test_knn_vs_stable_argsort_of_square_distance(B, N, S, k):
    """papc_knn_f32 == the first k columns of a STABLE argsort of square_distance(query, xyz) (the oracle's
    pinned arithmetic), distances bit-exact; fewer than k points -> index N / +inf."""
    xyz = _xyz(B, N, seed=3 * N + S)
    q = _xyz(B, S, seed=7 * S + N)
    if N >= 8:
        xyz[:, 5] = xyz[:, 2]            # duplicated points: equal distances, the lower index must come first
    d = capi.square_distance(q, xyz)                               # [B,S,N]
    order = np.argsort(d, axis=-1, kind="stable")[:, :, :k]
    idx, dist_ = layers.knn_points(k, _cu(xyz), _cu(q), return_dist=True)
    idx, dist_ = idx.cpu().numpy(), dist_.cpu().numpy()
    kk = min(k, N)
    np.testing.assert_array_equal(idx[:, :, :kk], order[:, :, :kk])
    np.testing.assert_array_equal(dist_[:, :, :kk], np.take_along_axis(d, order[:, :, :kk], -1))
    if kk < k:
        assert (idx[:, :, kk:] == N).all() and np.isinf(dist_[:, :, kk:]).all()


# ------------------------------------------------------------------ fused sampling: FPS + ball query + moments
@pytest.mark.parametrize("B,N,S,K,r", [(32, 1024, 512, 32, 0.2), (32, 512, 128, 64, 0.4), (5, 700, 100, 16, 0.15),
                                       (74, 300, 40, 8, 0.3), (3, 1000, 64, 48, 0.02)])
def test_fused_sample_group_bit_exact(B, N, S, K, r):
    """papc_sample_group_f32 (the ball query of centroid i overlapping the FPS recurrence, one launch) ==
    papc_fps_f32 + papc_ball_query_f32, bit for bit; its moment partial sums == the sums over the grouped,
    centred points (fp64 reference from the indices)."""
    xyz = _xyz(B, N, seed=B * 7 + N)
    start = synth.fps_start(B, N, seed=3)
    x = _cu(xyz)
    assert layers.L.lib().papc_sample_group_parts(B, N, S, K) >= 1
    got = layers._sample_group_fused(x, S, _cu(start), r, K, True)
    assert got is not None
    new_xyz, idx, mom = got
    fps_ref, nx_ref = layers.farthest_point_sample_idx(x, S, _cu(start), return_xyz=True)
    idx_ref = layers._ball_query(r, K, x, nx_ref, torch.int32)
    assert torch.equal(new_xyz, nx_ref)
    assert torch.equal(idx, idx_ref)
    # moments: rows = all B*S*K grouped points (padding rows repeat the first neighbour; an empty ball's N clamps)
    ii = np.minimum(idx.cpu().numpy().astype(np.int64), N - 1)
    g = np.take_along_axis(xyz[:, None, :, :].astype(np.float32), ii[..., None].repeat(3, -1), axis=2)
    cen = (g - new_xyz.cpu().numpy()[:, :, None, :]).astype(np.float64).reshape(-1, 3)
    want = np.array([cen[:, 0].sum(), cen[:, 1].sum(), cen[:, 2].sum(),
                     (cen[:, 0] ** 2).sum(), (cen[:, 0] * cen[:, 1]).sum(), (cen[:, 0] * cen[:, 2]).sum(),
                     (cen[:, 1] ** 2).sum(), (cen[:, 1] * cen[:, 2]).sum(), (cen[:, 2] ** 2).sum()])
    np.testing.assert_allclose(mom.cpu().numpy().sum(0), want, rtol=2e-6, atol=1e-6 * cen.shape[0] ** 0.5)


def test_fused_sampling_layer_equals_unfused(monkeypatch):
    """A SetAbstraction layer with the fused sampling launch == the same layer on the separate kernels:
    coordinates bit-exact, features within fp32 rounding of the two moment summation orders."""
    B, N = 8, 1024
    xyz = _cu(synth.clouds(B, N, seed=17))
    st = _cu(synth.fps_start(B, N, seed=18))
    sa = layers.PointNetSetAbstraction(512, 0.2, 32, 3, [64, 64, 128], False).to(DEV)
    outs = {}
    for flag in (True, False):
        monkeypatch.setattr(layers, "FUSED_SAMPLING", flag)
        ox, op = sa(xyz, None, start_idx=st)
        outs[flag] = (ox.clone(), op.clone())
    assert torch.equal(outs[True][0], outs[False][0])
    assert float((outs[True][1] - outs[False][1]).abs().max()) <= 1e-5
    assert layers.L.lib().papc_sample_group_parts(200, 1024, 512, 32) == 0   # too many clouds to be co-resident
