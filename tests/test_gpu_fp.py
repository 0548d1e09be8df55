"""GPU parity tests of PointNetFeaturePropagation (layers.py:284-335): CUDA
through the C ABI vs the NumPy restatement on the same seeded inputs."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import layers_np  # noqa: E402
from papc_b200 import layers, synth  # noqa: E402

DEV = "cuda:0"
TOL = dict(rtol=1e-5, atol=1e-5)


def _cu(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _case(B, N, S, D1, D2, seed):
    rng = np.random.default_rng(seed)
    xyz1 = synth.clouds(B, N, seed=seed)                       # [B,3,N]
    xyz2 = np.ascontiguousarray(xyz1[:, :, rng.permutation(N)[:S]])
    p1 = rng.standard_normal((B, D1, N)).astype(np.float32) if D1 else None
    p2 = rng.standard_normal((B, D2, S)).astype(np.float32)
    return xyz1, xyz2, p1, p2


@pytest.mark.parametrize("B,N,S,D1,D2", [(2, 256, 64, 6, 32), (3, 100, 16, 0, 20), (2, 64, 2, 5, 7),
                                         (2, 64, 1, 3, 9), (1, 2048, 512, 22, 128)])
def test_interpolation_matches_oracle(B, N, S, D1, D2):
    xyz1, xyz2, p1, p2 = _case(B, N, S, D1, D2, seed=B * 1000 + N)
    ref = layers_np.PointNetFeaturePropagation(D1 + D2, [8]).interpolate(xyz1, xyz2, p1, p2)   # [B,N,D1+D2]
    rows, cin = layers.feature_interpolate(_cu(xyz1.transpose(0, 2, 1)), _cu(xyz2.transpose(0, 2, 1)),
                                           _cu(p1.transpose(0, 2, 1)) if p1 is not None else None,
                                           _cu(p2.transpose(0, 2, 1)))
    assert cin == D1 + D2 and rows.shape[1] % 8 == 0
    got = rows.cpu().numpy().reshape(B, N, -1)
    assert np.all(got[:, :, cin:] == 0)                        # row padding
    if D1:
        np.testing.assert_array_equal(got[:, :, :D1], ref[:, :, :D1])
    np.testing.assert_allclose(got[:, :, D1:cin], ref[:, :, D1:], rtol=2e-6, atol=2e-6)


def test_interpolation_uses_first_three_sampled_points():
    """The reference quirk (argsort of the sorted distances = identity, layers.py:317-318): a point
    that coincides with sampled point 5 still receives (almost exactly) the features of point 0."""
    B, N, S, D2 = 1, 8, 8, 4
    xyz = np.zeros((B, 3, N), np.float32)
    xyz[0, 0, :] = np.arange(N, dtype=np.float32)              # points on a line, xyz2 == xyz1
    p2 = np.zeros((B, D2, S), np.float32)
    p2[0, :, 0] = 1.0
    p2[0, :, 5] = 100.0
    rows, _ = layers.feature_interpolate(_cu(xyz.transpose(0, 2, 1)), _cu(xyz.transpose(0, 2, 1)), None,
                                         _cu(p2.transpose(0, 2, 1)))
    got = rows.cpu().numpy().reshape(N, -1)[5, :D2]
    assert np.all(got > 0.99) and np.all(got <= 1.0)           # weight ~1 on sampled point 0, not on 5


@pytest.mark.parametrize("bn_mode", ["batch", "running"])
@pytest.mark.parametrize("B,N,S,D1,D2,mlp", [(2, 512, 128, 64, 128, [128, 64]),       # 8-aligned channels
                                              (2, 256, 64, 22, 128, [128, 128, 128]),  # 150 channels (fp1 of the seg models)
                                              (2, 128, 1, 6, 64, [32])])               # S == 1 (fp3)
def test_feature_propagation_layer(B, N, S, D1, D2, mlp, bn_mode):
    xyz1, xyz2, p1, p2 = _case(B, N, S, D1, D2, seed=7)
    rng = np.random.default_rng(11)
    g = layers.PointNetFeaturePropagation(D1 + D2, mlp)
    r = layers_np.PointNetFeaturePropagation(D1 + D2, mlp)
    c = D1 + D2
    for l, co in enumerate(mlp):
        w = (rng.standard_normal((co, c)) * np.sqrt(2.0 / c)).astype(np.float32)
        b = rng.uniform(-0.1, 0.1, co).astype(np.float32)
        gamma = rng.uniform(0.5, 1.5, co).astype(np.float32)
        beta = rng.uniform(-0.2, 0.2, co).astype(np.float32)
        rm = rng.uniform(-0.1, 0.1, co).astype(np.float32)
        rv = rng.uniform(0.5, 1.5, co).astype(np.float32)
        g.mlp_convs[l].weight, g.mlp_convs[l].bias = torch.from_numpy(w.reshape(co, c, 1)), torch.from_numpy(b)
        g.mlp_bns[l].weight, g.mlp_bns[l].bias = torch.from_numpy(gamma), torch.from_numpy(beta)
        g.mlp_bns[l]._mean, g.mlp_bns[l]._variance = torch.from_numpy(rm), torch.from_numpy(rv)
        r.mlp_convs[l].weight, r.mlp_convs[l].bias = w.reshape(co, c, 1, 1), b
        r.mlp_bns[l].weight, r.mlp_bns[l].bias = gamma, beta
        r.mlp_bns[l]._mean, r.mlp_bns[l]._variance = rm.copy(), rv.copy()
        r.mlp_bns[l].training = bn_mode == "batch"
        c = co
    g.bn_mode = bn_mode
    g.to(DEV)
    got = g(_cu(xyz1), _cu(xyz2), _cu(p1) if p1 is not None else None, _cu(p2))
    ref = r(xyz1, xyz2, p1, p2)
    assert tuple(got.shape) == ref.shape == (B, mlp[-1], N)
    np.testing.assert_allclose(got.cpu().numpy(), ref, **TOL)


def test_msg_seg_encoder_decoder_shapes_and_determinism():
    """sa1 -> sa2 -> sa3 -> fp3 -> fp2 -> fp1 wired as PointNet2_MSG_Seg (segment/pointnet2/pointnet2.py:
    62-67, 84-93) at a reduced size: output [B,128,N], finite, run-to-run identical."""
    from papc_b200 import sa_stack
    B, N = 2, 1024
    xyz = _cu(synth.clouds(B, N, seed=1))
    onehot = torch.zeros((B, 16, N), device=DEV)
    onehot[:, 2, :] = 1.0
    st1 = _cu(synth.fps_start(B, N, seed=2))
    st2 = torch.zeros(B, dtype=torch.int64, device=DEV)
    torch.manual_seed(0)
    m = sa_stack.MSGSegEncoderDecoder().to(DEV)
    a = m(xyz, onehot, start_idx=(st1, st2))
    b = m(xyz, onehot, start_idx=(st1, st2))
    assert tuple(a.shape) == (B, 128, N)
    assert torch.isfinite(a).all()
    assert torch.equal(a, b)
