"""CPU tests of the model-level oracle (oracle/models_np.py) and of the host side of papc_b200/models.py
(SURVEY.md 8f row N2): known answers for ``Categorical``, shapes / train-vs-eval semantics of the
restated heads, attribute names and parameter shapes of the product models, no CPU fallback."""
import numpy as np
import pytest

torch = pytest.importorskip("torch")

from oracle import models_np  # noqa: E402
from papc_b200 import _lib, models, synth  # noqa: E402


def test_categorical_known_answer():
    y = np.array([[2], [0]])
    got = models_np.Categorical(y, 4)                         # layers.py:7-14
    assert got.shape == (2, 4, 1) and got.dtype == np.float32
    np.testing.assert_array_equal(got[:, :, 0], [[0, 0, 1, 0], [1, 0, 0, 0]])
    np.testing.assert_array_equal(models.Categorical(y, 4).numpy(), got)
    np.testing.assert_array_equal(models.Categorical(y.reshape(-1), 4).numpy(), got)
    with pytest.raises(IndexError):
        models.Categorical(np.array([[4]]), 4)


def test_oracle_ssg_clas_head_semantics():
    B, N = 2, 512
    m = models_np.PointNet2_SSG_Clas(num_classes=16, rng=np.random.default_rng(1))
    xyz = synth.clouds(B, N, seed=0)
    st = (synth.fps_start(B, N, seed=1), np.zeros(B, dtype=np.int64))
    ev = m.eval()(xyz, start_idx=st)
    assert ev.shape == (B, 16) and np.isfinite(ev).all()
    np.testing.assert_array_equal(ev, m(xyz, start_idx=st))   # eval: no state moves
    mean0 = m.bn1._mean.copy()
    tr = m.train()(xyz, start_idx=st)
    assert tr.shape == (B, 16) and not np.allclose(tr, ev)    # batch statistics over the B rows
    assert not np.array_equal(m.bn1._mean, mean0)             # registered BatchNorm1D: running stats move
    # with B = 2 every channel normalises to +-1/sqrt(1+eps/var): relu(bn) is in [0, 1]
    h = models_np.LN.relu(m.bn1(m.fc1(np.ones((2, 1024), np.float32) * [[1.0], [2.0]])))
    assert h.max() <= 1.0 + 1e-6


@pytest.mark.parametrize("name,normal", [("PointNet2_SSG_Seg", False), ("PointNet2_MSG_Seg", True)])
def test_oracle_seg_shapes(name, normal):
    B, N = 1, 512
    m = getattr(models_np, name)(num_classes=16, num_parts=50, normal_channel=normal, rng=np.random.default_rng(2))
    xyz = synth.clouds(B, N, seed=3)
    if normal:
        xyz = np.concatenate([xyz, xyz], axis=1)
    out = m((xyz, np.array([[5]])), start_idx=(np.zeros(B, np.int64), np.zeros(B, np.int64)))
    assert out.shape == (B, N, 50) and np.isfinite(out).all()


def test_product_models_mirror_reference_attributes():
    m = models.PointNet2_SSG_Clas(num_classes=40, normal_channel=True)
    assert m.sa1.in_channel == 6 and m.sa3.group_all
    assert tuple(m.fc1.weight.shape) == (512, 1024) and tuple(m.fc3.weight.shape) == (40, 256)
    assert m.drop1.p == 0.4 and m.drop2.p == 0.4
    m = models.PointNet2_MSG_Clas()
    assert m.sa1.nsample_list == [16, 32, 128] and m.sa2.in_channel == 320 and m.drop2.p == 0.5
    m = models.PointNet2_SSG_Seg(normal_channel=True)
    assert m.sa1.in_channel == 9 and m.fp1.in_channel == 153 and len(m.fp1.mlp_convs) == 3
    m = models.PointNet2_MSG_Seg(num_parts=50)
    assert m.fp3.in_channel == 1536 and m.fp1.in_channel == 150 and len(m.fp1.mlp_convs) == 2
    assert tuple(m.conv2.weight.shape) == (50, 128, 1) and tuple(m.bn1._mean.shape) == (128,)


def test_seg_head_is_registered_like_the_reference():
    """segment/pointnet2/pointnet2.py:21-24 assigns conv1 / bn1 / conv2 as attributes, so Paddle registers them:
    their tensors must be in parameters() / state_dict() under Paddle's key names."""
    m = models.PointNet2_SSG_Seg()
    keys = set(m.state_dict().keys())
    assert {"conv1.weight", "conv1.bias", "bn1.weight", "bn1.bias", "bn1._mean", "bn1._variance",
            "conv2.weight", "conv2.bias"} <= keys
    sd = {k: v.clone() + 1.0 for k, v in m.state_dict().items()}
    m.load_state_dict(sd)
    assert float(m.bn1._mean[0]) == 1.0 and float(m.conv2.bias[0]) == 1.0
    c = models.PointNet2_SSG_Clas()
    assert {"fc1.weight", "bn1.weight", "fc3.bias"} <= set(c.state_dict().keys())


def test_product_models_have_no_cpu_fallback():
    m = models.PointNet2_SSG_Clas()
    with pytest.raises(_lib.PapcError):
        m(torch.zeros(2, 3, 1024))
    s = models.PointNet2_MSG_Seg()
    with pytest.raises(_lib.PapcError):
        s((torch.zeros(2, 3, 1024), np.array([[0], [1]])))
