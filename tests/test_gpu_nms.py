"""GPU parity tests of papc_b200.nms (csrc/nms.cu) through the C ABI: keep lists and IoU matrices bit-exact against
oracle/nms_oracle.c, and against the golden vectors of the reference's own kernels (tests/golden/nms_ref.npz)."""
import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

from oracle import capi  # noqa: E402
from papc_b200 import nms  # noqa: E402

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "nms_ref.npz"))
DEV = "cuda:0"


def _boxes(rng, n, extent):
    c = rng.uniform(0, extent, (n, 2))
    wh = rng.uniform(2.0, 12.0, (n, 2))
    return np.concatenate([c - wh / 2, c + wh / 2, rng.uniform(0.01, 1.0, (n, 1))], 1).astype(np.float32)


def _rboxes(rng, n, extent, score=True):
    cols = [rng.uniform(0, extent, (n, 2)), rng.uniform(1.5, 6.0, (n, 2)), rng.uniform(-np.pi, np.pi, (n, 1))]
    if score:
        cols.append(rng.uniform(0.01, 1.0, (n, 1)))
    return np.concatenate(cols, 1).astype(np.float32)


def test_golden_vectors_of_the_reference_kernels():
    for t in (0.3, 0.6):
        assert nms.nms_gpu(G["nms_dets"], t) == G[f"nms_keep_{t}"].tolist()
    for t in (0.1, 0.4):
        assert nms.rotate_nms_gpu(G["rnms_dets"], t) == G[f"rnms_keep_{t}"].tolist()
    for crit in (-1, 0, 1, 2):
        out = nms.rotate_iou_gpu_eval(G["riou_boxes"], G["riou_query"], crit)
        np.testing.assert_allclose(out, G[f"riou_eval_{crit}"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(nms.rotate_iou_gpu(G["riou_boxes"], G["riou_query"]), G["riou"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("n,extent,t", [(1000, 120.0, 0.5), (3000, 150.0, 0.3), (64, 30.0, 0.5), (65, 30.0, 0.1),
                                        (1, 10.0, 0.5), (777, 40.0, 0.7)])
def test_nms_bit_exact_vs_oracle(n, extent, t):
    d = _boxes(np.random.default_rng(n), n, extent)
    d[n // 2:, 4] = d[:n - n // 2, 4]            # duplicated scores: the tie rule matters
    assert nms.nms_gpu(d, t) == capi.nms(d, t).tolist()


@pytest.mark.parametrize("n,extent,t", [(1000, 80.0, 0.3), (500, 30.0, 0.1), (130, 20.0, 0.5), (2, 3.0, 0.01)])
def test_rotate_nms_bit_exact_vs_oracle(n, extent, t):
    d = _rboxes(np.random.default_rng(7 * n), n, extent)
    assert nms.rotate_nms_gpu(d, t) == capi.nms(d, t, rotated=True).tolist()


@pytest.mark.parametrize("N,K", [(300, 200), (64, 64), (65, 1), (1, 130)])
def test_rotate_iou_bit_exact_vs_oracle(N, K):
    rng = np.random.default_rng(N + K)
    a, b = _rboxes(rng, N, 25.0, score=False), _rboxes(rng, K, 25.0, score=False)
    for crit in (-1, 0, 1, 2):
        np.testing.assert_array_equal(nms.rotate_iou_gpu_eval(a, b, crit), capi.rotate_iou(a, b, crit))


def test_device_resident_form_and_edge_cases():
    d = _boxes(np.random.default_rng(3), 500, 60.0)
    keep, num = nms.nms_device(torch.from_numpy(d).to(DEV), 0.4)
    m = int(num.item())
    ref = capi.nms(d, 0.4)
    assert keep.dtype == torch.int32 and m == len(ref)
    np.testing.assert_array_equal(keep[:m].cpu().numpy(), ref)
    assert (keep[m:] == -1).all()
    assert nms.nms_gpu(np.zeros((0, 5), np.float32), 0.5) == []
    assert nms.rotate_iou_gpu(np.zeros((0, 5), np.float32), np.zeros((3, 5), np.float32)).shape == (0, 3)
    # idempotence: NMS of the kept boxes keeps all of them
    kept = d[ref]
    assert sorted(nms.nms_gpu(kept, 0.4)) == list(range(len(ref)))
    with pytest.raises(ValueError):
        nms.nms_gpu(np.zeros((4, 6), np.float32), 0.5)
