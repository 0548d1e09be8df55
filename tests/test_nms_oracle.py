"""CPU: oracle/nms_oracle.c against the golden vectors produced by the REFERENCE'S OWN numba.cuda kernels run under
numba's CUDA simulator (tests/golden/make_golden_nms.py -> nms_ref.npz): keep lists exact, IoU matrices <= 1e-6."""
import os

import numpy as np
import pytest

from oracle import capi

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "nms_ref.npz"))


@pytest.mark.parametrize("t", [0.3, 0.6])
def test_nms_keep_lists(t):
    np.testing.assert_array_equal(capi.nms(G["nms_dets"], t), G[f"nms_keep_{t}"])


@pytest.mark.parametrize("t", [0.1, 0.4])
def test_rotate_nms_keep_lists(t):
    np.testing.assert_array_equal(capi.nms(G["rnms_dets"], t, rotated=True), G[f"rnms_keep_{t}"])


@pytest.mark.parametrize("crit", [-1, 0, 1, 2])
def test_rotate_iou(crit):
    out = capi.rotate_iou(G["riou_boxes"], G["riou_query"], crit)
    np.testing.assert_allclose(out, G[f"riou_eval_{crit}"], rtol=0, atol=1e-6)
    if crit == -1:
        np.testing.assert_allclose(out, G["riou"], rtol=0, atol=1e-6)
        # reference quirk, reproduced: IDENTICAL boxes do not give IoU 1 -- both boxes' corners pass
        # point_in_quadrilateral, the clipped polygon holds every corner twice and the fan area comes out at
        # half the box (IoU 1/3) or, depending on the vertex order after the sort, 0
        assert np.all(out[np.arange(5), np.arange(5)] < 0.34)
        assert out[6, 6] == 0.25 and out[7, 7] == 0.0                            # contained box; disjoint boxes


def test_nms_known_answers():
    # three boxes: b overlaps a heavily (IoU 0.68), c is far away; scores a > b > c
    d = np.array([[0, 0, 10, 10, .9], [1, 1, 11, 11, .8], [50, 50, 60, 60, .7]], np.float32)
    assert capi.nms(d, 0.5).tolist() == [0, 2]
    assert capi.nms(d, 0.75).tolist() == [0, 1, 2]        # IoU = 100 / 142 = 0.704 with the reference's +1 pixel convention
    assert capi.nms(d[::-1].copy(), 0.5).tolist() == [2, 0]          # original indices, best score first
    assert capi.nms(np.zeros((0, 5), np.float32), 0.5).tolist() == []
    # equal scores: the higher index is visited first (numpy's stable argsort reversed)
    e = np.array([[0, 0, 10, 10, .5], [0, 0, 10, 10, .5]], np.float32)
    assert capi.nms(e, 0.5).tolist() == [1]
