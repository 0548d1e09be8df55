"""Generate tests/golden/pillars_ref.npz by EXECUTING THE REFERENCE'S OWN CLASSES
(PAPC/models/detect/pointpillars/models/bones/pillars.py: PFNLayer, PillarFeatureNet,
PointPillarsScatter; libs/tools/__init__.py: get_paddings_indicator; libs/functional.py: mask_select,
select_change), cut out of their modules with ``ast`` (the modules' own imports pull in the whole
detector) and run over the NumPy stand-in for paddle (tests/golden/paddle_stub.py).

What this pins: the reference's wiring -- decoration order [features, f_cluster, f_center(, dist)],
the (x, y) offsets from coors[:, 3] / coors[:, 2], the padding mask applied after the concat, the
max / tile / concat of PFNLayer, squeeze(), the per-sample scatter with idx = y * nx + x.  The
``nn.Linear`` / ``nn.BatchNorm1D`` arithmetic inside is the stub's restatement (fp64 accumulation), so
the decorated tensor (recorded as the first Linear's input) and the canvas are exact references, the PFN
output a reference up to that arithmetic.  Build-container only:  python tests/golden/make_golden_pillars.py
"""
import ast
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import paddle_stub  # noqa: E402
from papc_b200 import synth  # noqa: E402

PP = "/root/reference/PAPC/models/detect/pointpillars/"


def cut(path, names, ns):
    tree = ast.parse(open(path).read())
    body = [n for n in tree.body if isinstance(n, (ast.FunctionDef, ast.ClassDef)) and n.name in names]
    assert len(body) == len(names), (path, names)
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)


def load_reference():
    P = paddle_stub.install()
    ns = {"paddle": P, "nn": paddle_stub.nn, "F": paddle_stub.functional}
    cut(PP + "libs/tools/__init__.py", ["get_paddings_indicator"], ns)
    cut(PP + "libs/functional.py", ["mask_select", "select_change"], ns)
    cut(PP + "models/bones/pillars.py", ["PFNLayer", "PillarFeatureNet", "PointPillarsScatter"], ns)
    return P, ns


def pillars(rng, P, T, F, nx, ny, batch):
    """Zero-padded pillars as the voxeliser emits them, unique (y, x) cells per sample."""
    num = rng.integers(1, T + 1, P).astype(np.int32)
    feats = np.zeros((P, T, F), np.float32)
    coors = np.zeros((P, 4), np.int32)
    b = np.sort(rng.integers(0, batch, P)).astype(np.int32)
    for s in range(batch):
        m = np.nonzero(b == s)[0]
        cells = rng.permutation(nx * ny)[:len(m)]
        coors[m, 0], coors[m, 2], coors[m, 3] = s, cells // nx, cells % nx
    vs, rg = synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE
    for p in range(P):
        n = num[p]
        feats[p, :n, 0] = rg[0] + (coors[p, 3] + rng.random(n)) * vs[0]
        feats[p, :n, 1] = rg[1] + (coors[p, 2] + rng.random(n)) * vs[1]
        feats[p, :n, 2] = rng.uniform(rg[2], rg[5], n)
        feats[p, :n, 3:] = rng.random((n, F - 3))
    return feats, num, coors


if __name__ == "__main__":
    Pd, R = load_reference()
    T_ = Pd.to_tensor
    rng = np.random.default_rng(123)
    nx, ny = 40, 30
    out = {}
    feats, num, coors = pillars(rng, 200, 12, 4, nx, ny, batch=2)
    out.update(features=feats, num_voxels=num, coors=coors, nx=nx, ny=ny)
    for tag, filters, dist in (("one", (64,), False), ("two", (32, 64), False), ("dist", (16,), True)):
        net = R["PillarFeatureNet"](4, True, filters, dist, synth.KITTI_VOXEL_SIZE, synth.KITTI_PC_RANGE)
        for i, pfn in enumerate(net.pfn_layers):
            w = (rng.standard_normal(tuple(pfn.linear.weight.shape)) / 3.0).astype(np.float32)
            gam = rng.uniform(0.5, 1.5, w.shape[1]).astype(np.float32)
            bet = rng.uniform(-0.2, 0.2, w.shape[1]).astype(np.float32)
            pfn.linear.weight, pfn.norm.weight, pfn.norm.bias = T_(w), T_(gam), T_(bet)
            out[f"{tag}_w{i}"], out[f"{tag}_gamma{i}"], out[f"{tag}_beta{i}"] = w, gam, bet
        y = net(T_(feats), T_(num), T_(coors)).numpy()
        out[f"{tag}_decorated"] = net.pfn_layers[0].linear.inputs[0]      # what the first Linear received
        out[f"{tag}_out"] = y
    # scatter: two samples, then three with an empty one in the middle (the `else: pass` branch)
    vf = out["one_out"]
    sc = R["PointPillarsScatter"]([1, 1, ny, nx], num_input_features=vf.shape[1])
    out["canvas_b2"] = sc(T_(vf), T_(coors), 2).numpy()
    coors3 = coors.copy(); coors3[coors3[:, 0] == 1, 0] = 2
    out["coors_b3"] = coors3
    out["canvas_b3"] = sc(T_(vf), T_(coors3), 3).numpy()
    np.savez_compressed(os.path.join(HERE, "pillars_ref.npz"), **out)
    for k, v in out.items():
        v = np.asarray(v)
        print(f"{k:18s} {str(v.dtype):8s} {v.shape}")
