"""Generate tests/golden/models_ref.npz by EXECUTING THE REFERENCE'S OWN MODEL DEFINITIONS
(PAPC/models/classify/pointnet2/pointnet2.py, PAPC/models/segment/pointnet2/pointnet2.py, on top of
PAPC/models/layers/pointnet2_basic_layers.py -- all unmodified) over the NumPy stand-in for paddle
(tests/golden/paddle_stub.py).

What this pins: the reference's model wiring -- which layer feeds which, the SA / MSG / FP
configurations, ``Categorical`` and the [one_hot, xyz, points] concat, the heads, and the train / eval
behaviour that follows from how the reference registers its layers (head BatchNorms follow ``eval()``,
the list-held ones inside SetAbstraction / FeaturePropagation do not).  The Conv / Linear / BatchNorm
arithmetic is the stub's restatement (fp64 accumulation).  Parameters are a deterministic function of
their attribute path (tests/golden/param_gen.py), so the file holds inputs and outputs only.
Build-container only:
    python tests/golden/make_golden_models.py
"""
import ast
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import paddle_stub  # noqa: E402
import param_gen  # noqa: E402
from papc_b200 import synth  # noqa: E402

REF = "/root/reference/PAPC/models/"


def load_reference():
    P = paddle_stub.install()
    layers = types.ModuleType("ref_layers")
    exec(compile(open(REF + "layers/pointnet2_basic_layers.py").read(), "pointnet2_basic_layers.py", "exec"),
         layers.__dict__)
    ns = dict(layers.__dict__)          # the model files do `from PAPC.models.layers import ...`
    for path in ("classify/pointnet2/pointnet2.py", "segment/pointnet2/pointnet2.py"):
        tree = ast.parse(open(REF + path).read())
        body = [n for n in tree.body if isinstance(n, ast.ClassDef)]
        exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return P, ns


if __name__ == "__main__":
    P, R = load_reference()
    rng = np.random.default_rng(77)
    B, N = 2, 1024
    xyz = synth.clouds(B, N, seed=31)                                       # [B,3,N]
    nrm = rng.standard_normal((B, 3, N)).astype(np.float32)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    st1, st2 = synth.fps_start(B, N, seed=32), synth.fps_start(B, 512, seed=33)
    labels = np.array([[3], [15]], dtype=np.int64)
    out = dict(xyz=xyz, normals=nrm, start1=st1, start2=st2, labels=labels)

    def run(model, inputs):
        paddle_stub._next_randint.extend([st1, st2])                        # sa1's and sa2's FPS start draws
        y = model(inputs).numpy()
        assert not paddle_stub._next_randint
        return y

    for name, kw, seg in (("PointNet2_SSG_Clas", dict(normal_channel=False), False),
                          ("PointNet2_SSG_Clas", dict(normal_channel=True), False),
                          ("PointNet2_MSG_Clas", dict(normal_channel=False), False),
                          ("PointNet2_SSG_Seg", dict(normal_channel=False), True),
                          ("PointNet2_MSG_Seg", dict(normal_channel=True), True)):
        tag = name + ("_nc" if kw["normal_channel"] else "")
        model = R[name](**kw)
        n_params = len(param_gen.install(model, tag, wrap=paddle_stub.Tensor))
        x = np.concatenate([xyz, nrm], 1) if kw["normal_channel"] else xyz
        inputs = (x, labels) if seg else x
        model.eval()
        out[f"{tag}:eval"] = run(model, inputs)
        model.train()
        for d in ("drop1", "drop2"):
            if hasattr(model, d):
                getattr(model, d).p = 0                                     # the mask is not reproducible
        out[f"{tag}:train"] = run(model, inputs)
        out[f"{tag}:bn1_mean_after_train"] = model.bn1._mean.numpy()
        out[f"{tag}:bn1_var_after_train"] = model.bn1._variance.numpy()
        print(tag, n_params, "parameterised layers", out[f"{tag}:eval"].shape, float(np.abs(out[f"{tag}:eval"]).mean()),
              float(np.abs(out[f"{tag}:train"] - out[f"{tag}:eval"]).mean()))
    for k in [k for k in out if k.endswith((":eval", ":train")) and out[k].ndim == 3]:
        y = out.pop(k)                  # segmentation logits [B,N,50]: every 8th point + two checksums
        out[k + ":sub8"] = np.ascontiguousarray(y[:, ::8])
        out[k + ":sum"] = np.array([y.astype(np.float64).sum(), np.abs(y.astype(np.float64)).sum()])
    np.savez_compressed(os.path.join(HERE, "models_ref.npz"), **out)
    print(len(out), "arrays")
