"""The torch-backed ``paddle`` facade (papc_b200/compat) checked on the CPU: the reference's OWN files --
pointnet2_basic_layers.py and the two model files, staged unmodified under oracle/_ref/ by oracle/build.py --
run over the facade on torch-CPU and reproduce tests/golden/models_ref.npz (made by running the same files over
the NumPy stand-in, tests/golden/make_golden_models.py).  This is the configuration bench.py's reference arm
times; the GPU form (the same model files on papc_b200.layers) is tests/test_gpu_reference_models.py."""
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
import param_gen  # noqa: E402
from oracle import build as oracle_build  # noqa: E402

pytestmark = pytest.mark.skipif(oracle_build.ref_file("pointnet2_basic_layers.py") is None,
                                reason="oracle/_ref not staged (python oracle/build.py in the build container)")

CASES = [("PointNet2_SSG_Clas", False, False), ("PointNet2_SSG_Clas", True, False), ("PointNet2_SSG_Seg", False, True)]


@pytest.fixture()
def facade():
    from papc_b200 import compat
    saved = {k: sys.modules.get(k) for k in ("paddle", "paddle.nn", "paddle.nn.functional", "PAPC", "PAPC.models",
                                             "PAPC.models.layers")}
    compat.install(layers_file=oracle_build.ref_file("pointnet2_basic_layers.py"), device="cpu", force=True)
    ns = {}
    ns.update(compat.load_model_file(oracle_build.ref_file("classify_pointnet2.py")))
    ns.update(compat.load_model_file(oracle_build.ref_file("segment_pointnet2.py")))
    yield compat, ns
    compat.clear_fps_starts()
    compat.paddle_torch._DEVICE[0] = None
    for k, v in saved.items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v


@pytest.mark.parametrize("name,normal_channel,seg", CASES)
def test_reference_files_over_facade_cpu(facade, golden_dir, name, normal_channel, seg):
    compat, ns = facade
    g = np.load(os.path.join(golden_dir, "models_ref.npz"))
    tag = name + ("_nc" if normal_channel else "")
    torch.manual_seed(0)
    model = ns[name](normal_channel=normal_channel)
    n = len(param_gen.install(model, tag, wrap=lambda a: torch.from_numpy(np.ascontiguousarray(a))))
    assert n > 10
    x = np.concatenate([g["xyz"], g["normals"]], 1) if normal_channel else g["xyz"]
    inputs = (x, g["labels"]) if seg else x
    model.eval()
    compat.paddle_torch.queue_randint([g["start1"], g["start2"]])
    with torch.no_grad():
        y = model(inputs).numpy()
    if seg:
        np.testing.assert_allclose(y[:, ::8], g[f"{tag}:eval:sub8"], rtol=2e-4, atol=2e-4)
    else:
        np.testing.assert_allclose(y, g[f"{tag}:eval"], rtol=1e-4, atol=1e-4)
