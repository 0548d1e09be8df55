"""Run the reference's model files UNCHANGED on this library (north_star: "exposed behind
the same Paddle layer signatures ... so the existing model definitions call them unchanged").

    from papc_b200 import compat
    compat.install()                                  # paddle facade + PAPC.models.layers -> papc_b200.layers
    ns = compat.load_model_file("/path/to/PAPC/models/classify/pointnet2/pointnet2.py")
    model = ns["PointNet2_SSG_Clas"]()                # the reference's own class, on the sm_100a kernels

``install(layers_file=...)`` binds ``PAPC.models.layers`` to a layers FILE executed over the facade instead (the
reference's own pointnet2_basic_layers.py: bench.py's CPU reference arm).  With a real PaddlePaddle installed the
facade is not needed: INTEGRATION.md shows the ctypes / PD_BUILD_OP binding for that case.
"""
from __future__ import annotations

import sys
import types

from . import paddle_torch


def queue_fps_starts(values):
    """Seeded FPS start indices for the next sampled layers (the reference draws them with paddle.randint,
    layers.py:76): consumed by ``papc_b200.layers`` and by the facade's ``paddle.randint`` alike."""
    from .. import layers
    layers._START_QUEUE.extend(values)
    paddle_torch._RANDINT_QUEUE.extend(values)


def clear_fps_starts():
    from .. import layers
    del layers._START_QUEUE[:]
    del paddle_torch._RANDINT_QUEUE[:]


def install(layers_file=None, device=None, force=False):
    """Register the facade as ``paddle`` (unless a real paddle is importable and ``force`` is False) and bind
    ``PAPC.models.layers``: to ``papc_b200.layers`` (default) or to ``layers_file`` executed over the facade."""
    if not force and "paddle" in sys.modules and sys.modules["paddle"] is not paddle_torch:
        raise RuntimeError("a different 'paddle' module is already imported")
    sys.modules["paddle"] = paddle_torch
    sys.modules["paddle.nn"] = paddle_torch.nn
    sys.modules["paddle.nn.functional"] = paddle_torch.functional
    if device is not None:
        paddle_torch.set_device(device)
    mod = types.ModuleType("PAPC.models.layers")
    if layers_file is None:
        from .. import layers, models
        if paddle_torch.get_device().type == "cuda":
            layers.DEFAULT_DEVICE = paddle_torch.get_device()   # Paddle builds layers on the current device
        for name in ("PointNetSetAbstraction", "PointNetSetAbstractionMsg", "PointNetFeaturePropagation",
                     "square_distance", "index_points", "farthest_point_sample", "query_ball_point",
                     "sample_and_group", "sample_and_group_all"):
            setattr(mod, name, getattr(layers, name))
        mod.Categorical = lambda y, num_class=16: paddle_torch._wrap(
            models.Categorical(y, num_class, device=paddle_torch.get_device()))
        mod.pc_normalize = getattr(layers, "pc_normalize", None)
    else:
        mod.__file__ = layers_file
        exec(compile(open(layers_file).read(), layers_file, "exec"), mod.__dict__)
    papc, pm = types.ModuleType("PAPC"), types.ModuleType("PAPC.models")
    papc.models, pm.layers = pm, mod
    sys.modules["PAPC"], sys.modules["PAPC.models"], sys.modules["PAPC.models.layers"] = papc, pm, mod
    return mod


def load_model_file(path):
    """Execute a reference model file as it is (its ``import paddle`` / ``from PAPC.models.layers import ...``
    resolve to what ``install`` registered) and return its namespace."""
    ns = {"__name__": "papc_reference_model", "__file__": path}
    exec(compile(open(path).read(), path, "exec"), ns)
    return ns
