"""Deterministic parameters for the model-level golden vectors: every parameter is a function of its
attribute path (``<model>.sa1.mlp_convs.0.weight`` ...) and shape only, so the generator
(make_golden_models.py, over the reference's classes) and the tests (over the oracle's classes, which
use the reference's attribute names) install identical values without shipping ~30 MB of weights."""
import types
import zlib

import numpy as np


def _rng(path):
    return np.random.default_rng(zlib.crc32(path.encode()))


def install(model, prefix, wrap=lambda a: a):
    """Walk ``model`` (objects, lists) and overwrite every conv / linear / batch-norm parameter.
    ``wrap`` turns an ndarray into the holder's tensor type."""
    seen = []

    def shape_of(t):
        return tuple(t.shape)

    def visit(obj, path):
        if isinstance(obj, (list, tuple)):
            for i, o in enumerate(obj):
                visit(o, f"{path}.{i}")
            return
        if not hasattr(obj, "__dict__") or isinstance(obj, np.ndarray):
            return
        w = getattr(obj, "weight", None)
        if w is not None and hasattr(obj, "_mean"):                       # BatchNorm
            n = shape_of(w)[0]
            r = _rng(path)
            obj.weight = wrap(r.uniform(0.5, 1.5, n).astype(np.float32))
            obj.bias = wrap(r.uniform(-0.2, 0.2, n).astype(np.float32))
            obj._mean = wrap(r.uniform(-0.1, 0.1, n).astype(np.float32))
            obj._variance = wrap(r.uniform(0.5, 1.5, n).astype(np.float32))
            seen.append(path)
        elif w is not None:
            shp = shape_of(w)
            r = _rng(path)
            if len(shp) == 2:                                             # Linear [in,out]
                arr = (r.standard_normal(shp) / np.sqrt(shp[0])).astype(np.float32)
                nb = shp[1]
            else:                                                         # conv [out,in,1(,1)]
                arr = (r.standard_normal(shp[:2]) * np.sqrt(2.0 / shp[1])).astype(np.float32).reshape(shp)
                nb = shp[0]
            obj.weight = wrap(arr)
            if getattr(obj, "bias", None) is not None:
                obj.bias = wrap(r.uniform(-0.1, 0.1, nb).astype(np.float32))
            seen.append(path)
        for name, sub in list(vars(obj).items()):
            if name.startswith("_sub"):
                continue
            if name == "_modules" and isinstance(sub, dict):              # torch.nn.Module registry (the facade)
                for k, m in sub.items():
                    visit(m, f"{path}.{k}")
                continue
            if isinstance(sub, (type, types.FunctionType, types.MethodType, types.BuiltinFunctionType)):
                continue                                                  # dtypes, functions
            if isinstance(sub, (list, tuple)) or (hasattr(sub, "__dict__") and not isinstance(sub, np.ndarray)
                                                   and not callable(getattr(sub, "numpy", None))):
                visit(sub, f"{path}.{name}")

    visit(model, prefix)
    return seen
