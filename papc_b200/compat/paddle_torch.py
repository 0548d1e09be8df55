"""A torch-backed ``paddle`` facade: the ~40 paddle names the reference's PointNet++ files use
(PAPC/models/layers/pointnet2_basic_layers.py, PAPC/models/{classify,segment}/pointnet2/pointnet2.py), so that
those files run UNCHANGED in an environment without PaddlePaddle:

  * with ``PAPC.models.layers`` bound to ``papc_b200.layers`` (``papc_b200.compat.install()``), the reference's
    model definitions call this library's sm_100a kernels; only their heads (Linear / Conv1D / BatchNorm1D /
    Dropout over a few rows -- library work) run on the torch ops below;
  * with ``PAPC.models.layers`` bound to the reference's own layers file and the device set to 'cpu', the whole
    reference path runs on the host cores (bench.py's reference arm).

Paddle conventions kept: float32 default dtype, int64 indices, ``Tensor.transpose(perm)``, ``Tensor.numpy()``
returns a copy, ``paddle.max / sort`` return values only, ``nn.Linear.weight`` is [in,out], BatchNorm uses the
biased batch variance for normalisation AND for the running ``_variance`` with momentum 0.9 / epsilon 1e-5, only
sublayers assigned as attributes are registered (``eval()`` does not reach layers kept in Python lists).
This module is plumbing, not a re-implementation of Paddle: anything else raises AttributeError.
"""
from __future__ import annotations

import builtins
import types

import numpy as np
import torch

_DEVICE = [None]          # None = cuda when available (the product), else cpu
_RANDINT_QUEUE = []       # values handed out by paddle.randint before falling back to torch.randint


def set_device(name):
    """paddle.set_device('cpu' | 'gpu' | 'gpu:0')."""
    name = str(name)
    _DEVICE[0] = torch.device("cpu") if name.startswith("cpu") else torch.device(name.replace("gpu", "cuda"))
    return _DEVICE[0]


def get_device():
    if _DEVICE[0] is None:
        return torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    return _DEVICE[0]


_DT = {"float32": torch.float32, "float64": torch.float64, "int64": torch.int64, "int32": torch.int32,
       "bool": torch.bool, "float16": torch.float16, "uint8": torch.uint8}


def _dtype(dt):
    if isinstance(dt, torch.dtype):
        return dt
    if isinstance(dt, str):
        return _DT[dt]
    return _DT[np.dtype(dt).name]


class Tensor(torch.Tensor):
    """torch.Tensor with the Paddle spellings the reference uses."""

    @staticmethod
    def __new__(cls, data):
        return torch.Tensor._make_subclass(cls, data.detach() if isinstance(data, torch.Tensor) else torch.as_tensor(data))

    def transpose(self, *perm):
        if len(perm) == 1 and isinstance(perm[0], (list, tuple)):
            return self.permute(*perm[0])
        return super().transpose(*perm)

    def astype(self, dt):
        return self.to(_dtype(dt))

    def numpy(self):
        return self.detach().cpu().as_subclass(torch.Tensor).numpy().copy()

    def sort(self, axis=-1, descending=False):            # values only, stable
        return torch.sort(self, dim=axis, descending=descending, stable=True)[0]

    def sum(self, axis=None, keepdim=False, **kw):
        if "dim" in kw:
            axis = kw["dim"]
        if axis is None:
            return super().sum()
        return super().sum(dim=axis, keepdim=keepdim)

    def unsqueeze(self, axis):
        return super().unsqueeze(axis)

    def tile(self, reps):
        return super().tile(tuple(reps))


def _wrap(t):
    return t if isinstance(t, Tensor) else t.as_subclass(Tensor)


def to_tensor(x, dtype=None, place=None, stop_gradient=True):
    if isinstance(x, torch.Tensor):
        t = x
    else:
        a = np.asarray(x)
        if not isinstance(x, np.ndarray) and a.dtype == np.float64:
            a = a.astype(np.float32)              # Python floats -> the default float32
        t = torch.from_numpy(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(_dtype(dtype))
    dev = get_device()
    if t.device != dev:
        t = t.to(dev)
    return _wrap(t)


def matmul(x, y):
    return _wrap(torch.matmul(x, y))


def sum(x, axis=None, keepdim=False):                     # noqa: A001
    return _wrap(x).sum(axis=axis, keepdim=keepdim)


def max(x, axis=None, keepdim=False):                     # noqa: A001
    if axis is None:
        return _wrap(torch.max(x))
    return _wrap(torch.max(x, dim=axis, keepdim=keepdim)[0])


def argmax(x, axis=None):
    return _wrap(torch.argmax(x, dim=axis))


def sort(x, axis=-1):
    return _wrap(x).sort(axis=axis)


def argsort(x, axis=-1):
    return _wrap(torch.sort(x, dim=axis, stable=True)[1])


def tile(x, reps):
    return _wrap(x).tile(reps)


def arange(n, dtype="int64"):
    return _wrap(torch.arange(n, dtype=_dtype(dtype), device=get_device()))


def zeros(shape, dtype="float32"):
    return _wrap(torch.zeros(tuple(shape), dtype=_dtype(dtype), device=get_device()))


def ones(shape, dtype="float32"):
    return _wrap(torch.ones(tuple(shape), dtype=_dtype(dtype), device=get_device()))


def zeros_like(x):
    return _wrap(torch.zeros_like(x))


def concat(xs, axis=0):
    return _wrap(torch.cat([_wrap(x) for x in xs], dim=axis))


def stack(xs, axis=0):
    return _wrap(torch.stack(list(xs), dim=axis))


def unsqueeze(x, axis):
    return _wrap(x).unsqueeze(axis)


def reshape(x, shape):
    return _wrap(x).reshape(list(shape))


def transpose(x, perm):
    return _wrap(x).transpose(list(perm))


def randint(low, high, shape):
    if _RANDINT_QUEUE:
        v = torch.as_tensor(np.asarray(_RANDINT_QUEUE.pop(0)), dtype=torch.int64)
        assert tuple(v.shape) == tuple(shape)
        return _wrap(v.to(get_device()))
    return _wrap(torch.randint(low, high, tuple(shape), dtype=torch.int64, device=get_device()))


def queue_randint(values):
    """Seeded draws for the next ``paddle.randint`` calls (the reference draws the FPS start there, layers.py:76)."""
    _RANDINT_QUEUE.extend(values)


def no_grad():
    return torch.no_grad()


# ------------------------------------------------------------------------------------------------ paddle.nn
class Layer(torch.nn.Module):
    """paddle.nn.Layer as the reference uses it (torch.nn.Module has the same registration rule)."""

    def __init__(self, name_scope=None, dtype="float32"):
        super().__init__()

    def __setattr__(self, name, value):
        # Paddle code assigns plain tensors over parameters (``layer.weight = paddle.to_tensor(...)``)
        params = self.__dict__.get("_parameters")
        if (params is not None and name in params and isinstance(value, torch.Tensor)
                and not isinstance(value, torch.nn.Parameter)):
            value = torch.nn.Parameter(value.detach().as_subclass(torch.Tensor), requires_grad=False)
        super().__setattr__(name, value)

    def sublayers(self, include_self=False):
        mods = list(self.modules())
        return mods if include_self else mods[1:]


class Linear(Layer):
    """paddle.nn.Linear: x @ W + b with W [in,out]."""

    def __init__(self, in_features, out_features, weight_attr=None, bias_attr=None, name=None):
        super().__init__()
        w = torch.empty(in_features, out_features)
        torch.nn.init.xavier_uniform_(w)                                  # Paddle's default initialiser
        self.weight = torch.nn.Parameter(w.to(get_device()))
        self.bias = None if bias_attr is False else torch.nn.Parameter(torch.zeros(out_features, device=get_device()))

    def forward(self, x):
        y = torch.matmul(x, self.weight)
        return _wrap(y + self.bias if self.bias is not None else y)


class _ConvK1(Layer):
    """paddle.nn.Conv1D / Conv2D, kernel size 1 (all the reference uses): weight [out,in,1(,1)], channel axis 1.
    Evaluated as an fp32 matmul (no TF32), the library-GEMM form of a 1x1 convolution."""

    def __init__(self, in_channels, out_channels, kernel_size, nd):
        super().__init__()
        if kernel_size not in (1, (1,), (1, 1), [1], [1, 1]):
            raise NotImplementedError("the facade only carries the kernel-size-1 convolutions the reference uses")
        w = torch.empty((out_channels, in_channels) + (1,) * nd)
        torch.nn.init.kaiming_uniform_(w, a=5 ** 0.5)
        self.weight = torch.nn.Parameter(w.to(get_device()))
        self.bias = torch.nn.Parameter(torch.zeros(out_channels, device=get_device()))

    def forward(self, x):
        shp = x.shape
        w = self.weight.reshape(self.weight.shape[0], -1)
        y = torch.matmul(w, x.reshape(shp[0], shp[1], -1)) + self.bias.reshape(1, -1, 1)
        return _wrap(y.reshape((shp[0], w.shape[0]) + tuple(shp[2:])))


class Conv1D(_ConvK1):
    def __init__(self, in_channels, out_channels, kernel_size, **kw):
        super().__init__(in_channels, out_channels, kernel_size, 1)


class Conv2D(_ConvK1):
    def __init__(self, in_channels, out_channels, kernel_size, **kw):
        super().__init__(in_channels, out_channels, kernel_size, 2)


class _BatchNorm(Layer):
    """paddle.nn.BatchNorm1D / BatchNorm2D (channel axis 1 of a 2-, 3- or 4-D input)."""

    def __init__(self, num_features, momentum=0.9, epsilon=1e-5, **kw):
        super().__init__()
        dev = get_device()
        self.weight = torch.nn.Parameter(torch.ones(num_features, device=dev))
        self.bias = torch.nn.Parameter(torch.zeros(num_features, device=dev))
        self.register_buffer("_mean", torch.zeros(num_features, device=dev))
        self.register_buffer("_variance", torch.ones(num_features, device=dev))
        self._momentum, self._epsilon = momentum, epsilon

    def forward(self, x):
        axes = [i for i in range(x.dim()) if i != 1]
        shp = [1, -1] + [1] * (x.dim() - 2)
        if self.training:
            var, mean = torch.var_mean(x, dim=axes, unbiased=False)
            with torch.no_grad():
                self._mean.mul_(self._momentum).add_(mean, alpha=1.0 - self._momentum)
                self._variance.mul_(self._momentum).add_(var, alpha=1.0 - self._momentum)
        else:
            mean, var = self._mean, self._variance
        scale = self.weight * torch.rsqrt(var + self._epsilon)
        return _wrap(x * scale.reshape(shp) + (self.bias - mean * scale).reshape(shp))


class BatchNorm1D(_BatchNorm):
    pass


class BatchNorm2D(_BatchNorm):
    pass


class Dropout(Layer):
    """paddle.nn.Dropout (mode 'upscale_in_train' = torch's convention)."""

    def __init__(self, p=0.5, **kw):
        super().__init__()
        self.p = p

    def forward(self, x):
        return _wrap(torch.nn.functional.dropout(x, self.p, self.training))


class ReLU(Layer):
    def forward(self, x):
        return _wrap(torch.relu(x))


nn = types.ModuleType("paddle.nn")
for _n, _v in dict(Layer=Layer, LayerList=torch.nn.ModuleList, Sequential=torch.nn.Sequential, Linear=Linear,
                   Conv1D=Conv1D, Conv2D=Conv2D, BatchNorm1D=BatchNorm1D, BatchNorm2D=BatchNorm2D, Dropout=Dropout,
                   ReLU=ReLU).items():
    setattr(nn, _n, _v)
functional = types.ModuleType("paddle.nn.functional")
functional.relu = lambda x: _wrap(torch.relu(x))
functional.softmax = lambda x, axis=-1: _wrap(torch.softmax(x, dim=axis))
functional.log_softmax = lambda x, axis=-1: _wrap(torch.log_softmax(x, dim=axis))
nn.functional = functional
__version__ = "2.0.0-papc_b200-facade"
del builtins
