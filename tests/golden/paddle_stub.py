"""A NumPy-backed stand-in for the handful of ``paddle`` calls the reference's PointNet++ primitives
make (PAPC/models/layers/pointnet2_basic_layers.py:26-176, the forward passes :194-221, :243-281,
:296-330) -- TEST INFRASTRUCTURE for generating golden vectors, used by
tests/golden/make_golden_layers.py only.

Purpose: execute the REFERENCE'S OWN STATEMENTS (control flow, operation order, masks, sort / pad
logic, dtype round trips such as int64 indices stored in a float32 tensor) without PaddlePaddle,
which cannot be installed here.  Every stub op is the NumPy call of the same meaning, with Paddle's
dtype conventions: float32 default, int64 for ``arange`` / ``randint`` / ``argmax`` / ``argsort``,
Python scalars adopt the tensor's dtype (NumPy >= 2 does the same).  What this does NOT pin is the
arithmetic inside Paddle's kernels (e.g. the accumulation order of ``matmul``): vectors produced here
pin the oracle's *logic* to the reference's code; they are not outputs of Paddle itself.

Layers with parameters: the PointNet++ LAYER cases (make_golden_layers.py) build the reference's layers
with empty ``mlp`` lists, which exercises grouping, concat order, transposes and the max-pool of the
reference forward passes with no restated numerics at all.  The pillar and MODEL cases need
``nn.Linear`` / ``nn.Conv1D`` / ``nn.Conv2D`` (kernel 1) / ``nn.BatchNorm1D`` / ``nn.BatchNorm2D`` /
``nn.Dropout``; those ARE restated below (fp64 accumulation, Paddle 2.x BatchNorm semantics) and
labelled as such -- the vectors of those cases pin the reference's wiring around them, including which
BatchNorms follow ``eval()`` (attribute-registered ones) and which do not (the ones kept in lists).
"""
from __future__ import annotations

import types

import numpy as np

_next_randint = []          # values handed out by paddle.randint (the reference draws the FPS start)


def _unwrap(x):
    return x.a if isinstance(x, Tensor) else x


class Tensor:
    __array_priority__ = 100

    def __init__(self, a):
        self.a = np.asarray(a)

    # ---- introspection
    @property
    def shape(self):
        return list(self.a.shape)

    @property
    def dtype(self):
        return self.a.dtype

    def numpy(self):
        return self.a.copy()             # Tensor.numpy() copies in Paddle's dygraph mode

    # ---- shape ops (Paddle takes lists)
    def transpose(self, perm):
        return Tensor(self.a.transpose(perm))

    def reshape(self, shape):
        return Tensor(self.a.reshape(shape))

    def unsqueeze(self, axis):
        return Tensor(np.expand_dims(self.a, axis))

    def astype(self, dt):
        return Tensor(self.a.astype(dt))

    def tile(self, reps):
        return Tensor(np.tile(self.a, reps))

    def sort(self, axis=-1):
        return Tensor(np.sort(self.a, axis=axis, kind="stable"))

    def sum(self, axis=None, keepdim=False):
        return Tensor(self.a.sum(axis=axis, keepdims=keepdim, dtype=self.a.dtype))

    def squeeze(self, axis=None):
        return Tensor(np.squeeze(self.a, axis=axis))

    def t(self):
        return Tensor(self.a.T)

    def any(self):
        return Tensor(np.array([self.a.any()]))      # Paddle 2.0: a 1-element tensor

    # ---- indexing
    def __getitem__(self, idx):
        return Tensor(self.a[idx])

    def __setitem__(self, idx, value):
        self.a[idx] = _unwrap(value)     # casts to the destination dtype, as Paddle's set_value does

    # ---- arithmetic / comparisons
    def _bin(self, other, fn):
        return Tensor(fn(self.a, _unwrap(other)))

    def __add__(self, o): return self._bin(o, np.add)
    def __radd__(self, o): return Tensor(np.add(_unwrap(o), self.a))
    def __sub__(self, o): return self._bin(o, np.subtract)
    def __rsub__(self, o): return Tensor(np.subtract(_unwrap(o), self.a))
    def __mul__(self, o): return self._bin(o, np.multiply)
    def __rmul__(self, o): return Tensor(np.multiply(_unwrap(o), self.a))
    def __truediv__(self, o): return self._bin(o, np.divide)
    def __rtruediv__(self, o): return Tensor(np.divide(_unwrap(o), self.a))
    def __pow__(self, o): return self._bin(o, np.power)
    def __neg__(self): return Tensor(-self.a)
    def __lt__(self, o): return self._bin(o, np.less)
    def __gt__(self, o): return self._bin(o, np.greater)
    def __eq__(self, o): return self._bin(o, np.equal)          # noqa: PLW1641 (tensors are not hashed here)

    def __iadd__(self, o):
        self.a = np.add(self.a, _unwrap(o)).astype(self.a.dtype)
        return self

    def __imul__(self, o):
        self.a = np.multiply(self.a, _unwrap(o)).astype(self.a.dtype)
        return self

    def __isub__(self, o):
        self.a = np.subtract(self.a, _unwrap(o)).astype(self.a.dtype)
        return self


def to_tensor(x):
    a = np.array(_unwrap(x))              # a NumPy array keeps its dtype; Python ints become int64,
    if not isinstance(_unwrap(x), np.ndarray) and a.dtype == np.float64:
        a = a.astype(np.float32)          # Python floats the default float32
    return Tensor(a)


def matmul(x, y):
    return Tensor(np.matmul(_unwrap(x), _unwrap(y)))


def sum(x, axis=None, keepdim=False):     # noqa: A001 (paddle.sum)
    return x.sum(axis=axis, keepdim=keepdim)


def max(x, axis=None, keepdim=False):     # noqa: A001 (paddle.max)
    return Tensor(_unwrap(x).max(axis=axis, keepdims=keepdim))


def argmax(x, axis=None):
    return Tensor(np.argmax(_unwrap(x), axis=axis).astype(np.int64))   # first maximum, like paddle.argmax


def sort(x, axis=-1):
    return x.sort(axis=axis)


def argsort(x, axis=-1):
    return Tensor(np.argsort(_unwrap(x), axis=axis, kind="stable").astype(np.int64))


def tile(x, reps):
    return x.tile(reps)


def arange(n, dtype="int64"):
    return Tensor(np.arange(n, dtype=dtype))


def zeros(shape, dtype="float32"):
    return Tensor(np.zeros(shape, dtype=dtype))


def zeros_like(x):
    return Tensor(np.zeros_like(_unwrap(x)))


def unsqueeze(x, axis):
    return x.unsqueeze(axis)


def norm(x, p, axis, keepdim=False):
    assert p == 2
    a = _unwrap(x)
    return Tensor(np.sqrt((a * a).sum(axis=axis, keepdims=keepdim, dtype=a.dtype)))


def index_select(x, index, axis=0):
    return Tensor(np.take(_unwrap(x), _unwrap(index), axis=axis))


def stack(xs, axis=0):
    return Tensor(np.stack([_unwrap(x) for x in xs], axis=axis))


def ones(shape):
    return Tensor(np.ones(shape, dtype=np.float32))


def concat(xs, axis=0):
    return Tensor(np.concatenate([_unwrap(x) for x in xs], axis=axis))


def randint(low, high, shape):
    v = np.asarray(_next_randint.pop(0), dtype=np.int64)
    assert tuple(v.shape) == tuple(shape) and v.min() >= low and v.max() < high
    return Tensor(v)


_param_rng = np.random.default_rng(2024)   # initial weights of the restated layers (the generators overwrite
                                            # or export every parameter, so only reproducibility matters)


class _Layer:
    """paddle.nn.Layer as far as the reference uses it: sublayers assigned as ATTRIBUTES are registered and
    follow train() / eval(); layers kept in plain Python lists are not (which is why the reference's
    SetAbstraction / FeaturePropagation BatchNorms stay in training mode after model.eval())."""

    def __init__(self, *a, **k):
        object.__setattr__(self, "_sub", {})
        object.__setattr__(self, "training", True)

    def __setattr__(self, name, value):
        if isinstance(value, _Layer):
            self._sub[name] = value
        object.__setattr__(self, name, value)

    def train(self, mode=True):
        object.__setattr__(self, "training", mode)
        for sub in self._sub.values():
            sub.train(mode)
        return self

    def eval(self):
        return self.train(False)

    def __call__(self, *a, **k):
        return self.forward(*a, **k)


def _missing(name):
    def ctor(*a, **k):
        raise NotImplementedError(f"paddle.nn.{name} is not part of the stub (see the module docstring)")
    return ctor


# ---- layers with parameters: RESTATED numerics (fp64 accumulation, fp32 results), labelled as such in the
#      generators' docstrings; what the cases that use them pin is the reference's wiring around them.
class _Linear(_Layer):
    """paddle.nn.Linear: x @ W [+ b], W [in,out]."""

    def __init__(self, in_features, out_features, bias_attr=True):
        super().__init__()
        self.weight = Tensor((_param_rng.standard_normal((in_features, out_features)) / np.sqrt(in_features))
                             .astype(np.float32))
        self.bias = Tensor(np.zeros((out_features,), np.float32)) if bias_attr else None
        self.inputs = []                  # every input is recorded for the golden files

    def forward(self, x):
        self.inputs.append(x.numpy())
        y = x.a.astype(np.float64) @ self.weight.a.astype(np.float64)
        if self.bias is not None:
            y = y + self.bias.a.astype(np.float64)
        return Tensor(y.astype(np.float32))


class _ConvK1(_Layer):
    """paddle.nn.Conv1D / Conv2D with kernel size 1: weight [out,in,1(,1)], bias [out]; channel axis 1."""

    def __init__(self, in_channels, out_channels, kernel_size, nd):
        super().__init__()
        assert kernel_size == 1
        shape = (out_channels, in_channels) + (1,) * nd
        self.weight = Tensor((_param_rng.standard_normal(shape) * np.sqrt(2.0 / in_channels)).astype(np.float32))
        self.bias = Tensor(np.zeros((out_channels,), np.float32))

    def forward(self, x):
        a = x.a
        w = self.weight.a.reshape(self.weight.a.shape[0], -1).astype(np.float64)
        flat = np.ascontiguousarray(a).reshape(a.shape[0], a.shape[1], -1).astype(np.float64)
        y = np.matmul(w, flat) + self.bias.a.astype(np.float64)[None, :, None]
        return Tensor(y.reshape((a.shape[0], w.shape[0]) + a.shape[2:]).astype(np.float32))


class _BatchNorm(_Layer):
    """paddle.nn.BatchNorm1D / BatchNorm2D, channel axis 1 of a 2-, 3- or 4-D input: training mode normalises
    with the biased batch variance and moves ``_mean`` / ``_variance`` (momentum * running + (1 - momentum) *
    batch); eval mode uses the running statistics."""

    def __init__(self, num_features, momentum=0.9, epsilon=1e-5):
        super().__init__()
        self.weight = Tensor(np.ones((num_features,), np.float32))
        self.bias = Tensor(np.zeros((num_features,), np.float32))
        self._mean = Tensor(np.zeros((num_features,), np.float32))
        self._variance = Tensor(np.ones((num_features,), np.float32))
        self._momentum, self._epsilon = momentum, epsilon

    def forward(self, x):
        a = x.a.astype(np.float64)
        axes = tuple(i for i in range(a.ndim) if i != 1)
        shp = [1, -1] + [1] * (a.ndim - 2)
        if self.training:
            mean, var = a.mean(axis=axes), a.var(axis=axes)
            m = self._momentum
            self._mean = Tensor((m * self._mean.a + (1 - m) * mean).astype(np.float32))
            self._variance = Tensor((m * self._variance.a + (1 - m) * var).astype(np.float32))
        else:
            mean, var = self._mean.a.astype(np.float64), self._variance.a.astype(np.float64)
        y = (a - mean.reshape(shp)) / np.sqrt(var.reshape(shp) + np.float64(self._epsilon))
        y = y * self.weight.a.astype(np.float64).reshape(shp) + self.bias.a.astype(np.float64).reshape(shp)
        return Tensor(y.astype(np.float32))


class _Dropout(_Layer):
    """paddle.nn.Dropout (upscale_in_train): identity in eval mode or with p == 0; a random mask otherwise,
    which no golden case uses."""

    def __init__(self, p=0.5):
        super().__init__()
        self.p = p

    def forward(self, x):
        if not self.training or self.p == 0:
            return x
        raise NotImplementedError("Dropout with p > 0 in training mode is not reproducible outside Paddle")


nn = types.ModuleType("paddle.nn")
nn.Layer = _Layer
nn.LayerList = list            # (Paddle registers the members of a LayerList; the pillar cases never call eval())
nn.Linear = _Linear
nn.BatchNorm1D = _BatchNorm
nn.BatchNorm2D = _BatchNorm
nn.Conv1D = lambda i, o, k: _ConvK1(i, o, k, 1)
nn.Conv2D = lambda i, o, k: _ConvK1(i, o, k, 2)
nn.Dropout = _Dropout
functional = types.ModuleType("paddle.nn.functional")
functional.relu = lambda x: Tensor(np.maximum(_unwrap(x), 0))
nn.functional = functional


def install():
    """Register the stub as ``paddle`` / ``paddle.nn`` / ``paddle.nn.functional`` in sys.modules."""
    import sys
    me = sys.modules[__name__]
    sys.modules["paddle"] = me
    sys.modules["paddle.nn"] = nn
    sys.modules["paddle.nn.functional"] = functional
    return me
