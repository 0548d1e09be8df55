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

Convolution layers (``nn.Conv2D``, ``nn.BatchNorm2D`` ...) are deliberately NOT provided: the PointNet++
golden cases build the reference's layers with empty ``mlp`` lists, which exercises grouping, concat
order, transposes and the max-pool of the reference forward passes with no restated numerics.  The
pillar cases need ``nn.Linear`` / ``nn.BatchNorm1D``; those two ARE restated here (fp64 accumulation)
and labelled as such -- the pillar vectors pin the reference's wiring around them.
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


class _Layer:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self.forward(*a, **k)


def _missing(name):
    def ctor(*a, **k):
        raise NotImplementedError(f"paddle.nn.{name} is not part of the stub (see the module docstring)")
    return ctor


class _Linear(_Layer):
    """paddle.nn.Linear restated (x @ W [+ b], W [in,out]; fp64 accumulation, fp32 result) -- used by the
    PILLAR golden cases only, and recorded there as restated numerics: what those cases pin is the
    reference's wiring around it (decorations, mask, max / tile / concat)."""

    def __init__(self, in_features, out_features, bias_attr=True):
        self.weight = Tensor(np.zeros((in_features, out_features), np.float32))
        self.bias = Tensor(np.zeros((out_features,), np.float32)) if bias_attr else None
        self.inputs = []                  # every input is recorded for the golden file

    def forward(self, x):
        self.inputs.append(x.numpy())
        y = x.a.astype(np.float64) @ self.weight.a.astype(np.float64)
        if self.bias is not None:
            y = y + self.bias.a.astype(np.float64)
        return Tensor(y.astype(np.float32))


class _BatchNorm1D(_Layer):
    """paddle.nn.BatchNorm1D restated, training mode, input [N,C,L]: biased batch variance over (N, L)."""

    def __init__(self, num_features, momentum=0.9, epsilon=1e-5):
        self.weight = Tensor(np.ones((num_features,), np.float32))
        self.bias = Tensor(np.zeros((num_features,), np.float32))
        self._epsilon = epsilon

    def forward(self, x):
        a = x.a.astype(np.float64)
        mean, var = a.mean(axis=(0, 2)), a.var(axis=(0, 2))
        y = (a - mean[None, :, None]) / np.sqrt(var[None, :, None] + np.float64(self._epsilon))
        y = y * self.weight.a.astype(np.float64)[None, :, None] + self.bias.a.astype(np.float64)[None, :, None]
        return Tensor(y.astype(np.float32))


nn = types.ModuleType("paddle.nn")
nn.Layer = _Layer
nn.LayerList = list
nn.Linear = _Linear
nn.BatchNorm1D = _BatchNorm1D
for _n in ("Conv1D", "Conv2D", "BatchNorm2D", "Dropout"):
    setattr(nn, _n, _missing(_n))
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
