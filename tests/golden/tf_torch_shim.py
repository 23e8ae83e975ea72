"""A torch-backed stand-in for the handful of TensorFlow ops the reference's legacy physics code uses, so that the
reference's OWN source (cut out of /root/reference by AST, never copied into this repository) can be executed in the
build container, where TensorFlow is not installable.  Used only by the golden-vector generators in this directory.

Semantics implemented (and nothing else): float32 tensors as torch CPU tensors; python scalars promote to the tensor's
dtype as in TF; tf.pad(mode='SYMMETRIC') with width <= 1 (= edge replication); tf.scatter_nd sums duplicates;
tf.math.divide_no_nan returns 0 where the denominator is 0; tf.reduce_* over axis lists with keepdims;
tf.keras.backend.epsilon() = 1e-7.
"""
import contextlib
import types

import numpy as np
import torch

float32 = torch.float32
float64 = torch.float64
int32 = torch.int32
int64 = torch.int64
bool_ = torch.bool


def _t(x, dtype=None):
    if isinstance(x, torch.Tensor):
        return x if dtype is None else x.to(dtype)
    return torch.as_tensor(np.asarray(x), dtype=dtype)


def constant(value, dtype=None, shape=None, name=None):
    t = _t(value, dtype)
    if shape is not None and tuple(shape) != tuple(t.shape):
        t = t.reshape(tuple(shape)) if t.numel() == int(np.prod(shape)) else t.expand(tuple(shape)).clone()
    return t


def convert_to_tensor(value, dtype=None):
    return _t(value, dtype)


def cast(x, dtype):
    return _t(x).to(dtype)


def ones_like(x, dtype=None):
    return torch.ones_like(_t(x), dtype=dtype)


def zeros_like(x, dtype=None):
    return torch.zeros_like(_t(x), dtype=dtype)


def expand_dims(x, axis):
    return _t(x).unsqueeze(axis)


def shape(x):
    return torch.tensor(list(_t(x).shape), dtype=torch.int32)


def stack(values, axis=0):
    return torch.stack([_t(v) for v in values], dim=axis)


def pad(x, paddings, mode="CONSTANT"):
    x = _t(x)
    p = [[int(a), int(b)] for a, b in np.asarray(paddings).tolist()]
    assert len(p) == x.dim()
    if mode.upper() != "SYMMETRIC":
        raise NotImplementedError(mode)
    for d, (lo, hi) in enumerate(p):
        assert lo in (0, 1) and hi in (0, 1), "SYMMETRIC pad wider than 1 is not needed by the reference"
        parts = []
        if lo:
            parts.append(x.narrow(d, 0, 1))
        parts.append(x)
        if hi:
            parts.append(x.narrow(d, x.shape[d] - 1, 1))
        x = torch.cat(parts, dim=d) if len(parts) > 1 else x
    return x


def scatter_nd(indices, updates, shape):
    idx = _t(indices).long()
    upd = _t(updates)
    out = torch.zeros(tuple(int(s) for s in np.asarray(shape).tolist()), dtype=upd.dtype)
    out.index_put_(tuple(idx[:, k] for k in range(idx.shape[1])), upd, accumulate=True)     # duplicates sum
    return out


def _axes(axis):
    if axis is None:
        return None
    return tuple(int(a) for a in (axis if isinstance(axis, (list, tuple)) else [axis]))


def reduce_mean(x, axis=None, keepdims=False):
    x = _t(x)
    return x.mean() if axis is None else x.mean(dim=_axes(axis), keepdim=keepdims)


def reduce_sum(x, axis=None, keepdims=False):
    x = _t(x)
    return x.sum() if axis is None else x.sum(dim=_axes(axis), keepdim=keepdims)


def where(cond, a, b):
    return torch.where(cond, _t(a), _t(b))


def maximum(a, b):
    return torch.maximum(_t(a), _t(b))


def minimum(a, b):
    return torch.minimum(_t(a), _t(b))


def function(*a, **k):
    if a and callable(a[0]):
        return a[0]
    return lambda f: f


def device(_name):
    return contextlib.nullcontext()


def print(*a, **k):        # noqa: A001  (tf.print)
    return None


math = types.SimpleNamespace(
    pow=lambda x, y: torch.pow(_t(x), y),
    log=lambda x: torch.log(_t(x)),
    reduce_prod=lambda x, axis=None: torch.prod(_t(x)) if axis is None else torch.prod(_t(x), dim=axis),
    divide_no_nan=lambda x, y: torch.where(_t(y) == 0, torch.zeros_like(_t(x) * _t(y)), _t(x) / torch.where(_t(y) == 0, torch.ones_like(_t(y)), _t(y))),
)
random = types.SimpleNamespace(normal=lambda **k: (_ for _ in ()).throw(NotImplementedError("tf.random.normal")))
keras = types.SimpleNamespace(backend=types.SimpleNamespace(epsilon=lambda: 1e-7))
