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
    out.index_put_(tuple(idx[:, k] for k in builtins_range(idx.shape[1])), upd, accumulate=True)     # duplicates sum
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
    return torch.where(_t(cond).to(torch.bool) if not isinstance(cond, torch.Tensor) else cond, _t(a), _t(b))


def _pair(a, b):
    """python scalars take the tensor operand's dtype (TF's promotion of weak scalars)"""
    ta, tb = isinstance(a, torch.Tensor), isinstance(b, torch.Tensor)
    if ta and not tb:
        return a, torch.as_tensor(np.asarray(b), dtype=a.dtype)
    if tb and not ta:
        return torch.as_tensor(np.asarray(a), dtype=b.dtype), b
    return _t(a), _t(b)


def maximum(a, b):
    """value of tf.maximum; gradient as TensorFlow's MaximumGrad routes it: to x where x >= y (ties go to the FIRST
    argument), to y elsewhere -- torch.maximum would split a tie 50/50"""
    a, b = _pair(a, b)
    return torch.where(a >= b, a, b)


def minimum(a, b):
    """tf.minimum; MinimumGrad: to x where x <= y (ties to the first argument)"""
    a, b = _pair(a, b)
    return torch.where(a <= b, a, b)


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


# ---- additions for polyhm_splines.py / PVT_Layer_Subclassed.py / relative_permeability.py ----------------------------
class _Shape(tuple):
    @property
    def rank(self):
        return len(self)


class _T(torch.Tensor):
    """tensor whose .shape carries .rank (TensorShape), which the spline layer's constructor reads"""

    @property
    def shape(self):          # noqa: D401
        return _Shape(super().shape)


def convert_to_tensor(value, dtype=None):       # noqa: F811
    return _t(value, dtype).clone().as_subclass(_T)


class _SqrtTF(torch.autograd.Function):
    """correctly rounded sqrt (numpy; torch-CPU's vectorised sqrt is not), TF's SqrtGrad: (0.5 * dy) / y"""

    @staticmethod
    def forward(ctx, x):
        y = torch.from_numpy(np.sqrt(x.detach().numpy()))
        ctx.save_for_backward(y)
        return y

    @staticmethod
    def backward(ctx, g):
        (y,) = ctx.saved_tensors
        return (0.5 * g) / y


def sqrt(x):
    return _SqrtTF.apply(_t(x).as_subclass(torch.Tensor))


def square(x):
    x = _t(x)
    return x * x


def exp(x):
    return torch.exp(_t(x))


def reshape(x, shape_):
    sh = [int(s) for s in (shape_.tolist() if isinstance(shape_, torch.Tensor) else shape_)]
    return _t(x).reshape(sh)


def transpose(x, perm=None):
    x = _t(x)
    return x.permute(*perm) if perm is not None else x.t()


def unstack(x, num=None, axis=0):
    return [int(v) for v in _t(x).tolist()] if _t(x).dim() == 1 else list(torch.unbind(_t(x), dim=axis))


def eye(n, batch_shape=None, dtype=torch.float32):
    e = torch.eye(int(n), dtype=dtype)
    return e.expand(*[int(b) for b in batch_shape], int(n), int(n)).clone() if batch_shape else e


def ones(shape_, dtype=torch.float32):
    return torch.ones([int(s) for s in shape_], dtype=dtype)


def zeros(shape_, dtype=torch.float32):
    return torch.zeros([int(s) for s in shape_], dtype=dtype)


def concat(values, axis):
    return torch.cat([_t(v) for v in values], dim=axis)


def tile(x, multiples):
    return _t(x).repeat(*[int(m) for m in multiples])


def rank(x):
    return _t(x).dim()


def reduce_prod(x, axis=None):
    x = _t(x)
    return int(torch.prod(x)) if axis is None else torch.prod(x, dim=axis)


def matmul(a, b, transpose_b=False):
    """batched matmul with the inner index accumulated SEQUENTIALLY in fp32, one rounded multiply and one rounded add
    per term (no FMA): the accumulation order this repository's oracle pins where TensorFlow leaves it open"""
    a, b = _t(a), _t(b)
    if transpose_b:
        b = b.transpose(-1, -2)
    n = a.shape[-1]
    assert b.shape[-2] == n
    acc = a[..., :, 0:1] * b[..., 0:1, :]
    for i in range(1, n):
        acc = acc + a[..., :, i:i + 1] * b[..., i:i + 1, :]
    return acc


def pow(x, y):            # noqa: A001
    return torch.pow(_t(x), y)


def clip_by_value(x, lo, hi):
    """tf.clip_by_value; ClipByValueGrad: 1 to x on the closed interval [lo, hi], to lo where x < lo, to hi where x > hi"""
    x = _t(x)
    lo, hi = _t(lo, x.dtype), _t(hi, x.dtype)
    return torch.where(x < lo, lo, torch.where(x > hi, hi, x))


linalg = types.SimpleNamespace(solve=lambda a, b: torch.linalg.solve(_t(a), _t(b)))


class GradientTape:
    def __init__(self, persistent=False):
        self._watched = []

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def watch(self, x):
        x.requires_grad_(True)

    def gradient(self, y, x, unconnected_gradients=None):
        if isinstance(x, (list, tuple)):
            if len(x) == 0:
                return []
            if not y.requires_grad:
                return [torch.zeros_like(v) for v in x]
            g = torch.autograd.grad(y, list(x), grad_outputs=torch.ones_like(y), retain_graph=True, allow_unused=True)
            return [torch.zeros_like(v) if gi is None else gi for gi, v in zip(g, x)]
        # the derivative stays differentiable, as TensorFlow's nested tapes record it (PVTLayer's inner tape under the
        # loss tape; the Newton iterations of the blocking-factor integral, whose cost depends on the enclosing graph
        # even where the watched iterate is a fresh tensor)
        return torch.autograd.grad(y, x, grad_outputs=torch.ones_like(y), retain_graph=True, create_graph=True)[0]


class _Layer:
    def __init__(self, name=None, dtype=None, **kwargs):
        self.name = name
        self._built = False

    def build(self, input_shape):
        self._built = True

    def __call__(self, *a, **k):
        if not self._built:
            self.build(getattr(a[0], "shape", None) if a else None)
        return self.call(*a, **k)


keras.layers = types.SimpleNamespace(Layer=_Layer)
keras.initializers = types.SimpleNamespace(Constant=lambda v: v)
math.exp = exp
math.sqrt = sqrt
math.maximum = maximum
math.minimum = minimum
math.square = square


# ---- additions for welldata_processor.py / well_rate_bhp_Subclassed.py ------------------------------------------------
Tensor = torch.Tensor
string = object()


def size(x):
    return torch.tensor(_t(x).numel(), dtype=torch.int32)


def equal(a, b):
    if isinstance(a, str) or isinstance(b, str):
        return torch.tensor(a == b)
    return _t(a) == _t(b)


def less(a, b):
    return _t(a) < _t(b)


def greater(a, b):
    return _t(a) > _t(b)


def logical_and(a, b):
    return _t(a) & _t(b)


def logical_or(a, b):
    return _t(a) | _t(b)


def logical_not(a):
    return ~_t(a)


def reduce_any(x, axis=None):
    x = _t(x)
    return x.any() if axis is None else x.any(dim=axis)


def cond(pred, true_fn, false_fn):
    return true_fn() if _b.bool(pred) else false_fn()


def fill(shape_, value):
    return torch.full([int(s) for s in (shape_.tolist() if isinstance(shape_, torch.Tensor) else shape_)], float(value), dtype=torch.float32)


def range(*a, dtype=None):            # noqa: A001
    return torch.arange(*[int(v) for v in a], dtype=dtype or torch.int32)


def gather(params, indices, axis=0):
    idx = _t(indices).long()
    return _t(params).index_select(axis, idx.reshape(-1)).reshape(*_t(params).shape[:axis], *idx.shape, *_t(params).shape[axis + 1:]) \
        if idx.dim() != 0 else _t(params).select(axis, int(idx))


def gather_nd(params, indices):
    idx = _t(indices).long()
    return _t(params)[tuple(idx[..., k] for k in builtins_range(idx.shape[-1]))]


def tensor_scatter_nd_update(tensor, indices, updates):
    out = _t(tensor).clone()
    idx = _t(indices).long()
    out[tuple(idx[:, k] for k in builtins_range(idx.shape[1]))] = _t(updates).to(out.dtype)       # last writer wins
    return out


def linspace(start, stop, num):
    """tf.linspace with tensor end points along a new leading axis: start + delta*i for the interior, the END POINTS
    THEMSELVES at both ends (math_ops.linspace_nd concatenates (start, interior, stop))"""
    start, stop = torch.broadcast_tensors(_t(start), _t(stop))
    num = int(num)
    delta = (stop - start) / float(num - 1)
    interior = [start + delta * float(i) for i in builtins_range(1, num - 1)]
    return torch.stack([start] + interior + [stop], dim=0)


class TensorArray:
    def __init__(self, dtype=None, size=0, dynamic_size=False, clear_after_read=True):
        self._items = {}

    def write(self, i, v):
        self._items[int(i)] = v
        return self

    def stack(self):
        return torch.stack([self._items[k] for k in sorted(self._items)]) if self._items else torch.zeros(0)


import builtins as _b      # noqa: E402
builtins_range = _b.range
math.cumsum = lambda x, axis=0: torch.cumsum(_t(x), dim=axis)
debugging = types.SimpleNamespace(assert_shapes=lambda *a, **k: None, assert_equal=lambda *a, **k: None,
                                  assert_none_equal=lambda *a, **k: None, assert_greater_equal=lambda *a, **k: None,
                                  assert_less_equal=lambda *a, **k: None)
bool = torch.bool


def reduce_max(x, axis=None, keepdims=False):
    x = _t(x)
    return x.max() if axis is None else x.amax(dim=axis, keepdim=keepdims)


def abs(x):                # noqa: A001
    return torch.abs(_t(x))


strings = types.SimpleNamespace(format=lambda template, inputs=(), **k: str(template))      # log lines: never compared


def while_loop(cond, body, loop_vars, shape_invariants=None, **k):
    v = list(loop_vars)
    while _b.bool(cond(*v)):
        v = list(body(*v))
    return v


def less_equal(a, b):
    return _t(a) <= _t(b)


def greater_equal(a, b):
    return _t(a) >= _t(b)


torch.Tensor.get_shape = lambda self: tuple(self.shape)     # only read to fill tf.while_loop's shape_invariants


# ---- additions for DataSummary.nonormalize / normalize_diff (data_processing/data_processing_utils.py) -------------
def not_equal(a, b):
    return _t(a) != _t(b)


def broadcast_to(x, shape_):
    return _t(x).expand(*[int(s) for s in (shape_.tolist() if isinstance(shape_, torch.Tensor) else shape_)])


math.is_nan = lambda x: torch.isnan(_t(x))
math.is_inf = lambda x: torch.isinf(_t(x))
DType = object


# ---- additions for Hard_Layer_Subclassed.py ------------------------------------------------------------------------
def _add_weight(self, shape=None, initializer=None, constraint=None, trainable=True, name=None, **k):
    v = initializer(shape) if callable(initializer) else torch.full(tuple(int(s) for s in shape), float(initializer))
    return v.clone().requires_grad_(_b.bool(trainable))


_Layer.add_weight = _add_weight
constant_initializer = lambda value=0.0: (lambda shape: torch.full(tuple(int(s) for s in shape), float(value[0] if isinstance(value, (tuple, list)) else value)))
keras.constraints = types.SimpleNamespace(MinMaxNorm=lambda **k: None, UnitNorm=lambda **k: None)
keras.initializers.get = lambda name: None
nn = types.SimpleNamespace(sigmoid=torch.sigmoid, relu=torch.relu, tanh=torch.tanh)


class _Dense(_Layer):
    """tf.keras.layers.Dense on the last axis: activation(x @ kernel + bias); the generator sets kernel / bias"""

    def __init__(self, units, activation=None, kernel_initializer=None, kernel_constraint=None, name=None, **k):
        super().__init__(name=name)
        self.units, self.activation = int(units), activation
        self.kernel = self.bias = None

    def call(self, x):
        if self.kernel is None:
            self.kernel = torch.ones(x.shape[-1], self.units).requires_grad_(True)
            self.bias = torch.zeros(self.units).requires_grad_(True)
        y = matmul(x, self.kernel) + self.bias
        return self.activation(y) if self.activation is not None else y


keras.layers.Dense = _Dense
math.reduce_sum = reduce_sum
