"""CPU oracle for the physics-loss hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product path (the CUDA library behind ``include/srm_physics.h``)
never routes through it.

**PARITY PINNED AGAINST THE REFERENCE'S OWN CODE (forward values; TensorFlow's kernels stood in for).**  The
reference (molokwuvictor/3d-physics-based-ai-surrogate-reservoir-model) is pure Python/TensorFlow; TensorFlow is not
installable in the build container, the shipped reference is not import-clean, and it ships no test with an
assertion, no golden vector and no known-answer fixture for this path.  This file is a
*restatement* of the reference arithmetic, op by op, in torch-CPU / numpy (fp32 for parity, fp64 as the
exact-arithmetic twin), each function citing the reference lines it follows.  It is pinned by running the reference's
OWN source in the build container: ``tests/golden/make_reference_*_golden.py`` cut the functions / classes out of
``/root/reference`` by AST (nothing is copied into this repository) and execute them on seeded inputs with
``tests/golden/tf_torch_shim.py`` -- a torch-backed stand-in for the ~60 TensorFlow ops they use -- in place of
TensorFlow.  The committed goldens (``tests/golden/reference_*.npz``) hold the outputs, and ``tests/test_oracle*.py``
require this file to reproduce them:

  =====================================================  ===========================================  ==============
  reference code executed                                 what it pins                                 agreement
  =====================================================  ===========================================  ==============
  physics_error_gas_2D (physics_loss.py:9-224)            dry-gas dom, ibc, mbc (2-D grids)            bit for bit
  physics_error_gas_oil_2D (physics_loss.py:230-714)      gas-condensate dom, ibc, mbc, cmbc           bit for bit
  PolyharmonicSplineInterpolationLayer (polyhm_splines)   order-1 values of the 7 properties           bit for bit *
  PVTLayer.call (PVT_Layer_Subclassed.py:146-216)         clamp, layout, derivative w.r.t. clamped p   1e-5 / 1e-4 **
  RelativePermeability.compute_krog_krgo                  Corey values and end-point rules             <= 2 ulp ***
  WellRatesPressure.compute_rates_and_bhp (+ helpers)     DG, DG + blocking integral, GC rates, BHP    bit for bit
  WellDataProcessor.scatter_y / conn_shutins_idx          scatter positions, shut-in identity          exact
  DataSummary.nonormalize / normalize_diff                linear rows / log permeability row           exact / 2 ulp
  pinn_batch_sse_grad (physics_loss.py:742-870)           SSE per term, weights, counts, reported MSE  1e-5 / exact
  HardLayer (Hard_Layer_Subclassed.py:21-260)             layer output / cotangents                    bit for bit / 1e-6
  BatchGenerator._maybe_flatten (training.py)             sample order of the flattened batch axis     exact
  =====================================================  ===========================================  ==============
  *   with (w, v) as data and the stand-in's matmul accumulating the inner index sequentially (assumption below);
      the layer's own in-call solve (LAPACK through torch instead of numpy, cond ~ 3.6e6) moves values by <= 5e-6.
  **  torch autograd stands in for TF's tape: same formula chain, framework-specific accumulation order.
  *** tf.pow with the integer Corey exponents is pinned as a product here; libm's pow differs by <= 2 ulp.

What the stand-in cannot pin (it implements TF's op SEMANTICS, not Eigen's kernels): the accumulation order inside
tf.matmul / tf.reduce_sum and inside TF's gradient kernels, and ulp-level differences of pow / exp / log.  Those are
the assumptions listed next.  Gradients are validated against fp64 central finite differences of this file's own loss
(tests/test_oracle.py, tests/test_oracle_gc.py); the z faces are an extension (below) with nothing to pin against.

What is pinned here that TensorFlow leaves unspecified (so that "fp32, 1e-5" is well defined):
  * no FMA contraction anywhere: every ``*`` and ``+`` of the reference is a separately rounded op;
  * the 37-term RBF reduction (``tf.matmul`` in ``polyhm_splines.py:142``) is a sequential
    left-to-right fp32 sum over the knots in table order;
  * the reductions inside the spline's first and second derivative (TF's MatMul/Sum gradients) are
    the same sequential sums, combined as ``((S1*(2x)) + (-2*S2)) + v0`` -- see ``spline_eval_np``;
    the same pinned derivative serves TF's inner tape (PVTLayer) and the outer tape;
  * sqrt is IEEE correctly rounded (numpy) -- torch-CPU's vectorised sqrt is not;
  * ``tf.pow(dt2, 2.)`` in ``physics_loss.py:171`` is ``dt2*dt2``;
  * the spline weights ``(w, v)`` are solved once (fp32 LAPACK) and treated as data.

Extension beyond the shipped reference: the shipped stencil is 2-D 5-point on
``(B,H,W)``.  The 7-point 3-D form adds the ``k+-1`` faces with ``kz = kv_kh*kx`` and ``dz`` in
*difference form* ``a5*(p-pD) + a6*(p-pU)``; with ``Nz == 1`` and edge-replicating (SYMMETRIC)
padding that term is exactly ``0.0`` so the 3-D arithmetic reduces bit-for-bit to the shipped 2-D
arithmetic.

Layout conventions: fields are ``(B, D, H, W)`` with ``W`` (x, index ``i``) contiguous, ``H`` = y
(``j``), ``D`` = z (``k``).  ``kx`` is ``(R, D, H, W)`` per realisation, ``sample_real[b]`` maps a
sample to its realisation.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

EPSILON = 1e-10  # polyhm_splines.py:6

# loss-term slots written by the CUDA library and by this oracle (same order)
TERM_NAMES = ("dom", "ibc", "mbc", "tde", "obc", "ic", "td", "cmbc")
N_TERMS = len(TERM_NAMES)


# --------------------------------------------------------------------------------------------
# configuration (constants restated from default_configurations.py; values, not code)
# --------------------------------------------------------------------------------------------
@dataclass
class Well:
    """One connection (default_configurations.py:132-140; welldata_processor.py:39-70)."""
    i: int
    j: int
    k: int
    value: float = 500.0          # 'value' of the control mode (used as a gas-rate target)
    minimum_bhp: float = 4100.0
    wellbore_radius: float = 0.09525
    completion_ratio: float = 0.5
    producer: bool = True
    shutin_days: Tuple[float, float] = (1000.0, 0.0)   # [start, stop]; start>stop == never shut


@dataclass
class OracleConfig:
    D: int = 1
    H: int = 39
    W: int = 39
    length: float = 2900.0        # x extent (default_configurations.py:99)
    width: float = 2900.0         # y extent
    thickness: float = 80.0       # z extent
    phi: float = 0.2              # default_configurations.py:93
    kx_ky: float = 1.0            # horizontal_anisotropy
    kv_kh: float = 1.0            # vertical_anisotropy
    C: float = 0.001127           # default_configurations.py:450
    Dc: float = 5.6145833334
    # SCAL (default_configurations.py:262-266)
    Swmin: float = 0.22
    Sorg: float = 0.2
    Sgc: float = 0.05
    Socr: float = 0.2
    So_max: float = 0.28
    kro_Somax: float = 0.90
    krg_Sorg: float = 0.80
    krg_Swmin: float = 0.90
    nog: float = 3.0
    ng: float = 6.0
    # PVT clamp (PVT_Layer_Subclassed.py:165-167)
    p_min: float = 14.7
    p_max: float = 10000.0
    wells: List[Well] = field(default_factory=list)
    use_blocking_factor: bool = False     # well_rate_bhp_Subclassed.py:36
    n_intervals: int = 8                  # well_rate_bhp_Subclassed.py:39
    use_non_iterative: bool = True        # well_rate_bhp_Subclassed.py:44: False -> _iterative_method (:515-612)
    bhp_max_iters: int = 10               # well_rate_bhp_Subclassed.py:41
    bhp_tol: float = 1e-6                 # well_rate_bhp_Subclassed.py:42
    tde_in_dom: bool = True               # legacy DG folds trn_err into dom (physics_loss.py:175)

    @property
    def dx(self):
        return self.length / self.W

    @property
    def dy(self):
        return self.width / self.H

    @property
    def dz(self):
        return self.thickness / self.D


def default_wells(W=39, H=39, D=1, all_layers=False) -> List[Well]:
    """The five default connections (default_configurations.py:133-139) scaled to the grid."""
    def sc(v, n):
        return min(n - 1, int(round(v / 39.0 * n)))
    base = [(29, 29, 500.0, True), (29, 9, 1000.0, True), (9, 9, 500.0, True),
            (9, 29, 1000.0, True), (19, 19, 0.0, False)]
    out = []
    for (i, j, val, prod) in base:
        ks = range(D) if all_layers else (0,)
        for k in ks:
            out.append(Well(i=sc(i, W), j=sc(j, H), k=k, value=val, producer=prod))
    return out


# --------------------------------------------------------------------------------------------
# fp32 helpers
# --------------------------------------------------------------------------------------------
def f32(x) -> np.float32:
    return np.float32(x)


def rock_compressibility(phi: float) -> np.float32:
    """cf = 97.32e-6/(1+55.8721*phi**1.428586)  (physics_loss.py:64), fp32 scalar."""
    p = np.float32(phi)
    return np.float32(np.float32(97.32e-6) / (np.float32(1.0) + np.float32(55.8721) * np.power(p, np.float32(1.428586))))


def corey_krog_krgo_np(sg, cfg: OracleConfig, dtype=np.float32):
    """RelativePermeability.compute_krog_krgo (relative_permeability.py:49-75), numpy, elementwise."""
    t = dtype
    sg = np.asarray(sg, dtype=t)
    one = t(1.0)
    swmin, sorg, sgc, socr = t(cfg.Swmin), t(cfg.Sorg), t(cfg.Sgc), t(cfg.Socr)
    so = one - sg - swmin                                                     # :58
    with np.errstate(invalid="ignore"):
        krog = t(cfg.kro_Somax) * np.power((so - sorg) / (t(1.0) - swmin - sorg), t(cfg.nog))      # :59
        krgo = t(cfg.krg_Sorg) * np.power((sg - sgc) / (t(1.0) - sgc - swmin - sorg), t(cfg.ng))  # :60
    sorg_eff = max(sorg, socr)                                                # :66
    krog = np.where(so <= (swmin + sorg_eff), t(0.0), krog)                   # :67
    krgo = np.where(sg > (t(1.0) - (swmin + sorg)), t(cfg.krg_Swmin), krgo)   # :68
    krog = np.maximum(np.minimum(krog, t(cfg.kro_Somax)), t(0.0))             # :71
    krgo = np.maximum(np.minimum(krgo, t(cfg.krg_Swmin)), t(0.0))             # :72
    return krog.astype(t), krgo.astype(t)


# --------------------------------------------------------------------------------------------
# polyharmonic spline (polyhm_splines.py)
# --------------------------------------------------------------------------------------------
@dataclass
class SplineTable:
    """Knots + solved weights for P properties; (w, v) are DATA once solved."""
    c: np.ndarray            # (n,)  float32 knots (pressure, psi)
    f: np.ndarray            # (P,n) float32 knot values
    w: np.ndarray            # (P,n) float32 RBF weights
    v: np.ndarray            # (P,2) float32 linear term [slope, intercept]
    order: int
    lam: float
    names: Tuple[str, ...]


def _phi_np(r, order):
    """_phi (polyhm_splines.py:77-88)."""
    rs = np.maximum(r, r.dtype.type(EPSILON))
    if order == 1:
        return np.sqrt(rs)
    if order == 2:
        return r.dtype.type(0.5) * rs * np.log(rs)
    raise NotImplementedError("spline order %r" % order)


def spline_solve(c, f, order=1, lam=0.001, dtype=np.float32):
    """_solve_interpolation (polyhm_splines.py:103-135) for one property.

    The reference re-solves this (n+2)x(n+2) system inside every call (polyhm_splines.py:180);
    it depends on constants only, so it is solved once.  tf.linalg.solve -> LAPACK ?gesv here.
    """
    c = np.asarray(c, dtype=dtype).reshape(-1)
    f = np.asarray(f, dtype=dtype).reshape(-1)
    n = c.size
    cn = c * c                                              # :93-94 reduce_sum(square), d=1
    xy = (c[:, None] * c[None, :]).astype(dtype)            # :95 matmul with K=1
    r = (cn[:, None] - dtype(2) * xy) + cn[None, :]         # :96
    a = _phi_np(r.astype(dtype), order).astype(dtype)       # :112-115
    if lam > 0:
        a = a + dtype(lam) * np.eye(n, dtype=dtype)         # :116-118
    b = np.stack([c, np.ones(n, dtype=dtype)], axis=1)      # :120-121
    lhs = np.zeros((n + 2, n + 2), dtype=dtype)             # :123-127
    lhs[:n, :n] = a
    lhs[:n, n:] = b
    lhs[n:, :n] = b.T
    rhs = np.zeros((n + 2,), dtype=dtype)                   # :129-130
    rhs[:n] = f
    sol = np.linalg.solve(lhs, rhs).astype(dtype)           # :132
    return sol[:n].copy(), sol[n:].copy()


def load_pvt_table(path) -> Dict[str, np.ndarray]:
    z = np.load(path)
    return {str(k): z["table"][i].astype(np.float32) for i, k in enumerate(z["columns"])}


DG_PROPS = ("InvBg", "Invug")                                          # PVT_Layer_Subclassed.py:69-70
GC_PROPS = ("InvBg", "InvBo", "Invug", "Invuo", "Rs", "Rv", "Vro")     # PVT_Layer_Subclassed.py:71-72


def build_spline_table(columns: Dict[str, np.ndarray], props: Sequence[str], order=1, lam=0.001) -> SplineTable:
    """PVTLayer.build, spline branch (PVT_Layer_Subclassed.py:118-141): one spline per property on 'Pre'."""
    c = columns["Pre"].astype(np.float32)
    fs, ws, vs = [], [], []
    for p in props:
        w, v = spline_solve(c, columns[p], order=order, lam=lam)
        fs.append(columns[p].astype(np.float32))
        ws.append(w)
        vs.append(v)
    return SplineTable(c=c, f=np.stack(fs), w=np.stack(ws), v=np.stack(vs), order=order, lam=lam, names=tuple(props))


@dataclass
class PolyTable:
    """PVTLayer, fitting_method='polynomial' (PVT_Layer_Subclassed.py:77-87): coefficients [a_0, a_1, ...] per property"""
    coef: np.ndarray         # (P, n) float32
    names: Tuple[str, ...]
    order: int = 0


def poly_eval_np(x, tab: PolyTable, prop: int, dtype=np.float32, need=2):
    """PVTLayer.evaluate_polynomial (PVT_Layer_Subclassed.py:218-266), every line one rounded op:
        value = sum_i a_i * pow(x, i)               accumulated from zero, i ascending     :239-245
        d1    = sum_{i>=1} (i*a_i) * pow(x, i-1)    the layer's explicit derivative        :248-255
        d2    = TF's gradient of d1: sum_{i>=2} (i*a_i) * ((i-1) * pow(x, i-2))
    tf.pow with the integer exponents is pinned as the left-to-right product (pow(x,0)=1, pow(x,1)=x)."""
    t = dtype
    x = np.asarray(x, dtype=t)
    a = tab.coef[prop].astype(t)
    acc = np.zeros_like(x)
    a1 = np.zeros_like(x)
    a2 = np.zeros_like(x)
    pw, pwm1, pwm2 = np.ones_like(x), np.zeros_like(x), np.zeros_like(x)
    for i in range(a.size):
        acc = acc + a[i] * pw
        if i >= 1 and need >= 1:
            ic = t(i) * a[i]
            a1 = a1 + ic * pwm1
            if i >= 2 and need >= 2:
                a2 = a2 + ic * (t(i - 1) * pwm2)
        pwm2, pwm1 = pwm1, pw
        pw = x.copy() if i == 0 else pw * x
    return acc, (a1 if need >= 1 else None), (a2 if need >= 2 else None)


def prop_eval_np(x, tab, prop, dtype=np.float32, need=2):
    """value / d1 / d2 of one property: spline or polynomial fit, whichever table is given"""
    if isinstance(tab, PolyTable):
        return poly_eval_np(x, tab, prop, dtype, need)
    return spline_eval_np(x, tab, prop, dtype, need)


def _tf_maximum(x, y):
    """tf.maximum forward+gradient semantics: gradient to x where x >= y (ties -> first arg)."""
    return torch.where(x >= y, x, y)


def _tf_minimum(x, y):
    """tf.minimum: gradient to x where x <= y."""
    return torch.where(x <= y, x, y)


def _tf_clip(t, lo, hi):
    """tf.clip_by_value: value max(min(t,hi),lo) (the op's cwiseMin/cwiseMax); gradient per
    _ClipByValueGrad: to t where lo <= t <= hi, to lo where t < lo, to hi where t > hi."""
    val = torch.maximum(torch.minimum(t, hi), lo)
    below = (t < lo).to(t.dtype)
    above = (t > hi).to(t.dtype)
    route = (1.0 - below) * (1.0 - above) * t + below * lo + above * hi
    return route + (val - route).detach()


def _dnn(x, y):
    """tf.math.divide_no_nan: 0 where y == 0 (and zero gradient there)."""
    safe = torch.where(y == 0, torch.ones_like(y), y)
    return torch.where(y == 0, torch.zeros_like(x * y), x / safe)


def pvt_clamp(p, cfg: OracleConfig):
    """inputs_safe (PVT_Layer_Subclassed.py:165-167)."""
    lo = torch.full_like(p, cfg.p_min)
    hi = torch.full_like(p, cfg.p_max)
    return _tf_minimum(_tf_maximum(p, lo), hi)


def spline_eval_np(x, tab: SplineTable, prop: int, dtype=np.float32, need=2):
    """Value, first and second derivative of one property at clamped pressures ``x`` (numpy array).

    numpy, not torch: torch-CPU's vectorised ``sqrt`` (Sleef) is not correctly rounded (0.7 % of fp32
    inputs differ from IEEE), whereas TF/Eigen, numpy and CUDA ``sqrt.rn`` are.  Every line is one
    rounded fp32 op.

    value   _apply_interpolation (polyhm_splines.py:138-146):
        r_i   = (x*x - 2*(x*c_i)) + c_i*c_i                      :93-96
        phi_i = sqrt(max(r_i, EPS))                              :77-80   (order 1)
        val   = (sum_i phi_i*w_i, sequential) + (x*v0 + v1)      :142-146
    first derivative = what TF's tape returns for d val / d x (PVT_Layer_Subclassed.py:196-201),
    reductions pinned to sequential knot order:
        g_i = [r_i >= EPS] * (0.5*w_i)/phi_i         MatMulGrad -> w_i ; SqrtGrad (0.5*dy)/y ; MaximumGrad
        S1 = sum g_i ;  S2 = sum g_i*c_i             (grad reaching x_norm ; MatMulGrad of x c^T)
        d1 = ((S1*(x*2)) + (-2*S2)) + v0             SquareGrad = dy*(x*2) ; add_n
    second derivative = TF's gradient of the ops above (the nested-tape term of SURVEY H2):
        e_i   = (x*2) + (-2*c_i)                      cotangent reaching g_i
        dphi  = e_i * (-(g_i/phi_i))                  RealDivGrad  dy * (-x/y/y)
        drs   = (0.5*dphi)/phi_i ; dr = [r_i>=EPS]*drs           SqrtGrad ; MaximumGrad
        S1p = sum dr ; S2p = sum dr*c_i
        d2 = ((S1*2) + (S1p*(x*2))) + (-2*S2p)
    """
    t = dtype
    x = np.asarray(x, dtype=t)
    c = tab.c.astype(t)
    w = tab.w[prop].astype(t)
    v = tab.v[prop].astype(t)
    c2 = (c * c).astype(t)
    x2 = x * x
    tx = x * t(2)
    eps = t(EPSILON)
    zero = t(0)
    acc = np.zeros_like(x)
    s1 = np.zeros_like(x)
    s2 = np.zeros_like(x)
    s1p = np.zeros_like(x)
    s2p = np.zeros_like(x)
    with np.errstate(all="ignore"):
        for i in range(c.size):
            r = (x2 - t(2) * (x * c[i])) + c2[i]
            live = r >= eps
            rs = np.where(live, r, eps).astype(t)
            if tab.order == 1:
                ph = np.sqrt(rs)
            elif tab.order == 2:
                ph = t(0.5) * rs * np.log(rs)
            else:
                raise NotImplementedError
            acc = acc + ph * w[i]
            if need >= 1:
                if tab.order != 1:
                    raise NotImplementedError("pinned derivatives only for order 1")
                g = np.where(live, (t(0.5) * w[i]) / ph, zero).astype(t)
                s1 = s1 + g
                s2 = s2 + g * c[i]
                if need >= 2:
                    e = tx + (t(-2) * c[i])
                    dphi = e * (-(g / ph))
                    dr = np.where(live, (t(0.5) * dphi) / ph, zero).astype(t)
                    s1p = s1p + dr
                    s2p = s2p + dr * c[i]
    val = acc + (x * v[0] + v[1])
    d1 = ((s1 * tx) + (t(-2) * s2)) + v[0] if need >= 1 else None
    d2 = ((s1 * t(2)) + (s1p * tx)) + (t(-2) * s2p) if need >= 2 else None
    return val, d1, d2


class _SplineD1(torch.autograd.Function):
    """d val/d x as a differentiable node: backward multiplies by the pinned second derivative."""

    @staticmethod
    def forward(ctx, ph, tab, prop):
        npdt = np.float64 if ph.dtype == torch.float64 else np.float32
        _, d1, d2 = prop_eval_np(ph.detach().numpy(), tab, prop, npdt, need=2)
        ctx.save_for_backward(torch.from_numpy(np.ascontiguousarray(d2)))
        return torch.from_numpy(np.ascontiguousarray(d1))

    @staticmethod
    def backward(ctx, g):
        (d2,) = ctx.saved_tensors
        return g * d2, None, None


class _SplineVal(torch.autograd.Function):
    """spline value; backward multiplies by the pinned first derivative."""

    @staticmethod
    def forward(ctx, ph, tab, prop):
        npdt = np.float64 if ph.dtype == torch.float64 else np.float32
        val, d1, _ = prop_eval_np(ph.detach().numpy(), tab, prop, npdt, need=1)
        ctx.save_for_backward(torch.from_numpy(np.ascontiguousarray(d1)))
        return torch.from_numpy(np.ascontiguousarray(val))

    @staticmethod
    def backward(ctx, g):
        (d1,) = ctx.saved_tensors
        return g * d1, None, None


def spline_apply(ph, tab: SplineTable, prop: int, want_deriv: bool = False):
    """PolyharmonicSplineInterpolationLayer.call at clamped pressure ``ph`` (torch, differentiable);
    see ``spline_eval_np`` for the arithmetic."""
    val = _SplineVal.apply(ph, tab, prop)
    if not want_deriv:
        return val
    return val, _SplineD1.apply(ph, tab, prop)


def pvt_eval(p, tab: SplineTable, cfg: OracleConfig, props: Optional[Sequence[int]] = None,
             need_deriv: Sequence[int] = ()):
    """PVTLayer.call, spline branch (PVT_Layer_Subclassed.py:146-216).

    Returns (values{prop: tensor}, derivs{prop: tensor}).  The derivative is the inner-tape gradient
    of the value w.r.t. the *clamped* input (PVT_Layer_Subclassed.py:196-201); it is differentiable,
    which is what gives TF's outer tape the second-order term.
    """
    ph = pvt_clamp(p, cfg)
    if props is None:
        props = range(len(tab.names))
    values, derivs = {}, {}
    for q in props:
        if q in need_deriv:
            values[q], derivs[q] = spline_apply(ph, tab, q, want_deriv=True)
        else:
            values[q] = spline_apply(ph, tab, q)
    return values, derivs


def spline_closed_form_fp64(p, tab: SplineTable, cfg: OracleConfig, prop: int):
    """Exact-arithmetic value/slope of the order-1 interpolant: piecewise linear between knots.

    Used only to characterise the fp32 noise of the reference form."""
    assert tab.order == 1
    x = np.clip(np.asarray(p, dtype=np.float64), cfg.p_min, cfg.p_max)
    c = tab.c.astype(np.float64)
    w = tab.w[prop].astype(np.float64)
    v = tab.v[prop].astype(np.float64)
    d = x[..., None] - c
    val = (np.abs(d) * w).sum(-1) + x * v[0] + v[1]
    slope = (np.sign(d) * w).sum(-1) + v[0]
    return val, slope


# --------------------------------------------------------------------------------------------
# wells  (welldata_processor.py, well_rate_bhp_Subclassed.py)
# --------------------------------------------------------------------------------------------
def well_connection_index(wells: Sequence[Well]) -> np.ndarray:
    """(i,j,k) -> rows [k,j,i]  (welldata_processor.py:26-40, mode_order=('k','j','i')).  int32, bit-exact."""
    coords = np.array([[w.i, w.j, w.k] for w in wells], dtype=np.int32).reshape(-1, 3)
    perm = [2, 1, 0]
    return coords[:, perm].astype(np.int32)


def well_flat_index(wells: Sequence[Well], D: int, H: int, W: int) -> np.ndarray:
    kji = well_connection_index(wells)
    return ((kji[:, 0].astype(np.int64) * H + kji[:, 1]) * W + kji[:, 2]).astype(np.int32)


def well_control_value(w: Well) -> float:
    """sign rule: producer +, injector - (welldata_processor.py:89-97; ORAT column)."""
    return float(w.value) * (1.0 if w.producer else -1.0)


def shutin_open_mask(t_days: np.ndarray, wells: Sequence[Well]) -> np.ndarray:
    """conn_shutins_idx (welldata_processor.py:349-354): 1 where time is NOT within inclusive [start,stop].

    Returns int32 (B, n_wells).  Comparison is done in fp32 like the reference."""
    t = np.asarray(t_days, dtype=np.float32).reshape(-1, 1)
    start = np.array([w.shutin_days[0] for w in wells], dtype=np.float32).reshape(1, -1)
    stop = np.array([w.shutin_days[1] for w in wells], dtype=np.float32).reshape(1, -1)
    inside = (t >= start) & (t <= stop)
    return (~inside).astype(np.int32)


def peaceman_static(kx_cell, cfg: OracleConfig, w_rw, w_hc, dtype):
    """ro (well_rate_bhp_Subclassed.py:783-786) and the static part of Ck (:788, np.pi).

    kx_cell: tensor (..., n_wells) physical kx at the connection cells."""
    dt = dtype
    kx = kx_cell
    ky = torch.tensor(cfg.kx_ky, dtype=dt) * kx
    dx = torch.tensor(cfg.dx, dtype=dt)
    dy = torch.tensor(cfg.dy, dtype=dt)
    dz = torch.tensor(cfg.dz, dtype=dt)
    ro = 0.28 * torch.pow(torch.pow(ky / kx, 0.5) * torch.pow(dx, 2) + torch.pow(kx / ky, 0.5) * torch.pow(dy, 2), 0.5) \
        / (torch.pow(ky / kx, 0.25) + torch.pow(kx / ky, 0.25))
    two_pi = torch.tensor(2 * np.pi, dtype=dt)
    ck = (two_pi * w_hc * kx * dz * torch.tensor(cfg.C, dtype=dt)) / torch.log(ro / w_rw)
    return ck


def _mobility_gas_dg(p, tab, cfg, krg):
    vals, _ = pvt_eval(p, tab, cfg, props=(0, 1))
    return krg * vals[0] * vals[1]                         # well_rate_bhp_Subclassed.py:799


def blocking_integral_dg(p, pwf, tab, cfg, krg_p, krg_smax, n_intervals):
    """compute_blocking_integral_and_factor, DG branch (well_rate_bhp_Subclassed.py:840-960).

    tf.linspace(p, pwf, n+1): first/last points are start/stop exactly, interior = start + delta*i."""
    mg_n1 = _mobility_gas_dg(p, tab, cfg, krg_p)
    n = n_intervals
    delta = (pwf - p) / float(n)
    grid = [p] + [p + delta * float(i) for i in range(1, n)] + [pwf]
    sum_g = torch.zeros_like(p)
    mg_prev = mg_n1
    for i in range(n):
        p0 = grid[i]
        p1 = grid[i + 1]
        vals, _ = pvt_eval(p1, tab, cfg, props=(0, 1))
        mg1 = krg_smax * vals[0] * vals[1]                 # :917-918 with Sg1 = Sg_max (:912)
        dp = p0 - p1
        sum_g = sum_g + 0.5 * (mg_prev + mg1) * dp         # :920
        mg_prev = mg1
    return sum_g, mg_n1


def bhp_newton(gas_rate, p, pmin, q_t, cfg: OracleConfig):
    """WellRatesPressure._iterative_method (well_rate_bhp_Subclassed.py:515-612): Newton-Raphson on the bottom-hole
    pressure with a one-sided difference quotient (eps = 14.7 psi) for d qg / d pwf, clipped into [min_bhp, p] after
    every step.  The stopping test is the reference's: the WHOLE batch iterates while any connection misses its target
    by more than tol, at most max_iters times (tf.while_loop's cond, :548-562); the loop is differentiated through, as
    tf.while_loop's gradient does.  gas_rate(pwf) is _compute_phase_rates' qg."""
    pwf = pmin + 0.5 * (p - pmin)                                                # :537
    eps = 14.7                                                                   # :540
    it = 0
    while it < cfg.bhp_max_iters:
        qg = gas_rate(pwf)                                                       # :566-569 (the cond evaluates the same rate, :549-553)
        err = (qg - q_t).abs().detach()                                          # cond: evaluated, not differentiated
        if not bool((err > cfg.bhp_tol).any()):
            break
        qg_plus = gas_rate(pwf + eps)                                            # :571-574
        dq = (qg_plus - qg) / eps                                                # :576
        pwf_new = pwf - (qg - q_t) / (dq + 1e-12)                                # :584
        pwf = _tf_clip(pwf_new, pmin, p)                                         # :586
        it += 1
    return pwf


def wells_dg(p_cell, kx_cell, t_days, tab: SplineTable, cfg: OracleConfig, dtype):
    """WellRatesPressure.compute_rates_and_bhp, DG, non-iterative control
    (well_rate_bhp_Subclassed.py:727-837, 614-724, 963-1007), evaluated at the connection cells only.

    p_cell  : (B, nw) pressure at the connection cells (differentiable)
    kx_cell : (B, nw) physical permeability at those cells
    returns q (B,nw), pwf (B,nw)
    """
    dt = dtype
    wells = cfg.wells
    nw = len(wells)
    as_t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=dt).reshape(1, nw)
    rw = as_t([w.wellbore_radius for w in wells])
    hc = as_t([w.completion_ratio for w in wells])
    q_t = as_t([well_control_value(w) for w in wells])
    pmin = as_t([w.minimum_bhp for w in wells])
    well_id = torch.ones((1, nw), dtype=dt)
    shut = torch.as_tensor(shutin_open_mask(t_days, wells), dtype=dt)           # (B,nw)
    ck = shut * peaceman_static(kx_cell, cfg, rw, hc, dt)                        # :788
    _, krg = corey_krog_krgo_np(1.0 - cfg.Swmin, cfg, np.float64 if dt == torch.float64 else np.float32)
    krg = torch.tensor(float(krg), dtype=dt)                                     # Sg_n1 None -> 1-Swmin (:758)
    p = p_cell
    mg = _mobility_gas_dg(p, tab, cfg, krg)                                      # :799
    tiny = 1e-12

    def integral(pwf):
        if cfg.use_blocking_factor:
            ig, _ = blocking_integral_dg(p, pwf, tab, cfg, krg, krg, cfg.n_intervals)
            return ig
        return torch.ones_like(p)                                                # :955-959

    def phase_rates(pwf_):                                                       # _compute_phase_rates (:963-1007)
        ig = integral(pwf_)
        dp = p - pwf_ + tiny                                                     # :987
        if cfg.use_blocking_factor:
            blk = _dnn(ig, mg * dp)                                              # :991
        else:
            blk = ig
        qg_max2 = well_id * ck * blk * mg * dp                                   # :997
        return _tf_maximum(_tf_minimum(q_t.expand_as(p), qg_max2), torch.zeros_like(p))  # :1001

    if cfg.use_non_iterative:
        # ---- _non_iterative_method (:614-724)
        ig_max = integral(pmin.expand_as(p))
        dp_max = p - pmin + tiny                                                 # :650
        if cfg.use_blocking_factor:
            blk_max = _dnn(ig_max, mg * dp_max)                                  # :654
        else:
            blk_max = ig_max                                                     # :657
        qg_max = well_id * ck * blk_max * mg * dp_max                            # :662
        qg_opt = _tf_maximum(_tf_minimum(q_t.expand_as(p), qg_max), torch.zeros_like(p))   # :666
        lam = _tf_clip(_dnn(qg_opt, well_id * ck * blk_max * mg), torch.zeros_like(p), blk_max)  # :699
        dp_opt = lam * dp_max                                                    # :721
        pwf = p - dp_opt
        pwf = well_id * _tf_clip(pwf, pmin.expand_as(p), p)                      # :723
    else:
        pwf = bhp_newton(phase_rates, p, pmin.expand_as(p), q_t.expand_as(p), cfg)
    qg = phase_rates(pwf)
    return qg, pwf


# --------------------------------------------------------------------------------------------
# DG residual  (physics_loss.py:9-208)
# --------------------------------------------------------------------------------------------
def _nbr(x, dim, off):
    """Edge-replicated neighbour (tf.pad SYMMETRIC width 1 + slicing, physics_loss.py:18-21,33-38)."""
    n = x.shape[dim]
    idx = torch.clamp(torch.arange(n) + off, 0, n - 1)
    return x.index_select(dim, idx)


def _harm(kc, kn):
    """(2.*k1*k2)/(k1+k2)  (physics_loss.py:59-60)."""
    return (2.0 * kc * kn) / (kc + kn)


def dg_static(kx, cfg: OracleConfig):
    """Per-realisation face permeabilities (physics_loss.py:56-60; z faces by symmetry)."""
    dt = kx.dtype
    ky = torch.tensor(cfg.kx_ky, dtype=dt) * kx
    kz = torch.tensor(cfg.kv_kh, dtype=dt) * kx
    return dict(
        kW=_harm(kx, _nbr(kx, -1, -1)), kE=_harm(_nbr(kx, -1, +1), kx),
        kS=_harm(ky, _nbr(ky, -2, -1)), kN=_harm(_nbr(ky, -2, +1), ky),
        kD=_harm(kz, _nbr(kz, -3, -1)), kU=_harm(_nbr(kz, -3, +1), kz),
    )


def dg_residual(cfg: OracleConfig, tab: SplineTable, kx, p0, p1, dt1, dt2, t_days, sample_real,
                dtype=torch.float32):
    """physics_error_gas (physics_loss.py:79-208) given the network outputs.

    kx (R,D,H,W) physical; p0,p1 (B,D,H,W); dt1,dt2 (B,); t_days (B,) time at level n+1 (for
    shut-ins); sample_real (B,) int.  Returns a dict of tensors (differentiable w.r.t. p0,p1,dt1,dt2).
    """
    dt = dtype
    T = lambda v: torch.tensor(v, dtype=dt)
    B = p0.shape[0]
    sr = torch.as_tensor(np.asarray(sample_real), dtype=torch.long)
    kxb = kx.to(dt)
    st = {k: v.index_select(0, sr) for k, v in dg_static(kxb, cfg).items()}
    dx, dy, dz = T(cfg.dx), T(cfg.dy), T(cfg.dz)
    dv = dx * dy * dz                                                        # :36
    C, Dc = T(cfg.C), T(cfg.Dc)
    phi = T(cfg.phi)
    cf = torch.tensor(float(rock_compressibility(cfg.phi)), dtype=dt) if dt == torch.float32 else \
        T(97.32e-6 / (1 + 55.8721 * cfg.phi ** 1.428586))                    # :64
    Sgi = T(1.0 - cfg.Swmin) if dt == torch.float64 else torch.tensor(float(np.float32(1.0 - cfg.Swmin)), dtype=dt)  # :65
    _, krg_np = corey_krog_krgo_np(1.0 - cfg.Swmin, cfg, np.float64 if dt == torch.float64 else np.float32)
    krg = T(float(krg_np))                                                   # :129
    d1 = dt1.reshape(B, 1, 1, 1)
    d2 = dt2.reshape(B, 1, 1, 1)

    # PVT at n0 (value + dp-derivative of invBg) and at n1 (values)            :88-95, 111-117
    v0, dv0 = pvt_eval(p0, tab, cfg, props=(0,), need_deriv=(0,))
    A0, A0p = v0[0], dv0[0]
    v1, _ = pvt_eval(p1, tab, cfg, props=(0, 1))
    A1, M1 = v1[0], v1[1]

    p2 = (p1 - p0) * (1.0 + _dnn(d2, d1)) + p0                               # :126

    G1 = A1 * M1                                                             # :137
    pW, pE = _nbr(p1, -1, -1), _nbr(p1, -1, +1)                              # :132-133
    pS, pN = _nbr(p1, -2, -1), _nbr(p1, -2, +1)
    pD, pU = _nbr(p1, -3, -1), _nbr(p1, -3, +1)
    GW = (G1 + _nbr(G1, -1, -1)) / 2.0                                       # :147-148
    GE = (_nbr(G1, -1, +1) + G1) / 2.0
    GS = (G1 + _nbr(G1, -2, -1)) / 2.0
    GN = (_nbr(G1, -2, +1) + G1) / 2.0
    GD = (G1 + _nbr(G1, -3, -1)) / 2.0
    GU = (_nbr(G1, -3, +1) + G1) / 2.0
    cr = (phi * cf) * A0                                                     # :149
    cp = Sgi * ((phi * A0p) + cr)                                            # :150
    idx_, idy_, idz_ = 1.0 / dx, 1.0 / dy, 1.0 / dz
    a1 = C * st["kW"] * krg * GW * idx_ * idx_                               # :152
    a2 = C * st["kS"] * krg * GS * idy_ * idy_                               # :153
    a3 = C * st["kE"] * krg * GE * idx_ * idx_                               # :154
    a4 = C * st["kN"] * krg * GN * idy_ * idy_                               # :155
    a5 = C * st["kD"] * krg * GD * idz_ * idz_                               # 3-D extension
    a6 = C * st["kU"] * krg * GU * idz_ * idz_
    a5t = (1.0 / Dc) * (cp / d1)                                             # :156

    # wells: sparse evaluation at the connection cells, scattered (scatter_nd sums duplicates)
    nw = len(cfg.wells)
    q = torch.zeros_like(p1)
    pwf = torch.zeros_like(p1)
    mask = torch.zeros((cfg.D, cfg.H, cfg.W), dtype=dt)
    qw = pwfw = None
    if nw:
        flat = torch.as_tensor(well_flat_index(cfg.wells, cfg.D, cfg.H, cfg.W).astype(np.int64))
        p_cell = p1.reshape(B, -1).index_select(1, flat)
        kx_cell = kxb.index_select(0, sr).reshape(B, -1).index_select(1, flat)
        qw, pwfw = wells_dg(p_cell, kx_cell, np.asarray(t_days), tab, cfg, dt)
        q = torch.zeros((B, cfg.D * cfg.H * cfg.W), dtype=dt).index_add(1, flat, qw).reshape(p1.shape)
        pwf = torch.zeros((B, cfg.D * cfg.H * cfg.W), dtype=dt).index_add(1, flat, pwfw).reshape(p1.shape)
        mask = torch.zeros(cfg.D * cfg.H * cfg.W, dtype=dt).index_add(0, flat, torch.ones(nw, dtype=dt)).reshape(cfg.D, cfg.H, cfg.W)

    den = (d1 * d2) + d2 * d2                                                # :171  (dt2**2. pinned as dt2*dt2)
    num = ((d2 * p0) + (d1 * p2)) - ((d1 + d2) * p1)
    tde = (dv / Dc) * cp * ((2e-7 / d1) + (num / den))                       # :171
    s = (-a1 * pW) + (-a2 * pS) + ((a1 + a2 + a3 + a4) * p1) + (-a3 * pE) + (-a4 * pN)   # :174
    s = s + ((a5 * (p1 - pD)) + (a6 * (p1 - pU)))                            # 3-D extension, == 0 for Nz=1
    s = s + (q / dv)
    divq = dv * s                                                            # :174
    acc = dv * a5t * (p1 - p0)                                               # :175
    if cfg.tde_in_dom:
        dom = divq + (acc + tde)                                             # :175-176
    else:
        dom = divq + acc
    ibc = mask * divq                                                        # :189
    mb_cells = dv * Sgi * phi * (A1 - A0) * (1.0 / (Dc * d1))                # :193
    mbc = (-q.sum(dim=(1, 2, 3))) - mb_cells.sum(dim=(1, 2, 3))              # :193
    return dict(dom=dom, divq=divq, acc=acc, tde=tde, ibc=ibc, mbc=mbc, q=q, pwf=pwf, qw=qw, pwfw=pwfw,
                A0=A0, A0p=A0p, A1=A1, M1=M1, mask=mask)


def dg_loss_terms(res) -> torch.Tensor:
    """SSE per term (physics_loss.py:787-807), slots as TERM_NAMES."""
    z = torch.zeros((), dtype=res["dom"].dtype)
    sse = [
        (res["dom"] ** 2).sum(), (res["ibc"] ** 2).sum(), (res["mbc"] ** 2).sum(), (res["tde"] ** 2).sum(),
        z, z, z, z,
    ]
    return torch.stack(sse)


def dg_counts(cfg: OracleConfig, B: int) -> np.ndarray:
    """error counts (physics_loss.py:825-832): dom, ibc and the truncation term count every cell; mbc is counted with
    the ic FIELD's shape (physics_loss.py:830: reduce_sum(ones_like(ic_pinn_se))), i.e. every cell too."""
    n = B * cfg.D * cfg.H * cfg.W
    return np.array([n, n, n, n, 0, 0, 0, 0], dtype=np.float64)


def dg_forward_backward(cfg, tab, kx, p0, p1, dt1, dt2, t_days, sample_real, weights, dtype=torch.float32):
    """One evaluation of loss terms and the gradient of sum_k weights[k]*SSE_k w.r.t. p0,p1,dt1,dt2
    -- what TF's tape.gradient delivers to the nets' outputs (physics_loss.py:849-859)."""
    tt = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype).clone().requires_grad_(True)
    p0t, p1t, d1t, d2t = tt(p0), tt(p1), tt(dt1), tt(dt2)
    kxt = torch.as_tensor(np.asarray(kx), dtype=dtype)
    res = dg_residual(cfg, tab, kxt, p0t, p1t, d1t, d2t, t_days, sample_real, dtype=dtype)
    terms = dg_loss_terms(res)
    wt = torch.as_tensor(np.asarray(weights, dtype=np.float64), dtype=dtype)
    loss = (terms * wt).sum()
    gp0, gp1, gd1, gd2 = torch.autograd.grad(loss, [p0t, p1t, d1t, d2t], allow_unused=True)
    z = lambda g, ref: torch.zeros_like(ref) if g is None else g
    out = {k: (v.detach().numpy() if isinstance(v, torch.Tensor) else v) for k, v in res.items() if v is not None}
    out.update(terms=terms.detach().numpy(), gp0=z(gp0, p0t).numpy(), gp1=z(gp1, p1t).numpy(),
               gdt1=z(gd1, d1t).numpy(), gdt2=z(gd2, d2t).numpy())
    return out


# --------------------------------------------------------------------------------------------
# GC (gas-condensate, two-phase gas-oil) residual  (physics_loss.py:230-712)
# --------------------------------------------------------------------------------------------
# Additional pins for this path (TF leaves them open):
#   * tf.math.pow with an INTEGER-valued exponent n (Corey exponents nog = 3, ng = 6,
#     relative_permeability.py:59-60) is the left-to-right product ((x*x)*x)...  -- libm's pow and CUDA's
#     powf are not bit-identical, repeated multiplication is; non-integer exponents use pow and are
#     tolerance-checked only;
#   * python-float factors that multiply by exactly 1 (tdew_idx = 1, physics_loss.py:380) are dropped;
#   * 3-D extension as in the DG path: the k+-1 faces enter every component's divergence in
#     difference form a5*(p-pD) + a6*(p-pU) after the in-plane terms; the relative permeability on the
#     U face is selected like the E/N faces (pot = p_nbr - p_c), on the D face like the W/S faces
#     (pot = p_c - p_nbr)  (physics_loss.py:538-551 as written, including its W/S-face convention).
def _pow_pinned(x, n: float):
    if float(n).is_integer() and 1 <= n <= 16:
        y = x
        for _ in range(int(n) - 1):
            y = y * x
        return y
    return torch.pow(x, n)


def corey_krog_krgo_t(sg, cfg: OracleConfig, dtype=torch.float32):
    """RelativePermeability.compute_krog_krgo (relative_permeability.py:49-75), torch, TF gradient routing."""
    npdt = np.float64 if dtype == torch.float64 else np.float32
    t = npdt
    swmin, sorg, sgc, socr = t(cfg.Swmin), t(cfg.Sorg), t(cfg.Sgc), t(cfg.Socr)
    c = lambda v: torch.tensor(float(v), dtype=dtype)
    den_o = (t(1.0) - swmin) - sorg                                   # :59
    den_g = ((t(1.0) - sgc) - swmin) - sorg                           # :60
    so = (1.0 - sg) - c(swmin)                                        # :58
    krog = c(t(cfg.kro_Somax)) * _pow_pinned((so - c(sorg)) / c(den_o), cfg.nog)
    krgo = c(t(cfg.krg_Sorg)) * _pow_pinned((sg - c(sgc)) / c(den_g), cfg.ng)
    sorg_eff = max(sorg, socr)                                        # :66
    krog = torch.where(so <= c(swmin + sorg_eff), torch.zeros_like(krog), krog)             # :67
    krgo = torch.where(sg > c(t(1.0) - (swmin + sorg)), torch.ones_like(krgo) * c(t(cfg.krg_Swmin)), krgo)   # :68
    krog = _tf_maximum(_tf_minimum(krog, torch.ones_like(krog) * c(t(cfg.kro_Somax))), torch.zeros_like(krog))   # :71
    krgo = _tf_maximum(_tf_minimum(krgo, torch.ones_like(krgo) * c(t(cfg.krg_Swmin))), torch.zeros_like(krgo))   # :72
    return krog, krgo


def solve_newton_sg(cost, ref, max_iters, max_value):
    """_solve_newton (well_rate_bhp_Subclassed.py:236-269): start 0.1, Sg <- clip(Sg - f/(df + 1e-12), 0, max_value),
    df by the inner tape; every iteration stays in the graph (tf.while_loop is differentiated through)."""
    sg = torch.full_like(ref, 0.1)
    if not sg.requires_grad:
        sg.requires_grad_(True)
    for _ in range(int(max_iters)):
        f = cost(sg)
        (df,) = torch.autograd.grad(f, sg, grad_outputs=torch.ones_like(f), create_graph=True)
        sg = _tf_clip(sg - f / (df + 1e-12), torch.zeros_like(sg), torch.full_like(sg, max_value))
    return sg


def solve_bracket_sg(cost, ref, max_iters, tol, max_value):
    """_solve_chandrupatla (well_rate_bhp_Subclassed.py:272-324) as written: a regula-falsi bracket update on
    [0, max_value] (hi nudged to lo + 1e-3 when the end values do not bracket), at most max_iters steps while ANY element
    is wider than tol, result 0.5 (lo + hi)."""
    lo = torch.zeros_like(ref)
    hi = torch.ones_like(ref) * max_value
    f_lo, f_hi = cost(lo), cost(hi)
    bad = (f_lo * f_hi) > 0.0
    hi = torch.where(bad, lo + 1e-3, hi)
    f_hi = torch.where(bad, cost(hi), f_hi)
    it = 0
    while bool(((hi - lo) > tol).any()) and it < int(max_iters):
        d = (f_hi - f_lo) / (hi - lo + 1e-12)
        guess = hi - f_hi / d
        f_guess = cost(guess)
        rep = (f_lo * f_guess) < 0.0
        lo, f_lo, hi, f_hi = (torch.where(rep, lo, guess), torch.where(rep, f_lo, f_guess),
                              torch.where(rep, guess, hi), torch.where(rep, f_guess, f_hi))
        it += 1
    return 0.5 * (lo + hi)


def blocking_integral_gc(p, sg, pwf, tab, cfg, krog_n1, mg_n1, mo_n1, dtype, solver="newton", n_root_iter=20):
    """compute_blocking_integral_and_factor, GC branch (well_rate_bhp_Subclassed.py:857-950): trapezoid over
    tf.linspace(p, pwf, n+1); at every node the gas saturation solves mo(Sg) mg_n1 - mo_n1 mg(Sg) = 0 (the producing
    gas-oil ratio of the block is kept along the path), Sg_max where krog of the block is below 1e-3 (:912).
    Returns (Ig, Io)."""
    n = cfg.n_intervals
    delta = (pwf - p) / float(n)
    grid = [p] + [p + delta * float(i) for i in range(1, n)] + [pwf]
    sg_max = 1.0 - cfg.Swmin
    sum_g, sum_o = torch.zeros_like(p), torch.zeros_like(p)
    mg_prev, mo_prev = mg_n1, mo_n1
    cond = krog_n1 < 1e-3                                                        # :897
    for i in range(n):
        pa, pb = grid[i], grid[i + 1]
        v, _ = pvt_eval(pb, tab, cfg, props=(0, 1, 2, 3, 4, 5))
        invBg1, invBo1, invug1, invuo1, Rs1, Rv1 = (v[j] for j in range(6))

        def cost(s):                                                             # :899-908 (well_id == 1 at a connection)
            krog, krgo = corey_krog_krgo_t(s, cfg, dtype)
            mgg = krgo * invBg1 * invug1
            mgo = krog * invBo1 * invuo1 * Rs1
            moo = krog * invBo1 * invuo1
            mog = krgo * invBg1 * invug1 * Rv1
            return (moo + mog) * mg_n1 - mo_n1 * (mgg + mgo)

        if solver == "newton":
            s1 = solve_newton_sg(cost, sg, n_root_iter, sg_max)
        else:
            s1 = solve_bracket_sg(cost, sg, n_root_iter, 1e-6, sg_max)
        s1 = torch.where(cond, torch.ones_like(s1) * sg_max, s1)                 # :912
        krog1, krgo1 = corey_krog_krgo_t(s1, cfg, dtype)
        mg1 = krgo1 * invBg1 * invug1 + krog1 * invBo1 * invuo1 * Rs1            # :926-930
        mo1 = krog1 * invBo1 * invuo1 + krgo1 * invBg1 * invug1 * Rv1
        dp = pa - pb
        sum_g = sum_g + 0.5 * (mg_prev + mg1) * dp                               # :937
        sum_o = sum_o + 0.5 * (mo_prev + mo1) * dp * 1.0                         # :938
        mg_prev, mo_prev = mg1, mo1
    return sum_g, sum_o


def wells_gc(p_cell, sg_cell, kx_cell, t_days, tab: SplineTable, cfg: OracleConfig, dtype, solver="newton"):
    """WellRatesPressure.compute_rates_and_bhp, fluid_type 'GC', non-iterative control, with or without the
    blocking-factor integral (well_rate_bhp_Subclassed.py:727-837, 614-724, 840-960, 963-1034), at the connection cells.

    returns (qgg, qgo, qoo, qog) each (B,nw), pwf (B,nw)"""
    dt = dtype
    wells = cfg.wells
    nw = len(wells)
    as_t = lambda a: torch.as_tensor(np.asarray(a, dtype=np.float64), dtype=dt).reshape(1, nw)
    rw = as_t([w.wellbore_radius for w in wells])
    hc = as_t([w.completion_ratio for w in wells])
    q_t = as_t([well_control_value(w) for w in wells])
    pmin = as_t([w.minimum_bhp for w in wells])
    shut = torch.as_tensor(shutin_open_mask(t_days, wells), dtype=dt)
    ck = shut * peaceman_static(kx_cell, cfg, rw, hc, dt)                        # :788
    p = p_cell
    krog, krgo = corey_krog_krgo_t(sg_cell, cfg, dt)                             # :791
    v, _ = pvt_eval(p, tab, cfg, props=(0, 1, 2, 3, 4, 5))                       # :794-795
    invBg, invBo, invug, invuo, Rs, Rv = (v[i] for i in range(6))
    mgg = krgo * invBg * invug                                                   # :802-807
    mgo = krog * invBo * invuo * Rs
    moo = krog * invBo * invuo
    mog = krgo * invBg * invug * Rv
    mg = mgg + mgo
    mo = moo + mog
    tiny = 1e-12
    one = torch.ones_like(p)
    zero = torch.zeros_like(p)
    blk = bool(cfg.use_blocking_factor)

    def integrals(pwf_):
        if blk:
            return blocking_integral_gc(p, sg_cell, pwf_, tab, cfg, krog, mg, mo, dt, solver=solver)
        return one, one                                                          # :955-959
    def phase_rates(pwf_):                                                       # _compute_phase_rates (:963-1007)
        ig, io = integrals(pwf_)
        dp = p - pwf_ + tiny                                                     # :987
        blk_g = _dnn(ig, mg * dp) if blk else ig                                 # :990-995
        blk_o = _dnn(io, mo * dp) if blk else io
        qg_max2 = ck * blk_g * mg * dp                                           # :997
        qo_max2 = ck * blk_o * mo * dp                                           # :998
        qg_ = _tf_maximum(_tf_minimum(q_t.expand_as(p), qg_max2), zero)          # :1001
        qo_target = qg_ * (1.0 / (Rv + tiny))                                    # :1004
        return qg_, _tf_maximum(_tf_minimum(qo_target, qo_max2), zero)           # :1005

    if cfg.use_non_iterative:
        # ---- _non_iterative_method (:614-724)
        ig_max, _io_max = integrals(pmin.expand_as(p))
        dp_max = p - pmin + tiny                                                 # :650
        blk_g_max = _dnn(ig_max, mg * dp_max) if blk else ig_max                 # :654-657
        qg_max = ck * blk_g_max * mg * dp_max                                    # :662 (well_id == 1)
        qg_opt = _tf_maximum(_tf_minimum(q_t.expand_as(p), qg_max), zero)        # :666
        lam = _tf_clip(_dnn(qg_opt, ck * blk_g_max * mg), zero, blk_g_max)       # :699
        pwf = _tf_clip(p - lam * dp_max, pmin.expand_as(p), p)                   # :721-723
    else:
        pwf = bhp_newton(lambda pw: phase_rates(pw)[0], p, pmin.expand_as(p), q_t.expand_as(p), cfg)
    qg, qo = phase_rates(pwf)
    # ---- _split_condensate_components (:1010-1034)
    denom_g = mgg + mgo + tiny
    denom_o = moo + mog + tiny
    return (qg * (mgg / denom_g), qg * (mgo / denom_g), qo * (moo / denom_o), qo * (mog / denom_o)), pwf


def gc_residual(cfg: OracleConfig, tab: SplineTable, kx, p0, p1, sg0, sg1, so0, so1, dt1, dt2, t_days, sample_real,
                dtype=torch.float32):
    """physics_error_gas_oil (physics_loss.py:319-693) given the networks' outputs p, Sg, So at both time
    levels.  tab must hold the 7 GC properties in GC_PROPS order.  Differentiable w.r.t.
    p0, p1, sg0, sg1, so0, so1, dt1, dt2."""
    dt = dtype
    T = lambda v: torch.tensor(v, dtype=dt)
    B = p0.shape[0]
    sr = torch.as_tensor(np.asarray(sample_real), dtype=torch.long)
    kxb = kx.to(dt)
    st = {k: v.index_select(0, sr) for k, v in dg_static(kxb, cfg).items()}       # :283-284 (+ z faces)
    dx, dy, dz = T(cfg.dx), T(cfg.dy), T(cfg.dz)
    dv = dx * dy * dz                                                        # :255
    C, Dc = T(cfg.C), T(cfg.Dc)
    phi = T(cfg.phi)
    cf = torch.tensor(float(rock_compressibility(cfg.phi)), dtype=dt) if dt == torch.float32 else \
        T(97.32e-6 / (1 + 55.8721 * cfg.phi ** 1.428586))                    # :288
    d1 = dt1.reshape(B, 1, 1, 1)
    d2 = dt2.reshape(B, 1, 1, 1)

    # PVT: values at n0 and n1, dp-derivatives at n0 of invBg, invBo, Rs, Rv      :324-372, 506-514
    v0, dv0 = pvt_eval(p0, tab, cfg, props=(0, 1, 4, 5), need_deriv=(0, 1, 4, 5))
    invBg0, invBo0, Rs0, Rv0 = v0[0], v0[1], v0[4], v0[5]
    dinvBg0, dinvBo0, dRs0, dRv0 = dv0[0], dv0[1], dv0[4], dv0[5]
    v1, _ = pvt_eval(p1, tab, cfg, props=(0, 1, 2, 3, 4, 5))
    invBg1, invBo1, invug1, invuo1, Rs1, Rv1 = (v1[i] for i in range(6))
    RsinvBo0 = Rs0 * invBo0                                                  # :343
    RvinvBg0 = Rv0 * invBg0                                                  # :344
    Mgg = invBg1 * invug1                                                    # :386 invBgug_n1
    Moo = invBo1 * invuo1                                                    # :387 invBouo_n1
    RsinvBo1 = Rs1 * invBo1                                                  # :388
    RvinvBg1 = Rv1 * invBg1                                                  # :389
    Mgo = Rs1 * invBo1 * invuo1                                              # :390 RsinvBouo_n1
    Mog = Rv1 * invBg1 * invug1                                              # :391 RvinvBgug_n1

    # masses and truncation terms                                            :419-441
    rho1 = 1.0 + _dnn(d2, d1)
    mg0 = phi * ((invBg0 * sg0) + (RsinvBo0 * so0))
    mo0 = phi * ((invBo0 * so0) + (RvinvBg0 * sg0))
    mg1 = phi * ((invBg1 * sg1) + (RsinvBo1 * so1))
    mo1 = phi * ((invBo1 * so1) + (RvinvBg1 * sg1))
    mg2 = (mg1 - mg0) * rho1 + mg0
    mo2 = (mo1 - mo0) * rho1 + mo0
    rte = 1e-7 * (1 / 4)                                                     # :439
    den = (d1 * d2) + d2 * d2
    trn_g = (dv / Dc) * ((rte / d1) + ((((d2 * mg0) + (d1 * mg2)) - ((d1 + d2) * mg1)) / den))   # :440
    trn_o = (dv / Dc) * ((rte / d1) + ((((d2 * mo0) + (d1 * mo2)) - ((d1 + d2) * mo1)) / den))   # :441

    krog1, krgo1 = corey_krog_krgo_t(sg1, cfg, dt)                           # :457

    # wells (rates are model outputs in the legacy code, :461; here WellRatesPressure is evaluated at the connections)
    nw = len(cfg.wells)
    zf = torch.zeros_like(p1)
    q4 = [zf, zf, zf, zf]
    pwf = zf
    mask = torch.zeros((cfg.D, cfg.H, cfg.W), dtype=dt)
    qw4 = pwfw = None
    if nw:
        flat = torch.as_tensor(well_flat_index(cfg.wells, cfg.D, cfg.H, cfg.W).astype(np.int64))
        N = cfg.D * cfg.H * cfg.W
        p_cell = p1.reshape(B, -1).index_select(1, flat)
        sg_cell = sg1.reshape(B, -1).index_select(1, flat)
        kx_cell = kxb.index_select(0, sr).reshape(B, -1).index_select(1, flat)
        qw4, pwfw = wells_gc(p_cell, sg_cell, kx_cell, np.asarray(t_days), tab, cfg, dt)
        q4 = [torch.zeros((B, N), dtype=dt).index_add(1, flat, q).reshape(p1.shape) for q in qw4]
        pwf = torch.zeros((B, N), dtype=dt).index_add(1, flat, pwfw).reshape(p1.shape)
        mask = torch.zeros(N, dtype=dt).index_add(0, flat, torch.ones(nw, dtype=dt)).reshape(cfg.D, cfg.H, cfg.W)
    qfg, qdg, qfo, qvo = q4

    # chord slopes                                                           :465-466
    dpc = p1 - p0
    dSg = _dnn(sg1 - sg0, dpc)
    dSo = _dnn(so1 - so0, dpc)
    # product-rule PVT derivatives                                           :506-514
    dRsinvBo = (Rs0 * dinvBo0) + (invBo0 * dRs0)
    dRvinvBg = (Rv0 * dinvBg0) + (invBg0 * dRv0)

    # neighbours and face values
    def faces(X):                                                            # :517-525
        return dict(W=(X + _nbr(X, -1, -1)) / 2.0, E=(_nbr(X, -1, +1) + X) / 2.0,
                    S=(X + _nbr(X, -2, -1)) / 2.0, N=(_nbr(X, -2, +1) + X) / 2.0,
                    D=(X + _nbr(X, -3, -1)) / 2.0, U=(_nbr(X, -3, +1) + X) / 2.0)
    pn = dict(W=_nbr(p1, -1, -1), E=_nbr(p1, -1, +1), S=_nbr(p1, -2, -1), N=_nbr(p1, -2, +1),
              D=_nbr(p1, -3, -1), U=_nbr(p1, -3, +1))
    # potentials as written: "plus" faces nbr - cell, "minus" faces cell - nbr      :538-541
    pot = dict(E=pn["E"] - p1, W=p1 - pn["W"], N=pn["N"] - p1, S=p1 - pn["S"], U=pn["U"] - p1, D=p1 - pn["D"])
    off = dict(W=(-1, -1), E=(-1, +1), S=(-2, -1), N=(-2, +1), D=(-3, -1), U=(-3, +1))

    def upstream(kr):                                                        # :543-551 (selects carry no gradient)
        out = {}
        for f in "WESNDU":
            le = (pot[f] <= 0).to(dt)
            gt = (pot[f] > 0).to(dt)
            out[f] = le * kr + gt * _nbr(kr, *off[f])
        return out
    krg_f, kro_f = upstream(krgo1), upstream(krog1)
    fMgg, fMoo, fMog, fMgo = faces(Mgg), faces(Moo), faces(Mog), faces(Mgo)
    kf = dict(W=st["kW"], E=st["kE"], S=st["kS"], N=st["kN"], D=st["kD"], U=st["kU"])
    idl = dict(W=1.0 / dx, E=1.0 / dx, S=1.0 / dy, N=1.0 / dy, D=1.0 / dz, U=1.0 / dz)

    def coef(kr_f, M_f):                                                     # :563-583
        return {f: C * kf[f] * (kr_f[f] * M_f[f]) * idl[f] * idl[f] for f in "WESNDU"}
    agg, ago, aoo, aog = coef(krg_f, fMgg), coef(kro_f, fMgo), coef(kro_f, fMoo), coef(krg_f, fMog)

    phicf = phi * cf
    cprgg, cprgo, cproo, cprog = phicf * invBg0, phicf * RsinvBo0, phicf * invBo0, phicf * RvinvBg0   # :557-560
    idt = 1.0 / (Dc * d1)
    cpgg = idt * ((phi * invBg1 * dSg) + sg0 * ((phi * dinvBg0) + cprgg)) * dpc                        # :572
    cpgo = idt * ((phi * RsinvBo1 * dSo) + so0 * ((phi * dRsinvBo) + cprgo)) * dpc                     # :573
    cpoo = idt * ((phi * invBo1 * dSo) + so0 * ((phi * dinvBo0) + cproo)) * dpc                        # :585
    cpog = idt * ((phi * RvinvBg1 * dSg) + sg0 * ((phi * dRvinvBg) + cprog)) * dpc                     # :586

    def divq(a, q):                                                          # :590-611
        s = (-a["W"] * pn["W"]) + (-a["S"] * pn["S"]) + ((a["W"] + a["S"] + a["E"] + a["N"]) * p1) \
            + (-a["E"] * pn["E"]) + (-a["N"] * pn["N"])
        s = s + ((a["D"] * (p1 - pn["D"])) + (a["U"] * (p1 - pn["U"])))      # 3-D extension, == 0 for Nz = 1
        s = s + (q / dv)
        return dv * s
    divq_gg, divq_go, divq_oo, divq_og = divq(agg, qfg), divq(ago, qdg), divq(aoo, qfo), divq(aog, qvo)
    dom_gg = divq_gg + dv * cpgg                                             # :597-600
    dom_go = divq_go + dv * cpgo
    dom_oo = divq_oo + dv * cpoo                                             # :614-620
    dom_og = divq_og + dv * cpog
    dom = (dom_gg + dom_go) + (dom_oo + dom_og)                              # :638
    trn = trn_g + trn_o                                                      # :637
    ibc = mask * ((divq_gg + divq_go) + (divq_oo + divq_og))                 # :650
    mfac = dv * idt * phi                                                    # :655-659
    mbc_gg = mfac * ((sg1 * invBg1) - (sg0 * invBg0))
    mbc_go = mfac * ((so1 * RsinvBo1) - (so0 * RsinvBo0))
    mbc_oo = mfac * ((so1 * invBo1) - (so0 * invBo0))
    mbc_og = mfac * ((sg1 * RvinvBg1) - (sg0 * RvinvBg0))
    sm = lambda x: x.sum(dim=(1, 2, 3))
    mbc_g = (-sm(qfg + qdg)) - sm(mbc_gg + mbc_go)                           # :661
    mbc_o = (-sm(qfo + qvo)) - sm(mbc_oo + mbc_og)                           # :662
    mbc = mbc_g + mbc_o                                                      # :665
    return dict(dom=dom, ibc=ibc, mbc=mbc, cmbc=trn, divq=(divq_gg + divq_go) + (divq_oo + divq_og), q4=q4, pwf=pwf,
                qw4=qw4, pwfw=pwfw, krog1=krog1, krgo1=krgo1, mask=mask,
                parts=dict(divq_gg=divq_gg, divq_go=divq_go, divq_oo=divq_oo, divq_og=divq_og,
                           cpgg=cpgg, cpgo=cpgo, cpoo=cpoo, cpog=cpog, trn_g=trn_g, trn_o=trn_o))


def gc_loss_terms(res) -> torch.Tensor:
    """SSE per term (physics_loss.py:787-807): dom, ibc, mbc and cmbc (= truncation term, :680)."""
    z = torch.zeros((), dtype=res["dom"].dtype)
    return torch.stack([(res["dom"] ** 2).sum(), (res["ibc"] ** 2).sum(), (res["mbc"] ** 2).sum(), z, z, z, z,
                        (res["cmbc"] ** 2).sum()])


def gc_counts(cfg: OracleConfig, B: int) -> np.ndarray:
    n = B * cfg.D * cfg.H * cfg.W
    return np.array([n, n, n, 0, 0, 0, 0, n], dtype=np.float64)       # mbc counted with the ic field's shape (:830)


def gc_forward_backward(cfg, tab, kx, p0, p1, sg0, sg1, so0, so1, dt1, dt2, t_days, sample_real, weights,
                        dtype=torch.float32):
    """Loss terms and the gradient of sum_k weights[k]*SSE_k w.r.t. p0,p1,sg0,sg1,so0,so1,dt1,dt2."""
    tt = lambda a: torch.as_tensor(np.asarray(a), dtype=dtype).clone().requires_grad_(True)
    ins = [tt(a) for a in (p0, p1, sg0, sg1, so0, so1, dt1, dt2)]
    kxt = torch.as_tensor(np.asarray(kx), dtype=dtype)
    res = gc_residual(cfg, tab, kxt, *ins, t_days, sample_real, dtype=dtype)
    terms = gc_loss_terms(res)
    wt = torch.as_tensor(np.asarray(weights, dtype=np.float64), dtype=dtype)
    loss = (terms * wt).sum()
    grads = torch.autograd.grad(loss, ins, allow_unused=True)
    out = dict(dom=res["dom"].detach().numpy(), ibc=res["ibc"].detach().numpy(), mbc=res["mbc"].detach().numpy(),
               cmbc=res["cmbc"].detach().numpy(), terms=terms.detach().numpy(), loss=float(loss.detach()))
    if res["qw4"] is not None:
        out["qw4"] = np.stack([q.detach().numpy() for q in res["qw4"]])
        out["pwfw"] = res["pwfw"].detach().numpy()
    for name, g, ref in zip(("gp0", "gp1", "gsg0", "gsg1", "gso0", "gso1", "gdt1", "gdt2"), grads, ins):
        out[name] = (torch.zeros_like(ref) if g is None else g).numpy()
    return out


# --------------------------------------------------------------------------------------------
# (de)normalisation  (data_processing/data_processing_utils.py:1065-1183)
# --------------------------------------------------------------------------------------------
def denorm_linear(x, vmin, vmax, lo=-1.0, hi=1.0, dtype=np.float32):
    """_lin_rev (:1108-1109): (max-min)*((x-lo)/(hi-lo)) + min."""
    t = dtype
    x = np.asarray(x, dtype=t)
    return ((t(vmax) - t(vmin)) * ((x - t(lo)) / (t(hi) - t(lo))) + t(vmin)).astype(t)


def denorm_log(x, vmin, vmax, lo=-1.0, hi=1.0, dtype=np.float32):
    """_lnk_rev log branch (:1100-1102): exp(log(max/min)*((x-lo)/(hi-lo)) + log(min))."""
    t = dtype
    x = np.asarray(x, dtype=t)
    return np.exp(np.log(t(vmax) / t(vmin)) * ((x - t(lo)) / (t(hi) - t(lo))) + np.log(t(vmin))).astype(t)


def norm_diff_linear(d, vmin, vmax, lo=-1.0, hi=1.0, dtype=np.float32):
    """normalize_diff linear branch (:1166-1168): (hi-lo)/(max-min)*diff."""
    t = dtype
    return ((t(hi) - t(lo)) / (t(vmax) - t(vmin)) * np.asarray(d, dtype=t)).astype(t)


# ------------------------------------------------------------------------------------------------
# glue either side of the residual: HardLayer and the time-step mean
# ------------------------------------------------------------------------------------------------
def hard_layer_t(y, tn, expo, init_value=1.0, t_lo=-1.0, t_hi=1.0):
    """HardLayer.call (Hard_Layer_Subclassed.py:219-242) with the example's configuration (no rbf, no rectifier,
    identity activations, identity nonormalize_func): output = init_value - alpha_t ** kernel_exponent * p,
    alpha_t = (t - norm_limits[0]) / (norm_limits[1] - norm_limits[0]).  torch tensors (autograd-capable);
    y (B,D,H,W), tn (B,), expo (D,H,W).  tf.pow's gradient w.r.t. the exponent uses log of the SAFE base
    (where(x > 0, x, 1)): torch.pow's would give nan at alpha_t = 0, hence the explicit Function."""
    at = ((tn - t_lo) / (t_hi - t_lo)).view(-1, 1, 1, 1)
    alpha = _TfPow.apply(at.expand_as(y), expo.expand_as(y))
    return init_value - alpha * y


class _TfPow(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, e):
        z = torch.pow(x, e)
        ctx.save_for_backward(x, e, z)
        return z

    @staticmethod
    def backward(ctx, g):
        x, e, z = ctx.saved_tensors
        safe = torch.where(x > 0, x, torch.ones_like(x))
        gx = g * e * torch.pow(x, e - 1)
        ge = g * z * torch.where(x > 0, torch.log(safe), torch.zeros_like(x))
        return gx, ge


def time_step_mean_t(dtf):
    """tstep = reduce_mean(fac, axis=[1,2,3]) (physics_loss.py:102,122)"""
    return dtf.reshape(dtf.shape[0], -1).mean(dim=1)
