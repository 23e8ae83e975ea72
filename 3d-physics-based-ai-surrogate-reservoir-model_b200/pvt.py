"""Host side of the PVT path: table loading, the one-off polyharmonic solve, and a ``PVTLayer``
mirror of the reference interface (PVT_Layer_Subclassed.py:23-216) whose ``__call__`` runs the CUDA
kernel behind ``srm_pvt_eval``.

The reference solves the (n+2)x(n+2) interpolation system inside *every* call
(polyhm_splines.py:180); it depends on constants only, so it is solved once here, in fp32, and the
weights travel to the device as data.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, Optional, Sequence, Tuple

import numpy as np

EPSILON = 1e-10                                   # polyhm_splines.py:6
DG_PROPERTIES = ("invBg", "invug")                # PVT_Layer_Subclassed.py:69-70
GC_PROPERTIES = ("invBg", "invBo", "invug", "invuo", "Rs", "Rv", "Vro")   # :71-72
_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), "data", "pvt_table.npz")


class PVTTable:
    """Column store with the case-insensitive ``lookup`` the reference's DataSummary offers
    (data_processing/data_processing_utils.py:873-879), so DG resolves 'invBg' -> 'InvBg' etc."""

    def __init__(self, columns: Dict[str, np.ndarray]):
        self.columns = {str(k): np.asarray(v, dtype=np.float32) for k, v in columns.items()}
        self._lower = {k.lower(): k for k in self.columns}

    def lookup(self, key: str) -> np.ndarray:
        k = self._lower.get(str(key).lower())
        if k is None:
            raise KeyError(f"PVT table has no column {key!r}; columns: {sorted(self.columns)}")
        return self.columns[k]


def load_default_pvt_table(path: Optional[str] = None) -> PVTTable:
    """The reference's pvt_data.df (37 x 10 fp32), shipped bit-exactly as data/pvt_table.npz
    (exported by tests/golden/make_pvt_table.py)."""
    z = np.load(path or _DATA)
    return PVTTable({str(k): z["table"][i] for i, k in enumerate(z["columns"])})


def _phi(r: np.ndarray, order: int) -> np.ndarray:
    rs = np.maximum(r, r.dtype.type(EPSILON))     # polyhm_splines.py:78
    if order == 1:
        return np.sqrt(rs)
    if order == 2:
        return r.dtype.type(0.5) * rs * np.log(rs)
    raise ValueError(f"unsupported spline order {order} (1 or 2)")


def solve_polyharmonic(c: np.ndarray, f: np.ndarray, order: int = 1, regularization_weight: float = 0.001
                       ) -> Tuple[np.ndarray, np.ndarray]:
    """Solve [[A + lam*I, B], [B^T, 0]] [w; v] = [f; 0], B = [c, 1], in fp32
    (polyhm_splines.py:103-135).  Returns (w[n], v[2])."""
    t = np.float32
    c = np.asarray(c, dtype=t).reshape(-1)
    f = np.asarray(f, dtype=t).reshape(-1)
    n = c.size
    if f.size != n:
        raise ValueError("knots and values differ in length")
    cn = c * c
    xy = (c[:, None] * c[None, :]).astype(t)
    r = ((cn[:, None] - t(2) * xy) + cn[None, :]).astype(t)
    a = _phi(r, order).astype(t)
    if regularization_weight > 0:
        a = a + t(regularization_weight) * np.eye(n, dtype=t)
    lhs = np.zeros((n + 2, n + 2), dtype=t)
    lhs[:n, :n] = a
    lhs[:n, n] = c
    lhs[:n, n + 1] = 1
    lhs[n, :n] = c
    lhs[n + 1, :n] = 1
    rhs = np.zeros(n + 2, dtype=t)
    rhs[:n] = f
    sol = np.linalg.solve(lhs, rhs).astype(t)
    return sol[:n].copy(), sol[n:].copy()


@dataclass
class SplineTables:
    knots: np.ndarray        # (n,)
    w: np.ndarray            # (P, n)
    v: np.ndarray            # (P, 2)
    order: int
    properties: Tuple[str, ...]


def build_spline_tables(table: PVTTable, properties: Sequence[str], order: int = 1,
                        regularization_weight: float = 0.001) -> SplineTables:
    """PVTLayer.build, spline branch (PVT_Layer_Subclassed.py:118-141): one spline per property on 'pre'."""
    c = table.lookup("pre")
    ws, vs = [], []
    for p in properties:
        w, v = solve_polyharmonic(c, table.lookup(p), order, regularization_weight)
        ws.append(w)
        vs.append(v)
    return SplineTables(knots=c.astype(np.float32), w=np.stack(ws), v=np.stack(vs), order=order,
                        properties=tuple(properties))


@dataclass
class PolynomialTables:
    """PVTLayer, fitting_method='polynomial' (PVT_Layer_Subclassed.py:77-87,218-266): value = sum_i a_i p^i per
    property; ``coefficients[q]`` = [a_0, a_1, ...] of property q (all properties padded to one length)."""
    coefficients: np.ndarray          # (P, n) float32
    properties: Tuple[str, ...]

    @property
    def w(self):                      # the engine reads .w for the property count
        return self.coefficients


def build_polynomial_tables(polynomial_config: Dict[str, Sequence[float]], properties: Sequence[str]) -> PolynomialTables:
    """from a ``polynomial_config`` dict shaped like default_configurations.py:231-234 (keys case-insensitive)"""
    low = {k.lower(): v for k, v in polynomial_config.items()}
    rows = []
    for p in properties:
        if p.lower() not in low:
            raise ValueError(f"Polynomial coefficients missing for property: {p}")     # PVT_Layer_Subclassed.py:84-86
        rows.append(list(low[p.lower()]))
    n = max(len(r) for r in rows)
    coef = np.zeros((len(rows), n), dtype=np.float32)
    for q, r in enumerate(rows):
        coef[q, :len(r)] = np.asarray(r, dtype=np.float32)
    return PolynomialTables(coefficients=coef, properties=tuple(properties))


class PVTLayer:
    """Drop-in for the reference's ``PVTLayer`` call contract (PVT_Layer_Subclassed.py:146-216):

        out = layer(p)      # p: (B, *spatial, 1) pressure  ->  out: (2, n_prop, B, *spatial, 1)

    ``out[0]`` are the property values, ``out[1]`` their derivatives w.r.t. the clamped pressure.
    ``engine`` is the SrmPhysics handle that owns the device tables (so the loss and the layer share them); the
    fitting method (spline or polynomial) is the one its tables were built for.
    """

    def __init__(self, engine, fluid_type: str = "DG", fitting_method: Optional[str] = None, name: str = "pvt_layer"):
        have = "polynomial" if isinstance(engine.tables, PolynomialTables) else "spline"
        if fitting_method is not None and fitting_method.lower() != have:
            raise ValueError(f"PVTLayer: the engine's tables are {have!r}, not {fitting_method!r}")
        self.engine = engine
        self.fluid_type = fluid_type.upper()
        self.fitting_method = have
        self.properties = list(DG_PROPERTIES if self.fluid_type == "DG" else GC_PROPERTIES)
        self.name = name
        self.trainable_variables = []

    def __call__(self, inputs, training: bool = False):
        import torch
        p = inputs.contiguous().to(torch.float32)
        val, der = self.engine.pvt_eval(p.reshape(-1))
        n_prop = val.shape[0]
        shape = (n_prop,) + tuple(p.shape)
        return torch.stack([val.reshape(shape), der.reshape(shape)], dim=0)

    call = __call__
