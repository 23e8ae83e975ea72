"""ctypes binding of libsrm_physics.so (include/srm_physics.h).

This is the reference-side stub a maintainer would add: plain pointers and
sizes, no torch types in the signatures.  torch is used by the callers only to own device memory
and streams.  There is no CPU fallback: if the shared library is missing or no CUDA device is
usable, loading / handle creation raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional, Sequence

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# SRM_PHYSICS_LIB: kernel-tuning experiments load an alternative build of the same library
LIB_PATH = os.environ.get("SRM_PHYSICS_LIB") or os.path.join(_HERE, "libsrm_physics.so")

SRM_ABI_VERSION = 5
SRM_N_TERMS = 8
TERM_NAMES = ("dom", "ibc", "mbc", "tde", "obc", "ic", "td", "cmbc")
SRM_FLUID_DG, SRM_FLUID_GC = 0, 1
SRM_PVT_SPLINE, SRM_PVT_POLYNOMIAL = 0, 1
SRM_NUMERICS_REFERENCE, SRM_NUMERICS_CLOSED_FORM = 0, 1
SRM_FLAG_SAVE_FOR_BACKWARD = 1
SRM_ROOT_NEWTON, SRM_ROOT_BRACKET = 0, 1

EXPORTS = (
    "srm_version", "srm_last_error", "srm_create", "srm_destroy", "srm_workspace_bytes",
    "srm_pvt_eval", "srm_denormalize_log", "srm_selftest_rounding", "srm_wells", "srm_forward", "srm_backward",
    "srm_relperm", "srm_forward_gc", "srm_backward_gc", "srm_glue_workspace_bytes", "srm_glue_forward", "srm_glue_backward", "srm_gather_rows",
    "srm_features_forward", "srm_features_backward", "srm_weave_features",
)


class SrmWell(C.Structure):
    _fields_ = [("i", C.c_int32), ("j", C.c_int32), ("k", C.c_int32), ("q_target", C.c_float),
                ("pwf_min", C.c_float), ("rw", C.c_float), ("hc", C.c_float),
                ("shut_start", C.c_float), ("shut_stop", C.c_float)]


class SrmConfig(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("device", C.c_int32),
        ("D", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
        ("dx", C.c_float), ("dy", C.c_float), ("dz", C.c_float),
        ("C", C.c_float), ("Dc", C.c_float),
        ("phi", C.c_float), ("cf", C.c_float), ("Sgi", C.c_float), ("krg", C.c_float),
        ("kx_ky", C.c_float), ("kv_kh", C.c_float),
        ("fluid_type", C.c_int32),
        ("pvt_method", C.c_int32), ("spline_order", C.c_int32), ("n_knots", C.c_int32), ("n_props", C.c_int32),
        ("knots", C.POINTER(C.c_float)), ("spline_w", C.POINTER(C.c_float)), ("spline_v", C.POINTER(C.c_float)),
        ("p_min", C.c_float), ("p_max", C.c_float),
        ("n_wells", C.c_int32), ("wells", C.POINTER(SrmWell)),
        ("use_blocking_factor", C.c_int32), ("n_intervals", C.c_int32),
        ("numerics", C.c_int32), ("tde_in_dom", C.c_int32),
        ("pvt_lut", C.c_int32), ("lut_p_lo", C.c_float), ("lut_p_hi", C.c_float),
        ("Swmin", C.c_float), ("Sorg", C.c_float), ("Sgc", C.c_float), ("Socr", C.c_float),
        ("kro_Somax", C.c_float), ("krg_Sorg", C.c_float), ("krg_Swmin", C.c_float), ("nog", C.c_float), ("ng", C.c_float),
        ("root_solver", C.c_int32), ("n_root_iter", C.c_int32),
        ("bhp_iterative", C.c_int32), ("bhp_max_iters", C.c_int32), ("bhp_tol", C.c_float),
    ]


_lib = None


def load_library(path: Optional[str] = None):
    """dlopen the library and declare every prototype of include/srm_physics.h."""
    global _lib
    if _lib is not None and path is None:
        return _lib
    path = path or LIB_PATH
    if not os.path.exists(path):
        raise RuntimeError(
            f"{path} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a).  There is no CPU fallback for the physics-loss path.")
    lib = C.CDLL(path)
    vp, i32, i64, fp = C.c_void_p, C.c_int32, C.c_int64, C.c_float
    lib.srm_version.restype = C.c_int
    lib.srm_version.argtypes = []
    lib.srm_last_error.restype = C.c_char_p
    lib.srm_last_error.argtypes = []
    lib.srm_create.restype = C.c_int
    lib.srm_create.argtypes = [C.POINTER(SrmConfig), C.POINTER(vp)]
    lib.srm_destroy.restype = None
    lib.srm_destroy.argtypes = [vp]
    lib.srm_workspace_bytes.restype = C.c_size_t
    lib.srm_workspace_bytes.argtypes = [vp, i32, i32, i32]
    lib.srm_pvt_eval.restype = C.c_int
    lib.srm_pvt_eval.argtypes = [vp, i64, vp, vp, vp, vp]
    lib.srm_denormalize_log.restype = C.c_int
    lib.srm_denormalize_log.argtypes = [i32, i64, vp, fp, fp, fp, fp, vp, vp]
    lib.srm_selftest_rounding.restype = C.c_int
    lib.srm_selftest_rounding.argtypes = [i32, i64, C.c_uint64, C.POINTER(C.c_int64), vp]
    lib.srm_wells.restype = C.c_int
    lib.srm_wells.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.srm_forward.restype = C.c_int
    lib.srm_forward.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_size_t, i32, vp]
    lib.srm_backward.restype = C.c_int
    lib.srm_backward.argtypes = [vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, C.c_size_t, i32, vp]
    lib.srm_relperm.restype = C.c_int
    lib.srm_relperm.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp]
    lib.srm_forward_gc.restype = C.c_int
    lib.srm_forward_gc.argtypes = [vp, i32, i32] + [vp] * 11 + [vp] * 4 + [vp, C.c_size_t, i32, vp]
    lib.srm_backward_gc.restype = C.c_int
    lib.srm_backward_gc.argtypes = [vp, i32, i32] + [vp] * 11 + [vp] + [vp] * 8 + [vp, C.c_size_t, i32, vp]
    lib.srm_glue_workspace_bytes.restype = C.c_size_t
    lib.srm_glue_workspace_bytes.argtypes = [i32]
    lib.srm_glue_forward.restype = C.c_int
    lib.srm_glue_forward.argtypes = [vp, i32, fp, fp, fp] + [vp] * 11 + [vp, C.c_size_t, vp]
    lib.srm_glue_backward.restype = C.c_int
    lib.srm_glue_backward.argtypes = [vp, i32, fp, fp, fp] + [vp] * 16 + [vp]
    lib.srm_gather_rows.restype = C.c_int
    lib.srm_gather_rows.argtypes = [i32, vp, vp, i64, i64, i64, vp, vp]
    lib.srm_features_forward.restype = C.c_int
    lib.srm_features_forward.argtypes = [i32, vp, vp, i32, i64, i32, i32, i32, fp, fp, fp, fp, vp, vp, vp]
    lib.srm_features_backward.restype = C.c_int
    lib.srm_features_backward.argtypes = [i32, vp, i32, i64, i32, i32, vp, vp]
    lib.srm_weave_features.restype = C.c_int
    lib.srm_weave_features.argtypes = [i32, i32, i32, i64, vp, vp, vp, vp, vp, vp, fp, fp, vp, vp]
    if lib.srm_version() != SRM_ABI_VERSION:
        raise RuntimeError(f"libsrm_physics ABI {lib.srm_version()} != binding {SRM_ABI_VERSION}")
    if path == LIB_PATH:
        _lib = lib
    return lib


class SrmError(RuntimeError):
    pass


def check(lib, rc: int, what: str):
    if rc != 0:
        msg = lib.srm_last_error()
        raise SrmError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")


def _fptr(a: np.ndarray):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def make_config(*, device: int, D: int, H: int, W: int, dx: float, dy: float, dz: float, C_: float, Dc: float,
                phi: float, cf: float, Sgi: float, krg: float, kx_ky: float, kv_kh: float,
                knots: np.ndarray, spline_w: np.ndarray, spline_v: np.ndarray, spline_order: int,
                p_min: float, p_max: float, wells: Sequence[dict], use_blocking_factor: bool, n_intervals: int,
                numerics: int, tde_in_dom: bool, fluid_type: int = SRM_FLUID_DG, pvt_lut: bool = False,
                lut_range: Optional[Sequence[float]] = None, end_points: Optional[dict] = None,
                corey_exponents: Optional[dict] = None, pvt_method: int = SRM_PVT_SPLINE,
                root_solver: str = "newton", n_root_iter: int = 20,
                use_non_iterative: bool = True, bhp_max_iters: int = 10, bhp_tol: float = 1e-6):
    """Fill an SrmConfig; returns (cfg, keepalive) -- keepalive owns the host arrays cfg points into."""
    knots = np.ascontiguousarray(knots, dtype=np.float32)
    spline_w = np.ascontiguousarray(spline_w, dtype=np.float32)
    spline_v = np.ascontiguousarray(spline_v, dtype=np.float32)
    if pvt_method == SRM_PVT_SPLINE:
        assert spline_w.shape == (spline_v.shape[0], knots.size) and spline_v.shape[1] == 2
    warr = (SrmWell * max(1, len(wells)))()
    for n, w in enumerate(wells):
        warr[n] = SrmWell(int(w["i"]), int(w["j"]), int(w["k"]), float(w["q_target"]), float(w["pwf_min"]),
                          float(w["rw"]), float(w["hc"]), float(w["shut_start"]), float(w["shut_stop"]))
    cfg = SrmConfig(
        abi_version=SRM_ABI_VERSION, device=device, D=D, H=H, W=W, dx=dx, dy=dy, dz=dz, C=C_, Dc=Dc,
        phi=phi, cf=cf, Sgi=Sgi, krg=krg, kx_ky=kx_ky, kv_kh=kv_kh, fluid_type=fluid_type,
        pvt_method=int(pvt_method), spline_order=spline_order,
        n_knots=knots.size if pvt_method == SRM_PVT_SPLINE else spline_w.shape[1], n_props=spline_w.shape[0],
        knots=_fptr(knots), spline_w=_fptr(spline_w), spline_v=_fptr(spline_v), p_min=p_min, p_max=p_max,
        n_wells=len(wells), wells=warr, use_blocking_factor=int(bool(use_blocking_factor)),
        n_intervals=int(n_intervals), numerics=int(numerics), tde_in_dom=int(bool(tde_in_dom)),
        pvt_lut=int(bool(pvt_lut)), lut_p_lo=float(lut_range[0]) if lut_range else 0.0,
        lut_p_hi=float(lut_range[1]) if lut_range else 0.0,
        root_solver=SRM_ROOT_NEWTON if str(root_solver).lower() == "newton" else SRM_ROOT_BRACKET, n_root_iter=int(n_root_iter),
        bhp_iterative=int(not use_non_iterative), bhp_max_iters=int(bhp_max_iters), bhp_tol=float(bhp_tol))
    if end_points is not None:
        for k in ("Swmin", "Sorg", "Sgc", "Socr", "kro_Somax", "krg_Sorg", "krg_Swmin"):
            setattr(cfg, k, float(end_points[k]))
        cfg.nog = float(corey_exponents["nog"])
        cfg.ng = float(corey_exponents["ng"])
    return cfg, (knots, spline_w, spline_v, warr)
