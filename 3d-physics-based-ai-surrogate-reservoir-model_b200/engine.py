"""SrmPhysics: a handle of libsrm_physics.so bound to one CUDA device, driven with torch tensors.

torch is plumbing here (device memory, streams, autograd glue); every number comes out of the
CUDA kernels behind the C ABI.  All methods raise if a tensor is not a contiguous fp32 CUDA tensor
on the handle's device -- there is no host fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib as L
from .config import PhysicsSpec
from .pvt import PolynomialTables, SplineTables

NUMERICS = {"reference": L.SRM_NUMERICS_REFERENCE, "closed_form": L.SRM_NUMERICS_CLOSED_FORM}


def _ptr(t: Optional[torch.Tensor]):
    return C.c_void_p(0 if t is None else t.data_ptr())


class SrmPhysics:
    def __init__(self, spec: PhysicsSpec, tables: SplineTables, device: int = 0, numerics: str = "reference",
                 pvt_lut: bool = False, lut_range=None):
        """pvt_lut: tabulate the reference-order PVT spline for every fp32 pressure of lut_range
        (default: the whole clamp range, 2.5 GB) at create; bit-identical results (srm_physics.h)."""
        if not torch.cuda.is_available():
            raise RuntimeError("SrmPhysics needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = L.load_library()
        self.spec = spec
        self.tables = tables
        self.device = torch.device("cuda", device)
        self.numerics = numerics
        if spec.fluid_type not in ("DG", "GC"):
            raise ValueError(f"fluid_type {spec.fluid_type!r}: 'DG' (dry gas) or 'GC' (gas condensate)")
        poly = isinstance(tables, PolynomialTables)
        cfg, self._keep = L.make_config(
            device=device, D=spec.D, H=spec.H, W=spec.W, dx=spec.dx, dy=spec.dy, dz=spec.dz, C_=spec.C, Dc=spec.Dc,
            phi=spec.phi, cf=spec.cf, Sgi=spec.Sgi, krg=spec.krg, kx_ky=spec.kx_ky, kv_kh=spec.kv_kh,
            knots=np.zeros(1, np.float32) if poly else tables.knots, spline_w=tables.w,
            spline_v=np.zeros((tables.w.shape[0], 2), np.float32) if poly else tables.v,
            spline_order=1 if poly else tables.order, pvt_method=L.SRM_PVT_POLYNOMIAL if poly else L.SRM_PVT_SPLINE,
            p_min=spec.p_min, p_max=spec.p_max, wells=[w.as_dict() for w in spec.wells],
            use_blocking_factor=spec.use_blocking_factor, n_intervals=spec.n_intervals,
            numerics=NUMERICS[numerics], tde_in_dom=spec.tde_in_dom, pvt_lut=pvt_lut, lut_range=lut_range,
            fluid_type=L.SRM_FLUID_GC if spec.fluid_type == "GC" else L.SRM_FLUID_DG,
            end_points=spec.end_points, corey_exponents=spec.corey_exponents,
            root_solver=spec.root_solver, n_root_iter=spec.n_root_iter,
            use_non_iterative=spec.use_non_iterative, bhp_max_iters=spec.max_iters, bhp_tol=spec.tol)
        self.pvt_lut = bool(pvt_lut) and numerics == "reference"
        h = C.c_void_p()
        L.check(self.lib, self.lib.srm_create(C.byref(cfg), C.byref(h)), "srm_create")
        self._h = h
        self.n_wells = len(spec.wells)
        self.n_props = tables.w.shape[0]
        self._ws = None
        self._ws_B = None
        self.launches = 0       # kernels launched through this handle (bench's gpu_launches claim)
        self.fluid = spec.fluid_type

    def close(self):
        if getattr(self, "_h", None):
            self.lib.srm_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------------------------------
    def _check(self, t: torch.Tensor, name: str, dtype=torch.float32):
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.device == self.device and t.dtype == dtype
                and t.is_contiguous()):
            raise ValueError(f"{name}: need a contiguous {dtype} CUDA tensor on {self.device}, got "
                             f"{getattr(t, 'dtype', type(t))} on {getattr(t, 'device', '?')}")

    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def workspace_bytes(self, B: int, R: int) -> int:
        return int(self.lib.srm_workspace_bytes(self._h, B, R, L.SRM_FLAG_SAVE_FOR_BACKWARD))

    def new_workspace(self, B: int, R: int) -> torch.Tensor:
        """a workspace the caller owns (pass it as ws= to forward / backward): CUDA graphs and pipelines keep theirs"""
        return torch.empty(self.workspace_bytes(B, R), dtype=torch.uint8, device=self.device)

    def workspace(self, B: int, R: int) -> torch.Tensor:
        """the engine's own cached workspace, re-allocated when the batch shape changes"""
        if self._ws is None or self._ws_B != (B, R):
            self._ws = None
            self._ws = self.new_workspace(B, R)
            self._ws_B = (B, R)
        return self._ws

    def _ws_for(self, ws, B: int, R: int) -> torch.Tensor:
        if ws is None:
            return self.workspace(B, R)
        if not (ws.is_cuda and ws.device == self.device and ws.dtype == torch.uint8 and ws.is_contiguous()):
            raise ValueError("ws: need a contiguous uint8 CUDA tensor on the engine's device")
        if ws.numel() < self.workspace_bytes(B, R):
            raise ValueError(f"ws: {ws.numel()} bytes < {self.workspace_bytes(B, R)} needed for B={B}, R={R}")
        return ws

    def _check_batch(self, B: int, R: int, sample_real, *per_sample):
        """lengths of the per-sample inputs (a short one would be an out-of-bounds device read)"""
        for t, nm in per_sample:
            if t.numel() != B:
                raise ValueError(f"{nm}: {t.numel()} elements, need B = {B}")
        if sample_real is not None:
            self._check(sample_real, "sample_real", torch.int32)
            if sample_real.numel() != B:
                raise ValueError(f"sample_real: {sample_real.numel()} elements, need B = {B}")

    # ---------------------------------------------------------------------------------------
    def pvt_eval(self, p: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        """values, derivatives: each (n_props, n)."""
        self._check(p, "p")
        n = p.numel()
        val = torch.empty((self.n_props, n), dtype=torch.float32, device=self.device)
        der = torch.empty_like(val)
        L.check(self.lib, self.lib.srm_pvt_eval(self._h, n, _ptr(p), _ptr(val), _ptr(der), self._stream()), "srm_pvt_eval")
        self.launches += 1
        return val, der

    def denormalize_log(self, x_norm: torch.Tensor, kmin: float, kmax: float, lo: float = -1.0, hi: float = 1.0):
        self._check(x_norm, "x_norm")
        out = torch.empty_like(x_norm)
        L.check(self.lib, self.lib.srm_denormalize_log(self.device.index, x_norm.numel(), _ptr(x_norm), kmin, kmax, lo, hi, _ptr(out),
                                                       self._stream()), "srm_denormalize_log")
        self.launches += 1
        return out

    def features_forward(self, x: torch.Tensor, dn: Optional[torch.Tensor] = None, kx_range=None, lo: float = -1.0,
                         hi: float = 1.0, t_channel: int = 3, k_channel: int = 4):
        """one pass over the (B, ..., C) feature tensor: x1 = x with t_norm += dn[b] (if dn is given) and the
        de-normalised permeability channel (if kx_range = (kmin, kmax) is given).  Returns (x1, kx)."""
        self._check(x, "x")
        B, Cc = x.shape[0], x.shape[-1]
        cells = x.numel() // (B * Cc)
        x1 = torch.empty_like(x) if dn is not None else None
        kx = torch.empty(x.shape[:-1], dtype=torch.float32, device=self.device) if kx_range is not None else None
        if dn is not None:
            self._check(dn, "dn")
        kmin, kmax = kx_range if kx_range is not None else (1.0, 2.0)
        L.check(self.lib, self.lib.srm_features_forward(self.device.index, _ptr(x), _ptr(dn), B, cells, Cc, t_channel, k_channel,
                                                        kmin, kmax, lo, hi, _ptr(x1), _ptr(kx), self._stream()),
                "srm_features_forward")
        self.launches += 1
        return x1, kx

    def features_backward(self, gx1: torch.Tensor, t_channel: int = 3) -> torch.Tensor:
        """cotangent of dn: per-sample sum of the time channel of gx1"""
        self._check(gx1, "gx1")
        B, Cc = gx1.shape[0], gx1.shape[-1]
        gdn = torch.empty(B, dtype=torch.float32, device=self.device)
        L.check(self.lib, self.lib.srm_features_backward(self.device.index, _ptr(gx1), B, gx1.numel() // (B * Cc), Cc, t_channel,
                                                         _ptr(gdn), self._stream()), "srm_features_backward")
        self.launches += 1
        return gdn

    # -- glue either side of the physics kernels: HardLayer + per-sample mean of the dt field, both levels ------------
    def glue_forward(self, y0, y1, tn0, tn1, expo=None, dtf1=None, dtf2=None, init_value: float = 1.0,
                     t_lo: float = -1.0, t_hi: float = 1.0):
        """p_l = init_value - ((tn_l - t_lo)/(t_hi - t_lo)) ** expo * y_l  (Hard_Layer_Subclassed.py:219-242);
        dt_l = per-sample mean of dtf_l (physics_loss.py:102,122).  Returns (p0, p1, dt1, dt2); dt_l is None without dtf_l."""
        B = y0.shape[0]
        for t, nm in ((y0, "y0"), (y1, "y1"), (tn0, "tn0"), (tn1, "tn1")):
            self._check(t, nm)
        for t, nm in ((expo, "expo"), (dtf1, "dtf1"), (dtf2, "dtf2")):
            if t is not None:
                self._check(t, nm)
        p0, p1 = torch.empty_like(y0), torch.empty_like(y1)
        dt1 = torch.empty(B, dtype=torch.float32, device=self.device) if dtf1 is not None else None
        dt2 = torch.empty(B, dtype=torch.float32, device=self.device) if dtf2 is not None else None
        nbytes = self.lib.srm_glue_workspace_bytes(B)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        L.check(self.lib, self.lib.srm_glue_forward(self._h, B, init_value, t_lo, t_hi, _ptr(expo), _ptr(tn0), _ptr(tn1),
                                                    _ptr(y0), _ptr(y1), _ptr(dtf1), _ptr(dtf2), _ptr(p0), _ptr(p1),
                                                    _ptr(dt1), _ptr(dt2), _ptr(ws), nbytes, self._stream()),
                "srm_glue_forward")
        self.launches += 1 + (1 if (dtf1 is not None or dtf2 is not None) else 0)
        return p0, p1, dt1, dt2

    def glue_backward(self, y0, y1, tn0, tn1, gp0, gp1, expo=None, gdt1=None, gdt2=None, init_value: float = 1.0,
                      t_lo: float = -1.0, t_hi: float = 1.0, want_gexpo: bool = True, want_gtn=(False, False)):
        """cotangents of glue_forward: (gy0, gy1, gexpo, gdtf1, gdtf2), plus (gtn0, gtn1) -- the cotangents of the
        layer's time inputs, Hard_Layer_Subclassed.py:214-228 -- when want_gtn asks for either"""
        B = y0.shape[0]
        for t, nm in ((y0, "y0"), (y1, "y1"), (tn0, "tn0"), (tn1, "tn1"), (gp0, "gp0"), (gp1, "gp1")):
            self._check(t, nm)
        gy0, gy1 = torch.empty_like(y0), torch.empty_like(y1)
        gexpo = torch.empty(y0.shape[1:], dtype=torch.float32, device=self.device) if want_gexpo else None
        gdtf1 = torch.empty_like(y0) if gdt1 is not None else None
        gdtf2 = torch.empty_like(y0) if gdt2 is not None else None
        gtn0 = torch.empty(B, dtype=torch.float32, device=self.device) if want_gtn[0] else None
        gtn1 = torch.empty(B, dtype=torch.float32, device=self.device) if want_gtn[1] else None
        L.check(self.lib, self.lib.srm_glue_backward(self._h, B, init_value, t_lo, t_hi, _ptr(expo), _ptr(tn0), _ptr(tn1),
                                                     _ptr(y0), _ptr(y1), _ptr(gp0), _ptr(gp1), _ptr(gdt1), _ptr(gdt2),
                                                     _ptr(gy0), _ptr(gy1), _ptr(gexpo), _ptr(gdtf1), _ptr(gdtf2),
                                                     _ptr(gtn0), _ptr(gtn1), self._stream()), "srm_glue_backward")
        self.launches += 1
        if want_gtn[0] or want_gtn[1]:
            return gy0, gy1, gexpo, gdtf1, gdtf2, gtn0, gtn1
        return gy0, gy1, gexpo, gdtf1, gdtf2

    def wells(self, kx, sample_real, p, t_days, dense: bool = False):
        B = p.shape[0]
        R = kx.shape[0]
        self._check(kx, "kx"); self._check(p, "p"); self._check(t_days, "t_days")
        if sample_real is not None:
            self._check(sample_real, "sample_real", torch.int32)
        nw = max(self.n_wells, 1)
        qw = torch.zeros((B, nw), dtype=torch.float32, device=self.device)
        pwfw = torch.zeros_like(qw)
        dqdp = torch.zeros_like(qw)
        qd = torch.empty_like(p) if dense else None
        pd = torch.empty_like(p) if dense else None
        L.check(self.lib, self.lib.srm_wells(self._h, B, R, _ptr(kx), _ptr(sample_real), _ptr(p), _ptr(t_days),
                                             _ptr(qw), _ptr(pwfw), _ptr(dqdp), _ptr(qd), _ptr(pd), self._stream()),
                "srm_wells")
        self.launches += 4 + (2 if dense else 0)
        return dict(qw=qw, pwfw=pwfw, dqdp=dqdp, q=qd, pwf=pd)

    def forward(self, kx, sample_real, p0, p1, dt1, dt2, t1, want_dom: bool = False, want_wells: bool = False,
                save_for_backward: bool = True, ws=None):
        B = p0.shape[0]
        R = kx.shape[0]
        for t, nm in ((kx, "kx"), (p0, "p0"), (p1, "p1"), (dt1, "dt1"), (dt2, "dt2"), (t1, "t1")):
            self._check(t, nm)
        self._check_batch(B, R, sample_real, (dt1, "dt1"), (dt2, "dt2"), (t1, "t1"))
        if p0.numel() != B * self.spec.n_cells or p1.shape != p0.shape or kx.numel() != R * self.spec.n_cells:
            raise ValueError("field shapes do not match the handle's grid")
        ws = self._ws_for(ws, B, R)
        terms = torch.empty((2, L.SRM_N_TERMS), dtype=torch.float32, device=self.device)
        dom = torch.empty_like(p0) if want_dom else None
        qw = torch.empty((B, max(self.n_wells, 1)), dtype=torch.float32, device=self.device) if want_wells else None
        pwfw = torch.empty_like(qw) if want_wells else None
        flags = L.SRM_FLAG_SAVE_FOR_BACKWARD if save_for_backward else 0
        L.check(self.lib, self.lib.srm_forward(self._h, B, R, _ptr(kx), _ptr(sample_real), _ptr(p0), _ptr(p1),
                                               _ptr(dt1), _ptr(dt2), _ptr(t1), _ptr(terms), _ptr(dom), _ptr(qw),
                                               _ptr(pwfw), _ptr(ws), ws.numel(), flags, self._stream()), "srm_forward")
        # kernels per forward (memsets are not kernels): fused reference = faces, wells, qsum, residual, finalize;
        # staged reference = stage, wells, residual, finalize; closed form = grouping, wells, qsum, residual, finalize
        nk = 4 if (self.numerics == "reference" and not self.pvt_lut) else 5
        if self.n_wells == 0:
            nk -= 1 if nk == 4 else 2
        self.launches += nk + (2 if (want_wells and self.n_wells) else 0)
        return dict(terms=terms, dom=dom, qw=qw, pwfw=pwfw)

    def backward(self, kx, sample_real, p0, p1, dt1, dt2, t1, dterms, out=None, ws=None):
        """out: optional preallocated (gp0, gp1, gdt1, gdt2) to write into"""
        B = p0.shape[0]
        R = kx.shape[0]
        self._check(dterms, "dterms")
        if dterms.numel() != L.SRM_N_TERMS:
            raise ValueError(f"dterms: {dterms.numel()} elements, need {L.SRM_N_TERMS}")
        for t, nm in ((kx, "kx"), (p0, "p0"), (p1, "p1"), (dt1, "dt1"), (dt2, "dt2"), (t1, "t1")):
            self._check(t, nm)
        self._check_batch(B, R, sample_real, (dt1, "dt1"), (dt2, "dt2"), (t1, "t1"))
        ws = self._ws_for(ws, B, R)
        if out is not None:
            gp0, gp1, gdt1, gdt2 = out
            for t, nm in ((gp0, "gp0"), (gp1, "gp1"), (gdt1, "gdt1"), (gdt2, "gdt2")):
                self._check(t, nm)
        else:
            gp0 = torch.empty_like(p0)
            gp1 = torch.empty_like(p1)
            gdt1 = torch.empty_like(dt1)
            gdt2 = torch.empty_like(dt2)
        L.check(self.lib, self.lib.srm_backward(self._h, B, R, _ptr(kx), _ptr(sample_real), _ptr(p0), _ptr(p1),
                                                _ptr(dt1), _ptr(dt2), _ptr(t1), _ptr(dterms), _ptr(gp0), _ptr(gp1),
                                                _ptr(gdt1), _ptr(gdt2), _ptr(ws), ws.numel(), 0, self._stream()),
                "srm_backward")
        self.launches += 3 if self.n_wells else 2       # adjoint, inner-boundary scatter (wells only), finalize
        return gp0, gp1, gdt1, gdt2

    # ---------------------------------------------------------------------------------------
    # gas condensate (two-phase)                                   physics_loss.py:230-712
    def relperm(self, sg: torch.Tensor):
        """RelativePermeability.compute_krog_krgo (relative_permeability.py:49-75): krog, krgo, d/dSg of both."""
        self._check(sg, "sg")
        outs = [torch.empty_like(sg) for _ in range(4)]
        L.check(self.lib, self.lib.srm_relperm(self._h, sg.numel(), _ptr(sg), *[_ptr(o) for o in outs], self._stream()),
                "srm_relperm")
        self.launches += 1
        return tuple(outs)

    def wells_gc(self, kx, sample_real, p, sg, t_days):
        """WellRatesPressure.compute_rates_and_bhp, GC (well_rate_bhp_Subclassed.py:727-837): dense (qgg, qgo, qoo, qog)
        and pwf fields, zero off-well, duplicates summed (scatter_nd).  The well kernel runs inside srm_forward_gc; the
        rates do not depend on the time-level-n fields or the time steps, so neutral values stand in for them."""
        B, nw = p.shape[0], self.n_wells
        one = torch.ones(B, dtype=torch.float32, device=self.device)
        so = (1.0 - float(self.spec.end_points["Swmin"])) - sg
        fw = self.forward_gc(kx, sample_real, p, p, sg, sg, so, so, one, one, t_days, want_wells=True, save_for_backward=False)
        N = self.spec.n_cells
        dense = torch.zeros((5, B, N), dtype=torch.float32, device=self.device)
        if nw:
            cells = torch.tensor([(w.k * self.spec.H + w.j) * self.spec.W + w.i for w in self.spec.wells], dtype=torch.long,
                                 device=self.device)
            vals = torch.cat([fw["q4w"], fw["pwfw"].unsqueeze(0)], dim=0)           # (5, B, nw), caller's well order
            dense.index_add_(2, cells, vals)
        dense = dense.reshape((5,) + tuple(p.shape))
        return tuple(dense[:4]), dense[4]

    def forward_gc(self, kx, sample_real, p0, p1, sg0, sg1, so0, so1, dt1, dt2, t1, want_dom: bool = False,
                   want_wells: bool = False, save_for_backward: bool = True, ws=None):
        B, R = p0.shape[0], kx.shape[0]
        for t, nm in ((kx, "kx"), (p0, "p0"), (p1, "p1"), (sg0, "sg0"), (sg1, "sg1"), (so0, "so0"), (so1, "so1"),
                      (dt1, "dt1"), (dt2, "dt2"), (t1, "t1")):
            self._check(t, nm)
        self._check_batch(B, R, sample_real, (dt1, "dt1"), (dt2, "dt2"), (t1, "t1"))
        for t in (p1, sg0, sg1, so0, so1):
            if t.shape != p0.shape:
                raise ValueError("field shapes differ")
        if p0.numel() != B * self.spec.n_cells or kx.numel() != R * self.spec.n_cells:
            raise ValueError("field shapes do not match the handle's grid")
        ws = self._ws_for(ws, B, R)
        terms = torch.empty((2, L.SRM_N_TERMS), dtype=torch.float32, device=self.device)
        dom = torch.empty_like(p0) if want_dom else None
        nw = max(self.n_wells, 1)
        q4 = torch.zeros((4, B, nw), dtype=torch.float32, device=self.device) if want_wells else None
        pwfw = torch.zeros((B, nw), dtype=torch.float32, device=self.device) if want_wells else None
        flags = L.SRM_FLAG_SAVE_FOR_BACKWARD if save_for_backward else 0
        L.check(self.lib, self.lib.srm_forward_gc(
            self._h, B, R, _ptr(kx), _ptr(sample_real), _ptr(p0), _ptr(p1), _ptr(sg0), _ptr(sg1), _ptr(so0), _ptr(so1),
            _ptr(dt1), _ptr(dt2), _ptr(t1), _ptr(terms), _ptr(dom), _ptr(q4), _ptr(pwfw), _ptr(ws), ws.numel(), flags,
            self._stream()), "srm_forward_gc")
        self.launches += (4 if self.n_wells else 3) + (5 if (want_wells and self.n_wells) else 0)   # stage, wells, residual, finalize
        return dict(terms=terms, dom=dom, q4w=q4, pwfw=pwfw)

    def backward_gc(self, kx, sample_real, p0, p1, sg0, sg1, so0, so1, dt1, dt2, t1, dterms, out=None, ws=None):
        B, R = p0.shape[0], kx.shape[0]
        self._check(dterms, "dterms")
        if dterms.numel() != L.SRM_N_TERMS:
            raise ValueError(f"dterms: {dterms.numel()} elements, need {L.SRM_N_TERMS}")
        self._check_batch(B, R, sample_real, (dt1, "dt1"), (dt2, "dt2"), (t1, "t1"))
        ws = self._ws_for(ws, B, R)
        if out is not None:
            g = list(out)
            for t in g:
                self._check(t, "out")
        else:
            g = [torch.empty_like(p0) for _ in range(6)] + [torch.empty_like(dt1), torch.empty_like(dt2)]
        L.check(self.lib, self.lib.srm_backward_gc(
            self._h, B, R, _ptr(kx), _ptr(sample_real), _ptr(p0), _ptr(p1), _ptr(sg0), _ptr(sg1), _ptr(so0), _ptr(so1),
            _ptr(dt1), _ptr(dt2), _ptr(t1), _ptr(dterms), *[_ptr(t) for t in g], _ptr(ws), ws.numel(), 0, self._stream()),
            "srm_backward_gc")
        self.launches += 3 if self.n_wells else 2
        return tuple(g)        # gp0, gp1, gsg0, gsg1, gso0, gso1, gdt1, gdt2


class GraphedStep:
    """Forward + adjoint of one fixed-shape batch, captured once in a CUDA graph and replayed.

    At the reference's own sizes (39 x 39 x 1 cells, 32 samples: srm_training_examples/training_case_dry_gas_i.py:331)
    a step is a dozen launches of a few microseconds each and the host's launch path is the cost; the replay issues
    them as one graph launch.  The inputs are static device buffers (`inputs`; refill them with `load`), the outputs
    (`terms` [2][8], `grads`) are written in place by every replay.
    """

    def __init__(self, eng: "SrmPhysics", batch: dict, dterms: torch.Tensor, warmup: int = 2):
        self.eng = eng
        gc = eng.fluid == "GC"
        fwd, bwd = (eng.forward_gc, eng.backward_gc) if gc else (eng.forward, eng.backward)
        self.inputs = {k: v.clone() for k, v in batch.items()}
        self.dterms = dterms.clone()
        # the graph bakes the workspace address in: it owns one, so later calls on the engine (other batch shapes
        # re-allocate the engine's cached workspace) cannot pull it from under the replays
        self.ws = eng.new_workspace(self.inputs["p0"].shape[0], self.inputs["kx"].shape[0])
        cur = torch.cuda.current_stream(eng.device)
        side = torch.cuda.Stream(eng.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):           # workspace, function attributes and first-call work happen here, uncaptured
            for _ in range(max(1, warmup)):
                fwd(ws=self.ws, **self.inputs)
                bwd(dterms=self.dterms, ws=self.ws, **self.inputs)
        cur.wait_stream(side)
        torch.cuda.synchronize(eng.device)
        self.graph = torch.cuda.CUDAGraph()
        l0 = eng.launches
        with torch.cuda.graph(self.graph):
            fw = fwd(ws=self.ws, **self.inputs)
            self.grads = bwd(dterms=self.dterms, ws=self.ws, **self.inputs)
        self.terms = fw["terms"]
        self.kernels = eng.launches - l0        # kernels inside one replay

    def load(self, **fields):
        for k, v in fields.items():
            if k == "dterms":
                self.dterms.copy_(v)
            else:
                self.inputs[k].copy_(v)

    def replay(self):
        self.graph.replay()
        self.eng.launches += self.kernels
        return self.terms, self.grads


class HostPipeline:
    """The end-to-end call with HOST buffers: loss terms and gradients for a batch that lives in pinned host memory.

    The reference converts its numpy batch to device tensors every step (training.py:595-600) and pulls the loss
    terms back with .numpy() (training.py:608-610).  Here the batch is cut into chunks of whole realisations
    (samples are independent units and the loss terms are additive); chunk i+1 travels host->device on a copy
    stream while chunk i runs forward + adjoint and the gradients of chunk i-1 travel device->host on a third
    stream, so the PCIe link is used in both directions at once and the kernels hide behind the copies.
    Requires the realisation-major sample order BatchGenerator produces (training.py:187-204): sample_real ascending.
    """

    DG_FIELDS = ("p0", "p1")
    GC_FIELDS = ("p0", "p1", "sg0", "sg1", "so0", "so1")

    def __init__(self, eng: "SrmPhysics", host: dict, dterms, n_chunks: int = 8, grads_to_host: bool = True):
        """grads_to_host=False keeps the cotangents on the device (where the networks that consume them live, as in
        the reference's training loop): only the loss terms travel back, step() returns device gradient tensors."""
        self.eng = eng
        self.grads_to_host = bool(grads_to_host)
        self.gc = eng.fluid == "GC"
        self.fields = self.GC_FIELDS if self.gc else self.DG_FIELDS
        self.scalars = ("dt1", "dt2", "t1")
        dev = eng.device
        for k, v in host.items():
            if not (isinstance(v, torch.Tensor) and v.device.type == "cpu" and v.is_pinned() and v.is_contiguous()):
                raise ValueError(f"{k}: need a contiguous pinned host tensor")
        sr = host["sample_real"]
        if sr.numel() > 1 and bool((sr[1:] < sr[:-1]).any()):
            raise ValueError("HostPipeline needs realisation-major sample order (sample_real ascending)")
        self.host = host
        self.B = host["p0"].shape[0]
        R = host["kx"].shape[0]
        n_chunks = max(1, min(int(n_chunks), R))
        # chunk = a range of whole realisations and the (contiguous) samples that belong to it
        starts = torch.searchsorted(sr.to(torch.int64), torch.arange(R + 1, dtype=torch.int64)).tolist()
        self.chunks = []
        for c in range(n_chunks):
            r0, r1 = (c * R) // n_chunks, ((c + 1) * R) // n_chunks
            if r1 > r0 and starts[r1] > starts[r0]:
                self.chunks.append((r0, r1, starts[r0], starts[r1]))
        mb = max(b1 - b0 for _, _, b0, b1 in self.chunks)
        mr = max(r1 - r0 for r0, r1, _, _ in self.chunks)
        cell = host["p0"].shape[1:]
        f32 = dict(dtype=torch.float32, device=dev)
        self.din = [dict(kx=torch.empty((mr,) + tuple(cell), **f32), sample_real=torch.empty(mb, dtype=torch.int32, device=dev),
                         **{k: torch.empty((mb,) + tuple(cell), **f32) for k in self.fields},
                         **{k: torch.empty(mb, **f32) for k in self.scalars}) for _ in range(2)]
        self.gnames = tuple("g" + k for k in self.fields) + ("gdt1", "gdt2")
        if self.grads_to_host:
            self.dout = [[torch.empty((mb,) + tuple(cell), **f32) for _ in self.fields] + [torch.empty(mb, **f32), torch.empty(mb, **f32)]
                         for _ in range(2)]
            self.hout = {n: (torch.empty(self.B, dtype=torch.float32) if n.startswith("gdt") else torch.empty_like(host["p0"])).pin_memory()
                         for n in self.gnames}
        else:
            self.dgrads = {n: (torch.empty(self.B, **f32) if n.startswith("gdt") else torch.empty((self.B,) + tuple(cell), **f32))
                           for n in self.gnames}
            self.hout = {}
        self.ws = torch.empty(max(eng.workspace_bytes(b1 - b0, r1 - r0) for r0, r1, b0, b1 in self.chunks), dtype=torch.uint8, device=dev)
        self.hterms = torch.empty((2, L.SRM_N_TERMS), dtype=torch.float32).pin_memory()
        self.terms = torch.zeros((2, L.SRM_N_TERMS), **f32)
        self.dterms = dterms
        self.s_in, self.s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())
        self.d2h_bytes = sum(v.numel() * v.element_size() for v in self.hout.values()) + self.hterms.numel() * 4

    def step(self, reduce_terms=None):
        """one forward + adjoint over the whole host batch; returns (host terms [2][8], dict of host gradients)"""
        eng, host = self.eng, self.host
        comp = torch.cuda.current_stream(eng.device)
        self.s_in.wait_stream(comp)
        self.s_out.wait_stream(comp)
        self.terms.zero_()
        ev_in = [None, None]; ev_free_in = [None, None]; ev_comp = [None, None]; ev_free_out = [None, None]
        for c, (r0, r1, b0, b1) in enumerate(self.chunks):
            slot = c & 1
            nb, nr = b1 - b0, r1 - r0
            din = self.din[slot]
            with torch.cuda.stream(self.s_in):
                if ev_free_in[slot] is not None:
                    self.s_in.wait_event(ev_free_in[slot])
                din["kx"][:nr].copy_(host["kx"][r0:r1], non_blocking=True)
                din["sample_real"][:nb].copy_(host["sample_real"][b0:b1], non_blocking=True)
                for k in self.fields + self.scalars:
                    din[k][:nb].copy_(host[k][b0:b1], non_blocking=True)
                ev_in[slot] = self.s_in.record_event()
            comp.wait_event(ev_in[slot])
            if ev_free_out[slot] is not None:
                comp.wait_event(ev_free_out[slot])
            a = {k: din[k][:nb] for k in self.fields + self.scalars}
            a["kx"] = din["kx"][:nr]
            a["sample_real"] = din["sample_real"][:nb]
            if r0:
                a["sample_real"].sub_(r0)
            if self.grads_to_host:
                out = [t[:nb] for t in self.dout[slot]]
            else:
                out = [self.dgrads[n][b0:b1] for n in self.gnames]
            if self.gc:
                fw = eng.forward_gc(ws=self.ws, **a)
                self.terms.add_(fw["terms"])
                eng.backward_gc(dterms=self.dterms, out=out, ws=self.ws, **a)
            else:
                fw = eng.forward(ws=self.ws, **a)
                self.terms.add_(fw["terms"])
                eng.backward(dterms=self.dterms, out=out, ws=self.ws, **a)
            ev_comp[slot] = comp.record_event()
            ev_free_in[slot] = ev_comp[slot]
            if self.grads_to_host:
                with torch.cuda.stream(self.s_out):
                    self.s_out.wait_event(ev_comp[slot])
                    for n, t in zip(self.gnames, out):
                        self.hout[n][b0:b1].copy_(t, non_blocking=True)
                    ev_free_out[slot] = self.s_out.record_event()
        if reduce_terms is not None:
            reduce_terms(self.terms)
        self.hterms.copy_(self.terms, non_blocking=True)
        comp.wait_stream(self.s_out)
        torch.cuda.synchronize(eng.device)
        return self.hterms, (self.hout if self.grads_to_host else self.dgrads)
