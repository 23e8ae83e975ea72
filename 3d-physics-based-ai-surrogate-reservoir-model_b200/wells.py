"""Mirrors of the reference's well classes on top of the CUDA kernels.

``WellDataProcessor`` restates the integer bookkeeping of welldata_processor.py:18-389 (index rows
[k,j,i], sign rule, scatter_nd that sums duplicates, shut-in mask); ``WellRatesPressure`` keeps the
call signature of well_rate_bhp_Subclassed.py:727 and evaluates rates/BHP sparsely on the GPU
(srm_wells) instead of densely over every cell.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from .config import PhysicsSpec, WellSpec, wells_from_connections


class WellDataProcessor:
    def __init__(self, well_list: Sequence[dict], mode_order=("k", "j", "i")):
        self.well_list = list(well_list)
        self.perm = [{"i": 0, "j": 1, "k": 2}[d] for d in mode_order]           # welldata_processor.py:30
        self._specs: List[WellSpec] = wells_from_connections(self.well_list)

    def get_well_data(self):
        """welldata_processor.py:74-107"""
        coords = np.array([[w["i"], w["j"], w["k"]] for w in self.well_list], dtype=np.int32).reshape(-1, 3)
        return {
            "connection_index": torch.from_numpy(coords[:, self.perm].copy()),
            "control_mode_value": torch.tensor([w.q_target for w in self._specs], dtype=torch.float32),
            "names": [w.name for w in self._specs],
            "wellbore_radius": torch.tensor([w.rw for w in self._specs], dtype=torch.float32),
            "completion_ratio": torch.tensor([w.hc for w in self._specs], dtype=torch.float32),
            "minimum_bhp": torch.tensor([w.pwf_min for w in self._specs], dtype=torch.float32),
            "shutin_days": torch.tensor([[[w.shut_start, w.shut_stop]] for w in self._specs], dtype=torch.float32),
        }

    @staticmethod
    def scatter_y(target_shape, index_list, y, start_dim=1):
        """tf.scatter_nd into zeros; duplicates SUM (welldata_processor.py:170-224)."""
        idx = torch.as_tensor(np.asarray(index_list), dtype=torch.long).reshape(-1, 3)
        out = torch.zeros(tuple(target_shape), dtype=torch.float32)
        vals = torch.as_tensor(np.asarray(y, dtype=np.float32)).reshape(-1)
        if vals.numel() == 1:
            vals = vals.expand(idx.shape[0])
        lead = (0,) * start_dim
        tail = (0,) * (out.dim() - start_dim - 3)
        for n in range(idx.shape[0]):
            out[lead + tuple(idx[n].tolist()) + tail] += vals[n]
        return out

    @staticmethod
    def conn_shutins_idx(time_tensor, index_list, range_conditions, time_axis=0):
        """1 at connection cells whose time is NOT inside any inclusive [start, stop]; 0 elsewhere,
        incl. every non-well cell; duplicates: last writer (welldata_processor.py:228-389).
        time_tensor: (T, C, H, W[, ...]) with the time axis first."""
        t = torch.as_tensor(time_tensor)
        out = torch.zeros_like(t, dtype=torch.int32)
        idx = np.asarray(index_list).reshape(-1, 3)
        rc = torch.as_tensor(np.asarray(range_conditions, dtype=np.float32)).reshape(idx.shape[0], -1, 2)
        for n, (c, h, w) in enumerate(idx):
            v = t[:, c, h, w]
            v0 = v.reshape(v.shape[0], -1)[:, 0]
            inside = ((v0[:, None] >= rc[n, :, 0][None]) & (v0[:, None] <= rc[n, :, 1][None])).any(dim=1)
            upd = (~inside).to(torch.int32)
            out[:, c, h, w] = upd.reshape((-1,) + (1,) * (v.dim() - 1)).expand_as(v)
        return out


class WellRatesPressure:
    """compute_rates_and_bhp(x_n1, p_n1, Sg_n1, relperm_model, model_PVT, q_target=None, shutin_days=None)
    -> DG: (qg, pwf);  GC: ((qgg, qgo, qoo, qog), pwf) -- every field (B, D, H, W, 1), zero off-well
    (well_rate_bhp_Subclassed.py:727-837)."""

    def __init__(self, engine, fluid_type="DG", use_blocking_factor=None, n_intervals=None, use_non_iterative=True,
                 general_config=None, kx_stats=(0.26, 24.0), t_range=(0.0, 365.0), norm_limits=(-1.0, 1.0)):
        if fluid_type.upper() != engine.fluid:
            raise ValueError(f"fluid_type {fluid_type!r} does not match the engine ({engine.fluid})")
        spec: PhysicsSpec = engine.spec
        if bool(use_non_iterative) != bool(spec.use_non_iterative):      # the control method is fixed when the handle is made
            raise ValueError("use_non_iterative must match the engine's PhysicsSpec (well_rate_bhp_Subclassed.py:44, 813-822)")
        if use_blocking_factor is not None and bool(use_blocking_factor) != bool(spec.use_blocking_factor):
            raise ValueError("use_blocking_factor must match the engine's PhysicsSpec")
        self.engine = engine
        self.fluid_type = engine.fluid
        self.use_blocking_factor = spec.use_blocking_factor
        self.n_intervals = spec.n_intervals
        self.use_non_iterative, self.max_iters, self.tol = spec.use_non_iterative, spec.max_iters, spec.tol
        self.k_min, self.k_max = kx_stats
        self.t_min, self.t_max = t_range
        self.lo, self.hi = norm_limits
        self.trainable_variables = []

    def compute_rates_and_bhp(self, x_n1, p_n1, Sg_n1=None, relperm_model=None, model_PVT=None, q_target=None,
                              shutin_days=None):
        if q_target is not None or shutin_days is not None:
            raise NotImplementedError("dynamic q_target / shutin_days overrides are fixed at engine creation")
        eng = self.engine
        x = x_n1.to(eng.device, torch.float32)
        B = x.shape[0]
        kx = eng.denormalize_log(x[..., 4].contiguous(), self.k_min, self.k_max, self.lo, self.hi)
        tn = x[:, 0, 0, 0, 3]
        t = ((self.t_max - self.t_min) * ((tn - self.lo) / (self.hi - self.lo)) + self.t_min).contiguous()
        p = p_n1.to(eng.device, torch.float32).reshape(x.shape[:-1]).contiguous()
        sr = torch.arange(B, dtype=torch.int32, device=eng.device)
        if self.fluid_type == "GC":
            if Sg_n1 is None:
                raise ValueError("the gas-condensate well model needs Sg_n1")
            sg = Sg_n1.to(eng.device, torch.float32).reshape(x.shape[:-1]).contiguous()
            q4, pwf = eng.wells_gc(kx, sr, p, sg, t)
            return tuple(q.unsqueeze(-1) for q in q4), pwf.unsqueeze(-1)
        out = eng.wells(kx, sr, p, t, dense=True)
        return out["q"].unsqueeze(-1), out["pwf"].unsqueeze(-1)
