"""PhysicsLoss: host-side mirror of the class the reference's example instantiates
(srm_training_examples/training_case_dry_gas_i.py:357-364) and its training loop drives
(training.py:552-560,603-652).  The reference does not ship the class (physics_loss_Subclassed.py is
missing); the call contract is reconstructed from the caller and the arithmetic is the
legacy physics_loss.py:79-208,742-870 -- executed by the CUDA kernels behind the C ABI.

This is the torch-harness twin of the TensorFlow binding in INTEGRATION.md: torch modules stand in
for the Keras models and ``torch.autograd.Function`` for ``tf.custom_gradient``.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch

from . import _lib as L
from . import dist as sdist
from .config import DEFAULT_GENERAL, LOSS_KEYS

# loss key (default_configurations.py:63-83 order) -> slot of the library's terms vector
_SLOT = {"dom": L.TERM_NAMES.index("dom"), "ibc": L.TERM_NAMES.index("ibc"), "obc": L.TERM_NAMES.index("obc"),
         "ic": L.TERM_NAMES.index("ic"), "td": L.TERM_NAMES.index("td"), "mbc": L.TERM_NAMES.index("mbc"),
         "cmbc": L.TERM_NAMES.index("cmbc"), "tde": L.TERM_NAMES.index("tde")}


class _SrmPhysicsFn(torch.autograd.Function):
    """terms = srm_forward(...);  d terms / d (p0, p1, dt1, dt2) = srm_backward(..., dterms)"""

    @staticmethod
    def forward(ctx, engine, kx, sample_real, t1, p0, p1, dt1, dt2):
        p0c, p1c, d1c, d2c = (t.detach().contiguous() for t in (p0, p1, dt1, dt2))
        fw = engine.forward(kx, sample_real, p0c, p1c, d1c, d2c, t1, save_for_backward=True)
        terms = fw["terms"]                     # this rank's shard: the caller reduces AFTER it has queued the backward
        ctx.engine = engine
        ctx.save_for_backward(kx, sample_real, t1, p0c, p1c, d1c, d2c)
        return terms

    @staticmethod
    def backward(ctx, gterms):
        kx, sample_real, t1, p0, p1, dt1, dt2 = ctx.saved_tensors
        dterms = gterms[0].contiguous().to(torch.float32)
        gp0, gp1, gdt1, gdt2 = ctx.engine.backward(kx, sample_real, p0, p1, dt1, dt2, t1, dterms)
        return None, None, None, None, gp0, gp1, gdt1, gdt2


class _SrmPhysicsGcFn(torch.autograd.Function):
    """gas condensate: terms = srm_forward_gc(...); cotangents of p, Sg, So at both levels and of dt1, dt2"""

    @staticmethod
    def forward(ctx, engine, kx, sample_real, t1, p0, p1, sg0, sg1, so0, so1, dt1, dt2):
        c = [t.detach().contiguous() for t in (p0, p1, sg0, sg1, so0, so1, dt1, dt2)]
        fw = engine.forward_gc(kx, sample_real, *c, t1, save_for_backward=True)
        terms = fw["terms"]
        ctx.engine = engine
        ctx.save_for_backward(kx, sample_real, t1, *c)
        return terms

    @staticmethod
    def backward(ctx, gterms):
        kx, sample_real, t1, *c = ctx.saved_tensors
        dterms = gterms[0].contiguous().to(torch.float32)
        g = ctx.engine.backward_gc(kx, sample_real, *c, t1, dterms)
        return (None, None, None, None) + tuple(g)


class _OptimizerSlot:
    def __init__(self, optimizer):
        self.optimizer = optimizer


class _ShiftTimeFn(torch.autograd.Function):
    """x1 = x with channel 3 += dn[b]; d/dx = identity, d/d dn[b] = sum of the time channel of the cotangent"""

    @staticmethod
    def forward(ctx, eng, x, dn):
        ctx.eng = eng
        return eng.features_forward(x, dn.detach())[0]

    @staticmethod
    def backward(ctx, gx1):
        gx1 = gx1.contiguous()
        gdn = ctx.eng.features_backward(gx1) if ctx.needs_input_grad[2] else None
        return None, (gx1 if ctx.needs_input_grad[1] else None), gdn


class PhysicsLoss:
    """PhysicsLoss(main_model, pvt_model, time_step_model, well_rate_bhp_model, saturation_model=None,
                   optimizer_model_names_map=...)

    main_model / time_step_model: callables mapping the feature tensor x (B, D, H, W, 5), channels
    [z, y, x, t, k] normalised to [-1, 1] (data_processing/data_processing_utils.py:219-222), to a
    (B, D, H, W, 1) field (complete_trainable_module.py:142).  pvt_model and well_rate_bhp_model are
    the mirrors in pvt.py / wells.py; they carry the engine (device tables) the loss evaluates with.
    """

    def __init__(self, main_model, pvt_model, time_step_model, well_rate_bhp_model, saturation_model=None,
                 optimizer_model_names_map: Optional[Dict[str, str]] = None, *, optimizers: Optional[Dict[str, object]] = None,
                 general_config: Optional[dict] = None, kx_stats=(0.26, 24.0), weights: Optional[Dict[str, float]] = None):
        self.main_model = main_model
        self.pvt_model = pvt_model
        self.time_step_model = time_step_model
        self.well_rate_bhp_model = well_rate_bhp_model
        self.saturation_model = saturation_model
        self.engine = getattr(pvt_model, "engine", None) or getattr(well_rate_bhp_model, "engine", None)
        if self.engine is None:
            raise ValueError("pvt_model / well_rate_bhp_model must carry an SrmPhysics engine")
        if (saturation_model is not None) != (self.engine.fluid == "GC"):
            raise ValueError("saturation_model goes with a gas-condensate ('GC') engine, and only with one")
        g = {**DEFAULT_GENERAL, **(general_config or {})}
        self.general_config = g
        self.physics_mode_fraction = float(g["physics_mode_fraction"])           # training.py:605
        self.fluid_type = self.engine.fluid
        # training.py:559-560.  The legacy two-phase arithmetic sums the gas and oil equations into ONE residual
        # per term (physics_loss.py:638,650,665,680), so the GC terms are reported under 'gas' and 'oil' holds zeros.
        self.loss_keys = {"gas": list(LOSS_KEYS)}
        if self.fluid_type == "GC":
            self.loss_keys["oil"] = list(LOSS_KEYS)
        w = dict(g["default_weights"]["gas"])
        if weights:
            w.update(weights)
        self.weights = w
        default_map = {"pressure": "pressure", "time_step": "time_step"}
        if saturation_model is not None:
            default_map["saturation"] = "saturation"
        self.optimizer_model_names_map = optimizer_model_names_map or default_map
        self.trainable_models_keys = list(self.optimizer_model_names_map.keys())  # training.py:554
        by_key = {"pressure": main_model, "time_step": time_step_model, "saturation": saturation_model}
        self.trainable_models = [by_key[k] for k in self.trainable_models_keys]   # training.py:553
        self.optimizer_model_map = {k: _OptimizerSlot((optimizers or {}).get(k)) for k in self.trainable_models_keys}
        self.t_min, self.t_max = float(g["srm_start_time"]), float(g["srm_end_time"])
        self.norm_lo, self.norm_hi = (float(v) for v in g["data_normalization"]["normalization_limits"])
        self.k_min, self.k_max = (float(v) for v in kx_stats)

    # -- feature handling (data_processing/data_processing_utils.py:1065-1183) ---------------------
    def _time_days(self, x):
        tn = x[:, 0, 0, 0, 3]
        return (self.t_max - self.t_min) * ((tn - self.norm_lo) / (self.norm_hi - self.norm_lo)) + self.t_min

    def _shift_time(self, x, dt):
        """x_n1 = x_n0 with t_norm += normalize_diff(dt)   (physics_loss.py:105-110); one CUDA pass over x
        (srm_features_forward), its cotangent w.r.t. dt by srm_features_backward"""
        dn = (self.norm_hi - self.norm_lo) / (self.t_max - self.t_min) * dt
        return _ShiftTimeFn.apply(self.engine, x.contiguous(), dn.contiguous())

    @staticmethod
    def _params(model):
        return [p for p in getattr(model, "parameters", lambda: [])() if p.requires_grad]

    def pinn_batch_sse_grad(self, x, y=None):
        """returns (wmse, wmse_grad, wsse, error_count, y_model)   -- training.py:607

        wmse[0][i] / wsse[0][i] follow loss_keys['gas'][i]; wmse_grad[i] is the gradient list of
        trainable_models[i] of the total weighted SSE (physics_loss.py:787-859)."""
        eng = self.engine
        x = x.to(eng.device, torch.float32)
        B = x.shape[0]
        # permeability channel, de-normalised, straight from the feature tensor (no strided copy of the channel)
        kx = eng.features_forward(x.contiguous(), None, (self.k_min, self.k_max), self.norm_lo, self.norm_hi)[1]
        sample_real = torch.arange(B, dtype=torch.int32, device=eng.device)
        dt1 = self.time_step_model(x).reshape(B, -1).mean(dim=1)                  # physics_loss.py:102
        x1 = self._shift_time(x, dt1)
        from .hard_layer import CompleteTrainableModule, fused_two_level
        if isinstance(self.main_model, CompleteTrainableModule) and self.main_model.use_hard_layer:
            # both HardLayer evaluations and the second time-step mean in one CUDA pass (srm_glue_forward)
            p0, p1, dt2 = fused_two_level(self.main_model, self.time_step_model, x, x1)
        else:
            p0 = self.main_model(x)[..., 0]                                       # physics_loss.py:88-95
            p1 = self.main_model(x1)[..., 0]                                      # physics_loss.py:111-115
            dt2 = self.time_step_model(x1).reshape(B, -1).mean(dim=1)             # physics_loss.py:122
        t1 = self._time_days(x1).detach().contiguous()
        if self.fluid_type == "GC":
            top = 1.0 - float(eng.spec.end_points["Swmin"])
            sg0 = self.saturation_model(x)[..., 0]                                # physics_loss.py:331,373
            sg1 = self.saturation_model(x1)[..., 0]
            so0, so1 = top - sg0, top - sg1                                       # relative_permeability.py:58
            terms = _SrmPhysicsGcFn.apply(eng, kx, sample_real, t1, p0, p1, sg0, sg1, so0, so1, dt1, dt2)
        else:
            terms = _SrmPhysicsFn.apply(eng, kx, sample_real, t1, p0, p1, dt1, dt2)
        wvec = torch.tensor([self.weights[k] for k in self.loss_keys["gas"]], device=eng.device)
        slots = torch.tensor([_SLOT[k] for k in self.loss_keys["gas"]], device=eng.device)
        loss = (wvec * terms[0][slots]).sum()                                     # physics_loss.py:809-819
        params = [self._params(m) for m in self.trainable_models]
        flat = [p for ps in params for p in ps]
        grads = torch.autograd.grad(loss, flat, allow_unused=True) if flat else []
        out, i = [], 0
        for ps in params:
            out.append([torch.zeros_like(p) if g is None else g for p, g in zip(ps, grads[i:i + len(ps)])])
            i += len(ps)
        # Ranks hold disjoint sample shards.  The adjoint's upstream weights are the constants `wvec`, not the reduced
        # terms, so the 128-byte all-reduce of the REPORTED terms is queued only now, behind the backward kernels: no
        # rank waits for another in the middle of a step (reducing inside the forward re-synchronised the ranks before
        # every adjoint).  The handle is waited on where the values are first read.
        terms = terms.detach().clone()
        sdist.allreduce_terms_async(terms).wait()
        wsse = wvec * terms[0][slots]
        counts = terms[1][slots]
        wmse = wsse / torch.clamp(counts, min=1.0)                                # zeros_to_ones, :835-846
        phases = [wmse.detach()], [wsse.detach()], [counts.detach()]
        if self.fluid_type == "GC":
            for lst in phases:
                lst.append(torch.zeros_like(lst[0]))
        return phases[0], out, phases[1], phases[2], p0.detach().unsqueeze(-1)
