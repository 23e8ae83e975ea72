"""HardLayer / CompleteTrainableModule: host-side mirrors of the reference's output layer
(Hard_Layer_Subclassed.py:21-260) and of the wrapper the example builds its pressure model from
(complete_trainable_module.py:27-176), plus the fused "glue" of the physics loss: both time
levels of the layer and the per-sample means of the time-step field in ONE CUDA pass either side of the residual
kernels, with the cotangents tape.gradient would deliver.

torch modules stand in for the Keras layers (the torch-harness twin of the TF binding).  The arithmetic
runs in libsrm_physics.so (srm_glue_forward / srm_glue_backward); there is no CPU fallback.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch


class _GlueFn(torch.autograd.Function):
    """(p0, p1, dt1, dt2) = glue(tn0, tn1, y0, y1, expo, dtf1, dtf2).  The layer's time input is differentiable
    (Hard_Layer_Subclassed.py:214-228 reads it from the feature tensor with no stop_gradient): at level n+1 it is
    t_n + normalize_diff(dt1) (physics_loss.py:105-111), so tape.gradient reaches the time-step model through it."""

    @staticmethod
    def forward(ctx, engine, init_value, t_lo, t_hi, tn0, tn1, y0, y1, expo, dtf1, dtf2):
        c = lambda t: None if t is None else t.detach().contiguous()
        y0c, y1c, ec, d1c, d2c = c(y0), c(y1), c(expo), c(dtf1), c(dtf2)
        tn0, tn1 = tn0.detach().contiguous(), tn1.detach().contiguous()
        p0, p1, dt1, dt2 = engine.glue_forward(y0c, y1c, tn0, tn1, ec, d1c, d2c, init_value, t_lo, t_hi)
        ctx.engine, ctx.k = engine, (init_value, t_lo, t_hi)
        ctx.has = (expo is not None, dtf1 is not None, dtf2 is not None)
        ctx.save_for_backward(tn0, tn1, y0c, y1c, ec if ec is not None else torch.empty(0, device=y0c.device))
        z = torch.zeros(0, device=y0c.device)
        return p0, p1, (dt1 if dt1 is not None else z), (dt2 if dt2 is not None else z)

    @staticmethod
    def backward(ctx, gp0, gp1, gdt1, gdt2):
        tn0, tn1, y0, y1, expo = ctx.saved_tensors
        has_e, has1, has2 = ctx.has
        want_t = (bool(ctx.needs_input_grad[4]), bool(ctx.needs_input_grad[5]))
        res = ctx.engine.glue_backward(
            y0, y1, tn0, tn1, gp0.contiguous(), gp1.contiguous(), expo if has_e else None,
            gdt1.contiguous() if has1 else None, gdt2.contiguous() if has2 else None, *ctx.k, want_gexpo=has_e, want_gtn=want_t)
        gy0, gy1, gexpo, gdtf1, gdtf2 = res[:5]
        gtn0, gtn1 = (res[5], res[6]) if (want_t[0] or want_t[1]) else (None, None)
        return None, None, None, None, gtn0, gtn1, gy0, gy1, gexpo if has_e else None, gdtf1, gdtf2


_ACTIVATIONS = {"sigmoid": torch.sigmoid, "relu": torch.relu, "tanh": torch.tanh}


def _as_activation(a):
    """None / '' / a name the reference treats as identity -> None; a callable stays; a list keeps its last entry
    (Hard_Layer_Subclassed.py:96-103, 157-163, 231)"""
    if isinstance(a, (list, tuple)):
        a = a[-1] if len(a) else None
    return a if callable(a) else None


class HardLayer(torch.nn.Module):
    """HardLayer(norm_limits=[-1, 1], init_value=..., kernel_exponent_config={...}, use_rbf=False, rbf_config=None,
    kernel_activation=None, input_activation=None, rectifier=None)  -- Hard_Layer_Subclassed.py:29-118

    call([[time, property], p(, rect_input)]) -> init_value - alpha * input_activation(p),
        alpha = alpha_p * alpha_t ** kernel_activation(kernel_exponent) (* rbf_dense(property) with use_rbf),
        alpha_p = rectifier((rect_input - pdew) / (pmin - pdew)) when a rectifier and a third input are given, else 1
    (Hard_Layer_Subclassed.py:196-246).  kernel_exponent is trainable per cell (shape (D, H, W), constant-initialised,
    clipped to [min_value, max_value] after each optimiser step as Keras' MinMaxNorm does for a one-element axis).

    The power alpha_t ** e and the subtraction run in the CUDA glue (srm_glue_forward / _backward).  The options the
    example leaves off are element-wise factors of the network output -- init - alpha_t^e * (alpha_p * rbf * act(p)) is the
    same product -- so they are applied to it here, in torch, before the kernel; autograd carries their cotangents
    (dense kernel / bias, rectifier input)."""

    def __init__(self, engine, norm_limits: Sequence[float] = (-1.0, 1.0), init_value: float = 1.0,
                 kernel_exponent_config: Optional[dict] = None, use_rbf: bool = False, rbf_config: Optional[dict] = None,
                 kernel_activation=None, input_activation=None, rectifier=None, pdew: Optional[float] = None,
                 pmin: Optional[float] = None, name: str = "hard_layer"):
        super().__init__()
        self.engine = engine
        self.norm_limits = (float(norm_limits[0]), float(norm_limits[1]))
        self.init_value = float(init_value)
        cfg = {"initial_value": 0.5, "trainable": True, "min_value": 0.01, "max_value": 0.99, **(kernel_exponent_config or {})}
        iv = cfg["initial_value"]
        iv = float(iv[0] if isinstance(iv, (tuple, list)) else iv)       # the example passes a 1-tuple (training_case_dry_gas_i.py:94)
        self.kernel_exponent_config = cfg
        shape = (engine.spec.D, engine.spec.H, engine.spec.W)
        self.kernel_exponent = torch.nn.Parameter(torch.full(shape, iv, dtype=torch.float32, device=engine.device),
                                                  requires_grad=bool(cfg["trainable"]))
        self.kernel_activation = _as_activation(kernel_activation)
        self.input_activation = _as_activation(input_activation)
        self.rectifier = rectifier
        if rectifier is not None and (pdew is None or pmin is None):
            raise ValueError("a rectifier needs the dew point and the abandonment pressure (Hard_Layer_Subclassed.py:111-126, 223)")
        self.pdew, self.pmin = pdew, pmin
        self.use_rbf = bool(use_rbf)
        self.rbf_config = rbf_config or {"output_dim": 25, "activation": "sigmoid"}
        if self.use_rbf:                                                  # Dense(1, activation, glorot_normal, UnitNorm(axis=0))  :168-187
            self.rbf_dense = torch.nn.Linear(1, 1).to(engine.device)
            torch.nn.init.xavier_normal_(self.rbf_dense.weight)
            torch.nn.init.zeros_(self.rbf_dense.bias)
            self.rbf_activation = _ACTIVATIONS.get(self.rbf_config.get("activation"))
        self.name = name

    def apply_constraint(self):
        with torch.no_grad():
            self.kernel_exponent.clamp_(float(self.kernel_exponent_config["min_value"]), float(self.kernel_exponent_config["max_value"]))
            if self.use_rbf:                                              # UnitNorm(axis=0): w / (eps + ||w||)
                w = self.rbf_dense.weight
                w.div_(1e-7 + w.norm(dim=1, keepdim=True))

    def exponent(self):
        return self.kernel_exponent if self.kernel_activation is None else self.kernel_activation(self.kernel_exponent)

    def scaled_output(self, y, prop=None, rect_input=None):
        """input_activation(y) times the factors of alpha that do not depend on time: rectifier and rbf (fields like y)"""
        if self.input_activation is not None:
            y = self.input_activation(y)
        if self.rectifier is not None and rect_input is not None:
            r = rect_input[..., 0] if rect_input.dim() == y.dim() + 1 else rect_input
            y = self.rectifier((r - self.pdew) / (self.pmin - self.pdew)) * y
        if self.use_rbf:
            if prop is None:
                raise ValueError("use_rbf needs the property channel")
            f = self.rbf_dense(prop if prop.dim() == y.dim() + 1 else prop.unsqueeze(-1))[..., 0]
            y = (self.rbf_activation(f) if self.rbf_activation is not None else f) * y
        return y

    def forward(self, inputs, p=None):
        """inputs = [[time, property], p(, rect_input)] (the reference's list form) or (time, p); time (B,D,H,W,1) or (B,)"""
        prop = rect = None
        if p is None:
            (time, prop), p = inputs[0], inputs[1]
            rect = inputs[2] if len(inputs) > 2 else None
        else:
            time = inputs
        tn = time.reshape(time.shape[0], -1)[:, 0]
        y = self.scaled_output(p[..., 0] if p.dim() == 5 else p, prop, rect)
        # one level: the second slot re-uses the inputs detached (its outputs are dropped, its cotangents are zero)
        out, _, _, _ = _GlueFn.apply(self.engine, self.init_value, *self.norm_limits, tn, tn.detach(), y, y.detach(), self.exponent(), None, None)
        return out.unsqueeze(-1) if p.dim() == 5 else out


class CompleteTrainableModule(torch.nn.Module):
    """CompleteTrainableModule(main_network, hard_layer): call(inputs, rectifier_input=None, training=False)
    (complete_trainable_module.py:142-176): network output through the HardLayer, time = channel -2 and property =
    channel -1 of the feature tensor (DEFAULT_INPUT_SLICE_CONFIG, default_configurations.py:217-225)."""

    def __init__(self, main_network, hard_layer: Optional[HardLayer] = None, use_hard_layer: bool = True):
        super().__init__()
        self.main_network = main_network
        self.hard_layer = hard_layer
        self.use_hard_layer = bool(use_hard_layer and hard_layer is not None)

    def forward(self, inputs, rectifier_input=None, training: bool = False):
        y = self.main_network(inputs)
        if not self.use_hard_layer:
            return y
        hl_inputs = [[inputs[..., -2:-1], inputs[..., -1:]], y]
        if rectifier_input is not None and self.hard_layer.rectifier is not None:      # complete_trainable_module.py:176-179
            hl_inputs.append(rectifier_input)
        return self.hard_layer(hl_inputs)


def fused_two_level(module: CompleteTrainableModule, time_step_model, x0, x1):
    """Both evaluations of the physics loss' glue in one CUDA pass: p_n = module(x_n), p_n1 = module(x_n1) and
    dt1 = mean(time_step_model(x_n)), dt2 = mean(time_step_model(x_n1))   (physics_loss.py:88-122).
    x_n1 must already carry the shifted time (it depends on dt1: the caller evaluates time_step_model(x_n) first)."""
    hl = module.hard_layer
    y0 = hl.scaled_output(module.main_network(x0)[..., 0], x0[..., -1:])
    y1 = hl.scaled_output(module.main_network(x1)[..., 0], x1[..., -1:])
    tn0, tn1 = x0[:, 0, 0, 0, -2], x1[:, 0, 0, 0, -2]        # differentiable: x1 carries t_n + normalize_diff(dt1)
    dtf2 = time_step_model(x1)[..., 0]
    p0, p1, _, dt2 = _GlueFn.apply(hl.engine, hl.init_value, *hl.norm_limits, tn0, tn1, y0, y1, hl.exponent(), None, dtf2)
    return p0, p1, dt2
