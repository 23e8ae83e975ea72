"""Golden GRADIENTS of the dry-gas loss made by the REFERENCE'S OWN op graph.

What is executed (all of it cut out of /root/reference by AST or run as whole modules, nothing copied; TensorFlow
replaced by the torch-backed stand-in of this directory, whose GradientTape is torch autograd):

  * `pinn_batch_sse_grad` (physics_loss.py:742-870) with `physics_error_gas_2D` (physics_loss.py:9-224) behind it:
    the per-term weighted SSEs and `tape.gradient(term, model.trainable_variables)` exactly as the reference takes them
    (physics_loss.py:849-859);
  * the pipeline behind `model(x)` is the reference's too: `PVTLayer.call` (PVT_Layer_Subclassed.py:146-216: clamp,
    `PolyharmonicSplineInterpolationLayer` per property, derivative by the inner tape -- recorded by the outer tape, so
    the second derivative TensorFlow's nested tapes see is in the graph) and `WellRatesPressure.compute_rates_and_bhp`
    (well_rate_bhp_Subclassed.py:727-837 with the non-iterative control, the phase rates and the blocking-factor
    integral) evaluated on the level-(n+1) pressure with the same PVT layer.

`model.trainable_variables` are the NETWORK OUTPUTS themselves -- the pressure fields of both time levels and the two
time-step fields -- so the recorded gradients are what TensorFlow's tape hands back across the boundary this repository
implements (the cotangents of p0, p1 and of the time-step fields; the per-sample dL/ddt is the sum of a time-step field's
cotangent over the sample's cells, because the fragment takes `reduce_mean` of that field).

The spline layers take the oracle's (w, v) as their weights (the layer's own in-call solve differs from numpy's LAPACK
by <= 5e-6, pinned separately in make_reference_pvt_golden.py); tf.maximum / tf.minimum / tf.clip_by_value route ties
as TensorFlow does (first argument; inside the closed interval).

Output: tests/golden/reference_dg_grad.npz          python tests/golden/make_reference_grad_golden.py
"""
import logging
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import tf_torch_shim as tf                  # noqa: E402
import srm_oracle as O                      # noqa: E402
import make_reference_dg_golden as DG       # noqa: E402
import make_reference_pvt_golden as PV      # noqa: E402
import make_reference_wells_golden as WG    # noqa: E402

REF = "/root/reference"


def reference_pvt_layer(cols, tab, props, fluid):
    """the reference's PVTLayer over its own spline layers, with the oracle's (w, v) installed as the layer weights"""
    sp = PV.exec_module(os.path.join(REF, "polyhm_splines.py"))
    Layer = sp["PolyharmonicSplineInterpolationLayer"]
    knots = np.asarray(cols["Pre"], np.float32)
    lookup = {"pre": knots}
    names = {"InvBg": "invBg", "Invug": "invug", "InvBo": "invBo", "Invuo": "invuo", "Rs": "Rs", "Rv": "Rv", "Vro": "Vro"}
    for pi, prop in enumerate(props):
        lookup[names.get(prop, prop)] = np.asarray(tab.f[pi], np.float32)
    ns = {"tf": tf, "np": np, "PolyharmonicSplineInterpolationLayer": Layer}
    exec(PV.cut_class(os.path.join(REF, "PVT_Layer_Subclassed.py"), "PVTLayer"), ns)
    cfg = types.SimpleNamespace(lookup=lambda k: lookup[k])
    layer = ns["PVTLayer"](fluid_type=fluid, fitting_method="spline", spline_config=cfg, spline_order=tab.order, regularization_weight=0.001)
    layer(torch.full((1, 1, 1, 1), 4500.0))          # builds the spline layers
    for pi, prop in enumerate(layer.properties):
        sl = layer.spline_layers[prop]
        w = torch.as_tensor(tab.w[pi], dtype=torch.float32).reshape(1, -1, 1)
        v = torch.as_tensor(tab.v[pi], dtype=torch.float32).reshape(1, 2, 1)
        sl._solve_interpolation = (lambda w_, v_: (lambda *a, **k: (w_, v_)))(w, v)      # the layer re-solves in every call
    return layer


def reference_wells(cfg, D, H, W, fluid, blocking, relperm):
    """a bare WellRatesPressure instance with its attributes set as __init__ sets them (make_reference_wells_golden.py)"""
    wl = cfg.wells
    ns = {"tf": tf, "np": np, "logging": logging, "os": os}
    WDP = WG.build_class(os.path.join(REF, "welldata_processor.py"), "WellDataProcessor", ["scatter_y", "conn_shutins_idx"], ns)
    wdp = WDP.__new__(WDP)
    wdp.dtype = tf.float32
    conn = [(w.k, w.j, w.i) for w in wl]
    shape5 = (1, D, H, W, 1)
    ns2 = {"tf": tf, "np": np, "logging": logging, "os": os, "project_directory": "/tmp",
           "slice_tensor": lambda x, idx, dim=-1: x[..., idx[0]:idx[0] + 1]}
    WRP = WG.build_class(os.path.join(REF, "well_rate_bhp_Subclassed.py"), "WellRatesPressure",
                         ["compute_rates_and_bhp", "_non_iterative_method", "_compute_phase_rates", "compute_blocking_integral_and_factor",
                          "_split_condensate_components", "extract_pvt_properties", "_solve_newton"], ns2)
    w = WRP.__new__(WRP)
    w.fluid_type, w.use_blocking_factor, w.dtype, w.solver, w.n_intervals, w.n_root_iter = fluid, blocking, tf.float32, "newton", cfg.n_intervals, 20
    w.max_iters, w.tol, w.use_non_iterative, w.compute_mo = 10, 1e-6, True, fluid == "GC"
    w.kx_ky = tf.constant(cfg.kx_ky, dtype=tf.float32)
    w.dx = tf.constant(cfg.length, dtype=tf.float32) / W
    w.dy = tf.constant(cfg.width, dtype=tf.float32) / H
    w.dz = tf.constant(cfg.thickness, dtype=tf.float32) / D
    w.C = cfg.C
    w.well_data_processor = wdp
    w.well_data = {"connection_index": conn, "shutin_days": [[list(x.shutin_days)] for x in wl]}
    w.well_id = wdp.scatter_y(shape5, conn, 1.0)
    w.rw = wdp.scatter_y(shape5, conn, [x.wellbore_radius for x in wl])
    w.q0 = wdp.scatter_y(shape5, conn, [x.value for x in wl])
    w.pwf_min = wdp.scatter_y(shape5, conn, [x.minimum_bhp for x in wl])
    w.completion_ratio = wdp.scatter_y(shape5, conn, [x.completion_ratio for x in wl])
    w.scal_config = {"end_points": {"Swmin": cfg.Swmin}}
    w.relperm = types.SimpleNamespace(end_points={"Swmin": cfg.Swmin})
    w.norm_config = None
    w.data_summary = types.SimpleNamespace(get_key_index=lambda k: {"time": 3, "permx": 4}[k], nonormalize=lambda v, **kw: v)
    w.log_tensor_to_file = lambda *a, **k: None
    w._relperm_fn = relperm
    return w


class GraphModel:
    """stand-in for the Keras pipeline whose outputs are differentiable leaves: pressure and time-step field of each
    level; PVT and well rates are computed from them by the reference's own layers"""

    def __init__(self, cfg, P, DTF, pvt_layer, wells, kx5, cfd_type):
        self.dtype = tf.float32
        self.cfd_type = cfd_type
        self.P, self.DTF = P, DTF
        self.pvt, self.wells, self.kx5 = pvt_layer, wells, kx5
        self.calls = 0
        self.cf = float(O.rock_compressibility(cfg.phi))
        self.Sgi = float(np.float32(1.0 - cfg.Swmin))

    def __call__(self, x, training=True):
        lv = min(self.calls, 1)
        self.calls += 1
        p = self.P[lv]                                           # (B, H, W, 1)
        pv = self.pvt(p)                                         # [2, n_prop, B, H, W, 1]
        invBg, invug, dinvBg = pv[0][0], pv[0][1], pv[1][0]
        one = torch.ones_like(p)
        q = torch.zeros_like(p)
        if lv == 1:                                              # the rate field of level n+1 enters the residual (physics_loss.py:166)
            B, H, W, _ = p.shape
            x5 = torch.zeros(B, 1, H, W, 5)
            x5[..., 3] = x[3].reshape(B, 1, H, W)
            x5[..., 4] = self.kx5
            sg = torch.full((B, 1, H, W, 1), self.Sgi)
            rates, _pwf = self.wells.compute_rates_and_bhp(x5, p.reshape(B, 1, H, W, 1), sg, self.wells._relperm_fn, self.pvt)
            q = rates.reshape(B, H, W, 1)
        return [p, one, invBg, invug, torch.stack([dinvBg, one]), self.DTF[lv], one, q, one]


def run_case(W, H, B, R, seed, dts, blocking, bhp_limited=False):
    rng = np.random.default_rng(seed)
    wl = O.default_wells(W, H, 1)
    if bhp_limited:
        wl[0].value = 2.0e5                                      # a target the reservoir cannot deliver: dq/dp is live there
    cfg = O.OracleConfig(D=1, H=H, W=W, wells=wl, use_blocking_factor=blocking, n_intervals=8)
    cols = O.load_pvt_table(os.path.join(HERE, "pvt_table.npz"))
    tab = O.build_spline_table(cols, O.DG_PROPS, order=1, lam=0.001)
    kx = np.exp(rng.normal(np.log(3.0), 0.45, size=(R, 1, H, W))).astype(np.float32)
    sample_real = (np.arange(B) % R).astype(np.int32)
    p0 = (5000.0 - 300.0 * rng.random((B, 1, H, W)) - 2.0 * rng.standard_normal((B, 1, H, W))).astype(np.float32)
    p1 = (p0 - 30.0 * rng.random((B, 1, H, W))).astype(np.float32)
    dt1 = np.asarray([dts[(2 * b) % len(dts)] for b in range(B)], np.float32)
    dt2 = np.asarray([dts[(2 * b + 1) % len(dts)] for b in range(B)], np.float32)
    t_days = np.linspace(10.0, 300.0, B).astype(np.float32)     # level n; the wells of level n+1 see t_days + dt1
    tt = torch.as_tensor
    ch = lambda a: a.reshape(B, H, W, 1)
    field = lambda v: ch(torch.as_tensor(np.broadcast_to(v.reshape(B, 1, 1, 1), (B, 1, H, W)).copy()))
    leaf = lambda t: t.clone().requires_grad_(True)
    P = [leaf(ch(tt(p0))), leaf(ch(tt(p1)))]
    DTF = [leaf(field(dt1)), leaf(field(dt2))]
    _, krg = O.corey_krog_krgo_np(np.float32(1.0 - cfg.Swmin), cfg, np.float32)
    cfd = {
        "Dimension": {"Gridblock_Dim": [cfg.dx, cfg.dy, cfg.dz], "Dim": [H, W, 1], "Measurement": [cfg.length, cfg.width, cfg.thickness]},
        "Conn_Idx": torch.tensor([[w.j, w.i, 0] for w in wl], dtype=torch.int32),
        "Init_Grate": torch.tensor([w.value for w in wl], dtype=torch.float32),
        "Min_BHP": torch.tensor([w.minimum_bhp for w in wl], dtype=torch.float32),
        "Completion_Ratio": 0.5, "SCAL": {"End_Points": {"Swmin": cfg.Swmin}}, "Max_Train_Time": 365.0, "Pi": 5000.0,
        "Init_InvBg": 1.0, "Init_DinvBg": 0.0, "Init_Invug": 1.0,
        "Kr_gas_oil": lambda sg: (torch.tensor(0.0), torch.tensor(float(krg), dtype=torch.float32)),
        "Connection_Shutins": {"Days": [], "Shutins_Idx": [], "Shutins_Per_Conn_Idx": []},
    }
    relperm = lambda s: O.corey_krog_krgo_t(s if isinstance(s, torch.Tensor) else torch.tensor(np.float32(s)), cfg, torch.float32)
    pvt_layer = reference_pvt_layer(cols, tab, O.DG_PROPS, "DG")
    wells = reference_wells(cfg, 1, H, W, "DG", blocking, relperm)
    sr = torch.as_tensor(sample_real.astype(np.int64))
    kxb = tt(kx).index_select(0, sr)                             # (B, 1, H, W)
    model = GraphModel(cfg, P, DTF, pvt_layer, wells, kxb, cfd)
    nwt = [1.0, 0.0, 0.0, 0.5, 0.0, 2.0, 0.0, 0.0]               # dom, dbc, nbc, ibc, ic, mbc, cmbc, td
    model.nwt = torch.tensor(nwt, dtype=torch.float32)
    model.nT, model.nT_list = 1, [0]
    model.trainable_variables = [P[0], P[1], DTF[0], DTF[1]]
    ns = {"tf": tf, "nonormalize": lambda m, v, stat_idx=None, compute=True: v, "normalize_diff": lambda m, v, stat_idx=None, compute=True: v,
          "dnn": types.SimpleNamespace(conn_shutins_idx=lambda t, ci, days: torch.zeros_like(t)),
          "time_shifting": lambda m, x, **k: (x, 1.0, torch.tensor(1e30)),
          "zeros_to_ones": lambda c: torch.where(c == 0, torch.ones_like(c), c)}
    exec(DG.reference_function("physics_error_gas_2D"), ns)
    exec(DG.reference_function("pinn_batch_sse_grad"), ns)
    model.loss_func = {"Physics_Error": ns["physics_error_gas_2D"], "Reshape": lambda y: y, "Reduce_Axis": [1, 2, 3, 4], "Squeeze_Out": lambda y: y}
    x = [torch.zeros(B, H, W, 1), torch.zeros(B, H, W, 1), torch.zeros(B, H, W, 1), field(t_days),
         torch.full((B, H, W, 1), float(np.float32(cfg.phi))), ch(kxb)]
    y = [torch.zeros(B, H, W, 1)]
    wsse, wsse_grad, count, wmse, y_model = ns["pinn_batch_sse_grad"](model, x, y)
    assert model.calls == 2
    back = lambda a: a.detach().reshape(B, 1, H, W).numpy()
    out = dict(W=W, H=H, B=B, R=R, blocking=int(blocking), kx=kx, sample_real=sample_real, p0=p0, p1=p1, dt1=dt1, dt2=dt2,
               t_days=t_days, t1=(t_days + dt1).astype(np.float32), nwt=np.asarray(nwt, np.float32),
               wells=np.asarray([[w.i, w.j, w.k, w.value] for w in wl], np.float32),
               wsse=np.asarray([float(v) for v in wsse[:8]], np.float64))
    # term order of the returned lists: batch, dom, dbc, nbc, ibc, ic, mbc, cmbc, td (physics_loss.py:864-865)
    for name, i in (("batch", 0), ("dom", 1), ("ibc", 4), ("mbc", 6)):
        g = wsse_grad[i]
        out[f"g_{name}_p0"], out[f"g_{name}_p1"] = back(g[0]), back(g[1])
        out[f"g_{name}_dt1"] = g[2].detach().reshape(B, -1).sum(dim=1).numpy()       # d/d(per-sample mean) = sum over the field
        out[f"g_{name}_dt2"] = g[3].detach().reshape(B, -1).sum(dim=1).numpy()
    return out


def main():
    out = {}
    cases = {"a": dict(W=12, H=9, B=4, R=2, seed=5100, dts=[0.5, 2.25, 7.125, 1.0, 0.375, 9.5], blocking=False),
             "b": dict(W=16, H=11, B=3, R=3, seed=5102, dts=[4.0, 0.75, 1.5, 6.25], blocking=True),
             "c": dict(W=12, H=9, B=3, R=1, seed=5103, dts=[2.0, 0.5, 8.0], blocking=False, bhp_limited=True)}
    for name, kw in cases.items():
        r = run_case(**kw)
        print(name, "wsse", r["wsse"], " |g_batch_p1| max", np.abs(r["g_batch_p1"]).max(), " |g_batch_p0| max", np.abs(r["g_batch_p0"]).max())
        for k, v in r.items():
            out[f"{name}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "reference_dg_grad.npz"), **out)
    print("wrote reference_dg_grad.npz")


if __name__ == "__main__":
    main()
