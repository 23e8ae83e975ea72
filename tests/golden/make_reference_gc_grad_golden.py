"""Golden GRADIENTS of the gas-condensate (two-phase) loss made by the REFERENCE'S OWN op graph.

Same method as make_reference_grad_golden.py (dry gas): `pinn_batch_sse_grad` (physics_loss.py:742-870) over
`physics_error_gas_oil_2D` (physics_loss.py:230-714); behind `model(x)` the reference's own `PVTLayer` (seven
properties, derivatives by the nested tape), `RelativePermeability.compute_krog_krgo` (relative_permeability.py, run as a
whole module) and `WellRatesPressure.compute_rates_and_bhp` (GC branch with `_split_condensate_components`), all executed
through the torch-backed TensorFlow stand-in.  `model.trainable_variables` are the network outputs: pressure, gas and oil
saturation of both time levels and the two time-step fields; the recorded gradients are tape.gradient of each weighted
SSE term with respect to them.

Output: tests/golden/reference_gc_grad.npz          python tests/golden/make_reference_gc_grad_golden.py
"""
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
import make_reference_grad_golden as GG     # noqa: E402

REF = "/root/reference"


def reference_relperm(cfg):
    rp = PV.exec_module(os.path.join(REF, "relative_permeability.py"), {"np": np})
    ep = dict(Swmin=cfg.Swmin, Sorg=cfg.Sorg, Sgc=cfg.Sgc, Socr=cfg.Socr, So_max=cfg.So_max, kro_Somax=cfg.kro_Somax,
              krg_Sorg=cfg.krg_Sorg, krg_Swmin=cfg.krg_Swmin)
    model = rp["RelativePermeability"](end_points=ep, corey_exponents=dict(nog=cfg.nog, ng=cfg.ng), dtype=tf.float32)
    return model.compute_krog_krgo


class GraphModelGC:
    """differentiable stand-in of the Keras pipeline (two-phase): leaves P, SG, SO, DTF per level"""

    def __init__(self, cfg, P, SG, SO, DTF, pvt_layer, wells, relperm, kxb, cfd_type):
        self.dtype = tf.float32
        self.cfd_type = cfd_type
        self.P, self.SG, self.SO, self.DTF = P, SG, SO, DTF
        self.pvt, self.wells, self.relperm, self.kxb = pvt_layer, wells, relperm, kxb
        self.calls = 0
        self.cf = float(O.rock_compressibility(cfg.phi))
        self.PVT = None

    def __call__(self, x, training=True):
        lv = min(self.calls, 1)
        self.calls += 1
        p, sg, so = self.P[lv], self.SG[lv], self.SO[lv]
        B, H, W, _ = p.shape
        pv = self.pvt(p)                                          # [2, 7, B, H, W, 1]: invBg invBo invug invuo Rs Rv Vro
        val, der = pv[0], pv[1]
        one = torch.ones_like(p)
        z = torch.zeros_like(p)
        q4, pwf = [z, z, z, z], z
        if lv == 1 and len(self.wells.well_data["connection_index"]):
            x5 = torch.zeros(B, 1, H, W, 5)
            x5[..., 3] = x[3].reshape(B, 1, H, W)
            x5[..., 4] = self.kxb
            r5 = lambda t: t.reshape(B, 1, H, W, 1)
            rates, pw = self.wells.compute_rates_and_bhp(x5, r5(p), r5(sg), self.relperm, self.pvt)
            q4 = [r.reshape(B, H, W, 1) for r in rates]
            pwf = pw.reshape(B, H, W, 1)
        return [p, sg, so, val[0], val[1], val[2], val[3], val[4], val[5], one, der, self.DTF[lv], one, q4, pwf]


def run_case(seed, B, H, W, wells, R=1, sg_lo=0.2, sg_hi=0.75, small_dp=False, dts=(0.5, 2.25, 7.125, 1.0, 0.375, 9.5), blocking=False):
    cols = O.load_pvt_table(os.path.join(HERE, "pvt_table.npz"))
    tab = O.build_spline_table(cols, O.GC_PROPS, order=1, lam=0.001)
    if wells == "two":
        wl = [dict(i=2, j=2, k=0, value=500.0), dict(i=W - 2, j=H - 2, k=0, value=1000.0)]
    elif wells == "three":    # neighbouring well cells, one target the reservoir cannot deliver (BHP limited: dq/dp, dq/dSg live)
        wl = [dict(i=2, j=2, k=0, value=500.0), dict(i=6, j=4, k=0, value=300.0), dict(i=3, j=2, k=0, value=2.0e5)]
    else:
        wl = []
    cfg = O.OracleConfig(D=1, H=H, W=W, wells=[O.Well(**w) for w in wl], use_blocking_factor=blocking, n_intervals=8)
    rng = np.random.default_rng(seed)
    shp = (B, 1, H, W)
    d = dict(kx=rng.uniform(1, 6, (R, 1, H, W)).astype(np.float32))
    d["p0"] = (4700 + rng.uniform(-40, 40, shp)).astype(np.float32)
    d["p1"] = (d["p0"] - rng.uniform(-3 if small_dp else 1, 25, shp)).astype(np.float32)
    d["sg0"] = rng.uniform(sg_lo, sg_hi, shp).astype(np.float32)
    d["sg1"] = (d["sg0"] - rng.uniform(0.001, 0.02, shp)).astype(np.float32)
    d["so0"] = (np.float32(0.78) - d["sg0"]).astype(np.float32)
    d["so1"] = (np.float32(0.78) - d["sg1"]).astype(np.float32)
    d["dt1"] = np.asarray([dts[(2 * b) % len(dts)] for b in range(B)], np.float32)
    d["dt2"] = np.asarray([dts[(2 * b + 1) % len(dts)] for b in range(B)], np.float32)
    t0 = np.linspace(5, 50, B).astype(np.float32)
    d["t1"] = (t0 + d["dt1"]).astype(np.float32)                 # the wells of level n+1 see t + dt1
    d["sample_real"] = (np.arange(B) % R).astype(np.int32)
    tt = torch.as_tensor
    ch = lambda a: tt(a).reshape(B, H, W, 1)
    field = lambda v: torch.as_tensor(np.broadcast_to(v.reshape(B, 1, 1, 1), (B, H, W, 1)).copy())
    leaf = lambda t: t.clone().requires_grad_(True)
    P = [leaf(ch(d["p0"])), leaf(ch(d["p1"]))]
    SG = [leaf(ch(d["sg0"])), leaf(ch(d["sg1"]))]
    SO = [leaf(ch(d["so0"])), leaf(ch(d["so1"]))]
    DTF = [leaf(field(d["dt1"])), leaf(field(d["dt2"]))]
    relperm = reference_relperm(cfg)
    wells_o = cfg.wells
    cfd = {
        "Dimension": {"Gridblock_Dim": [cfg.dx, cfg.dy, cfg.dz], "Dim": [H, W, 1], "Measurement": [cfg.length, cfg.width, cfg.thickness]},
        "Conn_Idx": torch.tensor([[w.j, w.i, 0] for w in wells_o], dtype=torch.int32).reshape(-1, 3),
        "Init_Grate": torch.tensor([w.value for w in wells_o], dtype=torch.float32),
        "Min_BHP": torch.tensor([w.minimum_bhp for w in wells_o], dtype=torch.float32),
        "Completion_Ratio": 0.5, "SCAL": {"End_Points": {"Swmin": cfg.Swmin, "Sorg": cfg.Sorg}}, "Max_Train_Time": 365.0, "Pi": 5000.0,
        "Dew_Point": 4048.49, "Rhg_Std": 0.05, "Rho_Std": 50.0,
        "Init_InvBg": 1.0, "Init_DinvBg": 0.0, "Init_Invug": 1.0, "Init_InvBo": 1.0, "Init_Invuo": 1.0, "Init_Rs": 1.0, "Init_Rv": 1.0,
        "Kr_gas_oil": relperm,
    }
    pvt_layer = GG.reference_pvt_layer(cols, tab, O.GC_PROPS, "GC")
    wells_m = GG.reference_wells(cfg, 1, H, W, "GC", blocking, relperm) if wl else types.SimpleNamespace(well_data={"connection_index": []})
    sr = torch.as_tensor(d["sample_real"].astype(np.int64))
    kxb = tt(d["kx"]).index_select(0, sr)
    model = GraphModelGC(cfg, P, SG, SO, DTF, pvt_layer, wells_m, relperm, kxb, cfd)
    nwt = [1.0, 0.0, 0.0, 0.5, 0.0, 2.0, 0.25, 0.0]               # dom, dbc, nbc, ibc, ic, mbc, cmbc, td
    model.nwt = torch.tensor(nwt, dtype=torch.float32)
    model.nT, model.nT_list = 1, [0]
    model.trainable_variables = [P[0], P[1], SG[0], SG[1], SO[0], SO[1], DTF[0], DTF[1]]
    ident = lambda model, v, stat_idx=None, compute=True: v if isinstance(v, torch.Tensor) else torch.tensor(float(v))
    ns = {"tf": tf, "nonormalize": ident, "normalize_diff": ident, "normalize": ident,
          "dnn": types.SimpleNamespace(conn_shutins_idx=lambda t, ci, days: torch.zeros_like(t)),
          "time_shifting": lambda model, x, **k: (x, 1.0, torch.tensor(1e30)),
          "zeros_to_ones": lambda c: torch.where(c == 0, torch.ones_like(c), c)}
    exec(DG.reference_function("physics_error_gas_oil_2D"), ns)
    exec(DG.reference_function("pinn_batch_sse_grad"), ns)
    model.loss_func = {"Physics_Error": ns["physics_error_gas_oil_2D"], "Reshape": lambda y: y, "Reduce_Axis": [1, 2, 3, 4], "Squeeze_Out": lambda y: y}
    z = torch.zeros(B, H, W, 1)
    x = [z.clone(), z.clone(), z.clone(), field(t0), torch.full((B, H, W, 1), float(np.float32(cfg.phi))), kxb.reshape(B, H, W, 1)]
    wsse, wsse_grad, count, wmse, y_model = ns["pinn_batch_sse_grad"](model, x, [z.clone()])
    assert model.calls == 2
    back = lambda a: a.detach().reshape(B, 1, H, W).numpy()
    out = {k: v for k, v in d.items()}
    out.update(W=W, H=H, B=B, R=R, blocking=int(blocking), nwt=np.asarray(nwt, np.float32),
               wells=np.asarray([[w["i"], w["j"], w["k"], w["value"]] for w in wl], np.float32).reshape(-1, 4),
               wsse=np.asarray([float(v.detach()) if isinstance(v, torch.Tensor) else float(v) for v in wsse[:8]], np.float64))
    for name, i in (("batch", 0), ("dom", 1), ("ibc", 4), ("mbc", 6), ("cmbc", 7)):
        g = wsse_grad[i]
        for j, f in enumerate(("p0", "p1", "sg0", "sg1", "so0", "so1")):
            out[f"g_{name}_{f}"] = back(g[j])
        out[f"g_{name}_dt1"] = g[6].detach().reshape(B, -1).sum(dim=1).numpy()
        out[f"g_{name}_dt2"] = g[7].detach().reshape(B, -1).sum(dim=1).numpy()
    return out


def main():
    out = {}
    cases = {"a": dict(seed=5311, B=3, H=9, W=8, wells="two"),
             "b": dict(seed=5312, B=4, H=7, W=10, wells="three", R=2, sg_lo=0.2, sg_hi=0.35),
             "c": dict(seed=5313, B=2, H=6, W=7, wells="none", small_dp=True, sg_lo=0.2, sg_hi=0.5),
             # the blocking-factor integral: twenty Newton iterations per trapezoid node, all of them inside the tape
             "d": dict(seed=5314, B=2, H=7, W=10, wells="three", sg_lo=0.3, sg_hi=0.6, blocking=True)}
    for name, kw in cases.items():
        r = run_case(**kw)
        print(name, "wsse", r["wsse"], " max|g_batch|:", {f: float(np.abs(r[f"g_batch_{f}"]).max()) for f in ("p0", "p1", "sg0", "sg1", "so0", "so1", "dt1")})
        for k, v in r.items():
            out[f"{name}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "reference_gc_grad.npz"), **out)
    print("wrote reference_gc_grad.npz")


if __name__ == "__main__":
    main()
