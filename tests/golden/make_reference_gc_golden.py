"""Golden vectors for the gas-condensate (two-phase) residual made by the REFERENCE'S OWN code.

`physics_error_gas_oil_2D` (physics_loss.py:230-714) is cut out of /root/reference/physics_loss.py by AST and executed
on seeded inputs through the torch-backed TensorFlow stand-in of this directory (see make_reference_dg_golden.py for
the method).  The Keras pipeline is replaced by a stand-in returning GIVEN fields per time level: pressure, saturations,
the PVT values and dp-derivatives (evaluated by the oracle's spline -- itself pinned by make_reference_pvt_golden.py),
the time-step field and the four well-rate fields (oracle's WellRatesPressure restatement); `Kr_gas_oil` is the
oracle's Corey function (tf.pow pinned as a product; the reference class agrees to 2 ulp, same generator).  What is
recorded is the fragment's own arithmetic: masses and truncation terms, chord slopes, product-rule PVT derivatives,
upstream-weighted relative permeabilities, the sixteen face coefficients, four accumulation and four divergence terms,
dom, ibc, mbc, cmbc.

Output: tests/golden/reference_gc_residual.npz
"""
import ast
import os
import sys
import textwrap
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (HERE, os.path.join(ROOT, "tests"), os.path.join(ROOT, "oracle"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)
import tf_torch_shim as tf          # noqa: E402
import srm_oracle as O              # noqa: E402

REF = "/root/reference/physics_loss.py"


def reference_function(name):
    src = open(REF).read()
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == name)
    return textwrap.dedent(ast.get_source_segment(src, fn))


class FakeModel:
    def __init__(self, cfg, levels, cfd_type):
        self.dtype = tf.float32
        self.cfd_type = cfd_type
        self.levels = levels
        self.calls = 0
        self.cf = float(O.rock_compressibility(cfg.phi))
        self.PVT = None
        self.trainable_variables = []

    def __call__(self, x, training=True):
        lv = self.levels[min(self.calls, 1)]
        self.calls += 1
        one = torch.ones_like(lv["p"])
        return [lv["p"], lv["sg"], lv["so"], lv["invBg"], lv["invBo"], lv["invug"], lv["invuo"], lv["Rs"], lv["Rv"], one,
                lv["dpvt"], lv["dtf"], one, lv["q4"], lv["pwf"]]


def run_case(seed, B, H, W, wells, R=1, sg_lo=0.2, sg_hi=0.75, small_dp=False, dts=(0.5, 2.25, 7.125, 1.0, 0.375, 9.5)):
    cols = O.load_pvt_table(os.path.join(HERE, "pvt_table.npz"))
    tab = O.build_spline_table(cols, O.GC_PROPS, order=1, lam=0.001)
    if wells == "two":
        wl = [dict(i=2, j=2, k=0, value=500.0), dict(i=W - 2, j=H - 2, k=0, value=1000.0)]
    elif wells == "dup":
        wl = [dict(i=2, j=2, k=0, value=500.0), dict(i=2, j=2, k=0, value=300.0), dict(i=3, j=2, k=0, value=800.0)]
    else:
        wl = []
    cfg = O.OracleConfig(D=1, H=H, W=W, wells=[O.Well(**w) for w in wl])
    rng = np.random.default_rng(seed)
    shp = (B, 1, H, W)
    d = dict(kx=rng.uniform(1, 6, (R, 1, H, W)).astype(np.float32))
    d["p0"] = (4700 + rng.uniform(-40, 40, shp)).astype(np.float32)
    d["p1"] = (d["p0"] - rng.uniform(-3 if small_dp else 1, 25, shp)).astype(np.float32)
    if small_dp:
        d["p1"][0, 0, 0, :2] = d["p0"][0, 0, 0, :2]
    d["sg0"] = rng.uniform(sg_lo, sg_hi, shp).astype(np.float32)
    d["sg1"] = (d["sg0"] - rng.uniform(0.001, 0.02, shp)).astype(np.float32)
    d["so0"] = (np.float32(0.78) - d["sg0"]).astype(np.float32)
    d["so1"] = (np.float32(0.78) - d["sg1"]).astype(np.float32)
    d["dt1"] = np.asarray([dts[(2 * b) % len(dts)] for b in range(B)], np.float32)      # exact means, see the DG generator
    d["dt2"] = np.asarray([dts[(2 * b + 1) % len(dts)] for b in range(B)], np.float32)
    d["t1"] = np.linspace(5, 50, B).astype(np.float32)
    d["sample_real"] = (np.arange(B) % R).astype(np.int32)
    tt = lambda a: torch.as_tensor(a)
    res = O.gc_residual(cfg, tab, tt(d["kx"]), tt(d["p0"]), tt(d["p1"]), tt(d["sg0"]), tt(d["sg1"]), tt(d["so0"]), tt(d["so1"]),
                        tt(d["dt1"]), tt(d["dt2"]), d["t1"], d["sample_real"])
    v0, dv0 = O.pvt_eval(tt(d["p0"]), tab, cfg, props=(0, 1, 4, 5), need_deriv=(0, 1, 4, 5))
    v1, _ = O.pvt_eval(tt(d["p1"]), tab, cfg, props=(0, 1, 2, 3, 4, 5))
    ch = lambda a: a.detach().reshape(B, H, W, 1)
    z = torch.zeros(B, H, W, 1)
    one = torch.ones(B, H, W, 1)
    field = lambda v: torch.as_tensor(np.broadcast_to(v.reshape(B, 1, 1, 1), (B, H, W, 1)).copy())
    dp0 = torch.stack([ch(dv0[0]), ch(dv0[1]), z, z, ch(dv0[4]), ch(dv0[5]), z])
    levels = [dict(p=ch(tt(d["p0"])), sg=ch(tt(d["sg0"])), so=ch(tt(d["so0"])), invBg=ch(v0[0]), invBo=ch(v0[1]), invug=one, invuo=one,
                   Rs=ch(v0[4]), Rv=ch(v0[5]), dpvt=dp0, dtf=field(d["dt1"]), q4=[z, z, z, z], pwf=z),
              dict(p=ch(tt(d["p1"])), sg=ch(tt(d["sg1"])), so=ch(tt(d["so1"])), invBg=ch(v1[0]), invBo=ch(v1[1]), invug=ch(v1[2]), invuo=ch(v1[3]),
                   Rs=ch(v1[4]), Rv=ch(v1[5]), dpvt=torch.zeros_like(dp0), dtf=field(d["dt2"]), q4=[ch(q) for q in res["q4"]], pwf=ch(res["pwf"]))]
    wells_o = cfg.wells
    cfd = {
        "Dimension": {"Gridblock_Dim": [cfg.dx, cfg.dy, cfg.dz], "Dim": [H, W, 1], "Measurement": [cfg.length, cfg.width, cfg.thickness]},
        "Conn_Idx": torch.tensor([[w.j, w.i, 0] for w in wells_o], dtype=torch.int32).reshape(-1, 3),
        "Init_Grate": torch.tensor([w.value for w in wells_o], dtype=torch.float32),
        "Min_BHP": torch.tensor([w.minimum_bhp for w in wells_o], dtype=torch.float32),
        "Completion_Ratio": 0.5, "SCAL": {"End_Points": {"Swmin": cfg.Swmin, "Sorg": cfg.Sorg}}, "Max_Train_Time": 365.0, "Pi": 5000.0,
        "Dew_Point": 4048.49, "Rhg_Std": 0.05, "Rho_Std": 50.0,
        "Init_InvBg": 1.0, "Init_DinvBg": 0.0, "Init_Invug": 1.0, "Init_InvBo": 1.0, "Init_Invuo": 1.0, "Init_Rs": 1.0, "Init_Rv": 1.0,
        "Kr_gas_oil": lambda sg: O.corey_krog_krgo_t(sg, cfg, torch.float32),
    }
    model = FakeModel(cfg, levels, cfd)
    sr = torch.as_tensor(d["sample_real"].astype(np.int64))
    x = [z.clone(), z.clone(), z.clone(), field(d["t1"]), torch.full((B, H, W, 1), float(np.float32(cfg.phi))),
         tt(d["kx"]).index_select(0, sr).reshape(B, H, W, 1)]
    ident = lambda model, v, stat_idx=None, compute=True: v if isinstance(v, torch.Tensor) else torch.tensor(float(v))
    ns = {"tf": tf, "nonormalize": ident, "normalize_diff": ident, "normalize": ident,
          "dnn": types.SimpleNamespace(conn_shutins_idx=lambda t, ci, days: torch.zeros_like(t)),
          "time_shifting": lambda model, x, **k: (x, 1.0, torch.tensor(1e30))}
    exec(reference_function("physics_error_gas_oil_2D"), ns)
    errs, outs, checks, blks = ns["physics_error_gas_oil_2D"](model, x, None)
    dom, ibc, mbc, cmbc = errs[0], errs[3], checks[0], checks[1]
    assert model.calls == 2
    back = lambda a: a.detach().reshape(B, 1, H, W).numpy()
    out = {k: v for k, v in d.items()}
    out.update(W=W, H=H, B=B, R=R, wells=np.asarray([[w["i"], w["j"], w["k"], w["value"]] for w in wl], np.float32).reshape(-1, 4),
               ref_dom=back(dom), ref_ibc=back(ibc), ref_mbc=mbc.detach().reshape(B).numpy(), ref_cmbc=back(cmbc))
    o = dict(dom=res["dom"].detach().numpy(), ibc=res["ibc"].detach().numpy(), mbc=res["mbc"].detach().numpy(), cmbc=res["cmbc"].detach().numpy())
    return out, o


def ulp(a, b):
    ai = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    bi = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -2**31 - ai, ai)
    bi = np.where(bi < 0, -2**31 - bi, bi)
    return int(np.abs(ai - bi).max())


def main():
    out = {}
    cases = {"a": dict(seed=5301, B=3, H=9, W=8, wells="two"),
             "b": dict(seed=5302, B=4, H=7, W=10, wells="dup", R=2, sg_lo=0.2, sg_hi=0.35),
             "c": dict(seed=5303, B=2, H=6, W=7, wells="none", small_dp=True, sg_lo=0.2, sg_hi=0.5)}
    for name, kw in cases.items():
        r, o = run_case(**kw)
        print(name, "ulp distance reference fragment vs oracle: dom", ulp(r["ref_dom"], o["dom"]), "ibc", ulp(r["ref_ibc"], o["ibc"]),
              "cmbc", ulp(r["ref_cmbc"], o["cmbc"]), " mbc rel %.2e" % (np.abs(r["ref_mbc"] - o["mbc"]).max() / max(np.abs(o["mbc"]).max(), 1e-30)))
        for k, v in r.items():
            out[f"{name}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "reference_gc_residual.npz"), **out)
    print("wrote reference_gc_residual.npz")


if __name__ == "__main__":
    main()
