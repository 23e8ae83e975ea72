"""Golden vectors for the dry-gas residual made by the REFERENCE'S OWN code.

`physics_error_gas_2D` (physics_loss.py:9-224: the static part and the nested `physics_error_gas`) is cut out of
/root/reference/physics_loss.py by AST and executed here on seeded inputs, with TensorFlow replaced by the torch-backed
shim of this directory (TensorFlow is not installable in the build container; the shim implements the ~20 ops the
fragment uses and nothing else).  What the fragment expects from its surroundings is supplied by stand-ins:

  * `model(x, training=True)` -- the Keras pipeline -- returns, per time level, the pressure, invBg, invug, d(invBg)/dp,
    the time-step field and the well-rate field.  The stand-in returns GIVEN fields (seeded pressures; PVT values and
    well rates evaluated by the oracle at those pressures), so the fragment's own arithmetic -- face permeabilities,
    face averages, flux assembly, accumulation, truncation term, inner-boundary term, material balance -- is what is
    recorded.  PVT and wells are pinned separately (make_reference_pvt_golden.py, tests/test_oracle.py).
  * `nonormalize` / `normalize_diff` (not shipped with the fragment): identity -- physical inputs are passed.
  * `dnn.conn_shutins_idx`, `time_shifting`: results unused by the dry-gas residual; zeros / no shift.

Output: tests/golden/reference_dg_residual.npz (inputs + the fragment's dom, ibc, mbc).

    python tests/golden/make_reference_dg_golden.py      (build container: /root/reference must exist)
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
    tree = ast.parse(src)
    fn = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == name)
    return textwrap.dedent(ast.get_source_segment(src, fn))


class FakeModel:
    """stand-in for the Keras pipeline: returns the given fields, level n on the first call, n+1 on the second"""

    def __init__(self, cfg, levels, cfd_type):
        self.dtype = tf.float32
        self.cfd_type = cfd_type
        self.levels = levels
        self.calls = 0
        self.cf = float(O.rock_compressibility(cfg.phi))         # used as model.cf at physics_loss.py:149

    def __call__(self, x, training=True):
        lv = self.levels[min(self.calls, 1)]
        self.calls += 1
        one = torch.ones_like(lv["p"])
        return [lv["p"], one, lv["invBg"], lv["invug"], torch.stack([lv["dinvBg"], one]), lv["dtf"], one, lv["q"], one]


def run_case(W, H, B, R, seed, dts):
    rng = np.random.default_rng(seed)
    cfg = O.OracleConfig(D=1, H=H, W=W, wells=O.default_wells(W, H, 1))
    cols = O.load_pvt_table(os.path.join(HERE, "pvt_table.npz"))
    tab = O.build_spline_table(cols, O.DG_PROPS, order=1, lam=0.001)
    kx = np.exp(rng.normal(np.log(3.0), 0.45, size=(R, 1, H, W))).astype(np.float32)
    sample_real = (np.arange(B) % R).astype(np.int32)
    p0 = (5000.0 - 300.0 * rng.random((B, 1, H, W)) - 2.0 * rng.standard_normal((B, 1, H, W))).astype(np.float32)
    p1 = (p0 - 30.0 * rng.random((B, 1, H, W))).astype(np.float32)
    # time steps with few mantissa bits: the fragment takes the MEAN of a per-sample constant field, which is then exact
    dt1 = np.asarray([dts[(2 * b) % len(dts)] for b in range(B)], np.float32)
    dt2 = np.asarray([dts[(2 * b + 1) % len(dts)] for b in range(B)], np.float32)
    t_days = np.linspace(10.0, 300.0, B).astype(np.float32)
    tt = lambda a: torch.as_tensor(a)
    res = O.dg_residual(cfg, tab, tt(kx), tt(p0), tt(p1), tt(dt1), tt(dt2), t_days, sample_real)
    ch = lambda a: a.reshape(B, H, W, 1)                         # (B,1,H,W) -> the fragment's (B,H,W,1)
    A0, A0p, A1, M1, q = (res[k].detach() for k in ("A0", "A0p", "A1", "M1", "q"))
    field = lambda v: ch(torch.as_tensor(np.broadcast_to(v.reshape(B, 1, 1, 1), (B, 1, H, W)).copy()))
    levels = [dict(p=ch(tt(p0)), invBg=ch(A0), invug=torch.ones(B, H, W, 1), dinvBg=ch(A0p), dtf=field(dt1), q=torch.zeros(B, H, W, 1)),
              dict(p=ch(tt(p1)), invBg=ch(A1), invug=ch(M1), dinvBg=torch.zeros(B, H, W, 1), dtf=field(dt2), q=ch(q))]
    _, krg = O.corey_krog_krgo_np(np.float32(1.0 - cfg.Swmin), cfg, np.float32)
    wells = cfg.wells
    cfd = {
        "Dimension": {"Gridblock_Dim": [cfg.dx, cfg.dy, cfg.dz], "Dim": [H, W, 1], "Measurement": [cfg.length, cfg.width, cfg.thickness]},
        "Conn_Idx": torch.tensor([[w.j, w.i, 0] for w in wells], dtype=torch.int32),
        "Init_Grate": torch.tensor([w.value for w in wells], dtype=torch.float32),
        "Min_BHP": torch.tensor([w.minimum_bhp for w in wells], dtype=torch.float32),
        "Completion_Ratio": 0.5, "SCAL": {"End_Points": {"Swmin": cfg.Swmin}}, "Max_Train_Time": 365.0, "Pi": 5000.0,
        "Init_InvBg": 1.0, "Init_DinvBg": 0.0, "Init_Invug": 1.0,
        "Kr_gas_oil": lambda sg: (torch.tensor(0.0), torch.tensor(float(krg), dtype=torch.float32)),
        "Connection_Shutins": {"Days": [], "Shutins_Idx": [], "Shutins_Per_Conn_Idx": []},
    }
    model = FakeModel(cfg, levels, cfd)
    sr = torch.as_tensor(sample_real.astype(np.int64))
    x = [torch.zeros(B, H, W, 1), torch.zeros(B, H, W, 1), torch.zeros(B, H, W, 1),
         field(t_days), torch.full((B, H, W, 1), float(np.float32(cfg.phi))), ch(tt(kx).index_select(0, sr))]
    ns = {
        "tf": tf,
        "nonormalize": lambda model, v, stat_idx=None, compute=True: v,
        "normalize_diff": lambda model, v, stat_idx=None, compute=True: v,
        "dnn": types.SimpleNamespace(conn_shutins_idx=lambda t, ci, days: torch.zeros_like(t)),
        "time_shifting": lambda model, x, **k: (x, 1.0, torch.tensor(1e30)),
    }
    exec(reference_function("physics_error_gas_2D"), ns)
    errs, outs, checks, blks = ns["physics_error_gas_2D"](model, x, None)
    dom, ibc, mbc = errs[0], errs[3], checks[0]
    assert model.calls == 2
    back = lambda a: a.reshape(B, 1, H, W).numpy()
    return dict(W=W, H=H, B=B, R=R, kx=kx, sample_real=sample_real, p0=p0, p1=p1, dt1=dt1, dt2=dt2, t_days=t_days,
                ref_dom=back(dom), ref_ibc=back(ibc), ref_mbc=mbc.reshape(B).numpy(),
                oracle_dom=res["dom"].detach().numpy(), oracle_ibc=res["ibc"].detach().numpy(), oracle_mbc=res["mbc"].detach().numpy())


def main():
    out = {}
    for name, kw in {"a": dict(W=12, H=9, B=4, R=2, seed=5100, dts=[0.5, 2.25, 7.125, 1.0, 0.375, 9.5]),
                     "b": dict(W=39, H=39, B=3, R=3, seed=5101, dts=[4.0, 0.75, 1.5, 6.25])}.items():
        r = run_case(**kw)
        d = np.abs(r["ref_dom"].view(np.int32).astype(np.int64) - r["oracle_dom"].view(np.int32).astype(np.int64)).max()
        print(name, "max ulp distance reference fragment vs oracle: dom", d,
              " ibc equal:", np.array_equal(r["ref_ibc"], r["oracle_ibc"]),
              " mbc rel:", np.abs(r["ref_mbc"] - r["oracle_mbc"]).max() / np.abs(r["oracle_mbc"]).max())
        for k, v in r.items():
            if not k.startswith("oracle_"):
                out[f"{name}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "reference_dg_residual.npz"), **out)
    print("wrote reference_dg_residual.npz")


if __name__ == "__main__":
    main()
