"""Golden vectors for the loss assembly made by the REFERENCE'S OWN pinn_batch_sse_grad (physics_loss.py:742-870), executed
through the torch-backed TensorFlow stand-in with the reference's own physics_error_gas_2D as model.loss_func
['Physics_Error'] (same stand-ins as make_reference_dg_golden.py): squared errors, SSE per term, the weights nwt, the
error counts (mbc counted with the ic field's shape, :830) and the reported MSE.  `zeros_to_ones` is not shipped with the
fragment: where(c == 0, 1, c).   Output: tests/golden/reference_loss.npz
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import tf_torch_shim as tf                                   # noqa: E402
import make_reference_dg_golden as DG                        # noqa: E402


def main():
    out = {}
    captured = {}
    real_exec = DG.reference_function

    # run the dry-gas case "a" of make_reference_dg_golden.py, but through pinn_batch_sse_grad
    import srm_oracle as O
    W, H, B, R, seed, dts = 12, 9, 4, 2, 5100, [0.5, 2.25, 7.125, 1.0, 0.375, 9.5]
    # rebuild the same model / inputs by calling run_case's internals: simplest is to re-run it and capture the pieces
    orig_call = DG.FakeModel.__call__
    r = DG.run_case(W=W, H=H, B=B, R=R, seed=seed, dts=dts)
    # second pass: same construction, now driving the loss function
    rng = np.random.default_rng(seed)
    cfg = O.OracleConfig(D=1, H=H, W=W, wells=O.default_wells(W, H, 1))
    cols = O.load_pvt_table(os.path.join(HERE, "pvt_table.npz"))
    tab = O.build_spline_table(cols, O.DG_PROPS, order=1, lam=0.001)
    tt = torch.as_tensor
    res = O.dg_residual(cfg, tab, tt(r["kx"]), tt(r["p0"]), tt(r["p1"]), tt(r["dt1"]), tt(r["dt2"]), r["t_days"], r["sample_real"])
    ch = lambda a: a.reshape(B, H, W, 1)
    field = lambda v: ch(torch.as_tensor(np.broadcast_to(v.reshape(B, 1, 1, 1), (B, 1, H, W)).copy()))
    A0, A0p, A1, M1, q = (res[k].detach() for k in ("A0", "A0p", "A1", "M1", "q"))
    levels = [dict(p=ch(tt(r["p0"])), invBg=ch(A0), invug=torch.ones(B, H, W, 1), dinvBg=ch(A0p), dtf=field(r["dt1"]), q=torch.zeros(B, H, W, 1)),
              dict(p=ch(tt(r["p1"])), invBg=ch(A1), invug=ch(M1), dinvBg=torch.zeros(B, H, W, 1), dtf=field(r["dt2"]), q=ch(q))]
    _, krg = O.corey_krog_krgo_np(np.float32(1.0 - cfg.Swmin), cfg, np.float32)
    wells = cfg.wells
    cfd = {"Dimension": {"Gridblock_Dim": [cfg.dx, cfg.dy, cfg.dz], "Dim": [H, W, 1], "Measurement": [cfg.length, cfg.width, cfg.thickness]},
           "Conn_Idx": torch.tensor([[w.j, w.i, 0] for w in wells], dtype=torch.int32),
           "Init_Grate": torch.tensor([w.value for w in wells], dtype=torch.float32),
           "Min_BHP": torch.tensor([w.minimum_bhp for w in wells], dtype=torch.float32),
           "Completion_Ratio": 0.5, "SCAL": {"End_Points": {"Swmin": cfg.Swmin}}, "Max_Train_Time": 365.0, "Pi": 5000.0,
           "Init_InvBg": 1.0, "Init_DinvBg": 0.0, "Init_Invug": 1.0,
           "Kr_gas_oil": lambda sg: (torch.tensor(0.0), torch.tensor(float(krg), dtype=torch.float32)),
           "Connection_Shutins": {"Days": [], "Shutins_Idx": [], "Shutins_Per_Conn_Idx": []}}
    model = DG.FakeModel(cfg, levels, cfd)
    nwt = [1.0, 0.5, 0.25, 2.0, 1.5, 3.0, 0.75, 0.0]
    model.nwt = torch.tensor(nwt, dtype=torch.float32)
    model.nT, model.nT_list, model.trainable_variables = 1, [0], []
    ns = {"tf": tf, "nonormalize": lambda m, v, stat_idx=None, compute=True: v, "normalize_diff": lambda m, v, stat_idx=None, compute=True: v,
          "dnn": types.SimpleNamespace(conn_shutins_idx=lambda t, ci, days: torch.zeros_like(t)),
          "time_shifting": lambda m, x, **k: (x, 1.0, torch.tensor(1e30)),
          "zeros_to_ones": lambda c: torch.where(c == 0, torch.ones_like(c), c)}
    exec(DG.reference_function("physics_error_gas_2D"), ns)
    exec(DG.reference_function("pinn_batch_sse_grad"), ns)
    model.loss_func = {"Physics_Error": ns["physics_error_gas_2D"], "Reshape": lambda y: y, "Reduce_Axis": [1, 2, 3, 4], "Squeeze_Out": lambda y: y}
    sr = torch.as_tensor(r["sample_real"].astype(np.int64))
    x = [torch.zeros(B, H, W, 1), torch.zeros(B, H, W, 1), torch.zeros(B, H, W, 1), field(r["t_days"]),
         torch.full((B, H, W, 1), float(np.float32(cfg.phi))), ch(tt(r["kx"]).index_select(0, sr))]
    y = [torch.zeros(B, H, W, 1)]
    wsse, wsse_grad, count, wmse, y_model = ns["pinn_batch_sse_grad"](model, x, y)
    f = lambda v: float(v) if not isinstance(v, torch.Tensor) or v.numel() == 1 else v.detach().numpy()
    out.update({k: r[k] for k in ("W", "H", "B", "R", "kx", "sample_real", "p0", "p1", "dt1", "dt2", "t_days")})
    out["nwt"] = np.asarray(nwt, np.float32)
    out["wsse"] = np.asarray([f(v) for v in wsse[:8]], np.float64)           # batch, dom, dbc, nbc, ibc, ic, mbc, cmbc
    out["count"] = np.asarray([f(v) for v in count[:8]], np.float64)
    out["wmse"] = np.asarray([f(v) for v in wmse[:8]], np.float64)
    np.savez_compressed(os.path.join(HERE, "reference_loss.npz"), **out)
    print("wrote reference_loss.npz  wsse", out["wsse"], "count", out["count"])


if __name__ == "__main__":
    main()
