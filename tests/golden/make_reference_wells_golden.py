"""Golden vectors for the well source terms made by the REFERENCE'S OWN code, executed through the torch-backed
TensorFlow stand-in of this directory (TensorFlow is not installable here):

  * WellDataProcessor.scatter_y and .conn_shutins_idx (welldata_processor.py:170-224, 228-389), cut out by AST: the
    dense well masks / target fields and the shut-in identity (integer work);
  * WellRatesPressure.compute_rates_and_bhp with _non_iterative_method, _compute_phase_rates,
    compute_blocking_integral_and_factor, _split_condensate_components, extract_pvt_properties
    (well_rate_bhp_Subclassed.py:198-233, 614-724, 727-1034), cut out by AST and run on a bare instance whose attributes
    are set the way __init__ sets them (the constructor itself builds Keras models and reads files).
    model_PVT is the oracle's spline (pinned by make_reference_pvt_golden.py), relperm_model the oracle's Corey function,
    the de-normalisation of t and kx is the identity (physical inputs).

Cases: dry gas without and with the blocking-factor integral (n_intervals = 8), gas condensate without it and with it
(Newton and the bracketing solver, 20 root iterations per trapezoid node).
Output: tests/golden/reference_wells.npz (inputs + the dense rate / BHP fields the reference code returns).

Iterative BHP control (use_non_iterative=False: _iterative_method, well_rate_bhp_Subclassed.py:515-612, Newton-Raphson
on the bottom-hole pressure inside a tf.while_loop whose stopping test couples the batch): the same cases through the
reference's loop, plus the gradient of the summed rates w.r.t. the pressure (and gas saturation) field as autodiff
delivers it THROUGH the loop.  Output: tests/golden/reference_wells_iter.npz.
"""
import ast
import logging
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

REF = "/root/reference"


def cut_methods(path, cls, names):
    src = open(path).read()
    node = next(n for n in ast.parse(src).body if isinstance(n, ast.ClassDef) and n.name == cls)
    out = {}
    for fn in node.body:
        if isinstance(fn, ast.FunctionDef) and fn.name in names:
            fn.decorator_list = []
            out[fn.name] = textwrap.dedent(ast.get_source_segment(src, fn))
    missing = set(names) - set(out)
    assert not missing, missing
    return out


def strip_decorators(code):
    return "\n".join(l for l in code.splitlines() if not l.strip().startswith("@tf.function"))


def build_class(path, cls, names, ns):
    body = "\n\n".join(textwrap.indent(strip_decorators(c), "    ") for c in cut_methods(path, cls, names).values())
    exec(f"class {cls}:\n{body}\n", ns)
    return ns[cls]


def run_case(fluid, blocking, seed, B=4, D=2, H=7, W=9, solver="newton", iterative=False, max_iters=10):
    rng = np.random.default_rng(seed)
    wl = [O.Well(i=2, j=2, k=0, value=500.0), O.Well(i=W - 2, j=H - 3, k=D - 1, value=1000.0, shutin_days=(20.0, 35.0)),
          O.Well(i=4, j=1, k=0, value=2.0e5)]                    # the third target is BHP limited
    cfg = O.OracleConfig(D=D, H=H, W=W, wells=wl, use_blocking_factor=blocking, n_intervals=8, use_non_iterative=not iterative, bhp_max_iters=max_iters)
    cols = O.load_pvt_table(os.path.join(HERE, "pvt_table.npz"))
    props = O.GC_PROPS if fluid == "GC" else O.DG_PROPS
    tab = O.build_spline_table(cols, props, order=1, lam=0.001)
    shp = (B, D, H, W)
    kx = rng.uniform(1, 6, shp).astype(np.float32)
    p = (4700 + rng.uniform(-300, 250, shp)).astype(np.float32)
    sg = rng.uniform(0.2, 0.75, shp).astype(np.float32)
    t_days = np.asarray([5.0, 20.0, 30.0, 50.0][:B], np.float32)
    ns = {"tf": tf, "np": np, "logging": logging, "os": os}
    WDP = build_class(os.path.join(REF, "welldata_processor.py"), "WellDataProcessor", ["scatter_y", "conn_shutins_idx"], ns)
    wdp = WDP.__new__(WDP)
    wdp.dtype = tf.float32
    conn = [(w.k, w.j, w.i) for w in wl]                         # welldata_processor.py:26-40: [k, j, i] rows
    shape5 = (1, D, H, W, 1)
    ns2 = {"tf": tf, "np": np, "logging": logging, "os": os, "project_directory": "/tmp",
           "slice_tensor": lambda x, idx, dim=-1: x[..., idx[0]:idx[0] + 1]}
    WRP = build_class(os.path.join(REF, "well_rate_bhp_Subclassed.py"), "WellRatesPressure",
                      ["compute_rates_and_bhp", "_non_iterative_method", "_compute_phase_rates", "compute_blocking_integral_and_factor",
                       "_split_condensate_components", "extract_pvt_properties", "_solve_newton", "_solve_chandrupatla",
                       "_iterative_method"], ns2)
    w = WRP.__new__(WRP)
    w.fluid_type, w.use_blocking_factor, w.dtype, w.solver, w.n_intervals, w.n_root_iter = fluid, blocking, tf.float32, solver, 8, 20
    w.max_iters, w.tol, w.use_non_iterative, w.compute_mo = max_iters, 1e-6, not iterative, fluid == "GC"
    w.kx_ky = tf.constant(cfg.kx_ky, dtype=tf.float32)
    w.dx = tf.constant(cfg.length, dtype=tf.float32) / W          # :113-115
    w.dy = tf.constant(cfg.width, dtype=tf.float32) / H
    w.dz = tf.constant(cfg.thickness, dtype=tf.float32) / D
    w.C = cfg.C
    w.well_data_processor = wdp
    w.well_data = {"connection_index": conn, "shutin_days": [[list(x.shutin_days)] for x in wl]}
    w.well_id = wdp.scatter_y(shape5, conn, 1.0)                                                    # :128-132
    w.rw = wdp.scatter_y(shape5, conn, [x.wellbore_radius for x in wl])
    w.q0 = wdp.scatter_y(shape5, conn, [x.value for x in wl])
    w.pwf_min = wdp.scatter_y(shape5, conn, [x.minimum_bhp for x in wl])
    w.completion_ratio = wdp.scatter_y(shape5, conn, [x.completion_ratio for x in wl])
    w.scal_config = {"end_points": {"Swmin": cfg.Swmin}}
    w.relperm = types.SimpleNamespace(end_points={"Swmin": cfg.Swmin})
    w.norm_config = None
    w.data_summary = types.SimpleNamespace(get_key_index=lambda k: {"time": 3, "permx": 4}[k],
                                           nonormalize=lambda v, **kw: v)
    w.log_tensor_to_file = lambda *a, **k: None

    def model_PVT(pp):
        v, dv = O.pvt_eval(pp, tab, cfg, props=tuple(range(len(props))))
        return torch.stack([torch.stack([v[i] for i in range(len(props))]), torch.zeros(len(props), *pp.shape)])

    relperm = lambda s: O.corey_krog_krgo_t(s if isinstance(s, torch.Tensor) else torch.tensor(np.float32(s)), cfg, torch.float32)
    x = torch.zeros(B, D, H, W, 5)
    x[..., 3] = torch.as_tensor(t_days).view(B, 1, 1, 1)
    x[..., 4] = torch.as_tensor(kx)
    p5, sg5 = torch.as_tensor(p).unsqueeze(-1), torch.as_tensor(sg).unsqueeze(-1)
    if iterative:
        p5.requires_grad_(True)
        sg5.requires_grad_(True)
    # dry gas: the gas saturation is Sgi = 1 - Swmin everywhere (physics_loss.py:65,129); passed as a field because the
    # reference's blocking-factor branch calls ref.get_shape() on it (a python float, its own default, has none)
    sg_dg = torch.full_like(p5, float(np.float32(1.0 - cfg.Swmin)))
    rates, pwf = w.compute_rates_and_bhp(x, p5, sg5 if fluid == "GC" else sg_dg, relperm, model_PVT)
    shut = wdp.conn_shutins_idx(x[..., 3:4], conn, w.well_data["shutin_days"], time_axis=0)
    out = dict(D=D, H=H, W=W, B=B, kx=kx, p=p, sg=sg, t_days=t_days,
               wells=np.asarray([[x_.i, x_.j, x_.k, x_.value, x_.shutin_days[0], x_.shutin_days[1]] for x_ in wl], np.float32),
               well_id=w.well_id.numpy(), q0=w.q0.numpy(), pwf_min=w.pwf_min.numpy(), shut=shut.numpy().astype(np.int32),
               pwf=pwf[..., 0].detach().numpy(), max_iters=max_iters)
    if fluid == "GC":
        out["q4"] = np.stack([r[..., 0].detach().numpy() for r in rates])
    else:
        out["q"] = rates[..., 0].detach().numpy()
    # the oracle on the same inputs, at the connection cells
    flat = O.well_flat_index(wl, D, H, W).astype(np.int64)
    pc = torch.as_tensor(p).reshape(B, -1)[:, flat]
    kc = torch.as_tensor(kx).reshape(B, -1)[:, flat]
    if iterative:      # d(sum of all rates)/d(p, Sg) through the loop: the reference graph's and the oracle's
        pc.requires_grad_(True)
        tot = sum(r.sum() for r in rates) if fluid == "GC" else rates.sum()
        gr = torch.autograd.grad(tot, [p5, sg5] if fluid == "GC" else [p5], allow_unused=True)
        out["dq_dp"] = gr[0][..., 0].numpy()
        if fluid == "GC":
            out["dq_dsg"] = gr[1][..., 0].numpy()
    if fluid == "GC":
        sc = torch.as_tensor(sg).reshape(B, -1)[:, flat]
        if iterative:
            sc.requires_grad_(True)
        q4o, pwo = O.wells_gc(pc, sc, kc, t_days, tab, cfg, torch.float32, solver=solver)
        ref = [out["q4"][c].reshape(B, -1)[:, flat] for c in range(4)]
        d = max(ulp(q4o[c].detach().numpy(), ref[c]) for c in range(4))
    else:
        qo, pwo = O.wells_dg(pc, kc, t_days, tab, cfg, torch.float32)
        d = ulp(qo.detach().numpy(), out["q"].reshape(B, -1)[:, flat])
    dp = ulp(pwo.detach().numpy(), out["pwf"].reshape(B, -1)[:, flat])
    print(fluid, "blocking" if blocking else "plain", "iterative" if iterative else "", ": max ulp distance reference vs oracle at the connections: rates", d, " pwf", dp)
    if iterative:
        toto = sum(x_.sum() for x_ in q4o) if fluid == "GC" else qo.sum()
        go = torch.autograd.grad(toto, [pc, sc] if fluid == "GC" else [pc], allow_unused=True)
        a, b_ = go[0].numpy(), out["dq_dp"].reshape(B, -1)[:, flat]
        print("   d(sum q)/dp   oracle vs reference graph: max |diff| / max |ref| = %.2e" % (np.abs(a - b_).max() / max(np.abs(b_).max(), 1e-30)))
        if fluid == "GC":
            a, b_ = go[1].numpy(), out["dq_dsg"].reshape(B, -1)[:, flat]
            print("   d(sum q)/dSg  oracle vs reference graph: max |diff| / max |ref| = %.2e" % (np.abs(a - b_).max() / max(np.abs(b_).max(), 1e-30)))
    return out


def ulp(a, b):
    ai = np.asarray(a, np.float32).view(np.int32).astype(np.int64)
    bi = np.asarray(b, np.float32).view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, -2**31 - ai, ai)
    bi = np.where(bi < 0, -2**31 - bi, bi)
    return int(np.abs(ai - bi).max())


def main():
    out = {}
    for name, kw in {"dg": dict(fluid="DG", blocking=False, seed=5401), "dgblk": dict(fluid="DG", blocking=True, seed=5402),
                     "gc": dict(fluid="GC", blocking=False, seed=5403),
                     # the GC blocking-factor integral with its root finders (well_rate_bhp_Subclassed.py:857-950, 236-324)
                     "gcblk": dict(fluid="GC", blocking=True, seed=5404),
                     "gcblk_br": dict(fluid="GC", blocking=True, seed=5405, solver="chandrupatla")}.items():
        for k, v in run_case(**kw).items():
            out[f"{name}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "reference_wells.npz"), **out)
    print("wrote reference_wells.npz")
    out = {}
    for name, kw in {"dg": dict(fluid="DG", blocking=False, seed=5411), "dgblk": dict(fluid="DG", blocking=True, seed=5412),
                     "gc": dict(fluid="GC", blocking=False, seed=5413),
                     # three Newton steps on the BHP: each evaluates two trapezoid integrals of eight 20-step root finds
                     "gcblk": dict(fluid="GC", blocking=True, seed=5414, B=2, max_iters=3)}.items():
        for k, v in run_case(iterative=True, **kw).items():
            out[f"{name}_{k}"] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, "reference_wells_iter.npz"), **out)
    print("wrote reference_wells_iter.npz")


if __name__ == "__main__":
    main()
