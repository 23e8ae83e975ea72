"""Shared helpers for the tests: build the oracle configuration and the product spec from ONE
description, generate seeded synthetic inputs, compare fields."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import srm_b200 as srm  # noqa: E402
import srm_oracle as O  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
WEIGHTS = [1.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0]   # dom, ibc, mbc, tde (default_configurations.py:63-83)


def make_case(W, H, D, T, K, seed, wells="default", all_layers=False, use_blocking_factor=False, n_intervals=8,
              near_knots=True, order=1):
    """returns (oracle_cfg, oracle_table, spec, tables, batch)"""
    if wells == "default":
        wl = srm.config.scaled_default_wells(W, H, D, all_layers=all_layers)
    elif wells == "lattice":
        wl = srm.config.lattice_wells(W, H, D)
    elif wells == "none":
        wl = []
    else:
        wl = wells
    spec = srm.PhysicsSpec(D=D, H=H, W=W, wells=wl, use_blocking_factor=use_blocking_factor, n_intervals=n_intervals)
    cols = O.load_pvt_table(os.path.join(GOLDEN, "pvt_table.npz"))
    otab = O.build_spline_table(cols, O.DG_PROPS, order=order, lam=0.001)
    ptab = srm.build_spline_tables(srm.load_default_pvt_table(), srm.pvt.DG_PROPERTIES, order=order)
    ocfg = O.OracleConfig(
        D=D, H=H, W=W, use_blocking_factor=use_blocking_factor, n_intervals=n_intervals,
        wells=[O.Well(i=w.i, j=w.j, k=w.k, value=abs(w.q_target), producer=(w.q_target >= 0 and not (w.q_target == 0 and np.signbit(w.q_target))),
                      minimum_bhp=w.pwf_min, wellbore_radius=w.rw, completion_ratio=w.hc,
                      shutin_days=(w.shut_start, w.shut_stop)) for w in wl])
    batch = srm.synth.make_batch(W, H, D, T, K, [(w.i, w.j) for w in wl[:8]], seed=seed,
                                 near_knots=torch.as_tensor(otab.c) if near_knots else None)
    return ocfg, otab, spec, ptab, batch


def crowded_wells(D, n_cols, j=1, i0=2, step=2, duplicate=True, minimum_bhp=4100.0):
    """n_cols producer columns next to each other in grid row j, completed in every layer, plus (duplicate) a second
    connection in the first column's top cell: more columns than a tile's staged lists hold (well_tile.cuh) send the
    kernels to their search path, and the duplicate exercises the summed scatter (welldata_processor.py:170-224)"""
    conns = []
    for t in range(n_cols):
        for k in range(D):
            conns.append({"name": f"C{t}", "i": i0 + step * t, "j": j, "k": k, "type": "producer", "control": "ORAT",
                          "value": 400.0 + 50.0 * (t % 3), "minimum_bhp": minimum_bhp, "wellbore_radius": 0.09525,
                          "completion_ratio": 0.5, "shutin_days": [[1000.0, 0.0]]})
    if duplicate:
        conns.append(dict(conns[0], name="Cdup", value=250.0))
    return srm.config.wells_from_connections(conns)


def oracle_run(ocfg, otab, batch, weights=WEIGHTS, dtype=torch.float32):
    return O.dg_forward_backward(ocfg, otab, batch.kx.numpy(), batch.p0.numpy(), batch.p1.numpy(), batch.dt1.numpy(),
                                 batch.dt2.numpy(), batch.t1.numpy(), batch.sample_real.numpy(), weights, dtype=dtype)


def to_dev(batch, dev):
    c = lambda t: t.to(dev).contiguous()
    return dict(kx=c(batch.kx), sample_real=c(batch.sample_real), p0=c(batch.p0), p1=c(batch.p1), dt1=c(batch.dt1),
                dt2=c(batch.dt2), t1=c(batch.t1))


def cuda_run(spec, ptab, batch, weights=WEIGHTS, numerics="reference", want_dom=True, pvt_lut=False, lut_range=None):
    dev = torch.device("cuda", 0)
    eng = srm.SrmPhysics(spec, ptab, device=0, numerics=numerics, pvt_lut=pvt_lut, lut_range=lut_range)
    d = to_dev(batch, dev)
    fw = eng.forward(want_dom=want_dom, want_wells=True, **d)
    dterms = torch.tensor(weights, dtype=torch.float32, device=dev)
    gp0, gp1, gdt1, gdt2 = eng.backward(dterms=dterms, **d)
    torch.cuda.synchronize()
    out = dict(terms=fw["terms"][0].cpu().numpy(), counts=fw["terms"][1].cpu().numpy(),
               dom=fw["dom"].cpu().numpy() if want_dom else None, qw=fw["qw"].cpu().numpy(),
               pwfw=fw["pwfw"].cpu().numpy(), gp0=gp0.cpu().numpy(), gp1=gp1.cpu().numpy(),
               gdt1=gdt1.cpu().numpy(), gdt2=gdt2.cpu().numpy())
    eng.close()
    return out


def rel_to_max(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def ulp_diff(a, b):
    """max distance in units of fp32 ulps (of b)"""
    a = np.asarray(a, dtype=np.float32)
    b = np.asarray(b, dtype=np.float32)
    ai = a.view(np.int32).astype(np.int64)
    bi = b.view(np.int32).astype(np.int64)
    ai = np.where(ai < 0, np.int64(-2**31) - ai, ai)
    bi = np.where(bi < 0, np.int64(-2**31) - bi, bi)
    return int(np.abs(ai - bi).max())


def gc_case(seed, B=2, D=2, H=5, W=6, sg_lo=0.2, sg_hi=0.75, wells="two", R=1, small_dp=False, nog=None, ng=None):
    cols = O.load_pvt_table(os.path.join(GOLDEN, "pvt_table.npz"))
    otab = O.build_spline_table(cols, O.GC_PROPS, order=1, lam=0.001)
    if wells == "two":
        wl = [dict(i=2, j=2, k=0, value=500.0), dict(i=W - 2, j=H - 2, k=D - 1, value=1000.0)]
    elif wells == "dup":     # two connections in one cell + a neighbouring well cell
        wl = [dict(i=2, j=2, k=0, value=500.0), dict(i=2, j=2, k=0, value=300.0), dict(i=3, j=2, k=0, value=800.0)]
    elif isinstance(wells, tuple) and wells[0] == "columns":     # n well columns side by side, completed in every layer, one duplicate
        wl = [dict(i=1 + 2 * t, j=1, k=k, value=300.0 + 100.0 * (t % 3)) for t in range(wells[1]) for k in range(D)]
        wl.append(dict(i=1, j=1, k=D - 1, value=250.0))
    else:
        wl = []
    ocfg = O.OracleConfig(D=D, H=H, W=W, wells=[O.Well(**w) for w in wl])
    conns = [dict(i=w["i"], j=w["j"], k=w["k"], type="producer", control="ORAT", value=w["value"], minimum_bhp=4100.0,
                  wellbore_radius=0.09525, completion_ratio=0.5, shutin_days=[[1000.0, 0.0]]) for w in wl]
    spec = srm.PhysicsSpec(D=D, H=H, W=W, wells=srm.config.wells_from_connections(conns), fluid_type="GC")
    if nog is not None:          # other Corey exponents than the reference's defaults (3, 6): the kernels' run-time exponent path
        ocfg.nog, ocfg.ng = float(nog), float(ng)
        spec.corey_exponents = {"nog": float(nog), "ng": float(ng)}
    ptab = srm.pvt.SplineTables(knots=otab.c, w=otab.w, v=otab.v, order=1, properties=srm.pvt.GC_PROPERTIES)
    rng = np.random.default_rng(seed)
    shp = (B, D, H, W)
    d = dict(kx=rng.uniform(1, 6, (R, D, H, W)).astype(np.float32))
    d["p0"] = (4700 + rng.uniform(-40, 40, shp)).astype(np.float32)
    d["p1"] = (d["p0"] - rng.uniform(-3 if small_dp else 1, 25, shp)).astype(np.float32)
    if small_dp:
        d["p1"][0, 0, 0, :2] = d["p0"][0, 0, 0, :2]            # p1 == p0: divide_no_nan branch
    d["sg0"] = rng.uniform(sg_lo, sg_hi, shp).astype(np.float32)
    d["sg1"] = (d["sg0"] - rng.uniform(0.001, 0.02, shp)).astype(np.float32)
    d["so0"] = (np.float32(0.78) - d["sg0"]).astype(np.float32)
    d["so1"] = (np.float32(0.78) - d["sg1"]).astype(np.float32)
    d["dt1"] = rng.uniform(1, 6, B).astype(np.float32)
    d["dt2"] = rng.uniform(1, 6, B).astype(np.float32)
    d["t1"] = np.linspace(5, 50, B).astype(np.float32)
    d["sample_real"] = (np.arange(B) % R).astype(np.int32)
    return ocfg, otab, spec, ptab, d
