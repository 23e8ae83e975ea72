"""CPU checks of the gas-condensate (two-phase) oracle restatement (physics_loss.py:230-712).

The reference ships no known-answer test for this path (parity unpinned); what pins the
restatement here: fp64 central finite differences of the whole loss against the oracle's autograd
gradient (catches a wrong TF-gradient routing in min/max/clip/divide_no_nan/where and in the chord
slopes), structural identities of the residual, and the Corey end-point rules."""
import os

import numpy as np
import pytest
import torch

import util as U

O = U.O


def gc_case(seed=0, B=2, D=2, H=5, W=6, sg_lo=0.2, sg_hi=0.75, wells=True):
    cols = O.load_pvt_table(os.path.join(U.GOLDEN, "pvt_table.npz"))
    tab = O.build_spline_table(cols, O.GC_PROPS, order=1, lam=0.001)
    wl = [O.Well(i=2, j=2, k=0, value=500.0), O.Well(i=4, j=3, k=D - 1, value=1000.0)] if wells else []
    cfg = O.OracleConfig(D=D, H=H, W=W, wells=wl)
    rng = np.random.default_rng(seed)
    shp = (B, D, H, W)
    kx = rng.uniform(1, 6, (1, D, H, W)).astype(np.float32)
    p0 = (4700 + rng.uniform(-30, 30, shp)).astype(np.float32)
    p1 = (p0 - rng.uniform(1, 25, shp)).astype(np.float32)
    sg0 = rng.uniform(sg_lo, sg_hi, shp).astype(np.float32)
    sg1 = (sg0 - rng.uniform(0.001, 0.02, shp)).astype(np.float32)
    so0 = (np.float32(0.78) - sg0).astype(np.float32)
    so1 = (np.float32(0.78) - sg1).astype(np.float32)
    dt1 = rng.uniform(1, 6, B).astype(np.float32)
    dt2 = rng.uniform(1, 6, B).astype(np.float32)
    t = np.linspace(5, 50, B).astype(np.float32)
    sr = np.zeros(B, np.int32)
    return cfg, tab, dict(kx=kx, p0=p0, p1=p1, sg0=sg0, sg1=sg1, so0=so0, so1=so1, dt1=dt1, dt2=dt2, t_days=t,
                          sample_real=sr)


W_ALL = [1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0]


def test_corey_end_points_and_pinned_pow():
    cfg = O.OracleConfig()
    sg = torch.tensor([0.0, 0.05, 0.2, 0.36, 0.5, 0.58, 0.59, 0.78], dtype=torch.float32)
    krog, krgo = O.corey_krog_krgo_t(sg, cfg)
    kn, gn = O.corey_krog_krgo_np(sg.numpy(), cfg)
    assert np.allclose(krog.numpy(), kn, rtol=3e-7, atol=0) and np.allclose(krgo.numpy(), gn, rtol=6e-7, atol=0)
    assert krog[-1] == 0 and krog[4] == 0            # so <= Swmin + max(Sorg, Socr)  (relative_permeability.py:66-67)
    assert krgo[-1] == np.float32(0.9)               # sg > 1 - Swmin - Sorg -> krg_Swmin (:68)
    assert krgo[1] == 0 and krog[0] > 0


def test_gc_gradient_matches_fp64_finite_differences():
    cfg, tab, d = gc_case(seed=3)
    o = O.gc_forward_backward(cfg, tab, weights=W_ALL, dtype=torch.float64, **d)
    rng = np.random.default_rng(1)
    for name, g in (("p0", "gp0"), ("p1", "gp1"), ("sg0", "gsg0"), ("sg1", "gsg1"), ("so0", "gso0"), ("so1", "gso1"),
                    ("dt1", "gdt1")):
        base = d[name].astype(np.float64)
        for _ in range(4):
            idx = tuple(rng.integers(0, s) for s in base.shape)
            h = 1e-6 * max(1.0, abs(base[idx]))
            lp, lm = [], None
            vals = []
            for sgn in (+1, -1):
                x = base.copy()
                x[idx] += sgn * h
                dd = dict(d)
                dd[name] = x
                vals.append(O.gc_forward_backward(cfg, tab, weights=W_ALL, dtype=torch.float64, **dd)["loss"])
            fd = (vals[0] - vals[1]) / (2 * h)
            an = o[g][idx]
            assert abs(fd - an) <= 2e-5 * max(abs(an), 1e-6 * np.abs(o[g]).max()) + 1e-7 * np.abs(o[g]).max(), (name, idx, fd, an)


def test_gc_structure():
    """no wells, equal pressures everywhere and no change in time: every component vanishes except the
    rounding-level truncation term."""
    cfg, tab, d = gc_case(seed=5, wells=False)
    d["p0"][:] = 4700.0
    d["p1"][:] = 4700.0
    d["sg1"] = d["sg0"].copy()
    d["so1"] = d["so0"].copy()
    o = O.gc_forward_backward(cfg, tab, weights=W_ALL, dtype=torch.float64, **d)
    # the a*p flux form cancels to rounding only (-a*p_n + (sum a)*p_c with equal pressures)
    assert np.abs(o["dom"]).max() < 1e-9 and np.abs(o["mbc"]).max() < 1e-9 and np.abs(o["ibc"]).max() == 0
    # Nz == 1 reduces to the shipped 2-D arithmetic: the z faces contribute exactly zero
    cfg1, tab1, d1 = gc_case(seed=6, D=1)
    a = O.gc_forward_backward(cfg1, tab1, weights=W_ALL, dtype=torch.float32, **d1)
    assert np.isfinite(a["dom"]).all() and a["terms"][0] > 0


def test_gc_wells_split_sums_to_phase_rates():
    cfg, tab, d = gc_case(seed=7, sg_lo=0.2, sg_hi=0.35)          # mobile oil
    res = O.gc_residual(cfg, tab, torch.from_numpy(d["kx"]), *[torch.from_numpy(d[k]) for k in
                        ("p0", "p1", "sg0", "sg1", "so0", "so1", "dt1", "dt2")], d["t_days"], d["sample_real"])
    qgg, qgo, qoo, qog = res["qw4"]
    assert (res["krog1"] > 0).any()
    assert torch.all(qgg + qgo <= torch.tensor([500.0, 1000.0]) * (1 + 1e-6))
    assert torch.all(qoo >= 0) and torch.all(qog >= 0)


def test_gc_oracle_reproduces_golden():
    from golden.make_golden import GC_CASES, GC_WEIGHTS
    g = np.load(os.path.join(U.GOLDEN, "gc_3d.npz"))
    ocfg, otab, spec, ptab, d = U.gc_case(**GC_CASES["gc_3d"])
    for k in ("kx", "p0", "p1", "sg0", "sg1", "so0", "so1", "dt1", "dt2"):
        assert np.array_equal(d[k], g[k]), k                    # the seeded inputs are reproducible
    otab.w[:], otab.v[:] = g["w"], g["v"]                       # the golden's weights: this box's LAPACK does not enter
    o = O.gc_forward_backward(ocfg, otab, d["kx"], d["p0"], d["p1"], d["sg0"], d["sg1"], d["so0"], d["so1"], d["dt1"],
                              d["dt2"], d["t1"], d["sample_real"], GC_WEIGHTS)
    assert np.array_equal(o["dom"], g["o_dom"])
    assert np.allclose(o["terms"], g["o_terms"], rtol=1e-6)
    for k in ("gp0", "gp1", "gsg1"):
        assert np.allclose(o[k], g["o_" + k], rtol=1e-5, atol=1e-6 * np.abs(g["o_" + k]).max())


def _gc_ref_case(g, case):
    import srm_oracle as O
    W, H = int(g[f"{case}_W"]), int(g[f"{case}_H"])
    wl = [O.Well(i=int(r[0]), j=int(r[1]), k=int(r[2]), value=float(r[3])) for r in g[f"{case}_wells"]]
    blocking = bool(g[f"{case}_blocking"]) if f"{case}_blocking" in g.files else False
    return O.OracleConfig(D=1, H=H, W=W, wells=wl, use_blocking_factor=blocking, n_intervals=8)


@pytest.mark.parametrize("case", ["a", "b", "c"])
def test_oracle_gc_residual_equals_the_reference_fragment_bit_for_bit(case):
    """PIN: tests/golden/reference_gc_residual.npz holds dom, ibc, mbc, cmbc computed by the reference's OWN
    physics_error_gas_oil_2D (physics_loss.py:230-714, executed by tests/golden/make_reference_gc_golden.py through the
    torch-backed TF stand-in).  The oracle must reproduce the fields bit for bit on the same inputs (2-D grids)."""
    g = np.load(os.path.join(U.GOLDEN, "reference_gc_residual.npz"))
    cfg = _gc_ref_case(g, case)
    cols = O.load_pvt_table(os.path.join(U.GOLDEN, "pvt_table.npz"))
    tab = O.build_spline_table(cols, O.GC_PROPS, order=1, lam=0.001)
    tt = lambda k: torch.as_tensor(g[f"{case}_{k}"])
    res = O.gc_residual(cfg, tab, tt("kx"), tt("p0"), tt("p1"), tt("sg0"), tt("sg1"), tt("so0"), tt("so1"), tt("dt1"), tt("dt2"),
                        g[f"{case}_t1"], g[f"{case}_sample_real"])
    for k, rk in (("dom", "ref_dom"), ("ibc", "ref_ibc"), ("cmbc", "ref_cmbc")):
        assert np.array_equal(res[k].detach().numpy().view(np.uint32), g[f"{case}_{rk}"].view(np.uint32)), k
    assert np.allclose(res["mbc"].detach().numpy(), g[f"{case}_ref_mbc"], rtol=1e-6, atol=0)
    assert np.abs(g[f"{case}_ref_dom"]).max() > 0 and np.abs(g[f"{case}_ref_cmbc"]).max() > 0


# ---- gradients of the reference's OWN two-phase op graph (tests/golden/make_reference_gc_grad_golden.py) ---------------
# smallest passing H3 rtol of the oracle against them, over three cases and five term selections:
#   gp0 1.2e-5, gp1 1.4e-5, gsg*/gso* 5.0e-6, gdt1 3.2e-6; wider only where the reference's own fp32 autodiff is noise:
#   gp0 / gp1 of `dom` in case c (p1 == p0 cells: the 1/dp^2 pieces of the chord slopes, physics_loss.py:465-466) 2.1e-4, and
#   gdt1 of the cmbc term (a sum of truncation brackets that vanish analytically) 1e-3.
GC_GRAD_FIELDS = ("p0", "p1", "sg0", "sg1", "so0", "so1", "dt1")
GC_TERMS = ("dom", "ibc", "mbc", "tde", "obc", "ic", "td", "cmbc")


def gc_grad_gate(case, name, field):
    if name == "cmbc" and field == "dt1":
        return 2e-3
    if case == "c" and name == "dom" and field in ("p0", "p1"):       # chord slopes at p1 == p0 cells: 1/dp^2 pieces (oracle 2.1e-4, CUDA 3.6e-4)
        return 5e-4
    return 2e-5


def gc_grad_selections(nwt):
    sel = {"batch": dict(dom=nwt[0], ibc=nwt[3], mbc=nwt[5], cmbc=nwt[6]), "dom": dict(dom=nwt[0]), "ibc": dict(ibc=nwt[3]),
           "mbc": dict(mbc=nwt[5]), "cmbc": dict(cmbc=nwt[6])}
    return {k: [float(v.get(t, 0.0)) for t in GC_TERMS] for k, v in sel.items()}


def h3_min_rtol(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    den = np.abs(b) + np.abs(b).max()
    return float((np.abs(a - b) / np.maximum(den, 1e-300)).max()) if den.max() > 0 else float(np.abs(a).max())


@pytest.mark.parametrize("case", ["a", "b", "c", "d"])
def test_oracle_gc_gradients_equal_the_reference_graph_gradients(case):
    """PIN (adjoint, two-phase): tape.gradient of every weighted SSE term taken by the reference's OWN
    pinn_batch_sse_grad over physics_error_gas_oil_2D, PVTLayer (nested tape), RelativePermeability and
    WellRatesPressure (GC branch; case d with the blocking-factor integral, whose twenty Newton iterations per trapezoid
    node are all inside the tape), network outputs as trainable variables."""
    g = np.load(os.path.join(U.GOLDEN, "reference_gc_grad.npz"))
    cfg = _gc_ref_case(g, case)
    cols = O.load_pvt_table(os.path.join(U.GOLDEN, "pvt_table.npz"))
    tab = O.build_spline_table(cols, O.GC_PROPS, order=1, lam=0.001)
    a = lambda k: g[f"{case}_{k}"]
    live = 0
    for name, wts in gc_grad_selections(a("nwt")).items():
        r = O.gc_forward_backward(cfg, tab, a("kx"), a("p0"), a("p1"), a("sg0"), a("sg1"), a("so0"), a("so1"), a("dt1"), a("dt2"),
                                  a("t1"), a("sample_real"), wts)
        for f in GC_GRAD_FIELDS:
            ref = a(f"g_{name}_{f}")
            if name == "cmbc" and f == "dt1":
                # a sum of truncation brackets that vanish analytically: rounding noise, measured on the scale of the
                # batch gradient that reaches the same tensor
                assert np.abs(r["g" + f] - ref).max() <= 1e-6 * np.abs(a(f"g_batch_{f}")).max()
                continue
            assert h3_min_rtol(r["g" + f], ref) <= gc_grad_gate(case, name, f), (case, name, f, h3_min_rtol(r["g" + f], ref))
            live += int(np.abs(ref).max() > 0)
    assert live >= 20
    if case == "b":      # the BHP-limited connection: dq/dp and dq/dSg are live in the inner-boundary term
        assert np.abs(a("g_ibc_sg1")).max() > 0 and np.abs(a("g_ibc_p1")).max() > 0
