"""SRM_NUMERICS_CLOSED_FORM on the GPU.

This mode evaluates the reference's formulas the way exact arithmetic would (closed-form order-1
spline, flux in difference form), so its yardstick is the **fp64** twin of the oracle, not the
fp32 one: the reference's own fp32 evaluation differs from fp64 by 1e-4 .. 5e-2 of max on `dom`
and by more than the field's max on `gp0`.  Gates:
   |cuda - oracle64| <= 1e-5 * max|oracle64|   for dom, gp0, gp1, gdt1, and rel 1e-5 on the loss terms;
   and the closed form must be at least 10x closer to fp64 than the fp32 reference-order oracle is.
"""
import os

import numpy as np
import pytest
import torch

import util as U

srm, O = U.srm, U.O
pytestmark = pytest.mark.gpu

CASES = [
    dict(W=64, H=32, D=16, T=4, K=2, seed=2011, all_layers=True),        # TMA path, full z-chunk
    dict(W=39, H=39, D=1, T=8, K=4, seed=2001),                           # config-1 shape, cooperative loader
    dict(W=24, H=20, D=6, T=3, K=2, seed=2002, all_layers=True),         # partial tiles, partial z-chunk
    dict(W=16, H=16, D=4, T=2, K=2, seed=2003, all_layers=True, use_blocking_factor=True),
    dict(W=36, H=10, D=9, T=3, K=2, seed=2004),                           # two z-chunks of 8 + 1
    dict(W=33, H=7, D=2, T=1, K=1, seed=2005, wells="none"),
    dict(W=64, H=20, D=5, T=2, K=1, seed=2006, wells=("crowded", 4)),    # staged column lists (well_tile.cuh), a duplicate connection
    dict(W=64, H=20, D=5, T=2, K=1, seed=2007, wells=("crowded", 18)),   # more well columns in a tile than the lists hold
    # the lean TMA pair's rings: columns shorter than the four-plane pressure ring / the two-plane face ring, tiles cut by the
    # grid on both axes (64 x 16 tiles above 47 cells of width, 32 x 32 below), several tiles across
    dict(W=8, H=5, D=1, T=2, K=1, seed=2021),
    dict(W=72, H=33, D=2, T=2, K=1, seed=2022, all_layers=True),
    dict(W=128, H=17, D=3, T=1, K=2, seed=2023, all_layers=True),
    dict(W=200, H=16, D=7, T=1, K=1, seed=2024),
    dict(W=48, H=40, D=4, T=1, K=1, seed=2025),
]


@pytest.mark.parametrize("kw", CASES)
def test_closed_form_vs_fp64_oracle(kw):
    kw = dict(kw)
    if isinstance(kw.get("wells"), tuple):
        # rate-controlled connections only: at the BHP limit the fp64 yardstick and an fp32 evaluation may sit on different
        # sides of the switch (d rate / d p jumps there), which says nothing about the kernels
        kw["wells"] = U.crowded_wells(kw["D"], kw["wells"][1], minimum_bhp=3000.0)
    ocfg, otab, spec, ptab, batch = U.make_case(**kw)
    o64 = U.oracle_run(ocfg, otab, batch, dtype=torch.float64)
    o32 = U.oracle_run(ocfg, otab, batch)
    c = U.cuda_run(spec, ptab, batch, numerics="closed_form")
    for k in ("dom", "gp0", "gp1", "gdt1"):
        e_cf = U.rel_to_max(c[k], o64[k])
        e_32 = U.rel_to_max(o32[k], o64[k])
        assert e_cf <= 1e-5, (k, e_cf)
        assert e_cf * 10 <= e_32 or e_32 < 1e-6, (k, e_cf, e_32)
    assert np.allclose(c["terms"][:3], o64["terms"][:3], rtol=1e-5, atol=0)
    assert np.isclose(c["terms"][3], o64["terms"][3], rtol=1e-3)           # tde: 1e-12-sized, N == 0 vs fp64 residue
    assert np.all(c["gdt2"] == 0.0)                                         # analytically zero
    if ocfg.wells:
        for k in ("qw", "pwfw"):
            assert np.all(np.abs(c[k] - o64[k]) <= 1e-5 * np.abs(o64[k]) + 1e-5 * np.abs(o64[k]).max()), k


def test_closed_form_pvt_matches_exact_piecewise_linear_form():
    g = np.load(os.path.join(U.GOLDEN, "pvt_golden.npz"))
    tabs = srm.pvt.SplineTables(knots=g["knots"], w=g["w"], v=g["v"], order=1, properties=srm.pvt.DG_PROPERTIES)
    eng = srm.SrmPhysics(srm.PhysicsSpec(), tabs, numerics="closed_form")
    p = g["p"]
    val, der = eng.pvt_eval(torch.from_numpy(p).cuda())
    cols = O.load_pvt_table(os.path.join(U.GOLDEN, "pvt_table.npz"))
    otab = O.build_spline_table(cols, O.DG_PROPS)
    ocfg = O.OracleConfig()
    off_knot = np.abs(np.clip(p, 14.7, 1e4)[:, None] - otab.c[None, :]).min(1) > 1e-3
    for q in range(2):
        cv, cs = O.spline_closed_form_fp64(p, otab, ocfg, q)
        assert np.allclose(val[q].cpu().numpy(), cv, rtol=2e-6, atol=0)
        assert np.allclose(der[q].cpu().numpy()[off_knot], cs[off_knot], rtol=2e-6, atol=0)


def test_closed_form_on_knot_rule():
    """exactly on a knot TF's [r >= 1e-10] mask drops the knot's term: slope = mean of both sides"""
    tabs = srm.build_spline_tables(srm.load_default_pvt_table(), srm.pvt.DG_PROPERTIES)
    eng = srm.SrmPhysics(srm.PhysicsSpec(), tabs, numerics="closed_form")
    p = torch.tensor([4999.0, 5000.0, 5001.0], device="cuda")
    _, der = eng.pvt_eval(p)
    d = der.cpu().numpy()
    assert np.allclose(d[:, 1], 0.5 * (d[:, 0] + d[:, 2]), rtol=1e-6)
    ref = srm.SrmPhysics(srm.PhysicsSpec(), tabs, numerics="reference")
    _, der_ref = ref.pvt_eval(p)
    assert np.allclose(der_ref.cpu().numpy()[:, 1], d[:, 1], rtol=2e-2)     # reference order, with its fp32 noise


def test_closed_form_rejects_order_two():
    tabs = srm.build_spline_tables(srm.load_default_pvt_table(), srm.pvt.DG_PROPERTIES, order=2)
    with pytest.raises(srm._lib.SrmError, match="spline_order 1"):
        srm.SrmPhysics(srm.PhysicsSpec(), tabs, numerics="closed_form")
