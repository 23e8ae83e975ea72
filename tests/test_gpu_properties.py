"""Size-independent properties at BASELINE.json's full single-GPU size (config 2: 64x64x16, T=32,
K=16 -> 3.4e7 cell-timesteps), where the CPU oracle would take minutes."""
import numpy as np
import pytest
import torch

import util as U

srm = U.srm
pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cfg2():
    c = srm.synth.CONFIGS["cfg2"]
    wells = srm.config.scaled_default_wells(c["W"], c["H"], c["D"])
    spec = srm.PhysicsSpec(D=c["D"], H=c["H"], W=c["W"], wells=wells)
    tabs = srm.build_spline_tables(srm.load_default_pvt_table(), srm.pvt.DG_PROPERTIES)
    eng = srm.SrmPhysics(spec, tabs)
    b = srm.synth.make_batch(c["W"], c["H"], c["D"], c["T"], c["K"], [(w.i, w.j) for w in wells], seed=2002, device="cuda")
    d = dict(kx=b.kx, sample_real=b.sample_real, p0=b.p0, p1=b.p1, dt1=b.dt1, dt2=b.dt2, t1=b.t1)
    return eng, d


def test_terms_are_the_sums_of_squares_of_the_fields(cfg2):
    eng, d = cfg2
    fw = eng.forward(want_dom=True, **d)
    dom = fw["dom"].double()
    assert abs(float((dom * dom).sum()) / float(fw["terms"][0, 0]) - 1.0) < 1e-5
    B, N = d["p0"].shape[0], d["p0"][0].numel()
    assert fw["terms"][1].tolist() == [B * N, B * N, B * N, B * N, 0, 0, 0, 0]
    assert torch.isfinite(fw["dom"]).all()


def test_uniform_pressure_has_zero_flux_and_zero_accumulation(cfg2):
    """p0 == p1 == const and no wells: every flux difference, the accumulation and mbc vanish exactly
    in the 3-D extension's z faces; the reference's a*p form leaves only its rounding residue."""
    eng, d = cfg2
    spec = srm.PhysicsSpec(D=eng.spec.D, H=eng.spec.H, W=eng.spec.W, wells=[])
    e2 = srm.SrmPhysics(spec, eng.tables)
    p = torch.full_like(d["p0"][:4], 4700.0)
    fw = e2.forward(kx=d["kx"], sample_real=d["sample_real"][:4].contiguous(), p0=p, p1=p, dt1=d["dt1"][:4].contiguous(),
                    dt2=d["dt2"][:4].contiguous(), t1=d["t1"][:4].contiguous(), want_dom=True)
    dom = fw["dom"]
    scale = 4700.0 * 4 * 1e-3 * e2.spec.dx * e2.spec.dy * e2.spec.dz      # |a*p| * dv, a ~ 1e-3
    assert float(dom.abs().max()) < 1e-5 * scale
    assert float(fw["terms"][0, 2]) == 0.0                                  # mbc: A1 - A0 == 0 exactly


def test_loss_is_additive_over_sample_shards(cfg2):
    """the batch shards across ranks with no data-path exchange: terms(all) == sum of terms(shards)"""
    eng, d = cfg2
    full = eng.forward(**d)["terms"][0].double().clone()
    B = d["p0"].shape[0]
    acc = torch.zeros(8, dtype=torch.float64, device="cuda")
    for lo in range(0, B, B // 4):
        sl = slice(lo, lo + B // 4)
        part = {k: (v[sl].contiguous() if k != "kx" else v) for k, v in d.items()}
        acc += eng.forward(**part)["terms"][0].double()
    assert torch.allclose(acc, full, rtol=1e-6, atol=0)


def test_gradient_is_linear_in_the_upstream_weights(cfg2):
    eng, d = cfg2
    sub = {k: (v[:32].contiguous() if k != "kx" else v) for k, v in d.items()}
    eng.forward(**sub)
    w1 = torch.tensor([1.0, 0, 0, 0, 0, 0, 0, 0], device="cuda")
    w2 = torch.tensor([0, 0.5, 2.0, 0.25, 0, 0, 0, 0], device="cuda")
    g1 = [t.clone() for t in eng.backward(dterms=w1, **sub)]
    g2 = [t.clone() for t in eng.backward(dterms=w2, **sub)]
    g12 = eng.backward(dterms=w1 + w2, **sub)
    for a, b, c in zip(g1, g2, g12):
        ref = a.double() + b.double()
        assert float((c.double() - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_directional_derivative_of_the_loss(cfg2):
    """<grad, v> matches a central difference of the fp32 loss along a smooth direction v.  The
    direction avoids PVT-knot crossings by being small; tolerance reflects fp32 loss rounding."""
    eng, d = cfg2
    sub = {k: (v[:8].contiguous() if k != "kx" else v) for k, v in d.items()}
    w = torch.tensor([0.0, 0.0, 1.0, 0.0, 0, 0, 0, 0], device="cuda")        # mbc: smooth, well conditioned
    eng.forward(**sub)
    gp0, gp1, gdt1, gdt2 = eng.backward(dterms=w, **sub)
    gen = torch.Generator(device="cuda").manual_seed(5)
    v = torch.randn(sub["dt1"].shape, generator=gen, device="cuda")
    h = 1e-3

    def loss(dt1):
        t = eng.forward(**{**sub, "dt1": dt1.contiguous()})["terms"][0].double()
        return float((t * w.double()).sum())

    fd = (loss(sub["dt1"] + h * v) - loss(sub["dt1"] - h * v)) / (2 * h)
    an = float((gdt1.double() * v.double()).sum())
    assert abs(fd - an) <= 2e-3 * abs(an)


@pytest.mark.parametrize("to_host", [True, False])
@pytest.mark.parametrize("n_chunks", [1, 3])
def test_host_pipeline_matches_resident_run(n_chunks, to_host):
    """engine.HostPipeline (pinned host batch, chunks of whole realisations over three streams) returns the same
    gradients (bit for bit where no atomics are involved) and the same loss terms as one resident forward + adjoint: samples are independent."""
    ocfg, otab, spec, ptab, batch = U.make_case(W=16, H=12, D=3, T=4, K=5, seed=77, all_layers=True)
    eng = srm.SrmPhysics(spec, ptab, device=0, pvt_lut=True, lut_range=(4000.0, 5100.0))
    d = U.to_dev(batch, "cuda")
    w = torch.tensor(U.WEIGHTS, device="cuda")
    fw = eng.forward(**d)
    g = eng.backward(dterms=w, **d)
    host = {k: v.cpu().pin_memory() for k, v in d.items()}
    pipe = srm.engine.HostPipeline(eng, host, w, n_chunks=n_chunks, grads_to_host=to_host)
    for _ in range(2):                                   # the second pass reuses the slots
        hterms, hg = pipe.step()
    assert len(pipe.chunks) == n_chunks
    assert all(v.is_cuda != to_host for v in hg.values())
    assert pipe.d2h_bytes == hterms.numel() * 4 + (sum(v.numel() * 4 for v in hg.values()) if to_host else 0)
    hg = {k: v.cpu() for k, v in hg.items()}
    assert torch.allclose(hterms, fw["terms"].cpu(), rtol=1e-6)
    for name, t in zip(("gp0", "gp1", "gdt1", "gdt2"), g):
        if name == "gp1":     # the inner-boundary scatter uses float atomics where well cells are adjacent: order-dependent last bits
            assert torch.allclose(hg[name], t.cpu(), rtol=1e-5, atol=1e-6 * float(t.abs().max())), name
        else:
            assert torch.equal(hg[name], t.cpu()), name
    eng.close()


@pytest.mark.parametrize("W", [16, 39])       # even width: kernels_dg4.cu; odd: kernels_ref2.cu
def test_graphed_step_replays_the_eager_step(W):
    """engine.GraphedStep (forward + adjoint captured in one CUDA graph): a replay returns what the eager calls return,
    also after the static input buffers have been refilled."""
    ocfg, otab, spec, ptab, batch = U.make_case(W=W, H=12, D=2, T=4, K=2, seed=91)
    eng = srm.SrmPhysics(spec, ptab, device=0, pvt_lut=True)
    d = U.to_dev(batch, "cuda")
    w = torch.tensor(U.WEIGHTS, device="cuda")
    gs = srm.engine.GraphedStep(eng, d, w)
    assert gs.kernels >= 5
    for shift in (0.0, -7.25):
        cur = {**d, "p0": d["p0"] + shift, "p1": d["p1"] + shift}
        gs.load(p0=cur["p0"], p1=cur["p1"])
        terms, grads = gs.replay()
        torch.cuda.synchronize()
        fw = eng.forward(**cur)
        g = eng.backward(dterms=w, **cur)
        assert torch.allclose(terms, fw["terms"], rtol=1e-6)
        for name, a, b in zip(("gp0", "gp1", "gdt1", "gdt2"), grads, g):
            assert torch.allclose(a, b, rtol=1e-5, atol=1e-6 * float(b.abs().max())), name
    eng.close()


def test_lean_pair_at_full_size_equals_the_per_cell_path_and_repeats(cfg2):
    """The headline kernels (kernels_dg4.cu: exact table, z-marching tiles, split barrier) at BASELINE config 2's full
    size: residual field bit-identical to the per-cell spline path (no table, no tiles), gradients equal to rounding,
    and 20 repeats under load return the same bits (a hazard in the plane pipeline would show as run-to-run noise)."""
    eng0, d = cfg2
    w = torch.tensor(U.WEIGHTS, device="cuda")
    fw0 = eng0.forward(want_dom=True, **d)
    g0 = [t.clone() for t in eng0.backward(dterms=w, **d)]
    dom0 = fw0["dom"].clone()
    eng = srm.SrmPhysics(eng0.spec, eng0.tables, device=0, pvt_lut=True)
    ref = None
    for _ in range(20):
        dom = eng.forward(want_dom=True, **d)["dom"]
        g = eng.backward(dterms=w, **d)
        cur = [dom.clone(), g[0].clone()]                # dom and gp0 carry no atomics
        if ref is None:
            ref = cur + [g[1].clone(), g[2].clone()]
        else:
            for i, (a, c) in enumerate(zip(ref, cur)):
                assert torch.equal(a.view(torch.int32), c.view(torch.int32)), i
    assert torch.equal(ref[0].view(torch.int32), dom0.view(torch.int32))
    for name, a, b in zip(("gp0", "gp1", "gdt1"), (ref[1], ref[2], ref[3]), g0):
        assert torch.allclose(a, b, rtol=1e-5, atol=1e-5 * float(b.abs().max())), name
    eng.close()


@pytest.mark.parametrize("pvt_lut", [True, False])
def test_interleaved_batches_share_one_workspace(pvt_lut):
    """fwd(A), fwd(B), bwd(A), bwd(B) on ONE workspace (gradient accumulation with two .backward() calls): bwd(A)
    must recompute A's forward state -- the workspace holds B's -- and bwd(B) must then NOT take A's recomputed
    state for its own (the saved-state fingerprint went stale after a recompute)."""
    _, _, spec, ptab, b1 = U.make_case(W=24, H=10, D=3, T=2, K=2, seed=4101)
    _, _, _, _, b2 = U.make_case(W=24, H=10, D=3, T=2, K=2, seed=4102)
    eng = srm.SrmPhysics(spec, ptab, device=0, pvt_lut=pvt_lut)
    dev = torch.device("cuda", 0)
    dA, dB = U.to_dev(b1, dev), U.to_dev(b2, dev)
    w = torch.tensor(U.WEIGHTS, dtype=torch.float32, device=dev)
    ref = {}
    for nm, d in (("A", dA), ("B", dB)):
        eng.forward(**d)
        ref[nm] = [t.clone() for t in eng.backward(dterms=w, **d)]
    eng.forward(**dA)
    eng.forward(**dB)
    gA = [t.clone() for t in eng.backward(dterms=w, **dA)]
    gB = [t.clone() for t in eng.backward(dterms=w, **dB)]
    for got, want in ((gA, ref["A"]), (gB, ref["B"])):
        for x, y in zip(got, want):
            assert torch.equal(x, y)
    # same pointers, other per-sample scalars: the fingerprint covers dt1 / dt2 / t1 too
    d3 = dict(dA)
    d3["dt1"] = (dA["dt1"] * 1.5).contiguous()
    eng.forward(**dA)
    g3 = eng.backward(dterms=w, **d3)
    eng.forward(**d3)
    g3r = eng.backward(dterms=w, **d3)
    for x, y in zip(g3, g3r):
        assert torch.equal(x, y)
    eng.close()


def test_graphed_step_owns_its_workspace():
    """replays stay valid after the engine's cached workspace is re-allocated by calls with another batch shape"""
    _, _, spec, ptab, b1 = U.make_case(W=24, H=10, D=3, T=2, K=2, seed=4103)
    _, _, _, _, b2 = U.make_case(W=24, H=10, D=3, T=3, K=3, seed=4104)
    eng = srm.SrmPhysics(spec, ptab, device=0, pvt_lut=True)
    dev = torch.device("cuda", 0)
    dA, dB = U.to_dev(b1, dev), U.to_dev(b2, dev)
    w = torch.tensor(U.WEIGHTS, dtype=torch.float32, device=dev)
    gs = srm.engine.GraphedStep(eng, dA, w)
    t0, g0 = gs.replay()
    t0 = t0.clone(); g0 = [t.clone() for t in g0]
    for _ in range(3):                       # other shapes: the engine drops and re-allocates its own workspace
        eng.forward(**dB); eng.backward(dterms=w, **dB)
        junk = torch.full((eng.workspace_bytes(4, 2),), 255, dtype=torch.uint8, device=dev)   # recycle freed blocks
        eng.forward(**dA); eng.backward(dterms=w, **dA)
        del junk
    t1, g1 = gs.replay()
    torch.cuda.synchronize()
    assert torch.equal(t0, t1)
    for x, y in zip(g0, g1):
        assert torch.equal(x, y)
    eng.close()


def test_engine_rejects_short_per_sample_inputs():
    _, _, spec, ptab, b1 = U.make_case(W=24, H=10, D=3, T=2, K=2, seed=4105)
    eng = srm.SrmPhysics(spec, ptab, device=0)
    d = U.to_dev(b1, torch.device("cuda", 0))
    for k in ("dt1", "dt2", "t1", "sample_real"):
        bad = dict(d)
        bad[k] = d[k][:-1].contiguous()
        with pytest.raises(ValueError):
            eng.forward(**bad)
    # an out-of-range realisation index is clamped by the kernels, not an out-of-bounds read
    bad = dict(d)
    bad["sample_real"] = torch.full_like(d["sample_real"], 7)
    last = dict(d)
    last["sample_real"] = torch.full_like(d["sample_real"], d["kx"].shape[0] - 1)
    assert torch.equal(eng.forward(**bad)["terms"], eng.forward(**last)["terms"])
    eng.close()
