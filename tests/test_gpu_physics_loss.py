"""PhysicsLoss mirror end to end on the GPU: two tiny torch networks stand in for the Keras models;
the parameter gradients returned by pinn_batch_sse_grad must equal those of the same chain with the
oracle's differentiable residual in place of the CUDA op."""
import copy

import numpy as np
import pytest
import torch
import torch.nn as nn

import util as U

srm, O = U.srm, U.O
pytestmark = pytest.mark.gpu


class PressureNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.l1, self.l2 = nn.Linear(5, 8), nn.Linear(8, 1)

    def forward(self, x):
        return 4700.0 + 250.0 * torch.tanh(self.l2(torch.tanh(self.l1(x))))


class StepNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.l = nn.Linear(5, 1)

    def forward(self, x):
        return 0.1 + 9.9 * torch.sigmoid(self.l(x))


def features(B, D, H, W, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand((B, D, H, W, 5), generator=g) * 2 - 1
    x[..., 3] = (torch.rand((B, 1, 1, 1), generator=g) * 1.2 - 0.9).expand(B, D, H, W)    # one time per sample
    return x


def oracle_chain(pn, sn, x, ocfg, otab, weights, k_stats=(0.26, 24.0)):
    B = x.shape[0]
    kx = torch.from_numpy(O.denorm_log(x[..., 4].numpy(), *k_stats))
    p0 = pn(x)[..., 0]
    dt1 = sn(x).reshape(B, -1).mean(1)
    shift = torch.zeros_like(x)
    shift[..., 3] = (2.0 / 365.0 * dt1).view(-1, 1, 1, 1)
    x1 = x + shift
    p1 = pn(x1)[..., 0]
    dt2 = sn(x1).reshape(B, -1).mean(1)
    t1 = O.denorm_linear(x1[:, 0, 0, 0, 3].detach().numpy(), 0.0, 365.0)
    res = O.dg_residual(ocfg, otab, kx, p0, p1, dt1, dt2, t1, np.arange(B), dtype=torch.float32)
    terms = O.dg_loss_terms(res)
    loss = (terms * torch.tensor(weights)).sum()
    params = list(pn.parameters()) + list(sn.parameters())
    return terms.detach().numpy(), [g.numpy() for g in torch.autograd.grad(loss, params)]


def test_pinn_batch_sse_grad_contract_and_parity():
    torch.manual_seed(0)
    D, H, W, B = 2, 9, 12, 3
    wells = srm.config.scaled_default_wells(W, H, D)
    spec = srm.PhysicsSpec(D=D, H=H, W=W, wells=wells)
    tabs = srm.build_spline_tables(srm.load_default_pvt_table(), srm.pvt.DG_PROPERTIES)
    eng = srm.SrmPhysics(spec, tabs)
    pn, sn = PressureNet(), StepNet()
    pn_g, sn_g = copy.deepcopy(pn).cuda(), copy.deepcopy(sn).cuda()
    pvt = srm.PVTLayer(eng)
    wrb = srm.WellRatesPressure(eng)
    loss = srm.PhysicsLoss(pn_g, pvt, sn_g, wrb, saturation_model=None,
                           optimizer_model_names_map={"pressure": "pressure", "time_step": "time_step"})
    # attributes the training loop reads (training.py:552-560,605)
    assert loss.trainable_models == [pn_g, sn_g] and loss.trainable_models_keys == ["pressure", "time_step"]
    assert loss.loss_keys == {"gas": ["dom", "ibc", "obc", "ic", "td", "mbc", "cmbc", "tde"]}
    assert loss.physics_mode_fraction == 1.0 and set(loss.optimizer_model_map) == {"pressure", "time_step"}
    x = features(B, D, H, W, 5)
    wmse, wmse_grad, wsse, error_count, y_model = loss.pinn_batch_sse_grad(x.cuda(), None)
    assert wmse[0].shape == (8,) and len(wmse_grad) == 2 and y_model.shape == (B, D, H, W, 1)
    assert [len(g) for g in wmse_grad] == [4, 2]
    # oracle chain on the CPU with the same weights
    ocfg = O.OracleConfig(D=D, H=H, W=W, wells=[O.Well(i=w.i, j=w.j, k=w.k, value=abs(w.q_target), producer=not np.signbit(w.q_target))
                                                  for w in wells])
    cols = O.load_pvt_table(U.GOLDEN + "/pvt_table.npz")
    otab = O.build_spline_table(cols, O.DG_PROPS)
    oterms, ograds = oracle_chain(pn, sn, x, ocfg, otab, [1.0, 1.0, 1.0, 1.0, 0, 0, 0, 0])
    got = wsse[0].cpu().numpy()
    keys = loss.loss_keys["gas"]
    for name, slot in (("dom", 0), ("ibc", 1), ("mbc", 2), ("tde", 3)):
        assert np.isclose(got[keys.index(name)], oterms[slot], rtol=2e-4), name
    flat = [g for gs in wmse_grad for g in gs]
    for a, b in zip(flat, ograds):
        a = a.cpu().numpy()
        assert np.abs(a - b).max() <= 2e-4 * np.abs(b).max() + 1e-30, (np.abs(a - b).max(), np.abs(b).max())


def test_well_rates_pressure_mirror_matches_sparse_tables():
    D, H, W, B = 1, 10, 10, 2
    wells = srm.config.scaled_default_wells(W, H, D)
    spec = srm.PhysicsSpec(D=D, H=H, W=W, wells=wells)
    tabs = srm.build_spline_tables(srm.load_default_pvt_table(), srm.pvt.DG_PROPERTIES)
    eng = srm.SrmPhysics(spec, tabs)
    wrb = srm.WellRatesPressure(eng)
    x = features(B, D, H, W, 9).cuda()
    p = torch.full((B, D, H, W, 1), 4800.0, device="cuda")
    q, pwf = wrb.compute_rates_and_bhp(x, p, None, None, None)
    assert q.shape == (B, D, H, W, 1) and pwf.shape == q.shape
    cells = {(w.k, w.j, w.i) for w in wells}
    nz = {tuple(i[1:4]) for i in torch.nonzero(q[..., 0] != 0).tolist()}
    assert nz <= cells and len(nz) >= 4                      # zero off-well (scatter semantics)
    assert float(q.max()) <= 1000.0 + 1e-3


class SatNet(nn.Module):
    def __init__(self):
        super().__init__()
        self.l = nn.Linear(5, 1)

    def forward(self, x):
        return 0.25 + 0.45 * torch.sigmoid(self.l(x))          # Sg in (0.25, 0.70): mobile and immobile oil


def oracle_chain_gc(pn, sn, gn, x, ocfg, otab, weights, k_stats=(0.26, 24.0)):
    B = x.shape[0]
    kx = torch.from_numpy(O.denorm_log(x[..., 4].numpy(), *k_stats))
    p0 = pn(x)[..., 0]
    dt1 = sn(x).reshape(B, -1).mean(1)
    shift = torch.zeros_like(x)
    shift[..., 3] = (2.0 / 365.0 * dt1).view(-1, 1, 1, 1)
    x1 = x + shift
    p1 = pn(x1)[..., 0]
    dt2 = sn(x1).reshape(B, -1).mean(1)
    sg0, sg1 = gn(x)[..., 0], gn(x1)[..., 0]
    top = 1.0 - ocfg.Swmin
    t1 = O.denorm_linear(x1[:, 0, 0, 0, 3].detach().numpy(), 0.0, 365.0)
    res = O.gc_residual(ocfg, otab, kx, p0, p1, sg0, sg1, top - sg0, top - sg1, dt1, dt2, t1, np.arange(B))
    terms = O.gc_loss_terms(res)
    loss = (terms * torch.tensor(weights)).sum()
    params = list(pn.parameters()) + list(sn.parameters()) + list(gn.parameters())
    return terms.detach().numpy(), [g.numpy() for g in torch.autograd.grad(loss, params)]


def test_pinn_batch_sse_grad_gas_condensate():
    """saturation_model given -> the two-phase loss (physics_loss.py:230-712) through the GC kernels"""
    torch.manual_seed(1)
    D, H, W, B = 2, 7, 9, 3
    conns = [dict(i=2, j=2, k=0, type="producer", control="ORAT", value=500.0, minimum_bhp=4100.0, wellbore_radius=0.09525,
                  completion_ratio=0.5, shutin_days=[[1000.0, 0.0]]),
             dict(i=6, j=4, k=1, type="producer", control="ORAT", value=1000.0, minimum_bhp=4100.0, wellbore_radius=0.09525,
                  completion_ratio=0.5, shutin_days=[[1000.0, 0.0]])]
    spec = srm.PhysicsSpec(D=D, H=H, W=W, wells=srm.config.wells_from_connections(conns), fluid_type="GC")
    cols = O.load_pvt_table(U.GOLDEN + "/pvt_table.npz")
    otab = O.build_spline_table(cols, O.GC_PROPS)
    tabs = srm.pvt.SplineTables(knots=otab.c, w=otab.w, v=otab.v, order=1, properties=srm.pvt.GC_PROPERTIES)
    eng = srm.SrmPhysics(spec, tabs)
    pn, sn, gn = PressureNet(), StepNet(), SatNet()
    pn_g, sn_g, gn_g = (copy.deepcopy(m).cuda() for m in (pn, sn, gn))
    weights = {"dom": 1.0, "ibc": 1.0, "mbc": 1.0, "tde": 0.0, "cmbc": 1.0}
    loss = srm.PhysicsLoss(pn_g, srm.PVTLayer(eng, fluid_type="GC"), sn_g, srm.WellRatesPressure(eng, fluid_type="GC"),
                           saturation_model=gn_g, weights=weights)
    assert loss.trainable_models == [pn_g, sn_g, gn_g] and set(loss.loss_keys) == {"gas", "oil"}
    x = features(B, D, H, W, 6)
    wmse, wmse_grad, wsse, error_count, y_model = loss.pinn_batch_sse_grad(x.cuda(), None)
    assert len(wmse) == 2 and len(wmse_grad) == 3 and [len(g) for g in wmse_grad] == [4, 2, 2]
    ocfg = O.OracleConfig(D=D, H=H, W=W, wells=[O.Well(i=c["i"], j=c["j"], k=c["k"], value=c["value"]) for c in conns])
    oterms, ograds = oracle_chain_gc(pn, sn, gn, x, ocfg, otab, [1.0, 1.0, 1.0, 0, 0, 0, 0, 1.0])
    got = wsse[0].cpu().numpy()
    keys = loss.loss_keys["gas"]
    for name, slot in (("dom", 0), ("ibc", 1), ("mbc", 2)):
        assert np.isclose(got[keys.index(name)], oterms[slot], rtol=2e-4), name
    # the truncation term is rounding residue of the cell masses (its bracket vanishes identically): the networks'
    # outputs differ in the last bit between the CPU and the CUDA tanh/sigmoid, which re-rolls it
    assert np.isclose(got[keys.index("cmbc")], oterms[7], rtol=5e-2)
    flat = [g for gs in wmse_grad for g in gs]
    for a, b in zip(flat, ograds):
        a = a.cpu().numpy()
        assert np.abs(a - b).max() <= 5e-4 * np.abs(b).max() + 1e-30, (np.abs(a - b).max(), np.abs(b).max())
    # the well-model mirror: dense component rates, zero off-well
    sg = gn_g(x.cuda())
    (qgg, qgo, qoo, qog), pwf = loss.well_rate_bhp_model.compute_rates_and_bhp(x.cuda(), pn_g(x.cuda()), sg, None, None)
    assert qgg.shape == (B, D, H, W, 1) and int((qgg[..., 0] != 0).sum()) <= 2 * B and float(qgg.max()) <= 1000.0 + 1e-3
