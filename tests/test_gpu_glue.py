"""The fused glue either side of the physics kernels: HardLayer at both time levels and the
per-sample mean of the time-step field, through the C ABI (srm_glue_forward / srm_glue_backward), against the oracle
(oracle/srm_oracle.py: hard_layer_t, time_step_mean_t -- Hard_Layer_Subclassed.py:219-242, physics_loss.py:102,122).

Tolerances: alpha = alpha_t ** e is evaluated as 2^(e log2 alpha_t) with an fp64 logarithm per sample and a split
product (<= ~2 ulp against pow()); p = init - alpha*y then agrees to 1e-6 relative, cotangents to 1e-5 of their max.
"""
import os
import sys

import numpy as np
import pytest
import torch

import util as U

srm = U.srm
O = U.O
pytestmark = pytest.mark.gpu


def _case(W, H, D, B, seed, zero_time=False):
    ocfg, otab, spec, ptab, _ = U.make_case(W=W, H=H, D=D, T=2, K=1, seed=seed)
    g = torch.Generator().manual_seed(seed)
    y0 = 600.0 * torch.rand((B, D, H, W), generator=g)
    y1 = 600.0 * torch.rand((B, D, H, W), generator=g)
    expo = 0.1 + 0.89 * torch.rand((D, H, W), generator=g)
    tn0 = -1.0 + 1.9 * torch.rand(B, generator=g)
    if zero_time:
        tn0[0] = -1.0                                        # alpha_t = 0: the initial condition is enforced exactly
    tn1 = torch.clamp(tn0 + 0.05 * torch.rand(B, generator=g), max=1.0)
    dtf1 = 0.1 + 9.9 * torch.rand((B, D, H, W), generator=g)
    dtf2 = 0.1 + 9.9 * torch.rand((B, D, H, W), generator=g)
    return spec, ptab, dict(y0=y0, y1=y1, expo=expo, tn0=tn0, tn1=tn1, dtf1=dtf1, dtf2=dtf2)


def _oracle(d, init_value, w):
    t = {k: v.clone().requires_grad_(k in ("y0", "y1", "expo", "dtf1", "dtf2")) for k, v in d.items()}
    p0 = O.hard_layer_t(t["y0"], t["tn0"], t["expo"], init_value)
    p1 = O.hard_layer_t(t["y1"], t["tn1"], t["expo"], init_value)
    dt1, dt2 = O.time_step_mean_t(t["dtf1"]), O.time_step_mean_t(t["dtf2"])
    loss = (w["p0"] * p0).sum() + (w["p1"] * p1).sum() + (w["dt1"] * dt1).sum() + (w["dt2"] * dt2).sum()
    loss.backward()
    return dict(p0=p0.detach(), p1=p1.detach(), dt1=dt1.detach(), dt2=dt2.detach(), gy0=t["y0"].grad, gy1=t["y1"].grad,
                gexpo=t["expo"].grad, gdtf1=t["dtf1"].grad, gdtf2=t["dtf2"].grad)


@pytest.mark.parametrize("shape", [(64, 16, 4, 6), (39, 39, 1, 5), (6, 5, 3, 17)], ids=lambda s: "x".join(map(str, s)))
def test_glue_forward_backward_vs_oracle(shape):
    W, H, D, B = shape
    spec, ptab, d = _case(W, H, D, B, seed=3100 + W, zero_time=True)
    g = torch.Generator().manual_seed(7)
    w = dict(p0=torch.randn((B, D, H, W), generator=g), p1=torch.randn((B, D, H, W), generator=g),
             dt1=torch.randn(B, generator=g), dt2=torch.randn(B, generator=g))
    init_value = 5000.0
    o = _oracle(d, init_value, w)
    eng = srm.SrmPhysics(spec, ptab, device=0)
    dev = eng.device
    c = {k: v.to(dev) for k, v in d.items()}
    p0, p1, dt1, dt2 = eng.glue_forward(c["y0"], c["y1"], c["tn0"], c["tn1"], c["expo"], c["dtf1"], c["dtf2"], init_value)
    gy0, gy1, gexpo, gdtf1, gdtf2 = eng.glue_backward(c["y0"], c["y1"], c["tn0"], c["tn1"], w["p0"].to(dev), w["p1"].to(dev),
                                                      c["expo"], w["dt1"].to(dev), w["dt2"].to(dev), init_value)
    torch.cuda.synchronize()
    for name, a in (("p0", p0), ("p1", p1), ("dt1", dt1), ("dt2", dt2)):
        assert np.allclose(a.cpu().numpy(), o[name].numpy(), rtol=1e-6, atol=0), name
    assert np.all(p0[0].cpu().numpy() == np.float32(init_value))            # alpha_t = 0
    for name, a in (("gy0", gy0), ("gy1", gy1), ("gexpo", gexpo), ("gdtf1", gdtf1), ("gdtf2", gdtf2)):
        assert np.all(np.isfinite(a.cpu().numpy())), name
        assert U.rel_to_max(a.cpu().numpy(), o[name].numpy()) < 1e-5, (name, U.rel_to_max(a.cpu().numpy(), o[name].numpy()))
    # levels without a time-step field: no mean, no workspace
    q0, q1, n1, n2 = eng.glue_forward(c["y0"], c["y1"], c["tn0"], c["tn1"], c["expo"], None, None, init_value)
    assert n1 is None and n2 is None and torch.equal(q0, p0) and torch.equal(q1, p1)
    eng.close()


class _Net(torch.nn.Module):
    def __init__(self, seed):
        super().__init__()
        torch.manual_seed(seed)
        self.lin = torch.nn.Linear(5, 1)

    def forward(self, x):
        return 300.0 * torch.sigmoid(self.lin(x))


def test_complete_trainable_module_matches_torch_reference():
    """CompleteTrainableModule(main_network, HardLayer) -- complete_trainable_module.py:142-176 -- output and the
    gradients of the network weights and of kernel_exponent against the same module written in plain torch"""
    ocfg, otab, spec, ptab, batch = U.make_case(W=12, H=7, D=3, T=3, K=2, seed=3201)
    eng = srm.SrmPhysics(spec, ptab, device=0)
    dev = eng.device
    B = 6
    g = torch.Generator().manual_seed(11)
    x = (2.0 * torch.rand((B, 3, 7, 12, 5), generator=g) - 1.0)
    x[..., 3] = (2.0 * torch.rand(B, generator=g) - 1.0).view(B, 1, 1, 1)
    net = _Net(5).to(dev)
    hl = srm.HardLayer(eng, init_value=5000.0, kernel_exponent_config={"initial_value": (0.5,), "min_value": 0.1, "max_value": 1.0})
    with torch.no_grad():
        hl.kernel_exponent.add_(0.3 * torch.rand(hl.kernel_exponent.shape, generator=g).to(dev))
    mod = srm.CompleteTrainableModule(net, hl)
    wgt = torch.randn((B, 3, 7, 12, 1), generator=g).to(dev)
    out = mod(x.to(dev))
    (out * wgt).sum().backward()
    got = dict(out=out.detach().cpu(), ge=hl.kernel_exponent.grad.cpu(), gw=net.lin.weight.grad.cpu().clone())
    # plain torch twin
    net2 = _Net(5)
    e2 = hl.kernel_exponent.detach().cpu().clone().requires_grad_(True)
    y2 = net2(x)[..., 0]
    ref = O.hard_layer_t(y2, x[:, 0, 0, 0, 3], e2, 5000.0).unsqueeze(-1)
    (ref * wgt.cpu()).sum().backward()
    assert np.allclose(got["out"].numpy(), ref.detach().numpy(), rtol=1e-6, atol=0)
    assert U.rel_to_max(got["ge"].numpy(), e2.grad.numpy()) < 1e-5
    assert U.rel_to_max(got["gw"].numpy(), net2.lin.weight.grad.numpy()) < 1e-4        # fp32 reductions over B*N terms
    hl.apply_constraint()
    assert float(hl.kernel_exponent.detach().max()) <= 1.0 and float(hl.kernel_exponent.detach().min()) >= 0.1
    eng.close()


def test_physics_loss_with_fused_glue_equals_unfused():
    """PhysicsLoss.pinn_batch_sse_grad takes the fused glue when the pressure model is a CompleteTrainableModule with
    a HardLayer; the loss terms equal those of the same model evaluated op by op."""
    ocfg, otab, spec, ptab, batch = U.make_case(W=12, H=8, D=2, T=2, K=2, seed=3301)
    eng = srm.SrmPhysics(spec, ptab, device=0, pvt_lut=False)
    dev = eng.device
    B = 4
    g = torch.Generator().manual_seed(13)
    x = (2.0 * torch.rand((B, 2, 8, 12, 5), generator=g) - 1.0)
    x[..., 3] = (-0.9 + 1.5 * torch.rand(B, generator=g)).view(B, 1, 1, 1)
    x = x.to(dev)

    class Step(torch.nn.Module):
        def __init__(self):
            super().__init__()
            torch.manual_seed(3)
            self.lin = torch.nn.Linear(5, 1)

        def forward(self, x):
            return 0.1 + 9.9 * torch.sigmoid(self.lin(x))

    class Unfused(torch.nn.Module):          # same arithmetic, plain torch ops (not a CompleteTrainableModule)
        def __init__(self, net, hl):
            super().__init__()
            self.net, self.hl = net, hl

        def forward(self, x):
            y = self.net(x)[..., 0]
            return O.hard_layer_t(y, x[:, 0, 0, 0, 3], self.hl.kernel_exponent, self.hl.init_value).unsqueeze(-1)

    net, step = _Net(9).to(dev), Step().to(dev)
    hl = srm.HardLayer(eng, init_value=5000.0)
    pvt = srm.PVTLayer(eng)
    wells = srm.WellRatesPressure(eng)
    res = []
    for model in (srm.CompleteTrainableModule(net, hl), Unfused(net, hl)):
        loss = srm.PhysicsLoss(model, pvt, step, wells)
        wmse, grads, wsse, cnt, y_model = loss.pinn_batch_sse_grad(x)
        res.append((wsse[0].cpu().numpy(), [gg.cpu().numpy() for gl in grads for gg in gl]))
    # the truncation term (last slot) is rounding noise around an analytic zero: it re-rolls with any 1-ulp change of p
    assert np.allclose(res[0][0], res[1][0], rtol=2e-5, atol=1e-9 * float(np.abs(res[1][0]).max()))
    # The weight gradients are NOT compared across the two paths: in reference-order fp32 dL/dp0 is dominated by the
    # rounding noise of the spline's second derivative (18-46 x max against fp64), so the 1-ulp differences
    # between exp2-based and libm pow re-roll it.  The glue's own cotangents are pinned by the tests above, the physics
    # kernels' by test_gpu_parity.py on identical inputs; here: same structure, finite, non-trivial.
    assert len(res[0][1]) == len(res[1][1]) == 5                      # net (w, b), kernel_exponent, time-step net (w, b)
    for a, b in zip(res[0][1], res[1][1]):
        assert a.shape == b.shape and np.all(np.isfinite(a)) and np.abs(a).max() > 0
    eng.close()


def test_glue_against_the_reference_hard_layer_class():
    """srm_glue_forward / srm_glue_backward against values and cotangents produced by the reference's OWN HardLayer class
    (tests/golden/make_reference_hardlayer_golden.py): values 1e-6 relative (2^(e log2 a) vs libm pow), cotangents 1e-5."""
    g = np.load(os.path.join(U.GOLDEN, "reference_hardlayer.npz"))
    B, D, H, W = g["y"].shape
    ocfg, otab, spec, ptab, _ = U.make_case(W=W, H=H, D=D, T=1, K=1, seed=1)
    eng = srm.SrmPhysics(spec, ptab, device=0)
    dev = eng.device
    c = lambda k: torch.as_tensor(g[k]).to(dev).contiguous()
    p0, p1, _, _ = eng.glue_forward(c("y"), c("y"), c("tn"), c("tn"), c("expo"), None, None, 5000.0)
    assert np.allclose(p0.cpu().numpy(), g["out"], rtol=1e-6, atol=0) and torch.equal(p0, p1)
    zero = torch.zeros_like(c("wgt"))
    gy0, gy1, gexpo, _, _, gtn0, gtn1 = eng.glue_backward(c("y"), c("y"), c("tn"), c("tn"), c("wgt"), zero, c("expo"), None, None, 5000.0,
                                                          want_gtn=(True, True))
    assert U.rel_to_max(gy0.cpu().numpy(), g["gy"]) < 1e-5 and U.rel_to_max(gexpo.cpu().numpy(), g["gexpo"]) < 1e-5
    assert not gy1.any()
    # the cotangent of the layer's time input (the path to the time-step model, physics_loss.py:105-111); sample 0 sits at
    # alpha_t = 0 where the reference's own gradient is 0 * inf
    assert U.rel_to_max(gtn0.cpu().numpy()[1:], g["gtn"][1:]) < 1e-5
    assert not gtn1.cpu().numpy()[1:].any()
    eng.close()


def test_hard_layer_options_against_the_reference_class():
    """The mirror's HardLayer with every non-default option on -- use_rbf (Dense(1, sigmoid) on the property channel),
    the gas-condensate rectifier on a third input, activations on the kernel exponent and on the network output --
    against values and cotangents of the reference's OWN class with the same options (Hard_Layer_Subclassed.py:41-45,
    168-187, 219-246; tests/golden/make_reference_hardlayer_golden.py: reference_hardlayer_opts.npz)."""
    g = np.load(os.path.join(U.GOLDEN, "reference_hardlayer_opts.npz"))
    B, D, H, W = g["y"].shape
    ocfg, otab, spec, ptab, _ = U.make_case(W=W, H=H, D=D, T=1, K=1, seed=1)
    eng = srm.SrmPhysics(spec, ptab, device=0)
    dev = eng.device
    hl = srm.HardLayer(eng, norm_limits=[-1, 1], init_value=float(g["init_value"]),
                       kernel_exponent_config={"initial_value": (0.5,), "trainable": True, "min_value": 0.1, "max_value": 1.0},
                       use_rbf=True, rbf_config={"output_dim": 25, "activation": "sigmoid"}, rectifier=torch.relu,
                       kernel_activation=[torch.sigmoid], input_activation=torch.tanh, pdew=float(g["pdew"]), pmin=float(g["pmin"]))
    with torch.no_grad():
        hl.kernel_exponent.copy_(torch.as_tensor(g["expo"]).to(dev))
        hl.rbf_dense.weight.fill_(float(g["kernel"]))
        hl.rbf_dense.bias.fill_(float(g["bias"]))
    t5 = lambda k: torch.as_tensor(g[k]).to(dev).unsqueeze(-1)
    tn = torch.as_tensor(g["tn"]).to(dev).requires_grad_(True)
    time = tn.view(B, 1, 1, 1, 1).expand(B, D, H, W, 1)
    y, rect = t5("y").requires_grad_(True), t5("rect").requires_grad_(True)
    out = hl([[time, t5("prop")], y, rect])
    assert np.allclose(out.detach().cpu().numpy()[..., 0], g["out"], rtol=2e-6, atol=0)
    gy, ge, gt, gk, gb, gr = torch.autograd.grad((out * t5("wgt")).sum(), [y, hl.kernel_exponent, tn, hl.rbf_dense.weight, hl.rbf_dense.bias, rect])
    for got, key in ((gy[..., 0], "gy"), (ge, "gexpo"), (gt, "gtn"), (gk.reshape(1, 1), "gkernel"), (gb, "gbias"), (gr[..., 0], "grect")):
        assert U.rel_to_max(got.cpu().numpy(), g[key]) < 1e-5, key
    assert (g["rect"] < float(g["pdew"])).any() and (g["rect"] > float(g["pdew"])).any()      # both sides of the dew point
    eng.close()


def test_fused_two_level_carries_the_time_cotangent_to_the_time_step_model():
    """d p1 / d tn1 = -e alpha_t^(e-1) y1 / (t_hi - t_lo) must reach the time-step model through x1's time
    channel.  fused_two_level (one CUDA pass) against the same graph in plain torch ops on the same device."""
    ocfg, otab, spec, ptab, batch = U.make_case(W=12, H=8, D=2, T=2, K=2, seed=3302)
    eng = srm.SrmPhysics(spec, ptab, device=0, pvt_lut=False)
    dev = eng.device
    B = 4
    g = torch.Generator().manual_seed(17)
    x0 = (2.0 * torch.rand((B, 2, 8, 12, 5), generator=g) - 1.0)
    x0[..., 3] = (-0.8 + 1.2 * torch.rand(B, generator=g)).view(B, 1, 1, 1)
    x0 = x0.to(dev)
    net = _Net(21).to(dev)
    hl = srm.HardLayer(eng, init_value=5000.0)
    with torch.no_grad():
        hl.kernel_exponent.copy_(0.2 + 0.7 * torch.rand(hl.kernel_exponent.shape, generator=g).to(dev))
    mod = srm.CompleteTrainableModule(net, hl)
    step = torch.nn.Linear(5, 1).to(dev)
    w0 = torch.randn((B, 2, 8, 12), generator=g).to(dev)
    w1 = torch.randn((B, 2, 8, 12), generator=g).to(dev)
    idx = torch.zeros(5, device=dev); idx[3] = 1.0
    res = []
    for fused in (True, False):
        for m in (net, step, hl):
            m.zero_grad()
        dn = 0.01 * torch.sigmoid(step(x0)).reshape(B, -1).mean(dim=1)
        x1 = x0 + dn.view(B, 1, 1, 1, 1) * idx
        if fused:
            p0, p1, dt2 = srm.hard_layer.fused_two_level(mod, lambda x: torch.sigmoid(step(x)), x0, x1)
        else:
            e = hl.kernel_exponent
            p0 = O.hard_layer_t(net(x0)[..., 0], x0[:, 0, 0, 0, 3], e, 5000.0)
            p1 = O.hard_layer_t(net(x1)[..., 0], x1[:, 0, 0, 0, 3], e, 5000.0)
            dt2 = torch.sigmoid(step(x1))[..., 0].reshape(B, -1).mean(dim=1)
        ((p0 * w0).sum() + (p1 * w1).sum() + dt2.sum()).backward()
        res.append([step.weight.grad.clone().cpu().numpy(), step.bias.grad.clone().cpu().numpy(),
                    net.lin.weight.grad.clone().cpu().numpy(), hl.kernel_exponent.grad.clone().cpu().numpy()])
    for a, b in zip(*res):
        assert U.rel_to_max(a, b) < 1e-4, (a, b)          # fp32 reductions over B*N terms in different orders
    assert np.abs(res[0][0]).max() > 0
    eng.close()


@pytest.mark.parametrize("shape", [(3, 2, 6, 8, 5), (4, 1, 39, 39, 5), (2, 2, 5, 6, 7), (2, 1, 4, 4, 3)])   # vector rows, ragged rows (39 x 39 x 5), C != 5, C < 4
def test_feature_glue_time_shift_and_permeability_channel(shape):
    """srm_features_forward / _backward against the reference's torch-side construction (physics_loss.py:105-110:
    zeros_like + strided assign + add) and against srm_denormalize_log on the strided channel: bit for bit."""
    ocfg, otab, spec, ptab, batch = U.make_case(W=8, H=6, D=2, T=2, K=1, seed=5)
    eng = srm.SrmPhysics(spec, ptab, device=0)
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.rand(shape, generator=g, device="cuda") * 2.0 - 1.0
    dn = torch.rand(shape[0], generator=g, device="cuda") * 0.05
    tc, kc = (3, 4) if shape[-1] >= 5 else (1, 2)
    x1, kx = eng.features_forward(x, dn, (0.26, 24.0), -1.0, 1.0, t_channel=tc, k_channel=kc)
    shift = torch.zeros_like(x)
    shift[..., tc] = dn.view(-1, 1, 1, 1)
    assert torch.equal(x1, x + shift)
    assert torch.equal(kx, eng.denormalize_log(x[..., kc].contiguous(), 0.26, 24.0, -1.0, 1.0))
    assert eng.features_forward(x, None, (0.26, 24.0), t_channel=tc, k_channel=kc)[0] is None   # permeability only
    gx1 = torch.randn(shape, generator=g, device="cuda")
    gdn = eng.features_backward(gx1, t_channel=tc)
    ref = gx1[..., tc].double().sum(dim=(1, 2, 3)).float()
    assert torch.allclose(gdn, ref, rtol=1e-5, atol=1e-6)
    eng.close()


def test_shift_time_autograd_function_matches_torch():
    _ShiftTimeFn = srm.physics_loss._ShiftTimeFn
    ocfg, otab, spec, ptab, batch = U.make_case(W=8, H=6, D=2, T=2, K=1, seed=5)
    eng = srm.SrmPhysics(spec, ptab, device=0)
    g = torch.Generator(device="cuda").manual_seed(4)
    x = (torch.rand((3, 2, 6, 8, 5), generator=g, device="cuda") * 2 - 1).requires_grad_(True)
    dn = (torch.rand(3, generator=g, device="cuda") * 0.05).requires_grad_(True)
    wgt = torch.randn((3, 2, 6, 8, 5), generator=g, device="cuda")
    (_ShiftTimeFn.apply(eng, x, dn) * wgt).sum().backward()
    gx, gdn = x.grad.clone(), dn.grad.clone()
    x.grad = None; dn.grad = None
    shift = torch.zeros_like(x)
    idx = torch.zeros(5, device="cuda"); idx[3] = 1.0
    ((x + dn.view(-1, 1, 1, 1, 1) * idx) * wgt).sum().backward()
    assert torch.equal(gx, x.grad) and torch.allclose(gdn, dn.grad, rtol=1e-5, atol=1e-6)
    eng.close()
