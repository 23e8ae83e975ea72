"""Synthetic inputs of the BASELINE.json shapes.

Nothing here is on the measured path: it manufactures the tensors the networks of the reference
would hand to the physics loss (p0, p1, dt1, dt2) plus the static permeability realisations, with
the value ranges the reference's defaults imply (default_configurations.py:92-140):
log-normal kx (mean 3 mD, std 1.5, correlation 0.2*L, clipped to [0.26, 24] mD), pressure between
pwf_min = 4100 and Pi = 5000 psi with draw-down cones on the wells, time steps in [0.1, 10] days.
Runs on CPU (tests) or directly on the GPU (bench) -- torch ops only.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

# BASELINE.json configs -> (W, H, D, T, K)
CONFIGS = {
    "cfg1": dict(W=39, H=39, D=1, T=8, K=4),          # default grid, batch 32
    "cfg2": dict(W=64, H=64, D=16, T=32, K=16),
    "cfg3": dict(W=128, H=128, D=32, T=24, K=64),
    "cfg4": dict(W=128, H=128, D=32, T=24, K=32),     # two-phase (gas condensate)
    "cfg5": dict(W=256, H=256, D=64, T=32, K=8),
}


@dataclass
class SynthBatch:
    kx: torch.Tensor            # (R, D, H, W) physical mD
    p0: torch.Tensor            # (B, D, H, W)
    p1: torch.Tensor            # (B, D, H, W)
    dt1: torch.Tensor           # (B,)
    dt2: torch.Tensor           # (B,)
    t1: torch.Tensor            # (B,) time at level n+1, days
    sample_real: torch.Tensor   # (B,) int32, realisation-major: b = r*T + t


def _gauss_smooth(x, sigma_cells):
    """Separable Gaussian smoothing with edge replication along the last three axes of (N,D,H,W)."""
    out = x
    for axis, sig in zip((-1, -2, -3), sigma_cells):
        n = out.shape[axis]
        if n == 1 or sig <= 0:
            continue
        rad = max(1, min(int(3 * sig), n - 1, 48))
        k = torch.arange(-rad, rad + 1, device=x.device, dtype=x.dtype)
        w = torch.exp(-0.5 * (k / sig) ** 2)
        w = w / w.sum()
        o = out.movedim(axis, -1)
        shp = o.shape
        o = o.reshape(-1, 1, n)
        o = torch.nn.functional.pad(o, (rad, rad), mode="replicate")
        o = torch.nn.functional.conv1d(o, w.view(1, 1, -1))
        out = o.reshape(shp).movedim(-1, axis)
    return out


def make_kx(R, D, H, W, gen, device, mean=3.0, std=1.5, corr=0.2, kmin=0.26, kmax=24.0):
    sig_log = math.sqrt(math.log(1.0 + (std / mean) ** 2))     # KL_expansion.py:83-84
    mu_log = math.log(mean) - 0.5 * sig_log ** 2
    z = torch.randn((R, D, H, W), generator=gen, device=device, dtype=torch.float32)
    z = _gauss_smooth(z, (corr * W / 2.0, corr * H / 2.0, corr * D / 2.0))
    z = (z - z.mean(dim=(1, 2, 3), keepdim=True)) / (z.std(dim=(1, 2, 3), keepdim=True) + 1e-12)
    kx = torch.exp(mu_log + sig_log * z).clamp_(kmin, kmax)
    return kx.contiguous()


def make_batch(W, H, D, T, K, wells_ij, seed=2000, device="cpu", near_knots=None) -> SynthBatch:
    """K realisations x T time points, realisation-major sample order (b = r*T + t).

    wells_ij: list of (i, j) cell coordinates for the draw-down cones.
    near_knots: optional 1-D tensor of PVT knots; if given, a sprinkling of cells is placed within
    +-1 psi of the knots in [3700, 5000] (the parity set of SURVEY 8(d))."""
    dev = torch.device(device)
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed))
    B = K * T
    kx = make_kx(K, D, H, W, gen, dev)
    jj = torch.arange(H, device=dev, dtype=torch.float32).view(1, 1, H, 1)
    ii = torch.arange(W, device=dev, dtype=torch.float32).view(1, 1, 1, W)
    draw = torch.zeros((B, D, H, W), device=dev, dtype=torch.float32)
    rad = 0.12 * max(W, H)
    for (wi, wj) in wells_ij:
        depth = 50.0 + 550.0 * torch.rand((B, 1, 1, 1), generator=gen, device=dev)
        cone = torch.exp(-(((ii - wi) ** 2 + (jj - wj) ** 2) / (2.0 * rad * rad)))
        draw = draw + depth * cone
    # global depletion growing with time index + smooth background
    tfrac = (torch.arange(B, device=dev) % T).to(torch.float32).view(B, 1, 1, 1) / max(T - 1, 1)
    draw = draw * (0.25 + 0.75 * tfrac) + 40.0 * tfrac
    draw = draw.clamp_(0.0, 880.0)
    noise = 2.0 * torch.randn((B, D, H, W), generator=gen, device=dev)
    p0 = (5000.0 - draw + noise).clamp_(4105.0, 5000.0)
    dec = torch.rand((B, D, H, W), generator=gen, device=dev)
    dec = _gauss_smooth(dec, (2.0, 2.0, 1.0)) * 30.0
    p1 = (p0 - dec).clamp_(4101.0, 5000.0)
    if near_knots is not None:
        kn = near_knots.to(dev, torch.float32)
        kn = kn[(kn >= 3700.0) & (kn <= 5000.0)]
        n = p0.numel()
        m = max(1, n // 50)
        idx = torch.randint(0, n, (m,), generator=gen, device=dev)
        which = torch.randint(0, kn.numel(), (m,), generator=gen, device=dev)
        off = (torch.rand((m,), generator=gen, device=dev) * 2.0 - 1.0)
        p0.view(-1)[idx] = kn[which] + off
        idx1 = torch.randint(0, n, (m,), generator=gen, device=dev)
        which1 = torch.randint(0, kn.numel(), (m,), generator=gen, device=dev)
        off1 = (torch.rand((m,), generator=gen, device=dev) * 2.0 - 1.0)
        p1.view(-1)[idx1] = kn[which1] + off1
    dt1 = 0.1 + 9.9 * torch.rand((B,), generator=gen, device=dev)
    dt2 = 0.1 + 9.9 * torch.rand((B,), generator=gen, device=dev)
    t0 = torch.linspace(0.0, 365.0, T, device=dev).repeat(K)
    t1 = t0 + dt1
    sample_real = (torch.arange(B, device=dev) // T).to(torch.int32)
    return SynthBatch(kx=kx, p0=p0.contiguous(), p1=p1.contiguous(), dt1=dt1, dt2=dt2, t1=t1,
                      sample_real=sample_real)


def make_saturations(batch: SynthBatch, seed=2000, Swmin=0.22, sg_lo=0.50, sg_hi=0.78):
    """Gas / oil saturations at both time levels for the two-phase (GC) configs: smooth Sg in [sg_lo, sg_hi]
    (Sgc ... 1-Swmin), Sg1 slightly below Sg0 (liquid drop-out as pressure falls), So = 1-Swmin-Sg
    (relative_permeability.py:58).  Returns sg0, sg1, so0, so1 shaped like batch.p0."""
    dev = batch.p0.device
    gen = torch.Generator(device=dev)
    gen.manual_seed(int(seed) + 77)
    u = torch.rand(batch.p0.shape, generator=gen, device=dev)
    u = _gauss_smooth(u, (2.0, 2.0, 1.0))
    u = (u - u.amin()) / (u.amax() - u.amin() + 1e-12)
    sg0 = (sg_lo + (sg_hi - sg_lo) * u).to(torch.float32)
    drop = 0.02 * torch.rand(batch.p0.shape, generator=gen, device=dev)
    sg1 = (sg0 - drop).clamp_(0.0, 1.0 - Swmin)
    top = torch.tensor(float(1.0 - Swmin), dtype=torch.float32, device=dev)
    return sg0.contiguous(), sg1.contiguous(), (top - sg0).contiguous(), (top - sg1).contiguous()
