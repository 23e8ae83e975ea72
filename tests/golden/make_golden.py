#!/usr/bin/env python
"""Generate the golden vectors of the physics-loss path from the oracle.

PARITY UNPINNED: the reference ships no golden vectors or assertions for this path (SURVEY.md
section 4) and cannot be executed here (TensorFlow absent), so these goldens are the *oracle's*
outputs (oracle/srm_oracle.py, a line-cited restatement of the reference arithmetic) on seeded
synthetic inputs.  They pin (a) the oracle against regressions and (b) the CUDA kernels against
the oracle on the GPU box, where neither /root/reference nor a rerun of this script is required.

Usage:  python tests/golden/make_golden.py        (writes tests/golden/*.npz)
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util as U  # noqa: E402

O = U.O

CASES = {
    # name: make_case kwargs
    "dg_2d_default": dict(W=39, H=39, D=1, T=2, K=2, seed=2101),
    "dg_3d_layers": dict(W=12, H=10, D=4, T=2, K=2, seed=2102, all_layers=True),
    "dg_3d_blocking": dict(W=10, H=12, D=3, T=2, K=1, seed=2103, all_layers=True, use_blocking_factor=True, n_intervals=8),
}


def pvt_golden():
    cols = O.load_pvt_table(os.path.join(HERE, "pvt_table.npz"))
    tab = O.build_spline_table(cols, O.DG_PROPS, order=1, lam=0.001)
    cfg = O.OracleConfig()
    rng = np.random.default_rng(2100)
    p = np.concatenate([
        np.linspace(4000.0, 5100.0, 1501),                 # operating window
        rng.uniform(10.0, 12000.0, 1500),                  # whole table, beyond the clamp
        (tab.c[:, None] + np.array([-1.0, -0.25, -1e-3, 0.0, 1e-3, 0.25, 1.0])[None, :]).reshape(-1),  # at the knots
        np.array([-50.0, 0.0, 14.7, 14.69, 10000.0, 10000.5, 25000.0]),                                 # clamp edges
    ]).astype(np.float32)
    ph = O.pvt_clamp(torch.from_numpy(p), cfg).numpy()
    out = dict(p=p, knots=tab.c, w=tab.w, v=tab.v)
    for q, name in enumerate(tab.names):
        val, d1, d2 = O.spline_eval_np(ph, tab, q, np.float32, need=2)
        out[f"val_{name}"], out[f"d1_{name}"], out[f"d2_{name}"] = val, d1, d2
    np.savez_compressed(os.path.join(HERE, "pvt_golden.npz"), **out)
    print("pvt_golden:", p.size, "pressures")


def case_golden(name, kw):
    ocfg, otab, spec, ptab, batch = U.make_case(**kw)
    o = U.oracle_run(ocfg, otab, batch)
    keep = dict(kx=batch.kx.numpy(), p0=batch.p0.numpy(), p1=batch.p1.numpy(), dt1=batch.dt1.numpy(),
                dt2=batch.dt2.numpy(), t1=batch.t1.numpy(), sample_real=batch.sample_real.numpy(),
                weights=np.array(U.WEIGHTS, dtype=np.float32))
    for k in ("dom", "ibc", "mbc", "tde", "q", "pwf", "qw", "pwfw", "terms", "gp0", "gp1", "gdt1", "gdt2"):
        keep["o_" + k] = np.asarray(o[k], dtype=np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **keep)
    print(name, "B =", batch.p0.shape[0], "terms", o["terms"][:4])


GC_CASES = {
    "gc_3d": dict(seed=2201, B=2, D=2, H=6, W=8, sg_lo=0.2, sg_hi=0.6, wells="dup"),
}
GC_WEIGHTS = [1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0]
GC_GRADS = ("gp0", "gp1", "gsg0", "gsg1", "gso0", "gso1", "gdt1", "gdt2")


def gc_golden(name, kw):
    """gas-condensate case: inputs, the spline weights of the 7 properties, the fp32 oracle's outputs and the
    element-wise distance of its gradients to the fp64 twin (the reference's own fp32 noise, see tests/test_gpu_gc.py)"""
    ocfg, otab, spec, ptab, d = U.gc_case(**kw)
    args = (ocfg, otab, d["kx"], d["p0"], d["p1"], d["sg0"], d["sg1"], d["so0"], d["so1"], d["dt1"], d["dt2"], d["t1"],
            d["sample_real"], GC_WEIGHTS)
    o = O.gc_forward_backward(*args)
    o64 = O.gc_forward_backward(*args, dtype=torch.float64)
    keep = dict(d)
    keep.update(knots=otab.c, w=otab.w, v=otab.v, weights=np.array(GC_WEIGHTS, dtype=np.float32))
    for k in ("dom", "ibc", "mbc", "cmbc", "terms", "qw4", "pwfw") + GC_GRADS:
        keep["o_" + k] = np.asarray(o[k], dtype=np.float32)
    for k in GC_GRADS[:-1]:
        keep["noise_" + k] = (o[k].astype(np.float64) - o64[k]).astype(np.float32)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **keep)
    print(name, "terms", o["terms"])


if __name__ == "__main__":
    for n, kw in GC_CASES.items():
        gc_golden(n, kw)
    pvt_golden()
    for n, kw in CASES.items():
        case_golden(n, kw)
