"""closed-form mode against the fp64 oracle on a well next to the boundary: which loss term's gradient differs, lean vs generic kernels"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import util as U
srm = U.srm

def run(spec, ptab, batch, weights, misalign):
    dev = torch.device("cuda", 0)
    eng = srm.SrmPhysics(spec, ptab, device=0, numerics="closed_form")
    d = U.to_dev(batch, dev)
    if misalign:
        buf = torch.empty(d["p0"].numel() + 1, dtype=torch.float32, device=dev)
        v = buf[1:].view(d["p0"].shape); v.copy_(d["p0"]); d["p0"] = v
    fw = eng.forward(want_dom=True, want_wells=True, **d)
    dt = torch.tensor(weights, dtype=torch.float32, device=dev)
    gp0, gp1, gdt1, gdt2 = eng.backward(dterms=dt, **d)
    torch.cuda.synchronize()
    out = dict(gp1=gp1.cpu().numpy(), gp0=gp0.cpu().numpy(), dom=fw["dom"].cpu().numpy(), qw=fw["qw"].cpu().numpy())
    eng.close()
    return out

wl = U.crowded_wells(5, 1, duplicate=False)
ocfg, otab, spec, ptab, batch = U.make_case(W=64, H=20, D=5, T=2, K=1, seed=2006, wells=wl)
for wts in ([1, 0, 0, 0], [0, 1, 0, 0], [0, 0, 1, 0], [0, 0, 0, 1]):
    w8 = [float(x) for x in wts] + [0.0] * 4
    o64 = U.oracle_run(ocfg, otab, batch, weights=w8, dtype=torch.float64)
    lean = run(spec, ptab, batch, w8, False)
    gen = run(spec, ptab, batch, w8, True)
    for nm, c in (("lean", lean), ("generic", gen)):
        d = np.abs(np.asarray(c["gp1"], np.float64) - o64["gp1"])
        i = np.unravel_index(np.argmax(d), d.shape)
        print(wts, nm, "gp1 err/max %.3e" % (d.max() / max(np.abs(o64["gp1"]).max(), 1e-300)), "at", tuple(int(x) for x in i),
              "val", float(c["gp1"][i]), float(o64["gp1"][i]), flush=True)
        # the six neighbours and the well cells of sample 1, layer 1
    for (k, j, i) in [(1, 1, 2), (1, 1, 1), (1, 1, 3), (1, 0, 2), (1, 2, 2), (0, 1, 2), (2, 1, 2)]:
        print("   cell", (k, j, i), "lean %.6e generic %.6e o64 %.6e" % (lean["gp1"][1, k, j, i], gen["gp1"][1, k, j, i], o64["gp1"][1, k, j, i]))
