import sys, os, numpy as np, torch
sys.path.insert(0, "tests")
import util as U
import test_gpu_gc as T
O = U.O
for kw in T.CASES + [dict(seed=41, sg_lo=0.2, sg_hi=0.5, wells="dup", D=2, H=6, W=8)]:
    ocfg, otab, spec, ptab, d = T.gc_case(**kw)
    for mode in ("asis", "dpc>=1"):
        if mode == "dpc>=1":
            rng = np.random.default_rng(99)
            d["p1"] = (d["p0"] - rng.uniform(1, 25, d["p0"].shape)).astype(np.float32)
        o, c = T.run_both(ocfg, otab, spec, ptab, d)
        o64 = O.gc_forward_backward(ocfg, otab, d["kx"], d["p0"], d["p1"], d["sg0"], d["sg1"], d["so0"], d["so1"], d["dt1"],
                                    d["dt2"], d["t1"], d["sample_real"], T.W_ALL, dtype=torch.float64)
        line = [f"{kw} {mode} dom_ulp={U.ulp_diff(c['dom'], o['dom'])}"]
        for k in ("gp0", "gp1", "gsg0", "gsg1", "gso0", "gso1", "gdt1"):
            line.append(f"{k}: c-o32 {U.rel_to_max(c[k], o[k]):.1e} o32-o64 {U.rel_to_max(o[k], o64[k]):.1e} c-o64 {U.rel_to_max(c[k], o64[k]):.1e}")
        print("\n   ".join(line))
