"""forward + adjoint of one synthetic batch, a few times: the command ncu wraps (tools/gpu_prof.sh)
     python tools/prof_step.py <workload> <numerics> [K] [iters]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import srm_b200 as srm  # noqa: E402

name, numerics = sys.argv[1], sys.argv[2]
c, spec = bench.workload(name)
K = int(sys.argv[3]) if len(sys.argv) > 3 else c["K"]
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
gc = spec.fluid_type == "GC"
tabs = srm.build_spline_tables(srm.load_default_pvt_table(), srm.pvt.GC_PROPERTIES if gc else srm.pvt.DG_PROPERTIES, order=1)
eng = srm.SrmPhysics(spec, tabs, device=0, numerics=numerics, pvt_lut=(numerics == "reference"))
b = srm.synth.make_batch(spec.W, spec.H, spec.D, c["T"], K, [(w.i, w.j) for w in spec.wells[:8]], seed=2002, device="cuda")
d = dict(kx=b.kx, sample_real=b.sample_real, p0=b.p0, p1=b.p1, dt1=b.dt1, dt2=b.dt2, t1=b.t1)
if gc:
    d["sg0"], d["sg1"], d["so0"], d["so1"] = srm.synth.make_saturations(b, seed=2002)
w = torch.tensor([1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0] if gc else bench.WEIGHTS, dtype=torch.float32, device="cuda")
fwd, bwd = (eng.forward_gc, eng.backward_gc) if gc else (eng.forward, eng.backward)
for _ in range(iters):
    fw = fwd(**d)
    g = bwd(dterms=w, **d)
torch.cuda.synchronize()
print("terms", fw["terms"][0].tolist())
