"""time forward + adjoint of one resident synthetic batch under several builds of the library, in ONE process:
     python tools/tune.py <workload> <numerics> <K> <steps> lib1.so [lib2.so ...]
prints one line per library: fwd / bwd / step ms, fraction of the HBM roofline (28 + 8/T bytes per cell), loss terms"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import srm_b200 as srm  # noqa: E402

name, numerics, K, steps = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
libs = sys.argv[5:]
c, spec = bench.workload(name)
cones = [(w.i, w.j) for w in spec.wells[:8]]
if os.environ.get("TUNE_D"):            # the same x-y grid with another depth (per-CTA fixed costs against the column length)
    import dataclasses
    spec = dataclasses.replace(spec, D=int(os.environ["TUNE_D"]), wells=[w for w in spec.wells if w.k < int(os.environ["TUNE_D"])])
if os.environ.get("TUNE_LATTICE"):      # a 4 x 8 lattice completed in every layer instead of the config's connections
    import dataclasses
    spec = dataclasses.replace(spec, wells=srm.config.lattice_wells(spec.W, spec.H, spec.D))
    cones = [(w.i, w.j) for w in spec.wells[::spec.D][:8]]
if os.environ.get("TUNE_NOWELLS"):      # the same grid and pressures without connections: what the well code costs
    import dataclasses
    spec = dataclasses.replace(spec, wells=[])
gc = spec.fluid_type == "GC"
tabs = srm.build_spline_tables(srm.load_default_pvt_table(), srm.pvt.GC_PROPERTIES if gc else srm.pvt.DG_PROPERTIES, order=1)
b = srm.synth.make_batch(spec.W, spec.H, spec.D, c["T"], K, cones, seed=2002, device="cuda")
d = dict(kx=b.kx, sample_real=b.sample_real, p0=b.p0, p1=b.p1, dt1=b.dt1, dt2=b.dt2, t1=b.t1)
if gc:
    d["sg0"], d["sg1"], d["so0"], d["so1"] = srm.synth.make_saturations(b, seed=2002)
w = torch.tensor([1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0] if gc else bench.WEIGHTS, dtype=torch.float32, device="cuda")
N = b.p0.numel()
peak, _ = bench.peak_hbm()
ab = bench.alg_bytes_per_cell(c["T"], gc)
first = None
for lib in libs:
    srm._lib._lib = None
    srm._lib.LIB_PATH = os.path.abspath(lib)      # load_library() caches the library of LIB_PATH only
    srm._lib.load_library()
    try:
        eng = srm.SrmPhysics(spec, tabs, device=0, numerics=numerics, pvt_lut=(numerics == "reference"))
        fwd, bwd = (eng.forward_gc, eng.backward_gc) if gc else (eng.forward, eng.backward)
        out = None
        for _ in range(3):
            fw = fwd(**d)
            out = bwd(dterms=w, out=out, **d)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2 * steps + 1)]
        ev[0].record()
        for i in range(steps):
            fw = fwd(**d)
            ev[2 * i + 1].record()
            out = bwd(dterms=w, out=out, **d)
            ev[2 * i + 2].record()
        torch.cuda.synchronize()
        f = sum(ev[2 * i].elapsed_time(ev[2 * i + 1]) for i in range(steps)) / steps
        a = sum(ev[2 * i + 1].elapsed_time(ev[2 * i + 2]) for i in range(steps)) / steps
        terms = fw["terms"][0].tolist()
        chk = [float(t.double().abs().sum()) for t in out[:2]]
        if first is None:
            first = (terms, chk)
        same = all(abs(x - y) <= 1e-5 * max(abs(y), 1e-30) for x, y in zip(terms + chk, first[0] + first[1]))
        print(json.dumps({"lib": os.path.basename(lib), "fwd_ms": round(f, 4), "bwd_ms": round(a, 4), "step_ms": round(f + a, 4),
                          "frac": round(N * ab / ((f + a) * 1e-3) / 1e9 / peak, 4), "same_as_first": same,
                          "terms": [float("%.6g" % t) for t in terms[:4]]}), flush=True)
        eng.close()
        del eng, out, fw
        torch.cuda.empty_cache()
    except Exception as e:
        print(json.dumps({"lib": os.path.basename(lib), "error": repr(e)[:300]}), flush=True)
