#!/usr/bin/env python
"""bench.py -- residual+adjoint throughput of the physics-loss path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload cfg2] [--numerics reference]

One "step" = one forward (residual + loss terms) plus one adjoint (dL/dp0, dL/dp1, dL/ddt1,
dL/ddt2) over one batch of synthetic input of the named BASELINE config.  Prints ONE JSON line.

  value      cell-timesteps/s, whole job, inputs resident in HBM, CUDA-event timed, max over ranks
  e2e        same metric through the public API with HOST (pinned) buffers: H2D of the step's
             inputs and D2H of the loss terms and gradients inside the timed region
  roofline   the dominant kernel (the adjoint pass): algorithmic bytes (16 + 4/T per cell-timestep) /
             its CUDA-event-timed duration vs the measured HBM copy peak (MEASURED_PEAKS.json, else the 6650 GB/s
             fallback of B200_PROFILING.md); `forward` (12 + 4/T) and `step` (28 + 8/T, whole step) beside it;
             `traffic` = dram bytes per launch from the committed ncu capture (profiles/traffic.json)
  cpu_baseline  the oracle (a port of the reference's TF op graph; TensorFlow is not installable
             here) on the box's host cores, bounded sample of the same workload

Default workload: cfg5 (256 x 256 x 64, T = 32, K = 8 realisations per GPU, 2048 well connections with the
blocking-factor integral) -- the grid BASELINE.json quotes its 70 %-of-HBM target on.  N > 1 (torchrun): weak scaling,
every rank runs the same per-GPU workload on its own realisations (cfg5's batch sweep: K = 8 N in total);
`--strong` shards the named config's K realisations over the ranks instead (cfg3: K = 64 over 1/2/4/8 GPUs).  The only
collective is the all-reduce of the 16-float loss-term vector (NCCL), queued behind the adjoint and waited on one step
later, so no rank waits for another inside a step.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np
import torch

METRIC = "residual+adjoint cell-timesteps/sec"
UNIT = "cell-timesteps/s"
WEIGHTS = [1.0, 1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0]


def workload(name):
    import srm_b200 as srm
    c = dict(srm.synth.CONFIGS[name])
    if name == "cfg5":
        wells = srm.config.lattice_wells(c["W"], c["H"], c["D"])
        blocking = True
    else:
        wells = srm.config.scaled_default_wells(c["W"], c["H"], c["D"])
        blocking = False
    spec = srm.PhysicsSpec(D=c["D"], H=c["H"], W=c["W"], wells=wells, use_blocking_factor=blocking, n_intervals=8,
                           fluid_type="GC" if name == "cfg4" else "DG")
    return c, spec


def alg_bytes_per_cell(T, gc=False):
    # DG fwd + adjoint 28 + 8/T; GC (p, Sg, So at two levels; one dom) 68 + 8/T: forward reads 6 fields and
    # writes dom (28), adjoint reads 6 fields + dom and writes 6 gradients (52), minus nothing shared
    # (SURVEY.md 8(d) counts 52 + 8/T with So derived from Sg; the reference's So is a separate network output)
    return (80.0 if gc else 28.0) + 8.0 / T


def measured_traffic(workload, numerics):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the two dominant kernels, from the committed
    `ncu --set full` capture of this command (profiles/traffic.json, written by tools/ncu_traffic.py)"""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        t = json.load(open(path))
        return t.get(f"{workload}:{numerics}")
    except Exception:
        return None


def peak_hbm():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """samples SM clock / throttle reasons of one GPU through NVML while the timed region runs"""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._halt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:      # NVML unavailable: report that instead of inventing numbers
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._halt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.02)

    def stop(self):
        self._halt.set()
        if self.is_alive():
            self.join(timeout=1.0)
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "note": "NVML unavailable"}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU baseline: the oracle on the host cores (threads over independent samples)
# ------------------------------------------------------------------------------------------------
def cpu_oracle_throughput(name, target_seconds=15.0, max_samples=None):
    import srm_oracle as O
    import srm_b200 as srm
    from concurrent.futures import ThreadPoolExecutor
    c, spec = workload(name)
    # grids beyond ~1 M cells: the CPU unit of work is a z-slab of the grid (same x-y extent, same arithmetic per cell,
    # the lattice wells completed in the slab's layers), so that one round of one-unit-per-core ends in seconds
    slab = None
    if spec.n_cells > (1 << 21):
        Ds = max(1, (1 << 20) // (spec.W * spec.H))
        if Ds < spec.D:
            slab = Ds
            wl = srm.config.lattice_wells(spec.W, spec.H, Ds) if name == "cfg5" else srm.config.scaled_default_wells(spec.W, spec.H, Ds)
            spec = srm.PhysicsSpec(D=Ds, H=spec.H, W=spec.W, wells=wl, use_blocking_factor=spec.use_blocking_factor,
                                   n_intervals=spec.n_intervals, fluid_type=spec.fluid_type)
    cores = os.cpu_count() or 1
    torch.set_num_threads(1)        # parallelism comes from sample shards, one per core
    cols = O.load_pvt_table(os.path.join(ROOT, "tests", "golden", "pvt_table.npz"))
    gc = spec.fluid_type == "GC"
    otab = O.build_spline_table(cols, O.GC_PROPS if gc else O.DG_PROPS, order=1, lam=0.001)
    ocfg = O.OracleConfig(D=spec.D, H=spec.H, W=spec.W, use_blocking_factor=spec.use_blocking_factor,
                          n_intervals=spec.n_intervals,
                          wells=[O.Well(i=w.i, j=w.j, k=w.k, value=abs(w.q_target), producer=not np.signbit(w.q_target),
                                        minimum_bhp=w.pwf_min, wellbore_radius=w.rw, completion_ratio=w.hc,
                                        shutin_days=(w.shut_start, w.shut_stop)) for w in spec.wells])
    N = spec.n_cells
    # one sample per core per round; rounds until the time budget is spent
    nb = cores if max_samples is None else max(1, min(cores, max_samples))
    batch = srm.synth.make_batch(spec.W, spec.H, spec.D, 1, nb, [(w.i, w.j) for w in spec.wells[:8]], seed=2002)
    sat = [t.numpy() for t in srm.synth.make_saturations(batch, seed=2002)] if gc else None

    def one(b):
        sl = slice(b, b + 1)
        if gc:
            O.gc_forward_backward(ocfg, otab, batch.kx[sl].numpy(), batch.p0[sl].numpy(), batch.p1[sl].numpy(),
                                  sat[0][sl], sat[1][sl], sat[2][sl], sat[3][sl], batch.dt1[sl].numpy(),
                                  batch.dt2[sl].numpy(), batch.t1[sl].numpy(), np.zeros(1, np.int32),
                                  [1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0])
            return N
        O.dg_forward_backward(ocfg, otab, batch.kx[sl].numpy(), batch.p0[sl].numpy(), batch.p1[sl].numpy(),
                              batch.dt1[sl].numpy(), batch.dt2[sl].numpy(), batch.t1[sl].numpy(),
                              np.zeros(1, np.int32), WEIGHTS)
        return N

    done = 0
    t0 = time.perf_counter()
    with ThreadPoolExecutor(max_workers=cores) as ex:
        while True:
            done += sum(ex.map(one, range(nb)))
            el = time.perf_counter() - t0
            if el >= target_seconds or (max_samples is not None and done >= max_samples * N):
                break
    el = time.perf_counter() - t0
    return dict(value=done / el, unit=UNIT, cores=cores, kind="port",
                sample=(f"{done // N} samples of the {name} grid ({spec.W}x{spec.H}x{spec.D}) fwd+autograd backward, "
                        if slab is None else
                        f"{done // N} z-slabs of {slab} layers of the {name} grid ({spec.W}x{spec.H}x{slab} cells each, wells completed in "
                        f"the slab) fwd+autograd backward, ") + f"{el:.1f} s, {cores} threads over independent samples"), done, el


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(1, args.steps)
    per_step = max(2.0, min(20.0, 60.0 / (steps + args.warmup)))
    vals = []
    info = None
    for i in range(args.warmup + steps):
        info, done, el = cpu_oracle_throughput(args.workload, target_seconds=per_step)
        if i >= args.warmup:
            vals.append((done, el))
    tot = sum(d for d, _ in vals)
    tt = sum(e for _, e in vals)
    v = tot / tt
    c, spec = workload(args.workload)
    info["value"] = v
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tt / steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {'gas-condensate (two-phase)' if spec.fluid_type == 'GC' else 'dry-gas'} {spec.W}x{spec.H}x{spec.D}, T={c['T']}, K={c['K']} (same grid, fewer samples: a bounded sample per step)",
                       "note": "reference CPU path = oracle port of the TF op graph (TensorFlow not installable; reference not import-clean)"},
            "cpu_baseline": info,
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import srm_b200 as srm
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the physics-loss path has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    distributed = world > 1
    if distributed:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    c, spec = workload(args.workload)
    T, K = c["T"], c["K"]
    if args.K:
        K = args.K
    if args.strong:             # the config's K realisations are sharded over the ranks (dist.shard_realisations)
        lo_r, hi_r = srm.dist.shard_realisations(K, rank, world)
        if hi_r - lo_r != K // world or K % world:
            raise SystemExit(f"bench.py --strong: K = {K} realisations do not split evenly over {world} ranks")
        K_job = K
        K = hi_r - lo_r
    else:
        K_job = K * world
    # batches beyond ~8 GB per field (cfg5 at large K: 137 GB per field at K=256) run as `reps` chunks of K_res
    # realisations; the resident synthetic chunk is re-evaluated (same work per chunk, SURVEY 8(d): "large-K cases
    # stream device-generated chunks")
    K_total = K
    K_cap = max(1, int(8e9 // (4 * T * spec.n_cells)))
    reps = 1
    if K > K_cap:
        reps = -(-K // K_cap)
        while K % reps:
            reps += 1
        K = K // reps
    gc = spec.fluid_type == "GC"
    tabs = srm.build_spline_tables(srm.load_default_pvt_table(), srm.pvt.GC_PROPERTIES if gc else srm.pvt.DG_PROPERTIES, order=1)
    lut = args.numerics == "reference" and not args.no_pvt_lut
    torch.cuda.synchronize(dev)
    t_create = time.perf_counter()
    eng = srm.SrmPhysics(spec, tabs, device=local, numerics=args.numerics, pvt_lut=lut)
    torch.cuda.synchronize(dev)
    t_create = time.perf_counter() - t_create
    def make_inputs(seed):
        b_ = srm.synth.make_batch(spec.W, spec.H, spec.D, T, K, [(w.i, w.j) for w in spec.wells[:8]], seed=seed, device=dev)
        if args.p_window:       # stretch the synthetic pressures over [lo, hi] psi: how far the exact table's window reaches
            lo_p, hi_p = (float(v) for v in args.p_window.split(":"))
            for f in ("p0", "p1"):
                t = getattr(b_, f)
                t.sub_(4101.0).mul_((hi_p - lo_p) / (5000.0 - 4101.0)).add_(lo_p)
        d_ = dict(kx=b_.kx, sample_real=b_.sample_real, p0=b_.p0, p1=b_.p1, dt1=b_.dt1, dt2=b_.dt2, t1=b_.t1)
        if gc:
            d_["sg0"], d_["sg1"], d_["so0"], d_["so1"] = srm.synth.make_saturations(b_, seed=seed)
        return b_, d_
    b, d = make_inputs(2002 + rank)
    # rotating inputs: a second, differently seeded batch alternates with the first when both fit comfortably
    field_bytes = 4 * K * T * spec.n_cells
    sets = [d]
    if reps == 1 and field_bytes * (6 if gc else 2) <= 12e9:
        sets.append(make_inputs(3002 + rank)[1])
    fwd = eng.forward_gc if gc else eng.forward
    bwd = eng.backward_gc if gc else eng.backward
    B = b.p0.shape[0]
    N = B * spec.n_cells
    dterms = torch.tensor([1.0, 1.0, 1.0, 0.0, 0.0, 0.0, 0.0, 1.0] if gc else WEIGHTS, dtype=torch.float32, device=dev)

    gout = None

    def step(i=0):
        nonlocal gout
        di = sets[i % len(sets)]
        for _ in range(reps):
            fw = fwd(**di)
            gout = bwd(dterms=dterms, out=gout, **di)
        if distributed:
            srm.dist.allreduce_terms(fw["terms"])
        return fw["terms"], gout

    def barrier():
        if distributed:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize(dev)

    for i in range(max(args.warmup, 3)):
        step(i)
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ef = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps)]
    pending = None              # the previous step's loss-term all-reduce (NCCL's own stream)
    e0.record()
    for i in range(args.steps):
        di = sets[i % len(sets)]
        for rep in range(reps):
            fw = fwd(**di)
            if rep == reps - 1:
                ef[2 * i].record()
            gout = bwd(dterms=dterms, out=gout, **di)
        ef[2 * i + 1].record()
        if distributed:
            # The adjoint's upstream weights are constants, so nothing in a step reads the REDUCED terms: the 128-byte
            # all-reduce is queued behind the adjoint and waited on one step later (reporting only).  A rank never
            # waits for another rank inside a step, and NCCL's kernel never shares the SMs with the adjoint.
            if pending is not None:
                pending.wait()
            pending = srm.dist.allreduce_terms_async(fw["terms"])
    if pending is not None:
        pending.wait()
    e1.record()
    barrier()
    clocks = sampler.stop()
    launches = eng.launches - l0
    ms = e0.elapsed_time(e1)
    if reps == 1:
        fwd_ms = np.mean([(e0 if i == 0 else ef[2 * i - 1]).elapsed_time(ef[2 * i]) for i in range(args.steps)])
    else:
        fwd_ms = float("nan")       # per-pass split is only recorded for single-chunk steps
    bwd_ms = np.mean([ef[2 * i].elapsed_time(ef[2 * i + 1]) for i in range(args.steps)])
    print(f"[bench rank {rank}] timed region {ms:.3f} ms for {args.steps} steps (fwd {fwd_ms:.3f} + bwd {bwd_ms:.3f} ms per step)", file=sys.stderr)
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if distributed:
        import torch.distributed as dist
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    ms_step = ms / args.steps
    value = world * N * reps * args.steps / (ms * 1e-3)
    scaling = "strong" if args.strong else "weak"

    if args.kernels_only:       # tuning runs: the device-timed step alone
        if rank == 0:
            peak, _ = peak_hbm()
            print(json.dumps({"workload": args.workload, "numerics": args.numerics, "ms_per_step": ms_step, "fwd_ms": float(fwd_ms),
                              "bwd_ms": float(bwd_ms), "value": value,
                              "frac_step": N * reps * alg_bytes_per_cell(T, gc) / (ms_step * 1e-3) / 1e9 / peak}))
        if distributed:
            import torch.distributed as dist
            dist.destroy_process_group()
        return
    # ---- end to end through the public API with host buffers: srm.engine.HostPipeline (chunks of whole
    # realisations; H2D, kernels and D2H overlap on three streams)
    host = {k: v.cpu().pin_memory() for k, v in d.items()}
    pipe = srm.engine.HostPipeline(eng, host, dterms, n_chunks=args.e2e_chunks)
    h2d, d2h = pipe.h2d_bytes, pipe.d2h_bytes
    red = srm.dist.allreduce_terms if distributed else None
    e2e_steps = max(1, min(args.steps, 5))
    for _ in range(2):          # untimed: first touches of the pinned staging buffers
        pipe.step(red)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        for _rep in range(reps):
            hterms, hgrads = pipe.step(red)
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if distributed:
        import torch.distributed as dist
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * N * reps * e2e_steps / float(te.item())
    loss = float((hterms[0] * dterms.cpu()).sum())
    # the same call with the cotangents left on the device (where the reference's networks consume them): only the
    # loss terms cross back.  Reported beside the full round trip, not instead of it.
    del pipe, hgrads
    pipe2 = srm.engine.HostPipeline(eng, host, dterms, n_chunks=args.e2e_chunks, grads_to_host=False)
    for _ in range(2):
        pipe2.step(red)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        for _rep in range(reps):
            hterms2, _dg = pipe2.step(red)
    barrier()
    te2 = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if distributed:
        dist.all_reduce(te2, op=dist.ReduceOp.MAX)
    e2e_loss_only = world * N * reps * e2e_steps / float(te2.item())
    d2h_loss_only = pipe2.d2h_bytes
    if not torch.equal(hterms2, hterms):
        raise SystemExit("bench.py: loss-only pipeline terms differ from the full round trip")
    n_chunks_used = len(pipe2.chunks)
    del pipe2, _dg
    # the chunked pipeline must reproduce the resident run: terms are additive over samples
    ref_terms = step()[0]
    if distributed:
        pass        # step() already all-reduced
    torch.cuda.synchronize(dev)
    if not torch.allclose(hterms[0], ref_terms[0].cpu(), rtol=1e-5):
        raise SystemExit(f"bench.py: e2e pipeline terms {hterms[0].tolist()} != resident terms {ref_terms[0].cpu().tolist()}")

    # ---- the same step replayed from a CUDA graph (engine.GraphedStep): what a launch-bound case gains (cfg1, the
    # reference's own grid, is a dozen launches of a few microseconds); reported beside `value`, not instead of it
    graph = None
    if reps == 1 and B * spec.n_cells * 4 <= 2e9:
        gs = srm.engine.GraphedStep(eng, d, dterms)
        for _ in range(3):
            gs.replay()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        g0.record()
        for _ in range(args.steps):
            gs.replay()
        g1.record()
        torch.cuda.synchronize(dev)
        gms = g0.elapsed_time(g1) / args.steps
        if not distributed and not torch.allclose(gs.terms, ref_terms, rtol=1e-5):
            raise SystemExit("bench.py: graph replay terms differ from the eager step")
        graph = {"ms_per_step": gms, "value": world * N / (gms * 1e-3), "unit": UNIT, "kernels_per_replay": gs.kernels,
                 "api": "srm.engine.GraphedStep.replay: forward + adjoint captured once in a CUDA graph"}
        del gs

    # ---- the closed-form numerics mode on the same batch (the route past the exact-table gather floor):
    # same formulas evaluated as exact arithmetic would, <= 2.5e-6 from the fp64 oracle (tests/test_gpu_closed_form.py);
    # its distance to the fp32 reference-order results of THIS batch is measured here
    closed = None
    if not gc and args.numerics == "reference" and not args.no_closed_form:
        g_ref = [t.clone() for t in gout[:2]]
        eng_cf = srm.SrmPhysics(spec, tabs, device=local, numerics="closed_form")
        gcf = None
        for i in range(3):
            fwc = eng_cf.forward(**sets[0])
            gcf = eng_cf.backward(dterms=dterms, out=gcf, **sets[0])
        torch.cuda.synchronize(dev)
        dist_cf = {"terms_rel": [float(abs(a - b) / max(abs(b), 1e-30)) for a, b in zip(fwc["terms"][0].tolist()[:4], ref_terms[0].tolist()[:4])],
                   "gp0_rel_to_max": float((gcf[0] - g_ref[0]).abs().max() / g_ref[0].abs().max()),
                   "gp1_rel_to_max": float((gcf[1] - g_ref[1]).abs().max() / g_ref[1].abs().max())}
        del g_ref
        ec = [torch.cuda.Event(enable_timing=True) for _ in range(2 * args.steps + 1)]
        barrier()
        ec[0].record()
        for i in range(args.steps):
            di = sets[i % len(sets)]
            eng_cf.forward(**di)
            ec[2 * i + 1].record()
            gcf = eng_cf.backward(dterms=dterms, out=gcf, **di)
            ec[2 * i + 2].record()
        torch.cuda.synchronize(dev)
        cf_ms = ec[0].elapsed_time(ec[-1]) / args.steps
        cf_f = float(np.mean([ec[2 * i].elapsed_time(ec[2 * i + 1]) for i in range(args.steps)]))
        cf_a = float(np.mean([ec[2 * i + 1].elapsed_time(ec[2 * i + 2]) for i in range(args.steps)]))
        closed = {"ms_per_step": cf_ms, "fwd_ms": cf_f, "bwd_ms": cf_a, "value_per_gpu": N / (cf_ms * 1e-3),
                  "kernels": "k_fwd_cf2 / k_adj_cf2 (kernels_cf2.cu: 64 x 16 column tiles marching over z, every global read a TMA box copy "
                             "-- cp.async.bulk.tensor.4d -- into mbarrier rings four planes deep)",
                  "error_vs_fp64_oracle": "dom <= 1.9e-7, gp0 <= 2.5e-6, gp1 <= 8e-7 of max (gate of tests/test_gpu_closed_form.py); on the full "
                                          "cfg5 grid dom 1.8e-7, gp1 2.9e-7, gdt1 4e-8, gp0 1.6e-5 (tests/test_gpu_full_grid.py)",
                  "distance_to_fp32_reference_order": dist_cf}
        eng_cf.close()
        del eng_cf, gcf

    # ---- the fused glue either side of the kernels: HardLayer at both levels + dt means and
    # their cotangents, timed on the same resident batch (not part of `value`: the reference's metric is the residual)
    glue = None
    if not gc:
        ydev = torch.rand_like(d["p0"]) * 600.0
        expo = torch.full(d["p0"].shape[1:], 0.5, device=dev)
        tn = torch.linspace(-0.9, 0.9, B, device=dev)
        gsteps = max(3, min(args.steps, 10))

        def gstep():
            p0g, p1g, dt1g, dt2g = eng.glue_forward(ydev, d["p1"], tn, tn, expo, d["p0"], d["p1"], 5000.0)
            return eng.glue_backward(ydev, d["p1"], tn, tn, p0g, p1g, expo, dt1g, dt2g, 5000.0)
        for _ in range(3):
            gstep()
        g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        g0.record()
        for _ in range(gsteps):
            gstep()
        g1.record()
        torch.cuda.synchronize(dev)
        gms = g0.elapsed_time(g1) / gsteps
        gbytes = 56.0            # forward: read y0,y1,dtf1,dtf2, write p0,p1 (24 B); backward: read y0,y1,gp0,gp1, write gy0,gy1,gdtf1,gdtf2 (32 B)
        glue = {"ms_per_step": gms, "cell_timesteps_per_s": N / (gms * 1e-3), "alg_bytes_per_cell": gbytes,
                "achieved_GBps": N * gbytes / (gms * 1e-3) / 1e9,
                "kernels": "k_glue_fwd + k_glue_mean + k_glue_bwd (srm_glue_forward / srm_glue_backward)"}
        del ydev
        # batch gather of the device-resident feature tensor (BatchGenerator.__getitem__)
        try:
            lib = srm._lib.load_library()
            rowb = spec.n_cells * 5 * 4
            nb = min(B, max(1, int(2e9 // rowb)))
            xs = torch.rand((nb, spec.n_cells * 5), device=dev)
            xo = torch.empty_like(xs)
            perm = torch.randperm(nb, device=dev, dtype=torch.int32)
            st = torch.cuda.current_stream(dev).cuda_stream

            def gat():
                for a in range(0, nb, 65535):
                    e = min(nb, a + 65535)
                    srm._lib.check(lib, lib.srm_gather_rows(local, xs.data_ptr(), perm[a:e].data_ptr(), e - a, nb, rowb, xo[a:e].data_ptr(), st), "srm_gather_rows")
            for _ in range(3):
                gat()
            q0, q1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(dev)
            q0.record()
            for _ in range(gsteps):
                gat()
            q1.record()
            torch.cuda.synchronize(dev)
            qms = q0.elapsed_time(q1) / gsteps
            glue["batch_gather"] = {"ms": qms, "rows": nb, "row_bytes": rowb, "achieved_GBps": 2.0 * nb * rowb / (qms * 1e-3) / 1e9,
                                    "kernel": "k_gather_rows (srm_gather_rows): x_batch = x_all[batch_inds] on the device-resident data set"}
            # feature glue on the same rows: x_n1 = x with t_norm += dn[b], de-normalised permeability channel, and
            # the cotangent of dn (srm_features_forward / _backward): 20 + 20 + 4 B forward, 20 B backward per cell
            xf = xs.view(nb, spec.n_cells, 5)
            dnb = torch.rand(nb, device=dev) * 0.01

            def feat():
                x1f, kxf = eng.features_forward(xf, dnb, (0.26, 24.0))
                return eng.features_backward(x1f)
            for _ in range(3):
                feat()
            torch.cuda.synchronize(dev)
            q0.record()
            for _ in range(gsteps):
                feat()
            q1.record()
            torch.cuda.synchronize(dev)
            fms = q0.elapsed_time(q1) / gsteps
            glue["features"] = {"ms": fms, "rows": nb, "alg_bytes_per_cell": 64.0,
                                "achieved_GBps": nb * spec.n_cells * 64.0 / (fms * 1e-3) / 1e9,
                                "kernel": "k_features + k_features_bwd (srm_features_forward / srm_features_backward): time-shifted features, permeability channel, d/d dt"}
            del xs, xo, xf
        except Exception as e:      # never let an auxiliary measurement break the bench line
            glue["batch_gather"] = {"error": repr(e)}

    if rank == 0:
        peak, peak_src = peak_hbm()
        ab = alg_bytes_per_cell(T, gc)
        # the window of the exact table this batch touches: entries = representable fp32 values between min and max pressure
        p_lo = float(min(d["p0"].min(), d["p1"].min()))
        p_hi = float(max(d["p0"].max(), d["p1"].max()))
        n_entries = int(np.float32(p_hi).view(np.int32)) - int(np.float32(p_lo).view(np.int32)) + 1
        if glue:
            glue["frac"] = glue["achieved_GBps"] / peak
            if "achieved_GBps" in glue.get("batch_gather", {}):
                glue["batch_gather"]["frac"] = glue["batch_gather"]["achieved_GBps"] / peak
            if "achieved_GBps" in glue.get("features", {}):
                glue["features"]["frac"] = glue["features"]["achieved_GBps"] / peak
        achieved = N * reps * ab / (ms_step * 1e-3) / 1e9   # per GPU (each rank runs N * reps cells per step)
        # roofline of the dominant kernel (the adjoint pass) per the bench contract, with the forward pass and the whole
        # step beside it.  Pass durations are CUDA-event windows on the launching stream over the timed region: the
        # adjoint window holds k_adj4 (or the GC adjoint) plus its sparse inner-boundary and finalize kernels (~2 %).
        tr = measured_traffic(args.workload, args.numerics) or {}
        pk = tr.get("per_kernel", {})

        def kernel_traffic(*names):
            for n in names:
                for k, v in pk.items():          # template instances carry their arguments: k_fwd4<0>
                    if k == n or k.startswith(n + "<"):
                        return v["read"] + v["write"]
            return None
        ab_f = (28.0 if gc else 12.0) + 4.0 / T
        ab_a = (52.0 if gc else 16.0) + 4.0 / T
        one = N / 1e9                       # cells of one chunk (the pass windows cover one chunk)

        def pass_obj(name, ms_, abytes, traffic):
            ach = one * abytes / (ms_ * 1e-3) if ms_ == ms_ and ms_ > 0 else None
            return {"kernel": name, "ms": float(ms_), "alg_bytes_per_cell": abytes, "achieved": ach,
                    "frac": (ach / peak) if ach else None, "traffic": traffic}
        lean = lut and spec.W % 2 == 0          # kernels_dg4.cu takes even widths with the full-range table; else kernels_ref2.cu
        k_a = "k_adj_gc2" if gc else ("k_adj4" if lean else "k_adj_ref2" if lut else "adjoint kernels")
        k_f = "k_fwd_gc2" if gc else ("k_fwd4" if lean else "k_fwd_ref2" if lut else "forward kernels")
        adj = pass_obj("adjoint pass: " + k_a + " (+ inner-boundary scatter, finalize)", bwd_ms, ab_a,
                       kernel_traffic("k_adj4", "k_adj_gc2"))
        fwdp = pass_obj("forward pass: " + k_f + " (+ faces, wells, finalize)", fwd_ms, ab_f,
                        kernel_traffic("k_fwd4", "k_fwd_gc2"))
        roof = {"bound": "hbm", "unit": "GB/s", "peak": peak, "peak_source": peak_src,
                "kernel": adj["kernel"], "achieved": adj["achieved"], "frac": adj["frac"], "traffic": adj["traffic"],
                "alg_bytes_per_cell": ab_a, "ms": adj["ms"],
                "traffic_source": tr.get("source"),
                "forward": fwdp,
                "step": {"kernel": "whole step (forward + adjoint launches)", "ms": ms_step, "alg_bytes_per_cell": ab,
                         "achieved": achieved, "frac": achieved / peak, "traffic": tr.get("bytes_per_step")},
                "fwd_ms": float(fwd_ms), "bwd_ms": float(bwd_ms),
                "note": ("reference-order numerics: the exact-table gathers bound both passes at the L2 sector rate "
                         "(tools/gather_probe2.cu; floor = 25 % of the HBM roofline, DESIGN.md 5.1)") if (lut and not gc) else None}
        if closed is not None:
            closed["alg_bytes_per_cell"] = ab
            closed["achieved"] = closed["value_per_gpu"] * ab / 1e9
            closed["frac"] = closed["achieved"] / peak
            closed["forward_frac"] = one * ab_f / (closed["fwd_ms"] * 1e-3) / peak
            closed["adjoint_frac"] = one * ab_a / (closed["bwd_ms"] * 1e-3) / peak
            tr_cf = measured_traffic(args.workload, "closed_form")      # DRAM bytes of one forward + one adjoint launch at this batch
            closed["traffic"] = tr_cf.get("bytes_per_step") if tr_cf else None
            closed["traffic_source"] = tr_cf.get("source") if tr_cf else None
            roof["closed_form"] = closed
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cpu, _, _ = cpu_oracle_throughput(args.workload, target_seconds=12.0)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: {'gas-condensate (two-phase)' if gc else 'dry-gas'} {spec.W}x{spec.H}x{spec.D}, T={T}, K={K_total} per GPU" + (f" as {reps} chunks of K={K}" if reps > 1 else "") + f", B={B * reps}, "
                                   f"{len(spec.wells)} well connections" + (", blocking-factor integral" if spec.use_blocking_factor else ""),
                       "numerics": args.numerics, "cells_per_gpu_per_step": N * reps,
                       "pvt": ("reference-order spline tabulated per fp32 pressure over the clamp range at handle creation "
                               "(%.2f s, outside the timed region, bit-identical to direct evaluation); pressures of this batch span "
                               "[%.0f, %.0f] psi = %.2f M table entries" % (t_create, p_lo, p_hi, n_entries / 1e6)) if lut
                              else "evaluated per cell",
                       "l2": "inputs+workspace per step (%.0f MB) exceed the 126 MB L2; no explicit flush" % ((2 * N * 4 + eng.workspace(B, b.kx.shape[0]).numel()) / 1e6),
                       "parallelism": f"sample-sharded x{world} ({scaling}: K = {K_job} realisations in the job); the only collective is the all-reduce of the "
                                      "16-float loss-term vector, queued behind the adjoint and waited on one step later"},
            "roofline": roof,
            "cpu_baseline": cpu,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d) * reps, "d2h_bytes_per_step": int(d2h) * reps,
                    "steps": e2e_steps, "loss": loss,
                    "h2d_GBps_per_rank": int(h2d) * reps * e2e_steps / e2e_s / 1e9, "d2h_GBps_per_rank": int(d2h) * reps * e2e_steps / e2e_s / 1e9,
                    "api": f"srm.engine.HostPipeline.step: pinned host batch, {n_chunks_used} chunks of whole realisations, H2D / kernels / D2H on three streams; "
                           "value = full round trip (all cotangent fields back to pinned host memory)",
                    "grads_on_device": {"value": e2e_loss_only, "unit": UNIT, "d2h_bytes_per_step": int(d2h_loss_only) * reps,
                                        "note": "same call with grads_to_host=False: inputs from host, loss terms to host, cotangents stay in HBM for the networks' backward"}},
            "cuda_graph": graph,
            "gpu_launches": int(launches),
            "clocks": clocks,
            "glue": glue,
        }
        print(json.dumps(line))
    if distributed:
        import torch.distributed as dist
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg5", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--strong", action="store_true", help="shard the config's K realisations over the ranks (strong scaling) instead of K per rank")
    ap.add_argument("--p-window", default="", help="lo:hi psi -- stretch the synthetic pressures over this window (exact-table locality check)")
    ap.add_argument("--numerics", default="reference", choices=["reference", "closed_form"])
    ap.add_argument("--K", type=int, default=0, help="override realisations per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-closed-form", action="store_true", help="skip the closed-form numerics leg reported as roofline.closed_form")
    ap.add_argument("--kernels-only", action="store_true", help="tuning: print the device-timed step only (no e2e, graph, glue, cpu legs)")
    ap.add_argument("--e2e-chunks", type=int, default=8, help="realisation chunks of the host-buffer pipeline")
    ap.add_argument("--no-pvt-lut", action="store_true", help="reference numerics: evaluate the 37-term spline per cell instead of the exact table")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
