"""CPU tests of the host-side logic of the product package (no GPU, no compute through the .so)."""
import ctypes
import os
import re

import numpy as np
import pytest

import util as U

srm, O = U.srm, U.O


def test_product_spline_solve_is_bitwise_the_oracles():
    ptab = srm.build_spline_tables(srm.load_default_pvt_table(), srm.pvt.DG_PROPERTIES, order=1)
    cols = O.load_pvt_table(os.path.join(U.GOLDEN, "pvt_table.npz"))
    otab = O.build_spline_table(cols, O.DG_PROPS, order=1, lam=0.001)
    assert np.array_equal(ptab.knots, otab.c)
    assert np.array_equal(ptab.w, otab.w) and np.array_equal(ptab.v, otab.v)


def test_pvt_table_lookup_is_case_insensitive():
    t = srm.load_default_pvt_table()
    assert np.array_equal(t.lookup("invBg"), t.lookup("InvBg"))
    assert np.array_equal(t.lookup("pre"), t.lookup("Pre"))
    with pytest.raises(KeyError):
        t.lookup("nope")


def test_spec_scalars_match_oracle():
    spec = srm.PhysicsSpec()
    assert np.float32(spec.cf) == O.rock_compressibility(0.2)
    _, krg = O.corey_krog_krgo_np(1.0 - 0.22, O.OracleConfig(), np.float32)
    assert np.float32(spec.krg) == np.float32(krg) == np.float32(0.9)
    assert np.float32(spec.Sgi) == np.float32(0.78)
    assert abs(spec.dx - 2900.0 / 39) < 1e-12 and spec.dz == 80.0


def test_wells_from_connections_sign_rule_and_shutins():
    conns = [
        {"i": 1, "j": 2, "k": 0, "type": "producer", "control": "ORAT", "value": 500.0, "minimum_bhp": 4100.0,
         "wellbore_radius": 0.1, "completion_ratio": 0.5, "shutin_days": [[10.0, 20.0]]},
        {"i": 3, "j": 4, "k": 1, "type": "Injector", "control": "orat", "value": 250.0},
        {"i": 0, "j": 0, "k": 0, "type": "injector", "control": "BHP", "value": 3000.0, "shutin_days": [[1, 2], [3, 4]]},
    ]
    w = srm.config.wells_from_connections(conns)
    assert w[0].q_target == 500.0 and (w[0].shut_start, w[0].shut_stop) == (10.0, 20.0)
    assert w[1].q_target == -250.0 and (w[1].shut_start, w[1].shut_stop) == (0.0, 0.0)   # default [[0,0]]
    assert w[2].q_target == 3000.0                                                      # BHP stays positive
    assert (w[2].shut_start, w[2].shut_stop) == (0.0, 0.0)                              # malformed list -> default


def test_reference_shaped_config_dicts_are_accepted():
    spec = srm.spec_from_reference_configs()
    assert (spec.W, spec.H, spec.D) == (39, 39, 1) and len(spec.wells) == 5
    assert [(w.i, w.j, w.k) for w in spec.wells] == [(29, 29, 0), (29, 9, 0), (9, 9, 0), (9, 29, 0), (19, 19, 0)]
    assert spec.wells[4].q_target == 0.0 and np.signbit(spec.wells[4].q_target)         # injector: -0.0
    spec2 = srm.spec_from_reference_configs(reservoir={"Nx": 64, "Ny": 32, "Nz": 4, "porosity": 0.25})
    assert (spec2.W, spec2.H, spec2.D, spec2.phi) == (64, 32, 4, 0.25)


def test_header_symbols_are_all_exported():
    """the C-ABI library loads and exports every function include/srm_physics.h declares."""
    hdr = open(os.path.join(U.ROOT, "include", "srm_physics.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    names = set(re.findall(r"\b(srm_[a-z_]+)\s*\(", hdr))
    assert names == set(srm._lib.EXPORTS), names ^ set(srm._lib.EXPORTS)
    lib = srm._lib.load_library()
    for n in names:
        assert hasattr(lib, n), n
    assert lib.srm_version() == srm._lib.SRM_ABI_VERSION


def test_struct_layout_matches_header_field_order():
    hdr = open(os.path.join(U.ROOT, "include", "srm_physics.h")).read()
    body = hdr[hdr.index("typedef struct SrmConfig {"):hdr.index("} SrmConfig;")]
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        m = re.match(r"(?:const\s+)?(?:int32_t|float|SrmWell\*|float\*)\s*\*?\s*(.*)", decl.replace("typedef struct SrmConfig {", "").strip())
        if not m or not decl:
            continue
        for nm in m.group(1).split(","):
            nm = nm.strip().lstrip("*").strip()
            if nm:
                fields.append(nm)
    assert fields == [f[0] for f in srm._lib.SrmConfig._fields_], fields


def test_create_rejects_bad_configs_without_touching_a_gpu():
    """argument validation happens before any CUDA call, so it is testable on the CPU box"""
    lib = srm._lib.load_library()
    L = srm._lib
    tabs = srm.build_spline_tables(srm.load_default_pvt_table(), srm.pvt.DG_PROPERTIES)
    base = dict(device=0, D=1, H=4, W=4, dx=1.0, dy=1.0, dz=1.0, C_=0.001127, Dc=5.6145833334, phi=0.2, cf=1e-5,
                Sgi=0.78, krg=0.9, kx_ky=1.0, kv_kh=1.0, knots=tabs.knots, spline_w=tabs.w, spline_v=tabs.v,
                spline_order=1, p_min=14.7, p_max=1e4, wells=[], use_blocking_factor=False, n_intervals=8,
                numerics=0, tde_in_dom=True)

    def rc(**kw):
        cfg, keep = L.make_config(**{**base, **kw})
        h = ctypes.c_void_p()
        r = lib.srm_create(ctypes.byref(cfg), ctypes.byref(h))
        return r, lib.srm_last_error().decode()

    r, msg = rc(W=0)
    assert r == -1 and "grid" in msg
    r, msg = rc(spline_order=3)
    assert r == -1 and "spline_order" in msg
    r, msg = rc(numerics=7)
    assert r == -1 and "numerics" in msg
    r, msg = rc(wells=[dict(i=9, j=0, k=0, q_target=1.0, pwf_min=1.0, rw=0.1, hc=0.5, shut_start=0, shut_stop=0)])
    assert r == -1 and "outside the grid" in msg
    r, msg = rc(knots=tabs.knots[::-1].copy())
    assert r == -1 and "ascending" in msg
    r, msg = rc(fluid_type=1)
    assert r == -1


def test_engine_refuses_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    tabs = srm.build_spline_tables(srm.load_default_pvt_table(), srm.pvt.DG_PROPERTIES)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        srm.SrmPhysics(srm.PhysicsSpec(), tabs)


def test_synthetic_batch_ranges():
    b = srm.synth.make_batch(16, 12, 3, 4, 2, [(4, 4)], seed=1)
    assert b.p0.shape == (8, 3, 12, 16) and b.kx.shape == (2, 3, 12, 16)
    assert float(b.kx.min()) >= 0.2599 and float(b.kx.max()) <= 24.001
    assert float(b.p0.min()) > 4100 and float(b.p0.max()) <= 5000 and float((b.p0 - b.p1).max()) <= 31
    assert float(b.dt1.min()) >= 0.1 and float(b.dt1.max()) <= 10.0
    assert b.sample_real.tolist() == [0, 0, 0, 0, 1, 1, 1, 1]


def test_well_data_processor_integer_bookkeeping():
    wdp = srm.WellDataProcessor(srm.config.DEFAULT_WELLS["connections"])
    wd = wdp.get_well_data()
    assert wd["connection_index"].tolist() == [[0, 29, 29], [0, 9, 29], [0, 9, 9], [0, 29, 9], [0, 19, 19]]   # rows [k, j, i]
    assert wd["control_mode_value"].tolist()[:4] == [500.0, 1000.0, 500.0, 1000.0]
    m = wdp.scatter_y((1, 1, 39, 39, 1), wd["connection_index"], 1.0)
    assert float(m.sum()) == 5.0 and m[0, 0, 29, 29, 0] == 1.0 and m[0, 0, 9, 29, 0] == 1.0
    dup = wdp.scatter_y((1, 1, 4, 4, 1), [[0, 1, 1], [0, 1, 1]], [2.0, 3.0])
    assert dup[0, 0, 1, 1, 0] == 5.0                          # duplicates sum (tf.scatter_nd)
    import torch
    t = torch.tensor([5.0, 15.0, 25.0]).view(3, 1, 1, 1).expand(3, 1, 4, 4)
    s = wdp.conn_shutins_idx(t, [[0, 1, 1], [0, 2, 3]], [[[10.0, 20.0]], [[1000.0, 0.0]]])
    assert s[:, 0, 1, 1].tolist() == [1, 0, 1] and s[:, 0, 2, 3].tolist() == [1, 1, 1] and int(s.sum()) == 5


def test_committed_bench_line_meets_the_contract():
    """the bench line of the final build (profiles/r5_bench_cfg5.json, written by `python bench.py` on a B200) carries every key
    the contract names, and its roofline numbers follow from its own timings"""
    import json
    import math
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles", "r5_bench_cfg5.json")
    d = json.load(open(path))
    for k in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
              "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert k in d, k
    assert d["unit"] == "cell-timesteps/s" and d["higher_is_better"] is True and d["n_gpus"] == 1 and d["dtype"] == "f32"
    assert d["warmup"] >= 3 and "cfg5" in d["config"]["workload"] and "model" not in d["config"]
    cells = d["config"]["cells_per_gpu_per_step"]
    assert cells == 256 * 256 * 64 * 32 * 8
    assert math.isclose(d["value"], cells / (d["ms_per_step"] * 1e-3), rel_tol=1e-6)
    r = d["roofline"]
    assert r["bound"] == "hbm" and r["unit"] == "GB/s" and math.isclose(r["frac"], r["achieved"] / r["peak"], rel_tol=1e-9)
    assert math.isclose(r["achieved"], cells * r["alg_bytes_per_cell"] / (r["ms"] * 1e-3) / 1e9, rel_tol=1e-6)
    assert r["traffic"] and r["traffic"] >= cells * r["alg_bytes_per_cell"]          # measured DRAM bytes are not below the compulsory ones
    assert math.isclose(r["step"]["ms"], d["ms_per_step"], rel_tol=1e-9) and r["step"]["alg_bytes_per_cell"] == 28.25
    cf = r["closed_form"]
    assert math.isclose(cf["frac"], cells * 28.25 / (cf["ms_per_step"] * 1e-3) / 1e9 / r["peak"], rel_tol=1e-6)
    assert "k_fwd_cf2" in cf["kernels"] and cf["distance_to_fp32_reference_order"]["gp1_rel_to_max"] < 1e-3
    e = d["e2e"]
    assert e["unit"] == d["unit"] and e["h2d_bytes_per_step"] >= 2 * 4 * cells and e["d2h_bytes_per_step"] >= 2 * 4 * cells
    assert 0 < e["value"] < d["value"]                     # host buffers in the timed region: never faster than resident inputs
    c = d["cpu_baseline"]
    assert c["kind"] == "port" and c["cores"] >= 1 and c["unit"] == d["unit"] and c["sample"]
    assert d["gpu_launches"] > 0 and d["clocks"]["sm_mhz"] > 0
    assert not set(d["clocks"]["reasons"]) & {"hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown"}


def test_cpu_baseline_leg_of_the_bench_runs_the_oracle():
    """bench.py's cpu_baseline / --impl reference leg: the oracle port on host threads, a bounded sample of the named grid
    (config 1 in full; config 5 as 16-layer z-slabs of the same x-y extent)"""
    import bench
    info, done, el = bench.cpu_oracle_throughput("cfg1", target_seconds=0.2, max_samples=2)
    assert info["kind"] == "port" and info["unit"] == "cell-timesteps/s" and info["cores"] >= 1
    assert done >= 39 * 39 and info["value"] > 0 and "39x39x1" in info["sample"]
    c, spec = bench.workload("cfg5")
    assert (spec.W, spec.H, spec.D, c["T"], c["K"]) == (256, 256, 64, 32, 8) and len(spec.wells) == 2048 and spec.use_blocking_factor
    assert bench.alg_bytes_per_cell(32, False) == 28.25 and bench.alg_bytes_per_cell(24, True) == 80.0 + 8.0 / 24
