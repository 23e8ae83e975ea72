"""Feature construction and the on-disk formats: host-side readers on CPU, the device weave against
the reference's own weave_tensors + DataSummary.normalize (tests/golden/make_reference_weave_golden.py)."""
import os

import numpy as np
import pytest
import torch

import util as U

srm = U.srm


def test_permx_dat_round_trip_and_repeat_counts(tmp_path):
    from srm_b200 import data
    rng = np.random.default_rng(7)
    f = rng.uniform(0.26, 24.0, (2, 3, 4)).astype(np.float32)
    p = tmp_path / "PERMX_0001.dat"
    data.write_permx_dat(str(p), f, comments=["REALIZATION: 1", "GRID: 4x3x2"])
    txt = p.read_text().splitlines()
    assert txt[0].startswith("-- ") and txt[2] == "PERMX" and txt[-1] == "/" and len(txt) == 2 + 1 + f.size + 1
    g = data.read_permx_dat(str(p), shape=f.shape, keyword="PERMX")
    assert g.dtype == np.float32 and np.array_equal(g, f)                 # repr of a float32 round-trips
    assert np.array_equal(data.read_permx_dat(str(p)), f.reshape(-1))     # keyword not named: first block
    # simulator-style include: several values per line, n*value repeats, the slash on the data line, a trailing comment
    q = tmp_path / "perm.inc"
    q.write_text("-- header\nPORO\n 0.2 0.2 /\nPERMX\n 3*1.5 2.0  -- three cells\n 2*0.25 /\n")
    assert np.array_equal(data.read_permx_dat(str(q), keyword="permx"), np.asarray([1.5, 1.5, 1.5, 2.0, 0.25, 0.25], np.float32))
    with pytest.raises(ValueError):
        data.read_permx_dat(str(q), keyword="PERMZ")
    with pytest.raises(ValueError):
        data.read_permx_dat(str(q), keyword="PERMX", shape=(7,))


def test_kle_npy_reader(tmp_path):
    from srm_b200 import data
    a = np.random.default_rng(3).random((3, 2, 4, 5))
    np.save(tmp_path / "k.npy", a)
    np.savez_compressed(tmp_path / "k.npz", permeability=a[0])
    r = data.read_kle_npy(str(tmp_path / "k.npy"))
    assert r.shape == (3, 2, 4, 5) and r.dtype == np.float32 and np.array_equal(r, a.astype(np.float32))
    assert data.read_kle_npy(str(tmp_path / "k.npz")).shape == (1, 2, 4, 5)
    with pytest.raises(FileNotFoundError):
        data.read_kle_npy(str(tmp_path / "missing.npy"))


def test_positional_grids_match_the_golden_layout():
    from srm_b200 import data
    g = np.load(os.path.join(U.GOLDEN, "reference_weave.npz"))
    z, y, x = data.positional_grids(2, 5, 6, 2900.0, 2900.0, 80.0)
    assert np.array_equal(z, g["a_z"]) and np.array_equal(y, g["a_y"]) and np.array_equal(x, g["a_x"])


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["a", "b"])
def test_device_weave_equals_the_reference_weave_and_normalise(case):
    """srm_weave_features against the reference's OWN weave_tensors + DataSummary.normalize: sample order (k*T + t),
    channel order [z, y, x, t, k], linear channels bit for bit, the logarithmic permeability channel to 2 ulp
    (logf of CUDA vs torch)."""
    from srm_b200 import data
    g = np.load(os.path.join(U.GOLDEN, "reference_weave.npz"))
    dev = torch.device("cuda", 0)
    tt = lambda k: torch.as_tensor(g[f"{case}_{k}"]).to(dev).contiguous()
    out = data.weave_features(tt("permx"), tt("time"), tt("x"), tt("y"), tt("z"), g[f"{case}_stats"], limits=(-1.0, 1.0))
    torch.cuda.synchronize()
    ref = g[f"{case}_features"]
    got = out.cpu().numpy()
    assert got.shape == ref.shape
    assert np.array_equal(got[..., :4].view(np.uint32), ref[..., :4].view(np.uint32))
    assert U.ulp_diff(got[..., 4], ref[..., 4]) <= 2 or np.allclose(got[..., 4], ref[..., 4], rtol=0, atol=3e-7)


@pytest.mark.gpu
def test_device_weave_feeds_the_batch_generator_order():
    """the woven tensor is realisation-major (b = k*T + t): what batching.BatchGenerator flattens"""
    from srm_b200 import data
    dev = torch.device("cuda", 0)
    K, T, D, H, W = 3, 5, 2, 4, 8
    z, y, x = (torch.as_tensor(a).to(dev) for a in data.positional_grids(D, H, W, 2900.0, 2900.0, 80.0))
    permx = torch.rand(K, D, H, W, device=dev) * 20 + 0.3
    time = torch.linspace(0, 365, T, device=dev)
    stats = [[0, 80], [0, 2900], [0, 2900], [0, 365], [0.26, 24.0]]
    f = data.weave_features(permx, time, x, y, z, stats)
    assert f.shape == (K * T, D, H, W, 5)
    f5 = f.view(K, T, D, H, W, 5)
    assert torch.equal(f5[:, 0, ..., 4], f5[:, T - 1, ..., 4])            # permeability is time independent
    assert torch.equal(f5[0, :, ..., 3], f5[K - 1, :, ..., 3])            # time is realisation independent
    tn = f5[0, :, 0, 0, 0, 3].cpu().numpy()
    assert np.allclose(tn, np.linspace(-1, 1, T), atol=1e-6)


def _simfiles():
    return np.load(os.path.join(os.path.dirname(__file__), "golden", "reference_simfiles.npz"))


def test_restart_keywords_equal_the_reference_parser(tmp_path):
    """formatted restart file: blocks per keyword and report step, bit for bit what parse_continuous_file returns
    (simulation_data_process_pipeline.py:247-292), from the text and from a file on disk"""
    from srm_b200 import data
    g = _simfiles()
    text = str(g["restart_text"])
    keys = ["PRESSURE", "SGAS", "SOIL", "SWAT"]
    p = tmp_path / "CASE.FUNRST"
    p.write_text(text)
    for src in (text, str(p)):
        got = data.read_restart_keywords(src, keys)
        for k in keys:
            n = int(g[f"restart_{k}_n"])
            assert len(got[k]) == n, k
            for i in range(n):
                assert got[k][i].dtype == np.float32 and np.array_equal(got[k][i], g[f"restart_{k}_{i}"]), (k, i)
    assert got["SWAT"] == []
    grid = data.restart_to_grid(got["PRESSURE"], 2, 3, 5)
    assert grid.shape == (3, 2, 3, 5) and grid[1, 1, 2, 4] == got["PRESSURE"][1][-1] and grid[2, 0, 1, 0] == got["PRESSURE"][2][5]
    with pytest.raises(ValueError):
        data.restart_to_grid(got["PRESSURE"], 2, 3, 4)
    assert data.restart_to_grid([], 2, 3, 5).shape == (0, 2, 3, 5)


def test_rsm_columns_equal_the_reference_parser():
    """run summary: segmented tab-separated tables, merged multi-line titles, compound (name, qualifier) columns, NaN for a
    cell that is not a number, empty cells skipped, None for what is not there (parse_tabular_file_from_string, :148-245)"""
    from srm_b200 import data
    g = _simfiles()
    spec = [["TIME"], ["WOPR", "15 15 1"], ["WOPR", "20  20 1"], "WGPR", "WWPR", "WBHP", "FPR"]
    got = data.read_rsm_columns(str(g["rsm_text"]), spec)

    def same(a, key):
        if bool(g[key + "_none"]):
            assert a is None, key
        else:
            assert a is not None and a.dtype == np.float32 and np.array_equal(a, g[key], equal_nan=True), key

    for k in ("TIME", "WGPR", "WWPR", "WBHP", "FPR"):
        same(got[k], "rsm_" + k)
    for s in ("15 15 1", "20  20 1"):
        same(got["WOPR"][s], f"rsm_WOPR|{s}")
    assert got["TIME"].size == 12 and np.isnan(got["WOPR"]["15 15 1"][2]) and got["WGPR"].size == 5
    # dictionary form of the spec, and page-break lines of a real file (a numeric line with no header above it) are skipped
    paged = "1\n" + str(g["rsm_text"]).replace("\n\n\n", "\n\n1\n")
    again = data.read_rsm_columns(paged, {"FPR": ["FPR"], "WBHP": ["WBHP"]})
    assert np.array_equal(again["FPR"], got["FPR"]) and np.array_equal(again["WBHP"], got["WBHP"])
    assert data.read_rsm_columns("", ["TIME"]) == {"TIME": None}
