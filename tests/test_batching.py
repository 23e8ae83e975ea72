"""BatchGenerator mirror: the flattening of the collapsed axes against golden vectors made by the
REFERENCE'S OWN `_maybe_flatten` (tests/golden/make_batch_golden.py cuts it out of training.py), and -- on the GPU --
batches from the device-resident data set against numpy indexing, bit for bit (byte/index work)."""
import ctypes as C
import os

import numpy as np
import pytest
import torch

import util as U

srm = U.srm
G = np.load(os.path.join(U.GOLDEN, "batch_golden.npz"))


@pytest.mark.parametrize("name", ["kt_dhw5", "kt_only", "three", "inner"])
@pytest.mark.parametrize("order", ["F", "C"])
def test_maybe_flatten_matches_the_reference_function(name, order):
    from srm_b200.batching import maybe_flatten
    got = maybe_flatten(G[f"{name}_in"], list(G[f"{name}_axes"]), order)
    want = G[f"{name}_{order}"]
    assert got.shape == want.shape and np.array_equal(got, want)


def test_sample_order_is_realisation_fastest():
    """K x T collapsed in Fortran order: b = k + K*t ('flattened Fortran-order by BatchGenerator')"""
    from srm_b200.batching import maybe_flatten
    K, T = 3, 4
    a = np.arange(K * T, dtype=np.float32).reshape(K, T)         # a[k, t] = k*T + t
    f = maybe_flatten(a, [0, 1])
    for k in range(K):
        for t in range(T):
            assert f[k + K * t] == a[k, t]


@pytest.mark.gpu
def test_batches_equal_numpy_indexing_bit_for_bit():
    rng = np.random.default_rng(4200)
    feats = [rng.standard_normal((3, 4, 2, 5, 6, 5)).astype(np.float32), rng.standard_normal((2, 4, 2, 5, 6, 5)).astype(np.float32)]
    labels = [{"p": rng.standard_normal((3, 4, 2, 5, 6, 1)).astype(np.float32), "q": rng.standard_normal((3, 4, 2, 5, 6, 1)).astype(np.float32)},
              {"p": rng.standard_normal((2, 4, 2, 5, 6, 1)).astype(np.float32), "q": rng.standard_normal((2, 4, 2, 5, 6, 1)).astype(np.float32)}]
    pairs = list(zip(feats, labels))
    from srm_b200.batching import maybe_flatten
    x_all = np.concatenate([maybe_flatten(f, [0, 1]) for f in feats], axis=0)
    y_all = {k: np.concatenate([maybe_flatten(lb[k], [0, 1]) for lb in labels], axis=0) for k in ("p", "q")}
    for stack in (False, True):
        np.random.seed(4201)
        gen = srm.BatchGenerator(pairs, batch_size=7, shuffle=True, stack_labels=stack)
        np.random.seed(4201)
        ind = np.arange(x_all.shape[0])
        np.random.shuffle(ind)
        assert len(gen) == int(np.ceil(20 / 7)) and gen.N == 20
        for epoch in range(2):
            for i in range(len(gen)):
                x, y = gen[i]
                sel = ind[i * 7:min((i + 1) * 7, 20)]
                assert x.is_cuda and np.array_equal(x.cpu().numpy().view(np.uint32), x_all[sel].view(np.uint32))
                if stack:
                    want = np.stack([y_all[k][sel] for k in ("p", "q")], axis=0)
                    assert np.array_equal(y.cpu().numpy(), want)
                else:
                    for k in ("p", "q"):
                        assert np.array_equal(y[k].cpu().numpy(), y_all[k][sel])
            before = gen.indices.copy()
            gen.on_epoch_end()                                   # reshuffles with numpy's global generator, as the reference
            ind = gen.indices.copy()
            assert sorted(ind.tolist()) == list(range(20)) and not np.array_equal(ind, before)
    # the reference's batching arithmetic on its own shuffled index vector (golden)
    N, bs = int(G["idx_N"]), int(G["idx_bs"])
    data = np.arange(N * 3, dtype=np.float32).reshape(N, 3)
    np.random.seed(4101)
    gen = srm.BatchGenerator([(data, data.copy())], batch_size=bs, collapse_axes=None, shuffle=True)
    assert np.array_equal(gen.indices, G["idx_perm"])
    for i in range(len(gen)):
        rows = [r for r in G["idx_batches"][i] if r >= 0]
        assert np.array_equal(gen[i][0].cpu().numpy(), data[rows])
    empty = srm.BatchGenerator([], batch_size=4)
    assert len(empty) == 0 and empty.N == 0


@pytest.mark.gpu
def test_gather_rows_odd_row_sizes_and_out_of_range_index():
    lib = srm._lib.load_library()
    dev = torch.device("cuda", 0)
    g = torch.Generator().manual_seed(5)
    for row_bytes in (7, 12, 48, 4096 + 16):
        n_rows = 37
        src = torch.randint(0, 256, (n_rows, row_bytes), dtype=torch.uint8, generator=g).to(dev)
        idx = torch.tensor([5, 0, 36, 36, 12, -1, 37, 3], dtype=torch.int32, device=dev)
        dst = torch.full((idx.numel(), row_bytes), 255, dtype=torch.uint8, device=dev)
        rc = lib.srm_gather_rows(0, src.data_ptr(), idx.data_ptr(), idx.numel(), n_rows, row_bytes, dst.data_ptr(), None)
        assert rc == 0
        torch.cuda.synchronize()
        want = torch.zeros_like(dst)
        for r, s in enumerate(idx.tolist()):
            if 0 <= s < n_rows:
                want[r] = src[s]
        assert torch.equal(dst, want), row_bytes
    assert lib.srm_gather_rows(0, None, None, 3, 1, 4, None, None) == -1
    assert b"srm_gather_rows" in lib.srm_last_error()
