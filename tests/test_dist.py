"""N > 1 path on CPU: world_size-2 gloo.  Ranks take disjoint realisation shards, evaluate the loss
terms of their shard (the oracle stands in for the kernels here -- no GPU), all-reduce the 2x8
term vector with the product's dist.allreduce_terms, and must reproduce the single-process terms."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util as U

srm, O = U.srm, U.O


def test_shard_realisations_partition():
    for K in (1, 2, 5, 8, 64):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = srm.dist.shard_realisations(K, r, world)
                assert 0 <= lo <= hi <= K
                seen += list(range(lo, hi))
            assert seen == list(range(K))
            sizes = [srm.dist.shard_realisations(K, r, world) for r in range(world)]
            assert max(h - l for l, h in sizes) - min(h - l for l, h in sizes) <= 1
    assert srm.dist.shard_samples(4, 3, 1, 2) == (6, 12)
    with pytest.raises(ValueError):
        srm.dist.shard_realisations(4, 2, 2)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, K, T, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ocfg, otab, spec, ptab, batch = U.make_case(W=9, H=8, D=2, T=T, K=K, seed=77, all_layers=True)
        lo, hi = srm.dist.shard_samples(K, T, rank, world)
        rlo, rhi = srm.dist.shard_realisations(K, rank, world)
        o = O.dg_forward_backward(ocfg, otab, batch.kx[rlo:rhi].numpy(), batch.p0[lo:hi].numpy(), batch.p1[lo:hi].numpy(),
                                  batch.dt1[lo:hi].numpy(), batch.dt2[lo:hi].numpy(), batch.t1[lo:hi].numpy(),
                                  batch.sample_real[lo:hi].numpy() - rlo, U.WEIGHTS)
        terms = torch.zeros(2, 8, dtype=torch.float64)
        terms[0] = torch.from_numpy(o["terms"].astype(np.float64))
        terms[1] = torch.from_numpy(O.dg_counts(ocfg, hi - lo))
        srm.dist.allreduce_terms(terms)
        if rank == 0:
            np.save(out, terms.numpy())
    finally:
        dist.destroy_process_group()


def test_two_rank_sharded_terms_equal_single_process(tmp_path):
    K, T = 4, 2
    out = str(tmp_path / "terms.npy")
    mp.spawn(_worker, args=(2, _free_port(), K, T, out), nprocs=2, join=True)
    got = np.load(out)
    ocfg, otab, spec, ptab, batch = U.make_case(W=9, H=8, D=2, T=T, K=K, seed=77, all_layers=True)
    full = U.oracle_run(ocfg, otab, batch)
    assert np.allclose(got[0], full["terms"], rtol=1e-6, atol=0)
    assert np.array_equal(got[1], O.dg_counts(ocfg, K * T))
