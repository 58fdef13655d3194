"""World-size-2 gloo tests (CPU) of the data-parallel plumbing: contiguous walker sharding (main.py:231-238) and
the single statistics all-reduce that replaces jax.lax.pmean (src/MCMC.py:39, src/VMC.py:46-53).  The sharded
result must equal the single-process result on the concatenated walkers."""
import os
import sys
import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from coulombgas_b200 import dist as cd
    from oracle import potential as op, tables
    rng = np.random.default_rng(3)
    B, n, L = 8, 5, 4.0
    x = rng.uniform(0, L, (B, n, 2))                      # the same global batch on every rank
    lo, hi = cd.shard_bounds(B)
    G = tables.kpoints(2, 5)
    V = torch.as_tensor(op.potential_energy(x[lo:hi], 6, G, L, 2.0))      # this rank's shard (CPU oracle stands in)
    n_acc = float(np.arange(B)[lo:hi].sum())              # fake per-shard accept counts
    rate, vm, v2 = cd.accept_and_energy_stats(n_acc, 10 * (hi - lo), V)
    pm = cd.pmean_of_means(torch.tensor([V.mean()], dtype=torch.float64))
    Vall = op.potential_energy(x, 6, G, L, 2.0)
    ok = (abs(float(rate) - np.arange(B).sum() / (10 * B)) < 1e-15 and abs(float(vm) - Vall.mean()) < 1e-12
          and abs(float(v2) - (Vall ** 2).mean()) < 1e-10 and abs(float(pm) - Vall.mean()) < 1e-12
          and (lo, hi) == (rank * B // world, (rank + 1) * B // world))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_sharded_stats_equal_global():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_shard_bounds_rules():
    from coulombgas_b200 import dist as cd
    assert cd.shard_bounds(2048, 8, 3) == (768, 1024)
    with pytest.raises(ValueError):                        # main.py:231-233
        cd.shard_bounds(2050, 8, 0)
    assert cd.shard_bounds(16) == (0, 16)
