"""World-size-2 test (gloo, CPU) of the N>1 host logic: member sharding is a partition of the global
perturbation table, the timed-region reduction is a max over ranks, and the final diagnostics gather
reassembles the ensemble in global member order.  No timestep collective exists to test."""
import os
import socket

import numpy as np
import torch.distributed as dist
import torch.multiprocessing as mp

from cgenie_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, mpr, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tab = sharding.perturbation_table(world * mpr)
    mine = sharding.shard(tab, rank, world, mpr)
    lo, hi = sharding.shard_bounds(rank, world, mpr)
    # a per-member "diagnostic": something only the owner can compute from its shard
    diag = np.stack([mine["diff1"] * 1e-3 + mine["scf"], np.arange(lo, hi, dtype=float)], axis=1)
    allv = sharding.gather_diagnostics(diag, dist)
    tmax = sharding.max_over_ranks(10.0 + rank, dist)
    dist.barrier()
    if rank == 0:
        q.put((allv, tmax))
    dist.destroy_process_group()


def test_two_rank_sharding_and_gather():
    world, mpr = 2, 24
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, mpr, q)) for r in range(world)]
    for p in procs:
        p.start()
    allv, tmax = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    tab = sharding.perturbation_table(world * mpr)
    assert tmax == 11.0
    assert np.array_equal(allv[:, 1], np.arange(world * mpr))
    assert np.array_equal(allv[:, 0], tab["diff1"] * 1e-3 + tab["scf"])


def test_table_properties():
    t64 = sharding.perturbation_table(64)
    t1024 = sharding.perturbation_table(1024)
    for k in sharding.PERTURBED:
        assert t64[k][0] == sharding.BASE[k]                       # member 0 is the control
        assert np.array_equal(t64[k], t1024[k][:64])               # prefix-stable: shards agree across world sizes
        r = t1024[k] / sharding.BASE[k]
        assert r.min() >= 0.8 and r.max() <= 1.25
    # adrag is shared per group so barotropic factorisations are shared
    assert len(np.unique(t1024["adrag"])) <= 1024 // sharding.ADRAG_GROUP
    lo, hi = sharding.shard_bounds(3, 8, 128)
    assert (lo, hi) == (384, 512)


def test_ensemble_groups_argument_check():
    """group size outside 1..128 is refused before anything touches the device"""
    import pytest
    from cgenie_b200 import EnsembleGroups
    for bad in (0, 129):
        with pytest.raises(ValueError):
            EnsembleGroups("/nonexistent", n_members=4, group=bad)
