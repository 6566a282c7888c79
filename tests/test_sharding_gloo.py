"""N>1 path on CPU: world_size-2 gloo.  Each rank runs its block of chains (here through the CPU
oracle, which shares the chain_offset / Philox-keying contract with the CUDA engine), the sharded
driver gathers flags / steps / energies and must reproduce the unsharded result exactly."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from conftest import ROOT, golden_cnf
from sls_sat_solving_with_deep_learning_b200 import sharding


def test_chain_shard_partitions_exactly():
    for nruns in (0, 1, 7, 100, 4096, 4099):
        for world in (1, 2, 3, 8):
            blocks = [sharding.chain_shard(nruns, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and sum(c for _, c in blocks) == nruns
            for (o1, c1), (o2, _) in zip(blocks, blocks[1:]):
                assert o1 + c1 == o2
            assert max(c for _, c in blocks) - min(c for _, c in blocks) <= 1


def test_instance_shard_balances_cost():
    rng = np.random.default_rng(0)
    costs = rng.integers(100, 30000, size=600).tolist()
    parts = sharding.instance_shard(costs, 8)
    assert sorted(i for p in parts for i in p) == list(range(600))
    loads = [sum(costs[i] for i in p) for p in parts]
    assert max(loads) / (sum(loads) / 8) < 1.02
    assert sharding.instance_shard(costs, 8) == parts  # deterministic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, algo, nruns, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import sls_oracle as so

    oi = so.Instance.from_dimacs(golden_cnf("r3sat_test_random_KCNF3_200_869_9397183"))
    w = np.full(oi.n, 0.5)

    def runner(offset, count):
        return oi.run_sls(algo, w, w, 300, count, False, True, 0.0, seed=17, chain_offset=offset, nthreads=1)

    res = sharding.run_sls_sharded(runner, nruns, oi.n, oi.m)
    if rank == 0:
        np.savez(out, found=res.found, steps=res.steps, best_run=res.best_energy_run, flips=res.flips_run,
                 best=res.best_energy, assignment=res.best_assignment, total_flips=res.total_flips)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("algo,nruns", [("moser", 13), ("probsat", 8)])
def test_two_rank_run_equals_single_rank(tmp_path, algo, nruns):
    from oracle import sls_oracle as so

    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), algo, nruns, out), nprocs=2, join=True)
    got = np.load(out)
    oi = so.Instance.from_dimacs(golden_cnf("r3sat_test_random_KCNF3_200_869_9397183"))
    w = np.full(oi.n, 0.5)
    ref = oi.run_sls(algo, w, w, 300, nruns, False, True, 0.0, seed=17)
    assert (got["found"] == ref.found).all() and (got["steps"] == ref.steps).all()
    assert (got["best_run"] == ref.best_energy_run).all() and (got["flips"] == ref.flips_run).all()
    assert int(got["best"]) == ref.best_energy and int(got["total_flips"]) == ref.total_flips
    assert (got["assignment"] == ref.best_assignment).all()


def _batch_worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import json

    from oracle import sls_oracle as so

    names = sorted(json.load(open(os.path.join(ROOT, "tests", "golden", "expected.json"))))[:7]
    insts = [so.Instance.from_dimacs(golden_cnf(nm)) for nm in names]

    def runner(indices):  # instance i keeps seed 40 + i wherever it runs (the sls_run_batch convention)
        return [insts[i].run_sls("probsat", np.full(insts[i].n, 0.5), np.full(insts[i].n, 0.5), 200, 6, False, True, 0.0,
                                 seed=40 + i, nthreads=1) for i in indices]

    found, steps, best, flips = sharding.run_batch_sharded(runner, [i.num_lits for i in insts], 6)
    if rank == 0:
        np.savez(out, found=found, steps=steps, best=best, flips=flips)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_instance_sharding_equals_single_rank(tmp_path):
    import json

    from oracle import sls_oracle as so

    out = str(tmp_path / "batch.npz")
    mp.spawn(_batch_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    got = np.load(out)
    names = sorted(json.load(open(os.path.join(ROOT, "tests", "golden", "expected.json"))))[:7]
    for i, nm in enumerate(names):
        oi = so.Instance.from_dimacs(golden_cnf(nm))
        w = np.full(oi.n, 0.5)
        ref = oi.run_sls("probsat", w, w, 200, 6, False, True, 0.0, seed=40 + i, nthreads=1)
        assert (got["found"][i] == ref.found).all() and (got["steps"][i] == ref.steps).all()
        assert got["best"][i] == ref.best_energy and got["flips"][i] == ref.total_flips
