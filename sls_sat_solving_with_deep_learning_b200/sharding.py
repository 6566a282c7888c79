"""Sharding of the SLS path over the GPUs of one box (one process per GPU, torch.distributed).

The reference parallelises only over the runs of one call (rayon, rust/src/lib.rs:201) and loops
instances sequentially (python/src/evaluate_with_given_params.py:108).  Both axes are independent,
so here
  * chains of one instance are split into contiguous blocks of the global run index; a chain's
    Philox streams are keyed by that global index, so results do not depend on the number of ranks;
  * instances of a batch are split by estimated cost (longest-processing-time first).
The only collective of the path is the final gather of per-chain solved flags, step counts and best
energies (NCCL on GPUs, gloo in the CPU tests), plus one broadcast of the winning assignment.
"""
from __future__ import annotations

from typing import Callable, Sequence

import numpy as np


def chain_shard(nruns: int, rank: int, world: int) -> tuple[int, int]:
    """(offset, count) of the contiguous block of global run indices owned by `rank`."""
    base, rem = divmod(int(nruns), int(world))
    count = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, count


def instance_shard(costs: Sequence[float], world: int) -> list[list[int]]:
    """Longest-processing-time-first partition of instance indices; deterministic."""
    order = sorted(range(len(costs)), key=lambda i: (-float(costs[i]), i))
    loads = [0.0] * world
    parts: list[list[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: (loads[j], j))
        parts[r].append(i)
        loads[r] += float(costs[i])
    return [sorted(p) for p in parts]


def _all_gather_var(t, counts, dist, group=None):
    """all_gather of 1-D tensors with per-rank lengths `counts` (padded to the maximum)."""
    import torch

    mx = max(max(counts), 1)
    pad = torch.zeros(mx, dtype=t.dtype, device=t.device)
    pad[: t.numel()] = t
    out = [torch.zeros_like(pad) for _ in counts]
    dist.all_gather(out, pad, group=group)
    return torch.cat([o[:c] for o, c in zip(out, counts)])


def run_sls_sharded(
    runner: Callable,
    nruns: int,
    n_vars: int,
    m_clauses: int,
    *,
    device: str = "cpu",
    group=None,
):
    """Run `nruns` chains split over the ranks of the default process group and return the same
    result on every rank as one unsharded call would give.

    ``runner(offset, count)`` runs chains [offset, offset+count) on this rank (e.g.
    ``lambda off, cnt: eng.run_sls(inst, ..., nruns=cnt, seed=seed, chain_offset=off)``) and returns
    an object with found / steps / best_energy_run / flips_run / best_assignment.  Trajectories stay
    on the rank that produced them.
    """
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    counts = [chain_shard(nruns, r, world)[1] for r in range(world)]
    offset, count = chain_shard(nruns, rank, world)
    local = runner(offset, count)

    dev = torch.device(device)
    found = torch.as_tensor(np.asarray(local.found, dtype=np.uint8), device=dev)
    steps = torch.as_tensor(np.asarray(local.steps, dtype=np.int64), device=dev)
    best = torch.as_tensor(np.asarray(local.best_energy_run, dtype=np.int64), device=dev)
    flips = torch.as_tensor(np.asarray(local.flips_run, dtype=np.int64), device=dev)
    g_found = _all_gather_var(found, counts, dist, group).cpu().numpy().astype(bool)
    g_steps = _all_gather_var(steps, counts, dist, group).cpu().numpy().astype(np.uint64)
    g_best = _all_gather_var(best, counts, dist, group).cpu().numpy().astype(np.uint64)
    g_flips = _all_gather_var(flips, counts, dist, group).cpu().numpy().astype(np.uint64)

    # lib.rs:320-334: strict '<' arg-min over runs in run order, starting from formula.len()
    best_energy, winner = m_clauses, -1
    for r in range(nruns):
        if g_best[r] < best_energy:
            best_energy, winner = int(g_best[r]), r
    best_assignment = None
    if winner >= 0:
        bounds = np.cumsum([0] + counts)
        owner = int(np.searchsorted(bounds, winner, side="right") - 1)
        # the owner's local arg-min is the global winner (first strict minimum inside its block)
        buf = torch.zeros(max(n_vars, 1), dtype=torch.uint8, device=dev)
        if rank == owner:
            assert local.best_assignment is not None
            buf[:n_vars] = torch.as_tensor(np.asarray(local.best_assignment, dtype=np.uint8), device=dev)
        dist.broadcast(buf, src=dist.get_global_rank(group, owner) if group is not None else owner, group=group)
        best_assignment = buf[:n_vars].cpu().numpy()

    from .solver import SlsResult

    return SlsResult(
        best_assignment=best_assignment,
        found=g_found,
        best_energy=int(best_energy),
        steps=g_steps,
        best_energy_run=g_best,
        flips_run=g_flips,
        trajectories=list(getattr(local, "trajectories", [])),
        total_steps=int(g_steps.sum()),
        total_flips=int(g_flips.sum()),
        kernel_ms=float(getattr(local, "kernel_ms", 0.0)),
    )


def run_batch_sharded(runner: Callable, costs: Sequence[float], nruns: int, *, device: str = "cpu", group=None):
    """Partition a batch of instances over the ranks (:func:`instance_shard`), solve each rank's share
    with ``runner(indices) -> list of results`` (e.g. ``lambda idx: eng.run_sls_batch([insts[i] for i in
    idx], ...)`` -- pass ``seed`` handling inside the runner so that instance i keeps its own seed), and
    return on every rank the per-chain tables of the whole batch:
    ``found[n_inst, nruns]``, ``steps[n_inst, nruns]``, ``best_energy[n_inst]``, ``flips[n_inst]``.
    Every instance is owned by exactly one rank, so one sum-all_reduce of zero-initialised tables is
    the gather."""
    import torch
    import torch.distributed as dist

    rank, world = dist.get_rank(group), dist.get_world_size(group)
    n_inst = len(costs)
    mine = instance_shard(costs, world)[rank]
    results = runner(mine) if mine else []
    dev = torch.device(device)
    # this rank's rows are filled on the host and travel to the device as four tables (not two small copies per instance)
    h_found = np.zeros((n_inst, nruns), dtype=np.int32)
    h_steps = np.zeros((n_inst, nruns), dtype=np.int64)
    h_best = np.zeros(n_inst, dtype=np.int64)
    h_flips = np.zeros(n_inst, dtype=np.int64)
    for i, r in zip(mine, results):
        h_found[i] = np.asarray(r.found, dtype=np.int32)
        h_steps[i] = np.asarray(r.steps, dtype=np.int64)
        h_best[i] = int(r.best_energy)
        h_flips[i] = int(r.total_flips)
    found, steps, best, flips = (torch.from_numpy(a).to(dev) for a in (h_found, h_steps, h_best, h_flips))
    for t in (found, steps, best, flips):
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return found.cpu().numpy().astype(bool), steps.cpu().numpy().astype(np.uint64), best.cpu().numpy(), flips.cpu().numpy()
