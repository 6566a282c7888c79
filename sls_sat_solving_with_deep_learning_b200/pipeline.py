"""Oracle forward -> oracle-guided SLS over a large batch of instances (BASELINE configs[3]; the reference runs this loop
one instance at a time: evaluate_with_given_params.py:101-147 -- load the instance, network.apply, get_model_probabilities,
moser_rust.run_sls_python), with the three stages overlapped.

The stages of a chunk are (1) handles: CSR / DIMACS -> :class:`Instance` (host: occurrence lists, images), (2) the batched
forward (instance upload, LCG graphs on the device, the GNN kernels, probabilities back to the host) and (3) the batched
walk (weights -> thresholds on the host, one launch per kernel family, results back).  One call per stage over the whole
batch leaves the GPU idle while the host prepares the next stage -- on the config-4 shape the device is busy 3.2 s of 5.2 s.
Here the batch is cut into chunks that travel through three worker threads (the C library releases the GIL; its calls are
thread-safe: thread-local error state, a locked memory pool); the device work of one chunk runs under the host work of its
neighbours.  Results do not depend on the chunking: instance i is solved with ``instance_seeds[i]`` and chains are keyed
by (seed, chain id) only.
"""
from __future__ import annotations

import queue
import threading
from typing import Callable, Sequence

from . import solver as _solver

_STOP = object()


def staged(stages: Sequence[Callable], items, depth: int = 2) -> list:
    """Run ``stages[k](x)`` for every item, stage k of item i overlapping stage k+1 of item i-1 (one thread per stage, bounded
    queues of ``depth`` items between them).  Returns the last stage's outputs in item order; the first exception of any
    stage is re-raised after the threads have drained."""
    items = list(items)
    qs = [queue.Queue(maxsize=depth) for _ in range(len(stages) + 1)]
    failure: list[BaseException] = []

    def work(k):
        while True:
            x = qs[k].get()
            if x is _STOP:
                qs[k + 1].put(_STOP)
                return
            if failure:
                continue  # drain: keep the upstream producers from blocking on a full queue
            try:
                qs[k + 1].put(stages[k](x))
            except BaseException as e:  # noqa: BLE001 -- PanicException (the reference's panic) is a BaseException
                failure.append(e)

    threads = [threading.Thread(target=work, args=(k,), daemon=True) for k in range(len(stages))]
    for t in threads:
        t.start()
    out = []

    def feed():
        for x in items:
            if failure:
                break
            qs[0].put(x)
        qs[0].put(_STOP)

    feeder = threading.Thread(target=feed, daemon=True)
    feeder.start()
    while True:
        y = qs[-1].get()
        if y is _STOP:
            break
        out.append(y)
    feeder.join()
    for t in threads:
        t.join()
    if failure:
        raise failure[0]
    return out


def solve_with_oracle(net, sources, algo: str, nsteps: int, nruns: int, *, make_instance: Callable | None = None,
                      instance_seeds=None, chunk: int = 500, pre_compute_mapping: bool = True, prob_flip_best: float = 0.0,
                      device: int | None = None, keep_instances: bool = False, stats: dict | None = None):
    """``net.probabilities`` on every instance, then ``run_sls_batch`` with those probabilities as ``weights_initialize`` and
    ``weights_resample`` (the evaluation script's oracle-guided run, evaluate_with_given_params.py:116-147), chunk by chunk
    with the stages overlapped.

    ``sources``: one entry per instance -- an :class:`Instance`, or whatever ``make_instance`` turns into one (a DIMACS path,
    a ``(n, row_ptr, signed_lits)`` CSR triple ...).  ``instance_seeds``: one seed per instance (default 1 + index).
    Returns ``(results, probabilities)``: lists in source order of :class:`SlsResult` and of p[n_i]; with ``keep_instances``
    a third list holds the handles.  ``stats`` (optional dict) receives the summed stage times in seconds, the summed
    device times in ms and ``timeline``: (stage, first instance of the chunk, start s, end s) per stage call."""
    sources = list(sources)
    k = len(sources)
    seeds = list(range(1, k + 1)) if instance_seeds is None else [int(s) for s in instance_seeds]
    if len(seeds) != k:
        raise ValueError("instance_seeds needs one seed per instance")
    import time

    acc = {"build_s": 0.0, "forward_s": 0.0, "forward_device_ms": 0.0, "walk_s": 0.0, "walk_kernel_ms": 0.0, "timeline": []}
    t_origin = time.perf_counter()

    def stamp(stage, lo, t0):  # (stage, first instance of the chunk, start, end) in seconds since the call
        acc["timeline"].append((stage, lo, round(t0 - t_origin, 4), round(time.perf_counter() - t_origin, 4)))

    dev = _solver.current_device() if device is None else int(device)

    def pin_device():
        _solver.set_device(dev)  # the CUDA device is per thread: the workers follow the caller's

    def build(lo):
        t0 = time.perf_counter()
        hi = min(lo + chunk, k)
        insts = [s if isinstance(s, _solver.Instance) else make_instance(s) for s in sources[lo:hi]]
        acc["build_s"] += time.perf_counter() - t0
        stamp("build", lo, t0)
        return lo, insts

    def forward(x):
        lo, insts = x
        pin_device()
        t0 = time.perf_counter()
        probs = net.probabilities(insts)
        acc["forward_s"] += time.perf_counter() - t0
        acc["forward_device_ms"] += net.last_kernel_ms
        stamp("forward", lo, t0)
        return lo, insts, probs

    def walk(x):
        lo, insts, probs = x
        pin_device()
        t0 = time.perf_counter()
        res = _solver.run_sls_batch(insts, algo, probs, probs, nsteps, nruns, pre_compute_mapping, prob_flip_best,
                                    instance_seeds=seeds[lo:lo + len(insts)])
        acc["walk_s"] += time.perf_counter() - t0
        acc["walk_kernel_ms"] += res[0].kernel_ms if res else 0.0
        stamp("walk", lo, t0)
        return res, probs, (insts if keep_instances else None)

    if make_instance is None and any(not isinstance(s, _solver.Instance) for s in sources):
        raise ValueError("sources that are not Instance handles need make_instance")
    parts = staged([build, forward, walk], range(0, k, max(int(chunk), 1)))
    results = [r for p in parts for r in p[0]]
    probs = [q for p in parts for q in p[1]]
    if stats is not None:
        stats.update(acc)
    if keep_instances:
        return results, probs, [i for p in parts for i in p[2]]
    return results, probs
