"""Oracle forward -> oracle-guided SLS over a large batch of instances (BASELINE configs[3]; the reference runs this loop
one instance at a time: evaluate_with_given_params.py:101-147 -- load the instance, network.apply, get_model_probabilities,
moser_rust.run_sls_python), with the three stages overlapped.

The stages of a chunk are (1) handles: CSR / DIMACS -> :class:`Instance` (host: occurrence lists, images), (2) the batched
forward (instance upload, LCG graphs on the device, the GNN kernels, probabilities back to the host) and (3) the batched
walk (weights -> thresholds on the host, one launch per kernel family, results back).  One call per stage over the whole
batch leaves the GPU idle while the host prepares the next stage -- on the config-4 shape the device is busy 3.2 s of 5.2 s.
Here the batch is cut into chunks that travel through three (or, with two model handles, four) worker threads (the C library releases the GIL; its calls are
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


def staged(stages: Sequence, items, depth: int = 2) -> list:
    """Run the stages over every item, stage k of item i overlapping stage k+1 of item i-1.  A stage is a callable (one worker
    thread) or a list of callables (one worker thread each, taking items as they come -- e.g. two forward workers on two model
    handles); bounded queues of ``depth`` items connect the stages and every stage hands its results on in ITEM ORDER.  Returns
    the last stage's outputs in item order; the first exception of any stage is re-raised after the threads have drained."""
    items = list(items)
    stages = [list(s) if isinstance(s, (list, tuple)) else [s] for s in stages]
    qs = [queue.Queue(maxsize=depth) for _ in range(len(stages) + 1)]
    failure: list[BaseException] = []
    turn = [threading.Condition() for _ in stages]
    next_out = [0] * len(stages)      # sequence number the stage hands on next
    live = [len(s) for s in stages]   # workers of the stage that have not seen the end of the input yet

    def work(k, fn):
        while True:
            x = qs[k].get()
            if x is _STOP:
                qs[k].put(_STOP)  # for the stage's other workers
                with turn[k]:
                    live[k] -= 1
                    last = live[k] == 0
                if last:
                    qs[k].get()  # the marker this worker has just put back
                    qs[k + 1].put(_STOP)
                return
            seq, val = x
            y = None
            if not failure:
                try:
                    y = fn(val)
                except BaseException as e:  # noqa: BLE001 -- PanicException (the reference's panic) is a BaseException
                    failure.append(e)
            with turn[k]:  # hand on in item order (a failed or skipped item only advances the counter)
                while next_out[k] != seq:
                    turn[k].wait()
            if not failure:
                qs[k + 1].put((seq, y))
            with turn[k]:
                next_out[k] += 1
                turn[k].notify_all()

    threads = [threading.Thread(target=work, args=(k, fn), daemon=True) for k, fns in enumerate(stages) for fn in fns]
    for t in threads:
        t.start()
    out = []

    def feed():
        for seq, x in enumerate(items):
            if failure:
                break
            qs[0].put((seq, x))
        qs[0].put(_STOP)

    feeder = threading.Thread(target=feed, daemon=True)
    feeder.start()
    while True:
        y = qs[-1].get()
        if y is _STOP:
            break
        out.append(y[1])
    feeder.join()
    for t in threads:
        t.join()
    if failure:
        raise failure[0]
    return out


def solve_with_oracle(net, sources, algo: str, nsteps: int, nruns: int, *, make_instance: Callable | None = None,
                      instance_seeds=None, chunk: int = 500, pre_compute_mapping: bool = True, prob_flip_best: float = 0.0,
                      device: int | None = None, keep_instances: bool = False, stats: dict | None = None, reserve: bool = True):
    """``net.probabilities`` on every instance, then ``run_sls_batch`` with those probabilities as ``weights_initialize`` and
    ``weights_resample`` (the evaluation script's oracle-guided run, evaluate_with_given_params.py:116-147), chunk by chunk
    with the stages overlapped.

    ``net``: an :class:`~.gnn.InteractionNetworkLCG`, or a list of them built from the same parameters -- one forward worker
    thread per handle (a handle owns its activation workspace, so two handles let the forward of chunk i+2 be queued while
    the walk of chunk i+1 is still being prepared on the host).
    ``sources``: one entry per instance -- an :class:`Instance`, or whatever ``make_instance`` turns into one (a DIMACS path,
    a ``(n, row_ptr, signed_lits)`` CSR triple ...).  ``instance_seeds``: one seed per instance (default 1 + index).
    Returns ``(results, probabilities)``: lists in source order of :class:`SlsResult` and of p[n_i]; with ``keep_instances``
    a third list holds the handles.  ``reserve``: map the device memory pool's working set
    before the first upload (see :func:`~.solver.reserve_device_pool`).  ``stats`` (optional dict) receives the summed stage times in seconds, the summed
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
        if lo == 0 and reserve and insts:
            # the device images of all instances (about 24 bytes per literal occurrence, 16 per clause, 8 per variable) and the
            # per-chunk buffers of two chunks in flight, mapped now: growing the pool later would wait for the walk kernels
            per = sum(24 * i.num_lits + 16 * i.m + 8 * i.n + 4096 for i in insts) / len(insts)
            pin_device()
            _solver.reserve_device_pool(int(1.25 * per * k + 3 * per * len(insts)) + (256 << 20))
        acc["build_s"] += time.perf_counter() - t0
        stamp("build", lo, t0)
        return lo, insts

    lock = threading.Lock()

    def forward_on(one_net):
        def forward(x):
            lo, insts = x
            pin_device()
            t0 = time.perf_counter()
            probs = one_net.probabilities(insts)
            with lock:
                acc["forward_s"] += time.perf_counter() - t0
                acc["forward_device_ms"] += one_net.last_kernel_ms
                stamp("forward", lo, t0)
            return lo, insts, probs
        return forward

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
    nets = list(net) if isinstance(net, (list, tuple)) else [net]
    parts = staged([build, [forward_on(one) for one in nets], walk], range(0, k, max(int(chunk), 1)))
    results = [r for p in parts for r in p[0]]
    probs = [q for p in parts for q in p[1]]
    if stats is not None:
        stats.update(acc)
    if keep_instances:
        return results, probs, [i for p in parts for i in p[2]]
    return results, probs
