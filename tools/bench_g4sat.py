"""Generic-kernel throughput on pseudo-industrial instances (BASELINE config 3 shape) using the six
golden G4SAT fixtures replicated into a batch.  usage: python tools/bench_g4sat.py"""
import glob
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
import sls_sat_solving_with_deep_learning_b200 as eng  # noqa: E402

eng.set_device(0)
paths = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "instances", "g4sat_*.cnf")))
base = [eng.Instance.from_dimacs(p) for p in paths]
print([(os.path.basename(p)[:18], i.n, i.m, i.max_clause_len) for p, i in zip(paths, base)])
insts = [base[i % len(base)] for i in range(600)]
w = [np.full(i.n, 0.5) for i in insts]
out = {}
for algo in ("probsat", "moser"):
    eng.run_sls_batch(insts[:6], algo, w[:6], w[:6], 100, 100)
    t0 = time.perf_counter()
    res = eng.run_sls_batch(insts, algo, w, w, 10_000, 100, True, 0.0, seed=3)
    dt = time.perf_counter() - t0
    out[algo] = {"instances": 600, "chains": 100, "nsteps": 10000, "wall_s": dt, "kernel_ms": res[0].kernel_ms,
                 "steps_per_s": sum(r.total_steps for r in res) / (res[0].kernel_ms * 1e-3),
                 "flips_per_s": sum(r.total_flips for r in res) / (res[0].kernel_ms * 1e-3),
                 "solved": sum(bool(r.found.any()) for r in res)}
print(json.dumps(out, indent=1))
