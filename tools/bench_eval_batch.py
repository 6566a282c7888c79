"""Secondary measurements (not the headline bench): clause evaluation throughput (row a17) and the
batched many-instance solver (BASELINE config-1 shape).  usage: python tools/bench_eval_batch.py"""
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
from sls_sat_solving_with_deep_learning_b200 import synth  # noqa: E402

eng.set_device(0)
out = {}

# ---- clause evaluation, config-2 instance, 4096 assignments (host buffers in, host buffers out)
n, m = 10_000, 42_000
row, lits = synth.random_ksat_csr(n, m, 3, 20260102)
inst = eng.Instance.from_csr(n, row, lits)
rng = np.random.default_rng(0)
a = rng.integers(0, 2, size=(4096, n), dtype=np.uint8)
eng.eval_assignments(inst, a[:64])
t0 = time.perf_counter()
viol, energy = eng.eval_assignments(inst, a)
dt = time.perf_counter() - t0
out["eval_c2"] = {"assignments": 4096, "clauses": m, "wall_ms": dt * 1e3,
                  "clause_evals_per_s": 4096 * m / dt, "mean_energy": float(energy.mean())}

# ---- batched solver, config-1 shape: 2052 instances n in {100,200,300}, alpha 1.0..4.5, 100 chains each
insts = []
for i in range(2052):
    nn = 100 + (i % 3) * 100
    alpha = 1.0 + 3.5 * ((i * 7919) % 2052) / 2051
    r, l = synth.random_ksat_csr(nn, int(nn * alpha), 3, 500000 + i)
    insts.append(eng.Instance.from_csr(nn, r, l))
w = [np.full(i.n, 0.5) for i in insts]
eng.run_sls_batch(insts[:8], "probsat", w[:8], w[:8], 100, 100)
for algo in ("probsat", "moser"):
    t0 = time.perf_counter()
    res = eng.run_sls_batch(insts, algo, w, w, 100_000, 100, True, 0.0, seed=1)
    dt = time.perf_counter() - t0
    solved = sum(bool(r.found.any()) for r in res)
    flips = sum(r.total_flips for r in res)
    steps = sum(r.total_steps for r in res)
    out[f"batch_c1_{algo}"] = {"instances": len(insts), "chains_per_instance": 100, "nsteps": 100_000, "wall_s": dt,
                               "kernel_ms": res[0].kernel_ms, "instances_solved": solved,
                               "instances_solved_per_s": solved / dt, "flips_per_s_kernel": flips / (res[0].kernel_ms * 1e-3),
                               "steps_per_s_kernel": steps / (res[0].kernel_ms * 1e-3)}
print(json.dumps(out, indent=1))
