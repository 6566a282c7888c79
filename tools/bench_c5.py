"""BASELINE config 5 at reduced chain count: random 3-SAT n=10^6, alpha=4.0 (m=4*10^6), chain state in HBM.
usage: python tools/bench_c5.py [chains] [nsteps]"""
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
from oracle import sls_oracle as so  # noqa: E402

chains = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
eng.set_device(0)
n, m = 1_000_000, 4_000_000
t0 = time.perf_counter()
row, lits = synth.random_ksat_csr(n, m, 3, 20260105)
print(f"generated in {time.perf_counter()-t0:.1f} s", flush=True)
inst = eng.Instance.from_csr(n, row, lits)
w = np.full(n, 0.5)
out = {}
for algo in ("probsat", "moser"):
    s = eng.ResidentSolver(inst, algo, nsteps, chains, False, True, 0.0, seed=1)
    s.set_weights(w, w)
    s.launch()
    ms = s.sync()
    r = s.fetch()
    out[algo] = {"variant": s.variant, "chains": chains, "nsteps": nsteps, "kernel_ms": ms,
                 "flips_per_s": r.total_flips / (ms * 1e-3), "steps_per_s": r.total_steps / (ms * 1e-3),
                 "best_energy": r.best_energy, "mean_best": float(r.best_energy_run.mean())}
    # init-only cost (nsteps = 0) to separate the full scan from the walk
    s0 = eng.ResidentSolver(inst, algo, 0, chains, False, True, 0.0, seed=1)
    s0.set_weights(w, w)
    s0.launch()
    out[algo]["init_ms"] = s0.sync()
    s0.close()
    s.close()
# two chains against the CPU oracle (trajectory-exact)
oi = so.Instance.from_csr if hasattr(so.Instance, "from_csr") else None
g2 = eng.run_sls(inst, "probsat", w, w, 200, 2, True, True, 0.0, seed=9)
clauses_row, clauses_lits = row, lits
import ctypes as C
h = C.c_void_p()
so._check(so.lib().orc_instance_from_csr(n, m, row.ctypes.data, lits.ctypes.data, C.byref(h)))
o2 = so.Instance(h).run_sls("probsat", w, w, 200, 2, True, True, 0.0, seed=9)
out["oracle_match"] = bool(all((a == b).all() for a, b in zip(g2.trajectories, o2.trajectories)))
print(json.dumps(out, indent=1))
