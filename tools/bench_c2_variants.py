"""BASELINE config 2 (n=10^4, alpha=4.2, 4096 chains x 10^5 steps) for every flip rule, with and without the greedy
mix (prob_flip_best = 0.5), and with trajectory recording.  usage: python tools/bench_c2_variants.py [nsteps]"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as g  # noqa: E402

g.build()
import sls_sat_solving_with_deep_learning_b200 as eng  # noqa: E402
from sls_sat_solving_with_deep_learning_b200 import synth  # noqa: E402

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
eng.set_device(0)
n, m, R = 10_000, 42_000, 4096
row, lits = synth.random_ksat_csr(n, m, 3, 20260102)
inst = eng.Instance.from_csr(n, row, lits)
w = np.full(n, 0.5)
out = []
for algo, pfb, traj, steps in [("probsat", 0.0, False, nsteps), ("probsat", 0.5, False, nsteps // 4), ("moser", 0.0, False, nsteps),
                               ("moser_single", 0.0, False, nsteps), ("schoening", 0.0, False, nsteps),
                               ("probsat", 0.0, True, min(nsteps, 10_000))]:
    s = eng.ResidentSolver(inst, algo, steps, R, traj, True, pfb, seed=1)
    s.set_weights(w, w)
    for _ in range(2):
        s.launch()
        ms = s.sync()
    r = s.fetch_summary()
    out.append({"algo": algo, "prob_flip_best": pfb, "trajectories": traj, "nsteps": steps, "variant": s.variant, "kernel_ms": round(ms, 2),
                "flips_per_s": r.total_flips / (ms * 1e-3), "steps_per_s": r.total_steps / (ms * 1e-3),
                "mean_best_energy": float(r.best_energy_run.mean()), "solved": int(r.found.sum())})
    s.close()
    print(json.dumps(out[-1]), flush=True)
