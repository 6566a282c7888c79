"""Step latency of the C2 walk as a function of resident chains: probSAT on the config-2 instance with R chains
(R = 4 per SM ... 28 per SM).  One chain per warp, so kernel time / nsteps at small R is the dependent latency of one
step; at R = 4096 it is what the bench measures.  usage: python tools/bench_c2_occupancy.py [nsteps]"""
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
n, m = 10_000, 42_000
row, lits = synth.random_ksat_csr(n, m, 3, 20260102)
inst = eng.Instance.from_csr(n, row, lits)
w = np.full(n, 0.5)
for R in (148, 592, 1184, 2368, 4096, 8192):
    s = eng.ResidentSolver(inst, "probsat", nsteps, R, False, True, 0.0, seed=1)
    s.set_weights(w, w)
    for _ in range(2):
        s.launch()
        ms = s.sync()
    r = s.fetch_summary()
    s.close()
    print(json.dumps({"chains": R, "variant": s.variant if False else None, "kernel_ms": round(ms, 2), "ns_per_step_of_a_chain": ms * 1e6 / nsteps,
                      "flips_per_s": r.total_flips / (ms * 1e-3)}), flush=True)
