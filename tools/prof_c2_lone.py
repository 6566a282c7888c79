"""One chain per SM-quadrant on the config-2 instance (dev tool for ncu: with a lone warp the stall samples per SASS
instruction are the dependent latency of a step).  usage: python tools/prof_c2_lone.py [chains] [nsteps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sls_sat_solving_with_deep_learning_b200 as eng
from sls_sat_solving_with_deep_learning_b200 import synth
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 148
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
n, m = 10_000, 42_000
row, lits = synth.random_ksat_csr(n, m, 3, 20260102)
inst = eng.Instance.from_csr(n, row, lits)
w = np.full(n, 0.5)
s = eng.ResidentSolver(inst, "probsat", nsteps, chains, False, True, 0.0, seed=1)
s.set_weights(w, w)
s.launch()
print(s.variant, "ms", s.sync())
