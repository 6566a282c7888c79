"""HBM-tier walk on the config-5 instance (dev tool for ncu).  usage: python tools/prof_c5_walk.py [chains] [nsteps]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sls_sat_solving_with_deep_learning_b200 as eng
from sls_sat_solving_with_deep_learning_b200 import synth
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 300
n, m = 1_000_000, 4_000_000
row, lits = synth.random_ksat_csr(n, m, 3, 20260105)
inst = eng.Instance.from_csr(n, row, lits)
w = np.full(n, 0.5)
s = eng.ResidentSolver(inst, "probsat", nsteps, chains, False, True, 0.0, seed=1)
s.set_weights(w, w)
s.launch()
print(s.variant, "ms", s.sync())
