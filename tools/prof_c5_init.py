"""HBM-tier chain initialisation on the config-5 instance (n=10^6, m=4*10^6), init only (dev tool for ncu).
usage: python tools/prof_c5_init.py [chains]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sls_sat_solving_with_deep_learning_b200 as eng
from sls_sat_solving_with_deep_learning_b200 import synth
chains = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n, m = 1_000_000, 4_000_000
row, lits = synth.random_ksat_csr(n, m, 3, 20260105)
inst = eng.Instance.from_csr(n, row, lits)
w = np.full(n, 0.5)
s = eng.ResidentSolver(inst, "probsat", 0, chains, False, True, 0.0, seed=1)
s.set_weights(w, w)
for _ in range(2):
    s.launch()
    ms = s.sync()
print("variant", s.variant, "chains", chains, "init ms", ms, "clause evals/s", chains * m / (ms * 1e-3))
# clause evaluation of the same instance through sls_eval_assignments (energy only): 1024 assignments
a = np.random.default_rng(0).integers(0, 2, size=(1024, n)).astype(np.uint8)
for _ in range(2):
    _, e = eng.eval_assignments(inst, a, want_bits=False)
print("eval 1024 x 4e6: kernel ms", eng.eval_last_kernel_ms(), "clause evals/s", 1024 * m / (eng.eval_last_kernel_ms() * 1e-3), e[:3])
