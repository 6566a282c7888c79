"""Clause evaluation of 4096 assignments on the C2 instance, three calls (for an ncu launch list; dev tool)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sls_sat_solving_with_deep_learning_b200 as eng
from sls_sat_solving_with_deep_learning_b200 import synth
n, m, B = 10_000, 42_000, 4096
row, lits = synth.random_ksat_csr(n, m, 3, 20260102)
inst = eng.Instance.from_csr(n, row, lits)
a = np.random.default_rng(0).integers(0, 2, size=(B, n)).astype(np.uint8)
for _ in range(3):
    viol, energy = eng.eval_assignments(inst, a)
print(energy[:4], viol.shape)
