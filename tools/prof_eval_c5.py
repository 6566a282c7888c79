"""Clause evaluation of 1024 assignments on the C5 instance (n = 10^6, m = 4 * 10^6) through the C ABI, with and without
the violated-clause bits (for ncu; dev tool).  usage: python tools/prof_eval_c5.py [bits|energy]"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sls_sat_solving_with_deep_learning_b200 as eng
from sls_sat_solving_with_deep_learning_b200 import _lib, synth

mode = sys.argv[1] if len(sys.argv) > 1 else "bits"
n, m, B = 1_000_000, 4_000_000, 1024
row, lits = synth.random_ksat_csr(n, m, 3, 20260105)
inst = eng.Instance.from_csr(n, row, lits)
a = np.random.default_rng(3).integers(0, 2, size=(B, n), dtype=np.uint8)
energy = np.zeros(B, dtype=np.uint32)
bits = np.zeros((B, (m + 31) // 32), dtype=np.uint32) if mode == "bits" else None
lib = _lib.load()
for _ in range(2):
    rc = lib.sls_eval_assignments(inst._h, a.ctypes.data, B, bits.ctypes.data if bits is not None else None, energy.ctypes.data)
    assert rc == 0
    print(mode, "device ms", eng.eval_last_kernel_ms(), "evaluation kernel ms", eng.eval_last_eval_kernel_ms(), energy[:3])
