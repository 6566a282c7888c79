"""Where the time of one reference-facing call goes (dev tool): phases of run_sls_python on the C2 workload."""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sls_sat_solving_with_deep_learning_b200 as eng
from sls_sat_solving_with_deep_learning_b200 import synth
from sls_sat_solving_with_deep_learning_b200.solver import Instance, ResidentSolver

n, m, R, T = 10000, 42000, 4096, 100000
row, lits = synth.random_ksat_csr(n, m, 3, 20260102)
tmp = tempfile.mkdtemp(); path = os.path.join(tmp, "c2.cnf"); synth.write_dimacs(path, n, row, lits)
w = np.full(n, 0.5)
def now(): return time.perf_counter()
for it in range(8):
    t = [now()]
    inst = Instance.from_dimacs(path); t.append(now())
    s = ResidentSolver(inst, "probsat", T, R, False, True, 0.0, seed=100 + it); t.append(now())
    s.set_weights(w, w); t.append(now())
    s.launch(); t.append(now())
    kms = s.sync(); t.append(now())
    res = s.fetch_summary(); t.append(now())
    s.close(); inst.close(); t.append(now())
    tup = res.as_reference_tuple(); t.append(now())
    names = ["parse", "create", "weights", "launch", "sync", "fetch", "close", "tuple"]
    print(it, " ".join(f"{k}={1e3*(b-a):.2f}" for k, a, b in zip(names, t, t[1:])), f"kernel_ms={kms:.2f} total={1e3*(t[-1]-t[0]):.2f}")
import moser_rust
for it in range(8):
    t0 = now(); moser_rust.run_sls_python("probsat", path, w, w, T, R, False, True, 0); print("run_sls_python ms", 1e3 * (now() - t0))
