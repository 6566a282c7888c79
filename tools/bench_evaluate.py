"""The evaluation protocol with trajectories (load_model_and_test, keep_traj=True) on hard random 3-SAT instances:
wall time per instance.  usage: python tools/bench_evaluate.py [instances] [n_steps]"""
import os, pickle, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sls_sat_solving_with_deep_learning_b200 as eng
from sls_sat_solving_with_deep_learning_b200 import evaluate, synth

count = int(sys.argv[1]) if len(sys.argv) > 1 else 64
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 100_000
eng.set_device(0)
d = tempfile.mkdtemp()
for i in range(count):
    n, m = 200, 860  # alpha = 4.3: most chains use the whole step budget
    row, lits = synth.random_ksat_csr(n, m, 3, 1000 + i)
    synth.write_dimacs(os.path.join(d, f"i{i}.cnf"), n, row, lits)
    pickle.dump([1], open(os.path.join(d, f"i{i}_sol.pkl"), "wb"))
evaluate.load_model_and_test(d, "uniform", 1000, 100, "probsat", keep_traj=True, seed=1)  # warm
t0 = time.perf_counter()
out = evaluate.load_model_and_test(d, "uniform", n_steps, 100, "probsat", keep_traj=True, seed=1)
dt = time.perf_counter() - t0
print(f"{count} instances x 100 runs x {n_steps} steps with trajectories: {dt:.2f} s = {dt / count * 1e3:.1f} ms per instance; "
      f"mean final energy {float(out[4][-1]):.5f}")
