"""Oracle-forward throughput on the BASELINE config-4 shape: a batch of synthetic random 3-SAT
instances n=500, alpha=4.2 (N=3100 nodes, E=6800 edges each), haiku-style random parameters.
usage: python tools/bench_gnn.py [n_instances] [repeats]"""
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
from sls_sat_solving_with_deep_learning_b200 import gnn, synth  # noqa: E402
from oracle import gnn_oracle as go  # noqa: E402

n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 256
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
eng.set_device(0)
n, m = 500, 2100
insts = []
for i in range(n_inst):
    row, lits = synth.random_ksat_csr(n, m, 3, 20260104 + i)
    insts.append(eng.Instance.from_csr(n, row, lits))
params = go.init_params(seed=42)
net = gnn.InteractionNetworkLCG(params)
N, E = 2 * n + m, 3 * m + n
flops = n_inst * (1.3984e6 * E + 1.5328e6 * N)  # SURVEY section 8 a15 (dense FLOPs as the reference computes them)
net.probabilities(insts[:2])  # warm
best = 1e9
for r in range(reps):
    t0 = time.perf_counter()
    p = net.probabilities(insts)
    wall = time.perf_counter() - t0
    best = min(best, net.last_kernel_ms)
    print(f"rep {r}: device {net.last_kernel_ms:.2f} ms, wall {wall*1e3:.1f} ms, launches {net.launch_count()}")
net.set_profile(True)
net.probabilities(insts)
prof = net.profile()
net.set_profile(False)
print("profile (ms, events around every launch):", json.dumps(prof))
# spot-check one instance against the fp64 restatement
row, lits = insts[0].csr()
clauses = [lits[int(row[c]):int(row[c + 1])].tolist() for c in range(m)]
ref = go.model_probabilities(go.forward(params, go.lcg_graph(n, clauses), dtype=np.float64), n)
err = float(np.abs(p[0] - ref).max())
print(json.dumps({"instances": n_inst, "device_ms": best, "instances_per_s": n_inst / (best * 1e-3),
                  "dense_tflops": flops / (best * 1e-3) / 1e12, "precision": os.environ.get("SLS_B200_GNN_PRECISION", "fp16x2"), "mma_tflops_issued": 3 * flops / (best * 1e-3) / 1e12,
                  "max_abs_dp_vs_fp64": err}))
