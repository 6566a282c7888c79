"""BASELINE config 4 on one GPU: batched oracle forward on synthetic random 3-SAT instances (n=500, alpha=4.2),
then oracle-guided Moser-Tardos on all of them (S = 10^4 steps, R = 100 chains each), one batched call per stage.
usage: python tools/bench_c4.py [n_instances]      (the config names 10,000 instances over 8 GPUs = 1250 per GPU)"""
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
from oracle import gnn_oracle as go  # noqa: E402  (random haiku-style parameters only; nothing is computed by the oracle here)

n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 1250
eng.set_device(0)
n, m = 500, 2100
t0 = time.perf_counter()
insts = []
for i in range(n_inst):
    row, lits = synth.random_ksat_csr(n, m, 3, 20260104 + i)
    insts.append(eng.Instance.from_csr(n, row, lits))
t_build = time.perf_counter() - t0
net = gnn.InteractionNetworkLCG(go.init_params(seed=42))
net.probabilities(insts[:2])  # warm
eng.run_sls_batch(insts[:4], "moser", [np.full(n, 0.5)] * 4, [np.full(n, 0.5)] * 4, 100, 100)
def pipeline():
    t0 = time.perf_counter()
    probs = net.probabilities(insts)
    t_gnn = time.perf_counter() - t0
    gnn_ms = net.last_kernel_ms
    t0 = time.perf_counter()
    res = eng.run_sls_batch(insts, "moser", probs, probs, 10_000 - 1, 100)
    t_mt = time.perf_counter() - t0
    steps = sum(r.total_steps for r in res)
    flips = sum(r.total_flips for r in res)
    solved = sum(bool(r.found.any()) for r in res)
    return {
        "oracle_forward": {"wall_s": round(t_gnn, 3), "device_ms": round(gnn_ms, 2), "instances_per_s": n_inst / t_gnn},
        "moser_tardos": {"wall_s": round(t_mt, 3), "kernel_ms": round(res[0].kernel_ms, 2), "steps_per_s": steps / (res[0].kernel_ms * 1e-3),
                         "flips_per_s": flips / (res[0].kernel_ms * 1e-3), "instances_solved": solved},
        "end_to_end_instances_per_s": n_inst / (t_gnn + t_mt),
    }


# first pass: instances are uploaded and the activation workspace (8 GB chunks) is allocated inside the timed calls;
# second pass: the same handles and workspace again, as every later batch of a 10,000-instance job finds them
first = pipeline()
steady = pipeline()
print(json.dumps({"instances": n_inst, "build_s": round(t_build, 3), "first_pass": first, "steady_state": steady}, indent=1))
