"""Soak of the oracle forward: random batch sizes, shapes, cluster settings and product forms for a given number of seconds.
A hang shows as the caller's timeout (run it under `timeout`); progress goes to stdout line by line.
usage: python tools/gnn_soak.py [seconds] [seed]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sls_sat_solving_with_deep_learning_b200 as eng  # noqa: E402
from sls_sat_solving_with_deep_learning_b200 import gnn, synth  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 60.0
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 1)
eng.set_device(0)
net = gnn.InteractionNetworkLCG(synth.random_oracle_params(seed=42))
pool = {}


def instances(n, m, count):
    key = (n, m)
    have = pool.setdefault(key, [])
    while len(have) < count:
        have.append(eng.Instance.from_csr(n, *synth.random_ksat_csr(n, m, 3, 7000 + 31 * len(have) + n)))
    return have[:count]


t0, calls, first = time.perf_counter(), 0, {}
while time.perf_counter() - t0 < seconds:
    n = int(rng.choice([60, 200, 500, 500, 500, 1000]))
    m = int(n * rng.uniform(2.0, 4.5))
    count = int(rng.choice([1, 2, 3, 5, 8, 13, 21, 34, 55, 64, 89, 128, 144, 200, 256]))
    if n == 1000:
        count = min(count, 64)
    os.environ["SLS_B200_GNN_CLUSTER"] = str(rng.choice([1, 2]))
    os.environ["SLS_B200_GNN_PRECISION"] = str(rng.choice(["fp16x2", "fp16x2", "bf16x2", "tf32x3"]))
    insts = instances(n, m, count)
    p = net.probabilities(insts)
    assert all(np.isfinite(x).all() for x in p)
    key = (n, m, os.environ["SLS_B200_GNN_PRECISION"])
    if key in first:
        assert np.array_equal(p[0], first[key]), key  # instance 0 of a shape: same bits whatever the batch and cluster size
    else:
        first[key] = p[0]
    calls += 1
    if calls % 25 == 0:
        print(f"{calls} forwards, {time.perf_counter() - t0:.0f} s, last: n={n} m={m} x{count} {os.environ['SLS_B200_GNN_PRECISION']} "
              f"cluster {os.environ['SLS_B200_GNN_CLUSTER']} {net.last_kernel_ms:.2f} ms", flush=True)
print(f"done: {calls} forwards in {time.perf_counter() - t0:.0f} s, no hang, no mismatch")
