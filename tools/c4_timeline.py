"""Gantt view of pipeline.solve_with_oracle on the config-4 shape: which stage of which chunk ran when (host wall clock).
usage: python tools/c4_timeline.py [instances] [chunk] [model handles]"""
import gc
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import sls_sat_solving_with_deep_learning_b200 as eng  # noqa: E402
from sls_sat_solving_with_deep_learning_b200 import gnn, pipeline, synth  # noqa: E402

k = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
chunk = int(sys.argv[2]) if len(sys.argv) > 2 else 500
handles = int(sys.argv[3]) if len(sys.argv) > 3 else 1
eng.set_device(0)
n, m = 500, 2100
csrs = [synth.random_ksat_csr(n, m, 3, 20260104 + i) for i in range(k)]
params = synth.random_oracle_params(seed=42)
nets = [gnn.InteractionNetworkLCG(params) for _ in range(handles)]
net = nets[0]
warm = [eng.Instance.from_csr(n, *c) for c in csrs[:chunk]]  # a chunk of the real size: pool and workspace at working size
pw = [n_.probabilities(warm) for n_ in nets][0]
eng.run_sls_batch(warm, "moser", pw, pw, 9999, 100)
del warm, pw
if os.environ.get("C4_NO_GC"):
    gc.disable()
st = {}
t0 = time.perf_counter()
pipeline.solve_with_oracle(nets if handles > 1 else net, csrs, "moser", 9999, 100, make_instance=lambda c: eng.Instance.from_csr(n, *c), chunk=chunk, stats=st)
wall = time.perf_counter() - t0
print(f"{k} instances, chunk {chunk}, {handles} model handle(s): wall {wall:.3f} s = {k / wall:.0f} instances/s; forward device {st['forward_device_ms']:.0f} ms, "
      f"walk kernels {st['walk_kernel_ms']:.0f} ms")
for stage, lo, a, b in sorted(st["timeline"], key=lambda x: x[2]):
    if stage != "build" and (b - a) * 1e3 > 1.5 * (1e3 * wall / (k / chunk)) or os.environ.get("C4_ALL"):
        print(f"{stage:8s} chunk@{lo:6d}  {a:7.3f} -> {b:7.3f}  ({(b - a) * 1e3:6.0f} ms)")
