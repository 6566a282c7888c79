"""How far do the two product forms of the oracle-forward kernel (fp16x2, bf16x2, tf32x3) stay from an fp64 forward when the
parameters are NOT haiku's mild random init: weights scaled up, LayerNorm scales / offsets spread out, graphs with long
clauses (G4SAT shapes, k up to 297).  usage: python tools/gnn_harsh_numerics.py"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import __graft_entry__ as g  # noqa: E402

g.build()
import sls_sat_solving_with_deep_learning_b200 as eng  # noqa: E402
from conftest import golden_cnf, random_ksat  # noqa: E402
from oracle import gnn_oracle as go  # noqa: E402
from sls_sat_solving_with_deep_learning_b200 import gnn  # noqa: E402


def harsh_params(seed, weight_scale, ln_sigma):
    p = go.init_params(seed=seed, bias_scale=0.3)
    rng = np.random.default_rng(seed + 1)
    for name, mod in p.items():
        if "w" in mod:
            mod["w"] = (mod["w"] * weight_scale).astype(np.float32)
        else:
            mod["scale"] = (1.0 + ln_sigma * rng.standard_normal(mod["scale"].shape)).astype(np.float32)
            mod["offset"] = (ln_sigma * rng.standard_normal(mod["offset"].shape)).astype(np.float32)
    return p


def clauses_of(inst):
    row, lits = inst.csr()
    return [lits[int(row[c]):int(row[c + 1])].tolist() for c in range(inst.m)]


eng.set_device(0)
cases = []
for name in ("g4sat_hard_ps_00041", "ref_tests_medium", "g4sat_easy_ca_00092"):
    inst = eng.Instance.from_dimacs(golden_cnf(name))
    cases.append((name, inst, clauses_of(inst)))
cl = random_ksat(500, 2100, 3, seed=77)
cases.append(("random3sat_500", eng.Instance.from_clauses(500, cl), cl))
out = []
for wscale, ln_sigma in ((1.0, 0.0), (8.0, 0.0), (1.0, 0.5), (8.0, 0.5), (0.125, 0.5)):
    params = harsh_params(11, wscale, ln_sigma)
    net = gnn.InteractionNetworkLCG(params)
    for name, inst, clauses in cases:
        ref = go.model_probabilities(go.forward(params, go.lcg_graph(inst.n, clauses), dtype=np.float64), inst.n)
        ref32 = go.model_probabilities(go.forward(params, go.lcg_graph(inst.n, clauses), dtype=np.float32), inst.n)
        row = {"weights_x": wscale, "ln_sigma": ln_sigma, "graph": name, "max_k": inst.max_clause_len,
               "numpy_fp32": float(np.abs(ref32 - ref).max())}
        for prec in ("fp16x2", "bf16x2", "tf32x3"):
            os.environ["SLS_B200_GNN_PRECISION"] = prec
            p = net.probabilities([inst])[0]
            row[prec] = float(np.abs(p - ref).max())
        out.append(row)
        print(json.dumps(row), flush=True)
    net.close()
os.environ.pop("SLS_B200_GNN_PRECISION", None)
print("worst fp16x2", max(r["fp16x2"] for r in out), "worst bf16x2", max(r["bf16x2"] for r in out), "worst tf32x3", max(r["tf32x3"] for r in out))
