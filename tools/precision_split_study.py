"""CPU study (NumPy emulation, no GPU): how far do the oracle's probabilities drift from the fp64 restatement when the two
dense layers of every update are computed from split low-precision operands with fp32 accumulation?
Modes: tf32 x1 (single pass), 3xTF32 (what the CUDA kernel does), and two-/three-piece bf16 splits with the product
terms a kind::f16 tcgen05 kernel would issue.  usage: python tools/precision_split_study.py [n_instances]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import gnn_oracle as go  # noqa: E402
from sls_sat_solving_with_deep_learning_b200 import synth  # noqa: E402


def tf32(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)


def bf16(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) & np.uint32(0xFFFF0000)).view(np.float32)


def bf16_half_up(x):  # a cheaper split that was tried first: add half an ulp of the magnitude, truncate
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x8000)) & np.uint32(0xFFFF0000)).view(np.float32)


def bf16_trunc(x):
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32)
    return (u & np.uint32(0xFFFF0000)).view(np.float32)


def fp16(x):  # round to nearest even, subnormals and overflow as the hardware conversion
    with np.errstate(over="ignore"):
        return np.asarray(x, dtype=np.float32).astype(np.float16).astype(np.float32)


def split_kernel(x, rnd, pieces):  # first piece rounded half up, residual truncated
    a0 = bf16_half_up(x)
    return [a0, bf16_trunc((np.asarray(x, dtype=np.float32) - a0).astype(np.float32))]


def split(x, rnd, pieces):
    out, r = [], np.asarray(x, dtype=np.float32)
    for _ in range(pieces):
        p = rnd(r)
        out.append(p)
        r = (r - p).astype(np.float32)
    return out


MODES = {
    "fp32": None,
    "tf32 x1": (tf32, 1, [(0, 0)]),
    "3xTF32": (tf32, 2, [(0, 0), (1, 0), (0, 1)]),
    "bf16 x1": (bf16, 1, [(0, 0)]),
    "bf16 2 pieces, 3 terms": (bf16, 2, [(0, 0), (0, 1), (1, 0)]),
    "bf16 2 pieces, 3 terms, A: half-up + truncated residual": ("kernel", 2, [(0, 0), (0, 1), (1, 0)]),
    "bf16 2 pieces, 4 terms": (bf16, 2, [(0, 0), (0, 1), (1, 0), (1, 1)]),
    "bf16 3 pieces, 6 terms": (bf16, 3, [(0, 0), (0, 1), (1, 0), (1, 1), (0, 2), (2, 0)]),
    "fp16 x1": (fp16, 1, [(0, 0)]),
    "fp16 A 2 pieces x B 1 piece, 2 terms": (fp16, 2, [(0, 0), (1, 0)]),
    "fp16 2 pieces, 3 terms": (fp16, 2, [(0, 0), (0, 1), (1, 0)]),
}


def forward_mode(params, graph, mode):
    names = go.param_names(go.NUM_STEPS)
    P = {k: {kk: np.asarray(vv, dtype=np.float32) for kk, vv in v.items()} for k, v in params.items()}
    relu = lambda x: np.maximum(x, 0)

    def dense(name, x):
        w, b = P[name]["w"], P[name]["b"]
        if mode is None:
            return x @ w + b
        rnd, pieces, terms = mode
        if rnd == "kernel":
            xs, ws = split_kernel(x, None, 2), split(w, bf16, 2)
        else:
            xs, ws = split(x, rnd, pieces), split(w, rnd, pieces)
        acc = np.zeros((x.shape[0], w.shape[1]), dtype=np.float32)
        for i, j in terms:
            acc += xs[i] @ ws[j]
        return acc + b

    def ln(name, x):
        mean = x.mean(axis=-1, keepdims=True)
        var = ((x - mean) ** 2).mean(axis=-1, keepdims=True)
        return (x - mean) / np.sqrt(var + np.float32(go.LN_EPS)) * P[name]["scale"] + P[name]["offset"]

    def update(s, f):
        return ln(s["ln"], relu(dense(s["l2"], relu(dense(s["l1"], f)))))

    snd, rcv, N = graph["senders"], graph["receivers"], graph["n_node"]
    e = graph["edges"].astype(np.float32) @ P[names["edge_embed"]]["w"] + P[names["edge_embed"]]["b"]
    h = graph["nodes"].astype(np.float32) @ P[names["node_embed"]]["w"] + P[names["node_embed"]]["b"]
    for t in range(go.NUM_STEPS):
        s = names["steps"][t]
        e = update(s["edge"], np.concatenate([e, h[snd], h[rcv]], axis=1))
        h = update(s["node"], np.concatenate([h, go._segment_sum(e, snd, N), go._segment_sum(e, rcv, N)], axis=1))
    return h @ P[names["readout"]]["w"] + P[names["readout"]]["b"]


def harsh(params, weights_x, ln_sigma, seed=5):
    """the parameter families of tools/gnn_harsh_numerics.py: Linear weights scaled, LayerNorm scale ~ N(1, s), offset ~ N(0, s)"""
    rng = np.random.default_rng(seed)
    out = {}
    for k, v in params.items():
        v = {kk: np.array(vv, dtype=np.float32) for kk, vv in v.items()}
        if "w" in v:
            v["w"] = v["w"] * np.float32(weights_x)
        if "scale" in v and ln_sigma > 0:
            v["scale"] = (1.0 + ln_sigma * rng.standard_normal(v["scale"].shape)).astype(np.float32)
            v["offset"] = (ln_sigma * rng.standard_normal(v["offset"].shape)).astype(np.float32)
        out[k] = v
    return out


if __name__ == "__main__":
    n_inst = int(sys.argv[1]) if len(sys.argv) > 1 else 4
    n, m = 500, 2100
    params = go.init_params(seed=42)
    if len(sys.argv) > 3:
        params = harsh(params, float(sys.argv[2]), float(sys.argv[3]))
    worst = {k: 0.0 for k in MODES}
    for i in range(n_inst):
        row, lits = synth.random_ksat_csr(n, m, 3, 20260104 + i)
        clauses = [lits[int(row[c]):int(row[c + 1])].tolist() for c in range(m)]  # signed DIMACS literals
        g = go.lcg_graph(n, clauses)
        ref = go.model_probabilities(go.forward(params, g, dtype=np.float64), n)
        for k, mode in MODES.items():
            p = go.model_probabilities(forward_mode(params, g, mode), n)
            worst[k] = max(worst[k], float(np.abs(p - ref).max()))
    for k, v in worst.items():
        print(f"{k:56s} max |dp| vs fp64 = {v:.2e}")
