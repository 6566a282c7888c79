"""Synthetic random k-SAT instances of the shapes BASELINE.json names.

Mirrors the distribution of the reference's generator (python/src/generate_random_instances.py:28-92,
cnfgen RandomKCNF): every clause has k distinct variables in ascending order with independent fair
signs, and no clause appears twice.  Host-side NumPy only; produces the CSR the C ABI takes.
"""
from __future__ import annotations

import numpy as np


def random_ksat_csr(n: int, m: int, k: int = 3, seed: int = 20260102):
    """Returns (row_ptr uint64[m+1], signed_lits int32[m*k])."""
    rng = np.random.Generator(np.random.PCG64(seed))
    rows = np.zeros((0, k), dtype=np.int64)
    while rows.shape[0] < m:
        need = m - rows.shape[0]
        cand = rng.integers(0, n, size=(int(need * 1.05) + 16, k))
        cand.sort(axis=1)
        distinct = (np.diff(cand, axis=1) != 0).all(axis=1)
        cand = cand[distinct] + 1
        signs = rng.integers(0, 2, size=cand.shape) * 2 - 1
        rows = np.concatenate([rows, cand * signs])
        # drop duplicate clause contents, keeping first occurrences in order.  One integer key per clause (its literal
        # codes 2 * var + neg as digits of a base-(2n + 2) number) makes this a 1-D unique: 4x faster than a row-wise
        # unique at m = 4 * 10^6, same result
        base = 2 * n + 2
        if base ** k < 2 ** 63:
            code = 2 * np.abs(rows) + (rows < 0)
            key = code[:, 0]
            for j in range(1, k):
                key = key * base + code[:, j]
            _, first = np.unique(key, return_index=True)
        else:
            _, first = np.unique(rows, axis=0, return_index=True)
        rows = rows[np.sort(first)]
    rows = rows[:m]
    row_ptr = np.arange(m + 1, dtype=np.uint64) * k
    return row_ptr, rows.astype(np.int32).ravel()


def write_dimacs(path: str, n: int, row_ptr, signed_lits) -> None:
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    lits = np.asarray(signed_lits)
    with open(path, "w") as f:
        f.write(f"p cnf {n} {len(row_ptr) - 1}\n")
        lines = []
        for c in range(len(row_ptr) - 1):
            lines.append(" ".join(map(str, lits[row_ptr[c]:row_ptr[c + 1]])) + " 0\n")
        f.write("".join(lines))


def random_oracle_params(seed: int = 42, mlp_layers=(200, 200), num_steps: int = 5, embed: int = 32):
    """Random-init parameters of the reference's interaction network in haiku's layout ({module: {"w", "b"}} /
    {"scale", "offset"}}, module names in construction order): Linear weights ~ truncated normal (+-2 sigma) with
    sigma = 1 / sqrt(fan_in), zero biases, LayerNorm scale 1 / offset 0 -- haiku's defaults (model.py:13-16, :37-60, :144).
    Stand-in for the pre-trained Data/models files, which are not shipped with the reference checkout."""
    from .gnn import haiku_names

    rng = np.random.default_rng(seed)
    h1, h2 = mlp_layers

    def lin(fan_in, fan_out):
        w = np.clip(rng.standard_normal((fan_in, fan_out)), -2.0, 2.0) / np.sqrt(fan_in)
        return {"w": w.astype(np.float32), "b": np.zeros(fan_out, dtype=np.float32)}

    def ln(dim):
        return {"scale": np.ones(dim, dtype=np.float32), "offset": np.zeros(dim, dtype=np.float32)}

    names = haiku_names(num_steps)
    params = {names["edge_embed"]: lin(2, embed), names["node_embed"]: lin(2, embed)}
    e_dim = h_dim = embed
    for t in range(num_steps):
        st = names["steps"][t]
        params[st["edge"]["ln"]] = ln(h2)
        params[st["edge"]["l1"]] = lin(e_dim + 2 * h_dim, h1)
        params[st["edge"]["l2"]] = lin(h1, h2)
        params[st["node"]["ln"]] = ln(h2)
        params[st["node"]["l1"]] = lin(h_dim + 2 * h2, h1)
        params[st["node"]["l2"]] = lin(h1, h2)
        e_dim = h_dim = h2
    params[names["readout"]] = lin(h_dim, 1)
    return params

