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
        # drop duplicate clause contents, keeping first occurrences in order
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
