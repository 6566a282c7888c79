import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


@pytest.fixture(scope="session", autouse=True)
def _built():
    """Build the CUDA library (nvcc cross-compiles without a GPU) and the CPU oracle once."""
    import __graft_entry__ as g

    g.build()


@pytest.fixture(scope="session")
def expected():
    return json.load(open(os.path.join(GOLDEN, "expected.json")))


def golden_cnf(name):
    return os.path.join(GOLDEN, "instances", name + ".cnf")


def golden_assignments(name, n):
    packed = np.load(os.path.join(GOLDEN, "instances", name + "_assignments.npy"))
    return np.unpackbits(packed, axis=1, bitorder="little")[:, :n]


def random_ksat(n, m, k, seed, sort_vars=True):
    """Uniform random k-SAT as cnfgen emits it: k distinct variables per clause (ascending), fair
    signs, no duplicate clauses."""
    rng = np.random.default_rng(seed)
    seen = set()
    clauses = []
    while len(clauses) < m:
        vs = rng.choice(n, k, replace=False) + 1
        if sort_vars:
            vs = np.sort(vs)
        cl = tuple(int(v if rng.random() < 0.5 else -v) for v in vs)
        if cl not in seen:
            seen.add(cl)
            clauses.append(cl)
    return clauses
