"""LCG.get_constraint_graph (python/src/sat_representations.py:668-715) -- host-side, no GPU: the C++ builder behind
``sls_constraint_graph`` against the reference's own formulation (scipy sparse C^T C, restated here) as edge SETS, on
shipped instances and edge cases.  The reference's edge ORDER is whatever scipy's sparse product emits; the LLL loss only
segment-sums over the edges."""
import numpy as np
import pytest
import scipy.sparse

from conftest import golden_cnf, random_ksat
from oracle import gnn_oracle as go
from sls_sat_solving_with_deep_learning_b200 import gnn
import sls_sat_solving_with_deep_learning_b200 as eng


def reference_edges(n, m, senders, receivers):
    """The arithmetic of sat_representations.py:683-703: incidence matrix, C^T C, off-diagonal pattern, + 2n."""
    senders, receivers = np.asarray(senders), np.asarray(receivers)
    row = np.floor(senders[:-n] / 2).astype(np.int64)
    col = (receivers[:-n] - 2 * n).astype(np.int64)
    inc = scipy.sparse.csr_matrix((np.ones(len(row)), (row, col)), (n, m))
    adj = (inc.transpose() @ inc).tocoo()
    keep = adj.row != adj.col
    return set(zip((adj.row[keep] + 2 * n).tolist(), (adj.col[keep] + 2 * n).tolist()))


def clauses_of(inst):
    row, lits = inst.csr()
    return [lits[int(row[c]):int(row[c + 1])].tolist() for c in range(inst.m)]


@pytest.mark.parametrize("name", ["r3sat_test_random_KCNF3_100_210_6416893", "r3sat_test_random_KCNF3_300_660_144280",
                                  "g4sat_easy_ca_00092", "g4sat_hard_ps_00041", "ref_tests_medium"])
def test_constraint_graph_matches_scipy_formulation(name):
    inst = eng.Instance.from_dimacs(golden_cnf(name))
    graph = go.lcg_graph(inst.n, clauses_of(inst))
    m = graph["n_node"] - 2 * inst.n
    cg = gnn.get_constraint_graph(inst.n, m, graph["senders"], graph["receivers"])
    got = list(zip(cg["senders"].tolist(), cg["receivers"].tolist()))
    assert cg["n_node"] == m and cg["n_edge"] == len(got) == len(set(got))
    assert set(got) == reference_edges(inst.n, m, graph["senders"], graph["receivers"])
    assert got == sorted(got)  # clause by clause, neighbours in increasing index
    assert cg["edges"].shape == (len(got),) and cg["nodes"].shape == (m,)


def test_constraint_graph_small_cases():
    # (x1 v x2), (~x2 v x3), (x4): clauses 0 and 1 share x2, clause 2 is isolated
    g = go.lcg_graph(4, [(1, 2), (-2, 3), (4,)])
    cg = gnn.get_constraint_graph(4, 3, g["senders"], g["receivers"])
    assert list(zip(cg["senders"].tolist(), cg["receivers"].tolist())) == [(8, 9), (9, 8)]
    # a polarity does not matter, multiplicity of shared variables does not duplicate edges
    g = go.lcg_graph(3, [(1, 2, 3), (-1, -2, -3)])
    cg = gnn.get_constraint_graph(3, 2, g["senders"], g["receivers"])
    assert cg["n_edge"] == 2
    # no clauses at all
    cg = gnn.get_constraint_graph(2, 0, np.array([1, 3]), np.array([0, 2]))
    assert cg["n_edge"] == 0 and cg["n_node"] == 0
    with pytest.raises(eng.PanicException):  # clause index beyond n_clauses
        gnn.get_constraint_graph(2, 1, np.array([1, 2, 1, 3]), np.array([4, 5, 0, 2]))


def test_constraint_graph_config4_shape_is_fast():
    n, m = 500, 2100
    g = go.lcg_graph(n, random_ksat(n, m, 3, seed=4))
    cg = gnn.get_constraint_graph(n, m, g["senders"], g["receivers"])
    assert set(zip(cg["senders"].tolist(), cg["receivers"].tolist())) == reference_edges(n, m, g["senders"], g["receivers"])
