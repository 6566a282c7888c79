"""The CPU oracle against the reference's own known-answer tests and shipped fixtures (no GPU)."""
import glob
import itertools
import os
import pickle

import numpy as np
import pytest

from conftest import golden_assignments, golden_cnf
from oracle import sls_oracle as so


# ---- Philox4x32-10 known-answer vectors (Random123 kat_vectors) -------------------------------
@pytest.mark.parametrize(
    "ctr,key,out",
    [
        ([0, 0, 0, 0], [0, 0], [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]),
        ([0xFFFFFFFF] * 4, [0xFFFFFFFF] * 2, [0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD]),
        ([0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344], [0xA4093822, 0x299F31D0],
         [0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1]),
    ],
)
def test_philox_kat(ctr, key, out):
    assert so.philox4x32_10(ctr, key).tolist() == out


# ---- reference known-answer tests: python/tests/test_sat_instances.py --------------------------
def all_bitstrings(size):  # python/src/sat_instances.py:32-39
    return np.asarray(list(itertools.product([0, 1], repeat=size)))


def test_contradiction_one_violated():  # test_sat_instances.py:69-76
    inst = so.Instance.from_text("p cnf 1 2 \n 1 0 \n -1 0")
    assert (inst.n, inst.m) == (1, 2)
    assert [int(inst.eval_clauses(a).sum()) for a in all_bitstrings(1)] == [1, 1]


def test_sixteen_clause_formula_exactly_one_violated():  # test_sat_instances.py:86-111
    text = "p cnf 4 16 \n"
    lits = np.arange(4) + 1
    for bits in all_bitstrings(4):
        cl = ((bits - 0.5) * 2).astype(np.int32) * lits
        text += " ".join(str(int(x)) for x in cl) + " 0 \n "
    inst = so.Instance.from_text(text)
    assert (inst.n, inst.m) == (4, 16)
    assert all(int(inst.eval_clauses(a).sum()) == 1 for a in all_bitstrings(4))


def test_unequal_clause_lengths():  # test_sat_instances.py:121-135
    inst = so.Instance.from_text("p cnf 3 2 \n 1 2 0 \n 3 0")
    assert inst.eval_clauses([0, 0, 1]).tolist() == [1, 0]
    assert inst.eval_clauses([0, 1, 1]).tolist() == [0, 0]
    assert inst.eval_clauses([0, 1, 0]).tolist() == [0, 1]


def test_tautology_never_violated():  # instance of test_sat_instances.py:21-25 fed to lib.rs:13-18
    inst = so.Instance.from_text("p cnf 1 1 \n 1 -1 0")
    assert [int(inst.eval_clauses(a).sum()) for a in ([0], [1])] == [0, 0]


# ---- golden fixtures (subset of the shipped data, tests/golden/make_fixtures.py) --------------
def test_golden_energies(expected):
    for name, info in expected.items():
        inst = so.Instance.from_dimacs(golden_cnf(name))
        assert (inst.n, inst.m, inst.num_lits) == (info["n"], info["m"], info["num_lits"])
        rows = golden_assignments(name, inst.n)
        assert rows.shape[0] == info["rows"]
        assert [inst.energy(a) for a in rows] == info["energies"]
        assert info["energies"][0] == 0  # data_utils.py:130: the shipped solution satisfies the formula


REF = "/root/reference"


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference checkout only exists in the build container")
def test_every_shipped_solution_has_energy_zero():
    files = glob.glob(REF + "/Data/**/*_sol.pkl", recursive=True) + glob.glob(
        REF + "/python/tests/**/*_sol.pkl", recursive=True
    )
    assert len(files) > 3000
    for f in files:
        inst = so.Instance.from_dimacs(f.replace("_sol.pkl", ".cnf"))
        s = pickle.load(open(f, "rb"))
        if isinstance(s, dict):
            s = [x for _, x in sorted(s.items())]
        s = np.asarray(s)
        if (np.abs(s) >= 2).any() or (s < 0).any():
            s = (np.sign(s) + 1) // 2
        assert inst.energy(s) == 0, f
        assert inst.m_distinct == inst.m, f  # no duplicate clauses in the shipped data


# ---- DIMACS acceptance (varisat-like strictness) -----------------------------------------------
def test_parser_rules():
    ok = so.Instance.from_text("c comment\np cnf 3 2\n1 -2\n 3 0 -1 0\n")
    assert (ok.n, ok.m, ok.num_lits) == (3, 2, 4)
    assert so.Instance.from_text("1 2 0\n-3 0\n").n == 3  # no header: var_count = max variable
    assert so.Instance.from_text("p cnf 5 1\n1 0\n").n == 5  # header var count wins
    for bad in ("p cnf 2 2\n1 0\n", "p cnf 1 1\n2 0\n", "p cnf 2 1\n1 2\n", "p cnf 2 1\n1 x 0\n", "p dnf 1 1\n1 0\n"):
        with pytest.raises(so.OracleError):
            so.Instance.from_text(bad)
    with pytest.raises(so.OracleError):
        so.Instance.from_dimacs("/nonexistent/file.cnf")


def test_duplicate_clauses_are_one_set_element():  # SURVEY H1, lib.rs:176-181
    inst = so.Instance.from_text("p cnf 2 3\n1 2 0\n1 2 0\n-1 0\n")
    assert inst.m == 3 and inst.m_distinct == 2
    assert inst.energy([0, 0]) == 2  # Python semantics: every clause on its own
    assert inst.energy_distinct([0, 0]) == 1  # HashSet semantics
    r = inst.run_sls("schoening", [0, 0], [0.5, 0.5], 0, 1, True, seed=1)
    assert r.trajectories[0].tolist() == [1] and r.best_energy == 1


# ---- solver-level invariants -----------------------------------------------
@pytest.fixture(scope="module")
def medium():
    return so.Instance.from_dimacs(golden_cnf("ref_tests_medium"))


@pytest.mark.parametrize("algo", list(so.ALGOS))
@pytest.mark.parametrize("pfb", [0.0, 0.4])
@pytest.mark.parametrize("mapping", [True, False])
def test_walk_invariants(medium, algo, pfb, mapping):
    w = np.full(medium.n, 0.5)
    r = medium.run_sls(algo, w, w, 400, 12, True, mapping, pfb, seed=5)
    assert len(r.trajectories) == 12
    for run in range(12):
        t = r.trajectories[run]
        assert len(t) == r.steps[run] + 1 <= 401
        assert r.found[run] == (t[-1] == 0)
        assert r.best_energy_run[run] == t.min()
        assert (t[:-1] > 0).all()
    assert r.best_energy == r.best_energy_run.min()
    assert medium.energy_distinct(r.best_assignment) == r.best_energy
    assert r.total_steps == r.steps.sum()


def test_mapping_flag_only_changes_greedy(medium):
    w = np.full(medium.n, 0.5)
    for algo in so.ALGOS:
        a = medium.run_sls(algo, w, w, 300, 6, True, True, 0.0, seed=9)
        b = medium.run_sls(algo, w, w, 300, 6, True, False, 0.0, seed=9)
        assert all((x == y).all() for x, y in zip(a.trajectories, b.trajectories))


def test_exact_solution_as_init_takes_zero_steps(expected):
    name = "ref_tests_medium"
    inst = so.Instance.from_dimacs(golden_cnf(name))
    sol = golden_assignments(name, inst.n)[0].astype(float)
    r = inst.run_sls("moser", sol, sol, 100, 5, True, seed=3)
    assert r.found.all() and (r.steps == 0).all() and r.best_energy == 0
    assert (r.best_assignment == sol).all()


def test_no_run_below_m_returns_empty_best():
    # every clause violated at step 0 and nsteps = 0: best stays at m, best_assignment stays empty (lib.rs:203-219)
    inst = so.Instance.from_text("p cnf 2 2\n1 0\n2 0\n")
    r = inst.run_sls("moser", [0, 0], [0.5, 0.5], 0, 3, False, seed=1)
    assert r.best_energy == 2 and r.best_assignment is None and not r.found.any()
    assert r.as_reference_tuple()[0] == []


def test_threads_and_chain_offset_do_not_change_streams(medium):
    w = np.full(medium.n, 0.5)
    a = medium.run_sls("probsat", w, w, 200, 16, True, seed=4, nthreads=1)
    b = medium.run_sls("probsat", w, w, 200, 16, True, seed=4, nthreads=8)
    c = medium.run_sls("probsat", w, w, 200, 8, True, seed=4, chain_offset=8)
    assert all((x == y).all() for x, y in zip(a.trajectories, b.trajectories))
    assert all((x == y).all() for x, y in zip(a.trajectories[8:], c.trajectories))


def test_error_behaviour(medium):
    w = np.full(medium.n, 0.5)
    with pytest.raises(so.OracleError):  # lib.rs:125
        medium.run_sls("walksat", w, w, 10, 1)
    with pytest.raises(so.OracleError):  # lib.rs:209 index out of bounds
        medium.run_sls("moser", w[:-1], w, 10, 1)
    with pytest.raises(so.OracleError):  # gen_bool(p) outside [0,1]
        medium.run_sls("moser", w * 3, w, 10, 1)
    bad = w.copy()
    inst = so.Instance.from_text("p cnf 2 2\n1 2 0\n-1 -2 0\n")
    with pytest.raises(so.OracleError):  # WeightedIndex AllWeightsZero, lib.rs:71
        inst.run_sls("probsat", [0, 0], [0.0, 0.0], 10, 4, seed=2)
    del bad


# ---- distributional properties of the restated rand 0.8.5 pieces -------------------------------
def test_init_is_bernoulli_of_weights():
    n = 64
    inst = so.Instance.from_clauses(n, [(1, 2, 3)])
    w = np.linspace(0, 1, n)
    acc = np.zeros(n)
    R = 4000
    for chain in range(R):
        acc += inst.init_assignment(w, 123, chain)
    assert acc[0] == 0 and acc[-1] == R  # p = 0 never, p = 1 always
    z = (acc / R - w)[1:-1] / np.sqrt(w[1:-1] * (1 - w[1:-1]) / R)
    assert np.abs(z).max() < 4.5


def test_violated_clause_pick_is_uniform():
    # three disjoint violated clauses; schoening flips one variable of the picked clause, so the
    # first step's energy drop identifies nothing -- use the flipped variable's clause instead.
    inst = so.Instance.from_text("p cnf 9 3\n1 2 3 0\n4 5 6 0\n7 8 9 0\n")
    counts = np.zeros(3)
    R = 6000
    zeros = np.zeros(9)
    for chain in range(R):
        r = inst.run_sls("schoening", zeros, zeros + 0.5, 1, 1, False, seed=77, chain_offset=chain)
        counts[int(np.flatnonzero(r.best_assignment)[0]) // 3] += 1
    assert np.abs(counts / R - 1 / 3).max() < 0.03


def test_moser_resamples_each_variable_from_its_weight():
    inst = so.Instance.from_text("p cnf 3 1\n1 2 3 0\n")
    w = np.array([0.2, 0.5, 0.9])
    acc = np.zeros(3)
    R = 6000
    for chain in range(R):
        r = inst.run_sls("moser", np.zeros(3), w, 1, 1, False, seed=5, chain_offset=chain)
        a = r.best_assignment if r.found[0] else np.zeros(3)
        acc += a
    # one MT step from all-false: variable j ends up true with probability w_j (lib.rs:29-33)
    assert np.abs(acc / R - w).max() < 0.03


def test_probsat_picks_proportionally_to_flip_probability():
    inst = so.Instance.from_text("p cnf 3 1\n1 2 3 0\n")
    w = np.array([0.1, 0.3, 0.6])
    acc = np.zeros(3)
    R = 6000
    for chain in range(R):
        r = inst.run_sls("probsat", np.zeros(3), w, 1, 1, False, seed=6, chain_offset=chain)
        acc += r.best_assignment
    assert np.abs(acc / R - w / w.sum()).max() < 0.03


# ---- the C oracle against an independent, naive model of lib.rs (oracle/sls_model.py) ----------------
# Different data structures (hash set of clause contents, per-variable clause lists), different generator
# (Mersenne Twister), different element order: only distributions can agree.  Two-sample tests at a fixed
# false-alarm level; the chains of both sides are independent samples of what should be one process.
def _ks_two_sample(a, b):
    a, b = np.sort(np.asarray(a, float)), np.sort(np.asarray(b, float))
    grid = np.concatenate([a, b])
    d = np.abs(np.searchsorted(a, grid, side="right") / a.size - np.searchsorted(b, grid, side="right") / b.size).max()
    return d, np.sqrt(-0.5 * np.log(0.001 / 2) * (a.size + b.size) / (a.size * b.size))  # critical value at alpha = 0.001


@pytest.mark.parametrize("algo,pfb,mapping", [("moser", 0.0, True), ("probsat", 0.0, True), ("moser_single", 0.0, True),
                                              ("schoening", 0.0, False), ("probsat", 0.4, True), ("moser", 0.3, False)])
def test_c_oracle_and_naive_model_agree_in_distribution(algo, pfb, mapping):
    from conftest import random_ksat
    from oracle import sls_model

    n, m, S, R = 40, 150, 400, 300  # alpha = 3.75: a few hundred steps, some chains hit the budget
    clauses = random_ksat(n, m, 3, seed=31)
    inst = so.Instance.from_clauses(n, clauses)
    rng = np.random.default_rng(3)
    wi, wr = rng.uniform(0.2, 0.8, n), rng.uniform(0.3, 0.7, n)  # a non-uniform oracle: weights matter
    c = inst.run_sls(algo, wi, wr, S, R, True, mapping, pfb, seed=99)
    p = sls_model.run_sls(algo, n, clauses, wi.tolist(), wr.tolist(), S, R, True, mapping, pfb, seed=7)
    # solve fraction: two-proportion z test
    f1, f2 = float(np.mean(c.found)), float(np.mean(p.found))
    pool = (f1 + f2) / 2
    z = abs(f1 - f2) / max(np.sqrt(2 * pool * (1 - pool) / R), 1e-9)
    assert z < 4.0, (f1, f2)
    # step counts (capped at the budget), initial energy and energy after 20 steps: Kolmogorov-Smirnov
    for name, a, b in [("steps", c.steps, p.steps),
                       ("initial energy", [t[0] for t in c.trajectories], [t[0] for t in p.trajectories]),
                       ("energy after 20 steps", [t[min(20, len(t) - 1)] for t in c.trajectories],
                        [t[min(20, len(t) - 1)] for t in p.trajectories])]:
        d, crit = _ks_two_sample(a, b)
        assert d < crit, (name, d, crit)
    # and both honour the invariants of lib.rs:292-300
    for r in (c, p):
        assert all((t[-1] == 0) == bool(f) for t, f in zip(r.trajectories, r.found))
        assert all(len(t) == s + 1 for t, s in zip(r.trajectories, r.steps))


# ---- the two CPU models of the GNN forward ---------------------------------------------------------------------------
@pytest.mark.parametrize("layers,steps,bias", [((200, 200), 5, 0.0), ((200, 200), 5, 0.3), ((32, 32), 2, 0.2), ((64, 48), 3, 0.2)])
def test_numpy_and_torch_gnn_models_agree(layers, steps, bias):
    # oracle/gnn_oracle.py (NumPy, the checker of the CUDA kernels) and oracle/gnn_torch.py (torch fp32: graph from flat
    # literal arrays, names replayed from the construction order, index_add_ aggregation, F.layer_norm) share no code;
    # they must agree to fp32 rounding on mixed clause lengths and on a 3-SAT graph
    from oracle import gnn_oracle as go
    from oracle import gnn_torch as gt

    from conftest import random_ksat

    cases = [(60, random_ksat(60, 250, 3, seed=5)), (7, [(1, -2), (3,), (-1, -2, 3, 7), (4, 5, -6), (-7, 2)])]
    for n, clauses in cases:
        params = go.init_params(seed=42, mlp_layers=layers, num_steps=steps, bias_scale=bias)
        graph = go.lcg_graph(n, clauses)
        ref64 = go.forward(params, graph, dtype=np.float64, num_steps=steps)
        t = gt.forward(params, *gt.graph_from_clauses(n, clauses), mlp_layers=layers, num_message_passing_steps=steps).numpy()
        assert t.shape == ref64.shape
        assert np.abs(t - ref64).max() < 5e-5
        import torch

        p_t = gt.model_probabilities(torch.from_numpy(t), n)
        assert np.abs(p_t - go.model_probabilities(ref64, n)).max() < 1e-5
        # identical graph arrays from the two builders
        nodes, edges, snd, rcv = gt.graph_from_clauses(n, clauses)
        assert np.array_equal(nodes.numpy(), graph["nodes"]) and np.array_equal(edges.numpy(), graph["edges"])
        assert np.array_equal(snd.numpy(), graph["senders"]) and np.array_equal(rcv.numpy(), graph["receivers"])
