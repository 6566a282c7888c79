"""CUDA path vs the CPU oracle on identical inputs and identical Philox streams (needs a B200).

Bit-exact bar: violated-clause sets, energies, per-step trajectories, step counts, solved flags
and best assignments must be identical.  All calls go through the C ABI (libsls_b200.so).
"""
import os

import numpy as np
import pytest

from conftest import golden_assignments, golden_cnf, random_ksat
from oracle import sls_oracle as so
import sls_sat_solving_with_deep_learning_b200 as eng

pytestmark = pytest.mark.gpu

ALGOS = ["moser", "moser_single", "schoening", "probsat"]


@pytest.fixture(scope="module", autouse=True)
def _device():
    assert eng.device_count() > 0, "GPU tests need a CUDA device: the engine has no CPU fallback"
    eng.set_device(0)


def both(name):
    return eng.Instance.from_dimacs(golden_cnf(name)), so.Instance.from_dimacs(golden_cnf(name))


def assert_same_run(g, o):
    assert (g.steps == o.steps).all()
    assert (g.found == o.found).all()
    assert (g.best_energy_run == o.best_energy_run).all()
    assert (g.flips_run == o.flips_run).all()
    assert g.best_energy == o.best_energy
    assert (g.best_assignment is None) == (o.best_assignment is None)
    if g.best_assignment is not None:
        assert (g.best_assignment == o.best_assignment).all()
    assert len(g.trajectories) == len(o.trajectories)
    for a, b in zip(g.trajectories, o.trajectories):
        assert a.shape == b.shape and (a == b).all()


# ---- clause evaluation ------------------------------------------------------------------------
def test_eval_reference_known_answers():  # python/tests/test_sat_instances.py:69-135
    gi = eng.Instance.from_text("p cnf 1 2 \n 1 0 \n -1 0")
    assert [int(eng.get_violated_constraints(gi, [a]).sum()) for a in (0, 1)] == [1, 1]
    gi = eng.Instance.from_text("p cnf 3 2 \n 1 2 0 \n 3 0")
    assert eng.get_violated_constraints(gi, [0, 0, 1]).tolist() == [1, 0]
    assert eng.get_violated_constraints(gi, [0, 1, 1]).tolist() == [0, 0]
    assert eng.get_violated_constraints(gi, [0, 1, 0]).tolist() == [0, 1]
    import itertools

    text = "p cnf 4 16 \n"
    for bits in itertools.product([0, 1], repeat=4):
        text += " ".join(str((i + 1) * (1 if b else -1)) for i, b in enumerate(bits)) + " 0 \n"
    gi = eng.Instance.from_text(text)
    viol, energy = eng.eval_assignments(gi, np.asarray(list(itertools.product([0, 1], repeat=4))))
    assert (energy == 1).all() and (viol.sum(axis=1) == 1).all()


def test_eval_golden_fixtures(expected):
    for name, info in expected.items():
        gi, oi = both(name)
        rows = golden_assignments(name, gi.n)
        viol, energy = eng.eval_assignments(gi, rows)
        assert energy.tolist() == info["energies"], name
        for b in range(rows.shape[0]):
            assert (viol[b] == oi.eval_clauses(rows[b])).all(), (name, b)


def test_eval_ragged_batches_and_shapes():
    rng = np.random.default_rng(0)
    for n, m, k in [(1, 1, 1), (33, 65, 2), (100, 31, 3), (257, 1025, 3), (64, 2049, 5)]:
        clauses = random_ksat(n, m, min(k, n), seed=n + m)
        gi = eng.Instance.from_clauses(n, clauses)
        oi = so.Instance.from_clauses(n, clauses)
        for B in (1, 31, 32, 33, 100):
            a = rng.integers(0, 2, size=(B, n))
            viol, energy = eng.eval_assignments(gi, a)
            for b in range(B):
                assert (viol[b] == oi.eval_clauses(a[b])).all()
                assert energy[b] == oi.energy(a[b])
    gi = eng.Instance.from_text("p cnf 3 2\n0\n1 2 0\n")  # an empty clause is always violated (lib.rs:13-18)
    viol, energy = eng.eval_assignments(gi, [[1, 1, 1]])
    assert viol.tolist() == [[1, 0]] and energy.tolist() == [1]


def test_eval_batches_spanning_slices_and_candidate_energies(expected):
    # batches around the 256-assignment slice width, rows compared one by one with the oracle
    rng = np.random.default_rng(5)
    n, m = 300, 1300
    clauses = random_ksat(n, m, 3, seed=77)
    gi = eng.Instance.from_clauses(n, clauses)
    oi = so.Instance.from_clauses(n, clauses)
    for B in (255, 256, 257, 600):
        a = rng.integers(0, 2, size=(B, n))
        viol, energy = eng.eval_assignments(gi, a)
        assert eng.eval_last_kernel_ms() > 0
        for b in list(range(0, B, 37)) + [B - 1]:
            assert (viol[b] == oi.eval_clauses(a[b])).all()
        assert (viol.sum(axis=1) == energy).all()
        _, e2 = eng.eval_assignments(gi, a, want_bits=False)  # energy-only mode
        assert (e2 == energy).all()
    # data_utils.py:123-130: candidate energies, the first candidate of every shipped file is the solution
    for name, info in expected.items():
        gi, _ = both(name)
        rows = golden_assignments(name, gi.n)
        e = eng.candidate_energies(gi, rows)
        row, _ = gi.csr()
        m_nonempty = int(np.count_nonzero(np.diff(row.astype(np.int64)) > 0))
        want = (np.asarray(info["energies"]) - (gi.m - m_nonempty)) / m_nonempty
        assert np.allclose(e, want.astype(np.float32))
        assert e[0] == 0


def test_eval_full_size_c2_properties():
    # BASELINE config 2 shape (n=10k, m=42k): size-independent properties
    n, m = 10_000, 42_000
    clauses = random_ksat(n, m, 3, seed=20260102)
    gi = eng.Instance.from_clauses(n, clauses)
    oi = so.Instance.from_clauses(n, clauses)
    rng = np.random.default_rng(1)
    a = rng.integers(0, 2, size=(96, n))
    viol, energy = eng.eval_assignments(gi, a)
    assert (viol.sum(axis=1) == energy).all()
    for b in (0, 47, 95):
        assert (viol[b] == oi.eval_clauses(a[b])).all()
    # complementing an assignment and every literal sign gives the same violated set
    neg = eng.Instance.from_clauses(n, [tuple(-l for l in c) for c in clauses])
    viol2, energy2 = eng.eval_assignments(neg, 1 - a)
    assert (viol2 == viol).all() and (energy2 == energy).all()


# ---- the walk: trajectory-exact against the oracle ---------------------------------------------
@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("pfb", [0.0, 0.35])
@pytest.mark.parametrize("mapping", [True, False])
def test_walk_k3_kernel_matches_oracle(algo, pfb, mapping):
    name = "r3sat_test_random_KCNF3_200_869_9397183"
    gi, oi = both(name)
    rng = np.random.default_rng(3)
    w_init = rng.uniform(0.05, 0.95, gi.n)
    w_res = rng.uniform(0.05, 0.95, gi.n)
    g = eng.run_sls(gi, algo, w_init, w_res, 600, 40, True, mapping, pfb, seed=21)
    o = oi.run_sls(algo, w_init, w_res, 600, 40, True, mapping, pfb, seed=21)
    assert_same_run(g, o)


def _mixed_short_clauses(n, m, seed):
    """Distinct clauses of 1..3 literals over distinct variables: the fast kernel's absent-literal paths."""
    rng = np.random.default_rng(seed)
    seen, out = set(), []
    while len(out) < m:
        k = int(rng.integers(1, 4))
        vs = sorted(rng.choice(n, size=k, replace=False) + 1)
        c = tuple(int(v) * (1 if rng.integers(0, 2) else -1) for v in vs)
        if c not in seen:
            seen.add(c)
            out.append(c)
    return out


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("pfb", [0.0, 0.35])
def test_walk_k3_kernel_short_clauses_match_oracle(algo, pfb):
    # clauses with 1 and 2 literals next to 3-literal ones (absent literals = the dummy variable), both on an
    # instance a warp covers in one round and on one with several counter groups
    for n, m, seed in [(2, 2, 0), (40, 150, 1), (700, 2600, 2)]:
        clauses = [(1, 2), (-1, -2)] if n == 2 else _mixed_short_clauses(n, m, seed)
        gi = eng.Instance.from_clauses(n, clauses)
        oi = so.Instance.from_clauses(n, clauses)
        rng = np.random.default_rng(seed)
        w_init = rng.uniform(0.05, 0.95, n)
        w_res = rng.uniform(0.05, 0.95, n)
        s = eng.ResidentSolver(gi, algo, 300, 24, True, True, pfb, seed=5)
        assert s.variant.startswith("walk_k3s"), s.variant
        s.close()
        g = eng.run_sls(gi, algo, w_init, w_res, 300, 24, True, True, pfb, seed=5)
        o = oi.run_sls(algo, w_init, w_res, 300, 24, True, True, pfb, seed=5)
        assert_same_run(g, o)


@pytest.mark.parametrize("algo", ALGOS)
def test_walk_k3_kernel_boundary_weights_match_oracle(algo):
    # resample weights of exactly 0 and 1 on some variables: zero flip probabilities inside a clause
    # (WeightedIndex with zero entries, Bernoulli marks that never / always fire); probSAT fails as a whole when a
    # violated clause has only zero weights -- both sides must then fail
    name = "r3sat_test_random_KCNF3_100_210_6416893"
    gi, oi = both(name)
    rng = np.random.default_rng(9)
    w_res = rng.uniform(0.05, 0.95, gi.n)
    w_res[rng.choice(gi.n, 12, replace=False)] = 0.0
    w_res[rng.choice(gi.n, 12, replace=False)] = 1.0
    w_init = rng.uniform(0.05, 0.95, gi.n)
    try:
        o = oi.run_sls(algo, w_init, w_res, 400, 24, True, True, 0.0, seed=8)
    except Exception:
        with pytest.raises(eng.PanicException):
            eng.run_sls(gi, algo, w_init, w_res, 400, 24, True, True, 0.0, seed=8)
        return
    g = eng.run_sls(gi, algo, w_init, w_res, 400, 24, True, True, 0.0, seed=8)
    assert_same_run(g, o)


@pytest.mark.parametrize("algo", ALGOS)
@pytest.mark.parametrize("pfb", [0.0, 0.35])
@pytest.mark.parametrize("mapping", [True, False])
@pytest.mark.parametrize("name", ["ref_tests_medium", "g4sat_hard_ps_00041"])
def test_walk_generic_kernel_matches_oracle(algo, pfb, mapping, name):
    gi, oi = both(name)  # mixed clause lengths (k up to 6 / long pseudo-industrial clauses)
    rng = np.random.default_rng(4)
    w_init = rng.uniform(0.05, 0.95, gi.n)
    w_res = rng.uniform(0.05, 0.95, gi.n)
    g = eng.run_sls(gi, algo, w_init, w_res, 400, 24, True, mapping, pfb, seed=22)
    o = oi.run_sls(algo, w_init, w_res, 400, 24, True, mapping, pfb, seed=22)
    assert_same_run(g, o)


@pytest.mark.parametrize("force", [{"SLS_B200_FORCE_GENERIC": "1"}, {"SLS_B200_FORCE_STATE": "hbm"},
                                   {"SLS_B200_FORCE_GENERIC": "1", "SLS_B200_FORCE_STATE": "hbm"}])
@pytest.mark.parametrize("algo", ALGOS)
def test_walk_all_kernel_variants_agree(force, algo, monkeypatch):
    name = "r3sat_test_random_KCNF3_300_990_9917523"
    gi, oi = both(name)
    w = np.full(gi.n, 0.5)
    for k, v in force.items():
        monkeypatch.setenv(k, v)
    g = eng.run_sls(gi, algo, w, w, 500, 33, True, True, 0.2, seed=5)
    o = oi.run_sls(algo, w, w, 500, 33, True, True, 0.2, seed=5)
    assert_same_run(g, o)


def test_walk_all_golden_instances(expected):
    for name in expected:
        gi, oi = both(name)
        w = np.full(gi.n, 0.5)
        for algo in ("moser", "probsat"):
            g = eng.run_sls(gi, algo, w, w, 300, 10, True, True, 0.0, seed=8)
            o = oi.run_sls(algo, w, w, 300, 10, True, True, 0.0, seed=8)
            assert_same_run(g, o)


def test_walk_duplicate_clauses_and_repeated_variables():
    # duplicate contents are ONE violated-set element (lib.rs:176-181); a variable twice in a clause
    # appears twice in var_to_clauses (lib.rs:190-195)
    text = "p cnf 6 9\n1 2 3 0\n1 2 3 0\n-1 -2 0\n4 4 5 0\n-4 -5 6 0\n1 2 3 0\n-6 0\n3 -3 1 0\n-1 -1 0\n"
    gi, oi = eng.Instance.from_text(text), so.Instance.from_text(text)
    w = np.full(6, 0.5)
    for algo in ALGOS:
        for pfb in (0.0, 0.5):
            for mapping in (True, False):
                g = eng.run_sls(gi, algo, w * 0, w, 200, 50, True, mapping, pfb, seed=13)
                o = oi.run_sls(algo, w * 0, w, 200, 50, True, mapping, pfb, seed=13)
                assert_same_run(g, o)


def test_walk_edge_cases():
    gi, oi = both("ref_tests_medium")
    w = np.full(gi.n, 0.5)
    # nruns = 0 / nsteps = 0
    g = eng.run_sls(gi, "moser", w, w, 10, 0, True)
    assert g.best_energy == gi.m and g.best_assignment is None and len(g.found) == 0 and g.trajectories == []
    assert_same_run(eng.run_sls(gi, "moser", w, w, 0, 7, True, seed=2), oi.run_sls("moser", w, w, 0, 7, True, seed=2))
    # exact solution as initial weights: zero steps
    sol = golden_assignments("ref_tests_medium", gi.n)[0].astype(float)
    g = eng.run_sls(gi, "probsat", sol, sol, 50, 9, True, seed=1)
    assert g.found.all() and (g.steps == 0).all() and (g.best_assignment == sol).all()
    # every clause violated, nsteps = 0: the reference returns an empty best assignment
    gi2 = eng.Instance.from_text("p cnf 2 2\n1 0\n2 0\n")
    g = eng.run_sls(gi2, "moser", [0, 0], [0.5, 0.5], 0, 3)
    assert g.best_energy == 2 and g.best_assignment is None and g.as_reference_tuple()[0] == []
    # chain_offset keeps streams stable under sharding
    a = eng.run_sls(gi, "probsat", w, w, 200, 16, True, seed=4)
    b = eng.run_sls(gi, "probsat", w, w, 200, 8, True, seed=4, chain_offset=8)
    assert all((x == y).all() for x, y in zip(a.trajectories[8:], b.trajectories))
    # without trajectories the other outputs are unchanged
    c = eng.run_sls(gi, "probsat", w, w, 200, 16, False, seed=4)
    assert (a.steps == c.steps).all() and (a.best_assignment == c.best_assignment).all() and c.trajectories == []


def test_walk_error_behaviour():
    gi = eng.Instance.from_text("p cnf 2 2\n1 2 0\n-1 -2 0\n")
    with pytest.raises(eng.PanicException):  # WeightedIndex AllWeightsZero (lib.rs:71)
        eng.run_sls(gi, "probsat", [0, 0], [0.0, 0.0], 10, 4, seed=2)
    with pytest.raises(eng.PanicException):  # gen_bool outside [0,1] (lib.rs:209)
        eng.run_sls(gi, "moser", [0.5, 1.5], [0.5, 0.5], 10, 4)
    with pytest.raises(eng.PanicException):  # short weight vector
        eng.run_sls(gi, "moser", [0.5], [0.5, 0.5], 10, 4)
    gi = eng.Instance.from_text("p cnf 2 2\n0\n1 2 0\n")  # empty clause: choose() on it panics
    with pytest.raises(eng.PanicException):
        eng.run_sls(gi, "schoening", [0.5, 0.5], [0.5, 0.5], 10, 4, seed=1)
    # ... but Moser-Tardos just resamples nothing (lib.rs:24-35) and never terminates early
    g = eng.run_sls(gi, "moser", [0.5, 0.5], [0.5, 0.5], 20, 4, seed=1)
    assert (g.steps == 20).all() and not g.found.any()


def test_run_sls_python_drop_in_tuple():
    import moser_rust

    moser_rust.set_seed(1234)
    path = golden_cnf("r3sat_test_random_KCNF3_100_210_6416893")
    p = np.ones(100) / 2
    best, found, energy, steps, traj = moser_rust.run_sls_python("moser", path, p, p, 999, 20, True, True, 0)
    assert isinstance(best, list) and isinstance(found, list) and isinstance(energy, int) and isinstance(steps, list)
    assert len(found) == 20 and len(steps) == 20 and len(traj) == 20 and len(best) == 100
    padded = np.array([np.pad(t, (0, 1000 - len(t))) for t in traj])  # evaluate_with_given_params.py:18-23
    assert padded.shape == (20, 1000)
    _, _, e2, _, t2 = moser_rust.run_sls_python(
        "probsat", path, p, p, 99, 3, return_trajectories=False, pre_compute_mapping=True, prob_flip_best=0
    )  # train_utils.py:293-303
    assert len(t2) == 0 and isinstance(e2, int)
    moser_rust.set_seed(None)


def test_walk_c2_shape_statistics_and_invariants():
    # BASELINE config 2 shape at reduced chain/step counts: first 3 chains trajectory-exact, all
    # chains consistent with their own outputs
    n, m = 10_000, 42_000
    clauses = random_ksat(n, m, 3, seed=20260102)
    gi = eng.Instance.from_clauses(n, clauses)
    oi = so.Instance.from_clauses(n, clauses)
    w = np.full(n, 0.5)
    g = eng.run_sls(gi, "probsat", w, w, 3000, 256, True, True, 0.0, seed=99)
    o = oi.run_sls("probsat", w, w, 3000, 3, True, True, 0.0, seed=99)
    for r in range(3):
        assert (g.trajectories[r] == o.trajectories[r]).all()
    assert (g.flips_run == g.steps).all() and (g.steps == 3000).all()
    assert oi.energy_distinct(g.best_assignment) == g.best_energy == g.best_energy_run.min()
    assert all(t.min() == e for t, e in zip(g.trajectories, g.best_energy_run))


# ---- batched form: many instances per launch ---------------------------------------------------
@pytest.mark.parametrize("algo", ["schoening", "moser_single"])
def test_batch_table_building_covers_every_flip_rule(expected, algo):
    # a batch builds the per-weights tables of all its instances with one launch per kernel family (clause records for the
    # k <= 3 kernels, flip tables for the generic one, nothing for Schoening's rule): the rules the larger test below leaves out
    names = sorted(expected)[:6]
    insts = [eng.Instance.from_dimacs(golden_cnf(nm)) for nm in names]
    rng = np.random.default_rng(12)
    w = [rng.uniform(0.1, 0.9, i.n) for i in insts]
    batch = eng.run_sls_batch(insts, algo, w, w, 200, 12, True, 0.0, seed=90)
    for i, (nm, b) in enumerate(zip(names, batch)):
        o = so.Instance.from_dimacs(golden_cnf(nm)).run_sls(algo, w[i], w[i], 200, 12, False, True, 0.0, seed=90 + i)
        assert (b.steps == o.steps).all() and (b.found == o.found).all() and b.best_energy == o.best_energy, nm


@pytest.mark.parametrize("algo", ["moser", "probsat"])
@pytest.mark.parametrize("force", [{}, {"SLS_B200_FORCE_STATE": "hbm"}, {"SLS_B200_FORCE_GENERIC": "1"}])
def test_batch_equals_per_instance_runs(expected, algo, force, monkeypatch):
    for k, v in force.items():
        monkeypatch.setenv(k, v)
    names = sorted(expected)  # mixes k<=3 instances (fast kernel) with long-clause G4SAT ones (generic kernel)
    insts = [eng.Instance.from_dimacs(golden_cnf(nm)) for nm in names]
    rng = np.random.default_rng(11)
    w_init = [rng.uniform(0.1, 0.9, i.n) for i in insts]
    w_res = [rng.uniform(0.1, 0.9, i.n) for i in insts]
    batch = eng.run_sls_batch(insts, algo, w_init, w_res, 250, 20, True, 0.1, seed=40)
    assert len(batch) == len(insts)
    for i, (inst, b) in enumerate(zip(insts, batch)):
        one = eng.run_sls(inst, algo, w_init[i], w_res[i], 250, 20, False, True, 0.1, seed=40 + i)
        assert (b.steps == one.steps).all() and (b.found == one.found).all(), names[i]
        assert (b.best_energy_run == one.best_energy_run).all() and (b.flips_run == one.flips_run).all()
        assert b.best_energy == one.best_energy
        assert (b.best_assignment is None) == (one.best_assignment is None)
        if b.best_assignment is not None:
            assert (b.best_assignment == one.best_assignment).all()
    # ... and against the oracle for two of them
    for i in (0, len(insts) - 1):
        oi = so.Instance.from_dimacs(golden_cnf(names[i]))
        o = oi.run_sls(algo, w_init[i], w_res[i], 250, 20, False, True, 0.1, seed=40 + i)
        assert (batch[i].steps == o.steps).all() and batch[i].best_energy == o.best_energy


def test_batch_many_small_instances_config1_shape():
    # BASELINE config 1 shape: hundreds of small random 3-SAT instances x 100 chains in one launch
    insts, sols = [], []
    for i in range(300):
        n = 100 + (i % 3) * 100
        m = int(n * (1.0 + 3.0 * (i % 7) / 7))
        insts.append(eng.Instance.from_clauses(n, random_ksat(n, m, 3, seed=1000 + i)))
    w = [np.full(i.n, 0.5) for i in insts]
    res = eng.run_sls_batch(insts, "probsat", w, w, 2000, 100, True, 0.0, seed=7)
    for inst, r in zip(insts, res):
        assert r.found.shape == (100,) and (r.steps <= 2000).all()
        assert ((r.steps < 2000) <= r.found).all()  # a chain stops early only when it has solved the instance
        if r.found.any():
            viol, energy = eng.eval_assignments(inst, r.best_assignment[None, :])
            assert energy[0] == 0 and r.best_energy == 0
    one = eng.run_sls(insts[17], "probsat", w[17], w[17], 2000, 100, False, True, 0.0, seed=7 + 17)
    assert (res[17].steps == one.steps).all()
    # explicit per-instance seeds: a shard of the batch reproduces the same results (sharding.run_batch_sharded)
    shard = [3, 17, 200]
    part = eng.run_sls_batch([insts[i] for i in shard], "probsat", [w[i] for i in shard], [w[i] for i in shard], 2000, 100,
                             True, 0.0, instance_seeds=[7 + i for i in shard])
    for i, r in zip(shard, part):
        assert (r.steps == res[i].steps).all() and r.best_energy == res[i].best_energy
    with pytest.raises(eng.PanicException):
        eng.run_sls_batch(insts[:2], "walksat", w[:2], w[:2], 10, 1)


@pytest.mark.parametrize("force", [{}, {"SLS_B200_FORCE_STATE": "hbm"}])
def test_walk_more_than_65536_clauses_uses_second_counter_level(force, monkeypatch):
    # > 64 level-1 counters: the rank select goes through the level-2 counters (config-5 tier)
    for k, v in force.items():
        monkeypatch.setenv(k, v)
    n, m = 30_000, 130_000
    clauses = random_ksat(n, m, 3, seed=77)
    gi = eng.Instance.from_clauses(n, clauses)
    oi = so.Instance.from_clauses(n, clauses)
    w = np.full(n, 0.5)
    for algo in ("probsat", "moser"):
        g = eng.run_sls(gi, algo, w, w, 400, 6, True, True, 0.0, seed=3)
        o = oi.run_sls(algo, w, w, 400, 6, True, True, 0.0, seed=3)
        assert_same_run(g, o)


def test_walk_hbm_tier_more_chains_than_one_slice(monkeypatch):
    # the HBM tier initialises chains 256 at a time: 300 chains = one full slice + a partial one
    monkeypatch.setenv("SLS_B200_FORCE_STATE", "hbm")
    name = "r3sat_test_random_KCNF3_200_869_9397183"
    gi, oi = both(name)
    w = np.random.default_rng(2).uniform(0.1, 0.9, gi.n)
    for algo in ("probsat", "moser"):
        g = eng.run_sls(gi, algo, w, w, 60, 300, True, True, 0.0, seed=12)
        o = oi.run_sls(algo, w, w, 60, 300, True, True, 0.0, seed=12)
        assert_same_run(g, o)


@pytest.mark.parametrize("n", [217, 224, 256])  # ragged last assignment word, 32 variables + dummy bit, whole words only
def test_walk_hbm_tier_initialisation_with_certain_variables(n, monkeypatch):
    # weights_initialize with exact 0 and 1 (Bernoulli's ALWAYS_TRUE threshold) and without: the sliced init kernel has an
    # instantiation for either case, a whole-word path (128-bit threshold loads) and a tail path
    monkeypatch.setenv("SLS_B200_FORCE_STATE", "hbm")
    clauses = random_ksat(n, int(3.2 * n), 3, seed=400 + n)
    gi = eng.Instance.from_clauses(n, clauses)
    oi = so.Instance.from_clauses(n, clauses)
    rng = np.random.default_rng(n)
    for with_ones in (True, False):
        w = rng.uniform(0.05, 0.95, n)
        w[rng.choice(n, 9, replace=False)] = 0.0
        if with_ones:
            w[rng.choice(n, 11, replace=False)] = 1.0
            w[n - 1] = 1.0  # the last variable of the ragged word
        g = eng.run_sls(gi, "moser", w, np.clip(w, 0.05, 0.95), 40, 270, True, True, 0.0, seed=21)
        o = oi.run_sls("moser", w, np.clip(w, 0.05, 0.95), 40, 270, True, True, 0.0, seed=21)
        assert_same_run(g, o)


def test_walk_hbm_tier_flip_log_best_assignment(monkeypatch):
    # HBM tier with a short step budget: the best assignment is rebuilt from the flip log
    monkeypatch.setenv("SLS_B200_FORCE_STATE", "hbm")
    n, m = 30_000, 126_000
    clauses = random_ksat(n, m, 3, seed=78)
    gi = eng.Instance.from_clauses(n, clauses)
    oi = so.Instance.from_clauses(n, clauses)
    w = np.full(n, 0.5)
    for algo in ("probsat", "moser", "moser_single", "schoening"):
        for pfb in (0.0, 0.5):
            g = eng.run_sls(gi, algo, w, w, 300, 5, True, True, pfb, seed=4)  # 300 * k + 1 <= 4 * nwords: log in use
            o = oi.run_sls(algo, w, w, 300, 5, True, True, pfb, seed=4)
            assert_same_run(g, o)
    g = eng.run_sls(gi, "probsat", w, w, 5000, 3, False, True, 0.0, seed=4)  # too long for the log: direct copies
    o = oi.run_sls("probsat", w, w, 5000, 3, False, True, 0.0, seed=4)
    assert_same_run(g, o)


def test_trajectory_statistics_match_numpy_protocol():
    # python/src/evaluate_with_given_params.py:18-23, :141-156: pad with zeros, mean / median over runs
    gi, _ = both("r3sat_test_random_KCNF3_200_640_1042686")
    w = np.full(gi.n, 0.5)
    for nruns in (100, 37, 64):
        s = eng.ResidentSolver(gi, "probsat", 1500, nruns, True, True, 0.0, seed=6)
        s.set_weights(w, w)
        s.launch()
        res = s.fetch()
        mean, median = s.trajectory_stats()
        padded = np.array([np.pad(t, (0, 1501 - len(t))) for t in res.trajectories])
        assert res.found.any() and not res.found.all()  # some rows are padded, some are full length
        assert np.allclose(mean, padded.mean(axis=0), rtol=0, atol=1e-12)
        assert (median == np.median(padded, axis=0)).all()
        s.close()


def test_repeated_calls_do_not_leak_device_memory():
    # every call allocates and frees its working set through the stream-ordered pool (csrc/devmem.h)
    import torch

    torch.cuda.mem_get_info()  # creates torch's context before the baseline is taken
    gi, _ = both("r3sat_test_random_KCNF3_200_869_9397183")
    w = np.full(gi.n, 0.5)
    path = golden_cnf("r3sat_test_random_KCNF3_200_869_9397183")
    import moser_rust

    for i in range(3):
        eng.run_sls(gi, "probsat", w, w, 50, 64, True, True, 0.0, seed=i)
        moser_rust.run_sls_python("moser", path, w, w, 50, 64, False, True, 0)
    free0, _ = torch.cuda.mem_get_info()
    for i in range(40):
        eng.run_sls(gi, "probsat", w, w, 50, 64, True, True, 0.0, seed=i)
        moser_rust.run_sls_python("moser", path, w, w, 50, 64, False, True, 0)
        eng.eval_assignments(gi, np.zeros((70, gi.n), dtype=np.uint8))
    free1, _ = torch.cuda.mem_get_info()
    assert free0 - free1 < (64 << 20), (free0, free1)


def _fuzz_instance(rng):
    """Small random formulas covering every shape the kernels distinguish: pure 3-SAT (fast kernel), mixed short
    clauses, long clauses, duplicate clause contents, a variable twice in a clause (also with both signs)."""
    kind = rng.integers(0, 5)
    n = int(rng.integers(1, 60))
    m = int(rng.integers(1, 220))
    clauses = []
    for _ in range(m):
        if kind == 0:
            k = min(3, n)
        elif kind == 1:
            k = int(rng.integers(1, min(3, n) + 1))
        else:
            k = int(rng.integers(1, min(9, n) + 1))
        vs = rng.choice(n, size=k, replace=False) + 1
        c = [int(v) * (1 if rng.integers(0, 2) else -1) for v in vs]
        if kind == 3 and rng.random() < 0.2:
            c.append(c[0] if rng.random() < 0.5 else -c[0])  # repeated variable
        clauses.append(tuple(c))
        if kind >= 3 and rng.random() < 0.15:
            clauses.append(tuple(c))  # duplicate content
    return n, clauses


@pytest.mark.parametrize("seed", range(8))
def test_walk_fuzz_against_oracle(seed):
    rng = np.random.default_rng(1000 + seed)
    for case in range(8):
        n, clauses = _fuzz_instance(rng)
        gi = eng.Instance.from_clauses(n, clauses)
        oi = so.Instance.from_clauses(n, clauses)
        algo = ALGOS[int(rng.integers(0, 4))]
        pfb = float(rng.choice([0.0, 0.0, 0.3, 1.0]))
        mapping = bool(rng.integers(0, 2))
        w_init = rng.uniform(0, 1, n)
        w_res = rng.uniform(0.02, 0.98, n)
        if rng.random() < 0.3:
            w_init[rng.integers(0, n)] = float(rng.integers(0, 2))  # exact 0 / 1 initial probabilities
        nsteps, nruns = int(rng.integers(0, 400)), int(rng.integers(1, 40))
        sd = int(rng.integers(0, 2**62))
        g = eng.run_sls(gi, algo, w_init, w_res, nsteps, nruns, True, mapping, pfb, seed=sd)
        o = oi.run_sls(algo, w_init, w_res, nsteps, nruns, True, mapping, pfb, seed=sd)
        assert_same_run(g, o)
        a = rng.integers(0, 2, size=(int(rng.integers(1, 300)), n))
        viol, energy = eng.eval_assignments(gi, a)
        for b in (0, len(a) - 1):
            assert (viol[b] == oi.eval_clauses(a[b])).all()
            assert energy[b] == oi.energy(a[b])


@pytest.mark.parametrize("algo", ["moser", "probsat"])
@pytest.mark.parametrize("dual", ["quad", "dual", "single"])
def test_batch_two_chains_per_warp_matches_oracle(expected, algo, dual, monkeypatch):
    # without the greedy mix, small k<=3 instances of a batch run two chains per warp (walk_k3d.cuh): every chain
    # must still follow the oracle's trajectory (steps, solved flag, best energy, best assignment), also with a
    # chain count that leaves half-empty warps, boundary weights and short clauses
    # ("quad": the smallest instances -- at most 4096 clauses -- run four chains per warp, the rest two; "dual": two
    # everywhere; "single": walk_k3s_batch, one chain per warp)
    if dual == "single":
        monkeypatch.setenv("SLS_B200_NO_DUAL", "1")
    if dual == "dual":
        monkeypatch.setenv("SLS_B200_NO_QUAD", "1")
    rng = np.random.default_rng(17)
    names = [n for n in expected if n.startswith("r3sat")]
    insts, oracles = [], []
    for n in names:
        g, o = both(n)
        insts.append(g)
        oracles.append(o)
    for n, m, seed in [(40, 150, 1), (700, 2600, 2), (2, 2, 0), (1100, 4096, 3), (1200, 4500, 4)]:  # 4096 clauses: the last four-chain size
        clauses = [(1, 2), (-1, -2)] if n == 2 else _mixed_short_clauses(n, m, seed)
        insts.append(eng.Instance.from_clauses(n, clauses))
        oracles.append(so.Instance.from_clauses(n, clauses))
    w_init = [rng.uniform(0.05, 0.95, g.n) for g in insts]
    w_res = [rng.uniform(0.05, 0.95, g.n) for g in insts]
    if algo == "moser":
        w_res[0][:5] = 0.0
        w_res[0][5:10] = 1.0
    for nruns in (21, 8, 3):
        batch = eng.run_sls_batch(insts, algo, w_init, w_res, 300, nruns, True, 0.0, seed=70)
        for i, (oi, r) in enumerate(zip(oracles, batch)):
            o = oi.run_sls(algo, w_init[i], w_res[i], 300, nruns, False, True, 0.0, seed=70 + i)
            assert (r.steps == o.steps).all(), i
            assert (r.found == o.found).all(), i
            assert (r.best_energy_run == o.best_energy_run).all(), i
            assert (r.flips_run == o.flips_run).all(), i
            assert r.best_energy == o.best_energy
            assert (r.best_assignment is None) == (o.best_assignment is None)
            if r.best_assignment is not None:
                assert (r.best_assignment == o.best_assignment).all(), i


# ---- BASELINE config 5 at full instance size --------------------------------------------------
def test_walk_c5_full_size_hbm_tier_matches_oracle_on_sampled_chains():
    # n = 10^6, m = 4 * 10^6 (the C5 instance of bench.py): 512 chains = two slices of the sliced initialisation, chain state in
    # HBM (640 KB per chain), three-level rank select.  The CPU oracle re-runs single chains by their global index -- first /
    # last chain of each slice and two in the middle -- and every one must match step for step.
    from sls_sat_solving_with_deep_learning_b200 import synth

    n, m, R, S = 1_000_000, 4_000_000, 512, 120
    row, lits = synth.random_ksat_csr(n, m, 3, 20260105)
    gi = eng.Instance.from_csr(n, row, lits)
    oi = so.Instance.from_csr(n, row, lits)
    w = np.full(n, 0.5)
    for algo in ("probsat", "moser"):
        s = eng.ResidentSolver(gi, algo, S, R, True, True, 0.0, seed=9)
        assert s.variant == "walk_k3h<hbm>"
        s.set_weights(w, w)
        s.launch()
        g = s.fetch()
        s.close()
        assert (g.steps == S).all() and not g.found.any()  # alpha = 4.0: nobody solves it in 120 steps
        for j in (0, 131, 255, 256, 300, 511):
            o = oi.run_sls(algo, w, w, S, 1, True, True, 0.0, seed=9, chain_offset=j)
            assert (g.trajectories[j] == o.trajectories[0]).all(), (algo, j)
            assert g.best_energy_run[j] == o.best_energy_run[0] and g.flips_run[j] == o.flips_run[0], (algo, j)
        # the reduction over runs (lib.rs:320-334) and the winner's assignment, checked through the clause evaluation
        winner = int(np.argmin(g.best_energy_run))
        assert g.best_energy == g.best_energy_run[winner]
        assert oi.energy(g.best_assignment) == g.best_energy


# ---- sharding helpers on the CUDA engine --------------------------------------------------------
def test_sharding_helpers_drive_the_cuda_engine():
    # tests/test_sharding_gloo.py exercises sharding.py's collectives with world_size 2 on the CPU oracle; here the same
    # functions run with the CUDA engine behind them and NCCL as the backend (one rank per visible GPU box = 1 here;
    # bench.py --gpus N runs them over N ranks)
    import torch
    import torch.distributed as dist

    from sls_sat_solving_with_deep_learning_b200 import sharding

    torch.cuda.set_device(0)
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", init_method="tcp://127.0.0.1:29731", rank=0, world_size=1,
                                device_id=torch.device("cuda", 0))
        created = True
    try:
        name = "r3sat_test_random_KCNF3_200_869_9397183"
        gi, oi = both(name)
        w = np.random.default_rng(5).uniform(0.2, 0.8, gi.n)
        R = 37
        res = sharding.run_sls_sharded(
            lambda off, cnt: eng.run_sls(gi, "probsat", w, w, 3000, cnt, False, True, 0.0, seed=21, chain_offset=off),
            R, gi.n, gi.m, device="cuda")
        o = oi.run_sls("probsat", w, w, 3000, R, False, True, 0.0, seed=21)
        assert (res.steps == o.steps).all() and (res.found == o.found).all() and res.best_energy == o.best_energy
        assert (res.best_assignment is None) == (o.best_assignment is None)
        if o.best_assignment is not None:
            assert (res.best_assignment == o.best_assignment).all()
        # a batch split by instance: every instance keeps its own seed
        names = [name, "r3sat_test_random_KCNF3_100_210_6416893", "ref_tests_medium"]
        insts = [eng.Instance.from_dimacs(golden_cnf(nm)) for nm in names]
        oracles = [so.Instance.from_dimacs(golden_cnf(nm)) for nm in names]
        ws = [np.full(i.n, 0.5) for i in insts]

        def runner(idx):
            return eng.run_sls_batch([insts[i] for i in idx], "moser", [ws[i] for i in idx], [ws[i] for i in idx], 500, 16,
                                     instance_seeds=[100 + i for i in idx])

        found, steps, best, flips = sharding.run_batch_sharded(runner, [i.m for i in insts], 16, device="cuda")
        for i, oi_ in enumerate(oracles):
            o = oi_.run_sls("moser", ws[i], ws[i], 500, 16, False, True, 0.0, seed=100 + i)
            assert (steps[i] == o.steps).all() and (found[i] == o.found).all() and best[i] == o.best_energy and flips[i] == o.total_flips
    finally:
        if created:
            dist.destroy_process_group()


def test_instance_cannot_migrate_under_a_live_solver():
    # one device image per handle: with a solver holding pointers into it, moving the image to another GPU is refused
    # (single-GPU box: the guard's bookkeeping is checked through solver lifetimes on the same device)
    gi, _ = both("ref_tests_medium")
    w = np.full(gi.n, 0.5)
    s1 = eng.ResidentSolver(gi, "probsat", 50, 4, False, True, 0.0, seed=1)
    s2 = eng.ResidentSolver(gi, "moser", 50, 4, False, True, 0.0, seed=1)
    s1.set_weights(w, w)
    s1.launch()
    s1.close()
    s2.set_weights(w, w)
    s2.launch()
    r = s2.fetch()
    s2.close()
    assert r.steps.shape == (4,)
    if eng.device_count() > 1:
        s3 = eng.ResidentSolver(gi, "probsat", 50, 4, False, True, 0.0, seed=1)
        eng.set_device(1)
        with pytest.raises(eng.PanicException):
            eng.run_sls(gi, "probsat", w, w, 10, 2)
        eng.set_device(0)
        s3.close()
