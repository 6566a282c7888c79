"""The evaluation driver mirror (load_model_and_test, python/src/evaluate_with_given_params.py:26-174) against
the same protocol computed literally, in NumPy, from CPU-oracle trajectories with the same seeds."""
import os
import pickle
import shutil

import numpy as np
import pytest

from conftest import golden_cnf
from oracle import gnn_oracle as go
from oracle import sls_oracle as so
import sls_sat_solving_with_deep_learning_b200 as eng
from sls_sat_solving_with_deep_learning_b200 import evaluate

pytestmark = pytest.mark.gpu

NAMES = ["r3sat_test_random_KCNF3_100_210_6416893", "r3sat_test_random_KCNF3_100_325_6020176",
         "r3sat_test_random_KCNF3_200_640_1042686", "g4sat_easy_ca_00092"]


@pytest.fixture()
def data_dir(tmp_path):
    for nm in NAMES:
        shutil.copyfile(golden_cnf(nm), tmp_path / (nm + ".cnf"))
        pickle.dump([1], open(tmp_path / (nm + "_sol.pkl"), "wb"))  # only its presence matters (data_utils.py:47)
    return str(tmp_path)


def reference_protocol(data_path, probs_of, n_steps, n_runs, algo, keep_traj, mapping, pfb, seed):
    """evaluate_with_given_params.py:104-174 restated with the CPU oracle in place of moser_rust."""
    stems = evaluate.dataset_instances(data_path)
    n_array, alpha_array, total_steps, e_mean, e_median = [], [], [], [], []
    for idx, stem in enumerate(stems):
        oi = so.Instance.from_dimacs(stem + ".cnf")
        row, _ = oi.csr()
        m = int(np.count_nonzero(np.diff(row.astype(np.int64)) > 0))
        n_array.append(oi.n)
        alpha_array.append(m / oi.n)
        p = probs_of(stem, oi)
        r = oi.run_sls(algo, p, p, n_steps - 1, n_runs, keep_traj, mapping, pfb, seed=seed + idx)
        total_steps.append([int(x) for x in r.steps])
        if keep_traj:
            traj = np.array([np.pad(t, (0, n_steps - len(t))) for t in r.trajectories])  # get_padded_trajs, :18-23
            e_mean.append(np.mean(traj, axis=0) / m)
            e_median.append(np.median(traj, axis=0) / m)
    if e_mean:
        e_mean, e_median = np.mean(e_mean, axis=0), np.mean(e_median, axis=0)
    return n_array, alpha_array, e_mean, e_median, np.asarray(total_steps)


@pytest.mark.parametrize("keep_traj", [True, False])
@pytest.mark.parametrize("algo", ["moser", "probsat"])
def test_uniform_oracle_protocol(data_dir, tmp_path, keep_traj, algo):
    eng.set_device(0)
    out = str(tmp_path / "res.npy")
    got = evaluate.load_model_and_test(data_dir, "uniform", 400, 20, algo, path_save=out, keep_traj=keep_traj,
                                       pre_compute_mapping=True, prob_flip_best=0, seed=100)
    ref = reference_protocol(data_dir, lambda stem, oi: np.ones(oi.n) / 2, 400, 20, algo, keep_traj, True, 0.0, 100)
    assert got[0] == ["uniform"] and got[1] == [[]]
    assert got[2] == ref[0] and np.allclose(got[3], ref[1])
    assert (got[6] == ref[4]).all()
    if keep_traj:
        assert np.allclose(got[4], ref[2], atol=1e-12) and np.allclose(got[5], ref[3], atol=1e-12)
        assert len(got[4]) == 400
    else:
        assert len(got[4]) == 0 and len(got[5]) == 0
    saved = np.load(out, allow_pickle=True)  # the schema the notebook reads back (tutorial.ipynb)
    assert len(saved) == 7 and (np.asarray(saved[6]) == ref[4]).all()


def test_model_file_protocol(data_dir, tmp_path):
    # a model file as train.py:221-239 writes it (plain arrays instead of jax arrays), read by model_io
    eng.set_device(0)
    params = go.init_params(seed=3, bias_scale=0.05)
    lcg = type("LCG", (), {})
    lcg.__module__ = "python.src.sat_representations"
    import sys
    import types

    mod = types.ModuleType("python.src.sat_representations")
    mod.LCG = lcg
    sys.modules.update({"python": types.ModuleType("python"), "python.src": types.ModuleType("python.src"),
                        "python.src.sat_representations": mod})
    try:
        model = str(tmp_path / "model.npy")
        np.save(model, np.asarray([params, [2.0, 1.0, 0.0, 0.0, [200, 200], lcg, "interaction", False]], dtype=object))
    finally:
        for k in ("python", "python.src", "python.src.sat_representations"):
            sys.modules.pop(k, None)

    def probs_of(stem, oi):
        row, lits = oi.csr()
        clauses = [lits[int(row[c]):int(row[c + 1])].tolist() for c in range(oi.m)]
        dec = go.forward(params, go.lcg_graph(oi.n, clauses), dtype=np.float64)
        return go.model_probabilities(dec, oi.n)

    got = evaluate.load_model_and_test(data_dir, model, 200, 10, "moser", keep_traj=False, pre_compute_mapping=True, seed=5)
    assert got[1] == [[2.0, 1.0, 0.0, 0.0, [200, 200], "LCG", "interaction", False]]
    # the GPU probabilities differ from the fp64 restatement by ~1e-5, which can move a Bernoulli threshold:
    # compare solve statistics, not trajectories
    ref = reference_protocol(data_dir, probs_of, 200, 10, "moser", False, True, 0.0, 5)
    assert got[6].shape == ref[4].shape
    assert abs(got[6].mean() - ref[4].mean()) < 0.25 * ref[4].mean() + 5


def test_two_models_protocol(data_dir, tmp_path):
    # load_model_and_test_two_models (evaluate_with_given_params.py:177-345): separate oracles for the initial assignment
    # and for the flip rule.  Uniform / uniform must reproduce the one-model driver; the schema carries both paths.
    eng.set_device(0)
    out = str(tmp_path / "two.npy")
    two = evaluate.load_model_and_test_two_models(data_dir, "uniform", "uniform", 300, 12, "probsat", path_save=out,
                                                  keep_traj=True, pre_compute_mapping=True, prob_flip_best=0, seed=31)
    one = evaluate.load_model_and_test(data_dir, "uniform", 300, 12, "probsat", keep_traj=True, pre_compute_mapping=True,
                                       prob_flip_best=0, seed=31)
    assert two[0] == ["uniform", "uniform"] and two[1] == [[], []]
    assert two[2] == one[2] and two[3] == one[3] and (two[6] == one[6]).all()
    assert np.array_equal(two[4], one[4]) and np.array_equal(two[5], one[5])
    assert len(np.load(out, allow_pickle=True)) == 7
    # initial assignment from the exact solution (weights 0 / 1), uniform flip rule: every chain starts solved
    stems = evaluate.dataset_instances(data_dir)

    class Exact:  # stands in for an oracle file: _oracle() is the only consumer of model_io.load_oracle
        pass

    sols = []
    for stem in stems:
        oi = so.Instance.from_dimacs(stem + ".cnf")
        r = oi.run_sls("probsat", np.full(oi.n, 0.5), np.full(oi.n, 0.5), 200000, 8, False, True, 0.0, seed=1)
        assert r.best_energy == 0
        sols.append(r.best_assignment.astype(np.float64))
    orig = evaluate._oracle
    evaluate._oracle = lambda path, insts: (["exact"], sols) if path == "exact" else orig(path, insts)
    try:
        got = evaluate.load_model_and_test_two_models(data_dir, "exact", "uniform", 50, 6, "moser", keep_traj=False, seed=2)
    finally:
        evaluate._oracle = orig
    assert got[1] == [["exact"], []] and (got[6] == 0).all()


def test_training_probe_evaluate_moser_rust(data_dir):
    # train_utils.py:244-318: best energy of an oracle-guided MT run per instance / n_clauses, and the oracle's entropy
    from scipy.stats import entropy

    eng.set_device(0)
    stems = evaluate.dataset_instances(data_dir)
    paths = [s + ".cnf" for s in stems]
    e_gpu, h_gpu = evaluate.evaluate_moser_rust(paths, None, "uniform", n_steps_moser=60, n_runs_moser=3, seed=9)
    energies = []
    for idx, p in enumerate(paths):
        oi = so.Instance.from_dimacs(p)
        w = np.ones(oi.n) / 2
        r = oi.run_sls("moser", w, w, 59, 3, False, True, 0.0, seed=9 + idx)
        energies.append(r.best_energy / oi.m)
    assert e_gpu == pytest.approx(np.mean(energies), abs=1e-12) and h_gpu == pytest.approx(1.0)
    # with a network: entropies as scipy computes them from the GPU probabilities
    from sls_sat_solving_with_deep_learning_b200 import gnn

    net = gnn.InteractionNetworkLCG(go.init_params(seed=4, bias_scale=0.2))
    insts = [eng.Instance.from_dimacs(p) for p in paths]
    _, h_model = evaluate.evaluate_moser_rust(insts, net, "model", n_steps_moser=20, n_runs_moser=2, seed=1)
    probs = net.probabilities(insts)
    ref = np.mean([np.mean([entropy([1 - q, q], base=2) for q in p]) for p in probs])
    assert h_model == pytest.approx(ref, rel=1e-6)
    with pytest.raises(ValueError):
        evaluate.evaluate_moser_rust(paths, None, "something")
