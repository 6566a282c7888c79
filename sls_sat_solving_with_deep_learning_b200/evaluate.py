"""The reference's evaluation driver on the B200 engine.

Mirrors of ``load_model_and_test`` (python/src/evaluate_with_given_params.py:26-174) and
``load_model_and_test_two_models`` (:177-345): same arguments, same returned / saved ``total_array`` schema
``[[model_path], [model_details_list], n_array, alpha_array, energies_array_mean, energies_array_median,
total_steps_array]``.  What changes is how the work is scheduled: the reference loops over the instances and,
for each, runs an un-jitted GNN forward and one ``run_sls_python`` call; here all oracle probabilities come from
one batched forward (:meth:`gnn.InteractionNetworkLCG.probabilities`), all instances are solved by one batched
launch when no trajectories are requested (:func:`solver.run_sls_batch`), and when they are requested the
per-step mean / median over runs is reduced on the device (:meth:`solver.ResidentSolver.trajectory_stats`)
instead of shipping ``[n_runs, n_steps]`` arrays to NumPy (evaluate_with_given_params.py:18-23, :141-156).

Not mirrored: the dataset class's padding / candidate machinery (python/src/data_utils.py), which the evaluation
path does not use (it calls ``get_unpadded_problem`` only).
"""
from __future__ import annotations

import glob
import os
from os.path import join

import numpy as np

# The trajectory-keeping protocol runs waves of up to 32 small, latency-bound launches on their own streams.  The driver maps
# streams onto CUDA_DEVICE_MAX_CONNECTIONS hardware queues (default 8) and launches that share a queue serialise: measured
# 0.319 s per wave of 32 instances with 8 queues, 0.111 s with 32.  The variable is read when the CUDA context is created,
# so it is set when THIS driver is imported (import it before the first CUDA call) and only if the user has not set it;
# importing the package or loading libsls_b200.so alone changes no process-wide state.
os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")

from . import gnn, model_io  # noqa: E402
from .solver import Instance, ResidentSolver, run_sls_batch  # noqa: E402


def _streams(count: int):
    """Raw CUDA stream handles for concurrent launches (torch is used for plumbing only); the legacy default stream
    alone when torch is not importable."""
    try:
        import torch

        _streams.keep = [torch.cuda.Stream() for _ in range(count)]  # keep the objects alive
        return [st.cuda_stream for st in _streams.keep]
    except Exception:  # pragma: no cover
        return [0]


def dataset_instances(data_path: str):
    """File stems in the order SATTrainingDataset enumerates them (data_utils.py:47-60): every ``*_sol.pkl``."""
    return [p.split("_sol.pkl")[0] for p in glob.glob(join(data_path, "*_sol.pkl"))]


def _oracle(model_path, insts):
    """(model_details_list, probabilities per instance) of one oracle, ``"uniform"`` included (:67-88, :118-123)."""
    if model_path != "uniform":
        net, md = model_io.load_oracle(model_path)
        details = [md.inv_temp, md.alpha, md.beta, md.gamma, md.mlp_layers, md.graph_representation, md.network_type,
                   md.return_candidates]
        return details, net.probabilities(insts)  # one batched forward instead of one apply per instance (:116)
    return [], [np.ones(i.n) / 2 for i in insts]  # :123


def _evaluate(data_path, probs_of, n_steps, n_runs, algo_type, keep_traj, pre_compute_mapping, prob_flip_best, seed):
    """The part both drivers share: instances, oracle probabilities for initialisation and resampling, the walks, the
    per-step statistics.  ``probs_of(insts)`` returns (probabilities_initialize, probabilities_resample)."""
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    stems = dataset_instances(data_path)
    insts = [Instance.from_dimacs(s + ".cnf") for s in stems]
    probs_i, probs_r = probs_of(insts)

    n_array, alpha_array = [], []
    for inst in insts:
        row, _ = inst.csr()
        m_nonempty = int(np.count_nonzero(np.diff(row.astype(np.int64)) > 0))  # sat_instances.py:50 drops empty clauses
        n_array.append(inst.n)
        alpha_array.append(m_nonempty / inst.n)
    seeds = [(seed + idx) & 0xFFFFFFFFFFFFFFFF for idx in range(len(insts))]

    total_steps = []
    energies_array_mean, energies_array_median = [], []
    if not keep_traj:
        results = run_sls_batch(insts, algo_type, probs_i, probs_r, n_steps - 1, n_runs, pre_compute_mapping, prob_flip_best,
                                instance_seeds=seeds) if insts else []
        total_steps = [[int(s) for s in r.steps] for r in results]
    else:
        # One instance with n_runs chains is a small, latency-bound launch (a chain is a serial walk), so the
        # instances of a wave run concurrently, each on its own CUDA stream; the wave is bounded by the trajectory
        # buffers (n_runs x n_steps x 4 bytes per instance).
        wave = max(1, min(32, int((8 << 30) // max(1, 4 * n_runs * n_steps))))
        streams = _streams(wave)
        for w0 in range(0, len(insts), wave):
            solvers = []
            try:
                for inst, pi, pr, sd in zip(insts[w0:w0 + wave], probs_i[w0:w0 + wave], probs_r[w0:w0 + wave], seeds[w0:w0 + wave]):
                    s = ResidentSolver(inst, algo_type, n_steps - 1, n_runs, True, pre_compute_mapping, prob_flip_best, seed=sd)
                    solvers.append(s)
                    s.set_weights(pi, pr)
                for k, s in enumerate(solvers):
                    s.launch(streams[k % len(streams)])
                for s, inst, alpha in zip(solvers, insts[w0:w0 + wave], alpha_array[w0:w0 + wave]):
                    res = s.fetch_summary()
                    mean, median = s.trajectory_stats()  # zero-padded to n_steps entries, reduced over runs on the GPU
                    total_steps.append([int(x) for x in res.steps])
                    m_nonempty = alpha * inst.n
                    energies_array_mean.append(mean / m_nonempty)      # :143, :145-150
                    energies_array_median.append(median / m_nonempty)  # :144, :151-156
            finally:
                for s in solvers:
                    s.close()
    total_steps_array = np.asarray(total_steps)
    if energies_array_mean:
        energies_array_mean = np.mean(energies_array_mean, axis=0)      # :161
        energies_array_median = np.mean(energies_array_median, axis=0)  # :162
    return n_array, alpha_array, energies_array_mean, energies_array_median, total_steps_array


def load_model_and_test(
    data_path,
    model_path,
    n_steps: int,
    n_runs: int,
    algo_type: str,
    path_save=False,
    keep_traj: bool = True,
    pre_compute_mapping: bool = False,
    prob_flip_best: float = 0,
    seed: int | None = None,
):
    """See the module docstring.  ``seed`` (not in the reference, whose chains are seeded from OS entropy,
    rust/src/lib.rs:206) makes the run reproducible: instance ``idx`` uses ``seed + idx``."""
    details = []

    def probs_of(insts):
        d, p = _oracle(model_path, insts)
        details.append(d)
        return p, p

    rest = _evaluate(data_path, probs_of, n_steps, n_runs, algo_type, keep_traj, pre_compute_mapping, prob_flip_best, seed)
    total_array = [[model_path], [details[0]], *rest]
    if path_save:
        np.save(path_save, np.array(total_array, dtype=object))
    return total_array


def load_model_and_test_two_models(
    data_path,
    model_path_initialize,
    model_path_resample,
    n_steps: int,
    n_runs: int,
    algo_type,
    path_save=False,
    keep_traj: bool = True,
    pre_compute_mapping: bool = True,
    prob_flip_best: float = 0,
    seed: int | None = None,
):
    """Mirror of ``load_model_and_test_two_models`` (python/src/evaluate_with_given_params.py:177-345): one oracle (or
    ``"uniform"``) for the initial assignment, another for the resampling / flip rule.  Returned / saved schema:
    ``[[model_path_initialize, model_path_resample], [model_details_list_i, model_details_list_r], n_array, alpha_array,
    energies_array_mean, energies_array_median, total_steps_array]``."""
    details = []

    def probs_of(insts):
        di, pi = _oracle(model_path_initialize, insts)
        if model_path_resample == model_path_initialize:
            dr, pr = di, pi  # same file: one forward
        else:
            dr, pr = _oracle(model_path_resample, insts)
        details.extend([di, dr])
        return pi, pr

    rest = _evaluate(data_path, probs_of, n_steps, n_runs, algo_type, keep_traj, pre_compute_mapping, prob_flip_best, seed)
    total_array = [[model_path_initialize, model_path_resample], [details[0], details[1]], *rest]
    if path_save:
        np.save(path_save, np.array(total_array, dtype=object))
    return total_array


def evaluate_moser_rust(cnf_paths, net, mode_probabilities: str = "model", n_steps_moser: int = 100, n_runs_moser: int = 1,
                        seed: int | None = None):
    """The training loop's Moser-Tardos probe (python/src/train_utils.py:244-318) on the batched path: for every
    instance the oracle-guided MT walk (``"moser"``, ``n_steps_moser - 1`` steps, ``n_runs_moser`` runs, mapping on,
    no greedy mix -- the keyword call of :293-303), of which only the best energy is used (:293, :305), and the mean
    binary entropy of the oracle's probabilities (:306-316).  Returns ``(mean energy / n_clauses, mean entropy)``.

    The reference walks a ``data_subset`` of its dataset object with a haiku network; here the caller passes the
    ``.cnf`` paths (or parsed :class:`Instance` objects) of that subset and an :class:`gnn.InteractionNetworkLCG`
    (ignored for ``mode_probabilities="uniform"``).  One batched forward, one batched solve."""
    insts = [p if isinstance(p, Instance) else Instance.from_dimacs(p) for p in cnf_paths]
    if mode_probabilities == "uniform":
        probs = [np.ones(i.n) / 2 for i in insts]
    elif mode_probabilities == "model":
        probs = net.probabilities(insts)
    else:
        raise ValueError("not valid argument for mode_probabilities")  # the reference prints this and then fails (:284)
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    seeds = [(seed + idx) & 0xFFFFFFFFFFFFFFFF for idx in range(len(insts))]
    results = run_sls_batch(insts, "moser", probs, probs, n_steps_moser - 1, n_runs_moser, True, 0.0, instance_seeds=seeds)
    av_energies, av_entropies = [], []
    for inst, p, r in zip(insts, probs, results):
        row, _ = inst.csr()
        n_clauses = int(np.count_nonzero(np.diff(row.astype(np.int64)) > 0))  # problem.params[1]: empty clauses dropped
        av_energies.append(r.best_energy / n_clauses)
        q = np.clip(np.asarray(p, dtype=np.float64), 0.0, 1.0)
        with np.errstate(divide="ignore", invalid="ignore"):
            h = -(np.where(q > 0, q * np.log2(q), 0.0) + np.where(q < 1, (1 - q) * np.log2(1 - q), 0.0))
        av_entropies.append(float(np.mean(h)))  # scipy.stats.entropy(base=2) per variable, then the mean (:311-316)
    return (float(np.mean(av_energies)), float(np.mean(av_entropies)))
