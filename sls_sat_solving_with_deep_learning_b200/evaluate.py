"""The reference's evaluation driver on the B200 engine.

Mirror of ``load_model_and_test`` (python/src/evaluate_with_given_params.py:26-174): same arguments, same
returned / saved ``total_array`` schema
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

from . import gnn, model_io
from .solver import Instance, ResidentSolver, run_sls_batch


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
    if seed is None:
        seed = int.from_bytes(os.urandom(8), "little")
    stems = dataset_instances(data_path)
    insts = [Instance.from_dimacs(s + ".cnf") for s in stems]
    if model_path != "uniform":
        net, md = model_io.load_oracle(model_path)
        model_details_list = [md.inv_temp, md.alpha, md.beta, md.gamma, md.mlp_layers, md.graph_representation,
                              md.network_type, md.return_candidates]
        probs = net.probabilities(insts)  # one batched forward instead of one apply per instance (:116)
    else:
        model_details_list = []
        probs = [np.ones(i.n) / 2 for i in insts]  # :123

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
        results = run_sls_batch(insts, algo_type, probs, probs, n_steps - 1, n_runs, pre_compute_mapping, prob_flip_best,
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
                for k, (inst, p, sd) in enumerate(zip(insts[w0:w0 + wave], probs[w0:w0 + wave], seeds[w0:w0 + wave])):
                    s = ResidentSolver(inst, algo_type, n_steps - 1, n_runs, True, pre_compute_mapping, prob_flip_best, seed=sd)
                    solvers.append(s)
                    s.set_weights(p, p)
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
    total_array = [[model_path], [model_details_list], n_array, alpha_array, energies_array_mean, energies_array_median,
                   total_steps_array]
    if path_save:
        np.save(path_save, np.array(total_array, dtype=object))
    return total_array
