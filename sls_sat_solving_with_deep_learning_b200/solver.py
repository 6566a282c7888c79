"""Host-side mirror of the reference solver interface on top of the C ABI.

Reference interface mirrored here (relative to the reference repository):
  * ``moser_rust.run_sls_python``  rust/src/lib.rs:103-149  -> :func:`run_sls_python`
  * ``{LCG,VCG}.get_violated_constraints``  python/src/sat_representations.py:718-741, :353-373
    -> :func:`get_violated_constraints` / :func:`eval_assignments`

Everything computes on the GPU through ``libsls_b200.so``; nothing here falls back to the CPU.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from dataclasses import dataclass, field

import numpy as np

from . import _lib

ALGOS = {"moser": 0, "moser_single": 1, "schoening": 2, "probsat": 3}  # lib.rs:120-126


class PanicException(BaseException):
    """Stand-in for ``pyo3_runtime.PanicException``.

    Every failure of the reference is a Rust panic (lib.rs:115, :118, :125, :71, :209), which pyo3
    surfaces as an exception deriving from BaseException; call sites never catch it.
    """

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def _check(rc: int):
    if rc != 0:
        raise PanicException(rc, _lib.load().sls_last_error().decode())


def device_count() -> int:
    c = C.c_int(0)
    _lib.load().sls_device_count(C.byref(c))
    return c.value


_thread_device = threading.local()


def set_device(device: int) -> None:
    """Select the CUDA device of the calling thread (cudaSetDevice semantics: a new thread starts on device 0)."""
    _check(_lib.load().sls_set_device(device))
    _thread_device.index = int(device)


def reserve_device_pool(nbytes: int) -> None:
    """Grow the library's device memory pool to ``nbytes`` now (capped at its 4 GB release threshold) instead of step by step
    under running kernels (sls_device_pool_reserve)."""
    _check(_lib.load().sls_device_pool_reserve(int(max(nbytes, 0))))


def current_device() -> int:
    """The device this thread selected with :func:`set_device` (0 if it never did)."""
    return getattr(_thread_device, "index", 0)


class Instance:
    """A parsed CNF formula; its clause and occurrence CSR live in HBM once first used."""

    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_dimacs(cls, path) -> "Instance":
        h = C.c_void_p()
        _check(_lib.load().sls_instance_from_dimacs(os.fsencode(path), C.byref(h)))
        return cls(h)

    @classmethod
    def from_text(cls, text: str) -> "Instance":
        b = text.encode()
        h = C.c_void_p()
        _check(_lib.load().sls_instance_from_dimacs_text(b, len(b), C.byref(h)))
        return cls(h)

    @classmethod
    def from_clauses(cls, n: int, clauses) -> "Instance":
        row = np.zeros(len(clauses) + 1, dtype=np.uint64)
        if len(clauses):
            row[1:] = np.cumsum([len(c) for c in clauses])
        flat = np.asarray([l for c in clauses for l in c], dtype=np.int32)
        return cls.from_csr(n, row, flat)

    @classmethod
    def from_csr(cls, n: int, row_ptr, signed_lits) -> "Instance":
        row = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        flat = np.ascontiguousarray(signed_lits, dtype=np.int32)
        h = C.c_void_p()
        _check(_lib.load().sls_instance_from_csr(n, len(row) - 1, row.ctypes.data, flat.ctypes.data, C.byref(h)))
        return cls(h)

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().sls_instance_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:  # interpreter shutdown
            pass

    @property
    def n(self) -> int:
        return _lib.load().sls_instance_num_vars(self._h)

    @property
    def m(self) -> int:
        return _lib.load().sls_instance_num_clauses(self._h)

    @property
    def num_lits(self) -> int:
        return _lib.load().sls_instance_num_lits(self._h)

    @property
    def max_clause_len(self) -> int:
        return _lib.load().sls_instance_max_clause_len(self._h)

    def csr(self):
        row = np.zeros(self.m + 1, dtype=np.uint64)
        lits = np.zeros(max(self.num_lits, 1), dtype=np.int32)
        _check(_lib.load().sls_instance_get_csr(self._h, row.ctypes.data, lits.ctypes.data))
        return row, lits[: self.num_lits]


def eval_assignments(inst: Instance, assignments, want_bits: bool = True):
    """Batched clause evaluation on the GPU.

    assignments: [B, n] array of 0/1.  Returns (violated[B, m] uint8 or None, energy[B] uint32).
    """
    a = np.asarray(assignments)
    if a.ndim == 1:
        a = a[None, :]
    a = np.ascontiguousarray(a.astype(bool).astype(np.uint8))
    if a.shape[1] != inst.n:
        raise ValueError(f"assignments have {a.shape[1]} variables, the instance has {inst.n}")
    B = a.shape[0]
    m = inst.m
    mwords = (m + 31) // 32
    energy = np.zeros(max(B, 1), dtype=np.uint32)
    bits = np.zeros((max(B, 1), max(mwords, 1)), dtype=np.uint32) if want_bits else None
    _check(
        _lib.load().sls_eval_assignments(
            inst._h, a.ctypes.data, B, bits.ctypes.data if bits is not None else None, energy.ctypes.data
        )
    )
    violated = None
    if want_bits:
        violated = np.unpackbits(bits[:B, :mwords].view(np.uint8), axis=1, bitorder="little")[:, :m]
    return violated, energy[:B]


def eval_last_kernel_ms() -> float:
    """Device time of this thread's last :func:`eval_assignments` (transposition + evaluation kernels, no copies)."""
    return float(_lib.load().sls_eval_last_kernel_ms())


def eval_last_eval_kernel_ms() -> float:
    """... of its evaluation kernel alone (the transposition of the assignment bytes into bit slices excluded)."""
    return float(_lib.load().sls_eval_last_eval_kernel_ms())


def candidate_energies(inst: Instance, candidates) -> np.ndarray:
    """Energies of a stack of candidate assignments as the training loader computes them
    (python/src/data_utils.py:123-129: ``vmap(get_violated_constraints)`` over the candidates, summed and divided
    by the number of clauses).  The Python problem drops empty clauses when it reads the CNF
    (sat_instances.py:50); they are removed from both the count and the denominator here."""
    _, energy = eval_assignments(inst, candidates, want_bits=False)
    row, _ = inst.csr()
    n_empty = int(np.count_nonzero(np.diff(row.astype(np.int64)) == 0))
    m = inst.m - n_empty
    if m == 0:
        return np.zeros(len(energy), dtype=np.float32)
    return ((energy.astype(np.int64) - n_empty) / m).astype(np.float32)


def get_violated_constraints(inst: Instance, assignment) -> np.ndarray:
    """sat_representations.py:718-741: 1 for each violated clause, 0 otherwise."""
    v, _ = eval_assignments(inst, np.asarray(assignment)[None, :])
    return v[0].astype(np.int32)


@dataclass
class SlsResult:
    best_assignment: np.ndarray | None
    found: np.ndarray
    best_energy: int
    steps: np.ndarray
    best_energy_run: np.ndarray
    flips_run: np.ndarray
    trajectories: list = field(default_factory=list)
    total_steps: int = 0
    total_flips: int = 0
    kernel_ms: float = 0.0

    def as_reference_tuple(self):
        """The positional 5-tuple Python receives from lib.rs:141-147."""
        ba = [] if self.best_assignment is None else [bool(x) for x in self.best_assignment]
        return (
            ba,
            [bool(x) for x in self.found],
            int(self.best_energy),
            [int(x) for x in self.steps],
            self.trajectories,
        )


def _make_params(algo, nsteps, nruns, return_trajectories, pre_compute_mapping, prob_flip_best, seed, chain_offset):
    if algo not in ALGOS:
        raise PanicException(3, "Unknown algorithm type")  # lib.rs:125
    if nsteps < 0 or nruns < 0:
        raise OverflowError("can't convert negative int to unsigned")  # pyo3 usize extraction
    return _lib.RunParams(
        ALGOS[algo], int(bool(return_trajectories)), int(bool(pre_compute_mapping)), 0, int(nsteps), int(nruns),
        float(prob_flip_best), int(seed) & 0xFFFFFFFFFFFFFFFF, int(chain_offset),
    )


class _Buffers:
    def __init__(self, n, nruns, nsteps, want_traj):
        r1 = max(nruns, 1)
        self.best = np.zeros(max(n, 1), dtype=np.uint8)
        self.found = np.zeros(r1, dtype=np.uint8)
        self.steps = np.zeros(r1, dtype=np.uint64)
        self.ber = np.zeros(r1, dtype=np.uint64)
        self.flips = np.zeros(r1, dtype=np.uint64)
        self.traj = np.zeros((r1, nsteps + 1), dtype=np.uint32) if want_traj else None
        self.traj_len = np.zeros(r1, dtype=np.uint64) if want_traj else None
        self.res = _lib.RunResult(
            self.best.ctypes.data, self.found.ctypes.data, self.steps.ctypes.data, self.ber.ctypes.data,
            self.flips.ctypes.data,
            self.traj.ctypes.data if want_traj else None,
            self.traj_len.ctypes.data if want_traj else None,
            0, 0, 0, 0, 0, 0.0,
        )

    def to_result(self, n, nruns, want_traj) -> SlsResult:
        r = self.res
        trajs = []
        if want_traj:
            trajs = [self.traj[i, : int(self.traj_len[i])] for i in range(nruns)]
        return SlsResult(
            best_assignment=self.best[:n].copy() if r.has_best_assignment else None,
            found=self.found[:nruns].astype(bool),
            best_energy=int(r.best_energy),
            steps=self.steps[:nruns].copy(),
            best_energy_run=self.ber[:nruns].copy(),
            flips_run=self.flips[:nruns].copy(),
            trajectories=trajs,
            total_steps=int(r.total_steps),
            total_flips=int(r.total_flips),
            kernel_ms=float(r.kernel_ms),
        )


def _weights(w):
    return np.ascontiguousarray(np.asarray(w, dtype=np.float64).ravel())


def run_sls(
    inst: Instance,
    algo: str,
    weights_initialize,
    weights_resample,
    nsteps: int,
    nruns: int,
    return_trajectories: bool = False,
    pre_compute_mapping: bool = True,
    prob_flip_best: float = 0.0,
    seed: int = 0,
    chain_offset: int = 0,
) -> SlsResult:
    """run_sls (lib.rs:165-335) on an already parsed instance, with an explicit seed."""
    if return_trajectories:
        # resident path: only the valid prefix of every trajectory row crosses PCIe
        s = ResidentSolver(inst, algo, nsteps, nruns, True, pre_compute_mapping, prob_flip_best, seed, chain_offset)
        try:
            s.set_weights(weights_initialize, weights_resample)
            s.launch()
            return s.fetch()
        finally:
            s.close()
    p = _make_params(algo, nsteps, nruns, False, pre_compute_mapping, prob_flip_best, seed, chain_offset)
    wi, wr = _weights(weights_initialize), _weights(weights_resample)
    buf = _Buffers(inst.n, nruns, nsteps, False)
    _check(
        _lib.load().sls_run(inst._h, wi.ctypes.data, wi.size, wr.ctypes.data, wr.size, C.byref(p), C.byref(buf.res))
    )
    return buf.to_result(inst.n, nruns, False)


def run_sls_batch(
    instances,
    algo: str,
    weights_initialize,
    weights_resample,
    nsteps: int,
    nruns: int,
    pre_compute_mapping: bool = True,
    prob_flip_best: float = 0.0,
    seed: int = 0,
    chain_offset: int = 0,
    instance_seeds=None,
):
    """Many instances in one launch (sls_run_batch): ``weights_*`` are lists with one vector per instance.
    Returns a list of :class:`SlsResult`; entry i equals ``run_sls(instances[i], ..., seed=instance_seeds[i])``
    (``seed + i`` when ``instance_seeds`` is None)."""
    instances = list(instances)
    k = len(instances)
    p = _make_params(algo, nsteps, nruns, False, pre_compute_mapping, prob_flip_best, seed, chain_offset)
    same = weights_resample is weights_initialize  # the oracle-guided call: one list for both (no second conversion)
    wi = [_weights(w) for w in weights_initialize]
    wr = wi if same else [_weights(w) for w in weights_resample]
    if len(wi) != k or len(wr) != k:
        raise ValueError("one weight vector per instance is required")
    lens = [max(len(a), len(b)) for a, b in zip(wi, wr)]
    off = np.zeros(k + 1, dtype=np.uint64)
    off[1:] = np.cumsum(lens)
    for i in range(k):
        if len(wi[i]) < instances[i].n or (len(wr[i]) < instances[i].n and algo != "schoening"):
            raise PanicException(4, "weights shorter than var_count")
    if same:
        cat_i = cat_r = np.concatenate(wi) if k and int(off[-1]) else np.zeros(1, dtype=np.float64)
    else:
        cat_i = np.zeros(max(int(off[-1]), 1), dtype=np.float64)
        cat_r = np.zeros(max(int(off[-1]), 1), dtype=np.float64)
        for i in range(k):
            cat_i[int(off[i]): int(off[i]) + len(wi[i])] = wi[i]
            cat_r[int(off[i]): int(off[i]) + len(wr[i])] = wr[i]
    bufs = [_Buffers(inst.n, nruns, nsteps, False) for inst in instances]
    results = (_lib.RunResult * max(k, 1))(*[b.res for b in bufs])
    handles = (C.c_void_p * max(k, 1))(*[inst._h.value for inst in instances])
    seeds = None
    if instance_seeds is not None:
        seeds = np.ascontiguousarray(np.asarray(instance_seeds, dtype=np.uint64))
        if seeds.shape != (k,):
            raise ValueError("instance_seeds needs one seed per instance")
    _check(_lib.load().sls_run_batch(handles, k, cat_i.ctypes.data, cat_r.ctypes.data, off.ctypes.data,
                                     seeds.ctypes.data if seeds is not None else None, C.byref(p), results))
    out = []
    for i, b in enumerate(bufs):
        b.res = results[i]
        out.append(b.to_result(instances[i].n, nruns, False))
    return out


class ResidentSolver:
    """State for ``nruns`` chains kept in HBM between launches (sls_solver_* of the C ABI)."""

    def __init__(self, inst: Instance, algo: str, nsteps: int, nruns: int, return_trajectories=False,
                 pre_compute_mapping=True, prob_flip_best=0.0, seed=0, chain_offset=0):
        self.inst = inst
        self.nsteps, self.nruns, self.want_traj = int(nsteps), int(nruns), bool(return_trajectories)
        self._p = _make_params(algo, nsteps, nruns, return_trajectories, pre_compute_mapping, prob_flip_best, seed,
                               chain_offset)
        self._h = C.c_void_p()
        _check(_lib.load().sls_solver_create(inst._h, C.byref(self._p), C.byref(self._h)))

    @property
    def variant(self) -> str:
        return _lib.load().sls_solver_variant(self._h).decode()

    def set_weights(self, weights_initialize, weights_resample):
        wi, wr = _weights(weights_initialize), _weights(weights_resample)
        _check(_lib.load().sls_solver_set_weights(self._h, wi.ctypes.data, wi.size, wr.ctypes.data, wr.size))

    def set_seed(self, seed: int, chain_offset: int = 0):
        _check(_lib.load().sls_solver_set_seed(self._h, int(seed) & 0xFFFFFFFFFFFFFFFF, int(chain_offset)))

    def launch(self, stream: int | None = None):
        _check(_lib.load().sls_solver_launch(self._h, C.c_void_p(stream or 0)))

    def sync(self) -> float:
        _check(_lib.load().sls_solver_sync(self._h))
        ms = C.c_double(0)
        _lib.load().sls_solver_last_kernel_ms(self._h, C.byref(ms))
        return ms.value

    def launch_count(self) -> int:
        return _lib.load().sls_solver_launch_count(self._h)

    def export_device(self, found=0, steps=0, best_energy=0, flips=0, stream: int | None = None):
        """Device-to-device copy of per-chain results into caller-owned device buffers (raw pointers)."""
        _check(
            _lib.load().sls_solver_export_device(
                self._h, C.c_void_p(found), C.c_void_p(steps), C.c_void_p(best_energy), C.c_void_p(flips),
                C.c_void_p(stream or 0),
            )
        )

    def fetch_summary(self) -> SlsResult:
        """Everything but the trajectories (which stay on the device)."""
        return self.fetch(with_trajectories=False)

    def fetch(self, with_trajectories: bool = True) -> SlsResult:
        buf = _Buffers(self.inst.n, self.nruns, self.nsteps, False)  # trajectories are fetched packed, below
        _check(_lib.load().sls_solver_fetch(self._h, C.byref(buf.res)))
        res = buf.to_result(self.inst.n, self.nruns, False)
        if self.want_traj and with_trajectories:
            off = np.zeros(self.nruns + 1, dtype=np.uint64)
            _check(_lib.load().sls_solver_fetch_trajectories(self._h, None, 0, off.ctypes.data))
            packed = np.zeros(max(int(off[-1]), 1), dtype=np.uint32)
            _check(_lib.load().sls_solver_fetch_trajectories(self._h, packed.ctypes.data, packed.size, off.ctypes.data))
            res.trajectories = [packed[int(off[r]): int(off[r + 1])] for r in range(self.nruns)]
        return res

    def trajectory_stats(self):
        """(mean[nsteps+1], median[nsteps+1]) of the energy over runs per step, trajectories zero-padded after
        their last entry: the NumPy reduction of evaluate_with_given_params.py:18-23, :141-156 done on the GPU
        (divide by the clause count to get the reference's normalised curves)."""
        mean = np.zeros(self.nsteps + 1, dtype=np.float64)
        median = np.zeros(self.nsteps + 1, dtype=np.float64)
        _check(_lib.load().sls_solver_trajectory_stats(self._h, mean.ctypes.data, median.ctypes.data))
        return mean, median

    def close(self):
        if getattr(self, "_h", None):
            _lib.load().sls_solver_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


_seed_override: int | None = None


def set_seed(seed: int | None) -> None:
    """Make run_sls_python reproducible (the reference has no such knob: lib.rs:206 uses OS entropy)."""
    global _seed_override
    _seed_override = seed


def _fresh_seed() -> int:
    global _seed_override
    if _seed_override is not None:
        s = _seed_override
        _seed_override = (s * 6364136223846793005 + 1442695040888963407) & 0xFFFFFFFFFFFFFFFF
        return s
    env = os.environ.get("MOSER_RUST_SEED")
    if env is not None:
        return int(env)
    return int.from_bytes(os.urandom(8), "little")


def run_sls_python(
    algo_type_as_str,
    path,
    weights_initialize,
    weights_resample,
    nsteps,
    nruns,
    return_trajectories,
    pre_compute_mapping,
    prob_flip_best,
):
    """Drop-in for ``moser_rust.run_sls_python`` (rust/src/lib.rs:103-149).

    Same parameter names (train_utils.py:293-303 passes the last three by keyword), same positional
    5-tuple: (best_assignment, found_solutions, best_energy, num_steps, trajectories).  The file is
    parsed on every call like the reference does (lib.rs:115-118).  Trajectories are returned as a
    list of uint32 arrays (one per run, length steps+1): ``len()``, iteration and ``np.pad`` behave
    as with the reference's list of lists, without materialising 10^8 Python ints.
    """
    inst = Instance.from_dimacs(path)
    try:
        res = run_sls(
            inst, algo_type_as_str, weights_initialize, weights_resample, nsteps, nruns, return_trajectories,
            pre_compute_mapping, prob_flip_best, seed=_fresh_seed(),
        )
    finally:
        inst.close()
    return res.as_reference_tuple()
