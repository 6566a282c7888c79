"""ctypes binding of oracle/sls_oracle.c -- TEST INFRASTRUCTURE ONLY.

The oracle is the CPU restatement of the reference solver (rust/src/lib.rs) used
by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs as the checker / timed CPU baseline.  The product package never imports it.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle_sls.so")

ALGOS = {"moser": 0, "moser_single": 1, "schoening": 2, "probsat": 3}


def build(force: bool = False) -> str:
    """Compile the C restatement (gcc only; no reference sources involved)."""
    src = os.path.join(_HERE, "sls_oracle.c")
    hdr = os.path.join(_HERE, "sls_oracle.h")
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    )
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-B"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Params(C.Structure):
    _fields_ = [
        ("algo", C.c_int),
        ("nsteps", C.c_uint64),
        ("nruns", C.c_uint64),
        ("return_trajectories", C.c_int),
        ("pre_compute_mapping", C.c_int),
        ("prob_flip_best", C.c_double),
        ("seed", C.c_uint64),
        ("chain_offset", C.c_uint64),
        ("nthreads", C.c_int),
    ]


class _Result(C.Structure):
    _fields_ = [
        ("best_assignment", C.c_void_p),
        ("found", C.c_void_p),
        ("steps", C.c_void_p),
        ("best_energy_run", C.c_void_p),
        ("flips_run", C.c_void_p),
        ("traj", C.c_void_p),
        ("traj_len", C.c_void_p),
        ("best_energy", C.c_uint64),
        ("has_best_assignment", C.c_int),
        ("total_steps", C.c_uint64),
        ("total_flips", C.c_uint64),
    ]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        l = C.CDLL(_LIB_PATH)
        l.orc_instance_from_dimacs.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
        l.orc_instance_from_dimacs_text.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p)]
        l.orc_instance_from_csr.argtypes = [C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p, C.POINTER(C.c_void_p)]
        l.orc_instance_free.argtypes = [C.c_void_p]
        l.orc_instance_free.restype = None
        for f in ("orc_num_vars", "orc_num_clauses", "orc_num_distinct_clauses"):
            getattr(l, f).argtypes = [C.c_void_p]
            getattr(l, f).restype = C.c_uint32
        l.orc_num_lits.argtypes = [C.c_void_p]
        l.orc_num_lits.restype = C.c_uint64
        l.orc_get_csr.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_get_csr.restype = None
        l.orc_clause_is_violated.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        l.orc_eval_clauses.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_eval_clauses.restype = C.c_uint32
        l.orc_energy_distinct.argtypes = [C.c_void_p, C.c_void_p]
        l.orc_energy_distinct.restype = C.c_uint32
        l.orc_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        l.orc_philox4x32_10.restype = None
        l.orc_init_assignment.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_void_p]
        l.orc_run_sls.argtypes = [
            C.c_void_p,
            C.c_void_p,
            C.c_size_t,
            C.c_void_p,
            C.c_size_t,
            C.POINTER(_Params),
            C.POINTER(_Result),
        ]
        l.orc_last_error.restype = C.c_char_p
        _lib = l
    return _lib


class OracleError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"oracle error {code}: {msg}")
        self.code = code


def _check(rc: int):
    if rc != 0:
        raise OracleError(rc, lib().orc_last_error().decode())


def philox4x32_10(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(c.ctypes.data, k.ctypes.data, out.ctypes.data)
    return out


@dataclass
class RunResult:
    best_assignment: np.ndarray | None  # uint8[n] or None (the reference's empty Vec)
    found: np.ndarray  # bool[R]
    best_energy: int
    steps: np.ndarray  # uint64[R]
    best_energy_run: np.ndarray  # uint64[R]
    flips_run: np.ndarray  # uint64[R]
    trajectories: list  # list of uint32 arrays (len steps+1) or []
    total_steps: int
    total_flips: int

    def as_reference_tuple(self):
        """The positional 5-tuple of rust/src/lib.rs:141-147."""
        ba = [] if self.best_assignment is None else [bool(x) for x in self.best_assignment]
        return (
            ba,
            [bool(x) for x in self.found],
            int(self.best_energy),
            [int(x) for x in self.steps],
            [t.tolist() for t in self.trajectories],
        )


class Instance:
    def __init__(self, handle):
        self._h = handle

    @classmethod
    def from_dimacs(cls, path: str) -> "Instance":
        h = C.c_void_p()
        _check(lib().orc_instance_from_dimacs(os.fsencode(path), C.byref(h)))
        return cls(h)

    @classmethod
    def from_text(cls, text: str) -> "Instance":
        b = text.encode()
        h = C.c_void_p()
        _check(lib().orc_instance_from_dimacs_text(b, len(b), C.byref(h)))
        return cls(h)

    @classmethod
    def from_clauses(cls, n: int, clauses) -> "Instance":
        row = np.zeros(len(clauses) + 1, dtype=np.uint64)
        row[1:] = np.cumsum([len(c) for c in clauses])
        flat = np.asarray([l for c in clauses for l in c], dtype=np.int32)
        h = C.c_void_p()
        _check(lib().orc_instance_from_csr(n, len(clauses), row.ctypes.data, flat.ctypes.data, C.byref(h)))
        return cls(h)

    @classmethod
    def from_csr(cls, n: int, row_ptr, signed_lits) -> "Instance":
        row = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        flat = np.ascontiguousarray(signed_lits, dtype=np.int32)
        h = C.c_void_p()
        _check(lib().orc_instance_from_csr(n, len(row) - 1, row.ctypes.data, flat.ctypes.data, C.byref(h)))
        return cls(h)

    def __del__(self):
        if getattr(self, "_h", None) and _lib is not None:
            _lib.orc_instance_free(self._h)
            self._h = None

    @property
    def n(self) -> int:
        return lib().orc_num_vars(self._h)

    @property
    def m(self) -> int:
        return lib().orc_num_clauses(self._h)

    @property
    def num_lits(self) -> int:
        return lib().orc_num_lits(self._h)

    @property
    def m_distinct(self) -> int:
        return lib().orc_num_distinct_clauses(self._h)

    def csr(self):
        row = np.zeros(self.m + 1, dtype=np.uint64)
        lits = np.zeros(max(self.num_lits, 1), dtype=np.int32)
        lib().orc_get_csr(self._h, row.ctypes.data, lits.ctypes.data)
        return row, lits[: self.num_lits]

    def eval_clauses(self, assignment) -> np.ndarray:
        """sat_representations.py:718-741: 0/1 per clause."""
        a = np.ascontiguousarray(np.asarray(assignment).astype(bool).astype(np.uint8))
        assert a.shape == (self.n,)
        out = np.zeros(max(self.m, 1), dtype=np.uint8)
        lib().orc_eval_clauses(self._h, a.ctypes.data, out.ctypes.data)
        return out[: self.m]

    def energy(self, assignment) -> int:
        a = np.ascontiguousarray(np.asarray(assignment).astype(bool).astype(np.uint8))
        return int(lib().orc_eval_clauses(self._h, a.ctypes.data, None))

    def energy_distinct(self, assignment) -> int:
        a = np.ascontiguousarray(np.asarray(assignment).astype(bool).astype(np.uint8))
        return int(lib().orc_energy_distinct(self._h, a.ctypes.data))

    def init_assignment(self, w_init, seed: int, chain: int) -> np.ndarray:
        w = np.ascontiguousarray(w_init, dtype=np.float64)
        out = np.zeros(max(self.n, 1), dtype=np.uint8)
        _check(lib().orc_init_assignment(self._h, w.ctypes.data, seed, chain, out.ctypes.data))
        return out[: self.n]

    def run_sls(
        self,
        algo: str,
        w_init,
        w_resample,
        nsteps: int,
        nruns: int,
        return_trajectories: bool = False,
        pre_compute_mapping: bool = True,
        prob_flip_best: float = 0.0,
        seed: int = 0,
        chain_offset: int = 0,
        nthreads: int = 0,
    ) -> RunResult:
        if algo not in ALGOS:
            raise OracleError(3, "Unknown algorithm type")
        wi = np.ascontiguousarray(w_init, dtype=np.float64).ravel()
        wr = np.ascontiguousarray(w_resample, dtype=np.float64).ravel()
        n = self.n
        best = np.zeros(max(n, 1), dtype=np.uint8)
        found = np.zeros(max(nruns, 1), dtype=np.uint8)
        steps = np.zeros(max(nruns, 1), dtype=np.uint64)
        ber = np.zeros(max(nruns, 1), dtype=np.uint64)
        flips = np.zeros(max(nruns, 1), dtype=np.uint64)
        traj = traj_len = None
        if return_trajectories:
            traj = np.zeros((max(nruns, 1), nsteps + 1), dtype=np.uint32)
            traj_len = np.zeros(max(nruns, 1), dtype=np.uint64)
        p = _Params(
            ALGOS[algo], nsteps, nruns, int(return_trajectories), int(pre_compute_mapping), float(prob_flip_best),
            seed, chain_offset, nthreads,
        )
        r = _Result(
            best.ctypes.data, found.ctypes.data, steps.ctypes.data, ber.ctypes.data, flips.ctypes.data,
            traj.ctypes.data if traj is not None else None,
            traj_len.ctypes.data if traj_len is not None else None,
            0, 0, 0, 0,
        )
        _check(lib().orc_run_sls(self._h, wi.ctypes.data, wi.size, wr.ctypes.data, wr.size, C.byref(p), C.byref(r)))
        trajs = []
        if return_trajectories:
            trajs = [traj[i, : int(traj_len[i])] for i in range(nruns)]
        return RunResult(
            best_assignment=best[:n].copy() if r.has_best_assignment else None,
            found=found[:nruns].astype(bool),
            best_energy=int(r.best_energy),
            steps=steps[:nruns],
            best_energy_run=ber[:nruns],
            flips_run=flips[:nruns],
            trajectories=trajs,
            total_steps=int(r.total_steps),
            total_flips=int(r.total_flips),
        )
