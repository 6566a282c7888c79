"""ctypes loader for libsls_b200.so (the C ABI declared in include/sls_b200.h).

The library is built in-tree by ``__graft_entry__.build()`` / ``build_native()`` below and is the
only compute path of the package: there is no Python/NumPy fallback, and a missing library or a
missing GPU is an error, never a silent detour.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
LIB_PATH = os.path.join(_PKG, "lib", "libsls_b200.so")
if os.environ.get("SLS_B200_LIB"):  # development hook: A/B runs of differently built libraries (tools/)
    LIB_PATH = os.path.abspath(os.environ["SLS_B200_LIB"])
_SOURCES = ["csrc/api.cu", "csrc/gnn_api.cu", "csrc/instance.cpp"]
def _headers():
    """Every header the library is built from (staleness check): all of csrc/ plus the public C header."""
    import glob

    hs = sorted(glob.glob(os.path.join(_PKG, "csrc", "*.h")) + glob.glob(os.path.join(_PKG, "csrc", "*.cuh")))
    return [os.path.relpath(h, _PKG) for h in hs] + ["../include/sls_b200.h"]


NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


class NativeLibraryError(RuntimeError):
    """The CUDA library is missing or cannot be loaded."""


def _nvcc() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise NativeLibraryError("nvcc not found: cannot build libsls_b200.so")


def is_stale() -> bool:
    if os.environ.get("SLS_B200_LIB"):
        return False  # an explicitly chosen library is used as it is
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    return any(os.path.getmtime(os.path.join(_PKG, f)) > t for f in _SOURCES + _headers())


def build_native(force: bool = False, verbose: bool = False) -> str:
    """nvcc -gencode arch=compute_100a,code=sm_100a ... -> lib/libsls_b200.so (cross-compiles without a GPU)."""
    if not force and not is_stale():
        return LIB_PATH
    os.makedirs(os.path.dirname(LIB_PATH), exist_ok=True)
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", LIB_PATH, *_SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
    subprocess.check_call(cmd, cwd=_PKG)
    return LIB_PATH


class RunParams(C.Structure):
    _fields_ = [
        ("algo", C.c_int32),
        ("return_trajectories", C.c_int32),
        ("pre_compute_mapping", C.c_int32),
        ("reserved", C.c_int32),
        ("nsteps", C.c_uint64),
        ("nruns", C.c_uint64),
        ("prob_flip_best", C.c_double),
        ("seed", C.c_uint64),
        ("chain_offset", C.c_uint64),
    ]


class RunResult(C.Structure):
    _fields_ = [
        ("best_assignment", C.c_void_p),
        ("found", C.c_void_p),
        ("steps", C.c_void_p),
        ("best_energy_run", C.c_void_p),
        ("flips_run", C.c_void_p),
        ("traj", C.c_void_p),
        ("traj_len", C.c_void_p),
        ("best_energy", C.c_uint64),
        ("has_best_assignment", C.c_int32),
        ("reserved", C.c_int32),
        ("total_steps", C.c_uint64),
        ("total_flips", C.c_uint64),
        ("kernel_ms", C.c_double),
    ]


# every symbol include/sls_b200.h declares: (restype, argtypes)
_vp = C.c_void_p
SYMBOLS = {
    "sls_last_error": (C.c_char_p, []),
    "sls_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "sls_set_device": (C.c_int, [C.c_int]),
    "sls_device_pool_reserve": (C.c_int, [C.c_uint64]),
    "sls_instance_from_dimacs": (C.c_int, [C.c_char_p, C.POINTER(_vp)]),
    "sls_instance_from_dimacs_text": (C.c_int, [C.c_char_p, C.c_size_t, C.POINTER(_vp)]),
    "sls_instance_from_csr": (C.c_int, [C.c_uint32, C.c_uint32, _vp, _vp, C.POINTER(_vp)]),
    "sls_instance_free": (None, [_vp]),
    "sls_instance_num_vars": (C.c_uint32, [_vp]),
    "sls_instance_num_clauses": (C.c_uint32, [_vp]),
    "sls_instance_num_lits": (C.c_uint64, [_vp]),
    "sls_instance_max_clause_len": (C.c_uint32, [_vp]),
    "sls_instance_get_csr": (C.c_int, [_vp, _vp, _vp]),
    "sls_eval_assignments": (C.c_int, [_vp, _vp, C.c_uint64, _vp, _vp]),
    "sls_eval_last_kernel_ms": (C.c_double, []),
    "sls_eval_last_eval_kernel_ms": (C.c_double, []),
    "sls_constraint_graph": (C.c_int, [C.c_uint32, C.c_uint32, _vp, _vp, C.c_uint64, _vp, _vp, C.c_uint64, _vp]),
    "sls_run": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t, C.POINTER(RunParams), C.POINTER(RunResult)]),
    "sls_run_batch": (C.c_int, [C.POINTER(_vp), C.c_uint32, _vp, _vp, _vp, _vp, C.POINTER(RunParams), C.POINTER(RunResult)]),
    "sls_solver_create": (C.c_int, [_vp, C.POINTER(RunParams), C.POINTER(_vp)]),
    "sls_solver_free": (None, [_vp]),
    "sls_solver_set_weights": (C.c_int, [_vp, _vp, C.c_size_t, _vp, C.c_size_t]),
    "sls_solver_set_seed": (C.c_int, [_vp, C.c_uint64, C.c_uint64]),
    "sls_solver_launch": (C.c_int, [_vp, _vp]),
    "sls_solver_sync": (C.c_int, [_vp]),
    "sls_solver_last_kernel_ms": (C.c_int, [_vp, C.POINTER(C.c_double)]),
    "sls_solver_fetch": (C.c_int, [_vp, C.POINTER(RunResult)]),
    "sls_solver_fetch_trajectories": (C.c_int, [_vp, _vp, C.c_uint64, _vp]),
    "sls_solver_trajectory_stats": (C.c_int, [_vp, _vp, _vp]),
    "sls_solver_export_device": (C.c_int, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "sls_solver_launch_count": (C.c_int, [_vp]),
    "sls_solver_variant": (C.c_char_p, [_vp]),
    "sls_gnn_model_create": (C.c_int, [_vp, C.c_uint64, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.POINTER(_vp)]),
    "sls_gnn_model_free": (None, [_vp]),
    "sls_gnn_forward_graph": (C.c_int, [_vp, C.c_uint64, C.c_uint64, _vp, _vp, _vp, _vp, _vp, C.POINTER(C.c_double)]),
    "sls_gnn_forward_lcg": (C.c_int, [_vp, C.POINTER(_vp), C.c_uint32, _vp, _vp, C.POINTER(C.c_double)]),
    "sls_gnn_launch_count": (C.c_int, [_vp]),
    "sls_gnn_set_profile": (C.c_int, [_vp, C.c_int]),
    "sls_gnn_get_profile": (C.c_int, [_vp, C.POINTER(C.c_double)]),
}

_lib = None


def load():
    """Load libsls_b200.so; raises NativeLibraryError when it is missing (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise NativeLibraryError(
            f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). This package has no CPU fallback."
        )
    try:
        lib = C.CDLL(LIB_PATH)
    except OSError as e:  # pragma: no cover
        raise NativeLibraryError(f"cannot load {LIB_PATH}: {e}") from e
    for name, (res, args) in SYMBOLS.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise NativeLibraryError(f"{LIB_PATH} does not export {name}") from e
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
