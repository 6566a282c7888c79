"""The C-ABI library: loads, exports every symbol include/sls_b200.h declares, host-only entry
points agree with the oracle, compute entry points refuse to run without a GPU (no fallback)."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT, golden_cnf
from oracle import sls_oracle as so
from sls_sat_solving_with_deep_learning_b200 import _lib
import sls_sat_solving_with_deep_learning_b200 as eng


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "sls_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(sls_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_table_agree():
    assert declared_symbols() == sorted(_lib.SYMBOLS)


def test_library_exports_every_declared_symbol():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name


def test_library_is_self_contained():
    # the drop-in boundary must not drag torch or the oracle in
    import subprocess

    out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "torch" not in out and "oracle" not in out


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sls_sat_solving_with_deep_learning_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle/" not in src and "sls_oracle" not in src and "import oracle" not in src, f
    assert "oracle" not in open(os.path.join(ROOT, "moser_rust.py")).read()


def test_host_parser_matches_oracle_on_golden(expected):
    for name, info in expected.items():
        gi = eng.Instance.from_dimacs(golden_cnf(name))
        oi = so.Instance.from_dimacs(golden_cnf(name))
        assert (gi.n, gi.m, gi.num_lits) == (info["n"], info["m"], info["num_lits"])
        grow, glits = gi.csr()
        orow, olits = oi.csr()
        assert (grow == orow).all() and (glits == olits).all()


def test_host_parser_rules_match_oracle():
    ok = eng.Instance.from_text("c comment\np cnf 3 2\n1 -2\n 3 0 -1 0\n")
    assert (ok.n, ok.m, ok.num_lits, ok.max_clause_len) == (3, 2, 4, 3)
    assert eng.Instance.from_text("1 2 0\n-3 0\n").n == 3
    assert eng.Instance.from_text("p cnf 5 1\n1 0\n").n == 5
    for bad in ("p cnf 2 2\n1 0\n", "p cnf 1 1\n2 0\n", "p cnf 2 1\n1 2\n", "p cnf 2 1\n1 x 0\n", "p dnf 1 1\n1 0\n"):
        with pytest.raises(eng.PanicException):
            eng.Instance.from_text(bad)
    with pytest.raises(eng.PanicException) as e:  # lib.rs:115 "failed to read"
        eng.Instance.from_dimacs("/nonexistent/file.cnf")
    assert "failed to read" in str(e.value)


def test_from_csr_roundtrip():
    clauses = [(1, -2, 3), (-1, 4), (2,), (-3, -4, 1, 2)]
    gi = eng.Instance.from_clauses(4, clauses)
    row, lits = gi.csr()
    assert row.tolist() == [0, 3, 5, 6, 10]
    assert lits.tolist() == [l for c in clauses for l in c]
    with pytest.raises(eng.PanicException):
        eng.Instance.from_clauses(2, [(1, 3)])


def test_argument_errors_mirror_the_reference():
    with pytest.raises(eng.PanicException) as e:  # lib.rs:125
        eng.run_sls_python("walksat", golden_cnf("ref_tests_medium"), [0.5] * 116, [0.5] * 116, 10, 1, False, True, 0)
    assert "Unknown algorithm type" in str(e.value)
    with pytest.raises(OverflowError):  # pyo3 usize extraction of a negative int
        eng.run_sls_python("moser", golden_cnf("ref_tests_medium"), [0.5] * 116, [0.5] * 116, -1, 1, False, True, 0)
    import moser_rust

    assert moser_rust.run_sls_python is eng.run_sls_python
    import inspect

    assert list(inspect.signature(moser_rust.run_sls_python).parameters) == [
        "algo_type_as_str", "path", "weights_initialize", "weights_resample", "nsteps", "nruns",
        "return_trajectories", "pre_compute_mapping", "prob_flip_best",
    ]  # lib.rs:104-113; train_utils.py:293-303 passes the last three by keyword


@pytest.mark.skipif(eng.device_count() > 0, reason="checks the no-GPU behaviour")
def test_compute_fails_loudly_without_gpu():
    gi = eng.Instance.from_dimacs(golden_cnf("ref_tests_medium"))
    w = np.full(gi.n, 0.5)
    with pytest.raises(eng.PanicException):
        eng.run_sls(gi, "moser", w, w, 10, 2)
    with pytest.raises(eng.PanicException):
        eng.eval_assignments(gi, np.zeros((1, gi.n)))
    from sls_sat_solving_with_deep_learning_b200 import pipeline, solver

    with pytest.raises(eng.PanicException):
        solver.reserve_device_pool(1 << 20)
    with pytest.raises(eng.PanicException):  # the pipeline's first device call (the pool reservation) fails: no silent CPU path
        pipeline.solve_with_oracle(object(), [gi], "moser", 10, 2)
