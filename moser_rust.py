"""Drop-in replacement for the reference's Rust/pyo3 module ``moser_rust`` (Cargo.toml:7-10,
rust/src/lib.rs:151-155).  ``import moser_rust; moser_rust.run_sls_python(...)`` keeps working for
python/src/evaluate_with_given_params.py:10,128-138 and python/src/train_utils.py:5,293-303, and now
runs on the B200 through libsls_b200.so.
"""
from sls_sat_solving_with_deep_learning_b200.solver import PanicException, run_sls_python, set_seed  # noqa: F401

__all__ = ["run_sls_python", "PanicException", "set_seed"]
