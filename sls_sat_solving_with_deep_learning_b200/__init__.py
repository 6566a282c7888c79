"""B200-native engine for the oracle-guided SLS hot path (Moser-Tardos / WalkSAT chains + GNN oracle).

Public surface (mirrors the reference's interface for this path):
  run_sls_python          drop-in for moser_rust.run_sls_python (rust/src/lib.rs:103-149)
  run_sls, ResidentSolver the same solver on a parsed instance / with HBM-resident chain state
  run_sls_batch           many instances x nruns chains in one launch
  Instance                DIMACS / CSR -> clause + occurrence CSR in HBM
  get_violated_constraints, eval_assignments   clause evaluation (sat_representations.py:718-741)
  candidate_energies      the training loader's candidate energies (data_utils.py:123-129)
"""
from .solver import (  # noqa: F401
    ALGOS,
    Instance,
    PanicException,
    ResidentSolver,
    SlsResult,
    candidate_energies,
    device_count,
    eval_assignments,
    eval_last_eval_kernel_ms,
    eval_last_kernel_ms,
    get_violated_constraints,
    run_sls,
    run_sls_batch,
    run_sls_python,
    set_device,
    set_seed,
)
