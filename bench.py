#!/usr/bin/env python
"""bench.py -- headline benchmark of the SLS hot path (BASELINE.json, configs[1]).

Workload "C2": probSAT ("WalkSAT" in the reference README) with the uniform oracle (w = 1/2) on one
synthetic random 3-SAT instance n=10,000, alpha=4.2 (m=42,000), 4096 chains per GPU, 10^5 steps per
chain, prob_flip_best=0 (the reference default).  One bench "step" = one solve of all chains.
Metric: variable flips / s (whole job, all GPUs).

  python bench.py [--gpus N] [--steps K] [--warmup W]           own arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                        reference CPU arm (oracle port on host cores)

Under torchrun one process drives one GPU; chains are sharded by global chain id (weak scaling:
4096 chains per GPU), NCCL gathers the per-chain solved flags and step counts.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_VARS = 10_000
ALPHA = 4.2
N_CLAUSES = int(round(N_VARS * ALPHA))
CHAINS_PER_GPU = 4096
NSTEPS = 100_000
INSTANCE_SEED = 20260102
ALGO = "probsat"
# one variable flip, reference-faithful re-evaluation of every clause of the
# flipped variable: 4 (pick) + 8k (literals + weights) + d*(4 + 4k) bytes, k = 3, d = k*alpha
K = 3
D_OCC = K * ALPHA
BYTES_PER_FLIP = 4 + 8 * K + D_OCC * (4 + 4 * K)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.rows = []
        self._stop = threading.Event()
        self._t = None

    def _loop(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(
                    ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                    capture_output=True, text=True, timeout=5,
                ).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._loop, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i] == "Active"})
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": max(mx) if mx else None,
            "reasons": reasons,
            "samples": len(sm),
        }


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(path, nruns, nsteps, seed, nthreads):
    """One run_sls_python call of the reference CPU path, restated: parse the file (lib.rs:115-118)
    and run all chains on the host cores (lib.rs:201).  Uses oracle/ as the timed baseline only."""
    from oracle import sls_oracle

    t0 = time.perf_counter()
    oi = sls_oracle.Instance.from_dimacs(path)
    w = np.full(oi.n, 0.5)
    r = oi.run_sls(ALGO, w, w, nsteps, nruns, False, True, 0.0, seed=seed, nthreads=nthreads)
    return r.total_flips, time.perf_counter() - t0


def workload_config(extra=None):
    cfg = {
        "workload": "C2: probSAT (WalkSAT), uniform oracle w=1/2, synthetic random 3-SAT n=10000 alpha=4.2 "
                    "(m=42000), prob_flip_best=0",
        "chains_per_gpu": CHAINS_PER_GPU,
        "nsteps_per_chain": NSTEPS,
        "instance_seed": INSTANCE_SEED,
        "algo": ALGO,
    }
    if extra:
        cfg.update(extra)
    return cfg


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores."""
    if rank != 0:
        return
    from oracle import sls_oracle  # noqa: F401  (bench.py may execute oracle/ here only)
    from sls_sat_solving_with_deep_learning_b200 import synth

    cores = host_cores()
    row, lits = synth.random_ksat_csr(N_VARS, N_CLAUSES, K, INSTANCE_SEED)
    tmp = tempfile.mkdtemp(prefix="slsb200_")
    path = os.path.join(tmp, "c2.cnf")
    synth.write_dimacs(path, N_VARS, row, lits)
    # bounded sample of the workload: same instance, same steps per chain, fewer chains
    nruns = min(CHAINS_PER_GPU, cores * 16)
    for w in range(args.warmup):
        cpu_reference_run(path, max(cores, 1), NSTEPS // 10, 1000 + w, cores)
    flips = 0
    secs = 0.0
    for s in range(args.steps):
        f, dt = cpu_reference_run(path, nruns, NSTEPS, 2000 + s, cores)
        flips += f
        secs += dt
    value = flips / secs
    sample = f"{nruns} of {CHAINS_PER_GPU} chains x {NSTEPS} steps per step, {cores} threads, parse included"
    line = {
        "impl": "reference",
        "metric": "variable flips/sec", "value": value, "unit": "flips/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": secs / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u32 bit-ops + f64 probabilities", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": value, "unit": "flips/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "flips/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_own(args, rank, world, local_rank):
    import torch

    import sls_sat_solving_with_deep_learning_b200 as eng
    from sls_sat_solving_with_deep_learning_b200 import synth

    if not torch.cuda.is_available() or eng.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    eng.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    row, lits = synth.random_ksat_csr(N_VARS, N_CLAUSES, K, INSTANCE_SEED)
    inst = eng.Instance.from_csr(N_VARS, row, lits)
    w = np.full(N_VARS, 0.5)
    R = CHAINS_PER_GPU
    solver = eng.ResidentSolver(inst, ALGO, NSTEPS, R, False, True, 0.0, seed=1, chain_offset=rank * R)
    solver.set_weights(w, w)  # inputs resident in HBM before the timed region

    stream = torch.cuda.current_stream()
    d_found = torch.zeros(R, dtype=torch.uint8, device="cuda")
    d_steps = torch.zeros(R, dtype=torch.int64, device="cuda")
    d_flips = torch.zeros(R, dtype=torch.int64, device="cuda")
    g_found = torch.zeros(world * R, dtype=torch.uint8, device="cuda") if dist else None
    g_steps = torch.zeros(world * R, dtype=torch.int64, device="cuda") if dist else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def one_step(seed):
        solver.set_seed(seed, rank * R)
        solver.launch(stream.cuda_stream)
        solver.export_device(found=d_found.data_ptr(), steps=d_steps.data_ptr(), flips=d_flips.data_ptr(),
                             stream=stream.cuda_stream)
        if dist:  # the only collective of the path: per-chain solved flags and step counts
            dist.all_gather_into_tensor(g_found, d_found)
            dist.all_gather_into_tensor(g_steps, d_steps)

    for i in range(args.warmup):
        one_step(100 + i)
    torch.cuda.synchronize()

    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    total_flips = torch.zeros((), dtype=torch.int64, device="cuda")
    solved = 0
    kernel_ms = []
    if dist:
        dist.barrier()
    torch.cuda.synchronize()
    with ClockSampler(local_rank) as clocks:
        for i in range(args.steps):
            flush.fill_(i & 0xFF)  # L2 flush between timed iterations (outside the event pair)
            ev[i][0].record(stream)
            one_step(1000 + i)
            ev[i][1].record(stream)
            total_flips += d_flips.sum()
            torch.cuda.synchronize()
            kernel_ms.append(solver.sync())
            solved += int(d_found.sum().item())
        if dist:
            dist.barrier()
        torch.cuda.synchronize()
    ms = sum(a.elapsed_time(b) for a, b in ev)
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if dist:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(total_flips, op=dist.ReduceOp.SUM)
    ms = float(t.item())
    flips = int(total_flips.item())
    value = flips / (ms * 1e-3)

    # ---- e2e: the reference-facing call (parse + H2D + solve + D2H), host buffers, wall clock
    tmp = tempfile.mkdtemp(prefix="slsb200_")
    path = os.path.join(tmp, "c2.cnf")
    synth.write_dimacs(path, N_VARS, row, lits)
    import moser_rust

    e2e_steps = max(1, min(args.steps, 3))
    moser_rust.set_seed(5000 + rank)
    moser_rust.run_sls_python(ALGO, path, w, w, NSTEPS // 10, R, False, True, 0)  # warm
    if dist:
        dist.barrier()
    t0 = time.perf_counter()
    e2e_flips = 0
    for i in range(e2e_steps):
        _, found, _, steps, _ = moser_rust.run_sls_python(ALGO, path, w, w, NSTEPS, R, False, True, 0)
        e2e_flips += int(np.sum(steps))  # probSAT: one flip per step
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    tf = torch.tensor([e2e_flips], dtype=torch.int64, device="cuda")
    if dist:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
        dist.all_reduce(tf, op=dist.ReduceOp.SUM)
    e2e_value = float(tf.item()) / float(te.item())
    L = len(lits)
    h2d = (N_CLAUSES + 1) * 4 + L * 4 + (N_VARS + 1) * 4 + L * 4 + N_CLAUSES * 16 + (L + 32) * 16 + N_VARS * (8 + 8 + 16)
    d2h = R * (1 + 1 + 8 + 8 + 4) + 4 + (N_VARS + 31) // 32 * 4
    moser_rust.set_seed(None)

    line = None
    if rank == 0:
        peak, peak_src = peaks()
        kms = float(np.mean(kernel_ms))
        achieved = R * NSTEPS * BYTES_PER_FLIP / (kms * 1e-3) / 1e9
        line = {
            "metric": "variable flips/sec", "value": value, "unit": "flips/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 bit-ops + f64 probabilities", "data": "synthetic",
            "config": workload_config({
                "parallelism": f"chains sharded over {world} GPU(s), {R} per GPU",
                "kernel": solver.variant,
                "l2": "flushed between timed steps (256 MB write); the 4.7 MB instance is L2-resident by nature",
                "instances_solved_per_s": solved / (ms * 1e-3),
            }),
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "flips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "includes": "DIMACS parse + CSR build + H2D + solve + D2H (run_sls_python)"},
            "gpu_launches": int(args.steps * solver.launch_count()),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one walk_k3s launch (ncu --set full,
                         # profiles/r01_final_walk_k3s_summary.md): 5.51 MB + 0.86 MB, independent of the step count
                         # because instance and chain state never leave L2 / shared memory
                         "traffic": 6.37e6, "peak_source": peak_src,
                         "bytes_per_flip": BYTES_PER_FLIP, "kernel_ms": kms,
                         "note": "latency-bound walk: state in shared memory, instance in L2; see DESIGN.md",
                         # what the kernel is actually served from: L2 -> SM sectors per flip (ncu lts__t_sectors_srcunit_tex_op_read
                         # of the same launch: 300 B per flip) against the L2 rates measured on this pool with tools/ubench.cu
                         # (profiles/r01_ubench_l2_latency.md), and the measured step latency of a chain alone on an SM
                         "l2": {"achieved_gbs": R * NSTEPS * 300.0 / (kms * 1e-3) / 1e9, "random_sector_peak_gbs": 8193.7,
                                "stream_peak_gbs": 17526.5, "lone_chain_ns_per_step": 920.0,
                                "loaded_ns_per_step": kms * 1e6 / NSTEPS}},
        }
    if dist:
        dist.barrier()

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on the host cores
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = host_cores()
        nruns = min(CHAINS_PER_GPU, cores * 32)
        f, dt = cpu_reference_run(path, nruns, NSTEPS, 77, cores)
        line["cpu_baseline"] = {
            "value": f / dt, "unit": "flips/s", "cores": cores, "kind": "port",
            "sample": f"{nruns} of {CHAINS_PER_GPU} chains x {NSTEPS} steps, {cores} threads, parse included",
        }
    # ---- second leg of the north star, reported beside the headline (rank 0, N=1 only; never part of `value`): the batched
    # GNN oracle forward on BASELINE config 4's shape, random-init weights, device time by CUDA events inside the library
    if rank == 0 and world == 1 and not args.no_oracle_forward:
        try:
            line["oracle_forward"] = oracle_forward_leg()
        except Exception as exc:  # the headline line must survive a failure of the secondary leg
            line["oracle_forward"] = {"error": f"{type(exc).__name__}: {exc}"}
    if rank == 0:
        print(json.dumps(line), flush=True)
    if dist:
        dist.destroy_process_group()


def oracle_forward_leg(n_inst: int = 256, reps: int = 3):
    import sls_sat_solving_with_deep_learning_b200 as eng
    from sls_sat_solving_with_deep_learning_b200 import gnn, synth as _synth

    n, m = 500, 2100
    insts = []
    for i in range(n_inst):
        row, lits = _synth.random_ksat_csr(n, m, 3, 20260104 + i)
        insts.append(eng.Instance.from_csr(n, row, lits))
    net = gnn.InteractionNetworkLCG(_synth.random_oracle_params(seed=42))
    net.probabilities(insts)  # warm: uploads the instances, allocates the activation workspace
    best_ms, best_wall = 1e30, 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        net.probabilities(insts)
        best_wall = min(best_wall, time.perf_counter() - t0)
        best_ms = min(best_ms, net.last_kernel_ms)
    n_nodes, n_edges = 2 * n + m, 3 * m + n
    flops = n_inst * (1.3984e6 * n_edges + 1.5328e6 * n_nodes)  # dense FLOPs as the reference computes them
    return {
        "workload": f"C4 shape: {n_inst} synthetic 3-SAT instances n={n} m={m} (LCG: {n_nodes} nodes, {n_edges} edges each), random-init interaction network",
        "instances_per_s": n_inst / (best_ms * 1e-3), "device_ms": best_ms,
        "e2e_instances_per_s": n_inst / best_wall, "precision": os.environ.get("SLS_B200_GNN_PRECISION", "bf16x2"),
        "dense_tflops": flops / (best_ms * 1e-3) / 1e12, "gpu_launches": int(net.launch_count()),
        "kernel": "gnn::mlp_ln_persistent_ts_kernel (tcgen05, A operand in TMEM) + segment_sum_kernel",
    }


def main():
    global NSTEPS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-oracle-forward", action="store_true", help="skip the secondary GNN oracle-forward measurement")
    ap.add_argument("--nsteps-per-chain", type=int, default=NSTEPS,
                    help="profiling only: shorten the chains (the default is the BASELINE workload)")
    args = ap.parse_args()
    NSTEPS = args.nsteps_per_chain
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_own(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
