#!/usr/bin/env python
"""bench.py -- benchmark of the SLS hot path (BASELINE.json: "variable flips/sec & instances solved/sec").

Headline (`value`): workload "C2" = BASELINE configs[1]: probSAT ("WalkSAT" in the reference README) with the uniform
oracle (w = 1/2) on one synthetic random 3-SAT instance n=10,000, alpha=4.2 (m=42,000), 4096 chains per GPU, 10^5 steps
per chain, prob_flip_best=0 (the reference default).  One bench "step" = one solve of all chains.  Metric: variable
flips / s (whole job, all GPUs).

Beside it, in the same JSON line under "legs" (each with its own roofline / cpu_baseline / e2e; a failing leg fails the run):
  c2_moser      the same instance and chain count with Moser-Tardos resampling (rust/src/lib.rs:20-49)
  c1_batch      BASELINE configs[0] shape: 2052 synthetic random 3-SAT instances n 100-300, alpha 1.0-4.5, 100 chains each,
                through run_sls_batch: flips/s AND instances solved/s, with the CPU port on a sample of the same instances
                and the matched-solve-rate statistics (two-proportion z per instance, KS on step counts)
  c4_pipeline   BASELINE configs[3]: batched GNN oracle forward on 10,000 synthetic instances (n=500, alpha=4.2) followed
                by oracle-guided Moser-Tardos, sharded over the GPUs of the run; the oracle forward carries a tensor and an
                HBM roofline and a torch-CPU fp32 baseline
  c5            BASELINE configs[4]: n=10^6, alpha=4.0, 65,536 chains per GPU, chain state in HBM
  clause_eval   the bit-sliced clause evaluation against the L2 random-sector rate measured in this run
"ubench" holds the L2 / HBM ceilings and latencies measured on this box in this run (tools/ubench.cu).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--legs a,b,...|none]     own arm (CUDA, sm_100a)
  python bench.py --impl reference [...]                                         reference CPU arm (oracle port on host cores)

Under torchrun one process drives one GPU; C2 / C5 chains are sharded by global chain id (weak scaling: 4096 / 65,536
chains per GPU), the C1 / C4 batches are split by instance (strong scaling: fixed total work); NCCL gathers the per-chain
solved flags and step counts, and rank 0 re-computes chains of another rank's block to check the gathered tables.
"""
from __future__ import annotations

import argparse
import ctypes
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

# ---- C2 (headline) ---------------------------------------------------------------------------------------------------
N_VARS = 10_000
ALPHA = 4.2
N_CLAUSES = int(round(N_VARS * ALPHA))
CHAINS_PER_GPU = 4096
NSTEPS = 100_000
INSTANCE_SEED = 20260102
ALGO = "probsat"
K = 3
# SURVEY section 8(d), reference-faithful formulation (every clause of a flipped variable is re-evaluated):
#   one step that flips f variables moves 4 (pick) + 8k (literals + weights) + f * d * (4 + 4k) bytes, d = occurrences per variable


def bytes_per_step(k, d, flips_per_step):
    return 4 + 8 * k + flips_per_step * d * (4 + 4 * k)


BYTES_PER_FLIP = bytes_per_step(K, K * ALPHA, 1.0)  # probSAT: one flip per step -> 229.6 B
# ---- C1 / C4 / C5 shapes ("Config shorthands") ------------------------------------------------------
C1_INSTANCES, C1_RUNS, C1_NSTEPS, C1_CPU_STRIDE = 2052, 100, 100_000, 16
C4_INSTANCES, C4_N, C4_M, C4_RUNS, C4_NSTEPS, C4_SEED = 10_000, 500, 2100, 100, 10_000 - 1, 20260104
C4_CHUNK = int(os.environ.get("SLS_BENCH_C4_CHUNK", "1000"))  # instances per pipeline chunk (pipeline.solve_with_oracle)
C4_HANDLES = int(os.environ.get("SLS_BENCH_C4_HANDLES", "2"))  # model handles = forward worker threads of the pipeline
C5_N, C5_M, C5_CHAINS, C5_NSTEPS, C5_SEED = 1_000_000, 4_000_000, 65_536, 1000, 20260105
ALL_LEGS = ("c2_moser", "c1_batch", "c4_pipeline", "c5", "clause_eval")


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        j = json.load(open(p))
        return {"hbm_gbs": float(j["hbm_gbs"]), "bf16_tflops_sustained": float(j.get("bf16_tflops_sustained", 1399.4)),
                "source": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


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


def workload_config():
    """The `config` object: identical in the own arm and the reference arm (run-specific facts live under `run`)."""
    return {
        "workload": "C2: probSAT (WalkSAT), uniform oracle w=1/2, synthetic random 3-SAT n=10000 alpha=4.2 "
                    "(m=42000), prob_flip_best=0",
        "chains_per_gpu": CHAINS_PER_GPU,
        "nsteps_per_chain": NSTEPS,
        "instance_seed": INSTANCE_SEED,
        "algo": ALGO,
    }


# ======================================================================================================================
# workloads (synthetic, seeded; host NumPy)
# ======================================================================================================================
def c1_shapes():
    """(n, m, seed) of the 2052 synthetic C1-shaped instances (n in {100, 200, 300}, alpha spread over 1.0 .. 4.5)."""
    out = []
    for i in range(C1_INSTANCES):
        n = 100 + (i % 3) * 100
        alpha = 1.0 + 3.5 * ((i * 7919) % C1_INSTANCES) / (C1_INSTANCES - 1)
        out.append((n, int(n * alpha), 500_000 + i))
    return out


def csr_of(shape):
    from sls_sat_solving_with_deep_learning_b200 import synth

    n, m, seed = shape
    return synth.random_ksat_csr(n, m, 3, seed)


# ======================================================================================================================
# CPU side: the oracle port as the timed baseline (bench.py may execute oracle/ here and in --impl reference only)
# ======================================================================================================================
def cpu_c2(path, algo, nruns, nsteps, seed, nthreads):
    """One run_sls_python call of the reference CPU path, restated: parse the file (lib.rs:115-118) and run all chains
    on the host cores (lib.rs:201)."""
    from oracle import sls_oracle

    t0 = time.perf_counter()
    oi = sls_oracle.Instance.from_dimacs(path)
    w = np.full(oi.n, 0.5)
    r = oi.run_sls(algo, w, w, nsteps, nruns, False, True, 0.0, seed=seed, nthreads=nthreads)
    return r, time.perf_counter() - t0


def cpu_batch(shapes, csrs, algo, nsteps, nruns, seed0, nthreads, weights=None):
    """The reference's per-file loop (evaluate_with_given_params.py:108-138) on the CPU port: instances one after the
    other, the runs of one instance on all host threads."""
    from oracle import sls_oracle

    t0 = time.perf_counter()
    out = []
    for i, (shape, (row, lits)) in enumerate(zip(shapes, csrs)):
        oi = sls_oracle.Instance.from_csr(shape[0], row, lits)
        w = np.full(shape[0], 0.5) if weights is None else np.asarray(weights[i], dtype=np.float64)
        out.append(oi.run_sls(algo, w, w, nsteps, nruns, False, True, 0.0, seed=seed0 + i, nthreads=nthreads))
    return out, time.perf_counter() - t0


def cpu_oracle_forward(params, csrs, n_vars, threads):
    """torch-CPU fp32 restatement of model.py:127-144 (oracle/gnn_torch.py), one instance at a time like the reference's
    evaluation loop (evaluate_with_given_params.py:116)."""
    import torch

    from oracle import gnn_torch

    torch.set_num_threads(max(threads, 1))
    gnn_torch.probabilities(params, n_vars, *csrs[0])  # warm (thread pool, allocator)
    t0 = time.perf_counter()
    for row, lits in csrs:
        gnn_torch.probabilities(params, n_vars, row, lits)
    return len(csrs) / (time.perf_counter() - t0)


# ======================================================================================================================
# in-run microbenchmarks (tools/ubench.cu): the L2 / HBM ceilings the rooflines divide by are measured on THIS box
# ======================================================================================================================
def run_ubench(with_hbm=True):
    so = os.path.join(ROOT, "tools", "_build", "libubench.so")
    if not os.path.exists(so):
        import __graft_entry__ as g

        g.build_ubench()
    lib = ctypes.CDLL(so)
    lib.slsub_measure.restype = ctypes.c_int
    lib.slsub_measure.argtypes = [ctypes.c_char_p, ctypes.c_size_t, ctypes.c_int]
    buf = ctypes.create_string_buffer(1 << 14)
    rc = lib.slsub_measure(buf, len(buf), int(with_hbm))
    if rc != 0:
        raise RuntimeError(f"ubench failed (CUDA error {rc})")
    return json.loads(buf.value.decode())


# ======================================================================================================================
# legs
# ======================================================================================================================
def two_proportion_z(k1, n1, k2, n2):
    p = (k1 + k2) / (n1 + n2)
    if p <= 0.0 or p >= 1.0:
        return 0.0
    return ((k1 / n1) - (k2 / n2)) / np.sqrt(p * (1 - p) * (1 / n1 + 1 / n2))


def solve_rate_parity(gpu_found, gpu_steps, cpu_found, cpu_steps, nsteps):
    """SURVEY 8(d) "statistical parity": same instances, same weights, independent seeds.  Per instance a two-proportion z on
    the solved fraction and a two-sample KS on the step counts (unsolved chains count as nsteps); Bonferroni over the
    instances that can differ at all."""
    from scipy import stats

    zs, ks_p = [], []
    for gf, gs, cf, cs in zip(gpu_found, gpu_steps, cpu_found, cpu_steps):
        zs.append(two_proportion_z(int(gf.sum()), len(gf), int(cf.sum()), len(cf)))
        a, b = np.minimum(gs, nsteps).astype(np.float64), np.minimum(cs, nsteps).astype(np.float64)
        if a.min() != a.max() or b.min() != b.max():
            ks_p.append(float(stats.ks_2samp(a, b).pvalue))
    k = max(len(zs), 1)
    z_crit = float(stats.norm.isf(0.005 / k))
    ks_min = min(ks_p) if ks_p else 1.0
    ks_thr = 0.01 / max(len(ks_p), 1)
    return {
        "instances_compared": len(zs), "solve_fraction_max_abs_z": float(np.max(np.abs(zs))) if zs else 0.0,
        "z_critical_bonferroni_alpha_0.01": z_crit, "steps_ks_min_p": ks_min, "ks_p_threshold_bonferroni_alpha_0.01": ks_thr,
        "matched": bool((not zs or np.max(np.abs(zs)) < z_crit) and ks_min > ks_thr),
        "gross_mismatch": bool(zs and np.max(np.abs(zs)) > 8.0),
    }


def leg_c2_moser(ctx):
    """Moser-Tardos on the C2 instance: same chains, same budget, resident solver; per-GPU (rank 0 reports its own GPU)."""
    import moser_rust
    eng, torch = ctx["eng"], ctx["torch"]
    inst, w, R = ctx["c2_inst"], ctx["c2_w"], CHAINS_PER_GPU
    solver = eng.ResidentSolver(inst, "moser", NSTEPS, R, False, True, 0.0, seed=1, chain_offset=ctx["rank"] * R)
    solver.set_weights(w, w)
    reps = max(2, min(ctx["steps"], 5))
    for i in range(2):
        solver.set_seed(300 + i, ctx["rank"] * R)
        solver.launch()
        solver.sync()
    ms, flips, steps = [], 0, 0
    for i in range(reps):
        ctx["flush"].fill_(i)
        torch.cuda.synchronize()
        solver.set_seed(3000 + i, ctx["rank"] * R)
        solver.launch()
        ms.append(solver.sync())
        r = solver.fetch_summary()
        flips += r.total_flips
        steps += r.total_steps
    kms = float(np.mean(ms))
    f_per_step = flips / max(steps, 1)
    bps = bytes_per_step(K, K * ALPHA, f_per_step)
    variant, launches = solver.variant, solver.launch_count()
    solver.close()
    # end to end through the reference-facing call
    moser_rust.set_seed(7000 + ctx["rank"])
    moser_rust.run_sls_python("moser", ctx["c2_path"], w, w, NSTEPS // 10, R, False, True, 0)
    t0 = time.perf_counter()
    out = moser_rust.run_sls_python("moser", ctx["c2_path"], w, w, NSTEPS, R, False, True, 0)
    e2e_s = time.perf_counter() - t0
    moser_rust.set_seed(None)
    e2e_steps = int(np.sum(out[3]))
    leg = {
        "workload": "C2 instance, Moser-Tardos resampling (lib.rs:20-49), uniform oracle, 4096 chains x 10^5 steps",
        "value": flips / (sum(ms) * 1e-3), "unit": "flips/s", "steps_per_s": steps / (sum(ms) * 1e-3),
        "flips_per_step": f_per_step, "kernel_ms": kms, "kernel": variant, "gpu_launches": reps * launches,
        "e2e": {"value": e2e_steps * f_per_step / e2e_s, "unit": "flips/s", "steps_per_s": e2e_steps / e2e_s,
                "h2d_bytes_per_step": ctx["c2_h2d"], "d2h_bytes_per_step": ctx["c2_d2h"],
                "includes": "DIMACS parse + CSR build + H2D + solve + D2H (run_sls_python); flips = steps x the device-timed flips per step"},
        "roofline": {"bound": "latency", "contract_bound": "hbm", "achieved": steps / reps * bps / (kms * 1e-3) / 1e9,
                     "peak": ctx["peaks"]["hbm_gbs"], "unit": "GB/s", "frac": steps / reps * bps / (kms * 1e-3) / 1e9 / ctx["peaks"]["hbm_gbs"],
                     "traffic": None, "bytes_per_step": bps, "peak_source": ctx["peaks"]["source"],
                     "note": "same kernel family as the headline (walk_k3s): dependent-latency bound, see DESIGN.md section 4"},
    }
    if ctx["cpu"]:
        cores = host_cores()
        nruns = min(R, cores * 16)
        r, dt = cpu_c2(ctx["c2_path"], "moser", nruns, NSTEPS, 91, cores)
        leg["cpu_baseline"] = {"value": r.total_flips / dt, "unit": "flips/s", "steps_per_s": r.total_steps / dt, "cores": cores,
                               "kind": "port", "sample": f"{nruns} of {R} chains x {NSTEPS} steps, {cores} threads, parse included"}
    return leg


def sharded_batch(ctx, costs, runner, nruns):
    """run_batch_sharded over the ranks of the run (the product's sharding path, NCCL); a single process runs its runner on
    everything."""
    from sls_sat_solving_with_deep_learning_b200 import sharding

    if ctx["dist"] is None:
        res = runner(list(range(len(costs))))
        found = np.stack([np.asarray(r.found, dtype=bool) for r in res])
        steps = np.stack([np.asarray(r.steps, dtype=np.uint64) for r in res])
        flips = np.asarray([r.total_flips for r in res], dtype=np.int64)
        return found, steps, flips
    found, steps, _, flips = sharding.run_batch_sharded(runner, costs, nruns, device="cuda")
    return found, steps, flips


def barrier_max(ctx, seconds):
    """max over ranks of a per-rank duration"""
    if ctx["dist"] is None:
        return seconds
    t = ctx["torch"].tensor([seconds], dtype=ctx["torch"].float64, device="cuda")
    ctx["dist"].all_reduce(t, op=ctx["dist"].ReduceOp.MAX)
    return float(t.item())


def leg_c1_batch(ctx):
    """2052 C1-shaped instances x 100 chains, split over the ranks by instance (strong scaling).  Reports flips/s and
    instances solved/s for Moser-Tardos and probSAT; rank 0 at N = 1 also runs the CPU port on every 16th instance."""
    eng, sharding = ctx["eng"], ctx["sharding"]
    shapes = c1_shapes()
    costs = [m * C1_RUNS for (_, m, _) in shapes]
    mine = sharding.instance_shard(costs, ctx["world"])[ctx["rank"]]
    csrs = {i: csr_of(shapes[i]) for i in mine}  # synthetic data generation: outside every timed region
    out = {"workload": f"{C1_INSTANCES} synthetic random 3-SAT instances, n in {{100,200,300}}, alpha 1.0-4.5, uniform oracle, "
                       f"{C1_RUNS} chains x {C1_NSTEPS} steps each (C1 shape), instances split over {ctx['world']} GPU(s)"}
    # warm the kernels of this leg
    wi = [eng.Instance.from_csr(shapes[i][0], *csrs[i]) for i in mine[:4]]
    for algo in ("moser", "probsat"):
        eng.run_sls_batch(wi, algo, [np.full(x.n, 0.5) for x in wi], [np.full(x.n, 0.5) for x in wi], 100, C1_RUNS)
    for algo in ("moser", "probsat"):
        kernel_ms = [0.0]

        def runner(idx):
            insts = [eng.Instance.from_csr(shapes[i][0], *csrs[i]) for i in idx]  # parse-equivalent: CSR -> handle
            w = [np.full(shapes[i][0], 0.5) for i in idx]
            res = eng.run_sls_batch(insts, algo, w, w, C1_NSTEPS, C1_RUNS, True, 0.0, instance_seeds=[1 + i for i in idx])
            kernel_ms[0] = res[0].kernel_ms if res else 0.0
            return res

        if ctx["dist"] is not None:
            ctx["dist"].barrier()
        t0 = time.perf_counter()
        found, steps, flips = sharded_batch(ctx, costs, runner, C1_RUNS)
        wall = barrier_max(ctx, time.perf_counter() - t0)
        kms = barrier_max(ctx, kernel_ms[0] * 1e-3)
        solved = int(found.any(axis=1).sum())
        tot_flips, tot_steps = int(flips.sum()), int(steps.sum())
        leg = {
            "value": tot_flips / kms, "unit": "flips/s", "steps_per_s": tot_steps / kms, "kernel_ms_max_over_ranks": kms * 1e3,
            "instances_solved": solved, "instances_solved_per_s": solved / kms, "chains_solved": int(found.sum()),
            "e2e": {"value": tot_flips / wall, "unit": "flips/s", "instances_solved_per_s": solved / wall, "wall_s": wall,
                    "h2d_bytes_per_step": int(sum(16 * 3 * m + 8 * m + 24 * n for (n, m, _) in shapes)),
                    "d2h_bytes_per_step": int(C1_INSTANCES * C1_RUNS * 22),
                    "includes": "CSR -> instance handles (occurrence lists, images), H2D, kernels, D2H, result gather over the ranks"},
            "gpu_launches": 1 + len(mine),
        }
        f_per_step = tot_flips / max(tot_steps, 1)
        bps = bytes_per_step(3, 3 * float(np.mean([m / n for (n, m, _) in shapes])), f_per_step)
        leg["roofline"] = {"bound": "latency", "contract_bound": "hbm", "achieved": tot_steps * bps / kms / 1e9, "peak": ctx["peaks"]["hbm_gbs"],
                           "unit": "GB/s", "frac": tot_steps * bps / kms / 1e9 / ctx["peaks"]["hbm_gbs"], "traffic": None,
                           "bytes_per_step": bps, "peak_source": ctx["peaks"]["source"],
                           "note": "instances and chain state are shared-memory / L2 resident: throughput regime of walk_k3d (warp instructions per step)"}
        if ctx["cpu"]:
            cores = host_cores()
            sample = list(range(0, C1_INSTANCES, C1_CPU_STRIDE))
            cres, dt = cpu_batch([shapes[i] for i in sample], [csrs[i] for i in sample], algo, C1_NSTEPS, C1_RUNS, 1_000_003, cores)
            csolved = sum(bool(r.found.any()) for r in cres)
            leg["cpu_baseline"] = {
                "value": sum(r.total_flips for r in cres) / dt, "unit": "flips/s", "instances_solved_per_s": csolved / dt,
                "instances_solved": csolved, "cores": cores, "kind": "port",
                "sample": f"every {C1_CPU_STRIDE}th instance ({len(sample)} of {C1_INSTANCES}), all {C1_RUNS} chains x {C1_NSTEPS} steps, {cores} threads",
            }
            gpu_solved_same = int(found[sample].any(axis=1).sum())
            leg["cpu_baseline"]["gpu_instances_solved_on_the_same_sample"] = gpu_solved_same
            leg["solve_rate_parity"] = solve_rate_parity([found[i] for i in sample], [steps[i] for i in sample],
                                                         [r.found for r in cres], [r.steps for r in cres], C1_NSTEPS)
            if leg["solve_rate_parity"]["gross_mismatch"]:
                raise RuntimeError(f"c1_batch/{algo}: GPU and CPU-port solve fractions differ grossly: {leg['solve_rate_parity']}")
        out[algo] = leg
    return out


def leg_c4_pipeline(ctx):
    """BASELINE configs[3]: batched GNN oracle forward on 10,000 synthetic instances (n=500, alpha=4.2), then oracle-guided
    Moser-Tardos (100 chains x 10^4 steps), instances split over the ranks."""
    eng, gnn, synth, sharding, torch = ctx["eng"], ctx["gnn"], ctx["synth"], ctx["sharding"], ctx["torch"]
    shapes = [(C4_N, C4_M, C4_SEED + i) for i in range(C4_INSTANCES)]
    costs = [1.0] * C4_INSTANCES
    mine = sharding.instance_shard(costs, ctx["world"])[ctx["rank"]]
    csrs = {i: csr_of(shapes[i]) for i in mine}
    params = synth.random_oracle_params(seed=42)
    net = gnn.InteractionNetworkLCG(params)
    # warm-up on one chunk of the real size (handles that are dropped again): the library's memory pool and the forward's
    # workspace reach their working size before the timed pass -- a fresh process otherwise spends 0.5-1.5 s of the first
    # chunks in cudaMalloc (measured: 1620-1910 instances/s for the first runs on a box, 2430 once the allocations are warm)
    nets = [net] + [gnn.InteractionNetworkLCG(params) for _ in range(C4_HANDLES - 1)]
    warm = [eng.Instance.from_csr(C4_N, *csrs[i]) for i in mine[:C4_CHUNK]]
    pw = [one.probabilities(warm) for one in nets][0]
    eng.run_sls_batch(warm, "moser", pw, pw, C4_NSTEPS, C4_RUNS)
    del warm, pw
    stage = {"build_s": 0.0, "forward_s": 0.0, "forward_device_ms": 0.0, "walk_s": 0.0, "walk_kernel_ms": 0.0}
    keep = {}

    def runner(idx):
        # the product's pipeline: chunks of C4_CHUNK instances, handle build / forward / walk overlapped on three threads
        st = {}
        res, _, insts = ctx["pipeline"].solve_with_oracle(
            nets if len(nets) > 1 else net, [csrs[i] for i in idx], "moser", C4_NSTEPS, C4_RUNS, make_instance=lambda c: eng.Instance.from_csr(C4_N, *c),
            instance_seeds=[1 + i for i in idx], chunk=C4_CHUNK, keep_instances=True, stats=st)
        stage.update({k: v for k, v in st.items() if k != "timeline"})  # sums over the chunks; the Gantt view is tools/c4_timeline.py
        if os.environ.get("SLS_BENCH_C4_TIMELINE") and ctx["rank"] == 0:
            for row in sorted(st["timeline"], key=lambda x: x[2]):
                print("c4 timeline", row, file=sys.stderr)
        keep["insts"] = insts
        return res

    if ctx["dist"] is not None:
        ctx["dist"].barrier()
    t0 = time.perf_counter()
    found, steps, flips = sharded_batch(ctx, costs, runner, C4_RUNS)
    wall = barrier_max(ctx, time.perf_counter() - t0)
    per_stage = {k: barrier_max(ctx, v) for k, v in stage.items()}
    solved = int(found.any(axis=1).sum())
    dev_s = (per_stage["forward_device_ms"] + per_stage["walk_kernel_ms"]) * 1e-3
    leg = {
        "workload": f"{C4_INSTANCES} synthetic 3-SAT instances n={C4_N} m={C4_M}, random-init oracle -> p, Moser-Tardos {C4_RUNS} chains x "
                    f"{C4_NSTEPS + 1} steps, instances split over {ctx['world']} GPU(s) ({len(mine)} on rank 0)",
        "value": C4_INSTANCES / wall, "unit": "instances/s", "instances_solved": solved, "instances_solved_per_s": solved / wall,
        "wall_s": wall, "device_s_max_over_ranks": dev_s, "stages_max_over_ranks": per_stage,
        "walk_flips_per_s": int(flips.sum()) / (per_stage["walk_kernel_ms"] * 1e-3),
        "walk_steps_per_s": int(steps.sum()) / (per_stage["walk_kernel_ms"] * 1e-3),
        "e2e": {"value": C4_INSTANCES / wall, "unit": "instances/s", "h2d_bytes_per_step": int(C4_INSTANCES * (C4_M * 3 * 32 + C4_M * 80 + C4_N * 40)),
                "d2h_bytes_per_step": int(C4_INSTANCES * (C4_N * 4 + C4_RUNS * 22)),
                "includes": "CSR -> handles, instance upload, LCG graphs on the device, forward, p to the host and back as weights, walk, D2H, gather; "
                            f"pipeline.solve_with_oracle in chunks of {C4_CHUNK} instances, {C4_HANDLES} forward worker(s), the stages overlapped (stage times are sums over chunks)"},
        "gpu_launches": None,
    }
    # ---- the oracle forward by itself: one workspace chunk (256 instances), device time, kernel classes, rooflines
    n_fw = min(256, len(mine))
    insts = keep["insts"][:n_fw]
    net.probabilities(insts)
    best_ms, best_wall = 1e30, 1e30
    for _ in range(3):
        t1 = time.perf_counter()
        net.probabilities(insts)
        best_wall = min(best_wall, time.perf_counter() - t1)
        best_ms = min(best_ms, net.last_kernel_ms)
    launches = int(net.launch_count())
    net.set_profile(True)
    net.probabilities(insts)
    prof = net.profile()
    net.set_profile(False)
    n_nodes, n_edges = 2 * C4_N + C4_M, 3 * C4_M + C4_N
    flops = n_fw * (1.3984e6 * n_edges + 1.5328e6 * n_nodes)  # dense FLOPs as the reference computes them
    mlp_ms = prof["edge_update_ms"] + prof["node_update_ms"]
    H = 200
    # segment sums, algorithmic bytes per layer: every new edge row read once, one summed row written per (node, direction)
    # with two or more edges (clause nodes receive k = 3, literal nodes send their occurrences [+ the pair edge])
    agg_rows = C4_M + 2 * C4_N
    seg_bytes = 5 * n_fw * (n_edges + agg_rows) * H * 4
    pk = ctx["peaks"]
    fw = {
        "workload": f"C4 shape: {n_fw} instances n={C4_N} m={C4_M} (LCG: {n_nodes} nodes, {n_edges} edges each), random-init interaction network",
        "instances_per_s": n_fw / (best_ms * 1e-3), "device_ms": best_ms, "e2e_instances_per_s": n_fw / best_wall,
        "precision": os.environ.get("SLS_B200_GNN_PRECISION", "fp16x2"), "dense_tflops": flops / (best_ms * 1e-3) / 1e12,
        "gpu_launches": launches, "kernel_class_ms": prof,
        "kernel": "gnn::mlp_ln_persistent_ts_kernel (tcgen05 kind::f16, A operand in TMEM) + segment_sum_kernel",
        "roofline": {
            "bound": "tensor", "achieved": flops / (mlp_ms * 1e-3) / 1e12, "peak": pk["bf16_tflops_sustained"], "unit": "TFLOP/s",
            "frac": flops / (mlp_ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"], "traffic": None, "peak_source": pk["source"],
            "mma_issued_tflops": 3 * flops / (mlp_ms * 1e-3) / 1e12, "mma_issued_frac": 3 * flops / (mlp_ms * 1e-3) / 1e12 / pk["bf16_tflops_sustained"],
            "note": "dense FLOPs of the reference's fp32 contraction over the MLP kernels' device time; every product is three bf16 MMAs "
                    "(a0*b0 + a1*b0 + a0*b1), which is what mma_issued_* counts",
        },
        "segment_sum_roofline": {
            "bound": "hbm", "achieved": seg_bytes / (prof["segment_sum_ms"] * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
            "frac": seg_bytes / (prof["segment_sum_ms"] * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None, "peak_source": pk["source"],
            "algorithmic_bytes": seg_bytes,
        },
    }
    if ctx["cpu"]:
        cores = host_cores()
        sample = [csrs[i] for i in mine[:6]]
        fw["cpu_baseline"] = {"value": cpu_oracle_forward(params, sample, C4_N, cores), "unit": "instances/s", "cores": cores, "kind": "port",
                              "sample": f"{len(sample)} of the instances, one at a time (evaluate_with_given_params.py:116), torch-CPU fp32 "
                                        f"restatement oracle/gnn_torch.py, {cores} threads"}
        # the walk stage of the same pipeline on the CPU port, oracle probabilities from the GPU forward
        idx = mine[:8]
        probs = net.probabilities(keep["insts"][:8])
        cres, dt = cpu_batch([shapes[i] for i in idx], [csrs[i] for i in idx], "moser", C4_NSTEPS, C4_RUNS, 2_000_003, cores, weights=probs)
        fw_rate = fw["cpu_baseline"]["value"]
        walk_rate = len(idx) / dt
        leg["cpu_baseline"] = {"value": 1.0 / (1.0 / fw_rate + 1.0 / walk_rate), "unit": "instances/s", "cores": cores, "kind": "port",
                               "forward_instances_per_s": fw_rate, "walk_instances_per_s": walk_rate,
                               "sample": f"forward on {len(sample)} instances (torch-CPU fp32), walk on {len(idx)} instances (C port, {cores} threads); "
                                         "rates combined as a sequential pipeline"}
    leg["oracle_forward"] = fw
    leg["roofline"] = fw["roofline"]
    for one in nets:
        one.close()
    return leg


def leg_c5(ctx):
    """BASELINE configs[4]: n = 10^6, alpha = 4.0, 65,536 chains per GPU (weak scaling), chain state in HBM."""
    eng, torch = ctx["eng"], ctx["torch"]
    inst = ctx["c5_inst"]
    w = np.full(C5_N, 0.5)
    R = C5_CHAINS
    d = 3.0 * C5_M / C5_N
    out = {"workload": f"synthetic random 3-SAT n={C5_N} m={C5_M}, uniform oracle, {R} chains per GPU x {C5_NSTEPS} steps, chain state in HBM "
                       f"({ctx['world']} GPU(s), chains keyed by global id)"}
    pk = ctx["peaks"]
    for algo in ("probsat", "moser"):
        s = eng.ResidentSolver(inst, algo, C5_NSTEPS, R, False, True, 0.0, seed=1, chain_offset=ctx["rank"] * R)
        s.set_weights(w, w)
        s.launch()
        s.sync()  # warm
        s0 = eng.ResidentSolver(inst, algo, 0, 256, False, True, 0.0, seed=1)  # (variant probe only)
        s0.close()
        if ctx["dist"] is not None:
            ctx["dist"].barrier()
        torch.cuda.synchronize()
        s.set_seed(11, ctx["rank"] * R)
        s.launch()
        ms = s.sync()
        r = s.fetch_summary()
        variant, launches = s.variant, s.launch_count()
        s.close()
        # init only (nsteps = 0): separates the full scan (lib.rs:206-214) from the walk
        si = eng.ResidentSolver(inst, algo, 0, R, False, True, 0.0, seed=11, chain_offset=ctx["rank"] * R)
        si.set_weights(w, w)
        si.launch()
        si.sync()
        si.launch()
        init_ms = si.sync()
        si.close()
        ms_max = barrier_max(ctx, ms * 1e-3)
        tot_flips, tot_steps = r.total_flips, r.total_steps
        if ctx["dist"] is not None:
            t = torch.tensor([tot_flips, tot_steps], dtype=torch.int64, device="cuda")
            ctx["dist"].all_reduce(t)
            tot_flips, tot_steps = int(t[0].item()), int(t[1].item())
        walk_ms = max(ms - init_ms, 1e-6)
        f_per_step = r.total_flips / max(r.total_steps, 1)
        bps = bytes_per_step(3, d, f_per_step)
        # init: every slice of 256 chains streams the literals once (4k bytes per clause) and every chain writes m/8 bytes
        # of violated bits and reads n/8 of assignment bits
        init_bytes = (R / 256) * C5_M * 12 + R * (C5_M / 8 + C5_N / 8)
        leg = {
            "value": tot_flips / ms_max, "unit": "flips/s", "steps_per_s": tot_steps / ms_max, "kernel_ms": ms, "init_ms": init_ms,
            "walk_ms": walk_ms, "walk_flips_per_s_this_gpu": r.total_flips / (walk_ms * 1e-3), "kernel": variant, "gpu_launches": launches,
            "mean_best_energy": float(r.best_energy_run.mean()),
            "roofline": {"bound": "hbm", "achieved": r.total_steps * bps / (walk_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                         "frac": r.total_steps * bps / (walk_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None, "bytes_per_step": bps,
                         "peak_source": pk["source"], "kernel_part": "walk (total - init)",
                         "random_sector_peak_gbs": (ctx["ubench"] or {}).get("hbm_sector32_gbs")},
            "init_roofline": {"bound": "hbm", "achieved": init_bytes / (init_ms * 1e-3) / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                              "frac": init_bytes / (init_ms * 1e-3) / 1e9 / pk["hbm_gbs"], "traffic": None, "algorithmic_bytes": init_bytes,
                              "peak_source": pk["source"]},
        }
        if algo == "probsat":
            t0 = time.perf_counter()
            re = eng.run_sls(inst, algo, w, w, C5_NSTEPS, R, False, True, 0.0, seed=12, chain_offset=ctx["rank"] * R)
            e2e_s = barrier_max(ctx, time.perf_counter() - t0)
            leg["e2e"] = {"value": re.total_flips * ctx["world"] / e2e_s, "unit": "flips/s", "h2d_bytes_per_step": int(C5_N * 16),
                          "d2h_bytes_per_step": int(R * 22 + C5_N / 8),
                          "includes": "weights H2D, state allocation (42 GB), init, walk, results + best assignment D2H (instance handle resident)"}
            if ctx["cpu"]:
                from oracle import sls_oracle

                cores = host_cores()
                oi = sls_oracle.Instance.from_csr(C5_N, *ctx["c5_csr"])
                nruns = 2 * cores
                t0 = time.perf_counter()
                rc = oi.run_sls(algo, w, w, C5_NSTEPS, nruns, False, True, 0.0, seed=5, nthreads=cores)
                dt = time.perf_counter() - t0
                leg["cpu_baseline"] = {"value": rc.total_flips / dt, "unit": "flips/s", "cores": cores, "kind": "port",
                                       "sample": f"{nruns} of {R} chains x {C5_NSTEPS} steps, {cores} threads (the full scan of lib.rs:213 dominates, as on the GPU)"}
        out[algo] = leg
    return out


def leg_clause_eval(ctx):
    """sls_eval_assignments (clause_is_violated lib.rs:13-18 / get_violated_constraints) on the C5 instance: 1024 assignments
    x 4*10^6 clauses, energies only and with the violated-clause bits, against the L2 random-sector rate measured in this
    run (the bit-sliced kernels gather one 32-byte sector per literal per 256 assignments)."""
    eng = ctx["eng"]
    from sls_sat_solving_with_deep_learning_b200 import _lib

    lib = _lib.load()
    inst = ctx["c5_inst"]
    B = 1024
    rng = np.random.default_rng(3)
    a = rng.integers(0, 2, size=(B, C5_N), dtype=np.uint8)
    mwords = (C5_M + 31) // 32
    energy = np.zeros(B, dtype=np.uint32)
    bits = np.zeros((B, mwords), dtype=np.uint32)
    sector_bytes = B / 256 * 3 * C5_M * 32  # L2 -> SM sectors the algorithm needs
    ub = ctx["ubench"] or {}
    peak = ub.get("l2_sector32_gbs_8_in_flight") or ub.get("l2_sector32_gbs_4_in_flight")
    out = {"workload": f"{B} random assignments x the C5 instance ({C5_M} clauses), host buffers in / out"}
    for name, vb in (("energy_only", None), ("with_violated_bits", bits)):
        best, kern, wall = 1e30, 1e30, 1e30
        for _ in range(3):
            t0 = time.perf_counter()
            rc = lib.sls_eval_assignments(inst._h, a.ctypes.data, B, vb.ctypes.data if vb is not None else None, energy.ctypes.data)
            wall = min(wall, time.perf_counter() - t0)
            if rc != 0:
                raise RuntimeError(lib.sls_last_error().decode())
            if eng.eval_last_kernel_ms() < best:
                best, kern = eng.eval_last_kernel_ms(), eng.eval_last_eval_kernel_ms()
        out[name] = {
            # value: both device kernels (byte -> bit-slice transposition of the 1 GB of assignment bytes + evaluation);
            # the roofline is the evaluation kernel's own (kernel_ms)
            "value": B * C5_M / (best * 1e-3), "unit": "clause evaluations/s", "device_ms": best, "kernel_ms": kern,
            "kernel": "eval_energy_sliced_kernel" if vb is None else "eval_sliced_kernel<out>",
            "e2e": {"value": B * C5_M / wall, "unit": "clause evaluations/s", "h2d_bytes_per_step": int(a.nbytes),
                    "d2h_bytes_per_step": int(energy.nbytes + (vb.nbytes if vb is not None else 0))},
            "roofline": {"bound": "l2 random sectors", "contract_bound": "hbm", "achieved": sector_bytes / (kern * 1e-3) / 1e9, "peak": peak,
                         "unit": "GB/s", "frac": (sector_bytes / (kern * 1e-3) / 1e9 / peak) if peak else None, "traffic": None,
                         "peak_source": "measured in this run (tools/ubench.cu l2_sector32)",
                         "hbm_algorithmic_gbs": (B / 256 * C5_M * 12 + B * C5_N / 8 + (B * C5_M / 8 if vb is not None else 0)) / (kern * 1e-3) / 1e9},
            "mean_energy": float(energy.mean()),
        }
    if ctx["cpu"]:
        from oracle import sls_oracle

        oi = sls_oracle.Instance.from_csr(C5_N, *ctx["c5_csr"])
        t0 = time.perf_counter()
        es = [oi.energy(a[i]) for i in range(8)]
        dt = time.perf_counter() - t0
        assert es == [int(x) for x in energy[:8]], "clause evaluation differs from the CPU port"
        out["cpu_baseline"] = {"value": 8 * C5_M / dt, "unit": "clause evaluations/s", "cores": 1, "kind": "port",
                               "sample": "8 of the 1024 assignments, scalar port (energies equal the GPU's)"}
    return out


# ======================================================================================================================
# reference arm
# ======================================================================================================================
def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path on the host cores (the C port of rust/src/lib.rs:
    no Rust toolchain in this image).  Rank 0 alone works."""
    if rank != 0:
        return
    from oracle import sls_oracle  # noqa: F401  (bench.py may execute oracle/ here and in the cpu_baseline legs only)
    from sls_sat_solving_with_deep_learning_b200 import synth

    cores = host_cores()
    row, lits = synth.random_ksat_csr(N_VARS, N_CLAUSES, K, INSTANCE_SEED)
    tmp = tempfile.mkdtemp(prefix="slsb200_")
    path = os.path.join(tmp, "c2.cnf")
    synth.write_dimacs(path, N_VARS, row, lits)
    # bounded sample of the workload: same instance, same steps per chain, fewer chains
    nruns = min(CHAINS_PER_GPU, cores * 16)
    for w in range(args.warmup):
        cpu_c2(path, ALGO, max(cores, 1), NSTEPS // 10, 1000 + w, cores)
    flips = 0
    secs = 0.0
    for s in range(args.steps):
        r, dt = cpu_c2(path, ALGO, nruns, NSTEPS, 2000 + s, cores)
        flips += r.total_flips
        secs += dt
    value = flips / secs
    sample = f"{nruns} of {CHAINS_PER_GPU} chains x {NSTEPS} steps per step, {cores} threads, parse included"
    legs = {}
    if args.legs != "none":
        r, dt = cpu_c2(path, "moser", nruns, NSTEPS, 91, cores)
        legs["c2_moser"] = {"value": r.total_flips / dt, "unit": "flips/s", "steps_per_s": r.total_steps / dt,
                            "sample": f"{nruns} chains x {NSTEPS} steps, {cores} threads"}
        shapes = c1_shapes()
        idx = list(range(0, C1_INSTANCES, C1_CPU_STRIDE))
        csrs = [csr_of(shapes[i]) for i in idx]
        for algo in ("moser", "probsat"):
            cres, dt = cpu_batch([shapes[i] for i in idx], csrs, algo, C1_NSTEPS, C1_RUNS, 1_000_003, cores)
            solved = sum(bool(x.found.any()) for x in cres)
            legs[f"c1_batch_{algo}"] = {"value": sum(x.total_flips for x in cres) / dt, "unit": "flips/s", "instances_solved_per_s": solved / dt,
                                        "sample": f"every {C1_CPU_STRIDE}th of {C1_INSTANCES} instances, {C1_RUNS} chains x {C1_NSTEPS} steps, {cores} threads"}
        c4 = [synth.random_ksat_csr(C4_N, C4_M, 3, C4_SEED + i) for i in range(4)]
        legs["oracle_forward"] = {"value": cpu_oracle_forward(synth.random_oracle_params(seed=42), c4, C4_N, cores), "unit": "instances/s",
                                  "sample": f"4 C4-shaped instances, torch-CPU fp32 (oracle/gnn_torch.py), {cores} threads"}
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
        "legs": legs,
    }
    print(json.dumps(line), flush=True)


# ======================================================================================================================
# own arm
# ======================================================================================================================
def run_own(args, rank, world, local_rank):
    import torch

    import sls_sat_solving_with_deep_learning_b200 as eng
    from sls_sat_solving_with_deep_learning_b200 import gnn, pipeline, sharding, synth

    if not torch.cuda.is_available() or eng.device_count() < 1:
        raise RuntimeError("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local_rank)
    eng.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist_

        dist = dist_
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    legs = [] if args.legs == "none" else (list(ALL_LEGS) if args.legs == "all" else [x for x in args.legs.split(",") if x])
    for x in legs:
        if x not in ALL_LEGS:
            raise SystemExit(f"unknown leg {x!r}; choose from {ALL_LEGS}")
    pk = peaks()

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
    variant, launches_per_step = solver.variant, solver.launch_count()

    # ---- the gathered tables are USED: rank 0 re-computes chains of another rank's block from their global ids and
    # compares with what the collective delivered (the last timed step's seed)
    sharding_check = None
    if dist:
        last_seed = 1000 + args.steps - 1
        ok = True
        if rank == 0:
            other = world - 1
            for j0 in (0, R // 2 + 3, R - 8):
                chk = eng.run_sls(inst, ALGO, w, w, NSTEPS, 8, False, True, 0.0, seed=last_seed, chain_offset=other * R + j0)
                gs = g_steps[other * R + j0: other * R + j0 + 8].cpu().numpy().astype(np.uint64)
                gf = g_found[other * R + j0: other * R + j0 + 8].cpu().numpy().astype(bool)
                ok = ok and bool((gs == chk.steps).all() and (gf == chk.found).all())
            own = d_steps.cpu().numpy()
            ok = ok and bool((g_steps[:R].cpu().numpy() == own).all())
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device="cuda")
        dist.broadcast(flag, src=0)
        if int(flag.item()) != 1:
            raise RuntimeError("sharding check failed: gathered per-chain tables differ from a re-computation of the same global chain ids")
        sharding_check = "ok"

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
    d2h = R * (1 + 1 + 8 + 8 + 4) + 256 + (N_VARS + 31) // 32 * 4
    moser_rust.set_seed(None)
    kms = float(np.mean(kernel_ms))
    solver.close()

    cpu = rank == 0 and world == 1 and not args.no_cpu_baseline
    # ---- ceilings and latencies measured in this run, on this box (rank 0 at N = 1)
    ubench = None
    lone_ns = None
    if rank == 0 and world == 1 and not args.no_ubench:
        ubench = run_ubench(True)
        lone = eng.ResidentSolver(inst, ALGO, 20_000, 1, False, True, 0.0, seed=3)
        lone.set_weights(w, w)
        lone.launch()
        lone.sync()
        lone.launch()
        lone_ns = lone.sync() * 1e6 / 20_000
        lone.close()

    line = None
    if rank == 0:
        achieved = R * NSTEPS * BYTES_PER_FLIP / (kms * 1e-3) / 1e9
        # L2 -> SM sectors per flip of one walk_k3s launch (ncu lts__t_sectors_srcunit_tex_op_read, profiles/r02_walk_k3s_summary.md)
        l2_bytes_per_flip = 300.0
        line = {
            "metric": "variable flips/sec", "value": value, "unit": "flips/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32 bit-ops + f64 probabilities", "data": "synthetic",
            "config": workload_config(),
            "run": {
                "parallelism": f"chains sharded over {world} GPU(s), {R} per GPU; C1 / C4 legs split by instance",
                "kernel": variant,
                "l2": "flushed between timed steps (256 MB write); the 4.7 MB instance is L2-resident by nature",
                "legs": legs,
            },
            "clocks": clocks.summary(),
            "e2e": {"value": e2e_value, "unit": "flips/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "steps": e2e_steps, "includes": "DIMACS parse + CSR build + H2D + solve + D2H (run_sls_python)"},
            "gpu_launches": int(args.steps * launches_per_step),
            "roofline": {
                # the contract's figure: algorithmic bytes over kernel time against the measured HBM copy peak.  The kernel is
                # NOT bandwidth bound (state in shared memory, instance in L2, DRAM idle): what bounds it is the dependent
                # latency of one step -- "l2" / "latency" below give that picture with ceilings measured in this run
                "bound": "latency", "contract_bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                "frac": achieved / pk["hbm_gbs"],
                # dram__bytes_read.sum + dram__bytes_write.sum of one walk_k3s launch (ncu --set full,
                # profiles/r02_walk_k3s_summary.md): independent of the step count, instance and state never leave L2 / smem
                "traffic": 6.18e6, "traffic_source": "ncu capture of this kernel in the bench command, profiles/r02_walk_k3s_summary.md (not measurable inside a run)",
                "peak_source": pk["source"], "bytes_per_flip": BYTES_PER_FLIP, "kernel_ms": kms,
                "l2": {"achieved_gbs": R * NSTEPS * l2_bytes_per_flip / (kms * 1e-3) / 1e9,
                       "random_sector_peak_gbs": (ubench or {}).get("l2_sector32_gbs_8_in_flight"),
                       "stream_peak_gbs": (ubench or {}).get("l2_stream_gbs"), "ceilings": "measured in this run (ubench)" if ubench else None},
                "latency": {"loaded_ns_per_step": kms * 1e6 / NSTEPS, "lone_chain_ns_per_step": lone_ns,
                            "l2_hit_cycles": ((ubench or {}).get("latency_cycles") or {}).get("l2_hit"),
                            "dependent_l2_trip_ns_under_load": ((ubench or {}).get("l2_rec64_dependent") or {}).get("ns_per_link"),
                            "source": "measured in this run" if ubench else None},
            },
            "ubench": ubench,
        }
        if sharding_check:
            line["sharding_check"] = sharding_check
    if dist:
        dist.barrier()

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle port on the host cores
    if cpu:
        cores = host_cores()
        nruns = min(CHAINS_PER_GPU, cores * 32)
        r, dt = cpu_c2(path, ALGO, nruns, NSTEPS, 77, cores)
        line["cpu_baseline"] = {
            "value": r.total_flips / dt, "unit": "flips/s", "cores": cores, "kind": "port",
            "sample": f"{nruns} of {CHAINS_PER_GPU} chains x {NSTEPS} steps, {cores} threads, parse included",
        }

    # ---- the other legs: every rank takes part (sharded work, collectives); rank 0 reports.  A failing leg fails the run.
    ctx = {"eng": eng, "gnn": gnn, "synth": synth, "sharding": sharding, "pipeline": pipeline, "torch": torch, "dist": dist, "rank": rank, "world": world,
           "steps": args.steps, "cpu": cpu, "peaks": pk, "ubench": ubench, "flush": flush, "c2_inst": inst, "c2_w": w, "c2_path": path,
           "c2_h2d": h2d, "c2_d2h": d2h}
    results = {}
    if "c5" in legs or "clause_eval" in legs:
        ctx["c5_csr"] = synth.random_ksat_csr(C5_N, C5_M, 3, C5_SEED)
        ctx["c5_inst"] = eng.Instance.from_csr(C5_N, *ctx["c5_csr"])
    fns = {"c2_moser": leg_c2_moser, "c1_batch": leg_c1_batch, "c4_pipeline": leg_c4_pipeline, "c5": leg_c5, "clause_eval": leg_clause_eval}
    for name in legs:
        if name == "clause_eval" and (rank != 0 or world != 1):
            continue  # a per-GPU kernel measurement: reported at N = 1
        t0 = time.perf_counter()
        results[name] = fns[name](ctx)
        results[name]["leg_wall_s"] = time.perf_counter() - t0
    if rank == 0:
        line["legs"] = results
        c1 = results.get("c1_batch")
        if c1:
            line["instances_solved_per_s"] = {"probsat": c1["probsat"]["e2e"]["instances_solved_per_s"],
                                              "moser": c1["moser"]["e2e"]["instances_solved_per_s"],
                                              "workload": "c1_batch leg, end to end"}
        print(json.dumps(line), flush=True)
    if dist:
        dist.barrier()
        dist.destroy_process_group()


def main():
    global NSTEPS
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--legs", default="all", help="comma-separated subset of " + ",".join(ALL_LEGS) + ", or all / none")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ubench", action="store_true", help="skip the in-run L2 / HBM microbenchmarks")
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
