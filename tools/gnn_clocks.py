"""SM clock, power and throttle reasons WHILE the oracle forward runs back to back (NVML samples every 20 ms), next to the same
samples for the C2 walk kernel.  The MLP kernel's timeline is measured in SM cycles (clock64); this tells what a cycle is worth
under that kernel's load.  usage: python tools/gnn_clocks.py [seconds]"""
import json
import os
import sys
import threading
import time

import numpy as np
import pynvml

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import sls_sat_solving_with_deep_learning_b200 as eng  # noqa: E402
from sls_sat_solving_with_deep_learning_b200 import gnn, synth  # noqa: E402

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 3.0
eng.set_device(0)
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)


class Sampler:
    def __enter__(self):
        self.rows, self.stop = [], threading.Event()
        self.t = threading.Thread(target=self.loop, daemon=True)
        self.t.start()
        return self

    def loop(self):
        while not self.stop.is_set():
            self.rows.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                              pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)))
            self.stop.wait(0.02)

    def __exit__(self, *a):
        self.stop.set()
        self.t.join()

    def summary(self):
        r = self.rows[len(self.rows) // 5:]  # skip the ramp
        reasons = 0
        for x in r:
            reasons |= x[2]
        names = {pynvml.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap", pynvml.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
                 pynvml.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
                 pynvml.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
                 pynvml.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake"}
        return {"samples": len(r), "sm_mhz_median": float(np.median([x[0] for x in r])), "sm_mhz_min": min(x[0] for x in r),
                "power_w_median": float(np.median([x[1] for x in r])), "power_w_max": max(x[1] for x in r),
                "reasons": sorted(v for k, v in names.items() if reasons & k)}


n, m = 500, 2100
insts = [eng.Instance.from_csr(n, *synth.random_ksat_csr(n, m, 3, 20260104 + i)) for i in range(256)]
net = gnn.InteractionNetworkLCG(synth.random_oracle_params(seed=42))
net.probabilities(insts)
ms = []
with Sampler() as s:
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        net.probabilities(insts)
        ms.append(net.last_kernel_ms)
out = {"oracle_forward": dict(s.summary(), device_ms_median=float(np.median(ms)), device_ms_first=ms[0], forwards=len(ms)),
       "power_limit_w": pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1000.0,
       "sm_max_mhz": pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)}
row, lits = synth.random_ksat_csr(10_000, 42_000, 3, 20260102)
inst = eng.Instance.from_csr(10_000, row, lits)
w = np.full(10_000, 0.5)
sol = eng.ResidentSolver(inst, "probsat", 100_000, 4096, False, True, 0.0, seed=1)
sol.set_weights(w, w)
sol.launch()
sol.sync()
kms = []
with Sampler() as s:
    t0 = time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        sol.launch()
        kms.append(sol.sync())
out["c2_walk"] = dict(s.summary(), kernel_ms_median=float(np.median(kms)))
print(json.dumps(out))
