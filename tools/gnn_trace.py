"""Timeline of one persistent CTA of the oracle-forward MLP kernel (development builds with -DGNN_TRACE; see csrc/gnn_kernels.cuh
GNN_STAMP).  usage: python tools/gnn_trace.py trace.bin"""
import struct
import sys

import numpy as np

raw = open(sys.argv[1], "rb").read()
n1s, n2s, rows, words = struct.unpack("4i", raw[:16])
t = np.frombuffer(raw[16:], dtype=np.uint64).astype(np.int64)
S = n1s + n2s
ISS, PROD, EPI = 0, 4096, 8192
iss = t[ISS:ISS + 8 * 3 * S].reshape(3, S, 8)
t0 = iss[0, 0, 0]
print(f"slabs: {n1s} GEMM1 + {n2s} GEMM2 per tile; cycles relative to the first traced slab")
for it in range(3):
    print(f"--- tile {it}")
    for j in range(S):
        a = iss[it, j] - t0
        kind = "G1" if j < n1s else "G2"
        print(f"{kind} slab {j:2d}: start {a[0]:7d} | wait B {a[1]-a[0]:5d} | wait accfree {a[2]-a[1]:5d} | wait A0 {a[3]-a[2]:5d} | issue {a[4]-a[3]:4d} | "
              f"wait A1 {a[5]-a[4]:5d} | issue {a[6]-a[5]:4d} | copies {a[7]-a[6]:4d} | total {a[7]-a[0]:5d}")
    e = t[EPI + 4 * it: EPI + 4 * it + 4] - t0
    print(f"epilogue: acc2 full seen at {e[0]}, acc2 released +{e[1]-e[0]}, pass A done +{e[2]-e[0]}, pass B done +{e[3]-e[0]}")
    tile = iss[it, S - 1, 7] - iss[it, 0, 0]
    w = iss[it]
    print(f"tile time {tile}: wait B {int((w[:,1]-w[:,0]).sum())}, wait accfree {int((w[:,2]-w[:,1]).sum())}, wait A {int((w[:,3]-w[:,2]).sum()+(w[:,5]-w[:,4]).sum())}, "
          f"issue {int((w[:,4]-w[:,3]).sum()+(w[:,6]-w[:,5]).sum())}, copies {int((w[:,7]-w[:,6]).sum())}")
pr = t[PROD:PROD + 2 * 3 * S] - t0
print("producer publish times (unit: cycles):")
for it in range(3):
    print(" tile", it, [int(x) for x in pr[2 * it * S: 2 * (it + 1) * S]])


x = t[EPI + 32: EPI + 48]
if x[0]:
    names = ["TMEM wait returned", "ReLU done", "next request issued", "statistics done", "half 0 stored", "half 1 stored"]
    for ch in (3, 4):
        print(f"pass A, chunk {ch} of the first traced tile (stamps; arithmetic may float across them):")
        for i, nm in enumerate(names):
            print(f"  {nm:22s} +{int(x[8 * (ch - 3) + i] - x[0])}")
