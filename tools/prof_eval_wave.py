"""Phases of one wave of the evaluation driver (32 instances x 100 chains x 10^5 steps with trajectories, each instance on its
own CUDA stream): setup, launch, wait + summaries, trajectory statistics, teardown.  Used to find that 32 streams on the default
8 hardware queues ran 8 launches at a time (CUDA_DEVICE_MAX_CONNECTIONS).  usage: python tools/prof_eval_wave.py"""
import os, sys, time, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g; g.build()
import sls_sat_solving_with_deep_learning_b200 as eng
from sls_sat_solving_with_deep_learning_b200 import synth, evaluate
from sls_sat_solving_with_deep_learning_b200.solver import Instance, ResidentSolver
eng.set_device(0)
insts=[]
for i in range(32):
    row,lits=synth.random_ksat_csr(200,860,3,1000+i); insts.append(Instance.from_csr(200,row,lits))
streams=evaluate._streams(32)
w=np.ones(200)/2
for rep in range(2):
    t0=time.perf_counter()
    solvers=[ResidentSolver(inst,"probsat",100000-1,100,True,True,0.0,seed=5+k) for k,inst in enumerate(insts)]
    for s in solvers: s.set_weights(w,w)
    t1=time.perf_counter()
    for k,s in enumerate(solvers): s.launch(streams[k])
    t2=time.perf_counter()
    res=[s.fetch_summary() for s in solvers]
    t3=time.perf_counter()
    st=[s.trajectory_stats() for s in solvers]
    t4=time.perf_counter()
    for s in solvers: s.close()
    t5=time.perf_counter()
    print(f"setup {t1-t0:.3f} launch {t2-t1:.3f} fetch {t3-t2:.3f} stats {t4-t3:.3f} close {t5-t4:.3f}", [round(r.kernel_ms,1) for r in res[:4]])
