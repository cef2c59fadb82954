"""Experiment: consecutive 1000-replicate batches on two (or three) handles with their own streams — the short replicates of batch k+1
fill the CTA slots that the long epidemics of batch k leave idle."""
import sys, time; sys.path.insert(0, '.')
import numpy as np, covid19abm_b200 as cv
from bench import CFG, seeds_for, N_AGENTS, T_DAYS
R = 1000
ip = cv.ModelParameters(**CFG)
for nh in (1, 2, 3):
    pops = [cv.Population(R, N_AGENTS, T_DAYS, store_hmatrix=True) for _ in range(nh)]
    for k in range(3 * nh):
        p = pops[k % nh]; p.configure(ip, seeds_for(k, 0, R)); p.run(sync=False)
    for p in pops: p.sync()
    K = 24
    t0 = time.perf_counter()
    for k in range(K):
        p = pops[k % nh]; p.configure(ip, seeds_for(100 + k, 0, R)); p.run(sync=False)
    for p in pops: p.sync()
    dt = 1e3 * (time.perf_counter() - t0) / K
    print("handles", nh, "ms per step", round(dt, 3), "last_run_ms", [round(p.last_run_ms, 3) for p in pops])
    for p in pops: p.close()
