import os, subprocess, sys; sys.path.insert(0,'.'); sys.path.insert(0,'tests')
# needs the diagnostics build of the library: make -C covid19abm.jl_b200/csrc diag
os.environ.setdefault("CABM_LIB", os.path.abspath("covid19abm.jl_b200/libcabm_b200_diag.so"))
import numpy as np, covid19abm_b200 as cv
from bench import CFG, seeds_for
ip = cv.ModelParameters(**CFG)
pop = cv.Population(1000, 10000, 500, store_hmatrix=False)
pop.configure(ip, seeds_for(0,0,1000)); pop.run(); pop.run()
print("ms", pop.last_run_ms)
met = np.zeros(1000, np.int32); inf = np.zeros(1000, np.int32)
pop._chk(pop.L.cabm_get_dyntrans_counters(pop.h, met.ctypes.data, inf.ctypes.data))
inc, prev = pop.curves()
attack = 1 - prev[:,0,-1,0]/10000
order = np.argsort(-met)
print("kcycles sum", met.sum(), "max", met.max(), "mean", met.mean())
for r in order[:12]: print(r, "kcyc", met[r], "ms", met[r]*1024/1.965e6, "rounds", inf[r], "attack", attack[r], "peak inf", prev[r,0,:,2:8].sum(axis=1).max(), "cyc/round", met[r]*1024/max(inf[r],1))
print("takeoff count", (attack>0.1).sum(), "mean attack", attack.mean())
for r in order[500:503]: print(r, "kcyc", met[r], "rounds", inf[r], "attack", attack[r])

dbg = np.zeros((1000, 16), np.int64); pop._chk(pop.L.cabm_get_debug(pop.h, dbg.ctypes.data))
names = ["setup","list","budgets","(rounds tail)","time_update after S1","curves/idle","writeback","-","round: warp0 cand ->B1","round: events ->B2","round: dup compaction ->B3","round: apply ->B5","dyntrans end -> S1"]
big = attack > 0.1
print("phase kcycles  take-off mean | extinct mean")
for k in range(13): print(f"{names[k]:28s} {dbg[big,k].mean()/1e3:10.1f} | {dbg[~big,k].mean()/1e3:10.1f}")
print("rounds take-off mean", inf[big].mean(), "extinct", inf[~big].mean(), " active days?")
