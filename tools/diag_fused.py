"""Per-phase cycle counts of k_sim_fused (diagnostics build: make -C covid19abm.jl_b200/csrc diag).
Usage: python tools/diag_fused.py [B|REF|HOT|C1|E] [nrep]
  B    bench workload (BASELINE configs[1])          REF  the reference's timed scenario (test/runtests.jl:408-428)
  HOT  beta=0.06 ontario, no isolation, no tracing   C1   beta=0.06 with contact tracing ctstrat=1
  E    10^6 agents (BASELINE configs[4]; the fused kernel with per-agent arrays in global memory)"""
import os, sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
os.environ.setdefault("CABM_LIB", os.path.abspath("covid19abm.jl_b200/libcabm_b200_diag.so"))
import numpy as np, covid19abm_b200 as cv
from bench import CFG, seeds_for
which = sys.argv[1] if len(sys.argv) > 1 else "B"
R = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
cfgs = {"B": CFG,
        "REF": dict(β=0.0525, prov="newyork", modeltime=300),
        "HOT": dict(β=0.06, prov="ontario", initialinf=1, modeltime=500),
        "C1": dict(β=0.06, prov="ontario", initialinf=1, modeltime=500, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3),
        "E": dict(β=0.0345, prov="ontario", popsize=1000000, modeltime=500, initialinf=100, quar_grp_frac=0.5, quar_grp=5)}
ip = cv.ModelParameters(**cfgs[which])
T = ip.modeltime
NPOP = ip.popsize
pop = cv.Population(R, NPOP, T, store_hmatrix=False)
pop.configure(ip, seeds_for(0, 0, R)); pop.run(); pop.run()
print("config", which, "ms", pop.last_run_ms)
met = np.zeros(R, np.int32); inf = np.zeros(R, np.int32)
pop._chk(pop.L.cabm_get_dyntrans_counters(pop.h, met.ctypes.data, inf.ctypes.data))
inc, prev = pop.curves()
attack = 1 - prev[:, 0, -1, 0] / NPOP
order = np.argsort(-met)
print("kcycles sum", met.sum(), "max", met.max(), "mean", met.mean())
for r in order[:6]: print(r, "kcyc", met[r], "ms", met[r] * 1024 / 1.965e6, "chunks", inf[r], "attack", attack[r], "peak inf", prev[r, 0, :, 2:8].sum(axis=1).max(), "cyc/chunk", met[r] * 1024 / max(inf[r], 1))
print("takeoff count", (attack > 0.1).sum(), "mean attack", attack.mean())
dbg = np.zeros((R, 32), np.int64); pop._chk(pop.L.cabm_get_debug(pop.h, dbg.ctypes.data))
names = {22: "setup: initialize", 23: "setup: get_ag_dist", 24: "setup: insert_infected", 0: "setup: column 1 / resume reload", 1: "list",
         8: "chunk: budgets + prefix -> C1", 9: "chunk: event generation (rest)", 2: "  gen: owners", 3: "  gen: draws (philox)", 7: "  gen: grp + status loads", 20: "  gen: fresh budgets", 21: "  gen: marks/edges", 10: "chunk: C2 wait + DAG", 11: "chunk: apply singles + count shared",
         13: "chunk: C3 + compaction", 14: "chunk: C4 + resolve shared", 15: "chunk: C5 + link/clear", 12: "dyntrans end -> S1",
         27: "time_update: S1 -> scan start (nf reset, pass A)", 30: "time_update: scan + list", 31: "time_update: list barrier", 25: "time_update: listed events", 26: "time_update: min-reduce", 4: "time_update: S2 wait", 5: "curves/idle", 6: "writeback",
         28: "(sum over threads) cycles inside agent_event_core", 29: "(count) time_update events", 16: "(count) events", 17: "(count) shared-target events", 18: "(count) DAG edges", 19: "(count) DAG sweeps"}
big = attack > 0.1
print("phase kcycles  take-off mean | extinct mean | per chunk (take-off)")
nr = max(inf[big].sum(), 1)
for k in names: print(f"{k:2d} {names[k]:36s} {dbg[big,k].mean()/1e3:10.1f} | {dbg[~big,k].mean()/1e3:10.1f} | {dbg[big,k].sum()/nr:8.0f}")
print("chunks take-off mean", inf[big].mean(), "extinct", inf[~big].mean())
print("take-off total kcycles mean", met[big].mean(), "-> ms", met[big].mean() * 1024 / 1.965e6, " max ms", met.max() * 1024 / 1.965e6)
