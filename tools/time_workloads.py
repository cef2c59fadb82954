"""Device time of the fused path on a few workloads: one launch alone, and 12 steps pipelined over three handles.
Usage: python tools/time_workloads.py [B REF HOT C1 ...]"""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import json
import numpy as np, covid19abm_b200 as cv
from bench import CFG, seeds_for
cfgs = {"B": CFG,
        "REF": dict(β=0.0525, prov="newyork", modeltime=300),
        "REF500": dict(β=0.0525, prov="newyork", modeltime=500),
        "HOT": dict(β=0.06, prov="ontario", initialinf=1, modeltime=500),
        "C1": dict(β=0.06, prov="ontario", initialinf=1, modeltime=500, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3),
        "CB1": dict(CFG, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3)}
R = 1000
HM = "nohm" not in sys.argv
for which in ([a for a in sys.argv[1:] if a != "nohm"] or ["B", "REF", "HOT", "C1"]):
    ip = cv.ModelParameters(**cfgs[which])
    T = ip.modeltime
    pipe = cv.EnsemblePipeline(R, 10000, T, handles=3, store_hmatrix=HM, pinned=False)
    pops = pipe.pops
    for w in range(3):
        pops[w].configure(ip, seeds_for(1000 + w, 0, R)); pops[w].run()
    solos = []
    for w in range(3):
        pops[0].configure(ip, seeds_for(2000 + w, 0, R)); pops[0].run(); solos.append(pops[0].last_run_ms)
    solo = min(solos)
    K = 12
    pops[0].mark()
    for k in range(K):
        q = pops[k % 3]; q.configure(ip, seeds_for(k, 0, R)); q.run(sync=False)
    ms = max(q.ms_since_mark(pops[0]) for q in pops)
    inc, prev = pops[0].curves()
    attack = 1 - prev[:, 0, -1, 0] / 10000
    print(json.dumps({"workload": which, "hm": HM, "solo_ms": round(solo, 3), "pipelined_ms_per_step": round(ms / K, 3), "attack_mean": float(attack.mean()),
                      "takeoff_frac": float((attack > 0.1).mean()), "agent_days_per_sec_pipelined": R * 10000 * T / (ms / K * 1e-3)}), flush=True)
    pipe.close()
