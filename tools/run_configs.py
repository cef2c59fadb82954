"""Runs the BASELINE.json configurations other than the bench workload and prints one JSON line each:
 A  main(ip) for one replicate (configs[0]): latency on a kept handle, state matrix read back
 C  contact tracing ctstrat 1 and 3 (multikernel path), 1000 replicates
 D  calibration beta sweep: 20 beta points x 500 replicates, T = 50 (fused path)
 E  1,000,000 agents x 500 days x R replicates with a quarantined age group (multikernel path, native schedule)
Usage: python tools/run_configs.py [A] [C] [D] [E] [--e-reps 100] [--profile]"""
import json, sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import covid19abm_b200 as cv

which = [a for a in sys.argv[1:] if a in ("A", "C", "D", "E")] or ["A", "C", "D", "E"]
e_reps = int(sys.argv[sys.argv.index("--e-reps") + 1]) if "--e-reps" in sys.argv else 100
profile = "--profile" in sys.argv
peak = 6546.2


def run(name, ip, nrep, pof=None, mode=cv.MODE_EXACT, path="auto", store_hm=False, warm=True):
    first = ip[0] if isinstance(ip, list) else ip
    pop = cv.Population(nrep, first.popsize, first.modeltime, store_hmatrix=store_hm)
    pop.set_path(path)
    seeds = [726 * (i + 1) for i in range(nrep)]
    if warm:
        pop.configure(ip, seeds, pset_of_rep=pof, mode=mode); pop.run()
    pop.set_profiling(profile)
    pop.configure(ip, [s + 7 for s in seeds], pset_of_rep=pof, mode=mode)
    t0 = time.perf_counter(); pop.run(); wall = time.perf_counter() - t0
    ms = pop.last_run_ms
    sc = pop.scalars()
    inc, prev = pop.curves()
    units = nrep * first.popsize * first.modeltime
    out = {"config": name, "path": pop.last_path, "mode": "native" if mode else "exact", "replicates": nrep, "popsize": first.popsize,
           "modeltime": first.modeltime, "device_ms": round(ms, 3), "agent_days_per_sec": units / (ms * 1e-3), "replicates_per_sec": nrep / (ms * 1e-3),
           "errors": int((sc["err"] != 0).sum()), "attack_mean": float(1 - prev[:, 0, -1, 0].mean() / first.popsize),
           "hmatrix_in_hbm": store_hm, "columns_sum_to_N": bool((prev[:, 0].sum(axis=2) == first.popsize).all())}
    if profile:
        pr = pop.profile()
        out["kernels"] = {k: {"ms": round(v[0], 3), "launches": v[1]} for k, v in pr.items() if v[1]}
        if pr["time_update"][1]:
            ct = any(p.ctstrat for p in (ip if isinstance(ip, list) else [ip]))
            # k_time_update: next-event day 2 r (+ previous column 1 r + next column 1 w; + 0.25 B of mark bits with tracing)
            bpd = 2 + (2 if store_hm else 0) + (0.25 if ct else 0)
            tu = pr["time_update"]
            gbs = bpd * nrep * first.popsize / (tu[0] / tu[1] * 1e-3) / 1e9
            out["time_update_roofline"] = {"bytes_per_agent_day": bpd, "achieved_GBps": round(gbs, 1), "frac_of_measured_hbm_peak": round(gbs / peak, 4)}
    if name.startswith("D"):
        out["lat_inc_sum_mean_by_beta"] = [float(sc["lat_inc_sum"][np.array(pof) == k].mean()) for k in range(len(ip))]
    if any(p.ctstrat for p in (ip if isinstance(ip, list) else [ip])):
        out["totalisolated_mean"] = float(sc["totalisolated"].mean())
    print(json.dumps(out), flush=True)
    pop.close()


B = dict(prov="ontario", initialinf=1, modeltime=500, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3, tpreiso=0)
if "A" in which:
    ipa = cv.ModelParameters(β=0.0345, prov="ontario", initialinf=1, modeltime=500)
    with cv.Population(1, ipa.popsize, ipa.modeltime, store_hmatrix=True) as pop:
        dev, e2e, att = [], [], []
        for k in range(41):
            t0 = time.perf_counter()
            pop.configure(ipa, [726 * (k + 1)]); pop.run(sync=False); hm = pop.hmatrix(0)      # runsim's seed :78, main's return value :141
            e2e.append(1e3 * (time.perf_counter() - t0)); dev.append(pop.last_run_ms); att.append(float((hm[:, -1] != 0).mean()))
        dev, e2e, att = np.array(dev[1:]), np.array(e2e[1:]), np.array(att[1:])
        big = att > 0.1
        print(json.dumps({"config": "A main(ip), one replicate, 10000 agents x 500 days, no isolation", "path": pop.last_path, "seeds": 40,
                          "device_ms_median": round(float(np.median(dev)), 3), "device_ms_max": round(float(dev.max()), 3),
                          "host_to_host_ms_median (configure + run + Matrix{Int16} read back)": round(float(np.median(e2e)), 3), "host_to_host_ms_max": round(float(e2e.max()), 3),
                          "epidemics_that_take_off": int(big.sum()), "device_ms_when_taking_off_median": round(float(np.median(dev[big])), 3) if big.any() else None,
                          "device_ms_when_dying_out_median": round(float(np.median(dev[~big])), 3) if (~big).any() else None}), flush=True)
if "C" in which:
    run("C ctstrat=1", cv.ModelParameters(β=0.0345, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3, **B), 1000)
    run("C ctstrat=3 strat3qdays=14", cv.ModelParameters(β=0.0345, ctstrat=3, strat3qdays=14, fctcapture=0.5, fcontactst=0.5, cdaysback=3, **B), 1000)
    run("C ctstrat=1 (per-day kernels)", cv.ModelParameters(β=0.0345, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3, **B), 1000, path="multikernel")
    run("B (per-day kernels)", cv.ModelParameters(β=0.0345, **B), 1000, path="multikernel", store_hm=True)
    run("C ctstrat=1, beta=0.06 no isolation (larger epidemics)", cv.ModelParameters(β=0.06, prov="ontario", initialinf=1, modeltime=500, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3), 1000)
    run("C ctstrat=1 native", cv.ModelParameters(β=0.0345, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3, **B), 1000, path="multikernel", mode=cv.MODE_NATIVE)
if "D" in which:
    betas = np.linspace(0.0225, 0.085, 20)                     # scripts/local_run.jl:165 range, 20 points
    psets = [cv.ModelParameters(β=float(b), prov="ontario", calibration=True, modeltime=50, initialinf=1) for b in betas]
    pof = [i % 20 for i in range(10000)]                        # interleaved so every GPU/CTA sees every beta
    run("D calibration sweep 20 beta x 500", psets, 10000, pof=pof)
if "E" in which:
    ip = cv.ModelParameters(β=0.0345, prov="ontario", popsize=1000000, modeltime=500, initialinf=100, quar_grp_frac=0.5, quar_grp=5)
    run(f"E 1e6 agents x 500 d x {e_reps} native", ip, e_reps, mode=cv.MODE_NATIVE, path="multikernel", store_hm=False, warm=False)
