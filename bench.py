#!/usr/bin/env python3
"""bench.py — agent-days/s of covid19abm.jl's simulation path (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W                 # this repo's CUDA path, workload B (the headline)
    python bench.py --config {B,REF,C1,C3,D,E} ...                # the other BASELINE configurations, same JSON keys
    python bench.py --impl reference --gpus N --steps K ...       # the CPU restatement of the reference (oracle port)

Workloads (config.workload):
  B    BASELINE.json configs[1] — the 10,000-agent x 500-day Ontario scenario with isolation on (fmild=0.5, τmild=1,
       fsevere=1.0, fpreiso=0.3, tpreiso=0), 1000 replicates per GPU (weak scaling). The metric is quoted on this one.
  REF  the reference's own timed scenario (test/runtests.jl:408-428, README 2.13 s per replicate): β=0.0525, :newyork,
       no interventions, modeltime 300 (the source's default), 1000 replicates per GPU. Supercritical: 3 in 4 replicates take off.
  C1/C3  configs[2]: B plus contact tracing (fctcapture=0.5, fcontactst=0.5, cdaysback=3), ctstrat 1 / ctstrat 3 (strat3qdays=14)
  D    configs[3]: calibration sweep, 10,000 replicates x 50 days over a 20-point β grid, sharded over the GPUs (strong scaling)
  E    configs[4]: 1,000,000 agents x 500 days, sheltered age group, 100 replicates sharded over the GPUs (strong scaling)

A step = one full pass of the hot path (main(ip) :104-142 for every replicate of the batch: initialize, seed, modeltime x
(dyntrans, time_update + snapshot), reductions). Replicates are independent and shard over ranks with no data-path collective;
the one collective of the path — the end-of-run gather of the summary curves (cv.gather_ensemble, NCCL all_gather over NVLink) —
runs once after the timed region and its time is reported.

Steps alternate over `--pipeline` handles of the GPU (default 3; cv.EnsemblePipeline): each handle has its own stream, state
and pinned result buffers, so the replicates of step k+1 that die out run beside the few long epidemics of step k. Every step
is one full pass over its own replicates; all K steps lie inside the timed region.

`value` is device time (CUDA events on the handles' streams) with parameters and seeds resident; `e2e` goes through the public
API from host parameter structs to host result buffers (curves + scalars as runsim returns them, :77-101), copies included.
After the timed region a sample of the *timed* outputs is compared with the oracle bit for bit ("verified").
The Julia reference cannot run in this image (no julia): the CPU arm is the C restatement in oracle/ ("port").
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

N_AGENTS, T_DAYS = 10000, 500
CFG = dict(β=0.0345, prov="ontario", initialinf=1, modeltime=T_DAYS, popsize=N_AGENTS,
           fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3, tpreiso=0)
WORKLOAD = "configs[1]: Ontario beta=0.0345, 10000 agents x 500 days, isolation on (fmild=0.5 taumild=1 fsevere=1.0 fpreiso=0.3 tpreiso=0)"
METRIC, UNIT = "agent_days_per_sec", "agent-days/s"
_TRACE = dict(fctcapture=0.5, fcontactst=0.5, cdaysback=3)


def workload(name):
    """-> dict(desc, psets (list of kwargs), pof_cycle, N, T, R (per GPU if weak, total if strong), scaling, mode, hm)"""
    if name == "B":
        return dict(desc=WORKLOAD, psets=[CFG], N=N_AGENTS, T=T_DAYS, R=1000, scaling="weak", mode="exact", hm=True)
    if name == "REF":
        return dict(desc="reference's timed scenario (test/runtests.jl:408-428): newyork beta=0.0525, no interventions, 10000 agents x 300 days",
                    psets=[dict(β=0.0525, prov="newyork", modeltime=300, popsize=N_AGENTS)], N=N_AGENTS, T=300, R=1000, scaling="weak", mode="exact", hm=True)
    if name in ("C1", "C3"):
        ct = dict(ctstrat=1) if name == "C1" else dict(ctstrat=3, strat3qdays=14)
        return dict(desc=f"configs[2]: configs[1] + contact tracing {ct} fctcapture=0.5 fcontactst=0.5 cdaysback=3 (ctstrat=2 does not exist in the source, :165)",
                    psets=[dict(CFG, **_TRACE, **ct)], N=N_AGENTS, T=T_DAYS, R=1000, scaling="weak", mode="exact", hm=True)
    if name == "C1HOT":          # tools only: tracing under a supercritical epidemic (the stress case of profiles/r01_configs_CDE.jsonl)
        return dict(desc="beta=0.06 ontario, no isolation, contact tracing ctstrat=1 fctcapture=0.5 fcontactst=0.5 cdaysback=3, 10000 agents x 500 days",
                    psets=[dict(β=0.06, prov="ontario", initialinf=1, modeltime=T_DAYS, popsize=N_AGENTS, ctstrat=1, **_TRACE)], N=N_AGENTS, T=T_DAYS, R=1000, scaling="weak", mode="exact", hm=True)
    if name == "D":
        import numpy as np
        betas = np.linspace(0.0225, 0.085, 20)                   # scripts/local_run.jl:165 range, 20 points
        return dict(desc="configs[3]: calibration sweep, 10000 replicates x 50 days over 20 beta in [0.0225, 0.085], ontario, sharded over the GPUs",
                    psets=[dict(β=float(b), prov="ontario", calibration=True, modeltime=50, initialinf=1, popsize=N_AGENTS) for b in betas],
                    N=N_AGENTS, T=50, R=10000, scaling="strong", mode="exact", hm=False)
    if name == "E":
        return dict(desc="configs[4]: 1,000,000 agents x 500 days, quar_grp_frac=0.5 quar_grp=5, initialinf=100, 100 replicates sharded over the GPUs",
                    psets=[dict(β=0.0345, prov="ontario", popsize=1000000, modeltime=500, initialinf=100, quar_grp_frac=0.5, quar_grp=5)],
                    N=1000000, T=500, R=100, scaling="strong", mode=os.environ.get("CABM_BENCH_E_MODE", "exact"), hm=False)
    raise SystemExit(f"unknown --config {name}")


def seeds_for(step, rank, nrep):
    base = 1 + (step * 64 + rank) * 100000
    return [726 * (base + i) for i in range(nrep)]


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons during the timed region (B200_PROFILING.md recipe): NVML every 5 ms, or the
    nvidia-smi query of the recipe when pynvml is not importable."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag, self.max_mhz, self.source = index, [], False, None, "nvidia-smi"
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
            self.source = "nvml"
        except Exception:
            self.nvml = None

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    n = self.nvml
                    mhz = n.nvmlDeviceGetClockInfo(self.dev, n.NVML_CLOCK_SM)
                    rs = n.nvmlDeviceGetCurrentClocksEventReasons(self.dev) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") else n.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                    bits = [getattr(n, "nvmlClocksEventReasonHwSlowdown", 0x8), getattr(n, "nvmlClocksEventReasonHwThermalSlowdown", 0x40),
                            getattr(n, "nvmlClocksEventReasonSwThermalSlowdown", 0x20), getattr(n, "nvmlClocksEventReasonSwPowerCap", 0x4)]
                    self.samples.append([mhz] + [bool(rs & b) for b in bits])
                    time.sleep(0.005)
                else:
                    out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                         capture_output=True, text=True, timeout=5).stdout.strip()
                    if out:
                        f = [x.strip() for x in out.split(",")]
                        if f[0].isdigit():
                            self.max_mhz = int(f[1]) if f[1].isdigit() else self.max_mhz
                            self.samples.append([int(f[0])] + [x.lower().startswith("active") for x in f[2:6]])
                    time.sleep(0.05)
            except Exception:
                time.sleep(0.05)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = [n for k, n in enumerate(self.NAMES) if any(len(s) > 1 + k and s[1 + k] for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(self.samples), "source": self.source}


def oracle_psets(W):
    import oracle_api as O
    from test_gpu_parity import to_oracle
    import covid19abm_b200 as cv
    return O, [to_oracle(O, cv.ModelParameters(**kw)) for kw in W["psets"]]


def cpu_sample(W, nrep, threads):
    """Times the oracle port (oracle/abm_oracle.c) on `nrep` replicates of the workload with `threads` workers."""
    O, ops = oracle_psets(W)
    seeds = seeds_for(0, 999, nrep)
    pof = [i % len(ops) for i in range(nrep)]
    t0 = time.perf_counter()
    O.run_ensemble(ops, pof, seeds, nthreads=threads)
    dt = time.perf_counter() - t0
    return nrep * W["N"] * W["T"] / dt, dt


def cpu_sample_size(W, threads):
    """About 10-30 s of single-core work: replicates of the workload scaled to its cost per replicate."""
    per_rep = W["N"] * W["T"] / 2.5e7                      # ~2.5e7 agent-days/s per core for the port
    return max(threads, min(6 * threads, int(round(30 / max(per_rep, 1e-3))))) if W["N"] <= 100000 else max(2, threads // 4)


def run_reference(args, W, rank, world, emit):
    """--impl reference: the reference's CPU implementation of the path. Julia is absent from the image, so this is
    the oracle port (kind "port"), all host threads, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    nrep = max(threads, 16) if W["N"] <= 100000 else max(2, threads // 4)
    for _ in range(args.warmup):
        cpu_sample(W, min(nrep, threads), threads)
    tot = 0.0
    for _ in range(args.steps):
        _, dt = cpu_sample(W, nrep, threads)
        tot += dt
    value = args.steps * nrep * W["N"] * W["T"] / tot
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": W["scaling"],
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": W["desc"], "replicates_per_step": nrep, "note": "CPU restatement of covid19abm.jl in C (oracle/), not Julia: no julia binary in the image"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"{nrep} replicates per step, one per thread"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="B", choices=["B", "REF", "C1", "C3", "D", "E"], help="BASELINE configuration (default B: the one the metric is quoted on)")
    ap.add_argument("--replicates", type=int, default=None, help="replicates per GPU per step (weak) or in total (strong); default: the workload's")
    ap.add_argument("--no-hmatrix", action="store_true", help="do not keep the state matrices in HBM (curves only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle check of the timed outputs")
    ap.add_argument("--pipeline", type=int, default=3, help="handles (CUDA streams) the steps alternate over; 1 = one step after the other")
    ap.add_argument("--path", default="auto", choices=["auto", "multikernel", "fused", "fused-global"], help="execution path (include/cabm.h CABM_PATH_*)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    W = workload(args.config)
    if args.steps is None:
        args.steps = {"B": 50, "REF": 20, "C1": 30, "C3": 30, "D": 30, "E": 3}[args.config]
    if args.impl == "reference" and args.steps > 3 and W["N"] > 100000:
        args.steps = 3

    # the contract is ONE JSON line on stdout: libraries (NCCL's banner and INFO log, torchrun) also write there, so everything
    # else goes to stderr (dup2) and the line is written to the saved descriptor at the end. NCCL_DEBUG is left as the pool sets it.
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, W, rank, world, emit)
        return

    import numpy as np
    import torch
    import covid19abm_b200 as cv

    numa_bound = cv.bind_host_thread_to_gpu(local) if world > 1 else False     # before the pinned buffers are allocated
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    def max_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(x):
        if dist is None:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    N, T = W["N"], W["T"]
    Rtot = args.replicates if args.replicates is not None else W["R"]
    if W["scaling"] == "weak":
        R, my = Rtot, None                              # R replicates on every GPU
    else:
        my = cv.shard_indices(Rtot, rank, world)        # replicate r runs on rank r mod world
        R = len(my)
    Rmax = Rtot if W["scaling"] == "weak" else (Rtot + world - 1) // world
    psets = [cv.ModelParameters(**kw) for kw in W["psets"]]
    mode = cv.MODE_NATIVE if W["mode"] == "native" else cv.MODE_EXACT
    hm_on = W["hm"] and not args.no_hmatrix

    def batch_of(step):
        """(seeds, pset_of_rep) of this rank's replicates of one step"""
        if my is None:
            return seeds_for(step, rank, R), [i % len(psets) for i in range(R)]
        allseeds = seeds_for(step, 0, Rtot)
        return [allseeds[i] for i in my], [i % len(psets) for i in my]

    P = max(1, args.pipeline)
    pipe = cv.EnsemblePipeline(max(R, 1), N, T, handles=P, store_hmatrix=hm_on, device=local, path=args.path, copy=False, pinned=False)
    pops, pop = pipe.pops, pipe.pops[0]
    units = R * N * T                                   # agent-days per step on this GPU

    for w in range(max(args.warmup, P)):
        q = pops[w % P]
        sd, pof = batch_of(1000 + w)
        q.configure(psets, sd, pset_of_rep=pof, mode=mode)
        q.run()
    for q in pops:
        q.check_errors()
    one_ms = pop.last_run_ms                           # one step alone on the GPU (nothing else in flight)
    fused = pop.last_path in ("fused", "fused-global")

    sampler = ClockSampler(local)
    sampler.start()
    # ---- value: device-resident, CUDA-event time of exactly K steps (mark on the first handle's stream -> end of every handle's last run)
    barrier()
    l0 = sum(q.launch_count for q in pops)
    t0 = time.perf_counter()
    pop.mark()
    for k in range(args.steps):
        q = pops[k % P]
        sd, pof = batch_of(k)
        q.configure(psets, sd, pset_of_rep=pof, mode=mode)   # seeds + thresholds: ~10 KB, resident before the run that uses them
        q.run(sync=False)
    dev_ms = max(q.ms_since_mark(pop) for q in pops[:min(P, args.steps)])
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = sum(q.launch_count for q in pops) - l0
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)

    # ---- e2e: host params -> host results through the public API (cv.EnsemblePipeline.submit), copies inside the timed region.
    # Result form: Int16 incidence/prevalence curves [R][6][T][12] in pinned host arrays + per-replicate scalars. On the fused path
    # the kernel streams each finished replicate's curves over PCIe itself, and only the rows before the replicate's epidemic died out
    # (cabm_set_host_curve_rows: the rest repeat — incidence 0, prevalence of the last row sent; rows[r] tells how many). This
    # rows-only form is lossless (tests/test_full_size_parity.py::test_config_b_rows_only_transport_is_lossless) and is what e2e
    # times; `e2e.dense` repeats the loop with every row sent.
    direct = N <= 32767
    shape = (max(R, 1), 6, T, 12)
    slots = []
    for j in range(P):
        if direct:
            slots.append((cv.PinnedArray(shape, np.int16), cv.PinnedArray(shape, np.int16), cv.PinnedArray((max(R, 1),), np.int32)))
        else:
            slots.append(None)

    def e2e_loop(rows_only):
        barrier()
        t0 = time.perf_counter()
        got, last = 0, None
        for k in range(args.steps):
            sd, pof = batch_of(100 + k)
            s = slots[k % P]
            out = None if s is None else ((s[0].array, s[1].array, s[2].array) if rows_only else (s[0].array, s[1].array))
            r = pipe.submit(psets, sd, pset_of_rep=pof, mode=mode, out=out, tag=k)
            got += r is not None
        rest = pipe.drain()
        got += len(rest)
        barrier()
        ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
        assert got == args.steps
        return ms, rest[-1]

    rows_ok = direct and fused and T % 2 == 0
    e2e_ms, last = e2e_loop(rows_only=rows_ok)
    assert (last["err"] == 0).all()
    # the result arrays are views of the slot's pinned buffers: keep a dense copy of the last timed step for the checks below
    inc_l, prev_l = np.array(last["inc"][:R]), np.array(last["prev"][:R])
    scal_l = {k: np.array(last[k][:R]) for k in ("infectors", "totalisolated", "lat_inc_sum", "err")}
    if rows_ok:
        cv.expand_curves(inc_l, prev_l, np.array(last["rows"][:R]))
    attack_mean = float(1.0 - prev_l[:, 0, -1, 0].mean() / N) if R else None
    scal_bytes = R * (32 + 8 + 8 + 4)
    if rows_ok:
        nb_ = min(P, args.steps)
        sent_rows = sum(int(s[2].array[:R].sum()) for s in slots[:nb_]) / nb_
        d2h = int(sent_rows * 6 * 12 * 2 * 2 + R * 4 + scal_bytes)
    elif direct:
        d2h = 2 * int(np.prod(shape)) * 2 + scal_bytes
    else:
        d2h = 2 * int(np.prod(shape)) * 4 + scal_bytes
    h2d = R * 8 + R * 4 + len(psets) * (280 + 4 * T * 8)

    # ---- verify: a sample of the outputs of the LAST TIMED e2e step against the oracle, bit for bit
    verified = None
    if not args.no_verify and mode == cv.MODE_EXACT and N <= 100000 and R > 0:
        sd, pof = batch_of(100 + args.steps - 1)
        attack = 1.0 - prev_l[:, 0, -1, 0] / N
        order = np.argsort(-attack, kind="stable")
        pick = sorted(set(order[:32].tolist() + list(range(0, R, max(1, R // 48)))[:48]))     # the 32 largest epidemics + 48 spread evenly (>= 64 distinct)
        O, ops = oracle_psets(W)
        ref = O.run_ensemble(ops, [pof[i] for i in pick], [sd[i] for i in pick], nthreads=os.cpu_count() or 1)
        ok = bool((inc_l[pick] == ref["inc"]).all() and (prev_l[pick] == ref["prev"]).all()
                  and (scal_l["infectors"][pick] == ref["infectors"]).all() and (scal_l["totalisolated"][pick] == ref["totalisolated"]).all())
        if not ok:
            raise SystemExit("bench.py: the timed outputs differ from the oracle")
        verified = len(pick)

    dense_ms = None
    if rows_ok:
        dense_ms, _ = e2e_loop(rows_only=False)
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- the path's one collective: end-of-run gather of the summary curves (cv.gather_ensemble; NCCL all_gather over NVLink),
    # on the results of the last rows-only e2e step. Device time by CUDA events around the collectives, wall time for the whole gather.
    gather = None
    if dist is not None:
        nsims = Rtot * world if W["scaling"] == "weak" else Rtot
        loc = dict(inc=inc_l, prev=prev_l, **scal_l)
        barrier()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        full = cv.gather_ensemble(loc, nsims, rank, world, dist=dist, device=f"cuda:{local}", keep_on_device=True, events=(ev0, ev1))
        torch.cuda.synchronize()
        gather = {"collective": "all_gather (NCCL over NVLink) of the Int16 curves + scalars: every rank receives the whole ensemble, replicate order",
                  "replicates": nsims, "bytes_per_rank_in": int(sum(np.asarray(v).nbytes for v in loc.values())),
                  "ms_wall_incl_h2d": round(max_over_ranks(1e3 * (time.perf_counter() - t0)), 3), "ms_device_collectives": round(max_over_ranks(ev0.elapsed_time(ev1)), 3),
                  "checksum_prev": int(full["prev"].sum().item()), "expected_checksum_prev": int(sum_over_ranks(float(prev_l.sum())))}

    # ---- per-kernel-class device time of one more (profiled) step -> roofline of the dominant kernel
    pop.set_profiling(True)
    sd, pof = batch_of(200)
    pop.configure(psets, sd, pset_of_rep=pof, mode=mode)
    pop.run()
    prof = pop.profile()
    pop.set_profiling(False)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    kernels = {}
    for name, (ms, n) in prof.items():
        if n:
            kernels[name] = {"ms_per_step": round(ms, 3), "launches": n, "avg_us": round(1e3 * ms / n, 2)}
    ncu = {}
    npath = os.path.join(ROOT, "profiles", "r02_fused_ncu.json")          # this round's ncu --set full capture of the same command
    if os.path.exists(npath):
        ncu = json.load(open(npath))
    if fused:
        # k_sim_fused: per-agent state lives in shared memory; what must cross HBM per step:
        #   the state matrix, 1 B per agent-day (u8 health code, :221); per agent once: the HBM-resident fields written at
        #   initialize and the final status/expday (26 B); the curves [R][6][T][12] int32 x 2.
        hm_b = R * N * T if hm_on else 0
        alg_bytes = hm_b + R * N * 26 + 2 * R * 6 * T * 12 * 4
        ms, n = prof["fused"]
        gbs = alg_bytes / (ms / n * 1e-3) / 1e9
        cap = ncu.get(args.config, {}) if (Rtot == W["R"] and hm_on == W["hm"]) else {}
        roofline = {"kernel": "k_sim_fused" + ("<state in global memory>" if pop.last_path == "fused-global" else ""), "bound": "latency",
                    "bound_note": "per ncu the kernel is bound by dependent-instruction latency and barriers of the per-replicate day chain, not by a pipe: "
                                  "HBM is the roofline this line reports against (denominator), the kernel moves only the bytes it must",
                    "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s", "frac": round(gbs / peak, 4),
                    "traffic": cap.get("dram_bytes_per_launch"), "peak_source": peak_src,
                    "bytes_per_launch": alg_bytes, "bytes_per_agent_day": round(alg_bytes / max(units, 1), 3),
                    "issue_active_pct": cap.get("issue_active_pct"), "warps_active_pct": cap.get("warps_active_pct"),
                    "barrier_stall_share": cap.get("barrier_stall_share"), "dram_throughput_pct": cap.get("dram_throughput_pct"),
                    "ncu_capture": cap.get("source"),
                    "traffic_eliminated": {"survey_8d_bytes_per_agent_day": 14, "moved_bytes_per_agent_day": round(alg_bytes / max(units, 1), 3),
                                           "note": "SURVEY 8(d) priced reference-width fields streamed from HBM every day (14 B/agent-day); absolute event days, lazy "
                                                   "contact budgets and shared-memory residency remove that traffic (outputs stay bit-identical to the oracle's)"}}
        roofline["launch_ms_alone"] = round(ms / n, 3)
        roofline["achieved_in_timed_region"] = round(alg_bytes * args.steps / (dev_ms * 1e-3) / 1e9, 1)
        roofline["frac_in_timed_region"] = round(alg_bytes * args.steps / (dev_ms * 1e-3) / 1e9 / peak, 4)
        if pop.last_path == "fused-global" and R * N * 12 <= 126e6:
            roofline["bound_note"] += "; state in global memory but L2-resident at this size (%.0f MB of per-agent words): SURVEY 8(d)'s HBM yardstick applies only above L2 capacity" % (R * N * 12 / 1e6)
        if pop.last_path == "fused-global" and R * N * 12 > 126e6:
            # state in global memory (large populations): here the per-agent fields do live in HBM, so the yardstick is SURVEY 8(d)'s own
            # figure, 14 B per agent-day (25 B with contact tracing) at the reference's field widths, over one launch
            per_ad = 25 if any(kw.get("ctstrat", 0) for kw in W["psets"]) else 14
            alg_bytes = per_ad * units + hm_b
            gbs = alg_bytes / (ms / n * 1e-3) / 1e9
            roofline.update({"bound": "hbm", "achieved": round(gbs, 1), "frac": round(gbs / peak, 4), "bytes_per_launch": alg_bytes,
                             "bytes_per_agent_day": round(alg_bytes / max(units, 1), 3),
                             "bound_note": "state in HBM, touched as random 4-byte words (32-byte sectors) by one CTA per replicate: algorithmic bytes = SURVEY 8(d)'s "
                                           "14 B (25 B with tracing) per agent-day; `traffic` above that is sector and line over-fetch of the random accesses. ncu: long "
                                           "scoreboard and barriers are the top stalls — the chain of dependent HBM round trips per chunk, not bandwidth, bounds it",
                             "long_scoreboard_stall_share": cap.get("long_scoreboard_stall_share"),
                             "achieved_in_timed_region": round(alg_bytes * args.steps / (dev_ms * 1e-3) / 1e9, 1),
                             "frac_in_timed_region": round(alg_bytes * args.steps / (dev_ms * 1e-3) / 1e9 / peak, 4)})
            roofline.pop("traffic_eliminated", None)
    else:
        # algorithmic bytes per agent-day of k_time_update in the HBM layout: expday 2 r + previous column 1 r + next column 1 w
        tu_bytes = (2 + (2 if hm_on else 0)) * R * N
        dom = max(("dyntrans", "time_update"), key=lambda k: prof[k][0])
        tu_ms, tu_n = prof["time_update"]
        tu_gbs = tu_bytes / (tu_ms / tu_n * 1e-3) / 1e9 if tu_n else 0.0
        kernels["time_update"]["algorithmic_GBps"] = round(tu_gbs, 1)
        cap = ncu.get(args.config + "_time_update", {})
        roofline = {"kernel": "k_time_update", "bound": "hbm", "achieved": round(tu_gbs, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(tu_gbs / peak, 4), "traffic": cap.get("dram_bytes_per_launch"), "peak_source": peak_src,
                    "bytes_per_agent_day": tu_bytes / max(R * N, 1), "dominant_by_time": "k_" + dom,
                    "note": "time_update is the HBM-streaming per-agent kernel; dyntrans is a latency-bound gather/scatter (see kernels)"}

    cpu = None
    if rank == 0 and not args.no_cpu:
        threads = os.cpu_count() or 1
        nrep = cpu_sample_size(W, threads)
        v, dt = cpu_sample(W, nrep, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{nrep} replicates of the same workload, one per thread, {dt:.1f} s wall (C restatement of the reference, not Julia)"}

    tot_units = sum_over_ranks(float(units))
    if rank == 0:
        value = tot_units * args.steps / (dev_ms * 1e-3)
        e2e = tot_units * args.steps / (e2e_ms * 1e-3)
        reps = tot_units / (N * T)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": W["scaling"], "vs_baseline": None,
                "dtype": "u32", "data": "synthetic",
                "config": {"workload": W["desc"], "name": args.config,
                           ("replicates_per_gpu" if W["scaling"] == "weak" else "replicates_total"): Rtot,
                           "replicates_per_sec": reps * args.steps / (dev_ms * 1e-3),
                           "mode": ("exact: bit-identical to the sequential-order C oracle (keyed Philox draws); distribution-equivalent to the Julia reference, unpinned"
                                    if mode == cv.MODE_EXACT else "native: per-infector parallel with atomics, ensemble-equivalent to the exact schedule"),
                           "path": pop.last_path, "hmatrix_in_hbm": hm_on,
                           "l2": "inputs larger than L2: per-agent state, state matrix and curves of a step are rebuilt every step (B: 0.3 GB + 5 GB + 0.3 GB written per step)",
                           "wall_ms_per_step": wall_ms / args.steps, "attack_rate_mean": attack_mean,
                           "pipeline": f"steps alternate over {P} handles/streams on the GPU: the short replicates of step k+1 run beside the long epidemics of step k",
                           "ms_one_step_alone": one_ms},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                        "returns": ("Int16 incidence+prevalence curves [R][6][T][12] in pinned host arrays, rows-only transport (rows[r] leading rows sent per replicate, the rest repeat; "
                                    "lossless, cv.expand_curves), and per-replicate scalars (runsim :77-101)") if rows_ok else
                                   "incidence+prevalence curves [R][6][T][12] and per-replicate scalars (runsim :77-101) in host arrays",
                        "host_numa_bound": numa_bound,
                        "dense": None if dense_ms is None else {"value": tot_units * args.steps / (dense_ms * 1e-3), "ms_per_step": dense_ms / args.steps,
                                                                "d2h_bytes_per_step": 2 * int(np.prod(shape)) * 2 + scal_bytes,
                                                                "note": "the same loop with every row of the curves sent (the r01 form of e2e)"}},
                "verified": verified, "gather": gather,
                "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline, "kernels": kernels,
                "cpu_baseline": cpu}
        emit(line)
    pipe.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
