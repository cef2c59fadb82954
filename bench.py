#!/usr/bin/env python3
"""bench.py — agent-days/s of covid19abm.jl's simulation path (BASELINE.json metric) on N B200s.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # the CPU restatement of the reference (oracle port)

Workload (config.workload): BASELINE.json configs[1] — an ensemble of the 10,000-agent x 500-day Ontario
scenario with isolation on (fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3, tpreiso=0), 1000 replicates per GPU
(weak scaling: replicates are independent, sharded over ranks with no data-path collective; one end-of-run
NCCL gather of the summary curves). A step = one full pass of the hot path (main(ip) :104-142 for every
replicate of the batch: initialize, seed, 500 x (dyntrans, time_update + snapshot), reductions).

Steps alternate over `--pipeline` handles of the GPU (default 3; cv.EnsemblePipeline): each handle has its own stream, state
and pinned result buffers, so the replicates of step k+1 that die out run beside the few long epidemics of step k, which
would otherwise leave most SMs idle while they finish. Every step is one full pass over its own 1000 replicates; all K
steps lie inside the timed region. `--pipeline 1` runs one step after the other (config.ms_one_step_alone).

`value` is device time (CUDA events: a mark on the first handle's stream to the end of every handle's last run) with parameters and seeds already resident;
`e2e` goes through the public API from host parameter structs to host result buffers (curves + scalars as
runsim returns them, :77-101), copies included. The Julia reference cannot run in this image (no julia):
the CPU arm is the C restatement in oracle/ ("port"), one replicate per core on all host cores.
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
WORKLOAD = "configs[1]: Ontario beta=0.0345, 10000 agents x 500 days, isolation on (fmild=0.5 taumild=1 fsevere=1.0 fpreiso=0.3 tpreiso=0)"
CFG = dict(β=0.0345, prov="ontario", initialinf=1, modeltime=T_DAYS, popsize=N_AGENTS,
           fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3, tpreiso=0)
METRIC, UNIT = "agent_days_per_sec", "agent-days/s"


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


def cpu_sample(nrep, threads):
    """Times the oracle port (oracle/abm_oracle.c) on `nrep` replicates of the workload with `threads` workers."""
    import oracle_api as O
    from test_gpu_parity import to_oracle
    import covid19abm_b200 as cv
    p = to_oracle(O, cv.ModelParameters(**CFG))
    seeds = seeds_for(0, 999, nrep)
    t0 = time.perf_counter()
    O.run_ensemble([p], [0] * nrep, seeds, nthreads=threads)
    dt = time.perf_counter() - t0
    return nrep * N_AGENTS * T_DAYS / dt, dt


def run_reference(args, rank, world, emit):
    """--impl reference: the reference's CPU implementation of the path. Julia is absent from the image, so this is
    the oracle port (kind "port"), all host threads, each step a bounded sample of the same workload."""
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    nrep = max(threads, 16)
    for _ in range(args.warmup):
        cpu_sample(min(nrep, threads), threads)
    tot = 0.0
    for _ in range(args.steps):
        _, dt = cpu_sample(nrep, threads)
        tot += dt
    value = args.steps * nrep * N_AGENTS * T_DAYS / tot
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * tot / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "replicates_per_step": nrep, "note": "CPU restatement of covid19abm.jl in C (oracle/), not Julia: no julia binary in the image"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": f"{nrep} replicates per step, one per thread"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--replicates", type=int, default=1000, help="replicates per GPU per step")
    ap.add_argument("--no-hmatrix", action="store_true", help="do not keep the state matrices in HBM (curves only)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--pipeline", type=int, default=3, help="handles (CUDA streams) the steps alternate over; 1 = one step after the other")
    ap.add_argument("--path", default="auto", choices=["auto", "multikernel", "fused"], help="execution path (include/cabm.h CABM_PATH_*)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    # the contract is ONE JSON line on stdout: libraries (NCCL's version banner, torchrun) also write there, so everything
    # else goes to stderr and the line is written to the saved descriptor at the end
    real_stdout = os.dup(1)
    os.dup2(2, 1)
    os.environ["NCCL_DEBUG"] = "WARN"          # the pool sets VERSION, whose banner would land on stdout

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())

    if args.impl == "reference":
        run_reference(args, rank, world, emit)
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

    R = args.replicates
    ip = cv.ModelParameters(**CFG)
    # Steps alternate over `--pipeline` handles (each with its own stream, state and pinned result buffers): the replicates of
    # step k+1 that die out fill the SMs which the few long epidemics of step k leave idle (cv.EnsemblePipeline — the public API
    # for ensembles larger than one batch). Every step is still one full pass over its own 1000 replicates.
    P = max(1, args.pipeline)
    pipe = cv.EnsemblePipeline(R, N_AGENTS, T_DAYS, handles=P, store_hmatrix=not args.no_hmatrix, device=local, path=args.path, copy=False)
    pops, pop = pipe.pops, pipe.pops[0]
    pin_inc, pin_prev = pipe.bufs[0]
    for q in pops:
        q.stream_curves_to(None, None)                 # the device-resident leg leaves the results in HBM
    shape = (R, 6, T_DAYS, 12)
    units = R * N_AGENTS * T_DAYS                      # agent-days per step per GPU

    for w in range(max(args.warmup, P)):
        q = pops[w % P]
        q.configure(ip, seeds_for(1000 + w, rank, R))
        q.run()
    for q in pops:
        q.check_errors()
    one_ms = pop.last_run_ms                           # one step alone on the GPU (nothing else in flight)

    sampler = ClockSampler(local)
    sampler.start()
    # ---- value: device-resident, CUDA-event time of exactly K steps (mark on the first handle's stream -> end of every handle's last run)
    barrier()
    l0 = sum(q.launch_count for q in pops)
    t0 = time.perf_counter()
    pop.mark()
    for k in range(args.steps):
        q = pops[k % P]
        q.configure(ip, seeds_for(k, rank, R))         # seeds + thresholds: ~10 KB, resident before the run that uses them
        q.run(sync=False)
    dev_ms = max(q.ms_since_mark(pop) for q in pops[:min(P, args.steps)])
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    launches = sum(q.launch_count for q in pops) - l0
    dev_ms = max_over_ranks(dev_ms)
    wall_ms = max_over_ranks(wall_ms)
    # ---- e2e: host params -> host curves/scalars through the public API, copies inside the timed region
    # the result buffers are pinned host memory registered with the handles: the fused kernel writes each finished replicate's
    # curves into them over PCIe while the rest of the batch is still running (cabm_set_host_curves_i16), so collecting a batch
    # only waits for its run and reads the scalars; on the per-day-kernel path it is an ordinary device->host copy
    for q, bufs in zip(pops, pipe.bufs):
        q.stream_curves_to(bufs[0].array, bufs[1].array)
    barrier()
    t0 = time.perf_counter()
    got = 0
    for k in range(args.steps):
        r = pipe.submit(ip, seeds_for(100 + k, rank, R))
        got += r is not None
    rest = pipe.drain()
    got += len(rest)
    sc = rest[-1]
    barrier()
    e2e_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    assert got == args.steps
    # ---- the same loop with the rows-only form of the streamed curves (cabm_set_host_curve_rows): a replicate whose epidemic has
    # died out repeats its last row, so only the rows before that cross PCIe; reported beside `e2e`, which stays the dense form
    rows_bufs = [cv.PinnedArray((R,), np.int32) for _ in pops]
    for q, bufs, rb_ in zip(pops, pipe.bufs, rows_bufs):
        q.stream_curves_to(bufs[0].array, bufs[1].array, rb_.array)
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        pipe.submit(ip, seeds_for(100 + k, rank, R))
    pipe.drain()
    barrier()
    e2e_rows_ms = max_over_ranks(1e3 * (time.perf_counter() - t0))
    # rows sent per step: the row counts of the last batch on every handle (they stay in the rows arrays after the drain)
    nb_ = min(P, args.steps)
    sent_rows = sum(int(rb_.array.sum()) for rb_ in rows_bufs[:nb_]) / nb_
    rows_d2h = sent_rows * 6 * 12 * 2 * 2 + R * (32 + 8 + 8 + 4 + 4)
    rows_only_ok = pop.last_path == "fused"            # the per-day kernels copy the curves after the run, all rows
    for q in pops:
        q.stream_curves_to(None, None)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = R * 8 + R * 4 + 280 + 4 * T_DAYS * 8
    d2h = 2 * int(np.prod(shape)) * 2 + R * (32 + 8 + 8 + 4)
    assert (sc["err"] == 0).all()

    # ---- end-of-run gather of the summary curves (the path's only collective): mean prevalence per day/state
    mean_prev = torch.from_numpy(pin_prev.array[:, 0].astype(np.float64).mean(axis=0)).cuda()
    if dist is not None:
        gathered = [torch.empty_like(mean_prev) for _ in range(world)] if rank == 0 else None
        dist.gather(mean_prev, gathered, dst=0)
    attack = float(1.0 - pin_prev.array[:, 0, -1, 0].mean() / N_AGENTS)

    # ---- per-kernel-class device time of one more (profiled) step -> roofline of the dominant kernel
    pop.set_profiling(True)
    pop.configure(ip, seeds_for(200, rank, R))
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
    if pop.last_path == "fused":
        # k_sim_fused: per-agent state lives in shared memory; what must cross HBM per step:
        #   the state matrix, 1 B per agent-day (u8 health code, :221); per agent once: the HBM-resident fields written at
        #   initialize and the final status/expday (26 B); the curves [R][6][T][12] int32 x 2.
        hm_b = 0 if args.no_hmatrix else R * N_AGENTS * T_DAYS
        alg_bytes = hm_b + R * N_AGENTS * 26 + 2 * R * 6 * T_DAYS * 12 * 4
        ms, n = prof["fused"]
        gbs = alg_bytes / (ms / n * 1e-3) / 1e9
        traffic = None                                     # dram__bytes_read+write per launch from the committed ncu --set full capture
        tpath = os.path.join(ROOT, "profiles", "r01_fused_traffic.json")
        if os.path.exists(tpath) and R == 1000 and not args.no_hmatrix:
            traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
        roofline = {"kernel": "k_sim_fused", "bound": "hbm", "achieved": round(gbs, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(gbs / peak, 4), "traffic": traffic, "peak_source": peak_src,
                    "bytes_per_launch": alg_bytes, "bytes_per_agent_day": round(alg_bytes / (R * N_AGENTS * T_DAYS), 3),
                    "note": "algorithmic bytes = u8 state matrix + per-agent HBM fields + curves; agent state itself never leaves shared memory. "
                            "At the reference-width figure of SURVEY 8(d) (14 B/agent-day) the same launch would read "
                            f"{round(14 * R * N_AGENTS * T_DAYS / (ms / n * 1e-3) / 1e9, 1)} GB/s"}
        # `achieved` is this kernel timed alone (one launch on an otherwise idle GPU, as ncu sees it). In the timed region launches
        # of consecutive steps overlap; the bytes of the region over its duration are given beside it
        roofline["launch_ms_alone"] = round(ms / n, 3)
        roofline["achieved_in_timed_region"] = round(alg_bytes * args.steps / (dev_ms * 1e-3) / 1e9, 1)
        roofline["frac_in_timed_region"] = round(alg_bytes * args.steps / (dev_ms * 1e-3) / 1e9 / peak, 4)
    else:
        # algorithmic bytes per agent-day of k_time_update in the HBM layout: expday 2 r + previous column 1 r + next column 1 w
        tu_bytes = (2 + (0 if args.no_hmatrix else 2)) * R * N_AGENTS
        dom = max(("dyntrans", "time_update"), key=lambda k: prof[k][0])
        tu_ms, tu_n = prof["time_update"]
        tu_gbs = tu_bytes / (tu_ms / tu_n * 1e-3) / 1e9 if tu_n else 0.0
        kernels["time_update"]["algorithmic_GBps"] = round(tu_gbs, 1)
        roofline = {"kernel": "k_time_update", "bound": "hbm", "achieved": round(tu_gbs, 1), "peak": peak, "unit": "GB/s",
                    "frac": round(tu_gbs / peak, 4), "traffic": None, "peak_source": peak_src,
                    "bytes_per_agent_day": tu_bytes / (R * N_AGENTS), "dominant_by_time": "k_" + dom + ("_exact" if dom == "dyntrans" else ""),
                    "note": "time_update is the HBM-streaming per-agent kernel; dyntrans is a latency-bound gather/scatter (see kernels)"}

    cpu = None
    if rank == 0 and not args.no_cpu:
        threads = os.cpu_count() or 1
        nrep = max(6 * threads, 96)      # about 30 s of single-core work, ~2 s wall on 16 cores
        v, dt = cpu_sample(nrep, threads)
        cpu = {"value": v, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"{nrep} replicates of the same workload, one per thread, {dt:.1f} s wall (C restatement of the reference, not Julia)"}

    if rank == 0:
        value = world * units * args.steps / (dev_ms * 1e-3)
        e2e = world * units * args.steps / (e2e_ms * 1e-3)
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "u32", "data": "synthetic",
                "config": {"workload": WORKLOAD, "replicates_per_gpu": R, "replicates_per_sec": world * R * args.steps / (dev_ms * 1e-3),
                           "mode": "exact (bit-identical to the sequential reference order)", "path": pop.last_path, "hmatrix_in_hbm": not args.no_hmatrix,
                           "l2": "inputs larger than L2: 0.3 GB of agent state + 5 GB state matrix + 0.3 GB curves written per step, rebuilt every step",
                           "wall_ms_per_step": wall_ms / args.steps, "attack_rate_mean": attack,
                           "pipeline": f"steps alternate over {P} handles/streams on the GPU: the short replicates of step k+1 run beside the long epidemics of step k",
                           "ms_one_step_alone": one_ms},
                "e2e": {"value": e2e, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "ms_per_step": e2e_ms / args.steps, "returns": "incidence+prevalence curves [R][6][T][12] Int16 x2 (narrowed on device) and per-replicate scalars (runsim :77-101), pinned host buffers",
                        "host_numa_bound": numa_bound,
                        "rows_only": None if not rows_only_ok else {"value": world * units * args.steps / (e2e_rows_ms * 1e-3), "ms_per_step": e2e_rows_ms / args.steps, "d2h_bytes_per_step": int(rows_d2h),
                                      "note": "optional form of the result (cabm_set_host_curve_rows): only the rows of each replicate's curves before its epidemic died out are sent, "
                                              "plus their number; the rest repeat. Lossless; cv.expand_curves rebuilds the dense arrays on the host"}},
                "gpu_launches": int(launches), "clocks": sampler.summary(), "roofline": roofline, "kernels": kernels,
                "cpu_baseline": cpu}
        emit(line)
    pipe.close()
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
