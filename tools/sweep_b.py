"""Pipelined device time of the bench workload B for a few knobs: handles in flight, split day.
Usage: python tools/sweep_b.py"""
import os, subprocess, sys, json
def run(env, pipeline):
    e = dict(os.environ); e.update(env)
    out = subprocess.run([sys.executable, "bench.py", "--steps", "24", "--warmup", "3", "--no-cpu", "--no-verify", "--pipeline", str(pipeline)], capture_output=True, text=True, env=e)
    try:
        d = json.loads(out.stdout.strip().splitlines()[-1])
        return round(d["ms_per_step"], 3), round(d["e2e"]["ms_per_step"], 3)
    except Exception as ex:
        return ("ERR", out.stderr[-300:])
for pipeline in (2, 3, 4, 6):
    print("pipeline", pipeline, run({}, pipeline), flush=True)
for sd in (30, 45, 90, 120):
    print("split_day", sd, run({"CABM_FUSED_SPLIT_DAY": str(sd)}, 3), flush=True)
