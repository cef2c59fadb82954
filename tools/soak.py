"""Soak test of the work queues: many pipelined steps at several ensemble sizes and pipeline depths; every batch must finish and
repeat the results of the same seeds (run under `timeout`: a hang in the queues would show up as one)."""
import sys, time; sys.path.insert(0, '.')
import numpy as np, covid19abm_b200 as cv
from bench import CFG, N_AGENTS, T_DAYS
ip = cv.ModelParameters(**CFG)
ipc = cv.ModelParameters(ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3, **CFG)
for R, P, psets in ((1000, 3, ip), (700, 5, ip), (1500, 2, ip), (333, 4, ipc), (1000, 3, ipc)):
    ref = None
    t0 = time.perf_counter()
    with cv.EnsemblePipeline(R, N_AGENTS, T_DAYS, handles=P, store_hmatrix=(R <= 1000), copy=True) as pipe:
        seeds = [726 * (i + 1) for i in range(R)]
        outs = []
        for k in range(40):
            r = pipe.submit(psets, seeds)
            if r is not None: outs.append(r)
        outs += pipe.drain()
    assert len(outs) == 40
    for o in outs:
        assert (o["err"] == 0).all()
        key = (o["inc"].sum(), o["prev"][:, 0, -1, :].sum(), o["infectors"].sum(), o["lat_inc_sum"].sum())
        ref = ref or key
        assert key == ref, (key, ref)
    print("R", R, "handles", P, "ct" if psets is ipc else "  ", "40 steps ok,", round(1e3 * (time.perf_counter() - t0) / 40, 2), "ms per step incl. result copies")
# the whole-ensemble form: 6000 replicates in batches of 1000, curves written by the kernels into one pinned array
t0 = time.perf_counter()
big = cv.run_ensemble_pipelined(ip, 6000)
dt = time.perf_counter() - t0
one = cv.run_ensemble(ip, 1000, seeds=[726 * (i + 1) for i in range(3000, 4000)])
assert (big["inc"][3000:4000] == one["inc"]).all() and (big["prev"][3000:4000] == one["prev"]).all() and (big["infectors"][3000:4000] == one["infectors"]).all()
print("run_ensemble_pipelined(6000): %.1f ms per 1000 replicates incl. allocation of the pinned arrays; batch 3 equals a separate run" % (1e3 * dt / 6))
pipe = cv.EnsemblePipeline(1000, N_AGENTS, T_DAYS, pinned=False)
pinned = (cv.PinnedArray((6000, 6, T_DAYS, 12), np.int16), cv.PinnedArray((6000, 6, T_DAYS, 12), np.int16))
for rep in range(2):
    t0 = time.perf_counter(); big2 = cv.run_ensemble_pipelined(ip, 6000, pipe=pipe, pinned=pinned); dt = time.perf_counter() - t0
    assert (big2["inc"] == big["inc"]).all() if rep == 0 else True
    print("with a kept pipeline and pinned arrays: %.2f ms per 1000 replicates" % (1e3 * dt / 6))
big3 = cv.run_ensemble_pipelined(ip, 6000, pipe=pipe, rows_only=True)
ei, ep = cv.expand_curves(big3['inc'].copy(), big3['prev'].copy(), big3['rows'])
assert (ei == big['inc']).all() and (ep == big['prev']).all(); print('rows-only ensemble expands to the dense one; mean rows sent', big3['rows'].mean())
pipe.close()
