import os, sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np
import covid19abm_b200 as cv
import oracle_api as O; O.lib()
from test_gpu_parity import to_oracle, _random_scenario
rng = np.random.default_rng(20260 + 4)
for k in range(3):
    kw = _random_scenario(rng); seeds = [int(s) for s in rng.integers(1, 2 ** 62, size=6)]
ip = cv.ModelParameters(**kw)
print(kw)
ref = O.run_ensemble([to_oracle(O, ip)], [0] * 6, seeds, nthreads=6, want_hmatrix=True)
fails = {}
for it in range(12):
    for path in ("fused", "multikernel"):
        got = cv.run_ensemble(ip, 6, seeds=seeds, want_hmatrix=True, path=path)
        d = got["hmatrix"] != ref["hmatrix"]
        bad = d.any() or not (got["totalisolated"] == ref["totalisolated"]).all()
        fails[path] = fails.get(path, 0) + int(bad)
        if bad: print(it, path, "ndiff", int(d.sum()), "iso", (got["totalisolated"] - ref["totalisolated"]).tolist())
print(os.environ.get("CABM_LIB", "default")[-14:], "failures of 12:", fails)
