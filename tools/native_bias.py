"""Paired comparison (common random numbers) of the native schedule against the exact one on the same seeds."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, covid19abm_b200 as cv
N, T, R = 10000, 300, 1000
kw = dict(β=0.05, prov="ontario", initialinf=3, modeltime=T)
if len(sys.argv) > 1 and sys.argv[1] == "ct":
    kw = dict(β=0.06, prov="ontario", initialinf=3, modeltime=T, fmild=0.5, τmild=1, fsevere=1.0, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3)
ip = cv.ModelParameters(**kw)
seeds = [726 * (i + 1) for i in range(R)]
a = cv.run_ensemble(ip, R, seeds=seeds, mode=cv.MODE_EXACT)
b = cv.run_ensemble(ip, R, seeds=seeds, mode=cv.MODE_NATIVE, path="multikernel")
print("totalisolated mean exact", a["totalisolated"].mean(), "native", b["totalisolated"].mean())
print("paths", a["path"], b["path"], "ms", a["device_ms"], b["device_ms"])
for name, f in (("attack", lambda p: 1 - p[:, 0, -1, 0] / N), ("peak_day", lambda p: p[:, 0, :, 2:8].sum(axis=2).argmax(axis=1).astype(float)),
                ("peak_icu", lambda p: p[:, 0, :, 9].max(axis=1).astype(float))):
    x, y = f(a["prev"]), f(b["prev"])
    big = (1 - a["prev"][:, 0, -1, 0] / N > 0.02) & (1 - b["prev"][:, 0, -1, 0] / N > 0.02)
    d = (y - x)[big]
    print(f"{name:9s} exact {x[big].mean():.5f} native {y[big].mean():.5f} paired diff {d.mean():+.5f} +- {d.std(ddof=1)/np.sqrt(len(d)):.5f}  (n={big.sum()})")
print("extinct exact", (1 - a["prev"][:, 0, -1, 0] / N < 0.02).mean(), "native", (1 - b["prev"][:, 0, -1, 0] / N < 0.02).mean())
