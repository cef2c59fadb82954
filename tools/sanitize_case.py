"""Small end-to-end case for compute-sanitizer (memcheck / racecheck): both paths, both schedules, tracing on."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np, covid19abm_b200 as cv
ip = cv.ModelParameters(β=0.25, prov="ontario", popsize=2000, modeltime=40, initialinf=5, fmild=0.5, τmild=1, fsevere=0.8,
                        ctstrat=1, fctcapture=0.8, fcontactst=0.7, cdaysback=3, quar_grp_frac=0.3, quar_grp=3)
plain = cv.ModelParameters(β=0.25, prov="ontario", popsize=2000, modeltime=40, initialinf=5)
seeds = [7, 8, 9]
a = cv.run_ensemble(ip, 3, seeds=seeds, want_hmatrix=True, path="fused")
b = cv.run_ensemble(ip, 3, seeds=seeds, want_hmatrix=True, path="multikernel")
c = cv.run_ensemble(plain, 3, seeds=seeds, want_hmatrix=True, path="fused")
d = cv.run_ensemble(ip, 3, seeds=seeds, mode=cv.MODE_NATIVE, path="multikernel")
assert (a["hmatrix"] == b["hmatrix"]).all() and (a["inc"] == b["inc"]).all() and (a["totalisolated"] == b["totalisolated"]).all()
print("ok", a["path"], b["path"], int((a["hmatrix"][:, -1] != 0).sum()), int(a["totalisolated"].sum()), int((c["hmatrix"][:, -1] != 0).sum()), d["err"])
