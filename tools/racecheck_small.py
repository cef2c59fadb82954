"""Small tracing-heavy ensemble on the fused (shared-memory) path, for `compute-sanitizer --tool racecheck` (shared-memory hazards):
   compute-sanitizer --tool racecheck --racecheck-report analysis python tools/racecheck_small.py"""
import sys; sys.path.insert(0, '.')
import covid19abm_b200 as cv
ip = cv.ModelParameters(β=0.3, prov="ontario", popsize=3000, modeltime=45, initialinf=16, ctstrat=1, fctcapture=1.0, fcontactst=0.5, cdaysback=7,
                        fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3)
res = cv.run_ensemble(ip, 4, path="fused")
print("path", res["path"], "totalisolated", res["totalisolated"].tolist(), "attack", (1 - res["prev"][:, 0, -1, 0] / 3000).round(3).tolist())
