"""One warm-up and one measured launch of the fused kernel on a workload (for ncu captures).
Usage: python tools/one_launch.py [B|REF|HOT|C1] [nrep] [hm]"""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import covid19abm_b200 as cv
from bench import workload, seeds_for
which = sys.argv[1] if len(sys.argv) > 1 else "B"
W = workload(which)
R = int(sys.argv[2]) if len(sys.argv) > 2 else W["R"]
hm = (sys.argv[3] == "hm") if len(sys.argv) > 3 else W["hm"]
psets = [cv.ModelParameters(**kw) for kw in W["psets"]]
pop = cv.Population(R, W["N"], W["T"], store_hmatrix=hm)
pof = [i % len(psets) for i in range(R)]
for k in range(2):
    pop.configure(psets, seeds_for(k, 0, R), pset_of_rep=pof)
    pop.run()
    print(which, "launch", k, "ms", pop.last_run_ms, flush=True)
