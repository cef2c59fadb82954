"""Two launches of one workload on one handle (for ncu captures; the second is the warm one).
Usage: python tools/one_launch.py [B|REF|C1|C3|D|E] [nrep] [hm|nohm] [exact|native] [auto|fused|fused-global|multikernel]"""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import covid19abm_b200 as cv
from bench import workload, seeds_for
which = sys.argv[1] if len(sys.argv) > 1 else "B"
W = workload(which)
R = int(sys.argv[2]) if len(sys.argv) > 2 else W["R"]
hm = (sys.argv[3] == "hm") if len(sys.argv) > 3 else W["hm"]
mode = cv.MODE_NATIVE if (len(sys.argv) > 4 and sys.argv[4] == "native") else cv.MODE_EXACT
path = sys.argv[5] if len(sys.argv) > 5 else "auto"
psets = [cv.ModelParameters(**kw) for kw in W["psets"]]
pop = cv.Population(R, W["N"], W["T"], store_hmatrix=hm)
pop.set_path(path)
pof = [i % len(psets) for i in range(R)]
for k in range(2):
    pop.configure(psets, seeds_for(k, 0, R), pset_of_rep=pof, mode=mode)
    pop.run()
    print(which, "launch", k, "path", pop.last_path, "ms", pop.last_run_ms, flush=True)
