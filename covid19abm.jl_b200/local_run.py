"""Host-side mirror of scripts/local_run.jl — the callers either side of the simulation path.

`run` (ensemble + the .dat files), `compute_yearly_average`, `_calibrate` / `calibrate` / `calibrate_regression`.
The Slurm/pmap plumbing of the script (:17-21) is replaced by one device run per ensemble; all numbers come from the
device-side reductions of the CUDA path. File formats follow the script: CSV with a `time` column followed by one column
per simulation (`unstack(df, :time, :sim, c)`, :56), tab-delimited tuples for infectors.dat / ctnumbers.dat (:35-38)."""
import os

import numpy as np

from . import HEALTH, LAT, ModelParameters, Population, CabmError, collectdf

_STATES = [h.name for h in HEALTH][:12]
_SAVED = ("LAT", "ASYMP", "MILD", "MISO", "INF", "IISO", "HOS", "ICU", "DED")      # local_run.jl:50-51


def compute_yearly_average(inc_mean, prev_mean, nsims):
    """local_run.jl:75-103 for one group: columns time, cnt, <state>_prev, <state>_inc (lower case), mean over sims per day.
    inc_mean/prev_mean: [T][12] ensemble means (Population.mean_curves(), reduced on the device)."""
    T = inc_mean.shape[0]
    out = {"time": np.arange(1, T + 1), "cnt": np.full(T, nsims)}
    for s, name in enumerate(_STATES):
        out[name.lower() + "_prev"] = prev_mean[:, s]
    for s, name in enumerate(_STATES):
        out[name.lower() + "_inc"] = inc_mean[:, s]
    return out


def _write_csv(path, header, cols):
    with open(path, "w") as f:
        f.write(",".join(header) + "\n")
        for row in zip(*cols):
            f.write(",".join(repr(float(v)) if isinstance(v, (float, np.floating)) else str(int(v)) for v in row) + "\n")


def run(myp, nsims=500, folderprefix="./", writefiles=True, device=0):
    """run(myp, nsims, folderprefix, writefiles) local_run.jl:23-73. Returns {"all": dict of stacked columns} like `mydfs`."""
    if myp.calibration:
        raise CabmError("can not run simulation, calibration is on.")                    # :26
    seeds = [726 * x for x in range(1, nsims + 1)]                                         # runsim :78
    with Population(nsims, myp.popsize, myp.modeltime, device=device) as pop:
        pop.configure(myp, seeds)
        pop.run()
        pop.check_errors()
        inc, prev = pop.curves()
        sc = pop.scalars()
        minc, mprev = pop.mean_curves()
    T = myp.modeltime
    time = np.arange(1, T + 1)
    allag = {"sim": np.repeat(np.arange(1, nsims + 1), T), "time": np.tile(time, nsims)}   # vcat of cdr[i].a :40
    sec = np.zeros((nsims, T))
    for i in range(nsims):
        sec[i] = collectdf(inc[i, 0], prev[i, 0])["secondarys"]
    allag["secondarys"] = sec.reshape(-1)
    for s, name in enumerate(_STATES):
        allag[name + "_INC"] = inc[:, 0, :, s].reshape(-1)
        allag[name + "_PREV"] = prev[:, 0, :, s].reshape(-1)
    if writefiles:
        os.makedirs(folderprefix, exist_ok=True)
        with open(os.path.join(folderprefix, "infectors.dat"), "w") as f:                  # :35
            for i in range(nsims):
                f.write("\t".join(str(int(v)) for v in sc["infectors"][i]) + "\n")
        with open(os.path.join(folderprefix, "ctnumbers.dat"), "w") as f:                  # :38
            for i in range(nsims):
                f.write("\t".join(str(v) for v in (0, 0, int(sc["totalisolated"][i]), 0, 0, 0, 0)) + "\n")
        hdr = ["time"] + [str(i) for i in range(1, nsims + 1)]
        for name in _SAVED:                                                                # :53-59
            s = _STATES.index(name)
            for kind, arr in (("inc", inc), ("prev", prev)):
                fn = os.path.join(folderprefix, f"simlevel_{name.lower()}_{kind}_all.dat")
                _write_csv(fn, hdr, [time] + [arr[i, 0, :, s] for i in range(nsims)])
        ya = compute_yearly_average(minc[0], mprev[0], nsims)                              # :62-64
        _write_csv(os.path.join(folderprefix, "timelevel_all.dat"), list(ya.keys()), list(ya.values()))
        _write_csv(os.path.join(folderprefix, "secondarys_all.dat"), hdr, [time] + [sec[i] for i in range(nsims)])   # :67-69
    return {"all": allag}


def _calibrate(nsims, myp, device=0):
    """local_run.jl:135-147: mean and std (n-1) over sims of sum(_get_column_incidence(h, LAT))."""
    if not myp.calibration:
        raise CabmError("calibration parameter not turned on")
    with Population(nsims, myp.popsize, myp.modeltime, device=device) as pop:
        pop.configure(myp, [726 * x for x in range(1, nsims + 1)])
        pop.run()
        pop.check_errors()
        vals = pop.scalars()["lat_inc_sum"].astype(np.float64)
    return float(vals.mean()), float(vals.std(ddof=1)) if nsims > 1 else 0.0


def calibrate(beta, nsims=1000, init_inf=1, size=10000, prov="newyork", device=0):
    """local_run.jl:149-162."""
    myp = ModelParameters(β=beta, prov=prov, popsize=size, modeltime=50, calibration=True, initialinf=init_inf)
    m, sd = _calibrate(nsims, myp, device=device)
    return m


def calibrate_regression(betas=None, nsims=1000, prov="newyork", device=0):
    """local_run.jl:164-176: R0 over a β grid in ONE device run (every β is a parameter set of the same ensemble), then the
    least-squares line R0 = a + b·β. Returns (intercept, slope, betas, r0s)."""
    if betas is None:
        betas = np.arange(0.0225, 0.085 + 1e-12, 0.0005)                                   # :165
    betas = np.asarray(betas, dtype=np.float64)
    psets = [ModelParameters(β=float(b), prov=prov, modeltime=50, calibration=True, initialinf=1) for b in betas]
    nb = len(betas)
    pof = np.repeat(np.arange(nb), nsims)
    seeds = np.tile(726 * np.arange(1, nsims + 1), nb)
    with Population(nb * nsims, psets[0].popsize, 50, device=device) as pop:
        pop.configure(psets, seeds, pset_of_rep=pof)
        pop.run()
        pop.check_errors()
        vals = pop.scalars()["lat_inc_sum"].astype(np.float64).reshape(nb, nsims)
    ros = vals.mean(axis=1)
    A = np.stack([np.ones(nb), betas], axis=1)                                             # lm(hcat(ones, betas), ros) :172
    coef, *_ = np.linalg.lstsq(A, ros, rcond=None)
    return float(coef[0]), float(coef[1]), betas, ros


def calibrate_robustness(beta, reps, prov="ontario", nsims=(500, 1000), device=0):
    """local_run.jl:178-199: run calibration with the same beta `reps` times for each ensemble size to see the variation of R0.
    Returns means[reps][len(nsims)]. (The reference reuses its free-running RNG; here repetition k of size ns uses its own seeds.)"""
    means = np.zeros((reps, len(nsims)))
    for j, ns in enumerate(nsims):
        myp = ModelParameters(β=beta, prov=prov, modeltime=50, calibration=True, initialinf=1)
        with Population(ns, myp.popsize, 50, device=device) as pop:
            for k in range(reps):
                pop.configure(myp, [726 * (1 + x + ns * (k + reps * j)) for x in range(ns)])
                pop.run()
                pop.check_errors()
                means[k, j] = pop.scalars()["lat_inc_sum"].mean()
    return means
