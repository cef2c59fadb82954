"""The reference's own calibration design (scripts/local_run.jl:149-175): :newyork, beta in 0.0225:0.0005:0.085,
nsims replicates each, calibration=true, modeltime=50, initialinf=1; value = mean of sum(_get_column_incidence(h, LAT)).
Least-squares line R0 = a + b*beta against the author's fit b = 46.1761, a = -0.00285 (:174).
Usage: python tools/calibration_pin.py oracle [nsims] [threads]   (CPU oracle — test infrastructure)
       python tools/calibration_pin.py gpu [nsims]                (the CUDA path)"""
import sys, time, json; sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
import numpy as np

BETAS = 0.0225 + 0.0005 * np.arange(126)          # 0.0225:0.0005:0.085
REF_SLOPE, REF_ICPT = 46.1761, -0.00285


def fit(betas, means):
    A = np.stack([np.ones_like(betas), betas], axis=1)
    (a, b), *_ = np.linalg.lstsq(A, means, rcond=None)
    return float(a), float(b)


def bootstrap(vals, nboot=2000, seed=1):
    """vals: [n_beta][nsims] per-replicate R0 values. Resamples replicates within each beta; returns (a, b) draws."""
    rng = np.random.default_rng(seed)
    nb, ns = vals.shape
    out = np.zeros((nboot, 2))
    for k in range(nboot):
        idx = rng.integers(0, ns, size=(nb, ns))
        out[k] = fit(BETAS, np.take_along_axis(vals, idx, axis=1).mean(axis=1))
    return out


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "oracle"
    nsims = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
    t0 = time.perf_counter()
    if which == "oracle":
        import oracle_api as O
        threads = int(sys.argv[3]) if len(sys.argv) > 3 else 8
        psets = [O.make_params(beta=float(b), prov="newyork", calibration=True, modeltime=50, initialinf=1) for b in BETAS]
        pof = np.repeat(np.arange(126), nsims)
        seeds = [726 * (i + 1) for i in range(126 * nsims)]
        res = O.run_ensemble(psets, pof, seeds, nthreads=threads)
        vals = res["inc"][:, 0, :, 1].sum(axis=1).reshape(126, nsims).astype(np.float64)
    else:
        import covid19abm_b200 as cv
        psets = [cv.ModelParameters(β=float(b), prov="newyork", calibration=True, modeltime=50, initialinf=1) for b in BETAS]
        pof = np.repeat(np.arange(126), nsims)
        seeds = [726 * (i + 1) for i in range(126 * nsims)]
        res = cv.run_ensemble(psets, 126 * nsims, seeds=seeds, pset_of_rep=pof)
        vals = res["lat_inc_sum"].reshape(126, nsims).astype(np.float64)
    a, b = fit(BETAS, vals.mean(axis=1))
    bs = bootstrap(vals, 1000)
    lo, hi = np.percentile(bs, [0.5, 99.5], axis=0)
    print(json.dumps({"impl": which, "nsims": nsims, "intercept": a, "slope": b, "ci99_intercept": [lo[0], hi[0]], "ci99_slope": [lo[1], hi[1]],
                      "ref": [REF_ICPT, REF_SLOPE], "seconds": time.perf_counter() - t0}))
