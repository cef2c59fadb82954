"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle on the same seeded inputs — bit-exact.

The reference's own tests hold no golden vectors (test/runtests.jl is invariants only), so the
oracle (oracle/abm_oracle.c, pinned in tests/test_tables.py, test_invariants.py, test_golden.py and
test_calibration_pin.py) is the comparison; tolerance is zero:
every element of the Int16 state matrix, every curve entry and every scalar must be equal.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def to_oracle(O, ip):
    return O.make_params(beta=ip.β, frelasymp=ip.frelasymp, seasonal=ip.seasonal, popsize=ip.popsize, prov=ip.prov,
                         calibration=ip.calibration, modeltime=ip.modeltime, initialinf=ip.initialinf,
                         initialhi=ip.initialhi, asymp_props=ip.asymp_props, mild_props=ip.mild_props,
                         tau_mild=ip.τmild, fmild=ip.fmild, fsevere=ip.fsevere, tpreiso=ip.tpreiso,
                         fpreiso=ip.fpreiso, quar_grp=ip.quar_grp, quar_grp_frac=ip.quar_grp_frac,
                         ctstrat=ip.ctstrat, fctcapture=ip.fctcapture, fcontactst=ip.fcontactst,
                         cdaysback=ip.cdaysback, strat3qdays=ip.strat3qdays)


def first_mismatch(a, b):
    bad = np.argwhere(a != b)
    if len(bad) == 0:
        return None
    k = bad[np.argmin(bad[:, 1])] if a.ndim == 2 else bad[0]
    return tuple(int(x) for x in k), a[tuple(k)], b[tuple(k)]


CONFIGS = {
    # BASELINE.json configs[0]: single replicate Ontario, no isolation or tracing
    "A_ontario": dict(β=0.0345, prov="ontario", initialinf=1, modeltime=500),
    # configs[1]: isolation on (tpreiso=0 never activates fpreiso in the loop: SURVEY Q1; the tpreiso=1 variant does)
    "B_isolation": dict(β=0.0345, prov="ontario", initialinf=1, modeltime=500, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3, tpreiso=0),
    "B_isolation_tpre1": dict(β=0.05, prov="ontario", initialinf=5, modeltime=300, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3, tpreiso=1),
    # configs[2]: contact tracing (ctstrat 2 does not exist in the source, :165)
    "C_ct1": dict(β=0.05, prov="ontario", initialinf=5, modeltime=300, fmild=0.5, τmild=1, fsevere=1.0, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3),
    "C_ct3": dict(β=0.05, prov="ontario", initialinf=5, modeltime=300, fmild=0.5, τmild=1, fsevere=1.0, ctstrat=3, strat3qdays=7, fctcapture=0.5, fcontactst=0.5, cdaysback=3),
    "C_ct1_full": dict(β=0.06, prov="newyork", initialinf=10, modeltime=200, ctstrat=1, fctcapture=1.0, fcontactst=1.0, cdaysback=5),
    # configs[3]: calibration mode
    "D_calibration": dict(β=0.0525, prov="newyork", calibration=True, modeltime=50, initialinf=1),
    # configs[4] ingredients at the reference's population size: quarantine group, herd immunity, seasonality
    "E_quarantine": dict(β=0.06, prov="ontario", initialinf=10, modeltime=300, quar_grp_frac=0.5, quar_grp=5),
    "seasonal_hi": dict(β=0.08, prov="bc", initialinf=3, initialhi=2000, modeltime=250, seasonal=True, frelasymp=0.3),
    "hot": dict(β=0.2, prov="nunavut", initialinf=20, modeltime=120, fsevere=0.8, fmild=0.3, τmild=2),
}


def _cases():
    for name, kw in CONFIGS.items():
        for path in ("multikernel", "fused", "fused-global"):
            yield pytest.param(name, path, id=f"{name}-{path}")


@pytest.mark.parametrize("name,path", list(_cases()))
@pytest.mark.parametrize("seed", [726, 1452])
def test_main_bit_exact(cv, O, name, path, seed):
    ip = cv.ModelParameters(**CONFIGS[name])
    hm = cv.main(ip, seed=seed, path=path)
    assert cv.humans().last_path == path
    sim = O.Sim(to_oracle(O, ip), seed)
    ref = sim.main()
    mm = first_mismatch(np.asarray(hm), np.asarray(ref))
    assert mm is None, f"first mismatch (agent, day) {mm}"
    pop = cv.humans()
    for f in ["health", "swap", "sickfrom", "ag", "iso", "isovia", "nextday_meetcnt", "tis", "exp", "doi",
              "tracing", "tracestart", "traceend", "tracedxp"]:
        assert (pop.field(f) == sim.field(f)).all(), f
    assert (pop.field("dur") == sim.field("dur")).all()
    sc = pop.scalars()
    assert (sc["infectors"][0] == sim.count_infectors()).all()
    assert sc["totalisolated"][0] == sim.totalisolated
    inc, prev = pop.curves()
    rinc, rprev = O.collect_curves(ref, sim.field("ag"))
    assert (inc[0] == rinc).all() and (prev[0] == rprev).all()
    assert sc["lat_inc_sum"][0] == O.get_column_incidence(ref, O.LAT).sum()
    sim.close()


def test_ensemble_fused_bit_exact(cv, O):
    """400 replicates (more than the resident CTAs: exercises the work queue) x 2 parameter sets, fused kernel."""
    a = cv.ModelParameters(β=0.0345, prov="ontario", modeltime=120, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3)
    b = cv.ModelParameters(β=0.09, prov="quebec", modeltime=120, initialinf=4, initialhi=500, quar_grp_frac=0.4, quar_grp=2,
                           ctstrat=3, strat3qdays=5, fctcapture=0.7, fcontactst=0.6, cdaysback=4, fmild=0.3, τmild=1)
    nrep = 400
    seeds = [726 * (i + 1) for i in range(nrep)]
    pof = [i % 2 for i in range(nrep)]
    got = cv.run_ensemble([a, b], nrep, seeds=seeds, pset_of_rep=pof, want_hmatrix=True, path="fused")
    assert got["path"] == "fused"
    ref = O.run_ensemble([to_oracle(O, a), to_oracle(O, b)], pof, seeds, nthreads=16, want_hmatrix=True)
    assert (got["err"] == 0).all()
    assert (got["hmatrix"] == ref["hmatrix"]).all()
    assert (got["inc"] == ref["inc"]).all() and (got["prev"] == ref["prev"]).all()
    assert (got["infectors"] == ref["infectors"]).all()
    assert (got["totalisolated"] == ref["totalisolated"]).all() and ref["totalisolated"].sum() > 0
    assert (got["lat_inc_sum"] == ref["inc"][:, 0, :, 1].sum(axis=1)).all()


def test_ensemble_bit_exact(cv, O):
    """64 replicates x 2 parameter sets in one launch sequence, against the oracle's replicate loop."""
    a = cv.ModelParameters(β=0.0345, prov="ontario", modeltime=200, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3)
    b = cv.ModelParameters(β=0.07, prov="ontario", modeltime=200, initialinf=3, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3)
    nrep = 64
    seeds = [726 * (i + 1) for i in range(nrep)]
    pof = [i % 2 for i in range(nrep)]
    got = cv.run_ensemble([a, b], nrep, seeds=seeds, pset_of_rep=pof, want_hmatrix=True)
    ref = O.run_ensemble([to_oracle(O, a), to_oracle(O, b)], pof, seeds, nthreads=8, want_hmatrix=True)
    assert (got["err"] == 0).all()
    assert (got["hmatrix"] == ref["hmatrix"]).all()
    assert (got["inc"] == ref["inc"]).all() and (got["prev"] == ref["prev"]).all()
    assert (got["infectors"] == ref["infectors"]).all()
    assert (got["totalisolated"] == ref["totalisolated"]).all()


def test_reduce_hmatrix_matches_literal_reductions(cv, O):
    """_get_column_incidence / _get_column_prevalence (:270-293) on device for arbitrary matrices, including
    rows that revisit a state (first occurrence only) — against the oracle's literal findfirst/findall scans."""
    rng = np.random.default_rng(5)
    M, N, T = 3, 1000, 40
    hm = rng.integers(0, 12, size=(M, T, N)).astype(np.int16)
    ags = rng.integers(1, 6, size=(M, N)).astype(np.int32)
    with cv.Population(1, N, T) as pop:
        inc, prev = pop.reduce_hmatrix(hm, ags)
    for m in range(M):
        rinc, rprev = O.collect_curves(hm[m].T, ags[m])
        assert (inc[m] == rinc).all() and (prev[m] == rprev).all()
        for s in range(12):
            assert (inc[m, 0, :, s] == O.get_column_incidence(hm[m].T, s)).all()
            assert (prev[m, 0, :, s] == O.get_column_prevalence(hm[m].T, s)).all()


def test_stepwise_fields_follow_oracle(cv, O):
    """initialize / insert_infected / dyntrans / time_update one phase at a time, all agent fields compared."""
    ip = cv.ModelParameters(β=0.3, prov="ontario", modeltime=60, initialinf=1, fmild=0.5, τmild=1, fsevere=0.7,
                            ctstrat=1, fctcapture=1.0, fcontactst=0.7, cdaysback=3, quar_grp_frac=0.3, quar_grp=3)
    seed = 99
    names = ["health", "swap", "sickfrom", "ag", "iso", "isovia", "nextday_meetcnt", "tis", "exp", "doi",
             "tracing", "tracestart", "traceend", "tracedxp"]
    with cv.Population(1, ip.popsize, ip.modeltime) as pop:
        pop.configure(ip, [seed])
        sim = O.Sim(to_oracle(O, ip), seed)
        pop.initialize(); sim.initialize()
        for f in names:
            assert (pop.field(f) == sim.field(f)).all(), ("init", f)
        pop.insert_infected(cv.PRE, 30, 4, call_id=0); sim.insert_infected(O.PRE, 30, 4, 0)
        sim.p.fpreiso = 0.0
        for f in names:
            assert (pop.field(f) == sim.field(f)).all(), ("seed", f)
        for day in range(1, 31):
            met, inf = pop.dyntrans()
            rmet, rinf = sim.dyntrans()
            assert (int(met[0]), int(inf[0])) == (rmet, rinf), day
            for f in ["nextday_meetcnt", "swap", "sickfrom", "exp"]:
                assert (pop.field(f) == sim.field(f)).all(), (day, "dyntrans", f)
            pop.time_update(); sim.time_update()
            for f in names:
                g, r = pop.field(f), sim.field(f)
                assert (g == r).all(), (day, "time_update", f, np.nonzero(g != r)[0][:5])
        assert pop.totalisolated()[0] == sim.totalisolated
        sim.close()


def test_large_population_bit_exact(cv, O):
    """BASELINE config 5 ingredients beyond the reference's 10,000-agent limit (:113): 200,000 agents, 100 seeds, a
    quarantined age group — exact schedule; the population does not fit shared memory: the fused kernel keeps the per-agent arrays in
    global memory (default), or the per-day kernels run."""
    ip = cv.ModelParameters(β=0.06, prov="ontario", popsize=200000, modeltime=40, initialinf=100, quar_grp_frac=0.5, quar_grp=5)
    ref = O.run_ensemble([to_oracle(O, ip)], [0, 0], [726, 1452], nthreads=2, want_hmatrix=True)
    for path, expect in (("auto", "fused-global"), ("multikernel", "multikernel")):
        got = cv.run_ensemble(ip, 2, seeds=[726, 1452], want_hmatrix=True, path=path)
        assert got["path"] == expect
        assert (got["err"] == 0).all()
        assert (got["hmatrix"] == ref["hmatrix"]).all()
        assert (got["inc"] == ref["inc"]).all() and (got["prev"] == ref["prev"]).all()
        assert (got["infectors"] == ref["infectors"]).all()


def test_odd_population_sizes_bit_exact(cv, O):
    """popsize not a multiple of 8/16/32 (ragged tails of the vector loads and of the TMA column store)."""
    for n, path in ((9999, "fused"), (1003, "fused"), (1003, "multikernel"), (37, "fused")):
        ip = cv.ModelParameters(β=0.3, prov="ontario", popsize=n, modeltime=60, initialinf=3, fmild=0.4, τmild=1, fsevere=0.6)
        got = cv.run_ensemble(ip, 3, seeds=[5, 6, 7], want_hmatrix=True, path=path)
        ref = O.run_ensemble([to_oracle(O, ip)], [0, 0, 0], [5, 6, 7], nthreads=3, want_hmatrix=True)
        assert (got["hmatrix"] == ref["hmatrix"]).all(), (n, path)
        assert (got["inc"] == ref["inc"]).all() and (got["prev"] == ref["prev"]).all(), (n, path)


def test_int16_curves_equal_int32(cv):
    ip = cv.ModelParameters(β=0.2, prov="ontario", modeltime=80, initialinf=5)
    with cv.Population(8, ip.popsize, ip.modeltime) as pop:
        pop.configure(ip, [11 * (i + 1) for i in range(8)])
        pop.run()
        a, b = pop.curves()
        c, d = pop.curves(dtype=np.int16)
        assert c.dtype == np.int16 and (a == c).all() and (b == d).all() and a.sum() > 0


def test_local_run_outputs(cv, O, tmp_path):
    """scripts/local_run.jl mirror: run() writes the .dat files from device-reduced curves; compute_yearly_average is the
    device-side ensemble mean; _calibrate is the mean LAT-incidence sum (the author's R0 estimate, :135-147)."""
    ip = cv.ModelParameters(β=0.1, prov="ontario", modeltime=60, initialinf=2)
    nsims = 6
    dfs = cv.local_run.run(ip, nsims, str(tmp_path), True)
    ref = O.run_ensemble([to_oracle(O, ip)], [0] * nsims, [726 * x for x in range(1, nsims + 1)], nthreads=4)
    assert (dfs["all"]["LAT_INC"].reshape(nsims, 60) == ref["inc"][:, 0, :, 1]).all()
    tl = np.genfromtxt(tmp_path / "timelevel_all.dat", delimiter=",", names=True)
    assert np.allclose(tl["lat_inc"], ref["inc"][:, 0, :, 1].mean(axis=0)) and np.allclose(tl["sus_prev"], ref["prev"][:, 0, :, 0].mean(axis=0))
    sl = np.genfromtxt(tmp_path / "simlevel_icu_prev_all.dat", delimiter=",", skip_header=1)
    assert sl.shape == (60, nsims + 1) and (sl[:, 1:].T == ref["prev"][:, 0, :, 9]).all()
    inf = np.loadtxt(tmp_path / "infectors.dat")
    assert (inf == ref["infectors"]).all()
    for f in ("ctnumbers.dat", "secondarys_all.dat", "simlevel_lat_inc_all.dat", "simlevel_ded_prev_all.dat"):
        assert (tmp_path / f).exists()
    cal = cv.ModelParameters(β=0.05, prov="ontario", modeltime=50, calibration=True)
    m, sd = cv.local_run._calibrate(64, cal)
    rc = O.run_ensemble([to_oracle(O, cal)], [0] * 64, [726 * x for x in range(1, 65)], nthreads=8)
    v = rc["inc"][:, 0, :, 1].sum(axis=1).astype(float)
    assert m == v.mean() and abs(sd - v.std(ddof=1)) < 1e-12
    rb = cv.local_run.calibrate_robustness(0.05, 2, prov="ontario", nsims=(100, 200))
    assert rb.shape == (2, 2) and (rb > 1.0).all() and (rb < 4.0).all()
    a, b, betas, ros = cv.local_run.calibrate_regression(betas=[0.03, 0.05, 0.07], nsims=200, prov="newyork")
    assert 30 < b < 60 and abs(a) < 0.6        # the author's fit: R0 = 46.1761 beta - 0.00285 (scripts/local_run.jl:174)


def test_curves_streamed_to_pinned_host_memory(cv):
    """cabm_set_host_curves_i16: the fused kernel writes finished replicates' curves into registered pinned buffers."""
    ip = cv.ModelParameters(β=0.15, prov="ontario", modeltime=90, initialinf=4, fmild=0.4, τmild=1)
    R = 300
    with cv.Population(R, ip.popsize, ip.modeltime) as pop:
        a, b = cv.PinnedArray((R, 6, 90, 12), np.int16), cv.PinnedArray((R, 6, 90, 12), np.int16)
        a.array[:] = -1; b.array[:] = -1
        pop.stream_curves_to(a.array, b.array)
        pop.configure(ip, [3 * (i + 1) for i in range(R)])
        pop.run()
        assert pop.last_path == "fused"
        inc, prev = pop.curves()                                  # ordinary int32 read-back
        assert (a.array == inc).all() and (b.array == prev).all()
        c, d = pop.curves(a.array, b.array)                        # same pointers: synchronise only
        assert (c == inc).all() and (d == prev).all()
        pop.stream_curves_to(None, None)
        pop.configure(ip, [5 * (i + 1) for i in range(R)])
        pop.run()
        inc2, _ = pop.curves()
        assert (a.array == inc).all() and (inc2 != inc).any()    # unregistered: the old buffers are left alone
        a.free(); b.free()


def test_handle_reuse_leaves_no_residue(cv):
    """A handle is reconfigured and rerun many times (bench.py, β sweeps): results must not depend on what ran before."""
    a = cv.ModelParameters(β=0.3, prov="ontario", modeltime=70, initialinf=8, ctstrat=1, fctcapture=0.9, fcontactst=0.9, cdaysback=3)
    b = cv.ModelParameters(β=0.04, prov="bc", modeltime=70, initialinf=1)
    seeds = [9 * (i + 1) for i in range(40)]
    fresh = cv.run_ensemble(b, 40, seeds=seeds)
    with cv.Population(64, 10000, 70) as pop:
        for path in ("fused", "multikernel", "fused"):
            pop.set_path(path)
            pop.configure(a, [5 * (i + 1) for i in range(64)]); pop.run()
            pop.configure(b, seeds); pop.run()
            inc, prev = pop.curves()
            sc = pop.scalars()
            assert (inc == fresh["inc"]).all() and (prev == fresh["prev"]).all(), path
            assert (sc["infectors"] == fresh["infectors"]).all() and (sc["totalisolated"] == 0).all() and (sc["err"] == 0).all()


@pytest.mark.parametrize("n", [12000, 20000])
def test_fused_one_cta_per_sm_populations(cv, O, n):
    """Populations between ~10.4k and ~21k agents still fit shared memory, one CTA per SM instead of two."""
    ip = cv.ModelParameters(β=0.08, prov="ontario", popsize=n, modeltime=60, initialinf=6, fmild=0.4, τmild=1,
                            ctstrat=1, fctcapture=0.6, fcontactst=0.6, cdaysback=3)
    got = cv.run_ensemble(ip, 3, seeds=[21, 22, 23], want_hmatrix=True, path="fused")
    assert got["path"] == "fused"
    ref = O.run_ensemble([to_oracle(O, ip)], [0, 0, 0], [21, 22, 23], nthreads=3, want_hmatrix=True)
    assert (got["hmatrix"] == ref["hmatrix"]).all() and (got["inc"] == ref["inc"]).all() and (got["prev"] == ref["prev"]).all()
    assert (got["totalisolated"] == ref["totalisolated"]).all()


def test_fused_refuses_populations_that_do_not_fit(cv):
    ip = cv.ModelParameters(β=0.08, prov="ontario", popsize=40000, modeltime=20)
    with pytest.raises(cv.CabmError, match="fused path needs"):
        cv.run_ensemble(ip, 1, path="fused")
    assert cv.run_ensemble(ip, 1)["path"] == "fused-global"           # auto: the same kernel with the per-agent arrays in global memory
    assert cv.run_ensemble(ip, 1, mode=cv.MODE_NATIVE)["path"] == "multikernel"


@pytest.mark.parametrize("popsize,split", [(10000, None), (10007, None), (10000, "25")])
def test_fused_parked_replicates_and_tail_jobs_bit_exact(cv, O, popsize, split, monkeypatch):
    """More replicates than resident CTAs and a long run: the fused kernel runs every replicate for `split_day` days, parks the
    ones that are still active and queues the idle tail of the others (curve rows, state-matrix columns, host copy) as tail
    jobs. Everything must equal the oracle's sequential runs, with the curves also streamed to pinned host memory. popsize
    10007 takes the path without TMA bulk stores (columns are not 16-byte multiples); split 25 parks nearly everything that
    was seeded with several cases."""
    if split is not None:
        monkeypatch.setenv("CABM_FUSED_SPLIT_DAY", split)
    T = 160
    a = cv.ModelParameters(β=0.0345, prov="ontario", popsize=popsize, modeltime=T, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3)
    b = cv.ModelParameters(β=0.08, prov="alberta", popsize=popsize, modeltime=T, initialinf=3, ctstrat=1, fctcapture=0.6, fcontactst=0.6,
                           cdaysback=3, fmild=0.4, τmild=1)
    nrep = 330
    seeds = [41 * (i + 1) for i in range(nrep)]
    pof = [1 if i % 3 == 0 else 0 for i in range(nrep)]
    ref = O.run_ensemble([to_oracle(O, a), to_oracle(O, b)], pof, seeds, nthreads=16, want_hmatrix=True)
    with cv.Population(nrep, popsize, T, store_hmatrix=True) as pop:
        pi, pp = cv.PinnedArray((nrep, 6, T, 12), np.int16), cv.PinnedArray((nrep, 6, T, 12), np.int16)
        pi.array[:] = -1; pp.array[:] = -1
        pop.set_path("fused")
        pop.stream_curves_to(pi.array, pp.array)
        pop.configure([a, b], seeds, pof)
        pop.run()
        assert pop.last_path == "fused"
        inc, prev = pop.curves()
        sc = pop.scalars()
        hm = pop.hmatrix_all()
        assert (pop.hmatrix_all(dtype=np.uint8) == hm).all()                    # one byte per agent-day, as the device keeps it
        assert (sc["err"] == 0).all()
        assert (inc == ref["inc"]).all() and (prev == ref["prev"]).all()
        assert (pi.array == ref["inc"]).all() and (pp.array == ref["prev"]).all()
        assert (hm == ref["hmatrix"]).all()
        assert (sc["infectors"] == ref["infectors"]).all() and (sc["totalisolated"] == ref["totalisolated"]).all()
        extinct = ref["prev"][:, 0, -1, 2:8].sum(axis=1) == 0                   # nobody infectious at the end ...
        assert extinct.any() and (~extinct).any()                               # ... for some replicates but not for all: both kinds of job ran
        pop.stream_curves_to(None, None)
        pi.free(); pp.free()


@pytest.mark.parametrize("handles", [2, 3])
def test_pipelined_batches_equal_one_big_ensemble(cv, handles):
    """cv.EnsemblePipeline: consecutive batches run concurrently on several handles of one GPU (own stream, state and pinned result
    buffers each). Replicates are independent, so the concatenated batches must equal the same replicates run as one ensemble —
    whatever else shares the SMs while a batch runs."""
    ip = cv.ModelParameters(β=0.05, prov="ontario", modeltime=160, initialinf=2, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3)
    nsims, batch = 960, 320
    seeds = [17 * (i + 1) for i in range(nsims)]
    whole = cv.run_ensemble(ip, nsims, seeds=seeds)
    parts = cv.run_ensemble_pipelined(ip, nsims, batch=batch, handles=handles, seeds=seeds)
    assert whole["path"] == "fused" and (whole["err"] == 0).all() and (parts["err"] == 0).all()
    assert parts["inc"].dtype == np.int16                      # streamed into pinned host memory by the kernel
    for k in ("inc", "prev", "infectors", "totalisolated", "lat_inc_sum"):
        assert (parts[k] == whole[k]).all(), k
    attack = 1 - whole["prev"][:, 0, -1, 0] / ip.popsize
    assert (attack > 0.05).any() and (attack < 0.01).any()      # long and short replicates share the GPU


def test_second_device_in_the_same_process(cv):
    """Handles on two GPUs of one process (kernel attributes and constant tables are per device): same replicates, same bytes."""
    import ctypes
    try:
        n = ctypes.c_int(0)
        rt = ctypes.CDLL("libcudart.so")
        rt.cudaGetDeviceCount(ctypes.byref(n))
    except OSError:
        pytest.skip("no libcudart to count devices with")
    if n.value < 2:
        pytest.skip("needs two GPUs")
    ip = cv.ModelParameters(β=0.06, prov="ontario", modeltime=80, initialinf=3)
    a = cv.run_ensemble(ip, 64, device=0)
    b = cv.run_ensemble(ip, 64, device=1)
    assert a["path"] == b["path"] == "fused"
    for k in ("inc", "prev", "infectors", "lat_inc_sum"):
        assert (a[k] == b[k]).all(), k


@pytest.mark.parametrize("mode", ["exact", "native"])
def test_per_day_kernels_replayed_from_a_cuda_graph(cv, mode):
    """The per-day kernel path captures its launch sequence into a CUDA graph on the third run of a handle with the same arguments
    and replays it afterwards: runs 1, 2 (direct launches), 3 (capture) and 4, 5 (replay) must give what a fresh handle gives for
    the same seeds, also after the arguments change (tracing on: direct launches again, then a new graph)."""
    a = cv.ModelParameters(β=0.08, prov="ontario", modeltime=60, initialinf=3, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3)
    b = cv.ModelParameters(β=0.08, prov="ontario", modeltime=60, initialinf=3, ctstrat=1, fctcapture=0.7, fcontactst=0.7, cdaysback=3)
    m = cv.MODE_NATIVE if mode == "native" else cv.MODE_EXACT
    R = 48
    with cv.Population(R, a.popsize, 60) as pop:
        pop.set_path("multikernel")
        for k, ip in enumerate([a, a, a, a, a, b, b, b, b, a]):
            seeds = [31 * (i + 1) + k for i in range(R)]
            pop.configure(ip, seeds, mode=m); pop.run()
            assert pop.last_path == "multikernel"
            inc, prev = pop.curves(); sc = pop.scalars()
            if mode == "exact":
                fresh = cv.run_ensemble(ip, R, seeds=seeds, path="fused")          # the other path, a fresh handle: same bytes
                assert (inc == fresh["inc"]).all() and (prev == fresh["prev"]).all(), k
                assert (sc["infectors"] == fresh["infectors"]).all() and (sc["totalisolated"] == fresh["totalisolated"]).all(), k
            else:
                fresh = cv.run_ensemble(ip, R, seeds=seeds, path="multikernel", mode=m)
                assert (prev[:, 0].sum(axis=2) == a.popsize).all() and (sc["err"] == 0).all(), k
                # the native schedule is not bit-reproducible between runs (atomics decide), its size is
                assert abs(int(prev[:, 0, -1, 0].sum()) - int(fresh["prev"][:, 0, -1, 0].sum())) < 0.2 * R * a.popsize
    # launch accounting still counts kernels, replayed or not
    with cv.Population(8, a.popsize, 60) as pop:
        pop.set_path("multikernel")
        counts = []
        for k in range(5):
            l0 = pop.launch_count
            pop.configure(a, [5 * (i + 1) for i in range(8)], mode=m); pop.run()
            counts.append(pop.launch_count - l0)
        assert len(set(counts)) == 1 and counts[0] >= 2 * 60


@pytest.mark.parametrize("T", [160, 151])
def test_streamed_curves_rows_only(cv, T):
    """cabm_set_host_curve_rows: the kernel sends only the rows of a replicate's curves that differ (an epidemic that died out
    repeats its last row) and reports how many; expanded on the host they equal the dense curves. An odd modeltime sends
    everything (rows do not end on 16 bytes)."""
    ip = cv.ModelParameters(β=0.0345, prov="ontario", modeltime=T, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3)
    R = 340
    with cv.Population(R, ip.popsize, T) as pop:
        a, b, rows = cv.PinnedArray((R, 6, T, 12), np.int16), cv.PinnedArray((R, 6, T, 12), np.int16), cv.PinnedArray((R,), np.int32)
        a.array[:] = -7; b.array[:] = -7; rows.array[:] = -1
        pop.stream_curves_to(a.array, b.array, rows.array)
        pop.configure(ip, [13 * (i + 1) for i in range(R)])
        pop.run()
        pop.sync()
        inc, prev = pop.curves()                                   # dense int32 read-back of the same run
        n = rows.array.copy()
        assert (n >= 1).all() and (n <= T).all()
        if T % 2:
            assert (n == T).all()
        else:
            assert (n < T // 2).mean() > 0.5                       # most of these epidemics die out early ...
            sent = a.array != -7
            for r in (0, int(np.argmin(n)), int(np.argmax(n))):
                assert sent[r, :, :n[r]].all() and not sent[r, :, n[r]:].any()     # ... and only their first rows crossed the link
        ei, ep = cv.expand_curves(a.array.copy(), b.array.copy(), n)
        assert (ei == inc).all() and (ep == prev).all()
        pop.stream_curves_to(None, None)
        a.free(); b.free(); rows.free()


def _random_scenario(rng):
    provs = ["alberta", "bc", "manitoba", "newbruns", "newfdland", "nwterrito", "novasco", "nunavut", "ontario", "pei", "quebec", "saskat", "yukon", "newyork"]
    ct = int(rng.choice([0, 0, 1, 3]))
    kw = dict(β=float(rng.choice([0.0, 0.02, 0.05, 0.1, 0.3, 0.8])), prov=str(rng.choice(provs)),
              popsize=int(rng.choice([1999, 2048, 3000, 4097, 6000, 10000, 12345])), modeltime=int(rng.integers(2, 140)),
              initialinf=int(rng.integers(0, 25)), initialhi=int(rng.choice([0, 0, 50, 900])), seasonal=bool(rng.integers(0, 2)),
              frelasymp=float(rng.choice([0.0, 0.11, 0.5, 1.0])), calibration=bool(rng.random() < 0.15),
              τmild=int(rng.integers(0, 6)), fmild=float(rng.choice([0.0, 0.3, 1.0])), fsevere=float(rng.choice([0.0, 0.6, 1.0])),
              tpreiso=int(rng.integers(0, 30)), fpreiso=float(rng.choice([0.0, 0.3, 1.0])),
              quar_grp=int(rng.integers(1, 6)), quar_grp_frac=float(rng.choice([0.0, 0.0, 0.4, 1.0])),
              ctstrat=ct, fctcapture=float(rng.choice([0.2, 0.7, 1.0])) if ct else 0.0, fcontactst=float(rng.choice([0.2, 0.7, 1.0])) if ct else 0.0,
              cdaysback=int(rng.integers(0, 8)) if ct else 0, strat3qdays=int(rng.integers(0, 15)) if ct == 3 else 0)
    if kw["initialinf"] + kw["initialhi"] > kw["popsize"] - 10:
        kw["initialhi"] = 0
    return kw


@pytest.mark.parametrize("block", range(8))
def test_random_scenarios_bit_exact(cv, O, block):
    """Seeded random draws over the whole parameter space (every switch of ModelParameters, odd population sizes, short and long
    runs), six replicates each, fused kernel and per-day kernels against the oracle: state matrix, curves, scalars."""
    rng = np.random.default_rng(20260 + block)
    for k in range(8):
        kw = _random_scenario(rng)
        ip = cv.ModelParameters(**kw)
        seeds = [int(s) for s in rng.integers(1, 2 ** 62, size=6)]
        ref = O.run_ensemble([to_oracle(O, ip)], [0] * 6, seeds, nthreads=6, want_hmatrix=True)
        for path in (("fused", "multikernel") if k % 2 == 0 else ("fused",)):
            got = cv.run_ensemble(ip, 6, seeds=seeds, want_hmatrix=True, path=path)
            assert (got["hmatrix"] == ref["hmatrix"]).all(), (path, kw)
            assert (got["inc"] == ref["inc"]).all() and (got["prev"] == ref["prev"]).all(), (path, kw)
            assert (got["infectors"] == ref["infectors"]).all() and (got["totalisolated"] == ref["totalisolated"]).all(), (path, kw)


def test_long_contact_sets_are_identified_the_same_way_every_run(cv, O):
    """Index cases whose traced-contact set is longer than a group's gather buffer (seven days back at beta = 0.3: more than 128
    entries) take the walk over the log itself. Regression: a lane that ran ahead of its group there once let the next case's
    gather overwrite buffer slots of the case after it — a result that changed from run to run. Eight runs per path."""
    kw = dict(β=0.3, prov="nwterrito", popsize=6000, modeltime=112, initialinf=16, seasonal=True, frelasymp=1.0, τmild=2, fmild=0.3,
              fsevere=0.0, tpreiso=1, fpreiso=0.0, ctstrat=1, fctcapture=1.0, fcontactst=0.2, cdaysback=7)
    ip = cv.ModelParameters(**kw)
    seeds = [726 * (i + 1) for i in range(6)]
    ref = O.run_ensemble([to_oracle(O, ip)], [0] * 6, seeds, nthreads=6, want_hmatrix=True)
    assert ref["totalisolated"].sum() > 1000
    for it in range(8):
        for path in ("fused", "multikernel", "fused-global"):
            got = cv.run_ensemble(ip, 6, seeds=seeds, want_hmatrix=True, path=path)
            assert (got["totalisolated"] == ref["totalisolated"]).all(), (it, path)
            assert (got["hmatrix"] == ref["hmatrix"]).all(), (it, path)
            assert (got["inc"] == ref["inc"]).all() and (got["prev"] == ref["prev"]).all(), (it, path)


@pytest.mark.parametrize("kw", [
    dict(β=0.8, prov="newyork", popsize=4097, modeltime=59, initialinf=10, ctstrat=1, fctcapture=0.7, fcontactst=0.2, cdaysback=5),
    dict(β=0.3, prov="ontario", popsize=6000, modeltime=80, initialinf=16, ctstrat=3, strat3qdays=7, fctcapture=1.0, fcontactst=0.7, cdaysback=7),
    dict(β=0.1, prov="quebec", popsize=12345, modeltime=70, initialinf=12, initialhi=50, fmild=0.3, τmild=2, fsevere=0.6, fpreiso=0.3),
], ids=["ct1-dense", "ct3-long-sets", "no-tracing"])
def test_repeated_runs_are_identical(cv, kw):
    """Poor man's race detector (compute-sanitizer is not available on the test boxes): the same ensemble five times per path — every
    byte of the state matrices, the curves and the tracing counter must repeat. (Cross-warp hazards show up as run-to-run differences
    on heavy days; the comparison with the oracle is in the tests above.)"""
    ip = cv.ModelParameters(**kw)
    seeds = [991 * (i + 3) for i in range(6)]
    for path in ("fused", "fused-global", "multikernel"):
        first = cv.run_ensemble(ip, 6, seeds=seeds, want_hmatrix=True, path=path)
        for it in range(4):
            again = cv.run_ensemble(ip, 6, seeds=seeds, want_hmatrix=True, path=path)
            for k in ("hmatrix", "inc", "prev", "totalisolated", "infectors"):
                assert (again[k] == first[k]).all(), (path, it, k)
