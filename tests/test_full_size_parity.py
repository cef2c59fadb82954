"""Bit-exact parity with the oracle at BASELINE.json's stated sizes (the bench shape itself was never compared).

  B   1000 replicates x 10,000 agents x 500 days, isolation on, state matrix kept, three batches through the three-handle
      EnsemblePipeline that bench.py times (park/resume, fill jobs, copy jobs and overlapping launches all exercised):
      every curve entry, every scalar, and every byte of the state matrix of the first batch
  C   the same ensemble with contact tracing, ctstrat = 1 and ctstrat = 3
  D   10,000 calibration replicates over a 20-point beta grid, 50 days
  E   1,000,000 agents: exact schedule on the per-day kernels against the oracle, 60 days
Tolerance is zero everywhere. The oracle (test infrastructure) runs on all host cores; state matrices are compared in
chunks so that neither side ever holds the 10 GB Int16 ensemble."""
import os

import numpy as np
import pytest

from test_gpu_parity import to_oracle

pytestmark = pytest.mark.gpu
N, T = 10000, 500
B = dict(β=0.0345, prov="ontario", initialinf=1, modeltime=T, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3, tpreiso=0)
THREADS = os.cpu_count() or 8


def compare_against_oracle(O, pop, psets, pof, seeds, inc, prev, sc, hm_replicates=0, chunk=48):
    """Curves and scalars of every replicate; the state matrix (read back replicate by replicate through cabm_get_hmatrix)
    of the first `hm_replicates`."""
    op = [to_oracle(O, p) for p in psets]
    R = len(seeds)
    pof = np.zeros(R, dtype=np.int32) if pof is None else np.asarray(pof, dtype=np.int32)
    assert (sc["err"] == 0).all()
    for lo in range(0, R, chunk):
        hi = min(R, lo + chunk)
        want = lo < hm_replicates
        ref = O.run_ensemble(op, pof[lo:hi], seeds[lo:hi], nthreads=THREADS, want_hmatrix=want)
        assert (np.asarray(inc[lo:hi]) == ref["inc"]).all(), f"incidence differs in replicates {lo}..{hi}"
        assert (np.asarray(prev[lo:hi]) == ref["prev"]).all(), f"prevalence differs in replicates {lo}..{hi}"
        assert (sc["infectors"][lo:hi] == ref["infectors"]).all()
        assert (sc["totalisolated"][lo:hi] == ref["totalisolated"]).all()
        assert (sc["lat_inc_sum"][lo:hi] == ref["inc"][:, 0, :, 1].sum(axis=1)).all()
        if want:
            for r in range(lo, min(hi, hm_replicates)):
                got = pop.hmatrix(r)                       # (N, T) view of the [T][N] buffer
                assert (got.T == ref["hmatrix"][r - lo]).all(), f"state matrix of replicate {r} differs"


def test_config_b_bit_exact_through_the_bench_pipeline(cv, O):
    ip = cv.ModelParameters(**B)
    R, nb = 1000, 3
    seeds = [[726 * (k * R + i + 1) for i in range(R)] for k in range(nb)]
    with cv.EnsemblePipeline(R, N, T, handles=3, store_hmatrix=True, copy=False) as pipe:
        for k in range(nb):
            assert pipe.submit(ip, seeds[k], tag=k) is None
        res = pipe.drain()
        assert [r["tag"] for r in res] == [0, 1, 2] and all(p.last_path == "fused" for p in pipe.pops)
        for k, r in enumerate(res):
            compare_against_oracle(O, pipe.pops[k], [ip], None, seeds[k], r["inc"], r["prev"], r, hm_replicates=R if k == 0 else 0)
        attack = 1 - res[0]["prev"][:, 0, -1, 0] / N
        assert (attack > 0.1).sum() > 100          # the batch holds long epidemics (parked and resumed) as well as early extinctions


def test_config_b_rows_only_transport_is_lossless(cv, O):
    """The result form behind bench.py's e2e: only the rows before a replicate's epidemic died out cross PCIe."""
    ip = cv.ModelParameters(**B)
    R = 1000
    seeds = [726 * (7000 + i) for i in range(R)]
    res = cv.run_ensemble_pipelined(ip, R, batch=R, handles=1, seeds=seeds, rows_only=True)
    assert (res["rows"] <= T).all() and (res["rows"] < T).mean() > 0.5
    inc, prev = cv.expand_curves(res["inc"], res["prev"], res["rows"])
    ref = O.run_ensemble([to_oracle(O, ip)], [0] * R, seeds, nthreads=THREADS)
    assert (inc == ref["inc"]).all() and (prev == ref["prev"]).all()
    assert (res["infectors"] == ref["infectors"]).all()


@pytest.mark.parametrize("ct", [dict(ctstrat=1), dict(ctstrat=3, strat3qdays=14)], ids=["ctstrat1", "ctstrat3"])
def test_config_c_bit_exact_at_full_size(cv, O, ct):
    ip = cv.ModelParameters(fctcapture=0.5, fcontactst=0.5, cdaysback=3, **ct, **B)
    R = 1000
    seeds = [726 * (i + 1) for i in range(R)]
    with cv.Population(R, N, T, store_hmatrix=True) as pop:
        pop.configure(ip, seeds)
        pop.run()
        assert pop.last_path == "fused"
        inc, prev = pop.curves()
        sc = pop.scalars()
        compare_against_oracle(O, pop, [ip], None, seeds, inc, prev, sc, hm_replicates=240)
        assert sc["totalisolated"].sum() > 0


def test_config_d_bit_exact_at_full_size(cv, O):
    betas = np.linspace(0.0225, 0.085, 20)
    psets = [cv.ModelParameters(β=float(b), prov="ontario", calibration=True, modeltime=50, initialinf=1) for b in betas]
    R = 10000
    pof = [i % 20 for i in range(R)]
    seeds = [726 * (i + 1) for i in range(R)]
    with cv.Population(R, N, 50, store_hmatrix=True) as pop:
        pop.configure(psets, seeds, pset_of_rep=pof)
        pop.run()
        inc, prev = pop.curves()
        sc = pop.scalars()
        compare_against_oracle(O, pop, psets, pof, seeds, inc, prev, sc, hm_replicates=400, chunk=400)


def test_config_e_exact_schedule_at_a_million_agents(cv, O):
    """BASELINE configs[4] population (10^6 agents, sheltered age group), exact schedule (the fused kernel with its per-agent
    arrays in global memory, and the per-day kernels), against the oracle: 60 days of one replicate, every element of the state matrix."""
    ip = cv.ModelParameters(β=0.0345, prov="ontario", popsize=1000000, modeltime=60, initialinf=100, quar_grp_frac=0.5, quar_grp=5)
    hm = cv.main(ip, seed=726)
    assert cv.humans().last_path == "fused-global"
    hm2 = cv.main(ip, seed=726, path="multikernel")
    assert (np.asarray(hm) == np.asarray(hm2)).all()
    sim = O.Sim(to_oracle(O, ip), 726)
    ref = sim.main()
    assert (np.asarray(hm) == np.asarray(ref)).all()
    assert (np.asarray(hm)[:, -1] != 0).sum() > 300            # the epidemic is under way
    sc = cv.humans().scalars()
    assert (sc["infectors"][0] == sim.count_infectors()).all()
    sim.close()
