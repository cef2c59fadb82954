"""Edge cases of the path, both implementations side by side (GPU tests through the C ABI, oracle as the checker)."""
import numpy as np
import pytest

from test_gpu_parity import to_oracle

pytestmark = pytest.mark.gpu


def both(cv, O, seeds, path, **kw):
    ip = cv.ModelParameters(**kw)
    got = cv.run_ensemble(ip, len(seeds), seeds=seeds, want_hmatrix=True, path=path)
    ref = O.run_ensemble([to_oracle(O, ip)], [0] * len(seeds), seeds, nthreads=4, want_hmatrix=True)
    assert (got["hmatrix"] == ref["hmatrix"]).all(), kw
    assert (got["inc"] == ref["inc"]).all() and (got["prev"] == ref["prev"]).all() and (got["infectors"] == ref["infectors"]).all(), kw
    assert (got["err"] == 0).all()
    return got


@pytest.mark.parametrize("path", ["fused", "multikernel", "fused-global"])
def test_degenerate_scenarios(cv, O, path):
    seeds = [2 ** 63 + 12345, 2 ** 40 + 7, 1]                                   # 64-bit seeds use both halves of the Philox key
    g = both(cv, O, seeds, path, β=0.0, prov="ontario", popsize=3000, modeltime=40, initialinf=4)      # no transmission
    assert (g["prev"][:, 0, -1, 0] == 3000 - 4).all()
    g = both(cv, O, seeds, path, β=0.5, prov="ontario", popsize=3000, modeltime=40, initialinf=0)      # nobody seeded
    assert (g["prev"][:, 0, :, 0] == 3000).all() and (g["infectors"] == 0).all()
    both(cv, O, seeds, path, β=0.4, prov="ontario", popsize=3000, modeltime=1, initialinf=3)            # a single column
    both(cv, O, seeds, path, β=0.4, prov="ontario", popsize=3000, modeltime=2, initialinf=3)
    g = both(cv, O, seeds, path, β=1.0, prov="nunavut", popsize=3000, modeltime=60, initialinf=2999)     # all but one seeded
    assert (g["prev"][:, 0, 0, cv.PRE] == 2999).all()
    # seeding everybody leaves insert_infected(REC, 0, 4) without susceptibles: the reference errors (:371), and so do we
    ip = cv.ModelParameters(β=1.0, popsize=3000, modeltime=10, initialinf=3000)
    with pytest.raises(cv.CabmError, match="No susceptibles"):
        cv.main(ip, seed=1, path=path)
    with pytest.raises(O.OracleError, match="No susceptibles"):
        O.Sim(to_oracle(O, ip), 1).main()
    both(cv, O, seeds, path, β=1.0, prov="yukon", popsize=3000, modeltime=80, initialinf=50, initialhi=2900,          # almost all immune
         fmild=1.0, τmild=3, fsevere=1.0, fpreiso=1.0, tpreiso=5, frelasymp=1.0)
    both(cv, O, seeds, path, β=0.6, prov="pei", popsize=3000, modeltime=120, initialinf=30, quar_grp_frac=1.0, quar_grp=3,   # extremes of every switch
         ctstrat=3, strat3qdays=0, fctcapture=1.0, fcontactst=1.0, cdaysback=20, asymp_props=(0, 0, 0, 0, 0), mild_props=(1, 1, 0, 0, 0))
    both(cv, O, seeds, path, β=0.6, prov="saskat", popsize=3000, modeltime=120, initialinf=30, ctstrat=1, fctcapture=1.0, fcontactst=1.0,
         cdaysback=0, τmild=40, fmild=1.0)                                          # τmild longer than any infectious period (exp <= 0, :515)


@pytest.mark.parametrize("path", ["fused", "fused-global"])
def test_dense_order_dependences_halve_the_chunk(cv, O, path):
    """Small populations in which nearly everybody is infectious: most contact events of a chunk hit a later infector of the same
    chunk, far more than its edge list holds, so the kernel halves the chunk (repeatedly) and must still reproduce the sequential
    order bit for bit. Also long DAG chains (budgets reduced by infectors whose own budgets were reduced)."""
    seeds = [5, 6, 7, 8]
    both(cv, O, seeds, path, β=0.5, prov="ontario", popsize=400, modeltime=30, initialinf=399)
    both(cv, O, seeds, path, β=0.9, prov="quebec", popsize=700, modeltime=40, initialinf=650, fmild=0.5, τmild=1, fsevere=1.0,
         ctstrat=1, fctcapture=1.0, fcontactst=1.0, cdaysback=4)
    both(cv, O, seeds, path, β=0.3, prov="ontario", popsize=2500, modeltime=60, initialinf=1200, initialhi=300)


@pytest.mark.parametrize("path", ["fused", "multikernel"])
def test_too_many_seeds_is_the_reference_error(cv, O, path):
    ip = cv.ModelParameters(β=0.1, popsize=500, modeltime=10, initialinf=501)
    with pytest.raises(O.OracleError, match="No susceptibles"):
        O.Sim(to_oracle(O, ip), 1).main()
    with pytest.raises(cv.CabmError, match="No susceptibles"):
        cv.main(ip, seed=1, path=path)


@pytest.mark.parametrize("path", ["fused", "multikernel"])
def test_empty_age_group_is_flagged(cv, O, path):
    """rand(grps[i], g) on an empty group throws in the reference (:807); tiny populations can have one."""
    ip = cv.ModelParameters(β=0.9, prov="nunavut", popsize=12, modeltime=30, initialinf=6)
    flagged = 0
    for seed in range(1, 30):
        ags = O.Sim(to_oracle(O, ip), seed)
        ags.initialize()
        has_empty = len(set(ags.field("ag"))) < 5
        ags.close()
        out = cv.run_ensemble(ip, 1, seeds=[seed], path=path, strict=False)
        if has_empty:
            try:
                O.Sim(to_oracle(O, ip), seed).main()
                oracle_threw = False
            except O.OracleError:
                oracle_threw = True
            assert bool(out["err"][0] & 8) == oracle_threw, seed
            flagged += oracle_threw
        else:
            assert out["err"][0] == 0
    assert flagged > 0


def test_invalid_arguments_are_rejected(cv):
    with pytest.raises(cv.CabmError, match="invalid ct strategy"):
        cv.main(cv.ModelParameters(ctstrat=2, popsize=100, modeltime=5))
    with pytest.raises(cv.CabmError, match="canadian provinces"):
        cv.main(cv.ModelParameters(prov="atlantis", popsize=100, modeltime=5))
    with pytest.raises(cv.CabmError):
        cv.Population(1, 0, 10)
    with pytest.raises(cv.CabmError):
        cv.Population(1, 100, 1001)                       # exp = 999 sentinel: modeltime must stay below it (:179,:418)
    with cv.Population(2, 100, 5) as pop:
        with pytest.raises(cv.CabmError, match="has not been called"):
            pop.run()
        with pytest.raises(cv.CabmError):
            pop.configure(cv.ModelParameters(popsize=101, modeltime=5), [1])       # handle was sized for 100 agents
        with pytest.raises(cv.CabmError):
            pop.configure(cv.ModelParameters(popsize=100, modeltime=5), [1, 2, 3])  # more replicates than the handle holds
        pop.configure(cv.ModelParameters(popsize=100, modeltime=5), [1, 2])
        with pytest.raises(cv.CabmError, match="CABM_STORE_HMATRIX"):
            pop.run(); pop.hmatrix(0)
