"""BASELINE.json's full sizes (1000 replicates x 10,000 agents x 500 days; 10,000 calibration replicates), where the oracle
would take minutes: size-independent properties of the model instead of element-wise comparison, plus run-to-run
determinism of the exact schedule (a data race in the kernels would show up as differing bytes)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
N, T = 10000, 500
B = dict(β=0.0345, prov="ontario", initialinf=1, modeltime=T, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3, tpreiso=0)


def check_ensemble(cv, out, ip, nrep):
    inc, prev, sc = out["inc"], out["prev"], out
    assert (sc["err"] == 0).all()
    tot = prev[:, 0]
    assert (tot.sum(axis=2) == N).all()                                   # every column of hmatrix holds N agents
    assert (prev[:, 1:].sum(axis=1) == tot).all() and (inc[:, 1:].sum(axis=1) == inc[:, 0]).all()   # groups partition the population
    assert (np.diff(tot[:, :, cv.SUS], axis=1) <= 0).all()                # susceptibles only decrease
    assert (np.diff(tot[:, :, cv.REC] + tot[:, :, cv.DED], axis=1) >= 0).all()   # absorbing states only grow
    # prevalence is the running balance of entries and exits; a state is entered at most once per agent:
    ever = inc[:, 0].sum(axis=1)                                          # agents that ever showed state s
    assert (ever[:, cv.SUS] == N - ip.initialinf - (0 if ip.calibration else ip.initialhi)).all()   # seeds are never seen as SUS
    assert (ever[:, cv.REC] + ever[:, cv.DED] >= tot[:, -1, cv.REC] + tot[:, -1, cv.DED]).all()
    assert (ever[:, cv.REC] == tot[:, -1, cv.REC]).all() and (ever[:, cv.DED] == tot[:, -1, cv.DED]).all()
    # an agent infected on day t is first visible as LAT in column t+1 (:136-138): all but the last day's infections are counted
    nonsus_final = N - tot[:, -1, cv.SUS]
    assert (ever[:, cv.LAT] <= nonsus_final - ip.initialinf + 0).all()
    # _count_infectors (:295-313) runs on the final agents, i.e. one time_update after the last column: it can only exceed
    # the count of non-susceptibles in that column by the infections of the last day
    last_day = sc["infectors"].sum(axis=1) - nonsus_final
    assert (last_day >= 0).all() and (last_day <= 50).all()
    assert (sc["lat_inc_sum"] == ever[:, cv.LAT]).all()
    assert (inc >= 0).all() and (prev >= 0).all()


def test_config_b_full_size_properties_and_determinism(cv):
    ip = cv.ModelParameters(**B)
    seeds = [726 * (i + 1) for i in range(1000)]
    a = cv.run_ensemble(ip, 1000, seeds=seeds)
    b = cv.run_ensemble(ip, 1000, seeds=seeds)
    assert a["path"] == "fused"
    check_ensemble(cv, a, ip, 1000)
    for k in ("inc", "prev", "infectors", "totalisolated", "lat_inc_sum"):
        assert (a[k] == b[k]).all(), k
    attack = 1 - a["prev"][:, 0, -1, 0] / N
    assert 0.5 < (attack < 0.02).mean() < 0.95 and attack.max() > 0.1     # bimodal: most runs die out, some take off


def test_config_c_full_size_tracing(cv):
    ip = cv.ModelParameters(ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3, **B)
    a = cv.run_ensemble(ip, 1000)
    m = cv.run_ensemble(ip, 1000, path="multikernel")
    check_ensemble(cv, a, ip, 1000)
    for k in ("inc", "prev", "infectors", "totalisolated"):
        assert (a[k] == m[k]).all(), k                                    # fused and per-day kernels agree at full size
    assert a["totalisolated"].sum() > 0


def test_config_d_calibration_sweep(cv):
    betas = np.linspace(0.0225, 0.085, 20)
    psets = [cv.ModelParameters(β=float(b), prov="ontario", calibration=True, modeltime=50, initialinf=1) for b in betas]
    pof = [i % 20 for i in range(10000)]
    out = cv.run_ensemble(psets, 10000, pset_of_rep=pof)
    check_ensemble(cv, out, psets[0], 10000)
    r0 = np.array([out["lat_inc_sum"][np.array(pof) == k].mean() for k in range(20)])
    assert np.corrcoef(betas, r0)[0, 1] > 0.98                            # R0 grows linearly with beta (scripts/local_run.jl:174)
    slope = np.polyfit(betas, r0, 1)[0]
    assert 35 < slope < 60
    assert (out["prev"][:, 0, :, 2:10].sum(axis=(1, 2)) > 0).all()        # the seed is infectious; in calibration nobody else is


def test_config_e_million_agents_native(cv):
    ip = cv.ModelParameters(β=0.0345, prov="ontario", popsize=1000000, modeltime=120, initialinf=100, quar_grp_frac=0.5, quar_grp=5)
    out = cv.run_ensemble(ip, 4, mode=cv.MODE_NATIVE)
    assert out["path"] == "multikernel" and (out["err"] == 0).all()
    tot = out["prev"][:, 0]
    assert (tot.sum(axis=2) == 1000000).all()
    assert (np.diff(tot[:, :, cv.SUS], axis=1) <= 0).all()
    assert (tot[:, 0, cv.PRE] == 100).all()
    assert (tot[:, -1, cv.SUS] < 1000000 - 100).all()                      # 100 seeds in 10^6 agents do not die out
