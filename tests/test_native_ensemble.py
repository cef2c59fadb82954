"""Native (relaxed-order) mode: ensemble equivalence with the sequential reference order (BASELINE.json north_star).

The native dyntrans kernel resolves contact events with atomicCAS instead of the reference's ascending-infector order,
so single trajectories may differ from the oracle; the ensemble must not: two-sample KS tests (p > 0.01) on attack
rate, peak day and peak ICU, ensemble means within 3 standard errors of the difference, and the extinction fraction
(single-seed runs are bimodal: SURVEY §8c). The oracle ensemble uses different seeds, so this is a statistical
comparison of two independent samples, not a replay. A 2-SE criterion on independent samples raises a false alarm for
~5 % of seed sets per statistic (and early-growth luck is shared by all scenarios run on one seed set), so the bound here
is 3 SE; the north star's 2-SE requirement is checked in the much sharper paired form at the bottom of this file, where
both schedules run the same seeds and the mean difference must stay below half a standard error of the ensemble; and, to
the north star's letter, on BASELINE configs B and C at full size (1000 replicates x 500 days): two-sample KS p > 0.01 on
attack rate, peak day and peak ICU and ensemble means within 2 standard errors (test_native_full_size_configs)."""
import numpy as np
import pytest
from scipy import stats

from test_gpu_parity import to_oracle

pytestmark = pytest.mark.gpu

NREP, T, N = 1000, 300, 10000


def summary(prev):
    """prev: [R][6][T][12] -> attack rate, peak day, peak ICU per replicate (group 0 = all agents)."""
    tot = prev[:, 0]
    attack = 1.0 - tot[:, -1, 0] / N
    infectious = tot[:, :, 2:8].sum(axis=2)
    return attack, infectious.argmax(axis=1).astype(np.float64), tot[:, :, 9].max(axis=1).astype(np.float64)


@pytest.mark.parametrize("kw", [
    dict(β=0.05, prov="ontario", initialinf=3, modeltime=T),
    dict(β=0.06, prov="ontario", initialinf=3, modeltime=T, fmild=0.5, τmild=1, fsevere=1.0, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3),
], ids=["plain", "isolation+tracing"])
def test_native_ensemble_matches_sequential_order(cv, O, kw):
    ip = cv.ModelParameters(**kw)
    g = cv.run_ensemble(ip, NREP, seeds=[726 * (i + 1) for i in range(NREP)], mode=cv.MODE_NATIVE, path="multikernel")
    assert g["path"] == "multikernel" and (g["err"] == 0).all()
    # the oracle takes the next NREP replicate ids of the same seeding rule (runsim :78): disjoint, independent samples
    r = O.run_ensemble([to_oracle(O, ip)], [0] * NREP, [726 * (NREP + i + 1) for i in range(NREP)], nthreads=16)
    ga, gd, gi = summary(g["prev"])
    ra, rd, ri = summary(r["prev"])
    # extinction fraction (attack < 2 %): binomial SE of the difference
    ge, re_ = (ga < 0.02).mean(), (ra < 0.02).mean()
    assert abs(ge - re_) < 3 * np.sqrt(ge * (1 - ge) / NREP + re_ * (1 - re_) / NREP) + 1e-9
    big_g, big_r = ga >= 0.02, ra >= 0.02
    for name, a, b in (("attack", ga[big_g], ra[big_r]), ("peak_day", gd[big_g], rd[big_r]), ("peak_icu", gi[big_g], ri[big_r]),
                       ("attack_all", ga, ra)):
        p = stats.ks_2samp(a, b).pvalue
        assert p > 0.01, (name, p)
        se = np.sqrt(a.var(ddof=1) / len(a) + b.var(ddof=1) / len(b))
        assert abs(a.mean() - b.mean()) < 3 * se, (name, a.mean(), b.mean(), se)
    # conservation: every column of the state matrix sums to N (SURVEY §8c v)
    assert (g["prev"][:, 0].sum(axis=2) == N).all()
    assert (g["prev"][:, 1:].sum(axis=1) == g["prev"][:, 0]).all()


@pytest.mark.parametrize("kw", [
    dict(β=0.05, prov="ontario", initialinf=3, modeltime=200),
    dict(β=0.0345, prov="ontario", initialinf=1, modeltime=300, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3),
], ids=["plain", "configB"])
def test_native_vs_exact_common_random_numbers(cv, kw):
    """Same seeds through both schedules (common random numbers): the paired difference isolates any bias of the relaxed
    order far below the ensemble's own standard error. Also: trajectories are identical until the first day on which an
    exhausted budget or a contacted infector makes the order matter."""
    ip = cv.ModelParameters(**kw)
    R = 400
    seeds = [726 * (i + 1) for i in range(R)]
    a = cv.run_ensemble(ip, R, seeds=seeds, mode=cv.MODE_EXACT, path="multikernel", want_hmatrix=True)
    b = cv.run_ensemble(ip, R, seeds=seeds, mode=cv.MODE_NATIVE, path="multikernel", want_hmatrix=True)
    assert (b["err"] == 0).all()
    assert (a["hmatrix"][:, :3] == b["hmatrix"][:, :3]).all()          # seeding and the first days agree bit for bit
    xa = 1.0 - a["prev"][:, 0, -1, 0] / N
    xb = 1.0 - b["prev"][:, 0, -1, 0] / N
    assert ((xa < 0.02) == (xb < 0.02)).mean() > 0.97                   # same replicates go extinct
    d = xb - xa
    se_ens = xa.std(ddof=1) / np.sqrt(R)
    assert abs(d.mean()) < 0.5 * se_ens, (d.mean(), se_ens)
    assert (b["prev"][:, 0].sum(axis=2) == N).all()


C_TRACE = dict(ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3)
B_FULL = dict(β=0.0345, prov="ontario", initialinf=1, modeltime=500, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3, tpreiso=0)


@pytest.mark.parametrize("kw", [B_FULL, dict(B_FULL, **C_TRACE), dict(B_FULL, ctstrat=3, strat3qdays=14, fctcapture=0.5, fcontactst=0.5, cdaysback=3)],
                         ids=["configB", "configC-ctstrat1", "configC-ctstrat3"])
def test_native_full_size_configs(cv, kw):
    """BASELINE configs[1] and [2] as stated (1000 replicates x 10,000 agents x 500 days), native schedule against the exact one
    (which is bit-identical to the oracle, tests/test_full_size_parity.py) on the same seeds: KS p > 0.01 on attack rate, peak
    day and peak ICU; means within 2 standard errors of the ensemble mean; same replicates die out."""
    ip = cv.ModelParameters(**kw)
    R = 1000
    seeds = [726 * (i + 1) for i in range(R)]
    a = cv.run_ensemble(ip, R, seeds=seeds, mode=cv.MODE_EXACT)
    b = cv.run_ensemble(ip, R, seeds=seeds, mode=cv.MODE_NATIVE)
    assert a["path"] == "fused" and b["path"] == "multikernel" and (b["err"] == 0).all()
    sa, sb = summary(a["prev"]), summary(b["prev"])
    assert ((sa[0] < 0.02) == (sb[0] < 0.02)).mean() > 0.97
    for name, x, y in zip(("attack", "peak_day", "peak_icu"), sa, sb):
        p = stats.ks_2samp(x, y).pvalue
        assert p > 0.01, (name, p)
        se = np.sqrt(x.var(ddof=1) / R + y.var(ddof=1) / R)
        assert abs(x.mean() - y.mean()) < 2 * se, (name, x.mean(), y.mean(), se)
    assert (b["prev"][:, 0].sum(axis=2) == N).all()
    if ip.ctstrat:
        ta, tb = a["totalisolated"].astype(np.float64), b["totalisolated"].astype(np.float64)
        assert abs(ta.mean() - tb.mean()) < 2 * np.sqrt(ta.var(ddof=1) / R + tb.var(ddof=1) / R)


def test_native_vs_exact_at_100k_agents(cv):
    """BASELINE configs[4] ingredients at 100,000 agents (the exact schedule on the per-day kernels is the comparison there):
    paired common-random-number runs and KS tests, sheltered age group, 64 replicates x 150 days."""
    ip = cv.ModelParameters(β=0.0345, prov="ontario", popsize=100000, modeltime=150, initialinf=10, quar_grp_frac=0.5, quar_grp=5)
    R, Nl = 64, 100000
    seeds = [726 * (i + 1) for i in range(R)]
    a = cv.run_ensemble(ip, R, seeds=seeds, mode=cv.MODE_EXACT)
    b = cv.run_ensemble(ip, R, seeds=seeds, mode=cv.MODE_NATIVE)
    assert a["path"] == "fused-global" and b["path"] == "multikernel" and (b["err"] == 0).all() and (a["err"] == 0).all()
    xa, xb = 1.0 - a["prev"][:, 0, -1, 0] / Nl, 1.0 - b["prev"][:, 0, -1, 0] / Nl
    assert stats.ks_2samp(xa, xb).pvalue > 0.01
    assert abs(xa.mean() - xb.mean()) < 2 * np.sqrt(xa.var(ddof=1) / R + xb.var(ddof=1) / R)
    ia, ib = a["prev"][:, 0, :, 9].max(axis=1).astype(np.float64), b["prev"][:, 0, :, 9].max(axis=1).astype(np.float64)
    assert stats.ks_2samp(ia, ib).pvalue > 0.01
    assert (b["prev"][:, 0].sum(axis=2) == Nl).all()


@pytest.mark.parametrize("kw", [
    dict(β=0.0345, prov="ontario", initialinf=5, modeltime=200, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3),
    dict(β=0.06, prov="ontario", initialinf=1, modeltime=200, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3),
    dict(β=0.8, prov="newyork", popsize=2048, initialinf=600, modeltime=40),          # most agents infectious: the edge list overflows
    dict(β=0.0345, prov="ontario", popsize=200000, initialinf=50, modeltime=150),
], ids=["B", "tracing", "overflow", "200k"])
def test_native_edge_list_count_equals_the_full_second_sweep(cv, monkeypatch, kw):
    """The second-order budget count of the native schedule walks the list of the first sweep's events that hit a later infector.
    With CABM_NATIVE_VERIFY=1 the library also recomputes it by the full second sweep and compares on the device, every day
    (a difference raises ERR_INTERNAL for the replicate)."""
    monkeypatch.setenv("CABM_NATIVE_VERIFY", "1")
    ip = cv.ModelParameters(**kw)
    R = 8 if ip.popsize > 50000 else 64
    res = cv.run_ensemble(ip, R, mode=cv.MODE_NATIVE, strict=False)
    assert res["path"] == "multikernel"
    assert (res["err"] == 0).all(), res["err"]
    attack = 1 - res["prev"][:, 0, -1, 0] / ip.popsize
    assert attack.max() > 0.05                                                        # the check saw real epidemics
    monkeypatch.setenv("CABM_NATIVE_VERIFY", "0")
    monkeypatch.setenv("CABM_NATIVE_EDGES", "0")                                      # three full sweeps: the same distribution of outcomes
    res3 = cv.run_ensemble(ip, R, mode=cv.MODE_NATIVE, strict=False)
    assert (res3["err"] == 0).all()
    a3 = 1 - res3["prev"][:, 0, -1, 0] / ip.popsize
    assert abs(attack.mean() - a3.mean()) < 4 * np.sqrt((attack.var() + a3.var()) / R) + 1e-3
