"""The reference's own test suite (/root/reference/test/runtests.jl) re-expressed: every testset there is an
invariant under arbitrary draws (no seeds, no goldens), so each is stated once and run against the CPU oracle
(pins the oracle; runs without a GPU) and against the CUDA path through the C ABI (-m gpu)."""
import numpy as np
import pytest

import oracle_api as O
from backends import CudaBackend, OracleBackend

SUS, LAT, PRE, ASYMP, MILD, MISO, INF, IISO, HOS, ICU, REC, DED, UNDEF = range(13)
VIA_NULL, VIA_QU, VIA_PI, VIA_MI, VIA_CT = range(5)


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def make(request):
    made = []

    def _make(**kw):
        b = (OracleBackend if request.param == "oracle" else CudaBackend)(**kw)
        made.append(b)
        return b
    yield _make
    for b in made:
        b.close()


# ---- testset "init" (runtests.jl:27-89)
def test_init_fresh_agents(make):
    b = make(seed=11)
    b.initialize()
    assert len(b.field("health")) == 10000                                   # :32
    assert set(np.unique(b.field("ag"))) == {1, 2, 3, 4, 5}                  # :40
    assert (b.field("health") == SUS).all() and (b.field("swap") == UNDEF).all()      # :45-46
    assert (b.field("sickfrom") == UNDEF).all()
    assert (b.field("nextday_meetcnt") >= 0).all()
    assert (b.field("tis") == 0).all() and (b.field("exp") == 999).all()
    assert (b.field("iso") == 0).all() and (b.field("isovia") == VIA_NULL).all()
    assert b.field("dur").min() != 0                                         # :56
    assert (b.field("doi") == 999).all() and (b.field("tracing") == 0).all()
    assert (b.field("tracestart") == -1).all() and (b.field("traceend") == -1).all() and (b.field("tracedxp") == 0).all()


def test_init_age_distribution_and_durations(make):
    """Known answers the reference implies: Ontario age fractions (:328), NB means (:855), duration supports."""
    b = make(seed=3, prov="ontario")
    b.initialize()
    ag = b.field("ag")
    frac = np.bincount(ag, minlength=6)[1:] / 10000
    assert np.abs(frac - np.array([0.0519, 0.1727, 0.3930, 0.2150, 0.1674])).max() < 0.02
    dur = b.field("dur")
    assert dur[:, 0].min() >= 1 and dur[:, 0].max() <= 6          # latents - pres, latents in 4..7, pres in 1..3
    assert set(np.unique(dur[:, 2])) <= {1, 2, 3}
    assert dur[:, 1].min() >= 1 and dur[:, 3].min() >= 1
    cnt = b.field("nextday_meetcnt")
    means = np.array([10.21, 16.793, 13.7950, 11.2669, 8.0027])
    sds = np.array([7.65, 11.7201, 10.5045, 9.5935, 6.9638])
    for g in range(5):
        sel = cnt[ag == g + 1]
        assert abs(sel.mean() - means[g]) < 5 * sds[g] / np.sqrt(len(sel))


@pytest.mark.parametrize("ag", [1, 2, 3, 4, 5])
def test_insert_infected_into_age_group(make, ag):                           # :64-70
    b = make(seed=5)
    b.initialize()
    b.insert_infected(INF, 1, ag)
    h, a = b.field("health"), b.field("ag")
    assert ((h == INF) & (a == ag)).sum() == 1 and (h == INF).sum() == 1
    b.insert_infected(REC, 1, ag, call_id=1)
    assert (b.field("health") == REC).sum() == 1


def test_seed_inf_is_not_iiso(make):                                         # :73-76
    b = make(seed=5, fsevere=1.0)
    b.initialize()
    b.insert_infected(INF, 1, 1)
    h, s = b.field("health"), b.field("swap")
    assert ((h == INF) & (s == REC)).sum() == 1
    assert (b.field("sickfrom")[h == INF] == INF).all()                      # :391


def test_insert_more_than_susceptible_errors(make):                          # :371-372
    b = make(seed=5, popsize=64, modeltime=10)
    b.initialize()
    with pytest.raises(Exception, match="No susceptibles"):
        b.insert_infected(PRE, 65, 1)


# ---- testset "transitions" (:91-220)
def test_time_update_ages_susceptibles(make):                                # :96-103
    b = make(seed=2)
    b.initialize()
    b.time_update()
    assert (b.field("tis") == 1).all() and (b.field("exp") == 999).all()
    assert (b.field("health") == SUS).all() and (b.field("swap") == UNDEF).all()
    assert (b.field("doi") == 1000).all()


def test_expired_without_swap_is_an_error(make):                             # :105-108
    b = make(seed=2, popsize=256, modeltime=20)
    b.initialize()
    b.set_field("exp", 7, 0)
    with pytest.raises(Exception, match="swap expired, but no swap set"):
        b.time_update()


@pytest.mark.parametrize("h", list(range(1, 12)))
def test_forced_swaps_land_in_state(make, h):                                # :112-129 (h in 2:9 there; all 11 here)
    b = make(seed=4, popsize=2000, modeltime=30)
    b.initialize()
    who = list(range(0, 100))
    for i in who:
        b.set_field("swap", i, h)
        b.set_field("exp", i, 0)
    b.time_update()
    health, swap = b.field("health"), b.field("swap")
    assert (health[who] == h).all()
    assert (b.field("tis")[who] == 0).all()
    if h in (LAT, PRE, ASYMP, MILD, MISO, INF, IISO, HOS, ICU):
        assert (swap[who] != UNDEF).all()
    else:
        assert (swap[who] == UNDEF).all() and (b.field("exp")[who] == 999).all()
    assert (health[100:] == SUS).all()
    dur, exp = b.field("dur"), b.field("exp")
    if h == LAT:                                                             # :138-143
        assert np.isin(swap[who], (ASYMP, PRE)).all() and (b.field("doi")[who] == 0).all()
        assert (exp[who] == dur[who, 0]).all() and (b.field("iso")[who] == 0).all()
    if h == PRE:                                                             # :152-154
        assert np.isin(swap[who], (MILD, INF)).all() and (exp[who] == dur[who, 2]).all()
    if h == ASYMP:
        assert (swap[who] == REC).all() and (exp[who] == dur[who, 1]).all()
    if h == DED:
        assert (b.field("iso")[who] == 1).all() and (b.field("nextday_meetcnt")[who] == 0).all()    # :216-219
    if h in (HOS, ICU):
        assert (b.field("iso")[who] == 1).all() and np.isin(swap[who], (REC, DED)).all()
        lo, hi = (8, 17) if h == HOS else (10, 19)
        rec = swap[who] == REC
        assert ((exp[who][rec] >= lo) & (exp[who][rec] <= hi)).all()


def test_mild_to_miso_through_fmild_taumild(make):                           # :157-175
    b = make(seed=6, popsize=1000, modeltime=30, fmild=0.0)
    b.initialize()
    for i in range(50):
        b.set_field("swap", i, MILD); b.set_field("exp", i, 0)
    b.time_update()
    assert (b.field("swap")[:50] == REC).all()                              # fmild = 0 => REC
    b = make(seed=6, popsize=1000, modeltime=30, fmild=1.0, τmild=1)
    b.initialize()
    for i in range(50):
        b.set_field("swap", i, MILD); b.set_field("exp", i, 0)
    b.time_update()
    assert (b.field("health")[:50] == MILD).all() and (b.field("swap")[:50] == MISO).all() and (b.field("exp")[:50] == 1).all()
    b.time_update()
    assert (b.field("health")[:50] == MISO).all() and (b.field("swap")[:50] == REC).all()
    assert (b.field("iso")[:50] == 1).all() and (b.field("isovia")[:50] == VIA_MI).all()
    assert (b.field("exp")[:50] == b.field("dur")[:50, 3] - 1).all()        # :515


def test_quarantine_group(make):                                             # :179-192 (eldq/eldqag = quar_grp_frac/quar_grp)
    b = make(seed=7, quar_grp_frac=1.0, quar_grp=4)
    b.initialize()
    ag, iso, via = b.field("ag"), b.field("iso"), b.field("isovia")
    assert (iso[ag == 4] == 1).all() and (via[ag == 4] == VIA_QU).all()
    assert (iso[ag != 4] == 0).all() and (via[ag != 4] == VIA_NULL).all()
    assert (b.field("nextday_meetcnt")[ag == 4] <= 3).all()                 # :209-214


def test_fpreiso_isolates_presymptomatic(make):                              # :196-206
    b = make(seed=8, popsize=500, modeltime=30, fpreiso=1.0, tpreiso=1)
    b.initialize()
    b.insert_infected(PRE, 500, 1)
    assert (b.field("health") == PRE).all()
    assert (b.field("iso") == 1).all() and (b.field("isovia") == VIA_PI).all()
    b.time_update()
    assert (b.field("nextday_meetcnt") <= 3).all()


def test_tpreiso_zero_never_activates_in_loop(make):                         # SURVEY Q1 (:124-125,:131-135)
    b = make(seed=8, popsize=500, modeltime=30, fpreiso=1.0, tpreiso=0)
    b.initialize()
    for i in range(100):
        b.set_field("swap", i, PRE); b.set_field("exp", i, 0)
    b.time_update()
    assert (b.field("health")[:100] == PRE).all() and (b.field("iso")[:100] == 0).all()


def test_isolated_and_dead_contact_counts(make):                             # :209-219
    b = make(seed=9, popsize=2000, modeltime=60)
    b.initialize()
    for i in range(200):
        b.set_field("iso", i, 1)
    for d in range(5):
        b.time_update()
        c = b.field("nextday_meetcnt")
        assert (c[:200] <= 3).all() and (c[:200] >= 0).all()
    assert c[200:].max() > 3


# ---- testset "tranmission" (:222-259)
def test_dyntrans_accounting(make):
    b = make(seed=10)                                                        # β = 0 default
    b.initialize()
    before = b.field("nextday_meetcnt").copy()
    assert b.dyntrans() == (0, 0)                                            # :231-233 nobody infectious
    b.insert_infected(PRE, 5, 1)
    tm, ti = b.dyntrans()
    assert ti == 0                                                           # β = 0
    after = b.field("nextday_meetcnt")
    assert tm == int((before - after).sum()) and tm > 0                      # :236-242
    assert (after <= before).all()


def test_dyntrans_infects_with_beta_one(make):
    b = make(seed=10, β=1.0)
    b.initialize()
    b.insert_infected(PRE, 5, 1)
    tm, ti = b.dyntrans()
    swap, health, exp, tis = b.field("swap"), b.field("health"), b.field("exp"), b.field("tis")
    newly = (health == SUS) & (swap == LAT)
    assert ti == int(newly.sum()) and ti > 0
    assert (exp[newly] == tis[newly]).all()                                  # :826
    assert (b.field("sickfrom")[newly] == PRE).all()                         # :827
    b.time_update()
    assert (b.field("health")[newly] == LAT).all() and (b.field("doi")[newly] == 0).all()


def test_betavalue_multipliers():                                            # :251-258 (host-side table; oracle only)
    b = OracleBackend(seed=1, beta=1.0, seasonal=False, frelasymp=0.11)
    assert b.betavalue(1, ASYMP) == 1.0 * 0.11 and b.betavalue(1, MILD) == 1.0 * 0.44 and b.betavalue(1, INF) == 1.0 * 0.89
    assert b.betavalue(1, MISO) == 0.44 and b.betavalue(1, IISO) == 0.89 and b.betavalue(1, PRE) == 1.0
    b.close()


def test_betas_seasonal_and_flat():                                          # :79-88
    s = O.Sim(O.make_params(beta=1.0, seasonal=False, modeltime=300), 1)
    assert s.betas.sum() == 300
    s.close()
    s = O.Sim(O.make_params(beta=1.0, seasonal=True, modeltime=300), 1)
    t = np.arange(1, 301)
    temp = 6.261 + -11.81 * np.cos((80 - t) * 0.022) + 1.817 * np.sin((80 - t) * 0.022)
    temp = (temp - 2.5 * temp.min()) / (temp.max() - temp.min())
    assert np.allclose(s.betas, temp, rtol=0, atol=1e-12)
    s.close()


# ---- testset "calibration" (:261-292)
def test_calibration_latents_never_progress(make):
    b = make(seed=12, β=1.0, prov="ontario", calibration=True, initialinf=1)
    b.initialize()
    b.insert_infected(PRE, 1, 4)
    h = np.nonzero(b.field("health") == PRE)[0]
    assert len(h) == 1 and b.field("swap")[h[0]] in (ASYMP, MILD, INF)
    for _ in range(20):
        b.dyntrans(); b.time_update()
    health = b.field("health")
    assert health[h[0]] == REC and b.field("exp")[h[0]] == 999
    assert not np.isin(health, (PRE, ASYMP, MILD, MISO, INF, IISO, HOS, ICU, DED)).any()
    assert (health == LAT).sum() > 0
    lat = health == LAT
    assert (b.field("swap")[lat] == LAT).all() and (b.field("exp")[lat] == 999).all()        # :457-460


# ---- testset "contact trace" (:294-404)
def test_contact_trace_window_and_quarantine(make):
    b = make(seed=13, ctstrat=1, fctcapture=1.0, fcontactst=1.0, cdaysback=3, mild_props=(0, 0, 0, 0, 0),
             asymp_props=(0, 0, 0, 0, 0))
    b.initialize()
    assert b.field("tracestart")[0] == -1 and b.field("traceend")[0] == -1
    # asymptomatic-bound latent gets no window (:320-322): agent 1 is forced LAT with swap ASYMP on the same day
    b.set_field("swap", 0, LAT); b.set_field("exp", 0, 0)
    b.time_update()
    x = 0
    dur = b.field("dur")[x]
    assert b.field("health")[x] == LAT and b.field("swap")[x] == PRE and b.field("doi")[x] == 0
    ts, te = b.field("tracestart")[x], b.field("traceend")[x]
    assert ts > 0 and te > dur[0] + dur[2] and te <= dur[0] + dur[2] + dur[3]            # :326-330
    assert ts == max(0, te - 3)
    traced_days = 0
    for i in range(2, te + 3):
        b.dyntrans(); b.time_update()
        assert b.field("doi")[x] == i - 1                                                 # :347 (one update already done)
        if b.field("health")[x] == LAT:
            assert b.field("tracestart")[x] == ts and b.field("traceend")[x] == te and b.field("ncontacts")[x] == 0
        assert b.field("health")[x] not in (ASYMP, MILD)
        doi = b.field("doi")[x]
        if ts <= doi < te:
            assert b.field("tracing")[x] == 1
            traced_days += 1
        else:
            assert b.field("tracing")[x] == 0
    assert traced_days == 3                                                               # :375 == cdaysback
    iso = np.nonzero(b.field("iso") == 1)[0]
    assert x in iso and len(iso) > 1
    via, txp = b.field("isovia"), b.field("tracedxp")
    assert (via[iso] == VIA_CT).all() and (txp[iso] == 12).all()                          # two updates after identification
    assert b.totalisolated() >= len(iso)
    for t in range(1, 12):
        b.time_update()
        assert (b.field("tracedxp")[iso] == 12 - t).all() and (b.field("iso")[iso] == 1).all()
    b.time_update()
    assert (b.field("tracedxp")[iso] == 0).all() and (b.field("iso")[iso] == 0).all() and (b.field("isovia")[iso] == VIA_NULL).all()


def test_asymptomatic_latent_gets_no_window(make):                           # :320-322
    b = make(seed=14, popsize=500, modeltime=40, ctstrat=1, fctcapture=1.0, fcontactst=1.0, cdaysback=3,
             asymp_props=(1, 1, 1, 1, 1))
    b.initialize()
    for i in range(50):
        b.set_field("swap", i, LAT); b.set_field("exp", i, 0)
    b.time_update()
    assert (b.field("swap")[:50] == ASYMP).all()
    assert (b.field("tracestart")[:50] == -1).all() and (b.field("traceend")[:50] == -1).all()


def test_invalid_parameters(make):                                           # :164-165, :334
    with pytest.raises(Exception, match="invalid ct strategy"):
        make(ctstrat=2)
    with pytest.raises(Exception, match="canadian provinces"):
        make(prov="noknownlocation")
