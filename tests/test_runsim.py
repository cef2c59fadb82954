"""runsim(simnum, ip) (src/covid19abm.jl:77-101) — the caller everyone uses — against the oracle: seed simnum*726 (:78), the
six frames a, g1..g5 with time / secondarys / <S>_INC / <S>_PREV / sim (:226-245, :91-97), infectors (:295-313) and
ct_numbers (:86-87); and the secondarys_all.dat file of scripts/local_run.jl:67-69."""
import os

import numpy as np
import pytest

from test_gpu_parity import to_oracle

pytestmark = pytest.mark.gpu
STATES = ["SUS", "LAT", "PRE", "ASYMP", "MILD", "MISO", "INF", "IISO", "HOS", "ICU", "REC", "DED"]


def oracle_runsim(O, simnum, ip):
    sim = O.Sim(to_oracle(O, ip), simnum * 726)
    hm = sim.main()
    ags = sim.field("ag")
    inc, prev = O.collect_curves(hm, ags)
    out = dict(inc=inc, prev=prev, infectors=tuple(int(x) for x in sim.count_infectors()), totalisolated=int(sim.totalisolated), hm=np.asarray(hm))
    sim.close()
    return out


def secondarys(inc, prev):
    """:237-241 literally: LAT_INC ./ sum of the PRE..IISO prevalence columns, NaN and Inf replaced by 0."""
    T = inc.shape[0]
    sec = np.zeros(T)
    for t in range(T):
        den = int(prev[t, 2:8].sum())
        num = int(inc[t, 1])
        sec[t] = 0.0 if den == 0 else num / den
    return sec


@pytest.mark.parametrize("simnum,kw", [
    (1, dict(β=0.0525, prov="newyork")),                                                                  # the reference's own timed scenario (test/runtests.jl:408-428)
    (7, dict(β=0.0345, prov="ontario", modeltime=500, fmild=0.5, τmild=1, fsevere=1.0, fpreiso=0.3)),
    (3, dict(β=0.06, prov="ontario", modeltime=200, initialinf=3, ctstrat=1, fctcapture=0.5, fcontactst=0.5, cdaysback=3)),
], ids=["reference-main-run", "configB", "tracing"])
def test_runsim_matches_the_oracle(cv, O, simnum, kw):
    ip = cv.ModelParameters(**kw)
    res = cv.runsim(simnum, ip)
    ref = oracle_runsim(O, simnum, ip)
    assert (np.asarray(res["hmatrix"]) == ref["hm"]).all()
    for k, name in enumerate(["a", "g1", "g2", "g3", "g4", "g5"]):
        df = res[name]
        assert (df["time"] == np.arange(1, ip.modeltime + 1)).all() and (df["sim"] == simnum).all()
        for s, st in enumerate(STATES):
            assert (df[st + "_INC"] == ref["inc"][k, :, s]).all(), (name, st)
            assert (df[st + "_PREV"] == ref["prev"][k, :, s]).all(), (name, st)
        sec = secondarys(ref["inc"][k], ref["prev"][k])
        assert np.isfinite(df["secondarys"]).all()
        assert np.array_equal(df["secondarys"], sec), name
    assert res["infectors"] == ref["infectors"]
    assert res["ct_numbers"] == (0, 0, ref["totalisolated"], 0, 0, 0, 0)
    if ip.ctstrat:
        assert ref["totalisolated"] > 0
    # column order of _collectdf (:235-244): time, secondarys, 12 x _INC, 12 x _PREV, then sim (:92)
    assert list(res["a"].keys()) == ["time", "secondarys"] + [s + "_INC" for s in STATES] + [s + "_PREV" for s in STATES] + ["sim"]


def test_secondarys_file_matches_the_oracle(cv, O, tmp_path):
    local_run = cv.local_run
    ip = cv.ModelParameters(β=0.07, prov="ontario", modeltime=80, initialinf=2)
    nsims = 6
    local_run.run(ip, nsims=nsims, folderprefix=str(tmp_path))
    rows = [l.strip().split(",") for l in open(os.path.join(tmp_path, "secondarys_all.dat"))]
    assert rows[0] == ["time"] + [str(i) for i in range(1, nsims + 1)]
    data = np.array([[float(x) for x in r] for r in rows[1:]])
    assert data.shape == (80, nsims + 1) and (data[:, 0] == np.arange(1, 81)).all()
    for i in range(1, nsims + 1):
        ref = oracle_runsim(O, i, ip)
        assert np.array_equal(data[:, i], secondarys(ref["inc"][0], ref["prev"][0])), i
