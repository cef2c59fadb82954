"""No-GPU checks of the boundary: the C-ABI library loads and exports every symbol include/cabm.h declares, the
parameter struct layouts agree, the host mirror keeps the reference's ModelParameters semantics, and the product
path fails loudly (never falls back) when there is no CUDA device."""
import ctypes as C
import os
import re

import numpy as np
import pytest

import oracle_api as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "cabm.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(cabm_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol(cv):
    L = cv.lib()
    names = declared_symbols()
    assert len(names) >= 30
    for n in names:
        assert hasattr(L, n), n
    assert b"sm_100a" in L.cabm_version()


def test_param_struct_layouts_agree(cv):
    assert C.sizeof(cv.CabmParams) == C.sizeof(O.OrcParams) == 192
    for (n1, _), (n2, _) in zip(cv.CabmParams._fields_, O.OrcParams._fields_):
        assert getattr(cv.CabmParams, n1).offset == getattr(O.OrcParams, n2).offset


def test_model_parameters_defaults_and_aliases(cv):
    p = cv.ModelParameters()
    # defaults of src/covid19abm.jl:31-55
    assert (p.β, p.frelasymp, p.seasonal, p.popsize, p.prov, p.calibration, p.modeltime) == (0.0, 0.11, False, 10000, "newyork", False, 300)
    assert (p.initialinf, p.initialhi, p.τmild, p.fmild, p.fsevere, p.tpreiso, p.fpreiso) == (1, 0, 0, 0.0, 0.0, 0, 0.0)
    assert (p.quar_grp, p.quar_grp_frac, p.ctstrat, p.fctcapture, p.fcontactst, p.cdaysback, p.strat3qdays) == (5, 0.0, 0, 0.0, 0.0, 0, 0)
    assert p.asymp_props == (0.25, 0.25, 0.14, 0.07, 0.07) and p.mild_props == (0.95, 0.9, 0.85, 0.6, 0.2)
    q = cv.ModelParameters(beta=0.5, tau_mild=2, eldq=0.25, eldqag=4, fpre=1.0, cidtime=3, prov=":ontario")
    assert (q.β, q.τmild, q.quar_grp_frac, q.quar_grp) == (0.5, 2, 0.25, 4)
    c = q.to_c()
    assert c.prov_id == 9 and c.beta == 0.5 and c.tau_mild == 2 and c.quar_grp == 4
    with pytest.raises(TypeError):
        cv.ModelParameters(nosuchfield=1)
    # Float16 fields (:51-52): 0.3 is not representable in binary16
    assert cv.ModelParameters(fctcapture=0.3).to_c().fctcapture == float(np.float16(0.3)) == 0.300048828125


def test_collectdf_secondarys(cv):
    """_collectdf :226-245: secondarys = LAT_INC ./ sum(prev[PRE..IISO]) with NaN/Inf -> 0."""
    inc = np.zeros((4, 12), dtype=np.int32)
    prev = np.zeros((4, 12), dtype=np.int32)
    inc[:, cv.LAT] = [0, 3, 5, 2]
    prev[1, cv.PRE], prev[1, cv.INF] = 2, 4
    prev[3, cv.IISO], prev[3, cv.HOS] = 4, 100
    df = cv.collectdf(inc, prev)
    assert list(df["secondarys"]) == [0.0, 0.5, 0.0, 0.5]
    assert list(df["time"]) == [1, 2, 3, 4]
    assert "LAT_INC" in df and "DED_PREV" in df and len([k for k in df if k.endswith("_INC")]) == 12


def test_no_cpu_fallback_without_gpu(cv):
    """Without a CUDA device the product path must raise, not compute on the CPU."""
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        pytest.skip("a GPU is present")
    with pytest.raises(cv.CabmError, match="no CPU fallback"):
        cv.main(cv.ModelParameters(popsize=100, modeltime=5))


def test_product_does_not_reference_the_oracle():
    """oracle/ is test infrastructure: nothing under the package may import, include or link it."""
    pkg = os.path.join(ROOT, "covid19abm.jl_b200")
    for d, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h", "Makefile")):
                txt = open(os.path.join(d, f), errors="ignore").read()
                assert "oracle_api" not in txt and "abm_oracle" not in txt and "liboracle" not in txt, os.path.join(d, f)
