"""Committed known-answer vectors (tests/golden/oracle_golden.npz, made by tests/golden/gen_golden.py): the oracle must
reproduce them on the CPU (freezes the draw contract), and the CUDA path must reproduce them on the GPU without the
oracle in the loop."""
import hashlib
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import gen_golden as G  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.npz"))
CASES = [(n, s) for n in G.CASES for s in G.SEEDS]


@pytest.mark.parametrize("name,seed", CASES)
def test_oracle_reproduces_golden(name, seed):
    got = G.run_case(G.CASES[name], seed)
    for k, v in got.items():
        assert (np.array(v) == GOLD[f"{name}/{seed}/{k}"]).all(), k


def test_philox_counter_layout(O):
    assert (O.philox([5, 3, 4, 1], [726, 0]) == GOLD["philox/ctr_5_3_4_1_key_726"]).all()


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["fused", "multikernel"])
@pytest.mark.parametrize("name,seed", CASES)
def test_cuda_reproduces_golden(cv, name, seed, path):
    kw = dict(G.CASES[name])
    kw["β"] = kw.pop("beta")
    if "tau_mild" in kw:
        kw["τmild"] = kw.pop("tau_mild")
    ip = cv.ModelParameters(**kw)
    out = cv.run_ensemble(ip, 1, seeds=[seed], want_hmatrix=True, path=path)
    hm = out["hmatrix"][0]
    assert hashlib.sha256(hm.astype(np.int8).tobytes()).hexdigest() == str(GOLD[f"{name}/{seed}/sha"])
    assert (out["inc"][0] == GOLD[f"{name}/{seed}/inc"]).all() and (out["prev"][0] == GOLD[f"{name}/{seed}/prev"]).all()
    assert (out["infectors"][0] == GOLD[f"{name}/{seed}/infectors"]).all()
    assert out["totalisolated"][0] == GOLD[f"{name}/{seed}/totalisolated"]
    assert (hm[-1] == GOLD[f"{name}/{seed}/last_column"]).all()
