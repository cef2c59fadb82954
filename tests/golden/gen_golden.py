#!/usr/bin/env python3
"""Generates tests/golden/oracle_golden.npz — known-answer vectors of the draw contract.

The reference holds no golden vectors (test/runtests.jl is invariants only) and cannot run here (no julia), so these
are produced by the CPU oracle (oracle/abm_oracle.c) and committed: they freeze the keyed-draw contract, so that any
later change to the oracle, the tables or the Philox keying shows up as a diff against history, and they give the GPU
parity tests a fixture that does not depend on rebuilding the oracle. Regenerate only when the contract changes on purpose:
    python tests/golden/gen_golden.py
"""
import hashlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import oracle_api as O  # noqa: E402

CASES = {
    "plain": dict(beta=0.12, prov="ontario", popsize=2000, modeltime=80, initialinf=3),
    "isolation": dict(beta=0.15, prov="ontario", popsize=2000, modeltime=80, initialinf=3, fmild=0.5, tau_mild=1, fsevere=1.0, fpreiso=0.3, tpreiso=1),
    "tracing1": dict(beta=0.2, prov="newyork", popsize=2000, modeltime=80, initialinf=5, ctstrat=1, fctcapture=0.6, fcontactst=0.6, cdaysback=3, fmild=0.3, tau_mild=1),
    "tracing3": dict(beta=0.2, prov="bc", popsize=2000, modeltime=80, initialinf=5, ctstrat=3, strat3qdays=6, fctcapture=0.6, fcontactst=0.6, cdaysback=2),
    "calibration": dict(beta=0.3, prov="ontario", popsize=2000, modeltime=50, initialinf=1, calibration=True),
    "quarantine_hi_seasonal": dict(beta=0.25, prov="quebec", popsize=2000, modeltime=80, initialinf=4, initialhi=300, quar_grp_frac=0.5, quar_grp=5, seasonal=True, frelasymp=0.3),
}
SEEDS = (726, 1452)


def run_case(kw, seed):
    sim = O.Sim(O.make_params(**kw), seed)
    hm = np.ascontiguousarray(sim.main().T)                      # [T][N] int16, the hmatrix storage order
    inc, prev = O.collect_curves(hm.T, sim.field("ag"))
    out = dict(sha=hashlib.sha256(hm.astype(np.int8).tobytes()).hexdigest(), inc=inc.astype(np.int16), prev=prev.astype(np.int16),
               infectors=sim.count_infectors(), totalisolated=np.int64(sim.totalisolated), last_column=hm[-1].astype(np.int8),
               ag=sim.field("ag").astype(np.int8))
    sim.close()
    return out


def main():
    blob = {}
    for name, kw in CASES.items():
        for seed in SEEDS:
            for k, v in run_case(kw, seed).items():
                blob[f"{name}/{seed}/{k}"] = np.array(v)
    # first words of a few keyed draws (Philox counter/key layout of the contract)
    blob["philox/ctr_5_3_4_1_key_726"] = O.philox([5, 3, 4, 1], [726, 0])
    np.savez_compressed(os.path.join(HERE, "oracle_golden.npz"), **blob)
    print("wrote", len(blob), "arrays")


if __name__ == "__main__":
    main()
