"""Pins the draw contract on the CPU: Philox known answers, and every variate table of the oracle
(mpmath) against (i) the product's own long-double construction, (ii) scipy's distributions / the reference's
published moments. No GPU needed."""
import numpy as np
import pytest
from scipy import stats

import oracle_api as O

# Random123 kat_vectors, philox4x32-10
KAT = [([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
       ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
       ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1])]
NAMES = ["lat", "pre", "asy", "inf", "psi", "mu", "nb1", "nb2", "nb3", "nb4", "nb5"]


@pytest.mark.parametrize("ctr,key,out", KAT)
def test_philox_known_answers(ctr, key, out):
    assert [int(x) for x in O.philox(ctr, key)] == out


def pmf(name):
    base, thr = O.table(name)
    cum = np.concatenate([thr.astype(np.float64), [2.0 ** 32]]) / 2.0 ** 32
    return base, np.diff(np.concatenate([[0.0], cum]))


@pytest.mark.parametrize("name", NAMES)
def test_product_tables_equal_oracle_tables(cv, name):
    """cabm_get_table (C++ long double) == oracle_tables.h (mpmath, 60 digits), threshold for threshold."""
    b1, t1 = cv.get_table(name)
    b2, t2 = O.table(name)
    assert b1 == b2 and len(t1) == len(t2) and (t1 == t2).all()


def test_negative_binomial_tables_match_reference_moments():
    means = [10.21, 16.793, 13.7950, 11.2669, 8.0027]           # src/covid19abm.jl:855-856
    sds = [7.65, 11.7201, 10.5045, 9.5935, 6.9638]
    for g in range(5):
        base, p = pmf(f"nb{g + 1}")
        k = base + np.arange(len(p))
        m = (k * p).sum()
        assert abs(m - means[g]) < 1e-6 and abs(np.sqrt(((k - m) ** 2 * p).sum()) - sds[g]) < 1e-5
        pr = means[g] / sds[g] ** 2
        r = means[g] ** 2 / (sds[g] ** 2 - means[g])
        ref = stats.nbinom.pmf(k, r, pr)
        assert np.abs(p - ref).max() < 1e-9
        assert len(p) <= 256                                     # contact budgets fit the 8-bit field of the status word


def test_duration_tables_match_scipy():
    # lat: round(truncated(LogNormal(log 5.2, 0.1), 4, 7))  :347,352
    d = stats.lognorm(s=0.1, scale=5.2)
    z = d.cdf(7) - d.cdf(4)
    base, p = pmf("lat")
    ref = [(d.cdf(4.5) - d.cdf(4)) / z, (d.cdf(5.5) - d.cdf(4.5)) / z, (d.cdf(6.5) - d.cdf(5.5)) / z, (d.cdf(7) - d.cdf(6.5)) / z]
    assert base == 4 and np.abs(p - ref).max() < 1e-9
    assert np.abs(p - [0.070, 0.642, 0.276, 0.011]).max() < 2e-3          # SURVEY §8(c) known answers
    # pre: round(truncated(Gamma(1.058, 5/2.3), 0.8, 3))  :348,353
    d = stats.gamma(a=1.058, scale=5 / 2.3)
    z = d.cdf(3) - d.cdf(0.8)
    base, p = pmf("pre")
    ref = [(d.cdf(1.5) - d.cdf(0.8)) / z, (d.cdf(2.5) - d.cdf(1.5)) / z, (d.cdf(3) - d.cdf(2.5)) / z]
    assert base == 1 and np.abs(p - ref).max() < 1e-9
    # asy: ceil(Gamma(5,1)), inf: ceil(Gamma(3.2^2/3.7, 3.7/3.2))  :349-350,355-356
    for name, d in (("asy", stats.gamma(a=5, scale=1)), ("inf", stats.gamma(a=3.2 ** 2 / 3.7, scale=3.7 / 3.2))):
        base, p = pmf(name)
        k = base + np.arange(len(p))
        assert base == 1 and np.abs(p - (d.cdf(k) - d.cdf(k - 1))).max() < 1e-9
    # psi: round(truncated(Gamma(4.5, 2.75), 8, 17)), mu: round(truncated(Gamma(5.3, 2.1), 9, 15))  :590-593
    for name, d, lo, hi in (("psi", stats.gamma(a=4.5, scale=2.75), 8, 17), ("mu", stats.gamma(a=5.3, scale=2.1), 9, 15)):
        base, p = pmf(name)
        z = d.cdf(hi) - d.cdf(lo)
        ref = [(d.cdf(min(hi, k + 0.5)) - d.cdf(max(lo, k - 0.5))) / z for k in range(lo, hi + 1)]
        assert base == lo and len(p) == hi - lo + 1 and np.abs(p - ref).max() < 1e-9


def test_gpw_split_is_half_to_even_and_matches_product(cv):
    cm = np.array([[0.2287, 0.1839, 0.4219, 0.1116, 0.0539], [0.0276, 0.5964, 0.2878, 0.0591, 0.0291],
                   [0.0376, 0.1454, 0.6253, 0.1423, 0.0494], [0.0242, 0.1094, 0.4867, 0.2723, 0.1074],
                   [0.0207, 0.1083, 0.4071, 0.2193, 0.2446]])                 # :842-846
    for ag in range(1, 6):
        for cnt in range(0, 256):
            ref = np.rint(cm[ag - 1] * cnt).astype(np.int32)                  # numpy rint is half-to-even like Julia's round
            assert (O.gpw(ag, cnt) == ref).all() and (cv.get_gpw(ag, cnt) == ref).all()
    assert (O.gpw(1, 1) == 0).all()                                           # group 1 with budget 1 meets nobody


def test_oracle_is_deterministic_and_seed_sensitive():
    p = O.make_params(beta=0.08, prov="ontario", modeltime=60, initialinf=3)
    a = O.Sim(p, 77).main().copy()
    b = O.Sim(p, 77).main().copy()
    c = O.Sim(p, 78).main().copy()
    assert (a == b).all() and (a != c).any()
    assert set(np.unique(a)) <= set(range(12))
    assert (a[:, 0] != 0).sum() == 3                                          # column 1 is the post-seeding state
