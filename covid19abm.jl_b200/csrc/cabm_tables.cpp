// cabm_tables.cpp — host-side construction of the draw-contract tables.
//
// Every continuous variate of the reference is rounded to an integer as soon as it is drawn
// (src/covid19abm.jl:352-356, :546, :590-593) and the contact count is a negative binomial (:781,
// :853-866), so each becomes an inverse-CDF table over one 32-bit uniform word w:
//     value = base + #{k : thr[k] <= w},  thr[k] = floor(CDF(base+k) * 2^32 + 1/2).
// CDFs are evaluated here in long double (x87 80-bit on the host) with own series / continued-fraction
// code; tests/test_tables.py checks the result against the mpmath-generated oracle tables.
#include "cabm_tables.h"

#include <cmath>
#include <cstring>

namespace cabm {

typedef long double ld;

// regularised lower incomplete gamma P(a, x)
static ld gamma_p(ld a, ld x) {
    if (x <= 0) return 0;
    const ld eps = 1e-21L;
    ld lg = lgammal(a);
    if (x < a + 1) {                       // series
        ld ap = a, sum = 1 / a, del = sum;
        for (int n = 0; n < 100000; ++n) {
            ap += 1; del *= x / ap; sum += del;
            if (fabsl(del) < fabsl(sum) * eps) break;
        }
        return sum * expl(-x + a * logl(x) - lg);
    }
    // modified Lentz continued fraction for Q(a, x)
    const ld tiny = 1e-4000L;
    ld b = x + 1 - a, c = 1 / tiny, d = 1 / b, h = d;
    for (int i = 1; i < 100000; ++i) {
        ld an = -(ld)i * ((ld)i - a);
        b += 2;
        d = an * d + b; if (fabsl(d) < tiny) d = tiny;
        c = b + an / c; if (fabsl(c) < tiny) c = tiny;
        d = 1 / d;
        ld del = d * c; h *= del;
        if (fabsl(del - 1) < eps) break;
    }
    return 1 - expl(-x + a * logl(x) - lg) * h;
}
static ld gamma_cdf(ld shape, ld scale, ld x) { return x <= 0 ? 0 : gamma_p(shape, x / scale); }
static ld norm_cdf(ld z) { return 0.5L * erfcl(-z / sqrtl(2.0L)); }

static void push_thresholds(Table &t, const std::vector<ld> &cum) {
    const ld two32 = 4294967296.0L;
    for (ld c : cum) {
        ld v = floorl(c * two32 + 0.5L);
        if (v >= two32) break;
        t.thr.push_back((uint32_t)v);
    }
}

template <class F> static Table rounded_truncated(F cdf, ld lo, ld hi) {
    Table t;
    int kmin = (int)floorl(lo + 0.5L), kmax = (int)floorl(hi + 0.5L);
    t.base = kmin;
    ld flo = cdf(lo), z = cdf(hi) - flo;
    std::vector<ld> cum;
    for (int k = kmin; k <= kmax; ++k) {
        ld upper = (ld)k + 0.5L; if (upper > hi) upper = hi;
        cum.push_back((cdf(upper) - flo) / z);
    }
    push_thresholds(t, cum);
    return t;
}
static Table ceil_gamma(ld shape, ld scale) {
    Table t; t.base = 1;
    std::vector<ld> cum;
    for (int k = 1; k < 400; ++k) cum.push_back(gamma_cdf(shape, scale, (ld)k));
    push_thresholds(t, cum);
    return t;
}
static Table negbin(double mean, double sd) {
    // r and p exactly as the reference computes them in Float64 (:860-861)
    double p_d = 1 - (sd * sd - mean) / (sd * sd);
    double r_d = mean * mean / (sd * sd - mean);
    ld p = p_d, r = r_d, acc = 0, term = powl(p, r);
    Table t; t.base = 0;
    std::vector<ld> cum;
    for (int k = 0; k < 400; ++k) {
        acc += term; cum.push_back(acc);
        term = term * ((ld)k + r) / ((ld)k + 1) * (1 - p);
    }
    push_thresholds(t, cum);
    return t;
}

const double CONTACT_MATRIX[5][5] = {   // contact_matrix :838-848
    {0.2287, 0.1839, 0.4219, 0.1116, 0.0539}, {0.0276, 0.5964, 0.2878, 0.0591, 0.0291},
    {0.0376, 0.1454, 0.6253, 0.1423, 0.0494}, {0.0242, 0.1094, 0.4867, 0.2723, 0.1074},
    {0.0207, 0.1083, 0.4071, 0.2193, 0.2446} };

const double PROV_AG[15][5] = {    // get_province_ag :317-337
    {0.0655, 0.1851, 0.4331, 0.1933, 0.1230}, {0.0475, 0.1570, 0.3905, 0.2223, 0.1827},
    {0.0540, 0.1697, 0.3915, 0.2159, 0.1689}, {0.0634, 0.1918, 0.3899, 0.1993, 0.1556},
    {0.0460, 0.1563, 0.3565, 0.2421, 0.1991}, {0.0430, 0.1526, 0.3642, 0.2458, 0.1944},
    {0.0747, 0.2026, 0.4511, 0.1946, 0.0770}, {0.0455, 0.1549, 0.3601, 0.2405, 0.1990},
    {0.1157, 0.2968, 0.4321, 0.1174, 0.0380}, {0.0519, 0.1727, 0.3930, 0.2150, 0.1674},
    {0.0490, 0.1702, 0.3540, 0.2329, 0.1939}, {0.0545, 0.1615, 0.3782, 0.2227, 0.1831},
    {0.0666, 0.1914, 0.3871, 0.1997, 0.1552}, {0.0597, 0.1694, 0.4179, 0.2343, 0.1187},
    {0.064000, 0.163000, 0.448000, 0.181000, 0.144000} };

void gpw_split(int ag, int cnts, int out[5]) {   // :804, Julia round = half-to-even = nearbyint
    for (int g = 0; g < 5; ++g) out[g] = (int)std::nearbyint(CONTACT_MATRIX[ag - 1][g] * (double)cnts);
}

const Tables &tables() {
    static Tables T = [] {
        Tables t;
        const ld mu_ln = logl((ld)5.2), sg = (ld)0.1;
        t.lat = rounded_truncated([&](ld x) { return norm_cdf((logl(x) - mu_ln) / sg); }, 4, 7);       // :347
        const ld pre_scale = (ld)(5 / 2.3);
        t.pre = rounded_truncated([&](ld x) { return gamma_cdf((ld)1.058, pre_scale, x); }, (ld)0.8, 3); // :348
        t.asy = ceil_gamma(5, 1);                                                                      // :349
        t.inf = ceil_gamma((ld)((3.2) * (3.2) / 3.7), (ld)(3.7 / 3.2));                                // :350
        t.psi = rounded_truncated([&](ld x) { return gamma_cdf((ld)4.5, (ld)2.75, x); }, 8, 17);       // :590
        t.mu = rounded_truncated([&](ld x) { return gamma_cdf((ld)5.3, (ld)2.1, x); }, 9, 15);         // :592
        const double means[5] = {10.21, 16.793, 13.7950, 11.2669, 8.0027};                             // :855
        const double sds[5] = {7.65, 11.7201, 10.5045, 9.5935, 6.9638};                                // :856
        for (int g = 0; g < 5; ++g) t.nb[g] = negbin(means[g], sds[g]);
        return t;
    }();
    return T;
}

const Table *table_by_name(const char *name) {
    const Tables &t = tables();
    if (!strcmp(name, "lat")) return &t.lat;
    if (!strcmp(name, "pre")) return &t.pre;
    if (!strcmp(name, "asy")) return &t.asy;
    if (!strcmp(name, "inf")) return &t.inf;
    if (!strcmp(name, "psi")) return &t.psi;
    if (!strcmp(name, "mu")) return &t.mu;
    if (!strncmp(name, "nb", 2) && name[2] >= '1' && name[2] <= '5' && !name[3]) return &t.nb[name[2] - '1'];
    return nullptr;
}

uint64_t bernoulli_threshold(double p) {
    // rand() < p with rand() = w * 2^-32  <=>  w < ceil(p * 2^32)   (p * 2^32 is exact in double)
    if (!(p > 0)) return 0;
    double x = std::ceil(p * 4294967296.0);
    if (x >= 4294967296.0) return 4294967296ull;
    return (uint64_t)x;
}

}  // namespace cabm
