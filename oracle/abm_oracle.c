/* abm_oracle.c — CPU restatement of /root/reference/src/covid19abm.jl (see abm_oracle.h header).
 * TEST INFRASTRUCTURE ONLY — "parity unpinned" against Julia; pinned as described in abm_oracle.h.
 *
 * Structure mirrors the reference: array-of-structs `humans`, one global parameter block per sim,
 * strictly sequential agent loops in the reference's order. The only change is where random numbers
 * come from: every `rand(...)` call site is a keyed Philox4x32-10 draw, so that a
 * parallel implementation can reproduce the same trajectory bit for bit.
 */
#include "abm_oracle.h"
#include "oracle_tables.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------ draw contract */
enum { S_INIT = 1, S_COUNT = 2, S_SEED = 3, S_TRANS = 4, S_CTWIN = 5, S_CTCON = 6, S_CT3 = 7, S_DT = 8 };

static __thread char g_err[256];
const char *orc_last_error(void) { return g_err; }
static int fail(const char *msg) { snprintf(g_err, sizeof g_err, "%s", msg); return 1; }

void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    uint32_t c0 = ctr[0], c1 = ctr[1], c2 = ctr[2], c3 = ctr[3], k0 = key[0], k1 = key[1];
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1, n3 = (uint32_t)p0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

struct orc_sim {
    orc_params p;
    uint32_t key[2];
    orc_human *humans;
    double *betas;
    int64_t totalisolated;
    int day;                 /* st of the next dyntrans/time_update (1-based) */
    int *grp[5]; int grplen[5]; int grps_built;
};

/* Replay tape (include/cabm.h cabm_set_tape): a sorted table of draw words keyed by (seed, counter). A draw whose
 * key is on the tape takes the tape's four words instead of the Philox output; every other draw is unchanged. This is how an
 * external stream — e.g. the draws of an instrumented Julia run, dumped per logical draw site — is replayed. */
static orc_tape_entry *g_tape = NULL;
static uint64_t g_tape_n = 0;
static int tape_cmp(const void *a, const void *b) {
    const orc_tape_entry *x = (const orc_tape_entry *)a, *y = (const orc_tape_entry *)b;
    if (x->seed != y->seed) return x->seed < y->seed ? -1 : 1;
    for (int k = 0; k < 4; ++k) if (x->ctr[k] != y->ctr[k]) return x->ctr[k] < y->ctr[k] ? -1 : 1;
    return 0;
}
int orc_set_tape(const orc_tape_entry *entries, uint64_t n) {
    free(g_tape); g_tape = NULL; g_tape_n = 0;
    if (!entries || n == 0) return 0;
    g_tape = (orc_tape_entry *)malloc(sizeof(orc_tape_entry) * n);
    if (!g_tape) return fail("orc_set_tape: out of memory");
    memcpy(g_tape, entries, sizeof(orc_tape_entry) * n);
    qsort(g_tape, n, sizeof(orc_tape_entry), tape_cmp);
    g_tape_n = n;
    return 0;
}

static void draw4(const orc_sim *s, uint32_t c0, uint32_t day, uint32_t stream, uint32_t sub, uint32_t w[4]) {
    uint32_t ctr[4] = { c0, day, stream, sub };
    if (g_tape_n) {
        orc_tape_entry k;
        k.seed = (uint64_t)s->key[0] | ((uint64_t)s->key[1] << 32);
        memcpy(k.ctr, ctr, sizeof ctr);
        const orc_tape_entry *hit = (const orc_tape_entry *)bsearch(&k, g_tape, g_tape_n, sizeof(orc_tape_entry), tape_cmp);
        if (hit) { memcpy(w, hit->w, sizeof hit->w); return; }
    }
    orc_philox4x32_10(ctr, s->key, w);
}
static inline double u01(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }
static inline int below(uint32_t w, int n) { return (int)(((uint64_t)w * (uint64_t)n) >> 32); }
static inline int table_draw(uint32_t w, int base, const uint32_t *thr, int len) {
    int v = 0;
    while (v < len && thr[v] <= w) ++v;
    return base + v;
}

static const uint32_t *const nb_thr[5] = { orc_nb1_thr, orc_nb2_thr, orc_nb3_thr, orc_nb4_thr, orc_nb5_thr };
static const int nb_len[5] = { ORC_NB1_LEN, ORC_NB2_LEN, ORC_NB3_LEN, ORC_NB4_LEN, ORC_NB5_LEN };

int orc_table(const char *name, int *base, const uint32_t **thr) {
#define T(n, B, L, A) if (!strcmp(name, n)) { *base = B; *thr = A; return L; }
    T("lat", ORC_LAT_BASE, ORC_LAT_LEN, orc_lat_thr) T("pre", ORC_PRE_BASE, ORC_PRE_LEN, orc_pre_thr)
    T("asy", ORC_ASY_BASE, ORC_ASY_LEN, orc_asy_thr) T("inf", ORC_INF_BASE, ORC_INF_LEN, orc_inf_thr)
    T("psi", ORC_PSI_BASE, ORC_PSI_LEN, orc_psi_thr) T("mu", ORC_MU_BASE, ORC_MU_LEN, orc_mu_thr)
    T("nb1", 0, ORC_NB1_LEN, orc_nb1_thr) T("nb2", 0, ORC_NB2_LEN, orc_nb2_thr) T("nb3", 0, ORC_NB3_LEN, orc_nb3_thr)
    T("nb4", 0, ORC_NB4_LEN, orc_nb4_thr) T("nb5", 0, ORC_NB5_LEN, orc_nb5_thr)
#undef T
    return -1;
}

/* ------------------------------------------------------------------ constants */
/* get_province_ag :317-337 */
static const double PROV_AG[15][5] = {
    {0.0655, 0.1851, 0.4331, 0.1933, 0.1230}, {0.0475, 0.1570, 0.3905, 0.2223, 0.1827},
    {0.0540, 0.1697, 0.3915, 0.2159, 0.1689}, {0.0634, 0.1918, 0.3899, 0.1993, 0.1556},
    {0.0460, 0.1563, 0.3565, 0.2421, 0.1991}, {0.0430, 0.1526, 0.3642, 0.2458, 0.1944},
    {0.0747, 0.2026, 0.4511, 0.1946, 0.0770}, {0.0455, 0.1549, 0.3601, 0.2405, 0.1990},
    {0.1157, 0.2968, 0.4321, 0.1174, 0.0380}, {0.0519, 0.1727, 0.3930, 0.2150, 0.1674},
    {0.0490, 0.1702, 0.3540, 0.2329, 0.1939}, {0.0545, 0.1615, 0.3782, 0.2227, 0.1831},
    {0.0666, 0.1914, 0.3871, 0.1997, 0.1552}, {0.0597, 0.1694, 0.4179, 0.2343, 0.1187},
    {0.064000, 0.163000, 0.448000, 0.181000, 0.144000} };
/* contact_matrix :838-848 */
static const double CM[5][5] = {
    {0.2287, 0.1839, 0.4219, 0.1116, 0.0539}, {0.0276, 0.5964, 0.2878, 0.0591, 0.0291},
    {0.0376, 0.1454, 0.6253, 0.1423, 0.0494}, {0.0242, 0.1094, 0.4867, 0.2723, 0.1074},
    {0.0207, 0.1083, 0.4071, 0.2193, 0.2446} };

void orc_gpw(int ag, int cnts, int out[5]) {            /* :804  Int.(round.(cm[x.ag]*cnts)), half-to-even */
    for (int g = 0; g < 5; ++g) out[g] = (int)nearbyint(CM[ag - 1][g] * (double)cnts);
}

/* ------------------------------------------------------------------ reset_params / init_betas  :146-215 */
static void td_seasonality(int T, double *out) {        /* :204-215 */
    double mn = 0, mx = 0;
    for (int t = 1; t <= T; ++t) {
        double v = 6.261 + -11.81 * cos((80 - t) * 0.022) + 1.817 * sin((80 - t) * 0.022);
        out[t - 1] = v;
        if (t == 1 || v < mn) mn = v;
        if (t == 1 || v > mx) mx = v;
    }
    for (int t = 0; t < T; ++t) out[t] = (out[t] - 2.5 * mn) / (mx - mn);
}

orc_sim *orc_create(const orc_params *ip, uint64_t seed) {
    if (ip->popsize == 0) { fail("no population size given"); return NULL; }                       /* :164 */
    if (!(ip->ctstrat == 0 || ip->ctstrat == 1 || ip->ctstrat == 3)) { fail("invalid ct strategy"); return NULL; } /* :165 */
    if (ip->prov_id < 0 || ip->prov_id > 14) { fail("shame for not knowing your canadian provinces and territories"); return NULL; } /* :334 */
    orc_sim *s = (orc_sim *)calloc(1, sizeof *s);
    s->p = *ip;                                          /* :151-153 */
    s->key[0] = (uint32_t)seed; s->key[1] = (uint32_t)(seed >> 32);
    s->totalisolated = 0;                                /* :156-158 */
    s->betas = (double *)malloc(sizeof(double) * (size_t)(ip->modeltime > 0 ? ip->modeltime : 1));
    if (ip->seasonal) {                                  /* :193-197 */
        td_seasonality(ip->modeltime, s->betas);
        for (int t = 0; t < ip->modeltime; ++t) s->betas[t] = ip->beta * s->betas[t];
    } else {
        for (int t = 0; t < ip->modeltime; ++t) s->betas[t] = ip->beta * 1.0;
    }
    s->humans = (orc_human *)calloc((size_t)ip->popsize, sizeof(orc_human));   /* :167 */
    s->day = 1;
    return s;
}

void orc_destroy(orc_sim *s) {
    if (!s) return;
    for (int i = 0; i < s->p.popsize; ++i) free(s->humans[i].tracecontacts);
    for (int g = 0; g < 5; ++g) free(s->grp[g]);
    free(s->humans); free(s->betas); free(s);
}
orc_params *orc_params_ptr(orc_sim *s) { return &s->p; }
orc_human *orc_humans(orc_sim *s) { return s->humans; }
const double *orc_betas(orc_sim *s) { return s->betas; }
int64_t orc_totalisolated(orc_sim *s) { return s->totalisolated; }
int orc_day(orc_sim *s) { return s->day; }
void orc_set_day(orc_sim *s, int day) { s->day = day; }

/* ------------------------------------------------------------------ get_nextday_counts :772-788 */
static int nextday_counts(orc_sim *s, orc_human *x, int day_of_use) {
    uint32_t w4[4];
    uint32_t i = (uint32_t)x->idx - 1u;
    draw4(s, i >> 2, (uint32_t)day_of_use, S_COUNT, 0, w4);
    uint32_t w = w4[i & 3u];
    int cnt;
    if (x->iso) {                                        /* :778-779  rand()<0.5 ? 0 : rand(1:3), one word */
        cnt = (w < 0x80000000u) ? 0 : 1 + (int)((((uint64_t)(w & 0x7fffffffu)) * 3u) >> 31);
    } else {                                             /* :781 rand(nbs[ag]) */
        cnt = table_draw(w, 0, nb_thr[x->ag - 1], nb_len[x->ag - 1]);
    }
    if (x->health == ORC_DED) cnt = 0;                   /* :783-785 */
    x->nextday_meetcnt = cnt;
    return cnt;
}
int orc_get_nextday_counts(orc_sim *s, int i) { return nextday_counts(s, &s->humans[i], s->day); }

/* ------------------------------------------------------------------ initialize :171-188 */
int orc_initialize(orc_sim *s) {
    const double *agedist = PROV_AG[s->p.prov_id];
    for (int g = 0; g < 5; ++g) { free(s->grp[g]); s->grp[g] = NULL; }
    s->grps_built = 0;
    s->day = 1;
    for (int i = 0; i < s->p.popsize; ++i) {
        orc_human *x = &s->humans[i];
        free(x->tracecontacts);
        memset(x, 0, sizeof *x);                         /* Human() defaults :8-27 */
        x->swap = ORC_UNDEF; x->sickfrom = ORC_UNDEF; x->doi = 999; x->tracestart = -1; x->traceend = -1;
        x->idx = i + 1;
        uint32_t a[4], b[4];
        draw4(s, (uint32_t)i, 0, S_INIT, 0, a);
        draw4(s, (uint32_t)i, 0, S_INIT, 1, b);
        /* :177 rand(Categorical): walk the CDF, ag = 1 + #{k<4 : u >= cdf_k} */
        double u = u01(a[0]), cp = agedist[0];
        int ag = 1;
        while (ag < 5 && cp <= u) { cp += agedist[ag]; ++ag; }
        x->ag = ag;
        /* :178 age draw is never read — no key is spent on it */
        x->exp = 999;                                    /* :179 */
        /* :180 sample_epi_durations :345-358 */
        int latents = table_draw(a[1], ORC_LAT_BASE, orc_lat_thr, ORC_LAT_LEN);
        int pres = table_draw(a[2], ORC_PRE_BASE, orc_pre_thr, ORC_PRE_LEN);
        int asymps = table_draw(a[3], ORC_ASY_BASE, orc_asy_thr, ORC_ASY_LEN);
        int infs = table_draw(b[0], ORC_INF_BASE, orc_inf_thr, ORC_INF_LEN);
        latents = latents - pres;                        /* :354 */
        x->dur[0] = latents; x->dur[1] = asymps; x->dur[2] = pres; x->dur[3] = infs;
        /* :181 rand() < quar_grp_frac && ag == quar_grp */
        if (u01(b[1]) < s->p.quar_grp_frac && x->ag == s->p.quar_grp) { x->iso = 1; x->isovia = ORC_VIA_QU; }
        nextday_counts(s, x, 1);                         /* :186 */
    }
    return 0;
}

/* get_ag_dist :339-343 */
static void build_groups(orc_sim *s) {
    int n = s->p.popsize;
    for (int g = 0; g < 5; ++g) { free(s->grp[g]); s->grp[g] = (int *)malloc(sizeof(int) * (size_t)n); s->grplen[g] = 0; }
    for (int i = 0; i < n; ++i) { int g = s->humans[i].ag - 1; s->grp[g][s->grplen[g]++] = i; }
    s->grps_built = 1;
}

/* ------------------------------------------------------------------ state machine :431-644 */
static inline void set_isolation(orc_human *x, int iso, int via) {     /* :432-444 */
    x->iso = iso;
    if (x->isovia == ORC_VIA_NULL) x->isovia = via;
}
static inline void set_isolation_keep(orc_human *x, int iso) { set_isolation(x, iso, x->isovia); } /* :431 */

static void move_to_latent(orc_sim *s, orc_human *x, int day) {        /* :446-461 */
    uint32_t w[4]; draw4(s, (uint32_t)x->idx - 1u, (uint32_t)day, S_TRANS, 0, w);
    x->health = ORC_LAT; x->doi = 0; x->tis = 0; x->exp = x->dur[0];
    x->swap = (u01(w[0]) < s->p.asymp_props[x->ag - 1]) ? ORC_ASYMP : ORC_PRE;
    if (s->p.calibration) { x->swap = ORC_LAT; x->exp = 999; }
}
static void move_to_asymp(orc_human *x) {                              /* :464-472 */
    x->health = ORC_ASYMP; x->tis = 0; x->exp = x->dur[1]; x->swap = ORC_REC;
}
static void move_to_pre(orc_sim *s, orc_human *x, int day) {           /* :475-488 */
    uint32_t w[4]; draw4(s, (uint32_t)x->idx - 1u, (uint32_t)day, S_TRANS, 0, w);
    x->health = ORC_PRE; x->tis = 0; x->exp = x->dur[2];
    x->swap = (u01(w[0]) < s->p.mild_props[x->ag - 1]) ? ORC_MILD : ORC_INF;
    if (u01(w[1]) < s->p.fpreiso) set_isolation(x, 1, ORC_VIA_PI);
}
static void move_to_mild(orc_sim *s, orc_human *x, int day) {          /* :491-507 */
    x->health = ORC_MILD; x->tis = 0; x->exp = x->dur[3]; x->swap = ORC_REC;
    int go = x->iso;
    if (!go) { uint32_t w[4]; draw4(s, (uint32_t)x->idx - 1u, (uint32_t)day, S_TRANS, 0, w); go = u01(w[0]) < s->p.fmild; }
    if (go) { x->swap = ORC_MISO; x->exp = s->p.tau_mild; }
}
static void move_to_miso(orc_sim *s, orc_human *x) {                   /* :510-517 */
    x->health = ORC_MISO; x->swap = ORC_REC; x->tis = 0; x->exp = x->dur[3] - s->p.tau_mild;
    set_isolation(x, 1, ORC_VIA_MI);
}
static void move_to_infsimple(orc_human *x) {                          /* :520-528 */
    x->health = ORC_INF; x->tis = 0; x->exp = x->dur[3]; x->swap = ORC_REC;
    set_isolation(x, 0, ORC_VIA_NULL);
}
static int move_to_inf(orc_sim *s, orc_human *x, int day) {            /* :530-569 */
    static const double hlo[5] = {0.02, 0.02, 0.28, 0.28, 0.60}, hhi[5] = {0.03, 0.03, 0.34, 0.34, 0.68};
    static const double clo[5] = {0.01, 0.01, 0.03, 0.05, 0.05}, chi[5] = {0.015, 0.015, 0.05, 0.20, 0.15};
    const double mh[5] = {0.01 / 5, 0.01 / 5, 0.0135 / 3, 0.01225 / 1.5, 0.04 / 2};
    uint32_t a[4], b[4];
    draw4(s, (uint32_t)x->idx - 1u, (uint32_t)day, S_TRANS, 0, a);
    draw4(s, (uint32_t)x->idx - 1u, (uint32_t)day, S_TRANS, 1, b);
    int g = x->ag - 1;
    /* :535-536 only the agent's own age-group entries of h and c are ever read */
    double h = hlo[g] + (hhi[g] - hlo[g]) * u01(a[0]);
    double c = clo[g] + (chi[g] - clo[g]) * u01(a[1]);
    double m = mh[g];
    if (s->p.calibration) { h = 0; c = 0; m = 0; }      /* :540-544 */
    int time_to_hospital = (int)nearbyint(2.0 + 3.0 * u01(a[2]));      /* :546 */
    x->health = ORC_INF; x->swap = ORC_UNDEF; x->tis = 0;
    if (u01(a[3]) < h) {                                 /* :551 */
        x->exp = time_to_hospital;
        x->swap = (u01(b[0]) < c) ? ORC_ICU : ORC_HOS;   /* :553 */
    } else if (u01(b[1]) < m) {                          /* :555 */
        x->exp = x->dur[3]; x->swap = ORC_DED;
    } else {
        x->exp = x->dur[3]; x->swap = ORC_REC;
        if (x->iso || u01(b[2]) < s->p.fsevere) { x->exp = 1; x->swap = ORC_IISO; }   /* :561-564 */
    }
    if (x->swap == ORC_UNDEF) return fail("agent I -> ?");
    return 0;
}
static void move_to_iiso(orc_human *x) {                               /* :571-578 */
    x->health = ORC_IISO; x->swap = ORC_REC; x->tis = 0; x->exp = x->dur[3] - 1;
    set_isolation(x, 1, ORC_VIA_MI);
}
static int move_to_hospicu(orc_sim *s, orc_human *x, int day) {        /* :580-622 */
    const double mh[5] = {0.01 * 2, 0.01 * 2, 0.0135, 0.01225, 0.03 * 2};
    uint32_t w[4]; draw4(s, (uint32_t)x->idx - 1u, (uint32_t)day, S_TRANS, 0, w);
    int psi = table_draw(w[0], ORC_PSI_BASE, orc_psi_thr, ORC_PSI_LEN);   /* :590-591 (H and C share the law) */
    int mu = table_draw(w[1], ORC_MU_BASE, orc_mu_thr, ORC_MU_LEN);       /* :592-593 */
    int swaphealth = x->swap;
    x->health = swaphealth; x->swap = ORC_UNDEF; x->tis = 0;
    set_isolation_keep(x, 1);                            /* :599 */
    double u = u01(w[2]);
    if (swaphealth == ORC_HOS) {
        if (u < mh[x->ag - 1]) { x->exp = mu; x->swap = ORC_DED; } else { x->exp = psi; x->swap = ORC_REC; }
    }
    if (swaphealth == ORC_ICU) {
        if (u < 2 * mh[x->ag - 1]) { x->exp = mu + 2; x->swap = ORC_DED; } else { x->exp = psi + 2; x->swap = ORC_REC; }
    }
    if (x->swap == ORC_UNDEF) return fail("agent H -> ?");
    return 0;
}
static void move_to_dead(orc_human *x) {                               /* :624-633 */
    x->health = ORC_DED; x->swap = ORC_UNDEF; x->tis = 0; x->exp = 999; x->iso = 1;
    set_isolation_keep(x, 1);
}
static void move_to_recovered(orc_human *x) {                          /* :635-644 */
    x->health = ORC_REC; x->swap = ORC_UNDEF; x->tis = 0; x->exp = 999;
}

/* ------------------------------------------------------------------ insert_infected :360-395 */
int orc_insert_infected(orc_sim *s, int health, int num, int ag, int call_id, int32_t *chosen) {
    int n = s->p.popsize, L = 0;
    int *l = (int *)malloc(sizeof(int) * (size_t)n);
    for (int i = 0; i < n; ++i) {
        orc_human *x = &s->humans[i];
        if (x->health == ORC_SUS && (health != ORC_INF || x->ag == ag)) l[L++] = i;      /* :365-369 */
    }
    if (L == 0 || num > L) { free(l); return fail("No susceptibles in the population to change status"); }  /* :371-372 */
    int rc = 0, day = s->day - 1;
    /* :374 sample(l, num; replace=false): pick j takes the floor(u_j*(L-j))-th remaining susceptible */
    for (int j = 0; j < num && !rc; ++j) {
        uint32_t w[4]; draw4(s, (uint32_t)j, 0, S_SEED, (uint32_t)call_id, w);
        int k = below(w[0], L - j);
        int i = l[k];
        memmove(&l[k], &l[k + 1], sizeof(int) * (size_t)(L - j - 1 - k));
        orc_human *x = &s->humans[i];
        if (health == ORC_PRE) move_to_pre(s, x, day);
        else if (health == ORC_LAT) { move_to_latent(s, x, day); x->tis = x->exp; }
        else if (health == ORC_INF) move_to_infsimple(x);
        else if (health == ORC_REC) move_to_recovered(x);
        else rc = fail("can not insert human of this health");
        x->sickfrom = ORC_INF;                           /* :391 */
        if (chosen) chosen[j] = i + 1;
    }
    free(l);
    return rc;
}

/* ------------------------------------------------------------------ contact tracing :647-752 */
static void apply_ct_strategy(orc_sim *s, orc_human *y) {              /* :647-656 */
    set_isolation(y, 1, ORC_VIA_CT);
    s->totalisolated += 1;
    if (s->p.ctstrat == 1) y->tracedxp = 14;
    if (s->p.ctstrat == 3) y->tracedxp = s->p.strat3qdays;
}
static inline size_t tc_bytes(const orc_sim *s) { return ((size_t)s->p.popsize + 7) / 8; }
static void tc_clear(orc_sim *s, orc_human *x) { if (x->tracecontacts) memset(x->tracecontacts, 0, tc_bytes(s)); }
static void tc_set(orc_sim *s, orc_human *x, int j) {
    if (!x->tracecontacts) x->tracecontacts = (uint8_t *)calloc(tc_bytes(s), 1);
    x->tracecontacts[j >> 3] |= (uint8_t)(1u << (j & 7));
}

static int ct_dynamics(orc_sim *s, orc_human *x, int day) {            /* :658-752 */
    if (s->p.ctstrat == 0) return 0;                     /* :668 */
    int xh = x->health, xs = x->swap, doi = x->doi;
    uint32_t xi = (uint32_t)x->idx - 1u;
    if (xh == ORC_LAT && xs != ORC_ASYMP && doi == 0) {  /* :678 (&& chain: the draw only happens here) */
        uint32_t w[4]; draw4(s, xi, (uint32_t)day, S_CTWIN, 0, w);
        if (u01(w[0]) < (double)s->p.fctcapture) {
            int time_to_id = 1 + below(w[1], x->dur[3]); /* :680 rand(1:dur[4]) */
            int delta = x->dur[0] + x->dur[2] + time_to_id;
            int q = delta - s->p.cdaysback;
            x->tracestart = q > 0 ? q : 0;
            x->traceend = delta;
            x->tracing = 0;
            tc_clear(s, x);
            if (x->tracestart > x->traceend) return fail("tracestart < traceend");
        }
    }
    if (doi >= x->tracestart && doi < x->traceend) x->tracing = 1;     /* :699-701 */
    if (doi == x->traceend) {                            /* :705-716 */
        x->tracing = 0;
        apply_ct_strategy(s, x);
        if (x->tracecontacts) {
            for (int j = 0; j < s->p.popsize; ++j) {     /* findall(...) ascending */
                if (!(x->tracecontacts[j >> 3] & (1u << (j & 7)))) continue;
                uint32_t w[4]; draw4(s, xi, (uint32_t)day, S_CTCON, (uint32_t)j, w);
                if (u01(w[0]) < (double)s->p.fcontactst) apply_ct_strategy(s, &s->humans[j]);
            }
        }
        tc_clear(s, x);
    }
    if (x->tracedxp > 0) {                               /* :718-750 */
        if (!x->iso) return fail("a traced person should always be isolated until the trace is killed");
        x->tracedxp -= 1;
        if (x->tracedxp == 0) {
            set_isolation_keep(x, 0);
            x->isovia = ORC_VIA_NULL;
            if (s->p.ctstrat == 3) {
                uint32_t w[4]; draw4(s, xi, (uint32_t)day, S_CT3, 0, w);
                double rn = u01(w[0]);
                int h = x->health;
                if ((h == ORC_LAT && rn < 0.05) || (h == ORC_PRE && rn < 0.95) || (h == ORC_ASYMP && rn < 0.70) ||
                    h == ORC_MILD || h == ORC_INF || h == ORC_MISO || h == ORC_IISO) {
                    set_isolation(x, 1, ORC_VIA_CT);
                    x->tracedxp = 14;
                }
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------ time_update :399-428 */
int orc_time_update(orc_sim *s) {
    int day = s->day;
    for (int i = 0; i < s->p.popsize; ++i) {
        orc_human *x = &s->humans[i];
        x->tis += 1; x->doi += 1;
        if (x->tis >= x->exp) {
            int rc = 0;
            switch (x->swap) {
                case ORC_LAT: move_to_latent(s, x, day); break;
                case ORC_PRE: move_to_pre(s, x, day); break;
                case ORC_ASYMP: move_to_asymp(x); break;
                case ORC_MILD: move_to_mild(s, x, day); break;
                case ORC_MISO: move_to_miso(s, x); break;
                case ORC_INF: rc = move_to_inf(s, x, day); break;
                case ORC_IISO: move_to_iiso(x); break;
                case ORC_HOS: case ORC_ICU: rc = move_to_hospicu(s, x, day); break;
                case ORC_REC: move_to_recovered(x); break;
                case ORC_DED: move_to_dead(x); break;
                default: rc = fail("swap expired, but no swap set.");   /* :418 */
            }
            if (rc) return rc;
        }
        if (ct_dynamics(s, x, day)) return 1;            /* :422 */
        nextday_counts(s, x, day + 1);                   /* :425 */
    }
    s->day = day + 1;
    return 0;
}

/* ------------------------------------------------------------------ dyntrans :756-834 */
double orc_get_betavalue(orc_sim *s, int sys_time, int xhealth) {      /* :756-769 */
    if (s->p.modeltime == 0) return 0.0;
    double bf = s->betas[sys_time - 1];
    if (xhealth == ORC_ASYMP) bf = bf * s->p.frelasymp;
    else if (xhealth == ORC_MILD || xhealth == ORC_MISO) bf = bf * 0.44;
    else if (xhealth == ORC_INF || xhealth == ORC_IISO) bf = bf * 0.89;
    return bf;
}

int orc_dyntrans(orc_sim *s, int *totalmet_out, int *totalinf_out) {
    int totalmet = 0, totalinf = 0, day = s->day, n = s->p.popsize;
    if (!s->grps_built) build_groups(s);
    /* :794 the infectious set is fixed at the start of the day (health is not changed by dyntrans) */
    for (int xid = 0; xid < n; ++xid) {
        orc_human *x = &s->humans[xid];
        int xhealth = x->health;
        if (!(xhealth == ORC_PRE || xhealth == ORC_ASYMP || xhealth == ORC_MILD || xhealth == ORC_MISO ||
              xhealth == ORC_INF || xhealth == ORC_IISO)) continue;
        double beta = orc_get_betavalue(s, day, xhealth);               /* :800 */
        int cnts = x->nextday_meetcnt;                                   /* :802 */
        if (cnts == 0) continue;
        int gpw[5]; orc_gpw(x->ag, cnts, gpw);                           /* :804 */
        for (int g = 0; g < 5; ++g) {
            if (gpw[g] > 0 && s->grplen[g] == 0) return fail("empty age group sampled");
            for (int k = 0; k < gpw[g]; ++k) {
                uint32_t w[4]; draw4(s, (uint32_t)xid, (uint32_t)day, S_DT, ((uint32_t)g << 16) | (uint32_t)k, w);
                int j = s->grp[g][below(w[0], s->grplen[g])];            /* :807 rand(grps[i], g) */
                orc_human *y = &s->humans[j];
                if (y->nextday_meetcnt == 0) continue;                   /* :811-812 */
                y->nextday_meetcnt -= 1;                                 /* :813 */
                totalmet += 1;
                if (x->tracing) tc_set(s, x, j);                         /* :817-819 */
                if (y->health == ORC_SUS && y->swap == ORC_UNDEF) {      /* :822 */
                    if (u01(w[1]) < beta) {                              /* :823 */
                        totalinf += 1;
                        y->swap = ORC_LAT; y->exp = y->tis; y->sickfrom = xhealth;
                    }
                }
            }
        }
    }
    if (totalmet_out) *totalmet_out = totalmet;
    if (totalinf_out) *totalinf_out = totalinf;
    return 0;
}

/* ------------------------------------------------------------------ main :104-142 */
int orc_main(orc_sim *s, int16_t *hmatrix) {
    int n = s->p.popsize, T = s->p.modeltime;
    if (hmatrix) memset(hmatrix, 0, sizeof(int16_t) * (size_t)n * (size_t)T);   /* :111 */
    if (orc_initialize(s)) return 1;                                     /* :112 */
    if (orc_insert_infected(s, ORC_PRE, s->p.initialinf, 4, 0, NULL)) return 1;          /* :117/:119 */
    if (!s->p.calibration) {
        if (orc_insert_infected(s, ORC_REC, s->p.initialhi, 4, 1, NULL)) return 1;       /* :120 */
    }
    double fpreiso = s->p.fpreiso;                                       /* :124-125 */
    s->p.fpreiso = 0;
    build_groups(s);                                                     /* :128 */
    for (int st = 1; st <= T; ++st) {                                    /* :131 */
        if (st == s->p.tpreiso) s->p.fpreiso = fpreiso;                  /* :133-135 */
        if (hmatrix) {                                                   /* :136 _get_model_state :218-223 */
            int16_t *col = hmatrix + (size_t)(st - 1) * (size_t)n;
            for (int i = 0; i < n; ++i) col[i] = (int16_t)s->humans[i].health;
        }
        s->day = st;
        if (orc_dyntrans(s, NULL, NULL)) return 1;                       /* :137 */
        if (orc_time_update(s)) return 1;                                /* :138 */
    }
    return 0;
}

/* ------------------------------------------------------------------ reductions :259-313 */
void orc_get_column_incidence(const int16_t *hm, int n, int t, int hcol, int64_t *timevec) {   /* :270-280 */
    for (int d = 0; d < t; ++d) timevec[d] = 0;
    for (int i = 0; i < n; ++i) {                        /* eachrow */
        for (int d = 0; d < t; ++d) {                    /* findfirst */
            if (hm[(size_t)d * (size_t)n + (size_t)i] == hcol) { timevec[d] += 1; break; }
        }
    }
}
void orc_get_column_prevalence(const int16_t *hm, int n, int t, int hcol, int64_t *timevec) {  /* :282-293 */
    for (int d = 0; d < t; ++d) {
        int64_t c = 0;
        const int16_t *col = hm + (size_t)d * (size_t)n;
        for (int i = 0; i < n; ++i) c += (col[i] == hcol);
        timevec[d] = c;
    }
}
int orc_count_infectors(orc_sim *s, int64_t out[4]) {                  /* :295-313 */
    out[0] = out[1] = out[2] = out[3] = 0;
    for (int i = 0; i < s->p.popsize; ++i) {
        orc_human *x = &s->humans[i];
        if (x->health == ORC_SUS) continue;
        int f = x->sickfrom;
        if (f == ORC_PRE) out[0]++;
        else if (f == ORC_ASYMP) out[1]++;
        else if (f == ORC_MILD || f == ORC_MISO) out[2]++;
        else if (f == ORC_INF || f == ORC_IISO) out[3]++;
        else return fail("sickfrom not set right");
    }
    return 0;
}

/* runsim :90-97: _collectdf on the full matrix and on the five _splitstate row subsets.
 * One pass per row gives the same numbers as the 12 findfirst/findall scans. */
void orc_collect_curves(const int16_t *hm, int n, int t, const int32_t *ags, int32_t *inc, int32_t *prev) {
    size_t gstride = (size_t)t * 12;
    memset(inc, 0, sizeof(int32_t) * 6 * gstride);
    memset(prev, 0, sizeof(int32_t) * 6 * gstride);
    for (int i = 0; i < n; ++i) {
        int g = ags[i];                                  /* 1..5 */
        int seen[13] = {0};
        for (int d = 0; d < t; ++d) {
            int h = hm[(size_t)d * (size_t)n + (size_t)i];
            if (h < 0 || h > 11) continue;
            prev[(size_t)d * 12 + (size_t)h] += 1; prev[(size_t)g * gstride + (size_t)d * 12 + (size_t)h] += 1;
            if (!seen[h]) { seen[h] = 1; inc[(size_t)d * 12 + (size_t)h] += 1; inc[(size_t)g * gstride + (size_t)d * 12 + (size_t)h] += 1; }
        }
    }
}

/* ------------------------------------------------------------------ ensemble (pmap over runsim) */
typedef struct {
    const orc_params *psets; const int32_t *pset_of_rep; int nrep; const uint64_t *seeds;
    int32_t *inc, *prev; int64_t *infectors, *totalisolated; int16_t *hmatrix;
    int next; int failed; pthread_mutex_t mu;
} ens_job;

static void *ens_worker(void *arg) {
    ens_job *J = (ens_job *)arg;
    int16_t *hm_local = NULL; int32_t *ags = NULL; size_t hm_cap = 0, ag_cap = 0;
    for (;;) {
        pthread_mutex_lock(&J->mu);
        int r = J->next < J->nrep ? J->next++ : -1;
        pthread_mutex_unlock(&J->mu);
        if (r < 0) break;
        const orc_params *p = &J->psets[J->pset_of_rep ? J->pset_of_rep[r] : 0];
        size_t n = (size_t)p->popsize, T = (size_t)p->modeltime;
        int16_t *hm = J->hmatrix ? J->hmatrix + (size_t)r * n * T : NULL;
        if (!hm) {
            if (hm_cap < n * T) { free(hm_local); hm_local = (int16_t *)malloc(sizeof(int16_t) * n * T); hm_cap = n * T; }
            hm = hm_local;
        }
        if (ag_cap < n) { free(ags); ags = (int32_t *)malloc(sizeof(int32_t) * n); ag_cap = n; }
        orc_sim *s = orc_create(p, J->seeds[r]);
        int bad = (s == NULL) || orc_main(s, hm);
        if (!bad) {
            for (size_t i = 0; i < n; ++i) ags[i] = s->humans[i].ag;
            if (J->inc && J->prev) orc_collect_curves(hm, (int)n, (int)T, ags, J->inc + (size_t)r * 6 * T * 12, J->prev + (size_t)r * 6 * T * 12);
            if (J->infectors) bad |= orc_count_infectors(s, J->infectors + (size_t)r * 4);
            if (J->totalisolated) J->totalisolated[r] = s->totalisolated;
        }
        if (bad) { pthread_mutex_lock(&J->mu); J->failed++; pthread_mutex_unlock(&J->mu); }
        orc_destroy(s);
    }
    free(hm_local); free(ags);
    return NULL;
}

int orc_run_ensemble(const orc_params *psets, int n_psets, const int32_t *pset_of_rep, int nrep,
                     const uint64_t *seeds, int nthreads,
                     int32_t *inc, int32_t *prev, int64_t *infectors, int64_t *totalisolated, int16_t *hmatrix) {
    (void)n_psets;
    ens_job J = { psets, pset_of_rep, nrep, seeds, inc, prev, infectors, totalisolated, hmatrix, 0, 0, PTHREAD_MUTEX_INITIALIZER };
    if (nthreads < 1) nthreads = 1;
    if (nthreads > nrep) nthreads = nrep > 0 ? nrep : 1;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nthreads);
    for (int i = 0; i < nthreads; ++i) pthread_create(&th[i], NULL, ens_worker, &J);
    for (int i = 0; i < nthreads; ++i) pthread_join(th[i], NULL);
    free(th);
    return J.failed;
}
