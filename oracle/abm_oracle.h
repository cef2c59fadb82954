/* abm_oracle.h — CPU restatement of covid19abm.jl's simulation path.
 *
 * TEST INFRASTRUCTURE ONLY. Nothing under covid19abm.jl_b200/ may include, link or call this.
 * Allowed users: tests/, __graft_entry__.smoke(), bench.py's cpu_baseline / --impl reference legs.
 *
 * PARITY STATUS: "parity unpinned" with respect to a real Julia run — Julia is not installed in
 * the build image, the reference's dependencies are unpinned (Project.toml:6-14, no [compat]) and its
 * tests hold no fixed-seed golden vectors (test/runtests.jl is invariants only). This file follows
 * /root/reference/src/covid19abm.jl function by function (file:line cited at each function) with
 * the reference's sequential order and array-of-structs agents, and replaces "position in Julia's
 * global RNG stream" by the keyed Philox4x32-10 draw contract written down in DESIGN.md §3.
 * It is pinned by: Random123's published Philox known-answer vectors, the analytic moments of every
 * variate table (tests/test_tables.py), the invariants of test/runtests.jl re-expressed in
 * tests/test_invariants.py, and the one trajectory-level number the reference holds: the author's calibration
 * fit R0 = 46.1761*beta - 0.00285 (scripts/local_run.jl:164-175), reproduced by this oracle on the same design
 * (tests/test_calibration_pin.py; tools/calibration_pin.py). Parity with an actual Julia run remains unpinned
 * (no julia binary in the image, no seeds or goldens in the reference).
 */
#pragma once
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* HEALTH enum, src/covid19abm.jl:5 (Julia @enum is 0-based) */
enum { ORC_SUS = 0, ORC_LAT, ORC_PRE, ORC_ASYMP, ORC_MILD, ORC_MISO, ORC_INF, ORC_IISO,
       ORC_HOS, ORC_ICU, ORC_REC, ORC_DED, ORC_UNDEF };
/* isovia symbols, src/covid19abm.jl:21 */
enum { ORC_VIA_NULL = 0, ORC_VIA_QU, ORC_VIA_PI, ORC_VIA_MI, ORC_VIA_CT };

/* ModelParameters, src/covid19abm.jl:31-55 (field for field; prov as row index of :319-333) */
typedef struct orc_params {
    double beta;            /* β */
    double frelasymp;
    int32_t seasonal;
    int32_t popsize;
    int32_t prov_id;        /* 0 alberta .. 14 newyork, order of src/covid19abm.jl:319-333 */
    int32_t calibration;
    int32_t modeltime;
    int32_t initialinf;
    int32_t initialhi;
    int32_t tau_mild;       /* τmild */
    double asymp_props[5];
    double mild_props[5];
    double fmild;
    double fsevere;
    double fpreiso;
    double quar_grp_frac;
    int32_t tpreiso;
    int32_t quar_grp;
    int32_t ctstrat;
    int32_t cdaysback;
    int32_t strat3qdays;
    float fctcapture;       /* Float16 in Julia: caller passes the binary16-rounded value */
    float fcontactst;
    int32_t _pad;
} orc_params;

/* Human, src/covid19abm.jl:8-27 (age is never read by the model and is not kept) */
typedef struct orc_human {
    int32_t idx;
    int32_t health, swap, sickfrom;
    int32_t nextday_meetcnt;
    int32_t ag;             /* 1..5 */
    int32_t tis, exp;
    int32_t dur[4];         /* latents, asymps, pres, infs */
    int32_t doi;
    int32_t iso, isovia;
    int32_t tracing, tracestart, traceend, tracedxp;
    uint8_t *tracecontacts; /* bitset of popsize bits, allocated on first use (ref: Bool[10000], :25) */
} orc_human;

typedef struct orc_sim orc_sim;

const char *orc_last_error(void);

/* reset_params :146-168 — returns NULL on the reference's error() conditions */
orc_sim *orc_create(const orc_params *p, uint64_t seed);
void orc_destroy(orc_sim *s);
orc_params *orc_params_ptr(orc_sim *s);      /* the global `p` (mutable, e.g. fpreiso) */
orc_human *orc_humans(orc_sim *s);           /* the global `humans` */
const double *orc_betas(orc_sim *s);         /* the global BETAS, length modeltime */
int64_t orc_totalisolated(orc_sim *s);       /* ct_data.totalisolated :61,649 */
int orc_day(orc_sim *s);                     /* day index st the next dyntrans/time_update will use */
void orc_set_day(orc_sim *s, int day);

int orc_initialize(orc_sim *s);                                          /* :171-188 */
int orc_insert_infected(orc_sim *s, int health, int num, int ag, int call_id, int32_t *chosen); /* :360-395 */
int orc_dyntrans(orc_sim *s, int *totalmet, int *totalinf);              /* :790-834 */
int orc_time_update(orc_sim *s);                                         /* :399-428 */
int orc_get_nextday_counts(orc_sim *s, int i);                           /* :772-788 for humans[i], keyed on orc_day() */
double orc_get_betavalue(orc_sim *s, int sys_time, int xhealth);         /* :756-769 */
int orc_main(orc_sim *s, int16_t *hmatrix);                              /* :104-142 (after orc_create) */

/* result reductions :259-313 — literal restatements over a column-major popsize x modeltime matrix */
void orc_get_column_incidence(const int16_t *hm, int n, int t, int hcol, int64_t *timevec);  /* :270-280 */
void orc_get_column_prevalence(const int16_t *hm, int n, int t, int hcol, int64_t *timevec); /* :282-293 */
int orc_count_infectors(orc_sim *s, int64_t out[4]);                                         /* :295-313 */
/* curves for "all" + 5 age groups as runsim builds them (:90-97): inc/prev are [6][t][12] int32 */
void orc_collect_curves(const int16_t *hm, int n, int t, const int32_t *ags, int32_t *inc, int32_t *prev);

/* ensemble driver (the reference's pmap over runsim, scripts/local_run.jl:28-30), nthreads workers.
 * Any output pointer may be NULL. hmatrix (if given) is [nrep][t][n]. Returns 0 or the number of failed replicates. */
int orc_run_ensemble(const orc_params *psets, int n_psets, const int32_t *pset_of_rep, int nrep,
                     const uint64_t *seeds, int nthreads,
                     int32_t *inc, int32_t *prev, int64_t *infectors, int64_t *totalisolated,
                     int16_t *hmatrix);

/* draw-contract primitives, exposed for known-answer tests */
void orc_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]);
/* replay tape: same layout and meaning as cabm_tape_entry / cabm_set_tape (include/cabm.h). Process-wide; NULL or n = 0 clears it. */
typedef struct orc_tape_entry { uint64_t seed; uint32_t ctr[4]; uint32_t w[4]; } orc_tape_entry;
int orc_set_tape(const orc_tape_entry *entries, uint64_t n);
int orc_table(const char *name, int *base, const uint32_t **thr);       /* returns length or -1 */
void orc_gpw(int ag, int cnts, int out[5]);                              /* :804 split, ag 1..5 */

#ifdef __cplusplus
}
#endif
