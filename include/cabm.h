/* cabm.h — C ABI of libcabm_b200.so: covid19abm.jl's simulation path on B200 (sm_100a).
 *
 * The reference has no FFI: its boundary is the Julia API of /root/reference/src/covid19abm.jl.
 * Each entry point below names the Julia interface it replaces (file:line under /root/reference).
 * Plain C types only; no exceptions cross the ABI: every call returns 0 or a CABM_E_* code and
 * cabm_last_error() gives the message (the reference's error("...") strings where one exists).
 * One handle = one GPU = one stream; a handle is not thread-safe, distinct handles are independent (no unsynchronised
 * process-wide state: the creation error string is per calling thread, kernel attributes are set under a lock).
 * There is no CPU fallback: without a CUDA device cabm_create fails with CABM_E_CUDA.
 */
#ifndef CABM_H
#define CABM_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* HEALTH enum — src/covid19abm.jl:5 (0-based like Julia's @enum; these are the hmatrix codes :221) */
enum { CABM_SUS = 0, CABM_LAT, CABM_PRE, CABM_ASYMP, CABM_MILD, CABM_MISO, CABM_INF, CABM_IISO,
       CABM_HOS, CABM_ICU, CABM_REC, CABM_DED, CABM_UNDEF };
/* isovia symbols — src/covid19abm.jl:21 */
enum { CABM_VIA_NULL = 0, CABM_VIA_QU, CABM_VIA_PI, CABM_VIA_MI, CABM_VIA_CT };

/* error codes */
enum { CABM_OK = 0, CABM_E_ARG = 1, CABM_E_CUDA = 2, CABM_E_PARAM = 3, CABM_E_STATE = 4, CABM_E_MODEL = 5 };

/* per-replicate device error word (cabm_get_scalars): one bit per in-model error() of the reference */
enum { CABM_ERR_NO_SWAP = 1,        /* :418 "swap expired, but no swap set." */
       CABM_ERR_NO_SUSCEPTIBLE = 2, /* :371-372 */
       CABM_ERR_TRACED_NOT_ISO = 4, /* :719 */
       CABM_ERR_EMPTY_GROUP = 8,    /* rand(grps[i], g) on an empty group (:807 would throw) */
       CABM_ERR_CT_LOG_FULL = 16,   /* contact log capacity exceeded (sized from cdaysback at cabm_configure; implementation limit) */
       CABM_ERR_SICKFROM = 32,      /* :308 */
       CABM_ERR_INTERNAL = 64 };    /* an implementation limit of a kernel was exceeded; results of that replicate are invalid */

/* ModelParameters — src/covid19abm.jl:31-55, field for field.
 * prov_id indexes the rows of get_province_ag (:319-333): 0 alberta, 1 bc, 2 canada, 3 manitoba,
 * 4 newbruns, 5 newfdland, 6 nwterrito, 7 novasco, 8 nunavut, 9 ontario, 10 pei, 11 quebec,
 * 12 saskat, 13 yukon, 14 newyork. fctcapture/fcontactst are Float16 in Julia (:51-52): pass the
 * binary16-rounded value widened to float. */
typedef struct cabm_params {
    double beta;              /* β :32 */
    double frelasymp;         /* :33 */
    int32_t seasonal;         /* :34 */
    int32_t popsize;          /* :35 */
    int32_t prov_id;          /* :36 */
    int32_t calibration;      /* :37 */
    int32_t modeltime;        /* :38 */
    int32_t initialinf;       /* :39 */
    int32_t initialhi;        /* :40 */
    int32_t tau_mild;         /* τmild :43 */
    double asymp_props[5];    /* :41 */
    double mild_props[5];     /* :42 */
    double fmild;             /* :44 */
    double fsevere;           /* :45 */
    double fpreiso;           /* :47 */
    double quar_grp_frac;     /* :49 */
    int32_t tpreiso;          /* :46 */
    int32_t quar_grp;         /* :48 (1..5) */
    int32_t ctstrat;          /* :50 (0, 1 or 3; anything else is "invalid ct strategy" :165) */
    int32_t cdaysback;        /* :53 */
    int32_t strat3qdays;      /* :54 */
    float fctcapture;         /* :51 */
    float fcontactst;         /* :52 */
    int32_t _pad;
} cabm_params;

typedef struct cabm_handle cabm_handle;

/* cabm_create flags */
enum { CABM_STORE_HMATRIX = 1 };   /* keep the popsize x modeltime state matrix of every replicate in HBM */
/* dyntrans schedule */
enum { CABM_MODE_EXACT = 0,        /* serial-equivalent: bit-identical to the sequential-order C oracle (oracle/, keyed Philox draws);
                                      distribution-equivalent to the Julia reference, whose RNG streams are unpinned */
       CABM_MODE_NATIVE = 1 };     /* per-infector parallel with atomics: ensemble-equivalent to the exact schedule */

/* field ids for cabm_get_field / cabm_set_field — the members of `Human` (src/covid19abm.jl:8-27) */
enum { CABM_F_HEALTH = 0, CABM_F_SWAP, CABM_F_SICKFROM, CABM_F_NEXTDAY_MEETCNT, CABM_F_AG, CABM_F_TIS,
       CABM_F_EXP, CABM_F_DUR /* 4 per agent */, CABM_F_DOI, CABM_F_ISO, CABM_F_ISOVIA, CABM_F_TRACING,
       CABM_F_TRACESTART, CABM_F_TRACEEND, CABM_F_TRACEDXP, CABM_F_NCONTACTS /* |tracecontacts| */ };

/* message of the last failing call on this handle (or of the last failing cabm_create if h == NULL) */
const char *cabm_last_error(const cabm_handle *h);

/* Replaces the module-level state `humans`, `p`, `BETAS`, `ct_data` (:69-73) by device buffers sized for
 * max_replicates independent copies of a popsize-agent population. */
int cabm_create(cabm_handle **out, int device, int max_replicates, int popsize, int modeltime, int flags);
void cabm_destroy(cabm_handle *h);

/* reset_params(ip) :146-168 + init_betas :191-215 for n_rep replicates at once. psets[pset_of_rep[r]] is the
 * ModelParameters of replicate r (pset_of_rep == NULL: all use psets[0]); seeds[r] plays Random.seed! (:78).
 * All psets must share the handle's popsize and modeltime. Errors :164-165, :334 -> CABM_E_PARAM. */
int cabm_configure(cabm_handle *h, const cabm_params *psets, int n_psets, const int32_t *pset_of_rep,
                   int n_rep, const uint64_t *seeds, int mode);

/* main(ip) :104-142 for every configured replicate: initialize, seed, the day loop, then the
 * incidence/prevalence reductions (:259-293) and _count_infectors (:295-313), all on device.
 * Asynchronous on the handle's stream; cabm_sync or any cabm_get_* waits. */
int cabm_run(cabm_handle *h);

/* Replay tape — north_star's "replay mode feeds each kernel the reference's own uniform draws". The reference consumes one
 * sequential RNG (Random.seed!(simnum*726) :78, then rand() at :177-181,352-356,374,455,481,487,503,535-555,561,590-612,
 * 678-680,711,740,779-781,807,823); this library names every such draw by a key instead of a stream position (DESIGN.md §3:
 * Philox counter = (index, day, stream, sub)). A tape lists keys with the four 32-bit words the draw shall return; draws not
 * listed keep their Philox value. So a run of the Julia reference instrumented to log its draws per site can be replayed here
 * draw for draw: uniforms as w = floor(u * 2^32), integerised variates (durations, contact counts) as any w of the table bin
 * of the value (cabm_get_table). INTEGRATION.md gives the key of every draw site. The tape is consulted by the per-day
 * kernels (cabm_run then takes CABM_PATH_MULTIKERNEL; the cabm_step_* calls); entries may be in any order, duplicates are an
 * error; n = 0 removes the tape. The oracle has the same entry point (orc_set_tape) for parity tests. */
typedef struct cabm_tape_entry {
    uint64_t seed;        /* the replicate's seed as given to cabm_configure */
    uint32_t ctr[4];      /* (index, day, stream, sub) */
    uint32_t w[4];        /* the words the draw returns */
} cabm_tape_entry;
int cabm_set_tape(cabm_handle *h, const cabm_tape_entry *entries, uint64_t n);

/* Waits for the run. In-model error()s of the reference (:308 "sickfrom not set right", :371-372 "No susceptibles in the
 * population to change status", :418 "swap expired, but no swap set.", :719) and implementation limits (an empty age group
 * sampled, contact log full) are recorded per replicate on the device; after a run cabm_sync and every cabm_get_* result call
 * return CABM_E_MODEL with that message if any replicate is flagged (cabm_get_scalars only when its err argument is NULL: a
 * caller that reads the error words decides itself). cabm_set_strict(h, 0) switches the check off. */
int cabm_sync(cabm_handle *h);
int cabm_set_strict(cabm_handle *h, int on);
/* device time of the last cabm_run in milliseconds (CUDA events on the handle's stream) */
int cabm_last_run_ms(cabm_handle *h, float *ms);
/* Several handles on one GPU run concurrently (each has its own stream): consecutive batches of an ensemble overlap, the
 * short replicates of one batch filling the SMs that the long epidemics of the previous batch leave idle (the reference
 * gets the same effect from pmap's dynamic scheduling, scripts/local_run.jl:28-30). To time such a pipeline on the
 * device: cabm_mark records an event on the handle's stream now; cabm_ms_since_mark waits for the end of h's last
 * cabm_run and returns the milliseconds since the mark of `marked` (which may be another handle of the same GPU). */
int cabm_mark(cabm_handle *h);
int cabm_ms_since_mark(cabm_handle *marked, cabm_handle *h, float *ms);
/* Per-kernel-class device time of a run: with profiling on, cabm_run brackets every launch with CUDA events on
 * the handle's stream (and synchronises at the end); cabm_get_profile returns the summed milliseconds and the launch
 * count per class. Classes: dyntrans :790-834, time_update :399-428 streaming scan (+snapshot), ct_dynamics pass A :705-716,
 * everything else (initialize, seeding, group tables, reductions), the fused whole-run kernel, time_update's event kernel. */
enum { CABM_K_DYNTRANS = 0, CABM_K_TIME_UPDATE = 1, CABM_K_CT_IDENTIFY = 2, CABM_K_OTHER = 3, CABM_K_FUSED = 4, CABM_K_TU_EVENTS = 5, CABM_N_KCLASS = 6 };
int cabm_set_profiling(cabm_handle *h, int on);
int cabm_get_profile(cabm_handle *h, float *ms /*[CABM_N_KCLASS]*/, int32_t *launches /*[CABM_N_KCLASS]*/);
/* Execution path of cabm_run. FUSED: one persistent CTA per replicate runs all of main() with the per-agent state in
 * shared memory (needs the exact schedule and a popsize that fits, ~21k agents). FUSED_GLOBAL: the same kernel with the
 * per-agent arrays in global memory (L2) and only its work buffers in shared memory: exact schedule, any popsize.
 * MULTIKERNEL: one launch per phase per day over HBM-resident state (any popsize, exact or native schedule; also what the
 * cabm_step_* calls drive). AUTO (default): exact schedule -> FUSED when the population fits, else FUSED_GLOBAL; native
 * schedule -> MULTIKERNEL. All exact paths give bit-identical results. */
enum { CABM_PATH_AUTO = 0, CABM_PATH_MULTIKERNEL = 1, CABM_PATH_FUSED = 2, CABM_PATH_FUSED_GLOBAL = 3 };
int cabm_set_path(cabm_handle *h, int path);
int cabm_get_last_path(const cabm_handle *h);
/* diagnostics of the fused path: per replicate, cycles thread 0 spent in 16 phase slots (tools/diag_fused.py names them):
 * int64 [n_rep][16] */
int cabm_get_debug(cabm_handle *h, int64_t *out);
/* number of kernel launches issued by this handle so far */
int64_t cabm_launch_count(const cabm_handle *h);

/* The individual phases, as the reference's tests drive them on `cv.humans` (test/runtests.jl):
 * initialize() :171-188; insert_infected(health, num, ag) :360-395 (call_id names the seeding call in the
 * draw contract: 0 and 1 are the two calls main makes); dyntrans(st, grps) :790-834;
 * time_update() :399-428. The handle keeps the day index st (1-based), advanced by time_update. */
int cabm_step_initialize(cabm_handle *h);
int cabm_step_insert_infected(cabm_handle *h, int health, int num, int ag, int call_id);
int cabm_step_dyntrans(cabm_handle *h);
int cabm_step_time_update(cabm_handle *h);
int cabm_get_day(const cabm_handle *h);
/* (totalmet, totalinf) returned by the last dyntrans :833, per replicate. Kept by the per-day kernels (cabm_step_dyntrans, the
 * multikernel path); the fused paths do not count them — main discards them (:138) and the fused dyntrans does not even decide
 * the admission of events whose outcome cannot matter. */
int cabm_get_dyntrans_counters(cabm_handle *h, int32_t *totalmet, int32_t *totalinf);

/* hmatrix of one replicate: Int16, popsize x modeltime, column-major (:111,:141). Needs CABM_STORE_HMATRIX. */
int cabm_get_hmatrix(cabm_handle *h, int replicate, int16_t *out);
/* every replicate's matrix, [n_rep][modeltime][popsize] */
int cabm_get_hmatrix_all(cabm_handle *h, int16_t *out);
/* the same matrices as the device keeps them: one byte per agent-day (health codes are 0..11), half the PCIe traffic of the
 * Int16 form; out is [n_replicates][modeltime][popsize] */
int cabm_get_hmatrix_all_u8(cabm_handle *h, uint8_t *out);
/* _get_incidence_and_prev :259-268 for "all" (group 0) and the five _splitstate groups (:247-256), as
 * runsim collects them (:91-97): inc and prev are int32 [n_rep][6][modeltime][12]. */
int cabm_get_curves(cabm_handle *h, int32_t *inc, int32_t *prev);
/* the same curves narrowed to Int16 on the device (every count <= popsize <= 32767): half the device->host bytes */
int cabm_get_curves_i16(cabm_handle *h, int16_t *inc, int16_t *prev);
/* compute_yearly_average (scripts/local_run.jl:75-103): the mean over all replicates of every curve entry, reduced on the
 * device: inc_mean and prev_mean are double [6][modeltime][12] */
int cabm_get_mean_curves(cabm_handle *h, double *inc_mean, double *prev_mean);
/* Optional: register two pinned host buffers (cabm_host_alloc) of [max_replicates][6][modeltime][12] Int16. The fused kernel
 * then writes every finished replicate's curves straight into them while the other replicates are still running, and a
 * following cabm_get_curves_i16 with the same pointers only synchronises. Pass NULL, NULL to unregister. */
int cabm_set_host_curves_i16(cabm_handle *h, int16_t *inc, int16_t *prev);
/* Optional, with the buffers above: rows[replicate] (pinned, int32 [max_replicates]) receives the number of leading rows (days) of
 * that replicate's curves the kernel sent. The rows after them were not written: they repeat — incidence 0, prevalence equal to
 * the last row sent — because nobody was infectious and nothing was due any more (most replicates of a sub-critical scenario die
 * out within weeks: about a quarter of the bytes cross PCIe). NULL switches it off (every row is sent). modeltime must be even,
 * otherwise every row is sent and rows[r] = modeltime. */
int cabm_set_host_curve_rows(cabm_handle *h, int32_t *rows);
/* infectors = _count_infectors() :295-313 [n_rep][4]; totalisolated = ct_data.totalisolated :61,649 [n_rep];
 * lat_inc_sum = sum(_get_column_incidence(h, LAT)) (scripts/local_run.jl:143) [n_rep]; err = CABM_ERR_* bits.
 * Any pointer may be NULL. */
int cabm_get_scalars(cabm_handle *h, int64_t *infectors, int64_t *totalisolated, int64_t *lat_inc_sum,
                     uint32_t *err);
/* one member of every Human of one replicate, widened to int32 (CABM_F_DUR: 4 per agent) */
int cabm_get_field(cabm_handle *h, int field, int replicate, int32_t *out);
int cabm_set_field(cabm_handle *h, int field, int replicate, int agent, int32_t value);

/* _get_column_incidence :270-280 and _get_column_prevalence :282-293 over caller-supplied matrices:
 * hm is Int16 [n_mat][t][n] (each popsize x modeltime column-major), ags[n_mat][n] in 1..5;
 * inc/prev are int32 [n_mat][6][t][12]. Runs on the handle's device and stream. */
int cabm_reduce_hmatrix(cabm_handle *h, const int16_t *hm, int n_mat, int n, int t, const int32_t *ags,
                        int32_t *inc, int32_t *prev);

/* pinned host memory for result buffers (optional; any host pointer is accepted by cabm_get_*) */
void *cabm_host_alloc(uint64_t bytes);
void cabm_host_free(void *p);

/* draw-contract tables for cross-checking: name in {"lat","pre","asy","inf","psi","mu",
 * "nb1".."nb5"}; writes up to cap thresholds, returns the table length or -1. Host-only, needs no GPU. */
int cabm_get_table(const char *name, int32_t *base, uint32_t *thr, int cap);
/* gpw = Int.(round.(cm[ag]*cnts)) :804, ag 1..5. Host-only. */
void cabm_get_gpw(int ag, int cnts, int32_t out[5]);
const char *cabm_version(void);

#ifdef __cplusplus
}
#endif
#endif
