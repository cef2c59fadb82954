// cabm_device.cuh — HBM data layout, draw contract and per-agent device functions (sm_100a).
//
// Layout: the reference's array-of-structs `humans` (src/covid19abm.jl:8-27,69) becomes a
// struct-of-arrays with a replicate dimension, every array [R][Np] (Np = popsize padded to 32):
//   status  u32  : rem[0:8) health[8:12) swap[12:16) ag[16:19) iso[19] isob[20] tracing[21] tag[22:32)
//                  - everything dyntrans gathers about a contact target is in this one word, so one
//                    32-bit load (or one atomicCAS in native mode) resolves a contact event;
//                  - (tag, rem) is the agent's *remaining* contact budget, valid only when tag == day:
//                    nextday_meetcnt (:13,:786) is a pure function of the keyed draw, so it is evaluated
//                    lazily for infectors and their targets instead of once per agent-day (:425,:781);
//                  - isob is `iso` as of the agent's own last update step (the value :778 used).
//   expday  i16  : the agent's next event day = the scan key of time_update: min of texp and the contact-tracing days below
//   texp    i16  : absolute day on which tis >= exp (:405) becomes true   (replaces tis/exp :16-17)
//   entry   i16  : day of entry to the current state (tis = updates_done - entry)
//   dur     u32  : 4 x u8 (latents, asymps, pres, infs) :18
//   sickfrom/isovia u8, latday i16 (doi = updates_done - latday :19), tracestart/traceend i16 (:23-24, relative to latday),
//   txpday i16 (absolute day on which tracedxp :26 reaches 0; 0 = no countdown),
//   ct_head u32 -> newest chunk of the agent in the per-replicate contact log (replaces the Bool[10000] of :25).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace cabm {

enum : int { SUS = 0, LAT, PRE, ASYMP, MILD, MISO, INF, IISO, HOS, ICU, REC, DED, UNDEF };
enum : int { VIA_NULL = 0, VIA_QU, VIA_PI, VIA_MI, VIA_CT };
// draw-contract streams
enum : uint32_t { S_INIT = 1, S_COUNT = 2, S_SEED = 3, S_TRANS = 4, S_CTWIN = 5, S_CTCON = 6, S_CT3 = 7, S_DT = 8 };
enum : uint32_t { ERR_NO_SWAP = 1, ERR_NO_SUS = 2, ERR_TRACED_NOT_ISO = 4, ERR_EMPTY_GROUP = 8, ERR_CT_LOG_FULL = 16, ERR_SICKFROM = 32, ERR_INTERNAL = 64 };

constexpr uint32_t ST_ISO = 1u << 19, ST_ISOB = 1u << 20, ST_TRACING = 1u << 21;
constexpr uint32_t ST_TAGMASK = 0xFFC00000u, ST_REMMASK = 0xFFu;
constexpr uint32_t CT_NONE = 0xFFFFFFFFu;
constexpr unsigned FULL = 0xFFFFFFFFu;

__host__ __device__ inline int st_rem(uint32_t s) { return (int)(s & 0xFF); }
__host__ __device__ inline int st_health(uint32_t s) { return (int)((s >> 8) & 0xF); }
__host__ __device__ inline int st_swap(uint32_t s) { return (int)((s >> 12) & 0xF); }
__host__ __device__ inline int st_ag(uint32_t s) { return (int)((s >> 16) & 0x7); }   // 0..4
__host__ __device__ inline int st_tag(uint32_t s) { return (int)(s >> 22); }
__host__ __device__ inline uint32_t st_set_health(uint32_t s, int h) { return (s & ~0xF00u) | ((uint32_t)h << 8); }
__host__ __device__ inline uint32_t st_set_swap(uint32_t s, int h) { return (s & ~0xF000u) | ((uint32_t)h << 12); }
__host__ __device__ inline uint32_t st_make(int health, int swap, int ag0, bool iso) {
    return ((uint32_t)health << 8) | ((uint32_t)swap << 12) | ((uint32_t)ag0 << 16) | (iso ? (ST_ISO | ST_ISOB) : 0u);
}
__host__ __device__ inline bool is_infectious(int h) { return h >= PRE && h <= IISO; }  // :794
// _get_betavalue class :761-767: 0 PRE, 1 ASYMP, 2 MILD/MISO, 3 INF/IISO
__host__ __device__ inline int beta_class(int h) { return (int)((0xFFFFFA4Fu >> (2 * h)) & 3u); }    // two bits per health code, no branches

// per-parameter-set constants, thresholds are "w < thr  <=>  rand() < p"
struct DevPset {
    unsigned long long asymp_thr[5], mild_thr[5], age_thr[4];
    unsigned long long fmild_thr, fsevere_thr, fpreiso_thr, quar_thr, fctcap_thr, fcontact_thr;
    int tau_mild, tpreiso, quar_grp0, ctstrat, cdaysback, strat3qdays, calibration, initialinf, initialhi, qdays;
};

// small tables in constant memory
struct ConstTables {
    uint32_t lat_thr[4], pre_thr[4], asy_thr[40], inf_thr[40], psi_thr[12], mu_thr[8];
    int lat_len, pre_len, asy_len, inf_len, psi_len, mu_len;
    int lat_base, pre_base, asy_base, inf_base, psi_base, mu_base;
    int nb_len[5];
    double hlo[5], hspan[5], clo[5], cspan[5];                 // move_to_inf :535-536
    double cm[5][5];                                           // contact_matrix :838-848
    unsigned long long mh_inf_thr[5];                          // :537
    unsigned long long mh_hos_thr[5], mh_icu_thr[5];           // :585-586
    unsigned long long ct3_lat_thr, ct3_pre_thr, ct3_asymp_thr; // :741
};

struct Dev {
    int R, N, Np, T, P, store_hm, nwords;
    uint32_t ct_cap;
    uint32_t *status; int16_t *expday; int16_t *texp; int16_t *entry; uint32_t *dur; uint8_t *sickfrom; uint8_t *isovia;
    int16_t *latday, *tracestart, *traceend, *txpday; uint32_t *ct_head;
    uint32_t *mark_pre, *mark_post;          // [R][nwords] bitmasks
    uint32_t *ct_log; uint32_t *ct_cnt;      // [R][ct_cap] words: chunks [next chunk, count, targets...] per (traced agent, day); [R] words used
    uint32_t *ident_list, *ident_cnt;        // [R][2][Np], [R][2]: agents whose identification day is today / tomorrow
    uint32_t *ev_list, *ev_cnt;              // [R][Np], [R][2]: agents with an event today (k_time_update_scan -> k_time_update_events), count by day parity
    uint32_t *nat_edges, *nat_ecnt;          // native schedule: [R][Np][3] events of the day whose target is a later infector (x, j, k|g<<8|budget<<16|ag<<24), [R] their number
    uint32_t *nat_verify;                    // test aid (CABM_NATIVE_VERIFY=1): the second-order counts recomputed by a full sweep, compared on the device
    uint32_t *grp_idx; int *grp_off;         // [R][Np], [R][8]
    uint16_t *grp16;                         // [R][Np] the same index tables as 16-bit words: the fused kernel's copy (popsize <= 65535)
    uint32_t *infmask;                       // [R][nwords]
    uint8_t *col_g;                          // [R][Np] its current hmatrix column
    uint32_t *hits;                          // 2 x [R][Np/2] of 2 x u16: native mode, contacts received from lower-indexed infectors
    uint2 *key; int *pset_of; uint32_t *err; unsigned long long *totiso; int *totalmet, *totalinf;
    int *inc, *prev;                         // [R][6][T][12]
    long long *infectors, *latsum;           // [R][4], [R]
    long long *dbg;                          // [R][32] per-phase cycle counts of the fused kernel (diagnostics)
    int *resume_list, *resume_aux;           // [R], [R][2]: fused kernel's queue of parked replicates and their (nextfire, ninf)
    int *copy_list;                          // [R]: fused kernel's queue of finished replicates whose curves go to the host
    uint8_t *hm;                             // [R][T][N] health codes (widened to Int16 by cabm_get_hmatrix)
    int *host_rows;                          // optional pinned host array [R]: rows of a replicate's curves that were sent (the rest repeat)
    int16_t *host_inc, *host_prev;           // optional pinned host buffers [R][6][T][12]: the fused kernel streams each finished
                                             // replicate's curves there (Int16) so no device->host copy is left at the end
    const DevPset *psets; const unsigned long long *beta_thr;    // [P], [P][T][4]
    const uint32_t *nb_thr; const uint8_t *nb_guide;             // [5][256], [5][512] (top byte of w; next byte inside the top bucket)
    const uint32_t *tape_keys; const uint4 *tape_w; unsigned tape_n;   // replay tape (cabm_set_tape): sorted keys (seed lo, seed hi, counter x4), words
    const unsigned long long *gpw;                               // [5][256] five u8 counts packed
};

#ifdef CABM_KERNELS_TU   // device code below is compiled only in cabm_kernels.cu (no -rdc: constants are per-TU)
__constant__ ConstTables c_tab;

// ---------------------------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint2 key) {
    uint32_t k0 = key.x, k1 = key.y;
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t h0 = __umulhi(0xD2511F53u, c0), l0 = 0xD2511F53u * c0;
        uint32_t h1 = __umulhi(0xCD9E8D57u, c2), l1 = 0xCD9E8D57u * c2;
        c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    return make_uint4(c0, c1, c2, c3);
}
// Replay tape (cabm_set_tape): a draw whose (seed, counter) is on the tape takes the tape's words instead of the Philox
// output. Only the per-day kernels consult it (TP = true; the fused kernel never runs with a tape), through one uniform test.
__device__ __noinline__ bool tape_lookup(const uint32_t *keys, const uint4 *words, unsigned n, uint2 key, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint4 *out) {
    const uint32_t want[6] = {key.x, key.y, c0, c1, c2, c3};
    unsigned lo = 0, hi = n;
    while (lo < hi) {
        const unsigned mid = (lo + hi) >> 1;
        const uint32_t *k = keys + (size_t)mid * 6;
        int cmp = 0;
        // order: seed (hi word first, as a 64-bit number), then the counter words
        const uint32_t a[6] = {k[1], k[0], k[2], k[3], k[4], k[5]}, b[6] = {want[1], want[0], want[2], want[3], want[4], want[5]};
        for (int j = 0; j < 6 && cmp == 0; ++j) cmp = a[j] < b[j] ? -1 : a[j] > b[j] ? 1 : 0;
        if (cmp == 0) { *out = words[mid]; return true; }
        if (cmp < 0) lo = mid + 1; else hi = mid;
    }
    return false;
}
template <bool TP> __device__ __forceinline__ uint4 draw4(const Dev &S, uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint2 key) {
    if (TP && S.tape_n) { uint4 o; if (tape_lookup(S.tape_keys, S.tape_w, S.tape_n, key, c0, c1, c2, c3, &o)) return o; }
    return philox(c0, c1, c2, c3, key);
}
__device__ __forceinline__ uint32_t pick(uint4 w, uint32_t k) { return k == 0 ? w.x : k == 1 ? w.y : k == 2 ? w.z : w.w; }
__device__ __forceinline__ int below(uint32_t w, int n) { return (int)__umulhi(w, (uint32_t)n); }
__device__ __forceinline__ bool bern(uint32_t w, unsigned long long thr) { return (unsigned long long)w < thr; }
__device__ __forceinline__ double u01(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }
__device__ __forceinline__ int table_draw(uint32_t w, int base, const uint32_t *thr, int len) {
    int v = 0;
    while (v < len && thr[v] <= w) ++v;
    return base + v;
}

// the same count by a branch-free binary search (tables of up to 63 thresholds: no divergence between lanes)
__device__ __forceinline__ int table_draw_bs(uint32_t w, int base, const uint32_t *thr, int len) {
    int lo = 0;
#pragma unroll
    for (int step = 32; step; step >>= 1) { const int m = lo + step; if (m <= len && thr[m - 1] <= w) lo = m; }
    return base + lo;
}

// get_nextday_counts :772-788, evaluated on demand for day `day` from the agent's status word
template <bool TP = false> __device__ __forceinline__ int fresh_budget(const Dev &S, uint2 key, uint32_t i, uint32_t day, uint32_t st) {
    if (st_health(st) == DED) return 0;                                    // :783-785
    uint32_t w = pick(draw4<TP>(S, i >> 2, day, S_COUNT, 0, key), i & 3u);
    if (st & ST_ISOB)                                                      // :778-779
        return (w < 0x80000000u) ? 0 : 1 + (int)(((unsigned long long)(w & 0x7fffffffu) * 3ull) >> 31);
    int ag = st_ag(st);                                                    // :781
    const uint32_t *thr = S.nb_thr + ag * 256;
    int k = __ldg(&S.nb_guide[ag * 512 + ((w >> 24) == 255u ? 256u + ((w >> 16) & 0xFFu) : (w >> 24))]), len = c_tab.nb_len[ag];
    while (k < len && __ldg(&thr[k]) <= w) ++k;
    return k;
}
template <bool TP = false> __device__ __forceinline__ int budget_of(const Dev &S, uint2 key, uint32_t i, int day, uint32_t st) {
    return st_tag(st) == day ? st_rem(st) : fresh_budget<TP>(S, key, i, (uint32_t)day, st);
}

__device__ __forceinline__ void curve_move(const Dev &S, int r, int ag0, int col, int oldh, int newh) {
    // column `col` (0-based) is the first snapshot that shows the new state (:136,:218-223)
    if (col >= S.T || oldh == newh) return;
    size_t b = (((size_t)r * 6 + (size_t)(ag0 + 1)) * (size_t)S.T + (size_t)col) * 12;
    atomicAdd(&S.inc[b + newh], 1);
    atomicAdd(&S.prev[b + oldh], 1);   // exits are parked in prev[] until k_curves_finalize
}

// One expired agent: time_update's @match :405-419 and the move_to_* functions :446-644.
// transition_core returns the new status word and expiry day; it writes only the rarely-touched per-agent fields
// that always live in HBM (entry, latday, isovia). The caller stores status/expday (HBM arrays in the per-day
// kernels, shared memory in the fused kernel) and keeps the infectious bitmask and the incidence counters.
// `day` is st; for seeding it is 0.
struct Tr { uint32_t st; int expday; int oldh, newh; };

constexpr int DUR_CACHE = 256;          // direct-mapped shared-memory cache of `dur` words for infected agents (fused kernel)

template <bool TP = false> __device__ inline Tr transition_core(const Dev &S, const DevPset &P, int r, int i, int day, uint32_t st, uint2 key,
                                     int target /* swap to apply */, bool seed_simple_inf, unsigned long long *dcache = nullptr,
                                     int *via_io = nullptr /* the agent's isovia held by the caller: no load, no store here */) {
    size_t a = (size_t)r * S.Np + i;
    const int oldh = st_health(st), ag = st_ag(st);
    // dcache (fused kernel): direct-mapped shared-memory cache of the two rarely-touched fields a transition reads,
    // entry = agent [40:64) | isovia [32:35) | dur [0:32). An infected agent makes ~5 transitions: all but the first hit.
    int via0;
    uint32_t d4;
    if (dcache) {
        const unsigned long long e = dcache[i & (DUR_CACHE - 1)];
        if ((uint32_t)(e >> 40) == (uint32_t)i) { d4 = (uint32_t)e; via0 = (int)((e >> 32) & 7u); }
        else { d4 = S.dur[a]; via0 = (int)S.isovia[a]; }
        if (via_io) via0 = *via_io;      // (with contact tracing the caller holds isovia: the cached copy is not used)
    } else {
        via0 = via_io ? *via_io : (int)S.isovia[a];   // needed by a few targets only, but loaded here so that its latency hides behind the draws
        d4 = S.dur[a];
    }
    const int d_lat = d4 & 0xFF, d_asy = (d4 >> 8) & 0xFF, d_pre = (d4 >> 16) & 0xFF, d_inf = d4 >> 24;
    bool iso = st & ST_ISO;
    int via = -1;                       // -1: unchanged
    int newh = target, newsw = UNDEF, exp = 999;
    if (target < LAT || target > DED) {      // :418 "swap expired, but no swap set."
        atomicOr(&S.err[r], ERR_NO_SWAP);
        return Tr{st, 32767, oldh, oldh};
    }
    // Every move_to_* that draws uses the words of the same first call, so it is made once for all targets (entries without draws
    // ignore it), and the common entries are written as selects rather than a switch: agents of one warp that enter different
    // states on the same day then run in lockstep instead of one case after the other. Hospital-bound entries branch off.
    const uint4 w = draw4<TP>(S, i, day, S_TRANS, 0, key);
    if ((target == INF && !seed_simple_inf) || target == HOS || target == ICU) {
        if (target == INF) {                 // move_to_inf :530-569
            const uint4 wb = draw4<TP>(S, i, day, S_TRANS, 1, key);
            double h = __dadd_rn(c_tab.hlo[ag], __dmul_rn(c_tab.hspan[ag], u01(w.x)));
            double c = __dadd_rn(c_tab.clo[ag], __dmul_rn(c_tab.cspan[ag], u01(w.y)));
            unsigned long long mthr = c_tab.mh_inf_thr[ag];
            if (P.calibration) { h = 0; c = 0; mthr = 0; }
            const int tth = (int)rint(__dadd_rn(2.0, __dmul_rn(3.0, u01(w.z))));
            if (u01(w.w) < h) { exp = tth; newsw = (u01(wb.x) < c) ? ICU : HOS; }
            else if (bern(wb.y, mthr)) { exp = d_inf; newsw = DED; }
            else {
                exp = d_inf; newsw = REC;
                if (iso || bern(wb.z, P.fsevere_thr)) { exp = 1; newsw = IISO; }
            }
        } else {                             // move_to_hospicu :580-622
            const int psi = table_draw(w.x, c_tab.psi_base, c_tab.psi_thr, c_tab.psi_len);
            const int mu = table_draw(w.y, c_tab.mu_base, c_tab.mu_thr, c_tab.mu_len);
            iso = true;                                                                    // :599
            if (target == HOS) { if (bern(w.z, c_tab.mh_hos_thr[ag])) { exp = mu; newsw = DED; } else { exp = psi; newsw = REC; } }
            else { if (bern(w.z, c_tab.mh_icu_thr[ag])) { exp = mu + 2; newsw = DED; } else { exp = psi + 2; newsw = REC; } }
        }
    } else {
        const bool tLAT = target == LAT, tPRE = target == PRE, tMILD = target == MILD;
        const bool b = bern(w.x, tLAT ? P.asymp_thr[ag] : tPRE ? P.mild_thr[ag] : P.fmild_thr);   // :455, :481, :503
        exp = tLAT ? d_lat                                           // move_to_latent :446-461
            : target == ASYMP ? d_asy                                // move_to_asymp :464-472
            : tPRE ? d_pre                                           // move_to_pre :475-488
            : (tMILD || target == INF) ? d_inf                       // move_to_mild :491-507, move_to_infsimple :520-528 (seeds)
            : target == MISO ? d_inf - P.tau_mild                    // move_to_miso :510-517
            : target == IISO ? d_inf - 1                             // move_to_iiso :571-578
            : 999;                                                   // move_to_recovered :635-644, move_to_dead :624-633
        newsw = tLAT ? (b ? ASYMP : PRE) : tPRE ? (b ? MILD : INF) : (target == REC || target == DED) ? UNDEF : REC;
        if (tLAT) {
            if (P.calibration) { newsw = LAT; exp = 999; }           // :457-460
            S.latday[a] = (int16_t)day;
        }
        if (tPRE) {
            const bool active = (day == 0) || (P.tpreiso >= 1 && day >= P.tpreiso);   // main :124-125,:133-135
            if (active && bern(w.y, P.fpreiso_thr)) { iso = true; via = VIA_PI; }     // :487
        }
        if (tMILD && (iso || b)) { newsw = MISO; exp = P.tau_mild; }                  // :503-506
        if (target == MISO || target == IISO) { iso = true; via = VIA_MI; }
        if (target == INF) iso = false;                              // _set_isolation(x, false, :null) leaves isovia as is
        if (target == DED) iso = true;
    }
    int newvia = via0;
    if (via >= 0 && via0 == VIA_NULL) { newvia = via; if (via_io) *via_io = via; else S.isovia[a] = (uint8_t)via; }   // _set_isolation :432-444 only fills a :null isovia
    if (dcache) dcache[i & (DUR_CACHE - 1)] = ((unsigned long long)(uint32_t)i << 40) | ((unsigned long long)(uint32_t)newvia << 32) | d4;
    S.entry[a] = (int16_t)day;
    st = st_set_swap(st_set_health(st, newh), newsw);
    st = iso ? (st | ST_ISO | ST_ISOB) : (st & ~(ST_ISO | ST_ISOB));
    return Tr{st, day + exp, oldh, newh};
}

// HBM-resident variants used by the per-day kernels: flip the infectious bitmask, count the curves (transition_hbm) and
// store the expiry day both as texp and as the scan key (transition).
template <bool TP = false> __device__ inline Tr transition_hbm(const Dev &S, const DevPset &P, int r, int i, int day, uint32_t st, uint2 key,
                                    int target, bool seed_simple_inf, bool count_curves) {
    const Tr t = transition_core<TP>(S, P, r, i, day, st, key, target, seed_simple_inf);
    if (is_infectious(t.newh) != is_infectious(t.oldh)) atomicXor(&S.infmask[(size_t)r * S.nwords + (i >> 5)], 1u << (i & 31));
    if (count_curves) curve_move(S, r, st_ag(st), day, t.oldh, t.newh);
    return t;
}
template <bool TP = false> __device__ inline uint32_t transition(const Dev &S, const DevPset &P, int r, int i, int day, uint32_t st, uint2 key,
                                      int target, bool seed_simple_inf, bool count_curves) {
    const Tr t = transition_hbm<TP>(S, P, r, i, day, st, key, target, seed_simple_inf, count_curves);
    S.texp[(size_t)r * S.Np + i] = (int16_t)t.expday;
    S.expday[(size_t)r * S.Np + i] = (int16_t)t.expday;
    return t.st;
}

// time_update's body for one agent that has an event today (see k_time_update in cabm_kernels.cu for the ordering notes).
// Takes and returns the status word and the state-expiry day; touches the HBM-resident tracing fields directly; the caller
// stores status / texp / next event day and does the infectious-set and curve accounting from (oldh, newh).
struct Ev { uint32_t st; int texp, ne, oldh, newh; };

template <bool TP = false> __device__ inline Ev agent_event_core(const Dev &S, const DevPset &P, int r, int i, int day, uint2 key, uint32_t st, int texp,
                                      bool ct /* the replicate traces contacts */, bool pre, bool post,
                                      uint32_t *queue_cnt /* tomorrow's index cases */, uint32_t *queue, unsigned long long *dcache = nullptr) {
    const size_t a = (size_t)r * S.Np + i;
    int oldh = st_health(st), newh = oldh;
    // the agent's tracing fields, loaded together up front (independent loads: one round trip to L2) and written back at the end
    int txp = 0, txp0 = 0, via = 0, via_in = 0, latday = 0, ts = 0, te = 0, ts0 = 0, te0 = 0;
    if (ct) {
        txp = txp0 = S.txpday[a]; via = via_in = S.isovia[a]; latday = S.latday[a]; ts = ts0 = S.tracestart[a]; te = te0 = S.traceend[a];
        if (pre) {                                             // apply_ct_strategy :647-656 by an index case with a lower index
            st |= ST_ISO;
            if (via == VIA_NULL) via = VIA_CT;
            txp = P.qdays > 0 ? day + P.qdays - 1 : 0;         // counted down once more in this same update
        }
    }
    if (texp <= day || (st_health(st) == SUS && st_swap(st) == LAT)) {      // :405 (an infection today set exp = tis :826)
        const int target = st_swap(st);
        const Tr t = transition_core<TP>(S, P, r, i, day, st, key, target, false, dcache, ct ? &via : nullptr);
        st = t.st; texp = t.expday; oldh = t.oldh; newh = t.newh;
        if (target == LAT) latday = day;                       // move_to_latent set doi = 0 (:449; transition_core wrote latday)
    }
    int ne = texp;
    if (ct) {
        const int doi = day - latday;
        if (st_health(st) == LAT && st_swap(st) != ASYMP && doi == 0) {      // :678
            const uint4 w = draw4<TP>(S, i, day, S_CTWIN, 0, key);
            if (bern(w.x, P.fctcap_thr)) {
                const uint32_t d4 = S.dur[a];
                const int tid = 1 + below(w.y, (int)(d4 >> 24));              // :680
                const int delta = (int)(d4 & 0xFF) + (int)((d4 >> 16) & 0xFF) + tid;
                const int q = delta - P.cdaysback;
                ts = q > 0 ? q : 0; te = delta;
                st &= ~ST_TRACING;
                S.ct_head[a] = CT_NONE;                                        // :687
            }
        }
        if (doi >= ts && doi < te) st |= ST_TRACING;                          // :699-701
        if (doi == te) {                                                      // :705-716 (contacts were handled in pass A)
            st &= ~ST_TRACING;
            st |= ST_ISO;
            if (via == VIA_NULL) via = VIA_CT;
            txp = P.qdays > 0 ? day + P.qdays - 1 : 0;
            atomicAdd(&S.totiso[r], 1ull);
            S.ct_head[a] = CT_NONE;
        }
        if (txp != 0 && txp >= day) {                                         // :718 tracedxp > 0
            if (!(st & ST_ISO)) atomicOr(&S.err[r], ERR_TRACED_NOT_ISO);
            if (txp == day) {                                                 // :734 the countdown ends today
                st &= ~ST_ISO;
                via = VIA_NULL;
                txp = 0;
                if (P.ctstrat == 3) {
                    const uint4 w = draw4<TP>(S, i, day, S_CT3, 0, key);
                    const int h = st_health(st);
                    if ((h == LAT && bern(w.x, c_tab.ct3_lat_thr)) || (h == PRE && bern(w.x, c_tab.ct3_pre_thr)) ||
                        (h == ASYMP && bern(w.x, c_tab.ct3_asymp_thr)) || h == MILD || h == INF || h == MISO || h == IISO) {
                        st |= ST_ISO; via = VIA_CT; txp = day + 14;
                    }
                }
            }
        }
    }
    st = (st & ST_ISO) ? (st | ST_ISOB) : (st & ~ST_ISOB);
    if (ct) {
        if (post) {                                            // apply_ct_strategy by an index case with a higher index: after this agent's own step
            st |= ST_ISO;
            if (via == VIA_NULL) via = VIA_CT;
            txp = P.qdays > 0 ? day + P.qdays : 0;
        }
        if (txp != txp0) S.txpday[a] = (int16_t)txp;
        if (via != via_in) S.isovia[a] = (uint8_t)via;
        if (ts != ts0 || te != te0) { S.tracestart[a] = (int16_t)ts; S.traceend[a] = (int16_t)te; }
        // next event: state expiry, tracing on, the eve of identification (to queue for pass A), identification, end of quarantine
        if (te >= 0) {
            const int idd = latday + te, tron = latday + ts;
            if (ts < te && tron > day) ne = min(ne, tron);
            if (idd - 1 > day) ne = min(ne, idd - 1);
            if (idd > day) ne = min(ne, idd);
            if (idd == day + 1) {                              // identifies tomorrow (:705): queue for pass A
                const uint32_t slot = atomicAdd(queue_cnt, 1u);
                queue[slot] = (uint32_t)i;
            }
        }
        if (txp > day) ne = min(ne, txp);
        // a post mark isolates after the budget draw of today (:778 saw the old iso): tomorrow's own step must see the new one
        if (((st >> 19) ^ (st >> 20)) & 1u) ne = min(ne, day + 1);
    }
    return Ev{st, texp, ne, oldh, newh};
}


#endif  // CABM_KERNELS_TU

}  // namespace cabm
