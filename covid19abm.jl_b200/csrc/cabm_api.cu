// cabm_api.cu — the C ABI of include/cabm.h: handle, parameter validation, the day loop of main (:104-142).
// No CPU fallback: every entry point that computes needs the CUDA device the handle was created on.
#include "../../include/cabm.h"
#include "cabm_kernels.h"
#include "cabm_tables.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

using namespace cabm;

static_assert((int)CABM_UNDEF == (int)UNDEF && (int)CABM_DED == (int)DED, "HEALTH codes");

struct cabm_handle {
    int device = 0, R = 0, N = 0, Np = 0, T = 0, flags = 0;
    int n_rep = 0, n_psets = 0, mode = 0, day = 1, any_ct = 0, all_ct = 0, configured = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev_mark = nullptr;
    Dev S{};
    bool native_edges_tried = false;
    std::vector<void *> allocs;
    DevPset *d_psets = nullptr; unsigned long long *d_beta = nullptr; int psets_cap = 0;
    int *d_scratch = nullptr;          // 4*N ints for peeks
    unsigned *d_work = nullptr;        // work counter of the fused kernel
    int16_t *d_stage = nullptr; size_t stage_elems = 0;   // Int16 staging for hmatrix read-back
    int n_sm = 148, path = CABM_PATH_AUTO, last_path = 0;
    int16_t *host_inc = nullptr, *host_prev = nullptr;   // registered by cabm_set_host_curves_i16
    int64_t launches = 0;
    bool ran = false;
    // in-model error()s of the reference (:308,:371-372,:418,:719) surface as CABM_E_MODEL from cabm_sync / cabm_get_* after a run
    int strict = 1; bool err_checked = false; int err_rc = 0;
    size_t ct_words_per_agent = 0;     // contact-log capacity the current allocation was sized for
    double *d_mean = nullptr; size_t mean_elems = 0;   // scratch of cabm_get_mean_curves
    uint32_t *d_tape_keys = nullptr; uint4 *d_tape_w = nullptr;   // replay tape (cabm_set_tape)
    bool poked = false;                // state was edited through cabm_set_field: pass A falls back to the full scan
    // optional per-kernel-class timing (cabm_set_profiling): events around every launch of the next cabm_run
    int profiling = 0;
    std::vector<cudaEvent_t> pev; std::vector<int> pcls; size_t pev_used = 0;
    float prof_ms[CABM_N_KCLASS] = {0}; int prof_n[CABM_N_KCLASS] = {0};
    // per-day kernel path: the ~3 launches per day of a run are captured once into a CUDA graph and replayed while nothing the
    // kernels' by-value arguments depend on has changed (hash of the Dev block and of the run's switches)
    cudaGraphExec_t graph = nullptr; uint64_t graph_key = 0, seen_key = 0; int seen_count = 0; int64_t graph_launches = 0;
    std::string err;
};

static thread_local std::string g_create_err;     // per calling thread: distinct handles (and their creation) are independent

#define CU(call)                                                                                       \
    do {                                                                                               \
        cudaError_t e_ = (call);                                                                       \
        if (e_ != cudaSuccess) {                                                                       \
            h->err = std::string(#call) + ": " + cudaGetErrorString(e_);                               \
            return CABM_E_CUDA;                                                                        \
        }                                                                                              \
    } while (0)

static int fail(cabm_handle *h, int code, const std::string &msg) { h->err = msg; return code; }

extern "C" const char *cabm_last_error(const cabm_handle *h) { return h ? h->err.c_str() : g_create_err.c_str(); }
extern "C" const char *cabm_version(void) { return "covid19abm-b200 0.1 (sm_100a)"; }

template <class Tp> static cudaError_t dalloc(cabm_handle *h, Tp **p, size_t count) {
    void *q = nullptr;
    cudaError_t e = cudaMalloc(&q, count * sizeof(Tp) > 0 ? count * sizeof(Tp) : 16);
    if (e == cudaSuccess) { h->allocs.push_back(q); *p = (Tp *)q; }
    return e;
}

static ConstTables make_const_tables() {
    const Tables &t = tables();
    ConstTables c;
    memset(&c, 0, sizeof c);
    auto put = [](const Table &s, uint32_t *dst, int cap, int &len, int &base) {
        len = (int)s.thr.size(); base = s.base;
        for (int i = 0; i < len && i < cap; ++i) dst[i] = s.thr[i];
    };
    put(t.lat, c.lat_thr, 4, c.lat_len, c.lat_base); put(t.pre, c.pre_thr, 4, c.pre_len, c.pre_base);
    put(t.asy, c.asy_thr, 40, c.asy_len, c.asy_base); put(t.inf, c.inf_thr, 40, c.inf_len, c.inf_base);
    put(t.psi, c.psi_thr, 12, c.psi_len, c.psi_base); put(t.mu, c.mu_thr, 8, c.mu_len, c.mu_base);
    for (int g = 0; g < 5; ++g) c.nb_len[g] = (int)t.nb[g].thr.size();
    // move_to_inf :535-537 — Uniform(a,b) draws are a + (b-a)*rand()
    const double hlo[5] = {0.02, 0.02, 0.28, 0.28, 0.60}, hhi[5] = {0.03, 0.03, 0.34, 0.34, 0.68};
    const double clo[5] = {0.01, 0.01, 0.03, 0.05, 0.05}, chi[5] = {0.015, 0.015, 0.05, 0.20, 0.15};
    const double mh_inf[5] = {0.01 / 5, 0.01 / 5, 0.0135 / 3, 0.01225 / 1.5, 0.04 / 2};
    const double mh_hos[5] = {0.01 * 2, 0.01 * 2, 0.0135, 0.01225, 0.03 * 2};          // :585
    for (int g = 0; g < 5; ++g) {
        c.hlo[g] = hlo[g]; c.hspan[g] = hhi[g] - hlo[g]; c.clo[g] = clo[g]; c.cspan[g] = chi[g] - clo[g];
        c.mh_inf_thr[g] = bernoulli_threshold(mh_inf[g]);
        c.mh_hos_thr[g] = bernoulli_threshold(mh_hos[g]);
        c.mh_icu_thr[g] = bernoulli_threshold(2 * mh_hos[g]);                           // :586
    }
    for (int g = 0; g < 5; ++g) for (int q = 0; q < 5; ++q) c.cm[g][q] = CONTACT_MATRIX[g][q];
    c.ct3_lat_thr = bernoulli_threshold(0.05); c.ct3_pre_thr = bernoulli_threshold(0.95); c.ct3_asymp_thr = bernoulli_threshold(0.70);
    return c;
}

extern "C" int cabm_create(cabm_handle **out, int device, int max_replicates, int popsize, int modeltime, int flags) {
    if (!out || max_replicates < 1 || popsize < 1 || modeltime < 1 || modeltime > 1000) {
        g_create_err = popsize == 0 ? "no population size given" : "cabm_create: bad argument";
        return CABM_E_ARG;
    }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
        g_create_err = std::string("cabm_create: no usable CUDA device (") + (e != cudaSuccess ? cudaGetErrorString(e) : "device index out of range") + "); there is no CPU fallback";
        return CABM_E_CUDA;
    }
    cabm_handle *h = new cabm_handle();
    h->device = device; h->R = max_replicates; h->N = popsize; h->Np = (popsize + 31) / 32 * 32; h->T = modeltime; h->flags = flags;
    auto bail = [&](cudaError_t er, const char *what) {
        g_create_err = std::string("cabm_create: ") + what + ": " + cudaGetErrorString(er);
        cabm_destroy(h);
        return CABM_E_CUDA;
    };
    if ((e = cudaSetDevice(device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "stream");
    if ((e = cudaEventCreate(&h->ev0)) != cudaSuccess || (e = cudaEventCreate(&h->ev1)) != cudaSuccess || (e = cudaEventCreate(&h->ev_mark)) != cudaSuccess) return bail(e, "event");
    Dev &S = h->S;
    S.R = h->R; S.N = h->N; S.Np = h->Np; S.T = h->T; S.store_hm = (flags & CABM_STORE_HMATRIX) ? 1 : 0; S.nwords = h->Np / 32;
    S.ct_cap = 0;
    const size_t A = (size_t)h->R * h->Np, W = (size_t)h->R * S.nwords, C = (size_t)h->R * 6 * h->T * 12;
#define DA(ptr, count) if ((e = dalloc(h, &(ptr), (count))) != cudaSuccess) return bail(e, #ptr)
    DA(S.status, A); DA(S.expday, A); DA(S.texp, A); DA(S.entry, A); DA(S.dur, A); DA(S.sickfrom, A); DA(S.isovia, A);
    DA(S.latday, A); DA(S.tracestart, A); DA(S.traceend, A); DA(S.txpday, A); DA(S.ct_head, A);
    DA(S.mark_pre, W); DA(S.mark_post, W); DA(S.infmask, W); DA(S.ct_cnt, (size_t)h->R);
    DA(S.col_g, A); DA(S.ev_cnt, (size_t)h->R * 2);
    DA(S.grp_idx, A); DA(S.grp_off, (size_t)h->R * 8); DA(S.hits, A);
    if (h->N <= 65535) DA(S.grp16, A);
    DA(S.key, (size_t)h->R); DA(S.pset_of, (size_t)h->R); DA(S.err, (size_t)h->R); DA(S.totiso, (size_t)h->R);
    DA(S.totalmet, (size_t)h->R); DA(S.totalinf, (size_t)h->R);
    DA(S.inc, C); DA(S.prev, C); DA(S.infectors, (size_t)h->R * 4); DA(S.latsum, (size_t)h->R); DA(S.dbg, (size_t)h->R * 32); DA(S.resume_list, (size_t)h->R); DA(S.resume_aux, (size_t)h->R * 2); DA(S.copy_list, (size_t)h->R);
    if (S.store_hm) DA(S.hm, (size_t)h->R * h->T * h->N);
    DA(h->d_scratch, (size_t)4 * h->N);
    DA(h->d_work, 16);
    { cudaDeviceProp pr; if (cudaGetDeviceProperties(&pr, device) == cudaSuccess) h->n_sm = pr.multiProcessorCount; }
    // shared tables
    {
        const Tables &t = tables();
        std::vector<uint32_t> nb(5 * 256, 0xFFFFFFFFu);
        std::vector<uint8_t> guide(5 * 512, 0);     // per group: 256 entries on the top byte of w, 256 on the next byte for the top bucket
        std::vector<unsigned long long> gpw(5 * 256, 0);
        for (int g = 0; g < 5; ++g) {
            const auto &thr = t.nb[g].thr;
            if (thr.size() > 255) return bail(cudaErrorInvalidValue, "negative-binomial table too long");
            for (size_t k = 0; k < thr.size(); ++k) nb[g * 256 + k] = thr[k];
            for (int b = 0; b < 256; ++b) {
                uint32_t wmin = (uint32_t)b << 24; int k = 0;
                while (k < (int)thr.size() && thr[k] <= wmin) ++k;
                guide[g * 512 + b] = (uint8_t)k;
            }
            for (int b = 0; b < 256; ++b) {     // the top bucket holds the whole tail (~100 thresholds): second level
                uint32_t wmin = 0xFF000000u | ((uint32_t)b << 16); int k = 0;
                while (k < (int)thr.size() && thr[k] <= wmin) ++k;
                guide[g * 512 + 256 + b] = (uint8_t)k;
            }
            for (int cnt = 0; cnt < 256; ++cnt) {
                int o[5]; gpw_split(g + 1, cnt, o);
                unsigned long long v = 0;
                for (int q = 0; q < 5; ++q) v |= (unsigned long long)(o[q] & 0xFF) << (8 * q);
                gpw[g * 256 + cnt] = v;
            }
        }
        uint32_t *d_nb; uint8_t *d_guide; unsigned long long *d_gpw;
        DA(d_nb, nb.size()); DA(d_guide, guide.size()); DA(d_gpw, gpw.size());
        if ((e = cudaMemcpy(d_nb, nb.data(), nb.size() * 4, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "nb table upload");
        if ((e = cudaMemcpy(d_guide, guide.data(), guide.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "nb guide upload");
        if ((e = cudaMemcpy(d_gpw, gpw.data(), gpw.size() * 8, cudaMemcpyHostToDevice)) != cudaSuccess) return bail(e, "gpw table upload");
        S.nb_thr = d_nb; S.nb_guide = d_guide; S.gpw = d_gpw;
        ConstTables c = make_const_tables();
        upload_const_tables(c, h->stream);
        if ((e = cudaStreamSynchronize(h->stream)) != cudaSuccess) return bail(e, "table upload");
    }
#undef DA
    *out = h;
    return CABM_OK;
}

extern "C" void cabm_destroy(cabm_handle *h) {
    if (!h) return;
    cudaSetDevice(h->device);
    if (h->stream) cudaStreamSynchronize(h->stream);
    for (void *p : h->allocs) cudaFree(p);
    if (h->d_psets) cudaFree(h->d_psets);
    if (h->d_beta) cudaFree(h->d_beta);
    if (h->S.ct_log) cudaFree(h->S.ct_log);
    if (h->S.ident_list) cudaFree(h->S.ident_list);
    if (h->S.ident_cnt) cudaFree(h->S.ident_cnt);
    if (h->d_stage) cudaFree(h->d_stage);
    if (h->d_mean) cudaFree(h->d_mean);
    if (h->d_tape_keys) cudaFree(h->d_tape_keys);
    if (h->d_tape_w) cudaFree(h->d_tape_w);
    for (cudaEvent_t e : h->pev) cudaEventDestroy(e);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->ev_mark) cudaEventDestroy(h->ev_mark);
    if (h->graph) cudaGraphExecDestroy(h->graph);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

// td_seasonality :204-215
static void td_seasonality(int T, std::vector<double> &out) {
    out.resize(T);
    double mn = 0, mx = 0;
    for (int t = 1; t <= T; ++t) {
        double v = 6.261 + -11.81 * std::cos((80 - t) * 0.022) + 1.817 * std::sin((80 - t) * 0.022);
        out[t - 1] = v;
        if (t == 1 || v < mn) mn = v;
        if (t == 1 || v > mx) mx = v;
    }
    for (int t = 0; t < T; ++t) out[t] = (out[t] - 2.5 * mn) / (mx - mn);
}

extern "C" int cabm_configure(cabm_handle *h, const cabm_params *psets, int n_psets, const int32_t *pset_of_rep,
                              int n_rep, const uint64_t *seeds, int mode) {
    if (!h) return CABM_E_ARG;
    if (!psets || n_psets < 1 || n_rep < 1 || n_rep > h->R || !seeds) return fail(h, CABM_E_ARG, "cabm_configure: bad argument");
    if (mode != CABM_MODE_EXACT && mode != CABM_MODE_NATIVE) return fail(h, CABM_E_ARG, "cabm_configure: unknown mode");
    CU(cudaSetDevice(h->device));
    std::vector<DevPset> dp(n_psets);
    std::vector<unsigned long long> beta((size_t)n_psets * h->T * 4);
    int any_ct = 0, all_ct = 1;
    for (int q = 0; q < n_psets; ++q) {
        const cabm_params &p = psets[q];
        if (p.popsize == 0) return fail(h, CABM_E_PARAM, "no population size given");                       // :164
        if (!(p.ctstrat == 0 || p.ctstrat == 1 || p.ctstrat == 3)) return fail(h, CABM_E_PARAM, "invalid ct strategy");   // :165
        if (p.prov_id < 0 || p.prov_id > 14) return fail(h, CABM_E_PARAM, "shame for not knowing your canadian provinces and territories");  // :334
        if (p.popsize != h->N || p.modeltime != h->T) return fail(h, CABM_E_PARAM, "popsize/modeltime differ from the handle's");
        if (p.tau_mild < -100 || p.tau_mild > 100 || p.cdaysback < 0 || p.strat3qdays < 0 || p.initialinf < 0 || p.initialhi < 0)
            return fail(h, CABM_E_PARAM, "parameter out of range");
        DevPset &d = dp[q];
        memset(&d, 0, sizeof d);
        for (int g = 0; g < 5; ++g) { d.asymp_thr[g] = bernoulli_threshold(p.asymp_props[g]); d.mild_thr[g] = bernoulli_threshold(p.mild_props[g]); }
        double cp = 0;                                           // Categorical CDF walk :177: u >= cdf_k  <=>  w >= ceil(cdf_k 2^32)
        for (int k = 0; k < 4; ++k) { cp = (k == 0) ? PROV_AG[p.prov_id][0] : cp + PROV_AG[p.prov_id][k]; d.age_thr[k] = bernoulli_threshold(cp); }
        d.fmild_thr = bernoulli_threshold(p.fmild); d.fsevere_thr = bernoulli_threshold(p.fsevere);
        d.fpreiso_thr = bernoulli_threshold(p.fpreiso); d.quar_thr = bernoulli_threshold(p.quar_grp_frac);
        d.fctcap_thr = bernoulli_threshold((double)p.fctcapture); d.fcontact_thr = bernoulli_threshold((double)p.fcontactst);
        d.tau_mild = p.tau_mild; d.tpreiso = p.tpreiso; d.quar_grp0 = p.quar_grp - 1; d.ctstrat = p.ctstrat;
        d.cdaysback = p.cdaysback; d.strat3qdays = p.strat3qdays; d.calibration = p.calibration ? 1 : 0;
        d.initialinf = p.initialinf; d.initialhi = p.initialhi;
        d.qdays = p.ctstrat == 1 ? 14 : p.ctstrat == 3 ? p.strat3qdays : 0;                                  // :650-655
        any_ct |= p.ctstrat != 0;
        all_ct &= p.ctstrat != 0;
        // init_betas :191-202 and _get_betavalue :756-769
        std::vector<double> seas;
        if (p.seasonal) td_seasonality(h->T, seas);
        for (int t = 0; t < h->T; ++t) {
            double bf = p.seasonal ? p.beta * seas[t] : p.beta * 1.0;
            unsigned long long *b = &beta[((size_t)q * h->T + t) * 4];
            b[0] = bernoulli_threshold(bf); b[1] = bernoulli_threshold(bf * p.frelasymp);
            b[2] = bernoulli_threshold(bf * 0.44); b[3] = bernoulli_threshold(bf * 0.89);
        }
    }
    std::vector<int> pof(n_rep, 0);
    for (int r = 0; r < n_rep; ++r) {
        if (pset_of_rep) pof[r] = pset_of_rep[r];
        if (pof[r] < 0 || pof[r] >= n_psets) return fail(h, CABM_E_ARG, "pset_of_rep out of range");
    }
    CU(cudaStreamSynchronize(h->stream));
    if (n_psets > h->psets_cap) {
        if (h->d_psets) cudaFree(h->d_psets);
        if (h->d_beta) cudaFree(h->d_beta);
        h->d_psets = nullptr; h->d_beta = nullptr;
        CU(cudaMalloc((void **)&h->d_psets, sizeof(DevPset) * n_psets));
        CU(cudaMalloc((void **)&h->d_beta, sizeof(unsigned long long) * (size_t)n_psets * h->T * 4));
        h->psets_cap = n_psets;
    }
    CU(cudaMemcpy(h->d_psets, dp.data(), sizeof(DevPset) * n_psets, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_beta, beta.data(), beta.size() * 8, cudaMemcpyHostToDevice));
    std::vector<uint2> keys(n_rep);
    for (int r = 0; r < n_rep; ++r) keys[r] = make_uint2((uint32_t)seeds[r], (uint32_t)(seeds[r] >> 32));
    CU(cudaMemcpy(h->S.key, keys.data(), sizeof(uint2) * n_rep, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->S.pset_of, pof.data(), sizeof(int) * n_rep, cudaMemcpyHostToDevice));
    if (any_ct) {
        // contact log: one chunk [next, count, targets...] per (traced agent, day); a traced agent logs on at most cdaysback days
        // (:683-685, :699) and a day's chunk holds its admitted contacts (negative-binomial budgets, mean 8-17, :855). Sized from the
        // largest cdaysback of the parameter sets: 2 + 30 words per traced day (twice the mean budget), at least 128 words per
        // agent, bounded by a third of the free device memory. A log that still overflows raises CABM_ERR_CT_LOG_FULL, which
        // cabm_sync / cabm_get_* report as CABM_E_MODEL — never a silently truncated contact set.
        int maxback = 0;
        for (int q = 0; q < n_psets; ++q) if (psets[q].ctstrat != 0 && psets[q].cdaysback > maxback) maxback = psets[q].cdaysback;
        size_t per_agent = (size_t)32 * (size_t)(maxback > 0 ? maxback : 1);
        if (per_agent < 128) per_agent = 128;
        if (!h->S.ct_log || per_agent > h->ct_words_per_agent) {
            size_t free_b = 0, total_b = 0;
            CU(cudaMemGetInfo(&free_b, &total_b));
            if (h->S.ct_log) { free_b += sizeof(uint32_t) * (size_t)h->S.ct_cap * h->R; CU(cudaFree(h->S.ct_log)); h->S.ct_log = nullptr; h->S.ct_cap = 0; }
            size_t cap = per_agent * (size_t)h->N;          // 32-bit words per replicate
            const size_t budget = free_b / 3 / sizeof(uint32_t) / (size_t)h->R;
            if (cap > budget) cap = budget;
            if (cap > 0xFFFFFFF0ull) cap = 0xFFFFFFF0ull;
            if (cap < (size_t)128 * h->N && cap < budget) cap = (size_t)128 * h->N;
            CU(cudaMalloc((void **)&h->S.ct_log, sizeof(uint32_t) * cap * h->R));
            h->S.ct_cap = (uint32_t)cap; h->ct_words_per_agent = per_agent;
        }
        if (!h->S.ident_list) {
            CU(cudaMalloc((void **)&h->S.ident_list, sizeof(uint32_t) * 2 * (size_t)h->R * h->Np));
            CU(cudaMalloc((void **)&h->S.ident_cnt, sizeof(uint32_t) * 2 * (size_t)h->R));
        }
    }
    h->S.psets = h->d_psets; h->S.beta_thr = h->d_beta; h->S.P = n_psets;
    h->n_rep = n_rep; h->n_psets = n_psets; h->mode = mode; h->any_ct = any_ct; h->all_ct = all_ct; h->configured = 1; h->day = 1; h->ran = false;
    h->S.R = n_rep;
    return CABM_OK;
}

static int need_config(cabm_handle *h) {
    if (!h) return CABM_E_ARG;
    if (!h->configured) return fail(h, CABM_E_STATE, "cabm_configure has not been called");
    cudaError_t e = cudaSetDevice(h->device);
    if (e != cudaSuccess) return fail(h, CABM_E_CUDA, cudaGetErrorString(e));
    return CABM_OK;
}

// After a run has completed (stream synchronised): any replicate whose device error word is set fails the call with the
// reference's error("...") text, like the Julia exception would have ended that replicate (:308, :371-372, :418, :719).
static int model_check(cabm_handle *h) {
    if (!h->strict || !h->ran) return CABM_OK;
    if (h->err_checked) return h->err_rc ? fail(h, h->err_rc, h->err) : CABM_OK;
    std::vector<uint32_t> err((size_t)h->n_rep);
    CU(cudaMemcpy(err.data(), h->S.err, 4 * (size_t)h->n_rep, cudaMemcpyDeviceToHost));
    h->err_checked = true; h->err_rc = CABM_OK;
    for (int r = 0; r < h->n_rep; ++r) {
        const uint32_t e = err[r];
        if (!e) continue;
        const char *msg = (e & ERR_NO_SWAP) ? "swap expired, but no swap set." :
                          (e & ERR_NO_SUS) ? "No susceptibles in the population to change status" :
                          (e & ERR_TRACED_NOT_ISO) ? "a traced person should always be isolated until the trace is killed" :
                          (e & ERR_SICKFROM) ? "sickfrom not set right" :
                          (e & ERR_EMPTY_GROUP) ? "an empty age group was sampled (rand(grps[i], g) on an empty collection)" :
                          (e & ERR_CT_LOG_FULL) ? "contact log capacity exceeded: the traced contact set would be truncated" :
                          "an implementation limit of a kernel was exceeded";
        int nbad = 0;
        for (int q = r; q < h->n_rep; ++q) nbad += err[q] != 0;
        h->err_rc = CABM_E_MODEL;
        return fail(h, CABM_E_MODEL, std::string(msg) + " (replicate " + std::to_string(r) + ", error word " + std::to_string(e) + "; " + std::to_string(nbad) + " replicate(s) flagged)");
    }
    return CABM_OK;
}

static int reset_state(cabm_handle *h, bool curves_too = true) {
    Dev &S = h->S;
    const size_t W = (size_t)S.R * S.nwords, C = (size_t)S.R * 6 * S.T * 12;
    CU(cudaMemsetAsync(S.mark_pre, 0, W * 4, h->stream)); CU(cudaMemsetAsync(S.mark_post, 0, W * 4, h->stream));
    CU(cudaMemsetAsync(S.infmask, 0, W * 4, h->stream)); CU(cudaMemsetAsync(S.ct_cnt, 0, (size_t)S.R * 4, h->stream));
    if (S.ident_cnt) CU(cudaMemsetAsync(S.ident_cnt, 0, (size_t)S.R * 8, h->stream));
    h->poked = false;
    CU(cudaMemsetAsync(S.err, 0, (size_t)S.R * 4, h->stream)); CU(cudaMemsetAsync(S.totiso, 0, (size_t)S.R * 8, h->stream));
    CU(cudaMemsetAsync(S.totalmet, 0, (size_t)S.R * 4, h->stream)); CU(cudaMemsetAsync(S.totalinf, 0, (size_t)S.R * 4, h->stream));
    CU(cudaMemsetAsync(S.ev_cnt, 0, (size_t)S.R * 8, h->stream));
    if (curves_too) { CU(cudaMemsetAsync(S.inc, 0, C * 4, h->stream)); CU(cudaMemsetAsync(S.prev, 0, C * 4, h->stream)); }
    CU(cudaMemsetAsync(S.infectors, 0, (size_t)S.R * 32, h->stream)); CU(cudaMemsetAsync(S.latsum, 0, (size_t)S.R * 8, h->stream));
    h->day = 1;
    return CABM_OK;
}

// wraps one launcher call; in profiling mode brackets it with events on the handle's stream
template <class F> static void timed(cabm_handle *h, int cls, F f) {
    if (!h->profiling) { h->launches += f(); return; }
    while (h->pev.size() < h->pev_used + 2) { cudaEvent_t e; cudaEventCreate(&e); h->pev.push_back(e); }
    cudaEventRecord(h->pev[h->pev_used], h->stream);
    h->launches += f();
    cudaEventRecord(h->pev[h->pev_used + 1], h->stream);
    h->pev_used += 2; h->pcls.push_back(cls);
}

// the per-day kernels' event list [R][Np] is only needed on that path: allocated on first use
static int need_event_list(cabm_handle *h) {
    if (h->S.ev_list) return CABM_OK;
    CU(cudaMalloc((void **)&h->S.ev_list, sizeof(uint32_t) * (size_t)h->R * h->Np));
    h->allocs.push_back(h->S.ev_list);
    return CABM_OK;
}

// native schedule: list of the day's events that hit a later infector (second-order budget count walks it instead of re-drawing every
// event). Allocated on first use; without the memory for it the schedule falls back to three full sweeps.
static void need_native_edges(cabm_handle *h) {
    if (h->S.nat_edges || h->native_edges_tried) return;
    h->native_edges_tried = true;
    const char *off = getenv("CABM_NATIVE_EDGES");
    if (off && off[0] == '0' && off[1] == 0) return;                 // CABM_NATIVE_EDGES=0: the three-sweep form (for comparison)
    uint32_t *e = nullptr, *c = nullptr, *v = nullptr;
    const size_t words = (size_t)h->R * h->Np * 3;
    if (cudaMalloc((void **)&e, words * 4) != cudaSuccess) { cudaGetLastError(); return; }
    if (cudaMalloc((void **)&c, (size_t)h->R * 4) != cudaSuccess) { cudaGetLastError(); cudaFree(e); return; }
    const char *ver = getenv("CABM_NATIVE_VERIFY");
    if (ver && ver[0] == '1' && ver[1] == 0) {
        const size_t vb = (size_t)h->R * h->Np / 2 * 4;
        if (cudaMalloc((void **)&v, vb) == cudaSuccess) { cudaMemsetAsync(v, 0, vb, h->stream); h->allocs.push_back(v); } else { cudaGetLastError(); v = nullptr; }
    }
    cudaMemsetAsync(h->S.hits, 0, (size_t)h->R * h->Np * 4, h->stream);     // from here on k_native_clean keeps the counters at zero between days
    h->allocs.push_back(e); h->allocs.push_back(c);
    h->S.nat_edges = e; h->S.nat_ecnt = c; h->S.nat_verify = v;
}

static int do_dyntrans(cabm_handle *h, int day) {
    if (h->mode == CABM_MODE_NATIVE) need_native_edges(h);
    if (h->mode == CABM_MODE_NATIVE) timed(h, CABM_K_DYNTRANS, [&] { return launch_dyntrans_native(h->S, day, h->stream); });
    else timed(h, CABM_K_DYNTRANS, [&] { return launch_dyntrans_exact(h->S, day, h->stream); });
    return CABM_OK;
}
static int do_time_update(cabm_handle *h, int day) {
    if (h->any_ct) timed(h, CABM_K_CT_IDENTIFY, [&] { return launch_ct_identify(h->S, day, h->poked, h->stream); });
    timed(h, CABM_K_TIME_UPDATE, [&] { return launch_time_update_scan(h->S, day, h->any_ct != 0, h->stream); });
    timed(h, CABM_K_TU_EVENTS, [&] { return launch_time_update_events(h->S, day, h->stream); });
    return CABM_OK;
}

extern "C" int cabm_run(cabm_handle *h) {
    int rc = need_config(h); if (rc) return rc;
    const Dev &S = h->S;
    h->pev_used = 0; h->pcls.clear();
    // The fused kernel keeps a replicate's hot state in shared memory: usable when it fits two CTAs per SM budget
    // (popsize <= ~21k agents) with the exact schedule; otherwise one kernel per phase per day.
    // (shared-memory variant: popsize <= ~21k agents; otherwise the same kernel with the per-agent arrays in global memory, any popsize)
    const bool exact = h->mode == CABM_MODE_EXACT;
    if (S.tape_n && h->path != CABM_PATH_AUTO && h->path != CABM_PATH_MULTIKERNEL) return fail(h, CABM_E_STATE, "a replay tape is consulted by the per-day kernels only (CABM_PATH_AUTO or CABM_PATH_MULTIKERNEL)");
    const bool fused_ok = !S.tape_n && exact && h->N <= 65535 && fused_smem_bytes_host(h->Np, false) <= 227 * 1024 - 512;
    if (h->path == CABM_PATH_FUSED && !fused_ok) return fail(h, CABM_E_STATE, "fused path needs the exact schedule and a population that fits shared memory");
    if (h->path == CABM_PATH_FUSED_GLOBAL && !exact) return fail(h, CABM_E_STATE, "fused path needs the exact schedule");
    const bool use_smem = h->path == CABM_PATH_FUSED || (h->path == CABM_PATH_AUTO && fused_ok);
    const bool use_global = h->path == CABM_PATH_FUSED_GLOBAL || (h->path == CABM_PATH_AUTO && exact && !fused_ok && !S.tape_n);
    const bool use_fused = use_smem || use_global;
    CU(cudaEventRecord(h->ev0, h->stream));
    if ((rc = reset_state(h, /*curves_too=*/!use_fused))) return rc;      // the fused kernel writes every curve entry itself
    h->last_path = use_smem ? CABM_PATH_FUSED : use_global ? CABM_PATH_FUSED_GLOBAL : CABM_PATH_MULTIKERNEL;
    if (!use_smem && (rc = need_event_list(h))) return rc;        // per-day kernels, and the fused kernel with its state in global memory
    if (use_fused) {
        timed(h, CABM_K_FUSED, [&] { return launch_sim_fused(S, h->d_work, h->n_sm, h->any_ct != 0, use_global, h->stream); });  // :104-142 in one launch
    } else {
        auto enqueue = [&] {
            timed(h, CABM_K_OTHER, [&] { return launch_initialize(S, h->stream); });                                   // :112
            timed(h, CABM_K_OTHER, [&] { return launch_insert_infected(S, 0, 0, 0, 0, 0, 0, h->stream); });            // :117/:119
            timed(h, CABM_K_OTHER, [&] { return launch_insert_infected(S, 1, 0, 0, 0, 1, 0, h->stream); });            // :120
            timed(h, CABM_K_OTHER, [&] { return launch_build_groups(S, h->stream); });                                 // :128
            timed(h, CABM_K_OTHER, [&] { return launch_first_column(S, h->stream); });                                 // :136 for st = 1
            for (int st = 1; st <= h->T; ++st) {                                              // :131
                do_dyntrans(h, st);                                                           // :137
                do_time_update(h, st);                                                        // :138 (+ snapshot of st+1)
            }
            timed(h, CABM_K_OTHER, [&] { return launch_curves_finalize(S, h->stream); });                              // :259-293
            timed(h, CABM_K_OTHER, [&] { return launch_count_infectors(S, h->stream); });                              // :295-313
        };
        // The day loop is a fixed sequence of launches on one stream: the third run with the same arguments captures it into a
        // CUDA graph, later ones replay the graph (one submission instead of ~3 per day). Profiling runs launch directly.
        uint64_t key = 1469598103934665603ull;
        { const unsigned char *b = reinterpret_cast<const unsigned char *>(&S); for (size_t k = 0; k < sizeof(Dev); ++k) key = (key ^ b[k]) * 1099511628211ull; }
        key = (key ^ (uint64_t)(h->mode * 8 + h->any_ct * 4 + h->all_ct * 2 + (h->poked ? 1 : 0))) * 1099511628211ull;
        bool replayed = false;
        if (!h->profiling) {
            if (h->graph && h->graph_key == key) {
                if (cudaGraphLaunch(h->graph, h->stream) == cudaSuccess) { h->launches += h->graph_launches; replayed = true; }
            } else if (h->seen_key == key && h->seen_count >= 2) {      // worth the capture (a few ms on the host): a handle that keeps running this
                if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; }
                cudaGraph_t g = nullptr;
                const int64_t l0 = h->launches;
                if (cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
                    enqueue();
                    const cudaError_t ec = cudaStreamEndCapture(h->stream, &g);
                    h->graph_launches = h->launches - l0; h->launches = l0;
                    if (ec == cudaSuccess && g && cudaGraphInstantiate(&h->graph, g, 0) == cudaSuccess) {
                        h->graph_key = key;
                        if (cudaGraphLaunch(h->graph, h->stream) == cudaSuccess) { h->launches += h->graph_launches; replayed = true; }
                    } else h->graph = nullptr;
                    if (g) cudaGraphDestroy(g);
                }
                cudaGetLastError();         // a failed capture leaves an error behind: fall back to direct launches below
            }
            h->seen_count = h->seen_key == key ? h->seen_count + 1 : 1;
            h->seen_key = key;
        }
        if (!replayed) enqueue();
    }
    CU(cudaEventRecord(h->ev1, h->stream));
    CU(cudaGetLastError());
    h->day = h->T + 1; h->ran = true; h->err_checked = false; h->err_rc = CABM_OK;
    if (h->profiling) {
        CU(cudaStreamSynchronize(h->stream));
        for (int c = 0; c < CABM_N_KCLASS; ++c) { h->prof_ms[c] = 0; h->prof_n[c] = 0; }
        for (size_t k = 0; k < h->pcls.size(); ++k) {
            float ms = 0; cudaEventElapsedTime(&ms, h->pev[2 * k], h->pev[2 * k + 1]);
            h->prof_ms[h->pcls[k]] += ms; h->prof_n[h->pcls[k]] += 1;
        }
    }
    return CABM_OK;
}

extern "C" int cabm_set_path(cabm_handle *h, int path) {
    if (!h || path < CABM_PATH_AUTO || path > CABM_PATH_FUSED_GLOBAL) return CABM_E_ARG;
    h->path = path; return CABM_OK;
}
extern "C" int cabm_get_last_path(const cabm_handle *h) { return h ? h->last_path : -1; }
extern "C" int cabm_set_profiling(cabm_handle *h, int on) { if (!h) return CABM_E_ARG; h->profiling = on ? 1 : 0; return CABM_OK; }
extern "C" int cabm_get_profile(cabm_handle *h, float *ms, int32_t *launches) {
    if (!h || !ms || !launches) return CABM_E_ARG;
    for (int c = 0; c < CABM_N_KCLASS; ++c) { ms[c] = h->prof_ms[c]; launches[c] = h->prof_n[c]; }
    return CABM_OK;
}

extern "C" int cabm_sync(cabm_handle *h) {
    if (!h) return CABM_E_ARG;
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaGetLastError());
    return model_check(h);
}
extern "C" int cabm_set_strict(cabm_handle *h, int on) { if (!h) return CABM_E_ARG; h->strict = on ? 1 : 0; h->err_checked = false; return CABM_OK; }

// Replay tape: draws whose (seed, counter) key is listed take the listed words instead of the Philox output.
extern "C" int cabm_set_tape(cabm_handle *h, const cabm_tape_entry *entries, uint64_t n) {
    if (!h) return CABM_E_ARG;
    if (n && !entries) return fail(h, CABM_E_ARG, "cabm_set_tape: bad argument");
    if (n > 0xFFFFFFF0ull) return fail(h, CABM_E_ARG, "cabm_set_tape: too many entries");
    CU(cudaSetDevice(h->device));
    CU(cudaStreamSynchronize(h->stream));
    if (h->d_tape_keys) { cudaFree(h->d_tape_keys); h->d_tape_keys = nullptr; }
    if (h->d_tape_w) { cudaFree(h->d_tape_w); h->d_tape_w = nullptr; }
    h->S.tape_keys = nullptr; h->S.tape_w = nullptr; h->S.tape_n = 0;
    if (n == 0) return CABM_OK;
    std::vector<uint64_t> order(n);
    for (uint64_t i = 0; i < n; ++i) order[i] = i;
    auto less = [&](uint64_t a, uint64_t b) {
        const cabm_tape_entry &x = entries[a], &y = entries[b];
        if (x.seed != y.seed) return x.seed < y.seed;
        for (int k = 0; k < 4; ++k) if (x.ctr[k] != y.ctr[k]) return x.ctr[k] < y.ctr[k];
        return false;
    };
    std::sort(order.begin(), order.end(), less);
    std::vector<uint32_t> keys(n * 6); std::vector<uint4> words(n);
    for (uint64_t i = 0; i < n; ++i) {
        const cabm_tape_entry &e = entries[order[i]];
        if (i && !less(order[i - 1], order[i])) return fail(h, CABM_E_ARG, "cabm_set_tape: duplicate key");
        keys[i * 6 + 0] = (uint32_t)e.seed; keys[i * 6 + 1] = (uint32_t)(e.seed >> 32);
        for (int k = 0; k < 4; ++k) keys[i * 6 + 2 + k] = e.ctr[k];
        words[i] = make_uint4(e.w[0], e.w[1], e.w[2], e.w[3]);
    }
    CU(cudaMalloc((void **)&h->d_tape_keys, keys.size() * 4));
    CU(cudaMalloc((void **)&h->d_tape_w, words.size() * sizeof(uint4)));
    CU(cudaMemcpy(h->d_tape_keys, keys.data(), keys.size() * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(h->d_tape_w, words.data(), words.size() * sizeof(uint4), cudaMemcpyHostToDevice));
    h->S.tape_keys = h->d_tape_keys; h->S.tape_w = h->d_tape_w; h->S.tape_n = (unsigned)n;
    if (h->graph) { cudaGraphExecDestroy(h->graph); h->graph = nullptr; h->graph_key = 0; }
    return CABM_OK;
}

extern "C" int cabm_mark(cabm_handle *h) {
    if (!h) return CABM_E_ARG;
    cudaSetDevice(h->device);
    CU(cudaEventRecord(h->ev_mark, h->stream));
    return CABM_OK;
}

extern "C" int cabm_ms_since_mark(cabm_handle *marked, cabm_handle *h, float *ms) {
    if (!marked || !h || !ms) return CABM_E_ARG;
    cudaSetDevice(h->device);
    CU(cudaEventSynchronize(h->ev1));
    CU(cudaEventElapsedTime(ms, marked->ev_mark, h->ev1));
    return CABM_OK;
}

extern "C" int cabm_last_run_ms(cabm_handle *h, float *ms) {
    if (!h || !ms) return CABM_E_ARG;
    if (!h->ran) return fail(h, CABM_E_STATE, "no cabm_run yet");
    CU(cudaEventSynchronize(h->ev1));
    CU(cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return CABM_OK;
}
extern "C" int64_t cabm_launch_count(const cabm_handle *h) { return h ? h->launches : 0; }
extern "C" int cabm_get_debug(cabm_handle *h, int64_t *out) {
    if (!h || !out) return CABM_E_ARG;
    CU(cudaStreamSynchronize(h->stream));
    CU(cudaMemcpy(out, h->S.dbg, (size_t)h->n_rep * 256, cudaMemcpyDeviceToHost));
    return CABM_OK;
}
extern "C" int cabm_get_day(const cabm_handle *h) { return h ? h->day : -1; }

extern "C" int cabm_step_initialize(cabm_handle *h) {
    int rc = need_config(h); if (rc) return rc;
    if ((rc = need_event_list(h))) return rc;
    if ((rc = reset_state(h))) return rc;
    h->launches += launch_initialize(h->S, h->stream);
    h->launches += launch_build_groups(h->S, h->stream);
    CU(cudaGetLastError());
    return CABM_OK;
}
extern "C" int cabm_step_insert_infected(cabm_handle *h, int health, int num, int ag, int call_id) {
    int rc = need_config(h); if (rc) return rc;
    if (!(health == PRE || health == LAT || health == INF || health == REC)) return fail(h, CABM_E_ARG, "can not insert human of this health");  // :389
    if (num < 0 || ag < 1 || ag > 5) return fail(h, CABM_E_ARG, "cabm_step_insert_infected: bad argument");
    h->launches += launch_insert_infected(h->S, 2, health, num, ag, call_id, h->day - 1, h->stream);
    CU(cudaStreamSynchronize(h->stream));
    std::vector<uint32_t> err(h->n_rep);
    CU(cudaMemcpy(err.data(), h->S.err, 4 * (size_t)h->n_rep, cudaMemcpyDeviceToHost));
    for (uint32_t e : err) if (e & ERR_NO_SUS) return fail(h, CABM_E_MODEL, "No susceptibles in the population to change status");     // :371-372
    return CABM_OK;
}
extern "C" int cabm_step_dyntrans(cabm_handle *h) {
    int rc = need_config(h); if (rc) return rc;
    if (h->day > h->T) return fail(h, CABM_E_STATE, "day exceeds modeltime");
    do_dyntrans(h, h->day);
    CU(cudaGetLastError());
    return CABM_OK;
}
extern "C" int cabm_step_time_update(cabm_handle *h) {
    int rc = need_config(h); if (rc) return rc;
    if ((rc = need_event_list(h))) return rc;
    if (h->day > h->T) return fail(h, CABM_E_STATE, "day exceeds modeltime");
    do_time_update(h, h->day);
    h->day += 1;
    CU(cudaStreamSynchronize(h->stream));
    std::vector<uint32_t> err(h->n_rep);
    CU(cudaMemcpy(err.data(), h->S.err, 4 * (size_t)h->n_rep, cudaMemcpyDeviceToHost));
    for (uint32_t e : err) if (e & ERR_NO_SWAP) return fail(h, CABM_E_MODEL, "swap expired, but no swap set.");                          // :418
    return CABM_OK;
}
extern "C" int cabm_get_dyntrans_counters(cabm_handle *h, int32_t *totalmet, int32_t *totalinf) {
    int rc = need_config(h); if (rc) return rc;
    CU(cudaStreamSynchronize(h->stream));
    if (totalmet) CU(cudaMemcpy(totalmet, h->S.totalmet, 4 * (size_t)h->n_rep, cudaMemcpyDeviceToHost));
    if (totalinf) CU(cudaMemcpy(totalinf, h->S.totalinf, 4 * (size_t)h->n_rep, cudaMemcpyDeviceToHost));
    return CABM_OK;
}

// hmatrix is kept as one byte per element on the device and leaves as Int16 (:111): widen through a staging buffer
static int copy_hmatrix(cabm_handle *h, size_t first_elem, size_t n_elem, int16_t *out) {
    const size_t chunk = (size_t)32 << 20;
    if (!h->d_stage || h->stage_elems < (size_t)h->T * h->N) {
        if (h->d_stage) cudaFree(h->d_stage);
        h->d_stage = nullptr;
        h->stage_elems = n_elem < chunk ? n_elem : chunk;
        if (h->stage_elems < (size_t)h->T * h->N) h->stage_elems = (size_t)h->T * h->N;
        CU(cudaMalloc((void **)&h->d_stage, h->stage_elems * 2));
    }
    for (size_t o = 0; o < n_elem; o += h->stage_elems) {
        const size_t n = n_elem - o < h->stage_elems ? n_elem - o : h->stage_elems;
        h->launches += launch_widen(h->S.hm + first_elem + o, h->d_stage, n, h->stream);
        CU(cudaMemcpyAsync(out + o, h->d_stage, n * 2, cudaMemcpyDeviceToHost, h->stream));
        CU(cudaStreamSynchronize(h->stream));
    }
    return CABM_OK;
}
extern "C" int cabm_get_hmatrix(cabm_handle *h, int replicate, int16_t *out) {
    int rc = need_config(h); if (rc) return rc;
    if (!h->S.store_hm) return fail(h, CABM_E_STATE, "handle was created without CABM_STORE_HMATRIX");
    if (replicate < 0 || replicate >= h->n_rep || !out) return fail(h, CABM_E_ARG, "cabm_get_hmatrix: bad argument");
    const size_t n = (size_t)h->T * h->N;
    if ((rc = copy_hmatrix(h, (size_t)replicate * n, n, out))) return rc;
    return model_check(h);
}
extern "C" int cabm_get_hmatrix_all(cabm_handle *h, int16_t *out) {
    int rc = need_config(h); if (rc) return rc;
    if (!h->S.store_hm) return fail(h, CABM_E_STATE, "handle was created without CABM_STORE_HMATRIX");
    if (!out) return fail(h, CABM_E_ARG, "cabm_get_hmatrix_all: bad argument");
    if ((rc = copy_hmatrix(h, 0, (size_t)h->n_rep * h->T * h->N, out))) return rc;
    return model_check(h);
}
extern "C" int cabm_get_hmatrix_all_u8(cabm_handle *h, uint8_t *out) {
    int rc = need_config(h); if (rc) return rc;
    if (!h->S.store_hm) return fail(h, CABM_E_STATE, "handle was created without CABM_STORE_HMATRIX");
    if (!out) return fail(h, CABM_E_ARG, "cabm_get_hmatrix_all_u8: bad argument");
    CU(cudaMemcpyAsync(out, h->S.hm, (size_t)h->n_rep * h->T * h->N, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return model_check(h);
}
extern "C" int cabm_get_curves(cabm_handle *h, int32_t *inc, int32_t *prev) {
    int rc = need_config(h); if (rc) return rc;
    if (!h->ran) return fail(h, CABM_E_STATE, "curves exist after cabm_run");
    const size_t bytes = (size_t)h->n_rep * 6 * h->T * 12 * 4;
    if (inc) CU(cudaMemcpyAsync(inc, h->S.inc, bytes, cudaMemcpyDeviceToHost, h->stream));
    if (prev) CU(cudaMemcpyAsync(prev, h->S.prev, bytes, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return model_check(h);
}
extern "C" int cabm_set_host_curves_i16(cabm_handle *h, int16_t *inc, int16_t *prev) {
    if (!h) return CABM_E_ARG;
    if ((inc == nullptr) != (prev == nullptr)) return fail(h, CABM_E_ARG, "cabm_set_host_curves_i16: give both buffers or neither");
    if (inc) {
        if (h->N > 32767) return fail(h, CABM_E_ARG, "Int16 curves need popsize <= 32767");
        cudaPointerAttributes pa, pb;
        CU(cudaPointerGetAttributes(&pa, inc)); CU(cudaPointerGetAttributes(&pb, prev));
        if (pa.type != cudaMemoryTypeHost || pb.type != cudaMemoryTypeHost)
            return fail(h, CABM_E_ARG, "cabm_set_host_curves_i16: buffers must be pinned host memory (cabm_host_alloc)");
        CU(cudaStreamSynchronize(h->stream));
        CU(cudaHostGetDevicePointer((void **)&h->S.host_inc, inc, 0)); CU(cudaHostGetDevicePointer((void **)&h->S.host_prev, prev, 0));
        h->host_inc = inc; h->host_prev = prev;
    } else {
        CU(cudaStreamSynchronize(h->stream));
        h->S.host_inc = h->S.host_prev = nullptr; h->host_inc = h->host_prev = nullptr;
    }
    return CABM_OK;
}
extern "C" int cabm_set_host_curve_rows(cabm_handle *h, int32_t *rows) {
    if (!h) return CABM_E_ARG;
    CU(cudaStreamSynchronize(h->stream));
    if (rows) {
        cudaPointerAttributes pa;
        CU(cudaPointerGetAttributes(&pa, rows));
        if (pa.type != cudaMemoryTypeHost) return fail(h, CABM_E_ARG, "cabm_set_host_curve_rows: the array must be pinned host memory (cabm_host_alloc)");
        CU(cudaHostGetDevicePointer((void **)&h->S.host_rows, rows, 0));
    } else h->S.host_rows = nullptr;
    return CABM_OK;
}
extern "C" int cabm_get_curves_i16(cabm_handle *h, int16_t *inc, int16_t *prev) {
    int rc = need_config(h); if (rc) return rc;
    if (!h->ran) return fail(h, CABM_E_STATE, "curves exist after cabm_run");
    if (inc && inc == h->host_inc && prev == h->host_prev && (h->last_path == CABM_PATH_FUSED || h->last_path == CABM_PATH_FUSED_GLOBAL)) {   // streamed there by the kernel already
        CU(cudaStreamSynchronize(h->stream));
        return model_check(h);
    }
    if (h->N > 32767) return fail(h, CABM_E_ARG, "cabm_get_curves_i16 needs popsize <= 32767");
    if (!inc || !prev) return fail(h, CABM_E_ARG, "cabm_get_curves_i16: bad argument");
    const size_t n = (size_t)h->n_rep * 6 * h->T * 12;
    if (h->stage_elems < 2 * n) {
        if (h->d_stage) cudaFree(h->d_stage);
        h->d_stage = nullptr; h->stage_elems = 0;
        CU(cudaMalloc((void **)&h->d_stage, 2 * n * 2));
        h->stage_elems = 2 * n;
    }
    h->launches += launch_narrow(h->S.inc, h->S.prev, h->d_stage, n, h->stream);
    CU(cudaMemcpyAsync(inc, h->d_stage, n * 2, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(prev, h->d_stage + n, n * 2, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return model_check(h);
}
extern "C" int cabm_get_mean_curves(cabm_handle *h, double *inc_mean, double *prev_mean) {
    int rc = need_config(h); if (rc) return rc;
    if (!h->ran) return fail(h, CABM_E_STATE, "curves exist after cabm_run");
    if (!inc_mean || !prev_mean) return fail(h, CABM_E_ARG, "cabm_get_mean_curves: bad argument");
    const size_t per_rep = (size_t)6 * h->T * 12;
    if (h->mean_elems < 2 * per_rep) {                       // scratch kept with the handle
        if (h->d_mean) cudaFree(h->d_mean);
    if (h->d_tape_keys) cudaFree(h->d_tape_keys);
    if (h->d_tape_w) cudaFree(h->d_tape_w);
        h->d_mean = nullptr; h->mean_elems = 0;
        CU(cudaMalloc((void **)&h->d_mean, per_rep * 2 * sizeof(double)));
        h->mean_elems = 2 * per_rep;
    }
    double *d = h->d_mean;
    h->launches += launch_mean_curves(h->S.inc, h->S.prev, d, d + per_rep, h->n_rep, per_rep, h->stream);
    CU(cudaMemcpyAsync(inc_mean, d, per_rep * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaMemcpyAsync(prev_mean, d + per_rep, per_rep * 8, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return model_check(h);
}
extern "C" int cabm_get_scalars(cabm_handle *h, int64_t *infectors, int64_t *totalisolated, int64_t *lat_inc_sum, uint32_t *err) {
    int rc = need_config(h); if (rc) return rc;
    if ((infectors || lat_inc_sum) && !h->ran) return fail(h, CABM_E_STATE, "infectors/lat_inc_sum exist after cabm_run");
    const size_t R = (size_t)h->n_rep;
    if (infectors) CU(cudaMemcpyAsync(infectors, h->S.infectors, R * 32, cudaMemcpyDeviceToHost, h->stream));
    if (totalisolated) CU(cudaMemcpyAsync(totalisolated, h->S.totiso, R * 8, cudaMemcpyDeviceToHost, h->stream));
    if (lat_inc_sum) CU(cudaMemcpyAsync(lat_inc_sum, h->S.latsum, R * 8, cudaMemcpyDeviceToHost, h->stream));
    if (err) CU(cudaMemcpyAsync(err, h->S.err, R * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return err ? CABM_OK : model_check(h);       // a caller that collects the error words itself decides what they mean
}
extern "C" int cabm_get_field(cabm_handle *h, int field, int replicate, int32_t *out) {
    int rc = need_config(h); if (rc) return rc;
    if (field < 0 || field > CABM_F_NCONTACTS || replicate < 0 || replicate >= h->n_rep || !out) return fail(h, CABM_E_ARG, "cabm_get_field: bad argument");
    if (field >= CABM_F_NCONTACTS && !h->S.ct_log) { memset(out, 0, 4 * (size_t)h->N); return CABM_OK; }
    h->launches += launch_peek(h->S, field, replicate, h->day, h->d_scratch, h->stream);
    const size_t n = (size_t)h->N * (field == CABM_F_DUR ? 4 : 1);
    CU(cudaMemcpyAsync(out, h->d_scratch, n * 4, cudaMemcpyDeviceToHost, h->stream));
    CU(cudaStreamSynchronize(h->stream));
    return CABM_OK;
}
extern "C" int cabm_set_field(cabm_handle *h, int field, int replicate, int agent, int32_t value) {
    int rc = need_config(h); if (rc) return rc;
    if (field < 0 || field > CABM_F_TRACEDXP || field == CABM_F_AG || field == CABM_F_DUR || replicate < 0 || replicate >= h->n_rep || agent < 0 || agent >= h->N)
        return fail(h, CABM_E_ARG, "cabm_set_field: bad argument");
    h->poked = true;
    h->launches += launch_poke(h->S, field, replicate, agent, h->day, value, h->stream);
    CU(cudaStreamSynchronize(h->stream));
    return CABM_OK;
}

extern "C" int cabm_reduce_hmatrix(cabm_handle *h, const int16_t *hm, int n_mat, int n, int t, const int32_t *ags,
                                   int32_t *inc, int32_t *prev) {
    if (!h) return CABM_E_ARG;
    if (!hm || !ags || !inc || !prev || n_mat < 1 || n < 1 || t < 1) return fail(h, CABM_E_ARG, "cabm_reduce_hmatrix: bad argument");
    CU(cudaSetDevice(h->device));
    const size_t hb = (size_t)n_mat * n * t * 2, ab = (size_t)n_mat * n * 4, cb = (size_t)n_mat * 6 * t * 12 * 4;
    int16_t *d_hm = nullptr; int *d_ags = nullptr, *d_inc = nullptr, *d_prev = nullptr;
    cudaError_t e = cudaMalloc((void **)&d_hm, hb);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_ags, ab);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_inc, cb);
    if (e == cudaSuccess) e = cudaMalloc((void **)&d_prev, cb);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_hm, hm, hb, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(d_ags, ags, ab, cudaMemcpyHostToDevice, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_inc, 0, cb, h->stream);
    if (e == cudaSuccess) e = cudaMemsetAsync(d_prev, 0, cb, h->stream);
    if (e == cudaSuccess) { h->launches += launch_reduce_hmatrix(d_hm, d_ags, n_mat, n, t, d_inc, d_prev, h->stream); e = cudaGetLastError(); }
    if (e == cudaSuccess) e = cudaMemcpyAsync(inc, d_inc, cb, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(prev, d_prev, cb, cudaMemcpyDeviceToHost, h->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    cudaFree(d_hm); cudaFree(d_ags); cudaFree(d_inc); cudaFree(d_prev);
    if (e != cudaSuccess) return fail(h, CABM_E_CUDA, std::string("cabm_reduce_hmatrix: ") + cudaGetErrorString(e));
    return CABM_OK;
}

extern "C" void *cabm_host_alloc(uint64_t bytes) {
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 16) != cudaSuccess) return nullptr;
    return p;
}
extern "C" void cabm_host_free(void *p) { if (p) cudaFreeHost(p); }

extern "C" int cabm_get_table(const char *name, int32_t *base, uint32_t *thr, int cap) {
    const Table *t = name ? table_by_name(name) : nullptr;
    if (!t) return -1;
    if (base) *base = t->base;
    for (int i = 0; i < (int)t->thr.size() && i < cap && thr; ++i) thr[i] = t->thr[i];
    return (int)t->thr.size();
}
extern "C" void cabm_get_gpw(int ag, int cnts, int32_t out[5]) {
    int o[5]; gpw_split(ag, cnts, o);
    for (int g = 0; g < 5; ++g) out[g] = o[g];
}
