// cabm_kernels.cu — the sm_100a kernels of the simulation path.
// One launch advances every configured replicate; grid.y (or the warp index) is the replicate.
#define CABM_KERNELS_TU
#include "cabm_device.cuh"
#include "cabm_kernels.h"

namespace cabm {

void upload_const_tables(const ConstTables &t, cudaStream_t s) {
    cudaMemcpyToSymbolAsync(c_tab, &t, sizeof t, 0, cudaMemcpyHostToDevice, s);
}

// ============================================================================ initialize :171-188
// thread per (replicate, agent). Draw words: S_INIT block 0 = (ag, lat, pre, asymp), block 1 = (inf, quarantine).
template <bool TP> __global__ void __launch_bounds__(256) k_initialize(Dev S) {
    const int r = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.Np) return;
    const size_t a = (size_t)r * S.Np + i;
    const DevPset &P = S.psets[S.pset_of[r]];
    uint32_t st = st_make(SUS, UNDEF, 0, false), d4 = 0;
    if (i < S.N) {
        const uint2 key = S.key[r];
        uint4 wa = draw4<TP>(S, i, 0, S_INIT, 0, key), wb = draw4<TP>(S, i, 0, S_INIT, 1, key);
        int ag0 = 0;                                                   // :177 rand(Categorical)
#pragma unroll
        for (int k = 0; k < 4; ++k) ag0 += ((unsigned long long)wa.x >= P.age_thr[k]);
        int latents = table_draw(wa.y, c_tab.lat_base, c_tab.lat_thr, c_tab.lat_len);   // :347,352
        int pres = table_draw(wa.z, c_tab.pre_base, c_tab.pre_thr, c_tab.pre_len);      // :348,353
        int asymps = table_draw(wa.w, c_tab.asy_base, c_tab.asy_thr, c_tab.asy_len);    // :349,355
        int infs = table_draw(wb.x, c_tab.inf_base, c_tab.inf_thr, c_tab.inf_len);      // :350,356
        latents -= pres;                                                                // :354
        d4 = (uint32_t)latents | ((uint32_t)asymps << 8) | ((uint32_t)pres << 16) | ((uint32_t)infs << 24);
        bool quar = bern(wb.y, P.quar_thr) && ag0 == P.quar_grp0;                       // :181
        st = st_make(SUS, UNDEF, ag0, quar);
        S.isovia[a] = quar ? VIA_QU : VIA_NULL;                                          // :183
    } else {
        S.isovia[a] = VIA_NULL;
    }
    S.status[a] = st;
    S.dur[a] = d4;
    S.expday[a] = S.texp[a] = (i < S.N) ? (int16_t)999 : (int16_t)32767;                 // :179
    S.entry[a] = 0;
    S.sickfrom[a] = UNDEF;
    S.latday[a] = -999;                                                                   // doi = 999 (:19)
    S.tracestart[a] = -1; S.traceend[a] = -1; S.txpday[a] = 0;
    S.ct_head[a] = CT_NONE;
}

// ============================================================================ get_ag_dist :339-343
// CTA per replicate: ascending agent indices of each age group, concatenated; offsets in grp_off[r][0..5].
__global__ void __launch_bounds__(256) k_build_groups(Dev S) {
    const int r = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    __shared__ int wcnt[8][5];
    __shared__ int base[5];
    __shared__ int total[5];
    const size_t rb = (size_t)r * S.Np;
    if (tid < 5) total[tid] = 0;
    __syncthreads();
    // pass 1: group sizes
    int c[5] = {0, 0, 0, 0, 0};
    for (int i = tid; i < S.N; i += 256) c[st_ag(S.status[rb + i])]++;
#pragma unroll
    for (int g = 0; g < 5; ++g) {
        int v = c[g];
        for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
        if (lane == 0) atomicAdd(&total[g], v);
    }
    __syncthreads();
    if (tid == 0) {
        int o = 0;
        for (int g = 0; g < 5; ++g) { S.grp_off[r * 8 + g] = o; base[g] = o; o += total[g]; }
        S.grp_off[r * 8 + 5] = o;
    }
    __syncthreads();
    // pass 2: ordered scatter, 256 agents per round
    for (int i0 = 0; i0 < S.N; i0 += 256) {
        const int i = i0 + tid;
        const int ag = (i < S.N) ? st_ag(S.status[rb + i]) : -1;
        unsigned m[5];
#pragma unroll
        for (int g = 0; g < 5; ++g) { m[g] = __ballot_sync(FULL, ag == g); if (lane == 0) wcnt[wid][g] = __popc(m[g]); }
        __syncthreads();
        if (ag >= 0) {
            int pos = base[ag];
            for (int w = 0; w < wid; ++w) pos += wcnt[w][ag];
            unsigned mm = ag == 0 ? m[0] : ag == 1 ? m[1] : ag == 2 ? m[2] : ag == 3 ? m[3] : m[4];
            pos += __popc(mm & ((1u << lane) - 1u));
            S.grp_idx[rb + pos] = (uint32_t)i;
        }
        __syncthreads();
        if (tid < 5) { int s = 0; for (int w = 0; w < 8; ++w) s += wcnt[w][tid]; base[tid] += s; }
        __syncthreads();
    }
}

// ============================================================================ insert_infected :360-395
// CTA per replicate. Pick j takes the floor(u_j*(L-j))-th remaining susceptible in ascending index order.
// which: 0 = main's insert_infected(PRE, initialinf, 4) :117/:119; 1 = insert_infected(REC, initialhi, 4) :120
//        (skipped in calibration); 2 = explicit (health, num, ag) for the step API.
template <bool TP> __global__ void __launch_bounds__(256) k_insert_infected(Dev S, int which, int health, int num, int ag, int call_id, int day) {
    const int r = blockIdx.x, tid = threadIdx.x;
    const DevPset &P = S.psets[S.pset_of[r]];
    if (which == 0) { health = PRE; num = P.initialinf; call_id = 0; }
    if (which == 1) { if (P.calibration) return; health = REC; num = P.initialhi; call_id = 1; }
    const size_t rb = (size_t)r * S.Np;
    const uint2 key = S.key[r];
    const int seg = (S.N + 255) / 256, lo = tid * seg, hi = min(S.N, lo + seg);
    __shared__ int cnt[256];
    __shared__ int sel_tid, sel_k;
    int mine = 0;
    for (int i = lo; i < hi; ++i) {
        uint32_t st = S.status[rb + i];
        mine += (st_health(st) == SUS && (health != INF || st_ag(st) == ag - 1));       // :365-369
    }
    cnt[tid] = mine;
    __syncthreads();
    __shared__ int Ltot;
    if (tid == 0) { int L = 0; for (int t = 0; t < 256; ++t) L += cnt[t]; Ltot = L; }
    __syncthreads();
    if (Ltot == 0 || num > Ltot) { if (tid == 0 && (num > 0 || Ltot == 0)) atomicOr(&S.err[r], ERR_NO_SUS); return; }   // :371-372
    for (int j = 0; j < num; ++j) {
        if (tid == 0) {
            uint4 w = draw4<TP>(S, j, 0, S_SEED, call_id, key);
            int k = below(w.x, Ltot - j), t = 0;
            while (k >= cnt[t]) { k -= cnt[t]; ++t; }
            sel_tid = t; sel_k = k;
        }
        __syncthreads();
        if (tid == sel_tid) {
            int k = sel_k;
            for (int i = lo; i < hi; ++i) {
                uint32_t st = S.status[rb + i];
                if (!(st_health(st) == SUS && (health != INF || st_ag(st) == ag - 1))) continue;
                if (k-- > 0) continue;
                uint32_t ns = transition<TP>(S, P, r, i, day, st, key, health, /*seed_simple_inf=*/true, /*count_curves=*/false);
                if (health == LAT) {                                                     // x.tis = x.exp :383
                    const int ex = (int)S.texp[rb + i] - day;
                    S.entry[rb + i] = (int16_t)(day - ex); S.texp[rb + i] = S.expday[rb + i] = (int16_t)day;
                }
                // the contact budget already drawn for the coming day (:186 / :425) used the agent's old `iso`
                ns = (ns & ~ST_ISOB) | (st & ST_ISOB);
                S.status[rb + i] = ns;
                S.sickfrom[rb + i] = INF;                                                // :391
                break;
            }
            cnt[tid] -= 1;
        }
        __syncthreads();
    }
}

// ============================================================================ first column of hmatrix + day-1 curves
// _get_model_state(1, hmatrix) :136,:218-223 and the t=1 entries of :270-293 (every agent "first occurs" in
// its initial state on day 1).
__global__ void __launch_bounds__(256) k_first_column(Dev S) {
    const int r = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
    __shared__ int hist[5][12];
    if (threadIdx.x < 60) (&hist[0][0])[threadIdx.x] = 0;
    __syncthreads();
    if (i < S.N) {
        uint32_t st = S.status[(size_t)r * S.Np + i];
        int h = st_health(st);
        atomicAdd(&hist[st_ag(st)][h], 1);
        if (S.store_hm) S.hm[(size_t)r * S.T * S.N + i] = (uint8_t)h;
    }
    __syncthreads();
    if (threadIdx.x < 60) {
        int v = (&hist[0][0])[threadIdx.x];
        if (v) {
            int g = threadIdx.x / 12, s = threadIdx.x % 12;
            size_t b = (((size_t)r * 6 + (g + 1)) * (size_t)S.T + 0) * 12 + s;
            atomicAdd(&S.inc[b], v);
            atomicAdd(&S.prev[b], v);
        }
    }
}

// ============================================================================ dyntrans :790-834 (exact schedule)
// One warp per replicate walks the infectious agents in ascending index order (the reference's order, on which
// the result depends through the shared contact budgets :811-813); the lanes take that infector's contact
// events 32 at a time. Duplicate targets inside a batch are ranked with __match_any_sync so that the outcome
// equals processing the events one after another.
template <bool TP> __global__ void __launch_bounds__(32) k_dyntrans_exact(Dev S, int day) {
    const int r = blockIdx.x, lane = threadIdx.x;
    const DevPset &P = S.psets[S.pset_of[r]];
    const uint2 key = S.key[r];
    const size_t rb = (size_t)r * S.Np;
    const unsigned long long *bt = S.beta_thr + ((size_t)S.pset_of[r] * S.T + (size_t)(day - 1)) * 4;
    const int o0 = S.grp_off[r * 8 + 0], o1 = S.grp_off[r * 8 + 1], o2 = S.grp_off[r * 8 + 2],
              o3 = S.grp_off[r * 8 + 3], o4 = S.grp_off[r * 8 + 4], o5 = S.grp_off[r * 8 + 5];
    const bool ct = P.ctstrat != 0;
    uint32_t logn = ct ? S.ct_cnt[r] : 0;
    int met = 0, ninf = 0;
    const uint32_t lt = (1u << lane) - 1u;
    const uint32_t *mask = S.infmask + (size_t)r * S.nwords;
    for (int wb = 0; wb < S.nwords; wb += 32) {
        const uint32_t mw = (wb + lane < S.nwords) ? mask[wb + lane] : 0u;
        unsigned any = __ballot_sync(FULL, mw != 0);
        while (any) {
            const int L = __ffs(any) - 1; any &= any - 1;
            uint32_t m = __shfl_sync(FULL, mw, L);
            while (m) {
                const int b = __ffs(m) - 1; m &= m - 1;
                const int x = ((wb + L) << 5) + b;
                // ---- one infector (:798-803)
                const uint32_t sx = __ldcg(&S.status[rb + x]);
                const int xh = st_health(sx);
                const int cnts = budget_of<TP>(S, key, x, day, sx);
                if (cnts == 0) continue;
                const unsigned long long gp = __ldg(&S.gpw[st_ag(sx) * 256 + cnts]);     // :804
                const int p1 = (int)(gp & 0xFF), p2 = p1 + (int)((gp >> 8) & 0xFF), p3 = p2 + (int)((gp >> 16) & 0xFF),
                          p4 = p3 + (int)((gp >> 24) & 0xFF), E = p4 + (int)((gp >> 32) & 0xFF);
                const unsigned long long bthr = bt[beta_class(xh)];                      // :800
                bool tracing = ct && (sx & ST_TRACING);
                // contacts x makes while traced (:817-819) go to one chunk per (agent, day): [next chunk, count, targets...]
                uint32_t cbase = 0, ccount = 0;
                if (tracing) {
                    if (logn + 2u + (uint32_t)E <= S.ct_cap) { cbase = logn; logn += 2u + (uint32_t)E; }
                    else { if (lane == 0) atomicOr(&S.err[r], ERR_CT_LOG_FULL); tracing = false; }
                }
                for (int e0 = 0; e0 < E; e0 += 32) {
                    const int e = e0 + lane;
                    bool act = e < E;
                    const int g = (e >= p1) + (e >= p2) + (e >= p3) + (e >= p4);
                    const int k = e - (g == 0 ? 0 : g == 1 ? p1 : g == 2 ? p2 : g == 3 ? p3 : p4);
                    const int go = g == 0 ? o0 : g == 1 ? o1 : g == 2 ? o2 : g == 3 ? o3 : o4;
                    const int ge = g == 0 ? o1 : g == 1 ? o2 : g == 2 ? o3 : g == 3 ? o4 : o5;
                    const int len = ge - go;
                    if (act && len == 0) { atomicOr(&S.err[r], ERR_EMPTY_GROUP); act = false; }
                    uint32_t j = 0x80000000u + lane, sy = 0, w1 = 0;
                    int remy = 0;
                    if (act) {
                        const uint4 w = draw4<TP>(S, x, day, S_DT, ((uint32_t)g << 16) | (uint32_t)k, key);
                        w1 = w.y;
                        j = S.grp_idx[rb + go + below(w.x, len)];                        // :807
                        sy = __ldcg(&S.status[rb + j]);
                        remy = budget_of<TP>(S, key, j, day, sy);                            // :811
                    }
                    const unsigned peers = __match_any_sync(FULL, j);
                    const int rank = __popc(peers & lt), cntp = __popc(peers);
                    const bool adm = act && rank < remy;                                 // :812
                    const bool sus = st_health(sy) == SUS && st_swap(sy) == UNDEF;      // :822
                    const bool succ = adm && sus && bern(w1, bthr);                      // :823
                    const unsigned succm = __ballot_sync(FULL, succ) & peers;
                    const unsigned admm = __ballot_sync(FULL, adm);
                    if (act && rank == 0) {
                        const int used = min(cntp, remy);
                        uint32_t ns = (sy & ~(ST_REMMASK | ST_TAGMASK)) | (uint32_t)(remy - used) | ((uint32_t)day << 22);   // :813
                        if (succm) {                                                      // :824-827
                            ns = st_set_swap(ns, LAT);
                            S.sickfrom[rb + j] = (uint8_t)xh;
                            S.expday[rb + j] = (int16_t)day;                              // y.exp = y.tis: fires in today's time_update
                            ++ninf;
                        }
                        __stcg(&S.status[rb + j], ns);
                    }
                    met += adm;
                    if (tracing && admm) {                                                // :817-819
                        if (adm) S.ct_log[(size_t)r * S.ct_cap + cbase + 2u + ccount + (uint32_t)__popc(admm & lt)] = j;
                        ccount += (uint32_t)__popc(admm);
                    }
                    __syncwarp();
                }
                if (tracing) {
                    if (ccount == 0) logn = cbase;                                        // nothing admitted: give the chunk back
                    else if (lane == 0) {
                        uint32_t *c = S.ct_log + (size_t)r * S.ct_cap + cbase;
                        c[0] = S.ct_head[rb + x]; c[1] = ccount;
                        S.ct_head[rb + x] = cbase;
                    }
                }
            }
        }
    }
    for (int o = 16; o; o >>= 1) { met += __shfl_xor_sync(FULL, met, o); ninf += __shfl_xor_sync(FULL, ninf, o); }
    if (lane == 0) { S.totalmet[r] = met; S.totalinf[r] = ninf; if (ct) S.ct_cnt[r] = logn; }
}

// ============================================================================ dyntrans :790-834 (native schedule)
// One warp per infectious agent, lanes over its contact events, all replicates and all infectors in flight at once.
// The reference's sequential order (:797) is given up: a contact event claims one unit of the target's budget and, if
// the target is susceptible, infects it — both in a single atomicCAS on the target's status word, so the shared budget
// (:811-813) and "first infector wins" (:822-827) stay race-free. Which events are admitted when a budget runs out
// depends on timing: ensemble-equivalent to the reference (tests/test_native_ensemble.py), not bit-identical.
//
// In the reference an infector's own budget (:802) has already been reduced by the contacts it received from
// lower-indexed infectors (:813) when its turn comes. Launching everybody at once would hand out slightly too many
// contacts (measured: +0.2 % attack rate), so the kernel runs twice a day: COUNT draws every event from the undiminished
// budgets and counts, per infectious target, the hits coming from lower indices; APPLY gives infector x
// gpw(budget_x - hits_x) contacts and resolves them with CAS. COUNT runs twice (the second sweep uses the corrected
// budgets of the first), leaving a third-order difference in the infectious fraction.
template <bool COUNT, bool TP>
__global__ void __launch_bounds__(256) k_dyntrans_native(Dev S, int day, const uint32_t *hits_in, uint32_t *hits_out) {
    const int r = blockIdx.y, lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const DevPset &P = S.psets[S.pset_of[r]];
    const uint2 key = S.key[r];
    const size_t rb = (size_t)r * S.Np;
    const unsigned long long *bt = S.beta_thr + ((size_t)S.pset_of[r] * S.T + (size_t)(day - 1)) * 4;
    const int *off = S.grp_off + r * 8;
    const bool ct = P.ctstrat != 0;
    const uint32_t *mask = S.infmask + (size_t)r * S.nwords;
    // second COUNT sweep: only for a replicate whose edge list overflowed (k_native_edges did the others), or into the verify buffer
    if (COUNT && hits_in && S.nat_edges && S.nat_ecnt[r] <= (uint32_t)S.Np && hits_out != S.nat_verify) return;
    int met = 0, ninf = 0;
    // 64 mask words (2048 agents) per block, 8 per warp
    const int w_begin = blockIdx.x * 64 + wid * 8;
    for (int wi = w_begin; wi < w_begin + 8 && wi < S.nwords; ++wi) {
        uint32_t m = mask[wi];
        while (m) {
            const int b = __ffs(m) - 1; m &= m - 1;
            const int x = (wi << 5) + b;
            const uint32_t sx = __ldcg(&S.status[rb + x]);
            const int xh = st_health(sx);
            int cnts = st_health(sx) == DED ? 0 : fresh_budget<TP>(S, key, x, day, sx);    // :802, the day's undiminished budget
            if (hits_in) cnts = max(0, cnts - (int)((hits_in[(rb + x) >> 1] >> (16 * (x & 1))) & 0xFFFFu));
            if (cnts == 0) continue;
            const unsigned long long gp = __ldg(&S.gpw[st_ag(sx) * 256 + cnts]);     // :804
            const int p1 = (int)(gp & 0xFF), p2 = p1 + (int)((gp >> 8) & 0xFF), p3 = p2 + (int)((gp >> 16) & 0xFF),
                      p4 = p3 + (int)((gp >> 24) & 0xFF), E = p4 + (int)((gp >> 32) & 0xFF);
            const unsigned long long bthr = bt[beta_class(xh)];                      // :800
            const bool tracing = ct && (sx & ST_TRACING);
            uint32_t cbase = 0, ccount = 0;
            bool cfull = false;
            for (int e0 = 0; e0 < E; e0 += 32) {
                const int e = e0 + lane;
                bool adm = false;
                uint32_t j = 0;
                if (e < E) {
                    const int g = (e >= p1) + (e >= p2) + (e >= p3) + (e >= p4);
                    const int k = e - (g == 0 ? 0 : g == 1 ? p1 : g == 2 ? p2 : g == 3 ? p3 : p4);
                    const int go = off[g], len = off[g + 1] - go;
                    if (len == 0) atomicOr(&S.err[r], ERR_EMPTY_GROUP);
                    else {
                        const uint4 w = draw4<TP>(S, x, day, S_DT, ((uint32_t)g << 16) | (uint32_t)k, key);
                        j = S.grp_idx[rb + go + below(w.x, len)];                    // :807
                        if (COUNT) {
                            if ((int)j > x && ((mask[j >> 5] >> (j & 31)) & 1u)) {
                                atomicAdd(&hits_out[(rb + j) >> 1], 1u << (16 * (j & 1)));
                                if (!hits_in && S.nat_edges) {                        // first sweep: remember the event for the second-order count
                                    const uint32_t slot = atomicAdd(&S.nat_ecnt[r], 1u);
                                    if (slot < (uint32_t)S.Np) {
                                        uint32_t *rec = S.nat_edges + ((size_t)r * S.Np + slot) * 3;
                                        rec[0] = (uint32_t)x; rec[1] = j;
                                        rec[2] = (uint32_t)k | ((uint32_t)g << 8) | ((uint32_t)cnts << 16) | ((uint32_t)st_ag(sx) << 24);
                                    }
                                }
                            }
                            continue;
                        }
                        uint32_t old = __ldcg(&S.status[rb + j]);
                        int fresh = -1;
                        for (;;) {
                            int rem;
                            if (st_tag(old) == day) rem = st_rem(old);
                            else { if (fresh < 0) fresh = fresh_budget<TP>(S, key, j, day, old); rem = fresh; }
                            if (rem == 0) break;                                      // :812
                            uint32_t nw = (old & ~(ST_REMMASK | ST_TAGMASK)) | (uint32_t)(rem - 1) | ((uint32_t)day << 22);   // :813
                            const bool infect = st_health(old) == SUS && st_swap(old) == UNDEF && bern(w.y, bthr);           // :822-823
                            if (infect) nw = st_set_swap(nw, LAT);
                            const uint32_t seen = atomicCAS(&S.status[rb + j], old, nw);
                            if (seen == old) {
                                adm = true;
                                if (infect) { S.sickfrom[rb + j] = (uint8_t)xh; S.expday[rb + j] = (int16_t)day; ++ninf; }   // :825-827
                                break;
                            }
                            old = seen;
                        }
                    }
                }
                met += adm;
                if (!COUNT && tracing) {                                              // :817-819 — only this warp appends to x's chunk
                    const unsigned am = __ballot_sync(FULL, adm);
                    if (e0 == 0) {                                                        // reserve [next, count, E targets] once per (agent, day)
                        uint32_t b0 = 0;
                        if (lane == 0) b0 = atomicAdd(&S.ct_cnt[r], 2u + (uint32_t)E);
                        cbase = __shfl_sync(FULL, b0, 0);
                        if (cbase + 2u + (uint32_t)E > S.ct_cap) { if (lane == 0) atomicOr(&S.err[r], ERR_CT_LOG_FULL); cfull = true; }
                    }
                    if (!cfull) {
                        if (adm) S.ct_log[(size_t)r * S.ct_cap + cbase + 2u + ccount + (uint32_t)__popc(am & ((1u << lane) - 1u))] = j;
                        ccount += (uint32_t)__popc(am);
                    }
                }
            }
            if (!COUNT && tracing && !cfull && ccount > 0 && lane == 0) {
                uint32_t *c = S.ct_log + (size_t)r * S.ct_cap + cbase;
                c[0] = S.ct_head[rb + x]; c[1] = ccount;
                S.ct_head[rb + x] = cbase;
            }
        }
    }
    for (int o = 16; o; o >>= 1) { met += __shfl_xor_sync(FULL, met, o); ninf += __shfl_xor_sync(FULL, ninf, o); }
    if (!COUNT && lane == 0 && (met | ninf)) { atomicAdd(&S.totalmet[r], met); atomicAdd(&S.totalinf[r], ninf); }
}

// Second-order count of the native schedule from the first sweep's edge list: the event (g, k) of infector x still exists under
// x's corrected budget iff k < round(cm[ag][g] * (budget - hits0[x])) (:804 is monotone in the budget, and the event's draws are
// keyed by (x, day, g, k), so it meets the same target). One thread per edge instead of a second sweep over every event.
__global__ void __launch_bounds__(256) k_native_edges(Dev S, const uint32_t *h0, uint32_t *h1) {
    const int r = blockIdx.y;
    const uint32_t n = S.nat_ecnt[r];
    if (n > (uint32_t)S.Np) return;                                   // overflowed: the full sweep counts this replicate
    const size_t rb = (size_t)r * S.Np;
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint32_t *rec = S.nat_edges + (rb + e) * 3;
        const uint32_t x = rec[0], j = rec[1], w = rec[2];
        const int k = (int)(w & 0xFFu), g = (int)((w >> 8) & 7u), cnts = (int)((w >> 16) & 0xFFu), ag = (int)((w >> 24) & 7u);
        const int c2 = max(0, cnts - (int)((h0[(rb + x) >> 1] >> (16 * (x & 1))) & 0xFFFFu));
        if (c2 == 0) continue;
        const unsigned long long gp = __ldg(&S.gpw[ag * 256 + c2]);
        if (k < (int)((gp >> (8 * g)) & 0xFF)) atomicAdd(&h1[(rb + j) >> 1], 1u << (16 * (j & 1)));
    }
}
// end of the day's dyntrans: the hit counters go back to zero — only the words the edges touched (all of the replicate's after an overflow)
__global__ void __launch_bounds__(256) k_native_clean(Dev S, uint32_t *h0, uint32_t *h1) {
    const int r = blockIdx.y;
    const uint32_t n = S.nat_ecnt[r];
    const size_t rb = (size_t)r * S.Np;
    if (n > (uint32_t)S.Np) {
        for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < (size_t)S.Np / 2; k += (size_t)gridDim.x * blockDim.x) { h0[rb / 2 + k] = 0; h1[rb / 2 + k] = 0; }
        return;
    }
    for (uint32_t e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const uint32_t j = S.nat_edges[(rb + e) * 3 + 1];
        h0[(rb + j) >> 1] = 0; h1[(rb + j) >> 1] = 0;
    }
}
// test aid: the edge-list count must equal the full second sweep's
__global__ void __launch_bounds__(256) k_native_verify(Dev S, const uint32_t *h1) {
    const size_t n = (size_t)S.R * S.Np / 2;
    for (size_t k = blockIdx.x * (size_t)blockDim.x + threadIdx.x; k < n; k += (size_t)gridDim.x * blockDim.x) {
        if (h1[k] != S.nat_verify[k]) atomicOr(&S.err[(2 * k) / S.Np], ERR_INTERNAL);
        S.nat_verify[k] = 0;
    }
}

// ============================================================================ contact tracing, pass A
// The identification branch of ct_dynamics :705-716 writes *other* agents from inside the sequential agent loop.
// An index case x identifies on a day fixed when it became latent (doi == traceend), so the cross-agent writes are
// hoisted: this kernel draws Bernoulli(fcontactst) once per distinct traced contact y (:709-713) and leaves a mark —
// "pre" if y > x (y sees the isolation in its own update today), "post" if y < x (y was already updated today).
// One index case, handled by a warp or half a warp: gather its traced contacts (chunks are contiguous, so the reads coalesce),
// drop repeats (the reference's contact set is a Bool vector, :25) and draw Bernoulli(fcontactst) per distinct contact.
constexpr int CT_GATHER = 512;          // contacts gathered in shared memory per warp; longer sets take the slow path
// G lanes (a warp or half a warp) handle one index case; buf/htab are that group's gather buffer and hash table (hcap a power of two)
template <bool TP = false, int G = 32>
__device__ inline void ct_identify_warp(const Dev &S, const DevPset &P, int r, int x, int day, uint32_t *buf, int bufcap,
                                        uint32_t *htab, int hcap, uint32_t *mpre_row, uint32_t *mpost_row) {
    const int lane = threadIdx.x & 31, gl = lane & (G - 1);
    const unsigned gmask = G == 32 ? FULL : (((1u << (G & 31)) - 1u) << (lane & ~(G - 1)));
    const size_t a = (size_t)r * S.Np + x;
    const uint2 key = S.key[r];
    const uint32_t *log = S.ct_log + (size_t)r * S.ct_cap;
    const uint32_t head = S.ct_head[a];
    auto contact = [&](uint32_t y) {
        const uint4 w = draw4<TP>(S, x, day, S_CTCON, y, key);
        if (!bern(w.x, P.fcontact_thr)) return;               // :711
        atomicAdd(&S.totiso[r], 1ull);                        // :649
        if (y > (uint32_t)x) atomicOr(&mpre_row[y >> 5], 1u << (y & 31));
        else if (y < (uint32_t)x) atomicOr(&mpost_row[y >> 5], 1u << (y & 31));
    };
    // one walk along the chain of the case's chunks (newest first), gathering as it goes
    int n = 0;
    bool fits = true;
    for (uint32_t c = head; c != CT_NONE;) {
        const uint32_t next = log[c];
        const int cnt = (int)log[c + 1];
        if (n + cnt > bufcap) { fits = false; break; }
        for (int k = gl; k < cnt; k += G) buf[n + k] = log[c + 2 + k];
        n += cnt; c = next;
    }
    if (fits && 2 * n <= hcap) {
        // distinct contacts (the reference's contact set is a Bool vector, :25): every gathered contact is inserted into a small
        // open-addressing table; the lane whose insert claims an empty slot owns that contact
        for (int k = gl; k < hcap; k += G) htab[k] = 0xFFFFFFFFu;
        __syncwarp(gmask);
        for (int k = gl; k < n; k += G) {
            const uint32_t y = buf[k];
            uint32_t slot = ((y * 2654435761u) >> 16) & (uint32_t)(hcap - 1);
            for (;;) {
                const uint32_t old = atomicCAS(&htab[slot], 0xFFFFFFFFu, y);
                if (old == 0xFFFFFFFFu) { contact(y); break; }
                if (old == y) break;
                slot = (slot + 1u) & (uint32_t)(hcap - 1);
            }
        }
        __syncwarp(gmask);
    } else if (fits) {
        __syncwarp(gmask);
        for (int k = gl; k < n; k += G) {
            const uint32_t y = buf[k];
            bool dup = false;
            for (int q = 0; q < k; ++q) if (buf[q] == y) { dup = true; break; }
            if (!dup) contact(y);
        }
        __syncwarp(gmask);
    } else {
        // very long contact sets: the group walks the log itself, one entry per lane, each compared with every entry before it.
        // Control flow is uniform over the group. (No lane may run ahead to the next index case: its gather would land in buffer
        // slots that belong to other lanes of a case still being gathered — every branch ends with the group together.)
        for (uint32_t c = head; c != CT_NONE; c = log[c]) {
            const uint32_t cnt = log[c + 1];
            for (uint32_t k0 = 0; k0 < cnt; k0 += G) {
                const uint32_t k = k0 + (uint32_t)gl;
                const bool have = k < cnt;
                const uint32_t y = have ? log[c + 2 + k] : 0u;
                bool dup = !have;
                for (uint32_t c2 = head;; c2 = log[c2]) {
                    const uint32_t n2 = c2 == c ? min(cnt, k0 + (uint32_t)G) : log[c2 + 1];
                    for (uint32_t k2 = 0; k2 < n2; ++k2) {
                        const uint32_t z = log[c2 + 2 + k2];
                        if (z == y && (c2 != c || k2 < k)) dup = true;
                    }
                    if (c2 == c) break;
                }
                if (!dup) contact(y);
            }
        }
        __syncwarp(gmask);
    }
}

// scan variant (after state was poked by hand): every agent checks whether today is its identification day and queues itself
__global__ void __launch_bounds__(256) k_ct_scan_identifiers(Dev S, int day) {
    const int r = blockIdx.y, x = blockIdx.x * blockDim.x + threadIdx.x;
    const DevPset &P = S.psets[S.pset_of[r]];
    if (P.ctstrat == 0 || x >= S.N) return;
    const size_t a = (size_t)r * S.Np + x;
    const int te = S.traceend[a];
    if (te < 0 || day - (int)S.latday[a] != te) return;
    const uint32_t slot = atomicAdd(&S.ident_cnt[r * 2 + (day & 1)], 1u);
    S.ident_list[((size_t)r * 2 + (day & 1)) * S.Np + slot] = (uint32_t)x;
}

// Pass A over the queued index cases of today (queued by yesterday's agent_update_ct, or by the scan above): the work is
// proportional to the number of index cases, not to the population. CTA per replicate, warp per index case.
template <bool TP> __global__ void __launch_bounds__(128) k_ct_identify_list(Dev S, int day) {
    __shared__ uint32_t buf[4][CT_GATHER];
    const int r = blockIdx.x, wid = threadIdx.x >> 5;
    const DevPset &P = S.psets[S.pset_of[r]];
    if (P.ctstrat == 0) return;
    uint32_t *cnt = &S.ident_cnt[r * 2 + (day & 1)];
    const uint32_t n = *cnt;
    const uint32_t *list = S.ident_list + ((size_t)r * 2 + (day & 1)) * S.Np;
    const int grp = (threadIdx.x >> 4);                        // eight half-warps, one index case each at a time
    uint32_t *gb = buf[wid] + ((threadIdx.x >> 4) & 1) * (CT_GATHER / 2);
    for (uint32_t k = grp; k < n; k += 8)
        ct_identify_warp<TP, 16>(S, P, r, (int)list[k], day, gb, CT_GATHER / 4, gb + CT_GATHER / 4, CT_GATHER / 4,
                                 S.mark_pre + (size_t)r * S.nwords, S.mark_post + (size_t)r * S.nwords);
    __syncthreads();
    if (threadIdx.x == 0) *cnt = 0;
}

// ============================================================================ time_update :399-428 (+ snapshot :218-223)
// The per-agent streaming kernel. tis/doi/tracedxp are not stored as counters (absolute days) and the next-day contact
// count (:425) is drawn lazily by dyntrans, so an agent-day is one 16-bit compare of the agent's next event day; the next
// hmatrix column is the previous one with the (rare) changes patched in. Thread per 16 consecutive agents: two 128-bit
// loads of event days, one 128-bit load and one 128-bit store of column bytes = 4 B per agent-day (2 B without the state
// matrix; replicates with contact tracing also read their two mark bitmasks, 0.25 B). Everything else — the status word,
// durations, the tracing fields — is touched only by agents that have an event today (agent_event).
//
// agent_event: time_update's body for one agent whose next event day has come (:403-425), in the reference's order:
// [pre mark] -> state transition if expired (:405-420) -> ct_dynamics (:658-752: window set-up, tracing on,
// identification, quarantine expiry) -> isob := iso (the value :778 sees) -> [post mark]; then the next event day.
// The marks are the hoisted cross-agent writes of :705-716 (k_ct_identify_list). Quarantine countdown (:718-750):
// tracedxp == txpday - day after today's update.
template <bool TP> __device__ __noinline__ uint32_t agent_event(const Dev &S, int r, int i, int day, bool pre, bool post) {
    const size_t a = (size_t)r * S.Np + i;
    const uint32_t st0 = S.status[a];
    const int texp0 = S.texp[a];
    const int par = (day + 1) & 1;
    const DevPset &P = S.psets[S.pset_of[r]];
    const Ev e = agent_event_core<TP>(S, P, r, i, day, S.key[r], st0, texp0, P.ctstrat != 0, pre, post,
                                  S.ident_cnt ? &S.ident_cnt[r * 2 + par] : nullptr, S.ident_list ? S.ident_list + ((size_t)r * 2 + par) * S.Np : nullptr);
    if (e.newh != e.oldh) {
        if (is_infectious(e.newh) != is_infectious(e.oldh)) atomicXor(&S.infmask[(size_t)r * S.nwords + (i >> 5)], 1u << (i & 31));
        curve_move(S, r, st_ag(st0), day, e.oldh, e.newh);
    }
    if (e.texp != texp0) S.texp[a] = (int16_t)e.texp;
    S.expday[a] = (int16_t)e.ne;
    if (e.st != st0) S.status[a] = e.st;
    return (uint32_t)e.newh;
}

// Two launches per day. k_time_update_scan is the pure stream: thread per 32 consecutive agents, four 128-bit loads of event days,
// two 128-bit loads and stores of column bytes (the next column starts as a copy of the previous one), the mark bits;
// agents with an event today are appended to the replicate's event list (one atomicAdd per warp). k_time_update_events then runs
// agent_event for exactly those agents, one per thread in dense warps, and patches their byte of the new column. (A single kernel
// that handled events inside the stream stalled whole warps on one lane's dependent HBM loads: 0.10 of HBM peak on 10^6 agents.)
__global__ void __launch_bounds__(256) k_time_update_scan(const __grid_constant__ Dev S, int day, int any_ct) {
    // thread per 32 consecutive agents (one word of each mark bitmask): four 128-bit loads of event days in flight, two of column bytes
    const int r = blockIdx.y, lane = threadIdx.x & 31;
    const int i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 32;
    const bool valid = i0 < S.N;
    unsigned fire = 0, pre = 0, post = 0;
    if (valid) {
        const size_t a0 = (size_t)r * S.Np + i0;
        uint4 e[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) e[q] = *reinterpret_cast<const uint4 *>(S.expday + a0 + 8 * q);     // Np is padded to 32: in bounds
        const bool hm_on = S.store_hm && day < S.T;
        const bool vec = (S.N & 15) == 0;
        uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
        const uint8_t *prv = nullptr; uint8_t *cur = nullptr;
        if (hm_on) {                                                     // column st+1 = state at the start of the next day: copy, patched by the events
            prv = S.hm + ((size_t)r * S.T + (size_t)(day - 1)) * S.N + i0;
            cur = S.hm + ((size_t)r * S.T + (size_t)day) * S.N + i0;
            if (vec) { c0 = *reinterpret_cast<const uint4 *>(prv); if (i0 + 16 < S.N) c1 = *reinterpret_cast<const uint4 *>(prv + 16); }
        }
        if (any_ct) {                                                    // marks left by today's index cases (pass A)
            const size_t wi = (size_t)r * S.nwords + (i0 >> 5);
            pre = S.mark_pre[wi]; post = S.mark_post[wi];
            if (pre) S.mark_pre[wi] = 0;                                 // consumed by this update (this thread owns the word)
            if (post) S.mark_post[wi] = 0;
        }
        const uint32_t dd = (uint32_t)day * 0x10001u;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t m[4] = {__vcmples2(e[q].x, dd), __vcmples2(e[q].y, dd), __vcmples2(e[q].z, dd), __vcmples2(e[q].w, dd)};
#pragma unroll
            for (int k = 0; k < 4; ++k) fire |= ((m[k] & 1u) | ((m[k] >> 15) & 2u)) << (8 * q + 2 * k);
        }
        fire = (fire | pre | post) & (S.N - i0 >= 32 ? 0xFFFFFFFFu : ((1u << (S.N - i0)) - 1u));
        if (hm_on) {
            if (vec) { *reinterpret_cast<uint4 *>(cur) = c0; if (i0 + 16 < S.N) *reinterpret_cast<uint4 *>(cur + 16) = c1; }
            else for (int k = 0; k < 32 && i0 + k < S.N; ++k) cur[k] = prv[k];
        }
    }
    // append the agents with an event: exclusive prefix of the counts over the warp, one atomicAdd per warp
    const int c = __popc(fire);
    if (__any_sync(FULL, c != 0)) {
        int incl = c;
        for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, incl, o); if (lane >= o) incl += v; }
        const int total = __shfl_sync(FULL, incl, 31);
        unsigned base = 0;
        if (lane == 31) base = atomicAdd(&S.ev_cnt[r * 2 + (day & 1)], (unsigned)total);
        base = __shfl_sync(FULL, base, 31);
        uint32_t *out = S.ev_list + (size_t)r * S.Np + base + (unsigned)(incl - c);
        while (fire) {
            const int k = __ffs(fire) - 1; fire &= fire - 1;
            *out++ = (uint32_t)(i0 + k) | (((pre >> k) & 1u) << 30) | (((post >> k) & 1u) << 31);
        }
    }
}

template <bool TP> __global__ void __launch_bounds__(256) k_time_update_events(const __grid_constant__ Dev S, int day) {
    const int r = blockIdx.y;
    const unsigned n = S.ev_cnt[r * 2 + (day & 1)];
    const bool hm_on = S.store_hm && day < S.T;
    for (unsigned k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
        const uint32_t ent = S.ev_list[(size_t)r * S.Np + k];
        const int i = (int)(ent & 0x3FFFFFFFu);
        const uint32_t newh = agent_event<TP>(S, r, i, day, (ent >> 30) & 1u, ent >> 31);
        if (hm_on) S.hm[((size_t)r * S.T + (size_t)day) * S.N + i] = (uint8_t)newh;
    }
    if (blockIdx.x == 0 && threadIdx.x == 0) S.ev_cnt[r * 2 + ((day + 1) & 1)] = 0;      // tomorrow's counter (read two launches from now)
}

// Seeds isolated through fpreiso at seeding keep the day-1 contact budget that initialize drew for them (:186): their
// isob trails iso until the end of the first update. One pass after time_update(1).
__global__ void __launch_bounds__(256) k_resync_isob(Dev S) {
    const size_t n = (size_t)S.R * S.Np;
    for (size_t a = blockIdx.x * (size_t)blockDim.x + threadIdx.x; a < n; a += (size_t)gridDim.x * blockDim.x) {
        const uint32_t st = S.status[a];
        if (((st >> 19) ^ (st >> 20)) & 1u) S.status[a] = (st & ST_ISO) ? (st | ST_ISOB) : (st & ~ST_ISOB);
    }
}

// hmatrix leaves the device as Int16 (:111): widen a span of the stored health bytes
__global__ void k_widen_u8_i16(const uint8_t *src, int16_t *dst, size_t n) {
    for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 16; i < n; i += (size_t)gridDim.x * blockDim.x * 16) {
        if (i + 16 <= n && ((reinterpret_cast<uintptr_t>(src + i) | reinterpret_cast<uintptr_t>(dst + i)) & 15) == 0) {
            union { uint4 v; uint8_t b[16]; } in; in.v = *reinterpret_cast<const uint4 *>(src + i);
            union { uint4 v[2]; int16_t h[16]; } out;
#pragma unroll
            for (int k = 0; k < 16; ++k) out.h[k] = in.b[k];
            reinterpret_cast<uint4 *>(dst + i)[0] = out.v[0]; reinterpret_cast<uint4 *>(dst + i)[1] = out.v[1];
        } else {
            for (size_t k = i; k < n && k < i + 16; ++k) dst[k] = src[k];
        }
    }
}

// curves leave as Int16 when every count fits (popsize <= 32767): halves the device->host bytes of an ensemble
__global__ void k_narrow_i32_i16(const int *a, const int *b, int16_t *out, size_t n) {
    for (size_t i = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 4; i < 2 * n; i += (size_t)gridDim.x * blockDim.x * 4) {
        const int *src = i < n ? a + i : b + (i - n);          // n is a multiple of 4 (12 states per row)
        const int4 v = *reinterpret_cast<const int4 *>(src);
        union { uint2 u; int16_t h[4]; } o;
        o.h[0] = (int16_t)v.x; o.h[1] = (int16_t)v.y; o.h[2] = (int16_t)v.z; o.h[3] = (int16_t)v.w;
        *reinterpret_cast<uint2 *>(out + i) = o.u;
    }
}

// compute_yearly_average (scripts/local_run.jl:75-103): mean over the replicates of every curve entry, per day
__global__ void k_mean_curves(const int *inc, const int *prev, double *minc, double *mprev, int R, size_t per_rep) {
    const size_t e = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (e >= per_rep) return;
    long long a = 0, b = 0;
    for (int r = 0; r < R; ++r) { a += inc[(size_t)r * per_rep + e]; b += prev[(size_t)r * per_rep + e]; }
    minc[e] = (double)a / (double)R; mprev[e] = (double)b / (double)R;
}

// ============================================================================ curves: :259-293 from the transition counters
// During the run inc[r][g][t][s] counted entries into s first visible in column t and prev[] parked the exits.
// Every state is entered at most once per agent (SUS->LAT->PRE|ASYMP->...->REC|DED), so first-occurrence incidence
// (:270-280) is the entry count and prevalence (:282-293) is the running sum entries - exits. Group 0 = all agents.
__global__ void __launch_bounds__(64) k_curves_finalize(Dev S) {
    const int r = blockIdx.x, s = threadIdx.x;
    if (s >= 12) return;
    const size_t T = S.T;
    int *inc0 = S.inc + (size_t)r * 6 * T * 12, *prev0 = S.prev + (size_t)r * 6 * T * 12;
    for (size_t t = 0; t < T; ++t) { inc0[t * 12 + s] = 0; prev0[t * 12 + s] = 0; }
    long long lat = 0;
    for (int g = 1; g <= 5; ++g) {
        int *inc = inc0 + (size_t)g * T * 12, *prev = prev0 + (size_t)g * T * 12;
        int run = prev[s];
        inc0[s] += inc[s]; prev0[s] += run;
        for (size_t t = 1; t < T; ++t) {
            const int in = inc[t * 12 + s];
            run += in - prev[t * 12 + s];
            prev[t * 12 + s] = run;
            inc0[t * 12 + s] += in; prev0[t * 12 + s] += run;
        }
    }
    if (s == LAT) { for (size_t t = 0; t < T; ++t) lat += inc0[t * 12 + LAT]; S.latsum[r] = lat; }
}

// ============================================================================ _count_infectors :295-313
__global__ void __launch_bounds__(256) k_count_infectors(Dev S) {
    const int r = blockIdx.x;
    __shared__ int c[4];
    if (threadIdx.x < 4) c[threadIdx.x] = 0;
    __syncthreads();
    int l[4] = {0, 0, 0, 0};
    bool bad = false;
    for (int i = threadIdx.x; i < S.N; i += blockDim.x) {
        const size_t a = (size_t)r * S.Np + i;
        if (st_health(S.status[a]) == SUS) continue;
        const int f = S.sickfrom[a];
        if (f == PRE) l[0]++; else if (f == ASYMP) l[1]++; else if (f == MILD || f == MISO) l[2]++;
        else if (f == INF || f == IISO) l[3]++; else bad = true;
    }
    for (int k = 0; k < 4; ++k) if (l[k]) atomicAdd(&c[k], l[k]);
    if (bad) atomicOr(&S.err[r], ERR_SICKFROM);
    __syncthreads();
    if (threadIdx.x < 4) S.infectors[r * 4 + threadIdx.x] = c[threadIdx.x];
}

// ============================================================================ literal reductions over given matrices
// _get_column_incidence :270-280 (first occurrence along each row) and _get_column_prevalence :282-293 (count per
// column), for "all" and the five _splitstate groups :247-256. hm is [M][T][N]; thread per agent walks its row.
__global__ void __launch_bounds__(256) k_reduce_hmatrix(const int16_t *hm, const int *ags, int N, int T, int *inc, int *prev) {
    const int m = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x, lane = threadIdx.x & 31;
    const bool valid = i < N;
    const int g = valid ? ags[(size_t)m * N + i] : 0;          // 1..5
    const int16_t *row = hm + (size_t)m * T * N + i;
    int *incm = inc + (size_t)m * 6 * T * 12, *prevm = prev + (size_t)m * 6 * T * 12;
    unsigned seen = 0;
    for (int t = 0; t < T; ++t) {
        int h = valid ? row[(size_t)t * N] : -1;
        if (h < 0 || h > 11) h = -1;
        const bool first = h >= 0 && !((seen >> h) & 1);
        if (h >= 0) seen |= 1u << h;
        // warp-aggregate equal (group, state, first) triples before touching global memory
        const int keyv = h < 0 ? -1 : (g * 12 + h) * 2 + (first ? 1 : 0);
        const unsigned peers = __match_any_sync(FULL, keyv);
        if (h >= 0 && (__ffs(peers) - 1) == lane) {
            const int c = __popc(peers);
            atomicAdd(&prevm[((size_t)g * T + t) * 12 + h], c);
            atomicAdd(&prevm[((size_t)0 * T + t) * 12 + h], c);
            if (first) { atomicAdd(&incm[((size_t)g * T + t) * 12 + h], c); atomicAdd(&incm[((size_t)0 * T + t) * 12 + h], c); }
        }
    }
}

// ============================================================================ peek / poke (tests read `humans` like test/runtests.jl)
template <bool TP> __global__ void __launch_bounds__(256) k_peek(Dev S, int field, int r, int day, int *out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= S.N) return;
    const size_t a = (size_t)r * S.Np + i;
    const uint32_t st = S.status[a];
    const int done = day - 1;                    // number of time_updates so far
    const int tis = done - S.entry[a];
    int v = 0;
    switch (field) {
    case 0: v = st_health(st); break;
    case 1: v = st_swap(st); break;
    case 2: v = S.sickfrom[a]; break;
    case 3: v = budget_of<TP>(S, S.key[r], i, day, st); break;
    case 4: v = st_ag(st) + 1; break;
    case 5: v = tis; break;
    case 6: v = (st_health(st) == SUS && st_swap(st) == LAT) ? tis : (int)S.texp[a] - (int)S.entry[a]; break;
    case 7: { const uint32_t d4 = S.dur[a]; out[4 * i] = d4 & 0xFF; out[4 * i + 1] = (d4 >> 8) & 0xFF; out[4 * i + 2] = (d4 >> 16) & 0xFF; out[4 * i + 3] = d4 >> 24; return; }
    case 8: v = done - S.latday[a]; break;
    case 9: v = (st & ST_ISO) ? 1 : 0; break;
    case 10: v = S.isovia[a]; break;
    case 11: v = (st & ST_TRACING) ? 1 : 0; break;
    case 12: v = S.tracestart[a]; break;
    case 13: v = S.traceend[a]; break;
    case 14: { const int t = S.txpday[a]; v = (t != 0 && t > done) ? t - done : 0; break; }      // tracedxp :26
    case 15: {
        const uint32_t *log = S.ct_log + (size_t)r * S.ct_cap;
        const uint32_t head = S.ct_head[a];
        for (uint32_t c = head; c != CT_NONE; c = log[c]) {
            for (uint32_t k = 0; k < log[c + 1]; ++k) {
                const uint32_t y = log[c + 2 + k];
                bool dup = false;
                for (uint32_t c2 = head; c2 != CT_NONE && !dup; c2 = log[c2]) {
                    const uint32_t lim = c2 == c ? k : log[c2 + 1];
                    for (uint32_t k2 = 0; k2 < lim; ++k2) if (log[c2 + 2 + k2] == y) { dup = true; break; }
                    if (c2 == c) break;
                }
                v += !dup;
            }
        }
        break; }
    }
    out[i] = v;
}

__global__ void k_poke(Dev S, int field, int r, int i, int day, int value) {
    const size_t a = (size_t)r * S.Np + i;
    uint32_t st = S.status[a];
    const int done = day - 1;
    switch (field) {
    case 0: {
        const int oldh = st_health(st);
        if (is_infectious(oldh) != is_infectious(value)) atomicXor(&S.infmask[(size_t)r * S.nwords + (i >> 5)], 1u << (i & 31));
        st = st_set_health(st, value); break; }
    case 1: st = st_set_swap(st, value); break;
    case 2: S.sickfrom[a] = (uint8_t)value; break;
    case 3: st = (st & ~(ST_REMMASK | ST_TAGMASK)) | (uint32_t)(value & 0xFF) | ((uint32_t)day << 22); break;
    case 5: { const int exp = (int)S.texp[a] - (int)S.entry[a]; S.entry[a] = (int16_t)(done - value); S.texp[a] = (int16_t)(done - value + exp); break; }
    case 6: S.texp[a] = (int16_t)((int)S.entry[a] + value); break;
    case 8: S.latday[a] = (int16_t)(done - value); break;
    case 9: st = value ? (st | ST_ISO | ST_ISOB) : (st & ~(ST_ISO | ST_ISOB)); break;
    case 10: S.isovia[a] = (uint8_t)value; break;
    case 11: st = value ? (st | ST_TRACING) : (st & ~ST_TRACING); break;
    case 12: S.tracestart[a] = (int16_t)value; break;
    case 13: S.traceend[a] = (int16_t)value; break;
    case 14: S.txpday[a] = (int16_t)(value > 0 ? done + value : 0); break;
    default: break;
    }
    S.status[a] = st;
    S.expday[a] = 0;      // wake the agent at the next update: agent_event recomputes its next event day from the edited fields
}

// ============================================================================ launchers
static inline dim3 agent_grid(const Dev &S, int per_thread) {
    return dim3((unsigned)((S.Np + 256 * per_thread - 1) / (256 * per_thread)), (unsigned)S.R);
}

int launch_initialize(const Dev &S, cudaStream_t s) {
    if (S.tape_n) k_initialize<true><<<agent_grid(S, 1), 256, 0, s>>>(S); else k_initialize<false><<<agent_grid(S, 1), 256, 0, s>>>(S);
    return 1;
}
int launch_build_groups(const Dev &S, cudaStream_t s) { k_build_groups<<<S.R, 256, 0, s>>>(S); return 1; }
int launch_insert_infected(const Dev &S, int which, int health, int num, int ag, int call_id, int day, cudaStream_t s) {
    if (S.tape_n) k_insert_infected<true><<<S.R, 256, 0, s>>>(S, which, health, num, ag, call_id, day);
    else k_insert_infected<false><<<S.R, 256, 0, s>>>(S, which, health, num, ag, call_id, day);
    return 1;
}
int launch_first_column(const Dev &S, cudaStream_t s) { k_first_column<<<agent_grid(S, 1), 256, 0, s>>>(S); return 1; }
int launch_dyntrans_exact(const Dev &S, int day, cudaStream_t s) {
    if (S.tape_n) k_dyntrans_exact<true><<<S.R, 32, 0, s>>>(S, day); else k_dyntrans_exact<false><<<S.R, 32, 0, s>>>(S, day);
    return 1;
}
int launch_dyntrans_native(const Dev &S, int day, cudaStream_t s) {
    cudaMemsetAsync(S.totalmet, 0, (size_t)S.R * 4, s); cudaMemsetAsync(S.totalinf, 0, (size_t)S.R * 4, s);
    // two fixed-point counts for the budgets (first and second order in the infectious fraction), then the CAS pass. The first is a
    // sweep over every event, which also lists the events that hit a later infector; the second only walks that list.
    const size_t half = (size_t)S.R * S.Np / 2;
    uint32_t *h0 = S.hits, *h1 = S.hits + half;
    int n = 3;
    if (!S.nat_edges || day == 1) cudaMemsetAsync(S.hits, 0, half * 8, s);       // (with the edge list the counters are cleared by k_native_clean)
    if (S.nat_edges) cudaMemsetAsync(S.nat_ecnt, 0, (size_t)S.R * 4, s);
    const dim3 grid((unsigned)((S.nwords + 63) / 64), (unsigned)S.R);
    unsigned eb = (unsigned)(S.N / 8192); if (eb < 1) eb = 1; if (eb > 32) eb = 32;
    const dim3 egrid(eb, (unsigned)S.R);
    if (S.tape_n) k_dyntrans_native<true, true><<<grid, 256, 0, s>>>(S, day, nullptr, h0);
    else k_dyntrans_native<true, false><<<grid, 256, 0, s>>>(S, day, nullptr, h0);
    if (S.nat_edges) { k_native_edges<<<egrid, 256, 0, s>>>(S, h0, h1); ++n; }
    if (S.tape_n) k_dyntrans_native<true, true><<<grid, 256, 0, s>>>(S, day, h0, h1);            // returns at once unless the list overflowed
    else k_dyntrans_native<true, false><<<grid, 256, 0, s>>>(S, day, h0, h1);
    if (S.nat_edges && S.nat_verify) {
        if (S.tape_n) k_dyntrans_native<true, true><<<grid, 256, 0, s>>>(S, day, h0, S.nat_verify);
        else k_dyntrans_native<true, false><<<grid, 256, 0, s>>>(S, day, h0, S.nat_verify);
        k_native_verify<<<1024, 256, 0, s>>>(S, h1); n += 2;
    }
    if (S.tape_n) k_dyntrans_native<false, true><<<grid, 256, 0, s>>>(S, day, h1, nullptr);
    else k_dyntrans_native<false, false><<<grid, 256, 0, s>>>(S, day, h1, nullptr);
    if (S.nat_edges) { k_native_clean<<<egrid, 256, 0, s>>>(S, h0, h1); ++n; }
    return n;
}
int launch_ct_identify(const Dev &S, int day, bool scan, cudaStream_t s) {
    int n = 1;
    if (scan) {        // the queue may be stale: rebuild today's from the agent fields
        cudaMemsetAsync(S.ident_cnt, 0, (size_t)S.R * 8, s);
        k_ct_scan_identifiers<<<agent_grid(S, 1), 256, 0, s>>>(S, day); ++n;
    }
    if (S.tape_n) k_ct_identify_list<true><<<S.R, 128, 0, s>>>(S, day); else k_ct_identify_list<false><<<S.R, 128, 0, s>>>(S, day);
    return n;
}
int launch_time_update_scan(const Dev &S, int day, bool any_ct, cudaStream_t s) {
    k_time_update_scan<<<agent_grid(S, 32), 256, 0, s>>>(S, day, any_ct ? 1 : 0);
    return 1;
}
int launch_time_update_events(const Dev &S, int day, cudaStream_t s) {
    int n = 1;
    unsigned eb = (unsigned)(S.N / 16384); if (eb < 1) eb = 1; if (eb > 64) eb = 64;
    if (S.tape_n) k_time_update_events<true><<<dim3(eb, (unsigned)S.R), 256, 0, s>>>(S, day);
    else k_time_update_events<false><<<dim3(eb, (unsigned)S.R), 256, 0, s>>>(S, day);
    if (day == 1) { k_resync_isob<<<148 * 8, 256, 0, s>>>(S); ++n; }
    return n;
}
int launch_curves_finalize(const Dev &S, cudaStream_t s) { k_curves_finalize<<<S.R, 64, 0, s>>>(S); return 1; }
int launch_count_infectors(const Dev &S, cudaStream_t s) { k_count_infectors<<<S.R, 256, 0, s>>>(S); return 1; }
int launch_reduce_hmatrix(const int16_t *hm, const int *ags, int M, int N, int T, int *inc, int *prev, cudaStream_t s) {
    k_reduce_hmatrix<<<dim3((unsigned)((N + 255) / 256), (unsigned)M), 256, 0, s>>>(hm, ags, N, T, inc, prev); return 1;
}
int launch_widen(const uint8_t *src, int16_t *dst, size_t n, cudaStream_t s) {
    size_t blocks = (n / 16 + 255) / 256; if (blocks > 148 * 16) blocks = 148 * 16; if (blocks < 1) blocks = 1;
    k_widen_u8_i16<<<(unsigned)blocks, 256, 0, s>>>(src, dst, n); return 1;
}
int launch_narrow(const int *a, const int *b, int16_t *out, size_t n, cudaStream_t s) {
    k_narrow_i32_i16<<<148 * 8, 256, 0, s>>>(a, b, out, n); return 1;
}
int launch_mean_curves(const int *inc, const int *prev, double *minc, double *mprev, int R, size_t per_rep, cudaStream_t s) {
    k_mean_curves<<<(unsigned)((per_rep + 255) / 256), 256, 0, s>>>(inc, prev, minc, mprev, R, per_rep); return 1;
}
int launch_peek(const Dev &S, int field, int r, int day, int *out, cudaStream_t s) {
    if (S.tape_n) k_peek<true><<<(S.N + 255) / 256, 256, 0, s>>>(S, field, r, day, out); else k_peek<false><<<(S.N + 255) / 256, 256, 0, s>>>(S, field, r, day, out);
    return 1;
}
int launch_poke(const Dev &S, int field, int r, int i, int day, int value, cudaStream_t s) { k_poke<<<1, 1, 0, s>>>(S, field, r, i, day, value); return 1; }

}  // namespace cabm

#include "cabm_fused.cuh"
