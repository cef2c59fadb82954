// cabm_fused.cuh — k_sim_fused: one persistent CTA runs a whole replicate of main() (:104-142) out of shared memory.
// Included by cabm_kernels.cu (same translation unit: shares c_tab).
//
// Replicates are independent (the reference parallelises only over them, scripts/local_run.jl:28-30), so a CTA
// needs no grid-wide synchronisation: it initialises its population, seeds it, and loops over all modeltime days
// with the hot per-agent state resident in shared memory:
//     s_status u32[Np]  (the word dyntrans gathers/scatters; cabm_device.cuh)
//     s_expday i16[Np]  (time_update's only per-agent-day read)
//     s_col    u8[Np]   (the current hmatrix column; transitions patch it, one TMA bulk store per day writes it)
// HBM traffic per agent-day is therefore just the 1-byte hmatrix element (:221, widened to Int16 on the way out); the rarely-touched fields
// (dur, entry, sickfrom, isovia, latday) stay in HBM and are read on state transitions only. The age-group index tables of
// get_ag_dist (:339-343, u16, 2 B per agent) are written once per replicate and read through L2/L1: a target draw is one load of
// many that are in flight per thread, so its latency is hidden, and the 20 KB of shared memory pay for the chunk buffers.
// CTAs pull replicate indices from an atomic work counter (grid = resident CTAs), so long epidemics and early
// extinctions balance across SMs.
//
// dyntrans keeps the reference's sequential semantics (:797-832) without walking the infectors one after another. A day's
// infectious agents are listed in ascending index order and taken in CHUNKS of up to 256 infectors / FUSED_EVCAP contact events;
// inside a chunk everything is parallel:
//   * every potential event (infector x, age group g, slot k) of the chunk is drawn at once (:807) — targets and transmission
//     draws; a target's status word counts the day's contacts (tag = today) from the first event applied to it, and its own
//     lazy contact budget is drawn only when an outcome depends on it;
//   * the only order dependence BETWEEN infectors is x's own budget (:802), which earlier infectors reduce when they contact x
//     (:813). Such events (target infectious, later in the same chunk) are rare; they are collected as the edges of a DAG and
//     resolved by relaxation: hits on x -> smaller budget -> fewer events of x (the reduced budget's events are a subset of the
//     drawn ones, because round.(cm*cnts) :804 is monotone in cnts) -> fewer hits on later infectors ..., converging in
//     (longest dependency chain) sweeps;
//   * the order dependence ON A TARGET (its budget admits the first rem events :811-813, the first admitted event that transmits
//     infects :822-827) only matters for targets hit more than once in the chunk: those events are compacted in event order and
//     resolved per target by warps (each warp owns the targets of one residue class), everything else applies independently.
// Chunks are sequential (a later chunk sees the budgets the earlier ones left behind).
//
// Map of the kernel body (k_sim_fused):
//   work queues   sweep 0: fresh replicates for split_day days -> finished (fill job) or parked; sweep 1: parked replicates,
//                 copy jobs (curves -> pinned host memory, a few "copier" CTAs), fill jobs (idle tails of curves and state matrix)
//   per replicate initialize :171-188 -> get_ag_dist :339-343 -> insert_infected :360-395 -> column 1 -> day loop -> write-back
//   per active day infector list -> chunks { budgets + prefix | C1 | events | C2 | [DAG sweeps] | apply singles, count shared | C3 |
//                 compact shared | C4 | resolve shared | C5 } -> [pass A of contact tracing] -> time_update: scan lists the agents with
//                 an event | barrier | listed events, dense | S2 -> curves row + TMA store of the column
//   idle days     (nobody infectious, nothing due) repeat column and prevalence row; an idle tail is filled in one go
//
// Template GS ("global state", any population size): the same algorithm with status/expday/column/bitmasks left in global memory.
// Every per-agent access is then an L2/HBM round trip, so that variant keeps eight event windows in flight per warp, finds the
// targets a chunk hits more than once in a shared-memory table of the chunk's targets (M.htab) instead of population-wide bitmasks,
// and touches a target's status word once per event.
#pragma once
#include <cstdlib>
#include <mutex>
#include <type_traits>

namespace cabm {

constexpr int FUSED_NT = 256;          // threads per CTA
constexpr int FUSED_LCAP = 256;        // infectors listed per pass = infectors per chunk at most
constexpr int FUSED_EVCAP = 2048;      // contact events per chunk
constexpr int FUSED_HCAP = 8192;       // global-state variant: slots of the chunk's target table (four times the events of a chunk; a power of two)
constexpr int FUSED_NEDGE = 256;       // events of a chunk that hit a later infector of the same chunk (more: the chunk is halved)
constexpr unsigned FUSED_TAIL_CTAS = 8;  // CTAs that copy finished curves to the host at any one time (PCIe saturates with ~8)
constexpr int FUSED_SCAN_U = 5;        // chunks of 8 agents a thread scans at once in time_update (256 x 8 x 5 >= 10,000 agents in one go)
static_assert(FUSED_NT == 256 && FUSED_LCAP <= FUSED_NT && FUSED_EVCAP % 32 == 0 && FUSED_EVCAP <= 2048, "thread-index roles and record fields in k_sim_fused");
static_assert(FUSED_EVCAP / FUSED_NT <= 32, "one bit per window of a warp in the shared-target mask");
// event record: evy = target (EV_NONE: none); evm = transmits [0] | dropped by a reduced budget [1] | owner (chunk entry) [2:10)
constexpr uint32_t EV_NONE = 0xFFFFFFFFu;
constexpr uint16_t EVM_SUCC = 1, EVM_DROPPED = 2;

// Small per-CTA state. It lives at the start of the dynamic shared-memory block (not in static __shared__ arrays) so that every
// access is "block base + constant offset" from one uniform register.
struct FusedCtl {
    DevPset P;                               // this replicate's parameter set (read on every transition)
    double cm[25];                           // contact_matrix :838-848
    unsigned long long beta[4];              // _get_betavalue :756-769 thresholds of the day
    alignas(16) int run[72];                 // running prevalence of the six groups, for the idle-tail fill (16-byte aligned)
    int io[2][2][60];                        // [day parity][entries | exits][(group, state)] — curve deltas of one day
    uint32_t tlat[4], tpre[4], tasy[40], tinf[40];
    int r, off[8], wcnt[FUSED_NT / 32][5];
    int sel_tid, sel_k, Ltot, ninf;
    int nf[2], wsum[FUSED_NT / 32], wsum2[FUSED_NT / 32];
    int nb_len[5], kind;
    int infectors[4];
    int wfit[FUSED_NT / 32], wkept[FUSED_NT / 32], wneed[FUSED_NT / 32];   // per warp: entries that fit the chunk; entries with events and contact-log words up to its last fitting entry
    int nedge;                               // DAG edges of the chunk
    int wdup[FUSED_NT / 32];
    uint32_t logn, qcnt[2];                  // contact tracing (:647-752): the replicate's log fill, index-case queues by day parity
    unsigned evn;                            // agents with an event today, listed by the time_update scan
    int pad_[3];
};
static_assert(sizeof(FusedCtl) % 16 == 0, "the arrays behind FusedCtl need 16-byte alignment");

// IdxT: agent index type of the lists (u16 with the state in shared memory, u32 with the state in global memory)
template <class IdxT> struct FusedSmemT {
    uint32_t *status; int16_t *expday; uint8_t *col; uint32_t *inf;
    uint32_t *nb_thr; uint8_t *nb_guide; const int *nb_len;
    uint32_t *touch, *dupm, *mpre, *mpost;
    uint32_t *evy; uint16_t *evm;        // [EVCAP] events of the chunk, in event order: target (EV_NONE: no event); transmits | dropped << 1 | owner << 2
    uint16_t *D;                         // [EVCAP] events on shared targets, in event order
    uint32_t *edges;                     // [NEDGE] event [0:11) | entry of the target [11:19) | g [19:22) | k [22:30)
    IdxT *list;                          // [LCAP] the day's infectious agents from the cursor on, ascending
    IdxT *cx;                            // chunk entries (infectors with events): agent,
    uint8_t *ch, *cb, *cc, *cflag;       //   health | ag << 4 | tracing << 7; budget at chunk start; budget after hits; 1 = log contacts
    int *pref, *hits;                    //   event prefix [LCAP + 1]; hits by earlier entries of the chunk
    unsigned long long *gp0;             //   budget split over the five age groups (:804) at the chunk-start budget, five packed bytes
    uint32_t *gcum;                      //   the same split as cumulative counts of groups 1..4 (one byte each)
    uint32_t *cbase; int *clen;          //   contact-log chunk of a traced entry, contacts logged so far
    unsigned long long *dcache;          // (agent << 32 | dur) of recently infected agents
    int *cnt;                            // NT ints scratch (seeding, first column); aliases evy
    uint32_t *htab;                      // global-state variant: open-addressing table of the chunk's targets (bit 31: hit more than once)
};

// global_state: the per-agent arrays (status, expday, column, bitmasks) stay in global memory (any popsize); only the chunk
// buffers and tables are in shared memory
__host__ __device__ inline size_t fused_smem_bytes(int Np, bool global_state = false) {
    size_t b = sizeof(FusedCtl);
    if (!global_state) {
        b += (size_t)Np * 4;                 // status
        b += (size_t)Np * 2;                 // expday
        b += (size_t)Np;                     // col
        b += (size_t)(Np / 32) * 4;          // inf
        b += (size_t)(Np / 32) * 4 * 4;      // touch, dupm, mark_pre, mark_post bitmasks
    }
    b += 5 * 256 * 4 + 5 * 512;          // nb tables
    b += FUSED_EVCAP * 4 + FUSED_EVCAP * 2 * 2 + FUSED_NEDGE * 4;    // evy, evm, D, edges
    b += FUSED_LCAP * ((global_state ? 4 : 2) * 2 + 4 * 1 + 4 + 8 + 4 + 4 + 4) + (FUSED_LCAP + 1) * 4 + 12;   // list, cx, ch/cb/cc/cflag, hits, gp0, cbase, clen, gcum, pref
    b += DUR_CACHE * 8;                  // dcache
    if (global_state) b += FUSED_HCAP * 4;   // htab
    return (b + 15) / 16 * 16 + 64;
}

// get_nextday_counts :772-788 with the NB tables in shared memory
template <class MT> __device__ __forceinline__ int fresh_budget_s(const MT &M, uint2 key, uint32_t i, uint32_t day, uint32_t st) {
    if (st_health(st) == DED) return 0;
    uint32_t w = pick(philox(i >> 2, day, S_COUNT, 0, key), i & 3u);
    if (st & ST_ISOB) return (w < 0x80000000u) ? 0 : 1 + (int)(((unsigned long long)(w & 0x7fffffffu) * 3ull) >> 31);
    const int ag = st_ag(st);
    const uint32_t *thr = M.nb_thr + ag * 256;
    int k = M.nb_guide[ag * 512 + ((w >> 24) == 255u ? 256u + ((w >> 16) & 0xFFu) : (w >> 24))];
    const int len = M.nb_len[ag];
    while (k < len && thr[k] <= w) ++k;
    return k;
}

// gpw = Int.(round.(cm[ag]*cnts)) :804 (half-to-even) packed as five bytes; cm row in shared memory
__device__ __forceinline__ unsigned long long gpw_pack(const double *cm_row, int cnts) {
    unsigned long long gp = 0;
#pragma unroll
    for (int g = 0; g < 5; ++g) gp |= (unsigned long long)(unsigned)__double2int_rn(__dmul_rn(cm_row[g], (double)cnts)) << (8 * g);
    return gp;
}
__device__ __forceinline__ int gpw_sum(unsigned long long gp) {
    return (int)(gp & 0xFF) + (int)((gp >> 8) & 0xFF) + (int)((gp >> 16) & 0xFF) + (int)((gp >> 24) & 0xFF) + (int)((gp >> 32) & 0xFF);
}

// TMA bulk store of one hmatrix column from shared memory (UBLKCP in SASS)
__device__ __forceinline__ void bulk_store_g2s_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// The state matrix is written once and never read by the kernel: its lines are marked evict-first in L2 so that the 5 GB stream of
// a launch does not push out what the replicates keep re-reading through L2 (age-group tables, per-agent fields of transitions).
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, uint32_t bytes) {
    uint32_t s = (uint32_t)__cvta_generic_to_shared(ssrc);
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group.L2::cache_hint [%0], [%1], %2, %3;" ::"l"(gdst), "r"(s), "r"(bytes), "l"(pol) : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
// age-group index tables: read at random by every contact event for the whole life of the replicate -> evict-last in L2
__device__ __forceinline__ uint32_t ld_keep(const uint16_t *p) {
    uint64_t pol; unsigned short v;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm("ld.global.L2::cache_hint.u16 %0, [%1], %2;" : "=h"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ uint32_t ld_keep(const uint32_t *p) {
    uint64_t pol; uint32_t v;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm("ld.global.L2::cache_hint.u32 %0, [%1], %2;" : "=r"(v) : "l"(p), "l"(pol));
    return v;
}
__device__ __forceinline__ void bulk_store_wait_read_n() { asm volatile("cp.async.bulk.wait_group.read 6;" ::: "memory"); }   // bound the queue on idle days
__device__ __forceinline__ void bulk_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_store_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// CT = false: no replicate of the ensemble traces contacts (the tracing code is compiled out).
// GS = true: the per-agent state stays in global memory (L2) instead of shared memory — any popsize; no TMA column stores, no parking.
// state-matrix column without TMA (state in global memory, or a column that is not 16-byte aligned): 16 bytes per thread and pass
__device__ __forceinline__ void copy_column(uint8_t *dst, const uint8_t *src, int n, int tid, int nt) {
    if ((((uintptr_t)dst | (uintptr_t)src) & 15) == 0) {
        const int nv = n >> 4;
        for (int i = tid; i < nv; i += nt) reinterpret_cast<uint4 *>(dst)[i] = reinterpret_cast<const uint4 *>(src)[i];
        for (int i = (nv << 4) + tid; i < n; i += nt) dst[i] = src[i];
    } else {
        for (int i = tid; i < n; i += nt) dst[i] = src[i];
    }
}

template <bool CT, bool GS>
__global__ void __launch_bounds__(FUSED_NT, GS ? 1 : 2) k_sim_fused(Dev S, unsigned *work /* queue heads and tails, see below */, int split_day, unsigned tail_cta_limit) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5, NT = FUSED_NT, NW = FUSED_NT / 32;
    const int N = S.N, Np = S.Np, T = S.T, nwords = S.nwords;
    FusedCtl &Z = *reinterpret_cast<FusedCtl *>(smem_raw);
    using IdxT = typename std::conditional<GS, uint32_t, uint16_t>::type;
    FusedSmemT<IdxT> M;
    {
        unsigned char *p = smem_raw + sizeof(FusedCtl);
        M.dcache = (unsigned long long *)p; p += DUR_CACHE * 8;
        M.gp0 = (unsigned long long *)p; p += FUSED_LCAP * 8;
        M.htab = nullptr;
        if (GS) { M.htab = (uint32_t *)p; p += FUSED_HCAP * 4; }
        if (!GS) {
            M.status = (uint32_t *)p; p += (size_t)Np * 4;
            M.expday = (int16_t *)p; p += (size_t)Np * 2;      // Np is a multiple of 32: every array below stays 16-byte aligned
            M.col = (uint8_t *)p; p += (size_t)Np;
            M.inf = (uint32_t *)p; p += (size_t)nwords * 4;
            M.touch = (uint32_t *)p; p += (size_t)nwords * 4;
            M.dupm = (uint32_t *)p; p += (size_t)nwords * 4;
            M.mpre = (uint32_t *)p; p += (size_t)nwords * 4;
            M.mpost = (uint32_t *)p; p += (size_t)nwords * 4;
        } else {                                               // set per replicate (global memory)
            M.status = nullptr; M.expday = nullptr; M.col = nullptr; M.inf = nullptr; M.touch = M.dupm = M.mpre = M.mpost = nullptr;
        }
        M.nb_thr = (uint32_t *)p; p += 5 * 256 * 4;
        M.evy = (uint32_t *)p; M.cnt = (int *)p; p += FUSED_EVCAP * 4;
        M.edges = (uint32_t *)p; p += FUSED_NEDGE * 4;
        M.pref = (int *)p; p += (FUSED_LCAP + 1) * 4;
        M.hits = (int *)p; p += FUSED_LCAP * 4;
        M.cbase = (uint32_t *)p; p += FUSED_LCAP * 4;
        M.clen = (int *)p; p += FUSED_LCAP * 4;
        M.gcum = (uint32_t *)p; p += FUSED_LCAP * 4;
        M.list = (IdxT *)p; p += FUSED_LCAP * sizeof(IdxT);
        M.cx = (IdxT *)p; p += FUSED_LCAP * sizeof(IdxT);
        M.D = (uint16_t *)p; p += FUSED_EVCAP * 2;
        M.evm = (uint16_t *)p; p += FUSED_EVCAP * 2;
        M.nb_guide = (uint8_t *)p; p += 5 * 512;
        M.nb_len = Z.nb_len;
        M.ch = (uint8_t *)p; p += FUSED_LCAP;
        M.cb = (uint8_t *)p; p += FUSED_LCAP;
        M.cc = (uint8_t *)p; p += FUSED_LCAP;
        M.cflag = (uint8_t *)p; p += FUSED_LCAP;
    }

    for (int k = tid; k < 5 * 256; k += NT) M.nb_thr[k] = S.nb_thr[k];
    for (int k = tid; k < 5 * 512; k += NT) M.nb_guide[k] = S.nb_guide[k];
    if (tid < 5) Z.nb_len[tid] = c_tab.nb_len[tid];
    if (!GS) for (int k = tid; k < nwords; k += NT) { M.touch[k] = 0; M.dupm[k] = 0; M.mpre[k] = 0; M.mpost[k] = 0; }
    if (GS) for (int k = tid; k < FUSED_HCAP / 4; k += NT) reinterpret_cast<uint4 *>(M.htab)[k] = make_uint4(~0u, ~0u, ~0u, ~0u);
    if (tid == 0) Z.nedge = 0;
    if (tid < 4) { Z.tlat[tid] = c_tab.lat_thr[tid]; Z.tpre[tid] = c_tab.pre_thr[tid]; }
    if (tid < 40) { Z.tasy[tid] = c_tab.asy_thr[tid]; Z.tinf[tid] = c_tab.inf_thr[tid]; }
    if (tid < 25) Z.cm[tid] = c_tab.cm[tid / 5][tid % 5];
    int *const s_inc = Z.io[0][0], *const s_out = Z.io[0][1];   // first-column phase uses parity 0
    __syncthreads();
    const bool tma_ok = !GS && S.store_hm && (N % 16 == 0);     // cp.async.bulk needs a shared-memory source and 16-byte size and alignment

    // Two sweeps over an atomic work queue. Sweep 0 starts every replicate; with split_day > 0 a replicate that is still active
    // after split_day days is parked (state to HBM) and queued, so that the long epidemics — which bound the launch — all resume
    // early in sweep 1 instead of starting whenever the fresh queue happens to reach them. Everything else finishes in sweep 0.
    // Curve rows d0..T-1 of replicate r once nothing changes any more: incidence 0, prevalence = Z.run. The rows of a group are one
    // contiguous run of 12-int rows (48 B, 16-byte aligned), filled by the whole CTA with 16-byte stores.
    auto fill_idle_rows = [&](int r, int d0) {
        const int per = (T - d0) * 3;                                          // uint4 per group
        for (int idx = tid; idx < 6 * per; idx += NT) {
            const int G = idx / per, j = idx - G * per;
            const size_t b0 = (((size_t)r * 6 + G) * (size_t)T + (size_t)d0) * 12;
            reinterpret_cast<int4 *>(S.inc + b0)[j] = make_int4(0, 0, 0, 0);
            reinterpret_cast<int4 *>(S.prev + b0)[j] = *reinterpret_cast<const int4 *>(&Z.run[G * 12 + (j % 3) * 4]);
        }
    };
    // hmatrix columns d0..T-1 of replicate r = the column in M.col (TMA bulk copies queued by thread 0)
    auto fill_idle_columns = [&](uint8_t *hm, int d0) {
        for (int d = d0; d < T; ++d) {
            if (tma_ok) { if (tid == 0) { bulk_store_wait_read_n(); bulk_store(hm + (size_t)d * N, M.col, (uint32_t)N); } }
            else for (int i = tid; i < N; i += NT) hm[(size_t)d * N + i] = M.col[i];
        }
    };
    // replicate r's curves -> pinned host memory (Int16), 16-byte posted writes over PCIe. Rows d0..T-1 repeat (incidence 0,
    // prevalence of row d0-1): with a rows array registered (cabm_set_host_curve_rows) only the rows before them cross the link
    // and the host learns how many they are.
    auto stream_curves_to_host = [&](int r, int d0) {
        const size_t per = (size_t)6 * T * 12;                 // multiple of 8
        const int rows = (S.host_rows && (T & 1) == 0) ? min(T, (d0 + 1) & ~1) : T;      // even: every group's rows start and end on 16 bytes
        const int4 *gi = reinterpret_cast<const int4 *>(S.inc + (size_t)r * per), *gp = reinterpret_cast<const int4 *>(S.prev + (size_t)r * per);
        uint4 *hi = reinterpret_cast<uint4 *>(S.host_inc + (size_t)r * per), *hp = reinterpret_cast<uint4 *>(S.host_prev + (size_t)r * per);
        const int pg = rows * 12 / 8, full = T * 12 / 8;       // uint4 of Int16 per group: sent, and in the array (exact when rows < T: both even)
        const int nq = rows == T ? (int)(per / 8) : 6 * pg;
        for (int q = tid; q < nq; q += NT) {
            const int k = rows == T ? q : (q / pg) * full + q % pg;
            const int4 a = __ldcg(gi + 2 * k), b = __ldcg(gi + 2 * k + 1), c = __ldcg(gp + 2 * k), e = __ldcg(gp + 2 * k + 1);
            hi[k] = make_uint4((uint32_t)(a.x & 0xFFFF) | ((uint32_t)a.y << 16), (uint32_t)(a.z & 0xFFFF) | ((uint32_t)a.w << 16),
                               (uint32_t)(b.x & 0xFFFF) | ((uint32_t)b.y << 16), (uint32_t)(b.z & 0xFFFF) | ((uint32_t)b.w << 16));
            hp[k] = make_uint4((uint32_t)(c.x & 0xFFFF) | ((uint32_t)c.y << 16), (uint32_t)(c.z & 0xFFFF) | ((uint32_t)c.w << 16),
                               (uint32_t)(e.x & 0xFFFF) | ((uint32_t)e.y << 16), (uint32_t)(e.z & 0xFFFF) | ((uint32_t)e.w << 16));
        }
        if (S.host_rows && tid == 0) S.host_rows[r] = rows;
    };

    // Two sweeps over atomic work queues. Sweep 0 starts every replicate and runs it for split_day days (split_day = 0: to the end).
    // By then most epidemics have died out: their state is final, and what is left of them is pure output — the idle tail of the
    // curves and of the state matrix (a "fill job") and the copy of the curves to the host (a "copy job"). A replicate that is
    // still active is parked (state to HBM) and queued for sweep 1. Sweep 1 resumes the parked replicates first — they are the
    // long ones that bound the launch — and CTAs without one work the output jobs off meanwhile, so the output traffic (HBM,
    // PCIe) overlaps the latency-bound epidemics instead of delaying their start. A few CTAs saturate the PCIe link and more of
    // them only get in the way of the running epidemics, so at most `tail_cta_limit` CTAs become copiers (and stay so until
    // the copy queue has run dry); every other CTA leaves as soon as there is nothing else to do and frees its SM for the
    // next batch of the ensemble (cabm.h: several handles on one GPU).
    //   work[0] fresh head | [1],[2] parked head, tail | [3] fresh jobs finished | [4],[5] fill-job head, tail | [6] copiers
    //   [7] fill jobs finished | [8],[9] copy-job head, tail
    // resume_list holds the parked replicates from the bottom and the fill jobs from the top (each replicate is one or the other).
    bool copier = false;                                                        // (thread 0)
    for (int sweep = 0; sweep < 2; ++sweep) {
    if (sweep == 1 && split_day <= 0) break;
    for (;;) {
        __syncthreads();
        if (tid == 0) {
            int got = -1, kind = 0;
            if (sweep == 0) { const unsigned k = atomicAdd(&work[0], 1u); if (k < (unsigned)S.R) got = (int)k; }
            else {
                for (;;) {          // wait while jobs that might still queue something are in flight
                    const unsigned done = *(volatile unsigned *)&work[3];
                    __threadfence();
                    const unsigned t = *(volatile unsigned *)&work[2], h = *(volatile unsigned *)&work[1];
                    if (h < t && !copier) {                                  // (a copier would sit on its ticket for a whole epidemic)
                        if (atomicCAS(&work[1], h, h + 1) == h) {
                            while ((got = *(volatile int *)&S.resume_list[h]) < 0) __nanosleep(100);
                            break;
                        }
                        continue;
                    }
                    const unsigned filled = *(volatile unsigned *)&work[7];
                    __threadfence();
                    const unsigned ft = *(volatile unsigned *)&work[5], fh = *(volatile unsigned *)&work[4];
                    if (S.host_inc) {
                        const unsigned ct = *(volatile unsigned *)&work[9], ch = *(volatile unsigned *)&work[8];
                        if (ch < ct) {
                            if (!copier && *(volatile unsigned *)&work[6] < tail_cta_limit) {
                                if (atomicAdd(&work[6], 1u) < tail_cta_limit) copier = true; else atomicSub(&work[6], 1u);
                            }
                            if (copier) {
                                if (atomicCAS(&work[8], ch, ch + 1) == ch) {
                                    while ((got = *(volatile int *)&S.copy_list[ch]) < 0) __nanosleep(100);
                                    kind = 2;
                                    break;
                                }
                                continue;
                            }
                        } else if (copier && done >= (unsigned)S.R && filled >= ft) {      // nothing can queue a copy job any more
                            atomicSub(&work[6], 1u); copier = false;
                        }
                    }
                    if (fh < ft) {
                        if (atomicCAS(&work[4], fh, fh + 1) == fh) {
                            while ((got = *(volatile int *)&S.resume_list[S.R - 1 - (int)fh]) < 0) __nanosleep(100);
                            kind = 1;
                            break;
                        }
                        continue;
                    }
                    if (done >= (unsigned)S.R && !copier) break;
                    __nanosleep(500);
                }
                __threadfence();
            }
            Z.r = got; Z.kind = kind;
            if (tma_ok) bulk_store_wait_read();                                  // column copies of the previous job have left shared memory
        }
        __syncthreads();
        const int r = Z.r;
        if (r < 0) break;
        if (GS) {           // this replicate's per-agent arrays (the host zeroed the bitmasks before the launch)
            const size_t rb_ = (size_t)r * Np, rw_ = (size_t)r * nwords;
            M.status = S.status + rb_; M.expday = S.expday + rb_; M.col = S.col_g + rb_; M.inf = S.infmask + rw_;
            M.mpre = S.mark_pre + rw_; M.mpost = S.mark_post + rw_;     // (touch/dupm: the chunk's target table M.htab takes their place)
        }
        if (Z.kind == 2) {
            // -------------------------------------------------------- copy job: a finished replicate's curves -> pinned host memory
            stream_curves_to_host(r, __ldcg(&S.resume_aux[r * 2]));
            continue;
        }
        if (Z.kind == 1) {
            // -------------------------------------------------------- fill job: the idle tail of a replicate that died out
            const int d0 = __ldcg(&S.resume_aux[r * 2]);                       // first row that is not written yet (>= 1)
            uint8_t *hm = S.store_hm ? S.hm + (size_t)r * T * N : nullptr;
            if (tid < 72) Z.run[tid] = __ldcg(&S.prev[(((size_t)r * 6 + tid / 12) * (size_t)T + (size_t)(d0 - 1)) * 12 + tid % 12]);
            if (hm) {                                                          // the column of day d0 - 1 repeats to the end
                const uint8_t *src = hm + (size_t)(d0 - 1) * N;
                if (tma_ok) for (int k = tid; k < N / 16; k += NT) reinterpret_cast<uint4 *>(M.col)[k] = __ldcg(reinterpret_cast<const uint4 *>(src) + k);
                else for (int i = tid; i < N; i += NT) M.col[i] = __ldcg(src + i);
                if (tma_ok) bulk_store_g2s_fence();
            }
            __syncthreads();
            fill_idle_rows(r, d0);
            if (hm) fill_idle_columns(hm, d0);
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                if (S.host_inc) {
                    const unsigned slot = atomicAdd(&work[9], 1u);
                    __threadfence();
                    *(volatile int *)&S.copy_list[slot] = r;
                    __threadfence();
                }
                atomicAdd(&work[7], 1u);
            }
            continue;
        }
        const bool resumed = sweep == 1;
#ifdef CABM_FUSED_DIAG      // per-phase cycle counts seen by thread 0 (make diag; tools/diag_fused.py)
        const long long t_start = clock64();
        __shared__ long long t_ph[32], t_wdt[8];
        long long t_last = t_start;
        if (tid < 32) t_ph[tid] = 0;
#define PHASE(k) do { if (tid == 0) { t_ph[k] += clock64() - t_last; t_last = clock64(); } } while (0)
#define PHASEX(k) do { if (tid == FUSED_NT - 32) { t_ph[k] += clock64() - t_lastx; t_lastx = clock64(); } } while (0)
        long long t_lastx = 0;
        int n_rounds = 0;
#else
#define PHASE(k) do { } while (0)
#define PHASEX(k) do { } while (0)
#endif
        if (tid < (int)(sizeof(DevPset) / 4)) reinterpret_cast<uint32_t *>(&Z.P)[tid] = reinterpret_cast<const uint32_t *>(&S.psets[S.pset_of[r]])[tid];
        __syncthreads();
        const DevPset &P = Z.P;
        const bool ctr = CT && P.ctstrat != 0;        // this replicate traces contacts
        if (tid == 0) { Z.logn = 0; Z.qcnt[0] = 0; Z.qcnt[1] = 0; }
        const uint2 key = S.key[r];
        const size_t rb = (size_t)r * Np;
        IdxT *grp;                                      // get_ag_dist :339-343 index tables of this replicate (global memory, read through L2/L1)
        if constexpr (GS) grp = S.grp_idx + rb; else grp = S.grp16 + rb;
        const unsigned long long *beta = S.beta_thr + (size_t)S.pset_of[r] * T * 4;
        uint8_t *hm = S.store_hm ? S.hm + (size_t)r * T * N : nullptr;

        int run = 0;
        long long latsum = 0;
        int g0_in = 0;
        if (!resumed) {
        // ------------------------------------------------------------ initialize :171-188
        for (int i = tid; i < Np; i += NT) {
            const size_t a = rb + i;
            uint32_t st = st_make(SUS, UNDEF, 0, false), d4 = 0;
            uint8_t via = VIA_NULL;
            if (i < N) {
                const uint4 wa = philox(i, 0, S_INIT, 0, key), wb = philox(i, 0, S_INIT, 1, key);
                int ag0 = 0;
#pragma unroll
                for (int k = 0; k < 4; ++k) ag0 += ((unsigned long long)wa.x >= P.age_thr[k]);
                int latents = table_draw(wa.y, c_tab.lat_base, Z.tlat, c_tab.lat_len);
                const int pres = table_draw(wa.z, c_tab.pre_base, Z.tpre, c_tab.pre_len);
                const int asymps = table_draw_bs(wa.w, c_tab.asy_base, Z.tasy, c_tab.asy_len);
                const int infs = table_draw_bs(wb.x, c_tab.inf_base, Z.tinf, c_tab.inf_len);
                latents -= pres;
                d4 = (uint32_t)latents | ((uint32_t)asymps << 8) | ((uint32_t)pres << 16) | ((uint32_t)infs << 24);
                const bool quar = bern(wb.y, P.quar_thr) && ag0 == P.quar_grp0;
                st = st_make(SUS, UNDEF, ag0, quar);
                via = quar ? VIA_QU : VIA_NULL;
            }
            M.status[i] = st;
            M.expday[i] = (i < N) ? (int16_t)999 : (int16_t)32767;
            M.col[i] = SUS;
            S.dur[a] = d4; S.isovia[a] = via; S.entry[a] = 0; S.sickfrom[a] = UNDEF; S.latday[a] = -999;
            if (ctr) S.texp[a] = (i < N) ? (int16_t)999 : (int16_t)32767;
            S.tracestart[a] = -1; S.traceend[a] = -1; S.txpday[a] = 0; S.ct_head[a] = CT_NONE;
        }
        for (int k = tid; k < nwords; k += NT) M.inf[k] = 0;
        for (int k = tid; k < DUR_CACHE; k += NT) M.dcache[k] = ~0ull;
        for (int k = tid; k < 240; k += NT) (&Z.io[0][0][0])[k] = 0;
        if (tid < 4) Z.infectors[tid] = 0;
        if (tid == 0) Z.ninf = 0;
        __syncthreads();
        PHASE(22);

        // ------------------------------------------------------------ get_ag_dist :339-343 (ordered counting sort)
        // each warp owns a contiguous range of agents: count per group, one prefix over the warps, then every warp scatters
        // its range in ascending order with warp-local bases — two block barriers in total
        {
            const int per_warp = ((N + NW - 1) / NW + 31) & ~31, lo = wid * per_warp, hi = min(N, lo + per_warp);
            int c[5] = {0, 0, 0, 0, 0};
            for (int i0 = lo; i0 < hi; i0 += 32) {
                const int i = i0 + lane;
                const int ag = (i < hi) ? st_ag(M.status[i]) : -1;
#pragma unroll
                for (int g = 0; g < 5; ++g) c[g] += __popc(__ballot_sync(FULL, ag == g));
            }
            if (lane == 0) for (int g = 0; g < 5; ++g) Z.wcnt[wid][g] = c[g];
            __syncthreads();
            if (tid == 0) {
                int o = 0;
                for (int g = 0; g < 5; ++g) {
                    Z.off[g] = o;
                    for (int w = 0; w < NW; ++w) { const int v = Z.wcnt[w][g]; Z.wcnt[w][g] = o; o += v; }      // start of warp w's slice of group g
                }
                Z.off[5] = o;
            }
            __syncthreads();
            int base[5];
#pragma unroll
            for (int g = 0; g < 5; ++g) base[g] = Z.wcnt[wid][g];
            for (int i0 = lo; i0 < hi; i0 += 32) {
                const int i = i0 + lane;
                const int ag = (i < hi) ? st_ag(M.status[i]) : -1;
#pragma unroll
                for (int g = 0; g < 5; ++g) {
                    const unsigned m = __ballot_sync(FULL, ag == g);
                    if (ag == g) grp[base[g] + __popc(m & ((1u << lane) - 1u))] = (IdxT)i;
                    base[g] += __popc(m);
                }
            }
        }
        __syncthreads();
        PHASE(23);

        // ------------------------------------------------------------ insert_infected :360-395 (PRE seeds, then REC seeds)
        for (int which = 0; which < 2; ++which) {
            if (which == 1 && P.calibration) break;                                          // :116-121
            const int health = which == 0 ? PRE : REC, num = which == 0 ? P.initialinf : P.initialhi;
            if (which == 1 && num == 0 && N > P.initialinf) continue;                          // sample(l, 0) with l non-empty: nothing to do
            const int seg = (N + NT - 1) / NT, lo = tid * seg, hi = min(N, lo + seg);
            int mine = 0;
            for (int i = lo; i < hi; ++i) mine += (st_health(M.status[i]) == SUS);
            M.cnt[tid] = mine;
            { int v = mine; for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(FULL, v, o); if (lane == 0) Z.wsum[wid] = v; }
            __syncthreads();
            if (tid == 0) { int L = 0; for (int w = 0; w < NW; ++w) L += Z.wsum[w]; Z.Ltot = L; }
            __syncthreads();
            if (Z.Ltot == 0 || num > Z.Ltot) { if (tid == 0 && (num > 0 || Z.Ltot == 0)) atomicOr(&S.err[r], ERR_NO_SUS); continue; }
            for (int j = 0; j < num; ++j) {
                if (tid == 0) {
                    const uint4 w = philox(j, 0, S_SEED, which, key);
                    int k = below(w.x, Z.Ltot - j), t = 0;
                    while (k >= M.cnt[t]) { k -= M.cnt[t]; ++t; }
                    Z.sel_tid = t; Z.sel_k = k;
                }
                __syncthreads();
                auto seed_agent = [&](int i, uint32_t st) {
                    const Tr t = transition_core(S, P, r, i, 0, st, key, health, true, M.dcache);
                    M.status[i] = (t.st & ~ST_ISOB) | (st & ST_ISOB);    // day-1 budget was drawn with the old iso (:186)
                    M.expday[i] = (int16_t)t.expday;
                    if (ctr) S.texp[rb + i] = (int16_t)t.expday;
                    M.col[i] = (uint8_t)t.newh;
                    S.sickfrom[rb + i] = INF;                             // :391
                    if (is_infectious(t.newh)) { M.inf[i >> 5] |= 1u << (i & 31); Z.ninf += 1; }
                };
                if constexpr (GS) {
                    // state in global memory: the selected thread's segment (thousands of agents) is searched by the whole CTA — each warp
                    // counts the susceptibles of its slice with coalesced loads, the warp whose slice holds the k-th one locates it
                    const int slo = Z.sel_tid * seg, shi = min(N, slo + seg);
                    const int per = (((shi - slo + NW - 1) / NW) + 31) & ~31;
                    const int wlo = min(shi, slo + wid * per), whi = min(shi, wlo + per);
                    int c = 0;
                    for (int i0 = wlo; i0 < whi; i0 += 32) { const int i = i0 + lane; c += (i < whi && st_health(M.status[i]) == SUS); }
                    for (int o = 16; o; o >>= 1) c += __shfl_xor_sync(FULL, c, o);
                    if (lane == 0) Z.wsum2[wid] = c;
                    __syncthreads();
                    int k = Z.sel_k, w = 0;
                    while (w < NW - 1 && k >= Z.wsum2[w]) { k -= Z.wsum2[w]; ++w; }
                    if (wid == w) {
                        for (int i0 = wlo; i0 < whi; i0 += 32) {
                            const int i = i0 + lane;
                            const uint32_t st = i < whi ? M.status[i] : 0u;
                            const unsigned b = __ballot_sync(FULL, i < whi && st_health(st) == SUS);
                            const int n = __popc(b);
                            if (k < n) { if (lane == (int)__fns(b, 0, k + 1)) seed_agent(i, st); break; }
                            k -= n;
                        }
                        if (lane == 0) M.cnt[Z.sel_tid] -= 1;
                    }
                } else if (tid == Z.sel_tid) {
                    int k = Z.sel_k;
                    for (int i = lo; i < hi; ++i) {
                        const uint32_t st = M.status[i];
                        if (st_health(st) != SUS) continue;
                        if (k-- > 0) continue;
                        seed_agent(i, st);
                        break;
                    }
                    M.cnt[tid] -= 1;
                }
                __syncthreads();
            }
        }
        __syncthreads();
        PHASE(24);

        // ------------------------------------------------------------ column 1 (:136 at st = 1) and its curve entries
        // every agent first occurs in its initial state on day 1 (:270-280); nearly all are SUS, so count the others
        for (int i = tid; i < N; i += NT) {
            const uint32_t st = M.status[i];
            const int h = st_health(st);
            if (h != SUS) { atomicAdd(&s_inc[st_ag(st) * 12 + h], 1); atomicAdd(&s_out[st_ag(st) * 12 + SUS], 1); }
        }
        __syncthreads();
        run = 0;                                  // running prevalence of (group, state) = (tid/12, tid%12), tid < 60; group 0 in tid 64..75
        if (tid < 60) {
            const int g = tid / 12, s = tid % 12;
            run = s_inc[tid] + (s == SUS ? (Z.off[g + 1] - Z.off[g]) - s_out[tid] : 0);
            const size_t b = (((size_t)r * 6 + (g + 1)) * (size_t)T + 0) * 12 + s;
            S.inc[b] = run; S.prev[b] = run;
            M.cnt[tid] = run;
        }
        __syncthreads();
        if (tid >= 64 && tid < 76) {
            const int s = tid - 64;
            for (int g = 0; g < 5; ++g) run += M.cnt[g * 12 + s];
            const size_t b = (((size_t)r * 6 + 0) * (size_t)T + 0) * 12 + s;
            S.inc[b] = run; S.prev[b] = run;
            if (s == LAT) latsum = run;
        }
        __syncthreads();
        if (tid < 60) { s_inc[tid] = 0; s_out[tid] = 0; }
        if (hm) {
            if (tma_ok) { bulk_store_g2s_fence(); __syncthreads(); if (tid == 0) bulk_store(hm, M.col, (uint32_t)N); }
            else copy_column(hm, M.col, N, tid, NT);
        }
        __syncthreads();

        } else {
            // ---------------------------------------------------------- resume a parked replicate: reload the shared-memory state
            for (int i = tid; i < Np; i += NT) {
                const uint32_t st = __ldcg(&S.status[rb + i]);           // written by another CTA of this launch: read through L2
                M.status[i] = st; M.expday[i] = __ldcg(&S.expday[rb + i]); M.col[i] = (uint8_t)st_health(st);
            }
            for (int k = tid; k < nwords; k += NT) M.inf[k] = __ldcg(&S.infmask[(size_t)r * nwords + k]);
            for (int k = tid; k < DUR_CACHE; k += NT) M.dcache[k] = ~0ull;
            for (int k = tid; k < 240; k += NT) (&Z.io[0][0][0])[k] = 0;
            if (tid < 6) Z.off[tid] = __ldcg(&S.grp_off[r * 8 + tid]);
            if (tid < 4) Z.infectors[tid] = 0;
            if (tid == 0) {
                Z.ninf = __ldcg(&S.resume_aux[r * 2 + 1]);
                if (ctr) { Z.logn = __ldcg(&S.ct_cnt[r]); Z.qcnt[0] = __ldcg(&S.ident_cnt[r * 2]); Z.qcnt[1] = __ldcg(&S.ident_cnt[r * 2 + 1]); }
            }
            if (tid < 60) run = __ldcg(&S.prev[(((size_t)r * 6 + (tid / 12 + 1)) * (size_t)T + (size_t)split_day) * 12 + tid % 12]);
            else if (tid >= 64 && tid < 76) run = __ldcg(&S.prev[(((size_t)r * 6 + 0) * (size_t)T + (size_t)split_day) * 12 + (tid - 64)]);
            if (tid == 64 + LAT) latsum = __ldcg(&S.latsum[r]);
            __syncthreads();
        }
        PHASE(0);
        // ============================================================ the day loop :131-140
        // s_nextfire = earliest expiry day of any agent. A day with nobody infectious and nothing expiring changes no
        // state: it only repeats the column and the prevalence row, without a single barrier.
        if (tid < 2) Z.nf[tid] = 32767;
        if (tid == 0) Z.evn = 0;
        int nextfire = resumed ? __ldcg(&S.resume_aux[r * 2]) : 0, act = 0;    // uniform registers
        const int day_first = resumed ? split_day + 1 : 1;
        const int day_last = (!resumed && split_day > 0 && split_day < T) ? split_day : T;
        bool finished = day_last == T;                                         // set early by the fast-forward below
        int tail_day = 0;                                                      // > 0: rows tail_day..T-1 are left to a tail job
        int idle_from = 0;                                                     // > 0: rows idle_from..T-1 were filled here as an idle tail
        __syncthreads();
        for (int day = day_first; day <= day_last; ++day) {
            const bool active = Z.ninf > 0 || day >= nextfire;                 // uniform: Z.ninf only changes between S1 and S2
            if (!active && nextfire > T) {
                // nobody is infectious and no state expires before the end of the run: the remaining columns and curve rows
                // repeat the current ones (tis >= exp :405 never fires again) — write them without stepping through the days
                // curve rows day..T-1 of a group are one contiguous run of 12-int rows (48 B, 16-byte aligned): the whole CTA
                // fills them with 16-byte stores while thread 0 queues the column copies
                finished = true;
                if (!resumed && split_day > 0) { tail_day = day; break; }       // sweep 0: the tail is somebody else's job (see above)
                idle_from = day;
                if (tid < 60) Z.run[12 + tid] = run; else if (tid >= 64 && tid < 76) Z.run[tid - 64] = run;
                __syncthreads();
                fill_idle_rows(r, day);
                if (hm) fill_idle_columns(hm, day);
                break;
            }
            if (active) {
                if (hm && tma_ok && tid == 0) bulk_store_wait_read();          // earlier columns have left shared memory (ordered by the barriers below)
                // ---------------------------------------------------- dyntrans(st) :790-834
                const bool had_infectious = Z.ninf > 0;                        // uniform
                if (had_infectious) {
                    if (tid < 4) Z.beta[tid] = beta[(size_t)(day - 1) * 4 + tid];              // _get_betavalue :756-769 for today
                    int cursor = 0;                                                            // first agent index not yet listed
                    for (;;) {
                        // ordered list of (up to FUSED_LCAP) infectious agents at or after `cursor` (:794), built by the whole CTA
                        int lbase = 0, w0 = cursor >> 5;
                        for (; w0 < nwords && lbase < FUSED_LCAP; w0 += 2 * NT) {        // two mask words per thread: 10,000 agents in one pass
                            const int wi = w0 + 2 * tid;
                            uint32_t m0 = (wi < nwords) ? M.inf[wi] : 0u, m1 = (wi + 1 < nwords) ? M.inf[wi + 1] : 0u;
                            if (wi == (cursor >> 5)) m0 &= ~((1u << (cursor & 31)) - 1u);
                            const int c0 = __popc(m0), c = c0 + __popc(m1);
                            // exclusive prefix of c over the lanes, bit-sliced: one ballot per bit that occurs (c is 0..2 nearly always)
                            const int cmax = __reduce_max_sync(FULL, c);
                            int excl = 0, wtot = 0;
                            for (int b = 0; (cmax >> b) != 0; ++b) {
                                const unsigned bm = __ballot_sync(FULL, (c >> b) & 1);
                                excl += __popc(bm & ((1u << lane) - 1u)) << b; wtot += __popc(bm) << b;
                            }
                            if (lane == 0) Z.wsum[wid] = wtot;
                            __syncthreads();
                            int pos = lbase + excl, tot = 0;
                            for (int w = 0; w < NW; ++w) { const int v = Z.wsum[w]; tot += v; if (w < wid) pos += v; }
                            while (m0 && pos < FUSED_LCAP) { const int b = __ffs(m0) - 1; m0 &= m0 - 1; M.list[pos++] = (IdxT)((wi << 5) + b); }
                            while (m1 && pos < FUSED_LCAP) { const int b = __ffs(m1) - 1; m1 &= m1 - 1; M.list[pos++] = (IdxT)(((wi + 1) << 5) + b); }
                            lbase += tot;
                            __syncthreads();
                        }
                        const int nlist = min(lbase, FUSED_LCAP);
                        PHASE(1);
                        if (nlist == 0) break;
                        int lpos = 0, lim = FUSED_LCAP;       // lim: list entries the next chunk may take (halved when its DAG has too many edges)
                        while (lpos < nlist) {
#ifdef CABM_FUSED_DIAG
                            ++n_rounds;
#endif
                            // ---- (a) chunk = the longest stretch of the list from lpos whose events fit FUSED_EVCAP. One entry per thread:
                            //      today's budget (:802; hits by earlier chunks show up as tag == day), its split over the age groups (:804)
                            const int li = lpos + tid;
                            const bool have = li < nlist && tid < lim;
                            int E = 0, cnts = 0, x = 0, hx = 0; unsigned long long gp = 0; bool tr = false;
                            if (have) {
                                x = M.list[li];
                                const uint32_t sx = M.status[x];
                                cnts = fresh_budget_s(M, key, x, (uint32_t)day, sx);
                                if (st_tag(sx) == day) cnts = max(0, cnts - st_rem(sx));       // contacts received earlier today (:813) came off it
                                gp = gpw_pack(Z.cm + st_ag(sx) * 5, cnts);
                                E = gpw_sum(gp);
                                hx = st_health(sx) | (st_ag(sx) << 4) | ((sx & ST_TRACING) ? 0x80 : 0);
                                tr = ctr && (sx & ST_TRACING) && E > 0;                        // :817-819 logs its contacts: chunk [next, count, <= E targets]
                            }
                            // inclusive scan over the threads of (events [0:16) | entries with events [16:32)) and of the contact-log words
                            int v = E | ((E > 0) << 16), v2 = tr ? 2 + E : 0;
                            for (int o = 1; o < 32; o <<= 1) {
                                const int a = __shfl_up_sync(FULL, v, o), b = __shfl_up_sync(FULL, v2, o);
                                if (lane >= o) { v += a; v2 += b; }
                            }
                            if (lane == 31) { Z.wsum[wid] = v; Z.wsum2[wid] = v2; }
                            __syncthreads();                                                   // C0
                            for (int w = 0; w < wid; ++w) { v += Z.wsum[w]; v2 += Z.wsum2[w]; }
                            const int inclE = v & 0xFFFF, nkept = v >> 16;
                            const bool fits = have && inclE <= FUSED_EVCAP;                    // a prefix of the threads (the first entry always fits)
                            if (fits && E > 0) {
                                const int q = nkept - 1;
                                M.cx[q] = (IdxT)x; M.ch[q] = (uint8_t)hx; M.cb[q] = (uint8_t)cnts; M.cc[q] = (uint8_t)cnts; M.gp0[q] = gp;
                                {
                                    const uint32_t c1 = (uint32_t)(gp & 0xFF), c2 = c1 + (uint32_t)((gp >> 8) & 0xFF), c3 = c2 + (uint32_t)((gp >> 16) & 0xFF), c4 = c3 + (uint32_t)((gp >> 24) & 0xFF);
                                    M.gcum[q] = c1 | (c2 << 8) | (c3 << 16) | (c4 << 24);
                                }
                                M.pref[q + 1] = inclE;
                                for (int k = inclE - E; k < inclE; ++k) M.evm[k] = (uint16_t)(q << 2);      // owner of each of its events (read back by (b))
                                if (ctr) { M.cflag[q] = tr ? 1 : 0; M.cbase[q] = (uint32_t)(v2 - (tr ? 2 + E : 0)); M.clen[q] = 0; }
                            }
                            {
                                const unsigned fm = __ballot_sync(FULL, fits);
                                if (lane == 0) Z.wfit[wid] = __popc(fm);
                                if (fm && lane == 31 - __clz(fm)) { Z.wkept[wid] = nkept; Z.wneed[wid] = v2; }   // the warp's last fitting entry carries its totals
                            }
                            if (tid == 0) M.pref[0] = 0;
                            __syncthreads();                                                   // C1
                            int mlist = 0;
#pragma unroll
                            for (int w = 0; w < NW; ++w) mlist += Z.wfit[w];                     // fitting entries are a prefix of the threads
                            const int wl = (mlist - 1) >> 5;                                    // warp of the last fitting entry (mlist >= 1)
                            const int nb = Z.wkept[wl], Etot = nb ? M.pref[nb] : 0;
                            const uint32_t base0 = Z.logn;
                            const int needtot = ctr ? Z.wneed[wl] : 0;
                            const bool room = ctr && base0 + (uint32_t)needtot <= S.ct_cap;
                            if (ctr && !room && needtot && tid == 0) atomicOr(&S.err[r], ERR_CT_LOG_FULL);
                            PHASE(8);
                            if (nb == 0) {                                                     // nobody in this stretch of the list has contacts today
                                __syncthreads();                                               // (wfit/wkept are rewritten by the next chunk)
                                lpos += mlist; lim = FUSED_LCAP;
                                continue;
                            }
                            // ---- (b) all events of the chunk, in event order: warp w draws windows [win0, win1) of 32 consecutive events,
                            //      two windows at a time so that their (independent) draw chains overlap. The owner (chunk entry) of every
                            //      event was written into its record by the entry's thread during the chunk build.
                            const int nwin = (Etot + 31) >> 5, wpw = (nwin + NW - 1) / NW;
                            const int win0 = wid * wpw, win1 = min(nwin, win0 + wpw);
                            const int xlast = M.cx[nb - 1];
                            // draw of one event (branch-free: lanes past the end repeat the last event and are masked by `ok`)
                            auto draw = [&](int e, int q, int &idx, uint32_t &meta, bool &ok) {
                                const int qc = min(q, nb - 1), ec = min(e, Etot - 1);
                                const int el = ec - M.pref[qc];
                                const uint32_t pc = M.gcum[qc];                                 // cumulative round.(cm[ag]*cnts) :804 over the groups, one byte each
                                const int g = (el >= (int)(pc & 0xFF)) + (el >= (int)((pc >> 8) & 0xFF)) + (el >= (int)((pc >> 16) & 0xFF)) + (el >= (int)(pc >> 24));
                                const int k = el - (int)((((unsigned long long)pc << 8) >> (8 * g)) & 0xFF);
                                const int go = Z.off[g], len = Z.off[g + 1] - go;
                                const uint32_t hq = M.ch[qc];
                                const uint4 w = philox(M.cx[qc], day, S_DT, ((uint32_t)g << 16) | (uint32_t)k, key);
                                idx = min(go + below(w.x, len), N - 1);                         // :807
                                meta = (bern(w.y, Z.beta[beta_class(hq & 0xF)]) ? (uint32_t)EVM_SUCC : 0u) | ((uint32_t)qc << 2) | ((uint32_t)g << 26) | ((uint32_t)(len == 0) << 31);   // :800, :823
                                ok = e < Etot && len > 0;
                            };
                            // DAG edge of a drawn event whose target is a later infector of this chunk: that infector starts its turn with a
                            // smaller budget (:802, :813)
                            auto edge = [&](int e, uint32_t y, uint32_t meta) {
                                const int q = (int)((meta >> 2) & 255u);
                                int a = q + 1, b = nb - 1;
                                while (a < b) { const int mid = (a + b) >> 1; if ((int)M.cx[mid] < (int)y) a = mid + 1; else b = mid; }
                                if ((int)M.cx[a] == (int)y) {                                    // (a listed infector without events has nothing to lose)
                                    const int g = (int)((meta >> 26) & 7u), el = e - M.pref[q];
                                    const int k = el - (int)((((unsigned long long)M.gcum[q] << 8) >> (8 * g)) & 0xFF);
                                    const int slot = atomicAdd(&Z.nedge, 1);
                                    if (slot < FUSED_NEDGE) M.edges[slot] = (uint32_t)e | ((uint32_t)a << 11) | ((uint32_t)g << 19) | ((uint32_t)k << 22);
                                }
                            };
                            // NWIN windows of 32 events at once: every stage is a fully unrolled loop over the windows, so that the
                            // independent chains (Philox, table walks, loads) of the windows interleave in one basic block
                            auto windows = [&](int win, auto nwc) {
                                constexpr int NWIN = decltype(nwc)::value;
                                int q[NWIN], idx[NWIN]; uint32_t meta[NWIN], y[NWIN]; bool ok[NWIN];
#pragma unroll
                                for (int j = 0; j < NWIN; ++j) q[j] = (int)(M.evm[min(((win + j) << 5) + lane, Etot - 1)] >> 2);   // written by the chunk build
                                PHASE(2);
#pragma unroll
                                for (int j = 0; j < NWIN; ++j) draw(((win + j) << 5) + lane, q[j], idx[j], meta[j], ok[j]);
#pragma unroll
                                for (int j = 0; j < NWIN; ++j) y[j] = ld_keep(grp + idx[j]);       // loads in flight together
                                PHASE(7);
                                // (the target's status word is not touched here: its count of today's contacts starts when the first
                                // event is applied, step (d), where a tag older than today reads as zero contacts)
                                // records; targets hit more than once (by any event of the chunk, dropped later or not) take the ordered path
                                if constexpr (GS) {
                                    // state in global memory: which targets the chunk hits more than once is found in a table of the chunk's
                                    // targets in shared memory (slot = target, bit 31 = hit again) instead of two population-wide bitmasks
                                    // in global memory (an atomic, a load and two clearing stores per event)
                                    uint32_t slot[NWIN]; bool pend[NWIN];
#pragma unroll
                                    for (int j = 0; j < NWIN; ++j) { slot[j] = ((y[j] * 2654435761u) >> 16) & (uint32_t)(FUSED_HCAP - 1); pend[j] = ok[j]; }
                                    for (bool any = true; any;) {                                              // one probe of every unfinished window per round
                                        any = false;
#pragma unroll
                                        for (int j = 0; j < NWIN; ++j) {
                                            if (pend[j]) {
                                                const uint32_t old = atomicCAS(&M.htab[slot[j]], 0xFFFFFFFFu, y[j]);
                                                if (old == 0xFFFFFFFFu) pend[j] = false;                        // first event on this target
                                                else if ((old & 0x7FFFFFFFu) == y[j]) { if (!(old >> 31)) atomicOr(&M.htab[slot[j]], 0x80000000u); pend[j] = false; }
                                                else slot[j] = (slot[j] + 1u) & (uint32_t)(FUSED_HCAP - 1);
                                                any |= pend[j];
                                            }
                                        }
                                    }
#pragma unroll
                                    for (int j = 0; j < NWIN; ++j) {
                                        const int e = ((win + j) << 5) + lane;
                                        if (ok[j]) M.D[e] = (uint16_t)slot[j];                                  // read back by the apply pass (D is free until the compaction)
                                        M.evy[e] = ok[j] ? y[j] : EV_NONE;
                                        M.evm[e] = ok[j] ? (uint16_t)(meta[j] & 0x3FFu & ~(uint32_t)EVM_DROPPED) : (uint16_t)0;
                                        if (!ok[j] && (meta[j] >> 31) && e < Etot) atomicOr(&S.err[r], ERR_EMPTY_GROUP);
                                        if (ok[j] && (int)y[j] > (int)M.cx[(meta[j] >> 2) & 255u] && (int)y[j] <= xlast && ((M.inf[y[j] >> 5] >> (y[j] & 31)) & 1u))
                                            edge(e, y[j], meta[j]);
                                    }
                                } else {
                                    uint32_t old[NWIN];
#pragma unroll
                                    for (int j = 0; j < NWIN; ++j) old[j] = ok[j] ? atomicOr(&M.touch[y[j] >> 5], 1u << (y[j] & 31)) : 0u;
#pragma unroll
                                    for (int j = 0; j < NWIN; ++j) {
                                        const int e = ((win + j) << 5) + lane;
                                        M.evy[e] = ok[j] ? y[j] : EV_NONE;
                                        M.evm[e] = ok[j] ? (uint16_t)(meta[j] & 0x3FFu & ~(uint32_t)EVM_DROPPED) : (uint16_t)0;
                                        if (!ok[j] && (meta[j] >> 31) && e < Etot) atomicOr(&S.err[r], ERR_EMPTY_GROUP);
                                    }
#pragma unroll
                                    for (int j = 0; j < NWIN; ++j) {
                                        if (ok[j] && (old[j] >> (y[j] & 31)) & 1u) atomicOr(&M.dupm[y[j] >> 5], 1u << (y[j] & 31));
                                        // a later infector of this chunk? the index range first: with a large population hardly any target is inside it
                                        if (ok[j] && (int)y[j] > (int)M.cx[(meta[j] >> 2) & 255u] && (int)y[j] <= xlast && ((M.inf[y[j] >> 5] >> (y[j] & 31)) & 1u))
                                            edge(((win + j) << 5) + lane, y[j], meta[j]);
                                    }
                                }
                                PHASE(21);
                            };
                            {
                                int win = win0;
                                if constexpr (GS) {     // every load of a target is an L2 round trip: eight windows in flight per warp
                                    for (; win + 8 <= win1; win += 8) windows(win, std::integral_constant<int, 8>());
                                    if (win + 4 <= win1) { windows(win, std::integral_constant<int, 4>()); win += 4; }
                                }
                                for (; win + 2 <= win1; win += 2) windows(win, std::integral_constant<int, 2>());
                                if (win < win1) windows(win, std::integral_constant<int, 1>());
                            }
                            PHASE(9);
                            __syncthreads();                                                   // C2
                            const int ne = Z.nedge;
                            if (ne > FUSED_NEDGE) {
                                // too many order dependences for the edge list (tiny populations with most agents infectious): take half the entries
                                __syncthreads();
                                if (GS) { for (int k = tid; k < FUSED_HCAP / 4; k += NT) reinterpret_cast<uint4 *>(M.htab)[k] = make_uint4(~0u, ~0u, ~0u, ~0u); }
                                else for (int k = tid; k < nwords; k += NT) { M.touch[k] = 0; M.dupm[k] = 0; }
                                if (tid == 0) Z.nedge = 0;
                                lim = max(1, mlist >> 1);
                                __syncthreads();
                                continue;
                            }
                            // ---- (c) budgets of entries that earlier entries of the chunk contact: relax until nothing changes
                            if (ne > 0) {
#ifdef CABM_FUSED_DIAG
                                if (tid == 0) t_ph[18] += ne;
#endif
                                for (;;) {
#ifdef CABM_FUSED_DIAG
                                    if (tid == 0) t_ph[19] += 1;
#endif
                                    for (int j = tid; j < ne; j += NT) M.hits[(M.edges[j] >> 11) & 255u] = 0;      // only entries with edges count
                                    __syncthreads();
                                    for (int j = tid; j < ne; j += NT) {
                                        const uint32_t ed = M.edges[j];
                                        const int q = (int)(M.evm[ed & 2047u] >> 2), g = (int)((ed >> 19) & 7u), k = (int)(ed >> 22);
                                        const int cq = M.cc[q];
                                        bool valid = cq == (int)M.cb[q];
                                        if (!valid) valid = k < (int)((gpw_pack(Z.cm + ((M.ch[q] >> 4) & 7) * 5, cq) >> (8 * g)) & 0xFF);
                                        if (valid) atomicAdd(&M.hits[(ed >> 11) & 255u], 1);
                                    }
                                    __syncthreads();
                                    bool chg = false;
                                    for (int j = tid; j < ne; j += NT) {                            // (several edges into one entry compute the same value)
                                        const int t = (int)((M.edges[j] >> 11) & 255u);
                                        const int nc = max(0, (int)M.cb[t] - M.hits[t]);           // admitted hits = min(budget, hits) :811-813
                                        if (nc != (int)M.cc[t]) { M.cc[t] = (uint8_t)nc; chg = true; }
                                    }
                                    if (!__syncthreads_or(chg)) break;
                                }
                                // events a reduced budget no longer has: (g, k) with k >= round(cm[ag][g] * cnts'); one thread per edge walks the
                                // events of the entry the edge points to (an entry with several edges is walked more than once, to the same effect)
                                for (int j = tid; j < ne; j += NT) {
                                    const int t = (int)((M.edges[j] >> 11) & 255u), cq = M.cc[t];
                                    if (cq == (int)M.cb[t]) continue;
                                    const unsigned long long gnew = gpw_pack(Z.cm + ((M.ch[t] >> 4) & 7) * 5, cq), gold = M.gp0[t];
                                    int e = M.pref[t];
                                    for (int g = 0; g < 5; ++g) {
                                        const int n_old = (int)((gold >> (8 * g)) & 0xFF), n_new = (int)((gnew >> (8 * g)) & 0xFF);
                                        for (int k = n_new; k < n_old; ++k) M.evm[e + k] |= EVM_DROPPED;
                                        e += n_old;
                                    }
                                }
                                __syncthreads();
                            }
                            PHASE(10);
                            // ---- (d) apply. One event on a target: independent of everything else in the chunk (:809-830)
                            auto log_contact = [&](int q, uint32_t y) {                          // :817-819
                                S.ct_log[(size_t)r * S.ct_cap + base0 + M.cbase[q] + 2u + (uint32_t)atomicAdd(&M.clen[q], 1)] = y;
                            };
                            int ndl = 0;
                            unsigned smask = 0;                                                // bit j: this lane's event of window win0 + j shares its target
                            auto apply = [&](int win, auto nwc) {
                                constexpr int NWIN = decltype(nwc)::value;
                                uint32_t y[NWIN], rec[NWIN], dm[NWIN], sy[NWIN]; bool single[NWIN];
#pragma unroll
                                for (int j = 0; j < NWIN; ++j) {
                                    const int e = ((win + j) << 5) + lane;
                                    y[j] = e < Etot ? M.evy[e] : EV_NONE;
                                    rec[j] = e < Etot ? M.evm[e] : 0u;
                                }
#pragma unroll
                                for (int j = 0; j < NWIN; ++j) {
                                    single[j] = y[j] != EV_NONE && !(rec[j] & EVM_DROPPED);      // live so far
                                    if constexpr (GS) {
                                        dm[j] = single[j] ? M.htab[M.D[((win + j) << 5) + lane]] >> 31 : 0u;    // the slot found by the event's insert
                                        sy[j] = single[j] ? M.status[y[j]] : 0u;                 // (nearly every target is hit once: loaded without waiting for the answer)
                                    } else {
                                        dm[j] = single[j] ? (M.dupm[y[j] >> 5] >> (y[j] & 31)) & 1u : 0u;
                                    }
                                }
#pragma unroll
                                for (int j = 0; j < NWIN; ++j) {
                                    const bool shared = single[j] && dm[j];
                                    single[j] = single[j] && !shared;
                                    smask |= (unsigned)shared << (win + j - win0);
                                    ndl += __popc(__ballot_sync(FULL, shared));
                                    if constexpr (!GS) sy[j] = single[j] ? M.status[y[j]] : 0u;
                                }
#pragma unroll
                                for (int j = 0; j < NWIN; ++j) {
                                    if (!single[j]) continue;
                                    // The status word counts today's contacts of y (when its tag is today); the event is admitted iff that
                                    // count is still below y's budget (:812-813). The budget is only drawn when the answer matters: the event
                                    // would transmit to a susceptible, or its infector logs contacts.
                                    const int hits = st_tag(sy[j]) == day ? st_rem(sy[j]) : 0, q = (int)(rec[j] >> 2);   // today's contacts of y so far
                                    uint32_t ns = (sy[j] & ~(ST_REMMASK | ST_TAGMASK)) | ((uint32_t)day << 22) | (uint32_t)min(hits + 1, 255);
                                    const bool infects = st_health(sy[j]) == SUS && st_swap(sy[j]) == UNDEF && (rec[j] & EVM_SUCC);   // :822-823
                                    const bool logs = ctr && room && M.cflag[q];
                                    if ((infects || logs) && hits < fresh_budget_s(M, key, y[j], (uint32_t)day, sy[j])) {
                                        if (logs) log_contact(q, y[j]);
                                        if (infects) {                                              // :824-827
                                            ns = st_set_swap(ns, LAT);
                                            S.sickfrom[rb + y[j]] = (uint8_t)(M.ch[q] & 0xF);
                                            M.expday[y[j]] = (int16_t)day;                          // y.exp = y.tis :826 — fires in today's time_update
                                            prefetch_l2(S.dur + rb + y[j]); prefetch_l2(S.isovia + rb + y[j]);   // what that transition will read
                                        }
                                    }
                                    M.status[y[j]] = ns;
                                }
                            };
                            {
                                int win = win0;
                                if constexpr (GS) {
                                    for (; win + 8 <= win1; win += 8) apply(win, std::integral_constant<int, 8>());
                                    if (win + 4 <= win1) { apply(win, std::integral_constant<int, 4>()); win += 4; }
                                    for (; win + 2 <= win1; win += 2) apply(win, std::integral_constant<int, 2>());
                                }
                                for (; win < win1; ++win) apply(win, std::integral_constant<int, 1>());
                            }
                            if (lane == 0) Z.wdup[wid] = ndl;
                            PHASE(11);
                            __syncthreads();                                                   // C3
                            int dpos = 0, ndup = 0;
                            for (int w = 0; w < NW; ++w) { const int c = Z.wdup[w]; ndup += c; if (w < wid) dpos += c; }
                            if (ndup > 0) {
                                // events that share a target, compacted in event order
                                for (int win = win0; win < win1; ++win) {
                                    const int e = (win << 5) + lane;
                                    const bool shared = (smask >> (win - win0)) & 1u;          // found by the apply pass above
                                    const unsigned dm = __ballot_sync(FULL, shared);
                                    if (shared) M.D[dpos + __popc(dm & ((1u << lane) - 1u))] = (uint16_t)e;
                                    dpos += __popc(dm);
                                }
                                PHASE(13);
                                __syncthreads();                                               // C4
#ifdef CABM_FUSED_DIAG
                                if (tid == 0) t_ph[17] += ndup;
#endif
                                // warp w owns the targets y with y mod 8 == w: it walks the list 32 entries at a time and applies its
                                // targets' events in list (= event) order; lanes that share a target form a group led by its first event
                                for (int base = 0; base < ndup; base += 32) {
                                    const int idx = base + lane;
                                    const int ed = idx < ndup ? (int)M.D[idx] : 0;
                                    const uint32_t y = idx < ndup ? M.evy[ed] : EV_NONE;
                                    const uint32_t rec = M.evm[ed];
                                    const bool mine = y != EV_NONE && (int)(y & (NW - 1)) == wid;
                                    if (!__any_sync(FULL, mine)) continue;
                                    const int qd = (int)(rec >> 2);
                                    const unsigned peers = __match_any_sync(FULL, mine ? y : EV_NONE);
                                    const bool leader = mine && lane == __ffs(peers) - 1;
                                    const int n = __popc(peers), nmax = __reduce_max_sync(FULL, mine ? n : 0);
                                    const uint32_t sy = mine ? M.status[y] : 0u;
                                    const int hits0 = st_tag(sy) == day ? st_rem(sy) : 0;                    // contacts of y today before this group's
                                    int xh_inf = -1;
                                    bool sus = st_health(sy) == SUS && st_swap(sy) == UNDEF;                 // :822
                                    const int hit = (mine && (rec & EVM_SUCC)) ? (int)(M.ch[qd] & 0xF) : -1;     // would this event transmit (:823)?
                                    const bool logs = ctr && room && mine && M.cflag[qd];
                                    // y's budget is drawn only if some event of the group would transmit to it or is logged
                                    const bool wanted = (__ballot_sync(FULL, (hit >= 0 && sus) || logs) & peers) != 0u;
                                    const int budget = (leader && wanted) ? fresh_budget_s(M, key, y, (uint32_t)day, sy) : 0;
                                    unsigned m = peers;
                                    for (int k = 0; k < nmax; ++k) {                                          // k-th event of each target, in event order
                                        const int src = m ? __ffs(m) - 1 : lane; m &= m - 1u;
                                        const int hs = __shfl_sync(FULL, hit, src);
                                        const int qs = __shfl_sync(FULL, qd, src);
                                        if (leader && k < n && hits0 + k < budget) {                          // admitted :812-813
                                            if (ctr && room && M.cflag[qs]) log_contact(qs, y);
                                            if (sus && hs >= 0) { sus = false; xh_inf = hs; }
                                        }
                                    }
                                    if (leader) {
                                        uint32_t ns = (sy & ~(ST_REMMASK | ST_TAGMASK)) | ((uint32_t)day << 22) | (uint32_t)min(hits0 + n, 255);
                                        if (xh_inf >= 0) {
                                            ns = st_set_swap(ns, LAT);
                                            S.sickfrom[rb + y] = (uint8_t)xh_inf;
                                            M.expday[y] = (int16_t)day;
                                            prefetch_l2(S.dur + rb + y); prefetch_l2(S.isovia + rb + y);
                                        }
                                        M.status[y] = ns;
                                    }
                                    __syncwarp();
                                }
                            }
                            PHASE(14);
                            __syncthreads();                                                   // C5: status writes visible to the next chunk / time_update
                            if (ctr && room && tid < nb && M.cflag[tid] && M.clen[tid] > 0) {     // link the day's chunk of each traced infector
                                const int xq = M.cx[tid];
                                uint32_t *c = S.ct_log + (size_t)r * S.ct_cap + base0 + M.cbase[tid];
                                c[0] = S.ct_head[rb + xq]; c[1] = (uint32_t)M.clen[tid];
                                S.ct_head[rb + xq] = base0 + M.cbase[tid];
                            }
                            if (GS) {
                                for (int k = tid; k < FUSED_HCAP / 4; k += NT) reinterpret_cast<uint4 *>(M.htab)[k] = make_uint4(~0u, ~0u, ~0u, ~0u);     // next insert is after the next C1
                            } else {
                                for (int k = tid; k < nwords; k += NT) { M.touch[k] = 0; M.dupm[k] = 0; }   // next marking is after the next C1
                            }
                            if (tid == 0) { Z.nedge = 0; if (room) Z.logn = base0 + (uint32_t)needtot; }
#ifdef CABM_FUSED_DIAG
                            if (tid == 0) t_ph[16] += Etot;
#endif
                            PHASE(15);
                            lpos += mlist; lim = FUSED_LCAP;
                        }
                        PHASE(12);
                        if (w0 >= nwords && lbase <= FUSED_LCAP) break;                        // the list held every infectious agent
                        cursor = (int)M.list[FUSED_LCAP - 1] + 1;
                        __syncthreads();
                    }
                }

                // ---------------------------------------------------- time_update() :399-428
                if (!had_infectious) __syncthreads();                          // S1: col reusable (dyntrans, when it runs, ends with barriers of its own)
                PHASE(12);
                if (tid == 0) Z.nf[(act + 1) & 1] = 32767;                     // buffer of the next active day (last read after the previous S2)
                if (day == 1) {                                                // seeds isolated by fpreiso: resync isob (see k_time_update)
                    for (int i = tid; i < N; i += NT) { const uint32_t st = M.status[i]; if (((st >> 19) ^ (st >> 20)) & 1u) M.status[i] = (st & ST_ISO) ? (st | ST_ISOB) : (st & ~ST_ISOB); }
                }
                if (ctr) {
                    // pass A of ct_dynamics (:705-716): today's index cases (queued yesterday by agent_event_core) mark their traced
                    // contacts before anybody is updated (marks are bitmask ORs and keyed draws: the index cases are independent of each other).
                    const uint32_t nq = Z.qcnt[day & 1];
                    if (nq) {
                        __syncthreads();        // the contact chunks linked after today's last dyntrans chunk (by many warps) are visible to every warp
                        {   // one index case per half-warp at a time (the chain walks are latency-bound: sixteen in flight), each group
                            // gathering contacts in its own slice of the event buffer and deduplicating in its slice of the D/evm area
                            const uint32_t *qlist = S.ident_list + ((size_t)r * 2 + (day & 1)) * Np;
                            const int grp = tid >> 4;
                            constexpr int GB = FUSED_EVCAP / (2 * NW);                 // 128 words of gather buffer and of hash table per group
                            uint32_t *const htab_all = reinterpret_cast<uint32_t *>(M.D);   // D and evm are adjacent: 2 x EVCAP x 2 B = EVCAP words
                            for (uint32_t k = grp; k < nq; k += 2 * NW)
                                ct_identify_warp<false, 16>(S, P, r, (int)qlist[k], day, M.evy + grp * GB, GB, htab_all + grp * GB, GB, M.mpre, M.mpost);
                        }
                        __syncthreads();
                        if (tid == 0) Z.qcnt[day & 1] = 0;
                    }
                }
                // every thread looks at U chunks of 8 next-event days at once (independent 128-bit loads, packed 16-bit minima)
                PHASE(27);
                uint32_t mn = 0x7FFF7FFFu;                                     // running min of the event days that stay (2 x i16)
                {
                    // The scan only LISTS the agents with an event today (event day reached, or marked by pass A) and keeps the minimum
                    // of the event days that stay; the listed agents are then taken one per thread in dense warps. (Handled inside the
                    // stream, a lane's event stalls its whole warp for the length of the event's dependent chain — several times per pass
                    // when the state is in global memory, and once per event a lane holds on a heavy day in shared memory.)
                    constexpr int U = GS ? 8 : FUSED_SCAN_U;
                    uint32_t *const evl = GS ? S.ev_list + (size_t)r * Np : M.evy;          // (the event buffer is free between dyntrans and the next chunk build)
                    const unsigned ecap = GS ? (unsigned)Np : (unsigned)FUSED_EVCAP;        // shared memory: what does not fit is handled in place
                    const uint32_t dd = (uint32_t)day * 0x10001u;
                    int mne = 32767;
                    auto do_event = [&](int i, unsigned pb, unsigned qb) {
                        if (ctr) {
                            if (pb) atomicAnd(&M.mpre[i >> 5], ~(1u << (i & 31)));
                            if (qb) atomicAnd(&M.mpost[i >> 5], ~(1u << (i & 31)));
                        }
                        const uint32_t st = M.status[i];
                        const int texp = ctr ? (int)S.texp[rb + i] : (int)M.expday[i];      // without tracing the scan key is the expiry day
                        const int par = (day + 1) & 1;
#ifdef CABM_FUSED_DIAG
                        const long long t_e0 = clock64();
#endif
                        const Ev e = agent_event_core(S, P, r, i, day, key, st, texp, ctr, pb, qb,
                                                      &Z.qcnt[par], S.ident_list ? S.ident_list + ((size_t)r * 2 + par) * Np : nullptr, M.dcache);
#ifdef CABM_FUSED_DIAG
                        if (e.st != 0xFFFFFFFFu) { atomicAdd((unsigned long long *)&t_ph[28], (unsigned long long)(clock64() - t_e0)); atomicAdd((unsigned long long *)&t_ph[29], 1ull); }
#endif
                        M.status[i] = e.st;
                        M.expday[i] = (int16_t)e.ne;
                        mne = min(mne, e.ne);
                        if (ctr && e.texp != texp) S.texp[rb + i] = (int16_t)e.texp;
                        if (e.newh != e.oldh) {
                            M.col[i] = (uint8_t)e.newh;
                            atomicAdd(&Z.io[act & 1][0][st_ag(st) * 12 + e.newh], 1); atomicAdd(&Z.io[act & 1][1][st_ag(st) * 12 + e.oldh], 1);
                            if (is_infectious(e.newh) != is_infectious(e.oldh)) {
                                atomicXor(&M.inf[i >> 5], 1u << (i & 31));
                                atomicAdd(&Z.ninf, is_infectious(e.newh) ? 1 : -1);
                            }
                        }
                    };
                    for (int bb = 0; bb < N; bb += NT * 8 * U) {             // (the same number of passes for every thread: the append is a warp collective)
                        const int b0 = bb + tid * 8;
                        uint4 ex[U];
                        unsigned long long fire = 0, pre = 0, post = 0;        // 8 bits per chunk
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int i0 = b0 + u * NT * 8;
                            ex[u] = i0 < N ? *reinterpret_cast<const uint4 *>(M.expday + i0) : make_uint4(0x7FFF7FFFu, 0x7FFF7FFFu, 0x7FFF7FFFu, 0x7FFF7FFFu);
                            if (GS && i0 + NT * 8 * U < N) prefetch_l2(M.expday + i0 + NT * 8 * U);      // the next pass of this thread
                            if (ctr && i0 < N) {
                                const unsigned sh = i0 & 31;
                                pre |= (unsigned long long)((M.mpre[i0 >> 5] >> sh) & 0xFFu) << (8 * u);
                                post |= (unsigned long long)((M.mpost[i0 >> 5] >> sh) & 0xFFu) << (8 * u);
                            }
                        }
                        // earliest event day of these chunks with packed 16-bit minima (DPX VIMNMX: one instruction per pair of words); the
                        // per-agent compares are only needed when somebody is due
                        uint32_t cmin = 0x7FFF7FFFu;
#pragma unroll
                        for (int u = 0; u < U; ++u) cmin = __vimin3_s16x2(cmin, __vmins2(ex[u].x, ex[u].y), __vmins2(ex[u].z, ex[u].w));
                        if (min((int)(short)(cmin & 0xFFFFu), (int)(short)(cmin >> 16)) <= day || (pre | post) != 0) {
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                uint32_t wds[4] = {ex[u].x, ex[u].y, ex[u].z, ex[u].w};
                                const unsigned mk = (unsigned)((pre | post) >> (8 * u)) & 0xFFu;
                                unsigned f = 0;
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    uint32_t m = __vcmples2(wds[k], dd);                              // 0xFFFF per half whose day has come
                                    f |= ((m & 1u) | ((m >> 15) & 2u)) << (2 * k);
                                    m |= ((mk >> (2 * k)) & 1u ? 0xFFFFu : 0u) | ((mk >> (2 * k + 1)) & 1u ? 0xFFFF0000u : 0u);
                                    mn = __vmins2(mn, (wds[k] & ~m) | (0x7FFF7FFFu & m));             // listed agents report their new event day below
                                }
                                fire |= (unsigned long long)f << (8 * u);
                            }
                            fire |= pre | post;
                        } else {
                            mn = __vmins2(mn, cmin);
                        }
                        if (fire && b0 + (U - 1) * NT * 8 + 8 > N) {           // the last block of eight may run past the population: drop the padding
#pragma unroll
                            for (int u = 0; u < U; ++u) {
                                const int nv = N - (b0 + u * NT * 8);
                                if (nv < 8) fire &= ~((nv <= 0 ? 0xFFull : ((0xFFull << nv) & 0xFFull)) << (8 * u));
                            }
                        }
                        // append. Shared-memory variant: a thread with listed agents reserves its slots with one atomic (a few dozen a day).
                        // Global-state variant (thousands a day): one atomic per warp, slots by a prefix over the lanes.
                        unsigned wbase = 0;
                        if constexpr (GS) {
                            const int wc = __popcll(fire);
                            if (__any_sync(FULL, wc != 0)) {
                                int wincl = wc;
#pragma unroll
                                for (int o = 1; o < 32; o <<= 1) { const int v = __shfl_up_sync(FULL, wincl, o); if (lane >= o) wincl += v; }
                                if (lane == 31) wbase = atomicAdd(&Z.evn, (unsigned)wincl);
                                wbase = __shfl_sync(FULL, wbase, 31) + (unsigned)(wincl - wc);
                            }
                        }
                        if (fire) {
                            unsigned base = GS ? wbase : atomicAdd(&Z.evn, (unsigned)__popcll(fire));
                            while (fire) {
                                const int bit = __ffsll((long long)fire) - 1; fire &= fire - 1;
                                const int i = b0 + (bit >> 3) * NT * 8 + (bit & 7);
                                const unsigned pb = (unsigned)(pre >> bit) & 1u, qb = (unsigned)(post >> bit) & 1u;
                                if (base < ecap) evl[base++] = (uint32_t)i | (pb << 30) | (qb << 31);
                                else do_event(i, pb, qb);                      // (list full: a day with thousands of events in shared-memory mode)
                            }
                        }
                    }
                    PHASE(30);
                    __syncthreads();                                           // the list is complete
                    PHASE(31);
                    const unsigned nev = min(Z.evn, ecap);
                    // listed event k goes to warp k mod 8, lane k / 8: a few dozen events of different kinds (different branches of the
                    // state machine) spread over all warps instead of serialising in one
                    for (unsigned k = (unsigned)(wid + NW * lane); k < nev; k += NT) {
                        const uint32_t ent = evl[k];
                        do_event((int)(ent & 0x3FFFFFFFu), (ent >> 30) & 1u, ent >> 31);
                    }
                    mn = __vmins2(mn, (uint32_t)(mne & 0xFFFF) | 0x7FFF0000u);
                }
                PHASE(25);
                {
                    int v = min((int)(short)(mn & 0xFFFF), (int)(short)(mn >> 16));
                    v = __reduce_min_sync(FULL, v);
                    if (lane == 0) atomicMin(&Z.nf[act & 1], v);
                }
                if (hm && tma_ok) bulk_store_g2s_fence();
                PHASE(26);
                __syncthreads();                                               // S2
                nextfire = Z.nf[act & 1]; ++act;
                if (tid == 0) Z.evn = 0;                                 // every thread has read the count (before S2); the next append is behind a barrier
                PHASE(4);
            }
            // -------------------------------------------------------- column st+1 (:136 of the next day) and its curves
            if (day < T) {
                // threads 0..59 own one (age group, state) pair each and keep its running prevalence in a register;
                // threads 64..75 own group 0 ("all" = the sum over the five groups, :91-97). Idle days repeat the values.
                // curve deltas of active day k live in Z.io[k & 1] (k = act - 1 here); the other buffer, last read two active days ago,
                // is cleared for the next active day — no barrier needed between the group threads and the group-0 threads
                const int cur = (act - 1) & 1;
                if (tid < 60) {
                    const int g = tid / 12, s = tid % 12;
                    int in = 0;
                    if (active) {
                        in = Z.io[cur][0][tid];
                        run += in - Z.io[cur][1][tid];
                        Z.io[cur ^ 1][0][tid] = 0; Z.io[cur ^ 1][1][tid] = 0;
                    }
                    const size_t b = (((size_t)r * 6 + (g + 1)) * (size_t)T + (size_t)day) * 12 + s;
                    S.inc[b] = in; S.prev[b] = run;
                } else if (tid >= 64 && tid < 76) {
                    const int s = tid - 64;
                    g0_in = 0;
                    if (active) {
                        int out = 0;
                        for (int g = 0; g < 5; ++g) { g0_in += Z.io[cur][0][g * 12 + s]; out += Z.io[cur][1][g * 12 + s]; }
                        run += g0_in - out;
                    }
                    const size_t b = (((size_t)r * 6 + 0) * (size_t)T + (size_t)day) * 12 + s;
                    S.inc[b] = g0_in; S.prev[b] = run;
                    if (s == LAT) latsum += g0_in;
                }
                if (hm) {
                    if (tma_ok) { if (tid == 0) { bulk_store_wait_read_n(); bulk_store(hm + (size_t)day * N, M.col, (uint32_t)N); } }
                    else copy_column(hm + (size_t)day * N, M.col, N, tid, NT);
                }
            }
            PHASE(5);
        }
        __syncthreads();

        if (!finished) {
            // ---------------------------------------------------------- park: still active after split_day days -> state to HBM, queue for sweep 1
            for (int i = tid; i < Np; i += NT) {
                S.status[rb + i] = M.status[i]; S.expday[rb + i] = M.expday[i];
            }
            for (int k = tid; k < nwords; k += NT) S.infmask[(size_t)r * nwords + k] = M.inf[k];
            if (tid < 6) S.grp_off[r * 8 + tid] = Z.off[tid];
            if (tid == 64 + LAT) S.latsum[r] = latsum;
            if (tid == 0) {
                S.resume_aux[r * 2] = nextfire; S.resume_aux[r * 2 + 1] = Z.ninf;
                if (ctr) { S.ct_cnt[r] = Z.logn; S.ident_cnt[r * 2] = Z.qcnt[0]; S.ident_cnt[r * 2 + 1] = Z.qcnt[1]; }
                if (hm && tma_ok) bulk_store_wait_all();
            }
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                const unsigned slot = atomicAdd(&work[2], 1u);
                __threadfence();
                *(volatile int *)&S.resume_list[slot] = r;
                __threadfence();
                atomicAdd(&work[3], 1u);
            }
            continue;
        }
        // ------------------------------------------------------------ final state back to HBM, _count_infectors :295-313
        if (tid == 64 + LAT) S.latsum[r] = latsum;
        int l[4] = {0, 0, 0, 0};
        bool bad = false;
        for (int i = tid; i < Np; i += NT) {
            const uint32_t st = M.status[i];
            S.status[rb + i] = st;
            S.expday[rb + i] = M.expday[i];
            if (!ctr) S.texp[rb + i] = M.expday[i];
            if (i < N && st_health(st) != SUS) {
                const int f = S.sickfrom[rb + i];
                if (f == PRE) l[0]++; else if (f == ASYMP) l[1]++; else if (f == MILD || f == MISO) l[2]++;
                else if (f == INF || f == IISO) l[3]++; else bad = true;
            }
        }
        for (int k = tid; k < nwords; k += NT) S.infmask[(size_t)r * nwords + k] = M.inf[k];
        if (ctr && tid == 0) { S.ct_cnt[r] = Z.logn; S.ident_cnt[r * 2] = 0; S.ident_cnt[r * 2 + 1] = 0; }
        for (int k = 0; k < 4; ++k) if (l[k]) atomicAdd(&Z.infectors[k], l[k]);
        if (bad) atomicOr(&S.err[r], ERR_SICKFROM);
        __syncthreads();
        if (tid < 4) S.infectors[r * 4 + tid] = Z.infectors[tid];
        if (S.host_inc && tail_day == 0) {
            __syncthreads();                   // the rows written during the day loop (by threads 0..75) are visible block-wide
            stream_curves_to_host(r, idle_from > 0 ? idle_from : T);
        }
        // diagnostics of the fused path: kilo-cycles spent on this replicate and dyntrans batches issued
        PHASE(6);
#ifdef CABM_FUSED_DIAG
        if (tid == 0) { S.totalmet[r] = (int)((clock64() - t_start) >> 10); S.totalinf[r] = n_rounds; for (int k = 0; k < 32; ++k) S.dbg[(size_t)r * 32 + k] = t_ph[k]; }
#else
        if (tid == 0) { S.totalmet[r] = 0; S.totalinf[r] = 0; }
#endif
#undef PHASE
#undef PHASEX
        if (hm && tma_ok && tid == 0) bulk_store_wait_all();
        if (tail_day > 0) {                    // queue the tail job: rows < tail_day of the curves and the state matrix are in HBM
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                S.resume_aux[r * 2] = tail_day;
                const unsigned slot = atomicAdd(&work[5], 1u);
                __threadfence();
                *(volatile int *)&S.resume_list[S.R - 1 - (int)slot] = r;
            }
        }
        if (!resumed && tid == 0) { __threadfence(); atomicAdd(&work[3], 1u); }
    }
    }
    if (tid == 0 && S.store_hm) bulk_store_wait_all();     // column copies of the last fill job
}

size_t fused_smem_bytes_host(int Np, bool global_state) { return fused_smem_bytes(Np, global_state); }

// global_state = false: per-agent state in shared memory (popsize up to ~21k); true: in global memory (any popsize)
int launch_sim_fused(const Dev &S, unsigned *work, int n_sm, bool any_ct, bool global_state, cudaStream_t s) {
    const size_t bytes = fused_smem_bytes(S.Np, global_state);
    {   // per device: function attributes belong to the device's copy of the kernel. Handles of several host threads may get here
        // at once, so the flags sit behind a lock.
        static std::mutex mu;
        static bool attr_set[64] = {false};
        int dev = 0;
        cudaGetDevice(&dev);
        std::lock_guard<std::mutex> lock(mu);
        if (dev < 0 || dev >= 64 || !attr_set[dev]) {
            const int lim = 227 * 1024 - 512;        // the opt-in maximum per CTA minus this kernel's (diagnostic) static shared memory
            cudaFuncSetAttribute(k_sim_fused<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
            cudaFuncSetAttribute(k_sim_fused<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, lim);
            cudaFuncSetAttribute(k_sim_fused<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem_bytes(32, true));
            cudaFuncSetAttribute(k_sim_fused<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fused_smem_bytes(32, true));      // (independent of the population)
            if (dev >= 0 && dev < 64) attr_set[dev] = true;
        }
    }
    // Experiment overrides (environment) are validated: a value that would break the forward-progress argument of the work queues
    // is ignored. That argument: every counter a CTA spins on (work[3], work[7], the *_list slots) is advanced by a CTA that is already
    // resident and running (the grid never exceeds the resident CTA slots), and while copy jobs can still be queued at least one CTA
    // stays eligible to become a copier (tail_ctas >= 1 whenever curves are streamed to the host).
    auto env_int = [](const char *name, int lo, int hi, int dflt) {
        const char *e = getenv(name);
        if (!e || !*e) return dflt;
        char *end = nullptr;
        const long v = strtol(e, &end, 10);
        return (end && *end == 0 && v >= lo && v <= hi) ? (int)v : dflt;
    };
    int per_sm = bytes <= 113 * 1024 ? 2 : 1;
    per_sm = env_int("CABM_FUSED_CTAS_PER_SM", 1, per_sm, per_sm);
    int grid = n_sm * per_sm;
    if (grid > S.R) grid = S.R;
    // park-and-resume only pays when there are more replicates than resident CTAs and the run is long; with the state in global
    // memory a replicate runs from start to end on its CTA
    int split_day = (S.R > grid && S.T >= 150) ? 60 : 0;
    split_day = env_int("CABM_FUSED_SPLIT_DAY", 0, S.T - 1, split_day);
    if (split_day >= S.T || global_state) split_day = 0;
    cudaMemsetAsync(work, 0, 16 * sizeof(unsigned), s);
    if (split_day > 0) { cudaMemsetAsync(S.resume_list, 0xFF, (size_t)S.R * sizeof(int), s); cudaMemsetAsync(S.copy_list, 0xFF, (size_t)S.R * sizeof(int), s); }
    unsigned tail_ctas = FUSED_TAIL_CTAS;
    if (tail_ctas > (unsigned)grid / 4) tail_ctas = grid / 4 > 0 ? grid / 4 : 1;      // copiers do not resume parked replicates
    tail_ctas = (unsigned)env_int("CABM_FUSED_TAIL_CTAS", 1, grid, (int)tail_ctas);
    if (global_state) {
        if (any_ct) k_sim_fused<true, true><<<grid, FUSED_NT, bytes, s>>>(S, work, split_day, tail_ctas);
        else k_sim_fused<false, true><<<grid, FUSED_NT, bytes, s>>>(S, work, split_day, tail_ctas);
    } else {
        if (any_ct) k_sim_fused<true, false><<<grid, FUSED_NT, bytes, s>>>(S, work, split_day, tail_ctas);
        else k_sim_fused<false, false><<<grid, FUSED_NT, bytes, s>>>(S, work, split_day, tail_ctas);
    }
    return 1;
}

}  // namespace cabm
