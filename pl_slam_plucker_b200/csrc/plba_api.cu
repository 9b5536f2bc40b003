// Host side of the C ABI (include/plba.h).
//   upload   : index bookkeeping only — landmarks of each window are sorted by keyframe signature, cut into chunks and
//              segments, and flattened straight into ONE pinned staging buffer that reaches HBM with a single H2D copy;
//              device memory is one arena per handle (no per-call cudaMalloc).
//   run      : the whole LM schedule is a CUDA graph: WHILE(windows left){ IF(prep){gate, lambda init}; assemble; solve;
//              update+controller }.  The controller runs on the device and steers both conditional nodes, so the host
//              is not in the loop.  The graph is built once per (handle, profile): kernel parameters live at a fixed
//              device address and the grids are persistent, so nothing in the graph depends on the problem.
//              A host-driven loop over the same kernels remains for the tiled large-window solver, for the sharded
//              multi-GPU exchange step (all-reduce hook) and for per-kernel event timing.
//   download : one export kernel, one D2H copy, host-side un-permutation into the caller's buffers.
// There is NO CPU fallback in this file: every numeric step is a kernel launch on the handle's stream.
#include <vector>
#include <string>
#include <tuple>
#include <utility>
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <chrono>
#include "plba_solver.h"
#include "plba_warp.h"
#ifdef _OPENMP
#include <omp.h>
#endif
#include "plba_track.h"
#ifndef PLBA_HOST_EMU
#include <dlfcn.h>
#include <mutex>
#endif

using namespace plba;

// ---- NCCL, resolved at run time (no link-time dependency; the prototypes below are NCCL's public C API, nccl.h) ---------------
#ifndef PLBA_HOST_EMU
namespace {
typedef struct ncclComm *plba_ncclComm_t;
typedef struct { char internal[PLBA_COMM_ID_BYTES]; } plba_ncclUniqueId;
enum { PLBA_NCCL_SUM = 0, PLBA_NCCL_MAX = 2, PLBA_NCCL_F64 = 8 };
struct NcclApi {
    int (*GetUniqueId)(plba_ncclUniqueId *) = nullptr;
    int (*CommInitRank)(plba_ncclComm_t *, int, plba_ncclUniqueId, int) = nullptr;
    int (*CommInitAll)(plba_ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(plba_ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, plba_ncclComm_t, cudaStream_t) = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool ok = false; std::string why;
};
NcclApi &nccl_api() {
    static NcclApi api; static std::once_flag once;
    std::call_once(once, [] {
        void *hd = nullptr;
        const char *env = std::getenv("PLBA_NCCL_LIB");
        if (env && env[0]) hd = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
        if (!hd && dlsym(RTLD_DEFAULT, "ncclAllReduce")) hd = RTLD_DEFAULT;       // already mapped (e.g. by torch): use that copy
        if (!hd) hd = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
        if (!hd) hd = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
        if (!hd) { api.why = "libnccl.so.2 not found (set PLBA_NCCL_LIB)"; return; }
        auto sym = [&](const char *n) { void *q = dlsym(hd, n); if (!q) api.why = std::string("missing NCCL symbol ") + n; return q; };
        api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
        api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
        api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
        api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
        api.AllReduce = (decltype(api.AllReduce))sym("ncclAllReduce");
        api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
        api.ok = api.GetUniqueId && api.CommInitRank && api.CommInitAll && api.CommDestroy && api.AllReduce && api.GetErrorString;
    });
    return api;
}
}  // namespace
#endif
enum { PLBA_MAX_RANKS = 16 };

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return PLBA_E_CUDA; } } while (0)

// host loops of upload / download run in parallel above these sizes (a parallel region costs tens of microseconds: a config-2 window of
// 45 000 observations stays serial, config 4 with 1.15 M observations does not)
enum { PAR_LM = 40000, PAR_OBS = 200000 };
// ... and on about one thread per 150 000 observations, not on every core: config 4 (1.15 M observations) prepares in 8.8 ms on 8 threads
// and in 17.6 ms on the box's 16 (a parallel region's fork / join over 16 virtual CPUs costs more than a mid-size loop gains; config 5's
// 11 M observations take all of them).  PLBA_HOST_THREADS overrides.  Set at the top of plba_upload / plba_download.
static thread_local int g_host_nt = 1;
static int host_threads_for(int64_t n_obs) {
#ifdef _OPENMP
    const char *e = std::getenv("PLBA_HOST_THREADS"); const int env = e ? std::atoi(e) : 0;      // read per call: the tests switch it between uploads
    const int mx = omp_get_max_threads();
    if (env > 0) return std::min(env, std::max(mx, 1));
    return (int)std::max<int64_t>(1, std::min<int64_t>(mx, n_obs / 150000));
#else
    (void)n_obs; return 1;
#endif
}
struct WinInfo { int n_kf, n_free, n_pt, n_ls, n_pobs, n_lobs, kf0, slot0, pt0, ls0, po0, lo0; };

// bump allocator over a byte range (device arena or pinned staging): 256-byte aligned carve-outs
struct Carver {
    size_t off = 0;
    template <typename T> size_t take(size_t n) { const size_t o = off; off += (std::max<size_t>(n, 1) * sizeof(T) + 255) & ~(size_t)255; return o; }
};

struct plba_handle_s {
    int device = 0, n_sm = 148;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    // memory
    char *d_arena = nullptr; size_t d_cap = 0;
    char *h_in = nullptr; size_t h_in_cap = 0;       // pinned: staged inputs
    char *h_out = nullptr; size_t h_out_cap = 0;     // pinned: staged outputs
    DevP *d_P = nullptr;                             // fixed address: kernel parameters
    double *d_scratch = nullptr;                     // a few doubles of device memory for upload-time exchanges
    char *trk_dev = nullptr, *trk_host = nullptr; size_t trk_dev_cap = 0, trk_host_cap = 0;      // pose tracking (plba_track_solve): its own small arena
    DevP *h_P = nullptr;                             // pinned copy
    int *h_cnt = nullptr;                            // pinned counters
    DevP P{};
    size_t in_bytes = 0, out_off = 0, out_bytes = 0;
    // offsets inside the output region
    size_t o_T = 0, o_x = 0, o_pt = 0, o_ls = 0, o_plk = 0, o_pchi = 0, o_lchi = 0, o_pf = 0, o_lf = 0, o_ctrl = 0, o_trace = 0, o_cnt = 0;
    size_t i_pts0 = 0, i_lns0 = 0, i_lmap = 0;
    plba_options opt{};
    std::vector<WinInfo> wins;
    std::vector<int> pt_perm, ls_perm, po_perm, lo_perm;    // internal index -> caller's (global, window-offset) index
    std::vector<Chunk> ch_pt, ch_ls; std::vector<Seg> sg_pt, sg_ls; std::vector<int> fp_pt, fp_ls, pt_ptr, ls_ptr;   // upload scratch, capacity re-used
    std::vector<unsigned char> mark;
    int ls_dim = 4, max_nf = 0, solve_class = 1, band_blocks = 0;
    bool force_dense = false;
    int force_chunk = 0;             // 0 = route by size, 1 = always the CTA-chunk kernels, 2 = the warp kernels whenever every track fits a warp
    bool warp_path = true;           // this upload runs on the warp-autonomous kernels (plba_warp.h)
    std::vector<WItem> wi_pt, wi_ls;
    std::vector<long long> win_S_off; // offset of each window's reduced camera system inside P.S (host copy of DevP::win_S_off)
    std::vector<BcrW> bcr;           // large banded windows: node storage of the block cyclic reduction (per window)
    int large_solver = 0;            // 0 = block cyclic reduction (default), 1 = single-CTA banded Cholesky (A/B, tests)
    int grid_warp = 592, grid_warp_upd = 592;
    bool uploaded = false, small_path = true;
    bool bcr_v1 = false;             // PLBA_BCR_V1=1: the one-CTA-per-node elimination kernel of round 1 (A/B runs)
    bool bcr_layout = false;         // this upload stores the reduced camera system of its (large, block-banded) windows in the node form of the block cyclic reduction
    double *sysbuf = nullptr; size_t sys_doubles = 0, S_doubles = 0;
    int h_counters[CNT_N] = {0};
    int klaunch_seen = 0;            // device launch counter (CNT_KLAUNCH) at the last poll; 0 after a reset
    int64_t layout[8] = {0};
    plba_allreduce_fn allreduce = nullptr; void *allreduce_user = nullptr;
    void *nccl_comm = nullptr;       // ncclComm_t owned by the handle (plba_comm_init_rank / plba_create_group)
    int n_ranks = 1, rank = 0;
    int nccl_failed = 0;
    int *h_poll = nullptr; cudaEvent_t ev_poll[4]{};    // pinned copies of the device counters, one per in-flight round of the pipelined loop
    plba_timing timing{};
    cudaEvent_t ev[8]{};
    cudaStream_t stream_panel = nullptr; cudaEvent_t ev_la[2]{};   // dense tiled Cholesky: the look-ahead panel runs on a stream of its own
    bool no_lookahead = false;                                     // PLBA_NO_LOOKAHEAD=1 (A/B runs)
    int dense_group = 0;                                           // panels per trailing update of the dense tiled Cholesky (PLBA_DENSE_GROUP; 0 = by size)
    bool no_overlap = false;                                       // PLBA_NO_OVERLAP=1: solver and update kernel of a small window strictly one after the other (A/B runs)
    bool dense_k1 = false;                                         // PLBA_DENSE_K1=1: one trailing update per panel instead of per panel pair (A/B runs)
    cudaEvent_t ev_h2d = nullptr; bool h2d_pending = false;   // recorded after the H2D copies of an upload: the next upload waits for it before it rewrites the pinned staging
    bool detail_timing = false, no_graph = false;
    int grid_chunks = 296, grid_solve = 148;
#ifndef PLBA_HOST_EMU
    cudaGraph_t graph[12]{};       // [(profile * 2 + size class of the reduced system) * 2 + warp path]
    cudaGraphExec_t gexec[12]{};
    cudaGraphConditionalHandle cond_while[12]{}, cond_prep[12]{}, cond_prep2[12]{};
#endif
    void release() {
        if (d_arena) cudaFree(d_arena);
        if (h_in) cudaFreeHost(h_in);
        if (h_out) cudaFreeHost(h_out);
        if (trk_dev) cudaFree(trk_dev);
        if (trk_host) cudaFreeHost(trk_host);
        trk_dev = trk_host = nullptr; trk_dev_cap = trk_host_cap = 0;
        d_arena = nullptr; h_in = nullptr; h_out = nullptr; d_cap = h_in_cap = h_out_cap = 0; uploaded = false;
    }
};

static inline bool has_exchange(plba_handle h) { return h->nccl_comm != nullptr || h->allreduce != nullptr; }

static int ensure(plba_handle h, size_t dev_bytes, size_t in_bytes, size_t out_bytes) {
    if (dev_bytes > h->d_cap) {
        cudaStreamSynchronize(h->stream);
        if (h->d_arena) cudaFree(h->d_arena);
        h->d_arena = nullptr; h->d_cap = 0;
        const size_t want = dev_bytes + dev_bytes / 4 + (1 << 20);
        void *q = nullptr;
        if (cudaMalloc(&q, want) != cudaSuccess) { h->err = "cudaMalloc of the device arena failed"; return PLBA_E_CUDA; }
        h->d_arena = (char *)q; h->d_cap = want;
    }
    if (in_bytes > h->h_in_cap) {
        cudaStreamSynchronize(h->stream);
        if (h->h_in) cudaFreeHost(h->h_in);
        h->h_in = nullptr; h->h_in_cap = 0;
        const size_t want = in_bytes + in_bytes / 4 + (1 << 16);
        void *q = nullptr;
        if (cudaMallocHost(&q, want) != cudaSuccess) { h->err = "cudaMallocHost (input staging) failed"; return PLBA_E_CUDA; }
        h->h_in = (char *)q; h->h_in_cap = want;
    }
    if (out_bytes > h->h_out_cap) {
        cudaStreamSynchronize(h->stream);
        if (h->h_out) cudaFreeHost(h->h_out);
        h->h_out = nullptr; h->h_out_cap = 0;
        const size_t want = out_bytes + out_bytes / 4 + (1 << 16);
        void *q = nullptr;
        if (cudaMallocHost(&q, want) != cudaSuccess) { h->err = "cudaMallocHost (output staging) failed"; return PLBA_E_CUDA; }
        h->h_out = (char *)q; h->h_out_cap = want;
    }
    return PLBA_OK;
}

static int validate_problem(const plba_problem &p, const plba_options &o, std::string &err) {
    if (p.n_kf < 0 || p.n_free < 0 || p.n_pt < 0 || p.n_ls < 0 || p.n_pobs < 0 || p.n_lobs < 0) { err = "negative size"; return PLBA_E_ARG; }
    if (p.n_kf && (!p.kf_T_wc || !p.kf_slot)) { err = "kf arrays missing"; return PLBA_E_ARG; }
    if (p.n_pt && !p.pt_xyz) { err = "pt_xyz missing"; return PLBA_E_ARG; }
    if (p.n_pobs && (!p.po_lm || !p.po_kf || !p.po_uv)) { err = "point observation arrays missing"; return PLBA_E_ARG; }
    if (p.n_lobs && (!p.lo_lm || !p.lo_kf || !p.lo_ab)) { err = "line observation arrays missing"; return PLBA_E_ARG; }
    if (p.n_ls && o.profile == PLBA_PROFILE_H_END && !p.ls_end) { err = "ls_end missing"; return PLBA_E_ARG; }
    if (p.n_ls && o.profile != PLBA_PROFILE_H_END && !p.ls_plk) { err = "ls_plk missing"; return PLBA_E_ARG; }
    int nfree = 0, last = -1;
    for (int i = 0; i < p.n_kf; i++) {
        const int s = p.kf_slot[i];
        if (s < -1 || s >= p.n_free) { err = "kf_slot out of range"; return PLBA_E_ARG; }
        if (s >= 0) { if (s != last + 1) { err = "kf_slot must ascend 0..n_free-1"; return PLBA_E_ARG; } last = s; nfree++; }
    }
    if (nfree != p.n_free) { err = "n_free does not match kf_slot"; return PLBA_E_ARG; }
    int bad_p = 0, bad_l = 0;      // 1 = index out of range, 2 = not landmark-major
#pragma omp parallel for num_threads(g_host_nt) schedule(static) reduction(max : bad_p) if (p.n_pobs > PAR_OBS)
    for (int i = 0; i < p.n_pobs; i++) {
        if (p.po_lm[i] < 0 || p.po_lm[i] >= p.n_pt || p.po_kf[i] < 0 || p.po_kf[i] >= p.n_kf) bad_p = bad_p > 1 ? bad_p : 1;
        else if (i && p.po_lm[i] < p.po_lm[i - 1]) bad_p = 2;
    }
#pragma omp parallel for num_threads(g_host_nt) schedule(static) reduction(max : bad_l) if (p.n_lobs > PAR_OBS)
    for (int i = 0; i < p.n_lobs; i++) {
        if (p.lo_lm[i] < 0 || p.lo_lm[i] >= p.n_ls || p.lo_kf[i] < 0 || p.lo_kf[i] >= p.n_kf) bad_l = bad_l > 1 ? bad_l : 1;
        else if (i && p.lo_lm[i] < p.lo_lm[i - 1]) bad_l = 2;
    }
    if (bad_p) { err = bad_p == 1 ? "point observation index out of range" : "point observations not landmark-major"; return PLBA_E_ARG; }
    if (bad_l) { err = bad_l == 1 ? "line observation index out of range" : "line observations not landmark-major"; return PLBA_E_ARG; }
    return PLBA_OK;
}

// One landmark class of one window: signature sort.  Index work only.
struct ClassLayout {
    std::vector<int> perm;       // new local landmark -> old local landmark
    std::vector<int> optr;       // CSR of the caller's order (local)
    std::vector<int> group;      // signature group of each OLD landmark (equal group <=> identical keyframe sequence); empty = not grouped
};
static bool g_sig_prof = false;      // PLBA_HOST_PROF: phase times of the signature order of a large window
static void signature_order(int n_lm, int n_obs, const int32_t *lm, const int32_t *kf, bool permute, ClassLayout &L) {
    const bool par = n_obs > PAR_OBS;
    auto t_sp = std::chrono::steady_clock::now();
    auto SIGPROF = [&](const char *name) { if (g_sig_prof && par) { const auto t_ = std::chrono::steady_clock::now(); std::fprintf(stderr, "[sigorder] %-10s %.3f ms\n", name, std::chrono::duration<double, std::milli>(t_ - t_sp).count()); t_sp = t_; } };      // one large window: the loops below run in parallel (a region costs more than a small window's whole sort)
    // lm[] is non-decreasing (validated): optr[l] = first observation whose landmark is >= l
    if (par) {
        L.optr.assign(n_lm + 1, n_obs);
#pragma omp parallel for num_threads(g_host_nt) schedule(static)
        for (int i = 0; i < n_obs; i++) {
            const int lo = i ? lm[i - 1] + 1 : 0;
            for (int l = lo; l <= lm[i]; l++) L.optr[l] = i;
        }
    } else {
        // serial form without data-dependent branches (tracks hold ~5 observations: a loop per track mispredicts its exit every time —
        // measured 157 -> 42 us on 36 000 observations): backwards, unconditional stores, the last one of a landmark is its first observation;
        // landmarks without observations take their successor's entry
        L.optr.assign(n_lm + 1, -1);
        L.optr[n_lm] = n_obs;
        for (int i = n_obs - 1; i >= 0; i--) L.optr[lm[i]] = i;
        for (int l = n_lm - 1; l >= 0; l--) L.optr[l] = L.optr[l] < 0 ? L.optr[l + 1] : L.optr[l];
    }
    SIGPROF("optr");
    L.perm.resize(n_lm);
#pragma omp parallel for num_threads(g_host_nt) schedule(static) if (par)
    for (int l = 0; l < n_lm; l++) L.perm[l] = l;
    SIGPROF("iota");
    if (!permute || n_lm < 2) return;
    {   // fast path: every track is a contiguous keyframe run (the usual sliding-window case) => the signature IS (first KF, length):
        // one pass over the observations and a counting sort, no hashing
        int maxlen = 0, maxkf = 0, broken = 0;
        if (par) {
#pragma omp parallel for num_threads(g_host_nt) schedule(static) reduction(max : maxlen, maxkf, broken)
            for (int l = 0; l < n_lm; l++) {
                if (broken > 0) continue;      // (a max-reduction's private copy starts at INT_MIN, not at 0)
                const int a = L.optr[l], b = L.optr[l + 1];
                if (b - a > maxlen) maxlen = b - a;
                if (b > a && kf[a] > maxkf) maxkf = kf[a];
                for (int i = a + 1; i < b; i++) if (kf[i] != kf[i - 1] + 1) { broken = 1; break; }
            }
        } else {      // the same three results from branch-free passes (one over the observations, one over the landmarks)
            for (int i = 1; i < n_obs; i++) broken |= (int)(lm[i] == lm[i - 1]) & (int)(kf[i] != kf[i - 1] + 1);
            const int last = std::max(n_obs - 1, 0);
            for (int l = 0; l < n_lm && n_obs > 0; l++) {
                const int a = L.optr[l], b = L.optr[l + 1];
                maxlen = std::max(maxlen, b - a);
                const int v = kf[std::min(a, last)];
                maxkf = std::max(maxkf, b > a ? v : 0);
            }
        }
        SIGPROF("contig");
        const bool contiguous = broken <= 0;
        if (contiguous && (int64_t)(maxkf + 2) * (maxlen + 1) <= 4 * (int64_t)n_lm + 4096) {
            const int W = maxlen + 1, nkeys = (maxkf + 2) * W;
            L.group.resize(n_lm);
            if (par && g_host_nt > 1) {
                // one large window: the same stable counting sort over landmark slices — a histogram per slice, offsets in (key, slice) order,
                // every slice scatters its own landmarks (config 4: the serial sort was half of the host preparation)
                const int T = g_host_nt; const int per = (n_lm + T - 1) / T;
                std::vector<int> cnt((size_t)T * nkeys, 0);
#pragma omp parallel for num_threads(T) schedule(static, 1)
                for (int sidx = 0; sidx < T; sidx++) {
                    int *c = cnt.data() + (size_t)sidx * nkeys;
                    for (int l = sidx * per; l < std::min(n_lm, (sidx + 1) * per); l++) {
                        const int a = L.optr[l], b = L.optr[l + 1];
                        const int key = (b > a ? kf[a] : maxkf + 1) * W + (b - a);
                        L.group[l] = key; c[key]++;
                    }
                }
                SIGPROF("hist");
                int running = 0;
                for (int k = 0; k < nkeys; k++) for (int sidx = 0; sidx < T; sidx++) { int &c = cnt[(size_t)sidx * nkeys + k]; const int v = c; c = running; running += v; }
                SIGPROF("prefix");
#pragma omp parallel for num_threads(T) schedule(static, 1)
                for (int sidx = 0; sidx < T; sidx++) {
                    int *c = cnt.data() + (size_t)sidx * nkeys;
                    for (int l = sidx * per; l < std::min(n_lm, (sidx + 1) * per); l++) L.perm[c[L.group[l]]++] = l;
                }
                SIGPROF("scatter");
                return;
            }
            std::vector<int> cnt(nkeys + 1, 0);
            for (int l = 0; l < n_lm; l++) {
                const int a = L.optr[l], b = L.optr[l + 1];
                const int key = (b > a ? kf[a] : maxkf + 1) * W + (b - a);      // landmarks without observations go last
                L.group[l] = key; cnt[key + 1]++;
            }
            for (int k = 0; k < nkeys; k++) cnt[k + 1] += cnt[k];
            for (int l = 0; l < n_lm; l++) L.perm[cnt[L.group[l]]++] = l;
            return;
        }
    }
    // group landmarks by exact signature with an open-addressing table (hash -> group), then order the groups by
    // (first keyframe, first appearance) and counting-sort the landmarks by group: O(n) instead of a comparison sort
    struct Slot { uint64_t hsh; int group, rep; };
    struct GroupPart { std::vector<Slot> table; std::vector<int> rep, count, first; };
    std::vector<int> group_of(n_lm), g_first, g_rep, g_count;
    auto same_sig = [&](int l, int r) {
        const int a = L.optr[l], b = L.optr[l + 1], c = L.optr[r];
        if (b - a != L.optr[r + 1] - c) return false;
        for (int i = 0; i < b - a; i++) if (kf[a + i] != kf[c + i]) return false;
        return true;
    };
    auto sig_hash = [&](int l) {
        const int a = L.optr[l], b = L.optr[l + 1];
        uint64_t hsh = 1469598103934665603ULL ^ (uint64_t)(b - a);
        for (int i = a; i < b; i++) { hsh ^= (uint64_t)(uint32_t)kf[i] + 0x9e3779b97f4a7c15ULL + (hsh << 6) + (hsh >> 2); hsh *= 1099511628211ULL; }
        return hsh;
    };
    std::vector<uint64_t> hashes;
    // the table walk over the landmarks whose hash falls into part `part` of `nparts` (1 = all): group ids in order of first appearance
    auto walk = [&](int part, int nparts, GroupPart &G) {
        size_t cap = 1024;
        G.table.assign(cap, Slot{0, -1, -1});
        for (int l = 0; l < n_lm; l++) {
            const uint64_t hsh = hashes.empty() ? sig_hash(l) : hashes[l];
            if (nparts > 1 && (int)((hsh >> 48) % (uint64_t)nparts) != part) continue;
            const int a = L.optr[l], b = L.optr[l + 1];
            size_t pos = (size_t)(hsh >> 17) & (cap - 1);
            int g = -1;
            while (G.table[pos].group >= 0) {
                if (G.table[pos].hsh == hsh && same_sig(l, G.table[pos].rep)) { g = G.table[pos].group; break; }
                pos = (pos + 1) & (cap - 1);
            }
            if (g < 0) {
                g = (int)G.rep.size();
                G.rep.push_back(l); G.count.push_back(0); G.first.push_back(b > a ? kf[a] : 0x7fffffff);
                G.table[pos] = Slot{hsh, g, l};
                if (2 * G.rep.size() > cap) {      // grow: re-insert the representatives
                    cap *= 4;
                    std::vector<Slot> bigger(cap, Slot{0, -1, -1});
                    for (const Slot &sl : G.table) if (sl.group >= 0) { size_t q = (size_t)(sl.hsh >> 17) & (cap - 1); while (bigger[q].group >= 0) q = (q + 1) & (cap - 1); bigger[q] = sl; }
                    G.table.swap(bigger);
                }
            }
            group_of[l] = g; G.count[g]++;
        }
    };
    if (par && g_host_nt > 1) {
        // one large window: the hashes (the pass over every observation) in parallel, then one table per hash partition, each walked by its own
        // thread over ALL landmarks in order (equal signatures have equal hashes, i.e. meet in one partition); the partitions' groups are
        // concatenated — the order of the groups below depends on (first keyframe, first landmark of the group) only, not on the group ids
        const int T = g_host_nt;
        hashes.resize(n_lm);
#pragma omp parallel for num_threads(T) schedule(static)
        for (int l = 0; l < n_lm; l++) hashes[l] = sig_hash(l);
        SIGPROF("hashes");
        std::vector<GroupPart> parts(T);
#pragma omp parallel for num_threads(T) schedule(static, 1)
        for (int pi = 0; pi < T; pi++) walk(pi, T, parts[pi]);
        std::vector<int> off(T + 1, 0);
        for (int pi = 0; pi < T; pi++) off[pi + 1] = off[pi] + (int)parts[pi].rep.size();
#pragma omp parallel for num_threads(T) schedule(static)
        for (int l = 0; l < n_lm; l++) group_of[l] += off[(int)((hashes[l] >> 48) % (uint64_t)T)];
        for (int pi = 0; pi < T; pi++) {
            g_rep.insert(g_rep.end(), parts[pi].rep.begin(), parts[pi].rep.end()); g_count.insert(g_count.end(), parts[pi].count.begin(), parts[pi].count.end());
            g_first.insert(g_first.end(), parts[pi].first.begin(), parts[pi].first.end());
        }
    } else {
        GroupPart G;
        walk(0, 1, G);
        g_rep.swap(G.rep); g_count.swap(G.count); g_first.swap(G.first);
    }
    SIGPROF("table");
    const int ng = (int)g_rep.size();
    std::vector<int> order(ng);
    for (int g = 0; g < ng; g++) order[g] = g;
    std::sort(order.begin(), order.end(), [&](int x, int y) { return g_first[x] != g_first[y] ? g_first[x] < g_first[y] : g_rep[x] < g_rep[y]; });      // (first keyframe, first appearance)
    std::vector<int> start(ng, 0);
    int acc = 0;
    for (int r = 0; r < ng; r++) { start[order[r]] = acc; acc += g_count[order[r]]; }
    for (int l = 0; l < n_lm; l++) L.perm[start[group_of[l]]++] = l;       // stable inside a group
    L.group.swap(group_of);
    SIGPROF("groups");
}

extern "C" {

int plba_version(void) { return PLBA_VERSION; }

void plba_default_options(int32_t profile, plba_options *o) {
    std::memset(o, 0, sizeof(*o));
    o->profile = profile; o->quirks = PLBA_QUIRKS_FAITHFUL;
    o->lambda_lba_lm = 1e-5; o->lambda_lba_k = 10.0; o->max_iters_lba = 15;        // src/slamConfig.cpp:65-67
    o->homog_th = 1e-7; o->min_error = 1e-7; o->min_error_change = 1e-7;           // src2/config.cpp:80-85
    o->huber_delta = (double)(float)std::sqrt(5.991);                              // src/mapHandler.cpp:5978 (Q13)
    o->chi2_gate = 5.991; o->iters_stage1 = 5; o->iters_stage2 = 10;               // :6122, :6129, :6152
    o->lm_tau = 1e-5; o->lm_max_trials = 10;                                       // g2o defaults
}

int plba_create(int32_t device, void *stream, plba_handle *out) {
    if (!out) return PLBA_E_ARG;
    plba_handle h = new plba_handle_s();
    h->device = device;
#ifndef PLBA_HOST_EMU
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device >= ndev) {
        // no silent fallback: the library is CUDA-only
        delete h; *out = nullptr; return PLBA_E_CUDA;
    }
#endif
    if (cudaSetDevice(device) != cudaSuccess) { delete h; *out = nullptr; return PLBA_E_CUDA; }
    if (stream) { h->stream = (cudaStream_t)stream; h->own_stream = false; }
    else { if (cudaStreamCreate(&h->stream) != cudaSuccess) { delete h; *out = nullptr; return PLBA_E_CUDA; } h->own_stream = true; }
#ifndef PLBA_HOST_EMU
    int nsm = 0;
    if (cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && nsm > 0) h->n_sm = nsm;
    const char *ng = std::getenv("PLBA_NO_GRAPH");
    h->no_graph = ng && ng[0] == '1';
    const char *fd = std::getenv("PLBA_FORCE_DENSE");        // large windows: always take the dense DMMA Cholesky (tests, benchmarks)
    h->force_dense = fd && fd[0] == '1';
    const char *ls = std::getenv("PLBA_LARGE_SOLVER");       // "band": the single-CTA banded Cholesky instead of the block cyclic reduction
    h->large_solver = (ls && ls[0] == 'b' && ls[1] == 'a') ? 1 : 0;
    const char *b1 = std::getenv("PLBA_BCR_V1");
    h->bcr_v1 = b1 && b1[0] == '1';
    const char *fc = std::getenv("PLBA_FORCE_CHUNK");        // always take the CTA-chunk assembly / update kernels (tests, A/B runs)
    h->force_chunk = fc ? std::atoi(fc) : 0;
#else
    h->no_graph = true;
#endif
    h->grid_chunks = 2 * h->n_sm; h->grid_solve = h->n_sm;
    void *q = nullptr;
    auto fail = [&]() {      // every error exit releases what was acquired so far
        if (h->d_P) cudaFree(h->d_P);
        if (h->d_scratch) cudaFree(h->d_scratch);
        if (h->h_P) cudaFreeHost(h->h_P);
        if (h->h_cnt) cudaFreeHost(h->h_cnt);
        if (h->h_poll) cudaFreeHost(h->h_poll);
        if (h->own_stream && h->stream) cudaStreamDestroy(h->stream);
        delete h; *out = nullptr; return PLBA_E_CUDA;
    };
    if (cudaMalloc(&q, sizeof(DevP)) != cudaSuccess) return fail();
    h->d_P = (DevP *)q;
    if (cudaMalloc(&q, 64) != cudaSuccess) return fail();
    h->d_scratch = (double *)q;
    if (cudaMallocHost(&q, sizeof(DevP)) != cudaSuccess) return fail();
    h->h_P = (DevP *)q;
    if (cudaMallocHost(&q, sizeof(int) * CNT_N) != cudaSuccess) return fail();
    h->h_cnt = (int *)q;
    cudaEventCreate(&h->ev_h2d);
    if (cudaMallocHost(&q, sizeof(int) * CNT_N * 4) != cudaSuccess) return fail();
    h->h_poll = (int *)q;
    for (int i = 0; i < 4; i++) cudaEventCreate(&h->ev_poll[i]);
    for (int i = 0; i < 8; i++) cudaEventCreate(&h->ev[i]);
#ifndef PLBA_HOST_EMU
    {   // the look-ahead panel stream of the dense tiled Cholesky: created lazily would do, but a stream and two events cost nothing
        int lo_p = 0, hi_p = 0; cudaDeviceGetStreamPriorityRange(&lo_p, &hi_p);
        if (cudaStreamCreateWithPriority(&h->stream_panel, cudaStreamNonBlocking, hi_p) != cudaSuccess) { h->stream_panel = nullptr; cudaGetLastError(); }
        for (int i = 0; i < 2; i++) cudaEventCreateWithFlags(&h->ev_la[i], cudaEventDisableTiming);
        const char *nl = std::getenv("PLBA_NO_LOOKAHEAD");
        h->no_lookahead = nl && nl[0] == '1';
        const char *no = std::getenv("PLBA_NO_OVERLAP");
        h->no_overlap = no && no[0] == '1';
    }
#endif
    {   // launch geometry of the dense tiled Cholesky (also read by the host emulation, so that the CPU tests walk every variant)
        const char *k1 = std::getenv("PLBA_DENSE_K1");
        h->dense_k1 = k1 && k1[0] == '1';
        const char *dg = std::getenv("PLBA_DENSE_GROUP");
        if (dg && std::atoi(dg) >= 1 && std::atoi(dg) <= 8) h->dense_group = std::atoi(dg);
    }
    *out = h;
    return PLBA_OK;
}

void plba_destroy(plba_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    cudaStreamSynchronize(h->stream);
#ifndef PLBA_HOST_EMU
    for (int i = 0; i < 12; i++) { if (h->gexec[i]) cudaGraphExecDestroy(h->gexec[i]); if (h->graph[i]) cudaGraphDestroy(h->graph[i]); }
#endif
    h->release();
    if (h->d_P) cudaFree(h->d_P);
    if (h->d_scratch) cudaFree(h->d_scratch);
    if (h->h_P) cudaFreeHost(h->h_P);
    if (h->h_cnt) cudaFreeHost(h->h_cnt);
    for (int i = 0; i < 8; i++) cudaEventDestroy(h->ev[i]);
    if (h->ev_h2d) cudaEventDestroy(h->ev_h2d);
    for (int i = 0; i < 4; i++) if (h->ev_poll[i]) cudaEventDestroy(h->ev_poll[i]);
    if (h->h_poll) cudaFreeHost(h->h_poll);
#ifndef PLBA_HOST_EMU
    if (h->nccl_comm) { nccl_api().CommDestroy((plba_ncclComm_t)h->nccl_comm); h->nccl_comm = nullptr; }
#endif
#ifndef PLBA_HOST_EMU
    if (h->stream_panel) { cudaStreamSynchronize(h->stream_panel); cudaStreamDestroy(h->stream_panel); }
    for (int i = 0; i < 2; i++) if (h->ev_la[i]) cudaEventDestroy(h->ev_la[i]);
#endif
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char *plba_last_error(plba_handle h) { return h ? h->err.c_str() : "null handle"; }

int plba_set_allreduce(plba_handle h, plba_allreduce_fn fn, void *user) { if (!h) return PLBA_E_ARG; h->allreduce = fn; h->allreduce_user = user; if (!fn && !h->nccl_comm) { h->n_ranks = 1; h->rank = 0; } return PLBA_OK; }
int plba_set_allreduce_ranks(plba_handle h, int32_t nranks, int32_t rank) {
    if (!h || nranks < 1 || nranks > PLBA_MAX_RANKS || rank < 0 || rank >= nranks) return PLBA_E_ARG;
    h->n_ranks = nranks; h->rank = rank; return PLBA_OK;
}
int plba_comm_info(plba_handle h, int32_t *out2) { if (!h || !out2) return PLBA_E_ARG; out2[0] = has_exchange(h) ? h->n_ranks : 1; out2[1] = h->rank; return PLBA_OK; }
#ifndef PLBA_HOST_EMU
int plba_comm_unique_id(void *id_out) {
    if (!id_out) return PLBA_E_ARG;
    NcclApi &N = nccl_api();
    if (!N.ok) return PLBA_E_UNSUPPORTED;
    plba_ncclUniqueId id;
    if (N.GetUniqueId(&id) != 0) return PLBA_E_CUDA;
    std::memcpy(id_out, id.internal, PLBA_COMM_ID_BYTES);
    return PLBA_OK;
}
int plba_comm_init_rank(plba_handle h, int32_t nranks, int32_t rank, const void *id_in) {
    if (!h || !id_in || nranks < 1 || nranks > PLBA_MAX_RANKS || rank < 0 || rank >= nranks) return PLBA_E_ARG;
    NcclApi &N = nccl_api();
    if (!N.ok) { h->err = "NCCL unavailable: " + N.why; return PLBA_E_UNSUPPORTED; }
    CK(cudaSetDevice(h->device));
    if (h->nccl_comm) { N.CommDestroy((plba_ncclComm_t)h->nccl_comm); h->nccl_comm = nullptr; }
    plba_ncclUniqueId id; std::memcpy(id.internal, id_in, PLBA_COMM_ID_BYTES);
    plba_ncclComm_t c = nullptr;
    const int rc = N.CommInitRank(&c, nranks, id, rank);
    if (rc != 0) { h->err = std::string("ncclCommInitRank: ") + N.GetErrorString(rc); return PLBA_E_CUDA; }
    h->nccl_comm = c; h->n_ranks = nranks; h->rank = rank; h->nccl_failed = 0;
    return PLBA_OK;
}
int plba_comm_destroy(plba_handle h) {
    if (!h) return PLBA_E_ARG;
    if (h->nccl_comm) { cudaSetDevice(h->device); cudaStreamSynchronize(h->stream); nccl_api().CommDestroy((plba_ncclComm_t)h->nccl_comm); h->nccl_comm = nullptr; }
    if (!h->allreduce) { h->n_ranks = 1; h->rank = 0; }
    return PLBA_OK;
}
int plba_create_group(int32_t ndev, const int32_t *devs, plba_handle *out) {
    if (ndev < 1 || ndev > PLBA_MAX_RANKS || !devs || !out) return PLBA_E_ARG;
    NcclApi &N = nccl_api();
    if (!N.ok) return PLBA_E_UNSUPPORTED;
    for (int i = 0; i < ndev; i++) out[i] = nullptr;
    for (int i = 0; i < ndev; i++) {
        const int rc = plba_create(devs[i], nullptr, &out[i]);
        if (rc) { for (int j = 0; j < i; j++) { plba_destroy(out[j]); out[j] = nullptr; } return rc; }
    }
    std::vector<plba_ncclComm_t> comms(ndev, nullptr);
    std::vector<int> dl(devs, devs + ndev);
    if (N.CommInitAll(comms.data(), ndev, dl.data()) != 0) { for (int i = 0; i < ndev; i++) { plba_destroy(out[i]); out[i] = nullptr; } return PLBA_E_CUDA; }
    for (int i = 0; i < ndev; i++) { out[i]->nccl_comm = comms[i]; out[i]->n_ranks = ndev; out[i]->rank = i; }
    return PLBA_OK;
}
void plba_destroy_group(int32_t ndev, plba_handle *hs) { if (!hs) return; for (int i = 0; i < ndev; i++) { plba_destroy(hs[i]); hs[i] = nullptr; } }
#else
int plba_comm_unique_id(void *) { return PLBA_E_UNSUPPORTED; }
int plba_comm_init_rank(plba_handle, int32_t, int32_t, const void *) { return PLBA_E_UNSUPPORTED; }
int plba_comm_destroy(plba_handle) { return PLBA_OK; }
int plba_create_group(int32_t, const int32_t *, plba_handle *) { return PLBA_E_UNSUPPORTED; }
void plba_destroy_group(int32_t, plba_handle *) {}
#endif

}  // extern "C"

// ---- launch helpers ------------------------------------------------------------------------------------------
static inline dim3 grid1(int n, int b) { return dim3((unsigned)std::max(1, (n + b - 1) / b)); }
// function attributes and occupancy are PER DEVICE: one process may hold handles on several GPUs (plba_create_group), so the
// "already done" flags and the occupancy caches are indexed by the current device (callers have done cudaSetDevice(h->device))
enum { PLBA_MAX_DEV = 64 };
static inline int cur_dev() {
#ifndef PLBA_HOST_EMU
    int d = 0; cudaGetDevice(&d); return (d >= 0 && d < PLBA_MAX_DEV) ? d : 0;
#else
    return 0;
#endif
}
template <int PROF> static void set_smem_attr() {
#ifndef PLBA_HOST_EMU
    static bool done_dev[PLBA_MAX_DEV] = {false};
    bool &done = done_dev[cur_dev()];
    if (!done) {
        cudaFuncSetAttribute(k_assemble<PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemMax<PROF>::bytes());
        cudaFuncSetAttribute(k_update<PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmemMax<PROF>::bytes());
        cudaFuncSetAttribute(k_assemble_w<PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(WARPS_PER_CTA * WSmemMax<PROF>::bytes()));
        cudaFuncSetAttribute(k_update_w<PROF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(WARPS_PER_CTA * WSmemMax<PROF>::bytes()));
        cudaFuncSetAttribute(k_solve_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_small_smem());
        cudaFuncSetAttribute(k_solve_banded, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)solve_banded_smem());
        cudaFuncSetAttribute(k_bcr_elim, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bcr_elim_smem());
        cudaFuncSetAttribute(k_bcr_back, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bcr_back_smem());
        cudaFuncSetAttribute(k_bcr_factor, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bcr_factor_smem());
        cudaFuncSetAttribute(k_bcr_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bcr_schur_smem());
        cudaFuncSetAttribute(k_potrf_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)potrf_block_smem());
        cudaFuncSetAttribute(k_trsm_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)trsm_block_smem());
        cudaFuncSetAttribute(k_syrk_dmma, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)syrk_dmma_smem());
        cudaFuncSetAttribute(k_back_block, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)back_block_smem());
        done = true;
    }
#endif
}
static void set_all_attrs() { set_smem_attr<PLBA_PROFILE_G>(); set_smem_attr<PLBA_PROFILE_H_END>(); set_smem_attr<PLBA_PROFILE_H_PLK>(); }
// persistent grid = SMs x resident CTAs per SM of the heavier of the two chunk kernels
template <int PROF> static int chunk_occupancy() {
#ifndef PLBA_HOST_EMU
    int a = 1, b = 1;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_assemble<PROF>, OC, SmemMax<PROF>::bytes());
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, k_update<PROF>, OC, SmemMax<PROF>::bytes());
    return std::max(1, std::min(a, b));
#else
    return 2;
#endif
}
static int chunk_occupancy_for(int prof) {
    static int cache_dev[PLBA_MAX_DEV][3] = {{0}};
    int *cache = cache_dev[cur_dev()];
    if (cache[prof]) return cache[prof];
    return cache[prof] = prof == PLBA_PROFILE_G ? chunk_occupancy<PLBA_PROFILE_G>() : prof == PLBA_PROFILE_H_END ? chunk_occupancy<PLBA_PROFILE_H_END>() : chunk_occupancy<PLBA_PROFILE_H_PLK>();
}
// warp path: persistent grid of WARPS_PER_CTA-warp CTAs (emulation: one warp per "CTA")
// (work items are dealt dynamically, so each kernel simply gets as many CTAs as can be resident)
template <int PROF> static int warp_occupancy(int update) {
#ifndef PLBA_HOST_EMU
    int a = 1;
    if (update) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_update_w<PROF>, WNT, WARPS_PER_CTA * WSmemMax<PROF>::bytes());
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, k_assemble_w<PROF>, WNT, WARPS_PER_CTA * WSmemMax<PROF>::bytes());
    return std::max(1, a);
#else
    (void)update; return 1;
#endif
}
static int warp_occupancy_for(int prof, int update) {
    static int cache_dev[PLBA_MAX_DEV][6] = {{0}};
    int &c = cache_dev[cur_dev()][2 * prof + update];
    if (c) return c;
    return c = prof == PLBA_PROFILE_G ? warp_occupancy<PLBA_PROFILE_G>(update) : prof == PLBA_PROFILE_H_END ? warp_occupancy<PLBA_PROFILE_H_END>(update) : warp_occupancy<PLBA_PROFILE_H_PLK>(update);
}
static size_t warp_smem(int prof) {
    const size_t w = prof == PLBA_PROFILE_G ? WSmemMax<PLBA_PROFILE_G>::bytes() : prof == PLBA_PROFILE_H_END ? WSmemMax<PLBA_PROFILE_H_END>::bytes() : WSmemMax<PLBA_PROFILE_H_PLK>::bytes();
#ifndef PLBA_HOST_EMU
    return WARPS_PER_CTA * w;
#else
    return w;
#endif
}
#ifndef PLBA_HOST_EMU
static inline dim3 warp_block() { return dim3(WNT); }
#else
static inline dim3 warp_block() { return dim3(32); }
#endif
static size_t chunk_smem(int prof) {
    return prof == PLBA_PROFILE_G ? SmemMax<PLBA_PROFILE_G>::bytes() : prof == PLBA_PROFILE_H_END ? SmemMax<PLBA_PROFILE_H_END>::bytes() : SmemMax<PLBA_PROFILE_H_PLK>::bytes();
}
static void launch_assemble(plba_handle h, int mode) {
    const DevP *Pp = h->d_P;
    if (h->warp_path) {
        const dim3 g(h->grid_warp), b = warp_block(); const size_t sm = warp_smem(h->opt.profile);
        switch (h->opt.profile) {
        case PLBA_PROFILE_G: PLBA_LAUNCH((k_assemble_w<PLBA_PROFILE_G>), g, b, sm, h->stream, Pp, mode); break;
        case PLBA_PROFILE_H_END: PLBA_LAUNCH((k_assemble_w<PLBA_PROFILE_H_END>), g, b, sm, h->stream, Pp, mode); break;
        default: PLBA_LAUNCH((k_assemble_w<PLBA_PROFILE_H_PLK>), g, b, sm, h->stream, Pp, mode); break;
        }
        h->timing.n_launches++; if (mode == 1) h->timing.n_assemble++;
        return;
    }
    const dim3 g(h->grid_chunks), b(OC); const size_t sm = chunk_smem(h->opt.profile);
    switch (h->opt.profile) {
    case PLBA_PROFILE_G: PLBA_LAUNCH((k_assemble<PLBA_PROFILE_G>), g, b, sm, h->stream, Pp, mode); break;
    case PLBA_PROFILE_H_END: PLBA_LAUNCH((k_assemble<PLBA_PROFILE_H_END>), g, b, sm, h->stream, Pp, mode); break;
    default: PLBA_LAUNCH((k_assemble<PLBA_PROFILE_H_PLK>), g, b, sm, h->stream, Pp, mode); break;
    }
    h->timing.n_launches++; if (mode == 1) h->timing.n_assemble++;
}
static void launch_update(plba_handle h, int flags) {
    const DevP *Pp = h->d_P;
    if (h->warp_path) {
        const dim3 g(h->grid_warp_upd), b = warp_block(); const size_t sm = warp_smem(h->opt.profile);
        switch (h->opt.profile) {
        case PLBA_PROFILE_G: PLBA_LAUNCH((k_update_w<PLBA_PROFILE_G>), g, b, sm, h->stream, Pp, flags); break;
        case PLBA_PROFILE_H_END: PLBA_LAUNCH((k_update_w<PLBA_PROFILE_H_END>), g, b, sm, h->stream, Pp, flags); break;
        default: PLBA_LAUNCH((k_update_w<PLBA_PROFILE_H_PLK>), g, b, sm, h->stream, Pp, flags); break;
        }
        h->timing.n_launches++;
        return;
    }
    const dim3 g(h->grid_chunks), b(OC); const size_t sm = chunk_smem(h->opt.profile);
    switch (h->opt.profile) {
    case PLBA_PROFILE_G: PLBA_LAUNCH((k_update<PLBA_PROFILE_G>), g, b, sm, h->stream, Pp, flags); break;
    case PLBA_PROFILE_H_END: PLBA_LAUNCH((k_update<PLBA_PROFILE_H_END>), g, b, sm, h->stream, Pp, flags); break;
    default: PLBA_LAUNCH((k_update<PLBA_PROFILE_H_PLK>), g, b, sm, h->stream, Pp, flags); break;
    }
    h->timing.n_launches++;
}
static void allreduce(plba_handle h, double *p, size_t n, int op_max) {
    if (!n) return;
#ifndef PLBA_HOST_EMU
    if (h->nccl_comm) {      // in place, on the handle's stream: ordered with the kernels around it, no host synchronisation
        const int rc = nccl_api().AllReduce(p, p, n, PLBA_NCCL_F64, op_max ? PLBA_NCCL_MAX : PLBA_NCCL_SUM, (plba_ncclComm_t)h->nccl_comm, h->stream);
        if (rc != 0 && !h->nccl_failed) { h->nccl_failed = 1; h->err = std::string("ncclAllReduce: ") + nccl_api().GetErrorString(rc); }
        return;
    }
#endif
    if (!h->allreduce) return;
    // op is encoded in the sign of the count for the max reduction (only the band width at upload uses it)
    h->allreduce(p, op_max ? -(int64_t)n : (int64_t)n, (void *)h->stream, h->allreduce_user);
}
static bool bcr_is_active(plba_handle h) { return h->bcr_layout; }      // decided at upload: the window's S is stored (and assembled) in node form
static void launch_solve(plba_handle h) {
    const DevP &P = h->P; const DevP *Pp = h->d_P;
    if (P.n_free == 0) {
        // every keyframe fixed: nothing to factor, but the hand LM's pre-solve controller (error normalisation, stop tests, apply decision)
        // still runs, exactly as on the graph path where k_solve_small handles n == 0
        if (P.profile != PLBA_PROFILE_G) { PLBA_LAUNCH(k_control_h_pre, grid1(P.n_win, 128), dim3(128), 0, h->stream, Pp); h->timing.n_launches++; }
        return;
    }
    if (h->small_path) {
        PLBA_LAUNCH(k_solve_small, dim3(h->grid_solve), dim3(SS_NT), solve_small_smem(), h->stream, Pp);
        h->timing.n_launches++;
        return;
    }
    if (P.profile != PLBA_PROFILE_G) { PLBA_LAUNCH(k_control_h_pre, grid1(P.n_win, 128), dim3(128), 0, h->stream, Pp); h->timing.n_launches++; }
    if (h->bcr_layout) {
        // block-banded reduced camera system (no loop closure in the window): block cyclic reduction over nodes of >= band keyframes,
        // ceil(log2 N) levels down and up, every level one launch with a CTA per eliminated node
        for (int w = 0; w < P.n_win; w++) {
            const BcrW &B = h->bcr[w];
            if (B.N == 0 || (h->wins[w].n_pobs + h->wins[w].n_lobs == 0 && !has_exchange(h))) continue;      // (a rank with an EMPTY SHARD still solves the summed system: every rank must take the same step)
            // (the assembly kernels have accumulated straight into the node form, and the exchange step of the sharded path has summed
            //  it over the ranks: config 5 moves 17 MB, there is no dense S)
            int s_top = 0;
            for (int s = 1; s < B.N; s *= 2) {
                const int ne = (B.N + s - 1) / (2 * s);
                if (h->bcr_v1) { PLBA_LAUNCH(k_bcr_elim, dim3(ne), dim3(BCR_NT), bcr_elim_smem(), h->stream, Pp, w, B, s, 0); h->timing.n_launches++; }
                else {
                    // a node's elimination over seven CTAs: two factor its diagonal block with one coupling block each, five form the Schur products
                    PLBA_LAUNCH(k_bcr_factor, dim3(2 * ne), dim3(BF_NT), bcr_factor_smem(), h->stream, Pp, w, B, s, 0);
                    PLBA_LAUNCH(k_bcr_schur, dim3(BS_PARTS * ne), dim3(BS_NT), bcr_schur_smem(), h->stream, Pp, w, B, s); h->timing.n_launches += 2;
                }
                s_top = s;
            }
            if (h->bcr_v1) PLBA_LAUNCH(k_bcr_elim, dim3(1), dim3(BCR_NT), bcr_elim_smem(), h->stream, Pp, w, B, 0, 1);
            else PLBA_LAUNCH(k_bcr_factor, dim3(1), dim3(BF_NT), bcr_factor_smem(), h->stream, Pp, w, B, 0, 1);
            PLBA_LAUNCH(k_bcr_back, dim3(1), dim3(256), bcr_back_smem(), h->stream, Pp, w, B, 0, 1); h->timing.n_launches += 2;
            for (int s = s_top; s >= 1; s /= 2) {
                PLBA_LAUNCH(k_bcr_back, dim3((B.N + s - 1) / (2 * s)), dim3(256), bcr_back_smem(), h->stream, Pp, w, B, s, 0); h->timing.n_launches++;
            }
        }
        PLBA_LAUNCH(k_pose_update, grid1(P.n_free, 128), dim3(128), 0, h->stream, Pp); h->timing.n_launches++;
        return;
    }
    if (h->band_blocks <= BAND_MAX && !h->force_dense) {
        // the same, factored panel by panel by one CTA per window (kept for A/B runs: PLBA_LARGE_SOLVER=band)
        PLBA_LAUNCH(k_solve_banded, dim3(std::min(P.n_win, h->n_sm)), dim3(256), solve_banded_smem(), h->stream, Pp, h->band_blocks);
        PLBA_LAUNCH(k_pose_update, grid1(P.n_free, 128), dim3(128), 0, h->stream, Pp); h->timing.n_launches += 2;
        return;
    }
    // Dense tiled Cholesky with ONE STEP OF LOOK-AHEAD: the trailing update of step k is launched as (a) the tile row that holds the rows of
    // panel k + 1 and (b) the rest; panel k + 1 (diagonal block, block row, right-hand side) starts on a second, high-priority stream as soon
    // as (a) is done and runs under (b) — the panel chain (one CTA, then one wave of small CTAs) no longer idles the machine 125 times.
    // (b) writes rows >= lo + 128 only, the panel reads / writes rows lo .. lo + 95: disjoint; both read the finished rows of panel k.
    for (int w = 0; w < P.n_win; w++) {
        const int n = 6 * h->wins[w].n_free;
        if (n == 0 || (h->wins[w].n_pobs + h->wins[w].n_lobs == 0 && !has_exchange(h))) continue;
        cudaStream_t sm = h->stream, sp = h->stream;
#ifndef PLBA_HOST_EMU
        if (h->stream_panel && !h->no_lookahead) sp = h->stream_panel;
#endif
        const bool la = (sp != sm);
        double *Sw = P.S + h->win_S_off[w], *xw = P.xp + (size_t)6 * h->wins[w].slot0;      // y (forward substitution) and then x live in xp
        auto launch_panel = [&](cudaStream_t st, int k0) {
            const int nb = std::min((int)NBK, n - k0), lo = k0 + nb;
            PLBA_LAUNCH(k_potrf_block, dim3(1), dim3(PB_NT), potrf_block_smem(), st, Pp, Sw, n, P.ctrl + w, (const double *)(P.hpp_diag + (size_t)6 * h->wins[w].slot0), P.profile == PLBA_PROFILE_G ? 1 : 0, k0, nb);
            // columns right of the block + the right-hand side column
            PLBA_LAUNCH(k_trsm_block, dim3((n - lo + 1 + TRSM_COLS - 1) / TRSM_COLS), dim3(TRSM_COLS), trsm_block_smem(), st, Pp, Sw, n, xw, k0, nb);
            h->timing.n_launches += 2;
            if (lo < n) { PLBA_LAUNCH(k_rhs_update, dim3((n - lo + 255) / 256, (nb + RHS_ROWS - 1) / RHS_ROWS), dim3(256), 0, st, Pp, w, k0, nb); h->timing.n_launches++; }
        };
        auto launch_syrk = [&](cudaStream_t st, int k0, int nb, int lo, int tr0, int tr1, int rmax) {
            const int g = syrk_tiles(n - lo, tr0, tr1);
            if (g > 0) { PLBA_LAUNCH(k_syrk_dmma, dim3(g), dim3(256), syrk_dmma_smem(), st, Pp, Sw, n, k0, nb, lo, tr0, rmax); h->timing.n_launches++; }
        };
        PLBA_LAUNCH(k_rhs_init, grid1(n, 256), dim3(256), 0, sm, Pp, w); h->timing.n_launches++;
        launch_panel(sm, 0);
        if (h->dense_k1) {
            // first version of the look-ahead (A/B: PLBA_DENSE_K1=1): every panel updates the whole trailing matrix on its own (K = 96 per visit of a tile)
            for (int k0 = 0; k0 < n; k0 += NBK) {
                const int nb = std::min((int)NBK, n - k0), lo = k0 + nb;
                if (lo >= n) break;
                const int TM = (n - lo + STM - 1) / STM;
                launch_syrk(sm, k0, nb, lo, 0, 1, n);
#ifndef PLBA_HOST_EMU
                if (la) { cudaEventRecord(h->ev_la[0], sm); cudaStreamWaitEvent(sp, h->ev_la[0], 0); }
#endif
                launch_syrk(sm, k0, nb, lo, 1, TM, n);
                launch_panel(sp, lo);
#ifndef PLBA_HOST_EMU
                if (la) { cudaEventRecord(h->ev_la[1], sp); cudaStreamWaitEvent(sm, h->ev_la[1], 0); }
#endif
            }
        } else {
            // PANEL GROUPS of G: inside a group a panel updates only the rows of the group's remaining panels (<= 96 (G - 1) rows, K = 96);
            // the trailing matrix right of / below the group is visited ONCE per group with K = 96 G — 1 / G of the read-modify-write
            // traffic and of the pipeline fills of the C tiles.  Look-ahead: the tile rows that hold the next group go first, then the next
            // group (its panels and their in-group updates) runs on the panel stream under the rest of the update, which touches later rows only.
            const int G = h->dense_group > 0 ? h->dense_group : (n >= 4800 ? 3 : 2);      // measured at n = 12 000: G = 1 30.9 ms, 2 28.8, 3 28.3, 4 28.5, 6 28.7; n = 1 200: 2 is best
            cudaStream_t pw = sm;                                   // stream of the current group's panel work (the first group has nothing to hide under)
            for (int kG = 0; kG < n; kG += G * NBK) {
                const int gEnd = std::min(n, kG + G * NBK);
                for (int k0 = kG; k0 < gEnd; k0 += NBK) {
                    if (k0 > kG) launch_panel(pw, k0);              // (the group's first panel is already in flight: launched before the loop / at the end of the previous iteration)
                    const int lo = k0 + NBK;
                    if (lo < gEnd) launch_syrk(pw, k0, NBK, lo, 0, (gEnd - lo + STM - 1) / STM, gEnd);      // rows of the group's remaining panels only
                }
#ifndef PLBA_HOST_EMU
                if (pw != sm) { cudaEventRecord(h->ev_la[1], pw); cudaStreamWaitEvent(sm, h->ev_la[1], 0); }
#endif
                if (gEnd >= n) break;
                const int TM = (n - gEnd + STM - 1) / STM, TA = std::min(TM, (std::min(G * NBK, n - gEnd) + STM - 1) / STM);
                launch_syrk(sm, kG, gEnd - kG, gEnd, 0, TA, n);     // the tile rows of the next group
#ifndef PLBA_HOST_EMU
                if (la) { cudaEventRecord(h->ev_la[0], sm); cudaStreamWaitEvent(sp, h->ev_la[0], 0); }
#endif
                launch_syrk(sm, kG, gEnd - kG, gEnd, TA, TM, n);
                pw = sp;
                launch_panel(pw, gEnd);                             // first panel of the next group
            }
#ifndef PLBA_HOST_EMU
            if (pw != sm) { cudaEventRecord(h->ev_la[1], pw); cudaStreamWaitEvent(sm, h->ev_la[1], 0); }      // join (the last panel ran on the panel stream)
#endif
        }
        // backward substitution, column-oriented: solve block k (one CTA, the 96 x 96 triangle), then take it out of every row above
        for (int k0 = ((n - 1) / NBK) * NBK; k0 >= 0; k0 -= NBK) {
            const int nb = std::min((int)NBK, n - k0);
            PLBA_LAUNCH(k_back_block, dim3(1), dim3(256), back_block_smem(), sm, Pp, (const double *)Sw, n, xw, (const WinCtrl *)(P.ctrl + w), k0, nb);
            h->timing.n_launches++;
            if (k0 > 0) { PLBA_LAUNCH(k_back_update, dim3((k0 + BU_ROWS - 1) / BU_ROWS), dim3(2 * BU_ROWS), 0, sm, Pp, (const double *)Sw, n, xw, k0, nb); h->timing.n_launches++; }
        }
    }
    PLBA_LAUNCH(k_pose_update, grid1(P.n_free, 128), dim3(128), 0, h->stream, Pp); h->timing.n_launches++;
}

static int poll_counters(plba_handle h) {
    CK(cudaMemcpyAsync(h->h_cnt, h->P.counters, sizeof(int) * CNT_N, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < CNT_N; i++) h->h_counters[i] = h->h_cnt[i];
    return PLBA_OK;
}
PLBA_KERNEL void k_set_lambda(const DevP *Pp, double lambda) {
    PLBA_PARAMS(P, Pp);
    PHASE_BEGIN
        for (int w = PLBA_BID * PLBA_NT + tid; w < P.n_win; w += PLBA_NB * PLBA_NT) { P.ctrl[w].lambda = lambda; P.ctrl[w].need_init = 0; }
    PHASE_END
}

#ifndef PLBA_HOST_EMU
// ---- the LM loop as a CUDA graph --------------------------------------------------------------------------------
static cudaError_t add_kernel(cudaGraph_t g, cudaGraphNode_t *node, cudaGraphNode_t dep, void *fn, dim3 grid, dim3 block, size_t smem, void **args) {
    cudaKernelNodeParams kp{};
    kp.func = fn; kp.gridDim = grid; kp.blockDim = block; kp.sharedMemBytes = (unsigned)smem; kp.kernelParams = args; kp.extra = nullptr;
    return cudaGraphAddKernelNode(node, g, dep ? &dep : nullptr, dep ? 1 : 0, &kp);
}
template <int PROF> static int build_graph(plba_handle h) {
    const int pi = (PROF * 2 + h->solve_class) * 2 + (h->warp_path ? 1 : 0);
    if (h->gexec[pi]) return PLBA_OK;
    set_all_attrs();
    cudaGraph_t g = nullptr;
    CK(cudaGraphCreate(&g, 0));
    CK(cudaGraphConditionalHandleCreate(&h->cond_while[pi], g, 1, cudaGraphCondAssignDefault));
    CK(cudaGraphConditionalHandleCreate(&h->cond_prep[pi], g, 1, cudaGraphCondAssignDefault));
    cudaGraphNodeParams wp{}; wp.type = cudaGraphNodeTypeConditional;
    wp.conditional.handle = h->cond_while[pi]; wp.conditional.type = cudaGraphCondTypeWhile; wp.conditional.size = 1;
    cudaGraphNode_t wnode;
    CK(cudaGraphAddNode(&wnode, g, nullptr, 0, &wp));
    cudaGraph_t body = wp.conditional.phGraph_out[0];
    cudaGraphNodeParams ip{}; ip.type = cudaGraphNodeTypeConditional;
    ip.conditional.handle = h->cond_prep[pi]; ip.conditional.type = cudaGraphCondTypeIf; ip.conditional.size = 1;
    cudaGraphNode_t inode;
    CK(cudaGraphAddNode(&inode, body, nullptr, 0, &ip));
    cudaGraph_t prep = ip.conditional.phGraph_out[0];
    const DevP *Pp = h->d_P;
    const dim3 gc(h->warp_path ? h->n_sm * warp_occupancy<PROF>(0) : h->n_sm * chunk_occupancy<PROF>()), bc(h->warp_path ? (int)WNT : (int)OC);
    const dim3 gu(h->warp_path ? h->n_sm * warp_occupancy<PROF>(1) : h->n_sm * chunk_occupancy<PROF>());
    const size_t smc = h->warp_path ? WARPS_PER_CTA * WSmemMax<PROF>::bytes() : SmemMax<PROF>::bytes();
    void *f_asm = h->warp_path ? (void *)k_assemble_w<PROF> : (void *)k_assemble<PROF>, *f_upd = h->warp_path ? (void *)k_update_w<PROF> : (void *)k_update<PROF>;
    int mode0 = 0, mode1 = 1, fl = KF_FUSE_CONTROL | KF_IN_GRAPH;
    void *a_p[1] = {(void *)&Pp}, *a_m0[2] = {(void *)&Pp, (void *)&mode0}, *a_m1[2] = {(void *)&Pp, (void *)&mode1}, *a_fl[2] = {(void *)&Pp, (void *)&fl};
    cudaGraphNode_t n1, n2, n3, m1, m2, m3;
    CK(add_kernel(prep, &n1, nullptr, (void *)k_gate, dim3(h->grid_chunks), dim3(256), 0, a_p));
    CK(add_kernel(prep, &n2, n1, f_asm, gc, bc, smc, a_m0));
    CK(add_kernel(prep, &n3, n2, (void *)k_lambda_init, dim3(h->n_sm), dim3(128), 0, a_p));
    CK(add_kernel(body, &m1, inode, f_asm, gc, bc, smc, a_m1));
    CK(add_kernel(body, &m2, m1, (void *)k_solve_small, dim3(h->grid_solve), dim3(SS_NT), solve_small_smem(), a_p));
    // single small window, profile G, CTA-chunk kernels: the update kernel runs BESIDE the solver (programmatic edge: launched as soon as the
    // solver's one CTA is resident) and waits for the solver's flag where it first needs x_p — its re-linearisation at the old state and the
    // kernel boundary disappear from the trial's critical path.  PLBA_NO_OVERLAP=1 keeps the plain chain (A/B runs).
    const bool overlap = (PROF == PLBA_PROFILE_G) && h->solve_class == 0 && !h->warp_path && !h->no_overlap;
    int fl_ov = fl | KF_WAIT_SOLVE;
    void *a_flov[2] = {(void *)&Pp, (void *)&fl_ov};
    if (overlap) {
        CK(add_kernel(body, &m3, nullptr, f_upd, gu, bc, smc, a_flov));
        cudaGraphEdgeData ed{};
        ed.from_port = cudaGraphKernelNodePortProgrammatic; ed.to_port = 0; ed.type = cudaGraphDependencyTypeProgrammatic;
        CK(cudaGraphAddDependencies_v2(body, &m2, &m3, &ed, 1));
    } else
    CK(add_kernel(body, &m3, m2, f_upd, gu, bc, smc, a_fl));
#ifdef PLBA_GRAPH_TWO_TRIALS
    // a second LM trial inside the same WHILE iteration: the loop's per-iteration cost (evaluating the condition, re-launching the
    // body) is paid once per two trials.  Its IF(prep) node has a handle of its own, set by the controller of the first trial; when
    // the first trial finished the last window the second one runs as three empty launches, once per LBA.
    {
        CK(cudaGraphConditionalHandleCreate(&h->cond_prep2[pi], g, 1, cudaGraphCondAssignDefault));
        cudaGraphNodeParams ip2{}; ip2.type = cudaGraphNodeTypeConditional;
        ip2.conditional.handle = h->cond_prep2[pi]; ip2.conditional.type = cudaGraphCondTypeIf; ip2.conditional.size = 1;
        cudaGraphNode_t inode2;
        cudaGraphNode_t deps2[2] = {m3, m2};
        CK(cudaGraphAddNode(&inode2, body, deps2, 2, &ip2));
        cudaGraph_t prep2 = ip2.conditional.phGraph_out[0];
        cudaGraphNode_t q1, q2, q3, r1, r2, r3;
        CK(add_kernel(prep2, &q1, nullptr, (void *)k_gate, dim3(h->grid_chunks), dim3(256), 0, a_p));
        CK(add_kernel(prep2, &q2, q1, f_asm, gc, bc, smc, a_m0));
        CK(add_kernel(prep2, &q3, q2, (void *)k_lambda_init, dim3(h->n_sm), dim3(128), 0, a_p));
        CK(add_kernel(body, &r1, inode2, f_asm, gc, bc, smc, a_m1));
        CK(add_kernel(body, &r2, r1, (void *)k_solve_small, dim3(h->grid_solve), dim3(SS_NT), solve_small_smem(), a_p));
        if (overlap) {
            CK(add_kernel(body, &r3, nullptr, f_upd, gu, bc, smc, a_flov));
            cudaGraphEdgeData ed{};
            ed.from_port = cudaGraphKernelNodePortProgrammatic; ed.to_port = 0; ed.type = cudaGraphDependencyTypeProgrammatic;
            CK(cudaGraphAddDependencies_v2(body, &r2, &r3, &ed, 1));
        } else
        CK(add_kernel(body, &r3, r2, f_upd, gu, bc, smc, a_fl));
    }
#endif
    CK(cudaGraphInstantiate(&h->gexec[pi], g, 0));
    h->graph[pi] = g;
    return PLBA_OK;
}
static int build_graph_for(plba_handle h, int prof) {
    return prof == PLBA_PROFILE_G ? build_graph<PLBA_PROFILE_G>(h) : prof == PLBA_PROFILE_H_END ? build_graph<PLBA_PROFILE_H_END>(h) : build_graph<PLBA_PROFILE_H_PLK>(h);
}
#endif

static bool use_graph(plba_handle h) {
#ifdef PLBA_HOST_EMU
    (void)h; return false;
#else
    return h->small_path && !h->no_graph && !has_exchange(h) && !h->detail_timing;
#endif
}

static int launch_reset(plba_handle h, int gather = 0) {
    h->klaunch_seen = 0;
    PLBA_LAUNCH(k_reset, dim3(h->grid_chunks), dim3(256), 0, h->stream, (const DevP *)h->d_P, h->ls_dim, h->sys_doubles, h->sysbuf, gather);
    h->timing.n_launches++;
    return PLBA_OK;
}

extern "C" {

int plba_upload(plba_handle h, int32_t n, const plba_problem *probs, const plba_options *opt) {
    if (!h || n <= 0 || !probs || !opt) return PLBA_E_ARG;
    if (opt->profile < PLBA_PROFILE_G || opt->profile > PLBA_PROFILE_H_PLK) { h->err = "unknown profile"; return PLBA_E_ARG; }
    if (opt->shell != PLBA_SHELL_LBA && !(opt->shell == PLBA_SHELL_GBA && opt->profile == PLBA_PROFILE_H_END)) { h->err = "the GBA shell exists for profile H_END only"; return PLBA_E_ARG; }
    CK(cudaSetDevice(h->device));
    if (h->h2d_pending) { CK(cudaEventSynchronize(h->ev_h2d)); h->h2d_pending = false; }   // the previous upload's H2D copies may still be reading h_in / h_P
    const auto t_host0 = std::chrono::steady_clock::now();
    auto t_hp = t_host0; const bool hostprof = std::getenv("PLBA_HOST_PROF") != nullptr;
    g_sig_prof = hostprof;
#define HOSTPROF(name) do { if (hostprof) { const auto t_ = std::chrono::steady_clock::now(); std::fprintf(stderr, "[upload] %-10s %.3f ms\n", name, std::chrono::duration<double, std::milli>(t_ - t_hp).count()); t_hp = t_; } } while (0)
    h->uploaded = false;
    h->opt = *opt;
    h->timing = plba_timing{};
    const int prof = opt->profile;
    h->ls_dim = (prof == PLBA_PROFILE_H_END) ? 6 : 4;
    const int ld = h->ls_dim;
    h->wins.assign(n, WinInfo{});
    WinInfo tot{};
    int max_nf = 0;
    {   // host threads of this upload (see host_threads_for); a batch validates its windows in parallel, the first bad window (in order) reports
        int64_t nobs_hint = 0;
        for (int w = 0; w < n; w++) nobs_hint += (int64_t)std::max(probs[w].n_pobs, 0) + std::max(probs[w].n_lobs, 0);
        g_host_nt = host_threads_for(nobs_hint);
        int first_bad = n, rc_bad = 0;
#pragma omp parallel for num_threads(g_host_nt) schedule(dynamic, 8) if (n > 8)
        for (int w = 0; w < n; w++) {
            std::string e;
            const int rc = validate_problem(probs[w], *opt, e);
            if (rc) {
#pragma omp critical(plba_validate)
                if (w < first_bad) { first_bad = w; rc_bad = rc; h->err = e; }
            }
        }
        if (first_bad < n) return rc_bad;
    }
    for (int w = 0; w < n; w++) {
        const plba_problem &p = probs[w];
        WinInfo &wi = h->wins[w];
        wi.n_kf = p.n_kf; wi.n_free = p.n_free; wi.n_pt = p.n_pt; wi.n_ls = p.n_ls; wi.n_pobs = p.n_pobs; wi.n_lobs = p.n_lobs;
        wi.kf0 = tot.n_kf; wi.slot0 = tot.n_free; wi.pt0 = tot.n_pt; wi.ls0 = tot.n_ls; wi.po0 = tot.n_pobs; wi.lo0 = tot.n_lobs;
        if ((int64_t)tot.n_pobs + p.n_pobs > 0x7fffffff || (int64_t)tot.n_lobs + p.n_lobs > 0x7fffffff) { h->err = "batch too large"; return PLBA_E_ARG; }
        tot.n_kf += p.n_kf; tot.n_free += p.n_free; tot.n_pt += p.n_pt; tot.n_ls += p.n_ls; tot.n_pobs += p.n_pobs; tot.n_lobs += p.n_lobs;
        max_nf = std::max(max_nf, p.n_free);
        for (int i = 0; i < 4; i++) if (p.cam[i] != probs[0].cam[i]) { h->err = "all windows of a batch must share the camera"; return PLBA_E_UNSUPPORTED; }
    }
    HOSTPROF("validate");
    h->max_nf = max_nf;
    h->small_path = (6 * max_nf <= SMALL_NMAX);
    // solver grid: ONE CTA for a single window (the drop-in call), one CTA per SM for a batch; the two cases are separate cached graphs
    // ("solve class"), so that nothing inside a graph depends on the problem
    h->solve_class = (n == 1) ? 0 : 1;
    h->grid_solve = (n == 1) ? 1 : h->n_sm;
    set_all_attrs();
    h->grid_chunks = h->n_sm * chunk_occupancy_for(prof);
#ifndef PLBA_HOST_EMU
    h->grid_warp = h->n_sm * warp_occupancy_for(prof, 0); h->grid_warp_upd = h->n_sm * warp_occupancy_for(prof, 1);
#else
    h->grid_warp = h->grid_warp_upd = 24;               // emulation: 24 one-warp "CTAs"
#endif

    // ---- index work: signature order, chunks, segments --------------------------------------------------------
    std::vector<Chunk> &ch_pt = h->ch_pt, &ch_ls = h->ch_ls; std::vector<Seg> &sg_pt = h->sg_pt, &sg_ls = h->sg_ls; std::vector<int> &fp_pt = h->fp_pt, &fp_ls = h->fp_ls;
    ch_pt.clear(); ch_ls.clear(); sg_pt.clear(); sg_ls.clear(); fp_pt.clear(); fp_ls.clear();
    g_host_nt = host_threads_for((int64_t)tot.n_pobs + tot.n_lobs);
    h->pt_perm.resize(tot.n_pt); h->ls_perm.resize(tot.n_ls); h->po_perm.resize(tot.n_pobs); h->lo_perm.resize(tot.n_lobs);
    std::vector<int> &pt_ptr = h->pt_ptr, &ls_ptr = h->ls_ptr;
    pt_ptr.assign(tot.n_pt + 1, 0); ls_ptr.assign(tot.n_ls + 1, 0);
    bool too_long = false;
    std::vector<ClassLayout> Lps(n), Lls(n);
    // batches: parallel over windows; one LARGE window: parallel inside signature_order
#pragma omp parallel for num_threads(g_host_nt) schedule(dynamic, 1) if (n > 8)
    for (int wc = 0; wc < 2 * n; wc++) {
        const int w = wc >> 1;
        const plba_problem &p = probs[w];
        if (wc & 1) signature_order(p.n_ls, p.n_lobs, p.lo_lm, p.lo_kf, prof != PLBA_PROFILE_H_END, Lls[w]);   // Q3 addresses endpoint lines by position
        else signature_order(p.n_pt, p.n_pobs, p.po_lm, p.po_kf, true, Lps[w]);
    }
    HOSTPROF("sigorder");
    // route: the warp-autonomous kernels take windows whose longest track fits a warp; anything longer (or force_chunk) runs on the
    // CTA-chunk kernels
    int max_track = 0;
    for (int w = 0; w < n; w++) for (int cls = 0; cls < 2; cls++) {
        const ClassLayout &L = cls ? Lls[w] : Lps[w];
        const int nl1 = (int)L.optr.size() - 1;
#pragma omp parallel for num_threads(g_host_nt) schedule(static) reduction(max : max_track) if (nl1 > PAR_LM)
        for (int l = 0; l < nl1; l++) max_track = std::max(max_track, L.optr[l + 1] - L.optr[l]);
    }
    if (n == 1) HOSTPROF("ch.maxtrk");
    // routing by size, measured (profiles/README.md r01f-h, tools/route_sweep.py): up to ~0.7 M observations the two implementations are
    // within 4 % of each other with the CTA-chunk kernels ahead (a small upload is latency-bound and their warps share one instruction
    // stream); from ~1 M observations on the warp kernels win (config 3: 85 against 119 ms, config 5: assembly 1.9 against 2.8 ms)
    const int64_t n_obs_total = (int64_t)tot.n_pobs + tot.n_lobs;
    const bool big = n_obs_total >= (int64_t)6500 * h->n_sm;
    h->warp_path = h->force_chunk != 1 && max_track <= W_MAX_TRACK && (big || h->force_chunk == 2);
    std::vector<WItem> &wi_pt = h->wi_pt, &wi_ls = h->wi_ls;
    wi_pt.clear(); wi_ls.clear();
    // first segment of every (class, window) in sg_pt / sg_ls (the layout statistics walk the windows in parallel)
    std::vector<int> seg_begin[2]; seg_begin[0].assign(n + 1, 0); seg_begin[1].assign(n + 1, 0);
    // a batch builds the runs of its windows in parallel into per-(window, class) lists, concatenated in window order below
    std::vector<std::vector<Seg>> loc_sg(h->warp_path ? 2 * (size_t)n : 0); std::vector<std::vector<int>> loc_fp(h->warp_path ? 2 * (size_t)n : 0);
#pragma omp parallel for num_threads(g_host_nt) schedule(dynamic, 1) if (n > 8)
    for (int wc = 0; wc < (h->warp_path ? 2 * n : 0); wc++) {
        // runs of landmarks with identical keyframe sequence (kept as Seg records: n_lm unbounded, no chunk fields)
        const int w = wc >> 1;
        const plba_problem &p = probs[w]; const WinInfo &wi = h->wins[w];
        {
            const int cls = wc & 1;
            const ClassLayout &L = cls ? Lls[w] : Lps[w];
            const int nl = cls ? p.n_ls : p.n_pt, lm0 = cls ? wi.ls0 : wi.pt0, ob0 = cls ? wi.lo0 : wi.po0;
            const int32_t *kf = cls ? p.lo_kf : p.po_kf;
            std::vector<int> &perm = cls ? h->ls_perm : h->pt_perm, &operm = cls ? h->lo_perm : h->po_perm, &ptr = cls ? ls_ptr : pt_ptr;
            std::vector<Seg> &sgs = loc_sg[wc]; std::vector<int> &fps = loc_fp[wc];      // fp0 is relative to this list until the concatenation
            // (1) observation offsets of the re-ordered landmarks (serial prefix), (2) the observation permutation (parallel for large
            // windows: 12.5 M entries at config 5), (3) the runs (serial, index compares only)
            // One large window (first branch): L.optr is read in signature order, i.e. at scattered places, so everything that touches it
            // runs in parallel — track lengths and "a new run starts here" flags (a landmark joins the run of its PREDECESSOR, which has the
            // run's signature) — and the serial passes walk sequential arrays only.
            if (nl > PAR_LM && g_host_nt > 1) {
                std::vector<int> lens(nl); std::vector<unsigned char> brk(nl);
#pragma omp parallel for num_threads(g_host_nt) schedule(static)
                for (int nl_i = 0; nl_i < nl; nl_i++) {
                    const int old = L.perm[nl_i], a = L.optr[old], no = L.optr[old + 1] - a;
                    lens[nl_i] = no;
                    bool join = false;
                    if (nl_i > 0) {
                        const int pold = L.perm[nl_i - 1], pa = L.optr[pold];
                        if (L.optr[pold + 1] - pa == no) {
                            if (!L.group.empty()) join = (L.group[pold] == L.group[old]);
                            else { join = true; for (int i = 0; i < no; i++) if (kf[pa + i] != kf[a + i]) { join = false; break; } }
                        }
                    }
                    brk[nl_i] = join ? 0 : 1;
                }
                int ob = ob0;
                for (int nl_i = 0; nl_i < nl; nl_i++) { ptr[lm0 + nl_i] = ob; ob += lens[nl_i]; }
#pragma omp parallel for num_threads(g_host_nt) schedule(static)
                for (int nl_i = 0; nl_i < nl; nl_i++) {
                    const int old = L.perm[nl_i], a = L.optr[old];
                    perm[lm0 + nl_i] = lm0 + old;
                    int o = ptr[lm0 + nl_i];
                    for (int i = 0; i < lens[nl_i]; i++) operm[o++] = ob0 + a + i;
                }
                for (int nl_i = 0; nl_i < nl; nl_i++) {
                    if (!brk[nl_i]) { sgs.back().n_lm++; continue; }
                    const int old = L.perm[nl_i], a = L.optr[old], no = lens[nl_i];
                    Seg s{}; s.lm0 = lm0 + nl_i; s.n_lm = 1; s.nobs = no; s.fp0 = (int)fps.size(); s.pad1 = w;
                    for (int i = 0; i < no; i++) if (p.kf_slot[kf[a + i]] >= 0) { fps.push_back(i); s.nfree++; }
                    sgs.push_back(s);
                }
            } else {
                int ob = ob0;
                for (int nl_i = 0; nl_i < nl; nl_i++) { const int old = L.perm[nl_i]; ptr[lm0 + nl_i] = ob; ob += L.optr[old + 1] - L.optr[old]; }
#pragma omp parallel for num_threads(g_host_nt) schedule(static) if (nl > PAR_LM)
                for (int nl_i = 0; nl_i < nl; nl_i++) {
                    const int old = L.perm[nl_i], a = L.optr[old], b = L.optr[old + 1];
                    perm[lm0 + nl_i] = lm0 + old;
                    int o = ptr[lm0 + nl_i];
                    for (int i = a; i < b; i++) operm[o++] = ob0 + i;
                }
                int run_first_old = -1;
                for (int nl_i = 0; nl_i < nl; nl_i++) {
                    const int old = L.perm[nl_i], a = L.optr[old], b = L.optr[old + 1], no = b - a;
                    const int g = lm0 + nl_i;
                    bool join = false;
                    if (run_first_old >= 0) {
                        Seg &s = sgs.back();
                        if (s.nobs == no) {
                            const int fa = L.optr[run_first_old];
                            if (!L.group.empty()) join = (L.group[run_first_old] == L.group[old]);
                            else { join = true; for (int i = 0; i < no; i++) if (kf[fa + i] != kf[a + i]) { join = false; break; } }
                        }
                        if (join) s.n_lm++;
                    }
                    if (!join) {      // (a run of landmarks WITHOUT observations has nobs = 0: only the update kernel visits it)
                        Seg s{}; s.lm0 = g; s.n_lm = 1; s.nobs = no; s.fp0 = (int)fps.size(); s.pad1 = w;
                        for (int i = 0; i < no; i++) if (p.kf_slot[kf[a + i]] >= 0) { fps.push_back(i); s.nfree++; }
                        sgs.push_back(s); run_first_old = old;
                    }
                }
            }
            // (the entry after a window's last landmark = ob0 + its observations is the next window's first, written there; the very last one below)
        }
    }
    if (n == 1 && h->warp_path) HOSTPROF("ch.runs");
    if (h->warp_path) {
        pt_ptr[tot.n_pt] = tot.n_pobs; ls_ptr[tot.n_ls] = tot.n_lobs;
        for (int w = 0; w < n; w++) for (int cls = 0; cls < 2; cls++) {
            std::vector<Seg> &sgs = cls ? sg_ls : sg_pt; std::vector<int> &fps = cls ? fp_ls : fp_pt;
            seg_begin[cls][w] = (int)sgs.size();
            const int base = (int)fps.size();
            for (Seg sg : loc_sg[2 * (size_t)w + cls]) { sg.fp0 += base; sgs.push_back(sg); }
            fps.insert(fps.end(), loc_fp[2 * (size_t)w + cls].begin(), loc_fp[2 * (size_t)w + cls].end());
        }
        seg_begin[0][n] = (int)sg_pt.size(); seg_begin[1][n] = (int)sg_ls.size();
        if (n == 1) HOSTPROF("ch.concat");
        // cut the runs into items of up to W_ITEM_PASSES_MAX passes (fewer when the upload is small)
        int64_t total_passes = 0;
        for (int cls = 0; cls < 2; cls++) for (const Seg &sg : (cls ? sg_ls : sg_pt)) { const int lpp = sg.nobs ? 32 / sg.nobs : 32; total_passes += (sg.n_lm + lpp - 1) / lpp; }
        const int64_t warps = (int64_t)h->grid_warp * WARPS_PER_CTA;
        const int item_passes = (int)std::max<int64_t>(1, std::min<int64_t>(W_ITEM_PASSES_MAX, total_passes / (8 * warps)));   // items are dealt dynamically: >= 8 per warp keeps the tail short
        for (int cls = 0; cls < 2; cls++) {
            std::vector<WItem> &items = cls ? wi_ls : wi_pt; const std::vector<int> &ptr = cls ? ls_ptr : pt_ptr;
            for (const Seg &sg : (cls ? sg_ls : sg_pt)) {
                const int lpp = sg.nobs ? 32 / sg.nobs : 32, step = cls ? std::min(lpp * item_passes, 32) : lpp * item_passes;   // lines: the update kernel retracts an item's landmarks one per lane
                for (int l = 0; l < sg.n_lm; l += step) {
                    WItem it{}; it.lm0 = sg.lm0 + l; it.n_lm = std::min(step, sg.n_lm - l); it.ob0 = ptr[it.lm0]; it.k = sg.nobs; it.win = sg.pad1; it.nfree = sg.nfree; it.fp0 = sg.fp0;
                    items.push_back(it);
                }
            }
        }
    }
    // Chunk and segment size of the CTA-chunk kernels.  256 observations per chunk and 8 landmarks per (segment, pose pair) task are the
    // throughput optimum (profiles/README.md r01b).  An upload that fits ONE wave of CTAs is latency-bound instead: its observations are
    // spread over the SMs (smaller chunks when there are fewer than 256 observations per SM) and the segments are as short as keeps a chunk within one round of 256 Schur tasks
    // (the task loop runs seg_max iterations).  Measured on C2: assembly 31.4 us at (256, 8), 27.3 us at (256, 6), 43.9 us at (256, 4: two rounds).
    const int64_t n_obs_all = (int64_t)tot.n_pobs + tot.n_lobs;
    const int64_t est_chunks = ((int64_t)tot.n_pobs + OC - 1) / OC + ((int64_t)tot.n_lobs + OC - 1) / OC + 2 * n;
    int oc_cap = OC, seg_max = SEG_MAX;
    if (est_chunks <= h->grid_chunks && n_obs_all > 0) {
        // about one chunk per SM (two co-resident CTAs slow each other down: C2 at 298 chunks of 153 observations measured 35 us against
        // 28.6 us at 180 chunks of 256; C1 at 148 chunks of 64 observations 17.0 us against 20.5 us at 37 chunks of 256)
        oc_cap = (int)std::min<int64_t>(OC, std::max<int64_t>(64, (n_obs_all + h->n_sm - 1) / h->n_sm));
        // mean track length -> landmarks per chunk and Schur tasks per segment (pairs x 2 halves + diagonal)
        const double k_mean = (double)n_obs_all / std::max<int64_t>(1, (int64_t)tot.n_pt + tot.n_ls);
        const double lm_per_chunk = oc_cap / std::max(1.0, k_mean), tasks_per_seg = k_mean * (k_mean - 1.0) + k_mean;
        seg_max = 8;
        for (int cand = 3; cand <= 8; cand++) if (std::ceil(lm_per_chunk / cand) * tasks_per_seg <= 0.85 * OC) { seg_max = cand; break; }
    }
    const int lc_cap = std::max(1, oc_cap / 2);
    for (int w = 0; w < n && !h->warp_path; w++) {
        const plba_problem &p = probs[w]; const WinInfo &wi = h->wins[w];
        const ClassLayout &Lp = Lps[w], &Ll = Lls[w];
        for (int cls = 0; cls < 2; cls++) {
            const ClassLayout &L = cls ? Ll : Lp;
            const int nl = cls ? p.n_ls : p.n_pt, lm0 = cls ? wi.ls0 : wi.pt0, ob0 = cls ? wi.lo0 : wi.po0;
            const int32_t *kf = cls ? p.lo_kf : p.po_kf;
            std::vector<int> &perm = cls ? h->ls_perm : h->pt_perm, &operm = cls ? h->lo_perm : h->po_perm, &ptr = cls ? ls_ptr : pt_ptr;
            std::vector<Chunk> &chs = cls ? ch_ls : ch_pt; std::vector<Seg> &sgs = cls ? sg_ls : sg_pt; std::vector<int> &fps = cls ? fp_ls : fp_pt;
            seg_begin[cls][w] = (int)sgs.size();
            int chunk_tasks = 0, chunk_dtasks = 0;
            int ob = ob0;
            Chunk c{}; bool open = false; int seg_first_old = -1;
            auto close_chunk = [&](int lm_end, int ob_end) { if (open) { c.lm1 = lm_end; c.ob1 = ob_end; c.seg1 = (int)sgs.size(); chs.push_back(c); open = false; seg_first_old = -1; } };
            for (int nl_i = 0; nl_i < nl; nl_i++) {
                const int old = L.perm[nl_i], a = L.optr[old], b = L.optr[old + 1], no = b - a;
                if (no > OC) { too_long = true; break; }
                const int g = lm0 + nl_i;
                perm[g] = lm0 + old;
                if (open && ((g - c.lm0) >= lc_cap || (ob - c.ob0) + no > std::max(oc_cap, no))) close_chunk(g, ob);
                if (!open) { c = Chunk{}; c.lm0 = g; c.ob0 = ob; c.win = w; c.seg0 = (int)sgs.size(); open = true; seg_first_old = -1; chunk_tasks = 0; chunk_dtasks = 0; }
                bool join = false;
                if (seg_first_old >= 0) {
                    Seg &s = sgs.back();
                    const int fa = L.optr[seg_first_old];
                    if (s.n_lm < seg_max && s.nobs == no) {
                        if (!L.group.empty()) join = (L.group[seg_first_old] == L.group[old]);
                        else { join = true; for (int i = 0; i < no; i++) if (kf[fa + i] != kf[a + i]) { join = false; break; } }
                    }
                    if (join) s.n_lm++;
                }
                if (!join) {
                    Seg s{}; s.lm0 = g; s.n_lm = 1; s.nobs = no; s.fp0 = (int)fps.size(); s.task0 = chunk_tasks; s.dtask0 = chunk_dtasks;
                    for (int i = 0; i < no; i++) if (p.kf_slot[kf[a + i]] >= 0) { fps.push_back(i); s.nfree++; }
                    chunk_tasks += s.nfree * (s.nfree - 1) / 2; chunk_dtasks += s.nfree;
                    sgs.push_back(s); seg_first_old = old;
                }
                ptr[g] = ob;
                for (int i = a; i < b; i++) operm[ob++] = ob0 + i;
            }
            if (too_long) break;
            close_chunk(lm0 + nl, ob);
            ptr[lm0 + nl] = ob;
        }
        if (too_long) break;
    }
    if (too_long) { h->err = "a landmark has more than 256 observations"; return PLBA_E_UNSUPPORTED; }
    seg_begin[0][n] = (int)sg_pt.size(); seg_begin[1][n] = (int)sg_ls.size();

    HOSTPROF("chunks");
    {   // layout statistics (plba_layout_stats): structural non-zero 6x6 blocks of S = union over segments of their pose pairs
        std::vector<size_t> wbase(n + 1, 0);
        for (int w = 0; w < n; w++) wbase[w + 1] = wbase[w] + (size_t)h->wins[w].n_free * h->wins[w].n_free;
        std::vector<unsigned char> &mark = h->mark; mark.assign(wbase[n], 0);
        int64_t n_off = 0, n_diag = 0, nnzb = 0; int band = 0;
        // one (or few) large windows: the segments in parallel — marks are idempotent byte stores, the structural blocks are counted afterwards
        const bool seg_par = n <= 8 && g_host_nt > 1 && sg_pt.size() + sg_ls.size() > 4096;
        for (int w = 0; w < n && seg_par; w++) for (int cls = 0; cls < 2; cls++) {
            const std::vector<Seg> &sgs = cls ? sg_ls : sg_pt; const std::vector<int> &fps = cls ? fp_ls : fp_pt;
            const std::vector<int> &ptr = cls ? ls_ptr : pt_ptr; const std::vector<int> &operm = cls ? h->lo_perm : h->po_perm;
            const plba_problem &p = probs[w]; const WinInfo &wi = h->wins[w];
            const int32_t *kf = cls ? p.lo_kf : p.po_kf; const int ob0 = cls ? wi.lo0 : wi.po0;
            unsigned char *mk = mark.data() + wbase[w];
#pragma omp parallel for num_threads(g_host_nt) schedule(static) reduction(+ : n_off, n_diag) reduction(max : band)
            for (int si = seg_begin[cls][w]; si < seg_begin[cls][w + 1]; si++) {
                const Seg &sg = sgs[si];
                n_off += (int64_t)sg.nfree * (sg.nfree - 1) / 2; n_diag += sg.nfree;
                const int o0 = ptr[sg.lm0];
                int sl[OC];      // (a track holds at most OC observations: checked when the runs were built)
                for (int i = 0; i < sg.nfree; i++) sl[i] = p.kf_slot[kf[operm[o0 + fps[sg.fp0 + i]] - ob0]];
                for (int i = 0; i < sg.nfree; i++) for (int j = i; j < sg.nfree; j++) {
                    int a = sl[i], b = sl[j];
                    if (a > b) std::swap(a, b);
                    if (b - a > band) band = b - a;
                    __atomic_store_n(&mk[(size_t)a * wi.n_free + b], (unsigned char)1, __ATOMIC_RELAXED);
                }
            }
        }
        if (seg_par) {
            const int64_t nm = (int64_t)wbase[n];
#pragma omp parallel for num_threads(g_host_nt) schedule(static) reduction(+ : nnzb)
            for (int64_t i = 0; i < nm; i++) nnzb += mark[i];
        }
        // windows in parallel for a batch: the marks of a window are its own stretch of `mark`
#pragma omp parallel for num_threads(g_host_nt) schedule(dynamic, 4) reduction(+ : n_off, n_diag, nnzb) reduction(max : band) if (n > 8)
        for (int w = 0; w < (seg_par ? 0 : n); w++) for (int cls = 0; cls < 2; cls++) {
            const std::vector<Seg> &sgs = cls ? sg_ls : sg_pt; const std::vector<int> &fps = cls ? fp_ls : fp_pt;
            const std::vector<int> &ptr = cls ? ls_ptr : pt_ptr; const std::vector<int> &operm = cls ? h->lo_perm : h->po_perm;
            const plba_problem &p = probs[w]; const WinInfo &wi = h->wins[w];
            const int32_t *kf = cls ? p.lo_kf : p.po_kf; const int ob0 = cls ? wi.lo0 : wi.po0;
            std::vector<int> sl, sl_prev;
            for (int si = seg_begin[cls][w]; si < seg_begin[cls][w + 1]; si++) {
                const Seg &sg = sgs[si];
                n_off += (int64_t)sg.nfree * (sg.nfree - 1) / 2; n_diag += sg.nfree;
                const int o0 = ptr[sg.lm0];
                sl.resize(sg.nfree);
                for (int i = 0; i < sg.nfree; i++) sl[i] = p.kf_slot[kf[operm[o0 + fps[sg.fp0 + i]] - ob0]];
                if (sl == sl_prev) continue;                       // same free-keyframe set as the previous segment (a run cut into segments): nothing new to mark
                for (int i = 0; i < sg.nfree; i++) for (int j = i; j < sg.nfree; j++) {
                    int a = sl[i], b = sl[j];
                    if (a > b) std::swap(a, b);
                    if (b - a > band) band = b - a;
                    unsigned char &m = mark[wbase[w] + (size_t)a * wi.n_free + b];
                    if (!m) { m = 1; nnzb++; }
                }
                sl_prev = sl;
            }
        }
        h->layout[0] = (int64_t)ch_pt.size(); h->layout[1] = (int64_t)ch_ls.size(); h->layout[2] = (int64_t)sg_pt.size(); h->layout[3] = (int64_t)sg_ls.size();
        h->layout[4] = n_off; h->layout[5] = n_diag; h->layout[6] = nnzb; h->band_blocks = band;
        if (has_exchange(h)) {
            // landmark-sharded path: every rank must pick the same solver and the same node layout, i.e. agree on the half bandwidth
            // of the SUMMED reduced camera system = the largest over the shards (max all-reduce through the caller's hook)
            double v = (double)band;
            CK(cudaMemcpyAsync(h->d_scratch, &v, sizeof(double), cudaMemcpyHostToDevice, h->stream));
            allreduce(h, h->d_scratch, 1, 1);
            CK(cudaMemcpyAsync(&v, h->d_scratch, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
            CK(cudaStreamSynchronize(h->stream));
            h->band_blocks = (int)v;
        }
    }

    HOSTPROF("stats");
    // ---- memory plan ----------------------------------------------------------------------------------------
    Carver ci;   // staged inputs (same offsets in pinned host memory and at the start of the device arena)
    const size_t i_kf_slot = ci.take<int>(tot.n_kf), i_kf_win = ci.take<int>(tot.n_kf), i_slot_kf = ci.take<int>(tot.n_free);
    const size_t i_win_slot0 = ci.take<int>(n), i_win_nfree = ci.take<int>(n), i_win_ls0 = ci.take<int>(n), i_win_S = ci.take<long long>(n);
    const size_t i_win_bs = ci.take<int>(n), i_win_N = ci.take<int>(n);
    const size_t i_Tmap = ci.take<double>((size_t)12 * tot.n_kf), i_X0 = ci.take<double>((size_t)6 * tot.n_free);
    const size_t i_pts0 = ci.take<double>((size_t)3 * tot.n_pt), i_lns0 = ci.take<double>((size_t)ld * tot.n_ls), i_lmap = ci.take<double>((size_t)6 * tot.n_ls);
    const size_t i_pt_ptr = ci.take<int>(tot.n_pt + 1), i_ls_ptr = ci.take<int>(tot.n_ls + 1), i_pt_win = ci.take<int>(tot.n_pt), i_ls_win = ci.take<int>(tot.n_ls);
    // observations travel in the CALLER's order (straight copies on the host); the first kernel after the upload gathers them into the
    // internal order (k_reset, gather = 1): per landmark the first staged index of its run, per observation (u, v) / (a, b), the global
    // keyframe index and sigma^2 (1 where the caller gave none) — 28 / 44 bytes per observation + 4 per landmark
    const size_t i_pt_src = ci.take<int>(tot.n_pt), i_ls_src = ci.take<int>(tot.n_ls);
    const size_t i_r_po_kf = ci.take<int>(tot.n_pobs), i_r_lo_kf = ci.take<int>(tot.n_lobs);
    const size_t i_r_po_uv = ci.take<double>((size_t)2 * tot.n_pobs), i_r_lo_ab = ci.take<double>((size_t)4 * tot.n_lobs);
    const size_t i_r_po_s2 = ci.take<double>(tot.n_pobs), i_r_lo_s2 = ci.take<double>(tot.n_lobs);
    const size_t i_ch_pt = ci.take<Chunk>(ch_pt.size()), i_ch_ls = ci.take<Chunk>(ch_ls.size()), i_sg_pt = ci.take<Seg>(sg_pt.size()), i_sg_ls = ci.take<Seg>(sg_ls.size());
    const size_t i_fp_pt = ci.take<int>(fp_pt.size()), i_fp_ls = ci.take<int>(fp_ls.size());
    const size_t i_wi_pt = ci.take<WItem>(wi_pt.size()), i_wi_ls = ci.take<WItem>(wi_ls.size());
    const size_t i_ctrl0 = ci.take<WinCtrl>(n);
    h->in_bytes = ci.off; h->i_pts0 = i_pts0; h->i_lns0 = i_lns0; h->i_lmap = i_lmap;
    // reduced camera system: dense (6 nf)^2 per window, or — large block-banded windows on the block-cyclic-reduction solver — the
    // solver's node form [D | U], N nodes of m = 6 bs unknowns: 2 N m^2 doubles (config 5: 35 MB instead of 1.15 GB)
    h->bcr_layout = !h->small_path && h->band_blocks <= BAND_MAX && !h->force_dense && h->large_solver == 0;
    const int bcr_bs = std::min((int)BCR_BS_MAX, std::max(h->band_blocks, 6)), bcr_m = 6 * bcr_bs;
    long long S_off = 0;
    std::vector<long long> &win_S_off = h->win_S_off; win_S_off.assign(n, 0);
    for (int w = 0; w < n; w++) {
        win_S_off[w] = S_off;
        const long long nf_w = h->wins[w].n_free, N_w = (nf_w + bcr_bs - 1) / bcr_bs;
        S_off += h->bcr_layout ? 2 * N_w * bcr_m * bcr_m : 36 * nf_w * nf_w;
    }
    h->S_doubles = (size_t)S_off;
    h->sys_doubles = h->S_doubles + (size_t)18 * tot.n_free + (size_t)ACC_N * n + (size_t)n + (size_t)PLBA_MAX_RANKS * n;
    Carver cs = ci;   // device-only state follows the inputs
    const size_t i_po_kf = cs.take<int>(tot.n_pobs), i_po_lm = cs.take<int>(tot.n_pobs), i_lo_kf = cs.take<int>(tot.n_lobs), i_lo_lm = cs.take<int>(tot.n_lobs);
    const size_t i_po_uv = cs.take<double>((size_t)2 * tot.n_pobs), i_lo_ab = cs.take<double>((size_t)4 * tot.n_lobs);
    const size_t i_po_om = cs.take<double>(tot.n_pobs), i_lo_om = cs.take<double>(tot.n_lobs);
    size_t s_poseT[2], s_X[2], s_pts[2], s_lns[2], s_lpre[2];
    for (int b = 0; b < 2; b++) s_lpre[b] = cs.take<double>((size_t)LPRE_N * tot.n_ls);
    for (int b = 0; b < 2; b++) { s_poseT[b] = cs.take<double>((size_t)12 * tot.n_kf); s_X[b] = cs.take<double>((size_t)6 * tot.n_free); s_pts[b] = cs.take<double>((size_t)3 * tot.n_pt); s_lns[b] = cs.take<double>((size_t)ld * tot.n_ls); }
    const size_t s_po_lvl = cs.take<unsigned char>(tot.n_pobs), s_lo_lvl = cs.take<unsigned char>(tot.n_lobs);
    const size_t s_po_chi2 = cs.take<double>(tot.n_pobs), s_lo_chi2 = cs.take<double>(tot.n_lobs);      // cached chi2 of every edge (internal order; the output region holds the caller-order copy of k_export)
    const size_t s_sys = cs.take<double>(h->sys_doubles), s_xp = cs.take<double>((size_t)6 * tot.n_free);
    // node storage of the block cyclic reduction (large banded windows only)
    h->bcr.clear();
    std::vector<size_t> s_bcr;
    if (h->bcr_layout) {
        const int bs = bcr_bs, m = bcr_m;
        for (int w = 0; w < n; w++) {
            BcrW B{}; B.bs = bs; B.m = m; B.N = (h->wins[w].n_free + bs - 1) / bs;
            h->bcr.push_back(B);
            const size_t nn = (size_t)B.N * m * m;
            s_bcr.push_back(cs.take<double>(nn)); s_bcr.push_back(cs.take<double>(nn)); s_bcr.push_back(cs.take<double>((size_t)B.N * m));   // Xl, Xr, y (D, U, b, hd live in the system buffer)
            s_bcr.push_back(cs.take<double>(nn));                                                                                                // Lf
        }
    }
    const int trace_cap = (prof == PLBA_PROFILE_G) ? (opt->iters_stage1 + opt->iters_stage2) * opt->lm_max_trials + 2 : opt->max_iters_lba + 2;
    // output region (one D2H copy)
    h->out_off = cs.off;
    Carver co; co.off = cs.off;
    h->o_T = co.take<double>((size_t)12 * tot.n_kf); h->o_x = co.take<double>((size_t)6 * tot.n_free); h->o_pt = co.take<double>((size_t)3 * tot.n_pt);
    h->o_ls = co.take<double>((size_t)ld * tot.n_ls); h->o_plk = co.take<double>((size_t)6 * tot.n_ls);
    h->o_pchi = co.take<double>(tot.n_pobs); h->o_lchi = co.take<double>(tot.n_lobs);
    h->o_pf = co.take<unsigned char>(tot.n_pobs); h->o_lf = co.take<unsigned char>(tot.n_lobs);
    h->o_ctrl = co.take<WinCtrl>(n); h->o_trace = co.take<plba_trace_rec>((size_t)n * trace_cap); h->o_cnt = co.take<int>(CNT_N);
    h->out_bytes = co.off - h->out_off;
    int rc = ensure(h, co.off, h->in_bytes, h->out_bytes);
    if (rc) return rc;
    h->layout[7] = (int64_t)co.off;

    HOSTPROF("memplan");
    // ---- flatten into the pinned staging buffer ------------------------------------------------------------------
    char *hb = h->h_in;
    int *kf_slot = (int *)(hb + i_kf_slot), *kf_win = (int *)(hb + i_kf_win), *slot_kf = (int *)(hb + i_slot_kf);
    int *win_slot0 = (int *)(hb + i_win_slot0), *win_nfree = (int *)(hb + i_win_nfree), *win_ls0 = (int *)(hb + i_win_ls0);
    long long *winS = (long long *)(hb + i_win_S);
    int *win_bs = (int *)(hb + i_win_bs), *win_N = (int *)(hb + i_win_N);
    double *Tmap = (double *)(hb + i_Tmap), *X0 = (double *)(hb + i_X0), *pts0 = (double *)(hb + i_pts0), *lns0 = (double *)(hb + i_lns0), *lmap = (double *)(hb + i_lmap);
    int *d_pt_ptr = (int *)(hb + i_pt_ptr), *d_ls_ptr = (int *)(hb + i_ls_ptr), *pt_win = (int *)(hb + i_pt_win), *ls_win = (int *)(hb + i_ls_win);
    int *pt_src = (int *)(hb + i_pt_src), *ls_src = (int *)(hb + i_ls_src), *r_po_kf = (int *)(hb + i_r_po_kf), *r_lo_kf = (int *)(hb + i_r_lo_kf);
    double *r_po_uv = (double *)(hb + i_r_po_uv), *r_lo_ab = (double *)(hb + i_r_lo_ab), *r_po_s2 = (double *)(hb + i_r_po_s2), *r_lo_s2 = (double *)(hb + i_r_lo_s2);
    WinCtrl *ctrl0 = (WinCtrl *)(hb + i_ctrl0);
    std::memcpy(d_pt_ptr, pt_ptr.data(), sizeof(int) * pt_ptr.size());
    std::memcpy(d_ls_ptr, ls_ptr.data(), sizeof(int) * ls_ptr.size());
    if (!ch_pt.empty()) std::memcpy(hb + i_ch_pt, ch_pt.data(), sizeof(Chunk) * ch_pt.size());
    if (!ch_ls.empty()) std::memcpy(hb + i_ch_ls, ch_ls.data(), sizeof(Chunk) * ch_ls.size());
    if (!sg_pt.empty()) std::memcpy(hb + i_sg_pt, sg_pt.data(), sizeof(Seg) * sg_pt.size());
    if (!sg_ls.empty()) std::memcpy(hb + i_sg_ls, sg_ls.data(), sizeof(Seg) * sg_ls.size());
    if (!fp_pt.empty()) std::memcpy(hb + i_fp_pt, fp_pt.data(), sizeof(int) * fp_pt.size());
    if (!fp_ls.empty()) std::memcpy(hb + i_fp_ls, fp_ls.data(), sizeof(int) * fp_ls.size());
    if (!wi_pt.empty()) std::memcpy(hb + i_wi_pt, wi_pt.data(), sizeof(WItem) * wi_pt.size());
    if (!wi_ls.empty()) std::memcpy(hb + i_wi_ls, wi_ls.data(), sizeof(WItem) * wi_ls.size());
    HOSTPROF("fl.tables");
    const bool par_lm = (n <= 8);       // one (or few) big windows: parallel over landmarks; batches: parallel over windows
#pragma omp parallel for num_threads(g_host_nt) schedule(dynamic, 4) if (!par_lm)
    for (int w = 0; w < n; w++) {
        const plba_problem &p = probs[w]; const WinInfo &wi = h->wins[w];
        win_slot0[w] = wi.slot0; win_nfree[w] = wi.n_free; win_ls0[w] = wi.ls0; winS[w] = win_S_off[w];
        win_bs[w] = h->bcr_layout ? bcr_bs : 0; win_N[w] = h->bcr_layout ? (wi.n_free + bcr_bs - 1) / bcr_bs : 0;
        for (int k = 0; k < p.n_kf; k++) {
            const int s = p.kf_slot[k];
            kf_slot[wi.kf0 + k] = s < 0 ? -1 : wi.slot0 + s; kf_win[wi.kf0 + k] = w;
            inv_se3(p.kf_T_wc + 12 * (size_t)k, &Tmap[(size_t)(wi.kf0 + k) * 12]);     // T_cw (Q16: reference uses a general 4x4 inverse)
            if (s >= 0) {
                slot_kf[wi.slot0 + s] = wi.kf0 + k;
                if (p.x_pose) for (int i = 0; i < 6; i++) X0[(size_t)(wi.slot0 + s) * 6 + i] = p.x_pose[(size_t)6 * s + i];
                else log_se3(p.kf_T_wc + 12 * (size_t)k, &X0[(size_t)(wi.slot0 + s) * 6]);
            }
        }
        if (n == 1) HOSTPROF("fl.kf");
#pragma omp parallel for num_threads(g_host_nt) schedule(static) if (par_lm && p.n_pt > PAR_LM)
        for (int g = wi.pt0; g < wi.pt0 + p.n_pt; g++) {
            const int old = h->pt_perm[g] - wi.pt0;
            pt_win[g] = w;
            pt_src[g] = pt_ptr[g + 1] > pt_ptr[g] ? h->po_perm[pt_ptr[g]] : 0;      // first observation of the landmark's run in the caller's (staged) order
            for (int i = 0; i < 3; i++) pts0[(size_t)3 * g + i] = p.pt_xyz[(size_t)3 * old + i];
        }
#pragma omp parallel for num_threads(g_host_nt) schedule(static) if (par_lm && p.n_ls > PAR_LM)
        for (int g = wi.ls0; g < wi.ls0 + p.n_ls; g++) {
            const int old = h->ls_perm[g] - wi.ls0;
            ls_win[g] = w;
            ls_src[g] = ls_ptr[g + 1] > ls_ptr[g] ? h->lo_perm[ls_ptr[g]] : 0;
            if (prof == PLBA_PROFILE_H_END) for (int i = 0; i < 6; i++) lns0[(size_t)6 * g + i] = p.ls_end[(size_t)6 * old + i];
            else {
                // changePlukerToOrth (:6040, :1577) runs on the device (k_reset): four inverse trig calls per line are not host work
                for (int i = 0; i < 6; i++) lmap[(size_t)6 * g + i] = p.ls_plk[(size_t)6 * old + i];
            }
        }
        if (n == 1) HOSTPROF("fl.lm");
        // observations: straight copies in the caller's order, in slices so that a large window's copy runs on every host thread
        {
            const int SL = 65536, nsp = (p.n_pobs + SL - 1) / SL, nsl = (p.n_lobs + SL - 1) / SL;
#pragma omp parallel for num_threads(g_host_nt) schedule(static) if (par_lm && p.n_pobs + p.n_lobs > PAR_OBS)
            for (int sl = 0; sl < nsp + nsl; sl++) {
                if (sl < nsp) {
                    const size_t a = (size_t)sl * SL, c = std::min<size_t>(SL, (size_t)p.n_pobs - a), d = (size_t)wi.po0 + a;
                    std::memcpy(r_po_uv + 2 * d, p.po_uv + 2 * a, sizeof(double) * 2 * c);
                    if (wi.kf0 == 0) std::memcpy(r_po_kf + d, p.po_kf + a, sizeof(int) * c);
                    else for (size_t j = 0; j < c; j++) r_po_kf[d + j] = wi.kf0 + p.po_kf[a + j];
                    if (p.po_sig2) std::memcpy(r_po_s2 + d, p.po_sig2 + a, sizeof(double) * c);
                    else for (size_t j = 0; j < c; j++) r_po_s2[d + j] = 1.0;
                } else {
                    const size_t a = (size_t)(sl - nsp) * SL, c = std::min<size_t>(SL, (size_t)p.n_lobs - a), d = (size_t)wi.lo0 + a;
                    std::memcpy(r_lo_ab + 4 * d, p.lo_ab + 4 * a, sizeof(double) * 4 * c);
                    if (wi.kf0 == 0) std::memcpy(r_lo_kf + d, p.lo_kf + a, sizeof(int) * c);
                    else for (size_t j = 0; j < c; j++) r_lo_kf[d + j] = wi.kf0 + p.lo_kf[a + j];
                    if (p.lo_sig2) std::memcpy(r_lo_s2 + d, p.lo_sig2 + a, sizeof(double) * c);
                    else for (size_t j = 0; j < c; j++) r_lo_s2[d + j] = 1.0;
                }
            }
        }
        if (n == 1) HOSTPROF("fl.obs");
        WinCtrl c{};
        c.need_init = 1; c.apply = 1; c.ni = 2.0; c.err_prev = 999999999.9; c.n_lm_pt = p.n_pt; c.n_lm_ls = p.n_ls; c.n_obs = p.n_pobs + p.n_lobs;
        // nothing to do (src/mapHandler.cpp:1496-1500) — except on the landmark-sharded path: a rank whose shard is empty contributes zeros to
        // every exchange and takes the same LM decisions as the others (skipping it would leave them blocked in the first all-reduce)
        if (p.n_pobs + p.n_lobs == 0 && !has_exchange(h)) c.done = 1;
        ctrl0[w] = c;
    }

    HOSTPROF("flatten");
    // ---- kernel parameters ------------------------------------------------------------------------------------
    DevP &P = h->P;
    P = DevP{};
    char *db = h->d_arena;
    P.cam = Cam{probs[0].cam[0], probs[0].cam[1], probs[0].cam[2], probs[0].cam[3]};
    P.profile = prof; P.fixed_quirks = (opt->quirks == PLBA_QUIRKS_FIXED);
    P.n_win = n; P.n_kf = tot.n_kf; P.n_free = tot.n_free; P.n_pt = tot.n_pt; P.n_ls = tot.n_ls; P.n_pobs = tot.n_pobs; P.n_lobs = tot.n_lobs;
    P.n_chunks_pt = (int)ch_pt.size(); P.n_chunks_ls = (int)ch_ls.size();
    P.iters_stage1 = opt->iters_stage1; P.iters_stage2 = opt->iters_stage2; P.lm_max_trials = opt->lm_max_trials; P.max_iters_lba = opt->max_iters_lba;
    P.max_rounds = (prof == PLBA_PROFILE_G) ? (opt->iters_stage1 + opt->iters_stage2) * opt->lm_max_trials + 2 : opt->max_iters_lba + 1;
    P.huber_delta = opt->huber_delta; P.chi2_gate = opt->chi2_gate; P.homog_th = opt->homog_th; P.min_error = opt->min_error;
    P.min_error_change = opt->min_error_change; P.lm_tau = opt->lm_tau;
    P.gba = (prof == PLBA_PROFILE_H_END && opt->shell == PLBA_SHELL_GBA) ? 1 : 0;
    if (P.gba) { P.min_error = P.min_error_change = 2.220446049250313e-16; }      // numeric_limits<double>::epsilon() (src/mapHandler.cpp:3664, :3694)
    P.lambda_lba_lm = opt->lambda_lba_lm; P.lambda_lba_k = opt->lambda_lba_k;
    P.kf_slot = (int *)(db + i_kf_slot); P.kf_win = (int *)(db + i_kf_win); P.slot_kf = (int *)(db + i_slot_kf);
    P.win_slot0 = (int *)(db + i_win_slot0); P.win_nfree = (int *)(db + i_win_nfree); P.win_ls0 = (int *)(db + i_win_ls0); P.win_S_off = (long long *)(db + i_win_S);
    P.win_bcr_bs = (int *)(db + i_win_bs); P.win_bcr_N = (int *)(db + i_win_N);
    P.kf_Tmap = (double *)(db + i_Tmap); P.X0 = (double *)(db + i_X0); P.pts0 = (double *)(db + i_pts0); P.lns0 = (double *)(db + i_lns0); P.lns_map = (double *)(db + i_lmap);
    P.pt_ptr = (int *)(db + i_pt_ptr); P.ls_ptr = (int *)(db + i_ls_ptr); P.pt_win = (int *)(db + i_pt_win); P.ls_win = (int *)(db + i_ls_win);
    P.po_kf = (int *)(db + i_po_kf); P.po_lm = (int *)(db + i_po_lm); P.lo_kf = (int *)(db + i_lo_kf); P.lo_lm = (int *)(db + i_lo_lm);
    P.po_uv = (double *)(db + i_po_uv); P.lo_ab = (double *)(db + i_lo_ab); P.po_om = (double *)(db + i_po_om); P.lo_om = (double *)(db + i_lo_om);
    P.pt_src = (int *)(db + i_pt_src); P.ls_src = (int *)(db + i_ls_src); P.r_po_kf = (int *)(db + i_r_po_kf); P.r_lo_kf = (int *)(db + i_r_lo_kf);
    P.r_po_uv = (double *)(db + i_r_po_uv); P.r_lo_ab = (double *)(db + i_r_lo_ab); P.r_po_s2 = (double *)(db + i_r_po_s2); P.r_lo_s2 = (double *)(db + i_r_lo_s2);
    P.chunks_pt = (Chunk *)(db + i_ch_pt); P.chunks_ls = (Chunk *)(db + i_ch_ls); P.segs_pt = (Seg *)(db + i_sg_pt); P.segs_ls = (Seg *)(db + i_sg_ls);
    P.freepos_pt = (int *)(db + i_fp_pt); P.freepos_ls = (int *)(db + i_fp_ls);
    P.witems_pt = (WItem *)(db + i_wi_pt); P.witems_ls = (WItem *)(db + i_wi_ls); P.n_witems_pt = (int)wi_pt.size(); P.n_witems_ls = (int)wi_ls.size();
    P.ctrl0 = (WinCtrl *)(db + i_ctrl0);
    for (int b = 0; b < 2; b++) { P.poseT[b] = (double *)(db + s_poseT[b]); P.Xkf[b] = (double *)(db + s_X[b]); P.pts[b] = (double *)(db + s_pts[b]); P.lns[b] = (double *)(db + s_lns[b]); P.lpre[b] = (double *)(db + s_lpre[b]); }
    P.po_lvl = (unsigned char *)(db + s_po_lvl); P.lo_lvl = (unsigned char *)(db + s_lo_lvl);
    P.po_chi2 = (double *)(db + s_po_chi2); P.lo_chi2 = (double *)(db + s_lo_chi2);
    h->sysbuf = (double *)(db + s_sys);
    P.S = h->sysbuf; P.gs = P.S + h->S_doubles; P.hpp_diag = P.gs + (size_t)6 * tot.n_free; P.hpp_diag_init = P.hpp_diag + (size_t)6 * tot.n_free;
    P.acc = P.hpp_diag_init + (size_t)6 * tot.n_free; P.accB = P.acc + (size_t)4 * n; P.accmax = P.acc + (size_t)ACC_N * n; P.maxslots = P.accmax + n;
    P.n_ranks = has_exchange(h) ? h->n_ranks : 1; P.rank = h->rank;
    P.xp = (double *)(db + s_xp);
    for (size_t w = 0; w < h->bcr.size(); w++) {
        BcrW &B = h->bcr[w];
        const size_t nn = (size_t)B.N * B.m * B.m;
        B.D = P.S + win_S_off[w]; B.U = B.D + nn;
        B.b = P.gs + (size_t)6 * h->wins[w].slot0; B.hd = P.hpp_diag + (size_t)6 * h->wins[w].slot0;      // node i, unknown c <-> slot i bs + c / 6: the same index as in g
        B.Xl = (double *)(db + s_bcr[4 * w]); B.Xr = (double *)(db + s_bcr[4 * w + 1]); B.y = (double *)(db + s_bcr[4 * w + 2]);
        B.Lf = h->bcr_v1 ? nullptr : (double *)(db + s_bcr[4 * w + 3]);
    }
    P.solve_nf_max = std::min(max_nf, (int)SMALL_NMAX / 6);
    P.S_clear_doubles = h->small_path ? (long long)((h->S_doubles + 1) & ~(size_t)1) : 0;      // (the buffer continues with g: an odd tail would only clear g[0], which the solver has consumed too; S_doubles is even anyway)
    P.ctrl = (WinCtrl *)(db + h->o_ctrl); P.trace = (plba_trace_rec *)(db + h->o_trace); P.trace_cap = trace_cap; P.counters = (int *)(db + h->o_cnt);
    set_all_attrs();
#ifndef PLBA_HOST_EMU
    if (h->small_path && !h->no_graph) {
        if ((rc = build_graph_for(h, prof))) return rc;
        const int pi = (prof * 2 + h->solve_class) * 2 + (h->warp_path ? 1 : 0);
        P.cond_while = (unsigned long long)h->cond_while[pi]; P.cond_prep = (unsigned long long)h->cond_prep[pi]; P.cond_prep2 = (unsigned long long)h->cond_prep2[pi];
    }
#endif
    HOSTPROF("params");
    *h->h_P = P;
    h->timing.ms_host_prep = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
    CK(cudaMemcpyAsync(h->d_P, h->h_P, sizeof(DevP), cudaMemcpyHostToDevice, h->stream));
    CK(cudaMemcpyAsync(h->d_arena, h->h_in, h->in_bytes, cudaMemcpyHostToDevice, h->stream));
    h->timing.h2d_bytes += (int64_t)(h->in_bytes + sizeof(DevP));
    cudaEventRecord(h->ev_h2d, h->stream); h->h2d_pending = true;
    h->uploaded = true;
    return launch_reset(h, 1);
}

int plba_reset_state(plba_handle h) {
    if (!h || !h->uploaded) return PLBA_E_ARG;
    CK(cudaSetDevice(h->device));
    launch_reset(h);
    CK(cudaGetLastError());
    return PLBA_OK;
}

}  // extern "C"

// one LM round, host driven: (gate + lambda init) -> assemble -> [exchange] -> solve -> update -> [exchange] -> control.
// need_prep: the host KNOWS that a window waits for its chi2 gate / initial lambda (synchronous loop) or cannot know (pipelined
// loop: the prep kernels are launched every round and leave at once when nothing waits for them).
// Exchange steps of the landmark-sharded path (has_exchange): per LM trial ONE all-reduce of the reduced camera system with the cost
// sums riding along, plus 4 doubles per window after the update kernel; per prep block one all-reduce of [diag(H_pp) | slots of the
// landmark-diagonal maxima].  All of them on the handle's stream: no host synchronisation.
static int run_round(plba_handle h, bool need_prep) {
    DevP &P = h->P; cudaStream_t st = h->stream; const DevP *Pp = h->d_P;
    // (the large-window solvers factor in place: S, g and diag(H_pp) are cleared before every assembly; in node form that is 17 MB at config 5)
    const bool xch = has_exchange(h);
    if (!h->small_path) CK(cudaMemsetAsync(h->sysbuf, 0, sizeof(double) * (h->S_doubles + (size_t)12 * P.n_free), st));
    if (need_prep) {
        PLBA_LAUNCH(k_gate, dim3(h->grid_chunks), dim3(256), 0, st, Pp); h->timing.n_launches++;
        launch_assemble(h, 0);
        if (xch) {
            // [diag(H_pp) of the prep pass | acc | accB (both zero here) | accmax | one slot per rank]: the maximum travels as a sum of slots
            PLBA_LAUNCH(k_pack_max, grid1(P.n_ranks * P.n_win, 128), dim3(128), 0, st, Pp); h->timing.n_launches++;
            allreduce(h, P.hpp_diag_init, (size_t)6 * P.n_free + (size_t)ACC_N * P.n_win + (size_t)P.n_win + (size_t)P.n_ranks * P.n_win, 0);
        }
        PLBA_LAUNCH(k_lambda_init, dim3(h->n_sm), dim3(128), 0, st, Pp); h->timing.n_launches++;
    }
    if (h->detail_timing) cudaEventRecord(h->ev[0], st);
    launch_assemble(h, 1);
    if (h->detail_timing) cudaEventRecord(h->ev[1], st);
    if (xch) {
        // the exchange step: [S | g | diag(H_pp) | (zero) prep diagonal | cost sums] is ONE contiguous range, whatever the storage of S
        // (dense, or the band in node form for the block cyclic reduction: 17 MB at config 5)
        allreduce(h, h->sysbuf, h->S_doubles + (size_t)18 * P.n_free + (size_t)4 * P.n_win, 0);
    }
    launch_solve(h);
    if (h->detail_timing) cudaEventRecord(h->ev[2], st);
    launch_update(h, 0);
    if (h->detail_timing) cudaEventRecord(h->ev[3], st);
    if (xch) allreduce(h, P.accB, (size_t)4 * P.n_win, 0);                     // update-phase sums (new cost, scale, |dx|^2)
    PLBA_LAUNCH(k_control, dim3(1), dim3(256), 0, st, Pp, 0); h->timing.n_launches++;
    return PLBA_OK;
}

enum { PLBA_ROUNDS_IN_FLIGHT = 2 };
static int run_async(plba_handle h) {
    DevP &P = h->P; cudaStream_t st = h->stream;
    h->timing.ms_assemble = h->timing.ms_solve = h->timing.ms_update = 0;
#ifndef PLBA_HOST_EMU
    if (use_graph(h)) {
        CK(cudaGraphLaunch(h->gexec[(P.profile * 2 + h->solve_class) * 2 + (h->warp_path ? 1 : 0)], st));
        return PLBA_OK;
    }
#endif
    int rc;
#ifndef PLBA_HOST_EMU
    if (!h->detail_timing) {
        // Pipelined host-driven loop (large windows, sharded path): the host never waits for the round it has just launched.  The
        // controller runs on the device; the host only needs to learn WHEN every window is done, and reads the counters of the round
        // it launched PLBA_ROUNDS_IN_FLIGHT rounds ago (pinned copy + event), so the stream never drains.  At most that many rounds
        // are launched in vain; their kernels return at once for finished windows.  No per-round cudaStreamSynchronize.
        // (The kernels of the dense tiled Cholesky do not look at the windows' done flags — a check would sit on the critical path of 125
        //  panel steps — and one of its solves costs tens of milliseconds: windows on that solver are polled every round.)
        const bool dense_solver = !h->small_path && !h->bcr_layout && !(h->band_blocks <= BAND_MAX && !h->force_dense);
        const int in_flight = dense_solver ? 0 : (int)PLBA_ROUNDS_IN_FLIGHT;
        bool done = false;
        for (int round = 0; round < P.max_rounds && !done; round++) {
            if ((rc = run_round(h, true))) return rc;
            const int slot = round % 4;
            CK(cudaMemcpyAsync(h->h_poll + slot * CNT_N, P.counters, sizeof(int) * CNT_N, cudaMemcpyDeviceToHost, st));
            CK(cudaEventRecord(h->ev_poll[slot], st));
            if (round >= in_flight) {
                const int old = (round - in_flight) % 4;
                CK(cudaEventSynchronize(h->ev_poll[old]));
                done = h->h_poll[old * CNT_N + CNT_DONE] >= P.n_win;
            }
        }
        if (h->nccl_failed) return PLBA_E_CUDA;
        return PLBA_OK;
    }
#endif
    // synchronous loop (per-stage event timing, host emulation): one round, one poll
    if ((rc = poll_counters(h))) return rc;
    for (int round = 0; round < P.max_rounds; round++) {
        if (h->h_counters[CNT_DONE] >= P.n_win) break;
        const bool need_prep = h->h_counters[CNT_NEED_INIT] > 0 || h->h_counters[CNT_GATE] > 0;
        if ((rc = run_round(h, need_prep))) return rc;
        if ((rc = poll_counters(h))) return rc;
        if (h->detail_timing) {
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, h->ev[0], h->ev[1]); cudaEventElapsedTime(&b, h->ev[1], h->ev[2]); cudaEventElapsedTime(&c, h->ev[2], h->ev[3]);
            h->timing.ms_assemble += a; h->timing.ms_solve += b; h->timing.ms_update += c;
        }
    }
    return PLBA_OK;
}

// the chi2 gate of windows that left stage 0 in the very last round is still pending (src/mapHandler.cpp:6125-6147)
static int finish_pending_gate(plba_handle h) {
    if (h->P.profile != PLBA_PROFILE_G) return PLBA_OK;
    PLBA_LAUNCH(k_gate, dim3(h->grid_chunks), dim3(256), 0, h->stream, (const DevP *)h->d_P); h->timing.n_launches++;
    PLBA_LAUNCH(k_lambda_init, dim3(h->n_sm), dim3(128), 0, h->stream, (const DevP *)h->d_P); h->timing.n_launches++;
    return PLBA_OK;
}

static void account_graph_launches(plba_handle h, int64_t launches0) {
    // kernels that ran inside the graph: COUNTED on the device (every kernel of the library bumps CNT_KLAUNCH when it starts; the
    // counter restarts at every reset).  The count includes the host-launched kernels of the same span, so it replaces the host's tally.
    h->timing.n_launches = launches0 + (int64_t)(h->h_counters[CNT_KLAUNCH] - h->klaunch_seen);
    h->timing.n_assemble += h->h_counters[CNT_ROUNDS];
}

extern "C" {

int plba_run(plba_handle h) {
    if (!h || !h->uploaded) return PLBA_E_ARG;
    CK(cudaSetDevice(h->device));
    cudaStream_t st = h->stream;
    const int64_t launches0 = h->timing.n_launches, assemble0 = h->timing.n_assemble;
    const bool graph = use_graph(h);
    cudaEventRecord(h->ev[6], st);
    int rc;
    if ((rc = run_async(h))) return rc;
    if ((rc = finish_pending_gate(h))) return rc;
    cudaEventRecord(h->ev[7], st);
    if ((rc = poll_counters(h))) return rc;
    if (graph) account_graph_launches(h, launches0);
    h->klaunch_seen = h->h_counters[CNT_KLAUNCH];
    float ms = 0; cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]);
    h->timing.ms_total = ms;
    h->timing.ms_other = ms - h->timing.ms_assemble - h->timing.ms_solve - h->timing.ms_update;
    h->timing.n_launches_run = h->timing.n_launches - launches0;
    h->timing.n_assemble_run = h->timing.n_assemble - assemble0;
    h->timing.n_trials_run = h->h_counters[CNT_TRIALS];
    CK(cudaGetLastError());
    return PLBA_OK;
}

int plba_download(plba_handle h, int32_t n, plba_result *res) {
    if (!h || !h->uploaded || n != h->P.n_win || !res) return PLBA_E_ARG;
    CK(cudaSetDevice(h->device));
    DevP &P = h->P; cudaStream_t st = h->stream;
    g_host_nt = host_threads_for((int64_t)P.n_pobs + P.n_lobs);
    const bool G = (P.profile == PLBA_PROFILE_G);
    const int ld = h->ls_dim;
    char *db = h->d_arena;
    ExportP E{};
    E.T_wc = (double *)(db + h->o_T); E.x = (double *)(db + h->o_x); E.pt = (double *)(db + h->o_pt); E.ls = (double *)(db + h->o_ls); E.plk = (double *)(db + h->o_plk);
    E.pf = (unsigned char *)(db + h->o_pf); E.lf = (unsigned char *)(db + h->o_lf); E.pchi = (double *)(db + h->o_pchi); E.lchi = (double *)(db + h->o_lchi);
    E.ls_dim = ld; E.q9 = (P.profile == PLBA_PROFILE_H_PLK && !P.fixed_quirks) ? 1 : 0; E.do_final = G ? 1 : 0;
    PLBA_LAUNCH(k_export, dim3(h->grid_chunks), dim3(256), 0, st, (const DevP *)h->d_P, E); h->timing.n_launches++;
    CK(cudaMemcpyAsync(h->h_out, db + h->out_off, h->out_bytes, cudaMemcpyDeviceToHost, st));
    h->timing.d2h_bytes += (int64_t)h->out_bytes;
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    const auto t_host0 = std::chrono::steady_clock::now();
    auto t_hp = t_host0; const bool hostprof = std::getenv("PLBA_HOST_PROF") != nullptr;
#define HOSTPROF(name) do { if (hostprof) { const auto t_ = std::chrono::steady_clock::now(); std::fprintf(stderr, "[upload] %-10s %.3f ms\n", name, std::chrono::duration<double, std::milli>(t_ - t_hp).count()); t_hp = t_; } } while (0)
    const char *ob = h->h_out - h->out_off;      // so that ob + o_xxx addresses the host copy
    const double *T = (const double *)(ob + h->o_T), *X = (const double *)(ob + h->o_x), *pt = (const double *)(ob + h->o_pt), *ls = (const double *)(ob + h->o_ls);
    const double *plk = (const double *)(ob + h->o_plk), *pchi = (const double *)(ob + h->o_pchi), *lchi = (const double *)(ob + h->o_lchi);
    const unsigned char *pf = (const unsigned char *)(ob + h->o_pf), *lf = (const unsigned char *)(ob + h->o_lf);
    const WinCtrl *ctrl = (const WinCtrl *)(ob + h->o_ctrl); const plba_trace_rec *trace = (const plba_trace_rec *)(ob + h->o_trace);
    const int *cnt = (const int *)(ob + h->o_cnt);
    for (int i = 0; i < CNT_N; i++) h->h_counters[i] = cnt[i];
    const double *pt0 = (const double *)(h->h_in + h->i_pts0), *ls0 = (const double *)(h->h_in + h->i_lns0);
    int rc_all = PLBA_OK;
#pragma omp parallel for num_threads(g_host_nt) schedule(dynamic, 4) if (n > 8)
    for (int w = 0; w < n; w++) {
        const WinInfo &wi = h->wins[w]; plba_result &r = res[w];
        r.n_trace = ctrl[w].n_trace; r.n_trials = ctrl[w].n_trials;
        r.status = (wi.n_pobs + wi.n_lobs == 0 && !has_exchange(h)) ? PLBA_DISCARDED : (ctrl[w].numeric & (1 << 30)) ? PLBA_E_NUMERIC : PLBA_OK;
        if (r.trace) for (int i = 0; i < std::min(r.n_trace, std::min(r.trace_cap, P.trace_cap)); i++) r.trace[i] = trace[(size_t)w * P.trace_cap + i];
        if (r.kf_T_wc) for (int i = 0; i < 12 * wi.n_kf; i++) r.kf_T_wc[i] = T[(size_t)12 * wi.kf0 + i];
        if (r.x_pose) for (int i = 0; i < 6 * wi.n_free; i++) r.x_pose[i] = X[(size_t)6 * wi.slot0 + i];
        const bool par_lm = (n <= 8);       // one (or few) big windows: parallel over landmarks / observations; batches: parallel over windows
#pragma omp parallel for num_threads(g_host_nt) schedule(static) if (par_lm && wi.n_pt > PAR_LM)
        for (int g = wi.pt0; g < wi.pt0 + wi.n_pt; g++) {
            const int old = h->pt_perm[g] - wi.pt0;
            if (r.pt_xyz) for (int i = 0; i < 3; i++) r.pt_xyz[(size_t)3 * old + i] = pt[(size_t)3 * g + i];
            // inlier rule of the hand-LM write-back (src/mapHandler.cpp:2858-2860, 2871-2873); profile G leaves it to the final chi2 test
            if (r.pt_inlier) { double d2 = 0; for (int i = 0; i < 3; i++) { const double d = pt[(size_t)3 * g + i] - pt0[(size_t)3 * g + i]; d2 += d * d; } r.pt_inlier[old] = (!G && !P.gba && std::sqrt(d2) > 0.01) ? 0 : 1; }
        }
#pragma omp parallel for num_threads(g_host_nt) schedule(static) if (par_lm && wi.n_ls > PAR_LM)
        for (int g = wi.ls0; g < wi.ls0 + wi.n_ls; g++) {
            const int old = h->ls_perm[g] - wi.ls0;
            if (ld == 4) {
                if (r.ls_orth) for (int i = 0; i < 4; i++) r.ls_orth[(size_t)4 * old + i] = ls[(size_t)4 * g + i];
                if (r.ls_plk) for (int i = 0; i < 6; i++) r.ls_plk[(size_t)6 * old + i] = plk[(size_t)6 * g + i];
            } else if (r.ls_end) for (int i = 0; i < 6; i++) r.ls_end[(size_t)6 * old + i] = ls[(size_t)6 * g + i];
            if (r.ls_inlier) {
                double d2 = 0, o0[6];
                if (G) o0[0] = 0;
                else if (ld == 4) plk_to_orth((const double *)(h->h_in + h->i_lmap) + (size_t)6 * g, o0);     // initial orthonormal coordinates (computed on the device for the solve)
                else for (int i = 0; i < 6; i++) o0[i] = ls0[(size_t)6 * g + i];
                if (!G) for (int i = 0; i < ld; i++) { const double d = ls[(size_t)ld * g + i] - o0[i]; d2 += d * d; }
                r.ls_inlier[old] = (!G && !P.gba && std::sqrt(d2) > 0.01) ? 0 : 1;
            }
        }
        if (G) {
            // k_export wrote chi2 and flags in the caller's order: straight copies (in slices for a large window)
            const int SL = 262144, nsp = (wi.n_pobs + SL - 1) / SL, nsl = (wi.n_lobs + SL - 1) / SL;
#pragma omp parallel for num_threads(g_host_nt) schedule(static) if (par_lm && wi.n_pobs + wi.n_lobs > PAR_OBS)
            for (int sl = 0; sl < nsp + nsl; sl++) {
                if (sl < nsp) {
                    const size_t a = (size_t)sl * SL, c = std::min<size_t>(SL, (size_t)wi.n_pobs - a), d = (size_t)wi.po0 + a;
                    if (r.po_chi2) std::memcpy(r.po_chi2 + a, pchi + d, sizeof(double) * c);
                    if (r.po_flags) std::memcpy(r.po_flags + a, pf + d, c);
                } else {
                    const size_t a = (size_t)(sl - nsp) * SL, c = std::min<size_t>(SL, (size_t)wi.n_lobs - a), d = (size_t)wi.lo0 + a;
                    if (r.lo_chi2) std::memcpy(r.lo_chi2 + a, lchi + d, sizeof(double) * c);
                    if (r.lo_flags) std::memcpy(r.lo_flags + a, lf + d, c);
                }
            }
        }
    }
    for (int w = 0; w < n; w++) if (res[w].status < PLBA_DISCARDED) { rc_all = res[w].status; if (rc_all == PLBA_E_NUMERIC) h->err = "the reduced camera system was not positive definite in every trial of an LM iteration"; }
    h->timing.ms_host_unpack = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t_host0).count();
    return rc_all;
}

int plba_solve_batch(plba_handle h, int32_t n, const plba_problem *probs, const plba_options *opt, plba_result *res) {
    if (!h || !res) return PLBA_E_ARG;
    int rc = plba_upload(h, n, probs, opt);
    if (rc) { for (int w = 0; w < n; w++) { res[w].status = rc; res[w].n_trace = 0; res[w].n_trials = 0; } return rc; }
    const bool graph = use_graph(h);
    const int64_t launches0 = h->timing.n_launches, assemble0 = h->timing.n_assemble;
    cudaEventRecord(h->ev[6], h->stream);
    if ((rc = run_async(h))) return rc;
    if ((rc = finish_pending_gate(h))) return rc;
    cudaEventRecord(h->ev[7], h->stream);
    rc = plba_download(h, n, res);              // one synchronisation for the whole call
    if (graph) account_graph_launches(h, launches0);
    h->klaunch_seen = h->h_counters[CNT_KLAUNCH];
    float ms = 0; cudaEventElapsedTime(&ms, h->ev[6], h->ev[7]);
    h->timing.ms_total = ms; h->timing.ms_other = ms - h->timing.ms_assemble - h->timing.ms_solve - h->timing.ms_update;
    h->timing.n_launches_run = h->timing.n_launches - launches0;
    h->timing.n_assemble_run = h->timing.n_assemble - assemble0;
    h->timing.n_trials_run = h->h_counters[CNT_TRIALS];
    // fixed KFs: copy the caller's rows through untouched
    for (int w = 0; w < n; w++) if (res[w].kf_T_wc) for (int k = 0; k < probs[w].n_kf; k++) if (probs[w].kf_slot[k] < 0) for (int i = 0; i < 12; i++) res[w].kf_T_wc[(size_t)12 * k + i] = probs[w].kf_T_wc[(size_t)12 * k + i];
    return rc;
}

int plba_solve(plba_handle h, const plba_problem *prob, const plba_options *opt, plba_result *res) {
    if (!prob || !res) return PLBA_E_ARG;
    if (prob->n_pobs + prob->n_lobs == 0 && !(h && has_exchange(h))) { res->status = PLBA_DISCARDED; res->n_trace = 0; res->n_trials = 0; return PLBA_DISCARDED; }   // src/mapHandler.cpp:1496-1500
    int rc = plba_solve_batch(h, 1, prob, opt, res);
    return rc ? rc : res->status;
}

// ---- single-trial interface (bench roofline timing, sharded drivers) ----
int plba_trial_assemble(plba_handle h, double lambda) {
    if (!h || !h->uploaded) return PLBA_E_ARG;
    DevP &P = h->P; cudaStream_t st = h->stream;
    PLBA_LAUNCH(k_set_lambda, grid1(P.n_win, 128), dim3(128), 0, st, (const DevP *)h->d_P, lambda);
    CK(cudaMemsetAsync(h->sysbuf, 0, sizeof(double) * h->sys_doubles, st));
    launch_assemble(h, 1);
    CK(cudaGetLastError());
    return PLBA_OK;
}
int plba_trial_finish(plba_handle h, double lambda, double *chi_new, double *scale) {
    if (!h || !h->uploaded) return PLBA_E_ARG;
    DevP &P = h->P; cudaStream_t st = h->stream;
    (void)lambda;
    launch_solve(h);
    launch_update(h, 0);
    double acc[4];
    CK(cudaMemcpyAsync(acc, P.accB, sizeof(acc), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (chi_new) *chi_new = acc[ACC_CHI_NEW];
    if (scale) *scale = acc[ACC_SCALE];
    CK(cudaGetLastError());
    return PLBA_OK;
}
int plba_reduced_system(plba_handle h, void **dev_ptr, int64_t *n_doubles) {
    if (!h || !h->uploaded) return PLBA_E_ARG;
    if (dev_ptr) *dev_ptr = h->sysbuf;
    if (n_doubles) {
        *n_doubles = (int64_t)(h->S_doubles + (size_t)6 * h->P.n_free);
    }
    return PLBA_OK;
}
int plba_copy_reduced_system(plba_handle h, int32_t window, double *S_out, double *g_out) {
    if (!h || !h->uploaded || window < 0 || window >= h->P.n_win) return PLBA_E_ARG;
    const WinInfo &wi = h->wins[window];
    const size_t n = (size_t)6 * wi.n_free;
    if (g_out && n) CK(cudaMemcpyAsync(g_out, h->P.gs + (size_t)6 * wi.slot0, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    if (S_out && n && !h->bcr_layout) {
        size_t off = 0;
        for (int w = 0; w < window; w++) off += (size_t)36 * h->wins[w].n_free * h->wins[w].n_free;
        CK(cudaMemcpyAsync(S_out, h->P.S + off, sizeof(double) * n * n, cudaMemcpyDeviceToHost, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    if (S_out && n && h->bcr_layout) {
        // the window is stored in the node form of the block cyclic reduction (s_block()): unfold it into the dense upper storage
        const BcrW &B = h->bcr[window];
        const size_t nn = (size_t)B.N * B.m * B.m;
        std::vector<double> nodes(2 * nn);
        CK(cudaMemcpyAsync(nodes.data(), B.D, sizeof(double) * 2 * nn, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
        std::memset(S_out, 0, sizeof(double) * n * n);
        const int m = B.m;
        for (int i = 0; i < B.N; i++) {
            const int oi = i * m, mi = bcr_node_size(B, wi.n_free, i);
            const double *D = nodes.data() + (size_t)i * m * m, *U = nodes.data() + nn + (size_t)i * m * m;
            for (int r = 0; r < mi; r++) for (int c = r; c < mi; c++) S_out[(size_t)(oi + r) * n + oi + c] = D[(size_t)r * m + c];      // the node's upper triangle
            if (i > 0) for (int r = 0; r < m; r++) for (int c = 0; c < mi; c++) S_out[(size_t)(oi - m + r) * n + oi + c] = U[(size_t)r * m + c];
        }
    }
    return PLBA_OK;
}
// Average device time of one launch of a stage kernel on the resident problem: `reps` back-to-back launches bracketed by
// CUDA events on the handle's stream.  which: 0 assemble (mode 1), 1 reduced-system solve, 2 update.  State is not advanced.
int plba_time_kernel(plba_handle h, int32_t which, int32_t reps, double lambda, double *ms_avg) {
    if (!h || !h->uploaded || reps <= 0 || !ms_avg) return PLBA_E_ARG;
    cudaStream_t st = h->stream;
    PLBA_LAUNCH(k_set_lambda, grid1(h->P.n_win, 128), dim3(128), 0, st, (const DevP *)h->d_P, lambda);
    CK(cudaMemsetAsync(h->sysbuf, 0, sizeof(double) * h->sys_doubles, st));
    for (int w = 0; w < 2; w++) { if (which == 0) launch_assemble(h, 1); else if (which == 1) launch_solve(h); else launch_update(h, 0); }
    cudaEventRecord(h->ev[4], st);
    for (int r = 0; r < reps; r++) { if (which == 0) launch_assemble(h, 1); else if (which == 1) launch_solve(h); else launch_update(h, 0); }
    cudaEventRecord(h->ev[5], st);
    CK(cudaStreamSynchronize(st));
    float ms = 0; cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]);
    *ms_avg = (double)ms / reps;
    CK(cudaMemsetAsync(h->sysbuf, 0, sizeof(double) * h->sys_doubles, st));
    CK(cudaGetLastError());
    return PLBA_OK;
}
#ifdef PLBA_PROF
int plba_debug_prof(unsigned long long *out64, int reset) {
    if (out64) cudaMemcpyFromSymbol(out64, plba_prof_table, sizeof(unsigned long long) * 64);
    if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(plba_prof_table, z, sizeof(z)); }
    return 0;
}
#endif
#ifndef PLBA_HOST_EMU
}  // extern "C"
// ---- measured FP64 roof (bench.py: the roofline denominators of an FP64-bound path are measured, not nominal) ----
// DFMA: 8 independent chains per thread, 8 warps per CTA, 8 CTAs per SM.  DMMA: mma.sync.m8n8k4.f64, 4 independent accumulator pairs.
__global__ void __launch_bounds__(256) k_peak_dfma(double *out, int iters, double a, double b) {
    double x[8];
#pragma unroll
    for (int i = 0; i < 8; i++) x[i] = a + i * 1e-3 + threadIdx.x * 1e-6;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 8; i++) x[i] = fma(x[i], b, a);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += x[i];
    if (s == 1.2345e300) out[0] = s;
}
__global__ void __launch_bounds__(256) k_peak_dmma(double *out, int iters, double a, double b) {
    double c[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++) { c[i][0] = a * i; c[i][1] = b * i; }
    const double fa = a + threadIdx.x * 1e-6, fb = b;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 4; u++) {
#pragma unroll
            for (int i = 0; i < 4; i++) plba_dmma(c[i][0], c[i][1], fa, fb);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) s += c[i][0] + c[i][1];
    if (s == 1.2345e300) out[0] = s;
}
extern "C" {
// which: 0 = DFMA (vector pipe), 1 = DMMA (mma.sync.m8n8k4.f64).  Best of `reps` timed launches, CUDA events on the handle's stream.
int plba_measure_fp64_peak(plba_handle h, int32_t which, int32_t reps, double *tflops) {
    if (!h || !tflops || reps <= 0) return PLBA_E_ARG;
    CK(cudaSetDevice(h->device));
    const int iters = 20000, grid = 8 * h->n_sm;
    double best = 0.0;
    for (int r = 0; r < reps + 1; r++) {
        cudaEventRecord(h->ev[4], h->stream);
        if (which == 0) k_peak_dfma<<<grid, 256, 0, h->stream>>>(h->d_scratch, iters, 1.0000001, 0.9999999);
        else k_peak_dmma<<<grid, 256, 0, h->stream>>>(h->d_scratch, iters, 1.0000001, 0.9999999);
        cudaEventRecord(h->ev[5], h->stream);
        CK(cudaStreamSynchronize(h->stream));
        float ms = 0; cudaEventElapsedTime(&ms, h->ev[4], h->ev[5]);
        // DFMA: 32 FMAs per thread and iteration; DMMA: 16 MMAs per warp and iteration, 8 x 8 x 4 FMAs each
        const double flops = which == 0 ? 2.0 * 32.0 * iters * 256.0 * grid : 2.0 * 16.0 * 256.0 * iters * 8.0 * grid;
        if (r > 0 && ms > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    *tflops = best;
    CK(cudaGetLastError());
    return PLBA_OK;
}
#else
int plba_measure_fp64_peak(plba_handle, int32_t, int32_t, double *) { return PLBA_E_UNSUPPORTED; }
#endif
int plba_set_force_dense(plba_handle h, int on) { if (!h) return PLBA_E_ARG; h->force_dense = on != 0; return PLBA_OK; }
int plba_set_force_chunk(plba_handle h, int mode) { if (!h || mode < 0 || mode > 2) return PLBA_E_ARG; h->force_chunk = mode; return PLBA_OK; }
// ---- frame-to-frame pose tracking ---------------------------------------------------------------
void plba_track_default_options(plba_track_options *o) {
    std::memset(o, 0, sizeof(*o));
    o->homog_th = 1e-7; o->min_error = 1e-7; o->min_error_change = 1e-7;          // src2/config.cpp:80-85
    o->max_iters = 5;                                                             // Config::maxIters (src2/config.cpp:82)
}
int plba_track_solve(plba_handle h, int32_t n_frames, const plba_track_frame *frames, const plba_track_options *opt, plba_track_result *results) {
    if (!h || n_frames <= 0 || !frames || !opt || !results) return PLBA_E_ARG;
    CK(cudaSetDevice(h->device));
    size_t npt = 0, nls = 0;
    for (int f = 0; f < n_frames; f++) {
        const plba_track_frame &F = frames[f];
        if (F.n_pt < 0 || F.n_ls < 0 || F.n_pt > TRK_MAX || F.n_ls > TRK_MAX) { h->err = "a frame holds more than 1024 point or line matches"; return PLBA_E_UNSUPPORTED; }
        if ((F.n_pt && (!F.pt_P || !F.pt_obs)) || (F.n_ls && (!F.ls_sP || !F.ls_eP || !F.ls_NDc || !F.ls_obs || !F.ls_seg))) { h->err = "frame arrays missing"; return PLBA_E_ARG; }
        npt += F.n_pt; nls += F.n_ls;
    }
    Carver c;
    const size_t o_fr = c.take<TrackFrameDev>(n_frames), o_pP = c.take<double>(3 * npt), o_pO = c.take<double>(2 * npt), o_pI = c.take<unsigned char>(npt);
    const size_t o_lS = c.take<double>(3 * nls), o_lE = c.take<double>(3 * nls), o_lN = c.take<double>(6 * nls), o_lO = c.take<double>(4 * nls), o_lG = c.take<double>(4 * nls);
    const size_t o_l2 = c.take<double>(nls), o_lI = c.take<unsigned char>(nls);
    const size_t in_bytes = c.off, o_res = c.take<plba_track_result>(n_frames), total = c.off;
    if (total > h->trk_dev_cap) {
        cudaStreamSynchronize(h->stream);
        if (h->trk_dev) cudaFree(h->trk_dev);
        if (h->trk_host) cudaFreeHost(h->trk_host);
        h->trk_dev = h->trk_host = nullptr; h->trk_dev_cap = h->trk_host_cap = 0;
        const size_t want = total + total / 4 + (1 << 16);
        void *q = nullptr;
        if (cudaMalloc(&q, want) != cudaSuccess) { h->err = "cudaMalloc (tracking arena) failed"; return PLBA_E_CUDA; }
        h->trk_dev = (char *)q;
        if (cudaMallocHost(&q, want) != cudaSuccess) { h->err = "cudaMallocHost (tracking staging) failed"; return PLBA_E_CUDA; }
        h->trk_host = (char *)q; h->trk_dev_cap = h->trk_host_cap = want;
    }
    char *hb = h->trk_host, *db = h->trk_dev;
    TrackFrameDev *fr = (TrackFrameDev *)(hb + o_fr);
    size_t p0 = 0, l0 = 0;
    for (int f = 0; f < n_frames; f++) {
        const plba_track_frame &F = frames[f];
        fr[f].n_pt = F.n_pt; fr[f].n_ls = F.n_ls; fr[f].pt0 = (int)p0; fr[f].ls0 = (int)l0;
        for (int i = 0; i < 12; i++) fr[f].DT[i] = F.DT[i];
        if (F.n_pt) {
            std::memcpy(hb + o_pP + sizeof(double) * 3 * p0, F.pt_P, sizeof(double) * 3 * F.n_pt);
            std::memcpy(hb + o_pO + sizeof(double) * 2 * p0, F.pt_obs, sizeof(double) * 2 * F.n_pt);
            for (int i = 0; i < F.n_pt; i++) ((unsigned char *)(hb + o_pI))[p0 + i] = F.pt_inlier ? (F.pt_inlier[i] ? 1 : 0) : 1;
        }
        if (F.n_ls) {
            std::memcpy(hb + o_lS + sizeof(double) * 3 * l0, F.ls_sP, sizeof(double) * 3 * F.n_ls);
            std::memcpy(hb + o_lE + sizeof(double) * 3 * l0, F.ls_eP, sizeof(double) * 3 * F.n_ls);
            std::memcpy(hb + o_lN + sizeof(double) * 6 * l0, F.ls_NDc, sizeof(double) * 6 * F.n_ls);
            std::memcpy(hb + o_lO + sizeof(double) * 4 * l0, F.ls_obs, sizeof(double) * 4 * F.n_ls);
            std::memcpy(hb + o_lG + sizeof(double) * 4 * l0, F.ls_seg, sizeof(double) * 4 * F.n_ls);
            for (int i = 0; i < F.n_ls; i++) {
                ((double *)(hb + o_l2))[l0 + i] = F.ls_sigma2 ? F.ls_sigma2[i] : 1.0;
                ((unsigned char *)(hb + o_lI))[l0 + i] = F.ls_inlier ? (F.ls_inlier[i] ? 1 : 0) : 1;
            }
        }
        p0 += F.n_pt; l0 += F.n_ls;
    }
    CK(cudaMemcpyAsync(db, hb, in_bytes, cudaMemcpyHostToDevice, h->stream));
    TrackP T{};
    T.frames = (const TrackFrameDev *)(db + o_fr);
    T.pt_P = (const double *)(db + o_pP); T.pt_obs = (const double *)(db + o_pO); T.pt_in = (const unsigned char *)(db + o_pI);
    T.ls_sP = (const double *)(db + o_lS); T.ls_eP = (const double *)(db + o_lE); T.ls_NDc = (const double *)(db + o_lN);
    T.ls_obs = (const double *)(db + o_lO); T.ls_seg = (const double *)(db + o_lG); T.ls_s2 = (const double *)(db + o_l2); T.ls_in = (const unsigned char *)(db + o_lI);
    T.res = (plba_track_result *)(db + o_res);
    for (int i = 0; i < 4; i++) T.cam[i] = opt->cam[i];
    T.homog_th = opt->homog_th; T.min_error = opt->min_error; T.min_error_change = opt->min_error_change; T.max_iters = opt->max_iters; T.n_frames = n_frames;
    PLBA_LAUNCH(k_track_gn, dim3(std::min(n_frames, 8 * h->n_sm)), dim3(TRK_NT), track_smem(), h->stream, T);
    h->timing.n_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hb + o_res, db + o_res, sizeof(plba_track_result) * n_frames, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    std::memcpy(results, hb + o_res, sizeof(plba_track_result) * n_frames);
    return PLBA_OK;
}

int plba_create_lines(plba_handle h, const plba_newline_batch *B, double *NDc, double *NDw, double *err_first, double *err_curr, uint8_t *accept) {
    if (!h || !B || B->n < 0 || B->n_kf <= 0) return PLBA_E_ARG;
    if (B->n == 0) return PLBA_OK;
    if (!B->seg_l || !B->seg_r || !B->seg_curr || !B->kf_prev || !B->kf_curr || !B->kf_T_wc) { h->err = "batch arrays missing"; return PLBA_E_ARG; }
    for (int i = 0; i < B->n; i++) if (B->kf_prev[i] < 0 || B->kf_prev[i] >= B->n_kf || B->kf_curr[i] < 0 || B->kf_curr[i] >= B->n_kf) { h->err = "keyframe row out of range"; return PLBA_E_ARG; }
    CK(cudaSetDevice(h->device));
    const size_t n = (size_t)B->n;
    Carver c;
    const size_t o_sl = c.take<double>(4 * n), o_sr = c.take<double>(4 * n), o_sc = c.take<double>(4 * n), o_kp = c.take<int>(n), o_kc = c.take<int>(n), o_T = c.take<double>((size_t)12 * B->n_kf);
    const size_t in_bytes = c.off;
    const size_t o_c = c.take<double>(6 * n), o_w = c.take<double>(6 * n), o_e1 = c.take<double>(n), o_e2 = c.take<double>(n), o_ac = c.take<unsigned char>(n), total = c.off;
    if (total > h->trk_dev_cap) {
        cudaStreamSynchronize(h->stream);
        if (h->trk_dev) cudaFree(h->trk_dev);
        if (h->trk_host) cudaFreeHost(h->trk_host);
        h->trk_dev = h->trk_host = nullptr; h->trk_dev_cap = h->trk_host_cap = 0;
        const size_t want = total + total / 4 + (1 << 16);
        void *q = nullptr;
        if (cudaMalloc(&q, want) != cudaSuccess) { h->err = "cudaMalloc (tracking arena) failed"; return PLBA_E_CUDA; }
        h->trk_dev = (char *)q;
        if (cudaMallocHost(&q, want) != cudaSuccess) { h->err = "cudaMallocHost (tracking staging) failed"; return PLBA_E_CUDA; }
        h->trk_host = (char *)q; h->trk_dev_cap = h->trk_host_cap = want;
    }
    char *hb = h->trk_host, *db = h->trk_dev;
    std::memcpy(hb + o_sl, B->seg_l, sizeof(double) * 4 * n); std::memcpy(hb + o_sr, B->seg_r, sizeof(double) * 4 * n); std::memcpy(hb + o_sc, B->seg_curr, sizeof(double) * 4 * n);
    std::memcpy(hb + o_kp, B->kf_prev, sizeof(int) * n); std::memcpy(hb + o_kc, B->kf_curr, sizeof(int) * n); std::memcpy(hb + o_T, B->kf_T_wc, sizeof(double) * 12 * B->n_kf);
    CK(cudaMemcpyAsync(db, hb, in_bytes, cudaMemcpyHostToDevice, h->stream));
    NewLineP Q{};
    Q.n = B->n; for (int i = 0; i < 5; i++) Q.cam[i] = B->cam[i];
    Q.seg_l = (const double *)(db + o_sl); Q.seg_r = (const double *)(db + o_sr); Q.seg_curr = (const double *)(db + o_sc);
    Q.kf_prev = (const int *)(db + o_kp); Q.kf_curr = (const int *)(db + o_kc); Q.kf_T = (const double *)(db + o_T);
    Q.NDc = (double *)(db + o_c); Q.NDw = (double *)(db + o_w); Q.err_first = (double *)(db + o_e1); Q.err_curr = (double *)(db + o_e2); Q.accept = (unsigned char *)(db + o_ac);
    PLBA_LAUNCH(k_create_lines, grid1(B->n, 256), dim3(256), 0, h->stream, Q);
    h->timing.n_launches++;
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hb + o_c, db + o_c, total - o_c, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (NDc) std::memcpy(NDc, hb + o_c, sizeof(double) * 6 * n);
    if (NDw) std::memcpy(NDw, hb + o_w, sizeof(double) * 6 * n);
    if (err_first) std::memcpy(err_first, hb + o_e1, sizeof(double) * n);
    if (err_curr) std::memcpy(err_curr, hb + o_e2, sizeof(double) * n);
    if (accept) std::memcpy(accept, hb + o_ac, n);
    return PLBA_OK;
}

int plba_kernel_path(plba_handle h, int32_t *out4) {
    if (!h || !h->uploaded || !out4) return PLBA_E_ARG;
    out4[0] = h->warp_path ? 1 : 0;
    out4[1] = h->small_path ? 0 : h->bcr_layout ? 1 : (h->band_blocks <= BAND_MAX && !h->force_dense) ? 2 : 3;
    out4[2] = h->band_blocks;
    out4[3] = h->warp_path ? (int)(h->wi_pt.size() + h->wi_ls.size()) : (int)(h->ch_pt.size() + h->ch_ls.size());
    return PLBA_OK;
}
int plba_layout_stats(plba_handle h, int64_t *out8) { if (!h || !h->uploaded || !out8) return PLBA_E_ARG; for (int i = 0; i < 8; i++) out8[i] = h->layout[i]; return PLBA_OK; }
int plba_get_timing(plba_handle h, plba_timing *t) { if (!h || !t) return PLBA_E_ARG; *t = h->timing; return PLBA_OK; }
int plba_set_detail_timing(plba_handle h, int on) { if (!h) return PLBA_E_ARG; h->detail_timing = on != 0; return PLBA_OK; }

}  // extern "C"
