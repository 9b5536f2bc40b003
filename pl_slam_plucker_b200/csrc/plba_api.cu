// Host side of the C ABI (include/plba.h): flattening of one or many LBA windows into the resident SoA layout,
// the round loop that drives the kernels (no per-observation work on the host), and the write-back.
// There is NO CPU fallback in this file: every numeric step is a kernel launch on the handle's stream.
#include <vector>
#include <string>
#include <algorithm>
#include <cstdio>
#include "plba_solver.h"

using namespace plba;

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e_); return PLBA_E_CUDA; } } while (0)

struct WinInfo { int n_kf, n_free, n_pt, n_ls, n_pobs, n_lobs, kf0, slot0, pt0, ls0, po0, lo0; };

struct plba_handle_s {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    std::string err;
    std::vector<void *> allocs;
    DevP P{};
    plba_options opt{};
    std::vector<WinInfo> wins;
    int n_chunks_pt = 0, n_chunks_ls = 0, ls_dim = 4, max_nf = 0;
    bool uploaded = false;
    // initial state kept on the device for plba_reset_state()
    double *init_poseT = nullptr, *init_X = nullptr, *init_pts = nullptr, *init_lns = nullptr, *orth0 = nullptr;
    int *pt_win = nullptr, *ls_win = nullptr;
    double *sysbuf = nullptr; size_t sys_doubles = 0, S_doubles = 0;
    double *invbuf = nullptr;
    int *h_counters = nullptr;              // pinned
    WinCtrl *h_ctrl0 = nullptr;             // pinned, initial controller state
    std::vector<WinCtrl> ctrl_init;
    plba_allreduce_fn allreduce = nullptr; void *allreduce_user = nullptr;
    plba_timing timing{};
    cudaEvent_t ev[8]{};
    bool detail_timing = false;
    template <typename T> int dalloc(T **p, size_t n) {
        void *q = nullptr;
        if (cudaMalloc(&q, std::max<size_t>(n, 1) * sizeof(T)) != cudaSuccess) { err = "cudaMalloc failed"; return PLBA_E_CUDA; }
        allocs.push_back(q); *p = (T *)q; return PLBA_OK;
    }
    void free_all() { for (void *q : allocs) cudaFree(q); allocs.clear(); uploaded = false; }
};

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
    for (int i = 0; i < p.n_pobs; i++) {
        if (p.po_lm[i] < 0 || p.po_lm[i] >= p.n_pt || p.po_kf[i] < 0 || p.po_kf[i] >= p.n_kf) { err = "point observation index out of range"; return PLBA_E_ARG; }
        if (i && p.po_lm[i] < p.po_lm[i - 1]) { err = "point observations not landmark-major"; return PLBA_E_ARG; }
    }
    for (int i = 0; i < p.n_lobs; i++) {
        if (p.lo_lm[i] < 0 || p.lo_lm[i] >= p.n_ls || p.lo_kf[i] < 0 || p.lo_kf[i] >= p.n_kf) { err = "line observation index out of range"; return PLBA_E_ARG; }
        if (i && p.lo_lm[i] < p.lo_lm[i - 1]) { err = "line observations not landmark-major"; return PLBA_E_ARG; }
    }
    return PLBA_OK;
}

static void build_chunks(const std::vector<int> &ptr, int lm0, int lm1, int win, std::vector<Chunk> &out, bool &too_long) {
    int l = lm0;
    while (l < lm1) {
        Chunk c{}; c.lm0 = l; c.ob0 = ptr[l]; c.win = win;
        int e = l;
        while (e < lm1 && (e - l) < LC && (ptr[e + 1] - c.ob0) <= OC) e++;
        if (e == l) { too_long = true; e = l + 1; }
        c.lm1 = e; c.ob1 = ptr[e];
        out.push_back(c);
        l = e;
    }
}

template <typename T> static int h2d(plba_handle h, T *dst, const std::vector<T> &src) {
    if (src.empty()) return PLBA_OK;
    CK(cudaMemcpyAsync(dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, h->stream));
    h->timing.h2d_bytes += (int64_t)(src.size() * sizeof(T));
    return PLBA_OK;
}
#define UP(dptr, vec) do { int rc_ = h->dalloc(&dptr, (vec).size()); if (rc_) return rc_; rc_ = h2d(h, dptr, vec); if (rc_) return rc_; } while (0)

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
    cudaMallocHost((void **)&h->h_counters, sizeof(int) * CNT_N);
    for (int i = 0; i < 8; i++) cudaEventCreate(&h->ev[i]);
    *out = h;
    return PLBA_OK;
}

void plba_destroy(plba_handle h) {
    if (!h) return;
    cudaSetDevice(h->device);
    h->free_all();
    if (h->h_counters) cudaFreeHost(h->h_counters);
    for (int i = 0; i < 8; i++) cudaEventDestroy(h->ev[i]);
    if (h->own_stream) cudaStreamDestroy(h->stream);
    delete h;
}

const char *plba_last_error(plba_handle h) { return h ? h->err.c_str() : "null handle"; }

int plba_set_allreduce(plba_handle h, plba_allreduce_fn fn, void *user) { if (!h) return PLBA_E_ARG; h->allreduce = fn; h->allreduce_user = user; return PLBA_OK; }

int plba_upload(plba_handle h, int32_t n, const plba_problem *probs, const plba_options *opt) {
    if (!h || n <= 0 || !probs || !opt) return PLBA_E_ARG;
    if (opt->profile < PLBA_PROFILE_G || opt->profile > PLBA_PROFILE_H_PLK) { h->err = "unknown profile"; return PLBA_E_ARG; }
    CK(cudaSetDevice(h->device));
    h->free_all();
    h->opt = *opt;
    h->timing = plba_timing{};
    const int prof = opt->profile;
    h->ls_dim = (prof == PLBA_PROFILE_H_END) ? 6 : 4;
    const int ld = h->ls_dim;
    h->wins.assign(n, WinInfo{});
    WinInfo tot{};
    for (int w = 0; w < n; w++) {
        const plba_problem &p = probs[w];
        int rc = validate_problem(p, *opt, h->err);
        if (rc) return rc;
        WinInfo &wi = h->wins[w];
        wi.n_kf = p.n_kf; wi.n_free = p.n_free; wi.n_pt = p.n_pt; wi.n_ls = p.n_ls; wi.n_pobs = p.n_pobs; wi.n_lobs = p.n_lobs;
        wi.kf0 = tot.n_kf; wi.slot0 = tot.n_free; wi.pt0 = tot.n_pt; wi.ls0 = tot.n_ls; wi.po0 = tot.n_pobs; wi.lo0 = tot.n_lobs;
        if ((int64_t)tot.n_pobs + p.n_pobs > 0x7fffffff || (int64_t)tot.n_lobs + p.n_lobs > 0x7fffffff) { h->err = "batch too large"; return PLBA_E_ARG; }
        tot.n_kf += p.n_kf; tot.n_free += p.n_free; tot.n_pt += p.n_pt; tot.n_ls += p.n_ls; tot.n_pobs += p.n_pobs; tot.n_lobs += p.n_lobs;
    }
    // ---- flatten on the host (index arithmetic only) ----
    std::vector<int> kf_slot(tot.n_kf), kf_win(tot.n_kf), slot_kf(tot.n_free), win_slot0(n), win_nfree(n), win_ls0(n), pt_win(tot.n_pt), ls_win(tot.n_ls);
    std::vector<long long> win_S_off(n);
    std::vector<double> Tmap((size_t)tot.n_kf * 12), X0((size_t)tot.n_free * 6), pts((size_t)tot.n_pt * 3), lns((size_t)tot.n_ls * ld), lns_map((size_t)tot.n_ls * 6);
    std::vector<int> po_kf(tot.n_pobs), po_lm(tot.n_pobs), lo_kf(tot.n_lobs), lo_lm(tot.n_lobs), pt_ptr(tot.n_pt + 1, 0), ls_ptr(tot.n_ls + 1, 0);
    std::vector<double> po_uv((size_t)tot.n_pobs * 2), lo_ab((size_t)tot.n_lobs * 4), po_om(tot.n_pobs), lo_om(tot.n_lobs);
    std::vector<Chunk> ch_pt, ch_ls;
    h->ctrl_init.assign(n, WinCtrl{});
    long long S_off = 0; int max_nf = 0; bool too_long = false;
    for (int w = 0; w < n; w++) {
        const plba_problem &p = probs[w]; const WinInfo &wi = h->wins[w];
        win_slot0[w] = wi.slot0; win_nfree[w] = wi.n_free; win_ls0[w] = wi.ls0; win_S_off[w] = S_off;
        S_off += (long long)36 * wi.n_free * wi.n_free;
        max_nf = std::max(max_nf, wi.n_free);
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
        for (int l = 0; l < p.n_pt; l++) { pt_win[wi.pt0 + l] = w; for (int i = 0; i < 3; i++) pts[(size_t)(wi.pt0 + l) * 3 + i] = p.pt_xyz[(size_t)3 * l + i]; }
        for (int l = 0; l < p.n_ls; l++) {
            ls_win[wi.ls0 + l] = w;
            if (prof == PLBA_PROFILE_H_END) for (int i = 0; i < 6; i++) lns[(size_t)(wi.ls0 + l) * 6 + i] = p.ls_end[(size_t)6 * l + i];
            else {
                plk_to_orth(p.ls_plk + (size_t)6 * l, &lns[(size_t)(wi.ls0 + l) * 4]);                // changePlukerToOrth (:6040, :1577)
                for (int i = 0; i < 6; i++) lns_map[(size_t)(wi.ls0 + l) * 6 + i] = p.ls_plk[(size_t)6 * l + i];
            }
        }
        for (int i = 0; i < p.n_pobs; i++) {
            const int o = wi.po0 + i;
            po_kf[o] = wi.kf0 + p.po_kf[i]; po_lm[o] = wi.pt0 + p.po_lm[i]; pt_ptr[po_lm[o] + 1]++;
            po_uv[(size_t)2 * o] = p.po_uv[(size_t)2 * i]; po_uv[(size_t)2 * o + 1] = p.po_uv[(size_t)2 * i + 1];
            po_om[o] = (double)(float)(1.0 / (p.po_sig2 ? p.po_sig2[i] : 1.0));                      // const float& invSigma2 (:6009, Q13)
        }
        for (int i = 0; i < p.n_lobs; i++) {
            const int o = wi.lo0 + i;
            lo_kf[o] = wi.kf0 + p.lo_kf[i]; lo_lm[o] = wi.ls0 + p.lo_lm[i]; ls_ptr[lo_lm[o] + 1]++;
            for (int k = 0; k < 4; k++) lo_ab[(size_t)4 * o + k] = p.lo_ab[(size_t)4 * i + k];
            lo_om[o] = (double)(float)(1.0 / (p.lo_sig2 ? p.lo_sig2[i] : 1.0));
        }
        WinCtrl &c = h->ctrl_init[w];
        c.need_init = 1; c.apply = 1; c.ni = 2.0; c.err_prev = 999999999.9; c.n_lm_pt = p.n_pt; c.n_lm_ls = p.n_ls;
        if (p.n_pobs + p.n_lobs == 0) c.done = 1;      // nothing to do (src/mapHandler.cpp:1496-1500)
    }
    for (int l = 0; l < tot.n_pt; l++) pt_ptr[l + 1] += pt_ptr[l];
    for (int l = 0; l < tot.n_ls; l++) ls_ptr[l + 1] += ls_ptr[l];
    for (int w = 0; w < n; w++) {
        const WinInfo &wi = h->wins[w];
        build_chunks(pt_ptr, wi.pt0, wi.pt0 + wi.n_pt, w, ch_pt, too_long);
        build_chunks(ls_ptr, wi.ls0, wi.ls0 + wi.n_ls, w, ch_ls, too_long);
    }
    if (too_long) { h->err = "a landmark has more than 256 observations"; return PLBA_E_UNSUPPORTED; }
    h->n_chunks_pt = (int)ch_pt.size(); h->n_chunks_ls = (int)ch_ls.size(); h->max_nf = max_nf;

    DevP &P = h->P;
    P = DevP{};
    P.cam = Cam{probs[0].cam[0], probs[0].cam[1], probs[0].cam[2], probs[0].cam[3]};
    for (int w = 1; w < n; w++) for (int i = 0; i < 4; i++) if (probs[w].cam[i] != probs[0].cam[i]) { h->err = "all windows of a batch must share the camera"; return PLBA_E_UNSUPPORTED; }
    P.profile = prof; P.fixed_quirks = (opt->quirks == PLBA_QUIRKS_FIXED);
    P.n_win = n; P.n_kf = tot.n_kf; P.n_free = tot.n_free; P.n_pt = tot.n_pt; P.n_ls = tot.n_ls; P.n_pobs = tot.n_pobs; P.n_lobs = tot.n_lobs;
    P.iters_stage1 = opt->iters_stage1; P.iters_stage2 = opt->iters_stage2; P.lm_max_trials = opt->lm_max_trials; P.max_iters_lba = opt->max_iters_lba;
    P.huber_delta = opt->huber_delta; P.chi2_gate = opt->chi2_gate; P.homog_th = opt->homog_th; P.min_error = opt->min_error;
    P.min_error_change = opt->min_error_change; P.lm_tau = opt->lm_tau; P.lambda_lba_lm = opt->lambda_lba_lm; P.lambda_lba_k = opt->lambda_lba_k;
    int *d_i; double *d_d; long long *d_ll; Chunk *d_c;
    UP(d_i, kf_slot); P.kf_slot = d_i; UP(d_i, kf_win); P.kf_win = d_i; UP(d_i, slot_kf); P.slot_kf = d_i;
    UP(d_i, win_slot0); P.win_slot0 = d_i; UP(d_i, win_nfree); P.win_nfree = d_i; UP(d_i, win_ls0); P.win_ls0 = d_i;
    UP(d_ll, win_S_off); P.win_S_off = d_ll;
    UP(d_d, Tmap); P.kf_Tmap = d_d; h->init_poseT = d_d;
    UP(d_d, X0); h->init_X = d_d; UP(d_d, pts); h->init_pts = d_d; UP(d_d, lns); h->init_lns = d_d; h->orth0 = d_d;
    UP(d_d, lns_map); P.lns_map = d_d;
    UP(d_i, pt_ptr); P.pt_ptr = d_i; UP(d_i, ls_ptr); P.ls_ptr = d_i;
    UP(d_i, po_kf); P.po_kf = d_i; UP(d_i, po_lm); P.po_lm = d_i; UP(d_i, lo_kf); P.lo_kf = d_i; UP(d_i, lo_lm); P.lo_lm = d_i;
    UP(d_d, po_uv); P.po_uv = d_d; UP(d_d, lo_ab); P.lo_ab = d_d; UP(d_d, po_om); P.po_om = d_d; UP(d_d, lo_om); P.lo_om = d_d;
    UP(d_c, ch_pt); P.chunks_pt = d_c; UP(d_c, ch_ls); P.chunks_ls = d_c;
    UP(h->pt_win, pt_win); UP(h->ls_win, ls_win);
    int rc;
    for (int b = 0; b < 2; b++) {
        if ((rc = h->dalloc(&P.poseT[b], (size_t)tot.n_kf * 12))) return rc;
        if ((rc = h->dalloc(&P.Xkf[b], (size_t)tot.n_free * 6))) return rc;
        if ((rc = h->dalloc(&P.pts[b], (size_t)tot.n_pt * 3))) return rc;
        if ((rc = h->dalloc(&P.lns[b], (size_t)tot.n_ls * ld))) return rc;
    }
    if ((rc = h->dalloc(&P.po_lvl, (size_t)tot.n_pobs))) return rc;
    if ((rc = h->dalloc(&P.lo_lvl, (size_t)tot.n_lobs))) return rc;
    if ((rc = h->dalloc(&P.po_chi2, (size_t)tot.n_pobs))) return rc;
    if ((rc = h->dalloc(&P.lo_chi2, (size_t)tot.n_lobs))) return rc;
    // reduced-system buffer: [S | g | hpp_diag | hpp_diag_init | acc | accmax]  (one memset, one all-reduce range)
    h->S_doubles = (size_t)S_off;
    h->sys_doubles = h->S_doubles + (size_t)18 * tot.n_free + (size_t)ACC_N * n + (size_t)n;
    if ((rc = h->dalloc(&h->sysbuf, h->sys_doubles))) return rc;
    P.S = h->sysbuf; P.gs = P.S + h->S_doubles; P.hpp_diag = P.gs + (size_t)6 * tot.n_free; P.hpp_diag_init = P.hpp_diag + (size_t)6 * tot.n_free;
    P.acc = P.hpp_diag_init + (size_t)6 * tot.n_free; P.accB = P.acc + (size_t)4 * n; P.accmax = P.acc + (size_t)ACC_N * n;
    if ((rc = h->dalloc(&P.xp, (size_t)6 * tot.n_free))) return rc;
    if ((rc = h->dalloc(&P.ctrl, (size_t)n))) return rc;
    P.trace_cap = (prof == PLBA_PROFILE_G) ? (opt->iters_stage1 + opt->iters_stage2) * opt->lm_max_trials + 2 : opt->max_iters_lba + 2;
    if ((rc = h->dalloc(&P.trace, (size_t)n * P.trace_cap))) return rc;
    if ((rc = h->dalloc(&P.counters, (size_t)CNT_N))) return rc;
    if (6 * max_nf > SMALL_NMAX) { const int nt = (6 * max_nf + TB - 1) / TB; if ((rc = h->dalloc(&h->invbuf, (size_t)nt * TB * TB))) return rc; }
    h->uploaded = true;
    return plba_reset_state(h);
}

int plba_reset_state(plba_handle h) {
    if (!h || !h->uploaded) return PLBA_E_ARG;
    DevP &P = h->P; cudaStream_t st = h->stream;
    for (int b = 0; b < 2; b++) {
        CK(cudaMemcpyAsync(P.poseT[b], h->init_poseT, sizeof(double) * 12 * (size_t)P.n_kf, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(P.Xkf[b], h->init_X, sizeof(double) * 6 * (size_t)P.n_free, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(P.pts[b], h->init_pts, sizeof(double) * 3 * (size_t)P.n_pt, cudaMemcpyDeviceToDevice, st));
        CK(cudaMemcpyAsync(P.lns[b], h->init_lns, sizeof(double) * h->ls_dim * (size_t)P.n_ls, cudaMemcpyDeviceToDevice, st));
    }
    CK(cudaMemsetAsync(P.po_lvl, 0, (size_t)P.n_pobs, st));
    CK(cudaMemsetAsync(P.lo_lvl, 0, (size_t)P.n_lobs, st));
    CK(cudaMemsetAsync(P.po_chi2, 0, sizeof(double) * (size_t)P.n_pobs, st));
    CK(cudaMemsetAsync(P.lo_chi2, 0, sizeof(double) * (size_t)P.n_lobs, st));
    CK(cudaMemsetAsync(P.trace, 0, sizeof(plba_trace_rec) * (size_t)P.n_win * P.trace_cap, st));
    CK(cudaMemcpyAsync(P.ctrl, h->ctrl_init.data(), sizeof(WinCtrl) * (size_t)P.n_win, cudaMemcpyHostToDevice, st));
    int ndone = 0; for (const WinCtrl &c : h->ctrl_init) ndone += c.done;
    h->h_counters[CNT_DONE] = ndone; h->h_counters[CNT_NEED_INIT] = P.n_win - ndone; h->h_counters[CNT_GATE] = 0; h->h_counters[CNT_TRIALS] = 0;
    CK(cudaMemcpyAsync(P.counters, h->h_counters, sizeof(int) * CNT_N, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
    return PLBA_OK;
}

}  // extern "C"

// ---- launch helpers ------------------------------------------------------------------------------------------
template <int PROF, int LT> static void set_smem_attr() {
#ifndef PLBA_HOST_EMU
    static bool done = false;
    if (!done) {
        cudaFuncSetAttribute(k_assemble<PROF, LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<PROF, LT>::bytes());
        cudaFuncSetAttribute(k_update<PROF, LT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Smem<PROF, LT>::bytes());
        done = true;
    }
#endif
}
template <int PROF, int LT> static void launch_assemble(plba_handle h, int nchunks, int mode) {
    if (!nchunks) return;
    set_smem_attr<PROF, LT>();
    const size_t smem = Smem<PROF, LT>::bytes();
    PLBA_LAUNCH((k_assemble<PROF, LT>), dim3(nchunks), dim3(OC), smem, h->stream, h->P, mode);
    h->timing.n_launches++; if (mode == 1) h->timing.n_assemble++;
}
template <int PROF, int LT> static void launch_update(plba_handle h, int nchunks) {
    if (!nchunks) return;
    set_smem_attr<PROF, LT>();
    const size_t smem = Smem<PROF, LT>::bytes();
    PLBA_LAUNCH((k_update<PROF, LT>), dim3(nchunks), dim3(OC), smem, h->stream, h->P);
    h->timing.n_launches++;
}
static void do_assemble(plba_handle h, int mode) {
    switch (h->opt.profile) {
    case PLBA_PROFILE_G: launch_assemble<PLBA_PROFILE_G, LT_POINT>(h, h->n_chunks_pt, mode); launch_assemble<PLBA_PROFILE_G, LT_LINE_ORTH>(h, h->n_chunks_ls, mode); break;
    case PLBA_PROFILE_H_END: launch_assemble<PLBA_PROFILE_H_END, LT_POINT>(h, h->n_chunks_pt, mode); launch_assemble<PLBA_PROFILE_H_END, LT_LINE_END>(h, h->n_chunks_ls, mode); break;
    default: launch_assemble<PLBA_PROFILE_H_PLK, LT_POINT>(h, h->n_chunks_pt, mode); launch_assemble<PLBA_PROFILE_H_PLK, LT_LINE_ORTH>(h, h->n_chunks_ls, mode); break;
    }
}
static void do_update(plba_handle h) {
    switch (h->opt.profile) {
    case PLBA_PROFILE_G: launch_update<PLBA_PROFILE_G, LT_POINT>(h, h->n_chunks_pt); launch_update<PLBA_PROFILE_G, LT_LINE_ORTH>(h, h->n_chunks_ls); break;
    case PLBA_PROFILE_H_END: launch_update<PLBA_PROFILE_H_END, LT_POINT>(h, h->n_chunks_pt); launch_update<PLBA_PROFILE_H_END, LT_LINE_END>(h, h->n_chunks_ls); break;
    default: launch_update<PLBA_PROFILE_H_PLK, LT_POINT>(h, h->n_chunks_pt); launch_update<PLBA_PROFILE_H_PLK, LT_LINE_ORTH>(h, h->n_chunks_ls); break;
    }
}
static void do_solve(plba_handle h) {
    DevP &P = h->P;
    if (P.n_free == 0) return;
    if (6 * h->max_nf <= SMALL_NMAX) {
        const int n = 6 * h->max_nf, ldm = n + 1;
        const size_t sm = solve_small_smem(n, ldm);
#ifndef PLBA_HOST_EMU
        static size_t attr_set = 0;
        if (sm > attr_set) { cudaFuncSetAttribute(k_solve_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm); attr_set = sm; }
#endif
        PLBA_LAUNCH(k_solve_small, dim3(P.n_win), dim3(256), sm, h->stream, P, ldm);
        h->timing.n_launches++;
        return;
    }
    for (int w = 0; w < P.n_win; w++) {
        const int n = 6 * h->wins[w].n_free;
        if (n == 0 || h->ctrl_init[w].done) continue;
        const int nt = (n + TB - 1) / TB;
        for (int k = 0; k < nt; k++) {
            PLBA_LAUNCH(k_potrf_tile, dim3(1), dim3(256), sizeof(double) * (2 * TB * (TB + 1) + TB), h->stream, P, w, k, h->invbuf);
            h->timing.n_launches++;
            const int m = nt - k - 1;
            if (m > 0) {
                PLBA_LAUNCH(k_trsm_tiles, dim3(m), dim3(256), sizeof(double) * 2 * TB * TB, h->stream, P, w, k, (const double *)h->invbuf);
                PLBA_LAUNCH(k_syrk_tiles, dim3(m * (m + 1) / 2), dim3(256), sizeof(double) * 2 * TB * TB, h->stream, P, w, k, nt);
                h->timing.n_launches += 2;
            }
        }
        const int ntd = 1008;   // multiple of TB: G = 21 partial sums per row
        PLBA_LAUNCH(k_trisolve_large, dim3(1), dim3(ntd), sizeof(double) * (TB + TB * (ntd / TB)), h->stream, P, w, nt, (const double *)h->invbuf);
        h->timing.n_launches++;
    }
}
static void allreduce(plba_handle h, double *p, size_t n, int op_max) {
    if (!h->allreduce || !n) return;
    // op is encoded in the sign of the count for the max reduction (only the lambda-init scalar uses it)
    h->allreduce(p, op_max ? -(int64_t)n : (int64_t)n, (void *)h->stream, h->allreduce_user);
}
static int poll_counters(plba_handle h) {
    CK(cudaMemcpyAsync(h->h_counters, h->P.counters, sizeof(int) * CNT_N, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PLBA_OK;
}
PLBA_KERNEL void k_set_lambda(DevP P, double lambda) {
    PHASE_BEGIN
        const int w = PLBA_BID * PLBA_NT + tid;
        if (w < P.n_win) { P.ctrl[w].lambda = lambda; P.ctrl[w].need_init = 0; }
    PHASE_END
}
static inline dim3 grid1(int n, int b) { return dim3((unsigned)std::max(1, (n + b - 1) / b)); }

// one LM round for every window that is not finished: (gate) -> (lambda init) -> assemble -> solve -> update -> control
static int run_round(plba_handle h, bool need_init, bool need_gate) {
    DevP &P = h->P; cudaStream_t st = h->stream;
    const bool G = (P.profile == PLBA_PROFILE_G);
    if (need_gate) {
        if (P.n_pobs) { PLBA_LAUNCH(k_gate<LT_POINT>, grid1(P.n_pobs, 256), dim3(256), 0, st, P, P.n_pobs); h->timing.n_launches++; }
        if (P.n_lobs) { PLBA_LAUNCH(k_gate<LT_LINE_ORTH>, grid1(P.n_lobs, 256), dim3(256), 0, st, P, P.n_lobs); h->timing.n_launches++; }
    }
    CK(cudaMemsetAsync(h->sysbuf, 0, sizeof(double) * h->sys_doubles, st));
    if (need_init || need_gate) {
        if (need_init) {
            do_assemble(h, 0);
            allreduce(h, P.hpp_diag_init, (size_t)6 * P.n_free, 0);
            allreduce(h, P.accmax, (size_t)P.n_win, 1);
        }
        PLBA_LAUNCH(k_lambda_init, grid1(P.n_win, 128), dim3(128), 0, st, P); h->timing.n_launches++;
    }
    if (h->detail_timing) cudaEventRecord(h->ev[0], st);
    do_assemble(h, 1);
    if (h->detail_timing) cudaEventRecord(h->ev[1], st);
    allreduce(h, h->sysbuf, h->S_doubles + (size_t)12 * P.n_free, 0);            // the exchange step: S, g, hpp_diag
    allreduce(h, P.acc, (size_t)4 * P.n_win, 0);                                 // assemble-phase cost sums
    if (!G) { PLBA_LAUNCH(k_control_h_pre, grid1(P.n_win, 128), dim3(128), 0, st, P); h->timing.n_launches++; }
    do_solve(h);
    if (h->detail_timing) cudaEventRecord(h->ev[2], st);
    if (P.n_free) { PLBA_LAUNCH(k_pose_update, grid1(P.n_free, 128), dim3(128), 0, st, P); h->timing.n_launches++; }
    do_update(h);
    if (h->detail_timing) cudaEventRecord(h->ev[3], st);
    return PLBA_OK;
}

extern "C" {

int plba_run(plba_handle h) {
    if (!h || !h->uploaded) return PLBA_E_ARG;
    CK(cudaSetDevice(h->device));
    DevP &P = h->P; cudaStream_t st = h->stream;
    const bool G = (P.profile == PLBA_PROFILE_G);
    const int max_rounds = G ? (P.iters_stage1 + P.iters_stage2) * P.lm_max_trials + 2 : P.max_iters_lba + 1;
    const int64_t launches0 = h->timing.n_launches, assemble0 = h->timing.n_assemble;
    h->timing.ms_assemble = h->timing.ms_solve = h->timing.ms_update = 0;
    cudaEventRecord(h->ev[6], st);
    int rc;
    // h_counters holds the state left by plba_reset_state()
    for (int round = 0; round < max_rounds; round++) {
        if (h->h_counters[CNT_DONE] >= P.n_win) break;
        const bool need_init = h->h_counters[CNT_NEED_INIT] > 0, need_gate = h->h_counters[CNT_GATE] > 0;
        if ((rc = run_round(h, need_init, need_gate))) return rc;
        allreduce(h, P.accB, (size_t)4 * P.n_win, 0);                            // update-phase sums (new cost, scale, |dx|^2)
        PLBA_LAUNCH(k_control, grid1(P.n_win, 128), dim3(128), 0, st, P); h->timing.n_launches++;
        if ((rc = poll_counters(h))) return rc;
        if (h->detail_timing) {
            float a = 0, b = 0, c = 0;
            cudaEventElapsedTime(&a, h->ev[0], h->ev[1]); cudaEventElapsedTime(&b, h->ev[1], h->ev[2]); cudaEventElapsedTime(&c, h->ev[2], h->ev[3]);
            h->timing.ms_assemble += a; h->timing.ms_solve += b; h->timing.ms_update += c;
        }
    }
    if (G) {   // pending gate of windows that stopped right after stage 0, then the final per-edge test
        if (h->h_counters[CNT_GATE] > 0) {
            if (P.n_pobs) { PLBA_LAUNCH(k_gate<LT_POINT>, grid1(P.n_pobs, 256), dim3(256), 0, st, P, P.n_pobs); h->timing.n_launches++; }
            if (P.n_lobs) { PLBA_LAUNCH(k_gate<LT_LINE_ORTH>, grid1(P.n_lobs, 256), dim3(256), 0, st, P, P.n_lobs); h->timing.n_launches++; }
            PLBA_LAUNCH(k_lambda_init, grid1(P.n_win, 128), dim3(128), 0, st, P); h->timing.n_launches++;
        }
    }
    cudaEventRecord(h->ev[7], st);
    CK(cudaStreamSynchronize(st));
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
    const bool G = (P.profile == PLBA_PROFILE_G);
    const int ld = h->ls_dim;
    // device-side write-back into staging buffers
    double *d_T, *d_x, *d_pt, *d_ls, *d_plk; unsigned char *d_pf, *d_lf;
    int rc;
    if ((rc = h->dalloc(&d_T, (size_t)12 * P.n_kf))) return rc;
    if ((rc = h->dalloc(&d_x, (size_t)6 * P.n_free))) return rc;
    if ((rc = h->dalloc(&d_pt, (size_t)3 * P.n_pt))) return rc;
    if ((rc = h->dalloc(&d_ls, (size_t)ld * P.n_ls))) return rc;
    if ((rc = h->dalloc(&d_plk, (size_t)6 * P.n_ls))) return rc;
    if ((rc = h->dalloc(&d_pf, (size_t)P.n_pobs))) return rc;
    if ((rc = h->dalloc(&d_lf, (size_t)P.n_lobs))) return rc;
    if (G) {
        if (P.n_pobs) { PLBA_LAUNCH(k_final<LT_POINT>, grid1(P.n_pobs, 256), dim3(256), 0, st, P, P.n_pobs, d_pf); h->timing.n_launches++; }
        if (P.n_lobs) { PLBA_LAUNCH(k_final<LT_LINE_ORTH>, grid1(P.n_lobs, 256), dim3(256), 0, st, P, P.n_lobs, d_lf); h->timing.n_launches++; }
    }
    if (P.n_kf) { PLBA_LAUNCH(k_export_poses, grid1(P.n_kf, 128), dim3(128), 0, st, P, d_T, d_x); h->timing.n_launches++; }
    const int nl = std::max(P.n_pt, P.n_ls);
    if (nl) {
        const int q9 = (P.profile == PLBA_PROFILE_H_PLK && !P.fixed_quirks) ? 1 : 0;
        PLBA_LAUNCH(k_export_landmarks, grid1(nl, 256), dim3(256), 0, st, P, (const int *)h->pt_win, (const int *)h->ls_win, d_pt, d_ls, d_plk, ld, (const double *)h->orth0, q9);
        h->timing.n_launches++;
    }
    std::vector<double> T((size_t)12 * P.n_kf), X((size_t)6 * P.n_free), pt((size_t)3 * P.n_pt), ls((size_t)ld * P.n_ls), plk((size_t)6 * P.n_ls), pchi(P.n_pobs), lchi(P.n_lobs);
    std::vector<double> pt0((size_t)3 * P.n_pt), ls0((size_t)ld * P.n_ls);
    std::vector<unsigned char> pf(P.n_pobs), lf(P.n_lobs);
    std::vector<WinCtrl> ctrl(P.n_win);
    std::vector<plba_trace_rec> trace((size_t)P.n_win * P.trace_cap);
    auto d2h = [&](void *dst, const void *src, size_t bytes) -> int { if (!bytes) return 0; h->timing.d2h_bytes += (int64_t)bytes; return cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st) == cudaSuccess ? 0 : 1; };
    int bad = 0;
    bad |= d2h(T.data(), d_T, T.size() * 8); bad |= d2h(X.data(), d_x, X.size() * 8); bad |= d2h(pt.data(), d_pt, pt.size() * 8);
    bad |= d2h(ls.data(), d_ls, ls.size() * 8); bad |= d2h(plk.data(), d_plk, plk.size() * 8);
    bad |= d2h(pt0.data(), h->init_pts, pt0.size() * 8); bad |= d2h(ls0.data(), h->init_lns, ls0.size() * 8);
    if (G) { bad |= d2h(pchi.data(), P.po_chi2, pchi.size() * 8); bad |= d2h(lchi.data(), P.lo_chi2, lchi.size() * 8); bad |= d2h(pf.data(), d_pf, pf.size()); bad |= d2h(lf.data(), d_lf, lf.size()); }
    bad |= d2h(ctrl.data(), P.ctrl, ctrl.size() * sizeof(WinCtrl)); bad |= d2h(trace.data(), P.trace, trace.size() * sizeof(plba_trace_rec));
    if (bad) { h->err = "device to host copy failed"; return PLBA_E_CUDA; }
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    // staging buffers were the last 7 allocations
    for (int i = 0; i < 7; i++) { cudaFree(h->allocs.back()); h->allocs.pop_back(); }
    int rc_all = PLBA_OK;
    for (int w = 0; w < n; w++) {
        const WinInfo &wi = h->wins[w]; plba_result &r = res[w];
        r.n_trace = ctrl[w].n_trace; r.n_trials = ctrl[w].n_trials;
        r.status = (wi.n_pobs + wi.n_lobs == 0) ? PLBA_DISCARDED : PLBA_OK;
        if (r.trace) for (int i = 0; i < std::min(r.n_trace, std::min(r.trace_cap, P.trace_cap)); i++) r.trace[i] = trace[(size_t)w * P.trace_cap + i];
        if (r.x_pose) for (int i = 0; i < 6 * wi.n_free; i++) r.x_pose[i] = X[(size_t)6 * wi.slot0 + i];
        if (r.pt_xyz) for (int i = 0; i < 3 * wi.n_pt; i++) r.pt_xyz[i] = pt[(size_t)3 * wi.pt0 + i];
        if (ld == 4) {
            if (r.ls_orth) for (int i = 0; i < 4 * wi.n_ls; i++) r.ls_orth[i] = ls[(size_t)4 * wi.ls0 + i];
            if (r.ls_plk) for (int i = 0; i < 6 * wi.n_ls; i++) r.ls_plk[i] = plk[(size_t)6 * wi.ls0 + i];
        } else if (r.ls_end) for (int i = 0; i < 6 * wi.n_ls; i++) r.ls_end[i] = ls[(size_t)6 * wi.ls0 + i];
        // inlier rule of the hand-LM write-back (src/mapHandler.cpp:2858-2860, 2871-2873); profile G leaves it to the final chi2 test
        if (r.pt_inlier) for (int l = 0; l < wi.n_pt; l++) {
            double d2 = 0; for (int i = 0; i < 3; i++) { const double d = pt[(size_t)3 * (wi.pt0 + l) + i] - pt0[(size_t)3 * (wi.pt0 + l) + i]; d2 += d * d; }
            r.pt_inlier[l] = (!G && std::sqrt(d2) > 0.01) ? 0 : 1;
        }
        if (r.ls_inlier) for (int l = 0; l < wi.n_ls; l++) {
            double d2 = 0; for (int i = 0; i < ld; i++) { const double d = ls[(size_t)ld * (wi.ls0 + l) + i] - ls0[(size_t)ld * (wi.ls0 + l) + i]; d2 += d * d; }
            r.ls_inlier[l] = (!G && std::sqrt(d2) > 0.01) ? 0 : 1;
        }
        if (G) {
            if (r.po_chi2) for (int i = 0; i < wi.n_pobs; i++) r.po_chi2[i] = pchi[(size_t)wi.po0 + i];
            if (r.lo_chi2) for (int i = 0; i < wi.n_lobs; i++) r.lo_chi2[i] = lchi[(size_t)wi.lo0 + i];
            if (r.po_flags) for (int i = 0; i < wi.n_pobs; i++) r.po_flags[i] = pf[(size_t)wi.po0 + i];
            if (r.lo_flags) for (int i = 0; i < wi.n_lobs; i++) r.lo_flags[i] = lf[(size_t)wi.lo0 + i];
        }
        if (r.status < PLBA_DISCARDED) rc_all = r.status;
    }
    for (int w = 0; w < n; w++) {
        const WinInfo &wi = h->wins[w]; plba_result &r = res[w];
        if (!r.kf_T_wc) continue;
        for (int k = 0; k < wi.n_kf; k++) for (int i = 0; i < 12; i++) r.kf_T_wc[(size_t)12 * k + i] = T[(size_t)12 * (wi.kf0 + k) + i];
    }
    return rc_all;
}

int plba_solve_batch(plba_handle h, int32_t n, const plba_problem *probs, const plba_options *opt, plba_result *res) {
    if (!h || !res) return PLBA_E_ARG;
    int rc = plba_upload(h, n, probs, opt);
    if (rc) { for (int w = 0; w < n; w++) { res[w].status = rc; res[w].n_trace = 0; res[w].n_trials = 0; } return rc; }
    if ((rc = plba_run(h))) return rc;
    rc = plba_download(h, n, res);
    // fixed KFs: the exported buffer only holds free KFs; copy the caller's rows through
    for (int w = 0; w < n; w++) if (res[w].kf_T_wc) for (int k = 0; k < probs[w].n_kf; k++) if (probs[w].kf_slot[k] < 0) for (int i = 0; i < 12; i++) res[w].kf_T_wc[(size_t)12 * k + i] = probs[w].kf_T_wc[(size_t)12 * k + i];
    return rc;
}

int plba_solve(plba_handle h, const plba_problem *prob, const plba_options *opt, plba_result *res) {
    if (!prob || !res) return PLBA_E_ARG;
    if (prob->n_pobs + prob->n_lobs == 0) { res->status = PLBA_DISCARDED; res->n_trace = 0; res->n_trials = 0; return PLBA_DISCARDED; }   // src/mapHandler.cpp:1496-1500
    int rc = plba_solve_batch(h, 1, prob, opt, res);
    return rc ? rc : res->status;
}

// ---- single-trial interface (bench roofline timing, sharded drivers) ----
int plba_trial_assemble(plba_handle h, double lambda) {
    if (!h || !h->uploaded) return PLBA_E_ARG;
    DevP &P = h->P; cudaStream_t st = h->stream;
    PLBA_LAUNCH(k_set_lambda, grid1(P.n_win, 128), dim3(128), 0, st, P, lambda);
    CK(cudaMemsetAsync(h->sysbuf, 0, sizeof(double) * h->sys_doubles, st));
    do_assemble(h, 1);
    CK(cudaGetLastError());
    return PLBA_OK;
}
int plba_trial_finish(plba_handle h, double lambda, double *chi_new, double *scale) {
    if (!h || !h->uploaded) return PLBA_E_ARG;
    DevP &P = h->P; cudaStream_t st = h->stream;
    (void)lambda;
    if (P.profile != PLBA_PROFILE_G) { PLBA_LAUNCH(k_control_h_pre, grid1(P.n_win, 128), dim3(128), 0, st, P); }
    do_solve(h);
    if (P.n_free) PLBA_LAUNCH(k_pose_update, grid1(P.n_free, 128), dim3(128), 0, st, P);
    do_update(h);
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
    if (n_doubles) *n_doubles = (int64_t)(h->S_doubles + (size_t)6 * h->P.n_free);
    return PLBA_OK;
}
int plba_copy_reduced_system(plba_handle h, int32_t window, double *S_out, double *g_out) {
    if (!h || !h->uploaded || window < 0 || window >= h->P.n_win) return PLBA_E_ARG;
    const WinInfo &wi = h->wins[window];
    const size_t n = (size_t)6 * wi.n_free;
    std::vector<long long> off(1);
    CK(cudaMemcpyAsync(off.data(), h->P.win_S_off + window, sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (S_out && n) CK(cudaMemcpyAsync(S_out, h->P.S + off[0], sizeof(double) * n * n, cudaMemcpyDeviceToHost, h->stream));
    if (g_out && n) CK(cudaMemcpyAsync(g_out, h->P.gs + (size_t)6 * wi.slot0, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    return PLBA_OK;
}
int plba_get_timing(plba_handle h, plba_timing *t) { if (!h || !t) return PLBA_E_ARG; *t = h->timing; return PLBA_OK; }
int plba_set_detail_timing(plba_handle h, int on) { if (!h) return PLBA_E_ARG; h->detail_timing = on != 0; return PLBA_OK; }

}  // extern "C"
