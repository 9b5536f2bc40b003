// LBA kernels for sm_100a (FP64).  One CTA owns a CHUNK of consecutive landmarks (all their observations are
// contiguous: reference order is landmark-major, src/mapHandler.cpp:1424-1447) and runs, without leaving shared memory:
//   phase 0  per landmark   : landmark-only quantities (orth -> Plücker, U, W) hoisted out of the per-edge path
//   phase 1  per observation: residual + Jacobians + robust weight  (a12-a14, a17, a18 of SURVEY.md §8a)
//   phase 2  per landmark   : H_ll, b_l, damping, H_ll^-1                      (subsystem 2)
//   phase 3  per pose pair  : Schur update of the reduced camera system S, g   (subsystem 3)
// k_update re-linearises the same chunk, back-substitutes, retracts and evaluates the new cost (subsystem 4).
// W (H_pl) blocks are never materialised in HBM; every input is read once per kernel.
//
// All kernels are written as PHASE blocks (plba_port.h) so that the control flow can be debugged on a CPU.
#pragma once
#include "plba_math.h"
#include "../../include/plba.h"

namespace plba {

enum { LT_POINT = 0, LT_LINE_ORTH = 1, LT_LINE_END = 2 };
enum { OC = 256 /* observations (= threads) per chunk */, LC = 128 /* landmarks per chunk */ };

struct Chunk { int lm0, lm1, ob0, ob1, win, pad; };

struct WinCtrl {
    int cur, stage, iter, trial, need_init, do_gate, done, n_trace, solve_fail, apply, stop_code, pad;
    int n_lm_pt, n_lm_ls;            // landmark counts of the window (profile H normalisation)
    double lambda, ni, chi_cur, err_prev;
    double scale_pose, dx2_pose;     // pose part of computeScale() / ||DX||^2 (replicated on every rank, never all-reduced)
    int n_trials, pad2;
};
// per-window accumulators that ARE summed over ranks (landmark-sharded multi-GPU): P.acc[8*w + ...]
// assemble-phase sums live in P.acc[4*w + ...], update-phase sums in P.accB[4*w + ...] (two all-reduce ranges)
enum { ACC_CHI_LIN = 0, ACC_ERR_PT = 1, ACC_ERR_LS = 2, ACC_CHI_NEW = 0, ACC_SCALE = 1, ACC_DX2 = 2, ACC_N = 8 };
enum { CNT_DONE = 0, CNT_NEED_INIT = 1, CNT_GATE = 2, CNT_TRIALS = 3, CNT_N = 4 };

struct DevP {
    Cam cam;
    int profile, fixed_quirks;
    int n_win, n_kf, n_free, n_pt, n_ls, n_pobs, n_lobs;
    int iters_stage1, iters_stage2, lm_max_trials, max_iters_lba;
    double huber_delta, chi2_gate, homog_th, min_error, min_error_change, lm_tau, lambda_lba_lm, lambda_lba_k;
    // keyframes
    const int *kf_slot, *kf_win, *slot_kf;
    const double *kf_Tmap;            // [n_kf][12]  T_cw of the map pose (fixed KFs, profile-H pass 0, Q4)
    double *poseT[2];                 // [n_kf][12]  T_cw estimate, double buffered (push / pop)
    double *Xkf[2];                   // [n_free][6] profile H: log of T_kf_w
    const int *win_slot0, *win_nfree; // per window: first global slot, number of free KFs
    const int *win_ls0;               // per window: first global line index (Q3 indexing is window-relative)
    const long long *win_S_off;       // per window: offset (doubles) of its dense (6 nf)^2 S
    // landmarks (double buffered)
    double *pts[2], *lns[2];
    const double *lns_map;            // [n_ls][6] map Plücker (H_PLK pass 0)
    const int *pt_ptr, *ls_ptr;       // CSR landmark -> observations
    // observations (SoA)
    const int *po_kf, *lo_kf, *po_lm, *lo_lm;
    const double *po_uv, *lo_ab, *po_om, *lo_om;
    unsigned char *po_lvl, *lo_lvl;
    double *po_chi2, *lo_chi2;
    const Chunk *chunks_pt, *chunks_ls;
    // linear system
    double *S, *gs, *xp, *hpp_diag, *hpp_diag_init;
    double *acc, *accB;               // [n_win][4] each: sum-reduced accumulators (tail of the reduced-system buffer)
    double *accmax;                   // [n_win]         max-reduced: largest landmark diagonal (lambda init)
    WinCtrl *ctrl;
    plba_trace_rec *trace; int trace_cap;
    int *counters;                    // [CNT_N]
};

template <int PROF, int LT> struct KT;
template <> struct KT<PLBA_PROFILE_G, LT_POINT>     { enum { RANK = 2, D = 3, NPRE = 3 }; };
template <> struct KT<PLBA_PROFILE_G, LT_LINE_ORTH> { enum { RANK = 2, D = 4, NPRE = 21 }; };
template <> struct KT<PLBA_PROFILE_H_END, LT_POINT>    { enum { RANK = 1, D = 3, NPRE = 3 }; };
template <> struct KT<PLBA_PROFILE_H_END, LT_LINE_END> { enum { RANK = 1, D = 6, NPRE = 6 }; };
template <> struct KT<PLBA_PROFILE_H_PLK, LT_POINT>     { enum { RANK = 1, D = 3, NPRE = 3 }; };
template <> struct KT<PLBA_PROFILE_H_PLK, LT_LINE_ORTH> { enum { RANK = 1, D = 4, NPRE = 21 }; };

// shared-memory carve-up of a chunk (SoA: [component][slot] so that thread-per-slot accesses are conflict free)
template <int PROF, int LT>
struct Smem {
    typedef KT<PROF, LT> K;
    enum { NSYM = K::D * (K::D + 1) / 2 };
    double *A, *B, *E, *U, *Hinv, *V, *Pre, *Xl, *red;
    int *slot, *lmof, *freeidx, *nfree, *pairptr;
    static size_t bytes() {
        return sizeof(double) * ((size_t)(K::RANK * 6 + K::RANK * K::D + 2 * K::RANK) * OC + (size_t)(NSYM + 2 * K::D + K::NPRE) * LC + 8)
             + sizeof(int) * (3 * OC + 2 * LC + 8);
    }
    PLBA_HD explicit Smem(unsigned char *raw) {
        double *p = (double *)raw;
        A = p; p += K::RANK * 6 * OC;
        B = p; p += K::RANK * K::D * OC;
        E = p; p += K::RANK * OC;
        U = p; p += K::RANK * OC;
        Hinv = p; p += NSYM * LC;
        V = p; p += K::D * LC;
        Xl = p; p += K::D * LC;
        Pre = p; p += K::NPRE * LC;
        red = p; p += 8;
        int *q = (int *)p;
        slot = q; q += OC; lmof = q; q += OC; freeidx = q; q += OC; nfree = q; q += LC; pairptr = q; q += LC + 8;
    }
};

PLBA_HD int symidx(int r, int c, int D) { return r <= c ? r * D - r * (r - 1) / 2 + (c - r) : c * D - c * (c - 1) / 2 + (r - c); }

template <int LT> struct ObsAcc;
template <> struct ObsAcc<LT_POINT> {
    static PLBA_HD const Chunk *chunks(const DevP &P) { return P.chunks_pt; }
    static PLBA_HD const int *kf(const DevP &P) { return P.po_kf; }
    static PLBA_HD const int *lm(const DevP &P) { return P.po_lm; }
    static PLBA_HD const int *ptr(const DevP &P) { return P.pt_ptr; }
    static PLBA_HD const double *om(const DevP &P) { return P.po_om; }
    static PLBA_HD unsigned char *lvl(const DevP &P) { return P.po_lvl; }
    static PLBA_HD double *chi2(const DevP &P) { return P.po_chi2; }
    static PLBA_HD double *state(const DevP &P, int b) { return P.pts[b]; }
};
template <> struct ObsAcc<LT_LINE_ORTH> {
    static PLBA_HD const Chunk *chunks(const DevP &P) { return P.chunks_ls; }
    static PLBA_HD const int *kf(const DevP &P) { return P.lo_kf; }
    static PLBA_HD const int *lm(const DevP &P) { return P.lo_lm; }
    static PLBA_HD const int *ptr(const DevP &P) { return P.ls_ptr; }
    static PLBA_HD const double *om(const DevP &P) { return P.lo_om; }
    static PLBA_HD unsigned char *lvl(const DevP &P) { return P.lo_lvl; }
    static PLBA_HD double *chi2(const DevP &P) { return P.lo_chi2; }
    static PLBA_HD double *state(const DevP &P, int b) { return P.lns[b]; }
};
template <> struct ObsAcc<LT_LINE_END> : ObsAcc<LT_LINE_ORTH> {};

// ---- phase 0: per-landmark precompute into shared memory -------------------------------------------------
template <int PROF, int LT>
PLBA_HD void lm_precompute(const DevP &P, const WinCtrl &ctl, int win, int lm, int l, Smem<PROF, LT> &sm, const double *state) {
    typedef KT<PROF, LT> K;
    if (LT == LT_POINT) {
        for (int i = 0; i < 3; i++) sm.Pre[i * LC + l] = state[(size_t)3 * lm + i];
    } else if (LT == LT_LINE_ORTH) {
        double o[4], pl[6];
        for (int i = 0; i < 4; i++) o[i] = state[(size_t)4 * lm + i];
        if (PROF == PLBA_PROFILE_H_PLK && ctl.iter == 0) { for (int i = 0; i < 6; i++) pl[i] = P.lns_map[(size_t)6 * lm + i]; }   // pass 0 reads map NDw (:1744)
        else orth_to_plk(o, pl);
        LinePre L; line_pre_from_plk(pl, L);
        double *pre = sm.Pre + l;
        for (int i = 0; i < 3; i++) { pre[i * LC] = L.n[i]; pre[(3 + i) * LC] = L.d[i]; pre[(6 + i) * LC] = L.u1[i]; pre[(9 + i) * LC] = L.u2[i]; pre[(12 + i) * LC] = L.u3[i]; }
        pre[15 * LC] = L.w1; pre[16 * LC] = L.w2;
        for (int i = 0; i < 4; i++) pre[(17 + i) * LC] = o[i];
    } else {   // LT_LINE_END: endpoints; inside the loop the reference reads BOTH from offset 3*loc (Q3, :2697-2698)
        const bool q3 = (!P.fixed_quirks && ctl.iter > 0);
        const int l0 = P.win_ls0[win];
        const size_t q3off = (size_t)6 * l0 + (size_t)3 * (lm - l0);
        for (int i = 0; i < 3; i++) {
            sm.Pre[i * LC + l] = q3 ? state[q3off + i] : state[(size_t)6 * lm + i];
            sm.Pre[(3 + i) * LC + l] = q3 ? state[q3off + i] : state[(size_t)6 * lm + 3 + i];
        }
    }
    (void)K::D; (void)win;
}

template <int PROF, int LT>
PLBA_HD void load_line_pre(const Smem<PROF, LT> &sm, int l, LinePre &L) {
    const double *pre = sm.Pre + l;
    for (int i = 0; i < 3; i++) { L.n[i] = pre[i * LC]; L.d[i] = pre[(3 + i) * LC]; L.u1[i] = pre[(6 + i) * LC]; L.u2[i] = pre[(9 + i) * LC]; L.u3[i] = pre[(12 + i) * LC]; }
    L.w1 = pre[15 * LC]; L.w2 = pre[16 * LC];
}

// ---- phase 1: one observation -> scaled Jacobians At (RANK x 6), Bt (RANK x D), et (RANK) in shared memory -----
// Scaling: At = sqrt(w) J_pose, Bt = sqrt(w) J_lm, et = sqrt(w) e  with w = rho1 * Omega (G) or the Cauchy weight (H),
// so that H = J^T (w) J and b = -J^T (w e) are plain products of the stored rows (constructQuadraticForm).
// Returns the robust cost contribution; active = false leaves zero rows.
template <int PROF, int LT>
PLBA_HD double obs_linearize(const DevP &P, const WinCtrl &ctl, int o, int t, int l, Smem<PROF, LT> &sm, double &err_out) {
    typedef KT<PROF, LT> K;
    typedef ObsAcc<LT> OA;
    const int kf = OA::kf(P)[o];
    const int slot = P.kf_slot[kf];
    double A[K::RANK * 6], B[K::RANK * K::D], e[K::RANK];
    double cost = 0.0, wsq = 0.0;
    bool active = true;
    err_out = 0.0;
    if (PROF == PLBA_PROFILE_G) {
        if (ctl.stage == 1 && OA::lvl(P)[o]) active = false;
        if (active) {
            const double *T = P.poseT[ctl.cur] + (size_t)12 * kf;
            if (LT == LT_POINT) {
                const double Pw[3] = {sm.Pre[l], sm.Pre[LC + l], sm.Pre[2 * LC + l]};
                g_point_lin(P.cam, T, Pw, P.po_uv + (size_t)2 * o, e, A, B);
            } else {
                LinePre L; load_line_pre(sm, l, L);
                double head[3], tail[3];
                if (P.fixed_quirks) { for (int i = 0; i < 3; i++) { head[i] = L.n[i]; tail[i] = L.d[i]; } }
                else { for (int i = 0; i < 3; i++) { head[i] = sm.Pre[(17 + i) * LC + l]; tail[i] = sm.Pre[(18 + i) * LC + l]; } }   // Q12
                g_line_lin(P.cam, T, L, head, tail, P.lo_ab + (size_t)4 * o, e, A, B);
            }
            const double om = OA::om(P)[o];
            const double chi2 = om * (e[0] * e[0] + e[1] * e[1]);
            double rho0 = chi2, rho1 = 1.0;
            if (ctl.stage == 0) huber(P.huber_delta, chi2, rho0, rho1);
            cost = rho0;
            wsq = sqrt(rho1 * om);
        }
    } else {
        const bool pass0 = (ctl.iter == 0);
        double r, w, Jp[6], Jl[6];
        if (LT == LT_POINT) {
            const double *T = P.poseT[ctl.cur] + (size_t)12 * kf;
            const double Pw[3] = {sm.Pre[l], sm.Pre[LC + l], sm.Pre[2 * LC + l]};
            h_point(P.cam, T, Pw, P.po_uv + (size_t)2 * o, P.homog_th, Jp, Jl, r, w);
        } else {
            // Q4: inside the loop the line terms keep the MAP pose (:2700, :2010)
            const double *T = (pass0 || P.fixed_quirks) ? P.poseT[ctl.cur] + (size_t)12 * kf : P.kf_Tmap + (size_t)12 * kf;
            if (LT == LT_LINE_END) {
                const double Pw[3] = {sm.Pre[l], sm.Pre[LC + l], sm.Pre[2 * LC + l]}, Qw[3] = {sm.Pre[3 * LC + l], sm.Pre[4 * LC + l], sm.Pre[5 * LC + l]};
                const double th = (pass0 || P.fixed_quirks) ? P.homog_th : 0.0000001;
                h_endline(P.cam, T, Pw, Qw, P.lo_ab + (size_t)4 * o, th, P.fixed_quirks != 0, Jp, Jl, r, w);
            } else {
                LinePre L; load_line_pre(sm, l, L);
                h_plkline(P.cam, T, L, P.lo_ab + (size_t)4 * o, P.homog_th, P.fixed_quirks != 0, Jp, Jl, r, w);
            }
        }
        wsq = sqrt(w);
        for (int c = 0; c < 6; c++) A[c] = Jp[c];
        for (int c = 0; c < K::D; c++) B[c] = Jl[c];
        e[0] = -r;                 // g += J r w  ==  b = -J^T (w e) with e = -r
        cost = w * r * r;
        err_out = cost;
    }
    for (int i = 0; i < K::RANK * 6; i++) sm.A[i * OC + t] = active ? wsq * A[i] : 0.0;
    for (int i = 0; i < K::RANK * K::D; i++) sm.B[i * OC + t] = active ? wsq * B[i] : 0.0;
    for (int i = 0; i < K::RANK; i++) sm.E[i * OC + t] = active ? wsq * e[i] : 0.0;
    sm.slot[t] = active ? slot : -2;     // -1 fixed KF, -2 inactive edge
    return cost;
}

// ---- phase 2: per landmark: H_ll, b_l, damping, inverse ----------------------------------------------------
template <int PROF, int LT>
PLBA_HD void lm_blocks(const DevP &P, const WinCtrl &ctl, int l, int t0, int t1, Smem<PROF, LT> &sm, double *Hfull, double *bl, double &maxd) {
    typedef KT<PROF, LT> K;
    const int D = K::D;
    for (int i = 0; i < D * D; i++) Hfull[i] = 0.0;
    for (int i = 0; i < D; i++) bl[i] = 0.0;
    for (int t = t0; t < t1; t++) {
        if (sm.slot[t] == -2) continue;
        for (int k = 0; k < K::RANK; k++) {
            double b[D];
            for (int c = 0; c < D; c++) b[c] = sm.B[(k * D + c) * OC + t];
            const double ek = sm.E[k * OC + t];
            for (int r = 0; r < D; r++) { bl[r] -= b[r] * ek; for (int c = r; c < D; c++) Hfull[r * D + c] += b[r] * b[c]; }
        }
    }
    maxd = 0.0;
    for (int r = 0; r < D; r++) { if (fabs(Hfull[r * D + r]) > maxd) maxd = fabs(Hfull[r * D + r]); for (int c = 0; c < r; c++) Hfull[r * D + c] = Hfull[c * D + r]; }
    (void)l; (void)P; (void)ctl;
}
template <int PROF, int D>
PLBA_HD void damp_invert(const WinCtrl &ctl, double *Hfull) {
    for (int r = 0; r < D; r++) {
        if (PROF == PLBA_PROFILE_G) Hfull[r * D + r] += ctl.lambda;                 // g2o setLambda: additive
        else Hfull[r * D + r] += ctl.lambda * Hfull[r * D + r];                     // src/mapHandler.cpp:2564-2565
    }
    double keep[D * D];
    for (int i = 0; i < D * D; i++) keep[i] = Hfull[i];
    if (!spd_inverse<D>(Hfull)) { for (int i = 0; i < D * D; i++) Hfull[i] = keep[i]; gj_inverse<D>(Hfull); }
}

// ---------------------------------------------------------------------------------------------------------
// k_assemble: mode 0 = diagonal pass for the initial lambda (computeLambdaInit / Hmax, :2555-2561); mode 1 = full
// ---------------------------------------------------------------------------------------------------------
template <int PROF, int LT>
PLBA_KERNEL void k_assemble(DevP P, int mode) {
    typedef KT<PROF, LT> K;
    typedef ObsAcc<LT> OA;
    const int D = K::D, RANK = K::RANK;
    PLBA_SMEM(raw);
    Smem<PROF, LT> sm(raw);
    const Chunk ch = OA::chunks(P)[PLBA_BID];
    WinCtrl &ctl = P.ctrl[ch.win];
    if (ctl.done) return;
    if (mode == 0 && !ctl.need_init) return;
    const int nlm = ch.lm1 - ch.lm0, nob = ch.ob1 - ch.ob0;
    const int *ptr = OA::ptr(P);
    const double *state = OA::state(P, ctl.cur);

    PHASE_BEGIN
        if (tid < 8) sm.red[tid] = 0.0;
        if (tid < nlm) lm_precompute<PROF, LT>(P, ctl, ch.win, ch.lm0 + tid, tid, sm, state);
    PHASE_END
    PHASE_BEGIN
        double cost = 0.0;
        if (tid < nob) {
            const int o = ch.ob0 + tid;
            const int l = OA::lm(P)[o] - ch.lm0;
            sm.lmof[tid] = l;
            double err;
            cost = obs_linearize<PROF, LT>(P, ctl, o, tid, l, sm, err);
            const int slot = sm.slot[tid];
            // pose diagonal of H_pp (lambda init; multiplicative damping of profile H)
            if (slot >= 0 && (mode == 0 || PROF != PLBA_PROFILE_G)) {
                double *dst = (mode == 0 ? P.hpp_diag_init : P.hpp_diag) + (size_t)6 * slot;
                for (int c = 0; c < 6; c++) { double s = 0; for (int k = 0; k < RANK; k++) { const double a = sm.A[(k * 6 + c) * OC + tid]; s += a * a; } plba_atomic_add(dst + c, s); }
            }
        }
        plba_block_add(&sm.red[0], mode == 1 ? cost : 0.0);
    PHASE_END
    PHASE_BEGIN
        if (tid < nlm) {
            const int lm = ch.lm0 + tid;
            const int t0 = ptr[lm] - ch.ob0, t1 = ptr[lm + 1] - ch.ob0;
            double H[D * D], bl[D], maxd;
            lm_blocks<PROF, LT>(P, ctl, tid, t0, t1, sm, H, bl, maxd);
            if (mode == 0) { plba_atomic_max_pos(&P.accmax[ch.win], maxd); sm.nfree[tid] = 0; }
            else {
                damp_invert<PROF, D>(ctl, H);
                for (int r = 0; r < D; r++) for (int c = r; c < D; c++) sm.Hinv[symidx(r, c, D) * LC + tid] = H[r * D + c];
                for (int r = 0; r < D; r++) { double s = 0; for (int c = 0; c < D; c++) s += H[r * D + c] * bl[c]; sm.V[r * LC + tid] = s; }
                int nf = 0;
                for (int t = t0; t < t1; t++) if (sm.slot[t] >= 0) sm.freeidx[t0 + nf++] = t;
                sm.nfree[tid] = nf;
            }
        }
    PHASE_END
    if (mode == 0) return;
    PHASE_BEGIN
        if (tid == 0) {
            int acc = 0;
            for (int l = 0; l < nlm; l++) { sm.pairptr[l] = acc; acc += sm.nfree[l] * (sm.nfree[l] + 1) / 2; }
            sm.pairptr[nlm] = acc;
            double *accw = P.acc + (size_t)4 * ch.win;
            if (PROF == PLBA_PROFILE_G) plba_atomic_add(&accw[ACC_CHI_LIN], sm.red[0]);
            else plba_atomic_add(LT == LT_POINT ? &accw[ACC_ERR_PT] : &accw[ACC_ERR_LS], sm.red[0]);
        }
    PHASE_END
    PHASE_BEGIN
        const int npairs = sm.pairptr[nlm];
        const int slot0 = P.win_slot0[ch.win];
        const int ld = 6 * P.win_nfree[ch.win];
        double *Sw = P.S + P.win_S_off[ch.win];
        for (int p = tid; p < npairs; p += PLBA_NT) {
            int lo = 0, hi = nlm;                      // largest l with pairptr[l] <= p
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sm.pairptr[mid] <= p) lo = mid; else hi = mid; }
            const int l = lo, nf = sm.nfree[l];
            int q = p - sm.pairptr[l], i = 0;
            while (q >= nf - i) { q -= nf - i; i++; }
            const int j = i + q;
            const int t0 = ptr[ch.lm0 + l] - ch.ob0;
            const int ta = sm.freeidx[t0 + i], tb = sm.freeidx[t0 + j];
            // M = Bt_a Hinv Bt_b^T  (RANK x RANK)
            double Ta[RANK * D], M[RANK * RANK];
            for (int k = 0; k < RANK; k++) for (int c = 0; c < D; c++) {
                double s = 0; for (int m = 0; m < D; m++) s += sm.B[(k * D + m) * OC + ta] * sm.Hinv[symidx(m, c, D) * LC + l];
                Ta[k * D + c] = s;
            }
            for (int k = 0; k < RANK; k++) for (int k2 = 0; k2 < RANK; k2++) {
                double s = 0; for (int m = 0; m < D; m++) s += Ta[k * D + m] * sm.B[(k2 * D + m) * OC + tb];
                M[k * RANK + k2] = s;
            }
            double Aa[RANK * 6], Ab[RANK * 6];
            for (int k = 0; k < RANK * 6; k++) { Aa[k] = sm.A[k * OC + ta]; Ab[k] = sm.A[k * OC + tb]; }
            const int sa = sm.slot[ta] - slot0, sb = sm.slot[tb] - slot0;
            if (i == j) {
                // diagonal: S_aa += At^T (I - M) At ;  g_a += -At^T (et + Bt v)
                double N[RANK * RANK];
                for (int k = 0; k < RANK; k++) for (int k2 = 0; k2 < RANK; k2++) N[k * RANK + k2] = ((k == k2) ? 1.0 : 0.0) - M[k * RANK + k2];
                double NA[RANK * 6];
                for (int k = 0; k < RANK; k++) for (int c = 0; c < 6; c++) { double s = 0; for (int k2 = 0; k2 < RANK; k2++) s += N[k * RANK + k2] * Aa[k2 * 6 + c]; NA[k * 6 + c] = s; }
                for (int r = 0; r < 6; r++) for (int c = r; c < 6; c++) {
                    double s = 0; for (int k = 0; k < RANK; k++) s += Aa[k * 6 + r] * NA[k * 6 + c];
                    plba_atomic_add(&Sw[(size_t)(6 * sa + r) * ld + 6 * sa + c], s);
                }
                double ev[RANK];
                for (int k = 0; k < RANK; k++) { double s = sm.E[k * OC + ta]; for (int m = 0; m < D; m++) s += sm.B[(k * D + m) * OC + ta] * sm.V[m * LC + l]; ev[k] = s; }
                for (int r = 0; r < 6; r++) { double s = 0; for (int k = 0; k < RANK; k++) s += Aa[k * 6 + r] * ev[k]; plba_atomic_add(&P.gs[(size_t)6 * (slot0 + sa) + r], -s); }
            } else {
                double MA[RANK * 6];
                for (int k = 0; k < RANK; k++) for (int c = 0; c < 6; c++) { double s = 0; for (int k2 = 0; k2 < RANK; k2++) s += M[k * RANK + k2] * Ab[k2 * 6 + c]; MA[k * 6 + c] = s; }
                for (int r = 0; r < 6; r++) for (int c = 0; c < 6; c++) {
                    double s = 0; for (int k = 0; k < RANK; k++) s += Aa[k * 6 + r] * MA[k * 6 + c];     // block(a,b)(r,c)
                    if (sa < sb) plba_atomic_add(&Sw[(size_t)(6 * sa + r) * ld + 6 * sb + c], -s);
                    else if (sa > sb) plba_atomic_add(&Sw[(size_t)(6 * sb + c) * ld + 6 * sa + r], -s);
                    else {   // two observations of one landmark in the same KF: blk + blk^T lands on the diagonal block
                        if (r <= c) plba_atomic_add(&Sw[(size_t)(6 * sa + r) * ld + 6 * sa + c], -s);
                        if (c <= r) plba_atomic_add(&Sw[(size_t)(6 * sa + c) * ld + 6 * sa + r], -s);
                    }
                }
            }
        }
    PHASE_END
}

// ---------------------------------------------------------------------------------------------------------
// k_update: re-linearise, back-substitute x_l = H_ll^-1 (b_l - W^T x_p), retract, evaluate the new cost
// ---------------------------------------------------------------------------------------------------------
template <int PROF, int LT>
PLBA_KERNEL void k_update(DevP P) {
    typedef KT<PROF, LT> K;
    typedef ObsAcc<LT> OA;
    const int D = K::D, RANK = K::RANK;
    PLBA_SMEM(raw);
    Smem<PROF, LT> sm(raw);
    const Chunk ch = OA::chunks(P)[PLBA_BID];
    WinCtrl &ctl = P.ctrl[ch.win];
    if (ctl.done) return;
    const int nlm = ch.lm1 - ch.lm0, nob = ch.ob1 - ch.ob0;
    const int *ptr = OA::ptr(P);
    const double *state = OA::state(P, ctl.cur);
    double *state_new = OA::state(P, ctl.cur ^ 1);

    PHASE_BEGIN
        if (tid < 8) sm.red[tid] = 0.0;
        if (tid < nlm) lm_precompute<PROF, LT>(P, ctl, ch.win, ch.lm0 + tid, tid, sm, state);
    PHASE_END
    PHASE_BEGIN
        double sc = 0.0;
        if (tid < nob) {
            const int o = ch.ob0 + tid;
            const int l = OA::lm(P)[o] - ch.lm0;
            sm.lmof[tid] = l;
            double err;
            obs_linearize<PROF, LT>(P, ctl, o, tid, l, sm, err);
            const int slot = sm.slot[tid];
            for (int k = 0; k < RANK; k++) {
                double u = 0.0;
                if (slot >= 0) for (int c = 0; c < 6; c++) u += sm.A[(k * 6 + c) * OC + tid] * P.xp[(size_t)6 * slot + c];
                sm.U[k * OC + tid] = u;
                sc -= u * sm.E[k * OC + tid];          // x_p^T b_p restricted to this edge
            }
            if (slot < 0) sc = 0.0;
        }
        plba_block_add(&sm.red[1], sc);
    PHASE_END
    PHASE_BEGIN
        double sc = 0.0, d2 = 0.0;
        if (tid < nlm) {
            const int lm = ch.lm0 + tid;
            const int t0 = ptr[lm] - ch.ob0, t1 = ptr[lm + 1] - ch.ob0;
            double H[D * D], bl[D], maxd;
            lm_blocks<PROF, LT>(P, ctl, tid, t0, t1, sm, H, bl, maxd);
            damp_invert<PROF, D>(ctl, H);
            double rhs[D];
            for (int c = 0; c < D; c++) rhs[c] = bl[c];
            for (int t = t0; t < t1; t++) {
                if (sm.slot[t] < 0) continue;
                for (int k = 0; k < RANK; k++) { const double u = sm.U[k * OC + t]; for (int c = 0; c < D; c++) rhs[c] -= sm.B[(k * D + c) * OC + t] * u; }
            }
            double xl[D];
            for (int r = 0; r < D; r++) { double s = 0; for (int c = 0; c < D; c++) s += H[r * D + c] * rhs[c]; xl[r] = s; }
            for (int r = 0; r < D; r++) { sc += xl[r] * (ctl.lambda * xl[r] + bl[r]); d2 += xl[r] * xl[r]; }
            // retraction
            double cur[D], nw[D];
            for (int i = 0; i < D; i++) cur[i] = state[(size_t)D * lm + i];
            if (LT == LT_LINE_ORTH) orth_update(cur, xl, nw);                 // updateOrthCoord (a11)
            else for (int i = 0; i < D; i++) nw[i] = cur[i] + xl[i];
            if (PROF == PLBA_PROFILE_G || ctl.apply) for (int i = 0; i < D; i++) state_new[(size_t)D * lm + i] = nw[i];
            for (int i = 0; i < D; i++) sm.Xl[i * LC + tid] = nw[i];
        }
        plba_block_add(&sm.red[1], sc);
        plba_block_add(&sm.red[2], d2);
    PHASE_END
    if (PROF == PLBA_PROFILE_G) {
        // new cost at the trial state (computeActiveErrors + activeRobustChi2 after update)
        PHASE_BEGIN
            if (tid < nlm && LT == LT_LINE_ORTH) {
                double o4[4], pl[6];
                for (int i = 0; i < 4; i++) o4[i] = sm.Xl[i * LC + tid];
                orth_to_plk(o4, pl);
                for (int i = 0; i < 6; i++) sm.Pre[i * LC + tid] = pl[i];
            }
        PHASE_END
        PHASE_BEGIN
            double rho0 = 0.0;
            if (tid < nob && sm.slot[tid] != -2) {
                const int o = ch.ob0 + tid, l = sm.lmof[tid];
                const double *T = P.poseT[ctl.cur ^ 1] + (size_t)12 * OA::kf(P)[o];
                double e[2];
                if (LT == LT_POINT) { const double Pw[3] = {sm.Xl[l], sm.Xl[LC + l], sm.Xl[2 * LC + l]}; double zc; g_point_error(P.cam, T, Pw, P.po_uv + (size_t)2 * o, e, zc); }
                else { const double n[3] = {sm.Pre[l], sm.Pre[LC + l], sm.Pre[2 * LC + l]}, d[3] = {sm.Pre[3 * LC + l], sm.Pre[4 * LC + l], sm.Pre[5 * LC + l]}; g_line_error(P.cam, T, n, d, P.lo_ab + (size_t)4 * o, e); }
                const double chi2 = OA::om(P)[o] * (e[0] * e[0] + e[1] * e[1]);
                OA::chi2(P)[o] = chi2;                       // the cached _error of the edge (e->chi2())
                double rho1;
                rho0 = chi2;
                if (ctl.stage == 0) huber(P.huber_delta, chi2, rho0, rho1);
            }
            plba_block_add(&sm.red[0], rho0);
        PHASE_END
    }
    PHASE_BEGIN
        if (tid == 0) {
            double *acc = P.accB + (size_t)4 * ch.win;
            if (PROF == PLBA_PROFILE_G) plba_atomic_add(&acc[ACC_CHI_NEW], sm.red[0]);
            plba_atomic_add(&acc[ACC_SCALE], sm.red[1]);
            plba_atomic_add(&acc[ACC_DX2], sm.red[2]);
        }
    PHASE_END
}

// ---- per-window / per-keyframe small kernels ---------------------------------------------------------------
PLBA_KERNEL void k_lambda_init(DevP P) {
    PHASE_BEGIN
        const int w = PLBA_BID * PLBA_NT + tid;
        if (w < P.n_win) {
            WinCtrl &c = P.ctrl[w];
            if (!c.done && c.need_init) {
                double m = P.accmax[w];
                const int s0 = P.win_slot0[w], nf = P.win_nfree[w];
                for (int i = 0; i < 6 * nf; i++) { const double h = fabs(P.hpp_diag_init[(size_t)6 * s0 + i]); if (h > m) m = h; }
                if (P.profile == PLBA_PROFILE_G) { c.lambda = P.lm_tau * m; c.ni = 2.0; }
                else c.lambda = P.lambda_lba_lm * m;
                c.need_init = 0;
                plba_atomic_add_i(&P.counters[CNT_NEED_INIT], -1);
            }
            if (c.do_gate) { c.do_gate = 0; plba_atomic_add_i(&P.counters[CNT_GATE], -1); }
        }
    PHASE_END
}

// profile H, between assembly and solve: normalise err, stop tests, accept decision (src/mapHandler.cpp:2796-2814)
PLBA_KERNEL void k_control_h_pre(DevP P) {
    PHASE_BEGIN
        const int w = PLBA_BID * PLBA_NT + tid;
        if (w < P.n_win) {
            WinCtrl &c = P.ctrl[w];
            if (!c.done) {
                const double err_pt = P.acc[(size_t)4 * w + ACC_ERR_PT], err_ls = P.acc[(size_t)4 * w + ACC_ERR_LS];
                double err = err_pt + err_ls;
                const double zero = 0.0;
                const bool div0 = !P.fixed_quirks && (c.iter == 0 || P.profile == PLBA_PROFILE_H_PLK);    // Q1
                err = div0 ? err / zero : err / (double)(c.n_lm_pt + c.n_lm_ls);
                c.chi_cur = err;
                c.apply = 1; c.stop_code = 0;
                if (c.iter > 0) {
                    if (fabs(err - c.err_prev) < P.min_error_change || err < P.min_error) {
                        c.stop_code = 1; c.done = 1; c.apply = 0;
                        plba_atomic_add_i(&P.counters[CNT_DONE], 1);
                        if (c.n_trace < P.trace_cap) {
                            plba_trace_rec &t = P.trace[(size_t)w * P.trace_cap + c.n_trace];
                            t.window = w; t.stage = 0; t.iter = c.iter; t.trial = 0; t.accepted = 0; t.stop = 1;
                            t.chi = err; t.chi_new = 0; t.rho = 0; t.lambda = c.lambda; t.scale = 0; t.dx_norm = 0; t.err_pt = err_pt; t.err_ls = err_ls;
                        }
                        c.n_trace++;
                    } else c.apply = (err > c.err_prev) ? 0 : 1;
                }
            }
        }
    PHASE_END
}

PLBA_KERNEL void k_pose_update(DevP P) {
    PHASE_BEGIN
        const int s = PLBA_BID * PLBA_NT + tid;
        if (s < P.n_free) {
            const int kf = P.slot_kf[s];
            WinCtrl &c = P.ctrl[P.kf_win[kf]];
            if (!c.done) {
                double x[6], d2 = 0.0;
                for (int i = 0; i < 6; i++) { x[i] = P.xp[(size_t)6 * s + i]; d2 += x[i] * x[i]; }
                double Tn[12];
                if (P.profile == PLBA_PROFILE_G) {
                    pose_oplus_g(P.poseT[c.cur] + (size_t)12 * kf, x, Tn);
                    for (int i = 0; i < 12; i++) P.poseT[c.cur ^ 1][(size_t)12 * kf + i] = Tn[i];
                    plba_atomic_add(&c.scale_pose, c.lambda * d2);
                } else {
                    if (c.apply) {   // X <- log( exp(X) * exp(DX)^-1 )   (src/mapHandler.cpp:2571-2578)
                        double Tp[12], Td[12], Tdi[12], Tc[12], xn[6];
                        exp_se3(P.Xkf[c.cur] + (size_t)6 * s, Tp);
                        exp_se3(x, Td); inv_se3(Td, Tdi); mul_se3(Tp, Tdi, Tc);
                        log_se3(Tc, xn);
                        for (int i = 0; i < 6; i++) P.Xkf[c.cur ^ 1][(size_t)6 * s + i] = xn[i];
                        exp_se3(xn, Tp); inv_se3(Tp, Tn);      // T_iw used by the next linearisation (:2617-2621)
                        for (int i = 0; i < 12; i++) P.poseT[c.cur ^ 1][(size_t)12 * kf + i] = Tn[i];
                    }
                    plba_atomic_add(&c.dx2_pose, d2);
                }
            }
        }
    PHASE_END
}

PLBA_HD void write_trace(const DevP &P, int w, WinCtrl &c, int accepted, int stop, double chi, double chi_new, double rho, double lambda, double scale, double dx) {
    const double *acc = P.acc + (size_t)4 * w;
    if (c.n_trace < P.trace_cap) {
        plba_trace_rec &t = P.trace[(size_t)w * P.trace_cap + c.n_trace];
        t.window = w; t.stage = c.stage; t.iter = c.iter; t.trial = c.trial; t.accepted = accepted; t.stop = stop;
        t.chi = chi; t.chi_new = chi_new; t.rho = rho; t.lambda = lambda; t.scale = scale; t.dx_norm = dx; t.err_pt = acc[ACC_ERR_PT]; t.err_ls = acc[ACC_ERR_LS];
    }
    c.n_trace++;
}

// The LM controller, one thread per window: g2o OptimizationAlgorithmLevenberg::solve + SparseOptimizer::optimize
//  for profile G, src/mapHandler.cpp:2808-2837 for profile H.
PLBA_KERNEL void k_control(DevP P) {
    PHASE_BEGIN
        const int w = PLBA_BID * PLBA_NT + tid;
        if (w < P.n_win) {
            WinCtrl &c = P.ctrl[w];
            if (!c.done) {
                c.n_trials++;
                plba_atomic_add_i(&P.counters[CNT_TRIALS], 1);
                const double *acc = P.acc + (size_t)4 * w, *accB = P.accB + (size_t)4 * w;
                if (P.profile == PLBA_PROFILE_G) {
                    if (c.trial == 0) c.chi_cur = acc[ACC_CHI_LIN];                // currentChi = activeRobustChi2()
                    const double tempChi = c.solve_fail ? 1.7976931348623157e308 : accB[ACC_CHI_NEW];
                    const double scale = (accB[ACC_SCALE] + c.scale_pose) + 1e-3;
                    const double rho = (c.chi_cur - tempChi) / scale;
                    const double lam_used = c.lambda;
                    int accepted = 0;
                    const double chi_before = c.chi_cur;
                    if (rho > 0 && plba_isfinite(tempChi)) {
                        const double t = 2 * rho - 1;
                        double alpha = 1. - t * t * t;
                        alpha = alpha < 2. / 3. ? alpha : 2. / 3.;
                        const double sf = alpha > 1. / 3. ? alpha : 1. / 3.;
                        c.lambda *= sf; c.ni = 2; c.chi_cur = tempChi; c.cur ^= 1; accepted = 1;
                    } else { c.lambda *= c.ni; c.ni *= 2; }
                    const int qmax = c.trial + 1;
                    const bool again = (rho < 0 && qmax < P.lm_max_trials);
                    const bool terminate = !again && (qmax == P.lm_max_trials || rho == 0);
                    write_trace(P, w, c, accepted, terminate ? 1 : 0, chi_before, tempChi, rho, lam_used, scale, 0.0);
                    if (again) c.trial = qmax;
                    else {
                        c.trial = 0; c.iter++;
                        const int n_outer = c.stage == 0 ? P.iters_stage1 : P.iters_stage2;
                        if (terminate || c.iter >= n_outer) {
                            c.stage++; c.iter = 0;
                            if (c.stage == 1) { c.do_gate = 1; plba_atomic_add_i(&P.counters[CNT_GATE], 1); }
                            if (c.stage == 1 && P.iters_stage2 > 0) { c.need_init = 1; plba_atomic_add_i(&P.counters[CNT_NEED_INIT], 1); }
                            else { c.done = 1; plba_atomic_add_i(&P.counters[CNT_DONE], 1); }
                        }
                    }
                } else {
                    const double dx = sqrt(accB[ACC_DX2] + c.dx2_pose);
                    const double lam_used = c.lambda;
                    int stop = 0;
                    if (c.iter == 0) { c.cur ^= 1; }
                    else {
                        if (c.apply) { c.lambda *= P.lambda_lba_k; c.cur ^= 1; } else c.lambda /= P.lambda_lba_k;
                        if (dx < P.min_error_change) stop = 2;
                    }
                    write_trace(P, w, c, c.apply, stop, c.chi_cur, 0.0, 0.0, lam_used, 0.0, dx);
                    c.err_prev = c.chi_cur;
                    c.iter++;
                    if (stop || c.iter >= P.max_iters_lba) { c.done = 1; plba_atomic_add_i(&P.counters[CNT_DONE], 1); }
                }
                c.scale_pose = 0; c.dx2_pose = 0; c.solve_fail = 0;
            }
        }
    PHASE_END
}

// chi2 gate between the two g2o stages (src/mapHandler.cpp:6125-6147): level 1 <=> chi2 > 5.991 (or z_c <= 0 for points)
template <int LT>
PLBA_KERNEL void k_gate(DevP P, int n_obs) {
    typedef ObsAcc<LT> OA;
    PHASE_BEGIN
        const int o = PLBA_BID * PLBA_NT + tid;
        if (o < n_obs) {
            const int kf = OA::kf(P)[o];
            const WinCtrl &c = P.ctrl[P.kf_win[kf]];
            if (c.do_gate) {
                bool bad = OA::chi2(P)[o] > P.chi2_gate;
                if (LT == LT_POINT) {
                    const double *T = P.poseT[c.cur] + (size_t)12 * kf;
                    const double *Pw = P.pts[c.cur] + (size_t)3 * OA::lm(P)[o];
                    const double z = T[8] * Pw[0] + T[9] * Pw[1] + T[10] * Pw[2] + T[11];
                    if (!(z > 0.0)) bad = true;
                }
                OA::lvl(P)[o] = bad ? 1 : 0;
            }
        }
    PHASE_END
}

// final per-observation test (src/mapHandler.cpp:6156-6161, 6224-6230): level-1 edges are re-evaluated at the final estimate
template <int LT>
PLBA_KERNEL void k_final(DevP P, int n_obs, unsigned char *flags_out) {
    typedef ObsAcc<LT> OA;
    PHASE_BEGIN
        const int o = PLBA_BID * PLBA_NT + tid;
        if (o < n_obs) {
            const int kf = OA::kf(P)[o];
            const WinCtrl &c = P.ctrl[P.kf_win[kf]];
            const double *T = P.poseT[c.cur] + (size_t)12 * kf;
            const int lm = OA::lm(P)[o];
            const unsigned char lvl = OA::lvl(P)[o];
            double chi2 = OA::chi2(P)[o];
            unsigned char f = lvl ? PLBA_OBS_LEVEL1 : 0;
            if (LT == LT_POINT) {
                double e[2], zc;
                g_point_error(P.cam, T, P.pts[c.cur] + (size_t)3 * lm, P.po_uv + (size_t)2 * o, e, zc);
                if (lvl) chi2 = OA::om(P)[o] * (e[0] * e[0] + e[1] * e[1]);
                if (!(zc > 0.0)) f |= PLBA_OBS_NEGDEPTH | PLBA_OBS_BAD;
            } else if (lvl) {
                double pl[6], e[2];
                orth_to_plk(P.lns[c.cur] + (size_t)4 * lm, pl);
                g_line_error(P.cam, T, pl, pl + 3, P.lo_ab + (size_t)4 * o, e);
                chi2 = OA::om(P)[o] * (e[0] * e[0] + e[1] * e[1]);
            }
            if (chi2 > P.chi2_gate) f |= PLBA_OBS_BAD;
            OA::chi2(P)[o] = chi2;
            flags_out[o] = f;
        }
    PHASE_END
}

// write-back helpers: T_kf_w = estimate^-1 (:6302) / expmap_se3(X) (:2851-2852); NDw = changeOrthToPluker(orth) (:6318)
PLBA_KERNEL void k_export_poses(DevP P, double *T_wc_out, double *x_out) {
    PHASE_BEGIN
        const int kf = PLBA_BID * PLBA_NT + tid;
        if (kf < P.n_kf) {
            const int s = P.kf_slot[kf];
            const WinCtrl &c = P.ctrl[P.kf_win[kf]];
            double T[12];
            if (s < 0) inv_se3(P.kf_Tmap + (size_t)12 * kf, T);          // fixed observer: unchanged (plba_solve copies the caller's rows through)
            else if (P.profile == PLBA_PROFILE_G) inv_se3(P.poseT[c.cur] + (size_t)12 * kf, T);
            else { exp_se3(P.Xkf[c.cur] + (size_t)6 * s, T); for (int i = 0; i < 6; i++) x_out[(size_t)6 * s + i] = P.Xkf[c.cur][(size_t)6 * s + i]; }
            for (int i = 0; i < 12; i++) T_wc_out[(size_t)12 * kf + i] = T[i];
        }
    PHASE_END
}
// landmarks of every window gathered from that window's current buffer
PLBA_KERNEL void k_export_landmarks(DevP P, const int *pt_win, const int *ls_win, double *pt_out, double *ls_out, double *plk_out, int ls_dim, const double *orth0, int q9) {
    PHASE_BEGIN
        const int i = PLBA_BID * PLBA_NT + tid;
        if (i < P.n_pt) { const int b = P.ctrl[pt_win[i]].cur; for (int k = 0; k < 3; k++) pt_out[(size_t)3 * i + k] = P.pts[b][(size_t)3 * i + k]; }
        if (i < P.n_ls) {
            const int b = P.ctrl[ls_win[i]].cur;
            double v[6];
            for (int k = 0; k < ls_dim; k++) { v[k] = P.lns[b][(size_t)ls_dim * i + k]; ls_out[(size_t)ls_dim * i + k] = v[k]; }
            if (ls_dim == 4) {
                double pl[6];
                if (q9) { double dx[4]; for (int k = 0; k < 4; k++) dx[k] = v[k] - orth0[(size_t)4 * i + k]; orth_to_plk(dx, pl); }   // Q9 (:2194)
                else orth_to_plk(v, pl);
                for (int k = 0; k < 6; k++) plk_out[(size_t)6 * i + k] = pl[k];
            }
        }
    PHASE_END
}

}  // namespace plba
