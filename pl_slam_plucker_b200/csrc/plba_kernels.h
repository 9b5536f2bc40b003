// LBA kernels for sm_100a (FP64).
//
// Data layout (HBM, SoA, 256-byte aligned arrays carved from one arena): observations are stored landmark-major
// (the reference's own order, src/mapHandler.cpp:1424-1447) after the landmarks of each window have been sorted by
// their keyframe sequence ("signature"), so that runs of landmarks seen by exactly the same keyframes are adjacent.
// A CHUNK is <= OC observations / <= LC consecutive landmarks of one window; a SEGMENT is a run of <= SEG_MAX
// landmarks with identical signature inside a chunk.
//
// k_assemble (persistent grid, one chunk at a time per CTA, everything stays in shared memory):
//   phase 0  per landmark   : landmark-only quantities (orth -> Plücker, U, W) hoisted out of the per-edge path
//   phase 1  per observation: residual + Jacobians + robust weight  (a12-a14, a17, a18 of SURVEY.md §8a), 128-bit loads
//   phase 2  per landmark   : H_ll, b_l, damping, H_ll^-1                      (subsystem 2)
//   phase 3  per (segment, pose pair): Schur update of the reduced camera system S, g (subsystem 3): the 6x6 block is
//            accumulated IN REGISTERS over the landmarks of the segment and leaves the SM as one red.global.add.f64
//            per entry (FP64 atomics exist natively only for global memory on sm_100a; shared-memory ones are CAS loops).
// k_update re-linearises the same chunk, back-substitutes, retracts and evaluates the new cost (subsystem 4); its last
// CTA runs the LM controller and steers the WHILE / IF nodes of the CUDA graph that holds the whole LM loop.
// W (H_pl) blocks are never materialised in HBM; every input is read once per kernel.
//
// All kernels are written as PHASE blocks (plba_port.h) so that the control flow can be debugged on a CPU.
#pragma once
#include "plba_math.h"
#include "../../include/plba.h"

namespace plba {

enum { LT_POINT = 0, LT_LINE_ORTH = 1, LT_LINE_END = 2 };
enum { LPRE_N = 17 };   // n(3) d(3) u1(3) u2(3) u3(3) w1 w2
#ifndef PLBA_OC
#define PLBA_OC 256
#endif
#ifndef PLBA_SEG_MAX
#define PLBA_SEG_MAX 8
#endif
enum { OC = PLBA_OC /* observations (= threads) per chunk */, LC = PLBA_OC / 2 /* landmarks per chunk */, SEG_MAX = PLBA_SEG_MAX /* landmarks per segment */ };
enum { KF_FUSE_CONTROL = 1, KF_IN_GRAPH = 2, KF_WAIT_SOLVE = 4 };   // kernel flags (KF_WAIT_SOLVE: the update kernel runs concurrently with k_solve_small, see k_update)
#ifndef PLBA_CTAS_PER_SM
#define PLBA_CTAS_PER_SM 2   // measured: 2 resident CTAs (128 registers, small spills) beat 1 CTA with 254 registers on C2, C4 and C5
#endif

struct Chunk { int lm0, lm1, ob0, ob1, win, seg0, seg1, pad; };
// run of landmarks with identical keyframe sequence: lm0 = first landmark (internal order), nobs observations each,
// nfree of them in free keyframes at positions freepos[fp0 .. fp0+nfree); task0 = first (pose pair) task of the run in its chunk
struct Seg { int lm0, n_lm, nobs, nfree, task0, fp0, dtask0, pad1; };   // task0 / dtask0: first off-diagonal / diagonal task

// warp-path work item (plba_warp.h): a slice of a run of landmarks with identical keyframe sequence: n_lm consecutive landmarks
// (internal order) from lm0, k observations each starting at ob0; nfree of the k track positions are free keyframes:
// positions freepos[fp0 .. fp0 + nfree)
struct WItem { int lm0, n_lm, ob0, k, win, nfree, fp0, pad; };

struct WinCtrl {
    int cur, stage, iter, trial, need_init, do_gate, done, n_trace, solve_fail, apply, stop_code, numeric;   // numeric: low 16 bits = consecutive failed solves, bit 30 = PLBA_E_NUMERIC
    int n_lm_pt, n_lm_ls;            // landmark counts of the window (profile H normalisation)
    double lambda, ni, chi_cur, err_prev;
    double scale_pose, dx2_pose;     // pose part of computeScale() / ||DX||^2 (replicated on every rank, never all-reduced)
    int n_trials, n_obs;             // n_obs: observations of the window (GBA normalisation in FIXED mode)
};
// per-window accumulators that ARE summed over ranks (landmark-sharded multi-GPU):
// assemble-phase sums live in P.acc[4*w + ...], update-phase sums in P.accB[4*w + ...] (two all-reduce ranges)
enum { ACC_CHI_LIN = 0, ACC_ERR_PT = 1, ACC_ERR_LS = 2, ACC_CHI_NEW = 0, ACC_SCALE = 1, ACC_DX2 = 2, ACC_N = 8 };
enum { CNT_DONE = 0, CNT_NEED_INIT = 1, CNT_GATE = 2, CNT_TRIALS = 3, CNT_TICKET = 4, CNT_ROUNDS = 5, CNT_PREPS = 6,
       CNT_WORK_PT = 8, CNT_WORK_LS = 9, CNT_WTICKET = 10 /* warp path: next work item of each landmark class, CTA exit ticket */,
       CNT_SOLVE_DONE = 12 /* k_solve_small (single window) has published x_p and the trial poses: the update kernel that runs BESIDE it waits for this */,
       CNT_KLAUNCH = 11 /* kernels of the library that have started since the last reset: counted ON THE DEVICE (PLBA_PARAMS), so that launches inside the LM-loop graph are counted, not inferred */, CNT_N = 16 };

struct DevP {
    Cam cam;
    int profile, fixed_quirks;
    int n_win, n_kf, n_free, n_pt, n_ls, n_pobs, n_lobs;
    int n_chunks_pt, n_chunks_ls, max_rounds, gba;   // gba: Global BA shell around profile H_END (PLBA_SHELL_GBA)
    int iters_stage1, iters_stage2, lm_max_trials, max_iters_lba;
    double huber_delta, chi2_gate, homog_th, min_error, min_error_change, lm_tau, lambda_lba_lm, lambda_lba_k;
    // keyframes
    const int *kf_slot, *kf_win, *slot_kf;
    const double *kf_Tmap;            // [n_kf][12]  T_cw of the map pose (fixed KFs, profile-H pass 0, Q4)
    double *poseT[2];                 // [n_kf][12]  T_cw estimate, double buffered (push / pop)
    double *Xkf[2];                   // [n_free][6] profile H: log of T_kf_w
    const double *X0;                 // [n_free][6] initial X (reset)
    const int *win_slot0, *win_nfree; // per window: first global slot, number of free KFs
    const int *win_ls0;               // per window: first global line index (Q3 indexing is window-relative)
    const long long *win_S_off;       // per window: offset (doubles) of its reduced camera system inside S: dense (6 nf)^2, or node form [D | U] (see s_block)
    const int *win_bcr_bs, *win_bcr_N; // per window: keyframes per node / nodes of the block cyclic reduction (0 = dense storage)
    // landmarks (double buffered) in internal (signature-sorted) order
    double *pts[2], *lns[2];
    double *lpre[2];                  // [n_ls][LPRE_N] orthonormal lines: Plücker vector, U, W at the state of each buffer (warp kernels)
    const double *pts0, *lns0;        // initial values (reset, Q9)
    const double *lns_map;            // [n_ls][6] map Plücker (H_PLK pass 0)
    const int *pt_ptr, *ls_ptr;       // CSR landmark -> observations
    const int *pt_win, *ls_win;
    // observations (SoA)
    const int *po_kf, *lo_kf, *po_lm, *lo_lm;
    const double *po_uv, *lo_ab, *po_om, *lo_om;
    // the caller's observation arrays as staged by plba_upload (caller's order, keyframe indices already global) and, per landmark, the first
    // staged index of its run: k_reset(gather = 1) re-orders them into the arrays above once per upload
    const int *r_po_kf, *r_lo_kf, *pt_src, *ls_src;
    const double *r_po_uv, *r_lo_ab, *r_po_s2, *r_lo_s2;
    unsigned char *po_lvl, *lo_lvl;
    double *po_chi2, *lo_chi2;
    const Chunk *chunks_pt, *chunks_ls;
    const Seg *segs_pt, *segs_ls;
    const WItem *witems_pt, *witems_ls; int n_witems_pt, n_witems_ls;
    const int *freepos_pt, *freepos_ls;
    // linear system
    double *S, *gs, *xp, *hpp_diag, *hpp_diag_init;
    double *acc, *accB;               // [n_win][4] each: sum-reduced accumulators (tail of the reduced-system buffer)
    double *accmax;                   // [n_win]         largest landmark diagonal of this rank (lambda init)
    double *maxslots;                 // [16][n_win]     sharded path: accmax of every rank, one slot each, so that the maximum travels inside a SUM exchange
    int n_ranks, rank;
    WinCtrl *ctrl;
    const WinCtrl *ctrl0;             // initial controller state (reset)
    plba_trace_rec *trace; int trace_cap, solve_nf_max;   // solve_nf_max: most free keyframes of any window of the upload (thread-group size of k_solve_small)
    int *counters;                    // [CNT_N]
    long long S_clear_doubles;        // small-window path: doubles of S the update kernel clears (the solver consumed them); 0 = the solver path clears S itself
    unsigned long long cond_while, cond_prep, cond_prep2;   // cudaGraphConditionalHandle of the LM-loop graph (prep2: the IF node of the second trial of an iteration)
};

#ifndef PLBA_HOST_EMU
__shared__ DevP plba_params_smem;     // per-CTA copy of the kernel parameters (PLBA_PARAMS)
#endif

template <int PROF, int LT> struct KT;
template <> struct KT<PLBA_PROFILE_G, LT_POINT>     { enum { RANK = 2, D = 3, NPRE = 3 }; };
template <> struct KT<PLBA_PROFILE_G, LT_LINE_ORTH> { enum { RANK = 2, D = 4, NPRE = 21 }; };
template <> struct KT<PLBA_PROFILE_H_END, LT_POINT>    { enum { RANK = 1, D = 3, NPRE = 3 }; };
template <> struct KT<PLBA_PROFILE_H_END, LT_LINE_END> { enum { RANK = 1, D = 6, NPRE = 6 }; };
template <> struct KT<PLBA_PROFILE_H_PLK, LT_POINT>     { enum { RANK = 1, D = 3, NPRE = 3 }; };
template <> struct KT<PLBA_PROFILE_H_PLK, LT_LINE_ORTH> { enum { RANK = 1, D = 4, NPRE = 21 }; };
template <int PROF> struct LineOf { enum { LT = (PROF == PLBA_PROFILE_H_END) ? LT_LINE_END : LT_LINE_ORTH }; };

// shared-memory carve-up of a chunk (SoA: [component][slot] so that thread-per-slot accesses are conflict free)
template <int PROF, int LT>
struct Smem {
    typedef KT<PROF, LT> K;
    enum { NSYM = K::D * (K::D + 1) / 2 };
    double *A, *B, *E, *U, *TA, *Hinv, *V, *Pre, *Xl, *red;
    int *slot, *lmof, *act, *freepos, *seg_lm0, *seg_nlm, *seg_nobs, *seg_nfree, *seg_task0, *seg_dtask0, *seg_fp0, *lm_t0;
    static size_t bytes() {
        return sizeof(double) * ((size_t)(K::RANK * 6 + 2 * K::RANK * K::D + 2 * K::RANK) * OC + (size_t)(NSYM + 2 * K::D + K::NPRE) * LC + 8)
             + sizeof(int) * (4 * OC + 8 * LC + 48);
    }
    PLBA_HD explicit Smem(unsigned char *raw) {
        double *p = (double *)raw;
        A = p; p += K::RANK * 6 * OC;
        B = p; p += K::RANK * K::D * OC;
        E = p; p += K::RANK * OC;
        U = p; p += K::RANK * OC;
        TA = p; p += K::RANK * K::D * OC;
        Hinv = p; p += NSYM * LC;
        V = p; p += K::D * LC;
        Xl = p; p += K::D * LC;
        Pre = p; p += K::NPRE * LC;
        red = p; p += 8;
        int *q = (int *)p;
        slot = q; q += OC; lmof = q; q += OC; act = q; q += OC; freepos = q; q += OC;
        seg_lm0 = q; q += LC; seg_nlm = q; q += LC; seg_nobs = q; q += LC; seg_nfree = q; q += LC; seg_fp0 = q; q += LC; seg_task0 = q; q += LC + 16; seg_dtask0 = q; q += LC + 16; lm_t0 = q; q += LC + 16;
    }
};
template <int PROF> struct SmemMax {
    static size_t bytes() {
        const size_t a = Smem<PROF, LT_POINT>::bytes(), b = Smem<PROF, LineOf<PROF>::LT>::bytes();
        return a > b ? a : b;
    }
};

// Where the 6x6 block (ra, cb), ra <= cb (free-keyframe positions inside one window), of the reduced camera system lives: address of
// its entry (0, 0) and the strides of its rows and columns.
//   dense windows (one-CTA solver, dense DMMA Cholesky): row-major (6 nf)^2, upper block triangle;
//   block-banded large windows (block cyclic reduction): the assembly kernels accumulate straight into the solver's NODE form —
//     nodes of bs keyframes (m = 6 bs unknowns), D_i = upper triangle of the node's diagonal block (row stride m), U_j = the block between
//     node j - 1 (rows) and node j (columns).  No dense S exists for such a window (config 5: 35 MB instead of 1.15 GB) and no gather
//     kernel runs before the solve (round 1: 0.35 ms of a 1.77 ms solve).  The half bandwidth is <= bs, so cb / bs - ra / bs <= 1.
struct SBlk { double *p; long long sr, sc; };
struct SWin { double *Sw; long long ld; int bs, N; };      // one window's storage, read ONCE per chunk / work item (three dependent global loads)
PLBA_HD SWin s_window(const DevP &P, int win) {
    SWin w; w.Sw = P.S + P.win_S_off[win]; w.bs = P.win_bcr_bs[win]; w.N = P.win_bcr_N[win]; w.ld = 6 * (long long)P.win_nfree[win]; return w;
}
PLBA_HD SBlk s_block(const SWin &w, int ra, int cb) {
    SBlk b;
    if (w.bs == 0) { b.p = w.Sw + (size_t)(6 * ra) * w.ld + 6 * cb; b.sr = w.ld; b.sc = 1; return b; }
    const int bs = w.bs, m = 6 * bs, i = ra / bs, j = cb / bs, a = ra - i * bs, c = cb - j * bs;
    if (i == j) { b.p = w.Sw + (size_t)i * m * m + (size_t)(6 * a) * m + 6 * c; b.sr = m; b.sc = 1; }       // D_i[(6a + row) m + 6c + col]: UPPER triangle, the same strides as the dense storage (the two column halves of a block share cache lines: half the RED wavefronts of a transposed layout)
    else { b.p = w.Sw + (size_t)w.N * m * m + (size_t)j * m * m + (size_t)(6 * a) * m + 6 * c; b.sr = m; b.sc = 1; }   // U_j[(6a + row) m + 6c + col]
    return b;
}

PLBA_HD int symidx(int r, int c, int D) { return r <= c ? r * D - r * (r - 1) / 2 + (c - r) : c * D - c * (c - 1) / 2 + (r - c); }

template <int LT> struct ObsAcc;
template <> struct ObsAcc<LT_POINT> {
    static PLBA_HD const Seg *segs(const DevP &P) { return P.segs_pt; }
    static PLBA_HD const int *freepos(const DevP &P) { return P.freepos_pt; }
    static PLBA_HD const int *kf(const DevP &P) { return P.po_kf; }
    static PLBA_HD const int *lm(const DevP &P) { return P.po_lm; }
    static PLBA_HD const int *ptr(const DevP &P) { return P.pt_ptr; }
    static PLBA_HD const double *om(const DevP &P) { return P.po_om; }
    static PLBA_HD unsigned char *lvl(const DevP &P) { return P.po_lvl; }
    static PLBA_HD double *chi2(const DevP &P) { return P.po_chi2; }
    static PLBA_HD double *state(const DevP &P, int b) { return P.pts[b]; }
};
template <> struct ObsAcc<LT_LINE_ORTH> {
    static PLBA_HD const Seg *segs(const DevP &P) { return P.segs_ls; }
    static PLBA_HD const int *freepos(const DevP &P) { return P.freepos_ls; }
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
    if (LT == LT_POINT) {
        for (int i = 0; i < 3; i++) sm.Pre[i * LC + l] = state[(size_t)3 * lm + i];
    } else if (LT == LT_LINE_ORTH) {
        double o[4];
        for (int i = 0; i < 4; i++) o[i] = state[(size_t)4 * lm + i];
        LinePre L;
        if (PROF == PLBA_PROFILE_H_PLK && ctl.iter == 0) {      // pass 0 reads map NDw (:1744)
            double pl[6];
            for (int i = 0; i < 6; i++) pl[i] = P.lns_map[(size_t)6 * lm + i];
            line_pre_from_plk(pl, L);
        } else {
            // Plücker vector, U and W at the current state: written once per landmark and state (k_reset, the update kernels), not per kernel
            const double *c = P.lpre[ctl.cur] + (size_t)LPRE_N * lm;
            for (int i = 0; i < 3; i++) { L.n[i] = c[i]; L.d[i] = c[3 + i]; L.u1[i] = c[6 + i]; L.u2[i] = c[9 + i]; L.u3[i] = c[12 + i]; }
            L.w1 = c[15]; L.w2 = c[16];
        }
        double *pre = sm.Pre + l;
        for (int i = 0; i < 3; i++) { pre[i * LC] = L.n[i]; pre[(3 + i) * LC] = L.d[i]; pre[(6 + i) * LC] = L.u1[i]; pre[(9 + i) * LC] = L.u2[i]; pre[(12 + i) * LC] = L.u3[i]; }
        pre[15 * LC] = L.w1; pre[16 * LC] = L.w2;
        for (int i = 0; i < 4; i++) pre[(17 + i) * LC] = o[i];
    } else {   // LT_LINE_END: endpoints; inside the loop the reference reads BOTH from offset 3*loc (Q3, :2697-2698)
        const bool q3 = (!P.fixed_quirks && ctl.iter > 0);
        const int l0 = P.win_ls0[win];
        const size_t q3off = (size_t)6 * l0 + (size_t)3 * (lm - l0);      // endpoint lines are never re-ordered (see plba_upload)
        for (int i = 0; i < 3; i++) {
            sm.Pre[i * LC + l] = q3 ? state[q3off + i] : state[(size_t)6 * lm + i];
            sm.Pre[(3 + i) * LC + l] = q3 ? state[q3off + i] : state[(size_t)6 * lm + 3 + i];
        }
    }
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
// Returns the robust cost contribution.  Edges gated out of the active set (level 1 in stage 2) leave zero rows.
template <int PROF, int LT>
PLBA_HD double obs_linearize(const DevP &P, const WinCtrl &ctl, int o, int t, int l, Smem<PROF, LT> &sm) {
    typedef KT<PROF, LT> K;
    typedef ObsAcc<LT> OA;
    const int kf = OA::kf(P)[o];
    const int slot = P.kf_slot[kf];
    double A[K::RANK * 6], B[K::RANK * K::D], e[K::RANK];
    double cost = 0.0, wsq = 0.0;
    bool active = true;
    if (PROF == PLBA_PROFILE_G) {
        if (ctl.stage == 1 && OA::lvl(P)[o]) active = false;
        if (active) {
            const double *T = P.poseT[ctl.cur] + (size_t)12 * kf;
            if (LT == LT_POINT) {
                const double Pw[3] = {sm.Pre[l], sm.Pre[LC + l], sm.Pre[2 * LC + l]};
                const plba_d2 uv = *(const plba_d2 *)(P.po_uv + (size_t)2 * o);
                const double uv2[2] = {uv.x, uv.y};
                g_point_lin(P.cam, T, Pw, uv2, e, A, B);
            } else {
                LinePre L; load_line_pre(sm, l, L);
                double head[3], tail[3];
                if (P.fixed_quirks) { for (int i = 0; i < 3; i++) { head[i] = L.n[i]; tail[i] = L.d[i]; } }
                else { for (int i = 0; i < 3; i++) { head[i] = sm.Pre[(17 + i) * LC + l]; tail[i] = sm.Pre[(18 + i) * LC + l]; } }   // Q12
                const plba_d2 a0 = *(const plba_d2 *)(P.lo_ab + (size_t)4 * o), a1 = *(const plba_d2 *)(P.lo_ab + (size_t)4 * o + 2);
                const double ab[4] = {a0.x, a0.y, a1.x, a1.y};
                g_line_lin(P.cam, T, L, head, tail, ab, e, A, B);
            }
            const double om = OA::om(P)[o];
            const double chi2 = om * (e[0] * e[0] + e[1] * e[1]);
            double rho0 = chi2, rho1 = 1.0;
            if (ctl.stage == 0) huber(P.huber_delta, chi2, rho0, rho1);
            cost = rho0;
            { const double ww = rho1 * om; wsq = (ww > 1e-290 && ww < 1e290) ? ww * plba_rsqrt_fast(ww) : sqrt(ww); }      // sqrt(rho1 Omega)
        }
    } else {
        const bool pass0 = (ctl.iter == 0);
        double r, w, Jp[6], Jl[6];
        if (LT == LT_POINT) {
            const double *T = P.poseT[ctl.cur] + (size_t)12 * kf;
            const double Pw[3] = {sm.Pre[l], sm.Pre[LC + l], sm.Pre[2 * LC + l]};
            const plba_d2 uv = *(const plba_d2 *)(P.po_uv + (size_t)2 * o);
            const double uv2[2] = {uv.x, uv.y};
            h_point(P.cam, T, Pw, uv2, P.homog_th, Jp, Jl, r, w);
        } else {
            // Q4: inside the loop the line terms keep the MAP pose (:2700, :2010)
            const double *T = (pass0 || P.fixed_quirks) ? P.poseT[ctl.cur] + (size_t)12 * kf : P.kf_Tmap + (size_t)12 * kf;
            const plba_d2 a0 = *(const plba_d2 *)(P.lo_ab + (size_t)4 * o), a1 = *(const plba_d2 *)(P.lo_ab + (size_t)4 * o + 2);
            const double ab[4] = {a0.x, a0.y, a1.x, a1.y};
            if (LT == LT_LINE_END) {
                const double Pw[3] = {sm.Pre[l], sm.Pre[LC + l], sm.Pre[2 * LC + l]}, Qw[3] = {sm.Pre[3 * LC + l], sm.Pre[4 * LC + l], sm.Pre[5 * LC + l]};
                const double th = (pass0 || P.fixed_quirks) ? P.homog_th : 0.0000001;
                h_endline(P.cam, T, Pw, Qw, ab, th, P.fixed_quirks != 0, Jp, Jl, r, w);
            } else {
                LinePre L; load_line_pre(sm, l, L);
                h_plkline(P.cam, T, L, ab, P.homog_th, P.fixed_quirks != 0, Jp, Jl, r, w);
            }
        }
        wsq = sqrt(w);
        for (int c = 0; c < 6; c++) A[c] = Jp[c];
        for (int c = 0; c < K::D; c++) B[c] = Jl[c];
        e[0] = -r;                 // g += J r w  ==  b = -J^T (w e) with e = -r
        cost = w * r * r;
    }
    // (an edge gated out of the active set leaves zero rows: one branch instead of a select per entry)
    if (active) {
        for (int i = 0; i < K::RANK * 6; i++) sm.A[i * OC + t] = wsq * A[i];
        for (int i = 0; i < K::RANK * K::D; i++) sm.B[i * OC + t] = wsq * B[i];
        for (int i = 0; i < K::RANK; i++) sm.E[i * OC + t] = wsq * e[i];
    } else {
        for (int i = 0; i < K::RANK * 6; i++) sm.A[i * OC + t] = 0.0;
        for (int i = 0; i < K::RANK * K::D; i++) sm.B[i * OC + t] = 0.0;
        for (int i = 0; i < K::RANK; i++) sm.E[i * OC + t] = 0.0;
    }
    sm.slot[t] = slot;                   // structural: >= 0 free KF, -1 fixed observer
    sm.act[t] = active ? 1 : 0;
    return cost;
}

// ---- phase 2: per landmark: H_ll, b_l, damping, inverse ----------------------------------------------------
template <int PROF, int LT>
PLBA_HD void lm_blocks(int t0, int t1, Smem<PROF, LT> &sm, double *Hfull, double *bl, double &maxd) {
    typedef KT<PROF, LT> K;
    const int D = K::D;
    for (int i = 0; i < D * D; i++) Hfull[i] = 0.0;
    for (int i = 0; i < D; i++) bl[i] = 0.0;
    for (int t = t0; t < t1; t++) {
        for (int k = 0; k < K::RANK; k++) {
            double b[D];
            for (int c = 0; c < D; c++) b[c] = sm.B[(k * D + c) * OC + t];
            const double ek = sm.E[k * OC + t];
            for (int r = 0; r < D; r++) { bl[r] -= b[r] * ek; for (int c = r; c < D; c++) Hfull[r * D + c] += b[r] * b[c]; }
        }
    }
    maxd = 0.0;
    for (int r = 0; r < D; r++) { if (fabs(Hfull[r * D + r]) > maxd) maxd = fabs(Hfull[r * D + r]); for (int c = 0; c < r; c++) Hfull[r * D + c] = Hfull[c * D + r]; }
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
// assemble_chunk: mode 0 = diagonal pass for the initial lambda (computeLambdaInit / Hmax, :2555-2561); mode 1 = full
// ---------------------------------------------------------------------------------------------------------
template <int PROF, int LT>
PLBA_D void assemble_chunk(const DevP &Pin, const Chunk &ch, int mode) {
    PLBA_PARAMS_REF(P, Pin);
    PLBA_SMEM(raw);     // declared here (not passed in) so that every access below is a shared-space LDS / STS
    typedef KT<PROF, LT> K;
    typedef ObsAcc<LT> OA;
    const int D = K::D, RANK = K::RANK;
    Smem<PROF, LT> sm(raw);
    PROF_DECL;
    const WinCtrl &ctl = P.ctrl[ch.win];
    if (ctl.done) return;
    if (mode == 0 && !ctl.need_init) return;
    const int nlm = ch.lm1 - ch.lm0, nob = ch.ob1 - ch.ob0, nseg = ch.seg1 - ch.seg0;
    const int *ptr = OA::ptr(P);
    const double *state = OA::state(P, ctl.cur);

    PHASE_BEGIN
        if (tid < 8) sm.red[tid] = 0.0;
        if (tid < nlm) lm_precompute<PROF, LT>(P, ctl, ch.win, ch.lm0 + tid, tid, sm, state);
        if (tid <= nlm) sm.lm_t0[tid] = ptr[ch.lm0 + tid] - ch.ob0;      // CSR offsets of the chunk's landmarks, once, in shared memory
        if (tid < nseg) {
            const Seg sg = OA::segs(P)[ch.seg0 + tid];
            sm.seg_lm0[tid] = sg.lm0 - ch.lm0; sm.seg_nlm[tid] = sg.n_lm; sm.seg_nobs[tid] = sg.nobs;
            sm.seg_nfree[tid] = sg.nfree; sm.seg_task0[tid] = sg.task0; sm.seg_dtask0[tid] = sg.dtask0; sm.seg_fp0[tid] = sg.fp0;
            if (tid == nseg - 1) { sm.seg_task0[nseg] = sg.task0 + sg.nfree * (sg.nfree - 1) / 2; sm.seg_dtask0[nseg] = sg.dtask0 + sg.nfree; }
            const int *fp = OA::freepos(P) + sg.fp0;       // free-pose positions of the run's signature, staged next to its first landmark's slots
            const int t0s = ptr[sg.lm0] - ch.ob0;
            for (int i = 0; i < sg.nfree; i++) sm.freepos[t0s + i] = fp[i];
        }
    PHASE_END
    PROF_MARK(1);
    PHASE_BEGIN
        double cost = 0.0;
        if (tid < nob) {
            const int o = ch.ob0 + tid;
            const int l = OA::lm(P)[o] - ch.lm0;
            sm.lmof[tid] = l;
            cost = obs_linearize<PROF, LT>(P, ctl, o, tid, l, sm);
        }
        plba_block_add(&sm.red[0], mode == 1 ? cost : 0.0);
    PHASE_END
    PROF_MARK(2);
    PHASE_BEGIN
        if (tid < nlm) {

            const int t0 = sm.lm_t0[tid], t1 = sm.lm_t0[tid + 1];
            double H[D * D], bl[D], maxd;
            lm_blocks<PROF, LT>(t0, t1, sm, H, bl, maxd);
            if (mode == 0) plba_atomic_max_pos(&P.accmax[ch.win], maxd);
            else {
                damp_invert<PROF, D>(ctl, H);
                for (int r = 0; r < D; r++) for (int c = r; c < D; c++) sm.Hinv[symidx(r, c, D) * LC + tid] = H[r * D + c];
                for (int r = 0; r < D; r++) { double s = 0; for (int c = 0; c < D; c++) s += H[r * D + c] * bl[c]; sm.V[r * LC + tid] = s; }
            }
        }
        if (tid == 0 && mode == 1) {
            double *accw = P.acc + (size_t)4 * ch.win;
            if (PROF == PLBA_PROFILE_G) plba_atomic_add(&accw[ACC_CHI_LIN], sm.red[0]);
            else plba_atomic_add(LT == LT_POINT ? &accw[ACC_ERR_PT] : &accw[ACC_ERR_LS], sm.red[0]);
        }
    PHASE_END
    PROF_MARK(3);
    if (mode == 1) {
        // Ta = Bt H_ll^-1 per observation (hoisted out of the pair loop)
        PHASE_BEGIN
            if (tid < nob) {
                const int l = sm.lmof[tid];
                double Bt[RANK * D], Hi[Smem<PROF, LT>::NSYM];
#pragma unroll
                for (int k = 0; k < RANK * D; k++) Bt[k] = sm.B[k * OC + tid];
#pragma unroll
                for (int k = 0; k < Smem<PROF, LT>::NSYM; k++) Hi[k] = sm.Hinv[k * LC + l];
#pragma unroll
                for (int k = 0; k < RANK; k++) {
#pragma unroll
                    for (int c = 0; c < D; c++) {
                        double sum = 0;
#pragma unroll
                        for (int mm = 0; mm < D; mm++) sum += Bt[k * D + mm] * Hi[symidx(mm, c, D)];
                        sm.TA[(k * D + c) * OC + tid] = sum;
                    }
                }
            }
        PHASE_END
    PROF_MARK(4);
    }
    const int slot0 = P.win_slot0[ch.win];
    const int ld = 6 * P.win_nfree[ch.win];
    const SWin swin = s_window(P, ch.win);
    // ---- off-diagonal pose pairs (a,b), a before b in the landmark's track: S_ab -= At_a^T (Ta_a Bt_b^T) At_b ----
    PHASE_BEGIN
        const int ntasks = (nseg && mode == 1) ? sm.seg_task0[nseg] : 0;
        // every (segment, pose pair) is split into two tasks of three block rows each: twice the threads busy in this latency-bound
        // phase and half the accumulators per thread (18: no spills at 128 registers); the 2x2 / 2x6 inner products are recomputed
        for (int p2 = tid; p2 < 2 * ntasks; p2 += PLBA_NT) {
            const int p = p2 >> 1, hrow = 3 * (p2 & 1);
            int lo = 0, hi = nseg;                     // largest s with seg_task0[s] <= p
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sm.seg_task0[mid] <= p) lo = mid; else hi = mid; }
            const int s = lo, nf = sm.seg_nfree[s];
            int q = p - sm.seg_task0[s], i = 0;
            while (q >= nf - 1 - i) { q -= nf - 1 - i; i++; }
            const int j = i + 1 + q;
            const int l0 = sm.seg_lm0[s], nl = sm.seg_nlm[s], no = sm.seg_nobs[s];
            const int t0 = sm.lm_t0[l0];
            const int pa = sm.freepos[t0 + i], pb = sm.freepos[t0 + j];
            const int sa = sm.slot[t0 + pa] - slot0, sb = sm.slot[t0 + pb] - slot0;
            double blk[18];
#pragma unroll
            for (int k = 0; k < 18; k++) blk[k] = 0.0;
            for (int m = 0; m < nl; m++) {
                const int ta = t0 + m * no + pa, tb = t0 + m * no + pb;
                double M[RANK * RANK], MA[RANK * 6];
                {
                    double Ta[RANK * D], Bb[RANK * D];
#pragma unroll
                    for (int k = 0; k < RANK * D; k++) { Ta[k] = sm.TA[k * OC + ta]; Bb[k] = sm.B[k * OC + tb]; }
#pragma unroll
                    for (int k = 0; k < RANK; k++) {
#pragma unroll
                        for (int k2 = 0; k2 < RANK; k2++) {
                            double sum = 0;
#pragma unroll
                            for (int mm = 0; mm < D; mm++) sum += Ta[k * D + mm] * Bb[k2 * D + mm];
                            M[k * RANK + k2] = sum;
                        }
                    }
                }
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    double ab[RANK];
#pragma unroll
                    for (int k2 = 0; k2 < RANK; k2++) ab[k2] = sm.A[(k2 * 6 + c) * OC + tb];
#pragma unroll
                    for (int k = 0; k < RANK; k++) {
                        double sum = 0;
#pragma unroll
                        for (int k2 = 0; k2 < RANK; k2++) sum += M[k * RANK + k2] * ab[k2];
                        MA[k * 6 + c] = sum;
                    }
                }
#pragma unroll
                for (int r = 0; r < 3; r++) {
#pragma unroll
                    for (int k = 0; k < RANK; k++) {
                        const double av = sm.A[(k * 6 + hrow + r) * OC + ta];
#pragma unroll
                        for (int c = 0; c < 6; c++) blk[r * 6 + c] += av * MA[k * 6 + c];
                    }
                }
            }
            // block(a,b) lands at (sa,sb) if sa < sb, transposed at (sb,sa) if sa > sb; two observations of one landmark
            // in the same KF (sa == sb) put blk + blk^T on the diagonal block (upper part kept)
            const bool same = (sa == sb);
            const int npass = same ? 2 : 1;
            for (int pass = 0; pass < npass; pass++) {
                const bool tr = same ? (pass == 1) : (sa > sb);
                const int ra = tr ? sb : sa, cb = tr ? sa : sb;
                const SBlk sb_ = s_block(swin, ra, cb);
                double *base = sb_.p;
                // 32-bit strides (a row of the largest storage is 12 000 doubles): this address arithmetic sits on the latency path of a small window
                const int sr = (int)(tr ? sb_.sc : sb_.sr), sc = (int)(tr ? sb_.sr : sb_.sc);
#pragma unroll
                for (int r = 0; r < 3; r++) {
#pragma unroll
                    for (int c = 0; c < 6; c++) {
                        const int rr = hrow + r;
                        const bool on = !same || (pass == 0 ? (rr <= c) : (c <= rr));
                        if (on) plba_atomic_add(base + (rr * sr + c * sc), -blk[r * 6 + c]);
                    }
                }
            }
        }
    // ---- diagonal: S_aa += At^T (I - Ta Bt^T) At ;  g_a += -At^T (et + Bt v) ; diag(H_pp)_a += sum At^2 ----
    // (same phase as the off-diagonal tasks, dealt from the last thread downwards so that the two kinds of task sit in
    //  different warps whenever a chunk has fewer tasks than threads)
    {
        const int ndtasks = nseg ? sm.seg_dtask0[nseg] : 0;
        for (int p = PLBA_NT - 1 - tid; p < ndtasks; p += PLBA_NT) {
            int lo = 0, hi = nseg;
            while (hi - lo > 1) { const int mid = (lo + hi) >> 1; if (sm.seg_dtask0[mid] <= p) lo = mid; else hi = mid; }
            const int s = lo, i = p - sm.seg_dtask0[s];
            const int l0 = sm.seg_lm0[s], nl = sm.seg_nlm[s], no = sm.seg_nobs[s];
            const int t0 = sm.lm_t0[l0];
            const int pa = sm.freepos[t0 + i];
            const int sa = sm.slot[t0 + pa] - slot0;
            double Sd[21], gv[6], hd[6];
#pragma unroll
            for (int k = 0; k < 21; k++) Sd[k] = 0.0;
#pragma unroll
            for (int k = 0; k < 6; k++) { gv[k] = 0.0; hd[k] = 0.0; }
            for (int m = 0; m < nl; m++) {
                const int l = l0 + m, ta = t0 + m * no + pa;
                double Aa[RANK * 6];
#pragma unroll
                for (int k = 0; k < RANK * 6; k++) Aa[k] = sm.A[k * OC + ta];
#pragma unroll
                for (int c = 0; c < 6; c++) {
#pragma unroll
                    for (int k = 0; k < RANK; k++) hd[c] += Aa[k * 6 + c] * Aa[k * 6 + c];
                }
                if (mode == 0) continue;
                double Ba[RANK * D], N[RANK * RANK];
#pragma unroll
                for (int k = 0; k < RANK * D; k++) Ba[k] = sm.B[k * OC + ta];
#pragma unroll
                for (int k = 0; k < RANK; k++) {
#pragma unroll
                    for (int k2 = 0; k2 < RANK; k2++) {
                        double sum = 0;
#pragma unroll
                        for (int mm = 0; mm < D; mm++) sum += sm.TA[(k * D + mm) * OC + ta] * Ba[k2 * D + mm];
                        N[k * RANK + k2] = ((k == k2) ? 1.0 : 0.0) - sum;
                    }
                }
                double NA[RANK * 6], ev[RANK];
#pragma unroll
                for (int k = 0; k < RANK; k++) {
#pragma unroll
                    for (int c = 0; c < 6; c++) {
                        double sum = 0;
#pragma unroll
                        for (int k2 = 0; k2 < RANK; k2++) sum += N[k * RANK + k2] * Aa[k2 * 6 + c];
                        NA[k * 6 + c] = sum;
                    }
                    double sum = sm.E[k * OC + ta];
#pragma unroll
                    for (int mm = 0; mm < D; mm++) sum += Ba[k * D + mm] * sm.V[mm * LC + l];
                    ev[k] = sum;
                }
                {
                    int idx = 0;
#pragma unroll
                    for (int r = 0; r < 6; r++) {
#pragma unroll
                        for (int c = r; c < 6; c++) {
                            double sum = 0;
#pragma unroll
                            for (int k = 0; k < RANK; k++) sum += Aa[k * 6 + r] * NA[k * 6 + c];
                            Sd[idx++] += sum;
                        }
#pragma unroll
                        for (int k = 0; k < RANK; k++) gv[r] -= Aa[k * 6 + r] * ev[k];
                    }
                }
            }
            if (mode == 0) {
#pragma unroll
                for (int c = 0; c < 6; c++) plba_atomic_add(&P.hpp_diag_init[(size_t)6 * (slot0 + sa) + c], hd[c]);
            } else {
                int idx = 0;
                const SBlk sd_ = s_block(swin, sa, sa);
#pragma unroll
                for (int r = 0; r < 6; r++) {
#pragma unroll
                    for (int c = r; c < 6; c++) plba_atomic_add(sd_.p + (r * (int)sd_.sr + c * (int)sd_.sc), Sd[idx++]);
                    plba_atomic_add(&P.gs[(size_t)6 * (slot0 + sa) + r], gv[r]);
                    if (PROF != PLBA_PROFILE_G) plba_atomic_add(&P.hpp_diag[(size_t)6 * (slot0 + sa) + r], hd[r]);
                }
            }
        }
    }
    PHASE_END
    PROF_MARK(6);
}

// persistent grid: CTA b walks chunks b, b + gridDim, ... (points first, then lines)
template <int PROF>
PLBA_KERNEL void PLBA_BOUNDS(OC, PLBA_CTAS_PER_SM) k_assemble(const DevP *Pp, int mode) {
    if (mode == 0 && Pp->counters[CNT_NEED_INIT] == 0) { PLBA_COUNT_LAUNCH(Pp); return; }        // no window waits for its initial lambda (the counter only changes between launches)
    PLBA_PARAMS(P, Pp);
    if (PLBA_BID == 0 && (int)PLBA_TID0 == 0) P.counters[CNT_SOLVE_DONE] = 0;      // (this launch lies between the update kernel that read the flag and the solver that sets it next)
    const int ntot = P.n_chunks_pt + P.n_chunks_ls;
    for (int c = PLBA_BID; c < ntot; c += PLBA_NB) {
        if (c < P.n_chunks_pt) { const Chunk ch = P.chunks_pt[c]; assemble_chunk<PROF, LT_POINT>(P, ch, mode); }
        else { const Chunk ch = P.chunks_ls[c - P.n_chunks_pt]; assemble_chunk<PROF, LineOf<PROF>::LT>(P, ch, mode); }
    }
}

// ---------------------------------------------------------------------------------------------------------
// update_chunk: re-linearise, back-substitute x_l = H_ll^-1 (b_l - W^T x_p), retract, evaluate the new cost
// ---------------------------------------------------------------------------------------------------------
template <int PROF, int LT>
PLBA_D void update_chunk(const DevP &Pin, const Chunk &ch, bool wait_solve) {
    PLBA_PARAMS_REF(P, Pin);
    PLBA_SMEM(raw);
    typedef KT<PROF, LT> K;
    typedef ObsAcc<LT> OA;
    const int D = K::D, RANK = K::RANK;
    Smem<PROF, LT> sm(raw);
    PROF_DECL;
    const WinCtrl &ctl = P.ctrl[ch.win];
    if (ctl.done) return;
    const int nlm = ch.lm1 - ch.lm0, nob = ch.ob1 - ch.ob0;
    const int *ptr = OA::ptr(P);
    const double *state = OA::state(P, ctl.cur);
    double *state_new = OA::state(P, ctl.cur ^ 1);

    PHASE_BEGIN
        if (tid < 8) sm.red[tid] = 0.0;
        if (tid < nlm) lm_precompute<PROF, LT>(P, ctl, ch.win, ch.lm0 + tid, tid, sm, state);
        if (tid <= nlm) sm.lm_t0[tid] = ptr[ch.lm0 + tid] - ch.ob0;
    PHASE_END
    PROF_MARK(21);
    PHASE_BEGIN
        if (tid < nob) {
            const int o = ch.ob0 + tid;
            const int l = OA::lm(P)[o] - ch.lm0;
            sm.lmof[tid] = l;
            obs_linearize<PROF, LT>(P, ctl, o, tid, l, sm);
        }
    PHASE_END
    // per landmark: H_ll, b_l, the damped inverse — still the old state only; kept in the landmark thread's registers across the wait
    THR_ARR(double, Hs, D * D); THR_ARR(double, bls, D);
    PHASE_BEGIN
        THR_BIND(Hs); THR_BIND(bls);
        if (tid < nlm) {
            double maxd;
            lm_blocks<PROF, LT>(sm.lm_t0[tid], sm.lm_t0[tid + 1], sm, Hs, bls, maxd);
            damp_invert<PROF, D>(ctl, Hs);
        }
    PHASE_END
    PROF_MARK(27);
    // everything above reads the OLD state only; from here on the pose step x_p and the trial poses of the solver are needed
    if (wait_solve) plba_wait_flag(&P.counters[CNT_SOLVE_DONE]);
    PHASE_BEGIN
        double sc = 0.0;
        if (tid < nob) {
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
    PROF_MARK(22);
    PHASE_BEGIN
        THR_BIND(Hs); THR_BIND(bls);
        double sc = 0.0, d2 = 0.0;
        if (tid < nlm) {
            const int lm = ch.lm0 + tid;
            const int t0 = sm.lm_t0[tid], t1 = sm.lm_t0[tid + 1];
            const double *H = Hs, *bl = bls;
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
            if (PROF == PLBA_PROFILE_G || ctl.apply) {
                for (int i = 0; i < D; i++) state_new[(size_t)D * lm + i] = nw[i];
                if (LT == LT_LINE_ORTH) {
                    // the new state's Plücker vector / U / W for the next linearisation (both kernels read it instead of redoing the trigonometry),
                    // and, profile G, the Plücker vector for the cost evaluation below
                    double pl[6]; orth_to_plk(nw, pl);
                    LinePre Ln; line_pre_from_plk(pl, Ln);
                    double *c = P.lpre[ctl.cur ^ 1] + (size_t)LPRE_N * lm;
                    for (int i = 0; i < 3; i++) { c[i] = Ln.n[i]; c[3 + i] = Ln.d[i]; c[6 + i] = Ln.u1[i]; c[9 + i] = Ln.u2[i]; c[12 + i] = Ln.u3[i]; }
                    c[15] = Ln.w1; c[16] = Ln.w2;
                    if (PROF == PLBA_PROFILE_G) for (int i = 0; i < 6; i++) sm.Pre[i * LC + tid] = pl[i];
                }
            }
            for (int i = 0; i < D; i++) sm.Xl[i * LC + tid] = nw[i];
        }
        plba_block_add(&sm.red[1], sc);
        plba_block_add(&sm.red[2], d2);
    PHASE_END
    PROF_MARK(23);
    if (PROF == PLBA_PROFILE_G) {
        // new cost at the trial state (computeActiveErrors + activeRobustChi2 after update): needs the trial poses (flag value 2)
        if (wait_solve) plba_wait_flag(&P.counters[CNT_SOLVE_DONE], 2);
    PROF_MARK(24);
        PHASE_BEGIN
            double rho0 = 0.0;
            if (tid < nob && sm.act[tid]) {
                const int o = ch.ob0 + tid, l = sm.lmof[tid];
                const double *T = P.poseT[ctl.cur ^ 1] + (size_t)12 * OA::kf(P)[o];
                double e[2];
                if (LT == LT_POINT) {
                    const double Pw[3] = {sm.Xl[l], sm.Xl[LC + l], sm.Xl[2 * LC + l]}; double zc;
                    const plba_d2 uv = *(const plba_d2 *)(P.po_uv + (size_t)2 * o);
                    const double uv2[2] = {uv.x, uv.y};
                    g_point_error(P.cam, T, Pw, uv2, e, zc);
                } else {
                    const double n[3] = {sm.Pre[l], sm.Pre[LC + l], sm.Pre[2 * LC + l]}, d[3] = {sm.Pre[3 * LC + l], sm.Pre[4 * LC + l], sm.Pre[5 * LC + l]};
                    const plba_d2 a0 = *(const plba_d2 *)(P.lo_ab + (size_t)4 * o), a1 = *(const plba_d2 *)(P.lo_ab + (size_t)4 * o + 2);
                    const double ab[4] = {a0.x, a0.y, a1.x, a1.y};
                    g_line_error(P.cam, T, n, d, ab, e);
                }
                const double chi2 = OA::om(P)[o] * (e[0] * e[0] + e[1] * e[1]);
                OA::chi2(P)[o] = chi2;                       // the cached _error of the edge (e->chi2())
                double rho1;
                rho0 = chi2;
                if (ctl.stage == 0) huber(P.huber_delta, chi2, rho0, rho1);
            }
            plba_block_add(&sm.red[0], rho0);
        PHASE_END
    PROF_MARK(25);
    }
    PHASE_BEGIN
        if (tid == 0) {
            double *acc = P.accB + (size_t)4 * ch.win;
            if (PROF == PLBA_PROFILE_G) plba_atomic_add(&acc[ACC_CHI_NEW], sm.red[0]);
            plba_atomic_add(&acc[ACC_SCALE], sm.red[1]);
            plba_atomic_add(&acc[ACC_DX2], sm.red[2]);
        }
    PHASE_END
    PROF_MARK(26);
}

// ---- per-window controller steps (device functions; called from fused kernels and from the stand-alone ones) ----
// lambda init at the start of each g2o stage / hand-LM pass 0 (computeLambdaInit; src/mapHandler.cpp:2555-2561)
PLBA_D void lambda_init_window(const DevP &P, int w) {
    WinCtrl &c = P.ctrl[w];
    if (!c.done && c.need_init) {
        double m = plba_ld_l2(&P.accmax[w]);
        if (P.n_ranks > 1) { m = 0.0; for (int r = 0; r < P.n_ranks; r++) { const double v = plba_ld_l2(&P.maxslots[(size_t)r * P.n_win + w]); if (v > m) m = v; } }
        const int s0 = P.win_slot0[w], nf = P.win_nfree[w];
        for (int i0 = 0; i0 < 6 * nf; i0 += 12) {      // twelve loads in flight (one thread per window: the loads, not the maximum, are the latency)
            double hv[12];
#pragma unroll
            for (int u = 0; u < 12; u++) hv[u] = (i0 + u < 6 * nf) ? fabs(plba_ld_l2(&P.hpp_diag_init[(size_t)6 * s0 + i0 + u])) : 0.0;
#pragma unroll
            for (int u = 0; u < 12; u++) { if (hv[u] > m) m = hv[u]; if (i0 + u < 6 * nf) P.hpp_diag_init[(size_t)6 * s0 + i0 + u] = 0.0; }
        }
        P.accmax[w] = 0.0;
        if (P.profile == PLBA_PROFILE_G) { c.lambda = P.lm_tau * m; c.ni = 2.0; }
        else {
            if (P.gba && !P.fixed_quirks) m = (double)(long long)m;      // `int Hmax` of the GBA shell (src/mapHandler.cpp:3386-3391)
            c.lambda = P.lambda_lba_lm * m;
        }
        c.need_init = 0;
        plba_atomic_add_i(&P.counters[CNT_NEED_INIT], -1);
    }
    if (c.do_gate) { c.do_gate = 0; plba_atomic_add_i(&P.counters[CNT_GATE], -1); }
}

// profile H, between assembly and solve: normalise err, stop tests, accept decision (src/mapHandler.cpp:2796-2814)
PLBA_D void control_h_pre_window(const DevP &P, int w) {
    WinCtrl &c = P.ctrl[w];
    if (c.done) return;
    const double err_pt = plba_ld_l2(&P.acc[(size_t)4 * w + ACC_ERR_PT]), err_ls = plba_ld_l2(&P.acc[(size_t)4 * w + ACC_ERR_LS]);
    double err = err_pt + err_ls;
    const double zero = 0.0;
    const bool div0 = !P.fixed_quirks && (c.iter == 0 || P.profile == PLBA_PROFILE_H_PLK || P.gba);    // Q1 (GBA: every iteration, :3662)
    err = div0 ? err / zero : err / (P.gba ? (double)c.n_obs : (double)(c.n_lm_pt + c.n_lm_ls));
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
            P.acc[(size_t)4 * w + ACC_ERR_PT] = 0.0; P.acc[(size_t)4 * w + ACC_ERR_LS] = 0.0;
        } else c.apply = (err > c.err_prev) ? 0 : 1;
    }
}

// one free keyframe: apply the pose part of the step (g2o oplus / SE(3) retraction of the hand LM)
PLBA_D void pose_update_slot(const DevP &P, int s, double &scale_part, double &dx2_part) {
    const int kf = P.slot_kf[s];
    WinCtrl &c = P.ctrl[P.kf_win[kf]];
    scale_part = 0.0; dx2_part = 0.0;
    if (c.done) return;
    double x[6], d2 = 0.0;
    for (int i = 0; i < 6; i++) { x[i] = P.xp[(size_t)6 * s + i]; d2 += x[i] * x[i]; }
    double Tn[12];
    if (P.profile == PLBA_PROFILE_G) {
        pose_oplus_g(P.poseT[c.cur] + (size_t)12 * kf, x, Tn);
        for (int i = 0; i < 12; i++) P.poseT[c.cur ^ 1][(size_t)12 * kf + i] = Tn[i];
        scale_part = c.lambda * d2;
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
        dx2_part = d2;
    }
}

PLBA_D void write_trace(const DevP &P, int w, WinCtrl &c, int accepted, int stop, double chi, double chi_new, double rho, double lambda, double scale, double dx, double err_pt, double err_ls) {
    if (c.n_trace < P.trace_cap) {
        plba_trace_rec &t = P.trace[(size_t)w * P.trace_cap + c.n_trace];
        t.window = w; t.stage = c.stage; t.iter = c.iter; t.trial = c.trial; t.accepted = accepted; t.stop = stop;
        t.chi = chi; t.chi_new = chi_new; t.rho = rho; t.lambda = lambda; t.scale = scale; t.dx_norm = dx; t.err_pt = err_pt; t.err_ls = err_ls;
    }
    c.n_trace++;
}

// The LM controller for one window: g2o OptimizationAlgorithmLevenberg::solve + SparseOptimizer::optimize
//  for profile G, src/mapHandler.cpp:2808-2837 for profile H.  Consumes (and clears) the accumulators.
PLBA_D void control_window(const DevP &P, int w) {
    WinCtrl &c = P.ctrl[w];
    double *acc = P.acc + (size_t)4 * w, *accB = P.accB + (size_t)4 * w;
    const double a_chi = plba_ld_l2(&acc[ACC_CHI_LIN]), a_ept = plba_ld_l2(&acc[ACC_ERR_PT]), a_els = plba_ld_l2(&acc[ACC_ERR_LS]);
    const double b_chi = plba_ld_l2(&accB[ACC_CHI_NEW]), b_scale = plba_ld_l2(&accB[ACC_SCALE]), b_dx2 = plba_ld_l2(&accB[ACC_DX2]);
    for (int i = 0; i < 4; i++) { acc[i] = 0.0; accB[i] = 0.0; }
    if (c.done) return;
    c.n_trials++;
    plba_atomic_add_i(&P.counters[CNT_TRIALS], 1);
    // PLBA_E_NUMERIC: the reduced camera system was not positive definite in EVERY trial of an outer iteration (profile G) / in an
    // iteration of the hand LM.  The reference's SimplicialLDLT has no positivity requirement and would carry on with an indefinite
    // system; this library rejects such trials, so the caller is told that the result no longer follows the reference.
    int fail_streak = (c.trial == 0) ? 0 : (c.numeric & 0xffff);
    if (c.solve_fail) fail_streak++;
    c.numeric = (c.numeric & ~0xffff) | (fail_streak & 0xffff);
    if (c.solve_fail && P.profile != PLBA_PROFILE_G) c.numeric |= (1 << 30);
    if (P.profile == PLBA_PROFILE_G) {
        if (c.trial == 0) c.chi_cur = a_chi;                // currentChi = activeRobustChi2()
        const double tempChi = c.solve_fail ? 1.7976931348623157e308 : b_chi;
        const double scale = (b_scale + c.scale_pose) + 1e-3;
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
        write_trace(P, w, c, accepted, terminate ? 1 : 0, chi_before, tempChi, rho, lam_used, scale, 0.0, a_ept, a_els);
        if (again) c.trial = qmax;
        else {
            if (fail_streak == qmax) c.numeric |= (1 << 30);       // every trial of this outer iteration failed in the factorisation
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
        const double dx = sqrt(b_dx2 + c.dx2_pose);
        const double lam_used = c.lambda;
        int stop = 0;
        if (c.iter == 0) { c.cur ^= 1; }
        else {
            if (c.apply) { c.lambda *= P.lambda_lba_k; c.cur ^= 1; } else c.lambda /= P.lambda_lba_k;
            if (dx < P.min_error_change) stop = 2;
        }
        write_trace(P, w, c, c.apply, stop, c.chi_cur, 0.0, 0.0, lam_used, 0.0, dx, a_ept, a_els);
        c.err_prev = c.chi_cur;
        c.iter++;
        if (stop || c.iter >= P.max_iters_lba) { c.done = 1; plba_atomic_add_i(&P.counters[CNT_DONE], 1); }
    }
    c.scale_pose = 0; c.dx2_pose = 0; c.solve_fail = 0;
}

// end of an LM round: decide whether the graph's WHILE node iterates again and whether the next round starts with the
// "prep" block (chi2 gate + lambda init).  max_rounds bounds the loop whatever the data do.
PLBA_D void round_epilogue(const DevP &P, int flags) {
    const int rounds = P.counters[CNT_ROUNDS] + 1;
    P.counters[CNT_ROUNDS] = rounds;
    if (flags & KF_IN_GRAPH) {
        const int done = plba_ld_l2(&P.counters[CNT_DONE]);
        const int prep = plba_ld_l2(&P.counters[CNT_NEED_INIT]) + plba_ld_l2(&P.counters[CNT_GATE]);
        plba_graph_set(P.cond_while, (done < P.n_win && rounds < P.max_rounds) ? 1u : 0u);
        plba_graph_set(P.cond_prep, prep > 0 ? 1u : 0u);
        if (P.cond_prep2) plba_graph_set(P.cond_prep2, prep > 0 ? 1u : 0u);
    }
}

// The reduced camera system of small windows is consumed by k_solve_small (one CTA per window) but cleared HERE, by every CTA of the
// update kernel that runs next: clearing 115 KB (config 2) with the store path of the solver's single SM costs ~4 000 cycles of the
// latency-critical kernel (measured), spread over the grid it is free.  Coalesced 128-bit stores.
PLBA_D void clear_consumed_S(const DevP &P, int part, int nparts) {
    PHASE_BEGIN
        const long long n2 = P.S_clear_doubles / 2;
        plba_d2 *Sz = (plba_d2 *)P.S;
        const plba_d2 z = {0.0, 0.0};
        for (long long i = (long long)part * PLBA_NT + tid; i < n2; i += (long long)nparts * PLBA_NT) Sz[i] = z;
    PHASE_END
}

template <int PROF>
PLBA_KERNEL void PLBA_BOUNDS(OC, PLBA_CTAS_PER_SM) k_update(const DevP *Pp, int flags) {
    PLBA_SMEM(raw);
    PLBA_PARAMS(P, Pp);
    // KF_WAIT_SOLVE (single small window, profile G, inside the LM-loop graph): this kernel is launched BESIDE k_solve_small (programmatic
    // edge: it starts once the solver's CTA is resident), re-linearises its chunks at the old state while the solver factors, and waits
    // for the solver's flag only where x_p is first needed.  S is consumed by the solver, so it is cleared after the wait; CTAs without a
    // chunk leave at once (they must not hold SM slots while they spin).
    const bool wait_solve = (flags & KF_WAIT_SOLVE) != 0;
    if (P.S_clear_doubles && !wait_solve) clear_consumed_S(P, PLBA_BID, PLBA_NB);
    const int ntot = P.n_chunks_pt + P.n_chunks_ls;
    for (int c = PLBA_BID; c < ntot; c += PLBA_NB) {
        if (c < P.n_chunks_pt) { const Chunk ch = P.chunks_pt[c]; update_chunk<PROF, LT_POINT>(P, ch, wait_solve); }
        else { const Chunk ch = P.chunks_ls[c - P.n_chunks_pt]; update_chunk<PROF, LineOf<PROF>::LT>(P, ch, wait_solve); }
    }
    if (wait_solve && P.S_clear_doubles && PLBA_BID < ntot) {
        plba_wait_flag(&P.counters[CNT_SOLVE_DONE], 2);   // (a chunk of a finished window returns without waiting)
        clear_consumed_S(P, PLBA_BID, PLBA_NB < ntot ? PLBA_NB : ntot);
    }
    if (!(flags & KF_FUSE_CONTROL)) return;
    // the last CTA to arrive runs the controller for every window
    int *last = (int *)raw;
    PHASE_BEGIN
        if (tid == 0) { plba_fence(); *last = (plba_atomic_fetch_add_i(&P.counters[CNT_TICKET], 1) == PLBA_NB - 1) ? 1 : 0; }
    PHASE_END
    if (!*last) return;
    PHASE_BEGIN
        if (tid == 0) plba_fence();
        for (int w = tid; w < P.n_win; w += PLBA_NT) control_window(P, w);
    PHASE_END
    PHASE_BEGIN
        if (tid == 0) { P.counters[CNT_TICKET] = 0; plba_fence(); round_epilogue(P, flags); }
    PHASE_END
}

// ---- stand-alone per-window / per-keyframe kernels (host-driven loop: sharded multi-GPU and the tiled solver path) ----
// (the prep kernels are launched unconditionally by the pipelined host loop: they leave at once when no window asks for them)
PLBA_KERNEL void k_lambda_init(const DevP *Pp) {
    PLBA_PARAMS(P, Pp);
    if (P.counters[CNT_NEED_INIT] == 0 && P.counters[CNT_GATE] == 0) return;
    PHASE_BEGIN
        for (int w = PLBA_BID * PLBA_NT + tid; w < P.n_win; w += PLBA_NB * PLBA_NT) lambda_init_window(P, w);
        if (PLBA_BID == 0 && tid == 0) P.counters[CNT_PREPS]++;
    PHASE_END
}
// sharded path, before the prep exchange: this rank's landmark-diagonal maximum goes into its slot, the other slots are cleared
PLBA_KERNEL void k_pack_max(const DevP *Pp) {
    PLBA_PARAMS(P, Pp);
    PHASE_BEGIN
        for (int i = PLBA_BID * PLBA_NT + tid; i < P.n_ranks * P.n_win; i += PLBA_NB * PLBA_NT) {
            const int r = i / P.n_win, w = i - r * P.n_win;
            P.maxslots[i] = (r == P.rank) ? P.accmax[w] : 0.0;
        }
    PHASE_END
}
PLBA_KERNEL void k_control_h_pre(const DevP *Pp) {
    PLBA_PARAMS(P, Pp);
    PHASE_BEGIN
        for (int w = PLBA_BID * PLBA_NT + tid; w < P.n_win; w += PLBA_NB * PLBA_NT) control_h_pre_window(P, w);
    PHASE_END
}
PLBA_KERNEL void k_pose_update(const DevP *Pp) {
    PLBA_PARAMS(P, Pp);
    PHASE_BEGIN
        for (int s = PLBA_BID * PLBA_NT + tid; s < P.n_free; s += PLBA_NB * PLBA_NT) {
            double sc, d2; pose_update_slot(P, s, sc, d2);
            WinCtrl &c = P.ctrl[P.kf_win[P.slot_kf[s]]];
            if (sc != 0.0) plba_atomic_add(&c.scale_pose, sc);
            if (d2 != 0.0) plba_atomic_add(&c.dx2_pose, d2);
        }
    PHASE_END
}
PLBA_KERNEL void k_control(const DevP *Pp, int flags) {
    PLBA_PARAMS(P, Pp);
    PHASE_BEGIN
        for (int w = tid; w < P.n_win; w += PLBA_NT) control_window(P, w);
    PHASE_END
    PHASE_BEGIN
        if (tid == 0) round_epilogue(P, flags);
    PHASE_END
}

// chi2 gate between the two g2o stages (src/mapHandler.cpp:6125-6147): level 1 <=> chi2 > 5.991 (or z_c <= 0 for points)
template <int LT>
PLBA_D void gate_obs(const DevP &P, int o) {
    typedef ObsAcc<LT> OA;
    const int kf = OA::kf(P)[o];
    const WinCtrl &c = P.ctrl[P.kf_win[kf]];
    if (!c.do_gate) return;
    bool bad = OA::chi2(P)[o] > P.chi2_gate;
    if (LT == LT_POINT) {
        const double *T = P.poseT[c.cur] + (size_t)12 * kf;
        const double *Pw = P.pts[c.cur] + (size_t)3 * OA::lm(P)[o];
        const double z = T[8] * Pw[0] + T[9] * Pw[1] + T[10] * Pw[2] + T[11];
        if (!(z > 0.0)) bad = true;
    }
    OA::lvl(P)[o] = bad ? 1 : 0;
}
PLBA_KERNEL void k_gate(const DevP *Pp) {
    PLBA_PARAMS(P, Pp);
    if (P.counters[CNT_GATE] == 0) return;
    PHASE_BEGIN
        for (int o = PLBA_BID * PLBA_NT + tid; o < P.n_pobs; o += PLBA_NB * PLBA_NT) gate_obs<LT_POINT>(P, o);
        for (int o = PLBA_BID * PLBA_NT + tid; o < P.n_lobs; o += PLBA_NB * PLBA_NT) gate_obs<LT_LINE_ORTH>(P, o);
    PHASE_END
}

// final per-observation test (src/mapHandler.cpp:6156-6161, 6224-6230): level-1 edges are re-evaluated at the final estimate.
// chi2 and flags leave in the CALLER's observation order (run of the landmark in the staged arrays + position inside the run), so that
// plba_download copies them straight through
template <int LT>
PLBA_D void final_obs(const DevP &P, int o, unsigned char *flags_out, double *chi2_out) {
    typedef ObsAcc<LT> OA;
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
    const size_t i = (size_t)(LT == LT_POINT ? P.pt_src : P.ls_src)[lm] + (size_t)(o - (LT == LT_POINT ? P.pt_ptr : P.ls_ptr)[lm]);
    flags_out[i] = f; chi2_out[i] = chi2;
}

// write-back: T_kf_w = estimate^-1 (:6302) / expmap_se3(X) (:2851-2852); NDw = changeOrthToPluker(orth) (:6318);
// final chi2 test.  One launch, outputs land in the contiguous D2H staging region.
struct ExportP { double *T_wc, *x, *pt, *ls, *plk, *pchi, *lchi; unsigned char *pf, *lf; int ls_dim, q9, do_final, pad; };
PLBA_KERNEL void k_export(const DevP *Pp, ExportP E) {
    PLBA_PARAMS(P, Pp);
    PHASE_BEGIN
        const int g0 = PLBA_BID * PLBA_NT + tid, gs = PLBA_NB * PLBA_NT;
        for (int kf = g0; kf < P.n_kf; kf += gs) {
            const int s = P.kf_slot[kf];
            const WinCtrl &c = P.ctrl[P.kf_win[kf]];
            double T[12];
            if (s < 0) inv_se3(P.kf_Tmap + (size_t)12 * kf, T);          // fixed observer: unchanged (plba_solve copies the caller's rows through)
            else if (P.profile == PLBA_PROFILE_G) inv_se3(P.poseT[c.cur] + (size_t)12 * kf, T);
            else { exp_se3(P.Xkf[c.cur] + (size_t)6 * s, T); for (int i = 0; i < 6; i++) E.x[(size_t)6 * s + i] = P.Xkf[c.cur][(size_t)6 * s + i]; }
            for (int i = 0; i < 12; i++) E.T_wc[(size_t)12 * kf + i] = T[i];
        }
        for (int i = g0; i < P.n_pt; i += gs) { const int b = P.ctrl[P.pt_win[i]].cur; for (int k = 0; k < 3; k++) E.pt[(size_t)3 * i + k] = P.pts[b][(size_t)3 * i + k]; }
        for (int i = g0; i < P.n_ls; i += gs) {
            const int b = P.ctrl[P.ls_win[i]].cur;
            double v[6];
            for (int k = 0; k < E.ls_dim; k++) { v[k] = P.lns[b][(size_t)E.ls_dim * i + k]; E.ls[(size_t)E.ls_dim * i + k] = v[k]; }
            if (E.ls_dim == 4) {
                double pl[6];
                if (E.q9) { double dx[4]; for (int k = 0; k < 4; k++) dx[k] = v[k] - P.lns0[(size_t)4 * i + k]; orth_to_plk(dx, pl); }   // Q9 (:2194)
                else orth_to_plk(v, pl);
                for (int k = 0; k < 6; k++) E.plk[(size_t)6 * i + k] = pl[k];
            }
        }
        if (E.do_final) {
            for (int o = g0; o < P.n_pobs; o += gs) final_obs<LT_POINT>(P, o, E.pf, E.pchi);
            for (int o = g0; o < P.n_lobs; o += gs) final_obs<LT_LINE_ORTH>(P, o, E.lf, E.lchi);
        }
    PHASE_END
}

// state <- uploaded initial values (plba_reset_state): one launch.  gather = 1 (the launch that follows an upload): the observations are
// first brought from the caller's order into the internal one — a landmark's observations are ONE run in the caller's arrays (validated:
// landmark-major) and one run here, so a thread moves the run of one landmark; weights as the reference forms them:
// const float& invSigma2 = 1 / sigma2 (src/mapHandler.cpp:6009, Q13; double division and the float conversion round as on the host)
PLBA_KERNEL void k_reset(const DevP *Pp, int ls_dim, size_t sys_doubles, double *sysbuf, int gather) {
    PLBA_PARAMS(P, Pp);
    PHASE_BEGIN
        const size_t g0 = (size_t)PLBA_BID * PLBA_NT + tid, gs = (size_t)PLBA_NB * PLBA_NT;
        if (gather) {
            int *po_kf = const_cast<int *>(P.po_kf), *po_lm = const_cast<int *>(P.po_lm), *lo_kf = const_cast<int *>(P.lo_kf), *lo_lm = const_cast<int *>(P.lo_lm);
            double *po_uv = const_cast<double *>(P.po_uv), *po_om = const_cast<double *>(P.po_om), *lo_ab = const_cast<double *>(P.lo_ab), *lo_om = const_cast<double *>(P.lo_om);
            for (size_t g = g0; g < (size_t)P.n_pt; g += gs) {
                const int o0 = P.pt_ptr[g], no = P.pt_ptr[g + 1] - o0; const size_t i0 = (size_t)P.pt_src[g];
                for (int j = 0; j < no; j++) {
                    const size_t o = (size_t)o0 + j, i = i0 + j;
                    po_uv[2 * o] = P.r_po_uv[2 * i]; po_uv[2 * o + 1] = P.r_po_uv[2 * i + 1];
                    po_kf[o] = P.r_po_kf[i]; po_lm[o] = (int)g; po_om[o] = (double)(float)(1.0 / P.r_po_s2[i]);
                }
            }
            for (size_t g = g0; g < (size_t)P.n_ls; g += gs) {
                const int o0 = P.ls_ptr[g], no = P.ls_ptr[g + 1] - o0; const size_t i0 = (size_t)P.ls_src[g];
                for (int j = 0; j < no; j++) {
                    const size_t o = (size_t)o0 + j, i = i0 + j;
                    for (int k = 0; k < 4; k++) lo_ab[4 * o + k] = P.r_lo_ab[4 * i + k];
                    lo_kf[o] = P.r_lo_kf[i]; lo_lm[o] = (int)g; lo_om[o] = (double)(float)(1.0 / P.r_lo_s2[i]);
                }
            }
        }
        for (size_t i = g0; i < (size_t)12 * P.n_kf; i += gs) { const double v = P.kf_Tmap[i]; P.poseT[0][i] = v; P.poseT[1][i] = v; }
        for (size_t i = g0; i < (size_t)6 * P.n_free; i += gs) { const double v = P.X0[i]; P.Xkf[0][i] = v; P.Xkf[1][i] = v; P.xp[i] = 0.0; }
        for (size_t i = g0; i < (size_t)3 * P.n_pt; i += gs) { const double v = P.pts0[i]; P.pts[0][i] = v; P.pts[1][i] = v; }
        if (ls_dim == 4) {      // initial orthonormal coordinates from the map's Plücker vectors (MapLine::changePlukerToOrth, src/mapFeatures.cpp:186-201)
            double *lns0 = const_cast<double *>(P.lns0);
            for (size_t l = g0; l < (size_t)P.n_ls; l += gs) {
                double o[4]; plk_to_orth(P.lns_map + 6 * l, o);
                for (int i = 0; i < 4; i++) { lns0[4 * l + i] = o[i]; P.lns[0][4 * l + i] = o[i]; P.lns[1][4 * l + i] = o[i]; }
                double pl[6]; orth_to_plk_sc(o, pl);
                LinePre Ln; line_pre_from_plk(pl, Ln);
                for (int b = 0; b < 2; b++) {
                    double *c = P.lpre[b] + (size_t)LPRE_N * l;
                    for (int i = 0; i < 3; i++) { c[i] = Ln.n[i]; c[3 + i] = Ln.d[i]; c[6 + i] = Ln.u1[i]; c[9 + i] = Ln.u2[i]; c[12 + i] = Ln.u3[i]; }
                    c[15] = Ln.w1; c[16] = Ln.w2;
                }
            }
        } else
        for (size_t i = g0; i < (size_t)ls_dim * P.n_ls; i += gs) { const double v = P.lns0[i]; P.lns[0][i] = v; P.lns[1][i] = v; }
        for (size_t i = g0; i < (size_t)P.n_pobs; i += gs) { P.po_lvl[i] = 0; P.po_chi2[i] = 0.0; }
        for (size_t i = g0; i < (size_t)P.n_lobs; i += gs) { P.lo_lvl[i] = 0; P.lo_chi2[i] = 0.0; }
        for (size_t i = g0; i < sys_doubles; i += gs) sysbuf[i] = 0.0;
        for (size_t w = g0; w < (size_t)P.n_win; w += gs) { P.ctrl[w] = P.ctrl0[w]; }
        if (g0 == 0) {
            int ndone = 0;
            for (int w = 0; w < P.n_win; w++) ndone += P.ctrl0[w].done;
            P.counters[CNT_DONE] = ndone; P.counters[CNT_NEED_INIT] = P.n_win - ndone; P.counters[CNT_GATE] = 0;
            P.counters[CNT_TRIALS] = 0; P.counters[CNT_TICKET] = 0; P.counters[CNT_ROUNDS] = 0; P.counters[CNT_PREPS] = 0;
            P.counters[CNT_WORK_PT] = 0; P.counters[CNT_WORK_LS] = 0; P.counters[CNT_WTICKET] = 0; P.counters[CNT_KLAUNCH] = 0;
        }
    PHASE_END
}

}  // namespace plba
