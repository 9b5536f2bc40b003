// Damped solve of the reduced camera system  (S + damping) x_p = g_red   (subsystem 4; replaces
// Eigen::SimplicialLDLT of src/mapHandler.cpp:2566-2568 and g2o LinearSolverEigen, src/mapHandler.cpp:5925).
// S is symmetric positive definite after damping: Cholesky S = L L^T in FP64 on the vector pipe (tcgen05 has no
// f64 kind; DMMA is considered only for the large-window GEMM update).
//   k_solve_small : one CTA per window, whole system in shared memory (n = 6 Nkf <= 144: configs 1-3).  Right-looking
//                   Cholesky in panels of one pose block (6 columns): the 6x6 diagonal block is factored redundantly
//                   in the registers of every thread that owns a row below it, so a panel costs two block barriers
//                   instead of twelve; the right-hand side rides along as row n (L y = g for free).  The kernel also
//                   runs the hand-LM pre-solve controller, clears the accumulators it consumed and applies the pose step.
//   k_potrf_tile / k_trsm_tiles / k_syrk_tiles / k_trisolve_large : tiled right-looking factorisation in HBM/L2
//                   for large windows (configs 4-5), 48 x 48 tiles, many CTAs per step.
#pragma once
#include "plba_kernels.h"

namespace plba {

enum { SMALL_NMAX = 144 };


// Left-looking Cholesky of a symmetric positive definite matrix held as a LOWER triangle in shared memory (row stride ldm, odd),
// in panels of one pose block (6 columns).  nd = dimension (multiple of 6); rows nd..nr (if any) are extra rows that ride along
// (the right-hand side of k_solve_small: L y = g comes for free).  The factored 6x6 diagonal blocks are returned in Lblk (M keeps
// their un-factored values), 1/L_jj in dinv.  Per panel three phases:
//  U: every row r >= k0 (row n = right-hand side) accumulates sum_{q<k0} L[r][q] L[k0+c][q] for the panel's 6 columns.  One
//     thread per (row, half of the q range): its own row is read once per q (stride ldm, odd: conflict-free) and the six
//     rows k0..k0+5 are BROADCAST loads shared by the whole warp, so the phase is FMA-bound, not shared-memory-bound,
//     and no trailing matrix is ever re-written;
//  A: every row folds the partial sums in, the 6x6 diagonal block is factored redundantly in the registers of each row's
//     thread (no extra barrier), then the row is solved against it.
//  F: the partial sums are folded into the panel, one thread per entry.
// Must be called by all 256 threads of the CTA.
#ifndef PLBA_SOLVE_INPLACE_K0
#define PLBA_SOLVE_INPLACE_K0 36          // look-back (columns) up to which one thread per row updates the panel in place (measured: see profiles/README.md)
#endif
PLBA_D void chol_lower_panels(double *M, int ldm, int nd, int nr, double *dinv, double *Lblk, double *part, int *fail) {
    const int nf = nd / 6, n = nr;
for (int kb = 0; kb < nf; kb++) {
    const int k0 = 6 * kb, m = n - k0 + 1;
    // the q range of a row is split over as many threads as the 256-thread CTA allows: late panels have few rows but long rows
    // short look-back (first panels): one thread per row updates the panel in place and the fold phase is skipped
    const int mpad = (m + 31) & ~31, nsplit = (k0 <= PLBA_SOLVE_INPLACE_K0) ? 1 : (mpad <= 32) ? 8 : (mpad <= 64) ? 4 : (mpad <= 128) ? 2 : 1;
    PHASE_BEGIN
        const int sp = tid / mpad, rr = tid - sp * mpad;
        if (k0 > 0 && sp < nsplit && rr < m) {
            const int r = k0 + rr;
            const int qs = (k0 + nsplit - 1) / nsplit, q0 = sp * qs, q1 = (q0 + qs < k0) ? q0 + qs : k0;
            double acc[6] = {0, 0, 0, 0, 0, 0};
            const double *Mr = M + (size_t)r * ldm, *Mk = M + (size_t)k0 * ldm;
#pragma unroll 4
            for (int q = q0; q < q1; q++) {
                const double v = Mr[q];
#pragma unroll
                for (int c = 0; c < 6; c++) acc[c] += v * Mk[(size_t)c * ldm + q];
            }
            if (nsplit == 1) {
#pragma unroll
                for (int c = 0; c < 6; c++) M[(size_t)r * ldm + k0 + c] -= acc[c];
            } else {
#pragma unroll
                for (int c = 0; c < 6; c++) part[((size_t)sp * mpad + rr) * 6 + c] = acc[c];
            }
        }
    PHASE_END
    if (nsplit > 1) {
    PHASE_BEGIN
        if (k0 > 0) for (int idx = tid; idx < 6 * m; idx += PLBA_NT) {       // fold the partial sums into the panel, one thread per entry
            const int rr = idx / 6, c = idx - 6 * rr;
            double sum = 0.0;
            for (int sp = 0; sp < nsplit; sp++) sum += part[((size_t)sp * mpad + rr) * 6 + c];
            M[(size_t)(k0 + rr) * ldm + k0 + c] -= sum;
        }
    PHASE_END
    }
    PHASE_BEGIN
        const int r = k0 + 6 + tid;              // rows below the diagonal block; thread 0 factors the block even when no row is left
        if (r <= n || tid == 0) {
            double L[21], inv[6];
#pragma unroll
            for (int i = 0; i < 6; i++) {
#pragma unroll
                for (int j = 0; j <= i; j++) {
                    L[i * (i + 1) / 2 + j] = M[(size_t)(k0 + i) * ldm + k0 + j];
                }
            }
            bool bad = false;
#pragma unroll
            for (int j = 0; j < 6; j++) {
                double sd = L[j * (j + 1) / 2 + j];
#pragma unroll
                for (int k = 0; k < j; k++) sd -= L[j * (j + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
                if (!(sd > 0.0) || !plba_isfinite(sd)) { bad = true; sd = 1.0; }
                inv[j] = plba_rsqrt(sd);
                L[j * (j + 1) / 2 + j] = sd * inv[j];
#pragma unroll
                for (int i = j + 1; i < 6; i++) {
                    double v = L[i * (i + 1) / 2 + j];
#pragma unroll
                    for (int k = 0; k < j; k++) v -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
                    L[i * (i + 1) / 2 + j] = v * inv[j];
                }
            }
            if (r <= n) {
                double x[6];
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    double v = M[(size_t)r * ldm + k0 + c];
#pragma unroll
                    for (int k = 0; k < c; k++) v -= x[k] * L[c * (c + 1) / 2 + k];
                    x[c] = v * inv[c];
                }
#pragma unroll
                for (int c = 0; c < 6; c++) M[(size_t)r * ldm + k0 + c] = x[c];
            }
            if (tid == 0) {
                if (bad) *fail = 1;
#pragma unroll
                for (int c = 0; c < 6; c++) dinv[k0 + c] = inv[c];
#pragma unroll
                for (int i = 0; i < 21; i++) Lblk[kb * 21 + i] = L[i];
            }
        }
    PHASE_END
}
}

// ---- k_solve_small: block Cholesky with the matrix in REGISTERS --------------------------------------------------------------
// One thread per lower 6x6 block (bi >= bj) of the reduced camera system: the block stays in the thread's registers from the load
// to the backward substitution, so the factorisation never re-writes a trailing matrix in shared memory.  Right-looking, one pose
// block (6 columns) per panel, two block barriers per panel:
//   A(k): the threads of block column k solve their block against L_kk (X = M L_kk^-T) and publish it in shared memory; the
//         diagonal thread forwards the right-hand side it carries (y_k = L_kk^-1 g_k);
//   B(k): every thread right of the panel applies M_ij -= X_i X_j^T (216 FMAs on registers, 72 shared-memory loads of which 36 are
//         broadcasts); the diagonal threads also update their g_j; the diagonal thread of column k+1 factors its 6x6 block at once
//         (look-ahead: L_{k+1,k+1} is ready when phase A(k+1) starts).
// Threads are ordered by block column, so whole warps retire as the factorisation advances.  Backward substitution: one barrier per
// block row, x_i solved redundantly by the threads of row i.  A thread group = 32 * ceil(pairs / 32) threads handles one window;
// a 320-thread CTA holds floor(320 / group) windows (n = 60: 5 windows, n = 120: 1 window, n = 144: 300 blocks).
// The kernel also runs the hand-LM pre-solve controller, consumes S (cleared by the next update kernel), clears g and diag(H_pp) and
// applies the pose step.
enum { SS_NT = 320, SS_XLD = 37 /* odd stride of a published 6x6 block */, SS_LALL = 27 /* L_kk (21) + 1 / diag (6) */ };
PLBA_HD int ss_group_threads(int nf) { const int off = nf * (nf - 1) / 2; return 32 + ((off > 0 ? off + 31 : 32) & ~31); }   // warp 0: diagonal blocks; then the strictly lower blocks by column
PLBA_HD int ss_group_doubles(int nf) { return nf * (SS_XLD + SS_LALL + 12) + 48; }       // X panel, factors, y, x, L_cur + 1/diag, reductions
static inline size_t solve_small_smem(int = 0) { return (size_t)46 * 1024; }   // fixed (the LM-loop graph never depends on the problem); with the static parameter block it stays under the 48 KB that need no opt-in; worst case: 10 groups of nf = 7 = 46 656 bytes

// Cholesky of a 6x6 lower triangle held row-major in a 6x6 register tile (entries r >= c); inv = 1 / diagonal of L.
// The factorisation is a chain of six dependent pivots and FP64 operations have ~20 cycles of dependent-issue latency on sm_100a, so
// the chain is what a panel of k_solve_small waits for.  It is therefore run in the LDL^T form: per pivot one fast reciprocal (5
// operations), one multiply and one FMA are on the chain; the square roots that turn L D L^T into L L^T are taken off the chain.
PLBA_HD bool chol6_inplace(double *a, double *inv) {
    bool bad = false;
    double d[6];
#pragma unroll
    for (int j = 0; j < 6; j++) {
        double sd = a[j * 6 + j];
        if (!(sd > 0.0) || !plba_isfinite(sd)) { bad = true; sd = 1.0; }
        d[j] = sd;
        const double r = plba_rcp_fast(sd);
#pragma unroll
        for (int i = j + 1; i < 6; i++) {
            const double q = a[i * 6 + j] * r;
#pragma unroll
            for (int c = j + 1; c <= i; c++) a[i * 6 + c] -= q * a[c * 6 + j];
        }
    }
#pragma unroll
    for (int j = 0; j < 6; j++) {
        inv[j] = plba_rsqrt_fast(d[j]);
        a[j * 6 + j] = d[j] * inv[j];
#pragma unroll
        for (int i = j + 1; i < 6; i++) a[i * 6 + j] *= inv[j];
    }
    return !bad;
}

PLBA_KERNEL void PLBA_BOUNDS(SS_NT, 1) k_solve_small(const DevP *Pp) {
    PLBA_SMEM(raw);
    plba_launch_dependents();         // (an update kernel attached by a programmatic edge starts beside this kernel: see k_update, KF_WAIT_SOLVE)
    PLBA_PARAMS(P, Pp);
    PROF_DECL;
    const int nfmax = P.solve_nf_max > 0 ? P.solve_nf_max : 1;
    const int G = ss_group_threads(nfmax), gpc = SS_NT / G, gd = ss_group_doubles(nfmax);
    int *info = (int *)raw;                            // [gpc][4]: skip, nf, fail, -
    double *gbase = (double *)(raw + 256);
    THR_ARR(double, blk, 36); THR_ARR(double, gv, 6);
    THR_VAR(int, bi); THR_VAR(int, bj); THR_VAR(int, act);
    for (int w0 = PLBA_BID * gpc; w0 < P.n_win; w0 += PLBA_NB * gpc) {
        PHASE_BEGIN
            const int g = tid / G, t = tid - g * G, w = w0 + g;
            if (t == 0 && g < gpc) {
                int skip = 1, nf = 0;
                if (w < P.n_win) {
                    WinCtrl &ctl = P.ctrl[w];
                    if (!ctl.done && P.profile != PLBA_PROFILE_G) control_h_pre_window(P, w);
                    nf = P.win_nfree[w]; skip = (ctl.done || nf == 0) ? 1 : 0;
                }
                info[4 * g] = skip; info[4 * g + 1] = skip ? 0 : nf; info[4 * g + 2] = 0;
                double *red = gbase + (size_t)g * gd + gd - 8;
                red[0] = 0.0; red[1] = 0.0;
            }
        PHASE_END
        int nfloop = 0;
        for (int g = 0; g < gpc; g++) nfloop = info[4 * g + 1] > nfloop ? info[4 * g + 1] : nfloop;
        PHASE_BEGIN
        PHASE_END
        if (nfloop == 0) continue;
    PROF_MARK(42);
        // ---- load: thread (bi, bj) <- lower block (bi, bj) = transpose of the stored upper block (bj, bi); S, g, diag(H_pp) are consumed ----
        PHASE_BEGIN
            THR_BIND(blk); THR_BIND(gv); THR_BIND(bi); THR_BIND(bj); THR_BIND(act);
            const int g = tid / G, t = tid - g * G, w = w0 + g;
            act = 0; bi = 0; bj = 0;
            if (g < gpc && !info[4 * g]) {
                const int nf = info[4 * g + 1];
                int j = nf, rem = 0;
                if (t < 32) { if (t < nf) { j = t; rem = 0; } }                 // warp 0 of the group: the diagonal blocks
                else { j = 0; rem = t - 32; while (j < nf - 1 && rem >= nf - 1 - j) { rem -= nf - 1 - j; j++; } if (j >= nf - 1) j = nf; else rem += 1; }
                if (j < nf) {
                    act = 1; bj = j; bi = j + rem;
                    const int n = 6 * nf, slot0 = P.win_slot0[w];
                    double *Sw = P.S + P.win_S_off[w];
                    const double lambda = P.ctrl[w].lambda;
                    // 128-bit loads (rows of S are 16-byte aligned: n and 6 bi are even); S is cleared by the update kernel that runs next (clear_consumed_S)
#pragma unroll
                    for (int c = 0; c < 6; c++) {
                        const plba_d2 *row = (const plba_d2 *)(Sw + (size_t)(6 * bj + c) * n + 6 * bi);
#pragma unroll
                        for (int r2 = 0; r2 < 3; r2++) { const plba_d2 v = row[r2]; blk[(2 * r2) * 6 + c] = v.x; blk[(2 * r2 + 1) * 6 + c] = v.y; }
                    }
                    if (bi == bj) {
                        double *gs = P.gs + (size_t)6 * slot0 + 6 * bj, *hd = P.hpp_diag + (size_t)6 * slot0 + 6 * bj;
#pragma unroll
                        for (int c = 0; c < 6; c++) {
                            blk[c * 6 + c] += (P.profile == PLBA_PROFILE_G) ? lambda : lambda * hd[c];      // g2o setLambda: additive; hand LM: H_ii *= (1 + lambda)
                            gv[c] = gs[c]; gs[c] = 0.0; hd[c] = 0.0;
                        }
                        if (bj == 0) {      // look-ahead for the first panel
                            double *Lc = gbase + (size_t)g * gd + nf * (SS_XLD + SS_LALL + 12);
                            double inv[6];
                            if (!chol6_inplace(blk, inv)) info[4 * g + 2] = 1;
#pragma unroll
                            for (int r = 0; r < 6; r++) {
#pragma unroll
                                for (int c = 0; c <= r; c++) Lc[r * (r + 1) / 2 + c] = blk[r * 6 + c];
                            }
#pragma unroll
                            for (int c = 0; c < 6; c++) Lc[21 + c] = inv[c];
                        }
                    }
                }
            }
        PHASE_END
    PROF_MARK(43);
        for (int kb = 0; kb < nfloop; kb++) {
            // ---- A(kb): block column kb: X = M L_kk^-T, y_k = L_kk^-1 g_k ----
            PHASE_BEGIN
                THR_BIND(blk); THR_BIND(gv); THR_BIND(bi); THR_BIND(bj); THR_BIND(act);
                const int g = tid / G;
                if (act && bj == kb) {
                    const int nf = info[4 * g + 1];
                    double *gb = gbase + (size_t)g * gd;
                    double *Xp = gb, *Lall = gb + nf * SS_XLD, *ys = Lall + nf * SS_LALL, *Lc = ys + 12 * nf;
                    double L[21], inv[6];
#pragma unroll
                    for (int i = 0; i < 21; i++) L[i] = Lc[i];
#pragma unroll
                    for (int i = 0; i < 6; i++) inv[i] = Lc[21 + i];
                    if (bi == kb) {
                        double y[6];
#pragma unroll
                        for (int c = 0; c < 6; c++) y[c] = gv[c];
#pragma unroll
                        for (int c = 0; c < 6; c++) {         // column sweeps: two dependent operations per pivot
                            y[c] *= inv[c];
#pragma unroll
                            for (int m = c + 1; m < 6; m++) y[m] -= L[m * (m + 1) / 2 + c] * y[c];
                        }
#pragma unroll
                        for (int c = 0; c < 6; c++) ys[6 * kb + c] = y[c];
#pragma unroll
                        for (int i = 0; i < 21; i++) Lall[kb * SS_LALL + i] = L[i];
#pragma unroll
                        for (int i = 0; i < 6; i++) Lall[kb * SS_LALL + 21 + i] = inv[i];
                    } else {
#pragma unroll
                        for (int c = 0; c < 6; c++) {         // column sweeps over all six rows at once: two dependent operations per pivot
#pragma unroll
                            for (int r = 0; r < 6; r++) blk[r * 6 + c] *= inv[c];
#pragma unroll
                            for (int m = c + 1; m < 6; m++) {
#pragma unroll
                                for (int r = 0; r < 6; r++) blk[r * 6 + m] -= blk[r * 6 + c] * L[m * (m + 1) / 2 + c];
                            }
                        }
#pragma unroll
                        for (int i = 0; i < 36; i++) Xp[bi * SS_XLD + i] = blk[i];
                    }
                }
            PHASE_END
    PROF_MARK(44);
            // ---- B(kb): trailing update on registers; look-ahead factorisation of the next diagonal block ----
            PHASE_BEGIN
                THR_BIND(blk); THR_BIND(gv); THR_BIND(bi); THR_BIND(bj); THR_BIND(act);
                const int g = tid / G;
                if (act && bj > kb) {
                    const int nf = info[4 * g + 1];
                    double *gb = gbase + (size_t)g * gd;
                    const double *Xp = gb, *ys = gb + nf * (SS_XLD + SS_LALL);
                    double Xj[36];
#pragma unroll
                    for (int i = 0; i < 36; i++) Xj[i] = Xp[bj * SS_XLD + i];
                    if (bi != bj) {
#pragma unroll
                        for (int r = 0; r < 6; r++) {
                            double xi[6];
#pragma unroll
                            for (int m = 0; m < 6; m++) xi[m] = Xp[bi * SS_XLD + r * 6 + m];
#pragma unroll
                            for (int c = 0; c < 6; c++) {
                                double v = blk[r * 6 + c];
#pragma unroll
                                for (int m = 0; m < 6; m++) v -= xi[m] * Xj[c * 6 + m];
                                blk[r * 6 + c] = v;
                            }
                        }
                    } else {
#pragma unroll
                        for (int r = 0; r < 6; r++) {       // diagonal block: its lower triangle only
#pragma unroll
                            for (int c = 0; c <= r; c++) {
                                double v = blk[r * 6 + c];
#pragma unroll
                                for (int m = 0; m < 6; m++) v -= Xj[r * 6 + m] * Xj[c * 6 + m];
                                blk[r * 6 + c] = v;
                            }
                        }
#pragma unroll
                        for (int r = 0; r < 6; r++) {
                            double v = gv[r];
#pragma unroll
                            for (int m = 0; m < 6; m++) v -= Xj[r * 6 + m] * ys[6 * kb + m];
                            gv[r] = v;
                        }
                        if (bj == kb + 1) {
                            double *Lc = gb + nf * (SS_XLD + SS_LALL + 12);
                            double inv[6];
                            if (!chol6_inplace(blk, inv)) info[4 * g + 2] = 1;
#pragma unroll
                            for (int r = 0; r < 6; r++) {
#pragma unroll
                                for (int c = 0; c <= r; c++) Lc[r * (r + 1) / 2 + c] = blk[r * 6 + c];
                            }
#pragma unroll
                            for (int c = 0; c < 6; c++) Lc[21 + c] = inv[c];
                        }
                    }
                }
            PHASE_END
    PROF_MARK(45);
        }
    PROF_MARK(46);
        // ---- backward substitution L^T x = y: block row i at a time, x_i solved redundantly by the threads of the row ----
        for (int i = nfloop - 1; i >= 0; i--) {
            PHASE_BEGIN
                THR_BIND(blk); THR_BIND(bi); THR_BIND(bj); THR_BIND(act);
                const int g = tid / G;
                if (act && bi == i) {
                    const int nf = info[4 * g + 1];
                    double *gb = gbase + (size_t)g * gd;
                    const double *Lall = gb + nf * SS_XLD;
                    double *ys = gb + nf * (SS_XLD + SS_LALL), *xs = ys + 6 * nf;
                    double x[6];
#pragma unroll
                    for (int c = 0; c < 6; c++) x[c] = ys[6 * i + c];
#pragma unroll
                    for (int c = 5; c >= 0; c--) {          // column sweeps: two dependent operations per pivot
                        x[c] *= Lall[i * SS_LALL + 21 + c];
#pragma unroll
                        for (int m = 0; m < c; m++) x[m] -= Lall[i * SS_LALL + c * (c + 1) / 2 + m] * x[c];
                    }
                    if (bj == i) {
#pragma unroll
                        for (int c = 0; c < 6; c++) xs[6 * i + c] = x[c];
                    } else {
#pragma unroll
                        for (int c = 0; c < 6; c++) {
                            double v = ys[6 * bj + c];
#pragma unroll
                            for (int r = 0; r < 6; r++) v -= blk[r * 6 + c] * x[r];
                            ys[6 * bj + c] = v;
                        }
                    }
                }
            PHASE_END
        }
    PROF_MARK(47);
        PHASE_BEGIN
            const int g = tid / G, t = tid - g * G, w = w0 + g;
            if (g < gpc && !info[4 * g]) {
                const int nf = info[4 * g + 1], f = info[4 * g + 2], slot0 = P.win_slot0[w];
                const double *xs = gbase + (size_t)g * gd + nf * (SS_XLD + SS_LALL) + 6 * nf;
                for (int i = t; i < 6 * nf; i += G) P.xp[(size_t)6 * slot0 + i] = f ? 0.0 : xs[i];
                if (t == 0 && f) P.ctrl[w].solve_fail = 1;
            }
        PHASE_END
        if (PLBA_NB == 1) {           // single window: x_p is out — the update kernel running beside this one can go on (the trial poses follow: flag value 2)
            PHASE_BEGIN
                if (tid == 0) plba_set_flag(&P.counters[CNT_SOLVE_DONE], 1);
            PHASE_END
        }
        PHASE_BEGIN
            const int g = tid / G, t = tid - g * G, w = w0 + g;
            double sc = 0.0, d2 = 0.0;
            const bool on = (g < gpc && !info[4 * g]);
            if (on && t < info[4 * g + 1]) pose_update_slot(P, P.win_slot0[w] + t, sc, d2);
            double *red = gbase + (size_t)(g < gpc ? g : 0) * gd + gd - 8;
            plba_block_add(&red[0], on ? sc : 0.0);
            plba_block_add(&red[1], on ? d2 : 0.0);
        PHASE_END
        PHASE_BEGIN
            const int g = tid / G, t = tid - g * G, w = w0 + g;
            if (t == 0 && g < gpc && !info[4 * g]) { const double *red = gbase + (size_t)g * gd + gd - 8; P.ctrl[w].scale_pose = red[0]; P.ctrl[w].dx2_pose = red[1]; }
        PHASE_END
    PROF_MARK(50);
    }
    if (PLBA_NB == 1) {               // single window: x_p, the trial poses and the controller inputs are published
        PHASE_BEGIN
        PHASE_END
        PHASE_BEGIN
            if (tid == 0) plba_set_flag(&P.counters[CNT_SOLVE_DONE], 2);
        PHASE_END
    }
}

// ---- tiled path (6 Nkf > 144: BASELINE configs 4-5) ---------------------------------------------------------------------
// Right-looking block Cholesky S = U^T U on the dense upper-triangular storage in HBM, factor block NBK = 96 (16 pose blocks):
//   k_potrf_block : diagonal block (k0,k0) -> shared memory, panel Cholesky above, U_kk written back
//   k_trsm_block  : U_kj = U_kk^-T A_kj for every column right of the block, and y_k = U_kk^-T g_k; one thread per column,
//                   six accumulators per thread against broadcast rows of U_kk^T (same inner loop as the panel update)
//   k_rhs_update  : g_j -= U_kj^T y_k  (the right-hand side rides along the factorisation: forward substitution for free)
//   k_syrk_dmma   : trailing update A_ij -= U_ki^T U_kj on 128 x 128 tiles with FP64 tensor-core MMAs
//                   (mma.sync.m8n8k4.f64: the only dense contraction of the path — north star: "FP64 DMMA only for the dense
//                   reduced-camera-system Cholesky"); K = 96 streamed through shared memory in cp.async double-buffered chunks of 32
//   k_back_block  : backward substitution U x = y, one block row at a time
enum { NBK = 96 };

#ifndef PLBA_HOST_EMU
// Ampere-style asynchronous copies global -> shared (fire and forget: a thread can have dozens in flight without holding registers)
PLBA_D void plba_cp_async16(void *smem_dst, const void *gsrc, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 16 : 0;                    // src-size 0: the 16 bytes are zero-filled
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}
PLBA_D void plba_cp_async8(void *smem_dst, const void *gsrc, bool valid) {
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    const int sz = valid ? 8 : 0;
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(sa), "l"(gsrc), "r"(sz) : "memory");
}
// TMA bulk copies (cp.async.bulk, SASS UBLKCP): one instruction moves a whole operand row global -> shared, completion is counted in
// bytes on an mbarrier — no per-thread address arithmetic, no registers, no wait groups
PLBA_D unsigned plba_smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
PLBA_D void plba_mbar_init(void *bar, unsigned count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(plba_smem_u32(bar)), "r"(count) : "memory"); }
PLBA_D void plba_mbar_init_fence() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
PLBA_D void plba_mbar_expect_tx(void *bar, unsigned bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(plba_smem_u32(bar)), "r"(bytes) : "memory"); }
PLBA_D void plba_mbar_arrive(void *bar) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(plba_smem_u32(bar)) : "memory"); }
PLBA_D void plba_mbar_wait(void *bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nWAIT_%=:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra DONE_%=;\nbra WAIT_%=;\nDONE_%=:\n}" ::"r"(plba_smem_u32(bar)), "r"(parity) : "memory");
}
PLBA_D void plba_bulk_g2s(void *smem_dst, const void *gsrc, unsigned bytes, void *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(plba_smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(plba_smem_u32(bar)) : "memory");
}
PLBA_D void plba_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> PLBA_D void plba_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
#endif

// Diagonal block of a step (96 x 96 = 16 x 16 pose blocks): the register-resident block Cholesky of k_solve_small — one thread per lower
// 6 x 6 block, which stays in that thread's registers from the load to the write-back; per block column two block barriers: A the threads
// of column kb solve against L_kk and publish their block, B every thread right of it applies M_ij -= X_i X_j^T and the diagonal thread of
// column kb + 1 factors its block at once.  (The first version ran the left-looking shared-memory panels of round 1: 44 us per step on
// the factorisation's critical path.)
enum { PB_NT = 160, PB_XLD = 37 };
static inline size_t potrf_block_smem() { return sizeof(double) * ((size_t)(NBK / 6) * PB_XLD + 32) + 16; }
PLBA_KERNEL void PLBA_BOUNDS(PB_NT, 1) k_potrf_block(const DevP *Pp, double *Sw, int n, WinCtrl *ctl, const double *hd, int additive, int k0, int nb) {
    PLBA_SMEM(raw);
    PLBA_COUNT_LAUNCH(Pp);
    double *Xp = (double *)raw, *Lc = Xp + (size_t)(NBK / 6) * PB_XLD;      // published X blocks of the current column; L_kk (21) + 1 / diag (6)
    int *fail = (int *)(Lc + 32);
    const int nbk = nb / 6;
    THR_ARR(double, blk, 36);
    THR_VAR(int, bi); THR_VAR(int, bj); THR_VAR(int, act);
    PHASE_BEGIN
        THR_BIND(blk); THR_BIND(bi); THR_BIND(bj); THR_BIND(act);
        if (tid == 0) *fail = 0;
        act = 0; bi = 0; bj = 0;
        int j = nbk, rem = 0;
        if (tid < 32) { if (tid < nbk) { j = tid; rem = 0; } }                 // warp 0: the diagonal blocks
        else { j = 0; rem = tid - 32; while (j < nbk - 1 && rem >= nbk - 1 - j) { rem -= nbk - 1 - j; j++; } if (j >= nbk - 1) j = nbk; else rem += 1; }
        if (j < nbk) {
            act = 1; bj = j; bi = j + rem;
            // lower block (bi, bj) = transpose of the stored upper block (bj, bi); 128-bit loads (n, k0, 6 bi are even)
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const plba_d2 *row = (const plba_d2 *)(Sw + (size_t)(k0 + 6 * bj + c) * n + k0 + 6 * bi);
#pragma unroll
                for (int r2 = 0; r2 < 3; r2++) { const plba_d2 v = row[r2]; blk[(2 * r2) * 6 + c] = v.x; blk[(2 * r2 + 1) * 6 + c] = v.y; }
            }
            if (bi == bj) {
                const double lambda = ctl->lambda;
#pragma unroll
                for (int c = 0; c < 6; c++) blk[c * 6 + c] += additive ? lambda : lambda * hd[k0 + 6 * bj + c];      // g2o setLambda: additive; hand LM: H_ii *= (1 + lambda)
                if (bj == 0) {
                    double inv[6];
                    if (!chol6_inplace(blk, inv)) *fail = 1;
#pragma unroll
                    for (int r = 0; r < 6; r++) {
#pragma unroll
                        for (int c = 0; c <= r; c++) Lc[r * (r + 1) / 2 + c] = blk[r * 6 + c];
                    }
#pragma unroll
                    for (int c = 0; c < 6; c++) Lc[21 + c] = inv[c];
                }
            }
        }
    PHASE_END
    for (int kb = 0; kb < nbk; kb++) {
        PHASE_BEGIN
            THR_BIND(blk); THR_BIND(bi); THR_BIND(bj); THR_BIND(act);
            if (act && bj == kb && bi > kb) {
                double L[21], inv[6];
#pragma unroll
                for (int i = 0; i < 21; i++) L[i] = Lc[i];
#pragma unroll
                for (int i = 0; i < 6; i++) inv[i] = Lc[21 + i];
#pragma unroll
                for (int c = 0; c < 6; c++) {         // X = M L_kk^-T, column sweeps over all six rows at once
#pragma unroll
                    for (int r = 0; r < 6; r++) blk[r * 6 + c] *= inv[c];
#pragma unroll
                    for (int m = c + 1; m < 6; m++) {
#pragma unroll
                        for (int r = 0; r < 6; r++) blk[r * 6 + m] -= blk[r * 6 + c] * L[m * (m + 1) / 2 + c];
                    }
                }
#pragma unroll
                for (int i = 0; i < 36; i++) Xp[bi * PB_XLD + i] = blk[i];
            }
        PHASE_END
        PHASE_BEGIN
            THR_BIND(blk); THR_BIND(bi); THR_BIND(bj); THR_BIND(act);
            if (act && bj > kb) {
                double Xj[36];
#pragma unroll
                for (int i = 0; i < 36; i++) Xj[i] = Xp[bj * PB_XLD + i];
                if (bi != bj) {
#pragma unroll
                    for (int r = 0; r < 6; r++) {
                        double xi[6];
#pragma unroll
                        for (int m = 0; m < 6; m++) xi[m] = Xp[bi * PB_XLD + r * 6 + m];
#pragma unroll
                        for (int c = 0; c < 6; c++) {
                            double v = blk[r * 6 + c];
#pragma unroll
                            for (int m = 0; m < 6; m++) v -= xi[m] * Xj[c * 6 + m];
                            blk[r * 6 + c] = v;
                        }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 6; r++) {       // diagonal block: its lower triangle only
#pragma unroll
                        for (int c = 0; c <= r; c++) {
                            double v = blk[r * 6 + c];
#pragma unroll
                            for (int m = 0; m < 6; m++) v -= Xj[r * 6 + m] * Xj[c * 6 + m];
                            blk[r * 6 + c] = v;
                        }
                    }
                    if (bj == kb + 1) {                 // look-ahead: the next column's L_kk is ready when phase A(kb + 1) starts
                        double inv[6];
                        if (!chol6_inplace(blk, inv)) *fail = 1;
#pragma unroll
                        for (int r = 0; r < 6; r++) {
#pragma unroll
                            for (int c = 0; c <= r; c++) Lc[r * (r + 1) / 2 + c] = blk[r * 6 + c];
                        }
#pragma unroll
                        for (int c = 0; c < 6; c++) Lc[21 + c] = inv[c];
                    }
                }
            }
        PHASE_END
    }
    PHASE_BEGIN
        THR_BIND(blk); THR_BIND(bi); THR_BIND(bj); THR_BIND(act);
        if (act) {                                      // U[6 bj + c][6 bi + r] = L[6 bi + r][6 bj + c]
#pragma unroll
            for (int c = 0; c < 6; c++) {
                double *row = Sw + (size_t)(k0 + 6 * bj + c) * n + k0 + 6 * bi;
                if (bi != bj) {
#pragma unroll
                    for (int r2 = 0; r2 < 3; r2++) { plba_d2 v; v.x = blk[(2 * r2) * 6 + c]; v.y = blk[(2 * r2 + 1) * 6 + c]; ((plba_d2 *)row)[r2] = v; }
                } else {
#pragma unroll
                    for (int r = 0; r < 6; r++) if (r >= c) row[r] = blk[r * 6 + c];
                }
            }
        }
        if (tid == 0 && *fail) ctl->solve_fail = 1;
    PHASE_END
}

// columns k0+nb .. n-1 of block row k, plus the right-hand side as one more column (handled by the last CTA)
// Round 2, second half: ONE THREAD PER COLUMN WITH THE WHOLE COLUMN IN REGISTERS.  The first version kept the CTA's 96 x 128 columns in shared
// memory and paid ~1.3 shared loads per FMA plus 32 block barriers (66 us per step on the factorisation's critical path).  Here a thread
// loads its 96 entries once and runs the forward substitution right-looking, one pose block (6 pivots) per loop iteration: solve the 6 x 6
// triangle, then a_r -= sum_i U[p0+i][r] x_i for every later row — independent FMAs, the U rows are BROADCAST 128-bit shared loads (one
// load per two FMAs), no barrier at all.  Register indices must be compile-time, so the column SLIDES: after each block the live rows
// move down by six registers and the loop body always works on registers 0..95 (a fully unrolled nest, 112 KB of straight-line code,
// measured 100 us: every instruction line came from L2).
enum { TRSM_COLS = 64, TRSM_LD = NBK };
static inline size_t trsm_block_smem() { return sizeof(double) * ((size_t)NBK * TRSM_LD + NBK); }
PLBA_KERNEL void PLBA_BOUNDS(TRSM_COLS, 1) k_trsm_block(const DevP *Pp, double *Sw, int n, double *g, int k0, int nb) {
    PLBA_SMEM(raw);
    PLBA_COUNT_LAUNCH(Pp);
    double *Uk = (double *)raw, *Dk = Uk + (size_t)NBK * TRSM_LD;      // Uk[q][r] = U_kk[q][r] (r > q); rows / columns >= nb: identity
    const int c0 = k0 + nb + PLBA_BID * TRSM_COLS;                     // first column of this CTA; column index n = right-hand side
    THR_ARR(double, a, NBK);
    PHASE_BEGIN
        THR_BIND(a);
        // U_kk as asynchronous 16-byte copies (72 per thread, all in flight; rows / columns past nb are zero-filled), the thread's own
        // column as 96 plain loads issued behind them: one HBM round trip for everything
#ifndef PLBA_HOST_EMU
        for (int idx = tid; idx < NBK * NBK / 2; idx += PLBA_NT) {
            const int q = idx / (NBK / 2), r = (idx - q * (NBK / 2)) * 2;
            const bool ok = q < nb && r < nb;
            plba_cp_async16(Uk + (size_t)q * TRSM_LD + r, Sw + (size_t)(k0 + (ok ? q : 0)) * n + k0 + (ok ? r : 0), ok);
        }
        plba_cp_async_commit();
#else
        for (int idx = tid; idx < NBK * NBK; idx += PLBA_NT) {
            const int q = idx / NBK, r = idx - q * NBK;
            Uk[(size_t)q * TRSM_LD + r] = (q < nb && r < nb) ? Sw[(size_t)(k0 + q) * n + k0 + r] : 0.0;
        }
#endif
        const int c = c0 + tid;
        if (c <= n) {
#pragma unroll
            for (int r = 0; r < NBK; r++) a[r] = (r < nb) ? ((c < n) ? Sw[(size_t)(k0 + r) * n + c] : g[k0 + r]) : 0.0;
        }
#ifndef PLBA_HOST_EMU
        plba_cp_async_wait<0>();
#endif
    PHASE_END
    PHASE_BEGIN
        for (int q = tid; q < NBK; q += PLBA_NT) Dk[q] = (q < nb) ? 1.0 / Uk[(size_t)q * TRSM_LD + q] : 1.0;
    PHASE_END
    PHASE_BEGIN
        THR_BIND(a);
        const int c = c0 + tid;
        if (c <= n) {
#pragma unroll 1
            for (int p0 = 0; p0 < NBK; p0 += 6) {
                double xq[6];
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    xq[i] = a[i] * Dk[p0 + i];
#pragma unroll
                    for (int j = i + 1; j < 6; j++) a[j] -= Uk[(size_t)(p0 + i) * TRSM_LD + p0 + j] * xq[i];
                }
#pragma unroll
                for (int i = 0; i < 6; i++) {
                    if (p0 + i < nb) { if (c < n) Sw[(size_t)(k0 + p0 + i) * n + c] = xq[i]; else g[k0 + p0 + i] = xq[i]; }
                }
                // live rows p0 + 6 .. 95 sit in registers 6 .. 95 - p0, in blocks of 18 (a block past the end is skipped as a whole)
#pragma unroll
                for (int b = 0; b < 5; b++) {
                    if (p0 + 6 + 18 * b < NBK) {
#pragma unroll
                        for (int i = 0; i < 6; i++) {
                            const double *ur = Uk + (size_t)(p0 + i) * TRSM_LD + p0 + 6 + 18 * b;      // even offset: 16-byte aligned
#pragma unroll
                            for (int j = 0; j < 18; j += 2) {
                                const plba_d2 u2 = (p0 + 6 + 18 * b + j < NBK) ? *(const plba_d2 *)(ur + j) : plba_d2{0.0, 0.0};
                                a[6 + 18 * b + j] -= u2.x * xq[i]; a[6 + 18 * b + j + 1] -= u2.y * xq[i];
                            }
                        }
                    }
                }
#pragma unroll
                for (int m = 0; m < NBK - 6; m++) a[m] = a[m + 6];
            }
        }
    PHASE_END
}

// g_j -= sum_r U[k0+r][j] y[k0+r]   for every column j right of the block; the block's rows are dealt over gridDim.y groups of RHS_ROWS
// (a thread's chain is 12 products instead of 96: the kernel sits on the factorisation's critical path)
enum { RHS_ROWS = 12 };
PLBA_KERNEL void k_rhs_update(const DevP *Pp, int w, int k0, int nb) {
    PLBA_PARAMS(P, Pp);
    const int n = 6 * P.win_nfree[w], slot0 = P.win_slot0[w];
    const double *Sw = P.S + P.win_S_off[w];
    double *g = P.xp + (size_t)6 * slot0;
    PHASE_BEGIN
        const int j = k0 + nb + PLBA_BID * PLBA_NT + tid, r0 = PLBA_BIDY * RHS_ROWS;
        if (j < n && r0 < nb) {
            double s0 = 0.0, s1 = 0.0, s2 = 0.0;          // nb is a multiple of 6: three independent chains, loads issued ahead
#pragma unroll
            for (int r = r0; r < r0 + RHS_ROWS; r += 3) {
                if (r < nb) { s0 += Sw[(size_t)(k0 + r) * n + j] * g[k0 + r]; s1 += Sw[(size_t)(k0 + r + 1) * n + j] * g[k0 + r + 1]; s2 += Sw[(size_t)(k0 + r + 2) * n + j] * g[k0 + r + 2]; }
            }
            plba_atomic_add(&g[j], -((s0 + s1) + s2));
        }
    PHASE_END
}
PLBA_KERNEL void k_rhs_init(const DevP *Pp, int w) {
    PLBA_PARAMS(P, Pp);
    const int n = 6 * P.win_nfree[w], slot0 = P.win_slot0[w];
    PHASE_BEGIN
        for (int i = PLBA_BID * PLBA_NT + tid; i < n; i += PLBA_NB * PLBA_NT) P.xp[(size_t)6 * slot0 + i] = P.gs[(size_t)6 * slot0 + i];
    PHASE_END
}

#ifndef PLBA_HOST_EMU
PLBA_D void plba_dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}
#endif

// Trailing update, second generation (round 2).  What the first one measured (profiles/README.md r02j): 12.7 TFLOP/s although ONE warp
// alone reaches a quarter of the chip's FP64 tensor rate — with K = 96 per visit of a C tile, the 256 KB read-modify-write of the tile and
// the pipeline fill cost more than the 24 k-steps between them, and one 8-warp CTA per SM (128 accumulator registers per thread) has
// nothing to overlap them with.  Now: 128 x 64 tiles, 32 x 32 per warp (64 accumulator registers) -> TWO CTAs per SM, so that one CTA's
// epilogue / prologue runs under the other's k-steps; K chunks of 16 rows in a four-stage ring (three chunks in flight, one block barrier
// per chunk).  Tile rows [tr0, tr1) only: the factorisation's look-ahead launches the tile row that holds the next panel on its own.
enum { STM = 128, STN = 64, SKC2 = 16, SST = 4, SLDA = STM + 4, SLDB = STN + 4, SYRK_STAGE = SKC2 * (SLDA + SLDB) };
#ifndef PLBA_SYRK_TMA
#define PLBA_SYRK_TMA 1      // operand rows by TMA bulk copies + mbarriers (0: per-thread 16-byte cp.async, the first version of this kernel)
#endif
static inline size_t syrk_dmma_smem() { return sizeof(double) * (size_t)(SST * SYRK_STAGE) + 16 * SST; }
// number of 128 x 64 tiles of tile rows [tr0, tr1) of the upper-triangular trailing matrix of order m: tile row ti needs the column tiles 2 ti .. TN - 1
static inline int syrk_tiles(int m, int tr0, int tr1) {
    const int TN = (m + STN - 1) / STN;
    int c = 0;
    for (int ti = tr0; ti < tr1; ti++) c += std::max(0, TN - 2 * ti);
    return c;
}
// A_ij -= U_ki^T U_kj for the trailing tiles of step k; lo = first trailing row / column
PLBA_KERNEL void PLBA_BOUNDS(256, 2) k_syrk_dmma(const DevP *Pp, double *Sw, int n, int k0, int nb, int lo, int tr0, int rmax) {      // rows >= rmax are left alone
    PLBA_SMEM(raw);
    PLBA_COUNT_LAUNCH(Pp);            // (matrix pointer and order are kernel arguments: no parameter block, no dependent loads before the first copy)
    const int TN = (n - lo + STN - 1) / STN;
    int p = PLBA_BID, ti = tr0;
    while (p >= TN - 2 * ti) { p -= TN - 2 * ti; ti++; }
    const int tj = 2 * ti + p;
    const int i0 = lo + ti * STM, j0 = lo + tj * STN;
#ifdef PLBA_HOST_EMU
    (void)raw;
    PHASE_BEGIN     // plain loops on the host (the MMA fragments exist only on the GPU)
        for (int idx = tid; idx < STM * STN; idx += PLBA_NT) {
            const int r = i0 + idx / STN, c = j0 + idx % STN;
            if (r >= rmax || c >= n || c < r) continue;
            double sum = 0.0;
            for (int q = 0; q < nb; q++) sum += Sw[(size_t)(k0 + q) * n + r] * Sw[(size_t)(k0 + q) * n + c];
            Sw[(size_t)r * n + c] -= sum;
        }
    PHASE_END
#else
    double *ring = (double *)raw;                              // [SST stages][SKC2 rows][SLDA of A | SLDB of B]
    const int tid = (int)threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wm = warp & 3, wn = warp >> 2;                   // 4 x 2 warps: 32 x 32 per warp
    double acc[4][4][2];
#pragma unroll
    for (int a = 0; a < 4; a++)
#pragma unroll
        for (int b = 0; b < 4; b++) { acc[a][b][0] = 0.0; acc[a][b][1] = 0.0; }
    const int nchunk = (nb + SKC2 - 1) / SKC2;
#if PLBA_SYRK_TMA
    // warp 0 is the producer: lane r < 16 moves row r of the A operand (128 doubles), lane 16 + r row r of the B operand (64 doubles) of a chunk
    // as ONE bulk copy each; lane 0 announces the chunk's byte count on the stage's FULL mbarrier first.  Every warp waits on that mbarrier,
    // and says on the stage's EMPTY mbarrier when it is done with the chunk: the producer refills a stage once all eight warps have released
    // it — NO block barrier in the k-loop (the cp.async variant needs one per chunk to make the copies of other threads visible).
    unsigned long long *full = (unsigned long long *)(ring + (size_t)SST * SYRK_STAGE), *empty = full + SST;
    if (tid == 0) {
#pragma unroll
        for (int st = 0; st < SST; st++) { plba_mbar_init(&full[st], 1); plba_mbar_init(&empty[st], 8); }
        plba_mbar_init_fence();
    }
    __syncthreads();
    const unsigned a_bytes = 8u * (unsigned)((n - i0 < STM) ? n - i0 : STM), b_bytes = 8u * (unsigned)((n - j0 < STN) ? n - j0 : STN);      // (even counts: multiples of 16 bytes)
    auto load_chunk = [&](int ch) {
        if (warp == 0 && ch < nchunk) {
            const int st = ch % SST, rows = (nb - ch * SKC2 < SKC2) ? nb - ch * SKC2 : SKC2;
            double *As = ring + (size_t)st * SYRK_STAGE, *Bs = As + SKC2 * SLDA;
            if (ch >= SST) plba_mbar_wait(&empty[st], (unsigned)((ch / SST - 1) & 1));      // every warp has finished chunk ch - SST
            const int r = lane & 15;
            if (r >= rows) {
                // rows past the end of the panel (its last chunk only) are zeros: plain stores, released to the consumers by lane 0's arrive below
                if (lane < 16) { for (int c = 0; c < STM; c++) As[(size_t)r * SLDA + c] = 0.0; }
                else { for (int c = 0; c < STN; c++) Bs[(size_t)r * SLDB + c] = 0.0; }
            }
            __syncwarp();
            if (lane == 0) plba_mbar_expect_tx(&full[st], (unsigned)rows * (a_bytes + b_bytes));
            __syncwarp();
            const double *src = Sw + (size_t)(k0 + ch * SKC2 + r) * n;
            if (r < rows) {
                if (lane < 16) plba_bulk_g2s(As + (size_t)r * SLDA, src + i0, a_bytes, &full[st]);
                else plba_bulk_g2s(Bs + (size_t)r * SLDB, src + j0, b_bytes, &full[st]);
            }
        }
    };
#else
    auto load_chunk = [&](int ch) {
        if (ch < nchunk) {
            double *As = ring + (size_t)(ch % SST) * SYRK_STAGE, *Bs = As + SKC2 * SLDA;
            // 16 rows x (128 + 64) doubles = 1 536 16-byte copies: 4 + 2 per thread
#pragma unroll
            for (int it = 0; it < 4; it++) {
                const int e = tid + it * 256, r = e >> 6, c2 = (e & 63) * 2, kr = ch * SKC2 + r, c = i0 + c2;
                const bool ok = kr < nb && c < n;
                plba_cp_async16(As + (size_t)r * SLDA + c2, Sw + (size_t)(k0 + (ok ? kr : 0)) * n + (ok ? c : 0), ok);
            }
#pragma unroll
            for (int it = 0; it < 2; it++) {
                const int e = tid + it * 256, r = e >> 5, c2 = (e & 31) * 2, kr = ch * SKC2 + r, c = j0 + c2;
                const bool ok = kr < nb && c < n;
                plba_cp_async16(Bs + (size_t)r * SLDB + c2, Sw + (size_t)(k0 + (ok ? kr : 0)) * n + (ok ? c : 0), ok);
            }
        }
        plba_cp_async_commit();                                // (an empty group past the last chunk keeps the wait count uniform)
    };
#endif
    load_chunk(0); load_chunk(1); load_chunk(2);
    // the C tile is read only in the epilogue, ~10 us from now: pull its lines into L2 meanwhile (the trailing matrix does not fit L2)
#pragma unroll
    for (int a = 0; a < 4; a++) {
        const int r = i0 + wm * 32 + a * 8 + (lane >> 2), cb = j0 + wn * 32 + 8 * (lane & 3);      // the row's 32 columns as four 64-byte pieces
        if (r < rmax && cb < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(Sw + (size_t)r * n + cb));
    }
    for (int ch = 0; ch < nchunk; ch++) {
#if PLBA_SYRK_TMA
        plba_mbar_wait(&full[ch % SST], (unsigned)((ch / SST) & 1));      // chunk ch has landed (all its bytes)
#else
        plba_cp_async_wait<2>();                               // chunk ch has landed (this thread's copies); the barrier makes it everyone's
        __syncthreads();                                       // every warp is done with chunk ch - 1, whose stage is refilled now
        load_chunk(ch + 3);
#endif
        const double *Ac = ring + (size_t)(ch % SST) * SYRK_STAGE, *Bc = Ac + SKC2 * SLDA;
#pragma unroll
        for (int kk = 0; kk < SKC2; kk += 4) {
            double af[4], bf[4];
            const double *ar = Ac + (size_t)(kk + (lane & 3)) * SLDA + wm * 32 + (lane >> 2);
            const double *br = Bc + (size_t)(kk + (lane & 3)) * SLDB + wn * 32 + (lane >> 2);
#pragma unroll
            for (int a = 0; a < 4; a++) af[a] = ar[a * 8];
#pragma unroll
            for (int b = 0; b < 4; b++) bf[b] = br[b * 8];
#pragma unroll
            for (int a = 0; a < 4; a++)
#pragma unroll
                for (int b = 0; b < 4; b++) plba_dmma(acc[a][b][0], acc[a][b][1], af[a], bf[b]);
        }
#if PLBA_SYRK_TMA
        __syncwarp();
        if (lane == 0) plba_mbar_arrive(&empty[ch % SST]);     // this warp is done with the stage
        load_chunk(ch + 3);                                    // (producer warp: refill the stage of chunk ch - 1 once all warps have released it)
#endif
    }
    // epilogue: C -= acc (upper triangle only); two rows of four fragments (eight 16-byte loads) are in flight per thread
#pragma unroll
    for (int a2 = 0; a2 < 4; a2 += 2) {
        plba_d2 v[2][4];
        bool fast[2];
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int r = i0 + wm * 32 + (a2 + h) * 8 + (lane >> 2), cb = j0 + wn * 32 + 2 * (lane & 3);
            fast[h] = (r < rmax && cb >= r && cb + 24 + 1 < n);      // the common case: the lane's four 16-byte pieces lie inside the upper triangle
            if (fast[h]) {
#pragma unroll
                for (int b = 0; b < 4; b++) v[h][b] = *(const plba_d2 *)(Sw + (size_t)r * n + cb + b * 8);
            }
        }
#pragma unroll
        for (int h = 0; h < 2; h++) {
            const int a = a2 + h, r = i0 + wm * 32 + a * 8 + (lane >> 2), cb = j0 + wn * 32 + 2 * (lane & 3);
            if (r >= rmax) continue;
            double *row = Sw + (size_t)r * n;
            if (fast[h]) {
#pragma unroll
                for (int b = 0; b < 4; b++) { v[h][b].x -= acc[a][b][0]; v[h][b].y -= acc[a][b][1]; *(plba_d2 *)(row + cb + b * 8) = v[h][b]; }
            } else {
#pragma unroll
                for (int b = 0; b < 4; b++) {
                    const int c = cb + b * 8;
                    if (c < n && c >= r) row[c] -= acc[a][b][0];
                    if (c + 1 < n && c + 1 >= r) row[c + 1] -= acc[a][b][1];
                }
            }
        }
    }
#endif
}

// backward substitution U_kk x_k = y_k for block k (blocks are visited last to first; x overwrites y in xp; k_back_update has already
// taken the blocks right of k out of y_k).  The CTA loads the triangle (six loads in flight per thread), then ONE WARP solves it: lane l
// keeps y of rows l, l + 32, l + 64 in registers; column by column the owner publishes x_c and every lane takes it out of its rows
// (column reads of the odd-stride triangle are conflict-free): ~100 cycles per unknown, no block barrier inside the solve.
enum { BB_LD = NBK + 1 };
static inline size_t back_block_smem() { return sizeof(double) * ((size_t)NBK * BB_LD + 2 * NBK); }
PLBA_KERNEL void PLBA_BOUNDS(256, 1) k_back_block(const DevP *Pp, const double *Sw, int n, double *x, const WinCtrl *ctl, int k0, int nb) {
    PLBA_SMEM(raw);
    PLBA_COUNT_LAUNCH(Pp);
    double *Uk = (double *)raw, *Dk = Uk + (size_t)NBK * BB_LD, *xs = Dk + NBK;
    PHASE_BEGIN
#ifndef PLBA_HOST_EMU
        for (int idx = tid; idx < NBK * NBK; idx += PLBA_NT) {
            const int r = idx / NBK, c = idx - r * NBK;
            const bool ok = c >= r && c < nb;
            if (c >= r) plba_cp_async8(&Uk[(size_t)r * BB_LD + c], &Sw[(size_t)(k0 + (ok ? r : 0)) * n + k0 + (ok ? c : 0)], ok);
        }
        plba_cp_async_commit(); plba_cp_async_wait<0>();
#else
        for (int idx = tid; idx < NBK * NBK; idx += PLBA_NT) {
            const int r = idx / NBK, c = idx - r * NBK;
            if (c >= r) Uk[(size_t)r * BB_LD + c] = (c < nb) ? Sw[(size_t)(k0 + r) * n + k0 + c] : 0.0;
        }
#endif
    PHASE_END
    PHASE_BEGIN
        for (int r = tid; r < NBK; r += PLBA_NT) Dk[r] = (r < nb) ? 1.0 / Uk[(size_t)r * BB_LD + r] : 1.0;
    PHASE_END
    if (PLBA_WARP_IN_CTA == 0) {
        LANE_ARR(double, y, 3);
        WPHASE_BEGIN
            LANE_BIND(y);
#pragma unroll
            for (int j = 0; j < 3; j++) y[j] = (lane + 32 * j < nb) ? x[k0 + lane + 32 * j] : 0.0;
        WPHASE_END
#pragma unroll
        for (int c = NBK - 1; c >= 0; c--) {
            WPHASE_BEGIN
                LANE_BIND(y);
                if (lane == (c & 31)) xs[c] = y[c >> 5] * Dk[c];
            WPHASE_END
            WPHASE_BEGIN
                LANE_BIND(y);
                const double xc = xs[c];
#pragma unroll
                for (int j = 0; j < 3; j++) { if (lane + 32 * j < c) y[j] -= Uk[(size_t)(lane + 32 * j) * BB_LD + c] * xc; }
            WPHASE_END
        }
    }
    PHASE_BEGIN
    PHASE_END
    PHASE_BEGIN
        if (tid < nb) x[k0 + tid] = ctl->solve_fail ? 0.0 : xs[tid];
    PHASE_END
}

// Column-oriented half of the backward substitution: once block k is solved, y_i -= U[i][block k] x_k for EVERY row i above it — rows
// are independent (two threads per row, 48 contiguous columns each, 128-bit loads), so the whole machine reads the 96-column strip at
// once instead of one CTA reading a 96 x (n - k0) block row (round 2, first half: up to 370 us per step, 25.8 ms per solve at n = 12 000).
enum { BU_ROWS = 128 };
PLBA_KERNEL void PLBA_BOUNDS(2 * BU_ROWS, 1) k_back_update(const DevP *Pp, const double *Sw, int n, double *x, int k0, int nb) {
    PLBA_COUNT_LAUNCH(Pp);
    PLBA_SHARED double xs[NBK], part[2 * BU_ROWS];
    PHASE_BEGIN
        if (tid < NBK) xs[tid] = (tid < nb) ? x[k0 + tid] : 0.0;
    PHASE_END
    PHASE_BEGIN
        const int i = PLBA_BID * BU_ROWS + (tid >> 1), hf = tid & 1, c0 = hf * (NBK / 2);
        double s0 = 0.0, s1 = 0.0;
        if (i < k0) {
            const double *row = Sw + (size_t)i * n + k0 + c0;            // (k0 and n are even: 16-byte aligned)
#pragma unroll 8
            for (int c = 0; c < NBK / 2; c += 2) {
                if (c0 + c < nb) { const plba_d2 v = *(const plba_d2 *)(row + c); s0 += v.x * xs[c0 + c]; s1 += v.y * xs[c0 + c + 1]; }
            }
        }
        part[tid] = s0 + s1;
    PHASE_END
    PHASE_BEGIN
        const int i = PLBA_BID * BU_ROWS + tid;
        if (tid < BU_ROWS && i < k0) x[i] -= part[2 * tid] + part[2 * tid + 1];
    PHASE_END
}

// ---- banded path (6 Nkf > 144 and every landmark track shorter than BAND_MAX keyframes: the sliding-window shape of configs 4-5) ----
// The reduced camera system of a window without loop closures is block-banded (block (a,b) is non-zero only if some landmark is
// seen by free keyframes a and b), and Cholesky preserves the band.  One CTA per window walks the 6-column panels left-looking,
// exactly like chol_lower_panels, but every row only looks back over its band and the factor lives in a shared-memory RING of the
// last QB columns x RB rows; S is read once (band only), L is written back in place for the backward substitution.
// O(n * band^2) instead of O(n^3 / 3): n = 12 000, band 14 blocks -> 0.1 GFLOP instead of 576.
enum { BAND_MAX = 15 /* blocks */, BAND_RB = 6 * (BAND_MAX + 1), BAND_QB = 6 * BAND_MAX, BAND_LDR = BAND_RB + 1 };
static inline size_t solve_banded_smem() { return sizeof(double) * ((size_t)BAND_QB * BAND_LDR + BAND_QB + BAND_RB + 2 * (BAND_RB + 2) * 6 + 256 * 6 + 64) + 64; }

PLBA_KERNEL void k_solve_banded(const DevP *Pp, int bwb) {
    PLBA_SMEM(raw);
    PLBA_PARAMS(P, Pp);
    for (int w = PLBA_BID; w < P.n_win; w += PLBA_NB) {
        WinCtrl &ctl = P.ctrl[w];
        const int nf = P.win_nfree[w], n = 6 * nf, slot0 = P.win_slot0[w];
        if (ctl.done || n == 0) continue;                     // (profile H: k_control_h_pre has already run)
        const int RB = 6 * (bwb + 1), QB = 6 * bwb, ldr = BAND_LDR;
        double *ring = (double *)raw;                          // ring[(q % QB) * ldr + (r % RB)] = L[r][q]
        double *yring = ring + (size_t)BAND_QB * BAND_LDR;     // y[q % QB]   (the right-hand side rides along as one more row)
        double *xring = yring + BAND_QB;                       // backward substitution: x[r % RB]
        double *Pn0 = xring + BAND_RB;                         // two panel buffers [RB + 1][6] (row nr = right-hand side): the raw entries of
        double *Pn1 = Pn0 + (BAND_RB + 2) * 6;                 //   panel kb+1 are staged while panel kb is factored
        double *part = Pn1 + (BAND_RB + 2) * 6;                // [<= 256][6] partial sums
        double *Lb = part + 256 * 6;                           // [0..20] diagonal block staged for the backward pass, [24..29] rhs, [40] fail flag
        double *Sw = P.S + P.win_S_off[w];
        double *x = P.xp + (size_t)6 * slot0;
        const double lambda = ctl.lambda;
        PHASE_BEGIN
            if (tid == 0) Lb[40] = 0.0;
            const int nr0 = (n < RB) ? n : RB;
            for (int idx = tid; idx < 6 * (nr0 + 1); idx += PLBA_NT) {
                const int c = idx / (nr0 + 1), rr = idx - c * (nr0 + 1);
                double v;
                if (rr == nr0) v = P.gs[(size_t)6 * slot0 + c];
                else if (rr >= c) { v = Sw[(size_t)c * n + rr]; if (rr == c) v += (P.profile == PLBA_PROFILE_G) ? lambda : lambda * P.hpp_diag[(size_t)6 * slot0 + rr]; }
                else v = 0.0;
                Pn0[rr * 6 + c] = v;
            }
        PHASE_END
        for (int kb = 0; kb < nf; kb++) {
            double *Pn = (kb & 1) ? Pn1 : Pn0, *Pnext = (kb & 1) ? Pn0 : Pn1;
            const int k0 = 6 * kb;
            const int nr = (n - k0 < RB) ? n - k0 : RB;       // band rows of this panel: r = k0 + rr, rr < nr; rr == nr is the rhs row
            const int m = nr + 1;
            const int mpad = (m + 31) & ~31, nsplit = (mpad <= 32) ? 8 : (mpad <= 64) ? 4 : (mpad <= 128) ? 2 : 1;
            const int qlo_panel = (kb > bwb) ? 6 * (kb - bwb) : 0;
            const int kslot = k0 % RB;                         // RB and k0 are multiples of 6: rows k0 .. k0+5 sit at kslot .. kslot+5
            PHASE_BEGIN
                const int sp = tid / mpad, rr = tid - sp * mpad;
                if (k0 > 0 && sp < nsplit && rr < m) {
                    const bool rhs = (rr == nr);
                    const int r = k0 + rr;
                    const int rblk = r / 6;
                    const int qlo = rhs ? qlo_panel : ((rblk > bwb) ? 6 * (rblk - bwb) : 0);      // first column of row r's band
                    const int len = k0 - qlo, qs = (len + nsplit - 1) / nsplit;
                    const int q0 = qlo + sp * qs, q1 = (q0 + qs < k0) ? q0 + qs : k0;
                    double acc[6] = {0, 0, 0, 0, 0, 0};
                    int rslot = kslot + rr; if (rslot >= RB) rslot -= RB;
                    // the ring wraps at most once inside [q0, q1): two straight runs, each unrollable
                    int qc = (q0 < q1) ? q0 % QB : 0, left = q1 - q0;
                    while (left > 0) {
                        const int run = (QB - qc < left) ? QB - qc : left;
                        const double *col = ring + (size_t)qc * ldr;
                        const double *src = rhs ? yring + qc : col + rslot;
                        const int sstride = rhs ? 1 : ldr;
#pragma unroll 4
                        for (int t = 0; t < run; t++) {
                            const double v = src[(size_t)t * sstride];
#pragma unroll
                            for (int c = 0; c < 6; c++) acc[c] += v * col[(size_t)t * ldr + kslot + c];
                        }
                        left -= run; qc = 0;
                    }
#pragma unroll
                    for (int c = 0; c < 6; c++) part[((size_t)sp * mpad + rr) * 6 + c] = acc[c];
                }
            PHASE_END
            PHASE_BEGIN
                if (k0 > 0) for (int idx = tid; idx < 6 * m; idx += PLBA_NT) {       // fold the partial sums into the staged raw entries
                    const int rr = idx / 6, c = idx - 6 * rr;
                    double sum = 0.0;
                    for (int sp = 0; sp < nsplit; sp++) sum += part[((size_t)sp * mpad + rr) * 6 + c];
                    Pn[rr * 6 + c] -= sum;
                }
            PHASE_END
            PHASE_BEGIN
                const int rr = 6 + tid;                        // rows below the diagonal block (incl. the rhs row at rr == nr)
                if (tid >= 128) {
                    // the upper half of the CTA is idle while the lower half factors: it stages the raw entries of the NEXT panel
                    if (kb + 1 < nf) {
                        const int k1 = k0 + 6, nr1 = (n - k1 < RB) ? n - k1 : RB;
                        for (int idx = tid - 128; idx < 6 * (nr1 + 1); idx += PLBA_NT - 128) {
                            const int c = idx / (nr1 + 1), r1 = idx - c * (nr1 + 1);
                            double v;
                            if (r1 == nr1) v = P.gs[(size_t)6 * slot0 + k1 + c];
                            else if (r1 >= c) { v = Sw[(size_t)(k1 + c) * n + k1 + r1]; if (r1 == c) v += (P.profile == PLBA_PROFILE_G) ? lambda : lambda * P.hpp_diag[(size_t)6 * slot0 + k1 + r1]; }
                            else v = 0.0;
                            Pnext[r1 * 6 + c] = v;
                        }
                    }
                } else if (rr <= nr || tid == 0) {
                    double L[21], inv[6];
#pragma unroll
                    for (int i = 0; i < 6; i++) {
#pragma unroll
                        for (int j = 0; j <= i; j++) L[i * (i + 1) / 2 + j] = Pn[i * 6 + j];
                    }
                    bool bad = false;
#pragma unroll
                    for (int j = 0; j < 6; j++) {
                        double sd = L[j * (j + 1) / 2 + j];
#pragma unroll
                        for (int k = 0; k < j; k++) sd -= L[j * (j + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
                        if (!(sd > 0.0) || !plba_isfinite(sd)) { bad = true; sd = 1.0; }
                        inv[j] = plba_rsqrt(sd);
                        L[j * (j + 1) / 2 + j] = sd * inv[j];
#pragma unroll
                        for (int i = j + 1; i < 6; i++) {
                            double v = L[i * (i + 1) / 2 + j];
#pragma unroll
                            for (int k = 0; k < j; k++) v -= L[i * (i + 1) / 2 + k] * L[j * (j + 1) / 2 + k];
                            L[i * (i + 1) / 2 + j] = v * inv[j];
                        }
                    }
                    if (rr <= nr) {
                        double xr[6];
#pragma unroll
                        for (int c = 0; c < 6; c++) {
                            double v = Pn[rr * 6 + c];
#pragma unroll
                            for (int k = 0; k < c; k++) v -= xr[k] * L[c * (c + 1) / 2 + k];
                            xr[c] = v * inv[c];
                        }
                        // publish: ring for the panels to come, global (in place, upper storage) for the backward substitution
                        const int qc0 = (QB > 0) ? k0 % QB : 0;         // QB is a multiple of 6: columns k0 .. k0+5 sit at qc0 .. qc0+5
                        if (rr == nr) {
#pragma unroll
                            for (int c = 0; c < 6; c++) { if (QB > 0) yring[qc0 + c] = xr[c]; x[k0 + c] = xr[c]; }
                        } else {
                            int rslot = kslot + rr; if (rslot >= RB) rslot -= RB;
#pragma unroll
                            for (int c = 0; c < 6; c++) { if (QB > 0) ring[(size_t)(qc0 + c) * ldr + rslot] = xr[c]; Sw[(size_t)(k0 + c) * n + k0 + rr] = xr[c]; }
                        }
                    }
                    if (tid == 0) {
                        if (bad) Lb[40] = 1.0;
#pragma unroll
                        for (int i = 0; i < 6; i++) {
#pragma unroll
                            for (int j = 0; j <= i; j++) Sw[(size_t)(k0 + j) * n + k0 + i] = L[i * (i + 1) / 2 + j];     // factored diagonal block, upper storage
                        }
                    }
                }
            PHASE_END
        }
        // backward substitution L^T x = y, last panel first: x_k = L_kk^-T (y_k - sum over the band rows below of L[r][k-cols] x_r)
        for (int kb = nf - 1; kb >= 0; kb--) {
            const int k0 = 6 * kb;
            const int nr = (n - k0 < RB) ? n - k0 : RB;
            const int kslot = k0 % RB;
            PHASE_BEGIN
                // 6 columns x 32 lanes: lane-strided partial dot products over the rows below the diagonal block (contiguous in memory);
                // the seventh warp stages the factored diagonal block and the right-hand side
                const int c = tid >> 5, lane = tid & 31;
                if (c < 6) {
                    double sum = 0.0;
                    const double *col = Sw + (size_t)(k0 + c) * n + k0;
                    for (int rr = 6 + lane; rr < nr; rr += 32) { int xs_ = kslot + rr; if (xs_ >= RB) xs_ -= RB; sum += col[rr] * xring[xs_]; }
#ifndef PLBA_HOST_EMU
                    for (int o = 16; o > 0; o >>= 1) sum += __shfl_down_sync(0xffffffffu, sum, o);
                    if (lane == 0) part[c * 33] = sum;
#else
                    part[c * 33 + lane] = sum;
#endif
                } else if (c == 6) {
                    if (lane < 21) {
                        int i = 0; while ((i + 1) * (i + 2) / 2 <= lane) i++;
                        const int j = lane - i * (i + 1) / 2;
                        Lb[lane] = Sw[(size_t)(k0 + j) * n + k0 + i];
                    } else if (lane < 27) Lb[24 + lane - 21] = x[k0 + lane - 21];
                }
            PHASE_END
            PHASE_BEGIN
                if (tid == 0) {
                    double xs[6];
#pragma unroll
                    for (int c = 5; c >= 0; c--) {
                        double v = Lb[24 + c];
#ifndef PLBA_HOST_EMU
                        v -= part[c * 33];
#else
                        for (int l = 0; l < 32; l++) v -= part[c * 33 + l];
#endif
#pragma unroll
                        for (int mm = c + 1; mm < 6; mm++) v -= Lb[mm * (mm + 1) / 2 + c] * xs[mm];
                        xs[c] = v / Lb[c * (c + 1) / 2 + c];
                    }
                    const bool fail = (Lb[40] != 0.0);
#pragma unroll
                    for (int c = 0; c < 6; c++) { xring[kslot + c] = xs[c]; x[k0 + c] = fail ? 0.0 : xs[c]; }
                }
            PHASE_END
        }
        PHASE_BEGIN
            if (tid == 0 && Lb[40] != 0.0) ctl.solve_fail = 1;
        PHASE_END
    }
}

// ---- block cyclic reduction for block-banded reduced camera systems (6 Nkf > 144, sliding-window shape) -------------------------
// With half bandwidth `band` pose blocks the reduced camera system is block tridiagonal in NODES of bs >= band keyframes
// (m = 6 bs <= 90 unknowns).  Level l (stride s = 2^l) eliminates the nodes i = s, 3s, 5s, ... IN PARALLEL, one CTA each:
//     D_i = L L^T ;  Xl = A[i-s,i] L^-T ;  Xr = A[i+s,i] L^-T ;  y = b_i L^-T          (one shared-memory Cholesky: the coupling
//                                                                                      blocks and b ride along as extra rows)
//     D_{i-s} -= Xl Xl^T ;  D_{i+s} -= Xr Xr^T ;  A[i-s,i+s] = -Xl Xr^T ;  b_{i-s} -= Xl y ;  b_{i+s} -= Xr y
// which leaves a block tridiagonal system in the nodes that are multiples of 2s.  After ceil(log2 N) levels node 0 is solved and
// the levels are walked back:  L^T x_i = y - Xl^T x_{i-s} - Xr^T x_{i+s}.
// This is a Cholesky factorisation in a nested-dissection order of the keyframe chain: same system, same solution up to rounding,
// but the 6 Nkf sequential pivots of the banded factorisation become 90 ceil(log2 N) and every level fills the GPU.
struct BcrW { double *D, *U, *b, *hd, *Xl, *Xr, *y, *Lf; int N, bs, m, pad; };   // Lf: factors of the split elimination (k_bcr_factor: the two CTAs of a node both read D_i, so L cannot overwrite it)   // node storage: D, U, Xl, Xr: [N][m*m] (row stride m); b, hd, y: [N][m];
                                                                           // D (upper triangles), U: the window's slice of P.S (the assembly kernels accumulate into the node form directly, s_block()); b = the window's slice of g, hd of diag(H_pp)
enum { BCR_BS_MAX = 15, BCR_M_MAX = 6 * BCR_BS_MAX, BCR_NT = 512 };
static inline size_t bcr_elim_smem() { return sizeof(double) * ((size_t)(3 * BCR_M_MAX + 1) * (BCR_M_MAX + 1) + BCR_M_MAX + 21 * BCR_BS_MAX + 6 * 264 + 8) + 64; }
static inline size_t bcr_back_smem() { return sizeof(double) * ((size_t)BCR_M_MAX * (BCR_M_MAX + 1) + 11 * BCR_M_MAX + 8) + 64; }
PLBA_HD int bcr_node_size(const BcrW &B, int nf, int i) { const int k0 = i * B.bs, k1 = (k0 + B.bs < nf) ? k0 + B.bs : nf; return 6 * (k1 - k0); }

// eliminate the nodes (2 j + 1) s of one level (final != 0: node 0, nothing left to couple to)
PLBA_KERNEL void k_bcr_elim(const DevP *Pp, int w, BcrW B, int s, int final) {
    PLBA_SMEM(raw);
    PLBA_PARAMS(P, Pp);
    WinCtrl &ctl = P.ctrl[w];
    if (ctl.done) return;
    const int nf = P.win_nfree[w], m = B.m, ldm = m + 1;
    const int i = final ? 0 : (2 * PLBA_BID + 1) * s;
    const int left = final ? -1 : i - s, right = (final || i + s >= B.N) ? -1 : i + s;
    const int mi = bcr_node_size(B, nf, i), ml = left >= 0 ? m : 0, mr = right >= 0 ? bcr_node_size(B, nf, right) : 0;
    const int nr = mi + ml + mr;                                    // row nr = right-hand side
    double *M = (double *)raw, *dinv = M + (size_t)(3 * BCR_M_MAX + 1) * (BCR_M_MAX + 1), *Lblk = dinv + BCR_M_MAX, *part = Lblk + 21 * BCR_BS_MAX;
    int *fail = (int *)(part + 6 * 264);
    double *Di = B.D + (size_t)i * m * m, *Ui = B.U + (size_t)i * m * m;
    double *Ur = right >= 0 ? B.U + (size_t)right * m * m : nullptr;
    PHASE_BEGIN
        if (tid == 0) *fail = 0;
        for (int idx = tid; idx < mi * mi; idx += PLBA_NT) {
            const int r = idx / mi, c = idx - r * mi;
            if (c > r) continue;
            double v = Di[(size_t)c * m + r];                          // D is stored as an upper triangle (s_block): entry (r, c), c <= r, sits at (c, r)
            if (c == r) v += (P.profile == PLBA_PROFILE_G) ? ctl.lambda : ctl.lambda * B.hd[(size_t)i * m + r];      // g2o: additive; hand LM: H_ii (1 + lambda)
            M[(size_t)r * ldm + c] = v;
        }
        for (int idx = tid; idx < ml * mi; idx += PLBA_NT) { const int r = idx / mi, c = idx - r * mi; M[(size_t)(mi + r) * ldm + c] = Ui[(size_t)r * m + c]; }
        // A[i+s, i] = A[i, i+s]^T: rows = unknowns of the right neighbour (reads walk the rows of its U block: coalesced over r)
        for (int idx = tid; idx < mr * mi; idx += PLBA_NT) { const int c = idx / mr, r = idx - c * mr; M[(size_t)(mi + ml + r) * ldm + c] = Ur[(size_t)c * m + r]; }
        for (int c = tid; c < mi; c += PLBA_NT) M[(size_t)nr * ldm + c] = B.b[(size_t)i * m + c];
    PHASE_END
    chol_lower_panels(M, ldm, mi, nr, dinv, Lblk, part, fail);
    double *Xl = B.Xl + (size_t)i * m * m, *Xr = B.Xr + (size_t)i * m * m;
    PHASE_BEGIN
        // factor and solved rows kept for the way back (the factored 6x6 diagonal blocks live in Lblk)
        for (int idx = tid; idx < mi * mi; idx += PLBA_NT) {
            const int r = idx / mi, c = idx - r * mi;
            if (c > r) continue;
            double v;
            if (c / 6 == r / 6) { const int kb = r / 6, a = r - 6 * kb, b2 = c - 6 * kb; v = Lblk[kb * 21 + a * (a + 1) / 2 + b2]; }
            else v = M[(size_t)r * ldm + c];
            Di[(size_t)r * m + c] = v;
        }
        for (int idx = tid; idx < ml * mi; idx += PLBA_NT) { const int r = idx / mi, c = idx - r * mi; Xl[(size_t)r * m + c] = M[(size_t)(mi + r) * ldm + c]; }
        for (int idx = tid; idx < mr * mi; idx += PLBA_NT) { const int r = idx / mi, c = idx - r * mi; Xr[(size_t)r * m + c] = M[(size_t)(mi + ml + r) * ldm + c]; }
        for (int c = tid; c < mi; c += PLBA_NT) B.y[(size_t)i * m + c] = M[(size_t)nr * ldm + c];
        if (tid == 0 && *fail) ctl.solve_fail = 1;
    PHASE_END
    if (final) return;
    // Schur complement onto the neighbours: 3 x 3 register tiles over (rows of Xa) x (rows of Xb), operands in shared memory.
    // Two CTAs of a level update the same D (its left and its right neighbour are both being eliminated): red.global.add.
    PHASE_BEGIN
        for (int job = 0; job < 3; job++) {
            // job 0: D_left -= Xl Xl^T (lower); job 1: D_right -= Xr Xr^T (lower); job 2: A[left,right] = -Xl Xr^T
            const int ma = (job == 1) ? mr : ml, mb = (job == 0) ? ml : mr;
            if (ma == 0 || mb == 0) continue;
            const double *Xa = M + (size_t)(mi + (job == 1 ? ml : 0)) * ldm, *Xb = M + (size_t)(mi + (job == 0 ? 0 : ml)) * ldm;
            double *dst = (job == 0) ? B.D + (size_t)left * m * m : (job == 1) ? B.D + (size_t)right * m * m : Ur;
            const int nta = (ma + 2) / 3, ntb = (mb + 2) / 3;
            for (int t = tid; t < nta * ntb; t += PLBA_NT) {
                const int ra0 = 3 * (t / ntb), rb0 = 3 * (t - (t / ntb) * ntb);
                if (job < 2 && rb0 > ra0 + 2) continue;                 // symmetric update: lower triangle only
                double acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
                const double *a0 = Xa + (size_t)ra0 * ldm, *b0 = Xb + (size_t)rb0 * ldm;
                const bool a1 = ra0 + 1 < ma, a2 = ra0 + 2 < ma, b1 = rb0 + 1 < mb, b2 = rb0 + 2 < mb;
#pragma unroll 2
                for (int q = 0; q < mi; q++) {
                    const double x0 = a0[q], x1 = a1 ? a0[ldm + q] : 0.0, x2 = a2 ? a0[2 * ldm + q] : 0.0;
                    const double y0 = b0[q], y1 = b1 ? b0[ldm + q] : 0.0, y2 = b2 ? b0[2 * ldm + q] : 0.0;
                    acc[0] += x0 * y0; acc[1] += x0 * y1; acc[2] += x0 * y2;
                    acc[3] += x1 * y0; acc[4] += x1 * y1; acc[5] += x1 * y2;
                    acc[6] += x2 * y0; acc[7] += x2 * y1; acc[8] += x2 * y2;
                }
#pragma unroll
                for (int u = 0; u < 3; u++) {
#pragma unroll
                    for (int v = 0; v < 3; v++) {
                        const int r = ra0 + u, c = rb0 + v;
                        if (r >= ma || c >= mb) continue;
                        if (job < 2) { if (c <= r) plba_atomic_add(dst + (size_t)c * m + r, -acc[u * 3 + v]); }
                        else dst[(size_t)r * m + c] = -acc[u * 3 + v];
                    }
                }
            }
        }
        // right-hand sides of the neighbours
        for (int t = tid; t < ml + mr; t += PLBA_NT) {
            const double *Xa = M + (size_t)(mi + t) * ldm, *yy = M + (size_t)nr * ldm;
            double sum = 0.0;
            for (int q = 0; q < mi; q++) sum += Xa[q] * yy[q];
            plba_atomic_add(&B.b[(size_t)(t < ml ? left : right) * m + (t < ml ? t : t - ml)], -sum);
        }
    PHASE_END
}

// the way back: L^T x_i = y - Xl^T x_left - Xr^T x_right for the nodes (2 j + 1) s (final != 0: node 0)
PLBA_KERNEL void k_bcr_back(const DevP *Pp, int w, BcrW B, int s, int final) {
    PLBA_SMEM(raw);
    PLBA_PARAMS(P, Pp);
    const WinCtrl &ctl = P.ctrl[w];
    if (ctl.done) return;
    const int nf = P.win_nfree[w], m = B.m, ldm = m + 1, slot0 = P.win_slot0[w];
    const int i = final ? 0 : (2 * PLBA_BID + 1) * s;
    const int left = final ? -1 : i - s, right = (final || i + s >= B.N) ? -1 : i + s;
    const int mi = bcr_node_size(B, nf, i), ml = left >= 0 ? m : 0, mr = right >= 0 ? bcr_node_size(B, nf, right) : 0;
    double *L = (double *)raw, *rhs = L + (size_t)BCR_M_MAX * (BCR_M_MAX + 1), *xs = rhs + BCR_M_MAX, *dinv = xs + BCR_M_MAX, *xn = dinv + BCR_M_MAX, *part = xn + 2 * BCR_M_MAX;
    const double *Di = (B.Lf ? B.Lf : B.D) + (size_t)i * m * m, *Xl = B.Xl + (size_t)i * m * m, *Xr = B.Xr + (size_t)i * m * m;      // the factor: over D_i (one-CTA elimination) or in Lf (split elimination)
    double *x = P.xp + (size_t)6 * slot0;
    PHASE_BEGIN
        // the factor, 128-bit loads, four independent loads in flight per thread (this kernel is pure latency: 50 us per level before)
        const int npair = mi * mi / 2;                                   // mi is a multiple of 6: rows hold whole pairs
        for (int i0 = tid; i0 < npair; i0 += 4 * PLBA_NT) {
            plba_d2 v[4]; int rr[4], cc[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int idx = i0 + u * PLBA_NT;
                rr[u] = (2 * idx) / mi; cc[u] = 2 * idx - rr[u] * mi;
                const bool on = idx < npair && cc[u] <= rr[u];
                v[u] = on ? *(const plba_d2 *)(Di + (size_t)rr[u] * m + cc[u]) : plba_d2{0.0, 0.0};
                if (!on) rr[u] = -1;
            }
#pragma unroll
            for (int u = 0; u < 4; u++) if (rr[u] >= 0) {
                L[(size_t)rr[u] * ldm + cc[u]] = v[u].x;
                if (cc[u] == rr[u]) dinv[rr[u]] = 1.0 / v[u].x;
                if (cc[u] + 1 <= rr[u]) { L[(size_t)rr[u] * ldm + cc[u] + 1] = v[u].y; if (cc[u] + 1 == rr[u]) dinv[rr[u]] = 1.0 / v[u].y; }
            }
        }
        // the neighbours' solutions, staged once
        for (int r = tid; r < ml; r += PLBA_NT) xn[r] = x[(size_t)6 * left * B.bs + r];
        for (int r = tid; r < mr; r += PLBA_NT) xn[BCR_M_MAX + r] = x[(size_t)6 * right * B.bs + r];
    PHASE_END
    PHASE_BEGIN
        // rhs = y - Xl^T x_left - Xr^T x_right: thread = (pair of columns, fifth of the rows), 128-bit loads, six rows in flight
        const int ncp = mi / 2, cp = tid % 48, qt = tid / 48;           // 256 threads: 5 row groups x 48 column pairs (45 used at m = 90)
        if (cp < ncp && qt < 5) {
            double a0 = 0.0, a1 = 0.0;
            for (int pass = 0; pass < 2; pass++) {
                const double *X = pass ? Xr : Xl; const double *xv = pass ? xn + BCR_M_MAX : xn; const int rows = pass ? mr : ml;
                for (int r0 = qt; r0 < rows; r0 += 30) {
                    plba_d2 v[6];
#pragma unroll
                    for (int u = 0; u < 6; u++) { const int r = r0 + 5 * u; v[u] = r < rows ? *(const plba_d2 *)(X + (size_t)r * m + 2 * cp) : plba_d2{0.0, 0.0}; }
#pragma unroll
                    for (int u = 0; u < 6; u++) { const int r = r0 + 5 * u; if (r < rows) { a0 += v[u].x * xv[r]; a1 += v[u].y * xv[r]; } }
                }
            }
            part[qt * BCR_M_MAX + 2 * cp] = a0; part[qt * BCR_M_MAX + 2 * cp + 1] = a1;
        }
    PHASE_END
    PHASE_BEGIN
        for (int c = tid; c < mi; c += PLBA_NT) { double v = B.y[(size_t)i * m + c]; for (int q = 0; q < 5; q++) v -= part[q * BCR_M_MAX + c]; rhs[c] = v; }
    PHASE_END
    // column-oriented backward substitution by ONE warp (the chain of mi steps is latency: no block barrier inside):
    // x_j = rhs_j / L_jj, then rhs_r -= L[j][r] x_j for r < j (row j of L: contiguous)
#ifndef PLBA_HOST_EMU
    if (threadIdx.x < 32) {
        // rhs lives in registers (lane l owns entries l, l + 32, l + 64), x_j travels by shuffle: ~60 cycles per step
        const int lane = (int)threadIdx.x;
        double r0 = lane < mi ? rhs[lane] : 0.0, r1 = lane + 32 < mi ? rhs[lane + 32] : 0.0, r2 = lane + 64 < mi ? rhs[lane + 64] : 0.0;
        const double d0 = lane < mi ? dinv[lane] : 0.0, d1 = lane + 32 < mi ? dinv[lane + 32] : 0.0, d2 = lane + 64 < mi ? dinv[lane + 64] : 0.0;
        for (int j = mi - 1; j >= 0; j--) {
            const double *Lj = L + (size_t)j * ldm;
            const double l0 = lane < j ? Lj[lane] : 0.0, l1 = lane + 32 < j ? Lj[lane + 32] : 0.0, l2 = lane + 64 < j ? Lj[lane + 64] : 0.0;   // independent of the chain
            const int q = j >> 5;
            const double mine = (q == 0 ? r0 * d0 : q == 1 ? r1 * d1 : r2 * d2);
            const double xj = __shfl_sync(0xffffffffu, mine, j & 31);
            if (lane == (j & 31)) xs[j] = xj;
            r0 -= l0 * xj; r1 -= l1 * xj; r2 -= l2 * xj;
        }
    }
#else
    for (int j = mi - 1; j >= 0; j--) {
        WPHASE_BEGIN
            if (lane == 0) xs[j] = rhs[j] * dinv[j];
        WPHASE_END
        WPHASE_BEGIN
            const double xj = xs[j];
            for (int r = lane; r < j; r += 32) rhs[r] -= L[(size_t)j * ldm + r] * xj;
        WPHASE_END
    }
#endif
    PHASE_BEGIN
    PHASE_END
    PHASE_BEGIN
        for (int c = tid; c < mi; c += PLBA_NT) x[(size_t)6 * i * B.bs + c] = xs[c];
    PHASE_END
}


// ---- block cyclic reduction, second generation: the elimination of a node split over SEVEN CTAs in two launches ------------------
// k_bcr_elim above factors a node, solves both coupling blocks against the factor and forms the three Schur products in ONE CTA:
// ~4 M FMAs on one SM (>= 32 us at the FP64 peak of an SM, ~130 us measured) while most of the GPU idles at the deep levels.
//   k_bcr_factor : two CTAs per eliminated node (side 0: left coupling block + right-hand side, side 1: right coupling block).  Each
//                  factors the node's diagonal block itself (same instructions, same bits: no communication) with the register-
//                  resident block Cholesky of k_solve_small — one thread per 6x6 block, the coupling block rides along as extra
//                  block rows — and writes L, X = A L^-T and y = L^-1 b.
//   k_bcr_schur  : five CTAs per eliminated node: D_left -= Xl Xl^T (+ b_left -= Xl y), D_right -= Xr Xr^T (+ b_right), and three row
//                  groups of A[left, right] = -Xl Xr^T; operands staged in shared memory, one thread per 6x6 output block.
enum { BF_NT = 384, BF_LOW0 = 32, BF_X0 = 32 + BCR_BS_MAX * (BCR_BS_MAX - 1) / 2 /* 137 */, BF_XLD = 37, BS_NT = 256, BS_PARTS = 5 };   // 384 threads = 3 warps per scheduler: 168 registers
static inline size_t bcr_factor_smem() { return sizeof(double) * ((size_t)(2 * BCR_BS_MAX) * BF_XLD + 32 + BCR_M_MAX + 8) + 64; }
static inline size_t bcr_schur_smem() { return sizeof(double) * ((size_t)(BCR_M_MAX + 30) * (BCR_M_MAX + 1) + BCR_M_MAX + 8); }   // 88 KB (parts 0-1: one 90-row operand; parts 2-4: 30 + 90 rows); compiled for one CTA per SM: 164 registers, no spills (two CTAs at 128 registers measured 3 % slower)

PLBA_KERNEL void PLBA_BOUNDS(BF_NT, 1) k_bcr_factor(const DevP *Pp, int w, BcrW B, int s, int final) {
    PLBA_SMEM(raw);
    PLBA_PARAMS(P, Pp);
    WinCtrl &ctl = P.ctrl[w];
    if (ctl.done) return;
    const int nf = P.win_nfree[w], m = B.m;
    const int side = final ? 0 : (PLBA_BID & 1);
    const int i = final ? 0 : (2 * (PLBA_BID >> 1) + 1) * s;
    const int right = (final || i + s >= B.N) ? -1 : i + s;
    if (side == 1 && right < 0) return;
    const int mi = bcr_node_size(B, nf, i), nb = mi / 6;
    const int nx = final ? 0 : (side == 0 ? m / 6 : bcr_node_size(B, nf, right) / 6);      // block rows of the coupling block that rides along
    double *Xp = (double *)raw, *Lc = Xp + (size_t)(2 * BCR_BS_MAX) * BF_XLD, *ys = Lc + 32;
    int *fail = (int *)(ys + BCR_M_MAX);
    double *Di = B.D + (size_t)i * m * m;
    const double *Ui = B.U + (size_t)i * m * m, *Ur = right >= 0 ? B.U + (size_t)right * m * m : nullptr;
    const double lambda = ctl.lambda;
    THR_ARR(double, blk, 36); THR_ARR(double, gv, 6);
    THR_VAR(int, bi); THR_VAR(int, bj); THR_VAR(int, act);
    PHASE_BEGIN
        THR_BIND(blk); THR_BIND(gv); THR_BIND(bi); THR_BIND(bj); THR_BIND(act);
        if (tid == 0) *fail = 0;
        act = 0; bi = 0; bj = 0;
        // thread -> block: warp 0 the diagonal blocks, threads 32..136 the strictly lower blocks of the node by column, threads 137.. the
        // coupling block by column (block row nb + xi): 137 + 225 <= 384
        if (tid < BF_LOW0) { if (tid < nb) { act = 1; bi = bj = tid; } }
        else if (tid < BF_X0) {
            int j = 0, rem = tid - BF_LOW0;
            while (j < nb - 1 && rem >= nb - 1 - j) { rem -= nb - 1 - j; j++; }
            if (j < nb - 1) { act = 1; bj = j; bi = j + 1 + rem; }
        } else if (nx > 0) {
            const int u = tid - BF_X0;
            if (u < nx * nb) { act = 2; bj = u / nx; bi = nb + (u - bj * nx); }
        }
        if (act == 1) {
            // lower block (bi, bj) = transpose of the stored upper block (bj, bi)
#pragma unroll
            for (int c = 0; c < 6; c++) {
                const plba_d2 *col = (const plba_d2 *)(Di + (size_t)(6 * bj + c) * m + 6 * bi);
#pragma unroll
                for (int r2 = 0; r2 < 3; r2++) { const plba_d2 v = col[r2]; blk[(2 * r2) * 6 + c] = v.x; blk[(2 * r2 + 1) * 6 + c] = v.y; }
            }
            if (bi == bj) {
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    blk[c * 6 + c] += (P.profile == PLBA_PROFILE_G) ? lambda : lambda * B.hd[(size_t)i * m + 6 * bj + c];      // g2o: additive; hand LM: H_ii (1 + lambda)
                    gv[c] = B.b[(size_t)i * m + 6 * bj + c];
                }
                if (bj == 0) {      // look-ahead for the first panel
                    double inv[6];
                    if (!chol6_inplace(blk, inv)) *fail = 1;
#pragma unroll
                    for (int r = 0; r < 6; r++) {
#pragma unroll
                        for (int c = 0; c <= r; c++) Lc[r * (r + 1) / 2 + c] = blk[r * 6 + c];
                    }
#pragma unroll
                    for (int c = 0; c < 6; c++) Lc[21 + c] = inv[c];
                }
            }
        } else if (act == 2) {
            const int xi = bi - nb;
            if (side == 0) {      // A[i - s, i] = U_i: rows = unknowns of the left neighbour
#pragma unroll
                for (int r = 0; r < 6; r++) {
                    const plba_d2 *row = (const plba_d2 *)(Ui + (size_t)(6 * xi + r) * m + 6 * bj);
#pragma unroll
                    for (int c2 = 0; c2 < 3; c2++) { const plba_d2 v = row[c2]; blk[r * 6 + 2 * c2] = v.x; blk[r * 6 + 2 * c2 + 1] = v.y; }
                }
            } else {              // A[i + s, i] = U_right^T: rows = unknowns of the right neighbour
#pragma unroll
                for (int c = 0; c < 6; c++) {
                    const plba_d2 *col = (const plba_d2 *)(Ur + (size_t)(6 * bj + c) * m + 6 * xi);
#pragma unroll
                    for (int r2 = 0; r2 < 3; r2++) { const plba_d2 v = col[r2]; blk[(2 * r2) * 6 + c] = v.x; blk[(2 * r2 + 1) * 6 + c] = v.y; }
                }
            }
        }
    PHASE_END
    for (int kb = 0; kb < nb; kb++) {
        // ---- A(kb): block column kb: X = M L_kk^-T, y_k = L_kk^-1 b_k ----
        PHASE_BEGIN
            THR_BIND(blk); THR_BIND(gv); THR_BIND(bi); THR_BIND(bj); THR_BIND(act);
            if (act && bj == kb) {
                double L[21], inv[6];
#pragma unroll
                for (int q = 0; q < 21; q++) L[q] = Lc[q];
#pragma unroll
                for (int q = 0; q < 6; q++) inv[q] = Lc[21 + q];
                if (bi == kb) {
                    double y[6];
#pragma unroll
                    for (int c = 0; c < 6; c++) y[c] = gv[c];
#pragma unroll
                    for (int c = 0; c < 6; c++) {
                        y[c] *= inv[c];
#pragma unroll
                        for (int mm = c + 1; mm < 6; mm++) y[mm] -= L[mm * (mm + 1) / 2 + c] * y[c];
                    }
#pragma unroll
                    for (int c = 0; c < 6; c++) ys[6 * kb + c] = y[c];
                } else {
#pragma unroll
                    for (int c = 0; c < 6; c++) {
#pragma unroll
                        for (int r = 0; r < 6; r++) blk[r * 6 + c] *= inv[c];
#pragma unroll
                        for (int mm = c + 1; mm < 6; mm++) {
#pragma unroll
                            for (int r = 0; r < 6; r++) blk[r * 6 + mm] -= blk[r * 6 + c] * L[mm * (mm + 1) / 2 + c];
                        }
                    }
#pragma unroll
                    for (int q = 0; q < 36; q++) Xp[bi * BF_XLD + q] = blk[q];
                }
            }
        PHASE_END
        // ---- B(kb): trailing update on registers; look-ahead factorisation of the next diagonal block ----
        PHASE_BEGIN
            THR_BIND(blk); THR_BIND(gv); THR_BIND(bi); THR_BIND(bj); THR_BIND(act);
            if (act && bj > kb) {
                double Xj[36];
#pragma unroll
                for (int q = 0; q < 36; q++) Xj[q] = Xp[bj * BF_XLD + q];
                if (bi != bj) {
#pragma unroll
                    for (int r = 0; r < 6; r++) {
                        double xi[6];
#pragma unroll
                        for (int mm = 0; mm < 6; mm++) xi[mm] = Xp[bi * BF_XLD + r * 6 + mm];
#pragma unroll
                        for (int c = 0; c < 6; c++) {
                            double v = blk[r * 6 + c];
#pragma unroll
                            for (int mm = 0; mm < 6; mm++) v -= xi[mm] * Xj[c * 6 + mm];
                            blk[r * 6 + c] = v;
                        }
                    }
                } else {
#pragma unroll
                    for (int r = 0; r < 6; r++) {
#pragma unroll
                        for (int c = 0; c <= r; c++) {
                            double v = blk[r * 6 + c];
#pragma unroll
                            for (int mm = 0; mm < 6; mm++) v -= Xj[r * 6 + mm] * Xj[c * 6 + mm];
                            blk[r * 6 + c] = v;
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 6; r++) {
                        double v = gv[r];
#pragma unroll
                        for (int mm = 0; mm < 6; mm++) v -= Xj[r * 6 + mm] * ys[6 * kb + mm];
                        gv[r] = v;
                    }
                    if (bj == kb + 1) {
                        double inv[6];
                        if (!chol6_inplace(blk, inv)) *fail = 1;
#pragma unroll
                        for (int r = 0; r < 6; r++) {
#pragma unroll
                            for (int c = 0; c <= r; c++) Lc[r * (r + 1) / 2 + c] = blk[r * 6 + c];
                        }
#pragma unroll
                        for (int c = 0; c < 6; c++) Lc[21 + c] = inv[c];
                    }
                }
            }
        PHASE_END
    }
    // ---- results: L over D_i (side 0), X = A L^-T into Xl / Xr, y ----
    PHASE_BEGIN
        THR_BIND(blk); THR_BIND(bi); THR_BIND(bj); THR_BIND(act);
        if (act == 1 && side == 0) {
            double *Li = B.Lf + (size_t)i * m * m;      // (not over D_i: the other CTA of this node may still be loading it)
#pragma unroll
            for (int r = 0; r < 6; r++) {
#pragma unroll
                for (int c = 0; c < 6; c++) if (bi != bj || c <= r) Li[(size_t)(6 * bi + r) * m + 6 * bj + c] = blk[r * 6 + c];
            }
        } else if (act == 2) {
            double *X = (side == 0 ? B.Xl : B.Xr) + (size_t)i * m * m;
            const int xi = bi - nb;
#pragma unroll
            for (int r = 0; r < 6; r++) {
                plba_d2 *row = (plba_d2 *)(X + (size_t)(6 * xi + r) * m + 6 * bj);
#pragma unroll
                for (int c2 = 0; c2 < 3; c2++) { plba_d2 v; v.x = blk[r * 6 + 2 * c2]; v.y = blk[r * 6 + 2 * c2 + 1]; row[c2] = v; }
            }
        }
        if (side == 0) for (int c = tid; c < mi; c += PLBA_NT) B.y[(size_t)i * m + c] = ys[c];
        if (tid == 0 && *fail) ctl.solve_fail = 1;
    PHASE_END
}

PLBA_KERNEL void PLBA_BOUNDS(BS_NT, 1) k_bcr_schur(const DevP *Pp, int w, BcrW B, int s) {
    PLBA_SMEM(raw);
    PLBA_PARAMS(P, Pp);
    const WinCtrl &ctl = P.ctrl[w];
    if (ctl.done) return;
    const int nf = P.win_nfree[w], m = B.m, ld = m + 1;
    const int part = PLBA_BID % BS_PARTS, i = (2 * (PLBA_BID / BS_PARTS) + 1) * s;
    const int left = i - s, right = (i + s >= B.N) ? -1 : i + s;
    if (part >= 1 && right < 0) return;
    const int mi = bcr_node_size(B, nf, i), nb = mi / 6, ml = m, mr = right >= 0 ? bcr_node_size(B, nf, right) : 0;
    double *Xa = (double *)raw, *Xb = Xa + (size_t)30 * (BCR_M_MAX + 1), *ysm = Xa + (size_t)(BCR_M_MAX + 30) * (BCR_M_MAX + 1);      // parts 0-1 use Xa alone (up to 90 rows: it runs into Xb's space)
    const double *Xl = B.Xl + (size_t)i * m * m, *Xr = B.Xr + (size_t)i * m * m;
    // operands: part 0: Xa = Xl; part 1: Xa = Xr; parts 2-4: Xa = five block rows of Xl, Xb = Xr
    const int rg = part - 2;                                         // row group of the left neighbour (parts 2-4)
    const int ra0 = part >= 2 ? 30 * rg : 0;
    const int rows_a = part == 0 ? ml : part == 1 ? mr : ((ml - ra0) < 30 ? ((ml - ra0) > 0 ? ml - ra0 : 0) : 30);
    const double *srcA = part == 1 ? Xr : Xl + (size_t)ra0 * m;
    PHASE_BEGIN
        const int warp = tid >> 5, lane = tid & 31, nw = PLBA_NT >> 5;
        for (int r = warp; r < rows_a; r += nw) for (int c = lane; c < mi; c += 32) Xa[(size_t)r * ld + c] = srcA[(size_t)r * m + c];
        if (part >= 2) for (int r = warp; r < mr; r += nw) for (int c = lane; c < mi; c += 32) Xb[(size_t)r * ld + c] = Xr[(size_t)r * m + c];
        if (part < 2) for (int c = tid; c < mi; c += PLBA_NT) ysm[c] = B.y[(size_t)i * m + c];
    PHASE_END
    PHASE_BEGIN
        if (part < 2) {
            // D_n -= X X^T on the lower blocks (p >= q); two eliminations of one level meet in the same D: red.global.add
            const int nbo = rows_a / 6, nblk = nbo * (nbo + 1) / 2;
            double *Dn = B.D + (size_t)(part == 0 ? left : right) * m * m;
            if (tid < nblk) {
                int q = 0, rem = tid;
                while (rem >= nbo - q) { rem -= nbo - q; q++; }
                const int p = q + rem;
                double acc[36];
#pragma unroll
                for (int e = 0; e < 36; e++) acc[e] = 0.0;
                for (int c = 0; c < nb; c++) {
                    double Xq[36];
#pragma unroll
                    for (int r = 0; r < 6; r++) {
#pragma unroll
                        for (int k = 0; k < 6; k++) Xq[r * 6 + k] = Xa[(size_t)(6 * q + r) * ld + 6 * c + k];
                    }
#pragma unroll
                    for (int r = 0; r < 6; r++) {
                        double xp[6];
#pragma unroll
                        for (int k = 0; k < 6; k++) xp[k] = Xa[(size_t)(6 * p + r) * ld + 6 * c + k];
#pragma unroll
                        for (int cc = 0; cc < 6; cc++) {
#pragma unroll
                            for (int k = 0; k < 6; k++) acc[r * 6 + cc] += xp[k] * Xq[cc * 6 + k];
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < 6; r++) {
#pragma unroll
                    for (int cc = 0; cc < 6; cc++) if (p != q || cc <= r) plba_atomic_add(Dn + (size_t)(6 * q + cc) * m + 6 * p + r, -acc[r * 6 + cc]);      // (upper storage: entry (row, col) of the lower triangle sits at (col, row))
                }
            } else if (tid >= 128 && tid - 128 < rows_a) {
                // right-hand side of the neighbour: b_n -= X y
                const int r = tid - 128;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0;
                int q = 0;
                for (; q + 2 < mi; q += 3) { s0 += Xa[(size_t)r * ld + q] * ysm[q]; s1 += Xa[(size_t)r * ld + q + 1] * ysm[q + 1]; s2 += Xa[(size_t)r * ld + q + 2] * ysm[q + 2]; }
                for (; q < mi; q++) s0 += Xa[(size_t)r * ld + q] * ysm[q];
                plba_atomic_add(&B.b[(size_t)(part == 0 ? left : right) * m + r], -(s0 + s1 + s2));
            }
        } else {
            // A[left, right] = -Xl Xr^T (rows 30 rg .. of the left neighbour): plain stores, this CTA is the only writer
            double *Un = B.U + (size_t)right * m * m;
            const int nbp = rows_a / 6, nbq = mr / 6;
            if (tid < nbp * nbq) {
                const int p = tid / nbq, q = tid - p * nbq;
                double acc[36];
#pragma unroll
                for (int e = 0; e < 36; e++) acc[e] = 0.0;
                for (int c = 0; c < nb; c++) {
                    double Xq[36];
#pragma unroll
                    for (int r = 0; r < 6; r++) {
#pragma unroll
                        for (int k = 0; k < 6; k++) Xq[r * 6 + k] = Xb[(size_t)(6 * q + r) * ld + 6 * c + k];
                    }
#pragma unroll
                    for (int r = 0; r < 6; r++) {
                        double xp[6];
#pragma unroll
                        for (int k = 0; k < 6; k++) xp[k] = Xa[(size_t)(6 * p + r) * ld + 6 * c + k];
#pragma unroll
                        for (int cc = 0; cc < 6; cc++) {
#pragma unroll
                            for (int k = 0; k < 6; k++) acc[r * 6 + cc] += xp[k] * Xq[cc * 6 + k];
                        }
                    }
                }
#pragma unroll
                for (int r = 0; r < 6; r++) {
#pragma unroll
                    for (int cc = 0; cc < 6; cc++) Un[(size_t)(ra0 + 6 * p + r) * m + 6 * q + cc] = -acc[r * 6 + cc];
                }
            }
        }
    PHASE_END
}

}  // namespace plba
