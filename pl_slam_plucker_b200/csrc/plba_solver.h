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

enum { SMALL_NMAX = 144, TB = 48 };

static inline size_t solve_small_smem() { return sizeof(double) * ((size_t)(SMALL_NMAX + 1) * (SMALL_NMAX + 1) + 2 * SMALL_NMAX + 24 + 16 + 6 * 264 + 21 * (SMALL_NMAX / 6)) + 64; }

PLBA_KERNEL void k_solve_small(const DevP *Pp) {
    PLBA_SMEM(raw);
    PLBA_PARAMS(P, Pp);
    PROF_DECL;
    for (int w = PLBA_BID; w < P.n_win; w += PLBA_NB) {
        WinCtrl &ctl = P.ctrl[w];
        const int nf = P.win_nfree[w], n = 6 * nf, slot0 = P.win_slot0[w], ldm = n + 1;     // n even => ldm odd: conflict-free columns
        double *M = (double *)raw;                 // (n+1) x ldm lower triangle; row n = right-hand side
        double *dinv = M + (size_t)(n + 1) * ldm, *xs = dinv + n, *Ls = xs + n, *red = Ls + 24;
        int *flag = (int *)(red + 4);   // (red: 4 doubles, flag: 2 ints, then the partial-sum buffer)
        (void)0;              // [0] skip, [1] fail
        double *Sw = P.S + P.win_S_off[w];
        PHASE_BEGIN
            if (tid == 0) {
                if (!ctl.done && P.profile != PLBA_PROFILE_G) control_h_pre_window(P, w);
                flag[0] = ctl.done; flag[1] = 0; red[0] = 0.0; red[1] = 0.0;
            }
        PHASE_END
    PROF_MARK(41);
        const int skip = flag[0];
        PHASE_BEGIN
        PHASE_END
    PROF_MARK(42);
        if (skip || n == 0) continue;
        const double lambda = ctl.lambda;
        PHASE_BEGIN
            // upper triangle of S, one warp per row (coalesced), four loads in flight per thread; stored transposed as lower (cg,rg)
            const int warp = tid >> 5, lane = tid & 31, nwarp = PLBA_NT >> 5;
            for (int rg0 = warp; rg0 < n; rg0 += 2 * nwarp) {
                double v[2][4];
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int rg = rg0 + rr * nwarp;
#pragma unroll
                    for (int u = 0; u < 4; u++) { const int cg = lane + 32 * u; v[rr][u] = (rg < n && cg >= rg && cg < n) ? Sw[(size_t)rg * n + cg] : 0.0; }
                }
#pragma unroll
                for (int rr = 0; rr < 2; rr++) {
                    const int rg = rg0 + rr * nwarp;
                    if (rg >= n) continue;
                    const double dmp = (P.profile == PLBA_PROFILE_G) ? lambda : lambda * P.hpp_diag[(size_t)6 * slot0 + rg];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int cg = lane + 32 * u;
                        if (cg >= rg && cg < n) { Sw[(size_t)rg * n + cg] = 0.0; M[(size_t)cg * ldm + rg] = (cg == rg) ? v[rr][u] + dmp : v[rr][u]; }   // consumed: the next assembly accumulates into a clean S
                    }
                }
            }
            for (int rg = warp; rg < n && n > 128; rg += nwarp) {      // columns beyond 128 (n <= 144)
                const int cg = 128 + lane;
                if (cg < n && cg >= rg) { const double x = Sw[(size_t)rg * n + cg]; Sw[(size_t)rg * n + cg] = 0.0; M[(size_t)cg * ldm + rg] = (cg == rg) ? x + ((P.profile == PLBA_PROFILE_G) ? lambda : lambda * P.hpp_diag[(size_t)6 * slot0 + rg]) : x; }
            }
            for (int i = tid; i < n; i += PLBA_NT) { M[(size_t)n * ldm + i] = P.gs[(size_t)6 * slot0 + i]; P.gs[(size_t)6 * slot0 + i] = 0.0; }
        PHASE_END
    PROF_MARK(43);
        PHASE_BEGIN
            for (int i = tid; i < n; i += PLBA_NT) P.hpp_diag[(size_t)6 * slot0 + i] = 0.0;
        PHASE_END
    PROF_MARK(44);
        // Left-looking Cholesky in panels of one pose block (6 columns).  Per panel two phases:
        //  U: every row r >= k0 (row n = right-hand side) accumulates sum_{q<k0} L[r][q] L[k0+c][q] for the panel's 6 columns.  One
        //     thread per (row, half of the q range): its own row is read once per q (stride ldm, odd: conflict-free) and the six
        //     rows k0..k0+5 are BROADCAST loads shared by the whole warp, so the phase is FMA-bound, not shared-memory-bound,
        //     and no trailing matrix is ever re-written;
        //  A: every row folds the partial sums in, the 6x6 diagonal block is factored redundantly in the registers of each row's
        //     thread (no extra barrier), then the row is solved against it.
        double *part = Ls + 24 + 8;               // [nsplit][n+1][6] partial sums of phase U (nsplit * rows <= 256)
        double *Lblk = part + 6 * 264;            // [nf][21] factored diagonal blocks (M keeps their un-factored values)
        for (int kb = 0; kb < nf; kb++) {
            const int k0 = 6 * kb, m = n - k0 + 1;
            // the q range of a row is split over as many threads as the 256-thread CTA allows: late panels have few rows but long rows
            const int mpad = (m + 31) & ~31, nsplit = (mpad <= 32) ? 8 : (mpad <= 64) ? 4 : (mpad <= 128) ? 2 : 1;
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
#pragma unroll
                    for (int c = 0; c < 6; c++) part[((size_t)sp * mpad + rr) * 6 + c] = acc[c];
                }
            PHASE_END
            PROF_MARK(45);
            PHASE_BEGIN
                if (k0 > 0) for (int idx = tid; idx < 6 * m; idx += PLBA_NT) {       // fold the partial sums into the panel, one thread per entry
                    const int rr = idx / 6, c = idx - 6 * rr;
                    double sum = 0.0;
                    for (int sp = 0; sp < nsplit; sp++) sum += part[((size_t)sp * mpad + rr) * 6 + c];
                    M[(size_t)(k0 + rr) * ldm + k0 + c] -= sum;
                }
            PHASE_END
            PROF_MARK(51);
            PHASE_BEGIN
                const int r = k0 + 6 + tid;              // rows below the diagonal block
                if (r <= n) {
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
                    if (tid == 0) {
                        if (bad) flag[1] = 1;
#pragma unroll
                        for (int c = 0; c < 6; c++) dinv[k0 + c] = inv[c];
#pragma unroll
                        for (int i = 0; i < 21; i++) Lblk[kb * 21 + i] = L[i];
                    }
                }
            PHASE_END
            PROF_MARK(46);
        }
        // backward substitution L^T x = y, y = row n; column-oriented so that every step reads rows of L
        for (int kb = nf - 1; kb >= 0; kb--) {
            const int k0 = 6 * kb;
            PHASE_BEGIN
                if (tid <= k0) {       // threads 0..k0-1 own y[j]; thread k0 (or 0 when k0 == 0) records x
                    double x[6];
#pragma unroll
                    for (int c = 5; c >= 0; c--) {
                        double v = M[(size_t)n * ldm + k0 + c];
#pragma unroll
                        for (int m = c + 1; m < 6; m++) v -= Lblk[kb * 21 + m * (m + 1) / 2 + c] * x[m];
                        x[c] = v * dinv[k0 + c];
                    }
                    if (tid < k0) {
                        double y = M[(size_t)n * ldm + tid];
#pragma unroll
                        for (int c = 0; c < 6; c++) y -= M[(size_t)(k0 + c) * ldm + tid] * x[c];
                        M[(size_t)n * ldm + tid] = y;
                    } else {
#pragma unroll
                        for (int c = 0; c < 6; c++) xs[k0 + c] = x[c];
                    }
                }
            PHASE_END
    PROF_MARK(47);
        }
        PHASE_BEGIN
            const int f = flag[1];
            for (int i = tid; i < n; i += PLBA_NT) P.xp[(size_t)6 * slot0 + i] = f ? 0.0 : xs[i];
            if (tid == 0 && f) ctl.solve_fail = 1;
        PHASE_END
    PROF_MARK(48);
        PHASE_BEGIN
            double sc = 0.0, d2 = 0.0;
            if (tid < nf) pose_update_slot(P, slot0 + tid, sc, d2);
            plba_block_add(&red[0], sc);
            plba_block_add(&red[1], d2);
        PHASE_END
    PROF_MARK(49);
        PHASE_BEGIN
            if (tid == 0) { ctl.scale_pose = red[0]; ctl.dx2_pose = red[1]; }
        PHASE_END
    PROF_MARK(50);
    }
}

// ---- tiled path -----------------------------------------------------------------------------------------------
// Factor the diagonal tile k (damping added here: later tiles only receive additive updates) and store inv(U_kk).
PLBA_KERNEL void k_potrf_tile(const DevP *Pp, int w, int k, double *invbuf) {
    PLBA_SMEM(raw);
    PLBA_PARAMS(P, Pp);
    WinCtrl &ctl = P.ctrl[w];
    const int n = 6 * P.win_nfree[w], slot0 = P.win_slot0[w];
    double *Sw = P.S + P.win_S_off[w];
    const int k0 = k * TB, tb = (n - k0 < TB) ? n - k0 : TB;
    double *M = (double *)raw;                 // tb x (TB+1), holds U (upper) row-major
    double *Inv = M + TB * (TB + 1);
    double *dgt = Inv + TB * (TB + 1);         // diagonal of U (kept apart: the pivot row is read by every thread)
    const int ldm = TB + 1;
    PHASE_BEGIN
        for (int idx = tid; idx < tb * tb; idx += PLBA_NT) {
            const int r = idx / tb, c = idx % tb;
            double v = 0.0;
            if (c >= r) {
                v = Sw[(size_t)(k0 + r) * n + k0 + c];
                if (r == c) v += (P.profile == PLBA_PROFILE_G) ? ctl.lambda : ctl.lambda * P.hpp_diag[(size_t)6 * slot0 + k0 + r];
            }
            M[r * ldm + c] = v; Inv[r * ldm + c] = 0.0;
        }
    PHASE_END
    for (int q = 0; q < tb; q++) {             // row-oriented Cholesky of the upper triangle: U^T U
        PHASE_BEGIN
            const double p = M[q * ldm + q];
            const bool bad = !(p > 0.0) || !plba_isfinite(p);
            const double d = bad ? 1.0 : sqrt(p);
            if (tid == 0) { dgt[q] = d; if (bad) ctl.solve_fail = 1; }
            for (int c = q + 1 + tid; c < tb; c += PLBA_NT) M[q * ldm + c] /= d;
        PHASE_END
        PHASE_BEGIN
            const int ti = tid >> 4, tj = tid & 15;
            for (int r = q + 1 + ti; r < tb; r += PLBA_NT >> 4) {
                const double uqr = M[q * ldm + r];
                for (int c = r + tj; c < tb; c += 16) M[r * ldm + c] -= uqr * M[q * ldm + c];
            }
        PHASE_END
    }
    // inverse of the upper-triangular tile, column by column: U X = I
    PHASE_BEGIN
        if (tid < tb) {
            const int c = tid;
            for (int r = c; r >= 0; r--) {
                double s = (r == c) ? 1.0 : 0.0;
                for (int m = r + 1; m <= c; m++) s -= M[r * ldm + m] * Inv[m * ldm + c];
                Inv[r * ldm + c] = s / dgt[r];
            }
        }
    PHASE_END
    PHASE_BEGIN
        for (int idx = tid; idx < tb * tb; idx += PLBA_NT) {
            const int r = idx / tb, c = idx % tb;
            if (c >= r) Sw[(size_t)(k0 + r) * n + k0 + c] = (c == r) ? dgt[r] : M[r * ldm + c];
            invbuf[(size_t)k * TB * TB + r * TB + c] = (c >= r) ? Inv[r * ldm + c] : 0.0;
        }
    PHASE_END
}

// U_kj = inv(U_kk)^T A_kj for every tile j > k  (one CTA per tile)
PLBA_KERNEL void k_trsm_tiles(const DevP *Pp, int w, int k, const double *invbuf) {
    PLBA_SMEM(raw);
    PLBA_PARAMS(P, Pp);
    const int n = 6 * P.win_nfree[w];
    double *Sw = P.S + P.win_S_off[w];
    const int j = k + 1 + PLBA_BID;
    const int k0 = k * TB, tbk = (n - k0 < TB) ? n - k0 : TB;
    const int j0 = j * TB, tbj = (n - j0 < TB) ? n - j0 : TB;
    double *Inv = (double *)raw, *Aj = Inv + TB * TB;     // Inv[q][r], Aj[q][c]
    PHASE_BEGIN
        for (int idx = tid; idx < TB * TB; idx += PLBA_NT) {
            const int q = idx / TB, c = idx % TB;
            Inv[idx] = (q < tbk && c < tbk) ? invbuf[(size_t)k * TB * TB + idx] : 0.0;
            Aj[idx] = (q < tbk && c < tbj) ? Sw[(size_t)(k0 + q) * n + j0 + c] : 0.0;
        }
    PHASE_END
    PHASE_BEGIN
        const int ty = tid >> 4, tx = tid & 15;
        double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int q = 0; q < tbk; q++) {
            double a[3], bb[3];
            for (int u = 0; u < 3; u++) { a[u] = Inv[q * TB + ty * 3 + u]; bb[u] = Aj[q * TB + tx * 3 + u]; }   // (Inv^T)[r][q] = Inv[q][r]
            for (int u = 0; u < 3; u++) for (int v = 0; v < 3; v++) acc[u][v] += a[u] * bb[v];
        }
        for (int u = 0; u < 3; u++) for (int v = 0; v < 3; v++) {
            const int r = ty * 3 + u, c = tx * 3 + v;
            if (r < tbk && c < tbj) Sw[(size_t)(k0 + r) * n + j0 + c] = acc[u][v];
        }
    PHASE_END
}

// A_ij -= U_ki^T U_kj for k < i <= j  (one CTA per tile pair; grid.x enumerates the upper-triangular pairs)
PLBA_KERNEL void k_syrk_tiles(const DevP *Pp, int w, int k, int nt) {
    PLBA_SMEM(raw);
    PLBA_PARAMS(P, Pp);
    const int n = 6 * P.win_nfree[w];
    double *Sw = P.S + P.win_S_off[w];
    const int m = nt - k - 1;
    int p = PLBA_BID, a = 0;
    while (p >= m - a) { p -= m - a; a++; }
    const int i = k + 1 + a, j = i + p;
    const int k0 = k * TB, tbk = (n - k0 < TB) ? n - k0 : TB;
    const int i0 = i * TB, tbi = (n - i0 < TB) ? n - i0 : TB;
    const int j0 = j * TB, tbj = (n - j0 < TB) ? n - j0 : TB;
    double *Ui = (double *)raw, *Uj = Ui + TB * TB;
    PHASE_BEGIN
        for (int idx = tid; idx < TB * TB; idx += PLBA_NT) {
            const int q = idx / TB, c = idx % TB;
            Ui[idx] = (q < tbk && c < tbi) ? Sw[(size_t)(k0 + q) * n + i0 + c] : 0.0;
            Uj[idx] = (q < tbk && c < tbj) ? Sw[(size_t)(k0 + q) * n + j0 + c] : 0.0;
        }
    PHASE_END
    PHASE_BEGIN
        const int ty = tid >> 4, tx = tid & 15;
        double acc[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
        for (int q = 0; q < tbk; q++) {
            double av[3], bv[3];
            for (int u = 0; u < 3; u++) { av[u] = Ui[q * TB + ty * 3 + u]; bv[u] = Uj[q * TB + tx * 3 + u]; }
            for (int u = 0; u < 3; u++) for (int v = 0; v < 3; v++) acc[u][v] += av[u] * bv[v];
        }
        for (int u = 0; u < 3; u++) for (int v = 0; v < 3; v++) {
            const int r = ty * 3 + u, c = tx * 3 + v;
            if (r < tbi && c < tbj && (i != j || c >= r)) Sw[(size_t)(i0 + r) * n + j0 + c] -= acc[u][v];
        }
    PHASE_END
}

// U^T y = g ; U x = y with the stored inverse diagonal tiles (single CTA, 1024 threads)
PLBA_KERNEL void k_trisolve_large(const DevP *Pp, int w, int nt, const double *invbuf) {
    PLBA_SMEM(raw);
    PLBA_PARAMS(P, Pp);
    WinCtrl &ctl = P.ctrl[w];
    const int n = 6 * P.win_nfree[w], slot0 = P.win_slot0[w];
    const double *Sw = P.S + P.win_S_off[w];
    double *g = P.xp + (size_t)6 * slot0;       // solved in place in xp
    double *yt = (double *)raw, *part = yt + TB;   // part[TB][G]
    const int G = PLBA_NT / TB;
    PHASE_BEGIN
        for (int i = tid; i < n; i += PLBA_NT) g[i] = P.gs[(size_t)6 * slot0 + i];
    PHASE_END
    for (int k = 0; k < nt; k++) {
        const int k0 = k * TB, tb = (n - k0 < TB) ? n - k0 : TB;
        const double *Inv = invbuf + (size_t)k * TB * TB;
        PHASE_BEGIN
            if (tid < tb) { double s = 0; for (int q = 0; q <= tid; q++) s += Inv[q * TB + tid] * g[k0 + q]; yt[tid] = s; }
        PHASE_END
        PHASE_BEGIN
            if (tid < tb) g[k0 + tid] = yt[tid];
            for (int j = k0 + tb + tid; j < n; j += PLBA_NT) {
                double s = 0; for (int r = 0; r < tb; r++) s += Sw[(size_t)(k0 + r) * n + j] * yt[r];
                g[j] -= s;
            }
        PHASE_END
    }
    for (int k = nt - 1; k >= 0; k--) {
        const int k0 = k * TB, tb = (n - k0 < TB) ? n - k0 : TB;
        const double *Inv = invbuf + (size_t)k * TB * TB;
        PHASE_BEGIN
            const int r = tid / G, q = tid % G;
            if (r < tb) { double s = 0; for (int j = k0 + tb + q; j < n; j += G) s += Sw[(size_t)(k0 + r) * n + j] * g[j]; part[r * G + q] = s; }
        PHASE_END
        PHASE_BEGIN
            if (tid < tb) { double s = g[k0 + tid]; for (int q = 0; q < G; q++) s -= part[tid * G + q]; yt[tid] = s; }
        PHASE_END
        PHASE_BEGIN
            if (tid < tb) { double s = 0; for (int q = tid; q < tb; q++) s += Inv[tid * TB + q] * yt[q]; g[k0 + tid] = s; }
        PHASE_END
    }
    PHASE_BEGIN
        if (ctl.solve_fail) for (int i = tid; i < n; i += PLBA_NT) g[i] = 0.0;
    PHASE_END
}

}  // namespace plba
