// Damped solve of the reduced camera system  (S + damping) x_p = g_red   (subsystem 4; replaces
// Eigen::SimplicialLDLT of src/mapHandler.cpp:2566-2568 and g2o LinearSolverEigen, src/mapHandler.cpp:5925).
// S is symmetric positive definite after damping: Cholesky S = U^T U in FP64 on the vector pipe (tcgen05 has no
// f64 kind; DMMA is considered only for the large-window GEMM update).
//   k_solve_small : one CTA per window, whole system in shared memory (n = 6 Nkf <= 144: configs 1-3)
//   k_potrf_tile / k_trsm_tiles / k_syrk_tiles / k_trisolve_large : tiled right-looking factorisation in HBM/L2
//                   for large windows (configs 4-5), 48 x 48 tiles, many CTAs per step.
#pragma once
#include "plba_kernels.h"

namespace plba {

enum { SMALL_NMAX = 144, TB = 48 };

PLBA_KERNEL void k_solve_small(DevP P, int ldm) {
    PLBA_SMEM(raw);
    const int w = PLBA_BID;
    WinCtrl &ctl = P.ctrl[w];
    if (ctl.done) return;
    const int nf = P.win_nfree[w], n = 6 * nf, slot0 = P.win_slot0[w];
    double *M = (double *)raw;                 // n x ldm, lower triangle used
    double *dg = M + (size_t)n * ldm, *b = dg + n, *y = b + n, *x = y + n;
    int *fail = (int *)(x + n);
    const double *Sw = P.S + P.win_S_off[w];
    PHASE_BEGIN
        if (tid == 0) *fail = 0;
        for (int idx = tid; idx < n * n; idx += PLBA_NT) {
            const int r = idx / n, c = idx % n;
            if (c > r) continue;
            double v = Sw[(size_t)c * n + r];      // upper entry (c,r) mirrored into lower (r,c)
            if (r == c) v += (P.profile == PLBA_PROFILE_G) ? ctl.lambda : ctl.lambda * P.hpp_diag[(size_t)6 * slot0 + r];
            M[(size_t)r * ldm + c] = v;
        }
        for (int i = tid; i < n; i += PLBA_NT) b[i] = P.gs[(size_t)6 * slot0 + i];
    PHASE_END
    for (int k = 0; k < n; k++) {
        PHASE_BEGIN
            const double p = M[(size_t)k * ldm + k];
            const bool bad = !(p > 0.0) || !plba_isfinite(p);
            const double d = bad ? 1.0 : sqrt(p);
            if (tid == 0) { dg[k] = d; if (bad) *fail = 1; }
            for (int i = k + 1 + tid; i < n; i += PLBA_NT) M[(size_t)i * ldm + k] /= d;
        PHASE_END
        PHASE_BEGIN
            const int ti = tid >> 4, tj = tid & 15;
            for (int i = k + 1 + ti; i < n; i += PLBA_NT >> 4) {
                const double lik = M[(size_t)i * ldm + k];
                for (int j = k + 1 + tj; j <= i; j += 16) M[(size_t)i * ldm + j] -= lik * M[(size_t)j * ldm + k];
            }
        PHASE_END
    }
    for (int k = 0; k < n; k++) {              // L y = b
        PHASE_BEGIN
            const double yk = b[k] / dg[k];
            if (tid == 0) y[k] = yk;
            for (int i = k + 1 + tid; i < n; i += PLBA_NT) b[i] -= M[(size_t)i * ldm + k] * yk;
        PHASE_END
    }
    for (int k = n - 1; k >= 0; k--) {         // L^T x = y
        PHASE_BEGIN
            const double xk = y[k] / dg[k];
            if (tid == 0) x[k] = xk;
            for (int i = tid; i < k; i += PLBA_NT) y[i] -= M[(size_t)k * ldm + i] * xk;
        PHASE_END
    }
    PHASE_BEGIN
        const int f = *fail;
        for (int i = tid; i < n; i += PLBA_NT) P.xp[(size_t)6 * slot0 + i] = f ? 0.0 : x[i];
        if (tid == 0 && f) ctl.solve_fail = 1;
    PHASE_END
}
static inline size_t solve_small_smem(int n, int ldm) { return sizeof(double) * ((size_t)n * ldm + 4 * (size_t)n) + 16; }

// ---- tiled path -----------------------------------------------------------------------------------------------
// Factor the diagonal tile k (damping added here: later tiles only receive additive updates) and store inv(U_kk).
PLBA_KERNEL void k_potrf_tile(DevP P, int w, int k, double *invbuf) {
    PLBA_SMEM(raw);
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
PLBA_KERNEL void k_trsm_tiles(DevP P, int w, int k, const double *invbuf) {
    PLBA_SMEM(raw);
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
PLBA_KERNEL void k_syrk_tiles(DevP P, int w, int k, int nt) {
    PLBA_SMEM(raw);
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
PLBA_KERNEL void k_trisolve_large(DevP P, int w, int nt, const double *invbuf) {
    PLBA_SMEM(raw);
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
