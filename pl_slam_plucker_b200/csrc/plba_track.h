// Frame-to-frame pose tracking in Plücker mode (the step before the LBA path):
//   StereoFrameHandler::gaussNewtonOptimizationforPluker   src2/stereoFrameHandler.cpp:803-853
//   StereoFrameHandler::optimizeFunctionsUsingPluker       src2/stereoFrameHandler.cpp:564-801
// One CTA per frame (frames of a batch are independent), everything of a frame in shared memory / registers:
//   per Gauss-Newton iteration: residuals -> MAD scales (the two medians are found by RANK SELECTION, every thread counts how
//   many residuals lie below its own: no sort, two barriers) -> weighted 6x6 normal equations (block reduction) ->
//   thread 0: stop tests, pivoted 6x6 solve with |det| check, DT <- exp(dx)^-1 DT.
#pragma once
#include "plba_math.h"
#include "../../include/plba.h"

namespace plba {

enum { TRK_NT = 256, TRK_MAX = 1024 };

struct TrackFrameDev { int n_pt, n_ls, pt0, ls0; double DT[12]; };
struct TrackP {
    const TrackFrameDev *frames;
    const double *pt_P, *pt_obs, *ls_sP, *ls_eP, *ls_NDc, *ls_obs, *ls_seg, *ls_s2;
    const unsigned char *pt_in, *ls_in;
    plba_track_result *res;
    double cam[4], homog_th, min_error, min_error_change;
    int max_iters, n_frames;
};

// point term (:656-698): residual norm r and J_aux (already divided by max(homogTh, r))
PLBA_HD void trk_point(const double *cam, const double *DT, const double *P, const double *obs, double homog_th, double &r, double *J) {
    double q[3]; rot(DT, P, q);
    const double gx = q[0] + DT[3], gy = q[1] + DT[7], gz = q[2] + DT[11];
    const double dx = (cam[2] + cam[0] * gx / gz) - obs[0], dy = (cam[3] + cam[1] * gy / gz) - obs[1];
    r = sqrt(dx * dx + dy * dy);
    const double gz2 = gz * gz;
    const double fgz2 = cam[0] / dmax(homog_th, gz2);             // (fx for both axes, as the reference)
    const double den = dmax(homog_th, r);
    J[0] = +fgz2 * dx * gz / den;
    J[1] = +fgz2 * dy * gz / den;
    J[2] = -fgz2 * (gx * dx + gy * dy) / den;
    J[3] = -fgz2 * (gx * gy * dx + gy * gy * dy + gz * gz * dy) / den;
    J[4] = +fgz2 * (gx * gx * dx + gz * gz * dx + gx * gy * dy) / den;
    J[5] = +fgz2 * (gx * gz * dy - gy * gz * dx) / den;
}
// image line of the transformed Plücker line and the two endpoint distances (:712-720)
PLBA_HD void trk_line_err(const double *cam, const double *DT, const double *nd, const double *ob, double *l, double *Rn, double *Rd, double &err0, double &err1) {
    g_line_project(*(const Cam *)cam, DT, nd, nd + 3, l, Rn, Rd);
    const double f = sqrt(l[0] * l[0] + l[1] * l[1]);
    err0 = (ob[0] * l[0] + ob[1] * l[1] + l[2]) / f;
    err1 = (ob[2] * l[0] + ob[3] * l[1] + l[2]) / f;
}
// line term (:722-757): J_aux = (jac0^T err0 + jac1^T err1) / max(homogTh, |err|)
PLBA_HD void trk_line(const double *cam, const double *DT, const double *nd, const double *ob, double homog_th, double &en, double *J) {
    double l[3], Rn[3], Rd[3], e[2];
    trk_line_err(cam, DT, nd, ob, l, Rn, Rd, e[0], e[1]);
    en = sqrt(e[0] * e[0] + e[1] * e[1]);
    const double fm = 1.0 / sqrt(l[0] * l[0] + l[1] * l[1]);
    const double t[3] = {DT[3], DT[7], DT[11]};
    const double den = dmax(homog_th, en);
    for (int c = 0; c < 6; c++) J[c] = 0.0;
    for (int k = 0; k < 2; k++) {
        const double f0 = ob[2 * k] * fm - l[0] * e[k] * fm * fm, f1 = ob[2 * k + 1] * fm - l[1] * e[k] * fm * fm, f2 = fm;
        const double g[3] = {f0 * cam[1] - f2 * cam[1] * cam[2], f1 * cam[0] - f2 * cam[0] * cam[3], f2 * cam[0] * cam[1]};     // fai_e * K_L
        double jt[3], c1[3], tg[3], c2[3];
        cross3(Rd, g, jt);                    // g^T (-hat(R d))
        cross3(Rn, g, c1); cross3(t, g, tg); cross3(Rd, tg, c2);      // g^T (-hat(R n) - hat(t) hat(R d)) = Rn x g - Rd x (t x g)
        for (int c = 0; c < 3; c++) { J[c] += jt[c] * e[k]; J[3 + c] += (c1[c] - c2[c]) * e[k]; }
    }
    for (int c = 0; c < 6; c++) J[c] /= den;
}
// StereoFrame::lineSegmentOverlap (src2/stereoFrame.cpp:547-660)
PLBA_HD double trk_overlap(const double *seg, const double *sp, const double *ep) {
    const double sx = seg[0], sy = seg[1], ex = seg[2], ey = seg[3];
    double ls, le;
    if (fabs(sx - ex) < 1.0) { const double l1 = ey - sy; ls = (sp[1] - sy) / l1; le = (ep[1] - sy) / l1; }                 // vertical
    else if (fabs(sy - ey) < 1.0) { const double l0 = ex - sx; ls = (sp[0] - sx) / l0; le = (ep[0] - sx) / l0; }            // horizontal
    else {
        const double l0 = ex - sx, a = sy - ey, b = ex - sx, c = sx * ey - ex * sy, lxy = 1.0 / (a * a + b * b);
        const double spx = (b * (b * sp[0] - a * sp[1]) - a * c) * lxy, epx = (b * (b * ep[0] - a * ep[1]) - a * c) * lxy;
        ls = (spx - sx) / l0; le = (epx - sx) / l0;
    }
    const double lmin = ls < le ? ls : le, lmax = ls < le ? le : ls;
    if (lmin < 0.0 && lmax > 1.0) return 1.0;
    if (lmax < 0.0 || lmin > 1.0) return 0.0;
    if (lmin < 0.0) return lmax;
    if (lmax > 1.0) return 1.0 - lmin;
    return lmax - lmin;
}
// 6x6 solve by Gaussian elimination with partial pivoting; returns log|det| (ColPivHouseholderQR::logAbsDeterminant, :822)
PLBA_HD double trk_solve6(const double *H, const double *g, double *x) {
    double a[6][7];
    for (int r = 0; r < 6; r++) { for (int c = 0; c < 6; c++) a[r][c] = H[r * 6 + c]; a[r][6] = g[r]; }
    double logdet = 0.0;
    for (int k = 0; k < 6; k++) {
        int p = k; double best = fabs(a[k][k]);
        for (int r = k + 1; r < 6; r++) if (fabs(a[r][k]) > best) { best = fabs(a[r][k]); p = r; }
        if (!(best > 0.0)) { for (int i = 0; i < 6; i++) x[i] = 0.0; return -1.0e300; }
        if (p != k) for (int c = 0; c < 7; c++) { const double tt = a[k][c]; a[k][c] = a[p][c]; a[p][c] = tt; }
        logdet += log(best);
        for (int r = k + 1; r < 6; r++) { const double f = a[r][k] / a[k][k]; for (int c = k; c < 7; c++) a[r][c] -= f * a[k][c]; }
    }
    for (int r = 5; r >= 0; r--) { double s = a[r][6]; for (int c = r + 1; c < 6; c++) s -= a[r][c] * x[c]; x[r] = s / a[r][r]; }
    return logdet;
}

// median by rank selection over the inlier entries of v[0..n): the entry with exactly n_in/2 inlier entries ordered before it
PLBA_D void trk_select(const double *v, const unsigned char *in, int n, int n_in, int tid, int nt, double *out) {
    for (int i = tid; i < n; i += nt) {
        if (in && !in[i]) continue;
        const double vi = v[i];
        int rank = 0;
        for (int j = 0; j < n; j++) { if (in && !in[j]) continue; const double vj = v[j]; rank += (vj < vi || (vj == vi && j < i)) ? 1 : 0; }
        if (rank == n_in / 2) *out = vi;
    }
}

PLBA_KERNEL void k_track_gn(TrackP T) {
    PLBA_SMEM(raw);
    double *rp = (double *)raw, *rl = rp + TRK_MAX, *red = rl + TRK_MAX;     // residuals / deviations; red[0..27] sums, [32..] scalars
    double *st = red + 64;                                                  // [0..11] DT, [12] err_prev, [13] s_p, [14] s_l, [15] med_p, [16] med_l, [17] err
    int *fl = (int *)(st + 24);                                             // [0] stop, [1] good, [2] iters, [3] n_in_p, [4] n_in_l
    for (int fidx = PLBA_BID; fidx < T.n_frames; fidx += PLBA_NB) {
        const TrackFrameDev F = T.frames[fidx];
        const double *pP = T.pt_P + (size_t)3 * F.pt0, *pO = T.pt_obs + (size_t)2 * F.pt0;
        const double *lS = T.ls_sP + (size_t)3 * F.ls0, *lE = T.ls_eP + (size_t)3 * F.ls0, *lN = T.ls_NDc + (size_t)6 * F.ls0;
        const double *lO = T.ls_obs + (size_t)4 * F.ls0, *lG = T.ls_seg + (size_t)4 * F.ls0, *l2 = T.ls_s2 + F.ls0;
        const unsigned char *pI = T.pt_in + F.pt0, *lI = T.ls_in + F.ls0;
        PHASE_BEGIN
            if (tid < 12) st[tid] = F.DT[tid];
            if (tid == 0) { st[12] = 999999999.9; st[17] = 0.0; fl[0] = 0; fl[1] = 1; fl[2] = 0; fl[3] = 0; fl[4] = 0; }
        PHASE_END
        PHASE_BEGIN
            int cp = 0, cl = 0;
            for (int i = tid; i < F.n_pt; i += PLBA_NT) cp += pI[i] ? 1 : 0;
            for (int i = tid; i < F.n_ls; i += PLBA_NT) cl += lI[i] ? 1 : 0;
            if (cp) plba_atomic_add_i(&fl[3], cp);
            if (cl) plba_atomic_add_i(&fl[4], cl);
        PHASE_END
        for (int it = 0; it < T.max_iters; it++) {
            // ---- residuals at the current DT (:576-607) ----
            PHASE_BEGIN
                if (tid < 28) red[tid] = 0.0;
                for (int i = tid; i < F.n_pt; i += PLBA_NT) if (pI[i]) { double r, J[6]; trk_point(T.cam, st, pP + 3 * i, pO + 2 * i, T.homog_th, r, J); rp[i] = r; }
                for (int i = tid; i < F.n_ls; i += PLBA_NT) if (lI[i]) {
                    double l[3], Rn[3], Rd[3], e0, e1; trk_line_err(T.cam, st, lN + 6 * i, lO + 4 * i, l, Rn, Rd, e0, e1); rl[i] = sqrt(e0 * e0 + e1 * e1);
                }
            PHASE_END
            // ---- vector_stdv_mad (src2/auxiliar.cpp:444-460): median, |x - median| rounded to float, median again ----
            PHASE_BEGIN
                if (tid == 0) { st[15] = 0.0; st[16] = 0.0; }
            PHASE_END
            PHASE_BEGIN
                trk_select(rp, pI, F.n_pt, fl[3], tid, PLBA_NT, &st[15]);
                trk_select(rl, lI, F.n_ls, fl[4], tid, PLBA_NT, &st[16]);
            PHASE_END
            PHASE_BEGIN
                for (int i = tid; i < F.n_pt; i += PLBA_NT) if (pI[i]) rp[i] = (double)fabsf((float)(rp[i] - st[15]));
                for (int i = tid; i < F.n_ls; i += PLBA_NT) if (lI[i]) rl[i] = (double)fabsf((float)(rl[i] - st[16]));
                if (tid == 0) { st[13] = 0.0; st[14] = 0.0; }
            PHASE_END
            PHASE_BEGIN
                trk_select(rp, pI, F.n_pt, fl[3], tid, PLBA_NT, &st[13]);
                trk_select(rl, lI, F.n_ls, fl[4], tid, PLBA_NT, &st[14]);
            PHASE_END
            // ---- weighted normal equations (:653-793) ----
            PHASE_BEGIN
                const double th_min = 0.0001, th_max = sqrt(7.815);
                double s_p = 1.4826 * st[13], s_l = 1.4826 * st[14];
                if (s_p < th_min) s_p = th_min; if (s_p > th_max) s_p = th_max;
                if (s_l < th_min) s_l = th_min; if (s_l > th_max) s_l = th_max;
                double acc[28];
#pragma unroll
                for (int k = 0; k < 28; k++) acc[k] = 0.0;
                for (int i = tid; i < F.n_pt; i += PLBA_NT) if (pI[i]) {
                    double r, J[6]; trk_point(T.cam, st, pP + 3 * i, pO + 2 * i, T.homog_th, r, J);
                    const double x = r / s_p, w = 1.0 / (1.0 + x * x);
                    int q = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) {
#pragma unroll
                        for (int b = a; b < 6; b++) acc[q++] += J[a] * J[b] * w;
                        acc[21 + a] += J[a] * r * w;
                    }
                    acc[27] += r * r * w;
                }
                for (int i = tid; i < F.n_ls; i += PLBA_NT) if (lI[i]) {
                    double en, J[6]; trk_line(T.cam, st, lN + 6 * i, lO + 4 * i, T.homog_th, en, J);
                    double q3[3], sPc[3], ePc[3], sp[2], ep[2];
                    rot(st, lS + 3 * i, q3); sPc[0] = q3[0] + st[3]; sPc[1] = q3[1] + st[7]; sPc[2] = q3[2] + st[11];
                    rot(st, lE + 3 * i, q3); ePc[0] = q3[0] + st[3]; ePc[1] = q3[1] + st[7]; ePc[2] = q3[2] + st[11];
                    sp[0] = T.cam[2] + T.cam[0] * sPc[0] / sPc[2]; sp[1] = T.cam[3] + T.cam[1] * sPc[1] / sPc[2];
                    ep[0] = T.cam[2] + T.cam[0] * ePc[0] / ePc[2]; ep[1] = T.cam[3] + T.cam[1] * ePc[1] / ePc[2];
                    const double r = en * sqrt(l2[i]);
                    const double x = en / s_l;
                    const double w = (1.0 / (1.0 + x * x)) * trk_overlap(lG + 4 * i, sp, ep);
                    int q = 0;
#pragma unroll
                    for (int a = 0; a < 6; a++) {
#pragma unroll
                        for (int b = a; b < 6; b++) acc[q++] += J[a] * J[b] * w;
                        acc[21 + a] += J[a] * r * w;
                    }
                    acc[27] += en * en * w;
                }
#pragma unroll
                for (int k = 0; k < 28; k++) plba_block_add(&red[k], acc[k]);
            PHASE_END
            // ---- Gauss-Newton step (:812-836) ----
            PHASE_BEGIN
                if (tid == 0) {
                    double H[36], g[6], dx[6];
                    int q = 0;
                    for (int a = 0; a < 6; a++) { for (int b = a; b < 6; b++) { H[a * 6 + b] = red[q]; H[b * 6 + a] = red[q]; q++; } g[a] = red[21 + a]; }
                    const double err = red[27] / (double)(fl[3] + fl[4]);
                    st[17] = err; fl[2] = it + 1;
                    for (int k = 0; k < 36; k++) red[28 + k] = H[k];          // kept for DT_cov = H^-1 (:841)
                    if ((fabs(err - st[12]) < T.min_error_change) || (err < T.min_error)) fl[0] = 1;
                    else {
                        const double logdet = trk_solve6(H, g, dx);
                        if (logdet < 0.0) { fl[1] = 0; fl[0] = 1; }
                        else {
                            double Td[12], Tdi[12], Tn[12];
                            exp_se3(dx, Td); inv_se3(Td, Tdi); mul_se3(Tdi, st, Tn);       // DT << inverse_se3(expmap_se3(DT_inc)) * DT
                            for (int k = 0; k < 12; k++) st[k] = Tn[k];
                            double nn = 0; for (int k = 0; k < 6; k++) nn += dx[k] * dx[k];
                            if (sqrt(nn) < T.min_error_change) fl[0] = 1;
                            st[12] = err;
                        }
                    }
                }
            PHASE_END
            if (fl[0]) break;
        }
        PHASE_BEGIN
            if (tid == 0) {
                plba_track_result &R = T.res[fidx];
                if (fl[1]) {
                    double Hi[36];
                    for (int k = 0; k < 36; k++) Hi[k] = red[28 + k];
                    gj_inverse<6>(Hi);
                    for (int k = 0; k < 36; k++) R.DT_cov[k] = Hi[k];
                    for (int k = 0; k < 12; k++) R.DT[k] = st[k];
                    R.err = st[17];
                } else {
                    for (int k = 0; k < 36; k++) R.DT_cov[k] = (k % 7 == 0) ? 1.0 : 0.0;
                    for (int k = 0; k < 12; k++) R.DT[k] = F.DT[k];
                    R.err = -1.0;
                }
                R.iters = fl[2]; R.good = fl[1];
            }
        PHASE_END
    }
}
// ---- creation of Plücker line landmarks -----------------------------------------------------------
struct NewLineP {
    int n; double cam[5];
    const double *seg_l, *seg_r, *seg_curr, *kf_T; const int *kf_prev, *kf_curr;
    double *NDc, *NDw, *err_first, *err_curr; unsigned char *accept;
};
// StereoFrame::pi_from_ppp (src2/stereoFrame.cpp:870-875): plane through three points
PLBA_HD void nl_pi_from_ppp(const double *x1, const double *x2, const double *x3, double *pi) {
    const double a[3] = {x1[0] - x3[0], x1[1] - x3[1], x1[2] - x3[2]}, b[3] = {x2[0] - x3[0], x2[1] - x3[1], x2[2] - x3[2]};
    double c12[3]; cross3(a, b, pi); cross3(x1, x2, c12);
    pi[3] = -dot3(x3, c12);
}
// distance of two pixels to the image line of a world Plücker line seen from the keyframe with pose T_kf_w (:463-471, :477-486)
PLBA_HD double nl_reproj(const double *cam, const double *Twc, const double *NDw, const double *seg) {
    double Tcw[12], l[3], Rn[3], Rd[3];
    inv_se3(Twc, Tcw);
    g_line_project(*(const Cam *)cam, Tcw, NDw, NDw + 3, l, Rn, Rd);
    const double f = sqrt(l[0] * l[0] + l[1] * l[1]);
    const double e0 = (seg[0] * l[0] + seg[1] * l[1] + l[2]) / f, e1 = (seg[2] * l[0] + seg[3] * l[1] + l[2]) / f;
    return sqrt(e0 * e0 + e1 * e1);
}
PLBA_HD void nl_create(const NewLineP &Q, int i) {
    const double fx = Q.cam[0], fy = Q.cam[1], cx = Q.cam[2], cy = Q.cam[3], b = Q.cam[4];
    const double *sl = Q.seg_l + 4 * i, *sr = Q.seg_r + 4 * i;
    // back-projection to the normalised plane (backProjection_unit, src2/pinholeStereoCamera.cpp:215-223); right camera at (b, 0, 0)
    const double o1s[3] = {(sl[0] - cx) / fx, (sl[1] - cy) / fy, 1.0}, o1e[3] = {(sl[2] - cx) / fx, (sl[3] - cy) / fy, 1.0};
    const double o2s[3] = {(sr[0] - cx) / fx + b, (sr[1] - cy) / fy, 1.0}, o2e[3] = {(sr[2] - cx) / fx + b, (sr[3] - cy) / fy, 1.0};
    const double c1[3] = {0, 0, 0}, c2[3] = {b, 0, 0};
    double p1[4], p2[4];
    nl_pi_from_ppp(o1s, o1e, c1, p1); nl_pi_from_ppp(o2s, o2e, c2, p2);
    // pipi_plk (:877-883): dp = pi1 pi2^T - pi2 pi1^T ; plk = (dp03, dp13, dp23, -dp12, dp02, -dp01)
#define NL_DP(r, c) (p1[r] * p2[c] - p2[r] * p1[c])
    const double nd[6] = {NL_DP(0, 3), NL_DP(1, 3), NL_DP(2, 3), -NL_DP(1, 2), NL_DP(0, 2), -NL_DP(0, 1)};
#undef NL_DP
    if (Q.NDc) for (int k = 0; k < 6; k++) Q.NDc[6 * i + k] = nd[k];
    // into the world: TransformForPluker(T_kf_w, NDc), then |d| = 1 and |n| = |n| / |d| (src/mapHandler.cpp:449-459)
    const double *Twp = Q.kf_T + 12 * Q.kf_prev[i];
    double Rn[3], Rd[3], tx[3];
    rot(Twp, nd, Rn); rot(Twp, nd + 3, Rd);
    const double t[3] = {Twp[3], Twp[7], Twp[11]};
    cross3(t, Rd, tx);
    double nw[3] = {Rn[0] + tx[0], Rn[1] + tx[1], Rn[2] + tx[2]};
    const double nn = sqrt(dot3(nw, nw)), dn = sqrt(dot3(Rd, Rd)), d = nn / dn;
    double W[6];
    for (int k = 0; k < 3; k++) { W[k] = nw[k] / nn * d; W[3 + k] = Rd[k] / dn; }
    if (Q.NDw) for (int k = 0; k < 6; k++) Q.NDw[6 * i + k] = W[k];
    const double e1 = nl_reproj(Q.cam, Twp, W, sl);
    const double e2 = nl_reproj(Q.cam, Q.kf_T + 12 * Q.kf_curr[i], W, Q.seg_curr + 4 * i);
    if (Q.err_first) Q.err_first[i] = e1;
    if (Q.err_curr) Q.err_curr[i] = e2;
    if (Q.accept) Q.accept[i] = (e2 > sqrt(5.991)) ? 0 : 1;                       // :487
}
PLBA_KERNEL void k_create_lines(NewLineP Q) {
    PHASE_BEGIN
        for (int i = PLBA_BID * PLBA_NT + tid; i < Q.n; i += PLBA_NB * PLBA_NT) nl_create(Q, i);
    PHASE_END
}
static inline size_t track_smem() { return sizeof(double) * (2 * TRK_MAX + 64 + 64 + 24) + 64; }

}  // namespace plba
