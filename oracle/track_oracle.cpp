// TEST INFRASTRUCTURE ONLY.  CPU restatement of the Plücker-mode frame-to-frame pose tracking of kongan/PL-SLAM-plucker
// (SURVEY.md §8f row 3): StereoFrameHandler::optimizeFunctionsUsingPluker (src2/stereoFrameHandler.cpp:564-801) and
// gaussNewtonOptimizationforPluker (:803-853), with vector_stdv_mad (src2/auxiliar.cpp:444-460), robustWeightCauchy
// (:556-560), StereoFrame::lineSegmentOverlap (src2/stereoFrame.cpp:547-660), TransformForPluker
// (include2/stereoFrameHandler.h:114-122) and Eigen's ColPivHouseholderQR<Matrix6d> (restated from its published algorithm:
// Householder QR with column pivoting on the largest remaining column norm; logAbsDeterminant = sum log|R_ii|).
// PARITY UNPINNED: the reference ships no test or golden vector for this function either; pins are self-made (tests/).
// Only tests/ may load this; the product path (libplba.so) never does.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <vector>
#include "../include/plba.h"
#include "refmath.h"

using namespace oracle;

namespace {

// src2/auxiliar.cpp:444-460 (note the fabsf: the absolute deviations are rounded to float)
double vector_stdv_mad(std::vector<double> residues) {
    if (residues.size() != 0) {
        int n_samples = (int)residues.size();
        std::sort(residues.begin(), residues.end());
        double median = residues[n_samples / 2];
        for (int i = 0; i < n_samples; i++) residues[i] = fabsf(residues[i] - median);
        std::sort(residues.begin(), residues.end());
        double MAD = residues[n_samples / 2];
        return 1.4826 * MAD;
    }
    return 0.0;
}

// src2/stereoFrame.cpp:547-660
double lineSegmentOverlap(V2 spl_obs, V2 epl_obs, V2 spl_proj, V2 epl_proj) {
    double overlap = 1.f;
    auto by_lambda = [&](double lambda_s, double lambda_e) {
        double lambda_min = std::min(lambda_s, lambda_e), lambda_max = std::max(lambda_s, lambda_e);
        if (lambda_min < 0.f && lambda_max > 1.f) return 1.0;
        else if (lambda_max < 0.f || lambda_min > 1.f) return 0.0;
        else if (lambda_min < 0.f) return lambda_max;
        else if (lambda_max > 1.f) return 1.0 - lambda_min;
        return lambda_max - lambda_min;
    };
    if (std::fabs(spl_obs[0] - epl_obs[0]) < 1.0) {            // vertical lines
        V2 l = epl_obs - spl_obs;
        V2 spl_proj_line, epl_proj_line;
        spl_proj_line[0] = spl_obs[0]; spl_proj_line[1] = spl_proj[1];
        epl_proj_line[0] = epl_obs[0]; epl_proj_line[1] = epl_proj[1];
        overlap = by_lambda((spl_proj_line[1] - spl_obs[1]) / l[1], (epl_proj_line[1] - spl_obs[1]) / l[1]);
    } else if (std::fabs(spl_obs[1] - epl_obs[1]) < 1.0) {     // horizontal lines
        V2 l = epl_obs - spl_obs;
        V2 spl_proj_line, epl_proj_line;
        spl_proj_line[0] = spl_proj[0]; spl_proj_line[1] = spl_obs[1];
        epl_proj_line[0] = epl_proj[0]; epl_proj_line[1] = epl_obs[1];
        overlap = by_lambda((spl_proj_line[0] - spl_obs[0]) / l[0], (epl_proj_line[0] - spl_obs[0]) / l[0]);
    } else {                                                    // non-degenerate cases
        V2 l = epl_obs - spl_obs;
        double a = spl_obs[1] - epl_obs[1];
        double b = epl_obs[0] - spl_obs[0];
        double c = spl_obs[0] * epl_obs[1] - epl_obs[0] * spl_obs[1];
        V2 spl_proj_line, epl_proj_line;
        double lxy = 1.f / (a * a + b * b);
        spl_proj_line[0] = (b * (b * spl_proj[0] - a * spl_proj[1]) - a * c) * lxy;
        spl_proj_line[1] = (a * (-b * spl_proj[0] + a * spl_proj[1]) - b * c) * lxy;
        epl_proj_line[0] = (b * (b * epl_proj[0] - a * epl_proj[1]) - a * c) * lxy;
        epl_proj_line[1] = (a * (-b * epl_proj[0] + a * epl_proj[1]) - b * c) * lxy;
        overlap = by_lambda((spl_proj_line[0] - spl_obs[0]) / l[0], (epl_proj_line[0] - spl_obs[0]) / l[0]);
    }
    return overlap;
}

// include2/stereoFrameHandler.h:114-122
V6 TransformForPluker(const M4 &T, const V6 &vec) { return getTransformMatrixForPluker(T) * vec; }

struct Frame {
    const plba_track_frame *f; const plba_track_options *o;
    bool pt_in(int i) const { return !f->pt_inlier || f->pt_inlier[i]; }
    bool ls_in(int i) const { return !f->ls_inlier || f->ls_inlier[i]; }
};

// src2/stereoFrameHandler.cpp:564-801
void optimizeFunctionsUsingPluker(const Frame &F, const M4 &DT, M6 &H, V6 &g, double &e) {
    const plba_track_frame &f = *F.f; const double *cam = F.o->cam;
    M6 H_l, H_p; V6 g_l, g_p;
    double e_l = 0.0, e_p = 0.0;
    H = M6(); g = V6(); e = 0.0;
    std::vector<double> res_p, res_l;
    const M3 R = DT.block<3, 3>(0, 0); const V3 t = DT.block<3, 1>(0, 3);
    const M3 KL = plukerK(cam);
    auto V3at = [](const double *p, int i) { V3 v; v[0] = p[3 * i]; v[1] = p[3 * i + 1]; v[2] = p[3 * i + 2]; return v; };
    auto line_errs = [&](int i, double &err0, double &err1, V3 &pixel_line_curr) {
        V6 nd; for (int k = 0; k < 6; k++) nd[k] = f.ls_NDc[6 * i + k];
        V6 pluker_line_curr = TransformForPluker(DT, nd);
        pixel_line_curr = KL * pluker_line_curr.block<3, 1>(0, 0);
        const double *ob = f.ls_obs + 4 * i;
        err0 = ob[0] * pixel_line_curr[0] + ob[1] * pixel_line_curr[1] + pixel_line_curr[2];
        err0 = err0 / (std::sqrt(pixel_line_curr[0] * pixel_line_curr[0] + pixel_line_curr[1] * pixel_line_curr[1]));
        err1 = ob[2] * pixel_line_curr[0] + ob[3] * pixel_line_curr[1] + pixel_line_curr[2];
        err1 = err1 / (std::sqrt(pixel_line_curr[0] * pixel_line_curr[0] + pixel_line_curr[1] * pixel_line_curr[1]));
    };
    // pre-weight residuals (:576-607)
    for (int i = 0; i < f.n_pt; i++) if (F.pt_in(i)) {
        V3 P_ = R * V3at(f.pt_P, i) + t;
        V2 pl_proj = projection(cam, P_);
        V2 err_i; err_i[0] = pl_proj[0] - f.pt_obs[2 * i]; err_i[1] = pl_proj[1] - f.pt_obs[2 * i + 1];
        res_p.push_back(err_i.norm());
    }
    for (int i = 0; i < f.n_ls; i++) if (F.ls_in(i)) {
        double err0, err1; V3 l; line_errs(i, err0, err1, l);
        res_l.push_back(std::sqrt(err0 * err0 + err1 * err1));
    }
    // scale of the residuals (:609-651, the `else` branch)
    double s_p = 1.0, s_l = 1.0;
    const double th_min = 0.0001, th_max = std::sqrt(7.815);
    s_p = vector_stdv_mad(res_p); s_l = vector_stdv_mad(res_l);
    if (s_p < th_min) s_p = th_min;
    if (s_p > th_max) s_p = th_max;
    if (s_l < th_min) s_l = th_min;
    if (s_l > th_max) s_l = th_max;
    // point features (:653-698)
    int N_p = 0;
    for (int i = 0; i < f.n_pt; i++) if (F.pt_in(i)) {
        V3 P_ = R * V3at(f.pt_P, i) + t;
        V2 pl_proj = projection(cam, P_);
        V2 err_i; err_i[0] = pl_proj[0] - f.pt_obs[2 * i]; err_i[1] = pl_proj[1] - f.pt_obs[2 * i + 1];
        double err_i_norm = err_i.norm();
        double gx = P_[0], gy = P_[1], gz = P_[2], gz2 = gz * gz;
        double fgz2 = cam[0] / std::max(F.o->homog_th, gz2);
        double dx = err_i[0], dy = err_i[1];
        V6 J_aux;
        J_aux[0] = +fgz2 * dx * gz;
        J_aux[1] = +fgz2 * dy * gz;
        J_aux[2] = -fgz2 * (gx * dx + gy * dy);
        J_aux[3] = -fgz2 * (gx * gy * dx + gy * gy * dy + gz * gz * dy);
        J_aux[4] = +fgz2 * (gx * gx * dx + gz * gz * dx + gx * gy * dy);
        J_aux[5] = +fgz2 * (gx * gz * dy - gy * gz * dx);
        J_aux = J_aux / std::max(F.o->homog_th, err_i_norm);
        double r = err_i_norm;
        double x = r / s_p;
        double w = robustWeightCauchy(x);
        H_p += J_aux * J_aux.T() * w;
        g_p += J_aux * r * w;
        e_p += r * r * w;
        N_p++;
    }
    // line segment features (:700-782)
    int N_l = 0;
    for (int i = 0; i < f.n_ls; i++) if (F.ls_in(i)) {
        V3 sP_ = R * V3at(f.ls_sP, i) + t; V2 spl_proj = projection(cam, sP_);
        V3 eP_ = R * V3at(f.ls_eP, i) + t; V2 epl_proj = projection(cam, eP_);
        double err0, err1; V3 pixel_line_curr; line_errs(i, err0, err1, pixel_line_curr);
        const double *ob = f.ls_obs + 4 * i;
        double a0 = ob[0], b0 = ob[1], a1 = ob[2], b1 = ob[3];
        double lx = pixel_line_curr[0], ly = pixel_line_curr[1];
        double fm = 1.0 / std::sqrt(lx * lx + ly * ly);
        double err_i_norm = std::sqrt(err0 * err0 + err1 * err1);
        Mat<1, 3> fai_e0, fai_e1;
        fai_e0[0] = a0 * fm - lx * err0 * fm * fm; fai_e0[1] = b0 * fm - ly * err0 * fm * fm; fai_e0[2] = fm;
        fai_e1[0] = a1 * fm - lx * err1 * fm * fm; fai_e1[1] = b1 * fm - ly * err1 * fm * fm; fai_e1[2] = fm;
        Mat<3, 6> fai_pl_lc; fai_pl_lc.setBlock<3, 3>(0, 0, KL);
        V6 nd; for (int k = 0; k < 6; k++) nd[k] = f.ls_NDc[6 * i + k];
        V3 nn = nd.block<3, 1>(0, 0), dd = nd.block<3, 1>(3, 0);
        M6 fai_lc_RT;
        fai_lc_RT.setBlock<3, 3>(0, 3, -vechat(R * nn) - vechat(t) * vechat(R * dd));
        fai_lc_RT.setBlock<3, 3>(0, 0, -vechat(R * dd));
        Mat<1, 6> jac0 = fai_e0 * fai_pl_lc * fai_lc_RT;
        Mat<1, 6> jac1 = fai_e1 * fai_pl_lc * fai_lc_RT;
        V6 J_aux = (jac0.T() * err0 + jac1.T() * err1) / std::max(F.o->homog_th, err_i_norm);
        double s2 = f.ls_sigma2 ? f.ls_sigma2[i] : 1.0;
        double r = err_i_norm * std::sqrt(s2);
        double x = err_i_norm / s_l;
        double w = robustWeightCauchy(x);
        V2 spl, epl; spl[0] = f.ls_seg[4 * i]; spl[1] = f.ls_seg[4 * i + 1]; epl[0] = f.ls_seg[4 * i + 2]; epl[1] = f.ls_seg[4 * i + 3];
        double overlap = lineSegmentOverlap(spl, epl, spl_proj, epl_proj);
        w *= overlap;
        H_l += J_aux * J_aux.T() * w;
        g_l += J_aux * r * w;
        e_l += err_i_norm * err_i_norm * w;
        N_l++;
    }
    H = H_p + H_l; g = g_p + g_l; e = e_p + e_l;
    e /= (N_l + N_p);
}

// Eigen::ColPivHouseholderQR<Matrix6d>: solve + logAbsDeterminant
bool colPivHouseholderSolve(const M6 &A, const V6 &b, V6 &x, double &logAbsDet) {
    const int n = 6;
    double a[6][6], rhs[6]; int perm[6];
    for (int r = 0; r < n; r++) { rhs[r] = b[r]; perm[r] = r; for (int c = 0; c < n; c++) a[r][c] = A(r, c); }
    logAbsDet = 0.0;
    bool full_rank = true;
    for (int k = 0; k < n; k++) {
        int piv = k; double best = -1.0;
        for (int c = k; c < n; c++) { double s = 0; for (int r = k; r < n; r++) s += a[r][c] * a[r][c]; if (s > best) { best = s; piv = c; } }
        if (piv != k) { for (int r = 0; r < n; r++) std::swap(a[r][k], a[r][piv]); std::swap(perm[k], perm[piv]); }
        double norm = std::sqrt(best);
        if (norm == 0.0) { full_rank = false; break; }
        double alpha = (a[k][k] > 0) ? -norm : norm;
        double v[6] = {0, 0, 0, 0, 0, 0};
        for (int r = k; r < n; r++) v[r] = a[r][k];
        v[k] -= alpha;
        double vnorm2 = 0; for (int r = k; r < n; r++) vnorm2 += v[r] * v[r];
        if (vnorm2 > 0) {
            for (int c = k; c < n; c++) { double s = 0; for (int r = k; r < n; r++) s += v[r] * a[r][c]; s = 2.0 * s / vnorm2; for (int r = k; r < n; r++) a[r][c] -= s * v[r]; }
            double s = 0; for (int r = k; r < n; r++) s += v[r] * rhs[r]; s = 2.0 * s / vnorm2; for (int r = k; r < n; r++) rhs[r] -= s * v[r];
        }
        logAbsDet += std::log(std::fabs(a[k][k]));
    }
    if (!full_rank) { logAbsDet = -INFINITY; for (int i = 0; i < n; i++) x[i] = 0.0; return false; }
    double y[6];
    for (int r = n - 1; r >= 0; r--) { double s = rhs[r]; for (int c = r + 1; c < n; c++) s -= a[r][c] * y[c]; y[r] = s / a[r][r]; }
    for (int i = 0; i < n; i++) x[perm[i]] = y[i];
    return true;
}

}  // namespace

extern "C" {

// src2/stereoFrameHandler.cpp:803-853
int plba_track_oracle(const plba_track_frame *frame, const plba_track_options *opt, plba_track_result *res) {
    Frame F{frame, opt};
    M4 DT = M4::Identity(), DT_;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) DT(r, c) = frame->DT[4 * r + c];
    M6 H, DT_cov; V6 g, DT_inc;
    double err = 0, err_prev = 999999999.9;
    bool solution_is_good = true;
    DT_ = DT;
    int iters;
    for (iters = 0; iters < opt->max_iters; iters++) {
        optimizeFunctionsUsingPluker(F, DT, H, g, err);
        if ((std::fabs(err - err_prev) < opt->min_error_change) || (err < opt->min_error)) break;
        double logAbsDet;
        colPivHouseholderSolve(H, g, DT_inc, logAbsDet);
        if (logAbsDet < 0.0) { solution_is_good = false; break; }
        DT = inverse_se3(expmap_se3(DT_inc)) * DT;
        if (DT_inc.norm() < opt->min_error_change) break;
        err_prev = err;
    }
    double err_;
    if (solution_is_good) { DT_cov = inverse(H); err_ = err; }
    else { DT = DT_; err_ = -1.0; DT_cov = M6::Identity(); }
    for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) res->DT[4 * r + c] = DT(r, c);
    for (int i = 0; i < 36; i++) res->DT_cov[i] = DT_cov.m[i];
    res->err = err_; res->iters = (iters < opt->max_iters) ? iters + 1 : iters; res->good = solution_is_good ? 1 : 0;
    return 0;
}

}  // extern "C"

// ---- creation of Plücker line landmarks ------------------------------------------------------------
namespace {
// src2/stereoFrame.cpp:870-875
V4 pi_from_ppp(V3 x1, V3 x2, V3 x3) {
    V4 pi; V3 n = cross(x1 - x3, x2 - x3);
    pi[0] = n[0]; pi[1] = n[1]; pi[2] = n[2]; pi[3] = -dot(x3, cross(x1, x2));
    return pi;
}
// src2/stereoFrame.cpp:877-883
V6 pipi_plk(V4 pi1, V4 pi2) {
    V6 plk;
    M4 dp = pi1 * pi2.T() - pi2 * pi1.T();
    plk[0] = dp(0, 3); plk[1] = dp(1, 3); plk[2] = dp(2, 3); plk[3] = -dp(1, 2); plk[4] = dp(0, 2); plk[5] = -dp(0, 1);
    return plk;
}
}  // namespace

extern "C" int plba_create_lines_oracle(const plba_newline_batch *B, double *NDc, double *NDw, double *err_first, double *err_curr, uint8_t *accept) {
    const double *cam = B->cam; const double b = B->cam[4];
    const M3 KL = plukerK(cam);
    for (int i = 0; i < B->n; i++) {
        const double *sl = B->seg_l + 4 * i, *sr = B->seg_r + 4 * i;
        // backProjection_unit (src2/pinholeStereoCamera.cpp:215-223); src2/stereoFrame.cpp:383-397
        auto unit = [&](double u, double v) { V3 P; P[0] = (u - cam[2]) / cam[0]; P[1] = (v - cam[3]) / cam[1]; P[2] = 1.0; return P; };
        V3 obs1s = unit(sl[0], sl[1]), obs1e = unit(sl[2], sl[3]), obs2s = unit(sr[0], sr[1]), obs2e = unit(sr[2], sr[3]);
        obs2s[0] += b; obs2e[0] += b;
        V3 o1, o2; o2[0] = b;
        V6 line_pluker = pipi_plk(pi_from_ppp(obs1s, obs1e, o1), pi_from_ppp(obs2s, obs2e, o2));
        // src/mapHandler.cpp:449-459
        auto Tof = [&](int row) { M4 T = M4::Identity(); for (int r = 0; r < 3; r++) for (int c = 0; c < 4; c++) T(r, c) = B->kf_T_wc[12 * row + 4 * r + c]; return T; };
        M4 Tfw = Tof(B->kf_prev[i]);
        V6 plukerLW = getTransformMatrixForPluker(Tfw) * line_pluker;
        V3 nh = plukerLW.block<3, 1>(0, 0), dh = plukerLW.block<3, 1>(3, 0);
        double d = nh.norm() / dh.norm();
        dh = dh / dh.norm(); nh = nh / nh.norm();
        V6 new_pluker_lw; for (int k = 0; k < 3; k++) { new_pluker_lw[k] = nh[k] * d; new_pluker_lw[3 + k] = dh[k]; }
        auto reproj = [&](const M4 &Tkw, const double *pts) {      // :463-471 / :477-486
            V6 plukerLc = getTransformMatrixForPluker(inverse(Tkw)) * new_pluker_lw;
            V3 px = KL * plukerLc.block<3, 1>(0, 0);
            double fenmu = std::sqrt(px[0] * px[0] + px[1] * px[1]);
            V2 error; error[0] = (pts[0] * px[0] + pts[1] * px[1] + px[2]) / fenmu; error[1] = (pts[2] * px[0] + pts[3] * px[1] + px[2]) / fenmu;
            return error.norm();
        };
        double e1 = reproj(Tfw, sl), e2 = reproj(Tof(B->kf_curr[i]), B->seg_curr + 4 * i);
        if (NDc) for (int k = 0; k < 6; k++) NDc[6 * i + k] = line_pluker[k];
        if (NDw) for (int k = 0; k < 6; k++) NDw[6 * i + k] = new_pluker_lw[k];
        if (err_first) err_first[i] = e1;
        if (err_curr) err_curr[i] = e2;
        if (accept) accept[i] = (e2 > std::sqrt(5.991)) ? 0 : 1;   // :487
    }
    return 0;
}
