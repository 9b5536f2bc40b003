// Per-observation and per-landmark FP64 arithmetic of the LBA path, written out in scalar form for the GPU.
// Each function cites the reference lines whose RESULT it reproduces (the reference computes them with generic
// Eigen 6x6 / 6x4 products per edge; here the products are expanded by hand and everything that depends only on
// the landmark or only on the keyframe is hoisted out of the per-observation path).
#pragma once
#include "plba_port.h"

namespace plba {

struct Cam { double fx, fy, cx, cy; };

PLBA_HD void cross3(const double *a, const double *b, double *c) {
    c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
PLBA_HD double dot3(const double *a, const double *b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
// T = rows of a 3x4 [R|t]
PLBA_HD void rot(const double *T, const double *v, double *o) {
    o[0] = T[0] * v[0] + T[1] * v[1] + T[2] * v[2];
    o[1] = T[4] * v[0] + T[5] * v[1] + T[6] * v[2];
    o[2] = T[8] * v[0] + T[9] * v[1] + T[10] * v[2];
}
PLBA_HD void rotT(const double *T, const double *v, double *o) {   // R^T v
    o[0] = T[0] * v[0] + T[4] * v[1] + T[8] * v[2];
    o[1] = T[1] * v[0] + T[5] * v[1] + T[9] * v[2];
    o[2] = T[2] * v[0] + T[6] * v[1] + T[10] * v[2];
}
// [R|t]^-1 = [R^T | -R^T t]          (inverse_se3, src2/auxiliar.cpp:113-122)
PLBA_HD void inv_se3(const double *T, double *o) {
    double t[3] = {T[3], T[7], T[11]}, rt[3];
    rotT(T, t, rt);
    o[0] = T[0]; o[1] = T[4]; o[2] = T[8];  o[3] = -rt[0];
    o[4] = T[1]; o[5] = T[5]; o[6] = T[9];  o[7] = -rt[1];
    o[8] = T[2]; o[9] = T[6]; o[10] = T[10]; o[11] = -rt[2];
}
PLBA_HD void mul_se3(const double *A, const double *B, double *o) {   // o = A*B on 3x4 rows
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) o[r * 4 + c] = A[r * 4] * B[c] + A[r * 4 + 1] * B[4 + c] + A[r * 4 + 2] * B[8 + c];
        o[r * 4 + 3] = A[r * 4] * B[3] + A[r * 4 + 1] * B[7] + A[r * 4 + 2] * B[11] + A[r * 4 + 3];
    }
}
// expmap_se3, src2/auxiliar.cpp:124-141   x = [rho; phi]
PLBA_HD void exp_se3(const double *x, double *T) {
    const double w0 = x[3], w1 = x[4], w2 = x[5];
    const double th = sqrt(w0 * w0 + w1 * w1 + w2 * w2);
    if (th < 0.000001) {
        T[0] = 1; T[1] = 0; T[2] = 0; T[3] = x[0]; T[4] = 0; T[5] = 1; T[6] = 0; T[7] = x[1]; T[8] = 0; T[9] = 0; T[10] = 1; T[11] = x[2];
        return;
    }
    const double s[9] = {0, -w2 / th, w1 / th, w2 / th, 0, -w0 / th, -w1 / th, w0 / th, 0};
    double ss[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) ss[r * 3 + c] = s[r * 3] * s[c] + s[r * 3 + 1] * s[3 + c] + s[r * 3 + 2] * s[6 + c];
    const double sn = sin(th), cs = cos(th);
    double V[9];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
        const double id = (r == c) ? 1.0 : 0.0;
        T[r * 4 + c] = id + s[r * 3 + c] * sn + ss[r * 3 + c] * (1.0 - cs);
        V[r * 3 + c] = id + s[r * 3 + c] * (1.0 - cs) / th + ss[r * 3 + c] * (th - sn) / th;
    }
    for (int r = 0; r < 3; r++) T[r * 4 + 3] = V[r * 3] * x[0] + V[r * 3 + 1] * x[1] + V[r * 3 + 2] * x[2];
}
PLBA_HD void inv3(const double *V, double *o) {
    const double c00 = V[4] * V[8] - V[5] * V[7], c01 = V[5] * V[6] - V[3] * V[8], c02 = V[3] * V[7] - V[4] * V[6];
    const double det = V[0] * c00 + V[1] * c01 + V[2] * c02, id = 1.0 / det;
    o[0] = c00 * id; o[1] = (V[2] * V[7] - V[1] * V[8]) * id; o[2] = (V[1] * V[5] - V[2] * V[4]) * id;
    o[3] = c01 * id; o[4] = (V[0] * V[8] - V[2] * V[6]) * id; o[5] = (V[2] * V[3] - V[0] * V[5]) * id;
    o[6] = c02 * id; o[7] = (V[1] * V[6] - V[0] * V[7]) * id; o[8] = (V[0] * V[4] - V[1] * V[3]) * id;
}
// logmap_se3, src2/auxiliar.cpp:143-173
PLBA_HD void log_se3(const double *T, double *x) {
    double cosine = ((T[0] + T[5] + T[10]) - 1.0) / 2.0;
    if (cosine > 1.0) cosine = 1.0; else if (cosine < -1.0) cosine = -1.0;
    double sine = sqrt(1.0 - cosine * cosine);
    if (sine > 1.0) sine = 1.0;
    const double th = acos(cosine);
    double w[3] = {0, 0, 0};
    double V[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    if (th > 0.000001) {
        const double den = 2.0 * sine;
        w[0] = (T[9] - T[6]) * th / den; w[1] = (T[2] - T[8]) * th / den; w[2] = (T[4] - T[1]) * th / den;
        const double s[9] = {0, -w[2] / th, w[1] / th, w[2] / th, 0, -w[0] / th, -w[1] / th, w[0] / th, 0};
        for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) {
            const double ssrc = s[r * 3] * s[c] + s[r * 3 + 1] * s[3 + c] + s[r * 3 + 2] * s[6 + c];
            V[r * 3 + c] = ((r == c) ? 1.0 : 0.0) + s[r * 3 + c] * (1.0 - cosine) / th + ssrc * (th - sine) / th;
        }
    }
    double Vi[9]; inv3(V, Vi);
    const double t[3] = {T[3], T[7], T[11]};
    for (int r = 0; r < 3; r++) x[r] = Vi[r * 3] * t[0] + Vi[r * 3 + 1] * t[1] + Vi[r * 3 + 2] * t[2];
    x[3] = w[0]; x[4] = w[1]; x[5] = w[2];
}

// VertexLMPose::oplusImpl, g2o_types/g2o_types.h:172-203:  R <- dR(omega) R ; t <- t + dt   (delta = [dt; omega])
PLBA_HD void pose_oplus_g(const double *T, const double *dl, double *o) {
    const double ox = dl[3], oy = dl[4], oz = dl[5];
    const double th = sqrt(ox * ox + oy * oy + oz * oz), h = 0.5 * th;
    double im;
    const double re = cos(h);
    if (th < 1e-10) { const double t2 = th * th, t4 = t2 * t2; im = 0.5 - 0.0208333 * t2 + 0.000260417 * t4; }
    else im = sin(h) / th;
    const double x = im * ox, y = im * oy, z = im * oz;
    const double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    const double twx = tx * re, twy = ty * re, twz = tz * re, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    const double dR[9] = {1 - (tyy + tzz), txy - twz, txz + twy, txy + twz, 1 - (txx + tzz), tyz - twx, txz - twy, tyz + twx, 1 - (txx + tyy)};
    for (int r = 0; r < 3; r++) {
        for (int c = 0; c < 3; c++) o[r * 4 + c] = dR[r * 3] * T[c] + dR[r * 3 + 1] * T[4 + c] + dR[r * 3 + 2] * T[8 + c];
        o[r * 4 + 3] = T[r * 4 + 3] + dl[r];
    }
}

// R(theta) = Rz(t3) Ry(t2) Rx(t1), written out as src/mapFeatures.cpp:211-215
PLBA_HD void orth_R(const double *o, double *R) {
    double s1, c1, s2, c2, s3, c3;
    plba_sincos(o[0], &s1, &c1); plba_sincos(o[1], &s2, &c2); plba_sincos(o[2], &s3, &c3);      // one range reduction per angle
    R[0] = c2 * c3; R[1] = s1 * s2 * c3 - c1 * s3; R[2] = c1 * s2 * c3 + s1 * s3;
    R[3] = c2 * s3; R[4] = s1 * s2 * s3 + c1 * c3; R[5] = c1 * s2 * s3 - s1 * c3;
    R[6] = -s2;     R[7] = s1 * c2;                R[8] = c1 * c2;
}
// changeOrthToPluker, src/mapFeatures.cpp:203-224 / g2o_types.h:367-387 :  n = cos(phi) R(:,0), d = sin(phi) R(:,1)
PLBA_HD void orth_to_plk(const double *o, double *pl) {
    double R[9]; orth_R(o, R);
    double w1, w2; plba_sincos(o[3], &w2, &w1);
    pl[0] = w1 * R[0]; pl[1] = w1 * R[3]; pl[2] = w1 * R[6];
    pl[3] = w2 * R[1]; pl[4] = w2 * R[4]; pl[5] = w2 * R[7];
}
// changeOrthToPluker with one sincos per angle (src/mapFeatures.cpp:203-224)
PLBA_HD void orth_to_plk_sc(const double *o, double *pl) {
    double s1, c1, s2, c2, s3, c3, w1, w2;
    plba_sincos(o[0], &s1, &c1); plba_sincos(o[1], &s2, &c2); plba_sincos(o[2], &s3, &c3); plba_sincos(o[3], &w2, &w1);
    pl[0] = w1 * (c2 * c3); pl[1] = w1 * (c2 * s3); pl[2] = w1 * (-s2);
    pl[3] = w2 * (s1 * s2 * c3 - c1 * s3); pl[4] = w2 * (s1 * s2 * s3 + c1 * c3); pl[5] = w2 * (s1 * c2);
}
// Per-LANDMARK quantities of a Plücker line: the 6-vector, U = [n/|n|, d/|d|, nxd/|nxd|] and W = (|n|,|d|)/sqrt(|n|^2+|d|^2)
// (getOrhtRFromPluker / getOrthWFromPluker, g2o_types.h:472-495).  The reference recomputes these per edge.
struct LinePre { double n[3], d[3], u1[3], u2[3], u3[3], w1, w2; };
PLBA_HD void line_pre_from_plk(const double *pl, LinePre &L) {
    for (int i = 0; i < 3; i++) { L.n[i] = pl[i]; L.d[i] = pl[3 + i]; }
    // four reciprocal square roots instead of four square roots and eleven divisions (this sits on the latency path of every line chunk)
    const double n2 = dot3(L.n, L.n), d2 = dot3(L.d, L.d);
    const double ni = plba_rsqrt_hd(n2), di = plba_rsqrt_hd(d2);
    for (int i = 0; i < 3; i++) { L.u1[i] = L.n[i] * ni; L.u2[i] = L.d[i] * di; }
    double c[3]; cross3(L.n, L.d, c);
    const double ci = plba_rsqrt_hd(dot3(c, c));
    for (int i = 0; i < 3; i++) L.u3[i] = c[i] * ci;
    const double fi = plba_rsqrt_hd(n2 + d2);
    L.w1 = n2 * ni * fi; L.w2 = d2 * di * fi;
}
// changePlukerToOrth, src/mapFeatures.cpp:186-201
PLBA_HD void plk_to_orth(const double *pl, double *o) {
    LinePre L; line_pre_from_plk(pl, L);
    o[0] = atan2(L.u2[2], L.u3[2]);
    o[1] = asin(-L.u1[2]);
    o[2] = atan2(L.u1[1], L.u1[0]);
    o[3] = asin(L.w2);
}
// updateOrthCoord, include/mapHandler.h:252-335 / g2o_types.h:72-130:  U <- U Rx(d1) Ry(d2) Rz(d3), W <- W W(d4)
PLBA_HD void orth_update(const double *D, const double *dl, double *o) {
    double R[9]; orth_R(D, R);
    double w1, w2, sx, cx, sy, cy, sz, cz, s4, c4;
    plba_sincos(D[3], &w2, &w1); plba_sincos(dl[0], &sx, &cx); plba_sincos(dl[1], &sy, &cy); plba_sincos(dl[2], &sz, &cz); plba_sincos(dl[3], &s4, &c4);
    // M = Rx*Ry*Rz
    const double M[9] = {cy * cz, -cy * sz, sy,
                         sx * sy * cz + cx * sz, -sx * sy * sz + cx * cz, -sx * cy,
                         -cx * sy * cz + sx * sz, cx * sy * sz + sx * cz, cx * cy};
    // the reference multiplies left to right: ((R*Rx)*Ry)*Rz ; only rows needed for the extraction are formed
    double RX[9], RXY[9], Rn[9];
    const double Rx[9] = {1, 0, 0, 0, cx, -sx, 0, sx, cx}, Ry[9] = {cy, 0, sy, 0, 1, 0, -sy, 0, cy}, Rz[9] = {cz, -sz, 0, sz, cz, 0, 0, 0, 1};
    (void)M;
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) RX[r * 3 + c] = R[r * 3] * Rx[c] + R[r * 3 + 1] * Rx[3 + c] + R[r * 3 + 2] * Rx[6 + c];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) RXY[r * 3 + c] = RX[r * 3] * Ry[c] + RX[r * 3 + 1] * Ry[3 + c] + RX[r * 3 + 2] * Ry[6 + c];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) Rn[r * 3 + c] = RXY[r * 3] * Rz[c] + RXY[r * 3 + 1] * Rz[3 + c] + RXY[r * 3 + 2] * Rz[6 + c];
    const double W10 = w2 * c4 + w1 * s4;
    o[0] = atan2(Rn[7], Rn[8]);
    o[1] = asin(-Rn[6]);
    o[2] = atan2(Rn[3], Rn[0]);
    o[3] = asin(W10);
}

// RobustKernelHuber::robustify (g2o), rho0 and rho1 only
PLBA_HD void huber(double delta, double e, double &rho0, double &rho1) {
    const double dsqr = delta * delta;
    if (e <= dsqr) { rho0 = e; rho1 = 1.0; }
    else { const double isq = plba_rsqrt_safe(e), sq = e * isq; rho0 = 2 * sq * delta - dsqr; rho1 = delta * isq; }      // one reciprocal square root (7 dependent operations) instead of sqrt + division
}

// ---------------------------------------------------------------------------------------------------
// Profile G edges.  A = d e / d pose (2x6, [dt; omega]),  B = d e / d landmark (2xD)
// ---------------------------------------------------------------------------------------------------
// EdgePosePoint::computeError, g2o_types.h:224-230,232-263
PLBA_HD void g_point_error(const Cam &cam, const double *T, const double *Pw, const double *uv, double *e, double &zc) {
    double q[3]; rot(T, Pw, q);
    const double x = q[0] + T[3], y = q[1] + T[7], z = q[2] + T[11];
    const double invz = plba_rcp_safe(z);
    e[0] = uv[0] - (x * invz * cam.fx + cam.cx);
    e[1] = uv[1] - (y * invz * cam.fy + cam.cy);
    zc = z;
}
// EdgePosePoint::linearizeOplus, g2o_types.h:271-296
PLBA_HD void g_point_lin(const Cam &cam, const double *T, const double *Pw, const double *uv, double *e, double *A, double *B) {
    double q[3]; rot(T, Pw, q);                       // q = R Pw
    const double x = q[0] + T[3], y = q[1] + T[7], z = q[2] + T[11];
    // ONE reciprocal (hardware seed + two Newton steps) replaces the five IEEE divisions of the reference's expressions: each of those is
    // a ~15-instruction subroutine on the GPU, i.e. a quarter of this function; the results differ from the oracle's in the last bit
    const double invz = plba_rcp_safe(z), invz2 = invz * invz;
    e[0] = uv[0] - (x * invz * cam.fx + cam.cx);
    e[1] = uv[1] - (y * invz * cam.fy + cam.cy);
    const double a = cam.fx * invz, b = -cam.fx * x * invz2, c = cam.fy * invz, d = -cam.fy * y * invz2;
    // Jxi = -jpc R
    B[0] = -(a * T[0] + b * T[8]); B[1] = -(a * T[1] + b * T[9]); B[2] = -(a * T[2] + b * T[10]);
    B[3] = -(c * T[4] + d * T[8]); B[4] = -(c * T[5] + d * T[9]); B[5] = -(c * T[6] + d * T[10]);
    // Jxj = [-jpc , jpc * hat(R Pw)]
    A[0] = -a; A[1] = 0.0; A[2] = -b; A[3] = -b * q[1];           A[4] = -a * q[2] + b * q[0]; A[5] = a * q[1];
    A[6] = 0.0; A[7] = -c; A[8] = -d; A[9] = c * q[2] - d * q[1]; A[10] = d * q[0];            A[11] = -c * q[0];
}
// camera-frame moment n_c = R n + t x (R d), image line l = K_L n_c  (g2o_types.h:329-343, 349-365)
PLBA_HD void g_line_project(const Cam &cam, const double *T, const double *n, const double *d, double *l, double *Rn, double *Rd) {
    rot(T, n, Rn); rot(T, d, Rd);
    const double t[3] = {T[3], T[7], T[11]};
    double tx[3]; cross3(t, Rd, tx);
    const double nc0 = Rn[0] + tx[0], nc1 = Rn[1] + tx[1], nc2 = Rn[2] + tx[2];
    l[0] = cam.fy * nc0; l[1] = cam.fx * nc1; l[2] = -cam.fy * cam.cx * nc0 - cam.fx * cam.cy * nc1 + cam.fx * cam.fy * nc2;
}
PLBA_HD void g_line_error(const Cam &cam, const double *T, const double *n, const double *d, const double *ab, double *e) {
    double l[3], Rn[3], Rd[3]; g_line_project(cam, T, n, d, l, Rn, Rd);
    const double ifen = plba_rsqrt_safe(l[0] * l[0] + l[1] * l[1]);
    e[0] = (l[0] * ab[0] + l[1] * ab[1] + l[2]) * ifen;
    e[1] = (l[0] * ab[2] + l[1] * ab[3] + l[2]) * ifen;
}
// EdgePoseLine::linearizeOplus, g2o_types.h:389-453.  head/tail: the 3-vectors the pose Jacobian is built from
// (Q12: the reference feeds orth.head(3)/orth.tail(3); intended maths feeds n / d).
PLBA_HD void g_line_lin(const Cam &cam, const double *T, const LinePre &L, const double *head, const double *tail, const double *ab,
                        double *e, double *A, double *B) {
    double l[3], Rn[3], Rd[3]; g_line_project(cam, T, L.n, L.d, l, Rn, Rd);
    const double ifen = plba_rsqrt_safe(l[0] * l[0] + l[1] * l[1]), ifen2 = ifen * ifen;      // 1 / fenmu and 1 / fenmu^2: one reciprocal square root for the sqrt and the seven divisions of :406-421
    e[0] = (l[0] * ab[0] + l[1] * ab[1] + l[2]) * ifen;
    e[1] = (l[0] * ab[2] + l[1] * ab[3] + l[2]) * ifen;
    const double t[3] = {T[3], T[7], T[11]};
    double Rh[3], Rt[3];
    rot(T, head, Rh); rot(T, tail, Rt);
    for (int k = 0; k < 2; k++) {
        const double j0 = -l[0] * e[k] * ifen2 + ab[2 * k] * ifen, j1 = -l[1] * e[k] * ifen2 + ab[2 * k + 1] * ifen, j2 = ifen;
        // g = j * K_L  (1x3, derivative w.r.t. n_c)
        const double g[3] = {j0 * cam.fy - j2 * cam.fy * cam.cx, j1 * cam.fx - j2 * cam.fx * cam.cy, j2 * cam.fx * cam.fy};
        // pose:  g^T(-hat(Rt)) = Rt x g ;  g^T(-hat(Rh) - hat(t) hat(Rt)) = Rh x g - (g x t) x Rt
        double c0[3], gt[3], c1[3], c2[3];
        cross3(Rt, g, c0); cross3(Rh, g, c1); cross3(g, t, gt); cross3(gt, Rt, c2);
        A[6 * k + 0] = c0[0]; A[6 * k + 1] = c0[1]; A[6 * k + 2] = c0[2];
        A[6 * k + 3] = c1[0] - c2[0]; A[6 * k + 4] = c1[1] - c2[1]; A[6 * k + 5] = c1[2] - c2[2];
        // landmark:  g^T [R | hat(t) R] * jacobianFromPlukerToOrth(U, W)   (g2o_types.h:455-470)
        double gr[3], gd[3];
        rotT(T, g, gr); rotT(T, gt, gd);
        B[4 * k + 0] = L.w2 * dot3(gd, L.u3);
        B[4 * k + 1] = -L.w1 * dot3(gr, L.u3);
        B[4 * k + 2] = L.w1 * dot3(gr, L.u2) - L.w2 * dot3(gd, L.u1);
        B[4 * k + 3] = -L.w2 * dot3(gr, L.u1) + L.w1 * dot3(gd, L.u2);
    }
}

// ---------------------------------------------------------------------------------------------------
// Profile H terms (scalar residual r = |e|, one Jacobian row, Cauchy weight).  Tiw = inverse_se3(T_kf_w).
// ---------------------------------------------------------------------------------------------------
PLBA_HD double dmax(double a, double b) { return a > b ? a : b; }   // std::max semantics for finite values
// src/mapHandler.cpp:2368-2411 / :2610-2658
PLBA_HD void h_point(const Cam &cam, const double *T, const double *Xw, const double *uv, double th, double *Jp, double *Jl, double &r, double &w) {
    double q[3]; rot(T, Xw, q);
    const double gx = q[0] + T[3], gy = q[1] + T[7], gz = q[2] + T[11];
    const double dx = uv[0] - (cam.cx + cam.fx * gx / gz), dy = uv[1] - (cam.cy + cam.fy * gy / gz);
    r = sqrt(dx * dx + dy * dy);
    const double gz2 = 1.0 / dmax(th, gz * gz);
    const double fxdx = cam.fx * dx, fydy = cam.fy * dy;
    const double den = dmax(th, r);
    const double j0 = +gz2 * fxdx * gz, j1 = +gz2 * fydy * gz, j2 = -gz2 * (fxdx * gx + fydy * gy);
    Jp[0] = j0 / den; Jp[1] = j1 / den; Jp[2] = j2 / den;
    Jp[3] = -gz2 * (fxdx * gx * gy + fydy * gy * gy + fydy * gz * gz) / den;
    Jp[4] = +gz2 * (fxdx * gx * gx + fxdx * gz * gz + fydy * gx * gy) / den;
    Jp[5] = +gz2 * (fydy * gx * gz - fxdx * gy * gz) / den;
    Jl[0] = (j0 * T[0] + j1 * T[4] + j2 * T[8]) / den;
    Jl[1] = (j0 * T[1] + j1 * T[5] + j2 * T[9]) / den;
    Jl[2] = (j0 * T[2] + j1 * T[6] + j2 * T[10]) / den;
    w = 1.0 / (1.0 + r * r);                       // robustWeightCauchy, src2/auxiliar.cpp:556-560
}
// src/mapHandler.cpp:2450-2524 / :2694-2767 ; q5_fixed = upstream form (line coefficients instead of (e0,e1), Q5)
PLBA_HD void h_endline(const Cam &cam, const double *T, const double *Pw, const double *Qw, const double *lo, double th, bool q5_fixed,
                       double *Jp, double *Jl, double &r, double &w) {
    double qp[3], qq[3]; rot(T, Pw, qp); rot(T, Qw, qq);
    const double P[3] = {qp[0] + T[3], qp[1] + T[7], qp[2] + T[11]}, Q[3] = {qq[0] + T[3], qq[1] + T[7], qq[2] + T[11]};
    const double e0 = lo[0] * (cam.cx + cam.fx * P[0] / P[2]) + lo[1] * (cam.cy + cam.fy * P[1] / P[2]) + lo[2];
    const double e1 = lo[0] * (cam.cx + cam.fx * Q[0] / Q[2]) + lo[1] * (cam.cy + cam.fy * Q[1] / Q[2]) + lo[2];
    r = sqrt(e0 * e0 + e1 * e1);
    const double lx = q5_fixed ? lo[0] : e0, ly = q5_fixed ? lo[1] : e1;
    const double fxlx = cam.fx * lx, fyly = cam.fy * ly;
    const double den = dmax(th, r);
    double JP[6], JQ[6];
    for (int s = 0; s < 2; s++) {
        const double *G = s ? Q : P; double *J = s ? JQ : JP;
        const double gx = G[0], gy = G[1], gz = G[2];
        const double gz2 = 1.0 / dmax(th, gz * gz);
        J[0] = +gz2 * fxlx * gz; J[1] = +gz2 * fyly * gz; J[2] = -gz2 * (fxlx * gx + fyly * gy);
        J[3] = -gz2 * (fxlx * gx * gy + fyly * gy * gy + fyly * gz * gz);
        J[4] = +gz2 * (fxlx * gx * gx + fxlx * gz * gz + fyly * gx * gy);
        J[5] = +gz2 * (fyly * gx * gz - fxlx * gy * gz);
        const double es = s ? e1 : e0;
        for (int c = 0; c < 3; c++) Jl[3 * s + c] = (J[0] * T[c] + J[1] * T[4 + c] + J[2] * T[8 + c]) * es / den;
    }
    for (int c = 0; c < 6; c++) Jp[c] = (JP[c] * e0 + JQ[c] * e1) / den;
    w = 1.0 / (1.0 + r * r);
}
// src/mapHandler.cpp:1741-1812 / :2001-2075 ; fixed repairs Q6 (fenmu scaling), Q7 (sign in L.u2 term), Q8 (row sign)
PLBA_HD void h_plkline(const Cam &cam, const double *T, const LinePre &L, const double *ab, double th, bool fixed,
                       double *Jp, double *Jl, double &r, double &w) {
    double l[3], Rn[3], Rd[3]; g_line_project(cam, T, L.n, L.d, l, Rn, Rd);
    const double fen = sqrt(l[0] * l[0] + l[1] * l[1]);
    double e[2];
    e[0] = (ab[0] * l[0] + ab[1] * l[1] + l[2]) / fen;
    e[1] = (ab[2] * l[0] + ab[3] * l[1] + l[2]) / fen;
    r = sqrt(e[0] * e[0] + e[1] * e[1]);
    const double den = dmax(th, r);
    const double t[3] = {T[3], T[7], T[11]};
    const double q7 = fixed ? 1.0 : -1.0;
    double Ap[6] = {0, 0, 0, 0, 0, 0}, Bl[4] = {0, 0, 0, 0};
    for (int k = 0; k < 2; k++) {
        double j0, j1, j2;
        if (!fixed) { j0 = ab[2 * k] * fen - l[0] * e[k] * fen * fen; j1 = ab[2 * k + 1] * fen - l[1] * e[k] * fen * fen; j2 = fen; }
        else { j0 = -l[0] * e[k] / (fen * fen) + ab[2 * k] / fen; j1 = -l[1] * e[k] / (fen * fen) + ab[2 * k + 1] / fen; j2 = 1.0 / fen; }
        const double g[3] = {j0 * cam.fy - j2 * cam.fy * cam.cx, j1 * cam.fx - j2 * cam.fx * cam.cy, j2 * cam.fx * cam.fy};
        double c0[3], gt[3], c1[3], c2[3];
        cross3(Rd, g, c0); cross3(Rn, g, c1); cross3(g, t, gt); cross3(gt, Rd, c2);
        const double a[6] = {c0[0], c0[1], c0[2], c1[0] - c2[0], c1[1] - c2[1], c1[2] - c2[2]};
        double gr[3], gd[3];
        rotT(T, g, gr); rotT(T, gt, gd);
        const double b[4] = {L.w2 * dot3(gd, L.u3), -L.w1 * dot3(gr, L.u3), q7 * L.w1 * dot3(gr, L.u2) - L.w2 * dot3(gd, L.u1),
                             -L.w2 * dot3(gr, L.u1) + L.w1 * dot3(gd, L.u2)};
        for (int c = 0; c < 6; c++) Ap[c] += a[c] * e[k];
        for (int c = 0; c < 4; c++) Bl[c] += b[c] * e[k];
    }
    const double sg = fixed ? -1.0 : 1.0;
    for (int c = 0; c < 6; c++) Jp[c] = sg * Ap[c] / den;
    for (int c = 0; c < 4; c++) Jl[c] = sg * Bl[c] / den;
    w = 1.0 / (1.0 + r * r);
}

// ---------------------------------------------------------------------------------------------------
// Symmetric D x D inverse (D <= 6) by Cholesky; M is the full row-major matrix, overwritten by M^-1.
// Returns false if a pivot is not positive (the caller then falls back to a pivoted Gauss-Jordan).
// ---------------------------------------------------------------------------------------------------
template <int D>
PLBA_HD bool spd_inverse(double *M) {
    // Cholesky M = L L^T with ONE reciprocal square root per pivot (no sqrt followed by divisions: this sits on the latency
    // path of the per-landmark phase), then L^-1 and M^-1 = L^-T L^-1
    double Lm[D * D], ri[D];
    for (int i = 0; i < D * D; i++) Lm[i] = 0.0;
    for (int j = 0; j < D; j++) {
        double s = M[j * D + j];
        for (int k = 0; k < j; k++) s -= Lm[j * D + k] * Lm[j * D + k];
        if (!(s > 0.0)) return false;
        ri[j] = plba_rsqrt_safe(s);
        Lm[j * D + j] = s * ri[j];
        for (int i = j + 1; i < D; i++) {
            double v = M[i * D + j];
            for (int k = 0; k < j; k++) v -= Lm[i * D + k] * Lm[j * D + k];
            Lm[i * D + j] = v * ri[j];
        }
    }
    // Linv (lower)
    double Li[D * D];
    for (int i = 0; i < D * D; i++) Li[i] = 0.0;
    for (int j = 0; j < D; j++) {
        Li[j * D + j] = ri[j];
        for (int i = j + 1; i < D; i++) {
            double s = 0.0;
            for (int k = j; k < i; k++) s -= Lm[i * D + k] * Li[k * D + j];
            Li[i * D + j] = s * ri[i];
        }
    }
    for (int r = 0; r < D; r++) for (int c = 0; c <= r; c++) {
        double s = 0.0;
        for (int k = r; k < D; k++) s += Li[k * D + r] * Li[k * D + c];
        M[r * D + c] = s; M[c * D + r] = s;
    }
    return true;
}
template <int D>
PLBA_COLD void gj_inverse(double *M) {   // pivoted Gauss-Jordan, any non-singular matrix
    double a[D][2 * D];
    for (int r = 0; r < D; r++) for (int c = 0; c < D; c++) { a[r][c] = M[r * D + c]; a[r][D + c] = (r == c) ? 1.0 : 0.0; }
    for (int k = 0; k < D; k++) {
        int p = k; double best = fabs(a[k][k]);
        for (int r = k + 1; r < D; r++) if (fabs(a[r][k]) > best) { best = fabs(a[r][k]); p = r; }
        if (p != k) for (int c = 0; c < 2 * D; c++) { const double t = a[k][c]; a[k][c] = a[p][c]; a[p][c] = t; }
        const double inv = 1.0 / a[k][k];
        for (int c = 0; c < 2 * D; c++) a[k][c] *= inv;
        for (int r = 0; r < D; r++) if (r != k) { const double f = a[r][k]; if (f != 0.0) for (int c = 0; c < 2 * D; c++) a[r][c] -= f * a[k][c]; }
    }
    for (int r = 0; r < D; r++) for (int c = 0; c < D; c++) M[r * D + c] = a[r][D + c];
}

}  // namespace plba
