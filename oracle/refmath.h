// TEST INFRASTRUCTURE ONLY (see oracle/README.md).
// CPU restatement of the small pure-math helpers under the reference's LBA.
// Every function cites the reference file:line it follows.  PARITY UNPINNED: the reference ships no test,
// golden vector or fixture for any of these; pins are self-made (finite differences,
// round trips) in tests/.
#pragma once
#include "smallmat.h"

namespace oracle {

// src2/auxiliar.cpp:29-44
inline M3 skew(const V3 &v) {
    M3 s;
    s(0, 1) = -v[2]; s(0, 2) = v[1]; s(1, 2) = -v[0];
    s(1, 0) = v[2];  s(2, 0) = -v[1]; s(2, 1) = v[0];
    return s;
}
// src2/auxiliar.cpp:58-62
inline V3 skewcoords(const M3 &M) { V3 s; s[0] = M(2, 1); s[1] = M(0, 2); s[2] = M(1, 0); return s; }

// src2/auxiliar.cpp:113-122
inline M4 inverse_se3(const M4 &T) {
    M4 Tinv = M4::Identity();
    M3 R = T.block<3, 3>(0, 0);
    V3 t = T.block<3, 1>(0, 3);
    Tinv.setBlock<3, 3>(0, 0, R.T());
    Tinv.setBlock<3, 1>(0, 3, -(R.T() * t));
    return Tinv;
}

// src2/auxiliar.cpp:124-141   (x = [translation; rotation])
inline M4 expmap_se3(const V6 &x) {
    M3 R, V, s, I = M3::Identity();
    V3 t, w;
    M4 T = M4::Identity();
    for (int i = 0; i < 3; i++) { w[i] = x[3 + i]; t[i] = x[i]; }
    double theta = w.norm();
    if (theta < 0.000001)
        R = I;
    else {
        s = skew(w) / theta;
        R = I + s * std::sin(theta) + s * s * (1.0 - std::cos(theta));
        V = I + s * (1.0 - std::cos(theta)) / theta + s * s * (theta - std::sin(theta)) / theta;
        t = V * t;
    }
    T.setBlock<3, 3>(0, 0, R);
    T.setBlock<3, 1>(0, 3, t);
    return T;
}

// src2/auxiliar.cpp:143-173
inline V6 logmap_se3(const M4 &T) {
    M3 R, Id3 = M3::Identity();
    V3 Vt, t, w;
    M3 V = M3::Identity(), w_hat;
    V6 x;
    Vt[0] = T(0, 3); Vt[1] = T(1, 3); Vt[2] = T(2, 3);
    R = T.block<3, 3>(0, 0);
    double cosine = ((R(0, 0) + R(1, 1) + R(2, 2)) - 1.0) / 2.0;
    if (cosine > 1.0) cosine = 1.0;
    else if (cosine < -1.0) cosine = -1.0;
    double sine = std::sqrt(1.0 - cosine * cosine);
    if (sine > 1.0) sine = 1.0;
    else if (sine < -1.0) sine = -1.0;
    double theta = std::acos(cosine);
    if (theta > 0.000001) {
        w_hat = (R - R.T()) * theta / (2.0 * sine);
        w = skewcoords(w_hat);
        M3 s = skew(w) / theta;
        V = Id3 + s * (1.0 - cosine) / theta + s * s * (theta - sine) / theta;
    }
    t = inverse(V) * Vt;
    for (int i = 0; i < 3; i++) { x[i] = t[i]; x[3 + i] = w[i]; }
    return x;
}

// src2/auxiliar.cpp:556-560
inline double robustWeightCauchy(double norm_res) { return 1.0 / (1.0 + norm_res * norm_res); }

// src2/pinholeStereoCamera.cpp:235-241
inline V2 projection(const double cam[4], const V3 &P) {
    V2 uv; uv[0] = cam[2] + cam[0] * P[0] / P[2]; uv[1] = cam[3] + cam[1] * P[1] / P[2]; return uv;
}
// src2/pinholeStereoCamera.cpp:123-125 ; g2o_types/g2o_types.h:349-355
inline M3 plukerK(const double cam[4]) {
    double fx = cam[0], fy = cam[1], cx = cam[2], cy = cam[3];
    M3 K;
    K(0, 0) = fy; K(1, 1) = fx;
    K(2, 0) = -fy * cx; K(2, 1) = -fx * cy; K(2, 2) = fx * fy;
    return K;
}

// include/mapHandler.h:224-230 ; g2o_types/g2o_types.h:18-24
inline M3 vechat(const V3 &v) { return skew(v); }

// include/mapHandler.h:242-250 ; g2o_types/g2o_types.h:357-365
inline M6 getTransformMatrixForPluker(const M4 &T) {
    M6 temp;
    M3 R = T.block<3, 3>(0, 0);
    temp.setBlock<3, 3>(0, 0, R);
    temp.setBlock<3, 3>(0, 3, vechat(T.block<3, 1>(0, 3)) * R);
    temp.setBlock<3, 3>(3, 3, R);
    return temp;
}

// src/mapFeatures.cpp:226-240 ; g2o_types/g2o_types.h:482-495
inline M3 getOrhtRFromPluker(const V6 &pl) {
    M3 R;
    V3 n = pl.block<3, 1>(0, 0), d = pl.block<3, 1>(3, 0);
    V3 n0 = n, d0 = d;
    n = n / n.norm();
    d = d / d.norm();
    R.setCol(0, n);
    R.setCol(1, d);
    V3 c = cross(n0, d0);
    R.setCol(2, c / c.norm());
    return R;
}
// src/mapFeatures.cpp:242-249 ; g2o_types/g2o_types.h:472-480
inline M2 getOrthWFromPluker(const V6 &pl) {
    M2 temp;
    double nnorm = pl.block<3, 1>(0, 0).norm();
    double dnorm = pl.block<3, 1>(3, 0).norm();
    double fenmu = std::sqrt(nnorm * nnorm + dnorm * dnorm);
    temp(0, 0) = nnorm / fenmu; temp(0, 1) = -dnorm / fenmu; temp(1, 0) = dnorm / fenmu; temp(1, 1) = nnorm / fenmu;
    return temp;
}
// src/mapFeatures.cpp:186-201
inline V4 changePlukerToOrth(const V6 &pl) {
    V4 o;
    M3 R = getOrhtRFromPluker(pl);
    o[0] = std::atan2(R(2, 1), R(2, 2));
    o[1] = std::asin(-R(2, 0));
    o[2] = std::atan2(R(1, 0), R(0, 0));
    M2 W = getOrthWFromPluker(pl);
    o[3] = std::asin(W(1, 0));
    return o;
}
inline M3 orthR(double t0, double t1, double t2) {
    double s1 = std::sin(t0), c1 = std::cos(t0), s2 = std::sin(t1), c2 = std::cos(t1), s3 = std::sin(t2), c3 = std::cos(t2);
    M3 R;
    R(0, 0) = c2 * c3; R(0, 1) = s1 * s2 * c3 - c1 * s3; R(0, 2) = c1 * s2 * c3 + s1 * s3;
    R(1, 0) = c2 * s3; R(1, 1) = s1 * s2 * s3 + c1 * c3; R(1, 2) = c1 * s2 * s3 - s1 * c3;
    R(2, 0) = -s2;     R(2, 1) = s1 * c2;                R(2, 2) = c1 * c2;
    return R;
}
// src/mapFeatures.cpp:203-224 ; g2o_types/g2o_types.h:367-387
inline V6 changeOrthToPluker(const V4 &o) {
    M3 R = orthR(o[0], o[1], o[2]);
    double w1 = std::cos(o[3]), w2 = std::sin(o[3]);
    V6 pl;
    for (int i = 0; i < 3; i++) { pl[i] = w1 * R(i, 0); pl[3 + i] = w2 * R(i, 1); }
    return pl;
}
// g2o_types/g2o_types.h:455-470 (sign_q7=+1, FD-verified) ; src/mapFeatures.cpp:251-266 (sign_q7=-1, Q7)
inline Mat<6, 4> jacobianFromPlukerToOrth(const M3 &u, const M2 &w, double sign_q7) {
    double w1 = w(0, 0), w2 = w(1, 0);
    V3 u1 = u.col(0), u2 = u.col(1), u3 = u.col(2);
    Mat<6, 4> t;
    t.setBlock<3, 1>(0, 1, u3 * (-w1));
    t.setBlock<3, 1>(0, 2, u2 * (sign_q7 * w1));
    t.setBlock<3, 1>(0, 3, u1 * (-w2));
    t.setBlock<3, 1>(3, 0, u3 * w2);
    t.setBlock<3, 1>(3, 2, u1 * (-w2));
    t.setBlock<3, 1>(3, 3, u2 * w1);
    return t;
}
// include/mapHandler.h:252-335 ; g2o_types/g2o_types.h:72-130
inline V4 updateOrthCoord(const V4 &D, const V4 &dD) {
    M3 R = orthR(D[0], D[1], D[2]);
    double w1 = std::cos(D[3]), w2 = std::sin(D[3]);
    M3 Rz, Ry, Rx;
    Rz(0, 0) = std::cos(dD[2]); Rz(0, 1) = -std::sin(dD[2]); Rz(1, 0) = std::sin(dD[2]); Rz(1, 1) = std::cos(dD[2]); Rz(2, 2) = 1;
    Ry(0, 0) = std::cos(dD[1]); Ry(0, 2) = std::sin(dD[1]); Ry(1, 1) = 1; Ry(2, 0) = -std::sin(dD[1]); Ry(2, 2) = std::cos(dD[1]);
    Rx(0, 0) = 1; Rx(1, 1) = std::cos(dD[0]); Rx(1, 2) = -std::sin(dD[0]); Rx(2, 1) = std::sin(dD[0]); Rx(2, 2) = std::cos(dD[0]);
    R = R * Rx * Ry * Rz;
    M2 W, dW;
    W(0, 0) = w1; W(0, 1) = -w2; W(1, 0) = w2; W(1, 1) = w1;
    dW(0, 0) = std::cos(dD[3]); dW(0, 1) = -std::sin(dD[3]); dW(1, 0) = std::sin(dD[3]); dW(1, 1) = std::cos(dD[3]);
    W = W * dW;
    V4 p;
    p[0] = std::atan2(R(2, 1), R(2, 2));
    p[1] = std::asin(-R(2, 0));
    p[2] = std::atan2(R(1, 0), R(0, 0));
    p[3] = std::asin(W(1, 0));
    return p;
}

// g2o_types/g2o_types.h:172-203  VertexLMPose::oplusImpl  (update = [dt; omega]; R <- dR*R, t <- t + dt)
inline M4 poseOplusG2O(const M4 &est, const V6 &delta) {
    V3 omega; for (int i = 0; i < 3; i++) omega[i] = delta[3 + i];
    double theta = omega.norm();
    double half_theta = 0.5 * theta;
    double imag_factor;
    double real_factor = std::cos(half_theta);
    if (theta < 1e-10) {
        double theta_sq = theta * theta;
        double theta_po4 = theta_sq * theta_sq;
        imag_factor = 0.5 - 0.0208333 * theta_sq + 0.000260417 * theta_po4;
    } else {
        imag_factor = std::sin(half_theta) / theta;
    }
    // Eigen::Quaterniond(w,x,y,z).toRotationMatrix()
    double w = real_factor, x = imag_factor * omega[0], y = imag_factor * omega[1], z = imag_factor * omega[2];
    double tx = 2 * x, ty = 2 * y, tz = 2 * z;
    double twx = tx * w, twy = ty * w, twz = tz * w, txx = tx * x, txy = ty * x, txz = tz * x, tyy = ty * y, tyz = tz * y, tzz = tz * z;
    M3 dR;
    dR(0, 0) = 1 - (tyy + tzz); dR(0, 1) = txy - twz; dR(0, 2) = txz + twy;
    dR(1, 0) = txy + twz; dR(1, 1) = 1 - (txx + tzz); dR(1, 2) = tyz - twx;
    dR(2, 0) = txz - twy; dR(2, 1) = tyz + twx; dR(2, 2) = 1 - (txx + tyy);
    M4 out = est;
    out.setBlock<3, 3>(0, 0, dR * est.block<3, 3>(0, 0));
    for (int i = 0; i < 3; i++) out(i, 3) = est(i, 3) + delta[i];
    return out;
}

// g2o RobustKernelHuber::robustify (external g2o, restated in SURVEY.md §8c(3))
inline void huberRobustify(double delta, double e, double rho[3]) {
    double dsqr = delta * delta;
    if (e <= dsqr) { rho[0] = e; rho[1] = 1.; rho[2] = 0.; }
    else {
        double sqrte = std::sqrt(e);
        rho[0] = 2 * sqrte * delta - dsqr;
        rho[1] = delta / sqrte;
        rho[2] = -0.5 * rho[1] / e;
    }
}

// ---- g2o edges (g2o_types/g2o_types.h:206-300, 302-453) ----
struct PointEdgeLin { V2 e; Mat<2, 3> Jxi; Mat<2, 6> Jxj; double zc; };
inline V3 pointPc(const M4 &Tcw, const V3 &Pw) { return Tcw.block<3, 3>(0, 0) * Pw + Tcw.block<3, 1>(0, 3); }
inline V2 pointEdgeError(const double cam[4], const M4 &Tcw, const V3 &Pw, const V2 &obs) {
    V3 Pc = pointPc(Tcw, Pw);
    V2 proj; proj[0] = Pc[0] / Pc[2]; proj[1] = Pc[1] / Pc[2];
    V2 res; res[0] = proj[0] * cam[0] + cam[2]; res[1] = proj[1] * cam[1] + cam[3];
    return obs - res;
}
inline void pointEdgeLinearize(const double cam[4], const M4 &Tcw, const V3 &Pw, PointEdgeLin &o) {
    double fx = cam[0], fy = cam[1];
    M3 Rot = Tcw.block<3, 3>(0, 0);
    V3 Pc = pointPc(Tcw, Pw);
    double x = Pc[0], y = Pc[1], z = Pc[2];
    double invz = 1.0 / z, invz2 = invz * invz;
    Mat<2, 3> jpc;
    jpc(0, 0) = fx / z; jpc(0, 2) = -fx * x * invz2;
    jpc(1, 1) = fy / z; jpc(1, 2) = -fy * y * invz2;
    o.Jxi = -(jpc * Rot);
    o.Jxj.setBlock<2, 3>(0, 0, -jpc);
    o.Jxj.setBlock<2, 3>(0, 3, -(jpc * (-vechat(Rot * Pw))));
    o.zc = z;
}

struct LineEdgeLin { V2 e; Mat<2, 4> Jxi; Mat<2, 6> Jxj; };
inline V2 lineEdgeError(const double cam[4], const M4 &Tcw, const V4 &orth, const V4 &obs) {
    V6 Lw = changeOrthToPluker(orth);
    V6 Lc = getTransformMatrixForPluker(Tcw) * Lw;
    V3 l = plukerK(cam) * Lc.block<3, 1>(0, 0);
    double fenmu = std::sqrt(l[0] * l[0] + l[1] * l[1]);
    V2 e;
    e[0] = (l[0] * obs[0] + l[1] * obs[1] + l[2]) / fenmu;
    e[1] = (l[0] * obs[2] + l[1] * obs[3] + l[2]) / fenmu;
    return e;
}
// quirk_q12_faithful: pose Jacobian built from the orth 4-vector's tail(3)/head(3) (g2o_types.h:396,429-430)
inline void lineEdgeLinearize(const double cam[4], const M4 &Tcw, const V4 &orth, const V4 &obs, bool quirk_q12_faithful, LineEdgeLin &o) {
    M3 Rcw = Tcw.block<3, 3>(0, 0);
    V3 Pcw = Tcw.block<3, 1>(0, 3);
    V6 plukerLw = changeOrthToPluker(orth);
    V6 plukerLc = getTransformMatrixForPluker(Tcw) * plukerLw;
    V3 l = plukerK(cam) * plukerLc.block<3, 1>(0, 0);
    double lx = l[0], ly = l[1], lz = l[2];
    double fenmu = std::sqrt(lx * lx + ly * ly);
    V2 error;
    error[0] = (lx * obs[0] + ly * obs[1] + lz) / fenmu;
    error[1] = (lx * obs[2] + ly * obs[3] + lz) / fenmu;
    Mat<1, 3> j0, j1;
    j0(0, 0) = -lx * error[0] / (fenmu * fenmu) + obs[0] / fenmu;
    j0(0, 1) = -ly * error[0] / (fenmu * fenmu) + obs[1] / fenmu;
    j0(0, 2) = 1.0 / fenmu;
    j1(0, 0) = -lx * error[1] / (fenmu * fenmu) + obs[2] / fenmu;
    j1(0, 1) = -ly * error[1] / (fenmu * fenmu) + obs[3] / fenmu;
    j1(0, 2) = 1.0 / fenmu;
    Mat<3, 6> jac_lcPixel_lc;
    jac_lcPixel_lc.setBlock<3, 3>(0, 0, plukerK(cam));
    V3 tail3, head3;
    if (quirk_q12_faithful) { for (int i = 0; i < 3; i++) { tail3[i] = orth[1 + i]; head3[i] = orth[i]; } }
    else { for (int i = 0; i < 3; i++) { tail3[i] = plukerLw[3 + i]; head3[i] = plukerLw[i]; } }
    M6 jac_lc_rt;
    jac_lc_rt.setBlock<3, 3>(0, 0, -vechat(Rcw * tail3));
    jac_lc_rt.setBlock<3, 3>(0, 3, -vechat(Rcw * head3) - vechat(Pcw) * vechat(Rcw * tail3));
    Mat<1, 6> r0 = j0 * jac_lcPixel_lc * jac_lc_rt, r1 = j1 * jac_lcPixel_lc * jac_lc_rt;
    for (int c = 0; c < 6; c++) { o.Jxj(0, c) = r0(0, c); o.Jxj(1, c) = r1(0, c); }
    M6 jac_lc_lw = getTransformMatrixForPluker(Tcw);
    M3 U = getOrhtRFromPluker(plukerLw);
    M2 W = getOrthWFromPluker(plukerLw);
    Mat<6, 4> jac_lw_orth = jacobianFromPlukerToOrth(U, W, +1.0);
    Mat<1, 4> q0 = j0 * jac_lcPixel_lc * jac_lc_lw * jac_lw_orth, q1 = j1 * jac_lcPixel_lc * jac_lc_lw * jac_lw_orth;
    for (int c = 0; c < 4; c++) { o.Jxi(0, c) = q0(0, c); o.Jxi(1, c) = q1(0, c); }
    o.e = error;
}

}  // namespace oracle
