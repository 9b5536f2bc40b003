// TEST INFRASTRUCTURE ONLY — C wrapper around the REFERENCE's own vertex / edge classes.
//
// Compiles /root/reference/g2o_types/g2o_types.h UNMODIFIED, from where it lies (never copied into this repository), against
// the stand-in Eigen / g2o headers of oracle/ref_shim/ (written from scratch, see ref_shim/standin.h), into
// oracle/_ref/libref_g2o_types.so (recipe: oracle/Makefile, target `ref`).  The entry points below mirror the oracle's
// plba_oracle_point_edge / _line_edge / _pose_oplus / _update_orth / _orth_to_pluker so that tests/test_ref_pin.py can
// compare the restatement (oracle/refmath.h) with the reference's arithmetic value for value.  /root/reference does not exist on
// the GPU box: the outputs of this library on seeded inputs are committed as tests/golden/ref_g2o_types.npz
// (tests/golden/make_ref_golden.py) and the oracle is checked against those wherever the library itself is absent.
#include "g2o_types/g2o_types.h"

namespace {
Matrix4d mat4(const double *T16) { Matrix4d T; for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) T(i, j) = T16[4 * i + j]; return T; }
}

extern "C" {

// VertexLMPose::oplusImpl (g2o_types/g2o_types.h:172-203): T16 row-major 4x4, d6 = [dt; omega]
void ref_pose_oplus(const double *T16, const double *d6, double *out16) {
    VertexLMPose v; v.setEstimate(mat4(T16)); v.oplus(d6);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) out16[4 * i + j] = v.estimate()(i, j);
}
// VertexLMLineOrth::oplusImpl -> updateOrthCoord (:64-131)
void ref_line_oplus(const double *D4, const double *d4, double *out4) {
    VertexLMLineOrth v; Vector4d D; for (int i = 0; i < 4; i++) D(i) = D4[i];
    v.setEstimate(D); v.oplus(d4);
    for (int i = 0; i < 4; i++) out4[i] = v.estimate()(i);
}
// VertexLMPointXYZ::oplusImpl (:40-44)
void ref_point_oplus(const double *P3, const double *d3, double *out3) {
    VertexLMPointXYZ v; Vector3d P; for (int i = 0; i < 3; i++) P(i) = P3[i];
    v.setEstimate(P); v.oplus(d3);
    for (int i = 0; i < 3; i++) out3[i] = v.estimate()(i);
}
// EdgePosePoint::computeError / linearizeOplus / isDepthPositive (:224-296): e[2], Jxi[2x3], Jxj[2x6] row-major
int ref_point_edge(const double *cam, const double *Tcw16, const double *Pw3, const double *obs2, double *e2, double *Jxi6, double *Jxj12) {
    VertexLMPointXYZ vp; Vector3d P; for (int i = 0; i < 3; i++) P(i) = Pw3[i]; vp.setEstimate(P);
    VertexLMPose vT; vT.setEstimate(mat4(Tcw16));
    EdgePosePoint e; e.setVertex(0, &vp); e.setVertex(1, &vT); e.SetParams(cam[0], cam[1], cam[2], cam[3]);
    Vector2d ob; ob(0) = obs2[0]; ob(1) = obs2[1]; e.setMeasurement(ob);
    e.computeError(); e.linearizeOplus();
    for (int i = 0; i < 2; i++) e2[i] = e.error()(i);
    for (int i = 0; i < 2; i++) for (int j = 0; j < 3; j++) Jxi6[3 * i + j] = e.jacobianOplusXi()(i, j);
    for (int i = 0; i < 2; i++) for (int j = 0; j < 6; j++) Jxj12[6 * i + j] = e.jacobianOplusXj()(i, j);
    return e.isDepthPositive() ? 1 : 0;
}
// EdgePoseLine::computeError / linearizeOplus (:318-452): e[4], Jxi[4x4], Jxj[4x6] row-major (rows 2-3 are zero in the reference)
void ref_line_edge(const double *cam, const double *Tcw16, const double *orth4, const double *obs4, double *e4, double *Jxi16, double *Jxj24, double *chi2) {
    VertexLMLineOrth vl; Vector4d o, ob; for (int i = 0; i < 4; i++) { o(i) = orth4[i]; ob(i) = obs4[i]; } vl.setEstimate(o);
    VertexLMPose vT; vT.setEstimate(mat4(Tcw16));
    EdgePoseLine e; e.setVertex(0, &vl); e.setVertex(1, &vT); e.SetParams(cam[0], cam[1], cam[2], cam[3]);
    e.setMeasurement(ob);
    e.computeError(); e.linearizeOplus();
    for (int i = 0; i < 4; i++) e4[i] = e.error()(i);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) Jxi16[4 * i + j] = e.jacobianOplusXi()(i, j);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 6; j++) Jxj24[6 * i + j] = e.jacobianOplusXj()(i, j);
    if (chi2) *chi2 = e.chi2();
}
// EdgePoseLine::changeOrthToPluker (:367-387), getOrhtRFromPluker / getOrthWFromPluker (:472-495), jacobianFromPlukerToOrth (:455-470)
void ref_orth_to_pluker(const double *orth4, double *plk6) {
    EdgePoseLine e; Vector4d o; for (int i = 0; i < 4; i++) o(i) = orth4[i];
    Vector6d p = e.changeOrthToPluker(o);
    for (int i = 0; i < 6; i++) plk6[i] = p(i);
}
void ref_orth_UW_jac(const double *plk6, double *U9, double *W4, double *J24) {
    EdgePoseLine e; Vector6d p; for (int i = 0; i < 6; i++) p(i) = plk6[i];
    Matrix3d U = e.getOrhtRFromPluker(p); Matrix2d W = e.getOrthWFromPluker(p);
    Eigen::Matrix<double, 6, 4> J = e.jacobianFromPlukerToOrth(U, W);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) U9[3 * i + j] = U(i, j);
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) W4[2 * i + j] = W(i, j);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 4; j++) J24[4 * i + j] = J(i, j);
}
// EdgePoseLine::getTransformMatrix (:357-365) applied to a Plücker vector
void ref_transform_pluker(const double *Tcw16, const double *plk6, double *out6) {
    EdgePoseLine e; Vector6d p; for (int i = 0; i < 6; i++) p(i) = plk6[i];
    Vector6d q = e.getTransformMatrix(mat4(Tcw16)) * p;
    for (int i = 0; i < 6; i++) out6[i] = q(i);
}

}  // extern "C"
