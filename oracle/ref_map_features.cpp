// TEST INFRASTRUCTURE ONLY — C wrapper around the REFERENCE's MapLine orthonormal-representation helpers.
// Compiled together with /root/reference/src/mapFeatures.cpp (UNMODIFIED, from where it lies) against the stand-in headers of
// oracle/ref_shim/ into oracle/_ref/libref_g2o_types.so (oracle/Makefile, target `ref`).  Pins rows a6, a7, a8, a9 (the
// MapLine:: copies, incl. the wrong-sign Jacobian of quirk Q7) of SURVEY.md §8a.
#include "mapFeatures.h"

extern "C" {
// MapLine::changePlukerToOrth (src/mapFeatures.cpp:186-201)
void ref_ml_pluker_to_orth(const double *plk6, double *orth4) {
    Vector6d p; for (int i = 0; i < 6; i++) p(i) = plk6[i];
    Vector4d o = PLSLAM::MapLine::changePlukerToOrth(p);
    for (int i = 0; i < 4; i++) orth4[i] = o(i);
}
// MapLine::changeOrthToPluker (:203-224)
void ref_ml_orth_to_pluker(const double *orth4, double *plk6) {
    Vector4d o; for (int i = 0; i < 4; i++) o(i) = orth4[i];
    Vector6d p = PLSLAM::MapLine::changeOrthToPluker(o);
    for (int i = 0; i < 6; i++) plk6[i] = p(i);
}
// MapLine::getOrhtRFromPluker / getOrthWFromPluker / jacobianFromPlukerToOrth (:226-266; the Jacobian is the Q7 variant)
void ref_ml_UW_jac(const double *plk6, double *U9, double *W4, double *J24) {
    Vector6d p; for (int i = 0; i < 6; i++) p(i) = plk6[i];
    Matrix3d U = PLSLAM::MapLine::getOrhtRFromPluker(p); Matrix2d W = PLSLAM::MapLine::getOrthWFromPluker(p);
    Eigen::Matrix<double, 6, 4> J = PLSLAM::MapLine::jacobianFromPlukerToOrth(U, W);
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) U9[3 * i + j] = U(i, j);
    for (int i = 0; i < 2; i++) for (int j = 0; j < 2; j++) W4[2 * i + j] = W(i, j);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 4; j++) J24[4 * i + j] = J(i, j);
}
}
