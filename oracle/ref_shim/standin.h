// TEST INFRASTRUCTURE ONLY — minimal stand-ins for the parts of Eigen and g2o that the reference's g2o_types/g2o_types.h uses.
//
// Purpose: the reference (kongan/PL-SLAM-plucker) cannot be built in this image (Eigen, g2o, OpenCV, Boost, MRPT absent).  Its
// vertex / edge arithmetic for profile G, however, lives in ONE header, g2o_types/g2o_types.h, which needs nothing but fixed-size
// dense algebra and the g2o base-class members it touches.  These stand-ins are written from scratch (no Eigen / g2o code) so that
// oracle/ref_g2o_types.cpp can compile the reference header UNMODIFIED, from where it lies under /root/reference, into
// oracle/_ref/libref_g2o_types.so.  That library pins the oracle's restatement of rows a6-a11, a17-a19 of SURVEY.md §8a
// (tests/test_ref_pin.py).  What it cannot pin: the g2o LM / Schur shell and the hand-LM functions of
// src/mapHandler.cpp (OpenCV-entangled).
//
// Semantics kept: column vectors, (row, col) indexing, eager left-to-right evaluation of products (Eigen evaluates the same order
// for these fixed sizes; any difference is rounding-level), Quaternion::toRotationMatrix without normalisation.
// Limitation: head(n) / tail(n) with a run-time n are supported for n == 3 only (all the reference header uses).
#pragma once
#include <cmath>
#include <cassert>
#include <iostream>
#include <vector>
#include <algorithm>

namespace Eigen {

template <typename T, int R, int C> class Matrix;
template <int R, int C> struct BlockRef;

template <int R, int C> struct CommaInit {
    Matrix<double, R, C> *m; int k;
    CommaInit &operator,(double v) { assert(k < R * C); (*m)(k / C, k % C) = v; k++; return *this; }
};

template <typename T, int R, int C> class Matrix {
public:
    double a[R * C];    // row-major storage
    Matrix() { for (int i = 0; i < R * C; i++) a[i] = 0.0; }
    Matrix(const Matrix &) = default;
    Matrix &operator=(const Matrix &) = default;
    double &operator()(int i, int j) { return a[i * C + j]; }
    double operator()(int i, int j) const { return a[i * C + j]; }
    double &operator()(int i) { static_assert(R == 1 || C == 1, "vector"); return a[i]; }
    double operator()(int i) const { static_assert(R == 1 || C == 1, "vector"); return a[i]; }
    double &operator[](int i) { static_assert(R == 1 || C == 1, "vector"); return a[i]; }
    double operator[](int i) const { static_assert(R == 1 || C == 1, "vector"); return a[i]; }
    double x() const { return a[0]; } double y() const { return a[1]; } double z() const { return a[2]; }
    CommaInit<R, C> operator<<(double v) { a[0] = v; return CommaInit<R, C>{this, 1}; }
    void fill(double v) { for (int i = 0; i < R * C; i++) a[i] = v; }
    void setZero() { fill(0.0); }
    void setIdentity() { setZero(); for (int i = 0; i < (R < C ? R : C); i++) (*this)(i, i) = 1.0; }
    static Matrix Zero() { return Matrix(); }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    Matrix<T, C, R> transpose() const { Matrix<T, C, R> t; for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) t(j, i) = (*this)(i, j); return t; }
    double squaredNorm() const { double s = 0; for (int i = 0; i < R * C; i++) s += a[i] * a[i]; return s; }
    double norm() const { return std::sqrt(squaredNorm()); }
    void normalize() { const double n = norm(); for (int i = 0; i < R * C; i++) a[i] /= n; }
    double dot(const Matrix &o) const { double s = 0; for (int i = 0; i < R * C; i++) s += a[i] * o.a[i]; return s; }
    Matrix cross(const Matrix &o) const {
        static_assert(R * C == 3, "cross of 3-vectors");
        Matrix r; r.a[0] = a[1] * o.a[2] - a[2] * o.a[1]; r.a[1] = a[2] * o.a[0] - a[0] * o.a[2]; r.a[2] = a[0] * o.a[1] - a[1] * o.a[0]; return r;
    }
    Matrix &operator+=(const Matrix &o) { for (int i = 0; i < R * C; i++) a[i] += o.a[i]; return *this; }
    Matrix &operator-=(const Matrix &o) { for (int i = 0; i < R * C; i++) a[i] -= o.a[i]; return *this; }
    Matrix operator-() const { Matrix r; for (int i = 0; i < R * C; i++) r.a[i] = -a[i]; return r; }
    // blocks: the non-const forms return a value (so it can be used in expressions) that writes through on assignment
    template <int BR, int BC> BlockRef<BR, BC> block(int i0, int j0);
    template <int BR, int BC> Matrix<double, BR, BC> block(int i0, int j0) const {
        Matrix<double, BR, BC> r; for (int i = 0; i < BR; i++) for (int j = 0; j < BC; j++) r(i, j) = (*this)(i0 + i, j0 + j); return r;
    }
    BlockRef<R, 1> col(int j);
    Matrix<double, R, 1> col(int j) const { Matrix<double, R, 1> r; for (int i = 0; i < R; i++) r(i) = (*this)(i, j); return r; }
    BlockRef<3, 1> head(int n);
    BlockRef<3, 1> tail(int n);
    Matrix<double, 3, 1> head(int n) const { assert(n == 3 && C == 1); (void)n; Matrix<double, 3, 1> r; for (int i = 0; i < 3; i++) r(i) = a[i]; return r; }
    Matrix<double, 3, 1> tail(int n) const { assert(n == 3 && C == 1); (void)n; Matrix<double, 3, 1> r; for (int i = 0; i < 3; i++) r(i) = a[R - 3 + i]; return r; }
};

// a copy of a sub-block that remembers where it came from: IS-A Matrix (template deduction in products works), assignment writes back
template <int R, int C> struct BlockRef : public Matrix<double, R, C> {
    double *base; int ld;                 // address of element (0,0) inside the parent, parent's row stride
    BlockRef(double *b, int ld_) : base(b), ld(ld_) { for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) (*this)(i, j) = b[i * ld_ + j]; }
    BlockRef(const BlockRef &) = default;
    void store() { for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) base[i * ld + j] = (*this)(i, j); }
    BlockRef &operator=(const Matrix<double, R, C> &m) { Matrix<double, R, C>::operator=(m); store(); return *this; }
    BlockRef &operator=(const BlockRef &m) { Matrix<double, R, C>::operator=(static_cast<const Matrix<double, R, C> &>(m)); store(); return *this; }
    BlockRef &operator+=(const Matrix<double, R, C> &m) { Matrix<double, R, C>::operator+=(m); store(); return *this; }
    BlockRef &operator-=(const Matrix<double, R, C> &m) { Matrix<double, R, C>::operator-=(m); store(); return *this; }
};
template <typename T, int R, int C> template <int BR, int BC> BlockRef<BR, BC> Matrix<T, R, C>::block(int i0, int j0) { return BlockRef<BR, BC>(&a[i0 * C + j0], C); }
template <typename T, int R, int C> BlockRef<R, 1> Matrix<T, R, C>::col(int j) { return BlockRef<R, 1>(&a[j], C); }
template <typename T, int R, int C> BlockRef<3, 1> Matrix<T, R, C>::head(int n) { assert(n == 3 && C == 1); (void)n; return BlockRef<3, 1>(&a[0], 1); }
template <typename T, int R, int C> BlockRef<3, 1> Matrix<T, R, C>::tail(int n) { assert(n == 3 && C == 1); (void)n; return BlockRef<3, 1>(&a[R - 3], 1); }

template <int R, int K, int C> Matrix<double, R, C> operator*(const Matrix<double, R, K> &A, const Matrix<double, K, C> &B) {
    Matrix<double, R, C> r;
    for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) { double s = 0; for (int k = 0; k < K; k++) s += A(i, k) * B(k, j); r(i, j) = s; }
    return r;
}
template <int R, int C> Matrix<double, R, C> operator+(const Matrix<double, R, C> &A, const Matrix<double, R, C> &B) { Matrix<double, R, C> r(A); r += B; return r; }
template <int R, int C> Matrix<double, R, C> operator-(const Matrix<double, R, C> &A, const Matrix<double, R, C> &B) { Matrix<double, R, C> r(A); r -= B; return r; }
template <int R, int C> Matrix<double, R, C> operator*(double s, const Matrix<double, R, C> &A) { Matrix<double, R, C> r; for (int i = 0; i < R * C; i++) r.a[i] = s * A.a[i]; return r; }
template <int R, int C> Matrix<double, R, C> operator*(const Matrix<double, R, C> &A, double s) { return s * A; }
template <int R, int C> Matrix<double, R, C> operator/(const Matrix<double, R, C> &A, double s) { Matrix<double, R, C> r; for (int i = 0; i < R * C; i++) r.a[i] = A.a[i] / s; return r; }
template <int R, int C> std::ostream &operator<<(std::ostream &os, const Matrix<double, R, C> &A) { for (int i = 0; i < R; i++) { for (int j = 0; j < C; j++) os << A(i, j) << ' '; os << '\n'; } return os; }

typedef Matrix<double, 2, 1> Vector2d; typedef Matrix<double, 3, 1> Vector3d; typedef Matrix<double, 4, 1> Vector4d;
// run-time sized float matrix: only (n, n) construction and element access (descriptor-distance table of src/mapFeatures.cpp)
class MatrixXf { std::vector<float> v; int c; public: MatrixXf(int r, int c_) : v((size_t)r * c_, 0.f), c(c_) {} float &operator()(int i, int j) { return v[(size_t)i * c + j]; } };
typedef Matrix<double, 2, 2> Matrix2d; typedef Matrix<double, 3, 3> Matrix3d; typedef Matrix<double, 4, 4> Matrix4d;

// Map<const V>: a copy of the mapped doubles (read-only use in the reference header)
template <typename V> class Map;
template <int R, int C> class Map<const Matrix<double, R, C>> : public Matrix<double, R, C> {
public:
    explicit Map(const double *p) { for (int i = 0; i < R; i++) for (int j = 0; j < C; j++) (*this)(i, j) = (C == 1 || R == 1) ? p[i * C + j] : p[j * R + i]; }
};

// unit-less quaternion (w, x, y, z); toRotationMatrix as published (no normalisation)
class Quaterniond {
    double w_, x_, y_, z_;
public:
    Quaterniond(double w, double x, double y, double z) : w_(w), x_(x), y_(y), z_(z) {}
    Matrix3d toRotationMatrix() const {
        Matrix3d r;
        const double tx = 2.0 * x_, ty = 2.0 * y_, tz = 2.0 * z_;
        const double twx = tx * w_, twy = ty * w_, twz = tz * w_, txx = tx * x_, txy = ty * x_, txz = tz * x_, tyy = ty * y_, tyz = tz * y_, tzz = tz * z_;
        r(0, 0) = 1.0 - (tyy + tzz); r(0, 1) = txy - twz; r(0, 2) = txz + twy;
        r(1, 0) = txy + twz; r(1, 1) = 1.0 - (txx + tzz); r(1, 2) = tyz - twx;
        r(2, 0) = txz - twy; r(2, 1) = tyz + twx; r(2, 2) = 1.0 - (txx + tyy);
        return r;
    }
};
}  // namespace Eigen

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

// the reference's own headers bring these two typedefs (include2/auxiliar.h:42-43, include/mapFeatures.h:35-36) and `using namespace Eigen`
using namespace Eigen;
typedef Eigen::Matrix<double, 6, 1> Vector6d;
typedef Eigen::Matrix<double, 6, 6> Matrix6d;

namespace g2o {
using namespace Eigen;

struct HyperGraphVertex { virtual ~HyperGraphVertex() {} };

// the members of g2o::BaseVertex<D, T> the reference header touches
template <int D, typename T> class BaseVertex : public HyperGraphVertex {
public:
    static const int Dimension = D;
    typedef T EstimateType;
    BaseVertex() : _estimate() {}
    const T &estimate() const { return _estimate; }
    void setEstimate(const T &e) { _estimate = e; }
    void oplus(const double *v) { oplusImpl(v); }
    virtual bool read(std::istream &) = 0;
    virtual bool write(std::ostream &) const = 0;
    virtual void setToOriginImpl() = 0;
    virtual void oplusImpl(const double *update) = 0;
    virtual int estimateDimension() const { return D; }
protected:
    T _estimate;
};

// the members of g2o::BaseBinaryEdge<D, E, VertexXi, VertexXj> the reference header touches
template <int D, typename E, typename VertexXi, typename VertexXj> class BaseBinaryEdge {
public:
    typedef Matrix<double, D, 1> ErrorVector;
    BaseBinaryEdge() : _vertices(2, nullptr) { _information.setIdentity(); }
    virtual ~BaseBinaryEdge() {}
    void setVertex(int i, HyperGraphVertex *v) { _vertices[i] = v; }
    void setMeasurement(const E &m) { _measurement = m; }
    const ErrorVector &error() const { return _error; }
    double chi2() const { return _error.dot(_information * _error); }
    const Matrix<double, D, VertexXi::Dimension> &jacobianOplusXi() const { return _jacobianOplusXi; }
    const Matrix<double, D, VertexXj::Dimension> &jacobianOplusXj() const { return _jacobianOplusXj; }
    virtual void computeError() = 0;
    virtual void linearizeOplus() = 0;
protected:
    std::vector<HyperGraphVertex *> _vertices;
    E _measurement;
    ErrorVector _error;
    Matrix<double, D, D> _information;
    Matrix<double, D, VertexXi::Dimension> _jacobianOplusXi;
    Matrix<double, D, VertexXj::Dimension> _jacobianOplusXj;
};

template <typename M> class LinearSolverEigen {};
struct BlockSolverX { typedef int PoseMatrixType; };
}  // namespace g2o
