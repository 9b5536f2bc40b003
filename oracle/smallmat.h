// TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Tiny fixed-size dense matrices for the CPU oracle.
// Eigen is not available in this image; this gives the oracle the same "write the formula as in the
// reference" readability (Matrix<double,R,C>, block(), transpose(), products) with no dependency.
#pragma once
#include <cmath>
#include <cstring>

namespace oracle {

template <int R, int C>
struct Mat {
    double m[R * C];
    Mat() { for (int i = 0; i < R * C; i++) m[i] = 0.0; }
    double &operator()(int r, int c) { return m[r * C + c]; }
    double operator()(int r, int c) const { return m[r * C + c]; }
    double &operator[](int i) { return m[i]; }
    double operator[](int i) const { return m[i]; }
    static Mat Identity() { Mat I; for (int i = 0; i < (R < C ? R : C); i++) I(i, i) = 1.0; return I; }
    Mat<C, R> T() const { Mat<C, R> t; for (int r = 0; r < R; r++) for (int c = 0; c < C; c++) t(c, r) = (*this)(r, c); return t; }
    template <int BR, int BC> Mat<BR, BC> block(int r0, int c0) const {
        Mat<BR, BC> b; for (int r = 0; r < BR; r++) for (int c = 0; c < BC; c++) b(r, c) = (*this)(r0 + r, c0 + c); return b;
    }
    template <int BR, int BC> void setBlock(int r0, int c0, const Mat<BR, BC> &b) {
        for (int r = 0; r < BR; r++) for (int c = 0; c < BC; c++) (*this)(r0 + r, c0 + c) = b(r, c);
    }
    Mat<R, 1> col(int c) const { Mat<R, 1> v; for (int r = 0; r < R; r++) v[r] = (*this)(r, c); return v; }
    void setCol(int c, const Mat<R, 1> &v) { for (int r = 0; r < R; r++) (*this)(r, c) = v[r]; }
    double norm() const { double s = 0; for (int i = 0; i < R * C; i++) s += m[i] * m[i]; return std::sqrt(s); }
    double squaredNorm() const { double s = 0; for (int i = 0; i < R * C; i++) s += m[i] * m[i]; return s; }
    Mat &operator+=(const Mat &o) { for (int i = 0; i < R * C; i++) m[i] += o.m[i]; return *this; }
    Mat &operator-=(const Mat &o) { for (int i = 0; i < R * C; i++) m[i] -= o.m[i]; return *this; }
};

template <int R, int K, int C>
inline Mat<R, C> operator*(const Mat<R, K> &a, const Mat<K, C> &b) {
    Mat<R, C> o;
    for (int r = 0; r < R; r++)
        for (int c = 0; c < C; c++) { double s = 0; for (int k = 0; k < K; k++) s += a(r, k) * b(k, c); o(r, c) = s; }
    return o;
}
template <int R, int C> inline Mat<R, C> operator+(Mat<R, C> a, const Mat<R, C> &b) { a += b; return a; }
template <int R, int C> inline Mat<R, C> operator-(Mat<R, C> a, const Mat<R, C> &b) { a -= b; return a; }
template <int R, int C> inline Mat<R, C> operator-(Mat<R, C> a) { for (int i = 0; i < R * C; i++) a.m[i] = -a.m[i]; return a; }
template <int R, int C> inline Mat<R, C> operator*(Mat<R, C> a, double s) { for (int i = 0; i < R * C; i++) a.m[i] *= s; return a; }
template <int R, int C> inline Mat<R, C> operator*(double s, Mat<R, C> a) { return a * s; }
template <int R, int C> inline Mat<R, C> operator/(Mat<R, C> a, double s) { for (int i = 0; i < R * C; i++) a.m[i] /= s; return a; }

typedef Mat<2, 1> V2; typedef Mat<3, 1> V3; typedef Mat<4, 1> V4; typedef Mat<6, 1> V6;
typedef Mat<2, 2> M2; typedef Mat<3, 3> M3; typedef Mat<4, 4> M4; typedef Mat<6, 6> M6;

inline V3 cross(const V3 &a, const V3 &b) {
    V3 c; c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0]; return c;
}
template <int N> inline double dot(const Mat<N, 1> &a, const Mat<N, 1> &b) { double s = 0; for (int i = 0; i < N; i++) s += a[i] * b[i]; return s; }

// General inverse of a small dense matrix by Gauss-Jordan with partial pivoting
// (Eigen's dynamic-size inverse() is PartialPivLU based: same arithmetic class).
template <int N>
inline Mat<N, N> inverse(const Mat<N, N> &A) {
    double a[N][2 * N];
    for (int r = 0; r < N; r++) for (int c = 0; c < N; c++) { a[r][c] = A(r, c); a[r][N + c] = (r == c) ? 1.0 : 0.0; }
    for (int k = 0; k < N; k++) {
        int p = k; double best = std::fabs(a[k][k]);
        for (int r = k + 1; r < N; r++) if (std::fabs(a[r][k]) > best) { best = std::fabs(a[r][k]); p = r; }
        if (p != k) for (int c = 0; c < 2 * N; c++) { double t = a[k][c]; a[k][c] = a[p][c]; a[p][c] = t; }
        double inv = 1.0 / a[k][k];
        for (int c = 0; c < 2 * N; c++) a[k][c] *= inv;
        for (int r = 0; r < N; r++) if (r != k) { double f = a[r][k]; if (f != 0.0) for (int c = 0; c < 2 * N; c++) a[r][c] -= f * a[k][c]; }
    }
    Mat<N, N> o; for (int r = 0; r < N; r++) for (int c = 0; c < N; c++) o(r, c) = a[r][N + c]; return o;
}

}  // namespace oracle
