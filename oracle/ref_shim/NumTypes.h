// STAND-IN for the reference's include/NumTypes.h — TEST INFRASTRUCTURE ONLY, used by oracle/ref_pin to compile a few of the
// reference's OWN headers where they lie under /root/reference (MatrixAccumulators.h, GlobalFuncs.h, AffLight.h) plus its
// src/Setting.cc, so that the oracle restatement can be pinned against the reference's own code for those pieces.
// The real NumTypes.h pulls in Eigen3, Sophus, glog and DBoW3, none of which exist in this image. The three headers above use
// only: fixed-size Eigen::Matrix storage with operator()/operator[], setZero, +=, scalar*vector, vector+vector and one
// column * row^T outer product — that is all this file provides (eager, entry by entry, in the operand order written at the
// call site, which is also what Eigen's expression templates evaluate to for these expressions).
#pragma once
// the reference's own NumTypes.h sits next to AffLight.h and would win the quoted-include lookup: claim its include guard
#define LDSO_NUM_TYPES_H_
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace std;

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_ALIGN16 alignas(16)
#define EIGEN_ALWAYS_INLINE inline
#define EIGEN_STRONG_INLINE inline

namespace Eigen {
template<typename T, int R, int C> struct Matrix;
template<typename T, int N> struct RowView { const Matrix<T, N, 1> *v; };

template<typename T, int R, int C>
struct Matrix {
    T d[R * C];        // column-major like Eigen's default
    Matrix() {}
    Matrix(T a, T b) { static_assert(R * C == 2, "size"); d[0] = a; d[1] = b; }
    Matrix(T a, T b, T c) { static_assert(R * C == 3, "size"); d[0] = a; d[1] = b; d[2] = c; }
    Matrix(T a, T b, T c, T e) { static_assert(R * C == 4, "size"); d[0] = a; d[1] = b; d[2] = c; d[3] = e; }
    T &operator()(int r, int c) { return d[c * R + r]; }
    const T &operator()(int r, int c) const { return d[c * R + r]; }
    T &operator[](int i) { return d[i]; }
    const T &operator[](int i) const { return d[i]; }
    void setZero() { for (int i = 0; i < R * C; i++) d[i] = T(0); }
    Matrix &operator+=(const Matrix &o) { for (int i = 0; i < R * C; i++) d[i] += o.d[i]; return *this; }
    RowView<T, R> transpose() const { static_assert(C == 1, "only column vectors are transposed here"); return RowView<T, R>{this}; }
};
template<typename T, int R, int C> inline Matrix<T, R, C> operator*(T s, const Matrix<T, R, C> &m) {
    Matrix<T, R, C> o; for (int i = 0; i < R * C; i++) o.d[i] = s * m.d[i]; return o;
}
template<typename T, int R, int C> inline Matrix<T, R, C> operator+(const Matrix<T, R, C> &a, const Matrix<T, R, C> &b) {
    Matrix<T, R, C> o; for (int i = 0; i < R * C; i++) o.d[i] = a.d[i] + b.d[i]; return o;
}
template<typename T, int R, int C> inline Matrix<T, R, C> operator*(const Matrix<T, R, 1> &col, const RowView<T, C> &row) {
    Matrix<T, R, C> o; for (int c = 0; c < C; c++) for (int r = 0; r < R; r++) o.d[c * R + r] = col.d[r] * row.v->d[c]; return o;
}
template<typename T, int R, int C> inline Matrix<T, R, C> operator*(const Matrix<T, R, C> &m, T s) {
    Matrix<T, R, C> o; for (int i = 0; i < R * C; i++) o.d[i] = m.d[i] * s; return o;
}
// small fixed matrix * vector: each entry is the row-times-column sum accumulated left to right (what Eigen's unrolled
// coefficient-based product gives for these sizes)
template<typename T, int R, int K> inline Matrix<T, R, 1> operator*(const Matrix<T, R, K> &A, const Matrix<T, K, 1> &x) {
    Matrix<T, R, 1> o;
    for (int r = 0; r < R; r++) { T s = A(r, 0) * x.d[0]; for (int k = 1; k < K; k++) s += A(r, k) * x.d[k]; o.d[r] = s; }
    return o;
}
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
}  // namespace Eigen

const int CPARS = 4;
const int MAX_RES_PER_POINT = 8;
typedef Eigen::Matrix<double, 2, 1> Vec2;
typedef Eigen::Matrix<float, 2, 1> Vec2f;
typedef Eigen::Matrix<float, 3, 1> Vec3f;
typedef Eigen::Matrix<unsigned char, 3, 1> Vec3b;
typedef Eigen::Matrix<float, 3, 3> Mat33f;
typedef Eigen::Matrix<float, 9, 1> Vec9f;
typedef Eigen::Matrix<float, 14, 1> Vec14f;
typedef Eigen::Matrix<float, 9, 9> Mat99f;
typedef Eigen::Matrix<float, 13, 13> Mat1313f;
typedef Eigen::Matrix<float, 14, 14> Mat1414f;
// GlobalFuncs.h's eigenTestNan(const MatXX&) only needs rows(), cols() and operator(); it is not called by the pin
struct MatXX {
    int r = 0, c = 0;
    std::vector<double> d;
    int rows() const { return r; }
    int cols() const { return c; }
    double operator()(int i, int j) const { return d[(size_t) j * r + i]; }
};
