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
#include <type_traits>
#include <immintrin.h>
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <fstream>
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

// Eigen aligns fixed-size objects whose size is a multiple of 16 bytes to 16 bytes (SSE build); the reference's _mm_load_ps on
// RawResidualJacobian members relies on the struct layout that follows from it
template<typename T, int R, int C>
struct alignas((sizeof(T) * R * C) % 16 == 0 ? 16 : alignof(T)) Matrix {
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
    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    // Eigen lets a 1xN row initialise an Nx1 column (vector <- vector): EnergyFunctional::adHTdeltaF is Mat18f, read as Vec8f
    template<int R2, int C2> Matrix(const Matrix<T, R2, C2> &o) { static_assert(R2 * C2 == R * C && (R2 == 1 || C2 == 1) && (R == 1 || C == 1), "vector copy"); for (int i = 0; i < R * C; i++) d[i] = o.d[i]; }
    Matrix(const Matrix &) = default;
    Matrix &operator=(const Matrix &) = default;
    T dot(const Matrix &o) const { T s = d[0] * o.d[0]; for (int i = 1; i < R * C; i++) s += d[i] * o.d[i]; return s; }
    template<int R1 = R, int C1 = C, typename = typename std::enable_if<R1 * C1 == 1>::type> operator T() const { return d[0]; }      // 1x1 -> scalar
    template<int R2, int C2> Matrix<T, R2, C2> topLeftCorner() const { Matrix<T, R2, C2> o; for (int c = 0; c < C2; c++) for (int r = 0; r < R2; r++) o.d[c * R2 + r] = d[c * R + r]; return o; }
    template<int N> Matrix<T, N, 1> tail() const { Matrix<T, N, 1> o; for (int i = 0; i < N; i++) o.d[i] = d[R * C - N + i]; return o; }
    T squaredNorm() const { T s = d[0] * d[0]; for (int i = 1; i < R * C; i++) s += d[i] * d[i]; return s; }
    template<int N> Matrix<T, N, 1> head() const { Matrix<T, N, 1> o; for (int i = 0; i < N; i++) o.d[i] = d[i]; return o; }
    template<int N> struct Seg { T *p; Seg &operator=(const Matrix<T, N, 1> &v) { for (int i = 0; i < N; i++) p[i] = v.d[i]; return *this; } };
    template<int N> Seg<N> segment(int i0) { return Seg<N>{d + i0}; }
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
template<typename T, int R, int C, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline Matrix<T, R, C> operator*(const Matrix<T, R, C> &m, S s_) {
    const T s = (T) s_;     // Eigen converts the scalar to the matrix's scalar type first
    Matrix<T, R, C> o; for (int i = 0; i < R * C; i++) o.d[i] = m.d[i] * s; return o;
}
// row^T * matrix -> row (entries accumulated left to right)
template<typename T, int R, int C> inline Matrix<T, 1, C> operator*(const RowView<T, R> &row, const Matrix<T, R, C> &M) {
    Matrix<T, 1, C> o;
    for (int c = 0; c < C; c++) { T s = row.v->d[0] * M(0, c); for (int r = 1; r < R; r++) s += row.v->d[r] * M(r, c); o.d[c] = s; }
    return o;
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
typedef Eigen::Matrix<float, 2, 2> Mat22f;
typedef Eigen::Matrix<float, 1, 8> Mat18f;
typedef Eigen::Matrix<float, 4, 1> Vec4f;
typedef Eigen::Matrix<float, 6, 1> Vec6f;
typedef Eigen::Matrix<float, 8, 1> Vec8f;
typedef Eigen::Matrix<float, CPARS, 1> VecCf;
typedef Eigen::Matrix<float, MAX_RES_PER_POINT, 1> VecNRf;
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
