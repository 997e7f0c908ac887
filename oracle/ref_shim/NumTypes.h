// STAND-IN for the reference's include/NumTypes.h — TEST INFRASTRUCTURE ONLY, used by oracle/ref_pin to compile parts of the
// reference's OWN sources where they lie under /root/reference (Residuals.cc, ImmaturePoint.cc, CoarseTracker.cc, MatrixAccumulators.h,
// GlobalFuncs.h, ResidualProjections.h, AffLight.h, Setting.cc), so that the oracle restatement can be pinned against the reference's own
// code. The real NumTypes.h pulls in Eigen3, Sophus, glog and DBoW3, none of which exist in this image. Those sources use a small
// part of Eigen: fixed-size matrices with element access, +, -, scalar and matrix products, dot / norm / sum, head / tail / segment /
// block views, cast, a 3x3 inverse, the comma initialiser and (in the tracker's LM loop only) ldlt().solve(). This file provides
// exactly that, evaluated eagerly, entry by entry, sums accumulated left to right in the operand order written at the call site —
// which is what Eigen's expression templates evaluate to for these fixed small sizes. ldlt().solve() and Sophus::SE3d are forwarded
// to the oracle's own restatements (omath.h), i.e. those two pieces are NOT pinned by anything compiled against this header.
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
#include <algorithm>
#include <map>
#include <iostream>
#include <fstream>
#include <memory>
#include <string>
#include <vector>

using namespace std;

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_ALIGN16 alignas(16)
#define EIGEN_ALWAYS_INLINE inline
#define EIGEN_STRONG_INLINE inline

// implemented by the pin harness with the oracle's restatement of Eigen::LDLT (oracle/omath.h)
extern "C" void ref_shim_ldlt_solve(int n, const double *A_colmajor, const double *b, double *x);
extern "C" void ref_shim_inverse_lu(int n, const double *A_colmajor, double *Ainv_colmajor);      // Eigen PartialPivLU inverse (sizes > 4)
extern "C" void ref_shim_jacobi_svd(int m, int n, const double *A_colmajor, double *U, double *S, double *V);

namespace Eigen {
template<typename T, int R, int C> struct Matrix;
template<typename T, int R, int C> struct Block;
template<typename T, int N> struct RowView { const Matrix<T, N, 1> *v; };
template<typename T, int R, int C> struct CommaInit {
    Matrix<T, R, C> *m; int k;
    template<typename S> CommaInit &operator,(S v) { m->d[(k % C) * R + (k / C)] = (T) v; k++; return *this; }     // row by row
};
template<typename T, int N> struct LDLTOf {
    Matrix<T, N, N> A;
    Matrix<T, N, 1> solve(const Matrix<T, N, 1> &b) const;
};

// Eigen aligns fixed-size objects whose size is a multiple of 16 bytes to 16 bytes (SSE build); the reference's _mm_load_ps on
// RawResidualJacobian members relies on the struct layout that follows from it
template<typename T, int R, int C>
struct alignas((sizeof(T) * R * C) % 16 == 0 ? 16 : alignof(T)) Matrix {
    T d[R * C];        // column-major like Eigen's default
    Matrix() {}
    Matrix(T a, T b) { static_assert(R * C == 2, "size"); d[0] = a; d[1] = b; }
    Matrix(T a, T b, T c) { static_assert(R * C == 3, "size"); d[0] = a; d[1] = b; d[2] = c; }
    Matrix(T a, T b, T c, T e) { static_assert(R * C == 4, "size"); d[0] = a; d[1] = b; d[2] = c; d[3] = e; }
    // Eigen lets a 1xN row initialise an Nx1 column (vector <- vector): EnergyFunctional::adHTdeltaF is Mat18f, read as Vec8f
    template<int R2, int C2, typename = typename std::enable_if<(R2 != R || C2 != C) && R2 * C2 == R * C && (R2 == 1 || C2 == 1) && (R == 1 || C == 1)>::type>
    Matrix(const Matrix<T, R2, C2> &o) { for (int i = 0; i < R * C; i++) d[i] = o.d[i]; }
    Matrix(const Matrix &) = default;
    Matrix &operator=(const Matrix &) = default;
    T &operator()(int r, int c) { return d[c * R + r]; }
    const T &operator()(int r, int c) const { return d[c * R + r]; }
    T &operator[](int i) { return d[i]; }
    const T &operator[](int i) const { return d[i]; }
    void setZero() { for (int i = 0; i < R * C; i++) d[i] = T(0); }
    void setIdentity() { setZero(); for (int i = 0; i < (R < C ? R : C); i++) d[i * R + i] = T(1); }
    static Matrix Zero() { Matrix m; m.setZero(); return m; }
    static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
    static Matrix Constant(T v) { Matrix m; for (int i = 0; i < R * C; i++) m.d[i] = v; return m; }
    void setConstant(T v) { for (int i = 0; i < R * C; i++) d[i] = v; }
    Matrix &operator+=(const Matrix &o) { for (int i = 0; i < R * C; i++) d[i] += o.d[i]; return *this; }
    Matrix &operator-=(const Matrix &o) { for (int i = 0; i < R * C; i++) d[i] -= o.d[i]; return *this; }
    template<typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    Matrix &operator*=(S s_) { const T s = (T) s_; for (int i = 0; i < R * C; i++) d[i] *= s; return *this; }
    Matrix operator-() const { Matrix o; for (int i = 0; i < R * C; i++) o.d[i] = -d[i]; return o; }
    template<typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    Matrix &operator/=(S s_) { const T s = (T) s_; for (int i = 0; i < R * C; i++) d[i] /= s; return *this; }
    template<typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    Matrix operator/(S s_) const { Matrix o = *this; o /= s_; return o; }
    template<typename U> Matrix<U, R, C> cast() const { Matrix<U, R, C> o; for (int i = 0; i < R * C; i++) o.d[i] = (U) d[i]; return o; }
    template<int R1 = R, int C1 = C, typename = typename std::enable_if<R1 * C1 == 1>::type> operator T() const { return d[0]; }      // 1x1 -> scalar
    T dot(const Matrix &o) const { T s = d[0] * o.d[0]; for (int i = 1; i < R * C; i++) s += d[i] * o.d[i]; return s; }
    T squaredNorm() const { T s = d[0] * d[0]; for (int i = 1; i < R * C; i++) s += d[i] * d[i]; return s; }
    T norm() const { return std::sqrt(squaredNorm()); }
    T sum() const { T s = d[0]; for (int i = 1; i < R * C; i++) s += d[i]; return s; }
    // a column vector's transpose stays a view (row * matrix, column * row products below); any other shape transposes by value
    template<int C1 = C> typename std::enable_if<C1 == 1, RowView<T, R>>::type transpose() const { return RowView<T, R>{this}; }
    template<int C1 = C> typename std::enable_if<C1 != 1, Matrix<T, C, R>>::type transpose() const {
        Matrix<T, C, R> o; for (int c = 0; c < C; c++) for (int r = 0; r < R; r++) o(c, r) = (*this)(r, c); return o;
    }
    T *data() { return d; }
    const T *data() const { return d; }
    Matrix cwiseProduct(const Matrix &o) const { Matrix m; for (int i = 0; i < R * C; i++) m.d[i] = d[i] * o.d[i]; return m; }
    CommaInit<T, R, C> operator<<(T v) { d[0] = v; return CommaInit<T, R, C>{this, 1}; }
    LDLTOf<T, R> ldlt() const { static_assert(R == C, "square"); return LDLTOf<T, R>{*this}; }
    // 3x3 inverse by cofactors / determinant (Eigen's compute_inverse for size 3)
    template<int R1 = R> typename std::enable_if<(R1 > 4), Matrix>::type inverse() const {       // PartialPivLU path
        static_assert(R == C && std::is_same<T, double>::value, "square double"); Matrix o; ref_shim_inverse_lu(R, d, o.d); return o;
    }
    template<int R1 = R> typename std::enable_if<(R1 <= 4), Matrix>::type inverse() const {
        static_assert(R == 3 && C == 3, "only the 3x3 cofactor inverse is used");
        const Matrix &m = *this; Matrix o;
        const T c00 = m(1, 1) * m(2, 2) - m(1, 2) * m(2, 1), c10 = m(1, 2) * m(2, 0) - m(1, 0) * m(2, 2), c20 = m(1, 0) * m(2, 1) - m(1, 1) * m(2, 0);
        const T invdet = T(1) / (m(0, 0) * c00 + m(0, 1) * c10 + m(0, 2) * c20);
        o(0, 0) = c00 * invdet; o(1, 0) = c10 * invdet; o(2, 0) = c20 * invdet;
        o(0, 1) = (m(0, 2) * m(2, 1) - m(0, 1) * m(2, 2)) * invdet; o(1, 1) = (m(0, 0) * m(2, 2) - m(0, 2) * m(2, 0)) * invdet; o(2, 1) = (m(2, 0) * m(0, 1) - m(0, 0) * m(2, 1)) * invdet;
        o(0, 2) = (m(0, 1) * m(1, 2) - m(0, 2) * m(1, 1)) * invdet; o(1, 2) = (m(1, 0) * m(0, 2) - m(0, 0) * m(1, 2)) * invdet; o(2, 2) = (m(0, 0) * m(1, 1) - m(1, 0) * m(0, 1)) * invdet;
        return o;
    }
    // views: on a non-const object a Block (a value that also writes through on =, *=, setZero), on a const object a value
    template<int BR, int BC> Block<T, BR, BC> block(int r, int c) { return Block<T, BR, BC>(d + c * R + r, R); }
    template<int BR, int BC> Matrix<T, BR, BC> block(int r, int c) const { return Block<T, BR, BC>(const_cast<T *>(d) + c * R + r, R); }
    template<int BR, int BC> Block<T, BR, BC> topLeftCorner() { return block<BR, BC>(0, 0); }
    template<int BR, int BC> Matrix<T, BR, BC> topLeftCorner() const { return block<BR, BC>(0, 0); }
    template<int BR, int BC> Block<T, BR, BC> topRightCorner() { return block<BR, BC>(0, C - BC); }
    template<int BR, int BC> Matrix<T, BR, BC> topRightCorner() const { return block<BR, BC>(0, C - BC); }
    Block<T, R, 1> col(int c) { return Block<T, R, 1>(d + c * R, R); }
    Block<T, 1, C> row(int r) { return Block<T, 1, C>(d + r, R); }
    template<int N> Block<T, N, 1> segment(int i0) { static_assert(C == 1 || R == 1, "vector"); return Block<T, N, 1>(d + i0, N); }
    template<int N> Matrix<T, N, 1> segment(int i0) const { Matrix<T, N, 1> o; for (int i = 0; i < N; i++) o.d[i] = d[i0 + i]; return o; }
    template<int N> Block<T, N, 1> head() { return segment<N>(0); }
    template<int N> Matrix<T, N, 1> head() const { return segment<N>(0); }
    template<int N> Block<T, N, 1> tail() { return segment<N>(R * C - N); }
    template<int N> Matrix<T, N, 1> tail() const { return segment<N>(R * C - N); }
};

template<typename T, int BR, int BC>
struct Block : Matrix<T, BR, BC> {
    T *p; int ld;      // top-left element of the viewed region and the column stride of its matrix
    Block(T *p_, int ld_) : p(p_), ld(ld_) { for (int c = 0; c < BC; c++) for (int r = 0; r < BR; r++) this->d[c * BR + r] = p[c * ld + r]; }
    void push() { for (int c = 0; c < BC; c++) for (int r = 0; r < BR; r++) p[c * ld + r] = this->d[c * BR + r]; }
    Block &operator=(const Matrix<T, BR, BC> &m) { for (int i = 0; i < BR * BC; i++) this->d[i] = m.d[i]; push(); return *this; }
    Block &operator=(const Block &m) { return *this = static_cast<const Matrix<T, BR, BC> &>(m); }
    template<typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
    Block &operator*=(S s) { Matrix<T, BR, BC>::operator*=(s); push(); return *this; }
    void setZero() { Matrix<T, BR, BC>::setZero(); push(); }
    Block &noalias() { return *this; }
    struct Diag { Block *b; Diag &operator+=(const Matrix<T, BR, 1> &v) { for (int i = 0; i < BR; i++) b->d[i * BR + i] += v.d[i]; b->push(); return *this; } };
    Diag diagonal() { static_assert(BR == BC, "square"); return Diag{this}; }
    Block &operator+=(const Matrix<T, BR, BC> &m) { Matrix<T, BR, BC>::operator+=(m); push(); return *this; }
    Block &operator-=(const Matrix<T, BR, BC> &m) { Matrix<T, BR, BC>::operator-=(m); push(); return *this; }
};
// ---- dynamic-size vectors and matrices (VecX, VecXf, MatXX) as EnergyFunctional.cc and the Hessian stitchers use them. Everything is
// evaluated eagerly into values; products are row-times-column sums accumulated from 0 in increasing index order.
template<typename T> struct DynVecT;
struct DynMat;
template<typename T> struct DynSeg {        // v.head(n) / v.tail(n) / v.segment(i, n): a view that assigns through
    DynVecT<T> *v; int i0, n;
    operator DynVecT<T>() const;
    DynSeg &operator=(const DynVecT<T> &o);
    DynSeg &operator=(const DynSeg &o) { return *this = (DynVecT<T>) o; }
    DynSeg &noalias() { return *this; }
    DynSeg &operator-=(const DynVecT<T> &o);
};
template<typename T> struct DiagWrap { const DynVecT<T> *v; };      // v.asDiagonal()
template<typename T> struct DynVecT {
    std::vector<T> d;
    DynVecT() {}
    explicit DynVecT(int n) : d((size_t) n, T(0)) {}
    template<int N> DynVecT(const Matrix<T, N, 1> &m) : d(m.d, m.d + N) {}
    static DynVecT Zero(int n) { return DynVecT(n); }
    static DynVecT Constant(int n, T v) { DynVecT o(n); for (auto &x : o.d) x = v; return o; }
    int size() const { return (int) d.size(); }
    int rows() const { return (int) d.size(); }
    T &operator[](int i) { return d[i]; }
    const T &operator[](int i) const { return d[i]; }
    T &operator()(int i) { return d[i]; }
    const T &operator()(int i) const { return d[i]; }
    void conservativeResize(int n) { d.resize((size_t) n, T(0)); }
    template<int N> Block<T, N, 1> segment(int i0) { return Block<T, N, 1>(&d[i0], N); }
    template<int N> Matrix<T, N, 1> segment(int i0) const { Matrix<T, N, 1> o; for (int i = 0; i < N; i++) o.d[i] = d[i0 + i]; return o; }
    template<int N> Block<T, N, 1> head() { return segment<N>(0); }
    template<int N> Matrix<T, N, 1> head() const { return segment<N>(0); }
    template<int N> Block<T, N, 1> tail() { return segment<N>((int) d.size() - N); }
    template<int N> Matrix<T, N, 1> tail() const { return segment<N>((int) d.size() - N); }
    DynSeg<T> segment(int i0, int n) { return DynSeg<T>{this, i0, n}; }
    DynSeg<T> head(int n) { return DynSeg<T>{this, 0, n}; }
    DynSeg<T> tail(int n) { return DynSeg<T>{this, (int) d.size() - n, n}; }
    DynVecT &noalias() { return *this; }
    DynVecT &operator+=(const DynVecT &o) { for (size_t i = 0; i < d.size(); i++) d[i] += o.d[i]; return *this; }
    DynVecT &operator-=(const DynVecT &o) { for (size_t i = 0; i < d.size(); i++) d[i] -= o.d[i]; return *this; }
    DynVecT operator-() const { DynVecT o = *this; for (auto &x : o.d) x = -x; return o; }
    T dot(const DynVecT &o) const { T s = T(0); for (size_t i = 0; i < d.size(); i++) s += d[i] * o.d[i]; return s; }
    T norm() const { return std::sqrt(dot(*this)); }
    DynVecT normalized() const { const T n = norm(); DynVecT o = *this; for (auto &x : o.d) x = x / n; return o; }
    DynVecT cwiseAbs() const { DynVecT o = *this; for (auto &x : o.d) x = std::fabs(x); return o; }
    DynVecT cwiseSqrt() const { DynVecT o = *this; for (auto &x : o.d) x = std::sqrt(x); return o; }
    DynVecT cwiseInverse() const { DynVecT o = *this; for (auto &x : o.d) x = T(1) / x; return o; }
    template<typename U> DynVecT<U> cast() const { DynVecT<U> o((int) d.size()); for (size_t i = 0; i < d.size(); i++) o.d[i] = (U) d[i]; return o; }
    DiagWrap<T> asDiagonal() const { return DiagWrap<T>{this}; }
};
template<typename T> DynSeg<T>::operator DynVecT<T>() const { DynVecT<T> o(n); for (int i = 0; i < n; i++) o.d[i] = v->d[i0 + i]; return o; }
template<typename T> DynSeg<T> &DynSeg<T>::operator=(const DynVecT<T> &o) { for (int i = 0; i < n; i++) v->d[i0 + i] = o.d[i]; return *this; }
template<typename T> DynSeg<T> &DynSeg<T>::operator-=(const DynVecT<T> &o) { for (int i = 0; i < n; i++) v->d[i0 + i] -= o.d[i]; return *this; }
typedef DynVecT<double> DynVec;
inline DynVec operator+(const DynVec &a, const DynVec &b) { DynVec o = a; o += b; return o; }
inline DynVec operator-(const DynVec &a, const DynVec &b) { DynVec o = a; o -= b; return o; }
inline DynVec operator*(double s, const DynVec &a) { DynVec o = a; for (auto &x : o.d) x = s * x; return o; }
inline DynVec operator*(const DiagWrap<double> &D, const DynVec &a) { DynVec o = a; for (size_t i = 0; i < o.d.size(); i++) o.d[i] = D.v->d[i] * a.d[i]; return o; }

// N consecutive diagonal entries of a dynamic matrix (H.diagonal().segment<8>(i) += v)
template<int N> struct DiagSeg {
    double *p; int stride;
    DiagSeg &operator+=(const Matrix<double, N, 1> &v) { for (int i = 0; i < N; i++) p[(size_t) i * stride] += v.d[i]; return *this; }
};
struct DiagView {
    double *p; int stride, n;
    template<int N> DiagSeg<N> segment(int i0) { return DiagSeg<N>{p + (size_t) i0 * stride, stride}; }
    template<int N> DiagSeg<N> head() { return segment<N>(0); }
    operator DynVec() const { DynVec o(n); for (int i = 0; i < n; i++) o.d[i] = p[(size_t) i * stride]; return o; }
    DynVec cwiseAbs() const { return ((DynVec) *this).cwiseAbs(); }
    DynVec cwiseSqrt() const { return ((DynVec) *this).cwiseSqrt(); }
};
struct DynBlk {          // M.block(i, j, r, c) and the corner / rows / cols variants: a view that assigns through
    DynMat *m; int i0, j0, r, c;
    operator DynMat() const;
    DynBlk &operator=(const DynMat &o);
    DynBlk &operator=(const DynBlk &o);
    DynBlk &noalias() { return *this; }
    DynBlk &operator-=(const DynMat &o);
    DynMat transpose() const;
    void setZero();
};
struct DynCol { DynMat *m; int j; DynCol &operator=(const DynVec &v); };
struct DynLDLT;
struct DynMat {
    int r = 0, c = 0;
    std::vector<double> d;       // column-major
    DynMat() {}
    template<typename I1, typename I2, typename = typename std::enable_if<std::is_integral<I1>::value && std::is_integral<I2>::value>::type>
    DynMat(I1 r_, I2 c_) : r((int) r_), c((int) c_), d((size_t) r_ * (size_t) c_, 0.0) {}
    template<int R, int C, typename = typename std::enable_if<(C > 1)>::type> DynMat(const Matrix<double, R, C> &m) : r(R), c(C), d(m.d, m.d + R * C) {}
    static DynMat Zero(int r_, int c_) { return DynMat(r_, c_); }
    int rows() const { return r; }
    int cols() const { return c; }
    double &operator()(int i, int j) { return d[(size_t) j * r + i]; }
    double operator()(int i, int j) const { return d[(size_t) j * r + i]; }
    void conservativeResize(int nr, int nc) {
        DynMat o(nr, nc);
        for (int j = 0; j < std::min(c, nc); j++) for (int i = 0; i < std::min(r, nr); i++) o(i, j) = (*this)(i, j);
        *this = o;
    }
    template<int BR, int BC> Block<double, BR, BC> block(int i, int j) { return Block<double, BR, BC>(&d[(size_t) j * r + i], r); }
    template<int BR, int BC> Block<double, BR, BC> topLeftCorner() { return block<BR, BC>(0, 0); }
    template<int BR, int BC> Block<double, BR, BC> bottomRightCorner() { return block<BR, BC>(r - BR, c - BC); }
    template<int N> DynBlk rightCols() { return DynBlk{this, 0, c - N, r, N}; }
    template<int N> DynBlk bottomRows() { return DynBlk{this, r - N, 0, N, c}; }
    DynBlk block(int i, int j, int br, int bc) { return DynBlk{this, i, j, br, bc}; }
    DynBlk topLeftCorner(int br, int bc) { return DynBlk{this, 0, 0, br, bc}; }
    DynBlk bottomLeftCorner(int br, int bc) { return DynBlk{this, r - br, 0, br, bc}; }
    DynBlk rightCols(int n) { return DynBlk{this, 0, c - n, r, n}; }
    DynBlk bottomRows(int n) { return DynBlk{this, r - n, 0, n, c}; }
    DynCol col(int j) { return DynCol{this, j}; }
    DiagView diagonal() { return DiagView{d.data(), r + 1, std::min(r, c)}; }
    DynMat &noalias() { return *this; }
    DynMat &operator+=(const DynMat &o) { for (size_t i = 0; i < d.size(); i++) d[i] += o.d[i]; return *this; }
    DynMat &operator-=(const DynMat &o) { for (size_t i = 0; i < d.size(); i++) d[i] -= o.d[i]; return *this; }
    DynMat transpose() const { DynMat o(c, r); for (int j = 0; j < c; j++) for (int i = 0; i < r; i++) o(j, i) = (*this)(i, j); return o; }
    inline DynLDLT ldlt() const;
};
inline DynBlk::operator DynMat() const { DynMat o(r, c); for (int j = 0; j < c; j++) for (int i = 0; i < r; i++) o(i, j) = (*m)(i0 + i, j0 + j); return o; }
inline DynBlk &DynBlk::operator=(const DynMat &o) { for (int j = 0; j < c; j++) for (int i = 0; i < r; i++) (*m)(i0 + i, j0 + j) = o(i, j); return *this; }
inline DynBlk &DynBlk::operator=(const DynBlk &o) { return *this = (DynMat) o; }
inline DynBlk &DynBlk::operator-=(const DynMat &o) { for (int j = 0; j < c; j++) for (int i = 0; i < r; i++) (*m)(i0 + i, j0 + j) -= o(i, j); return *this; }
inline DynMat DynBlk::transpose() const { return ((DynMat) *this).transpose(); }
inline void DynBlk::setZero() { for (int j = 0; j < c; j++) for (int i = 0; i < r; i++) (*m)(i0 + i, j0 + j) = 0.0; }
inline DynCol &DynCol::operator=(const DynVec &v) { for (int i = 0; i < m->r; i++) (*m)(i, j) = v.d[i]; return *this; }
inline DynMat operator+(const DynMat &a, const DynMat &b) { DynMat o = a; o += b; return o; }
inline DynMat operator-(const DynMat &a, const DynMat &b) { DynMat o = a; o -= b; return o; }
template<typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline DynMat operator*(S s_, const DynMat &a) { const double s = (double) s_; DynMat o = a; for (auto &x : o.d) x = s * x; return o; }
template<typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline DynMat operator*(const DynMat &a, S s_) { const double s = (double) s_; DynMat o = a; for (auto &x : o.d) x = x * s; return o; }
inline DynMat operator*(const DynMat &A, const DynMat &B) {
    DynMat C(A.r, B.c);
    for (int j = 0; j < B.c; j++) for (int i = 0; i < A.r; i++) { double s = 0.0; for (int k = 0; k < A.c; k++) s += A(i, k) * B(k, j); C(i, j) = s; }
    return C;
}
inline DynVec operator*(const DynMat &A, const DynVec &v) {
    DynVec o(A.r);
    for (int i = 0; i < A.r; i++) { double s = 0.0; for (int k = 0; k < A.c; k++) s += A(i, k) * v.d[k]; o.d[i] = s; }
    return o;
}
inline DynMat operator*(const DiagWrap<double> &D, const DynMat &A) { DynMat o = A; for (int j = 0; j < A.c; j++) for (int i = 0; i < A.r; i++) o(i, j) = D.v->d[i] * A(i, j); return o; }
inline DynMat operator*(const DynMat &A, const DiagWrap<double> &D) { DynMat o = A; for (int j = 0; j < A.c; j++) for (int i = 0; i < A.r; i++) o(i, j) = A(i, j) * D.v->d[j]; return o; }
struct DynLDLT { DynMat A; DynVec solve(const DynVec &b) const { DynVec x(A.r); ref_shim_ldlt_solve(A.r, A.d.data(), b.d.data(), x.d.data()); return x; } };
inline DynLDLT DynMat::ldlt() const { return DynLDLT{*this}; }
// Eigen::JacobiSVD<MatXX>(M, ComputeThinU | ComputeThinV): forwarded to the oracle's restatement (omath.h jacobi_svd) by the harness
enum { ComputeThinU = 1, ComputeThinV = 2 };
template<typename M> struct JacobiSVD {
    DynMat U, V; DynVec S;
    JacobiSVD(const DynMat &A, unsigned) {
        const int k = std::min(A.r, A.c);
        U = DynMat(A.r, k); V = DynMat(A.c, k); S = DynVec(k);
        ref_shim_jacobi_svd(A.r, A.c, A.d.data(), U.d.data(), S.d.data(), V.d.data());
    }
    const DynVec &singularValues() const { return S; }
    const DynMat &matrixU() const { return U; }
    const DynMat &matrixV() const { return V; }
};
template<typename T> using aligned_allocator = std::allocator<T>;
template<typename T, int R, int C, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline Matrix<T, R, C> operator*(S s_, const Matrix<T, R, C> &m) {
    const T s = (T) s_;     // Eigen converts the scalar to the matrix's scalar type first
    Matrix<T, R, C> o; for (int i = 0; i < R * C; i++) o.d[i] = s * m.d[i]; return o;
}
template<typename T, int R, int C, typename S, typename = typename std::enable_if<std::is_arithmetic<S>::value>::type>
inline Matrix<T, R, C> operator*(const Matrix<T, R, C> &m, S s_) {
    const T s = (T) s_;
    Matrix<T, R, C> o; for (int i = 0; i < R * C; i++) o.d[i] = m.d[i] * s; return o;
}
template<typename T, int R, int C> inline Matrix<T, R, C> operator+(const Matrix<T, R, C> &a, const Matrix<T, R, C> &b) {
    Matrix<T, R, C> o; for (int i = 0; i < R * C; i++) o.d[i] = a.d[i] + b.d[i]; return o;
}
template<typename T, int R, int C> inline Matrix<T, R, C> operator-(const Matrix<T, R, C> &a, const Matrix<T, R, C> &b) {
    Matrix<T, R, C> o; for (int i = 0; i < R * C; i++) o.d[i] = a.d[i] - b.d[i]; return o;
}
template<typename T> inline T &operator-=(T &a, const Matrix<T, 1, 1> &m) { a -= m.d[0]; return a; }      // scalar -= (row * column)
template<typename T> inline T &operator+=(T &a, const Matrix<T, 1, 1> &m) { a += m.d[0]; return a; }
// column * row^T (outer product)
template<typename T, int R, int C> inline Matrix<T, R, C> operator*(const Matrix<T, R, 1> &col, const RowView<T, C> &row) {
    Matrix<T, R, C> o; for (int c = 0; c < C; c++) for (int r = 0; r < R; r++) o.d[c * R + r] = col.d[r] * row.v->d[c]; return o;
}
// row^T * matrix -> row (entries accumulated left to right)
template<typename T, int R, int C> inline Matrix<T, 1, C> operator*(const RowView<T, R> &row, const Matrix<T, R, C> &M) {
    Matrix<T, 1, C> o;
    for (int c = 0; c < C; c++) { T s = row.v->d[0] * M(0, c); for (int r = 1; r < R; r++) s += row.v->d[r] * M(r, c); o.d[c] = s; }
    return o;
}
// small fixed matrix * matrix / vector: each entry is the row-times-column sum accumulated left to right (what Eigen's unrolled
// coefficient-based product gives for these sizes)
template<typename T, int R, int K, int C> inline Matrix<T, R, C> operator*(const Matrix<T, R, K> &A, const Matrix<T, K, C> &B) {
    Matrix<T, R, C> o;
    for (int c = 0; c < C; c++) for (int r = 0; r < R; r++) { T s = A(r, 0) * B(0, c); for (int k = 1; k < K; k++) s += A(r, k) * B(k, c); o(r, c) = s; }
    return o;
}
template<typename T, int N> inline Matrix<T, N, 1> LDLTOf<T, N>::solve(const Matrix<T, N, 1> &b) const {
    // double systems go to the oracle's LDLT restatement; the float systems of CoarseInitializer::trackFrame are solved through it in
    // double and rounded (nothing that is pinned depends on them)
    double Ad[N * N], bd[N], xd[N];
    for (int i = 0; i < N * N; i++) Ad[i] = (double) A.d[i];
    for (int i = 0; i < N; i++) bd[i] = (double) b.d[i];
    ref_shim_ldlt_solve(N, Ad, bd, xd);
    Matrix<T, N, 1> x; for (int i = 0; i < N; i++) x.d[i] = (T) xd[i]; return x;
}
// Eigen::DiagonalMatrix<T, N> as CoarseInitializer uses it (wM: the SCALE_* weights)
template<typename T, int N> struct DiagonalMatrix {
    Matrix<T, N, 1> dg;
    Matrix<T, N, 1> &diagonal() { return dg; }
    const Matrix<T, N, 1> &diagonal() const { return dg; }
    Matrix<T, N, N> toDenseMatrix() const { Matrix<T, N, N> m; m.setZero(); for (int i = 0; i < N; i++) m(i, i) = dg.d[i]; return m; }
};
template<typename T, int N, int C> inline Matrix<T, N, C> operator*(const DiagonalMatrix<T, N> &D, const Matrix<T, N, C> &M) {
    Matrix<T, N, C> o; for (int c = 0; c < C; c++) for (int r = 0; r < N; r++) o(r, c) = D.dg.d[r] * M(r, c); return o;
}
template<typename T, int R, int N> inline Matrix<T, R, N> operator*(const Matrix<T, R, N> &M, const DiagonalMatrix<T, N> &D) {
    Matrix<T, R, N> o; for (int c = 0; c < N; c++) for (int r = 0; r < R; r++) o(r, c) = M(r, c) * D.dg.d[c]; return o;
}
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<float, 4, 1> Vector4f;
typedef Matrix<int, 2, 1> Vector2i;
typedef Matrix<int, 3, 1> Vector3i;
}  // namespace Eigen

const int CPARS = 4;
const int MAX_RES_PER_POINT = 8;
typedef Eigen::Matrix<double, 2, 1> Vec2;
typedef Eigen::Matrix<double, 3, 1> Vec3;
typedef Eigen::Matrix<double, 5, 1> Vec5;
typedef Eigen::Matrix<double, 6, 1> Vec6;
typedef Eigen::Matrix<double, 7, 1> Vec7;
typedef Eigen::Matrix<double, 8, 1> Vec8;
typedef Eigen::Matrix<double, 10, 1> Vec10;
typedef Eigen::Matrix<double, 3, 3> Mat33;
typedef Eigen::Matrix<double, 6, 6> Mat66;
typedef Eigen::Matrix<double, 4, 2> Mat42;
typedef Eigen::Matrix<double, 7, 7> Mat77;
typedef Eigen::Matrix<double, 8, 8> Mat88;
typedef Eigen::Matrix<float, 8, 8> Mat88f;
typedef Eigen::Matrix<double, 8, CPARS> Mat8C;
typedef Eigen::Matrix<double, CPARS, CPARS> MatCC;
typedef Eigen::Matrix<double, 8 + CPARS + 1, 8 + CPARS + 1> MatPCPC;
typedef Eigen::Matrix<double, CPARS, 1> VecC;
typedef Eigen::DynMat MatXX;
typedef Eigen::DynVec VecX;
typedef Eigen::DynVecT<float> VecXf;
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
typedef Eigen::Matrix<float, 10, 1> Vec10f;
typedef Eigen::Matrix<float, 14, 1> Vec14f;
typedef Eigen::Matrix<float, 9, 9> Mat99f;
typedef Eigen::Matrix<float, 13, 13> Mat1313f;
typedef Eigen::Matrix<float, 14, 14> Mat1414f;
// Sophus::SE3d as the tracker uses it (exp, *, inverse, rotationMatrix, translation): forwarded to the oracle's restatement of
// Sophus (oracle/omath.h) — so the SE(3) algebra is shared by both sides of the pin and is NOT itself pinned
#include "../omath.h"
namespace Sophus {
class SE3d {
public:
    oracle::Quat q;       // unit quaternion (w, x, y, z)
    Vec3 t = Vec3(0, 0, 0);
    SE3d() {}
    explicit SE3d(const oracle::SE3 &s) : q(s.q), t(s.t[0], s.t[1], s.t[2]) {}
    SE3d(const Mat33 &R, const Vec3 &tr) {      // Sophus::SE3d(rotation_matrix, translation) (thirdparty/sophus/se3.hpp)
        double Rm[9], tv[3];
        for (int i = 0; i < 3; i++) { tv[i] = tr[i]; for (int j = 0; j < 3; j++) Rm[i * 3 + j] = R(i, j); }
        const oracle::SE3 s = oracle::SE3::fromRt(Rm, tv);
        q = s.q; t = Vec3(s.t[0], s.t[1], s.t[2]);
    }
    oracle::SE3 o() const { oracle::SE3 s; s.q = q; s.t = oracle::V3{{t[0], t[1], t[2]}}; return s; }
    Mat33 rotationMatrix() const { const oracle::M3 R = oracle::qmat(q); Mat33 m; for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) m(i, j) = R(i, j); return m; }
    Vec3 &translation() { return t; }
    const Vec3 &translation() const { return t; }
    static SE3d exp(const Vec6 &a) { return SE3d(oracle::SE3::exp(a.d)); }
    Vec6 log() const { Vec6 v; o().log(v.d); return v; }
    Mat66 Adj() const { double A[36]; o().Adj(A); Mat66 m; for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) m(i, j) = A[i * 6 + j]; return m; }
    SE3d operator*(const SE3d &b) const { return SE3d(o() * b.o()); }
    SE3d inverse() const { return SE3d(o().inverse()); }
};
}
typedef Sophus::SE3d SE3;
