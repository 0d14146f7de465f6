// mini_eigen.hpp -- TEST INFRASTRUCTURE ONLY (part of oracle/, never linked into the product).
//
// A small, eagerly evaluated stand-in for the subset of the Eigen 3 API that the reference's hot-path sources
// touch (pop_planar_slam/Thirdparty/isam/include/isam/*.h, isamlib/*.cpp, pop_planar_slam/src/isam_plane3d.{h,cpp}).
// Eigen3 itself is not installed in the build container and cannot be fetched (no network), so `oracle/_ref` compiles
// the UNMODIFIED reference sources against this header instead (recipe: oracle/Makefile, target _ref).  Written from the
// published Eigen API / algorithms (Quaternion <-> rotation matrix, AngleAxis(Quaternion) in its >= 3.3 atan2 form,
// partial-pivot LU inverse, LLT); no Eigen code is copied.  Everything evaluates immediately: no expression templates,
// so aliasing is never an issue, at the price of temporaries.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <limits>
#include <sstream>
#include <type_traits>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 3
#define EIGEN_MINOR_VERSION 0

namespace Eigen {

const int Dynamic = -1;
const int Infinity = -1;
enum { ColMajor = 0, RowMajor = 1, AutoAlign = 0, DontAlign = 2 };
enum { Lower = 1, Upper = 2 };
enum { ComputeFullU = 4, ComputeThinU = 8, ComputeFullV = 16, ComputeThinV = 32 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
typedef std::ptrdiff_t Index;

template <typename S, int R, int C, int Opt = 0, int MR = R, int MC = C> class Matrix;
template <typename P, int BR, int BC> class Block;
template <typename S> class Quaternion;
template <typename S> class AngleAxis;
template <typename D> class ArrayWrap;
template <typename M> class CommaInit;
template <typename M> class LLT;

namespace internal {
[[noreturn]] inline void fail(const char* what) {
  std::fprintf(stderr, "mini_eigen: %s\n", what);
  std::abort();
}
template <int A, int B> struct pick { static const int value = (A != Dynamic ? A : B); };
template <typename T> struct traits;
template <typename S, int R, int C, int O, int MR, int MC> struct traits<Matrix<S, R, C, O, MR, MC>> {
  typedef S Scalar;
  static const int Rows = R, Cols = C, Options = O;
};
template <typename P, int BR, int BC> struct traits<Block<P, BR, BC>> {
  typedef typename traits<typename std::remove_const<P>::type>::Scalar Scalar;
  static const int Rows = BR, Cols = BC, Options = 0;
};
}  // namespace internal

// ----------------------------------------------------------------------------------------------------------------
// read-only interface shared by matrices and block views (CRTP)
// ----------------------------------------------------------------------------------------------------------------
template <typename D>
class MatrixBase {
 public:
  typedef typename internal::traits<D>::Scalar Scalar;
  static const int RowsAtCompileTime = internal::traits<D>::Rows, ColsAtCompileTime = internal::traits<D>::Cols;
  typedef Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> Plain;
  typedef Matrix<Scalar, ColsAtCompileTime, RowsAtCompileTime> PlainT;

  const D& derived() const { return *static_cast<const D*>(this); }
  D& derived() { return *static_cast<D*>(this); }
  Index rows() const { return derived().rows(); }
  Index cols() const { return derived().cols(); }
  Index size() const { return rows() * cols(); }
  Scalar coeff(Index i, Index j) const { return derived().coeff(i, j); }
  Scalar coeff(Index i) const { return cols() == 1 ? coeff(i, 0) : coeff(0, i); }
  Scalar operator()(Index i, Index j) const { return coeff(i, j); }
  Scalar operator()(Index i) const { return coeff(i); }
  Scalar operator[](Index i) const { return coeff(i); }
  Scalar x() const { return coeff(0); }
  Scalar y() const { return coeff(1); }
  Scalar z() const { return coeff(2); }
  Scalar w() const { return coeff(3); }

  Plain eval() const {
    Plain r(rows(), cols());
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) r(i, j) = coeff(i, j);
    return r;
  }
  PlainT transpose() const {
    PlainT r(cols(), rows());
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) r(j, i) = coeff(i, j);
    return r;
  }
  PlainT adjoint() const { return transpose(); }
  Scalar squaredNorm() const {
    Scalar s = 0;
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) s += coeff(i, j) * coeff(i, j);
    return s;
  }
  Scalar norm() const { return std::sqrt(squaredNorm()); }
  Plain normalized() const {
    Plain r = eval();
    Scalar n = norm();
    if (n > Scalar(0)) r /= n;
    return r;
  }
  Scalar sum() const {
    Scalar s = 0;
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) s += coeff(i, j);
    return s;
  }
  Scalar prod() const {
    Scalar s = 1;
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) s *= coeff(i, j);
    return s;
  }
  Scalar mean() const { return sum() / Scalar(size()); }
  Scalar trace() const {
    Scalar s = 0;
    for (Index i = 0; i < std::min(rows(), cols()); i++) s += coeff(i, i);
    return s;
  }
  Scalar maxCoeff() const {
    Scalar m = coeff(0, 0);
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) m = std::max(m, coeff(i, j));
    return m;
  }
  Scalar minCoeff() const {
    Scalar m = coeff(0, 0);
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) m = std::min(m, coeff(i, j));
    return m;
  }
  template <typename I> Scalar maxCoeff(I* idx) const {
    Scalar m = coeff(0); *idx = 0;
    for (Index i = 1; i < size(); i++) if (coeff(i) > m) { m = coeff(i); *idx = (I)i; }
    return m;
  }
  template <typename I> Scalar minCoeff(I* idx) const {
    Scalar m = coeff(0); *idx = 0;
    for (Index i = 1; i < size(); i++) if (coeff(i) < m) { m = coeff(i); *idx = (I)i; }
    return m;
  }
  template <int P> Scalar lpNorm() const {
    if (P == Infinity) {
      Scalar m = 0;
      for (Index j = 0; j < cols(); j++)
        for (Index i = 0; i < rows(); i++) m = std::max(m, (Scalar)std::fabs(coeff(i, j)));
      return m;
    }
    if (P == 1) {
      Scalar m = 0;
      for (Index j = 0; j < cols(); j++)
        for (Index i = 0; i < rows(); i++) m += std::fabs(coeff(i, j));
      return m;
    }
    return norm();
  }
  template <typename O> Scalar dot(const MatrixBase<O>& o) const {
    Scalar s = 0;
    for (Index i = 0; i < size(); i++) s += coeff(i) * o.coeff(i);
    return s;
  }
  template <typename O> Matrix<Scalar, 3, 1> cross(const MatrixBase<O>& o) const {
    Matrix<Scalar, 3, 1> r;
    r(0) = coeff(1) * o.coeff(2) - coeff(2) * o.coeff(1);
    r(1) = coeff(2) * o.coeff(0) - coeff(0) * o.coeff(2);
    r(2) = coeff(0) * o.coeff(1) - coeff(1) * o.coeff(0);
    return r;
  }
  Plain cwiseAbs() const {
    Plain r = eval();
    for (Index i = 0; i < r.size(); i++) r.data()[i] = std::fabs(r.data()[i]);
    return r;
  }
  Plain cwiseSqrt() const {
    Plain r = eval();
    for (Index i = 0; i < r.size(); i++) r.data()[i] = std::sqrt(r.data()[i]);
    return r;
  }
  Plain cwiseInverse() const {
    Plain r = eval();
    for (Index i = 0; i < r.size(); i++) r.data()[i] = Scalar(1) / r.data()[i];
    return r;
  }
  template <typename O> Plain cwiseProduct(const MatrixBase<O>& o) const {
    Plain r = eval();
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) r(i, j) *= o.coeff(i, j);
    return r;
  }
  template <typename O> Plain cwiseQuotient(const MatrixBase<O>& o) const {
    Plain r = eval();
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) r(i, j) /= o.coeff(i, j);
    return r;
  }
  template <typename T> Matrix<T, RowsAtCompileTime, ColsAtCompileTime> cast() const {
    Matrix<T, RowsAtCompileTime, ColsAtCompileTime> r(rows(), cols());
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) r(i, j) = (T)coeff(i, j);
    return r;
  }
  bool allFinite() const {
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) if (!std::isfinite(coeff(i, j))) return false;
    return true;
  }
  bool isZero(Scalar prec = 1e-12) const { return lpNorm<Infinity>() <= prec; }
  template <typename O> bool isApprox(const MatrixBase<O>& o, Scalar prec = 1e-12) const {
    return ((*this) - o).squaredNorm() <= prec * prec * std::min(squaredNorm(), o.squaredNorm());
  }
  ArrayWrap<Plain> array() const { return ArrayWrap<Plain>(eval()); }
  const D& matrix() const { return derived(); }
  static const int DiagDim = internal::pick<RowsAtCompileTime, ColsAtCompileTime>::value;
  Matrix<Scalar, DiagDim, 1> diagonal() const;
  Matrix<Scalar, Dynamic, Dynamic> asDiagonal() const;
  Plain inverse() const;
  Scalar determinant() const;
  LLT<Plain> llt() const;
  template <int RR, int CC> Matrix<Scalar, Dynamic, Dynamic> replicate() const;
  Matrix<Scalar, Dynamic, Dynamic> replicate(Index rr, Index cc) const;
  Matrix<Scalar, 1, ColsAtCompileTime> colwise_sum() const;

  // ---- read-only sub-blocks (copies) ----
  Matrix<Scalar, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) const;
  template <int BR, int BC> Matrix<Scalar, BR, BC> block(Index i, Index j) const;
  Matrix<Scalar, Dynamic, 1> head(Index n) const;
  Matrix<Scalar, Dynamic, 1> tail(Index n) const;
  Matrix<Scalar, Dynamic, 1> segment(Index i, Index n) const;
  template <int N> Matrix<Scalar, N, 1> head() const;
  template <int N> Matrix<Scalar, N, 1> tail() const;
  template <int N> Matrix<Scalar, N, 1> segment(Index i) const;
  Matrix<Scalar, RowsAtCompileTime, 1> col(Index j) const;
  Matrix<Scalar, 1, ColsAtCompileTime> row(Index i) const;
  Matrix<Scalar, Dynamic, Dynamic> topLeftCorner(Index r, Index c) const { return block(0, 0, r, c); }
  Matrix<Scalar, Dynamic, Dynamic> topRightCorner(Index r, Index c) const { return block(0, cols() - c, r, c); }
  Matrix<Scalar, Dynamic, Dynamic> bottomLeftCorner(Index r, Index c) const { return block(rows() - r, 0, r, c); }
  Matrix<Scalar, Dynamic, Dynamic> bottomRightCorner(Index r, Index c) const { return block(rows() - r, cols() - c, r, c); }
  Matrix<Scalar, Dynamic, Dynamic> topRows(Index r) const { return block(0, 0, r, cols()); }
  Matrix<Scalar, Dynamic, Dynamic> bottomRows(Index r) const { return block(rows() - r, 0, r, cols()); }
  Matrix<Scalar, Dynamic, Dynamic> leftCols(Index c) const { return block(0, 0, rows(), c); }
  Matrix<Scalar, Dynamic, Dynamic> rightCols(Index c) const { return block(0, cols() - c, rows(), c); }
  Matrix<Scalar, Dynamic, Dynamic> middleRows(Index i, Index r) const { return block(i, 0, r, cols()); }
  Matrix<Scalar, Dynamic, Dynamic> middleCols(Index j, Index c) const { return block(0, j, rows(), c); }
  template <unsigned Mode> Matrix<Scalar, Dynamic, Dynamic> triangularView() const;
};

// ----------------------------------------------------------------------------------------------------------------
// storage
// ----------------------------------------------------------------------------------------------------------------
namespace internal {
template <typename S, int R, int C, bool Fixed = (R != Dynamic && C != Dynamic)>
struct Storage;
template <typename S, int R, int C>
struct Storage<S, R, C, true> {
  S v[R * C > 0 ? R * C : 1];
  Storage() { for (int i = 0; i < R * C; i++) v[i] = S(); }
  Index rows() const { return R; }
  Index cols() const { return C; }
  void resize(Index r, Index c) { if (r != R || c != C) fail("resize of a fixed-size matrix"); }
  S* data() { return v; }
  const S* data() const { return v; }
};
template <typename S, int R, int C>
struct Storage<S, R, C, false> {
  typedef typename std::conditional<std::is_same<S, bool>::value, unsigned char, S>::type Cell;   // (std::vector<bool> has no data())
  std::vector<Cell> v;
  Index r_, c_;
  Storage() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C) {}
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  void resize(Index r, Index c) {
    if (R != Dynamic && r != R) fail("bad row count");
    if (C != Dynamic && c != C) fail("bad column count");
    if ((size_t)(r * c) != v.size()) v.assign((size_t)(r * c), Cell());   // same size: coefficients kept (reinterpreted), as Eigen
    r_ = r; c_ = c;
  }
  S* data() { return reinterpret_cast<S*>(v.data()); }
  const S* data() const { return reinterpret_cast<const S*>(v.data()); }
};
}  // namespace internal

// writable interface shared by Matrix and Block
template <typename D>
class WritableBase : public MatrixBase<D> {
 public:
  typedef typename MatrixBase<D>::Scalar Scalar;
  using MatrixBase<D>::derived;
  using MatrixBase<D>::rows;
  using MatrixBase<D>::cols;
  using MatrixBase<D>::size;
  using MatrixBase<D>::operator();
  using MatrixBase<D>::operator[];
  using MatrixBase<D>::x;
  using MatrixBase<D>::y;
  using MatrixBase<D>::z;
  using MatrixBase<D>::w;
  using MatrixBase<D>::block;
  using MatrixBase<D>::head;
  using MatrixBase<D>::tail;
  using MatrixBase<D>::segment;
  using MatrixBase<D>::col;
  using MatrixBase<D>::row;
  using MatrixBase<D>::diagonal;
  using MatrixBase<D>::topLeftCorner;
  using MatrixBase<D>::topRightCorner;
  using MatrixBase<D>::bottomLeftCorner;
  using MatrixBase<D>::bottomRightCorner;
  using MatrixBase<D>::topRows;
  using MatrixBase<D>::bottomRows;
  using MatrixBase<D>::leftCols;
  using MatrixBase<D>::rightCols;
  using MatrixBase<D>::array;
  Scalar& coeffRef(Index i, Index j) { return derived().coeffRef(i, j); }
  Scalar& coeffRef(Index i) { return cols() == 1 ? coeffRef(i, 0) : coeffRef(0, i); }
  Scalar& operator()(Index i, Index j) { return coeffRef(i, j); }
  Scalar& operator()(Index i) { return coeffRef(i); }
  Scalar& operator[](Index i) { return coeffRef(i); }
  Scalar& x() { return coeffRef(0); }
  Scalar& y() { return coeffRef(1); }
  Scalar& z() { return coeffRef(2); }
  Scalar& w() { return coeffRef(3); }
  template <typename O> D& assign_from(const MatrixBase<O>& o) {
    // (the source is evaluated first: eager semantics make aliasing harmless)
    typename MatrixBase<O>::Plain t = o.eval();
    derived().resize_like(t.rows(), t.cols());
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) coeffRef(i, j) = t(i, j);
    return derived();
  }
  template <typename O> D& operator+=(const MatrixBase<O>& o) {
    typename MatrixBase<O>::Plain t = o.eval();
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) coeffRef(i, j) += t(i, j);
    return derived();
  }
  template <typename O> D& operator-=(const MatrixBase<O>& o) {
    typename MatrixBase<O>::Plain t = o.eval();
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) coeffRef(i, j) -= t(i, j);
    return derived();
  }
  template <typename O> D& operator*=(const MatrixBase<O>& o) { return assign_from((*this) * o); }
  D& operator*=(Scalar s) {
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) coeffRef(i, j) *= s;
    return derived();
  }
  D& operator/=(Scalar s) {
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) coeffRef(i, j) /= s;
    return derived();
  }
  D& setZero() { return setConstant(Scalar(0)); }
  D& setOnes() { return setConstant(Scalar(1)); }
  D& setConstant(Scalar s) {
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) coeffRef(i, j) = s;
    return derived();
  }
  D& fill(Scalar s) { return setConstant(s); }
  D& setIdentity() {
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) coeffRef(i, j) = (i == j) ? Scalar(1) : Scalar(0);
    return derived();
  }
  void normalize() {
    Scalar n = this->norm();
    if (n > Scalar(0)) (*this) /= n;
  }
  void transposeInPlace() { assign_from(this->transpose()); }
  CommaInit<D> operator<<(Scalar s);
  template <typename O> CommaInit<D> operator<<(const MatrixBase<O>& o);
  template <typename A> CommaInit<D> operator<<(const ArrayWrap<A>& a) { return (*this) << a.matrix(); }

  // ---- writable views ----
  Block<D, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) { return Block<D, Dynamic, Dynamic>(derived(), i, j, r, c); }
  template <int BR, int BC> Block<D, BR, BC> block(Index i, Index j) { return Block<D, BR, BC>(derived(), i, j, BR, BC); }
  Block<D, Dynamic, 1> head(Index n) { return vec_block<Dynamic>(0, n); }
  Block<D, Dynamic, 1> tail(Index n) { return vec_block<Dynamic>(size() - n, n); }
  Block<D, Dynamic, 1> segment(Index i, Index n) { return vec_block<Dynamic>(i, n); }
  template <int N> Block<D, N, 1> head() { return vec_block<N>(0, N); }
  template <int N> Block<D, N, 1> tail() { return vec_block<N>(size() - N, N); }
  template <int N> Block<D, N, 1> segment(Index i) { return vec_block<N>(i, N); }
  Block<D, MatrixBase<D>::RowsAtCompileTime, 1> col(Index j) { return Block<D, MatrixBase<D>::RowsAtCompileTime, 1>(derived(), 0, j, rows(), 1); }
  Block<D, 1, MatrixBase<D>::ColsAtCompileTime> row(Index i) { return Block<D, 1, MatrixBase<D>::ColsAtCompileTime>(derived(), i, 0, 1, cols()); }
  Block<D, Dynamic, Dynamic> topLeftCorner(Index r, Index c) { return block(0, 0, r, c); }
  Block<D, Dynamic, Dynamic> topRightCorner(Index r, Index c) { return block(0, cols() - c, r, c); }
  Block<D, Dynamic, Dynamic> bottomLeftCorner(Index r, Index c) { return block(rows() - r, 0, r, c); }
  Block<D, Dynamic, Dynamic> bottomRightCorner(Index r, Index c) { return block(rows() - r, cols() - c, r, c); }
  Block<D, Dynamic, Dynamic> topRows(Index r) { return block(0, 0, r, cols()); }
  Block<D, Dynamic, Dynamic> bottomRows(Index r) { return block(rows() - r, 0, r, cols()); }
  Block<D, Dynamic, Dynamic> leftCols(Index c) { return block(0, 0, rows(), c); }
  Block<D, Dynamic, Dynamic> rightCols(Index c) { return block(0, cols() - c, rows(), c); }
  Block<D, Dynamic, 1> diagonal() { return Block<D, Dynamic, 1>(derived(), 0, 0, std::min(rows(), cols()), 1, true); }

 private:
  template <int N> Block<D, N, 1> vec_block(Index i, Index n) {
    // a vector view of a row vector or a column vector (stored as a column view with a stride flag for rows)
    if (cols() == 1) return Block<D, N, 1>(derived(), i, 0, n, 1);
    return Block<D, N, 1>(derived(), 0, i, n, 1, false, true);
  }
};

// ----------------------------------------------------------------------------------------------------------------
// Matrix
// ----------------------------------------------------------------------------------------------------------------
template <typename S, int R, int C, int Opt, int MR, int MC>
class Matrix : public WritableBase<Matrix<S, R, C, Opt, MR, MC>> {
  internal::Storage<S, R, C> st_;
  static const bool kRowMajor = (Opt & RowMajor) != 0 && R != 1 && C != 1;

 public:
  typedef S Scalar;
  typedef WritableBase<Matrix> Base;
  using Base::operator();
  using Base::operator=;

  Matrix() {}
  Matrix(const Matrix& o) = default;
  Matrix& operator=(const Matrix& o) = default;
  // vector of given length / dynamic matrix of given size / fixed 2-vector from two coefficients
  explicit Matrix(Index n) {
    if (R == Dynamic && C == 1) st_.resize(n, 1);
    else if (C == Dynamic && R == 1) st_.resize(1, n);
    else if (R != Dynamic && C != Dynamic && R * C == 1) st_.v[0] = (S)n;
    else if (R == Dynamic && C == Dynamic) st_.resize(n, n);
    else if (n != (Index)(R * C)) internal::fail("bad size constructor");
  }
  explicit Matrix(int n) : Matrix((Index)n) {}
  explicit Matrix(unsigned n) : Matrix((Index)n) {}
  explicit Matrix(unsigned long n) : Matrix((Index)n) {}
  template <typename T0, typename T1, typename std::enable_if<std::is_integral<T0>::value && std::is_integral<T1>::value, int>::type = 0>
  Matrix(T0 r, T1 c) { init2((double)r, (double)c, (Index)r, (Index)c, true); }
  Matrix(double a, double b) { init2(a, b, (Index)a, (Index)b, false); }
  Matrix(float a, float b) { init2(a, b, (Index)a, (Index)b, false); }
  Matrix(S a, S b, S c) { need(3); st_.data()[0] = a; st_.data()[1] = b; st_.data()[2] = c; }
  Matrix(S a, S b, S c, S d) { need(4); st_.data()[0] = a; st_.data()[1] = b; st_.data()[2] = c; st_.data()[3] = d; }
  explicit Matrix(const S* p) { for (Index i = 0; i < this->size(); i++) st_.data()[i] = p[i]; }
  template <typename O> Matrix(const MatrixBase<O>& o) { this->assign_from(o); }
  template <typename O> Matrix& operator=(const MatrixBase<O>& o) { return this->assign_from(o); }
  template <typename A> Matrix(const ArrayWrap<A>& a);
  template <typename A> Matrix& operator=(const ArrayWrap<A>& a);
  Matrix(const Quaternion<S>& q);     // rotation matrix of a quaternion (RotationBase conversion)
  Matrix(const AngleAxis<S>& a);
  Matrix& operator=(const Quaternion<S>& q);

  // a fixed 1x1 matrix converts to its coefficient (inner products written as a^T * b)
  template <int RR = R, int CC = C, typename std::enable_if<RR == 1 && CC == 1, int>::type = 0>
  operator S() const { return st_.data()[0]; }
  Index rows() const { return st_.rows(); }
  Index cols() const { return st_.cols(); }
  S coeff(Index i, Index j) const { return st_.data()[kRowMajor ? i * cols() + j : i + j * rows()]; }
  S& coeffRef(Index i, Index j) { return st_.data()[kRowMajor ? i * cols() + j : i + j * rows()]; }
  using Base::coeff;
  using Base::coeffRef;
  S* data() { return st_.data(); }
  const S* data() const { return st_.data(); }
  void resize(Index r, Index c) { st_.resize(r, c); }
  void resize(Index n) { if (C == 1 || (C == Dynamic && R != 1 && false)) st_.resize(n, 1); else if (R == 1) st_.resize(1, n); else st_.resize(n, 1); }
  void resize_like(Index r, Index c) {
    if (r == rows() && c == cols()) return;
    // assigning a column to a row-vector type (or vice versa) transposes implicitly, as Eigen does for vectors
    if (R == 1 && c == 1 && C == Dynamic) { st_.resize(1, r); return; }
    if (C == 1 && r == 1 && R == Dynamic) { st_.resize(c, 1); return; }
    st_.resize(r, c);
  }
  void conservativeResize(Index r, Index c) {
    Matrix t(*this);
    st_.resize(r, c);
    for (Index j = 0; j < std::min(c, t.cols()); j++)
      for (Index i = 0; i < std::min(r, t.rows()); i++) coeffRef(i, j) = t(i, j);
  }
  void conservativeResize(Index n) { if (C == 1) conservativeResize(n, 1); else conservativeResize(1, n); }
  void swap(Matrix& o) { std::swap(st_, o.st_); }

  static Matrix Zero() { Matrix m; m.setZero(); return m; }
  static Matrix Zero(Index n) { Matrix m(n); m.setZero(); return m; }
  static Matrix Zero(Index r, Index c) { Matrix m(r, c); m.setZero(); return m; }
  static Matrix Ones() { Matrix m; m.setOnes(); return m; }
  static Matrix Ones(Index n) { Matrix m(n); m.setOnes(); return m; }
  static Matrix Ones(Index r, Index c) { Matrix m(r, c); m.setOnes(); return m; }
  static Matrix Constant(S v) { Matrix m; m.setConstant(v); return m; }
  static Matrix Constant(Index n, S v) { Matrix m(n); m.setConstant(v); return m; }
  static Matrix Constant(Index r, Index c, S v) { Matrix m(r, c); m.setConstant(v); return m; }
  static Matrix Identity() { Matrix m; m.setIdentity(); return m; }
  static Matrix Identity(Index r, Index c) { Matrix m(r, c); m.setIdentity(); return m; }
  static Matrix Random() { Matrix m; for (Index i = 0; i < m.size(); i++) m.data()[i] = S(2) * std::rand() / RAND_MAX - S(1); return m; }
  static Matrix Random(Index n) { Matrix m(n); for (Index i = 0; i < m.size(); i++) m.data()[i] = S(2) * std::rand() / RAND_MAX - S(1); return m; }
  static Matrix Random(Index r, Index c) { Matrix m(r, c); for (Index i = 0; i < m.size(); i++) m.data()[i] = S(2) * std::rand() / RAND_MAX - S(1); return m; }
  static Matrix UnitX() { Matrix m; m.setZero(); m(0) = 1; return m; }
  static Matrix UnitY() { Matrix m; m.setZero(); m(1) = 1; return m; }
  static Matrix UnitZ() { Matrix m; m.setZero(); m(2) = 1; return m; }
  Matrix& setZero() { Base::setZero(); return *this; }
  Matrix& setZero(Index n) { resize(n); Base::setZero(); return *this; }
  Matrix& setZero(Index r, Index c) { resize(r, c); Base::setZero(); return *this; }
  Matrix& setIdentity() { Base::setIdentity(); return *this; }
  Matrix& setIdentity(Index r, Index c) { resize(r, c); Base::setIdentity(); return *this; }

 private:
  void need(Index n) {
    if (R == Dynamic && C == 1) st_.resize(n, 1);
    else if (C == Dynamic && R == 1) st_.resize(1, n);
    else if (this->size() != n) internal::fail("coefficient constructor on a matrix of another size");
  }
  void init2(double a, double b, Index r, Index c, bool integral) {
    if (R != Dynamic && C != Dynamic && R * C == 2) { st_.data()[0] = (S)a; st_.data()[1] = (S)b; return; }
    if (!integral) internal::fail("two-coefficient constructor on a matrix that is not a 2-vector");
    st_.resize(r, c);
  }
};

// ----------------------------------------------------------------------------------------------------------------
// Block: writable view into a matrix (or into another view)
// ----------------------------------------------------------------------------------------------------------------
template <typename P, int BR, int BC>
class Block : public WritableBase<Block<P, BR, BC>> {
  P& p_;
  Index i0_, j0_, r_, c_;
  bool diag_, rowvec_;

 public:
  typedef typename internal::traits<Block>::Scalar Scalar;
  typedef WritableBase<Block> Base;
  using Base::operator();
  Block(P& p, Index i0, Index j0, Index r, Index c, bool diag = false, bool rowvec = false)
      : p_(p), i0_(i0), j0_(j0), r_(r), c_(c), diag_(diag), rowvec_(rowvec) {
    if (i0 < 0 || j0 < 0 || r < 0 || c < 0) internal::fail("negative block");
    if (!diag && !rowvec && (i0 + r > p.rows() || j0 + c > p.cols())) internal::fail("block out of range");
    if (rowvec && (j0 + r > p.cols())) internal::fail("segment out of range");
  }
  Block(const Block&) = default;
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Scalar coeff(Index i, Index j) const {
    if (diag_) return p_.coeff(i, i);
    if (rowvec_) return p_.coeff(i0_, j0_ + i);
    return p_.coeff(i0_ + i, j0_ + j);
  }
  Scalar& coeffRef(Index i, Index j) {
    if (diag_) return p_.coeffRef(i, i);
    if (rowvec_) return p_.coeffRef(i0_, j0_ + i);
    return p_.coeffRef(i0_ + i, j0_ + j);
  }
  using Base::coeff;
  using Base::coeffRef;
  void resize_like(Index r, Index c) {
    if (r == r_ && c == c_) return;
    if (r_ * c_ == r * c && (r_ == 1 || c_ == 1) && (r == 1 || c == 1)) return;  // vector <-> row vector
    internal::fail("assignment to a block of another size");
  }
  Block& operator=(const Block& o) { return this->assign_from(o); }
  template <typename O> Block& operator=(const MatrixBase<O>& o) {
    typename MatrixBase<O>::Plain t = o.eval();
    if (t.rows() == r_ && t.cols() == c_) {
      for (Index j = 0; j < c_; j++)
        for (Index i = 0; i < r_; i++) coeffRef(i, j) = t(i, j);
    } else if (t.size() == r_ * c_ && (t.rows() == 1 || t.cols() == 1)) {
      for (Index k = 0; k < t.size(); k++) Base::coeffRef(k) = t(k);
    } else {
      internal::fail("assignment to a block of another size");
    }
    return *this;
  }
  template <typename A> Block& operator=(const ArrayWrap<A>& a) { return (*this) = a.matrix(); }
};

// ----------------------------------------------------------------------------------------------------------------
// comma initialiser
// ----------------------------------------------------------------------------------------------------------------
template <typename M>
class CommaInit {
  M& m_;
  Index row_, col_, blockRows_;

 public:
  typedef typename M::Scalar Scalar;
  CommaInit(M& m) : m_(m), row_(0), col_(0), blockRows_(1) {}
  void put(Scalar s) {
    if (col_ == m_.cols()) { row_ += blockRows_; col_ = 0; blockRows_ = 1; }
    if (row_ >= m_.rows()) internal::fail("too many coefficients in a comma initialiser");
    m_.coeffRef(row_, col_++) = s;
  }
  template <typename O> void put(const MatrixBase<O>& o) {
    if (col_ == m_.cols()) { row_ += blockRows_; col_ = 0; blockRows_ = 1; }
    if (col_ == 0) blockRows_ = o.rows();
    for (Index j = 0; j < o.cols(); j++)
      for (Index i = 0; i < o.rows(); i++) m_.coeffRef(row_ + i, col_ + j) = o.coeff(i, j);
    col_ += o.cols();
  }
  template <typename A> void put(const ArrayWrap<A>& a) { put(a.matrix()); }
  template <typename A> CommaInit& operator,(const ArrayWrap<A>& a) { put(a.matrix()); return *this; }
  CommaInit& operator,(Scalar s) { put(s); return *this; }
  template <typename O> CommaInit& operator,(const MatrixBase<O>& o) { put(o); return *this; }
  M& finished() { return m_; }
};
template <typename D> CommaInit<D> WritableBase<D>::operator<<(Scalar s) { CommaInit<D> c(derived()); c.put(s); return c; }
template <typename D> template <typename O> CommaInit<D> WritableBase<D>::operator<<(const MatrixBase<O>& o) { CommaInit<D> c(derived()); c.put(o); return c; }

// ----------------------------------------------------------------------------------------------------------------
// typedefs
// ----------------------------------------------------------------------------------------------------------------
#define MINI_EIGEN_TYPEDEFS(S, sfx)                         \
  typedef Matrix<S, 2, 2> Matrix2##sfx;                     \
  typedef Matrix<S, 3, 3> Matrix3##sfx;                     \
  typedef Matrix<S, 4, 4> Matrix4##sfx;                     \
  typedef Matrix<S, Dynamic, Dynamic> MatrixX##sfx;         \
  typedef Matrix<S, 2, 1> Vector2##sfx;                     \
  typedef Matrix<S, 3, 1> Vector3##sfx;                     \
  typedef Matrix<S, 4, 1> Vector4##sfx;                     \
  typedef Matrix<S, Dynamic, 1> VectorX##sfx;               \
  typedef Matrix<S, 1, 2> RowVector2##sfx;                  \
  typedef Matrix<S, 1, 3> RowVector3##sfx;                  \
  typedef Matrix<S, 1, 4> RowVector4##sfx;                  \
  typedef Matrix<S, 1, Dynamic> RowVectorX##sfx;
MINI_EIGEN_TYPEDEFS(double, d)
MINI_EIGEN_TYPEDEFS(float, f)
MINI_EIGEN_TYPEDEFS(int, i)
#undef MINI_EIGEN_TYPEDEFS

// ----------------------------------------------------------------------------------------------------------------
// arithmetic (results sized at compile time where both operands are)
// ----------------------------------------------------------------------------------------------------------------
template <typename A, typename B>
Matrix<typename A::Scalar, internal::pick<A::RowsAtCompileTime, B::RowsAtCompileTime>::value, internal::pick<A::ColsAtCompileTime, B::ColsAtCompileTime>::value>
operator+(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  if (a.rows() != b.rows() || a.cols() != b.cols()) internal::fail("operator+: size mismatch");
  Matrix<typename A::Scalar, internal::pick<A::RowsAtCompileTime, B::RowsAtCompileTime>::value, internal::pick<A::ColsAtCompileTime, B::ColsAtCompileTime>::value> r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); j++)
    for (Index i = 0; i < a.rows(); i++) r(i, j) = a.coeff(i, j) + b.coeff(i, j);
  return r;
}
template <typename A, typename B>
Matrix<typename A::Scalar, internal::pick<A::RowsAtCompileTime, B::RowsAtCompileTime>::value, internal::pick<A::ColsAtCompileTime, B::ColsAtCompileTime>::value>
operator-(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  if (a.rows() != b.rows() || a.cols() != b.cols()) internal::fail("operator-: size mismatch");
  Matrix<typename A::Scalar, internal::pick<A::RowsAtCompileTime, B::RowsAtCompileTime>::value, internal::pick<A::ColsAtCompileTime, B::ColsAtCompileTime>::value> r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); j++)
    for (Index i = 0; i < a.rows(); i++) r(i, j) = a.coeff(i, j) - b.coeff(i, j);
  return r;
}
template <typename A>
typename MatrixBase<A>::Plain operator-(const MatrixBase<A>& a) {
  typename MatrixBase<A>::Plain r(a.rows(), a.cols());
  for (Index j = 0; j < a.cols(); j++)
    for (Index i = 0; i < a.rows(); i++) r(i, j) = -a.coeff(i, j);
  return r;
}
template <typename A, typename B>
Matrix<typename A::Scalar, A::RowsAtCompileTime, B::ColsAtCompileTime> operator*(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  if (a.cols() != b.rows()) internal::fail("operator*: inner dimensions differ");
  typename MatrixBase<A>::Plain ea = a.eval();
  typename MatrixBase<B>::Plain eb = b.eval();
  Matrix<typename A::Scalar, A::RowsAtCompileTime, B::ColsAtCompileTime> r(a.rows(), b.cols());
  const Index n = a.rows(), m = b.cols(), k = a.cols();
  for (Index j = 0; j < m; j++)
    for (Index i = 0; i < n; i++) {
      typename A::Scalar s = 0;
      for (Index t = 0; t < k; t++) s += ea(i, t) * eb(t, j);
      r(i, j) = s;
    }
  return r;
}
#define MINI_EIGEN_SCALAR_OPS(T)                                                                                   \
  template <typename A> typename MatrixBase<A>::Plain operator*(const MatrixBase<A>& a, T s) {                     \
    typename MatrixBase<A>::Plain r = a.eval();                                                                    \
    r *= (typename A::Scalar)s;                                                                                    \
    return r;                                                                                                      \
  }                                                                                                                \
  template <typename A> typename MatrixBase<A>::Plain operator*(T s, const MatrixBase<A>& a) { return a * s; }     \
  template <typename A> typename MatrixBase<A>::Plain operator/(const MatrixBase<A>& a, T s) {                     \
    typename MatrixBase<A>::Plain r = a.eval();                                                                    \
    r /= (typename A::Scalar)s;                                                                                    \
    return r;                                                                                                      \
  }
MINI_EIGEN_SCALAR_OPS(double)
MINI_EIGEN_SCALAR_OPS(float)
MINI_EIGEN_SCALAR_OPS(int)
#undef MINI_EIGEN_SCALAR_OPS

template <typename A, typename B> bool operator==(const MatrixBase<A>& a, const MatrixBase<B>& b) {
  if (a.rows() != b.rows() || a.cols() != b.cols()) return false;
  for (Index j = 0; j < a.cols(); j++)
    for (Index i = 0; i < a.rows(); i++) if (a.coeff(i, j) != b.coeff(i, j)) return false;
  return true;
}
template <typename A, typename B> bool operator!=(const MatrixBase<A>& a, const MatrixBase<B>& b) { return !(a == b); }

template <typename D> std::ostream& operator<<(std::ostream& os, const MatrixBase<D>& m) {
  for (Index i = 0; i < m.rows(); i++) {
    for (Index j = 0; j < m.cols(); j++) os << (j ? " " : "") << m.coeff(i, j);
    if (i + 1 < m.rows()) os << "\n";
  }
  return os;
}

// ---- out-of-line members of MatrixBase ----
template <typename D> Matrix<typename MatrixBase<D>::Scalar, Dynamic, Dynamic> MatrixBase<D>::block(Index i, Index j, Index r, Index c) const {
  if (i < 0 || j < 0 || i + r > rows() || j + c > cols()) internal::fail("block out of range");
  Matrix<Scalar, Dynamic, Dynamic> o(r, c);
  for (Index b = 0; b < c; b++)
    for (Index a = 0; a < r; a++) o(a, b) = coeff(i + a, j + b);
  return o;
}
template <typename D> template <int BR, int BC> Matrix<typename MatrixBase<D>::Scalar, BR, BC> MatrixBase<D>::block(Index i, Index j) const {
  if (i < 0 || j < 0 || i + BR > rows() || j + BC > cols()) internal::fail("block out of range");
  Matrix<Scalar, BR, BC> o;
  for (Index b = 0; b < BC; b++)
    for (Index a = 0; a < BR; a++) o(a, b) = coeff(i + a, j + b);
  return o;
}
template <typename D> Matrix<typename MatrixBase<D>::Scalar, Dynamic, 1> MatrixBase<D>::segment(Index i, Index n) const {
  if (i < 0 || i + n > size()) internal::fail("segment out of range");
  Matrix<Scalar, Dynamic, 1> o(n);
  for (Index a = 0; a < n; a++) o(a) = coeff(i + a);
  return o;
}
template <typename D> Matrix<typename MatrixBase<D>::Scalar, Dynamic, 1> MatrixBase<D>::head(Index n) const { return segment(0, n); }
template <typename D> Matrix<typename MatrixBase<D>::Scalar, Dynamic, 1> MatrixBase<D>::tail(Index n) const { return segment(size() - n, n); }
template <typename D> template <int N> Matrix<typename MatrixBase<D>::Scalar, N, 1> MatrixBase<D>::segment(Index i) const {
  if (i < 0 || i + N > size()) internal::fail("segment out of range");
  Matrix<Scalar, N, 1> o;
  for (Index a = 0; a < N; a++) o(a) = coeff(i + a);
  return o;
}
template <typename D> template <int N> Matrix<typename MatrixBase<D>::Scalar, N, 1> MatrixBase<D>::head() const { return segment<N>(0); }
template <typename D> template <int N> Matrix<typename MatrixBase<D>::Scalar, N, 1> MatrixBase<D>::tail() const { return segment<N>(size() - N); }
template <typename D> Matrix<typename MatrixBase<D>::Scalar, MatrixBase<D>::RowsAtCompileTime, 1> MatrixBase<D>::col(Index j) const {
  Matrix<Scalar, RowsAtCompileTime, 1> o(rows());
  for (Index a = 0; a < rows(); a++) o(a) = coeff(a, j);
  return o;
}
template <typename D> Matrix<typename MatrixBase<D>::Scalar, 1, MatrixBase<D>::ColsAtCompileTime> MatrixBase<D>::row(Index i) const {
  Matrix<Scalar, 1, ColsAtCompileTime> o(cols());
  for (Index a = 0; a < cols(); a++) o(a) = coeff(i, a);
  return o;
}
template <typename D> Matrix<typename MatrixBase<D>::Scalar, MatrixBase<D>::DiagDim, 1> MatrixBase<D>::diagonal() const {
  Matrix<Scalar, DiagDim, 1> o(std::min(rows(), cols()));
  for (Index a = 0; a < o.size(); a++) o(a) = coeff(a, a);
  return o;
}
template <typename D> Matrix<typename MatrixBase<D>::Scalar, Dynamic, Dynamic> MatrixBase<D>::asDiagonal() const {
  Matrix<Scalar, Dynamic, Dynamic> o(size(), size());
  for (Index a = 0; a < size(); a++) o(a, a) = coeff(a);
  return o;
}
template <typename D> template <int RR, int CC> Matrix<typename MatrixBase<D>::Scalar, Dynamic, Dynamic> MatrixBase<D>::replicate() const { return replicate(RR, CC); }
template <typename D> Matrix<typename MatrixBase<D>::Scalar, Dynamic, Dynamic> MatrixBase<D>::replicate(Index rr, Index cc) const {
  Matrix<Scalar, Dynamic, Dynamic> o(rows() * rr, cols() * cc);
  for (Index j = 0; j < o.cols(); j++)
    for (Index i = 0; i < o.rows(); i++) o(i, j) = coeff(i % rows(), j % cols());
  return o;
}
template <typename D> template <unsigned Mode> Matrix<typename MatrixBase<D>::Scalar, Dynamic, Dynamic> MatrixBase<D>::triangularView() const {
  Matrix<Scalar, Dynamic, Dynamic> o(rows(), cols());
  for (Index j = 0; j < cols(); j++)
    for (Index i = 0; i < rows(); i++) o(i, j) = ((Mode == Upper && i <= j) || (Mode == Lower && i >= j)) ? coeff(i, j) : Scalar(0);
  return o;
}
// inverse by Gauss-Jordan elimination with partial pivoting
template <typename D> typename MatrixBase<D>::Plain MatrixBase<D>::inverse() const {
  const Index n = rows();
  if (n != cols()) internal::fail("inverse of a non-square matrix");
  Matrix<Scalar, Dynamic, Dynamic> a = eval(), b = Matrix<Scalar, Dynamic, Dynamic>::Identity(n, n);
  for (Index k = 0; k < n; k++) {
    Index p = k;
    for (Index i = k + 1; i < n; i++) if (std::fabs(a(i, k)) > std::fabs(a(p, k))) p = i;
    if (p != k) for (Index j = 0; j < n; j++) { std::swap(a(k, j), a(p, j)); std::swap(b(k, j), b(p, j)); }
    const Scalar piv = Scalar(1) / a(k, k);
    for (Index j = 0; j < n; j++) { a(k, j) *= piv; b(k, j) *= piv; }
    for (Index i = 0; i < n; i++) {
      if (i == k) continue;
      const Scalar f = a(i, k);
      if (f == Scalar(0)) continue;
      for (Index j = 0; j < n; j++) { a(i, j) -= f * a(k, j); b(i, j) -= f * b(k, j); }
    }
  }
  Plain r(n, n);
  for (Index j = 0; j < n; j++)
    for (Index i = 0; i < n; i++) r(i, j) = b(i, j);
  return r;
}
template <typename D> typename MatrixBase<D>::Scalar MatrixBase<D>::determinant() const {
  const Index n = rows();
  Matrix<Scalar, Dynamic, Dynamic> a = eval();
  Scalar det = 1;
  for (Index k = 0; k < n; k++) {
    Index p = k;
    for (Index i = k + 1; i < n; i++) if (std::fabs(a(i, k)) > std::fabs(a(p, k))) p = i;
    if (a(p, k) == Scalar(0)) return 0;
    if (p != k) { for (Index j = 0; j < n; j++) std::swap(a(k, j), a(p, j)); det = -det; }
    det *= a(k, k);
    for (Index i = k + 1; i < n; i++) {
      const Scalar f = a(i, k) / a(k, k);
      for (Index j = k; j < n; j++) a(i, j) -= f * a(k, j);
    }
  }
  return det;
}

// ---- LLT (Cholesky, lower) ----
template <typename M>
class LLT {
  Matrix<typename M::Scalar, Dynamic, Dynamic> L_;
  bool ok_;

 public:
  typedef typename M::Scalar Scalar;
  template <typename D> explicit LLT(const MatrixBase<D>& a) : ok_(true) {
    const Index n = a.rows();
    L_ = Matrix<Scalar, Dynamic, Dynamic>::Zero(n, n);
    for (Index j = 0; j < n; j++) {
      Scalar d = a.coeff(j, j);
      for (Index k = 0; k < j; k++) d -= L_(j, k) * L_(j, k);
      if (!(d > Scalar(0))) { ok_ = false; d = std::fabs(d); }
      const Scalar l = std::sqrt(d);
      L_(j, j) = l;
      for (Index i = j + 1; i < n; i++) {
        Scalar s = a.coeff(i, j);
        for (Index k = 0; k < j; k++) s -= L_(i, k) * L_(j, k);
        L_(i, j) = s / l;
      }
    }
  }
  Matrix<Scalar, Dynamic, Dynamic> matrixL() const { return L_; }
  Matrix<Scalar, Dynamic, Dynamic> matrixU() const { return L_.transpose(); }
  ComputationInfo info() const { return ok_ ? Success : NumericalIssue; }
  template <typename D> Matrix<Scalar, Dynamic, Dynamic> solve(const MatrixBase<D>& b) const {
    const Index n = L_.rows();
    Matrix<Scalar, Dynamic, Dynamic> x = b.eval();
    for (Index c = 0; c < x.cols(); c++) {
      for (Index i = 0; i < n; i++) { Scalar s = x(i, c); for (Index k = 0; k < i; k++) s -= L_(i, k) * x(k, c); x(i, c) = s / L_(i, i); }
      for (Index i = n - 1; i >= 0; i--) { Scalar s = x(i, c); for (Index k = i + 1; k < n; k++) s -= L_(k, i) * x(k, c); x(i, c) = s / L_(i, i); }
    }
    return x;
  }
};
template <typename D> LLT<typename MatrixBase<D>::Plain> MatrixBase<D>::llt() const { return LLT<Plain>(*this); }

// ---- decompositions that the reference only uses off the hot path (covariance recovery / GLC): declared so that the
// headers parse, aborting if ever executed ----
template <typename M>
class JacobiSVD {
 public:
  typedef typename M::Scalar Scalar;
  JacobiSVD() {}
  template <typename D> JacobiSVD(const MatrixBase<D>&, unsigned = 0) { internal::fail("JacobiSVD is not provided by the shim (off the hot path)"); }
  Matrix<Scalar, Dynamic, Dynamic> matrixU() const { return Matrix<Scalar, Dynamic, Dynamic>(); }
  Matrix<Scalar, Dynamic, Dynamic> matrixV() const { return Matrix<Scalar, Dynamic, Dynamic>(); }
  Matrix<Scalar, Dynamic, 1> singularValues() const { return Matrix<Scalar, Dynamic, 1>(); }
  template <typename D> Matrix<Scalar, Dynamic, Dynamic> solve(const MatrixBase<D>&) const { return Matrix<Scalar, Dynamic, Dynamic>(); }
};
template <typename M>
class SelfAdjointEigenSolver {
 public:
  typedef typename M::Scalar Scalar;
  SelfAdjointEigenSolver() {}
  template <typename D> SelfAdjointEigenSolver(const MatrixBase<D>&, int = 0) { internal::fail("SelfAdjointEigenSolver is not provided by the shim (off the hot path)"); }
  Matrix<Scalar, Dynamic, Dynamic> eigenvectors() const { return Matrix<Scalar, Dynamic, Dynamic>(); }
  Matrix<Scalar, Dynamic, 1> eigenvalues() const { return Matrix<Scalar, Dynamic, 1>(); }
  ComputationInfo info() const { return Success; }
};

// ----------------------------------------------------------------------------------------------------------------
// coefficient-wise wrapper (.array())
// ----------------------------------------------------------------------------------------------------------------
template <typename M>
class ArrayWrap {
  M m_;

 public:
  typedef typename M::Scalar Scalar;
  explicit ArrayWrap(const M& m) : m_(m) {}
  const M& matrix() const { return m_; }
  Index rows() const { return m_.rows(); }
  Index cols() const { return m_.cols(); }
  Index size() const { return m_.size(); }
  Scalar operator()(Index i) const { return m_(i); }
  Scalar operator()(Index i, Index j) const { return m_(i, j); }
  template <typename F> ArrayWrap map(F f) const {
    M r = m_;
    for (Index i = 0; i < r.size(); i++) r.data()[i] = f(r.data()[i]);
    return ArrayWrap(r);
  }
  template <typename O, typename F> ArrayWrap zip(const ArrayWrap<O>& o, F f) const {
    if (o.rows() != rows() || o.cols() != cols()) internal::fail("array operation: size mismatch");
    M r = m_;
    for (Index j = 0; j < cols(); j++)
      for (Index i = 0; i < rows(); i++) r(i, j) = f(m_(i, j), o(i, j));
    return ArrayWrap(r);
  }
  ArrayWrap abs() const { return map([](Scalar v) { return (Scalar)std::fabs(v); }); }
  ArrayWrap sqrt() const { return map([](Scalar v) { return (Scalar)std::sqrt(v); }); }
  ArrayWrap square() const { return map([](Scalar v) { return v * v; }); }
  ArrayWrap inverse() const { return map([](Scalar v) { return Scalar(1) / v; }); }
  Scalar sum() const { return m_.sum(); }
  Scalar maxCoeff() const { return m_.maxCoeff(); }
  Scalar minCoeff() const { return m_.minCoeff(); }
  template <typename O> ArrayWrap operator*(const ArrayWrap<O>& o) const { return zip(o, [](Scalar a, Scalar b) { return a * b; }); }
  template <typename O> ArrayWrap operator/(const ArrayWrap<O>& o) const { return zip(o, [](Scalar a, Scalar b) { return a / b; }); }
  template <typename O> ArrayWrap operator+(const ArrayWrap<O>& o) const { return zip(o, [](Scalar a, Scalar b) { return a + b; }); }
  template <typename O> ArrayWrap operator-(const ArrayWrap<O>& o) const { return zip(o, [](Scalar a, Scalar b) { return a - b; }); }
  ArrayWrap operator*(Scalar s) const { return map([s](Scalar v) { return v * s; }); }
  ArrayWrap operator/(Scalar s) const { return map([s](Scalar v) { return v / s; }); }
  ArrayWrap operator+(Scalar s) const { return map([s](Scalar v) { return v + s; }); }
  ArrayWrap operator-(Scalar s) const { return map([s](Scalar v) { return v - s; }); }
  ArrayWrap operator-() const { return map([](Scalar v) { return -v; }); }
  friend ArrayWrap operator*(Scalar s, const ArrayWrap& a) { return a * s; }
  friend ArrayWrap operator/(Scalar s, const ArrayWrap& a) { return a.map([s](Scalar v) { return s / v; }); }
  friend ArrayWrap operator+(Scalar s, const ArrayWrap& a) { return a + s; }
  friend ArrayWrap operator-(Scalar s, const ArrayWrap& a) { return a.map([s](Scalar v) { return s - v; }); }
};
template <typename S, int R, int C, int O, int MR, int MC> template <typename A>
Matrix<S, R, C, O, MR, MC>::Matrix(const ArrayWrap<A>& a) { this->assign_from(a.matrix()); }
template <typename S, int R, int C, int O, int MR, int MC> template <typename A>
Matrix<S, R, C, O, MR, MC>& Matrix<S, R, C, O, MR, MC>::operator=(const ArrayWrap<A>& a) { return this->assign_from(a.matrix()); }

// ----------------------------------------------------------------------------------------------------------------
// geometry: Quaternion, AngleAxis, Rotation2D, Isometry3d (the last two only so that off-path headers parse)
// ----------------------------------------------------------------------------------------------------------------
template <typename S>
class Quaternion {
  Matrix<S, 4, 1> c_;  // x, y, z, w

 public:
  typedef S Scalar;
  Quaternion() {}
  Quaternion(S w, S x, S y, S z) { c_(0) = x; c_(1) = y; c_(2) = z; c_(3) = w; }
  explicit Quaternion(const S* p) { for (int i = 0; i < 4; i++) c_(i) = p[i]; }
  // from the coefficient vector (x, y, z, w)
  template <typename D, typename std::enable_if<internal::traits<D>::Cols == 1 || internal::traits<D>::Rows == Dynamic, int>::type = 0>
  explicit Quaternion(const MatrixBase<D>& v) {
    if (v.rows() == 3 && v.cols() == 3) { from_matrix(v); return; }
    if (v.size() != 4) internal::fail("Quaternion from a vector that has not 4 coefficients");
    for (int i = 0; i < 4; i++) c_(i) = v.coeff(i);
  }
  // from a 3x3 rotation matrix
  template <typename D, typename std::enable_if<internal::traits<D>::Cols == 3 && internal::traits<D>::Rows == 3, int>::type = 0>
  explicit Quaternion(const MatrixBase<D>& m) { from_matrix(m); }
  explicit Quaternion(const AngleAxis<S>& aa);
  Quaternion& operator=(const AngleAxis<S>& aa) { *this = Quaternion(aa); return *this; }
  template <typename D> Quaternion& operator=(const MatrixBase<D>& m) { from_matrix(m); return *this; }

  S x() const { return c_(0); }
  S y() const { return c_(1); }
  S z() const { return c_(2); }
  S w() const { return c_(3); }
  S& x() { return c_(0); }
  S& y() { return c_(1); }
  S& z() { return c_(2); }
  S& w() { return c_(3); }
  const Matrix<S, 4, 1>& coeffs() const { return c_; }
  Matrix<S, 4, 1>& coeffs() { return c_; }
  Matrix<S, 3, 1> vec() const { return Matrix<S, 3, 1>(c_(0), c_(1), c_(2)); }
  static Quaternion Identity() { return Quaternion(1, 0, 0, 0); }
  Quaternion& setIdentity() { *this = Identity(); return *this; }
  S squaredNorm() const { return c_.squaredNorm(); }
  S norm() const { return c_.norm(); }
  void normalize() { c_.normalize(); }
  Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
  Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
  Quaternion inverse() const {
    const S n2 = squaredNorm();
    if (n2 > S(0)) return Quaternion(w() / n2, -x() / n2, -y() / n2, -z() / n2);
    return Quaternion(0, 0, 0, 0);
  }
  S dot(const Quaternion& o) const { return c_.dot(o.c_); }
  Quaternion operator*(const Quaternion& b) const {   // Hamilton product
    const Quaternion& a = *this;
    return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(),
                      a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                      a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(),
                      a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
  }
  Quaternion& operator*=(const Quaternion& b) { *this = (*this) * b; return *this; }
  Matrix<S, 3, 3> toRotationMatrix() const {
    Matrix<S, 3, 3> R;
    const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
    const S twx = tx * w(), twy = ty * w(), twz = tz * w();
    const S txx = tx * x(), txy = ty * x(), txz = tz * x();
    const S tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
    R(0, 0) = S(1) - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = S(1) - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = S(1) - (txx + tyy);
    return R;
  }
  Matrix<S, 3, 3> matrix() const { return toRotationMatrix(); }
  template <typename D> Matrix<S, 3, 1> operator*(const MatrixBase<D>& v) const { return toRotationMatrix() * v; }
  template <typename D> Matrix<S, 3, 1> _transformVector(const MatrixBase<D>& v) const { return toRotationMatrix() * v; }
  template <typename T> Quaternion<T> cast() const { return Quaternion<T>((T)w(), (T)x(), (T)y(), (T)z()); }

 private:
  // rotation matrix -> quaternion: the branch on the trace / largest diagonal element of Shepperd's method, as published
  template <typename D> void from_matrix(const MatrixBase<D>& m) {
    if (m.rows() != 3 || m.cols() != 3) internal::fail("Quaternion from a matrix that is not 3x3");
    S t = m.coeff(0, 0) + m.coeff(1, 1) + m.coeff(2, 2);
    if (t > S(0)) {
      t = std::sqrt(t + S(1));
      w() = S(0.5) * t;
      t = S(0.5) / t;
      x() = (m.coeff(2, 1) - m.coeff(1, 2)) * t;
      y() = (m.coeff(0, 2) - m.coeff(2, 0)) * t;
      z() = (m.coeff(1, 0) - m.coeff(0, 1)) * t;
    } else {
      int i = 0;
      if (m.coeff(1, 1) > m.coeff(0, 0)) i = 1;
      if (m.coeff(2, 2) > m.coeff(i, i)) i = 2;
      const int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(m.coeff(i, i) - m.coeff(j, j) - m.coeff(k, k) + S(1));
      c_(i) = S(0.5) * t;
      t = S(0.5) / t;
      w() = (m.coeff(k, j) - m.coeff(j, k)) * t;
      c_(j) = (m.coeff(j, i) + m.coeff(i, j)) * t;
      c_(k) = (m.coeff(k, i) + m.coeff(i, k)) * t;
    }
  }
};
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;

template <typename S>
class AngleAxis {
  Matrix<S, 3, 1> axis_;
  S angle_;

 public:
  AngleAxis() : angle_(0) { axis_(0) = 1; }
  template <typename D> AngleAxis(S angle, const MatrixBase<D>& axis) : axis_(axis), angle_(angle) {}
  // quaternion -> angle / axis in the atan2 form of Eigen >= 3.3 (angle in [0, pi], axis flipped when w < 0); Eigen 3.2's
  // acos form gives an angle in [0, 2 pi] and the reference's own wrap (isam_plane3d.h:291-292) maps both to the same vector
  explicit AngleAxis(const Quaternion<S>& q) {
    S n = q.vec().norm();
    if (n < std::numeric_limits<S>::epsilon()) n = q.vec().template lpNorm<Infinity>() > S(0) ? std::sqrt(q.vec().squaredNorm()) : S(0);
    if (n != S(0)) {
      angle_ = S(2) * std::atan2(n, std::fabs(q.w()));
      if (q.w() < S(0)) n = -n;
      axis_ = q.vec() / n;
    } else {
      angle_ = S(0);
      axis_ = Matrix<S, 3, 1>(1, 0, 0);
    }
  }
  template <typename D> explicit AngleAxis(const MatrixBase<D>& m) { *this = AngleAxis(Quaternion<S>(m)); }
  S angle() const { return angle_; }
  S& angle() { return angle_; }
  const Matrix<S, 3, 1>& axis() const { return axis_; }
  Matrix<S, 3, 1>& axis() { return axis_; }
  Matrix<S, 3, 3> toRotationMatrix() const {
    Matrix<S, 3, 3> R;
    const S s = std::sin(angle_), c = std::cos(angle_);
    const Matrix<S, 3, 1> cu = axis_ * (S(1) - c), su = axis_ * s;
    S t;
    t = cu.x() * axis_.y(); R(0, 1) = t - su.z(); R(1, 0) = t + su.z();
    t = cu.x() * axis_.z(); R(0, 2) = t + su.y(); R(2, 0) = t - su.y();
    t = cu.y() * axis_.z(); R(1, 2) = t - su.x(); R(2, 1) = t + su.x();
    R(0, 0) = cu.x() * axis_.x() + c; R(1, 1) = cu.y() * axis_.y() + c; R(2, 2) = cu.z() * axis_.z() + c;
    return R;
  }
  Matrix<S, 3, 3> matrix() const { return toRotationMatrix(); }
  AngleAxis inverse() const { return AngleAxis(-angle_, axis_); }
  Quaternion<S> operator*(const AngleAxis& o) const { return Quaternion<S>(*this) * Quaternion<S>(o); }
  Quaternion<S> operator*(const Quaternion<S>& o) const { return Quaternion<S>(*this) * o; }
  template <typename D> Matrix<S, 3, Dynamic> operator*(const MatrixBase<D>& v) const { return toRotationMatrix() * v; }
};
typedef AngleAxis<double> AngleAxisd;
typedef AngleAxis<float> AngleAxisf;

template <typename S> Quaternion<S>::Quaternion(const AngleAxis<S>& aa) {
  const S ha = S(0.5) * aa.angle();
  const S s = std::sin(ha);
  w() = std::cos(ha); x() = s * aa.axis().x(); y() = s * aa.axis().y(); z() = s * aa.axis().z();
}
template <typename S, int R, int C, int O, int MR, int MC> Matrix<S, R, C, O, MR, MC>::Matrix(const Quaternion<S>& q) { this->assign_from(q.toRotationMatrix()); }
template <typename S, int R, int C, int O, int MR, int MC> Matrix<S, R, C, O, MR, MC>::Matrix(const AngleAxis<S>& a) { this->assign_from(a.toRotationMatrix()); }
template <typename S, int R, int C, int O, int MR, int MC> Matrix<S, R, C, O, MR, MC>& Matrix<S, R, C, O, MR, MC>::operator=(const Quaternion<S>& q) { return this->assign_from(q.toRotationMatrix()); }

template <typename S>
class Rotation2D {
  S a_;

 public:
  explicit Rotation2D(S a = 0) : a_(a) {}
  S angle() const { return a_; }
  S& angle() { return a_; }
  Matrix<S, 2, 2> toRotationMatrix() const {
    Matrix<S, 2, 2> R;
    R(0, 0) = std::cos(a_); R(0, 1) = -std::sin(a_); R(1, 0) = std::sin(a_); R(1, 1) = std::cos(a_);
    return R;
  }
  Matrix<S, 2, 2> matrix() const { return toRotationMatrix(); }
  Rotation2D inverse() const { return Rotation2D(-a_); }
  template <typename D> Matrix<S, 2, 1> operator*(const MatrixBase<D>& v) const { return toRotationMatrix() * v; }
  Rotation2D operator*(const Rotation2D& o) const { return Rotation2D(a_ + o.a_); }
};
typedef Rotation2D<double> Rotation2Dd;
typedef Rotation2D<float> Rotation2Df;

// 4x4 rigid transform (only what off-path headers need to parse / the simplest operations)
template <typename S>
class Isometry3 {
  Matrix<S, 4, 4> m_;

 public:
  Isometry3() { m_.setIdentity(); }
  template <typename D> explicit Isometry3(const MatrixBase<D>& m) : m_(m) {}
  static Isometry3 Identity() { return Isometry3(); }
  const Matrix<S, 4, 4>& matrix() const { return m_; }
  Matrix<S, 4, 4>& matrix() { return m_; }
  Matrix<S, 3, 3> rotation() const { return m_.template block<3, 3>(0, 0); }
  Matrix<S, 3, 3> linear() const { return m_.template block<3, 3>(0, 0); }
  Matrix<S, 3, 1> translation() const { return m_.template block<3, 1>(0, 3); }
  Isometry3 inverse() const { return Isometry3(m_.inverse()); }
  Isometry3 operator*(const Isometry3& o) const { return Isometry3(m_ * o.m_); }
  template <typename D> Matrix<S, 3, 1> operator*(const MatrixBase<D>& v) const { return rotation() * v + translation(); }
  S operator()(Index i, Index j) const { return m_(i, j); }
  S& operator()(Index i, Index j) { return m_(i, j); }
};
typedef Isometry3<double> Isometry3d;
typedef Isometry3<float> Isometry3f;

}  // namespace Eigen
