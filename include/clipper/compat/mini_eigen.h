/*
 * mini_eigen.h -- the few Eigen names the CLIPPER public API is written in, for builds where
 * Eigen3 is not installed (this repository's build box has no Eigen and no network).
 *
 * When <Eigen/Dense> IS available, clipper/types.h includes the real Eigen and this file is not
 * used; the C++ shell (clipper_shell.cpp) only touches the subset of the Eigen interface that
 * exists identically here: column-major storage, data(), rows(), cols(), size(), operator(),
 * (rows,cols) construction, Zero/Ones/Identity, SparseMatrix::{outerIndexPtr,innerIndexPtr,
 * valuePtr,nonZeros,setFromTriplets,makeCompressed}.  It is a type vocabulary, not a linear
 * algebra library: no expression templates, no products.
 */
#pragma once
#include <algorithm>
#include <cassert>
#include <cstddef>
#include <initializer_list>
#include <vector>

namespace Eigen {

constexpr int Dynamic = -1;
using Index = std::ptrdiff_t;

template <typename M> class CommaInit {
 public:
  CommaInit(M& m, typename M::Scalar first) : m_(m), k_(0) { put(first); }
  CommaInit& operator,(typename M::Scalar v) { put(v); return *this; }
 private:
  void put(typename M::Scalar v) {  // row-major fill order, like Eigen's comma initialiser
    const Index r = k_ / m_.cols(), c = k_ % m_.cols();
    m_(r, c) = v; ++k_;
  }
  M& m_;
  Index k_;
};

template <typename T, int R, int C>
class Matrix {
 public:
  using Scalar = T;
  Matrix() : rows_(R == Dynamic ? 0 : R), cols_(C == Dynamic ? 0 : C), d_((size_t)rows_ * cols_) {}
  explicit Matrix(Index n) : rows_(C == 1 ? n : (R == Dynamic ? n : R)), cols_(C == 1 ? 1 : (R == 1 ? n : (C == Dynamic ? 0 : C))), d_((size_t)rows_ * cols_) {}
  Matrix(Index r, Index c) : rows_(r), cols_(c), d_((size_t)r * c) {}
  template <int R2, int C2>
  Matrix(const Matrix<T, R2, C2>& o) : rows_(o.rows()), cols_(o.cols()), d_(o.data(), o.data() + o.size()) {}
  template <int R2, int C2>
  Matrix& operator=(const Matrix<T, R2, C2>& o) { rows_ = o.rows(); cols_ = o.cols(); d_.assign(o.data(), o.data() + o.size()); return *this; }

  Index rows() const { return rows_; }
  Index cols() const { return cols_; }
  Index size() const { return rows_ * cols_; }
  T* data() { return d_.data(); }
  const T* data() const { return d_.data(); }
  T& operator()(Index i, Index j) { return d_[(size_t)(i + j * rows_)]; }
  const T& operator()(Index i, Index j) const { return d_[(size_t)(i + j * rows_)]; }
  T& operator()(Index i) { return d_[(size_t)i]; }
  const T& operator()(Index i) const { return d_[(size_t)i]; }
  T& operator[](Index i) { return d_[(size_t)i]; }
  const T& operator[](Index i) const { return d_[(size_t)i]; }

  void resize(Index r, Index c) { rows_ = r; cols_ = c; d_.assign((size_t)r * c, T()); }
  void resize(Index n) { if (C == 1) resize(n, 1); else resize(1, n); }
  void conservativeResize(Index r, Index c) {
    Matrix t(r, c);
    for (Index j = 0; j < std::min(c, cols_); ++j)
      for (Index i = 0; i < std::min(r, rows_); ++i) t(i, j) = (*this)(i, j);
    *this = t;
  }
  void setZero() { std::fill(d_.begin(), d_.end(), T()); }
  void setOnes() { std::fill(d_.begin(), d_.end(), T(1)); }

  static Matrix Zero(Index r, Index c) { Matrix m(r, c); return m; }
  static Matrix Zero(Index n) { Matrix m(n); return m; }
  static Matrix Ones(Index r, Index c) { Matrix m(r, c); m.setOnes(); return m; }
  static Matrix Ones(Index n) { Matrix m(n); m.setOnes(); return m; }
  static Matrix Identity(Index r, Index c) { Matrix m(r, c); for (Index i = 0; i < std::min(r, c); ++i) m(i, i) = T(1); return m; }

  Matrix<T, Dynamic, 1> col(Index j) const {
    Matrix<T, Dynamic, 1> v(rows_);
    for (Index i = 0; i < rows_; ++i) v(i) = (*this)(i, j);
    return v;
  }
  void setCol(Index j, std::initializer_list<T> vals) { Index i = 0; for (T v : vals) (*this)(i++, j) = v; }
  Matrix<T, 1, Dynamic> row(Index i) const {
    Matrix<T, 1, Dynamic> v(1, cols_);
    for (Index j = 0; j < cols_; ++j) v(0, j) = (*this)(i, j);
    return v;
  }
  Matrix<T, Dynamic, Dynamic> transpose() const {
    Matrix<T, Dynamic, Dynamic> t(cols_, rows_);
    for (Index j = 0; j < cols_; ++j) for (Index i = 0; i < rows_; ++i) t(j, i) = (*this)(i, j);
    return t;
  }
  Matrix<T, Dynamic, 1> diagonal() const {
    Matrix<T, Dynamic, 1> v(std::min(rows_, cols_));
    for (Index i = 0; i < v.size(); ++i) v(i) = (*this)(i, i);
    return v;
  }
  T sum() const { T s = T(); for (const T& v : d_) s += v; return s; }
  CommaInit<Matrix> operator<<(T first) { return CommaInit<Matrix>(*this, first); }

  template <int R2, int C2>
  bool operator==(const Matrix<T, R2, C2>& o) const {
    return rows_ == o.rows() && cols_ == o.cols() && std::equal(d_.begin(), d_.end(), o.data());
  }
  template <int R2, int C2> bool operator!=(const Matrix<T, R2, C2>& o) const { return !(*this == o); }

 private:
  Index rows_, cols_;
  std::vector<T> d_;
};

using MatrixXd = Matrix<double, Dynamic, Dynamic>;
using MatrixXi = Matrix<int, Dynamic, Dynamic>;
using Matrix3Xd = Matrix<double, 3, Dynamic>;
using VectorXd = Matrix<double, Dynamic, 1>;
using VectorXi = Matrix<int, Dynamic, 1>;
using Vector3d = Matrix<double, 3, 1>;
using RowVectorXd = Matrix<double, 1, Dynamic>;

template <typename T>
class Triplet {
 public:
  Triplet() : r_(0), c_(0), v_(T()) {}
  Triplet(Index r, Index c, const T& v) : r_(r), c_(c), v_(v) {}
  Index row() const { return r_; }
  Index col() const { return c_; }
  const T& value() const { return v_; }
 private:
  Index r_, c_;
  T v_;
};

/* compressed column storage, StorageIndex = int (Eigen's default) */
template <typename T>
class SparseMatrix {
 public:
  using Scalar = T;
  using StorageIndex = int;
  SparseMatrix() : rows_(0), cols_(0), outer_(1, 0) {}
  SparseMatrix(Index r, Index c) : rows_(r), cols_(c), outer_((size_t)c + 1, 0) {}
  Index rows() const { return rows_; }
  Index cols() const { return cols_; }
  Index nonZeros() const { return (Index)val_.size(); }
  void resize(Index r, Index c) { rows_ = r; cols_ = c; outer_.assign((size_t)c + 1, 0); inner_.clear(); val_.clear(); }
  template <typename It>
  void setFromTriplets(It first, It last) {  // duplicates are summed, like Eigen
    std::vector<Triplet<T>> t(first, last);
    std::sort(t.begin(), t.end(), [](const Triplet<T>& a, const Triplet<T>& b) {
      return a.col() != b.col() ? a.col() < b.col() : a.row() < b.row(); });
    outer_.assign((size_t)cols_ + 1, 0); inner_.clear(); val_.clear();
    for (size_t k = 0; k < t.size(); ++k) {
      if (k > 0 && t[k].col() == t[k - 1].col() && t[k].row() == t[k - 1].row()) { val_.back() += t[k].value(); continue; }
      inner_.push_back((int)t[k].row()); val_.push_back(t[k].value()); ++outer_[(size_t)t[k].col() + 1];
    }
    for (Index j = 0; j < cols_; ++j) outer_[(size_t)j + 1] += outer_[(size_t)j];
  }
  void makeCompressed() {}
  bool isCompressed() const { return true; }
  const int* outerIndexPtr() const { return outer_.data(); }
  const int* innerIndexPtr() const { return inner_.data(); }
  const T* valuePtr() const { return val_.data(); }
  T coeff(Index i, Index j) const {
    for (int q = outer_[(size_t)j]; q < outer_[(size_t)j + 1]; ++q) if (inner_[(size_t)q] == i) return val_[(size_t)q];
    return T();
  }
 private:
  Index rows_, cols_;
  std::vector<int> outer_, inner_;
  std::vector<T> val_;
};

}  // namespace Eigen
