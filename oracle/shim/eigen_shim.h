// eigen_shim.h — a minimal, self-written stand-in for the part of Eigen 3 that the reference's EKF,
// transform and laser-detector translation units use.  TEST INFRASTRUCTURE (oracle/): it exists so that
// /root/reference/src/reflector_ekf_slam/reflector_ekf_slam{,_gps}.cc, src/reflector_detect/laser/*.cc
// and their headers compile UNMODIFIED here (Eigen itself is not installed and there is no network).
// Nothing under reflector_ekf_slam_b200/ includes it.
//
// What is modelled, because the reference's results or costs depend on it:
//   * column-major storage, Dynamic = -1, fixed and dynamic sizes, comma initialiser, blocks;
//   * LAZY expressions: `const auto K_t = A * B.transpose() * (...).inverse()` keeps an expression that
//     is re-evaluated at every use (reflector_ekf_slam.cc:305-308, :354-357) — operands that are plain
//     matrices nest by reference, expressions nest by value, like Eigen;
//   * every matrix product is evaluated into a temporary (Eigen's "products alias") and assignment of
//     an expression evaluates the right-hand side completely before the destination is touched;
//   * dynamic .inverse() = partial-pivot LU + solve against the identity (PartialPivLU::inverse);
//   * 1x1 fixed-size expressions convert to their scalar (:411, :437).
// What is NOT modelled: Eigen's exact blocked summation order (unspecified even between Eigen
// versions, so agreement with a real Eigen build is tolerance-level for anyone), SIMD packets,
// alignment, and everything the reference does not call.
#ifndef REKF_ORACLE_EIGEN_SHIM_H
#define REKF_ORACLE_EIGEN_SHIM_H

#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <type_traits>
#include <utility>
#include <vector>

namespace Eigen
{
const int Dynamic = -1;
typedef std::ptrdiff_t Index;

template <class T, int R, int C> class Matrix;
template <class X> class Transpose;
template <class X> class Inverse;
template <class X, class U> class Cast;
template <class L, class Rr> class Product;
template <class L, class Rr, int Sign> class AddSub;
template <class X, int Op> class ScalarOp; // Op 0: x*s, 1: x/s, 2: -x
template <class M> class Block;

namespace internal
{
template <class D> struct traits;
template <class X> struct dtraits : traits<typename std::decay<X>::type> {};
template <class T, int R, int C> struct traits<Matrix<T, R, C>> { typedef T Scalar; enum { Rows = R, Cols = C }; };
template <class X> struct traits<Transpose<X>> { typedef typename dtraits<X>::Scalar Scalar; enum { Rows = dtraits<X>::Cols, Cols = dtraits<X>::Rows }; };
template <class X> struct traits<Inverse<X>> { typedef typename dtraits<X>::Scalar Scalar; enum { Rows = dtraits<X>::Rows, Cols = dtraits<X>::Cols }; };
template <class X, class U> struct traits<Cast<X, U>> { typedef U Scalar; enum { Rows = dtraits<X>::Rows, Cols = dtraits<X>::Cols }; };
template <class L, class Rr> struct traits<Product<L, Rr>> { typedef typename dtraits<L>::Scalar Scalar; enum { Rows = dtraits<L>::Rows, Cols = dtraits<Rr>::Cols }; };
template <class L, class Rr, int S> struct traits<AddSub<L, Rr, S>>
{
  typedef typename dtraits<L>::Scalar Scalar;
  enum { Rows = int(dtraits<L>::Rows) != Dynamic ? int(dtraits<L>::Rows) : int(dtraits<Rr>::Rows), Cols = int(dtraits<L>::Cols) != Dynamic ? int(dtraits<L>::Cols) : int(dtraits<Rr>::Cols) };
};
template <class X, int Op> struct traits<ScalarOp<X, Op>> { typedef typename dtraits<X>::Scalar Scalar; enum { Rows = dtraits<X>::Rows, Cols = dtraits<X>::Cols }; };
template <class M> struct traits<Block<M>> { typedef typename dtraits<M>::Scalar Scalar; enum { Rows = Dynamic, Cols = dtraits<M>::Cols == 1 ? 1 : Dynamic }; };

// how an operand is stored inside an expression node: plain matrices by reference when they are
// lvalues, by value when they are temporaries; expressions always by value
template <class A> struct nest
{
  typedef typename std::decay<A>::type D;
  typedef D type;
};
template <class T, int R, int C> struct nest<Matrix<T, R, C> &> { typedef const Matrix<T, R, C> &type; };
template <class T, int R, int C> struct nest<const Matrix<T, R, C> &> { typedef const Matrix<T, R, C> &type; };

// read-only strided view of evaluated data (trans: element (i,j) lives at p[j + i*ld])
template <class T> struct View
{
  const T *p;
  Index ld, rows, cols;
  bool trans;
  T operator()(Index i, Index j) const { return trans ? p[j + i * ld] : p[i + j * ld]; }
};

template <class T> void gemm(const View<T> &A, const View<T> &B, T *C, Index ldc);
template <class T> bool lu_inverse(Index n, T *A);
} // namespace internal

// ------------------------------------------------------------------------------------------------
// common base of every expression (CRTP)
template <class D> class Base
{
public:
  typedef typename internal::traits<D>::Scalar Scalar;
  enum { Rows = internal::traits<D>::Rows, Cols = internal::traits<D>::Cols };
  typedef Matrix<Scalar, Rows, Cols> Plain;

  const D &derived() const { return *static_cast<const D *>(this); }
  D &derived() { return *static_cast<D *>(this); }

  Index rows() const { return derived().rows(); }
  Index cols() const { return derived().cols(); }
  Index size() const { return rows() * cols(); }

  Transpose<typename internal::nest<const D &>::type> transpose() const { return Transpose<typename internal::nest<const D &>::type>(derived()); }
  Inverse<typename internal::nest<const D &>::type> inverse() const { return Inverse<typename internal::nest<const D &>::type>(derived()); }
  template <class U> Cast<typename internal::nest<const D &>::type, U> cast() const { return Cast<typename internal::nest<const D &>::type, U>(derived()); }

  // coefficient access on an arbitrary expression evaluates it (plain matrices override these)
  Scalar operator()(Index i, Index j) const { return derived().eval()(i, j); }
  Scalar operator()(Index i) const { return derived().eval()(i); }
  Scalar x() const { return (*this)(0); }
  Scalar y() const { return (*this)(1); }
  Scalar z() const { return (*this)(2); }
  Scalar w() const { return (*this)(3); }

  Scalar squaredNorm() const
  {
    const auto &m = derived().eval();
    Scalar s = Scalar(0);
    for (Index k = 0; k < m.size(); ++k) s += m.data()[k] * m.data()[k];
    return s;
  }
  Scalar norm() const { return std::sqrt(squaredNorm()); }
  Plain normalized() const
  {
    Plain m = derived().eval();
    const Scalar n = m.norm();
    for (Index k = 0; k < m.size(); ++k) m.data()[k] /= n;
    return m;
  }
  template <int N> Matrix<Scalar, N, 1> head() const
  {
    const auto &m = derived().eval();
    Matrix<Scalar, N, 1> r;
    for (int k = 0; k < N; ++k) r(k) = m(k);
    return r;
  }

  // 1x1 fixed-size expressions convert to their scalar, as in Eigen
  template <class S, class = typename std::enable_if<std::is_same<S, Scalar>::value && Rows == 1 && Cols == 1>::type>
  operator S() const { return derived().eval()(0, 0); }
};

// ------------------------------------------------------------------------------------------------
namespace internal
{
template <class T, int R, int C, bool Fixed = (R != Dynamic && C != Dynamic)> struct Storage;
template <class T, int R, int C> struct Storage<T, R, C, true>
{
  T d[R * C];
  Storage() : d() {}
  T *data() { return d; }
  const T *data() const { return d; }
  Index rows() const { return R; }
  Index cols() const { return C; }
  void resize(Index r, Index c) { assert(r == R && c == C); (void)r; (void)c; }
};
template <class T, int R, int C> struct Storage<T, R, C, false>
{
  std::vector<T> d;
  Index r_, c_;
  Storage() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C) {}
  T *data() { return d.data(); }
  const T *data() const { return d.data(); }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  void resize(Index r, Index c)
  {
    assert((R == Dynamic || r == R) && (C == Dynamic || c == C));
    if (r * c != r_ * c_) d.resize(static_cast<size_t>(r * c));
    r_ = r;
    c_ = c;
  }
};

template <class M> class CommaInit
{
public:
  template <class S> CommaInit(M &m, const S &first) : m_(m), k_(0) { put(first); }
  template <class S> CommaInit &operator,(const S &s)
  {
    put(s);
    return *this;
  }

private:
  template <class S> void put(const S &s)
  {
    const Index c = m_.cols();
    assert(k_ < m_.size());
    m_(k_ / c, k_ % c) = static_cast<typename M::Scalar>(s); // row by row, like Eigen
    ++k_;
  }
  M &m_;
  Index k_;
};
} // namespace internal

template <class T, int R, int C> class Matrix : public Base<Matrix<T, R, C>>
{
public:
  typedef T Scalar;
  enum { Rows = R, Cols = C, IsVector = (R == 1 || C == 1), Fixed = (R != Dynamic && C != Dynamic) };

  Matrix() {}
  Matrix(const Matrix &) = default;
  Matrix(Matrix &&) = default;
  Matrix &operator=(const Matrix &) = default;
  Matrix &operator=(Matrix &&) = default;

  // Vector2(x, y) for fixed size-2 vectors; (rows, cols) otherwise
  template <class A, class B, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value>::type>
  Matrix(const A &a, const B &b) { init2(a, b, std::integral_constant<bool, (Fixed && R * C == 2)>()); }
  template <class A, class B, class Cc, class = typename std::enable_if<std::is_arithmetic<A>::value && std::is_arithmetic<B>::value && std::is_arithmetic<Cc>::value>::type>
  Matrix(const A &a, const B &b, const Cc &c)
  {
    static_assert(Fixed && R * C == 3, "3-coefficient constructor needs a fixed size-3 vector");
    s_.d[0] = static_cast<T>(a), s_.d[1] = static_cast<T>(b), s_.d[2] = static_cast<T>(c);
  }
  Matrix(const T &a, const T &b, const T &c, const T &d)
  {
    static_assert(Fixed && R * C == 4, "4-coefficient constructor needs a fixed size-4 vector");
    s_.d[0] = a, s_.d[1] = b, s_.d[2] = c, s_.d[3] = d;
  }
  // VectorXd(n)
  template <class A, class = typename std::enable_if<std::is_integral<A>::value>::type, class = void>
  explicit Matrix(const A &n)
  {
    static_assert(!Fixed && IsVector, "size constructor needs a dynamic vector");
    s_.resize(C == 1 ? Index(n) : 1, C == 1 ? 1 : Index(n));
  }
  // from any expression
  template <class D> Matrix(const Base<D> &e) { assign(e.derived().eval()); }

  template <class D> Matrix &operator=(const Base<D> &e)
  {
    take(e.derived().eval());
    return *this;
  }

  Index rows() const { return s_.rows(); }
  Index cols() const { return s_.cols(); }
  Index size() const { return rows() * cols(); }
  T *data() { return s_.data(); }
  const T *data() const { return s_.data(); }
  const Matrix &eval() const { return *this; }

  T &operator()(Index i, Index j) { return s_.data()[i + j * rows()]; }
  const T &operator()(Index i, Index j) const { return s_.data()[i + j * rows()]; }
  T &operator()(Index i) { return s_.data()[i]; }
  const T &operator()(Index i) const { return s_.data()[i]; }
  T &operator[](Index i) { return s_.data()[i]; }
  const T &operator[](Index i) const { return s_.data()[i]; }
  T &x() { return s_.data()[0]; }
  T &y() { return s_.data()[1]; }
  T &z() { return s_.data()[2]; }
  T &w() { return s_.data()[3]; }
  const T &x() const { return s_.data()[0]; }
  const T &y() const { return s_.data()[1]; }
  const T &z() const { return s_.data()[2]; }
  const T &w() const { return s_.data()[3]; }

  void resize(Index r, Index c) { s_.resize(r, c); }
  void resize(Index n) { s_.resize(C == 1 ? n : 1, C == 1 ? 1 : n); }
  Matrix &setZero()
  {
    std::fill(data(), data() + size(), T(0));
    return *this;
  }
  Matrix &setZero(Index n)
  {
    resize(n);
    return setZero();
  }

  static Matrix Zero()
  {
    Matrix m;
    m.setZero();
    return m;
  }
  static Matrix Zero(Index n)
  {
    Matrix m;
    m.resize(n);
    m.setZero();
    return m;
  }
  static Matrix Zero(Index r, Index c)
  {
    Matrix m;
    m.resize(r, c);
    m.setZero();
    return m;
  }
  static Matrix Identity()
  {
    Matrix m = Zero();
    for (Index k = 0; k < std::min(m.rows(), m.cols()); ++k) m(k, k) = T(1);
    return m;
  }
  static Matrix Identity(Index r, Index c)
  {
    Matrix m = Zero(r, c);
    for (Index k = 0; k < std::min(r, c); ++k) m(k, k) = T(1);
    return m;
  }
  static Matrix Unit(int k)
  {
    Matrix m = Zero();
    m(k) = T(1);
    return m;
  }
  static Matrix UnitX() { return Unit(0); }
  static Matrix UnitY() { return Unit(1); }
  static Matrix UnitZ() { return Unit(2); }

  template <class S> internal::CommaInit<Matrix> operator<<(const S &s) { return internal::CommaInit<Matrix>(*this, s); }

  Block<Matrix> block(Index i, Index j, Index r, Index c) { return Block<Matrix>(*this, i, j, r, c); }
  Block<const Matrix> block(Index i, Index j, Index r, Index c) const { return Block<const Matrix>(*this, i, j, r, c); }
  Block<Matrix> topRows(Index n) { return block(0, 0, n, cols()); }
  Block<const Matrix> topRows(Index n) const { return block(0, 0, n, cols()); }
  Block<Matrix> bottomRows(Index n) { return block(rows() - n, 0, n, cols()); }
  Block<const Matrix> bottomRows(Index n) const { return block(rows() - n, 0, n, cols()); }
  Block<Matrix> topLeftCorner(Index r, Index c) { return block(0, 0, r, c); }
  Block<const Matrix> topLeftCorner(Index r, Index c) const { return block(0, 0, r, c); }

  template <class D> Matrix &operator+=(const Base<D> &e)
  {
    const auto &v = e.derived().eval();
    assert(v.rows() == rows() && v.cols() == cols());
    for (Index k = 0; k < size(); ++k) data()[k] += v.data()[k];
    return *this;
  }
  template <class D> Matrix &operator-=(const Base<D> &e)
  {
    const auto &v = e.derived().eval();
    assert(v.rows() == rows() && v.cols() == cols());
    for (Index k = 0; k < size(); ++k) data()[k] -= v.data()[k];
    return *this;
  }
  Matrix &operator*=(const T &s)
  {
    for (Index k = 0; k < size(); ++k) data()[k] *= s;
    return *this;
  }
  Matrix &operator/=(const T &s)
  {
    for (Index k = 0; k < size(); ++k) data()[k] /= s;
    return *this;
  }

  // raw construction used by expression evaluation
  static Matrix Uninit(Index r, Index c)
  {
    Matrix m;
    m.resize(r, c);
    return m;
  }

private:
  template <class A, class B> void init2(const A &a, const B &b, std::true_type) { s_.d[0] = static_cast<T>(a), s_.d[1] = static_cast<T>(b); }
  template <class A, class B> void init2(const A &a, const B &b, std::false_type) { s_.resize(Index(a), Index(b)); }
  template <class U, int R2, int C2> void assign(const Matrix<U, R2, C2> &m)
  {
    static_assert(std::is_same<U, T>::value, "implicit scalar conversion between matrices is not allowed (use cast<>())");
    s_.resize(m.rows(), m.cols());
    std::copy(m.data(), m.data() + m.size(), data());
  }
  template <class U, int R2, int C2> void take(const Matrix<U, R2, C2> &m) { assign(m); }
  void take(Matrix &&m) { *this = std::move(m); }
  internal::Storage<T, R, C> s_;
};

template <class T, int R, int C> std::ostream &operator<<(std::ostream &os, const Matrix<T, R, C> &m)
{
  for (Index i = 0; i < m.rows(); ++i)
  {
    for (Index j = 0; j < m.cols(); ++j) os << (j ? " " : "") << m(i, j);
    if (i + 1 < m.rows()) os << "\n";
  }
  return os;
}

// ------------------------------------------------------------------------------------------------
// a rectangular window onto a plain matrix: readable expression and assignable lvalue
template <class M> class Block : public Base<Block<M>>
{
public:
  typedef typename std::remove_const<M>::type PlainM;
  typedef typename PlainM::Scalar Scalar;
  typedef Matrix<Scalar, internal::traits<Block>::Rows, internal::traits<Block>::Cols> Plain;
  Block(M &m, Index i, Index j, Index r, Index c) : m_(&m), i_(i), j_(j), r_(r), c_(c) { assert(i >= 0 && j >= 0 && i + r <= m.rows() && j + c <= m.cols()); }
  Index rows() const { return r_; }
  Index cols() const { return c_; }
  Scalar operator()(Index i, Index j) const { return (*m_)(i_ + i, j_ + j); }
  Scalar operator()(Index i) const { return c_ == 1 ? (*m_)(i_ + i, j_) : (*m_)(i_, j_ + i); }
  Plain eval() const
  {
    Plain out = Plain::Uninit(r_, c_);
    for (Index j = 0; j < c_; ++j)
      for (Index i = 0; i < r_; ++i) out(i, j) = (*m_)(i_ + i, j_ + j);
    return out;
  }
  internal::View<Scalar> view() const { return internal::View<Scalar>{m_->data() + i_ + j_ * m_->rows(), m_->rows(), r_, c_, false}; }

  template <class D> Block &operator=(const Base<D> &e)
  {
    store(e.derived().eval(), 0);
    return *this;
  }
  Block &operator=(const Block &e)
  {
    store(e.eval(), 0);
    return *this;
  }
  template <class D> Block &operator+=(const Base<D> &e)
  {
    store(e.derived().eval(), 1);
    return *this;
  }
  template <class D> Block &operator-=(const Base<D> &e)
  {
    store(e.derived().eval(), -1);
    return *this;
  }

private:
  template <class V> void store(const V &v, int mode)
  {
    assert(v.rows() == r_ && v.cols() == c_);
    for (Index j = 0; j < c_; ++j)
      for (Index i = 0; i < r_; ++i)
      {
        Scalar &dst = (*m_)(i_ + i, j_ + j);
        dst = mode == 0 ? v(i, j) : (mode > 0 ? dst + v(i, j) : dst - v(i, j));
      }
  }
  M *m_;
  Index i_, j_, r_, c_;
};

// ------------------------------------------------------------------------------------------------
namespace internal
{
// Operand<E>: evaluated data of an expression as a View, without copying plain matrices, their
// transposes or their blocks (Eigen hands those to its GEMM directly as well)
template <class E> struct Operand
{
  typedef typename traits<E>::Scalar T;
  typename Base<E>::Plain tmp;
  explicit Operand(const E &e) : tmp(e.eval()) {}
  View<T> view() const { return View<T>{tmp.data(), tmp.rows(), tmp.rows(), tmp.cols(), false}; }
};
template <class T, int R, int C> struct Operand<Matrix<T, R, C>>
{
  const Matrix<T, R, C> &m;
  explicit Operand(const Matrix<T, R, C> &e) : m(e) {}
  View<T> view() const { return View<T>{m.data(), m.rows(), m.rows(), m.cols(), false}; }
};
template <class M> struct Operand<Block<M>>
{
  Block<M> b;
  explicit Operand(const Block<M> &e) : b(e) {}
  View<typename Block<M>::Scalar> view() const { return b.view(); }
};
template <class X> struct Operand<Transpose<X>>
{
  typedef typename std::decay<X>::type XD;
  typedef typename traits<XD>::Scalar T;
  Operand<XD> inner;
  explicit Operand(const Transpose<X> &e) : inner(e.nested()) {}
  View<T> view() const
  {
    View<T> v = inner.view();
    std::swap(v.rows, v.cols);
    v.trans = !v.trans;
    return v;
  }
};
} // namespace internal

template <class X> class Transpose : public Base<Transpose<X>>
{
public:
  typedef typename std::decay<X>::type XD;
  typedef typename Base<Transpose>::Plain Plain;
  typedef typename Base<Transpose>::Scalar Scalar;
  explicit Transpose(const XD &x) : x_(x) {}
  const XD &nested() const { return x_; }
  Index rows() const { return x_.cols(); }
  Index cols() const { return x_.rows(); }
  Plain eval() const
  {
    const auto &m = x_.eval();
    Plain out = Plain::Uninit(m.cols(), m.rows());
    for (Index j = 0; j < m.cols(); ++j)
      for (Index i = 0; i < m.rows(); ++i) out(j, i) = m(i, j);
    return out;
  }

private:
  X x_;
};

template <class X, class U> class Cast : public Base<Cast<X, U>>
{
public:
  typedef typename std::decay<X>::type XD;
  typedef typename Base<Cast>::Plain Plain;
  explicit Cast(const XD &x) : x_(x) {}
  Index rows() const { return x_.rows(); }
  Index cols() const { return x_.cols(); }
  Plain eval() const
  {
    const auto &m = x_.eval();
    Plain out = Plain::Uninit(m.rows(), m.cols());
    for (Index k = 0; k < m.size(); ++k) out.data()[k] = static_cast<U>(m.data()[k]);
    return out;
  }

private:
  X x_;
};

template <class X> class Inverse : public Base<Inverse<X>>
{
public:
  typedef typename std::decay<X>::type XD;
  typedef typename Base<Inverse>::Plain Plain;
  typedef typename Base<Inverse>::Scalar Scalar;
  explicit Inverse(const XD &x) : x_(x) {}
  Index rows() const { return x_.rows(); }
  Index cols() const { return x_.cols(); }
  Plain eval() const
  {
    Plain a = x_.eval(); // copy: factorised in place
    assert(a.rows() == a.cols());
    const Index n = a.rows();
    if (Plain::Fixed && n == 2)
    { // Eigen's fixed-size 2x2 path: closed form
      const Scalar det = a(0, 0) * a(1, 1) - a(1, 0) * a(0, 1), inv = Scalar(1) / det;
      Plain r = a;
      r(0, 0) = a(1, 1) * inv, r(1, 1) = a(0, 0) * inv, r(0, 1) = -a(0, 1) * inv, r(1, 0) = -a(1, 0) * inv;
      return r;
    }
    internal::lu_inverse<Scalar>(n, a.data());
    return a;
  }

private:
  X x_;
};

template <class L, class Rr> class Product : public Base<Product<L, Rr>>
{
public:
  typedef typename std::decay<L>::type LD;
  typedef typename std::decay<Rr>::type RD;
  typedef typename Base<Product>::Plain Plain;
  typedef typename Base<Product>::Scalar Scalar;
  Product(const LD &l, const RD &r) : l_(l), r_(r) {}
  Index rows() const { return l_.rows(); }
  Index cols() const { return r_.cols(); }
  Plain eval() const
  {
    const internal::Operand<LD> a(l_);
    const internal::Operand<RD> b(r_);
    const internal::View<Scalar> va = a.view(), vb = b.view();
    assert(va.cols == vb.rows);
    Plain out = Plain::Uninit(va.rows, vb.cols);
    internal::gemm<Scalar>(va, vb, out.data(), va.rows);
    return out;
  }

private:
  L l_;
  Rr r_;
};

template <class L, class Rr, int Sign> class AddSub : public Base<AddSub<L, Rr, Sign>>
{
public:
  typedef typename std::decay<L>::type LD;
  typedef typename std::decay<Rr>::type RD;
  typedef typename Base<AddSub>::Plain Plain;
  AddSub(const LD &l, const RD &r) : l_(l), r_(r) {}
  Index rows() const { return l_.rows(); }
  Index cols() const { return l_.cols(); }
  Plain eval() const
  {
    const auto &a = l_.eval();
    const auto &b = r_.eval();
    assert(a.rows() == b.rows() && a.cols() == b.cols());
    Plain out = Plain::Uninit(a.rows(), a.cols());
    const Index n = a.size();
    if (Sign > 0)
      for (Index k = 0; k < n; ++k) out.data()[k] = a.data()[k] + b.data()[k];
    else
      for (Index k = 0; k < n; ++k) out.data()[k] = a.data()[k] - b.data()[k];
    return out;
  }

private:
  L l_;
  Rr r_;
};

template <class X, int Op> class ScalarOp : public Base<ScalarOp<X, Op>>
{
public:
  typedef typename std::decay<X>::type XD;
  typedef typename Base<ScalarOp>::Plain Plain;
  typedef typename Base<ScalarOp>::Scalar Scalar;
  ScalarOp(const XD &x, Scalar s) : x_(x), s_(s) {}
  Index rows() const { return x_.rows(); }
  Index cols() const { return x_.cols(); }
  Plain eval() const
  {
    const auto &a = x_.eval();
    Plain out = Plain::Uninit(a.rows(), a.cols());
    const Index n = a.size();
    for (Index k = 0; k < n; ++k) out.data()[k] = Op == 0 ? a.data()[k] * s_ : (Op == 1 ? a.data()[k] / s_ : -a.data()[k]);
    return out;
  }

private:
  X x_;
  Scalar s_;
};

// ------------------------------------------------------------------------------------------------
// operators.  Forwarding references so that temporaries are nested by value (no dangling `auto`).
namespace internal
{
template <class A> struct is_expr : std::is_base_of<Base<typename std::decay<A>::type>, typename std::decay<A>::type> {};
template <class A, class B> struct both_expr : std::integral_constant<bool, is_expr<A>::value && is_expr<B>::value> {};
} // namespace internal

template <class A, class B, class = typename std::enable_if<internal::both_expr<A, B>::value>::type>
Product<typename internal::nest<A>::type, typename internal::nest<B>::type> operator*(A &&a, B &&b)
{
  return Product<typename internal::nest<A>::type, typename internal::nest<B>::type>(a, b);
}
template <class A, class B, class = typename std::enable_if<internal::both_expr<A, B>::value>::type>
AddSub<typename internal::nest<A>::type, typename internal::nest<B>::type, 1> operator+(A &&a, B &&b)
{
  return AddSub<typename internal::nest<A>::type, typename internal::nest<B>::type, 1>(a, b);
}
template <class A, class B, class = typename std::enable_if<internal::both_expr<A, B>::value>::type>
AddSub<typename internal::nest<A>::type, typename internal::nest<B>::type, -1> operator-(A &&a, B &&b)
{
  return AddSub<typename internal::nest<A>::type, typename internal::nest<B>::type, -1>(a, b);
}
template <class A, class S, class = typename std::enable_if<internal::is_expr<A>::value && std::is_arithmetic<S>::value>::type>
ScalarOp<typename internal::nest<A>::type, 0> operator*(A &&a, const S &s)
{
  typedef typename std::decay<A>::type::Scalar T;
  return ScalarOp<typename internal::nest<A>::type, 0>(a, static_cast<T>(s));
}
template <class A, class S, class = typename std::enable_if<internal::is_expr<A>::value && std::is_arithmetic<S>::value>::type>
ScalarOp<typename internal::nest<A>::type, 0> operator*(const S &s, A &&a)
{
  typedef typename std::decay<A>::type::Scalar T;
  return ScalarOp<typename internal::nest<A>::type, 0>(a, static_cast<T>(s));
}
template <class A, class S, class = typename std::enable_if<internal::is_expr<A>::value && std::is_arithmetic<S>::value>::type>
ScalarOp<typename internal::nest<A>::type, 1> operator/(A &&a, const S &s)
{
  typedef typename std::decay<A>::type::Scalar T;
  return ScalarOp<typename internal::nest<A>::type, 1>(a, static_cast<T>(s));
}
template <class A, class = typename std::enable_if<internal::is_expr<A>::value>::type>
ScalarOp<typename internal::nest<A>::type, 2> operator-(A &&a)
{
  typedef typename std::decay<A>::type::Scalar T;
  return ScalarOp<typename internal::nest<A>::type, 2>(a, T(0));
}

// ------------------------------------------------------------------------------------------------
namespace internal
{
// C(M x N, column-major, ld = ldc) = A * B.  Small products: plain loops in Eigen's "lazy product"
// coefficient order.  Large ones: packed panels and a register-blocked micro-kernel (GotoBLAS
// scheme) — single-threaded, like the reference's OpenMP-less build (CMakeLists.txt:4-6).
#if defined(__AVX512F__)
enum { GEMM_VBYTES = 64, GEMM_NR = 12 }; // 2 x 12 = 24 of 32 zmm accumulators
#elif defined(__AVX__)
enum { GEMM_VBYTES = 32, GEMM_NR = 6 }; // 2 x 6 = 12 of 16 ymm accumulators
#else
enum { GEMM_VBYTES = 16, GEMM_NR = 4 }; // SSE2 (the reference's flag-less build): 2 x 4 = 8 of 16 xmm
#endif
enum { GEMM_MC = 96, GEMM_KC = 256, GEMM_NC = 2040 }; // MC, NC multiples of every MR / NR above
template <class T> struct gemm_cfg { enum { VL = GEMM_VBYTES / sizeof(T), MR = 2 * VL, NR = GEMM_NR }; };

// acc (MR x NR, column-major) = Ap panel (kc x MR) times Bp panel (kc x NR); GCC vector extensions
template <class T> inline void gemm_micro(Index kc, const T *__restrict ap, const T *__restrict bp, T *__restrict acc)
{
  enum { VL = gemm_cfg<T>::VL, MR = gemm_cfg<T>::MR, NR = gemm_cfg<T>::NR };
  typedef T vec __attribute__((vector_size(GEMM_VBYTES), aligned(sizeof(T))));
  vec c0[NR], c1[NR];
  for (int j = 0; j < NR; ++j) c0[j] = c1[j] = vec{};
  for (Index p = 0; p < kc; ++p)
  {
    const vec a0 = *reinterpret_cast<const vec *>(ap + p * MR);
    const vec a1 = *reinterpret_cast<const vec *>(ap + p * MR + VL);
    const T *b = bp + p * NR;
#pragma GCC unroll 12
    for (int j = 0; j < NR; ++j)
    {
      c0[j] += a0 * b[j];
      c1[j] += a1 * b[j];
    }
  }
  for (int j = 0; j < NR; ++j)
  {
    *reinterpret_cast<vec *>(acc + j * MR) = c0[j];
    *reinterpret_cast<vec *>(acc + j * MR + VL) = c1[j];
  }
}

template <class T> void gemm(const View<T> &A, const View<T> &B, T *C, Index ldc)
{
  enum { GEMM_MR = gemm_cfg<T>::MR };
  const Index M = A.rows, N = B.cols, K = A.cols;
  if (M * N * K <= 4096 || M < 4 || N < 4)
  {
    for (Index j = 0; j < N; ++j)
      for (Index i = 0; i < M; ++i)
      {
        T s = T(0);
        for (Index p = 0; p < K; ++p) s += A(i, p) * B(p, j);
        C[i + j * ldc] = s;
      }
    return;
  }
  for (Index j = 0; j < N; ++j) std::fill(C + j * ldc, C + j * ldc + M, T(0));
  std::vector<T> Ap(static_cast<size_t>((GEMM_MC + GEMM_MR) * GEMM_KC)), Bp(static_cast<size_t>((GEMM_NC + GEMM_NR) * GEMM_KC));
  T acc[GEMM_MR * GEMM_NR];
  for (Index jc = 0; jc < N; jc += GEMM_NC)
  {
    const Index nc = std::min<Index>(GEMM_NC, N - jc);
    for (Index pc = 0; pc < K; pc += GEMM_KC)
    {
      const Index kc = std::min<Index>(GEMM_KC, K - pc);
      // pack B(pc:pc+kc, jc:jc+nc) into NR-wide row panels, zero padded
      for (Index jr = 0; jr < nc; jr += GEMM_NR)
      {
        T *dst = Bp.data() + (jr / GEMM_NR) * kc * GEMM_NR;
        const Index nr = std::min<Index>(GEMM_NR, nc - jr);
        for (Index p = 0; p < kc; ++p)
          for (Index j = 0; j < GEMM_NR; ++j) dst[p * GEMM_NR + j] = j < nr ? B(pc + p, jc + jr + j) : T(0);
      }
      for (Index ic = 0; ic < M; ic += GEMM_MC)
      {
        const Index mc = std::min<Index>(GEMM_MC, M - ic);
        for (Index ir = 0; ir < mc; ir += GEMM_MR)
        {
          T *dst = Ap.data() + (ir / GEMM_MR) * kc * GEMM_MR;
          const Index mr = std::min<Index>(GEMM_MR, mc - ir);
          if (!A.trans && mr == GEMM_MR)
            for (Index p = 0; p < kc; ++p) std::memcpy(dst + p * GEMM_MR, A.p + (ic + ir) + (pc + p) * A.ld, sizeof(T) * GEMM_MR);
          else
            for (Index p = 0; p < kc; ++p)
              for (Index i = 0; i < GEMM_MR; ++i) dst[p * GEMM_MR + i] = i < mr ? A(ic + ir + i, pc + p) : T(0);
        }
        for (Index jr = 0; jr < nc; jr += GEMM_NR)
        {
          const Index nr = std::min<Index>(GEMM_NR, nc - jr);
          for (Index ir = 0; ir < mc; ir += GEMM_MR)
          {
            const Index mr = std::min<Index>(GEMM_MR, mc - ir);
            gemm_micro<T>(kc, Ap.data() + (ir / GEMM_MR) * kc * GEMM_MR, Bp.data() + (jr / GEMM_NR) * kc * GEMM_NR, acc);
            for (Index j = 0; j < nr; ++j)
            {
              T *c = C + (jc + jr + j) * ldc + ic + ir;
              for (Index i = 0; i < mr; ++i) c[i] += acc[j * GEMM_MR + i];
            }
          }
        }
      }
    }
  }
}

// in-place inverse by partial-pivot LU and a solve against the identity (PartialPivLU::inverse)
template <class T> bool lu_inverse(Index n, T *A)
{
  if (n <= 0) return true;
  std::vector<Index> piv(static_cast<size_t>(n));
  bool ok = true;
  for (Index k = 0; k < n; ++k)
  {
    Index p = k;
    T best = std::abs(A[k + k * n]);
    for (Index i = k + 1; i < n; ++i)
      if (std::abs(A[i + k * n]) > best) best = std::abs(A[i + k * n]), p = i;
    piv[k] = p;
    if (best == T(0)) { ok = false; continue; }
    if (p != k)
      for (Index j = 0; j < n; ++j) std::swap(A[k + j * n], A[p + j * n]);
    const T pivot = A[k + k * n];
    for (Index i = k + 1; i < n; ++i) A[i + k * n] /= pivot;
    for (Index j = k + 1; j < n; ++j)
    {
      const T u = A[k + j * n];
      if (u == T(0)) continue;
      T *cj = A + j * n;
      const T *lk = A + k * n;
      for (Index i = k + 1; i < n; ++i) cj[i] -= lk[i] * u;
    }
  }
  std::vector<T> X(static_cast<size_t>(n * n), T(0));
  for (Index c = 0; c < n; ++c) X[c + c * n] = T(1);
  for (Index k = 0; k < n; ++k) // rows of the identity follow the interchanges: X = P
    if (piv[k] != k)
      for (Index c = 0; c < n; ++c) std::swap(X[k + c * n], X[piv[k] + c * n]);
  for (Index c = 0; c < n; ++c)
  {
    T *x = X.data() + c * n;
    for (Index k = 0; k < n; ++k) // forward substitution, unit lower
    {
      const T xk = x[k];
      if (xk == T(0)) continue;
      const T *lk = A + k * n;
      for (Index i = k + 1; i < n; ++i) x[i] -= lk[i] * xk;
    }
    for (Index k = n - 1; k >= 0; --k) // backward substitution, upper
    {
      x[k] /= A[k + k * n];
      const T xk = x[k];
      const T *uk = A + k * n;
      for (Index i = 0; i < k; ++i) x[i] -= uk[i] * xk;
    }
  }
  std::copy(X.begin(), X.end(), A);
  return ok;
}
} // namespace internal

// ------------------------------------------------------------------------------------------------
// geometry: only what transform/rigid_transform.h, transform/transform.h and the detector touch
template <class T> class Rotation2D
{
public:
  typedef Matrix<T, 2, 1> Vector2;
  Rotation2D() : a_(T(0)) {}
  explicit Rotation2D(const T &a) : a_(a) {}
  static Rotation2D Identity() { return Rotation2D(T(0)); }
  T angle() const { return a_; }
  T &angle() { return a_; }
  Rotation2D inverse() const { return Rotation2D(-a_); }
  Rotation2D operator*(const Rotation2D &o) const { return Rotation2D(a_ + o.a_); }
  template <class D> Vector2 operator*(const Base<D> &v) const
  {
    const Vector2 p = v.derived().eval();
    const T s = std::sin(a_), c = std::cos(a_); // Eigen: toRotationMatrix() * vec
    return Vector2(c * p.x() - s * p.y(), s * p.x() + c * p.y());
  }
  template <class U> Rotation2D<U> cast() const { return Rotation2D<U>(static_cast<U>(a_)); }

private:
  T a_;
};
typedef Rotation2D<double> Rotation2Dd;
typedef Rotation2D<float> Rotation2Df;

template <class T> class AngleAxis
{
public:
  typedef Matrix<T, 3, 1> Vector3;
  AngleAxis() : angle_(T(0)), axis_(Vector3::UnitX()) {}
  template <class D> AngleAxis(const T &angle, const Base<D> &axis) : angle_(angle), axis_(axis.derived().eval()) {}
  T angle() const { return angle_; }
  const Vector3 &axis() const { return axis_; }

private:
  T angle_;
  Vector3 axis_;
};

template <class T> class Quaternion
{
public:
  typedef Matrix<T, 3, 1> Vector3;
  Quaternion() : w_(T(1)), x_(T(0)), y_(T(0)), z_(T(0)) {}
  Quaternion(const T &w, const T &x, const T &y, const T &z) : w_(w), x_(x), y_(y), z_(z) {}
  Quaternion(const AngleAxis<T> &aa)
  {
    const T h = aa.angle() / T(2), s = std::sin(h);
    w_ = std::cos(h), x_ = s * aa.axis().x(), y_ = s * aa.axis().y(), z_ = s * aa.axis().z();
  }
  static Quaternion Identity() { return Quaternion(); }
  T &w() { return w_; }
  T &x() { return x_; }
  T &y() { return y_; }
  T &z() { return z_; }
  const T &w() const { return w_; }
  const T &x() const { return x_; }
  const T &y() const { return y_; }
  const T &z() const { return z_; }
  Vector3 vec() const { return Vector3(x_, y_, z_); }
  T squaredNorm() const { return w_ * w_ + x_ * x_ + y_ * y_ + z_ * z_; }
  T norm() const { return std::sqrt(squaredNorm()); }
  Quaternion normalized() const
  {
    const T n = norm();
    return Quaternion(w_ / n, x_ / n, y_ / n, z_ / n);
  }
  Quaternion conjugate() const { return Quaternion(w_, -x_, -y_, -z_); }
  Quaternion inverse() const
  {
    const T n2 = squaredNorm();
    return Quaternion(w_ / n2, -x_ / n2, -y_ / n2, -z_ / n2);
  }
  Quaternion operator*(const Quaternion &b) const
  {
    return Quaternion(w_ * b.w_ - x_ * b.x_ - y_ * b.y_ - z_ * b.z_, w_ * b.x_ + x_ * b.w_ + y_ * b.z_ - z_ * b.y_,
                      w_ * b.y_ + y_ * b.w_ + z_ * b.x_ - x_ * b.z_, w_ * b.z_ + z_ * b.w_ + x_ * b.y_ - y_ * b.x_);
  }
  template <class D> Vector3 operator*(const Base<D> &vv) const
  { // Eigen's _transformVector: v + w*t + q.vec x t, t = 2 q.vec x v
    const Vector3 v = vv.derived().eval();
    const T tx = T(2) * (y_ * v.z() - z_ * v.y()), ty = T(2) * (z_ * v.x() - x_ * v.z()), tz = T(2) * (x_ * v.y() - y_ * v.x());
    return Vector3(v.x() + w_ * tx + (y_ * tz - z_ * ty), v.y() + w_ * ty + (z_ * tx - x_ * tz), v.z() + w_ * tz + (x_ * ty - y_ * tx));
  }
  template <class U> Quaternion<U> cast() const { return Quaternion<U>(static_cast<U>(w_), static_cast<U>(x_), static_cast<U>(y_), static_cast<U>(z_)); }

private:
  T w_, x_, y_, z_;
};
template <class T> Quaternion<T> operator*(const AngleAxis<T> &a, const AngleAxis<T> &b) { return Quaternion<T>(a) * Quaternion<T>(b); }
template <class T> Quaternion<T> operator*(const Quaternion<T> &a, const AngleAxis<T> &b) { return a * Quaternion<T>(b); }

template <class V> class Map : public std::remove_const<V>::type
{
public:
  typedef typename std::remove_const<V>::type Plain;
  explicit Map(const typename Plain::Scalar *p)
  {
    static_assert(Plain::Fixed, "Map shim covers fixed-size types only");
    std::copy(p, p + Plain::Rows * Plain::Cols, this->data());
  }
};

typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<float, Dynamic, Dynamic> MatrixXf;
typedef Matrix<double, Dynamic, 1> VectorXd;
typedef Matrix<float, Dynamic, 1> VectorXf;
typedef Matrix<double, 2, 2> Matrix2d;
typedef Matrix<float, 2, 2> Matrix2f;
typedef Matrix<double, 3, 3> Matrix3d;
typedef Matrix<float, 3, 3> Matrix3f;
typedef Matrix<double, 2, 1> Vector2d;
typedef Matrix<float, 2, 1> Vector2f;
typedef Matrix<int, 2, 1> Vector2i;
typedef Matrix<double, 3, 1> Vector3d;
typedef Matrix<float, 3, 1> Vector3f;
typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<float, 4, 1> Vector4f;
typedef Quaternion<double> Quaterniond;
typedef Quaternion<float> Quaternionf;
typedef AngleAxis<double> AngleAxisd;
typedef AngleAxis<float> AngleAxisf;
} // namespace Eigen

#endif // REKF_ORACLE_EIGEN_SHIM_H
