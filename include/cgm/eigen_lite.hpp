// Minimal stand-ins for the Eigen types that appear in the reference's host sources (the matcher
// and g2o-facing signatures of src/matcher, src/slam, src/mrslam): Vector2i/2f/2d, Vector3f/3d,
// Matrix2d/3d/3f, MatrixXd, Matrix<unsigned char, Dynamic, Dynamic>, Rotation2Dd, Isometry2d.
// Used ONLY when the real Eigen is not installed (it is not in this image); with Eigen present the
// real headers are used and this file defines nothing. Semantics follow Eigen where the reference
// relies on them: construction from mixed scalars converts like a C++ cast (double -> int
// truncates), `m << a, b, c, d` fills row by row, the 2x2 / 3x3 inverse is the cofactor formula
// Eigen uses for fixed sizes up to 4. One deliberate difference: fixed-size objects are
// zero-initialised (Eigen leaves them uninitialised; src/mrslam/mr_graph_slam.cpp:374-383 and
// :461-467 read the never-written lower triangle of such a matrix).
#ifndef CGM_EIGEN_LITE_HPP
#define CGM_EIGEN_LITE_HPP

#if defined(__has_include)
#if __has_include(<Eigen/Core>) && !defined(CGM_FORCE_EIGEN_LITE)
#define CGM_HAVE_EIGEN 1
#endif
#endif

#ifdef CGM_HAVE_EIGEN
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <Eigen/StdVector>
#else
#include <algorithm>  // (the real Eigen/Core pulls it in; src/matcher/chargrid.cpp:307 relies on that)
#include <cmath>
#include <cstddef>
#include <memory>
#include <ostream>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

const int Dynamic = -1;

template <typename T>
using aligned_allocator = std::allocator<T>;

template <typename S, int R, int C>
struct FMat;

// `m << a, b, c;` (row-major fill)
template <typename S>
struct CommaInit {
  S* p;
  int n, k;
  CommaInit(S* p_, int n_, S first) : p(p_), n(n_), k(0) { p[k++] = first; }
  template <typename T>
  CommaInit& operator,(T v) {
    if (k < n) p[k++] = static_cast<S>(v);
    return *this;
  }
};

template <typename S, int N>
struct Vec {
  S v[N];
  Vec() {
    for (int i = 0; i < N; ++i) v[i] = S(0);
  }
  template <typename A, typename B>
  Vec(A a, B b) {
    static_assert(N == 2, "two-argument constructor needs a 2-vector");
    v[0] = static_cast<S>(a);
    v[1] = static_cast<S>(b);
  }
  template <typename A, typename B, typename D>
  Vec(A a, B b, D c) {
    static_assert(N == 3, "three-argument constructor needs a 3-vector");
    v[0] = static_cast<S>(a);
    v[1] = static_cast<S>(b);
    v[2] = static_cast<S>(c);
  }
  static Vec Zero() { return Vec(); }
  S& x() { return v[0]; }
  S& y() { return v[1]; }
  S& z() { return v[2]; }
  const S& x() const { return v[0]; }
  const S& y() const { return v[1]; }
  const S& z() const { return v[2]; }
  S& operator[](int i) { return v[i]; }
  const S& operator[](int i) const { return v[i]; }
  S& operator()(int i) { return v[i]; }
  const S& operator()(int i) const { return v[i]; }
  int size() const { return N; }
  void setZero() {
    for (int i = 0; i < N; ++i) v[i] = S(0);
  }
  template <typename T>
  CommaInit<S> operator<<(T first) {
    return CommaInit<S>(v, N, static_cast<S>(first));
  }
  Vec operator+(const Vec& o) const {
    Vec r;
    for (int i = 0; i < N; ++i) r.v[i] = v[i] + o.v[i];
    return r;
  }
  Vec operator-(const Vec& o) const {
    Vec r;
    for (int i = 0; i < N; ++i) r.v[i] = v[i] - o.v[i];
    return r;
  }
  Vec operator-() const {
    Vec r;
    for (int i = 0; i < N; ++i) r.v[i] = -v[i];
    return r;
  }
  Vec operator*(S s) const {
    Vec r;
    for (int i = 0; i < N; ++i) r.v[i] = v[i] * s;
    return r;
  }
  Vec operator/(S s) const {
    Vec r;
    for (int i = 0; i < N; ++i) r.v[i] = v[i] / s;
    return r;
  }
  Vec& operator+=(const Vec& o) {
    for (int i = 0; i < N; ++i) v[i] += o.v[i];
    return *this;
  }
  Vec& operator-=(const Vec& o) {
    for (int i = 0; i < N; ++i) v[i] -= o.v[i];
    return *this;
  }
  S dot(const Vec& o) const {
    S s = S(0);
    for (int i = 0; i < N; ++i) s += v[i] * o.v[i];
    return s;
  }
  S squaredNorm() const { return dot(*this); }
  S norm() const { return std::sqrt(squaredNorm()); }
  FMat<S, 1, N> transpose() const;
  template <typename T>
  Vec<T, N> cast() const {
    Vec<T, N> r;
    for (int i = 0; i < N; ++i) r.v[i] = static_cast<T>(v[i]);
    return r;
  }
};

template <typename T, typename S, int N>
Vec<S, N> operator*(T s, const Vec<S, N>& a) {
  return a * static_cast<S>(s);
}
template <typename S, int N>
std::ostream& operator<<(std::ostream& os, const Vec<S, N>& a) {
  for (int i = 0; i < N; ++i) os << (i ? "\n" : "") << a.v[i];
  return os;
}

typedef Vec<int, 2> Vector2i;
typedef Vec<float, 2> Vector2f;
typedef Vec<double, 2> Vector2d;
typedef Vec<float, 3> Vector3f;
typedef Vec<double, 3> Vector3d;

struct MatrixXd;

// matrix x column vector: a column vector, or (row vector x column vector, where Eigen yields a
// 1x1 matrix that converts to its entry) the scalar itself
template <typename S, int R>
struct ColResult {
  typedef Vec<S, R> type;
  static type make(const S* v) {
    type r;
    for (int i = 0; i < R; ++i) r[i] = v[i];
    return r;
  }
};
template <typename S>
struct ColResult<S, 1> {
  typedef S type;
  static type make(const S* v) { return v[0]; }
};

// Fixed-size matrix, row-major storage m[C * i + j].
template <typename S, int R, int C>
struct FMat {
  S m[R * C];
  FMat() {
    for (int i = 0; i < R * C; ++i) m[i] = S(0);
  }
  FMat(const MatrixXd& o);  // Matrix3d Pv = PvX  (graph_slam.cpp:600)
  static FMat Zero() { return FMat(); }
  static FMat Identity() {
    FMat r;
    for (int i = 0; i < (R < C ? R : C); ++i) r.m[C * i + i] = S(1);
    return r;
  }
  int rows() const { return R; }
  int cols() const { return C; }
  S& operator()(int i, int j) { return m[C * i + j]; }
  const S& operator()(int i, int j) const { return m[C * i + j]; }
  void setZero() {
    for (int i = 0; i < R * C; ++i) m[i] = S(0);
  }
  void setIdentity() { *this = Identity(); }
  template <typename T>
  CommaInit<S> operator<<(T first) {
    return CommaInit<S>(m, R * C, static_cast<S>(first));
  }
  FMat operator+(const FMat& o) const {
    FMat r;
    for (int i = 0; i < R * C; ++i) r.m[i] = m[i] + o.m[i];
    return r;
  }
  FMat operator-(const FMat& o) const {
    FMat r;
    for (int i = 0; i < R * C; ++i) r.m[i] = m[i] - o.m[i];
    return r;
  }
  FMat& operator+=(const FMat& o) {
    for (int i = 0; i < R * C; ++i) m[i] += o.m[i];
    return *this;
  }
  FMat operator*(S s) const {
    FMat r;
    for (int i = 0; i < R * C; ++i) r.m[i] = m[i] * s;
    return r;
  }
  template <int K>
  FMat<S, R, K> operator*(const FMat<S, C, K>& o) const {
    FMat<S, R, K> r;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < K; ++j) {
        S s = S(0);
        for (int k = 0; k < C; ++k) s += m[C * i + k] * o.m[K * k + j];
        r.m[K * i + j] = s;
      }
    return r;
  }
  typename ColResult<S, R>::type operator*(const Vec<S, C>& v) const {
    S r[R];
    for (int i = 0; i < R; ++i) {
      S s = S(0);
      for (int k = 0; k < C; ++k) s += m[C * i + k] * v[k];
      r[i] = s;
    }
    return ColResult<S, R>::make(r);
  }
  FMat<S, C, R> transpose() const {
    FMat<S, C, R> r;
    for (int i = 0; i < R; ++i)
      for (int j = 0; j < C; ++j) r.m[R * j + i] = m[C * i + j];
    return r;
  }
  template <typename T>
  FMat<T, R, C> cast() const {
    FMat<T, R, C> r;
    for (int i = 0; i < R * C; ++i) r.m[i] = static_cast<T>(m[i]);
    return r;
  }
  S determinant() const {
    static_assert(R == C && (R == 2 || R == 3), "determinant: 2x2 or 3x3");
    if (R == 2) return m[0] * m[3] - m[1] * m[2];
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
           m[2] * (m[3] * m[7] - m[4] * m[6]);
  }
  FMat inverse() const {  // cofactors times 1 / det, as Eigen does for fixed sizes <= 4
    static_assert(R == C && (R == 2 || R == 3), "inverse: 2x2 or 3x3");
    FMat r;
    const S invdet = S(1) / determinant();
    if (R == 2) {
      r.m[0] = m[3] * invdet;
      r.m[1] = -m[1] * invdet;
      r.m[2] = -m[2] * invdet;
      r.m[3] = m[0] * invdet;
      return r;
    }
    r.m[0] = (m[4] * m[8] - m[5] * m[7]) * invdet;
    r.m[1] = (m[2] * m[7] - m[1] * m[8]) * invdet;
    r.m[2] = (m[1] * m[5] - m[2] * m[4]) * invdet;
    r.m[3] = (m[5] * m[6] - m[3] * m[8]) * invdet;
    r.m[4] = (m[0] * m[8] - m[2] * m[6]) * invdet;
    r.m[5] = (m[2] * m[3] - m[0] * m[5]) * invdet;
    r.m[6] = (m[3] * m[7] - m[4] * m[6]) * invdet;
    r.m[7] = (m[1] * m[6] - m[0] * m[7]) * invdet;
    r.m[8] = (m[0] * m[4] - m[1] * m[3]) * invdet;
    return r;
  }
};

template <typename S, int R, int C, typename T>
FMat<S, R, C> operator*(T s, const FMat<S, R, C>& a) {
  return a * static_cast<S>(s);
}
template <typename S, int R, int C>
std::ostream& operator<<(std::ostream& os, const FMat<S, R, C>& a) {
  for (int i = 0; i < R; ++i) {
    for (int j = 0; j < C; ++j) os << (j ? " " : "") << a(i, j);
    if (i + 1 < R) os << "\n";
  }
  return os;
}
template <typename S, int N>
FMat<S, 1, N> Vec<S, N>::transpose() const {
  FMat<S, 1, N> r;
  for (int i = 0; i < N; ++i) r.m[i] = v[i];
  return r;
}

typedef FMat<double, 2, 2> Matrix2d;
typedef FMat<double, 3, 3> Matrix3d;
typedef FMat<float, 3, 3> Matrix3f;
typedef FMat<float, 2, 2> Matrix2f;

// Dynamic double matrix, just enough for SparseBlockMatrix<MatrixXd>::block(r, c) and the
// covariance plumbing of graph_manipulator.cpp:147-157.
struct MatrixXd {
  int r_ = 0, c_ = 0;
  std::vector<double> d;
  MatrixXd() {}
  MatrixXd(int r, int c) : r_(r), c_(c), d(static_cast<size_t>(r) * c, 0.0) {}
  template <int R, int C>
  MatrixXd(const FMat<double, R, C>& o) : r_(R), c_(C), d(o.m, o.m + R * C) {}
  int rows() const { return r_; }
  int cols() const { return c_; }
  void resize(int r, int c) {
    r_ = r;
    c_ = c;
    d.assign(static_cast<size_t>(r) * c, 0.0);
  }
  void setZero() { d.assign(d.size(), 0.0); }
  double& operator()(int i, int j) { return d[static_cast<size_t>(i) * c_ + j]; }
  const double& operator()(int i, int j) const { return d[static_cast<size_t>(i) * c_ + j]; }
};
template <typename S, int R, int C>
FMat<S, R, C>::FMat(const MatrixXd& o) {
  for (int i = 0; i < R; ++i)
    for (int j = 0; j < C; ++j) m[C * i + j] = static_cast<S>(o(i, j));
}
inline std::ostream& operator<<(std::ostream& os, const MatrixXd& a) {
  for (int i = 0; i < a.rows(); ++i) {
    for (int j = 0; j < a.cols(); ++j) os << (j ? " " : "") << a(i, j);
    if (i + 1 < a.rows()) os << "\n";
  }
  return os;
}

// Matrix<unsigned char, Dynamic, Dynamic> (the matcher's distance stamp, chargrid.h:46):
// column-major like Eigen's default, so data() has the layout the C ABI expects.
template <typename S, int R, int C>
class Matrix;
template <typename S>
class Matrix<S, Dynamic, Dynamic> {
 public:
  Matrix() : r_(0), c_(0) {}
  Matrix(int r, int c) : r_(r), c_(c), d_(static_cast<size_t>(r) * c) {}
  void resize(int r, int c) {
    r_ = r;
    c_ = c;
    d_.resize(static_cast<size_t>(r) * c);
  }
  void fill(S v) { d_.assign(d_.size(), v); }
  int rows() const { return r_; }
  int cols() const { return c_; }
  S& operator()(int i, int j) { return d_[static_cast<size_t>(j) * r_ + i]; }
  const S& operator()(int i, int j) const { return d_[static_cast<size_t>(j) * r_ + i]; }
  const S* data() const { return d_.data(); }
  S* data() { return d_.data(); }

 private:
  int r_, c_;
  std::vector<S> d_;
};

struct Rotation2Dd {
  double a;
  explicit Rotation2Dd(double angle = 0.0) : a(angle) {}
  double angle() const { return a; }
  double& angle() { return a; }
  Vector2d operator*(const Vector2d& v) const {
    const double c = std::cos(a), s = std::sin(a);
    return Vector2d(c * v.x() - s * v.y(), s * v.x() + c * v.y());
  }
  Rotation2Dd inverse() const { return Rotation2Dd(-a); }
  Matrix2d toRotationMatrix() const {
    Matrix2d r;
    const double c = std::cos(a), s = std::sin(a);
    r(0, 0) = c;
    r(0, 1) = -s;
    r(1, 0) = s;
    r(1, 1) = c;
    return r;
  }
};

// 2-D rigid transform as a 3x3 homogeneous matrix (CharGrid::integrateScan, chargrid.h:218-227)
struct Isometry2d {
  Matrix3d h;
  Isometry2d() : h(Matrix3d::Identity()) {}
  static Isometry2d Identity() { return Isometry2d(); }
  Vector2d operator*(const Vector2d& p) const {
    return Vector2d(h(0, 0) * p.x() + h(0, 1) * p.y() + h(0, 2), h(1, 0) * p.x() + h(1, 1) * p.y() + h(1, 2));
  }
};

}  // namespace Eigen
#endif  // CGM_HAVE_EIGEN
#endif
