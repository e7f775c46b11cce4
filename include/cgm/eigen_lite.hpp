// Minimal stand-ins for the few Eigen types that appear in the reference's matcher / g2o-facing
// signatures (Vector2i/2f/2d, Vector3f/3d, Matrix3d, Rotation2Dd, Isometry-free). Used ONLY when
// the real Eigen is not installed (it is not in this image); with Eigen present the real headers
// are used and this file defines nothing. Semantics follow Eigen where the reference relies on
// them: construction from mixed scalars converts like a C++ cast (double -> int truncates).
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
#include <cmath>
#include <cstddef>
#include <memory>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW

namespace Eigen {

template <typename T>
using aligned_allocator = std::allocator<T>;

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
  template <typename A, typename B, typename C>
  Vec(A a, B b, C c) {
    static_assert(N == 3, "three-argument constructor needs a 3-vector");
    v[0] = static_cast<S>(a);
    v[1] = static_cast<S>(b);
    v[2] = static_cast<S>(c);
  }
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
  void setZero() {
    for (int i = 0; i < N; ++i) v[i] = S(0);
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
  Vec& operator+=(const Vec& o) {
    for (int i = 0; i < N; ++i) v[i] += o.v[i];
    return *this;
  }
  Vec& operator-=(const Vec& o) {
    for (int i = 0; i < N; ++i) v[i] -= o.v[i];
    return *this;
  }
  S squaredNorm() const {
    S s = S(0);
    for (int i = 0; i < N; ++i) s += v[i] * v[i];
    return s;
  }
  S norm() const { return std::sqrt(squaredNorm()); }
};

typedef Vec<int, 2> Vector2i;
typedef Vec<float, 2> Vector2f;
typedef Vec<double, 2> Vector2d;
typedef Vec<float, 3> Vector3f;
typedef Vec<double, 3> Vector3d;

struct Matrix3d {
  double m[9];  // row-major
  Matrix3d() {
    for (int i = 0; i < 9; ++i) m[i] = 0.0;
  }
  static Matrix3d Identity() {
    Matrix3d r;
    r.m[0] = r.m[4] = r.m[8] = 1.0;
    return r;
  }
  static Matrix3d Zero() { return Matrix3d(); }
  double& operator()(int i, int j) { return m[3 * i + j]; }
  const double& operator()(int i, int j) const { return m[3 * i + j]; }
  void setZero() {
    for (int i = 0; i < 9; ++i) m[i] = 0.0;
  }
  Vector3d operator*(const Vector3d& v) const {
    Vector3d r;
    for (int i = 0; i < 3; ++i) r[i] = m[3 * i] * v[0] + m[3 * i + 1] * v[1] + m[3 * i + 2] * v[2];
    return r;
  }
  Matrix3d operator*(const Matrix3d& o) const {
    Matrix3d r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j)
        r(i, j) = m[3 * i] * o(0, j) + m[3 * i + 1] * o(1, j) + m[3 * i + 2] * o(2, j);
    return r;
  }
  Matrix3d transpose() const {
    Matrix3d r;
    for (int i = 0; i < 3; ++i)
      for (int j = 0; j < 3; ++j) r(i, j) = (*this)(j, i);
    return r;
  }
  double determinant() const {
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) +
           m[2] * (m[3] * m[7] - m[4] * m[6]);
  }
  Matrix3d inverse() const {
    const double d = determinant();
    Matrix3d r;
    r.m[0] = (m[4] * m[8] - m[5] * m[7]) / d;
    r.m[1] = (m[2] * m[7] - m[1] * m[8]) / d;
    r.m[2] = (m[1] * m[5] - m[2] * m[4]) / d;
    r.m[3] = (m[5] * m[6] - m[3] * m[8]) / d;
    r.m[4] = (m[0] * m[8] - m[2] * m[6]) / d;
    r.m[5] = (m[2] * m[3] - m[0] * m[5]) / d;
    r.m[6] = (m[3] * m[7] - m[4] * m[6]) / d;
    r.m[7] = (m[1] * m[6] - m[0] * m[7]) / d;
    r.m[8] = (m[0] * m[4] - m[1] * m[3]) / d;
    return r;
  }
};

// Dynamic double matrix, just enough for SparseBlockMatrix<MatrixXd>::block(r, c).
struct MatrixXd {
  int r_ = 0, c_ = 0;
  std::vector<double> d;
  MatrixXd() {}
  MatrixXd(int r, int c) : r_(r), c_(c), d(static_cast<size_t>(r) * c, 0.0) {}
  int rows() const { return r_; }
  int cols() const { return c_; }
  double& operator()(int i, int j) { return d[static_cast<size_t>(i) * c_ + j]; }
  const double& operator()(int i, int j) const { return d[static_cast<size_t>(i) * c_ + j]; }
};

struct Rotation2Dd {
  double a;
  explicit Rotation2Dd(double angle = 0.0) : a(angle) {}
  double angle() const { return a; }
  Vector2d operator*(const Vector2d& v) const {
    const double c = std::cos(a), s = std::sin(a);
    return Vector2d(c * v.x() - s * v.y(), s * v.x() + c * v.y());
  }
};

}  // namespace Eigen
#endif  // CGM_HAVE_EIGEN
#endif
