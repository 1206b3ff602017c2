// oracle/shim/Eigen/src_rest.h -- Geometry (Transform, Quaternion), Eigenvalues, Cholesky, SVD of the Eigen stand-in.
// See Core for what this is and what it does not pin.
#ifndef ORACLE_SHIM_EIGEN_REST
#define ORACLE_SHIM_EIGEN_REST
#include "Core"

namespace Eigen {

// ---- Quaternion: construction from a rotation matrix as in Eigen ----
template <class S>
class Quaternion {
 public:
  S q[4];  // x y z w
  Quaternion() { q[0] = q[1] = q[2] = 0; q[3] = 1; }
  Quaternion(const S &w, const S &x, const S &y, const S &z) { q[0] = x; q[1] = y; q[2] = z; q[3] = w; }
  template <class D>
  explicit Quaternion(const MatrixBase<D> &R) {
    S t = (R.coeff(0, 0) + R.coeff(1, 1)) + R.coeff(2, 2);
    if (t > S(0)) {
      t = std::sqrt(t + S(1.0));
      q[3] = S(0.5) * t;
      t = S(0.5) / t;
      q[0] = (R.coeff(2, 1) - R.coeff(1, 2)) * t;
      q[1] = (R.coeff(0, 2) - R.coeff(2, 0)) * t;
      q[2] = (R.coeff(1, 0) - R.coeff(0, 1)) * t;
    } else {
      int i = 0;
      if (R.coeff(1, 1) > R.coeff(0, 0)) i = 1;
      if (R.coeff(2, 2) > R.coeff(i, i)) i = 2;
      int j = (i + 1) % 3, k = (j + 1) % 3;
      t = std::sqrt(R.coeff(i, i) - R.coeff(j, j) - R.coeff(k, k) + S(1.0));
      q[i] = S(0.5) * t;
      t = S(0.5) / t;
      q[3] = (R.coeff(k, j) - R.coeff(j, k)) * t;
      q[j] = (R.coeff(j, i) + R.coeff(i, j)) * t;
      q[k] = (R.coeff(k, i) + R.coeff(i, k)) * t;
    }
  }
  void normalize() {
    S n = std::sqrt(((q[0] * q[0] + q[1] * q[1]) + q[2] * q[2]) + q[3] * q[3]);
    for (int i = 0; i < 4; i++) q[i] = q[i] / n;
  }
  S x() const { return q[0]; }
  S y() const { return q[1]; }
  S z() const { return q[2]; }
  S w() const { return q[3]; }
  S &x() { return q[0]; }
  S &y() { return q[1]; }
  S &z() { return q[2]; }
  S &w() { return q[3]; }
  // Eigen's QuaternionBase::toRotationMatrix
  Matrix<S, 3, 3> toRotationMatrix() const {
    Matrix<S, 3, 3> R;
    const S tx = S(2) * q[0], ty = S(2) * q[1], tz = S(2) * q[2];
    const S twx = tx * q[3], twy = ty * q[3], twz = tz * q[3];
    const S txx = tx * q[0], txy = ty * q[0], txz = tz * q[0];
    const S tyy = ty * q[1], tyz = tz * q[1], tzz = tz * q[2];
    R(0, 0) = S(1) - (tyy + tzz); R(0, 1) = txy - twz; R(0, 2) = txz + twy;
    R(1, 0) = txy + twz; R(1, 1) = S(1) - (txx + tzz); R(1, 2) = tyz - twx;
    R(2, 0) = txz - twy; R(2, 1) = tyz + twx; R(2, 2) = S(1) - (txx + tyy);
    return R;
  }
};
typedef Quaternion<float> Quaternionf;
typedef Quaternion<double> Quaterniond;

// ---- Transform<Scalar, 3, Isometry> (Appendix A1) ----
template <class S, int Dim, int Mode>
class Transform {
 public:
  typedef S Scalar;
  typedef Matrix<S, 4, 4> MatrixType;
  MatrixType m;
  Transform() { m.setIdentity(); }  // Eigen leaves it uninitialised; callers here always assign
  template <class D> explicit Transform(const MatrixBase<D> &o) { m = o; }
  static Transform Identity() { return Transform(); }
  void setIdentity() { m.setIdentity(); }
  MatrixType &matrix() { return m; }
  const MatrixType &matrix() const { return m; }
  Ref<S, 3, 3> linear() { return m.template block<3, 3>(0, 0); }
  Matrix<S, 3, 3> linear() const { return m.template block<3, 3>(0, 0); }
  Matrix<S, 3, 3> rotation() const { return linear(); }
  Ref<S, 3, 1> translation() { return m.template block<3, 1>(0, 3); }
  Matrix<S, 3, 1> translation() const { return m.template block<3, 1>(0, 3); }
  S &operator()(int r, int c) { return m(r, c); }
  S operator()(int r, int c) const { return m(r, c); }
  template <class D> Transform &operator=(const MatrixBase<D> &o) { m = o; return *this; }
  void makeAffine() { m(3, 0) = 0; m(3, 1) = 0; m(3, 2) = 0; m(3, 3) = 1; }
  // Isometry: R' = R^T, t' = -(R^T t)
  Transform inverse() const {
    Transform o;
    Matrix<S, 3, 3> Rt = linear().transpose();
    o.linear() = Rt;
    o.translation() = -(Rt * translation());
    o.makeAffine();
    return o;
  }
  // R = Ra Rb, t = Ra tb + ta
  Transform operator*(const Transform &b) const {
    Transform o;
    o.linear() = linear() * b.linear();
    o.translation() = linear() * b.translation() + translation();
    o.makeAffine();
    return o;
  }
  // homogeneous 4-vectors / 4xN: the full matrix product
  template <class D>
  typename std::enable_if<(int)traits<D>::Rows == 4, Matrix<S, 4, traits<D>::Cols> >::type operator*(const MatrixBase<D> &v) const {
    return m * v;
  }
  // 3-vectors: affine action
  template <class D>
  typename std::enable_if<(int)traits<D>::Rows == 3, Matrix<S, 3, traits<D>::Cols> >::type operator*(const MatrixBase<D> &v) const {
    static_assert((int)traits<D>::Cols == 1, "3-vector");
    return linear() * v + translation();
  }
  const S *data() const { return m.data(); }
  S *data() { return m.data(); }
};
typedef Transform<float, 3, Isometry> Isometry3f;
typedef Transform<double, 3, Isometry> Isometry3d;
typedef Transform<float, 3, Affine> Affine3f;

// ---- SelfAdjointEigenSolver<Matrix3f>::computeDirect: the oracle's restatement (unpinned) ----
template <class M>
class SelfAdjointEigenSolver {
 public:
  SelfAdjointEigenSolver() {}
  template <class D>
  SelfAdjointEigenSolver &computeDirect(const MatrixBase<D> &A, int /*options*/ = ComputeEigenvectors) {
    static_assert((int)traits<D>::Rows == 3 && (int)traits<D>::Cols == 3, "3x3 only");
    Matrix<float, 3, 3> a = A;
    orc_eigen3(a.data(), _w.data(), _v.data());
    return *this;
  }
  const Matrix<float, 3, 1> &eigenvalues() const { return _w; }
  const Matrix<float, 3, 3> &eigenvectors() const { return _v; }
  ComputationInfo info() const { return Success; }
 private:
  Matrix<float, 3, 1> _w;
  Matrix<float, 3, 3> _v;
};

// ---- H.ldlt().solve(b): the oracle's restatement (unpinned) ----
template <class M>
class LDLT {
 public:
  M A;
  explicit LDLT(const M &a) : A(a) {}
  template <class D>
  Matrix<float, 6, 1> solve(const MatrixBase<D> &b) const {
    Matrix<float, 6, 1> bb = b, x;
    orc_ldlt_solve6(A.data(), bb.data(), x.data());
    return x;
  }
};
// ---- LLT (unscented.h): textbook lower Cholesky, float32, row by row like Eigen's unblocked llt_inplace ----
template <class M>
class LLT {
 public:
  M L;
  ComputationInfo _info;
  LLT() : _info(Success) {}
  template <class D>
  LLT &compute(const MatrixBase<D> &a_) {
    M a = a_;
    const int n = traits<M>::Rows;
    L.setZero();
    _info = Success;
    for (int j = 0; j < n; j++) {
      float s = a(j, j);
      for (int k = 0; k < j; k++) s -= L(j, k) * L(j, k);
      if (!(s > 0.0f)) _info = NumericalIssue;
      float d = std::sqrt(s);
      L(j, j) = d;
      for (int i = j + 1; i < n; i++) {
        float t = a(i, j);
        for (int k = 0; k < j; k++) t -= L(i, k) * L(j, k);
        L(i, j) = t / d;
      }
    }
    return *this;
  }
  ComputationInfo info() const { return _info; }
  const M &matrixL() const { return L; }
};

// ---- JacobiSVD as used by Aligner::_computeStatistics (symmetric input): pseudo-inverse and singular values through
// the oracle's float64 Jacobi eigen-solver (unpinned, tolerance-level: DESIGN.md known deviations) ----
template <class M>
class JacobiSVD {
 public:
  M A;
  Matrix<float, traits<M>::Rows, 1> sv;
  JacobiSVD() {}
  JacobiSVD(const M &a, unsigned int = 0) { compute(a); }
  template <class D>
  JacobiSVD &compute(const MatrixBase<D> &a, unsigned int = 0) {
    A = a;
    if ((int)traits<M>::Rows == 3) {
      Matrix<float, 3, 3> t = A.template block<3, 3>(0, 0);
      float s[3];
      orc_sym_singular_values3(t.data(), s);
      for (int i = 0; i < 3; i++) sv(i) = s[i];
    }
    return *this;
  }
  template <class D>
  M solve(const MatrixBase<D> &rhs) const {
    static_assert((int)traits<M>::Rows == 6, "pseudo-inverse of the 6x6 system only");
    Matrix<float, 6, 6> in = A, out;
    orc_sym_pinv6(in.data(), out.data());
    return out * rhs;
  }
  const Matrix<float, traits<M>::Rows, 1> &singularValues() const { return sv; }
};

// Matrix6f::inverse(): the oracle's float64 Gauss-Jordan
template <> struct inverse_impl<float, 6> {
  static Matrix<float, 6, 6> run(const Matrix<float, 6, 6> &a) {
    Matrix<float, 6, 6> o;
    orc_mat6_inverse(a.data(), o.data());
    return o;
  }
};

}  // namespace Eigen

// member templates that need the decompositions
namespace Eigen {
template <class D> LDLT<typename MatrixBase<D>::PlainObject> MatrixBase<D>::ldlt() const { return LDLT<PlainObject>(eval()); }
template <class D> LLT<typename MatrixBase<D>::PlainObject> MatrixBase<D>::llt() const { LLT<PlainObject> l; l.compute(eval()); return l; }
}
#endif
