// pwn/compat.h -- the handful of Eigen / cv::Mat_ types the pwn:: API surface is written in.
//
// The reference declares its API with Eigen3 fixed-size types and cv::Mat_ images
// (g2o_frontend/pwn_core/pwn_typedefs.h:3-62, homogeneousvector4f.h:17-93).  Neither library exists
// in this image, so these are minimal stand-ins with the same names, storage order (column-major
// matrices, row-major images) and the members the trackers call (matrix(), linear(), translation(),
// inverse(), operator*, create(), setTo(), operator()(r,c)).  A caller that has real Eigen/OpenCV
// converts through data(): the memory layouts are identical.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace pwn {

template <int R, int C>
struct Matrix {
  float m[R * C];  // column-major, like Eigen
  Matrix() { std::memset(m, 0, sizeof m); }
  static Matrix Zero() { return Matrix(); }
  static Matrix Identity() {
    Matrix a;
    for (int i = 0; i < (R < C ? R : C); i++) a(i, i) = 1.0f;
    return a;
  }
  void setZero() { std::memset(m, 0, sizeof m); }
  void setIdentity() { *this = Identity(); }
  float &operator()(int r, int c) { return m[c * R + r]; }
  float operator()(int r, int c) const { return m[c * R + r]; }
  float &operator()(int i) { return m[i]; }
  float operator()(int i) const { return m[i]; }
  float &operator[](int i) { return m[i]; }
  float operator[](int i) const { return m[i]; }
  float *data() { return m; }
  const float *data() const { return m; }
  int rows() const { return R; }
  int cols() const { return C; }
  Matrix<C, R> transpose() const {
    Matrix<C, R> t;
    for (int r = 0; r < R; r++)
      for (int c = 0; c < C; c++) t(c, r) = (*this)(r, c);
    return t;
  }
  template <int K>
  Matrix<R, K> operator*(const Matrix<C, K> &o) const {
    Matrix<R, K> out;
    for (int r = 0; r < R; r++)
      for (int k = 0; k < K; k++) {
        float s = 0.0f;
        for (int c = 0; c < C; c++) s += (*this)(r, c) * o(c, k);
        out(r, k) = s;
      }
    return out;
  }
  Matrix operator*(float f) const {
    Matrix o;
    for (int i = 0; i < R * C; i++) o.m[i] = m[i] * f;
    return o;
  }
  Matrix operator+(const Matrix &b) const {
    Matrix o;
    for (int i = 0; i < R * C; i++) o.m[i] = m[i] + b.m[i];
    return o;
  }
  Matrix operator-(const Matrix &b) const {
    Matrix o;
    for (int i = 0; i < R * C; i++) o.m[i] = m[i] - b.m[i];
    return o;
  }
  bool operator==(const Matrix &b) const { return std::memcmp(m, b.m, sizeof m) == 0; }
  bool operator!=(const Matrix &b) const { return !(*this == b); }
  float squaredNorm() const {
    float s = 0;
    for (int i = 0; i < R * C; i++) s += m[i] * m[i];
    return s;
  }
  float x() const { return m[0]; }
  float y() const { return m[1]; }
  float z() const { return m[2]; }
};

typedef Matrix<3, 3> Matrix3f;
typedef Matrix<4, 4> Matrix4f;
typedef Matrix<6, 6> Matrix6f;
typedef Matrix<3, 1> Vector3f;
typedef Matrix<4, 1> Vector4f;
typedef Matrix<6, 1> Vector6f;

// Eigen::Isometry3f stand-in (4x4 column-major)
struct Isometry3f {
  Matrix4f mat;
  Isometry3f() { mat.setIdentity(); }
  static Isometry3f Identity() { return Isometry3f(); }
  void setIdentity() { mat.setIdentity(); }
  Matrix4f &matrix() { return mat; }
  const Matrix4f &matrix() const { return mat; }
  const float *data() const { return mat.m; }
  float *data() { return mat.m; }
  Matrix3f linear() const {
    Matrix3f r;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) r(i, j) = mat(i, j);
    return r;
  }
  void setLinear(const Matrix3f &r) {
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) mat(i, j) = r(i, j);
  }
  Vector3f translation() const {
    Vector3f t;
    for (int i = 0; i < 3; i++) t(i) = mat(i, 3);
    return t;
  }
  void setTranslation(float x, float y, float z) { mat(0, 3) = x; mat(1, 3) = y; mat(2, 3) = z; }
  void fixLastRow() { mat(3, 0) = 0.f; mat(3, 1) = 0.f; mat(3, 2) = 0.f; mat(3, 3) = 1.f; }
  // Eigen Isometry inverse: R^T, -(R^T t)
  Isometry3f inverse() const {
    Isometry3f o;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) o.mat(r, c) = mat(c, r);
    for (int r = 0; r < 3; r++)
      o.mat(r, 3) = -((o.mat(r, 0) * mat(0, 3) + o.mat(r, 1) * mat(1, 3)) + o.mat(r, 2) * mat(2, 3));
    return o;
  }
  Isometry3f operator*(const Isometry3f &b) const {
    Isometry3f o;
    for (int r = 0; r < 3; r++) {
      for (int c = 0; c < 3; c++)
        o.mat(r, c) = (mat(r, 0) * b.mat(0, c) + mat(r, 1) * b.mat(1, c)) + mat(r, 2) * b.mat(2, c);
      o.mat(r, 3) = ((mat(r, 0) * b.mat(0, 3) + mat(r, 1) * b.mat(1, 3)) + mat(r, 2) * b.mat(2, 3)) + mat(r, 3);
    }
    return o;
  }
};

// cv::Mat_<T> stand-in: row-major rows x cols (pwn_typedefs.h:17-62)
template <typename T>
struct Image {
  int rows, cols;
  std::vector<T> buf;
  Image() : rows(0), cols(0) {}
  Image(int r, int c) : rows(0), cols(0) { create(r, c); }
  void create(int r, int c) {
    if (r != rows || c != cols) {
      rows = r;
      cols = c;
      buf.assign((size_t)r * c, T());
    }
  }
  void setTo(T v) { std::fill(buf.begin(), buf.end(), v); }
  T &operator()(int r, int c) { return buf[(size_t)r * cols + c]; }
  const T &operator()(int r, int c) const { return buf[(size_t)r * cols + c]; }
  T *data() { return buf.data(); }
  const T *data() const { return buf.data(); }
  size_t total() const { return buf.size(); }
  bool empty() const { return buf.empty(); }
};

typedef Image<float> DepthImage;         // metres
typedef Image<int> IntImage;
typedef Image<int> IndexImage;
typedef Image<uint16_t> RawDepthImage;   // millimetres

// homogeneousvector4f.h:17-83
struct Point : Vector4f {
  Point() { m[3] = 1.0f; }
  Point(float x, float y, float z) { m[0] = x; m[1] = y; m[2] = z; m[3] = 1.0f; }
};
struct Normal : Vector4f {
  Normal() { m[3] = 0.0f; }
  Normal(float x, float y, float z) { m[0] = x; m[1] = y; m[2] = z; m[3] = 0.0f; }
};
typedef std::vector<Point> PointVector;
typedef std::vector<Normal> NormalVector;

// informationmatrix.h:13-84 (4x4, last row/column zero)
struct InformationMatrix : Matrix4f {
  InformationMatrix() {}
  InformationMatrix(const Matrix4f &o) : Matrix4f(o) {
    for (int i = 0; i < 4; i++) { (*this)(3, i) = 0.f; (*this)(i, 3) = 0.f; }
  }
  void setDiagonal(float a, float b, float c) { setZero(); (*this)(0, 0) = a; (*this)(1, 1) = b; (*this)(2, 2) = c; }
};
typedef std::vector<InformationMatrix> InformationMatrixVector;

// stats.h:13-121
struct Stats : Matrix4f {
  int _n;
  Vector3f _eigenValues;
  // stats.h:98-119: the curvature is cached (and settable); a cloud that was downloaded without its Stats still
  // carries the per-point curvature the CorrespondenceFinder gates on
  mutable bool _curvatureComputed;
  mutable float _curvature;
  Stats() : _n(0), _curvatureComputed(false), _curvature(0.0f) { setIdentity(); }
  int n() const { return _n; }
  Vector3f eigenValues() const { return _eigenValues; }
  Matrix3f eigenVectors() const {
    Matrix3f r;
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) r(i, j) = (*this)(i, j);
    return r;
  }
  Point mean() const { return Point((*this)(0, 3), (*this)(1, 3), (*this)(2, 3)); }
  float curvature() const {
    if (!_curvatureComputed)
      _curvature = (float)((double)_eigenValues(0) / ((double)((_eigenValues(0) + _eigenValues(1)) + _eigenValues(2)) + 1e-9));
    _curvatureComputed = true;
    return _curvature;
  }
  void setCurvature(float curvature_) { _curvature = curvature_; _curvatureComputed = true; }
  void setN(int n_) { _n = n_; }
  void setEigenValues(const Vector3f &eigenValues_) { _eigenValues = eigenValues_; }
};
typedef std::vector<Stats> StatsVector;

// correspondencefinder.h:14-27
struct Correspondence {
  Correspondence(int referenceIndex_ = -1, int currentIndex_ = -1) : referenceIndex(referenceIndex_), currentIndex(currentIndex_) {}
  int referenceIndex, currentIndex;
};
typedef std::vector<Correspondence> CorrespondenceVector;

}  // namespace pwn
