// oracle/shim/opencv2/core/core.hpp -- TEST INFRASTRUCTURE.  The reference uses OpenCV on this path only as an image
// container (cv::Mat_<T>: rows, cols, create, setTo, operator()(r, c), data); this stand-in provides exactly that, with
// OpenCV's semantics where they matter: row-major storage, copies share the pixel buffer (reference counted), create()
// keeps the buffer when size and type are unchanged.  New buffers are zero-filled (OpenCV leaves them uninitialised).
#ifndef ORACLE_SHIM_OPENCV_CORE
#define ORACLE_SHIM_OPENCV_CORE
#include <cstring>
#include <memory>
#include <vector>

#define CV_8UC1 0
#define CV_8SC1 1
#define CV_16UC1 2
#define CV_16SC1 3
#define CV_32SC1 4
#define CV_32FC1 5
#define CV_64FC1 6

namespace cv {
typedef unsigned char uchar;
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
};
struct Rect { int x, y, width, height; Rect(int x_ = 0, int y_ = 0, int w = 0, int h = 0) : x(x_), y(y_), width(w), height(h) {} };
inline int cvElemSize(int type) { static const int s[] = {1, 1, 2, 2, 4, 4, 8}; return s[type]; }
template <class T> struct DataType;
template <> struct DataType<unsigned char> { enum { type = CV_8UC1 }; };
template <> struct DataType<char> { enum { type = CV_8SC1 }; };
template <> struct DataType<unsigned short> { enum { type = CV_16UC1 }; };
template <> struct DataType<short> { enum { type = CV_16SC1 }; };
template <> struct DataType<int> { enum { type = CV_32SC1 }; };
template <> struct DataType<unsigned int> { enum { type = CV_32SC1 }; };
template <> struct DataType<float> { enum { type = CV_32FC1 }; };
template <> struct DataType<double> { enum { type = CV_64FC1 }; };

class Mat {
 public:
  int rows, cols;
  uchar *data;
  size_t step;  // bytes per row of the underlying buffer (a region of interest keeps its parent's)
  Mat() : rows(0), cols(0), data(0), step(0), _type(CV_8UC1) {}
  Mat(int r, int c, int type) : rows(0), cols(0), data(0), step(0), _type(type) { create(r, c, type); }
  int type() const { return _type; }
  bool empty() const { return rows * cols == 0; }
  size_t total() const { return (size_t)rows * cols; }
  void create(int r, int c, int type) {
    if (r == rows && c == cols && type == _type && data) return;
    _buf.reset(new std::vector<uchar>((size_t)r * c * cvElemSize(type), 0));
    rows = r;
    cols = c;
    _type = type;
    step = (size_t)c * cvElemSize(type);
    data = _buf->empty() ? 0 : &(*_buf)[0];
  }
  // region of interest: a header onto the same pixels (Rect is x = first column, y = first row, width, height)
  Mat roi(const Rect &q) const {
    Mat m(*this);
    m.rows = q.height;
    m.cols = q.width;
    m.data = data + (size_t)q.y * step + (size_t)q.x * cvElemSize(_type);
    return m;
  }
  Mat clone() const {
    Mat m;
    m.create(rows, cols, _type);
    for (int r = 0; r < rows && data; r++) std::memcpy(m.data + (size_t)r * m.step, data + (size_t)r * step, (size_t)cols * cvElemSize(_type));
    return m;
  }
  Mat &setTo(const Scalar &s) {
    const size_t n = (size_t)rows * cols;
    switch (_type) {
      case CV_8UC1: fill<unsigned char>(n, s.val[0]); break;
      case CV_8SC1: fill<char>(n, s.val[0]); break;
      case CV_16UC1: fill<unsigned short>(n, s.val[0]); break;
      case CV_16SC1: fill<short>(n, s.val[0]); break;
      case CV_32SC1: fill<int>(n, s.val[0]); break;
      case CV_32FC1: fill<float>(n, s.val[0]); break;
      default: fill<double>(n, s.val[0]); break;
    }
    return *this;
  }
 protected:
  template <class T> void fill(size_t, double v) {
    for (int r = 0; r < rows; r++) {
      T *p = (T *)(data + (size_t)r * step);
      for (int c = 0; c < cols; c++) p[c] = (T)v;
    }
  }
  int _type;
  std::shared_ptr<std::vector<uchar> > _buf;
};

template <class T>
class Mat_ : public Mat {
 public:
  Mat_() : Mat() { _type = DataType<T>::type; }
  Mat_(int r, int c) : Mat(r, c, DataType<T>::type) {}
  Mat_(const Mat &m) : Mat(m) {}
  Mat_ &operator=(const Mat &m) { Mat::operator=(m); return *this; }
  void create(int r, int c) { Mat::create(r, c, DataType<T>::type); }
  T &operator()(int r, int c) { return ((T *)(data + (size_t)r * step))[c]; }
  const T &operator()(int r, int c) const { return ((const T *)(data + (size_t)r * step))[c]; }
  Mat_ operator()(const Rect &q) const { return Mat_(roi(q)); }
  Mat_ &setTo(const Scalar &s) { Mat::setTo(s); return *this; }
  Mat_ &setTo(double v) { Mat::setTo(Scalar(v)); return *this; }
  Mat_ clone() const { return Mat_(Mat::clone()); }
};
}  // namespace cv
#endif
