// oracle/shim/opencv2/highgui/highgui.hpp -- TEST INFRASTRUCTURE.  cv::imread for the one thing the reference's CLI
// drivers read with it (pwn_core/pwn_simple_aligner.cpp:137, pwn_aligner.cpp:150): binary PGM depth images (P5, 8 or 16
// bit, 16-bit samples big-endian as the Netpbm format defines and OpenCV decodes them).  Anything else returns an empty Mat.
#ifndef ORACLE_SHIM_OPENCV_HIGHGUI
#define ORACLE_SHIM_OPENCV_HIGHGUI
#include <cctype>
#include <fstream>
#include <string>
#include <vector>

#include "../core/core.hpp"

#define CV_LOAD_IMAGE_UNCHANGED -1

namespace cv {
inline Mat imread(const std::string &filename, int = CV_LOAD_IMAGE_UNCHANGED) {
  std::ifstream is(filename.c_str(), std::ios::binary);
  if (!is) return Mat();
  std::vector<char> buf((std::istreambuf_iterator<char>(is)), std::istreambuf_iterator<char>());
  size_t i = 0;
  long field[3];
  if (buf.size() < 2 || buf[0] != 'P' || buf[1] != '5') return Mat();
  i = 2;
  for (int k = 0; k < 3; k++) {
    for (;;) {
      while (i < buf.size() && std::isspace((unsigned char)buf[i])) i++;
      if (i < buf.size() && buf[i] == '#') {
        while (i < buf.size() && buf[i] != '\n') i++;
        continue;
      }
      break;
    }
    long v = 0;
    while (i < buf.size() && std::isdigit((unsigned char)buf[i])) v = 10 * v + (buf[i++] - '0');
    field[k] = v;
  }
  i++;  // the single whitespace after maxval
  const int cols = (int)field[0], rows = (int)field[1];
  const bool wide = field[2] > 255;
  if (buf.size() < i + (size_t)rows * cols * (wide ? 2 : 1)) return Mat();
  Mat m(rows, cols, wide ? CV_16UC1 : CV_8UC1);
  const unsigned char *p = (const unsigned char *)&buf[i];
  if (wide) {
    unsigned short *d = (unsigned short *)m.data;
    for (size_t k = 0; k < (size_t)rows * cols; k++) d[k] = (unsigned short)((p[2 * k] << 8) | p[2 * k + 1]);
  } else {
    std::memcpy(m.data, p, (size_t)rows * cols);
  }
  return m;
}
}  // namespace cv
#endif
