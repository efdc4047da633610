// Stand-in for <opencv4/opencv2/core.hpp>: a dense 2-D array with the handful of cv::Mat operations the reference's
// hot-path sources use (BatchMultiBevGen.cpp:123,133-136,163,180-181,191,201-210,237-238,244,271-291,311-316,340-355,371;
// CloudManip.cpp:83-108).  Third-party behaviour restated by us, NOT pinned by the reference (../README.md):
//   Mat::zeros / Mat::ones, `double * Mat` (saturate_cast of the double product), `Mat / Mat` (IEEE element-wise
//   divide for CV_32F — what cv::divide does in OpenCV 4), shared-data assignment, clone(), at<T>(), ptr(), elemSize();
//   Formatter::FMT_CSV text: "%3d" for 8-bit, "%.<prec>g" for float (default precision 8), ", " between values,
//   "\n" between rows and after the last row.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <ostream>
#include <string>
#include <vector>

#define CV_8U 0
#define CV_8S 1
#define CV_16U 2
#define CV_16S 3
#define CV_32S 4
#define CV_32F 5
#define CV_64F 6
#define CV_8UC1 CV_8U
#define CV_8SC1 CV_8S
#define CV_32FC1 CV_32F
#define CV_64FC1 CV_64F

namespace cv {

typedef unsigned char uchar;

template <class T> using Ptr = std::shared_ptr<T>;

class Mat {
 public:
  int rows = 0, cols = 0;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  void create(int r, int c, int type) {
    rows = r; cols = c; type_ = type;
    buf_ = std::make_shared<std::vector<uchar> >(static_cast<std::size_t>(r) * c * elemSize(), uchar(0));
  }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
  static Mat ones(int r, int c, int type) { Mat m(r, c, type); m.fill(1.0); return m; }
  int type() const { return type_; }
  int depth() const { return type_; }
  int channels() const { return 1; }
  bool empty() const { return !buf_ || buf_->empty(); }
  std::size_t elemSize() const { static const std::size_t sz[7] = {1, 1, 2, 2, 4, 4, 8}; return sz[type_]; }
  uchar* ptr(int r = 0) { return buf_->data() + static_cast<std::size_t>(r) * cols * elemSize(); }
  const uchar* ptr(int r = 0) const { return buf_->data() + static_cast<std::size_t>(r) * cols * elemSize(); }
  template <class T> T& at(int r, int c) { return reinterpret_cast<T*>(ptr(r))[c]; }
  template <class T> const T& at(int r, int c) const { return reinterpret_cast<const T*>(ptr(r))[c]; }
  Mat clone() const { Mat m; m.rows = rows; m.cols = cols; m.type_ = type_; if (buf_) m.buf_ = std::make_shared<std::vector<uchar> >(*buf_); return m; }
  double get(int r, int c) const {
    switch (type_) {
      case CV_8U: return at<std::uint8_t>(r, c); case CV_8S: return at<std::int8_t>(r, c);
      case CV_16U: return at<std::uint16_t>(r, c); case CV_16S: return at<std::int16_t>(r, c);
      case CV_32S: return at<std::int32_t>(r, c); case CV_32F: return at<float>(r, c); default: return at<double>(r, c);
    }
  }
  // saturate_cast<T>(double): round half to even (cvRound) and clamp for the integer depths, plain narrowing for float
  void set(int r, int c, double v) {
    switch (type_) {
      case CV_32F: at<float>(r, c) = static_cast<float>(v); return;
      case CV_64F: at<double>(r, c) = v; return;
      default: break;
    }
    double q = std::nearbyint(v);
    switch (type_) {
      case CV_8U: at<std::uint8_t>(r, c) = static_cast<std::uint8_t>(q < 0 ? 0 : q > 255 ? 255 : q); break;
      case CV_8S: at<std::int8_t>(r, c) = static_cast<std::int8_t>(q < -128 ? -128 : q > 127 ? 127 : q); break;
      case CV_16U: at<std::uint16_t>(r, c) = static_cast<std::uint16_t>(q < 0 ? 0 : q > 65535 ? 65535 : q); break;
      case CV_16S: at<std::int16_t>(r, c) = static_cast<std::int16_t>(q < -32768 ? -32768 : q > 32767 ? 32767 : q); break;
      default: at<std::int32_t>(r, c) = static_cast<std::int32_t>(q); break;
    }
  }
  void fill(double v) { for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) set(r, c, v); }
  void convertTo(Mat& dst, int type) const {
    Mat m(rows, cols, type);
    for (int r = 0; r < rows; r++) for (int c = 0; c < cols; c++) {
      double v = get(r, c);
      if (type != CV_32F && type != CV_64F && !(v == v)) v = 0;   // NaN -> 0 for the integer depths
      m.set(r, c, v);
    }
    dst = m;
  }
 private:
  int type_ = CV_8U;
  std::shared_ptr<std::vector<uchar> > buf_;
};

inline Mat operator*(double s, const Mat& a) {
  Mat m(a.rows, a.cols, a.type());
  for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.set(r, c, a.get(r, c) * s);
  return m;
}
inline Mat operator*(const Mat& a, double s) { return s * a; }

// element-wise divide; CV_32F operands divide in float (IEEE), as cv::divide does
inline Mat operator/(const Mat& a, const Mat& b) {
  Mat m(a.rows, a.cols, a.type());
  if (a.type() == CV_32F && b.type() == CV_32F) {
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.at<float>(r, c) = a.at<float>(r, c) / b.at<float>(r, c);
  } else {
    for (int r = 0; r < a.rows; r++) for (int c = 0; c < a.cols; c++) m.set(r, c, a.get(r, c) / b.get(r, c));
  }
  return m;
}

class Formatted {
 public:
  std::string text;
};
inline std::ostream& operator<<(std::ostream& os, const Ptr<Formatted>& f) { return os << f->text; }

namespace stub { inline bool& format_enabled() { static bool on = true; return on; } }   // false: format() returns "" (timing runs)

class Formatter {
 public:
  enum FormatType { FMT_DEFAULT = 0, FMT_MATLAB = 1, FMT_CSV = 2, FMT_PYTHON = 3, FMT_NUMPY = 4, FMT_C = 5 };
  static Ptr<Formatter> get(FormatType = FMT_DEFAULT) { return std::make_shared<Formatter>(); }
  void set16fPrecision(int p = 4) { (void)p; }
  void set32fPrecision(int p = 8) { prec32f_ = p; }
  void set64fPrecision(int p = 16) { prec64f_ = p; }
  void setMultiline(bool = true) {}
  Ptr<Formatted> format(const Mat& m) const {
    Ptr<Formatted> out = std::make_shared<Formatted>();
    if (!stub::format_enabled()) return out;
    std::string& s = out->text;
    char fl[16], buf[80];
    std::snprintf(fl, sizeof fl, "%%.%dg", m.type() == CV_64F ? prec64f_ : prec32f_);
    for (int r = 0; r < m.rows; r++) {
      for (int c = 0; c < m.cols; c++) {
        if (m.type() == CV_32F || m.type() == CV_64F) {
          double v = m.get(r, c);
          if (v != v) std::snprintf(buf, sizeof buf, "nan");
          else if (std::isinf(v)) std::snprintf(buf, sizeof buf, "%s", v > 0 ? "inf" : "-inf");
          else std::snprintf(buf, sizeof buf, fl, v);
        } else if (m.type() == CV_8U || m.type() == CV_8S) {
          std::snprintf(buf, sizeof buf, "%3d", static_cast<int>(m.get(r, c)));
        } else {
          std::snprintf(buf, sizeof buf, "%d", static_cast<int>(m.get(r, c)));
        }
        s += buf;
        if (c + 1 < m.cols) s += ", ";
      }
      if (r + 1 < m.rows) s += "\n";
    }
    if (m.cols > 1) s += "\n";
    return out;
  }
 private:
  int prec32f_ = 8, prec64f_ = 16;
};

inline Ptr<Formatted> format(const Mat& m, Formatter::FormatType t) { return Formatter::get(t)->format(m); }

}  // namespace cv
