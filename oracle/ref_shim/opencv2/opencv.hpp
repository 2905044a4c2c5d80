#pragma once
// Stand-in for the slice of OpenCV the reference's src/vio.cpp, include/frame.h, include/feature.h touch: an owning 8-bit matrix
// (data / rows / cols / step.p[0] / clone / at / ptr) and no-op drawing / display calls (visualisation is off the pinned path).
#include <cstring>
#include <memory>
#include <string>
#include <vector>
#define CV_8UC1 0
#define CV_8U 0
#define CV_8UC3 16
#define CV_32FC1 5
#define CV_BGR2GRAY 6
#define CV_GRAY2BGR 8
#define CV_INTER_LINEAR 1
namespace cv {
struct Size { int width = 0, height = 0; Size() {} Size(int w, int h) : width(w), height(h) {} };
template <class T> struct Point_ { T x = 0, y = 0; Point_() {} Point_(T a, T b) : x(a), y(b) {} };
typedef Point_<float> Point2f;
typedef Point_<int> Point;
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {} };
struct Vec3b { unsigned char v[3]; unsigned char &operator[](int i) { return v[i]; } const unsigned char &operator[](int i) const { return v[i]; } };
enum { FONT_HERSHEY_COMPLEX = 3, FONT_HERSHEY_SIMPLEX = 0, COLOR_BGR2GRAY = 6, COLOR_GRAY2BGR = 8, INTER_LINEAR = 1, LINE_AA = 16 };
class Mat {
  std::shared_ptr<std::vector<unsigned char>> buf_;
 public:
  unsigned char *data = nullptr;
  int rows = 0, cols = 0, type_ = 0;
  struct Step { size_t p[2] = {0, 0}; operator size_t() const { return p[0]; } } step;
  Mat() {}
  Mat(int r, int c, int type) { create(r, c, type); }
  Mat(int r, int c, int type, void *ext) : data((unsigned char *)ext), rows(r), cols(c), type_(type) { step.p[0] = (size_t)c * elemSize(), step.p[1] = elemSize(); }
  Mat(Size s, int type) { create(s.height, s.width, type); }
  size_t elemSize() const { return type_ == CV_8UC3 ? 3 : (type_ == CV_32FC1 ? 4 : 1); }
  void create(int r, int c, int type) {
    rows = r, cols = c, type_ = type;
    buf_ = std::make_shared<std::vector<unsigned char>>((size_t)r * c * elemSize(), 0);
    data = buf_->data(), step.p[0] = (size_t)c * elemSize(), step.p[1] = elemSize();
  }
  Mat clone() const { Mat m; m.create(rows, cols, type_); if (data) memcpy(m.data, data, (size_t)rows * step.p[0]); return m; }
  void copyTo(Mat &o) const { o = clone(); }
  bool empty() const { return data == nullptr || rows == 0; }
  int type() const { return type_; }
  int channels() const { return type_ == CV_8UC3 ? 3 : 1; }
  Size size() const { return Size(cols, rows); }
  template <class T> T &at(int r, int c) { return *reinterpret_cast<T *>(data + (size_t)r * step.p[0] + (size_t)c * sizeof(T)); }
  template <class T> const T &at(int r, int c) const { return *reinterpret_cast<const T *>(data + (size_t)r * step.p[0] + (size_t)c * sizeof(T)); }
  template <class T> T *ptr(int r = 0) { return reinterpret_cast<T *>(data + (size_t)r * step.p[0]); }
  template <class T> const T *ptr(int r = 0) const { return reinterpret_cast<const T *>(data + (size_t)r * step.p[0]); }
  unsigned char *ptr(int r = 0) { return data + (size_t)r * step.p[0]; }
  static Mat zeros(int r, int c, int type) { return Mat(r, c, type); }
};
inline void circle(Mat &, Point2f, int, Scalar, int = 1, int = 8, int = 0) {}
inline void line(Mat &, Point2f, Point2f, Scalar, int = 1, int = 8, int = 0) {}
inline void rectangle(Mat &, Point2f, Point2f, Scalar, int = 1, int = 8, int = 0) {}
inline void putText(Mat &, const std::string &, Point2f, int, double, Scalar, int = 1, int = 8, bool = false) {}
inline bool imwrite(const std::string &, const Mat &) { return true; }
inline void imshow(const std::string &, const Mat &) {}
inline int waitKey(int = 0) { return -1; }
inline void hconcat(const Mat &a, const Mat &, Mat &o) { o = a.clone(); }
inline void cvtColor(const Mat &a, Mat &o, int, int = 0) { o = a.clone(); }
inline void resize(const Mat &a, Mat &o, Size, double = 0, double = 0, int = 1) { o = a.clone(); }
inline void absdiff(const Mat &a, const Mat &, Mat &o) { o = a.clone(); }
}  // namespace cv
