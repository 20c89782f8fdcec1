// ref_shim: just enough of OpenCV for HDLFrame::dumpToImage to compile (debug dump, unused)
#pragma once
#include <string>
#include <vector>
#define CV_8UC1 0
namespace cv {
struct Scalar { double v; Scalar(double x = 0) : v(x) {} };
class Mat {
 public:
  Mat() : r_(0), c_(0) {}
  Mat(int r, int c, int, const Scalar& s = Scalar()) : r_(r), c_(c), d_((size_t)r * c, (unsigned char)s.v) {}
  template <class T> T& at(int i, int j) { return reinterpret_cast<T&>(d_[(size_t)i * c_ + j]); }
  int rows() const { return r_; }
  int cols() const { return c_; }
 private:
  int r_, c_;
  std::vector<unsigned char> d_;
};
inline bool imwrite(const std::string&, const Mat&) { return false; }
}  // namespace cv
