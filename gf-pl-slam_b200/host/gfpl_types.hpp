// gfpl_types.hpp — the few third-party types the reference's class surface is written in, reduced to what
// the tracking hot path touches, so that the facades (stereoFrame.hpp, stereoFrameHandler.hpp,
// mapHandler.hpp) keep the reference's member and method signatures without Eigen / OpenCV installed.
//
//   Eigen::Matrix<double,R,C>  ->  gfpl::Mat<R,C>   column-major storage like Eigen's default, (i,j) access
//   cv::Mat (N x 32 CV_8UC1)   ->  gfpl::DescMat
//   cv::KeyPoint, cv::DMatch, cv::line_descriptor::KeyLine  ->  field-compatible PODs
//
// A maintainer of the reference swaps these typedefs for the real Eigen / OpenCV types:
// the facades only use operator()(i,j), data(), rows/cols and the listed fields.
#pragma once
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../include/gfpl_c.h"

namespace gfpl {

template <int R, int C>
struct Mat {
  double v[R * C];  // column-major (Eigen default): element (i,j) at v[j*R + i]
  Mat() { std::memset(v, 0, sizeof(v)); }
  double& operator()(int i, int j) { return v[j * R + i]; }
  double operator()(int i, int j) const { return v[j * R + i]; }
  double& operator()(int i) { return v[i]; }
  double operator()(int i) const { return v[i]; }
  double* data() { return v; }
  const double* data() const { return v; }
  static Mat Identity() {
    Mat m;
    for (int i = 0; i < (R < C ? R : C); i++) m(i, i) = 1.0;
    return m;
  }
  static Mat Zero() { return Mat(); }
  // row-major export / import: the C ABI is row-major
  void to_row_major(double* out) const {
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C; j++) out[i * C + j] = (*this)(i, j);
  }
  void from_row_major(const double* in) {
    for (int i = 0; i < R; i++)
      for (int j = 0; j < C; j++) (*this)(i, j) = in[i * C + j];
  }
};

template <int RA, int K, int CB>
inline Mat<RA, CB> operator*(const Mat<RA, K>& a, const Mat<K, CB>& b) {
  Mat<RA, CB> c;
  for (int i = 0; i < RA; i++)
    for (int j = 0; j < CB; j++) {
      double s = 0.0;
      for (int k = 0; k < K; k++) s += a(i, k) * b(k, j);
      c(i, j) = s;
    }
  return c;
}

typedef Mat<2, 1> Vector2d;
typedef Mat<3, 1> Vector3d;
typedef Mat<6, 1> Vector6d;
typedef Mat<3, 3> Matrix3d;
typedef Mat<4, 4> Matrix4d;
typedef Mat<6, 6> Matrix6d;

// N x 32 CV_8UC1 descriptor matrix (ORB: ORBextractor.cc:1037,1068; LBD: binary_descriptor_custom.cpp:639)
struct DescMat {
  std::vector<uint8_t> bytes;
  int rows = 0;
  static const int cols = GFPL_DESC_BYTES;
  const uint8_t* ptr(int r = 0) const { return bytes.data() + (size_t)r * cols; }
  uint8_t* ptr(int r = 0) { return bytes.data() + (size_t)r * cols; }
  void resize(int n) { rows = n; bytes.resize((size_t)n * cols); }
  void push_back_row(const uint8_t* row) { bytes.insert(bytes.end(), row, row + cols); rows++; }
  bool empty() const { return rows == 0; }
};

struct Point2f { float x = 0, y = 0; };
struct KeyPoint {  // cv::KeyPoint
  Point2f pt;
  float size = 0, angle = -1, response = 0;
  int octave = 0, class_id = -1;
};
struct DMatch {  // cv::DMatch
  int queryIdx = -1, trainIdx = -1, imgIdx = 0;
  float distance = 0;
};
struct KeyLine {  // cv::line_descriptor::KeyLine (descriptor_custom.hpp:105-170)
  float angle = 0;
  int class_id = -1, octave = 0;
  float startPointX = 0, startPointY = 0, endPointX = 0, endPointY = 0;
  float lineLength = 0, response = 0;
};

// cv::NormTypes
enum { NORM_HAMMING = GFPL_NORM_HAMMING, NORM_HAMMING2 = GFPL_NORM_HAMMING2 };

// stand-in for `cv::BFMatcher*`: the facade only needs the norm the matcher was constructed with
struct BFMatcher {
  int normType;
  bool crossCheck;
  explicit BFMatcher(int norm = NORM_HAMMING, bool cross = false) : normType(norm), crossCheck(cross) {}
};

}  // namespace gfpl
