// stereoFrame.hpp — StVO::StereoFrame / PointFeature / LineFeature / PinholeStereoCamera facade.
//
// Keeps the member and method names of the reference (include/stereoFrame.h:89-260,
// include/stereoFeatures.h:36-128, include/pinholeStereoCamera.h) for the matching half of the class and
// forwards the work to libgfpl.so through the C ABI (include/gfpl_c.h).  Feature DETECTION (ORB, LSD/LBD)
// stays with the host extractors: key points, key lines and descriptors are injected with setFeatures().
#pragma once
#include <list>
#include <stdexcept>
#include <string>
#include <vector>

#include "gfpl_types.hpp"

namespace StVO {

using gfpl::DescMat;
using gfpl::DMatch;
using gfpl::KeyLine;
using gfpl::KeyPoint;
using gfpl::Matrix3d;
using gfpl::Matrix4d;
using gfpl::Matrix6d;
using gfpl::Vector2d;
using gfpl::Vector3d;
using gfpl::Vector6d;
using gfpl::BFMatcher;
using std::vector;

// thrown when a C-ABI call fails (the reference has no error channel: assert / unchecked indexing)
struct GfplError : std::runtime_error {
  int code;
  GfplError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// owns the gfpl_ctx shared by the frames and the handler of one tracker (one CUDA device, one stream)
class GfplContext {
 public:
  explicit GfplContext(const gfpl_camera& cam, int device = 0, const gfpl_config* cfg = nullptr);
  ~GfplContext();
  GfplContext(const GfplContext&) = delete;
  gfpl_ctx* get() const { return ctx_; }
  gfpl_config& config() { return cfg_; }   // mutable like Config::x() = v; call apply() afterwards
  void apply();
  void check(int rc) const;

 private:
  gfpl_ctx* ctx_ = nullptr;
  gfpl_config cfg_;
};

// include/pinholeStereoCamera.h — the rectified model only (src/pinholeStereoCamera.cpp:133-170)
class PinholeStereoCamera {
 public:
  PinholeStereoCamera(int width, int height, double fx, double fy, double cx, double cy, double b)
      : width(width), height(height), fx(fx), fy(fy), cx(cx), cy(cy), b(b) {}
  Vector3d backProjection(const double& u, const double& v, const double& disp) const {
    Vector3d P;
    const double bd = b / disp;
    P(0) = bd * (u - cx); P(1) = bd * (v - cy); P(2) = bd * fx;
    return P;
  }
  Vector2d projection(const Vector3d& P) const {
    Vector2d uv;
    uv(0) = cx + fx * P(0) / P(2); uv(1) = cy + fy * P(1) / P(2);
    return uv;
  }
  double getDisparity(const double pZ) const { return fx * b / pZ; }
  int getWidth() const { return width; }
  int getHeight() const { return height; }
  double getB() const { return b; }
  double getFx() const { return fx; }
  double getFy() const { return fy; }
  double getCx() const { return cx; }
  double getCy() const { return cy; }
  gfpl_camera c_abi() const { return gfpl_camera{fx, fy, cx, cy, b, width, height}; }

 private:
  int width, height;
  double fx, fy, cx, cy, b;
};

// include/stereoFeatures.h:36-62
class PointFeature {
 public:
  int idx = -1;
  Vector2d pl, pl_obs;
  double disp = 0;
  Vector3d P;
  bool inlier = true;
  int level = 0;
  double sigma2 = 1.0;
  bool frame_matched = false;
};

// include/stereoFeatures.h:64-128 (fields on the tracking path)
class LineFeature {
 public:
  int idx = -1;
  Vector2d spl, epl, spl_obs, epl_obs;
  double sdisp = 0, edisp = 0, angle = 0, sdisp_obs = 0, edisp_obs = 0;
  Vector3d sP, eP;
  Vector3d le, le_obs;
  bool inlier = true;
  int level = 0;
  double sigma2 = 1.0;
  bool frame_matched = false;
  Matrix3d covSpt3D, covEpt3D;
  double cutRatio[2] = {0, 0};
  gfpl::Mat<6, 6> invCovPose;
  // written by estimateProjUncertainty_descent (stereoFeatures.h:116-123)
  vector<double> volumeArr;        // here: {volume at cut {0,0} is not kept} -> the accepted volume, volumeArr.back()
  vector<Vector2d> lPt2D;
  vector<Matrix3d> covlPt2D;
  int fst_idx = 0, lst_idx = 0;
};

// include/stereoFrame.h:89-260
class StereoFrame {
 public:
  StereoFrame(int idx_, PinholeStereoCamera* cam_, GfplContext* gfpl_, double time_stamp_ = 0.0);
  ~StereoFrame();

  // injection point replacing detectFeatures (stereoFrame.cpp:1019-1123): output of the host ORB / LSD+LBD extractors
  void setFeatures(const vector<KeyPoint>& pl, const DescMat& dl, const vector<KeyPoint>& pr, const DescMat& dr,
                   const vector<KeyLine>& ll, const DescMat& ldl, const vector<KeyLine>& lr, const DescMat& ldr);

  // stereoFrame.h:104-106 — the matching half (knn both ways, accept loops, back-projection)
  void extractInitialStereoFeatures(int fast_th = 20);   // stereoFrame.cpp:148-336
  void extractStereoFeatures_Brute(int fast_th = 20);    // stereoFrame.cpp:770-1016
  // stereoFrame.cpp:410-768 (the default of stereoFrameHandler.cpp:91); needs orbExtractor_r->mvImagePyramid:
  void setRightPyramid(const gfpl_pyramid& pyr) { right_pyramid = pyr; }   // one frame (data -> that frame's block)
  void extractStereoFeatures_ORBSLAM(int fast_th = 20);
  gfpl_pyramid right_pyramid{};

  // stereoFrame.h:116-132 — cv::BFMatcher wrappers (stereoFrame.cpp:1229-1257)
  void matchPointFeatures(BFMatcher* bfm, const DescMat& pdesc_1, const DescMat& pdesc_2,
                          vector<vector<DMatch>>& pmatches_12);
  void matchLineFeatures(BFMatcher* bfm, const DescMat& ldesc_1, const DescMat& ldesc_2,
                         vector<vector<DMatch>>& lmatches_12);
  void matchPointFeatures_radius(BFMatcher* bfm, const DescMat& pdesc_1, const DescMat& pdesc_2,
                                 vector<vector<DMatch>>& pmatches_12);
  void matchLineFeatures_radius(BFMatcher* bfm, const DescMat& ldesc_1, const DescMat& ldesc_2,
                                vector<vector<DMatch>>& lmatches_12);

  void estimateStereoUncertainty();                      // stereoFrame.cpp:1448-1484

  // members with the reference's names (stereoFrame.h:205-228)
  int frame_idx;
  double time_stamp;
  PinholeStereoCamera* cam;
  Matrix4d Tfw, DT;
  Matrix6d Tfw_cov, DT_cov;
  Vector6d DT_cov_eig;
  double err_norm = 0;
  vector<PointFeature*> stereo_pt;
  vector<LineFeature*> stereo_ls;
  vector<KeyPoint> points_l, points_r;
  vector<KeyLine> lines_l, lines_r;
  DescMat pdesc_l, pdesc_r, ldesc_l, ldesc_r;

  // C-ABI views of this frame (B = 1) used by the handler
  void exportStereo(vector<gfpl_point_feature>& pt, vector<gfpl_line_feature>& ln) const;

 private:
  void stereoMatch(int variant);
  void knn(int norm, const DescMat& a, const DescMat& b, vector<vector<DMatch>>& out);
  void radius(int norm, float r, const DescMat& a, const DescMat& b, vector<vector<DMatch>>& out);
  GfplContext* gfpl_;
};

}  // namespace StVO
