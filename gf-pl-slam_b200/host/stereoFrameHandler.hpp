// stereoFrameHandler.hpp — StVO::StereoFrameHandler facade: frame-to-frame matching, good line cutting and
// the Gauss-Newton pose optimiser of the reference (include/stereoFrameHandler.h:38-174), forwarded to
// libgfpl.so through the C ABI.  Member and method names are the reference's; key-frame decision, adaptive
// FAST and visualisation stay with the host application (out of scope).
#pragma once
#include <list>

#include "stereoFrame.hpp"

namespace StVO {

class StereoFrameHandler {
 public:
  StereoFrameHandler(PinholeStereoCamera* cam_, GfplContext* gfpl_);
  ~StereoFrameHandler();

  // stereoFrameHandler.h:45-50 — the frames arrive with their detector output injected (setFeatures)
  void initialize(StereoFrame* first_frame);        // stereoFrameHandler.cpp:45-81
  void insertStereoPair(StereoFrame* next_frame);   // stereoFrameHandler.cpp:83-151
  void predictFramePose();                          // :153-157
  void updateFrame();                               // :926-961 (prev <- curr)

  // stereoFrameHandler.h:55-57
  void crossFrameMatching_Hybrid();                 // :451-695 (active in the reference, :109)
  void crossFrameMatching_Brute();                  // :698-860

  // stereoFrameHandler.h:59-60
  void optimizePose();                              // :1864-1937
  void optimizePose(Matrix4d DT_ini);               // :1939-2030

  // stereoFrameHandler.h:76-104
  void estimateProjUncertainty_submodular(const double stepCutRatio, const double rngCutRatio[2]);  // :1618-1764
  void estimateProjUncertainty_descent(const double stepCutRatio, const double rngCutRatio[2]);     // :1476-1610
  void getPoseInfoOnLine(const Matrix4d DT_inv, const Vector2d J_loss, const LineFeature* line, const double cutRatio[2],
                         gfpl::Mat<6, 6>& info_pose);   // :1342-1411 (one evaluation = a one-line cut problem)
  void updateEndPointByRatio(LineFeature* line);    // :1451-1470

  // stereoFrameHandler.h:165-167 — exposed for parity checks
  void optimizeFunctions(Matrix4d DT, Matrix6d& H, Vector6d& g, double& e);  // :2118-2245

  // members (stereoFrameHandler.h:131-160)
  std::list<PointFeature*> matched_pt;
  std::list<LineFeature*> matched_ls;
  StereoFrame* prev_frame = nullptr;
  StereoFrame* curr_frame = nullptr;
  PinholeStereoCamera* cam;
  int n_inliers = 0, n_inliers_pt = 0, n_inliers_ls = 0;
  int numFrameLoss = 0;
  int last_status = 0;   // gfpl_pose_result::status of the last optimizePose (data, not an error)

 private:
  void crossFrameMatching(int mode);
  void optimizePoseImpl(const Matrix4d* DT_ini);
  void buildObs(std::vector<gfpl_pt_obs>& pt, std::vector<uint8_t>& pin, std::vector<gfpl_ln_obs>& ln,
                std::vector<uint8_t>& lin) const;
  gfpl_frame_state prevState() const;
  GfplContext* gfpl_;
  std::vector<int> matched_ls_prev_idx_;   // index of each matched line in prev_frame->stereo_ls
};

}  // namespace StVO
