// mapHandler.hpp — PLSLAM::MapHandler facade for local bundle adjustment (reference include/mapHandler.h:73-157,
// src/mapHandler.cpp:1108-1712).  The residual / Jacobian / block-H assembly of every Levenberg-Marquardt pass
// runs on the GPU (gfpl_lba_assemble); the gather of the local map (pointer chasing), the damping, the linear
// solve and the state update stay on the host, as in the reference.  The reference solves the damped system
// with Eigen::SimplicialLDLT on a dense->sparse copy (:1439-1441); here the same system is solved by
// eliminating the block-diagonal landmark part (Schur complement), which is algebraically the same solution.
#pragma once
#include <array>
#include <vector>

#include "stereoFrame.hpp"

namespace PLSLAM {

using gfpl::Matrix4d;
using gfpl::Vector2d;
using gfpl::Vector3d;
using gfpl::Vector6d;
using std::vector;
typedef std::array<int, 6> Vector6i;  // (lm_map, lm_loc, obs_idx, kf_map, kf_loc or -1, inlier)

// the fields of include/keyFrame.h and include/mapFeatures.h that local BA reads and writes
struct KeyFrame {
  int kf_idx = 0;
  bool local = false;
  Matrix4d T_kf_w = Matrix4d::Identity();
  Vector6d x_kf_w;
};
struct MapPoint {
  int idx = 0;
  bool local = false;
  Vector3d point3D;
  vector<Vector2d> obs_list;
  vector<int> kf_obs_list;
  vector<double> sigma_list;
};
struct MapLine {
  int idx = 0;
  bool local = false;
  Vector6d line3D;
  vector<Vector3d> obs_list;
  vector<int> kf_obs_list;
  vector<double> sigma_list;
};

class MapHandler {
 public:
  MapHandler(StVO::PinholeStereoCamera* cam_, StVO::GfplContext* gfpl_) : cam(cam_), gfpl(gfpl_) {}

  void localBundleAdjustment();  // mapHandler.cpp:1108-1215
  void levMarquardtOptimizationLBA(vector<double> X_aux, vector<int> kf_list, vector<int> pt_list, vector<int> ls_list,
                                   vector<Vector6i> pt_obs_list, vector<Vector6i> ls_obs_list);  // :1217-1712

  vector<KeyFrame*> map_keyframes;
  vector<MapPoint*> map_points;
  vector<MapLine*> map_lines;
  StVO::PinholeStereoCamera* cam;
  StVO::GfplContext* gfpl;
  int lba_iterations = 0;      // what the reference prints as "LBA Iterations"
  double lba_last_err = 0.0;
};

// SE(3) helpers of src/auxiliar.cpp:154-214 used by the KF update (:1443-1448)
Matrix4d inverse_se3(const Matrix4d& T);
Matrix4d expmap_se3(const Vector6d& x);
Vector6d logmap_se3(const Matrix4d& T);

}  // namespace PLSLAM
