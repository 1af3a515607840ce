// mapHandler.hpp — PLSLAM::MapHandler facade for map matching and local bundle adjustment (reference
// include/mapHandler.h:73-157, src/mapHandler.cpp:189-772, 1108-1712).  Every Levenberg-Marquardt pass — residuals,
// Jacobians, block-H assembly and the damped solve — runs on the GPU (gfpl_lba_step); the gather of the local map
// (pointer chasing), the error accounting, the accept / reject of a step and the state update stay on the host, as in
// the reference.  The reference solves the damped system with Eigen::SimplicialLDLT on a dense->sparse copy
// (:1439-1441); the device eliminates the block-diagonal landmark part (Schur complement): the same solution.
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
typedef std::array<uint8_t, GFPL_DESC_BYTES> Descriptor;
struct KeyFrame {
  int kf_idx = 0;
  bool local = false;
  Matrix4d T_kf_w = Matrix4d::Identity();
  Vector6d x_kf_w;
  StVO::StereoFrame* stereo_frame = nullptr;   // shares the feature pointers like the reference (keyFrame.cpp:38-39)
};
struct MapPoint {  // include/mapFeatures.h:36-62
  int idx = 0;
  bool local = false;
  Vector3d point3D;
  vector<Vector2d> obs_list;
  vector<Vector3d> dir_list;
  vector<int> kf_obs_list;
  vector<double> sigma_list;
  vector<Descriptor> desc_list;
  Descriptor med_desc{};          // representative descriptor (mapFeatures.cpp:51-82), refreshed by MapHandler::refreshMedoids
  Vector3d med_obs_dir;
  void addMapPointObservation(const uint8_t* desc, int kf_idx, const Vector2d& obs, const Vector3d& dir) {  // mapFeatures.cpp:42-49
    Descriptor d;
    std::memcpy(d.data(), desc, GFPL_DESC_BYTES);
    desc_list.push_back(d); obs_list.push_back(obs); kf_obs_list.push_back(kf_idx); dir_list.push_back(dir);
    sigma_list.push_back(1.0);   // default sigma2_ = 1.f (mapFeatures.h:46,49)
  }
};
struct MapLine {   // include/mapFeatures.h:64-95
  int idx = 0;
  bool local = false;
  Vector6d line3D;
  vector<Vector3d> obs_list;
  vector<Vector3d> dir_list;
  vector<gfpl::Mat<4, 1>> pts_list;
  vector<int> kf_obs_list;
  vector<double> sigma_list;
  vector<Descriptor> desc_list;
  Descriptor med_desc{};
  Vector3d med_obs_dir;
  void addMapLineObservation(const uint8_t* desc, int kf_idx, const Vector3d& obs, const Vector3d& dir,
                             const gfpl::Mat<4, 1>& pts) {  // mapFeatures.cpp:118-126
    Descriptor d;
    std::memcpy(d.data(), desc, GFPL_DESC_BYTES);
    desc_list.push_back(d); obs_list.push_back(obs); kf_obs_list.push_back(kf_idx); dir_list.push_back(dir);
    pts_list.push_back(pts); sigma_list.push_back(1.0);
  }
};

class MapHandler {
 public:
  MapHandler(StVO::PinholeStereoCamera* cam_, StVO::GfplContext* gfpl_) : cam(cam_), gfpl(gfpl_) {}

  // lookForCommonMatches (mapHandler.cpp:189-772): matching + gates on the GPU (gfpl_kf_match for the KF <-> KF half,
  // gfpl_map_match for the local-map <-> KF half), landmark / graph bookkeeping on the host exactly in the
  // reference's loop order.  The optional pose refinement (:495-507, has_refinement is hard-wired false) is omitted.
  void lookForCommonMatches(KeyFrame* kf0, KeyFrame*& kf1);
  void matchLocalMap(KeyFrame* kf1, int& common_pt, int& common_ls);   // :509-770
  // MapPoint / MapLine::updateAverageDescDir (mapFeatures.cpp:51-92, 128-170) for the landmarks that gained an
  // observation, batched into one gfpl_medoid_desc call per feature type.  The reference refreshes inside every
  // add*Observation; med_desc is only read by the NEXT selection of local landmarks (:532, :656), which skips the
  // landmarks just observed by kf1 (:520, :642), so refreshing once per half gives the same values.
  void refreshMedoids(const vector<int>& pts, const vector<int>& lines);
  void localBundleAdjustment();  // mapHandler.cpp:1108-1215
  void levMarquardtOptimizationLBA(vector<double> X_aux, vector<int> kf_list, vector<int> pt_list, vector<int> ls_list,
                                   vector<Vector6i> pt_obs_list, vector<Vector6i> ls_obs_list);  // :1217-1712

  vector<KeyFrame*> map_keyframes;
  vector<MapPoint*> map_points;
  vector<MapLine*> map_lines;
  StVO::PinholeStereoCamera* cam;
  StVO::GfplContext* gfpl;
  int max_pt_idx = 0, max_ls_idx = 0;
  vector<vector<int>> full_graph;                       // covisibility counts (mapHandler.h:134)
  vector<vector<int>> map_points_kf_idx, map_lines_kf_idx;
  int last_common_pt = 0, last_common_ls = 0;
  int lba_iterations = 0;      // what the reference prints as "LBA Iterations"
  double lba_last_err = 0.0;
};

// SE(3) helpers of src/auxiliar.cpp:154-214 used by the KF update (:1443-1448)
Matrix4d inverse_se3(const Matrix4d& T);
Matrix4d expmap_se3(const Vector6d& x);
Vector6d logmap_se3(const Matrix4d& T);

}  // namespace PLSLAM
