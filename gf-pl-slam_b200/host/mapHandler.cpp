// mapHandler.cpp — PLSLAM::MapHandler local-BA facade over gfpl_lba_assemble (see mapHandler.hpp).
#include "mapHandler.hpp"

#include <cmath>
#include <limits>

namespace PLSLAM {

// ---- SE(3) glue (auxiliar.cpp:154-214) --------------------------------------------------------------------
Matrix4d inverse_se3(const Matrix4d& T) {
  Matrix4d Ti = Matrix4d::Identity();
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) Ti(i, j) = T(j, i);
    Ti(i, 3) = -(T(0, i) * T(0, 3) + T(1, i) * T(1, 3) + T(2, i) * T(2, 3));
  }
  return Ti;
}

static void skew(const double* v, double S[3][3]) {
  S[0][0] = 0; S[0][1] = -v[2]; S[0][2] = v[1];
  S[1][0] = v[2]; S[1][1] = 0; S[1][2] = -v[0];
  S[2][0] = -v[1]; S[2][1] = v[0]; S[2][2] = 0;
}

Matrix4d expmap_se3(const Vector6d& x) {
  double w[3] = {x(3), x(4), x(5)}, t[3] = {x(0), x(1), x(2)};
  double R[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  if (!(theta < 0.000001)) {
    double s[3][3], ss[3][3], V[3][3];
    skew(w, s);
    for (auto& r : s) for (double& v : r) v /= theta;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { ss[i][j] = 0; for (int k = 0; k < 3; k++) ss[i][j] += s[i][k] * s[k][j]; }
    const double sn = std::sin(theta), cs = std::cos(theta);
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        const double I = i == j ? 1.0 : 0.0;
        R[i][j] = I + s[i][j] * sn + ss[i][j] * (1.0 - cs);
        V[i][j] = I + s[i][j] * (1.0 - cs) / theta + ss[i][j] * (theta - sn) / theta;
      }
    double tn[3];
    for (int i = 0; i < 3; i++) tn[i] = V[i][0] * t[0] + V[i][1] * t[1] + V[i][2] * t[2];
    for (int i = 0; i < 3; i++) t[i] = tn[i];
  }
  Matrix4d T = Matrix4d::Identity();
  for (int i = 0; i < 3; i++) { for (int j = 0; j < 3; j++) T(i, j) = R[i][j]; T(i, 3) = t[i]; }
  return T;
}

static bool inv3(const double m[3][3], double o[3][3]) {
  const double c00 = m[1][1] * m[2][2] - m[1][2] * m[2][1], c10 = m[1][2] * m[2][0] - m[1][0] * m[2][2],
               c20 = m[1][0] * m[2][1] - m[1][1] * m[2][0];
  const double det = m[0][0] * c00 + m[0][1] * c10 + m[0][2] * c20;
  const double id = 1.0 / det;
  o[0][0] = c00 * id; o[0][1] = (m[0][2] * m[2][1] - m[0][1] * m[2][2]) * id; o[0][2] = (m[0][1] * m[1][2] - m[0][2] * m[1][1]) * id;
  o[1][0] = c10 * id; o[1][1] = (m[0][0] * m[2][2] - m[0][2] * m[2][0]) * id; o[1][2] = (m[0][2] * m[1][0] - m[0][0] * m[1][2]) * id;
  o[2][0] = c20 * id; o[2][1] = (m[0][1] * m[2][0] - m[0][0] * m[2][1]) * id; o[2][2] = (m[0][0] * m[1][1] - m[0][1] * m[1][0]) * id;
  return det != 0.0;
}

Vector6d logmap_se3(const Matrix4d& T) {
  double w[3] = {0, 0, 0}, V[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  double cosine = (T(0, 0) + T(1, 1) + T(2, 2) - 1.0) / 2.0;
  cosine = std::fmin(1.0, std::fmax(-1.0, cosine));
  double sine = std::sqrt(1.0 - cosine * cosine);
  const double theta = std::acos(cosine);
  if (theta > 0.000001) {
    w[0] = theta * (T(2, 1) - T(1, 2)) / (2.0 * sine);
    w[1] = theta * (T(0, 2) - T(2, 0)) / (2.0 * sine);
    w[2] = theta * (T(1, 0) - T(0, 1)) / (2.0 * sine);
    double s[3][3], ss[3][3];
    skew(w, s);
    for (auto& r : s) for (double& v : r) v /= theta;
    for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { ss[i][j] = 0; for (int k = 0; k < 3; k++) ss[i][j] += s[i][k] * s[k][j]; }
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++)
        V[i][j] = (i == j ? 1.0 : 0.0) + s[i][j] * (1.0 - cosine) / theta + ss[i][j] * (theta - sine) / theta;
  }
  double Vi[3][3];
  inv3(V, Vi);
  Vector6d x;
  for (int i = 0; i < 3; i++) x(i) = Vi[i][0] * T(0, 3) + Vi[i][1] * T(1, 3) + Vi[i][2] * T(2, 3);
  x(3) = w[0]; x(4) = w[1]; x(5) = w[2];
  return x;
}

// ---- lookForCommonMatches, KF <-> KF half (mapHandler.cpp:189-490) ---------------------------------------------
static Vector3d transform_point(const Matrix4d& T, const Vector3d& P) {
  Vector3d r;
  for (int i = 0; i < 3; i++) r(i) = T(i, 0) * P(0) + T(i, 1) * P(1) + T(i, 2) * P(2) + T(i, 3);
  return r;
}
static Vector3d normalized(const Vector3d& v) {
  const double n = std::sqrt(v(0) * v(0) + v(1) * v(1) + v(2) * v(2));
  Vector3d r;
  for (int i = 0; i < 3; i++) r(i) = v(i) / n;
  return r;
}

void MapHandler::lookForCommonMatches(KeyFrame* kf0, KeyFrame*& kf1) {
  const int kf0_idx = kf0->kf_idx, kf1_idx = kf1->kf_idx;
  const Matrix4d DT = inverse_se3(kf1->T_kf_w) * kf0->T_kf_w;   // :197
  StVO::StereoFrame* f0 = kf0->stereo_frame;
  StVO::StereoFrame* f1 = kf1->stereo_frame;
  const int need = std::max(kf0_idx, kf1_idx) + 1;
  if ((int)full_graph.size() < need) {
    full_graph.resize(need);
    map_points_kf_idx.resize(need);
    map_lines_kf_idx.resize(need);
  }
  for (auto& row : full_graph) row.resize(full_graph.size(), 0);

  // ---- matching + gates on the GPU ------------------------------------------------------------------
  vector<gfpl_point_feature> p0, p1;
  vector<gfpl_line_feature> l0, l1;
  f0->exportStereo(p0, l0);
  f1->exportStereo(p1, l1);
  int32_t n0[2] = {(int32_t)f0->stereo_pt.size(), (int32_t)f0->stereo_ls.size()};
  int32_t n1[2] = {(int32_t)f1->stereo_pt.size(), (int32_t)f1->stereo_ls.size()};
  vector<uint8_t> pd0(p0.size() * 32), ld0(l0.size() * 32), pd1(p1.size() * 32), ld1(l1.size() * 32);
  if (n0[0]) std::memcpy(pd0.data(), f0->pdesc_l.ptr(), (size_t)n0[0] * 32);
  if (n0[1]) std::memcpy(ld0.data(), f0->ldesc_l.ptr(), (size_t)n0[1] * 32);
  if (n1[0]) std::memcpy(pd1.data(), f1->pdesc_l.ptr(), (size_t)n1[0] * 32);
  if (n1[1]) std::memcpy(ld1.data(), f1->ldesc_l.ptr(), (size_t)n1[1] * 32);
  gfpl_stereo_frames v0{1, (int32_t)p0.size(), (int32_t)l0.size(), &n0[0], &n0[1], p0.data(), pd0.data(), l0.data(), ld0.data()};
  gfpl_stereo_frames v1{1, (int32_t)p1.size(), (int32_t)l1.size(), &n1[0], &n1[1], p1.data(), pd1.data(), l1.data(), ld1.data()};
  vector<gfpl_pt_obs> opt(p0.size());
  vector<gfpl_ln_obs> oln(l0.size());
  vector<uint8_t> pin(p0.size()), lin(l0.size());
  int32_t on[2] = {0, 0};
  gfpl_track_obs obs{1, (int32_t)p0.size(), (int32_t)l0.size(), &on[0], &on[1], opt.data(), pin.data(), oln.data(), lin.data()};
  double DTr[16];
  DT.to_row_major(DTr);
  gfpl->check(gfpl_kf_match(gfpl->get(), &v0, &v1, DTr, &obs));

  // ---- landmark / graph bookkeeping in the reference's loop order (ascending kf0 index) -------------------------
  int common_pt = 0;
  for (int k = 0; k < on[0]; k++) {  // :268-334
    const int lr_qdx = opt[k].prev_idx, lr_tdx = opt[k].curr_idx;
    StVO::PointFeature* a = f0->stereo_pt[lr_qdx];
    StVO::PointFeature* b = f1->stereo_pt[lr_tdx];
    common_pt++;
    if (a->idx == -1) {
      a->idx = max_pt_idx;
      b->idx = max_pt_idx;
      MapPoint* mp = new MapPoint();
      mp->idx = max_pt_idx;
      mp->point3D = transform_point(kf0->T_kf_w, a->P);
      mp->addMapPointObservation(f0->pdesc_l.ptr(lr_qdx), kf0->kf_idx, a->pl, normalized(mp->point3D));
      map_points_kf_idx.at(kf0->kf_idx).push_back(max_pt_idx);
      mp->addMapPointObservation(f1->pdesc_l.ptr(lr_tdx), kf1->kf_idx, b->pl, normalized(transform_point(kf1->T_kf_w, b->P)));
      if ((int)map_points.size() <= max_pt_idx) map_points.resize(max_pt_idx + 1, nullptr);
      map_points[max_pt_idx] = mp;
      max_pt_idx++;
      full_graph[kf1_idx][kf0_idx]++;
      full_graph[kf0_idx][kf1_idx]++;
    } else {
      const int lm_idx = a->idx;
      if (lm_idx < (int)map_points.size() && map_points[lm_idx] != nullptr) {
        b->idx = lm_idx;
        map_points[lm_idx]->addMapPointObservation(f1->pdesc_l.ptr(lr_tdx), kf1->kf_idx, b->pl,
                                                   normalized(transform_point(kf1->T_kf_w, b->P)));
        for (int o : map_points[lm_idx]->kf_obs_list)
          if (o != kf1_idx && o < (int)full_graph.size()) { full_graph[kf1_idx][o]++; full_graph[o][kf1_idx]++; }
      }
    }
  }
  int common_ls = 0;
  for (int k = 0; k < on[1]; k++) {  // :405-487
    const int lr_qdx = oln[k].prev_idx, lr_tdx = oln[k].curr_idx;
    StVO::LineFeature* a = f0->stereo_ls[lr_qdx];
    StVO::LineFeature* b = f1->stereo_ls[lr_tdx];
    common_ls++;
    auto mid_dir = [&](const Matrix4d& T, const StVO::LineFeature* l) {
      Vector3d m;
      for (int i = 0; i < 3; i++) m(i) = 0.5 * (l->sP(i) + l->eP(i));
      return normalized(transform_point(T, m));
    };
    auto pts_of = [](const StVO::LineFeature* l) {
      gfpl::Mat<4, 1> p;
      p(0) = l->spl(0); p(1) = l->spl(1); p(2) = l->epl(0); p(3) = l->epl(1);
      return p;
    };
    if (a->idx == -1) {
      a->idx = max_ls_idx;
      b->idx = max_ls_idx;
      MapLine* ml = new MapLine();
      ml->idx = max_ls_idx;
      const Vector3d s3 = transform_point(kf0->T_kf_w, a->sP), e3 = transform_point(kf0->T_kf_w, a->eP);
      for (int i = 0; i < 3; i++) { ml->line3D(i) = s3(i); ml->line3D(3 + i) = e3(i); }
      Vector3d m;
      for (int i = 0; i < 3; i++) m(i) = 0.5 * (s3(i) + e3(i));
      ml->addMapLineObservation(f0->ldesc_l.ptr(lr_qdx), kf0->kf_idx, a->le, normalized(m), pts_of(a));
      ml->addMapLineObservation(f1->ldesc_l.ptr(lr_tdx), kf1->kf_idx, b->le, mid_dir(kf1->T_kf_w, b), pts_of(b));
      map_lines_kf_idx.at(kf0->kf_idx).push_back(max_ls_idx);
      if ((int)map_lines.size() <= max_ls_idx) map_lines.resize(max_ls_idx + 1, nullptr);
      map_lines[max_ls_idx] = ml;
      max_ls_idx++;
      full_graph[kf1_idx][kf0_idx]++;
      full_graph[kf0_idx][kf1_idx]++;
    } else {
      const int lm_idx = a->idx;
      if (lm_idx < (int)map_lines.size() && map_lines[lm_idx] != nullptr) {
        b->idx = lm_idx;
        map_lines[lm_idx]->addMapLineObservation(f1->ldesc_l.ptr(lr_tdx), kf1->kf_idx, b->le, mid_dir(kf1->T_kf_w, b), pts_of(b));
        for (int o : map_lines[lm_idx]->kf_obs_list)
          if (o != kf1_idx && o < (int)full_graph.size()) { full_graph[kf1_idx][o]++; full_graph[o][kf1_idx]++; }
      }
    }
  }
  {
    vector<int> tp, tl;
    for (int k = 0; k < on[0]; k++) tp.push_back(f0->stereo_pt[opt[k].prev_idx]->idx);
    for (int k = 0; k < on[1]; k++) tl.push_back(f0->stereo_ls[oln[k].prev_idx]->idx);
    refreshMedoids(tp, tl);
  }
  matchLocalMap(kf1, common_pt, common_ls);
  last_common_pt = common_pt;
  last_common_ls = common_ls;
}

// ---- lookForCommonMatches, local map <-> KF half (mapHandler.cpp:509-770) ------------------------------------------
void MapHandler::matchLocalMap(KeyFrame* kf1, int& common_pt, int& common_ls) {
  const int kf1_idx = kf1->kf_idx;
  StVO::StereoFrame* f1 = kf1->stereo_frame;
  const Matrix4d Twf = inverse_se3(kf1->T_kf_w);  // :200
  // host walk over the map: `local` and "last observer is not kf1" (:516-521, :636-643); the in-image test runs on the GPU
  vector<MapPoint*> cand_pt;
  vector<MapLine*> cand_ls;
  for (MapPoint* mp : map_points)
    if (mp && mp->local && mp->kf_obs_list.back() != kf1_idx) cand_pt.push_back(mp);
  for (MapLine* ml : map_lines)
    if (ml && ml->local && ml->kf_obs_list.back() != kf1_idx) cand_ls.push_back(ml);
  const int32_t np = (int32_t)cand_pt.size(), nl = (int32_t)cand_ls.size();
  const int32_t cp = std::max(np, 1), cl = std::max(nl, 1);
  vector<double> X((size_t)cp * 3, 0.0), L((size_t)cl * 6, 0.0);
  vector<uint8_t> dX((size_t)cp * 32, 0), dL((size_t)cl * 32, 0);
  for (int i = 0; i < np; i++) {
    for (int k = 0; k < 3; k++) X[3 * i + k] = cand_pt[i]->point3D(k);
    std::memcpy(&dX[32 * (size_t)i], cand_pt[i]->med_desc.data(), 32);
  }
  for (int i = 0; i < nl; i++) {
    for (int k = 0; k < 6; k++) L[6 * i + k] = cand_ls[i]->line3D(k);
    std::memcpy(&dL[32 * (size_t)i], cand_ls[i]->med_desc.data(), 32);
  }
  vector<gfpl_point_feature> p1;
  vector<gfpl_line_feature> l1;
  f1->exportStereo(p1, l1);
  int32_t n1[2] = {(int32_t)f1->stereo_pt.size(), (int32_t)f1->stereo_ls.size()};
  if (p1.empty()) p1.resize(1);
  if (l1.empty()) l1.resize(1);
  vector<uint8_t> pd1(p1.size() * 32, 0), ld1(l1.size() * 32, 0);
  if (n1[0]) std::memcpy(pd1.data(), f1->pdesc_l.ptr(), (size_t)n1[0] * 32);
  if (n1[1]) std::memcpy(ld1.data(), f1->ldesc_l.ptr(), (size_t)n1[1] * 32);
  gfpl_local_map lm{1, cp, cl, &np, &nl, X.data(), dX.data(), L.data(), dL.data()};
  gfpl_stereo_frames v1{1, (int32_t)p1.size(), (int32_t)l1.size(), &n1[0], &n1[1], p1.data(), pd1.data(), l1.data(), ld1.data()};
  vector<gfpl_map_pair> mpt(cp), mln(cl);
  int32_t mn[2] = {0, 0};
  gfpl_map_matches mm{1, cp, cl, &mn[0], &mn[1], mpt.data(), mln.data()};
  double Tr[16];
  Twf.to_row_major(Tr);
  gfpl->check(gfpl_map_match(gfpl->get(), &lm, &v1, Tr, &mm));

  // position of every stereo line inside the reference's `unmatched_lines` list (built before the loop, :663-674)
  vector<int> un_pos(f1->stereo_ls.size(), -1);
  {
    int c = 0;
    for (size_t i = 0; i < f1->stereo_ls.size(); i++)
      if (f1->stereo_ls[i]->idx == -1) un_pos[i] = c++;
  }
  auto bump_graph = [&](const vector<int>& kf_obs_list) {
    for (int o : kf_obs_list)
      if (o != kf1_idx && o < (int)full_graph.size()) { full_graph[kf1_idx][o]++; full_graph[o][kf1_idx]++; }
  };
  vector<int> tp, tl;
  for (int k = 0; k < mn[0]; k++) {  // :606-625
    MapPoint* mp = cand_pt[mpt[k].map_idx];
    StVO::PointFeature* b = f1->stereo_pt[mpt[k].kf_idx];
    common_pt++;
    b->idx = mp->idx;
    const Vector3d dir_kf = transform_point(kf1->T_kf_w, normalized(b->P));   // R * dir + t as written at :613
    mp->addMapPointObservation(f1->pdesc_l.ptr(mpt[k].kf_idx), kf1_idx, b->pl, dir_kf);
    bump_graph(mp->kf_obs_list);
    tp.push_back(mp->idx);
  }
  for (int k = 0; k < mn[1]; k++) {  // :745-765
    MapLine* ml = cand_ls[mln[k].map_idx];
    StVO::LineFeature* b = f1->stereo_ls[mln[k].kf_idx];
    common_ls++;
    b->idx = ml->idx;
    Vector3d m;
    for (int i = 0; i < 3; i++) m(i) = 0.5 * (b->sP(i) + b->eP(i));
    const Vector3d mP3d = normalized(transform_point(kf1->T_kf_w, m));
    gfpl::Mat<4, 1> pts;
    pts(0) = b->spl(0); pts(1) = b->spl(1); pts(2) = b->epl(0); pts(3) = b->epl(1);
    // the reference stores row `lr_tdx` of the FULL ldesc_l, lr_tdx being the position in the unmatched list (:755)
    ml->addMapLineObservation(f1->ldesc_l.ptr(un_pos[mln[k].kf_idx]), kf1_idx, b->le, mP3d, pts);
    bump_graph(ml->kf_obs_list);
    tl.push_back(ml->idx);
  }
  refreshMedoids(tp, tl);
}

// ---- updateAverageDescDir for a set of landmarks (mapFeatures.cpp:51-92, 128-170) ----------------------------------
void MapHandler::refreshMedoids(const vector<int>& pts, const vector<int>& lines) {
  auto run = [&](auto& landmarks, const vector<int>& ids) {
    vector<uint8_t> desc;
    vector<int32_t> off{0};
    vector<int> live;
    for (int id : ids) {
      if (id < 0 || id >= (int)landmarks.size() || !landmarks[id]) continue;
      for (const Descriptor& d : landmarks[id]->desc_list) desc.insert(desc.end(), d.begin(), d.end());
      off.push_back((int32_t)(desc.size() / 32));
      live.push_back(id);
    }
    if (live.empty()) return;
    vector<int32_t> med(live.size());
    gfpl->check(gfpl_medoid_desc(gfpl->get(), desc.data(), off.data(), (int32_t)live.size(), med.data()));
    for (size_t k = 0; k < live.size(); k++) {
      auto* lm = landmarks[live[k]];
      lm->med_desc = lm->desc_list[med[k]];
      Vector3d s;                                   // the reference sums into an uninitialised Vector3d (:85-88): 0 here
      for (int i = 0; i < 3; i++) s(i) = 0.0;
      for (const Vector3d& d : lm->dir_list) for (int i = 0; i < 3; i++) s(i) += d(i);
      const double n = (double)lm->dir_list.size();
      for (int i = 0; i < 3; i++) lm->med_obs_dir(i) = s(i) / n;
    }
  };
  run(map_points, pts);
  run(map_lines, lines);
}

// ---- localBundleAdjustment: gather (mapHandler.cpp:1108-1215) ------------------------------------------------------
void MapHandler::localBundleAdjustment() {
  vector<double> X_aux;
  vector<int> kf_list, pt_list, ls_list;
  vector<Vector6i> pt_obs_list, ls_obs_list;
  for (KeyFrame* kf : map_keyframes)
    if (kf && kf->local && kf->kf_idx != 0) {
      for (int i = 0; i < 6; i++) X_aux.push_back(kf->x_kf_w(i));
      kf_list.push_back(kf->kf_idx);
    }
  auto local_kf = [&](int kf_map) {
    for (size_t j = 0; j < kf_list.size(); j++) if (kf_list[j] == kf_map) return (int)j;
    return -1;
  };
  int lm_local_idx = 0;
  for (MapPoint* pt : map_points)
    if (pt && pt->local) {
      for (int i = 0; i < 3; i++) X_aux.push_back(pt->point3D(i));
      for (size_t i = 0; i < pt->obs_list.size(); i++)
        pt_obs_list.push_back(Vector6i{pt->idx, lm_local_idx, (int)i, pt->kf_obs_list[i], local_kf(pt->kf_obs_list[i]), 1});
      lm_local_idx++;
      pt_list.push_back(pt->idx);
    }
  lm_local_idx = 0;
  for (MapLine* ls : map_lines)
    if (ls && ls->local) {
      for (int i = 0; i < 6; i++) X_aux.push_back(ls->line3D(i));
      for (size_t i = 0; i < ls->obs_list.size(); i++)
        ls_obs_list.push_back(Vector6i{ls->idx, lm_local_idx, (int)i, ls->kf_obs_list[i], local_kf(ls->kf_obs_list[i]), 1});
      lm_local_idx++;
      ls_list.push_back(ls->idx);
    }
  if (pt_obs_list.size() + ls_obs_list.size() != 0)
    levMarquardtOptimizationLBA(X_aux, kf_list, pt_list, ls_list, pt_obs_list, ls_obs_list);
}

// ---- levMarquardtOptimizationLBA (mapHandler.cpp:1217-1712) --------------------------------------------------------
void MapHandler::levMarquardtOptimizationLBA(vector<double> X_aux, vector<int> kf_list, vector<int> pt_list,
                                             vector<int> ls_list, vector<Vector6i> pt_obs_list,
                                             vector<Vector6i> ls_obs_list) {
  const gfpl_config& cfg = gfpl->config();
  const int Nkf = (int)kf_list.size(), N = (int)X_aux.size();
  const int Npt = pt_obs_list.empty() ? 0 : pt_obs_list.back()[1] + 1;   // :1241-1242
  const int Nls = ls_obs_list.empty() ? 0 : ls_obs_list.back()[1] + 1;   // :1320-1321
  const int npo = (int)pt_obs_list.size(), nlo = (int)ls_obs_list.size();
  vector<double> X = X_aux, DX(N, 0.0);

  // flat, pass-independent observation arrays
  vector<int32_t> pt_lm(npo), pt_kf(npo), pt_fix(npo), ls_lm(nlo), ls_kf(nlo), ls_fix(nlo);
  vector<double> pt_uv(2 * (size_t)npo), pt_s2(npo), ls_le(3 * (size_t)nlo), ls_s2(nlo), fix_T;
  vector<int> fix_ids;
  auto fix_slot = [&](int kf_map) {
    for (size_t i = 0; i < fix_ids.size(); i++) if (fix_ids[i] == kf_map) return (int)i;
    fix_ids.push_back(kf_map);
    double T[16];
    map_keyframes[kf_map]->T_kf_w.to_row_major(T);
    fix_T.insert(fix_T.end(), T, T + 16);
    return (int)fix_ids.size() - 1;
  };
  for (int o = 0; o < npo; o++) {
    const Vector6i& ob = pt_obs_list[o];
    const MapPoint* mp = map_points[ob[0]];
    pt_lm[o] = ob[1]; pt_kf[o] = ob[4]; pt_fix[o] = ob[4] < 0 ? fix_slot(ob[3]) : 0;
    pt_uv[2 * o] = mp->obs_list[ob[2]](0); pt_uv[2 * o + 1] = mp->obs_list[ob[2]](1);
    pt_s2[o] = mp->sigma_list[ob[2]];
  }
  for (int o = 0; o < nlo; o++) {
    const Vector6i& ob = ls_obs_list[o];
    const MapLine* ml = map_lines[ob[0]];
    ls_lm[o] = ob[1]; ls_kf[o] = ob[4]; ls_fix[o] = ob[4] < 0 ? fix_slot(ob[3]) : 0;
    for (int i = 0; i < 3; i++) ls_le[3 * o + i] = ml->obs_list[ob[2]](i);
    ls_s2[o] = ml->sigma_list[ob[2]];
  }
  if (fix_T.empty()) fix_T.assign(16, 0.0);
  // quirk (i) of SURVEY.md §8a L2: line landmarks are always linearised at the MAP values (:1555-1558)
  vector<double> X_ls_map(6 * (size_t)std::max(Nls, 1), 0.0);
  for (int j = 0; j < Nls; j++)
    for (int i = 0; i < 6; i++) X_ls_map[6 * j + i] = map_lines[ls_list[j]]->line3D(i);

  vector<double> kf_T(16 * (size_t)std::max(Nkf, 1)), X_pt(3 * (size_t)std::max(Npt, 1));
  double err_sum = 0.0;

  // one LM pass on the device: assembly at the current state + damped solve (gfpl_lba_step).  DX is computed in every
  // pass, also in the one where the reference leaves the loop before solving (:1652-1653) — it is simply not used then.
  double lambda_eff = 0.0;
  auto step = [&](bool first_pass, double lambda, int scale_by_hmax) {
    for (int k = 0; k < Nkf; k++) {
      Matrix4d T;
      if (first_pass) T = map_keyframes[kf_list[k]]->T_kf_w;   // :1255
      else { Vector6d x; for (int i = 0; i < 6; i++) x(i) = X[6 * k + i]; T = expmap_se3(x); }   // :1480
      T.to_row_major(&kf_T[16 * (size_t)k]);
    }
    for (int j = 0; j < Npt; j++)
      for (int i = 0; i < 3; i++)
        X_pt[3 * j + i] = first_pass ? map_points[pt_list[j]]->point3D(i) : X[6 * Nkf + 3 * j + i];   // :1253 / :1476
    gfpl_lba_problem pb{Nkf, Npt, Nls, npo, nlo, kf_T.data(), X_pt.data(), X_ls_map.data(), pt_lm.data(), pt_kf.data(),
                        pt_fix.data(), pt_uv.data(), pt_s2.data(), ls_lm.data(), ls_kf.data(), ls_fix.data(), ls_le.data(),
                        ls_s2.data(), fix_T.data(), (int32_t)fix_ids.size()};
    gfpl->check(gfpl_lba_step(gfpl->get(), &pb, lambda, scale_by_hmax, DX.data(), &err_sum, &lambda_eff));
  };

  auto update = [&]() {  // :1443-1451
    for (int i = 0; i < Nkf; i++) {
      Vector6d x, dx;
      for (int k = 0; k < 6; k++) { x(k) = X[6 * i + k]; dx(k) = DX[6 * i + k]; }
      const Matrix4d Tcurr = expmap_se3(x) * inverse_se3(expmap_se3(dx));
      const Vector6d xn = logmap_se3(Tcurr);
      for (int k = 0; k < 6; k++) X[6 * i + k] = xn(k);
    }
    for (int i = 6 * Nkf; i < N; i++) X[i] += DX[i];
  };

  double err, err_prev;
  double lambda = cfg.lambda_lba_lm;
  const double lambda_k = cfg.lambda_lba_k;
  const double eps = std::numeric_limits<double>::epsilon();
  step(true, lambda, 1);                     // :1243-1441; lambda *= (int)Hmax (quirk iii) happens inside
  err = err_sum / 0.0;                       // quirk (ii): Npt_obs + Nls_obs are never incremented (:1427)
  lambda = lambda_eff;
  update();
  err_prev = err;
  int iters;
  for (iters = 1; iters < cfg.max_iters_lba; iters++) {
    step(false, lambda, 0);
    err = err_sum / (Npt + Nls);             // :1650
    if (std::fabs(err - err_prev) < eps || err < eps) break;
    if (err > err_prev) lambda /= lambda_k;
    else { lambda *= lambda_k; update(); }
    double n2 = 0.0;
    for (double v : DX) n2 += v * v;
    if (std::sqrt(n2) < eps) break;
    err_prev = err;
  }
  lba_iterations = iters;
  lba_last_err = err;
  // write back (:1691-1712); x_kf_w is NOT refreshed (quirk vi)
  for (int i = 0; i < Nkf; i++) {
    Vector6d x;
    for (int k = 0; k < 6; k++) x(k) = X[6 * i + k];
    map_keyframes[kf_list[i]]->T_kf_w = expmap_se3(x);
  }
  for (int i = 0; i < Npt; i++) for (int k = 0; k < 3; k++) map_points[pt_list[i]]->point3D(k) = X[6 * Nkf + 3 * i + k];
  for (int i = 0; i < Nls; i++) for (int k = 0; k < 6; k++) map_lines[ls_list[i]]->line3D(k) = X[6 * Nkf + 3 * Npt + 6 * i + k];
}

}  // namespace PLSLAM
