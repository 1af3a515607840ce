// stereoFrameHandler.cpp — StVO::StereoFrameHandler facade over the C ABI (see stereoFrameHandler.hpp).
#include "stereoFrameHandler.hpp"

#include <cmath>

namespace StVO {

using std::vector;

StereoFrameHandler::StereoFrameHandler(PinholeStereoCamera* cam_, GfplContext* gfpl) : cam(cam_), gfpl_(gfpl) {}
StereoFrameHandler::~StereoFrameHandler() {}

void StereoFrameHandler::initialize(StereoFrame* first) {
  prev_frame = first;
  prev_frame->extractInitialStereoFeatures();
  prev_frame->Tfw = Matrix4d::Identity();
  prev_frame->Tfw_cov = Matrix6d::Identity();
  prev_frame->DT = Matrix4d::Identity();
}

void StereoFrameHandler::predictFramePose() { curr_frame->Tfw = prev_frame->Tfw * prev_frame->DT; }

void StereoFrameHandler::insertStereoPair(StereoFrame* next) {
  curr_frame = next;
  curr_frame->extractStereoFeatures_Brute();   // the brute-force stereo stage (:90); detection was injected
  predictFramePose();
  if (gfpl_->config().use_line_conf_cut) prev_frame->estimateStereoUncertainty();   // :103-106
  crossFrameMatching_Hybrid();                                                      // :109
  if (gfpl_->config().use_line_conf_cut) {                                          // :120-135
    double rngCutRatio[2] = {gfpl_->config().cut_rng[0], gfpl_->config().cut_rng[1]};
    estimateProjUncertainty_submodular(gfpl_->config().cut_step, rngCutRatio);
  }
}

void StereoFrameHandler::updateFrame() {
  matched_pt.clear();
  matched_ls.clear();
  delete prev_frame;
  prev_frame = curr_frame;
  curr_frame = nullptr;
}

gfpl_frame_state StereoFrameHandler::prevState() const {
  gfpl_frame_state st{};
  prev_frame->Tfw.to_row_major(st.Tfw);
  prev_frame->DT.to_row_major(st.DT);
  prev_frame->DT_cov.to_row_major(st.DT_cov);
  prev_frame->Tfw_cov.to_row_major(st.Tfw_cov);
  st.time_stamp = prev_frame->time_stamp;
  st.time_stamp_curr = curr_frame ? curr_frame->time_stamp : prev_frame->time_stamp;
  return st;
}

static gfpl_stereo_frames view(vector<gfpl_point_feature>& pt, vector<gfpl_line_feature>& ln, vector<uint8_t>& pd,
                               vector<uint8_t>& ld, int32_t* n_pt, int32_t* n_ln, const StereoFrame* f) {
  f->exportStereo(pt, ln);
  *n_pt = (int32_t)f->stereo_pt.size();
  *n_ln = (int32_t)f->stereo_ls.size();
  pd.assign(pt.size() * 32, 0);
  ld.assign(ln.size() * 32, 0);
  if (*n_pt) std::memcpy(pd.data(), f->pdesc_l.ptr(), (size_t)*n_pt * 32);
  if (*n_ln) std::memcpy(ld.data(), f->ldesc_l.ptr(), (size_t)*n_ln * 32);
  return gfpl_stereo_frames{1, (int32_t)pt.size(), (int32_t)ln.size(), n_pt, n_ln, pt.data(), pd.data(), ln.data(), ld.data()};
}

void StereoFrameHandler::crossFrameMatching(int mode) {
  matched_pt.clear();
  matched_ls.clear();
  matched_ls_prev_idx_.clear();
  vector<gfpl_point_feature> ppt, cpt;
  vector<gfpl_line_feature> pln, cln;
  vector<uint8_t> ppd, pld, cpd, cld;
  int32_t pn_pt, pn_ln, cn_pt, cn_ln;
  gfpl_stereo_frames pv = view(ppt, pln, ppd, pld, &pn_pt, &pn_ln, prev_frame);
  gfpl_stereo_frames cv = view(cpt, cln, cpd, cld, &cn_pt, &cn_ln, curr_frame);
  gfpl_frame_state st = prevState();
  const int cap_pt = gfpl_->config().max_point_match_num, cap_ln = gfpl_->config().max_line_match_num;
  vector<gfpl_pt_obs> opt(cap_pt);
  vector<gfpl_ln_obs> oln(cap_ln);
  vector<uint8_t> pin(cap_pt), lin(cap_ln);
  int32_t n_pt = 0, n_ln = 0;
  gfpl_track_obs obs{1, cap_pt, cap_ln, &n_pt, &n_ln, opt.data(), pin.data(), oln.data(), lin.data()};
  gfpl_->check(gfpl_f2f_match(gfpl_->get(), &pv, &cv, &st, mode, &obs));
  for (size_t i = 0; i < prev_frame->stereo_pt.size(); i++) prev_frame->stereo_pt[i]->frame_matched = false;
  for (size_t i = 0; i < curr_frame->stereo_pt.size(); i++) curr_frame->stereo_pt[i]->frame_matched = false;
  for (int k = 0; k < n_pt; k++) {  // stereoFrameHandler.cpp:576-586
    PointFeature* point_ = prev_frame->stereo_pt[opt[k].prev_idx];
    point_->pl_obs(0) = opt[k].pl_obs[0]; point_->pl_obs(1) = opt[k].pl_obs[1];
    point_->inlier = true;
    matched_pt.push_back(point_);
    curr_frame->stereo_pt[opt[k].curr_idx]->idx = point_->idx;
    point_->frame_matched = true;
    curr_frame->stereo_pt[opt[k].curr_idx]->frame_matched = true;
  }
  for (int k = 0; k < n_ln; k++) {  // :666-674
    LineFeature* line_ = prev_frame->stereo_ls[oln[k].prev_idx];
    const LineFeature* c = curr_frame->stereo_ls[oln[k].curr_idx];
    line_->sdisp_obs = c->sdisp; line_->edisp_obs = c->edisp;
    line_->spl_obs = c->spl; line_->epl_obs = c->epl; line_->le_obs = c->le;
    line_->inlier = true;
    matched_ls.push_back(line_);
    matched_ls_prev_idx_.push_back(oln[k].prev_idx);
    curr_frame->stereo_ls[oln[k].curr_idx]->idx = line_->idx;
  }
  n_inliers_pt = (int)matched_pt.size();
  n_inliers_ls = (int)matched_ls.size();
  n_inliers = n_inliers_pt + n_inliers_ls;
}

void StereoFrameHandler::crossFrameMatching_Hybrid() { crossFrameMatching(GFPL_F2F_HYBRID); }
void StereoFrameHandler::crossFrameMatching_Brute() { crossFrameMatching(GFPL_F2F_BRUTE); }

void StereoFrameHandler::buildObs(vector<gfpl_pt_obs>& pt, vector<uint8_t>& pin, vector<gfpl_ln_obs>& ln,
                                  vector<uint8_t>& lin) const {
  pt.assign(std::max<size_t>(matched_pt.size(), 1), gfpl_pt_obs{});
  ln.assign(std::max<size_t>(matched_ls.size(), 1), gfpl_ln_obs{});
  pin.assign(pt.size(), 0);
  lin.assign(ln.size(), 0);
  size_t k = 0;
  for (auto it = matched_pt.begin(); it != matched_pt.end(); ++it, ++k) {
    for (int i = 0; i < 3; i++) pt[k].P[i] = (*it)->P(i);
    pt[k].pl_obs[0] = (*it)->pl_obs(0); pt[k].pl_obs[1] = (*it)->pl_obs(1);
    pt[k].sigma2 = (*it)->sigma2;
    pin[k] = (*it)->inlier ? 1 : 0;
  }
  k = 0;
  for (auto it = matched_ls.begin(); it != matched_ls.end(); ++it, ++k) {
    for (int i = 0; i < 3; i++) { ln[k].sP[i] = (*it)->sP(i); ln[k].eP[i] = (*it)->eP(i); ln[k].le_obs[i] = (*it)->le_obs(i); }
    ln[k].sigma2 = (*it)->sigma2;
    ln[k].prev_idx = (int32_t)k;   // covariances are passed in matched order
    lin[k] = (*it)->inlier ? 1 : 0;
  }
}

void StereoFrameHandler::optimizeFunctions(Matrix4d DT, Matrix6d& H, Vector6d& g, double& e) {
  vector<gfpl_pt_obs> pt; vector<gfpl_ln_obs> ln; vector<uint8_t> pin, lin;
  buildObs(pt, pin, ln, lin);
  int32_t n_pt = (int32_t)matched_pt.size(), n_ln = (int32_t)matched_ls.size();
  gfpl_track_obs obs{1, (int32_t)pt.size(), (int32_t)ln.size(), &n_pt, &n_ln, pt.data(), pin.data(), ln.data(), lin.data()};
  double DTr[16], Hr[36], gr[6];
  DT.to_row_major(DTr);
  gfpl_->check(gfpl_pose_hg(gfpl_->get(), DTr, &obs, Hr, gr, &e));
  H.from_row_major(Hr);
  for (int i = 0; i < 6; i++) g(i) = gr[i];
}

void StereoFrameHandler::optimizePoseImpl(const Matrix4d* DT_ini) {
  vector<gfpl_pt_obs> pt; vector<gfpl_ln_obs> ln; vector<uint8_t> pin, lin;
  buildObs(pt, pin, ln, lin);
  int32_t n_pt = (int32_t)matched_pt.size(), n_ln = (int32_t)matched_ls.size();
  gfpl_track_obs obs{1, (int32_t)pt.size(), (int32_t)ln.size(), &n_pt, &n_ln, pt.data(), pin.data(), ln.data(), lin.data()};
  gfpl_frame_state st = prevState();
  double ini[16];
  if (DT_ini) DT_ini->to_row_major(ini);
  gfpl_pose_result r{};
  gfpl_->check(gfpl_pose_optimize(gfpl_->get(), &obs, &st, DT_ini ? ini : nullptr,
                                  DT_ini ? GFPL_POSE_SEED_GIVEN : GFPL_POSE_SEED_PREV_DT, &r));
  size_t k = 0;
  for (auto it = matched_pt.begin(); it != matched_pt.end(); ++it, ++k) (*it)->inlier = pin[k] != 0;   // :2096-2104
  k = 0;
  for (auto it = matched_ls.begin(); it != matched_ls.end(); ++it, ++k) (*it)->inlier = lin[k] != 0;   // :2106-2114
  n_inliers = r.n_inliers; n_inliers_pt = r.n_inliers_pt; n_inliers_ls = r.n_inliers_ls;
  curr_frame->DT.from_row_major(r.DT);
  curr_frame->Tfw.from_row_major(r.Tfw);
  curr_frame->DT_cov.from_row_major(r.DT_cov);
  curr_frame->Tfw_cov.from_row_major(r.Tfw_cov);
  for (int i = 0; i < 6; i++) curr_frame->DT_cov_eig(i) = r.DT_cov_eig[i];
  curr_frame->err_norm = r.err_norm;
  last_status = r.status;
  if (DT_ini) {  // numFrameLoss bookkeeping of the second overload (:2014,2027)
    if (r.status & GFPL_POSE_NONFINITE_2) numFrameLoss++;
    else numFrameLoss = 0;
  }
}

void StereoFrameHandler::optimizePose() { optimizePoseImpl(nullptr); }
void StereoFrameHandler::optimizePose(Matrix4d DT_ini) { optimizePoseImpl(&DT_ini); }

static void cov_of(const LineFeature* l, gfpl_line_cov& c) {
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) { c.covS[i * 3 + j] = l->covSpt3D(i, j); c.covE[i * 3 + j] = l->covEpt3D(i, j); }
}

static void dt_inv_of(const StereoFrame* prev, const StereoFrame* curr, double* out) {
  // Matrix4d DT_inv = curr_frame->Tfw.inverse() * prev_frame->Tfw (:1635): rigid transforms -> closed-form inverse
  double Tc[16], Tp[16], Ti[16];
  curr->Tfw.to_row_major(Tc);
  prev->Tfw.to_row_major(Tp);
  for (int i = 0; i < 16; i++) Ti[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int i = 0; i < 3; i++) {
    for (int j = 0; j < 3; j++) Ti[i * 4 + j] = Tc[j * 4 + i];
    Ti[i * 4 + 3] = -(Tc[0 * 4 + i] * Tc[3] + Tc[1 * 4 + i] * Tc[7] + Tc[2 * 4 + i] * Tc[11]);
  }
  for (int i = 0; i < 4; i++)
    for (int j = 0; j < 4; j++) {
      double s = 0;
      for (int k = 0; k < 4; k++) s += Ti[i * 4 + k] * Tp[k * 4 + j];
      out[i * 4 + j] = s;
    }
}

void StereoFrameHandler::estimateProjUncertainty_submodular(const double stepCutRatio, const double rngCutRatio[2]) {
  if (matched_ls.empty()) return;  // :1621
  gfpl_config& cfg = gfpl_->config();
  cfg.cut_step = stepCutRatio; cfg.cut_rng[0] = rngCutRatio[0]; cfg.cut_rng[1] = rngCutRatio[1];
  gfpl_->apply();
  vector<gfpl_pt_obs> pt; vector<gfpl_ln_obs> ln; vector<uint8_t> pin, lin;
  buildObs(pt, pin, ln, lin);
  int32_t n_pt = (int32_t)matched_pt.size(), n_ln = (int32_t)matched_ls.size();
  gfpl_track_obs obs{1, (int32_t)pt.size(), (int32_t)ln.size(), &n_pt, &n_ln, pt.data(), pin.data(), ln.data(), lin.data()};
  vector<gfpl_line_cov> cov(ln.size());
  size_t k = 0;
  for (auto it = matched_ls.begin(); it != matched_ls.end(); ++it, ++k) cov_of(*it, cov[k]);
  double DT_inv[16];
  dt_inv_of(prev_frame, curr_frame, DT_inv);
  vector<gfpl_line_cut> cut(ln.size());
  gfpl_->check(gfpl_linecut_submodular(gfpl_->get(), DT_inv, &obs, cov.data(), (int32_t)cov.size(), cut.data()));
  k = 0;
  for (auto it = matched_ls.begin(); it != matched_ls.end(); ++it, ++k) {
    LineFeature* l = *it;
    l->cutRatio[0] = cut[k].cutRatio[0]; l->cutRatio[1] = cut[k].cutRatio[1];
    l->invCovPose.from_row_major(cut[k].invCovPose);
    // updateEndPointByRatio's in-place edits (:1456-1469), applied exactly where the reference applies them
    if (std::fabs(l->cutRatio[0]) > 0.0001) {
      for (int i = 0; i < 3; i++) l->sP(i) = ln[k].sP[i];
      l->spl(0) = cut[k].spl[0]; l->spl(1) = cut[k].spl[1]; l->sdisp = cut[k].sdisp;
    }
    if (std::fabs(l->cutRatio[1]) > 0.0001) {
      for (int i = 0; i < 3; i++) l->eP(i) = ln[k].eP[i];
      l->epl(0) = cut[k].epl[0]; l->epl(1) = cut[k].epl[1]; l->edisp = cut[k].edisp;
    }
  }
}

// stereoFrameHandler.cpp:1476-1610: every matched line on its own (gfpl_linecut_descent: one device thread per line)
void StereoFrameHandler::estimateProjUncertainty_descent(const double stepCutRatio, const double rngCutRatio[2]) {
  if (matched_ls.empty()) return;  // :1479
  gfpl_config& cfg = gfpl_->config();
  cfg.cut_step = stepCutRatio; cfg.cut_rng[0] = rngCutRatio[0]; cfg.cut_rng[1] = rngCutRatio[1];
  gfpl_->apply();
  vector<gfpl_pt_obs> pt; vector<gfpl_ln_obs> ln; vector<uint8_t> pin, lin;
  buildObs(pt, pin, ln, lin);
  int32_t n_pt = (int32_t)matched_pt.size(), n_ln = (int32_t)matched_ls.size();
  gfpl_track_obs obs{1, (int32_t)pt.size(), (int32_t)ln.size(), &n_pt, &n_ln, pt.data(), pin.data(), ln.data(), lin.data()};
  vector<gfpl_line_cov> cov(ln.size());
  size_t k = 0;
  for (auto it = matched_ls.begin(); it != matched_ls.end(); ++it, ++k) cov_of(*it, cov[k]);
  double DT_inv[16];
  dt_inv_of(prev_frame, curr_frame, DT_inv);
  vector<gfpl_line_cut_descent> cut(ln.size());
  gfpl_->check(gfpl_linecut_descent(gfpl_->get(), DT_inv, &obs, cov.data(), (int32_t)cov.size(), cut.data()));
  k = 0;
  for (auto it = matched_ls.begin(); it != matched_ls.end(); ++it, ++k) {
    LineFeature* l = *it;
    const gfpl_line_cut_descent& c = cut[k];
    l->cutRatio[0] = c.cutRatio[0]; l->cutRatio[1] = c.cutRatio[1];
    l->volumeArr.assign(1, c.volume);              // the track's last entry; its length is c.n_volume
    l->covlPt2D.assign(4, Matrix3d());             // :1578-1588: identity, start point, end point, identity
    l->lPt2D.assign(4, Vector2d());
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        l->covlPt2D[0](i, j) = l->covlPt2D[3](i, j) = (i == j) ? 1.0 : 0.0;
        l->covlPt2D[1](i, j) = c.covlPt2D[0][i * 3 + j];
        l->covlPt2D[2](i, j) = c.covlPt2D[1][i * 3 + j];
      }
    for (int q = 0; q < 4; q++) { l->lPt2D[q](0) = c.lPt2D[q][0]; l->lPt2D[q](1) = c.lPt2D[q][1]; }
    l->fst_idx = c.fst_idx; l->lst_idx = c.lst_idx;
    for (int i = 0; i < 3; i++) { l->sP(i) = ln[k].sP[i]; l->eP(i) = ln[k].eP[i]; }   // :1601-1602, unconditional
    l->spl(0) = c.spl[0]; l->spl(1) = c.spl[1]; l->epl(0) = c.epl[0]; l->epl(1) = c.epl[1];
    l->sdisp = c.sdisp; l->edisp = c.edisp;
  }
}

void StereoFrameHandler::getPoseInfoOnLine(const Matrix4d DT_inv, const Vector2d J_loss, const LineFeature* line,
                                           const double cutRatio[2], gfpl::Mat<6, 6>& info_pose) {
  // one evaluation == a cut problem with a single line, no points and the search range pinned to `cutRatio`:
  // pass 1 of the kernel evaluates cut {0,0} on endpoints pre-cut on the host exactly as :1347-1352 do.
  gfpl_ln_obs o{};
  gfpl_line_cov c{};
  const double c0 = cutRatio[0], c1 = cutRatio[1];
  for (int i = 0; i < 3; i++) {
    o.sP[i] = (1 - c0) * line->sP(i) + c0 * line->eP(i);
    o.eP[i] = (1 - c1) * line->eP(i) + c1 * line->sP(i);
  }
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      c.covS[i * 3 + j] = (1 - c0) * (1 - c0) * line->covSpt3D(i, j) + c0 * c0 * line->covEpt3D(i, j);
      c.covE[i * 3 + j] = (1 - c1) * (1 - c1) * line->covEpt3D(i, j) + c1 * c1 * line->covSpt3D(i, j);
    }
  o.le_obs[0] = J_loss(0); o.le_obs[1] = J_loss(1); o.le_obs[2] = 0.0; o.sigma2 = 1.0; o.prev_idx = 0;
  gfpl_pt_obs nopt{};
  uint8_t pin = 0, lin = 1;
  int32_t n_pt = 0, n_ln = 1;
  gfpl_track_obs obs{1, 1, 1, &n_pt, &n_ln, &nopt, &pin, &o, &lin};
  gfpl_config saved = gfpl_->config();
  gfpl_->config().cut_rng[0] = 0.0; gfpl_->config().cut_rng[1] = -1.0;   // empty range: no neighbour is admissible
  gfpl_->apply();
  double DTr[16];
  DT_inv.to_row_major(DTr);
  gfpl_line_cut cut{};
  int rc = gfpl_linecut_submodular(gfpl_->get(), DTr, &obs, &c, 1, &cut);
  gfpl_->config() = saved;
  gfpl_->apply();
  gfpl_->check(rc);
  info_pose.from_row_major(cut.invCovPose);
}

void StereoFrameHandler::updateEndPointByRatio(LineFeature* line) {  // :1451-1470 (scalar host bookkeeping)
  if (std::fabs(line->cutRatio[0]) < 0.0001 && std::fabs(line->cutRatio[1]) < 0.0001) return;
  if (std::fabs(line->cutRatio[0]) > 0.0001) {
    Vector3d sP_cut;
    for (int i = 0; i < 3; i++) sP_cut(i) = (1 - line->cutRatio[0]) * line->sP(i) + line->cutRatio[0] * line->eP(i);
    line->sP = sP_cut;
    line->spl = cam->projection(sP_cut);
    line->sdisp = cam->getDisparity(sP_cut(2));
  }
  if (std::fabs(line->cutRatio[1]) > 0.0001) {
    Vector3d eP_cut;
    for (int i = 0; i < 3; i++) eP_cut(i) = (1 - line->cutRatio[1]) * line->eP(i) + line->cutRatio[1] * line->sP(i);
    line->eP = eP_cut;
    line->epl = cam->projection(eP_cut);
    line->edisp = cam->getDisparity(eP_cut(2));
  }
}

}  // namespace StVO
