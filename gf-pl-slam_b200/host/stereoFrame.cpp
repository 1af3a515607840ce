// stereoFrame.cpp — StVO::StereoFrame facade over the C ABI (see stereoFrame.hpp).
#include "stereoFrame.hpp"

#include <cmath>

namespace StVO {

GfplContext::GfplContext(const gfpl_camera& cam, int device, const gfpl_config* cfg) {
  if (cfg) cfg_ = *cfg;
  else gfpl_config_default(&cfg_);
  int rc = gfpl_ctx_create(device, &cam, &cfg_, &ctx_);
  if (rc != GFPL_OK) throw GfplError(rc, gfpl_last_error(nullptr));
}
GfplContext::~GfplContext() { gfpl_ctx_destroy(ctx_); }
void GfplContext::apply() { check(gfpl_ctx_set_config(ctx_, &cfg_)); }
void GfplContext::check(int rc) const {
  if (rc != GFPL_OK) throw GfplError(rc, gfpl_last_error(ctx_));
}

StereoFrame::StereoFrame(int idx_, PinholeStereoCamera* cam_, GfplContext* gfpl, double time_stamp_)
    : frame_idx(idx_), time_stamp(time_stamp_), cam(cam_), gfpl_(gfpl) {
  Tfw = Matrix4d::Identity();
  DT = Matrix4d::Identity();
}

// the reference never frees its features (`~StereoFrame(){}`, stereoFrame.cpp:146); the facade owns them
StereoFrame::~StereoFrame() {
  for (auto* p : stereo_pt) delete p;
  for (auto* l : stereo_ls) delete l;
}

void StereoFrame::setFeatures(const vector<KeyPoint>& pl, const DescMat& dl, const vector<KeyPoint>& pr,
                              const DescMat& dr, const vector<KeyLine>& ll, const DescMat& ldl,
                              const vector<KeyLine>& lr, const DescMat& ldr) {
  points_l = pl; points_r = pr; lines_l = ll; lines_r = lr;
  pdesc_l = dl; pdesc_r = dr; ldesc_l = ldl; ldesc_r = ldr;
}

static void to_abi(const vector<KeyPoint>& in, vector<gfpl_keypoint>& out, size_t cap) {
  out.assign(cap, gfpl_keypoint{});
  for (size_t i = 0; i < in.size(); i++) out[i] = gfpl_keypoint{in[i].pt.x, in[i].pt.y, in[i].octave, in[i].class_id};
}
static void to_abi(const vector<KeyLine>& in, vector<gfpl_keyline>& out, size_t cap) {
  out.assign(cap, gfpl_keyline{});
  for (size_t i = 0; i < in.size(); i++)
    out[i] = gfpl_keyline{in[i].startPointX, in[i].startPointY, in[i].endPointX, in[i].endPointY,
                          in[i].angle, in[i].octave, in[i].lineLength, in[i].response};
}
static void pad_desc(const DescMat& in, vector<uint8_t>& out, size_t cap) {
  out.assign(cap * 32, 0);
  if (in.rows) std::memcpy(out.data(), in.ptr(), (size_t)in.rows * 32);
}

void StereoFrame::stereoMatch(int variant) {
  const int cap_pt = (int)std::max<size_t>(std::max(points_l.size(), points_r.size()), 1);
  const int cap_ln = (int)std::max<size_t>(std::max(lines_l.size(), lines_r.size()), 1);
  vector<gfpl_keypoint> kl, kr;
  vector<gfpl_keyline> ll, lr;
  vector<uint8_t> dl, dr, ldl, ldr;
  to_abi(points_l, kl, cap_pt); to_abi(points_r, kr, cap_pt);
  to_abi(lines_l, ll, cap_ln); to_abi(lines_r, lr, cap_ln);
  pad_desc(pdesc_l, dl, cap_pt); pad_desc(pdesc_r, dr, cap_pt);
  pad_desc(ldesc_l, ldl, cap_ln); pad_desc(ldesc_r, ldr, cap_ln);
  int32_t npl = (int32_t)points_l.size(), npr = (int32_t)points_r.size();
  int32_t nll = (int32_t)lines_l.size(), nlr = (int32_t)lines_r.size();
  gfpl_raw_frames raw{1, cap_pt, cap_ln, &npl, &npr, &nll, &nlr, kl.data(), kr.data(), dl.data(), dr.data(),
                      ll.data(), lr.data(), ldl.data(), ldr.data()};
  vector<gfpl_point_feature> pt(cap_pt);
  vector<gfpl_line_feature> ln(cap_ln);
  vector<uint8_t> pd((size_t)cap_pt * 32), ld((size_t)cap_ln * 32);
  int32_t n_pt = 0, n_ln = 0;
  gfpl_stereo_frames out{1, cap_pt, cap_ln, &n_pt, &n_ln, pt.data(), pd.data(), ln.data(), ld.data()};
  if (variant == GFPL_STEREO_ORBSLAM)
    gfpl_->check(gfpl_stereo_match_orbslam(gfpl_->get(), &raw, &right_pyramid, &out));
  else
    gfpl_->check(gfpl_stereo_match(gfpl_->get(), &raw, variant, &out));

  for (auto* p : stereo_pt) delete p;
  for (auto* l : stereo_ls) delete l;
  stereo_pt.clear(); stereo_ls.clear();
  for (int k = 0; k < n_pt; k++) {  // stereoFrame.cpp:862-872
    auto* f = new PointFeature();
    f->pl(0) = pt[k].pl[0]; f->pl(1) = pt[k].pl[1];
    f->disp = pt[k].disp;
    for (int i = 0; i < 3; i++) f->P(i) = pt[k].P[i];
    f->idx = pt[k].idx; f->level = pt[k].level; f->sigma2 = pt[k].sigma2;
    stereo_pt.push_back(f);
    if (variant != GFPL_STEREO_ORBSLAM) {  // commented out in _ORBSLAM (:616-620)
      points_l[pt[k].src_l].class_id = k;
      points_r[pt[k].src_r].class_id = k;
    }
  }
  pdesc_l.resize(n_pt);  // pdesc_l_.copyTo(pdesc_l) (:878): row k <-> stereo_pt[k]
  if (n_pt) std::memcpy(pdesc_l.ptr(), pd.data(), (size_t)n_pt * 32);
  for (int k = 0; k < n_ln; k++) {  // stereoFrame.cpp:1002-1007
    auto* f = new LineFeature();
    for (int i = 0; i < 2; i++) { f->spl(i) = ln[k].spl[i]; f->epl(i) = ln[k].epl[i]; }
    for (int i = 0; i < 3; i++) { f->sP(i) = ln[k].sP[i]; f->eP(i) = ln[k].eP[i]; f->le(i) = ln[k].le[i]; }
    f->sdisp = ln[k].sdisp; f->edisp = ln[k].edisp; f->angle = ln[k].angle;
    f->idx = ln[k].idx; f->level = ln[k].level; f->sigma2 = ln[k].sigma2;
    stereo_ls.push_back(f);
  }
  ldesc_l.resize(n_ln);  // ldesc_l_.copyTo(ldesc_l) (:1013)
  if (n_ln) std::memcpy(ldesc_l.ptr(), ld.data(), (size_t)n_ln * 32);
}

void StereoFrame::extractInitialStereoFeatures(int) { stereoMatch(GFPL_STEREO_INITIAL); }
void StereoFrame::extractStereoFeatures_Brute(int) { stereoMatch(GFPL_STEREO_BRUTE); }
void StereoFrame::extractStereoFeatures_ORBSLAM(int) { stereoMatch(GFPL_STEREO_ORBSLAM); }

void StereoFrame::knn(int norm, const DescMat& a, const DescMat& b, vector<vector<DMatch>>& out) {
  out.clear();
  if (a.rows == 0) return;
  vector<int32_t> idx((size_t)a.rows * 2);
  vector<float> dist((size_t)a.rows * 2);
  gfpl_->check(gfpl_knn2_hamming256(gfpl_->get(), a.ptr(), a.rows, b.ptr(), b.rows, norm, idx.data(), dist.data()));
  out.resize(a.rows);
  for (int i = 0; i < a.rows; i++)
    for (int k = 0; k < 2; k++) {
      DMatch m;
      m.queryIdx = i; m.trainIdx = idx[2 * i + k]; m.imgIdx = 0; m.distance = dist[2 * i + k];
      out[i].push_back(m);
    }
}

void StereoFrame::radius(int norm, float r, const DescMat& a, const DescMat& b, vector<vector<DMatch>>& out) {
  out.assign(a.rows, vector<DMatch>());
  if (a.rows == 0) return;
  vector<int32_t> off((size_t)a.rows + 1), idx;
  vector<float> dist;
  int cap = a.rows * 4;
  for (;;) {
    idx.resize(cap > 0 ? cap : 1);
    dist.resize(cap > 0 ? cap : 1);
    int rc = gfpl_radius_hamming256(gfpl_->get(), a.ptr(), a.rows, b.ptr(), b.rows, norm, r, off.data(), idx.data(),
                                    dist.data(), cap);
    if (rc == GFPL_E_CAPACITY) { cap = off[a.rows]; continue; }  // offsets are filled: re-size and retry
    gfpl_->check(rc);
    break;
  }
  for (int i = 0; i < a.rows; i++)
    for (int e = off[i]; e < off[i + 1]; e++) {
      DMatch m;
      m.queryIdx = i; m.trainIdx = idx[e]; m.imgIdx = 0; m.distance = dist[e];
      out[i].push_back(m);
    }
}

void StereoFrame::matchPointFeatures(BFMatcher* bfm, const DescMat& d1, const DescMat& d2, vector<vector<DMatch>>& m) {
  knn(bfm->normType, d1, d2, m);
}
void StereoFrame::matchLineFeatures(BFMatcher* bfm, const DescMat& d1, const DescMat& d2, vector<vector<DMatch>>& m) {
  knn(bfm->normType, d1, d2, m);
}
void StereoFrame::matchPointFeatures_radius(BFMatcher* bfm, const DescMat& d1, const DescMat& d2,
                                            vector<vector<DMatch>>& m) {
  radius(bfm->normType, (float)gfpl_->config().point_match_radius, d1, d2, m);
}
void StereoFrame::matchLineFeatures_radius(BFMatcher* bfm, const DescMat& d1, const DescMat& d2,
                                           vector<vector<DMatch>>& m) {
  radius(bfm->normType, (float)gfpl_->config().line_match_radius, d1, d2, m);
}

void StereoFrame::exportStereo(vector<gfpl_point_feature>& pt, vector<gfpl_line_feature>& ln) const {
  pt.assign(std::max<size_t>(stereo_pt.size(), 1), gfpl_point_feature{});
  ln.assign(std::max<size_t>(stereo_ls.size(), 1), gfpl_line_feature{});
  for (size_t k = 0; k < stereo_pt.size(); k++) {
    const PointFeature& f = *stereo_pt[k];
    gfpl_point_feature& o = pt[k];
    for (int i = 0; i < 3; i++) o.P[i] = f.P(i);
    o.pl[0] = f.pl(0); o.pl[1] = f.pl(1);
    o.disp = f.disp; o.sigma2 = f.sigma2; o.level = f.level; o.idx = f.idx; o.src_l = o.src_r = -1;
  }
  for (size_t k = 0; k < stereo_ls.size(); k++) {
    const LineFeature& f = *stereo_ls[k];
    gfpl_line_feature& o = ln[k];
    for (int i = 0; i < 3; i++) { o.sP[i] = f.sP(i); o.eP[i] = f.eP(i); o.le[i] = f.le(i); }
    for (int i = 0; i < 2; i++) { o.spl[i] = f.spl(i); o.epl[i] = f.epl(i); }
    o.sdisp = f.sdisp; o.edisp = f.edisp; o.angle = f.angle; o.sigma2 = f.sigma2; o.level = f.level; o.idx = f.idx;
    o.src_l = o.src_r = -1;
  }
}

void StereoFrame::estimateStereoUncertainty() {
  if (stereo_ls.empty()) return;  // :1450
  vector<gfpl_point_feature> pt;
  vector<gfpl_line_feature> ln;
  exportStereo(pt, ln);
  int32_t n_pt = (int32_t)stereo_pt.size(), n_ln = (int32_t)stereo_ls.size();
  vector<uint8_t> pd(pt.size() * 32), ld(ln.size() * 32);
  gfpl_stereo_frames fr{1, (int32_t)pt.size(), (int32_t)ln.size(), &n_pt, &n_ln, pt.data(), pd.data(), ln.data(), ld.data()};
  vector<gfpl_line_cov> cov(ln.size());
  gfpl_->check(gfpl_stereo_uncertainty(gfpl_->get(), &fr, cov.data()));
  for (int k = 0; k < n_ln; k++)
    for (int i = 0; i < 3; i++)
      for (int j = 0; j < 3; j++) {
        stereo_ls[k]->covSpt3D(i, j) = cov[k].covS[i * 3 + j];
        stereo_ls[k]->covEpt3D(i, j) = cov[k].covE[i * 3 + j];
      }
}

}  // namespace StVO
