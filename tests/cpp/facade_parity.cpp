// facade_parity.cpp — drives the C++ facades (StVO::StereoFrame / StereoFrameHandler, PLSLAM::MapHandler) the way
// the reference's main loop drives its classes (app/plslam_mod.cpp:377-408) and checks the results against the
// CPU oracle on the same synthetic detector output.  TEST code: it links the oracle, the product never does.
//
//   facade_parity --no-gpu : only checks that the facade fails loudly without a CUDA device
//   facade_parity          : full run on cuda:0
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../gf-pl-slam_b200/host/mapHandler.hpp"
#include "../../gf-pl-slam_b200/host/stereoFrameHandler.hpp"
#include "../../oracle/gfpl_oracle.h"

using namespace StVO;

static uint64_t rng_state = 0x9E3779B97F4A7C15ull;
static uint32_t rnd() { rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng_state >> 33); }
static double uni(double a, double b) { return a + (b - a) * (rnd() / 2147483648.0); }
static double gauss() { double s = 0; for (int i = 0; i < 12; i++) s += uni(0, 1); return s - 6.0; }

struct Raw {
  std::vector<KeyPoint> pl, pr;
  std::vector<KeyLine> ll, lr;
  DescMat dl, dr, ldl, ldr;
};

static int fails = 0;
#define EXPECT(c, ...) do { if (!(c)) { fails++; printf("FAIL %s:%d: ", __FILE__, __LINE__); printf(__VA_ARGS__); printf("\n"); } } while (0)

static void flip(uint8_t* d) { for (int i = 0; i < 32; i++) d[i] ^= (uint8_t)(rnd() & rnd() & rnd() & rnd() & 0xFF); }

// scene in the first camera frame; frame k observes it after the rigid motion T (R about y + translation)
static void observe(const PinholeStereoCamera& cam, const std::vector<gfpl::Vector3d>& P, const std::vector<std::vector<uint8_t>>& code,
                    const std::vector<gfpl::Vector3d>& Ls, const std::vector<gfpl::Vector3d>& Le,
                    const std::vector<std::vector<uint8_t>>& lcode, double ang, const double t[3], Raw& raw) {
  const double c = std::cos(ang), s = std::sin(ang);
  auto xf = [&](const gfpl::Vector3d& p) { gfpl::Vector3d q; q(0) = c * p(0) + s * p(2) + t[0]; q(1) = p(1) + t[1]; q(2) = -s * p(0) + c * p(2) + t[2]; return q; };
  for (size_t i = 0; i < P.size(); i++) {
    gfpl::Vector3d q = xf(P[i]);
    gfpl::Vector2d uv = cam.projection(q);
    double d = cam.getDisparity(q(2));
    KeyPoint kl, kr;
    kl.pt.x = (float)(uv(0) + 0.3 * gauss()); kl.pt.y = (float)(uv(1) + 0.3 * gauss()); kl.octave = rnd() % 4;
    kr.pt.x = (float)(uv(0) - d + 0.3 * gauss()); kr.pt.y = (float)(uv(1) + 0.3 * gauss()); kr.octave = kl.octave;
    raw.pl.push_back(kl); raw.pr.push_back(kr);
    std::vector<uint8_t> a = code[i], b = code[i];
    flip(a.data()); flip(b.data());
    raw.dl.push_back_row(a.data()); raw.dr.push_back_row(b.data());
  }
  for (size_t i = 0; i < Ls.size(); i++) {
    gfpl::Vector3d a = xf(Ls[i]), b = xf(Le[i]);
    gfpl::Vector2d ua = cam.projection(a), ub = cam.projection(b);
    double da = cam.getDisparity(a(2)), db = cam.getDisparity(b(2));
    KeyLine l, r;
    l.startPointX = (float)(ua(0) + 0.3 * gauss()); l.startPointY = (float)(ua(1) + 0.3 * gauss());
    l.endPointX = (float)(ub(0) + 0.3 * gauss()); l.endPointY = (float)(ub(1) + 0.3 * gauss());
    r.startPointX = (float)(ua(0) - da + 0.3 * gauss()); r.startPointY = (float)(ua(1) + 0.3 * gauss());
    r.endPointX = (float)(ub(0) - db + 0.3 * gauss()); r.endPointY = (float)(ub(1) + 0.3 * gauss());
    l.angle = std::atan2(l.endPointY - l.startPointY, l.endPointX - l.startPointX);
    l.lineLength = r.lineLength = 50; l.response = r.response = 50;
    raw.ll.push_back(l); raw.lr.push_back(r);
    std::vector<uint8_t> x = lcode[i], y = lcode[i];
    flip(x.data()); flip(y.data());
    raw.ldl.push_back_row(x.data()); raw.ldr.push_back_row(y.data());
  }
}

static gfpl_raw_frames raw_view(const Raw& r, std::vector<gfpl_keypoint>& kl, std::vector<gfpl_keypoint>& kr, std::vector<gfpl_keyline>& ll,
                                std::vector<gfpl_keyline>& lr, int32_t* n) {
  for (auto& k : r.pl) kl.push_back(gfpl_keypoint{k.pt.x, k.pt.y, k.octave, -1});
  for (auto& k : r.pr) kr.push_back(gfpl_keypoint{k.pt.x, k.pt.y, k.octave, -1});
  for (auto& k : r.ll) ll.push_back(gfpl_keyline{k.startPointX, k.startPointY, k.endPointX, k.endPointY, k.angle, k.octave, k.lineLength, k.response});
  for (auto& k : r.lr) lr.push_back(gfpl_keyline{k.startPointX, k.startPointY, k.endPointX, k.endPointY, k.angle, k.octave, k.lineLength, k.response});
  n[0] = (int32_t)kl.size(); n[1] = (int32_t)kr.size(); n[2] = (int32_t)ll.size(); n[3] = (int32_t)lr.size();
  return gfpl_raw_frames{1, n[0], n[2], &n[0], &n[1], &n[2], &n[3], kl.data(), kr.data(), r.dl.ptr(), r.dr.ptr(), ll.data(), lr.data(), r.ldl.ptr(), r.ldr.ptr()};
}

static int test_orbslam_stereo(const Raw& r1, PinholeStereoCamera& cam, GfplContext& gf);

static int test_tracking() {
  PinholeStereoCamera cam(752, 480, 458.654, 458.654, 367.215, 248.375, 0.110078);
  GfplContext gf(cam.c_abi());
  const int NP = 300, NL = 100;
  std::vector<gfpl::Vector3d> P(NP), Ls(NL), Le(NL);
  std::vector<std::vector<uint8_t>> code(NP, std::vector<uint8_t>(32)), lcode(NL, std::vector<uint8_t>(32));
  for (int i = 0; i < NP; i++) {
    double z = uni(2, 12), u = uni(60, 690), v = uni(60, 420);
    P[i](0) = (u - 367.215) * z / 458.654; P[i](1) = (v - 248.375) * z / 458.654; P[i](2) = z;
    for (auto& b : code[i]) b = (uint8_t)rnd();
  }
  for (int i = 0; i < NL; i++) {
    double z = uni(2, 7), u = uni(100, 650), v = uni(100, 380);
    Ls[i](0) = (u - 367.215) * z / 458.654; Ls[i](1) = (v - 248.375) * z / 458.654; Ls[i](2) = z;
    Le[i](0) = Ls[i](0) + uni(-0.4, 0.4); Le[i](1) = Ls[i](1) + uni(0.4, 1.0) * (rnd() & 1 ? 1 : -1); Le[i](2) = z + uni(-0.3, 0.3);
    for (auto& b : lcode[i]) b = (uint8_t)rnd();
  }
  Raw r0, r1;
  const double t0[3] = {0, 0, 0}, t1[3] = {0.02, -0.01, -0.04};
  observe(cam, P, code, Ls, Le, lcode, 0.0, t0, r0);
  observe(cam, P, code, Ls, Le, lcode, 0.004, t1, r1);

  // ---- the facade, driven like app/plslam_mod.cpp:377-408 ------------------------------------------------
  StereoFrameHandler StVO(&cam, &gf);
  StereoFrame* f0 = new StereoFrame(0, &cam, &gf, 0.0);
  f0->setFeatures(r0.pl, r0.dl, r0.pr, r0.dr, r0.ll, r0.ldl, r0.lr, r0.ldr);
  StVO.initialize(f0);
  StereoFrame* f1 = new StereoFrame(1, &cam, &gf, 0.05);
  f1->setFeatures(r1.pl, r1.dl, r1.pr, r1.dr, r1.ll, r1.ldl, r1.lr, r1.ldr);
  StVO.insertStereoPair(f1);
  // BFMatcher wrapper (stereoFrame.cpp:1229-1234): 2 matches per query, ascending distance
  std::vector<std::vector<DMatch>> mm;
  BFMatcher bfm(gfpl::NORM_HAMMING, false);
  f1->matchPointFeatures(&bfm, StVO.prev_frame->pdesc_l, f1->pdesc_l, mm);
  EXPECT(mm.size() == StVO.prev_frame->stereo_pt.size() && mm[0].size() == 2 && mm[0][0].distance <= mm[0][1].distance, "knn wrapper");
  std::vector<std::vector<DMatch>> mr;
  f1->matchPointFeatures_radius(&bfm, StVO.prev_frame->pdesc_l, f1->pdesc_l, mr);
  EXPECT(mr.size() == mm.size(), "radius wrapper size");
  for (size_t i = 0; i < mr.size(); i++)
    if (mm[i][0].distance <= 50) EXPECT(!mr[i].empty() && mr[i][0].trainIdx == mm[i][0].trainIdx, "radius list head == nearest neighbour");
  Matrix6d H; Vector6d g; double e;
  StVO.optimizeFunctions(StVO.prev_frame->DT, H, g, e);
  StVO.optimizePose(StVO.prev_frame->DT);
  printf("facade: stereo pts %zu/%zu lines %zu/%zu  matched %zu pts %zu lines  inliers %d  err %.4f  status %d\n",
         f0->stereo_pt.size(), f1->stereo_pt.size(), f0->stereo_ls.size(), f1->stereo_ls.size(), StVO.matched_pt.size(),
         StVO.matched_ls.size(), StVO.n_inliers, f1->err_norm, StVO.last_status);

  // ---- the oracle on the same detector output ------------------------------------------------------------------
  gfpl_camera ccam = cam.c_abi();
  gfpl_config cfg;
  orc_config_default(&cfg);
  std::vector<gfpl_keypoint> kl0, kr0, kl1, kr1;
  std::vector<gfpl_keyline> ll0, lr0, ll1, lr1;
  int32_t n0[4], n1[4];
  gfpl_raw_frames raw0 = raw_view(r0, kl0, kr0, ll0, lr0, n0), raw1 = raw_view(r1, kl1, kr1, ll1, lr1, n1);
  std::vector<gfpl_point_feature> ppt(NP), cpt(NP);
  std::vector<gfpl_line_feature> pln(NL), cln(NL);
  std::vector<uint8_t> ppd(NP * 32), pld(NL * 32), cpd(NP * 32), cld(NL * 32);
  int32_t pn[2], cn[2];
  gfpl_stereo_frames prev{1, NP, NL, &pn[0], &pn[1], ppt.data(), ppd.data(), pln.data(), pld.data()};
  gfpl_stereo_frames curr{1, NP, NL, &cn[0], &cn[1], cpt.data(), cpd.data(), cln.data(), cld.data()};
  orc_stereo_match(&ccam, &cfg, &raw0, GFPL_STEREO_INITIAL, &prev);
  gfpl_frame_state st{};
  for (int i = 0; i < 16; i++) st.Tfw[i] = st.DT[i] = (i % 5 == 0) ? 1.0 : 0.0;
  for (int i = 0; i < 36; i++) st.Tfw_cov[i] = (i % 7 == 0) ? 1.0 : 0.0;
  st.time_stamp = 0.0; st.time_stamp_curr = 0.05;
  std::vector<gfpl_pt_obs> opt(500); std::vector<gfpl_ln_obs> oln(300); std::vector<uint8_t> pin(500), lin(300);
  int32_t on[2];
  gfpl_track_obs obs{1, 500, 300, &on[0], &on[1], opt.data(), pin.data(), oln.data(), lin.data()};
  std::vector<gfpl_line_cut> cut(300);
  gfpl_pose_result pose{};
  orc_track(&ccam, &cfg, &prev, &st, &raw1, GFPL_TRACK_LINECUT | GFPL_TRACK_SEED_GIVEN, &curr, &obs, cut.data(), &pose, 1);

  EXPECT((int)f0->stereo_pt.size() == pn[0] && (int)f0->stereo_ls.size() == pn[1], "prev stereo counts %zu/%d %zu/%d", f0->stereo_pt.size(), pn[0], f0->stereo_ls.size(), pn[1]);
  EXPECT((int)f1->stereo_pt.size() == cn[0] && (int)f1->stereo_ls.size() == cn[1], "curr stereo counts");
  EXPECT((int)StVO.matched_pt.size() == on[0] && (int)StVO.matched_ls.size() == on[1], "matched counts %zu/%d %zu/%d", StVO.matched_pt.size(), on[0], StVO.matched_ls.size(), on[1]);
  EXPECT(on[0] > 150 && on[1] > 5, "enough matches (%d, %d)", on[0], on[1]);
  EXPECT(StVO.n_inliers == pose.n_inliers && StVO.last_status == pose.status, "inliers %d vs %d", StVO.n_inliers, pose.n_inliers);
  size_t k = 0;
  for (auto it = StVO.matched_ls.begin(); it != StVO.matched_ls.end(); ++it, ++k) {
    EXPECT((*it)->cutRatio[0] == cut[k].cutRatio[0] && (*it)->cutRatio[1] == cut[k].cutRatio[1], "cut ratio of line %zu", k);
    EXPECT((*it)->inlier == (lin[k] != 0), "line inlier %zu", k);
  }
  k = 0;
  for (auto it = StVO.matched_pt.begin(); it != StVO.matched_pt.end(); ++it, ++k) EXPECT((*it)->inlier == (pin[k] != 0), "point inlier %zu", k);
  double dmax = 0;
  for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++) dmax = std::fmax(dmax, std::fabs(f1->DT(i, j) - pose.DT[i * 4 + j]));
  EXPECT(dmax < 1e-9, "DT differs by %g", dmax);
  EXPECT(std::fabs(f1->err_norm - pose.err_norm) < 1e-9 * std::fabs(pose.err_norm), "err_norm");
  // the recovered motion is the simulated one (DT maps curr -> prev: inverse of the scene motion)
  EXPECT(std::fabs(f1->DT(0, 3) + 0.02) < 0.02 && std::fabs(f1->DT(2, 3) - 0.04) < 0.03, "recovered translation (%.3f %.3f %.3f)", f1->DT(0, 3), f1->DT(1, 3), f1->DT(2, 3));
  // ---- C6: estimateProjUncertainty_descent through the facade vs the oracle, on the lines as the submodular cut left them ----
  {
    const size_t nl = StVO.matched_ls.size();
    std::vector<gfpl_ln_obs> dl(std::max<size_t>(nl, 1));
    std::vector<gfpl_line_cov> dc(dl.size());
    std::vector<gfpl_pt_obs> dp(1); std::vector<uint8_t> dpin(1), dlin(dl.size(), 1);
    size_t q = 0;
    for (auto it = StVO.matched_ls.begin(); it != StVO.matched_ls.end(); ++it, ++q) {
      for (int i = 0; i < 3; i++) { dl[q].sP[i] = (*it)->sP(i); dl[q].eP[i] = (*it)->eP(i); dl[q].le_obs[i] = (*it)->le_obs(i); }
      dl[q].sigma2 = (*it)->sigma2; dl[q].prev_idx = (int32_t)q;
      for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++) { dc[q].covS[i * 3 + j] = (*it)->covSpt3D(i, j); dc[q].covE[i * 3 + j] = (*it)->covEpt3D(i, j); }
    }
    double Tc[16], Tp[16], Ti[16], DTi[16];   // curr->Tfw^-1 * prev->Tfw with the facade's own operations (:1635)
    f1->Tfw.to_row_major(Tc); f0->Tfw.to_row_major(Tp);
    for (int i = 0; i < 16; i++) Ti[i] = (i % 5 == 0) ? 1.0 : 0.0;
    for (int i = 0; i < 3; i++) {
      for (int j = 0; j < 3; j++) Ti[i * 4 + j] = Tc[j * 4 + i];
      Ti[i * 4 + 3] = -(Tc[0 * 4 + i] * Tc[3] + Tc[1 * 4 + i] * Tc[7] + Tc[2 * 4 + i] * Tc[11]);
    }
    for (int i = 0; i < 4; i++)
      for (int j = 0; j < 4; j++) { double a = 0; for (int kk = 0; kk < 4; kk++) a += Ti[i * 4 + kk] * Tp[kk * 4 + j]; DTi[i * 4 + j] = a; }
    int32_t dn[2] = {0, (int32_t)nl};
    gfpl_track_obs dobs{1, 1, (int32_t)dl.size(), &dn[0], &dn[1], dp.data(), dpin.data(), dl.data(), dlin.data()};
    std::vector<gfpl_line_cut_descent> dout(dl.size());
    orc_linecut_descent(&ccam, &cfg, DTi, &dobs, dc.data(), (int32_t)dc.size(), dout.data());
    double rng[2] = {cfg.cut_rng[0], cfg.cut_rng[1]};
    StVO.estimateProjUncertainty_descent(cfg.cut_step, rng);
    int moved = 0;
    q = 0;
    for (auto it = StVO.matched_ls.begin(); it != StVO.matched_ls.end(); ++it, ++q) {
      const LineFeature* l = *it;
      EXPECT(l->cutRatio[0] == dout[q].cutRatio[0] && l->cutRatio[1] == dout[q].cutRatio[1], "descent cut ratio of line %zu", q);
      EXPECT(l->volumeArr.size() == 1 && l->volumeArr.back() == dout[q].volume, "descent volume of line %zu", q);
      EXPECT(l->sP(0) == dl[q].sP[0] && l->eP(2) == dl[q].eP[2] && l->spl(1) == dout[q].spl[1] && l->edisp == dout[q].edisp, "descent end points of line %zu", q);
      EXPECT(l->covlPt2D.size() == 4 && l->covlPt2D[1](2, 2) == dout[q].covlPt2D[0][8] && l->covlPt2D[2](0, 1) == dout[q].covlPt2D[1][1], "descent covlPt2D of line %zu", q);
      EXPECT(l->lPt2D.size() == 4 && l->lPt2D[3](0) == dout[q].lPt2D[3][0] && l->fst_idx == 0 && l->lst_idx == 3, "descent lPt2D of line %zu", q);
      moved += (dout[q].cutRatio[0] > 0 || dout[q].cutRatio[1] > 0) ? 1 : 0;
    }
    printf("facade estimateProjUncertainty_descent: %zu lines, %d moved\n", nl, moved);
  }
  // ---- MapHandler::lookForCommonMatches (KF <-> KF half) on the same two frames as key frames 0 and 1 ------------
  {
    for (auto* pf : f0->stereo_pt) pf->idx = -1;    // no landmark yet (the handler reused idx for f2f bookkeeping)
    for (auto* pf : f1->stereo_pt) pf->idx = -1;
    for (auto* lf : f0->stereo_ls) lf->idx = -1;
    PLSLAM::MapHandler map(&cam, &gf);
    PLSLAM::KeyFrame k0, k1;
    k0.kf_idx = 0; k0.stereo_frame = f0; k0.T_kf_w = f0->Tfw;
    k1.kf_idx = 1; k1.stereo_frame = f1; k1.T_kf_w = f1->Tfw;
    PLSLAM::KeyFrame* k1p = &k1;
    map.lookForCommonMatches(&k0, k1p);
    // oracle: same matching + gates
    gfpl::Matrix4d DTk = PLSLAM::inverse_se3(k1.T_kf_w) * k0.T_kf_w;
    double DTr[16];
    DTk.to_row_major(DTr);
    std::vector<gfpl_pt_obs> kpt(NP); std::vector<gfpl_ln_obs> kln(NL); std::vector<uint8_t> kpin(NP), klin(NL);
    int32_t kn[2];
    gfpl_track_obs kobs{1, NP, NL, &kn[0], &kn[1], kpt.data(), kpin.data(), kln.data(), klin.data()};
    // stereo frames as the facade sees them after tracking (line endpoints of f0 were cut in place)
    std::vector<gfpl_point_feature> a0, a1; std::vector<gfpl_line_feature> b0, b1;
    f0->exportStereo(a0, b0); f1->exportStereo(a1, b1);
    int32_t m0[2] = {(int32_t)f0->stereo_pt.size(), (int32_t)f0->stereo_ls.size()}, m1[2] = {(int32_t)f1->stereo_pt.size(), (int32_t)f1->stereo_ls.size()};
    gfpl_stereo_frames w0{1, (int32_t)a0.size(), (int32_t)b0.size(), &m0[0], &m0[1], a0.data(), const_cast<uint8_t*>(f0->pdesc_l.ptr()), b0.data(), const_cast<uint8_t*>(f0->ldesc_l.ptr())};
    gfpl_stereo_frames w1{1, (int32_t)a1.size(), (int32_t)b1.size(), &m1[0], &m1[1], a1.data(), const_cast<uint8_t*>(f1->pdesc_l.ptr()), b1.data(), const_cast<uint8_t*>(f1->ldesc_l.ptr())};
    orc_kf_match(&ccam, &cfg, &w0, &w1, DTr, &kobs);
    printf("facade lookForCommonMatches: %d common points, %d common lines (oracle %d, %d), %d map points, %d map lines\n",
           map.last_common_pt, map.last_common_ls, kn[0], kn[1], map.max_pt_idx, map.max_ls_idx);
    EXPECT(map.last_common_pt == kn[0] && map.last_common_ls == kn[1], "common matches vs oracle");
    EXPECT(kn[0] > 100, "enough common points (%d)", kn[0]);
    EXPECT(map.max_pt_idx == kn[0] && (int)map.map_points.size() == kn[0], "one new MapPoint per accepted pair");
    EXPECT(map.full_graph[0][1] == kn[0] + kn[1] && map.full_graph[1][0] == kn[0] + kn[1], "covisibility graph");
    for (int k = 0; k < kn[0]; k++) {
      EXPECT(f0->stereo_pt[kpt[k].prev_idx]->idx == k && f1->stereo_pt[kpt[k].curr_idx]->idx == k, "landmark index of pair %d", k);
      EXPECT(map.map_points[k]->kf_obs_list.size() == 2 && map.map_points[k]->kf_obs_list[1] == 1, "observations of landmark %d", k);
    }
    // ---- local-map <-> KF half (:509-770): kf1 forgets every second landmark, then finds it again through the map ----
    {
      std::vector<int> all_p, all_l;
      for (auto* mp : map.map_points) {
        mp->local = true;
        all_p.push_back(mp->idx);
        if (mp->idx % 2 == 0) { mp->desc_list.pop_back(); mp->obs_list.pop_back(); mp->kf_obs_list.pop_back(); mp->dir_list.pop_back(); mp->sigma_list.pop_back(); }
      }
      for (auto* ml : map.map_lines) {
        ml->local = true;
        all_l.push_back(ml->idx);
        if (ml->idx % 2 == 0) { ml->desc_list.pop_back(); ml->obs_list.pop_back(); ml->kf_obs_list.pop_back(); ml->dir_list.pop_back(); ml->pts_list.pop_back(); ml->sigma_list.pop_back(); }
      }
      for (auto* pf : f1->stereo_pt) if (pf->idx >= 0 && pf->idx % 2 == 0) pf->idx = -1;
      for (auto* lf : f1->stereo_ls) if (lf->idx >= 0 && lf->idx % 2 == 0) lf->idx = -1;
      map.refreshMedoids(all_p, all_l);
      std::vector<PLSLAM::MapPoint*> cp_; std::vector<PLSLAM::MapLine*> cl_;
      for (auto* mp : map.map_points) if (mp->kf_obs_list.back() != 1) cp_.push_back(mp);
      for (auto* ml : map.map_lines) if (ml->kf_obs_list.back() != 1) cl_.push_back(ml);
      int32_t np_ = (int32_t)cp_.size(), nl_ = (int32_t)cl_.size();
      std::vector<double> X(3 * (size_t)std::max(np_, 1)), L(6 * (size_t)std::max(nl_, 1));
      std::vector<uint8_t> dX(32 * (size_t)std::max(np_, 1)), dL(32 * (size_t)std::max(nl_, 1));
      for (int i = 0; i < np_; i++) { for (int k = 0; k < 3; k++) X[3 * i + k] = cp_[i]->point3D(k); std::memcpy(&dX[32 * i], cp_[i]->med_desc.data(), 32); }
      for (int i = 0; i < nl_; i++) { for (int k = 0; k < 6; k++) L[6 * i + k] = cl_[i]->line3D(k); std::memcpy(&dL[32 * i], cl_[i]->med_desc.data(), 32); }
      gfpl_local_map lm{1, std::max(np_, 1), std::max(nl_, 1), &np_, &nl_, X.data(), dX.data(), L.data(), dL.data()};
      f1->exportStereo(a1, b1);
      gfpl_stereo_frames w1b{1, (int32_t)a1.size(), (int32_t)b1.size(), &m1[0], &m1[1], a1.data(), const_cast<uint8_t*>(f1->pdesc_l.ptr()), b1.data(), const_cast<uint8_t*>(f1->ldesc_l.ptr())};
      std::vector<gfpl_map_pair> mp_(lm.cap_pt), ml_(lm.cap_ln);
      int32_t mn[2];
      gfpl_map_matches mm{1, lm.cap_pt, lm.cap_ln, &mn[0], &mn[1], mp_.data(), ml_.data()};
      double Tr[16];
      PLSLAM::inverse_se3(k1.T_kf_w).to_row_major(Tr);
      orc_map_match(&ccam, &cfg, &lm, &w1b, Tr, &mm);
      int cpn = 0, cln = 0;
      map.matchLocalMap(k1p, cpn, cln);
      printf("facade matchLocalMap: %d of %d forgotten points and %d of %d lines found again (oracle %d, %d)\n", cpn, np_, cln, nl_, mn[0], mn[1]);
      EXPECT(cpn == mn[0] && cln == mn[1], "local-map matches vs oracle");
      EXPECT(mn[0] > 10, "enough local-map point matches (%d)", mn[0]);
      for (int k = 0; k < mn[0]; k++) {
        EXPECT(f1->stereo_pt[mp_[k].kf_idx]->idx == cp_[mp_[k].map_idx]->idx, "landmark id of local-map pair %d", k);
        EXPECT(cp_[mp_[k].map_idx]->kf_obs_list.back() == 1 && cp_[mp_[k].map_idx]->desc_list.size() == 2, "observation added for pair %d", k);
      }
      for (int k = 0; k < mn[1]; k++) EXPECT(f1->stereo_ls[ml_[k].kf_idx]->idx == cl_[ml_[k].map_idx]->idx, "landmark id of local-map line pair %d", k);
    }
    for (auto* mp : map.map_points) delete mp;
    for (auto* ml : map.map_lines) delete ml;
  }
  if (test_orbslam_stereo(r1, cam, gf)) return 1;
  StVO.updateFrame();
  EXPECT(StVO.prev_frame == f1 && StVO.curr_frame == nullptr, "updateFrame");
  delete f1;
  return 0;
}

// StereoFrame::extractStereoFeatures_ORBSLAM through the facade vs the oracle (same detector output + right pyramid)
static int test_orbslam_stereo(const Raw& r1, PinholeStereoCamera& cam, GfplContext& gf) {
  gfpl_camera ccam = cam.c_abi();
  gfpl_config cfg;
  orc_config_default(&cfg);
  gfpl_pyramid pyr{};
  pyr.n_levels = cfg.orb_nlevels;
  float sf = 1.0f;
  int64_t off = 0;
  for (int l = 0; l < pyr.n_levels; l++) {
    const float isf = 1.0f / sf;
    pyr.width[l] = (int)std::lround(752.0f * isf); pyr.height[l] = (int)std::lround(480.0f * isf);
    pyr.row_stride[l] = pyr.width[l] + 38;          // a bordered Mat like ORBextractor::ComputePyramid's (step > cols)
    pyr.offset[l] = off;
    off += (int64_t)pyr.row_stride[l] * pyr.height[l];
    sf = sf * (float)cfg.orb_scale_factor;
  }
  pyr.frame_stride = off;
  std::vector<uint8_t> img((size_t)off);
  uint32_t x = 12345u;
  for (size_t i = 0; i < img.size(); i++) { x = x * 1664525u + 1013904223u; img[i] = (uint8_t)(((i / 3) * 7 + (x >> 28)) & 0xff); }
  pyr.data = img.data();
  StereoFrame f(2, &cam, &gf, 0.1);
  f.setFeatures(r1.pl, r1.dl, r1.pr, r1.dr, r1.ll, r1.ldl, r1.lr, r1.ldr);
  f.setRightPyramid(pyr);
  f.extractStereoFeatures_ORBSLAM();
  std::vector<gfpl_keypoint> kl, kr; std::vector<gfpl_keyline> ll, lr;
  int32_t n[4];
  gfpl_raw_frames raw = raw_view(r1, kl, kr, ll, lr, n);
  std::vector<gfpl_point_feature> pt(n[0]); std::vector<gfpl_line_feature> ln(n[2]);
  std::vector<uint8_t> pd((size_t)n[0] * 32), ld((size_t)n[2] * 32);
  int32_t on[2];
  gfpl_stereo_frames out{1, n[0], n[2], &on[0], &on[1], pt.data(), pd.data(), ln.data(), ld.data()};
  EXPECT(orc_stereo_match_orbslam(&ccam, &cfg, &raw, &pyr, &out) == 0, "oracle call");
  printf("facade extractStereoFeatures_ORBSLAM: %zu points, %zu lines (oracle %d, %d)\n", f.stereo_pt.size(), f.stereo_ls.size(), on[0], on[1]);
  EXPECT((int)f.stereo_pt.size() == on[0] && (int)f.stereo_ls.size() == on[1] && on[0] > 10, "counts vs oracle");
  for (int k = 0; k < on[0]; k++) {
    EXPECT(f.stereo_pt[k]->disp == pt[k].disp && f.stereo_pt[k]->P(2) == pt[k].P[2] && f.stereo_pt[k]->pl(0) == pt[k].pl[0], "stereo point %d", k);
    EXPECT(std::memcmp(f.pdesc_l.ptr(k), &pd[(size_t)k * 32], 32) == 0, "descriptor row %d", k);
  }
  return 0;
}

static int test_lba() {
  PinholeStereoCamera cam(1241, 376, 718.856, 718.856, 607.1928, 185.2157, 0.537166);
  GfplContext gf(cam.c_abi());
  PLSLAM::MapHandler map(&cam, &gf);
  const int NKF = 6, NPT = 120, NLS = 30;
  for (int k = 0; k < NKF; k++) {
    auto* kf = new PLSLAM::KeyFrame();
    kf->kf_idx = k; kf->local = k >= 1;      // KF 0 is fixed (:1119); it still observes
    PLSLAM::Vector6d x; x(2) = 0.8 * k; x(0) = 0.05 * gauss(); x(4) = 0.01 * gauss();
    kf->T_kf_w = PLSLAM::expmap_se3(x);
    kf->x_kf_w = PLSLAM::logmap_se3(kf->T_kf_w);
    map.map_keyframes.push_back(kf);
  }
  auto project = [&](const PLSLAM::KeyFrame* kf, const gfpl::Vector3d& Xw) {
    gfpl::Matrix4d Tiw = PLSLAM::inverse_se3(kf->T_kf_w);
    gfpl::Vector3d Xc;
    for (int i = 0; i < 3; i++) Xc(i) = Tiw(i, 0) * Xw(0) + Tiw(i, 1) * Xw(1) + Tiw(i, 2) * Xw(2) + Tiw(i, 3);
    return cam.projection(Xc);
  };
  std::vector<gfpl::Vector3d> truth;
  for (int j = 0; j < NPT; j++) {
    auto* mp = new PLSLAM::MapPoint();
    mp->idx = j; mp->local = true;
    gfpl::Vector3d X; X(0) = uni(-6, 6); X(1) = uni(-2, 2); X(2) = uni(12, 40);
    truth.push_back(X);
    for (int k = 0; k < NKF; k++)
      if ((j + k) % 2 == 0 || k == 0) {
        gfpl::Vector2d uv = project(map.map_keyframes[k], X);
        uv(0) += 0.2 * gauss(); uv(1) += 0.2 * gauss();
        mp->obs_list.push_back(uv); mp->kf_obs_list.push_back(k); mp->sigma_list.push_back(1.0);
      }
    for (int i = 0; i < 3; i++) mp->point3D(i) = X(i) + 0.15 * gauss();   // perturbed initial landmark
    map.map_points.push_back(mp);
  }
  for (int j = 0; j < NLS; j++) {
    auto* ml = new PLSLAM::MapLine();
    ml->idx = j; ml->local = true;
    gfpl::Vector3d A, B; A(0) = uni(-5, 5); A(1) = uni(-2, 2); A(2) = uni(12, 30); B(0) = A(0) + uni(-1, 1); B(1) = A(1) + uni(0.5, 1.5); B(2) = A(2) + uni(-1, 1);
    for (int k = 0; k < NKF; k += 2) {
      gfpl::Vector2d ua = project(map.map_keyframes[k], A), ub = project(map.map_keyframes[k], B);
      gfpl::Vector3d le; le(0) = ua(1) - ub(1); le(1) = ub(0) - ua(0); le(2) = ua(0) * ub(1) - ua(1) * ub(0);
      const double n = std::sqrt(le(0) * le(0) + le(1) * le(1));
      for (int i = 0; i < 3; i++) le(i) /= n;
      ml->obs_list.push_back(le); ml->kf_obs_list.push_back(k); ml->sigma_list.push_back(1.0);
    }
    for (int i = 0; i < 3; i++) { ml->line3D(i) = A(i) + 0.05 * gauss(); ml->line3D(3 + i) = B(i) + 0.05 * gauss(); }
    map.map_lines.push_back(ml);
  }
  auto rms = [&]() { double s = 0; for (int j = 0; j < NPT; j++) for (int i = 0; i < 3; i++) { double d = map.map_points[j]->point3D(i) - truth[j](i); s += d * d; } return std::sqrt(s / NPT); };

  // ---- replica of the LM loop on the ORACLE assembly with a plain dense solve (checks the GPU assembly, the
  //      Schur-complement solve and the loop logic of the facade against an independent route) -------------------
  const int Nkf = NKF - 1, N = 6 * Nkf + 3 * NPT + 6 * NLS;
  std::vector<double> X(N), Xls_map(6 * NLS);
  std::vector<int32_t> pt_lm, pt_kf, pt_fix, ls_lm, ls_kf, ls_fix;
  std::vector<double> pt_uv, pt_s2, ls_le, ls_s2, fixT(16);
  map.map_keyframes[0]->T_kf_w.to_row_major(fixT.data());
  for (int k = 0; k < Nkf; k++) for (int i = 0; i < 6; i++) X[6 * k + i] = map.map_keyframes[k + 1]->x_kf_w(i);
  for (int j = 0; j < NPT; j++) {
    for (int i = 0; i < 3; i++) X[6 * Nkf + 3 * j + i] = map.map_points[j]->point3D(i);
    for (size_t o = 0; o < map.map_points[j]->obs_list.size(); o++) {
      const int kf = map.map_points[j]->kf_obs_list[o];
      pt_lm.push_back(j); pt_kf.push_back(kf - 1); pt_fix.push_back(0);
      pt_uv.push_back(map.map_points[j]->obs_list[o](0)); pt_uv.push_back(map.map_points[j]->obs_list[o](1)); pt_s2.push_back(1.0);
    }
  }
  for (int j = 0; j < NLS; j++) {
    for (int i = 0; i < 6; i++) { X[6 * Nkf + 3 * NPT + 6 * j + i] = map.map_lines[j]->line3D(i); Xls_map[6 * j + i] = map.map_lines[j]->line3D(i); }
    for (size_t o = 0; o < map.map_lines[j]->obs_list.size(); o++) {
      const int kf = map.map_lines[j]->kf_obs_list[o];
      ls_lm.push_back(j); ls_kf.push_back(kf - 1); ls_fix.push_back(0);
      for (int i = 0; i < 3; i++) ls_le.push_back(map.map_lines[j]->obs_list[o](i));
      ls_s2.push_back(1.0);
    }
  }
  const int npo = (int)pt_lm.size(), nlo = (int)ls_lm.size();
  gfpl_camera ccam = cam.c_abi();
  gfpl_config cfg;
  orc_config_default(&cfg);
  std::vector<double> kfT(16 * Nkf), Xpt(3 * NPT), Hkk(36 * Nkf), Hpp(9 * NPT), Hll(36 * NLS), Hpk(18 * npo), Hlk(36 * nlo), g(N), H((size_t)N * N), DX(N);
  double errsum = 0;
  auto assemble = [&](bool first) {
    for (int k = 0; k < Nkf; k++) {
      gfpl::Matrix4d T;
      if (first) T = map.map_keyframes[k + 1]->T_kf_w;
      else { PLSLAM::Vector6d x; for (int i = 0; i < 6; i++) x(i) = X[6 * k + i]; T = PLSLAM::expmap_se3(x); }
      T.to_row_major(&kfT[16 * k]);
    }
    for (int i = 0; i < 3 * NPT; i++) Xpt[i] = X[6 * Nkf + i];
    gfpl_lba_problem pb{Nkf, NPT, NLS, npo, nlo, kfT.data(), Xpt.data(), Xls_map.data(), pt_lm.data(), pt_kf.data(), pt_fix.data(), pt_uv.data(),
                        pt_s2.data(), ls_lm.data(), ls_kf.data(), ls_fix.data(), ls_le.data(), ls_s2.data(), fixT.data(), 1};
    gfpl_lba_blocks out{Hkk.data(), Hpp.data(), Hll.data(), Hpk.data(), Hlk.data(), g.data(), &errsum};
    orc_lba_assemble(&ccam, &cfg, &pb, &out);
    std::fill(H.begin(), H.end(), 0.0);
    for (int k = 0; k < Nkf; k++) for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H[(size_t)(6 * k + i) * N + 6 * k + j] = Hkk[36 * k + i * 6 + j];
    for (int l = 0; l < NPT; l++) for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) H[(size_t)(6 * Nkf + 3 * l + i) * N + 6 * Nkf + 3 * l + j] = Hpp[9 * l + i * 3 + j];
    for (int l = 0; l < NLS; l++) for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) H[(size_t)(6 * Nkf + 3 * NPT + 6 * l + i) * N + 6 * Nkf + 3 * NPT + 6 * l + j] = Hll[36 * l + i * 6 + j];
    for (int o = 0; o < npo; o++) if (pt_kf[o] >= 0) for (int i = 0; i < 3; i++) for (int j = 0; j < 6; j++) {
      H[(size_t)(6 * Nkf + 3 * pt_lm[o] + i) * N + 6 * pt_kf[o] + j] += Hpk[18 * o + i * 6 + j];
      H[(size_t)(6 * pt_kf[o] + j) * N + 6 * Nkf + 3 * pt_lm[o] + i] += Hpk[18 * o + i * 6 + j];
    }
    for (int o = 0; o < nlo; o++) if (ls_kf[o] >= 0) for (int i = 0; i < 6; i++) for (int j = 0; j < 6; j++) {
      H[(size_t)(6 * Nkf + 3 * NPT + 6 * ls_lm[o] + i) * N + 6 * ls_kf[o] + j] += Hlk[36 * o + i * 6 + j];
      H[(size_t)(6 * ls_kf[o] + j) * N + 6 * Nkf + 3 * NPT + 6 * ls_lm[o] + i] += Hlk[36 * o + i * 6 + j];
    }
  };
  auto solve = [&](double lambda) {   // dense Gaussian elimination with partial pivoting on (H + lambda diag(H))
    std::vector<double> A = H, b = g;
    for (int i = 0; i < N; i++) A[(size_t)i * N + i] += lambda * H[(size_t)i * N + i];
    for (int k = 0; k < N; k++) {
      int p = k;
      for (int i = k + 1; i < N; i++) if (std::fabs(A[(size_t)i * N + k]) > std::fabs(A[(size_t)p * N + k])) p = i;
      if (p != k) { for (int j = 0; j < N; j++) std::swap(A[(size_t)k * N + j], A[(size_t)p * N + j]); std::swap(b[k], b[p]); }
      for (int i = k + 1; i < N; i++) {
        const double f = A[(size_t)i * N + k] / A[(size_t)k * N + k];
        if (f == 0.0) continue;
        for (int j = k; j < N; j++) A[(size_t)i * N + j] -= f * A[(size_t)k * N + j];
        b[i] -= f * b[k];
      }
    }
    for (int i = N - 1; i >= 0; i--) { double sv = b[i]; for (int j = i + 1; j < N; j++) sv -= A[(size_t)i * N + j] * DX[j]; DX[i] = sv / A[(size_t)i * N + i]; }
  };
  auto update = [&]() {
    for (int k = 0; k < Nkf; k++) {
      PLSLAM::Vector6d x, dx;
      for (int i = 0; i < 6; i++) { x(i) = X[6 * k + i]; dx(i) = DX[6 * k + i]; }
      PLSLAM::Vector6d xn = PLSLAM::logmap_se3(PLSLAM::expmap_se3(x) * PLSLAM::inverse_se3(PLSLAM::expmap_se3(dx)));
      for (int i = 0; i < 6; i++) X[6 * k + i] = xn(i);
    }
    for (int i = 6 * Nkf; i < N; i++) X[i] += DX[i];
  };
  double lambda = cfg.lambda_lba_lm, err, err_prev;
  assemble(true);
  err = errsum / 0.0;
  int Hmax = 0;
  for (int i = 0; i < N; i++) { const double h = H[(size_t)i * N + i]; if (h > Hmax || h < -Hmax) Hmax = (int)std::fabs(h); }
  lambda *= Hmax;
  solve(lambda); update(); err_prev = err;
  int iters;
  for (iters = 1; iters < cfg.max_iters_lba; iters++) {
    assemble(false);
    err = errsum / (NPT + NLS);
    if (std::fabs(err - err_prev) < 2.220446049250313e-16 || err < 2.220446049250313e-16) break;
    solve(lambda);
    if (err > err_prev) lambda /= cfg.lambda_lba_k; else { lambda *= cfg.lambda_lba_k; update(); }
    double n2 = 0; for (double v : DX) n2 += v * v;
    if (std::sqrt(n2) < 2.220446049250313e-16) break;
    err_prev = err;
  }

  const double before = rms();
  map.localBundleAdjustment();
  const double after = rms();
  printf("facade LBA: %d LM iterations (replica %d), landmark RMS error %.6f -> %.6f m, err %.6f (replica %.6f), lambda0 = %d * 0.001\n",
         map.lba_iterations, iters, before, after, map.lba_last_err, err, Hmax);
  EXPECT(map.lba_iterations == iters && iters >= 2, "LM iteration count %d vs %d", map.lba_iterations, iters);
  EXPECT(std::fabs(map.lba_last_err - err) <= 1e-9 * std::fabs(err), "final error %.12g vs %.12g", map.lba_last_err, err);
  double dmax = 0, scale = 0;
  for (int j = 0; j < NPT; j++) for (int i = 0; i < 3; i++) {
    dmax = std::fmax(dmax, std::fabs(map.map_points[j]->point3D(i) - X[6 * Nkf + 3 * j + i]));
    scale = std::fmax(scale, std::fabs(X[6 * Nkf + 3 * j + i]));
  }
  EXPECT(dmax <= 1e-9 * scale, "landmarks after LBA differ from the dense replica by %g (scale %g)", dmax, scale);
  for (int k = 0; k < Nkf; k++) {
    PLSLAM::Vector6d x; for (int i = 0; i < 6; i++) x(i) = X[6 * k + i];
    gfpl::Matrix4d T = PLSLAM::expmap_se3(x);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 4; j++)
      EXPECT(std::fabs(T(i, j) - map.map_keyframes[k + 1]->T_kf_w(i, j)) < 1e-9, "KF %d pose after LBA", k + 1);
  }
  EXPECT(after <= before, "heavily damped LM step does not increase the landmark error (%.6f -> %.6f)", before, after);
  return 0;
}

int main(int argc, char** argv) {
  if (argc > 1 && std::string(argv[1]) == "--no-gpu") {
    PinholeStereoCamera cam(752, 480, 458.654, 458.654, 367.215, 248.375, 0.110078);
    try {
      GfplContext gf(cam.c_abi());
      printf("a CUDA device is present; nothing to check in --no-gpu mode\n");
      return 0;
    } catch (const GfplError& e) {
      printf("no device -> GfplError %d: %s\n", e.code, e.what());
      return e.code == GFPL_E_NO_DEVICE ? 0 : 1;
    }
  }
  test_tracking();
  test_lba();
  printf(fails ? "FACADE PARITY FAILED (%d)\n" : "FACADE PARITY OK\n", fails);
  return fails ? 1 : 0;
}
