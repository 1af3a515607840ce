/*
 * gfpl_c.h — C ABI of the B200-native tracking hot path of GF-PL-SLAM.
 *
 * This is the drop-in boundary.  The reference has no FFI of its
 * own: its boundary is the C++ class surface of StVO::StereoFrame,
 * StVO::StereoFrameHandler and PLSLAM::MapHandler.  Every entry point below
 * names the reference member(s) it replaces (file:line under the reference
 * tree).  The C++ facades in gf-pl-slam_b200/host/ keep the reference's class
 * and member names and forward to these functions.
 *
 * Conventions
 *   - plain C, plain pointers and sizes, no torch / Eigen / OpenCV types;
 *   - all matrices are ROW-MAJOR fp64 (Eigen's default is column-major: the
 *     facades transpose);
 *   - every function returns 0 (GFPL_OK) or a negative GFPL_E_* code and
 *     records a message retrievable with gfpl_last_error();
 *   - numerical failure of a frame pair is DATA (gfpl_pose_result::status),
 *     not an error;
 *   - the caller owns every host buffer; a gfpl_ctx owns all device memory,
 *     streams and events it creates; nothing allocated on one side is freed on
 *     the other;
 *   - "batched" = B independent problems with a common capacity (stride)
 *     `cap_*`; per-problem counts live in `n_*[B]` arrays;
 *   - there is NO CPU fallback: if no sm_100 device is present the context
 *     cannot be created and every compute entry point fails with
 *     GFPL_E_NO_DEVICE.
 */
#ifndef GFPL_C_H
#define GFPL_C_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GFPL_ABI_VERSION 1
/* Preconditions checked by every batched entry point (GFPL_E_INVALID otherwise): each per-problem count n_*[p] lies in
 * [0, cap_*] (the kernels size shared memory and output rows by the capacities), and no side of a matching problem
 * has GFPL_MAX_KNN_ROWS rows or more (the (distance, row) sort key holds the row index in 20 bits). */
#define GFPL_MAX_KNN_ROWS (1 << 20)
#define GFPL_DESC_BYTES 32 /* ORB (ORBextractor.cc:1037,1068) and LBD (binary_descriptor_custom.cpp:639) rows */

/* cv::NormTypes values used by the reference's BFMatcher constructions */
#define GFPL_NORM_HAMMING 6  /* stereoFrame.cpp:176,798; stereoFrameHandler.cpp:465,609,703,777 */
#define GFPL_NORM_HAMMING2 7 /* stereoFrame.cpp:636,885 */

enum gfpl_status_code {
  GFPL_OK = 0,
  GFPL_E_INVALID = -1,     /* bad argument (null pointer, negative size, Nt < 2 for knn k=2 ...) */
  GFPL_E_NO_DEVICE = -2,   /* no CUDA device / not sm_100 */
  GFPL_E_CUDA = -3,        /* a CUDA runtime call failed; see gfpl_last_error */
  GFPL_E_NOMEM = -4,
  GFPL_E_CAPACITY = -5     /* an output capacity was too small (count still reported) */
};

/* ------------------------------------------------------------------------- */
/* Camera and configuration                                                  */
/* ------------------------------------------------------------------------- */

/* PinholeStereoCamera after rectification (pinholeStereoCamera.cpp:133-170). */
typedef struct gfpl_camera {
  double fx, fy, cx, cy, b;
  int32_t width, height;
} gfpl_camera;

/* Mirror of the hot-path fields of the reference's Config singleton
 * (config.cpp:26-154, accessors config.h:36-143) with identical defaults. */
typedef struct gfpl_config {
  /* flags */
  int32_t has_points;          /* config.cpp:75 */
  int32_t has_lines;           /* :76 */
  int32_t lr_in_parallel;      /* :77 (selects radius vs knn in crossFrameMatching_Hybrid) */
  int32_t best_lr_matches;     /* :79 */
  int32_t use_line_conf_cut;   /* :83 */
  int32_t max_vol_line_cut;    /* :86  (cutWithMaxVol) */
  int32_t max_line_match_num;  /* :94  300 */
  int32_t max_point_match_num; /* :95  500 */
  int32_t min_features;        /* :122 10 */
  int32_t max_iters;           /* :124 5 */
  int32_t max_iters_ref;       /* :125 10 */
  int32_t max_iters_lba;       /* :55  20 */
  /* matching thresholds */
  double ratio_disp_std;       /* :84 0.15 */
  double ratio_disp_std_hor;   /* :85 0.9 */
  double max_dist_epip;        /* :101 2.0 */
  double min_disp;             /* :102 1.0 */
  double max_ratio_12_p;       /* :103 0.9 */
  double point_match_radius;   /* :104 50 */
  double stereo_overlap_th;    /* :106 0.5 */
  double line_horiz_th;        /* :108 0.1 */
  double desc_th_l;            /* :109 0.1 */
  double line_cov_th;          /* :110 10 */
  double line_match_radius;    /* :111 80 */
  /* optimisation */
  double homog_th;             /* :121 1e-7 */
  double min_error;            /* :126 1e-7 */
  double min_error_change;     /* :127 1e-7 */
  double inlier_k;             /* :128 2 */
  double motion_step_th;       /* :129 10 */
  double orb_scale_factor;     /* :135 1.2 */
  double lsd_scale;            /* :146 1 */
  double lambda_lba_lm;        /* :53 0.001 */
  double lambda_lba_k;         /* :54 10 */
  /* constants that are literals in the reference source */
  double f2f_point_window;     /* 10.0, stereoFrameHandler.cpp:534 */
  double budget_slack;         /* 1.2,  stereoFrameHandler.cpp:659,748,827 */
  double cut_step;             /* 0.05, stereoFrameHandler.cpp:135 */
  double cut_rng[2];           /* {0,1}, stereoFrameHandler.cpp:134 */
  /* defined semantics for reference UB */
  double mad_uninit_value;     /* D1: value of the uninitialised nn_dist_median, default 0.0 */
  /* map matching (MapHandler::lookForCommonMatches) */
  double max_kf_epip_p;        /* config.cpp:42 1.0 */
  double max_kf_epip_l;        /* config.cpp:43 1.0 */
  int32_t orb_nlevels;         /* config.cpp:136 4 */
  int32_t _pad0;
} gfpl_config;

void gfpl_config_default(gfpl_config* cfg);
/* EuRoC (config/euroc_params.yaml) and KITTI (config/kitti/kitti00-02.yaml) rectified cameras */
void gfpl_camera_euroc(gfpl_camera* cam);
void gfpl_camera_kitti(gfpl_camera* cam);

/* ------------------------------------------------------------------------- */
/* Feature records                                                           */
/* ------------------------------------------------------------------------- */

/* Subset of cv::KeyPoint the path reads (pt.x, pt.y, octave). 16 B */
typedef struct gfpl_keypoint {
  float x, y;
  int32_t octave;
  int32_t class_id; /* written by the stereo stage like points_l[].class_id (stereoFrame.cpp:867-869) */
} gfpl_keypoint;

/* Subset of cv::line_descriptor::KeyLine the path reads (descriptor_custom.hpp:105-170). 32 B */
typedef struct gfpl_keyline {
  float startPointX, startPointY, endPointX, endPointY;
  float angle;
  int32_t octave;
  float lineLength;
  float response;
} gfpl_keyline;

/* cv::DMatch */
typedef struct gfpl_dmatch {
  int32_t queryIdx, trainIdx, imgIdx;
  float distance;
} gfpl_dmatch;

/* StVO::PointFeature (stereoFeatures.h:36-62), fields on the path. 72 B */
typedef struct gfpl_point_feature {
  double P[3];
  double pl[2];
  double disp;
  double sigma2;
  int32_t level;
  int32_t idx;          /* reference `idx` */
  int32_t src_l, src_r; /* raw left / right keypoint index the feature came from */
} gfpl_point_feature;

/* StVO::LineFeature (stereoFeatures.h:64-128), fields on the path. 152 B */
typedef struct gfpl_line_feature {
  double sP[3], eP[3];
  double spl[2], epl[2];
  double sdisp, edisp;
  double le[3];
  double angle;
  double sigma2;
  int32_t level;
  int32_t idx;
  int32_t src_l, src_r;
} gfpl_line_feature;

/* covSpt3D / covEpt3D of a LineFeature (stereoFeatures.h:108), row-major 3x3. 144 B */
typedef struct gfpl_line_cov {
  double covS[9], covE[9];
} gfpl_line_cov;

/* What optimizeFunctions reads from one entry of matched_pt. 56 B */
typedef struct gfpl_pt_obs {
  double P[3];
  double pl_obs[2];
  double sigma2;
  int32_t prev_idx, curr_idx;
} gfpl_pt_obs;

/* What optimizeFunctions reads from one entry of matched_ls. 88 B */
typedef struct gfpl_ln_obs {
  double sP[3], eP[3];
  double le_obs[3];
  double sigma2;
  int32_t prev_idx, curr_idx;
} gfpl_ln_obs;

/* Per matched line outputs of the cut search (LineFeature::cutRatio, invCovPose and
 * the fields updateEndPointByRatio rewrites, stereoFrameHandler.cpp:1451-1470). 352 B */
typedef struct gfpl_line_cut {
  double cutRatio[2];
  double invCovPose[36];
  double spl[2], epl[2];
  double sdisp, edisp;
} gfpl_line_cut;

/* Per matched line outputs of StereoFrameHandler::estimateProjUncertainty_descent (stereoFrameHandler.cpp:1476-1610):
 * cutRatio, the last entry and the length of volumeArr, and the fields :1567-1606 rewrite — covlPt2D[1] / [2] (entries 0
 * and 3 are the identity), lPt2D[0..3], fst_idx / lst_idx, spl / epl / sdisp / edisp of the cut end points.  296 B */
typedef struct gfpl_line_cut_descent {
  double cutRatio[2];
  double volume;              /* volumeArr.back(): product of the two largest eigenvalues of the pose covariance */
  int32_t n_volume;           /* volumeArr.size() */
  int32_t fst_idx, lst_idx;   /* 0, 3 (:1599-1600) */
  int32_t pad;
  double covlPt2D[2][9];      /* row-major 3x3 */
  double lPt2D[4][2];
  double spl[2], epl[2];
  double sdisp, edisp;
} gfpl_line_cut_descent;

/* The fields of a PREVIOUS frame's stereo features that the tracking step reads (f2f matching: P, sigma2, descriptor;
 * line uncertainty / cut / GN: sP, eP, spl, epl, sdisp, edisp, le[0], sigma2).  A streaming replay uploads these slim
 * records instead of the full gfpl_point_feature / gfpl_line_feature (32 B vs 72 B, 112 B vs 152 B per feature). */
typedef struct gfpl_prev_point {
  double P[3];
  double sigma2;
} gfpl_prev_point;
typedef struct gfpl_prev_line {
  double sP[3], eP[3];
  double spl[2], epl[2];
  double sdisp, edisp;
  double le0;     /* le[0]: selects ratioDispSTD vs ratioDispSTDHor (stereoFrame.cpp:1462) */
  double sigma2;
} gfpl_prev_line;
typedef struct gfpl_prev_frames {
  int32_t B, cap_pt, cap_ln;
  const int32_t *n_pt, *n_ln;    /* [B] */
  const gfpl_prev_point* pt;     /* [B][cap_pt] */
  const uint8_t* pdesc;          /* [B][cap_pt][32] */
  const gfpl_prev_line* ln;      /* [B][cap_ln] */
  const uint8_t* ldesc;          /* [B][cap_ln][32] */
} gfpl_prev_frames;

/* Compact host-side input of one tracking step (streaming replay): the same information as gfpl_raw_frames +
 * gfpl_prev_frames in fewer bytes — the host-to-device copy is what bounds the end-to-end rate.  Pyramid levels
 * (octave / level, 0..255) travel as separate byte arrays.
 *  - key point: pt.x, pt.y (8 B) + octave (1 B); class_id is not read by the path,
 *  - key line: start / end point, angle (20 B) + octave (1 B); lineLength, response are consumed by the detector-side
 *    truncation of stereoFrame.cpp:1093-1102 only,
 *  - previous stereo point: pl.x, pl.y, disp as the floats the stereo stage produced them from (12 B) + level (1 B):
 *    P = backProjection(pl, disp) (pinholeStereoCamera.cpp:133-141) and sigma2 (stereoFeatures.cpp:41-48) are
 *    recomputed on the device by the same expressions, bit for bit.  Only valid for stereo points whose pl / disp are
 *    float values (extractStereoFeatures_* output); otherwise use gfpl_prev_frames,
 *  - previous stereo line: the left end points as the floats of the key line and the two end-point disparities (32 B)
 *    + level (1 B): sP / eP = backProjection(end point, disparity), le[0] (normalised line equation, stereoFrame.cpp:
 *    924-927) and sigma2 are recomputed on the device by the stereo stage's expressions.  Only valid for stereo lines
 *    that were not cut (updateEndPointByRatio edits sP / eP / spl / epl); otherwise use gfpl_prev_frames. */
typedef struct gfpl_keypoint8 { float x, y; } gfpl_keypoint8;
typedef struct gfpl_keyline20 { float startPointX, startPointY, endPointX, endPointY, angle; } gfpl_keyline20;
typedef struct gfpl_prev_point12 { float x, y, disp; } gfpl_prev_point12;
typedef struct gfpl_prev_line32 { float spl[2], epl[2]; double sdisp, edisp; } gfpl_prev_line32;
typedef struct gfpl_compact_input {
  int32_t B, cap_pt, cap_ln, prev_cap_pt, prev_cap_ln;
  const int32_t *n_pt_l, *n_pt_r, *n_ln_l, *n_ln_r;   /* [B] */
  const gfpl_keypoint8 *pt_l, *pt_r;                   /* [B][cap_pt] */
  const uint8_t *pt_oct_l, *pt_oct_r;                  /* [B][cap_pt] */
  const uint8_t *pdesc_l, *pdesc_r;                    /* [B][cap_pt][32] */
  const gfpl_keyline20 *ln_l, *ln_r;                   /* [B][cap_ln] */
  const uint8_t *ln_oct_l, *ln_oct_r;                  /* [B][cap_ln] */
  const uint8_t *ldesc_l, *ldesc_r;                    /* [B][cap_ln][32] */
  const int32_t *prev_n_pt, *prev_n_ln;                /* [B] */
  const gfpl_prev_point12* prev_pt;                    /* [B][prev_cap_pt] */
  const uint8_t* prev_pt_level;                        /* [B][prev_cap_pt] */
  const uint8_t* prev_pdesc;                           /* [B][prev_cap_pt][32] */
  const gfpl_prev_line32* prev_ln;                     /* [B][prev_cap_ln] */
  const uint8_t* prev_ln_level;                        /* [B][prev_cap_ln] */
  const uint8_t* prev_ldesc;                           /* [B][prev_cap_ln][32] */
} gfpl_compact_input;

/* prev_frame->{Tfw,DT,DT_cov,Tfw_cov,time_stamp} (stereoFrame.h:212-221) + curr time stamp */
typedef struct gfpl_frame_state {
  double Tfw[16];
  double DT[16];
  double DT_cov[36];
  double Tfw_cov[36];
  double time_stamp;
  double time_stamp_curr;
} gfpl_frame_state;

/* status bits of gfpl_pose_result::status */
#define GFPL_POSE_OK 0
#define GFPL_POSE_FEW_FEATURES 1  /* n_inliers <= min_features before the first stage (:1878,1953) */
#define GFPL_POSE_NONFINITE_1 2   /* first-stage result not finite (:1884,1959) */
#define GFPL_POSE_FEW_INLIERS 4   /* too few inliers after removeOutliers (:1888,1963) */
#define GFPL_POSE_NONFINITE_2 8   /* final DT not finite -> identity, err_norm=-1, numFrameLoss++ (:2016-2028) */
#define GFPL_POSE_STEP_REJECTED 16 /* motion-step gate failed (:1992-2012) */

/* curr_frame->{DT,Tfw,DT_cov,DT_cov_eig,Tfw_cov,err_norm} + counters */
typedef struct gfpl_pose_result {
  double DT[16];
  double Tfw[16];
  double DT_cov[36];
  double DT_cov_eig[6];
  double Tfw_cov[36];
  double err_norm;
  int32_t status;
  int32_t n_inliers, n_inliers_pt, n_inliers_ls;
  int32_t iters_first, iters_ref;
} gfpl_pose_result;

/* ------------------------------------------------------------------------- */
/* Batched views. Pointers are HOST pointers in every gfpl_* call below.      */
/* ------------------------------------------------------------------------- */

/* points_l/r, pdesc_l/r, lines_l/r, ldesc_l/r of B StereoFrames (stereoFrame.h:223-228) */
typedef struct gfpl_raw_frames {
  int32_t B, cap_pt, cap_ln;
  const int32_t *n_pt_l, *n_pt_r, *n_ln_l, *n_ln_r; /* [B] */
  const gfpl_keypoint *pt_l, *pt_r;                 /* [B][cap_pt] */
  const uint8_t *pdesc_l, *pdesc_r;                 /* [B][cap_pt][32] */
  const gfpl_keyline *ln_l, *ln_r;                  /* [B][cap_ln] */
  const uint8_t *ldesc_l, *ldesc_r;                 /* [B][cap_ln][32] */
} gfpl_raw_frames;

/* stereo_pt / stereo_ls and the compacted pdesc_l / ldesc_l of B StereoFrames */
typedef struct gfpl_stereo_frames {
  int32_t B, cap_pt, cap_ln;
  int32_t *n_pt, *n_ln;     /* [B] */
  gfpl_point_feature* pt;   /* [B][cap_pt] */
  uint8_t* pdesc;           /* [B][cap_pt][32], row k <-> pt[k] (stereoFrame.cpp:878) */
  gfpl_line_feature* ln;    /* [B][cap_ln] */
  uint8_t* ldesc;           /* [B][cap_ln][32] (stereoFrame.cpp:1013) */
} gfpl_stereo_frames;

/* matched_pt / matched_ls of B StereoFrameHandlers, as flat observation lists */
typedef struct gfpl_track_obs {
  int32_t B, cap_pt, cap_ln;
  int32_t *n_pt, *n_ln;     /* [B] */
  gfpl_pt_obs* pt;          /* [B][cap_pt] */
  uint8_t* pt_inlier;       /* [B][cap_pt] */
  gfpl_ln_obs* ln;          /* [B][cap_ln] */
  uint8_t* ln_inlier;       /* [B][cap_ln] */
} gfpl_track_obs;

/* ------------------------------------------------------------------------- */
/* Context                                                                   */
/* ------------------------------------------------------------------------- */

typedef struct gfpl_ctx gfpl_ctx;

int gfpl_abi_version(void);
/* device = CUDA ordinal.  Fails with GFPL_E_NO_DEVICE when CUDA is unavailable. */
int gfpl_ctx_create(int device, const gfpl_camera* cam, const gfpl_config* cfg, gfpl_ctx** out);
void gfpl_ctx_destroy(gfpl_ctx* ctx);
int gfpl_ctx_set_config(gfpl_ctx* ctx, const gfpl_config* cfg);
int gfpl_ctx_set_camera(gfpl_ctx* ctx, const gfpl_camera* cam);
const char* gfpl_last_error(const gfpl_ctx* ctx); /* ctx may be NULL: thread-local last error of ctx-less calls */
int gfpl_ctx_sync(gfpl_ctx* ctx);
/* the context's cudaStream_t: every entry point orders its device work on it (the pipelined replay included), so a caller
 * can enqueue its own work behind the library's without a host synchronisation — e.g. the NCCL gather of the pose
 * records of a sharded replay */
void* gfpl_ctx_stream(gfpl_ctx* ctx);
/* number of kernels this context has launched since creation (bench.py gpu_launches) */
int64_t gfpl_ctx_launch_count(const gfpl_ctx* ctx);

/* ------------------------------------------------------------------------- */
/* M1/M2/M3 — brute-force 256-bit Hamming matching                            */
/* ------------------------------------------------------------------------- */

/* Replaces cv::BFMatcher(norm,false)::knnMatch(q,t,out,2) behind
 * StereoFrame::matchPointFeatures (stereoFrame.cpp:1229-1234) and matchLineFeatures
 * (:1236-1241) and the direct call sites listed in SURVEY.md §8b.
 * out_idx[2*i+k], out_dist[2*i+k]: k-th nearest train row of query i, ascending
 * distance, ties -> ascending trainIdx.  Requires nt >= 2. */
int gfpl_knn2_hamming256(gfpl_ctx* ctx, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt,
                         int32_t norm, int32_t* out_idx, float* out_dist);

/* Batched form: problem p has nq[p] rows at q + p*cap_q*32 and nt[p] rows at t + p*cap_t*32;
 * outputs at out_* + p*cap_q*2.  Problems with nt[p] < 2 produce idx = -1, dist = +inf. */
int gfpl_knn2_hamming256_batch(gfpl_ctx* ctx, int32_t B, const uint8_t* q, const int32_t* nq, int32_t cap_q,
                               const uint8_t* t, const int32_t* nt, int32_t cap_t, int32_t norm,
                               int32_t* out_idx, float* out_dist);

/* Both match directions of a mutual-best-match test from one pass over the distance matrix: the two knnMatch calls of
 * stereoFrame.cpp:807-808 / stereoFrameHandler.cpp:713-716 / mapHandler.cpp:225-226, of which the accept loops read
 * [i][0], [i][1] of the forward direction and only [j][0] of the reverse one (:836, :735, :253).
 * out_idx / out_dist: [B][cap_q][2] as gfpl_knn2_hamming256_batch; out_best_idx / out_best_dist: [B][cap_t] nearest query
 * row of every train row (ties -> lowest query index), -1 / +inf when nq[p] == 0.  NORM_HAMMING runs on the tensor cores
 * (int8 +-64 product, exact); problems of at most 128 rows per side run there for both norms (NORM_HAMMING2 as a
 * K = 384 product), larger NORM_HAMMING2 problems and tiny ones on the integer pipes. */
int gfpl_knn2_mutual_batch(gfpl_ctx* ctx, int32_t B, const uint8_t* q, const int32_t* nq, int32_t cap_q,
                           const uint8_t* t, const int32_t* nt, int32_t cap_t, int32_t norm, int32_t* out_idx,
                           float* out_dist, int32_t* out_best_idx, float* out_best_dist);

/* Replaces cv::BFMatcher::radiusMatch behind matchPointFeatures_radius
 * (stereoFrame.cpp:1243-1250) and matchLineFeatures_radius (:1252-1257).
 * CSR output: query i owns entries [offsets[i], offsets[i+1]); each list holds every
 * train row with distance <= radius in ascending (distance, trainIdx) order.
 * `cap` = capacity of out_idx/out_dist; if the total exceeds it the call returns
 * GFPL_E_CAPACITY with offsets filled (so the caller can re-size). */
int gfpl_radius_hamming256(gfpl_ctx* ctx, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt,
                           int32_t norm, float radius, int32_t* offsets /*nq+1*/, int32_t* out_idx,
                           float* out_dist, int32_t cap);

/* ------------------------------------------------------------------------- */
/* M4/M5/M6 — stereo L/R matching and accept loops                            */
/* ------------------------------------------------------------------------- */

#define GFPL_STEREO_INITIAL 0 /* StereoFrame::extractInitialStereoFeatures, stereoFrame.cpp:173-334 */
#define GFPL_STEREO_BRUTE 1   /* StereoFrame::extractStereoFeatures_Brute,  stereoFrame.cpp:795-1014
                                 (its line half is identical to _ORBSLAM :633-767) */

/* Matching half of extractInitialStereoFeatures / extractStereoFeatures_Brute for B frames:
 * knn k=2 L->R and R->L, mutual check, ratio / gap test, epipolar, disparity, overlap and
 * covariance gates, back-projection, descriptor compaction.  Detection is injected through `raw`. */
int gfpl_stereo_match(gfpl_ctx* ctx, const gfpl_raw_frames* raw, int32_t variant, gfpl_stereo_frames* out);

/* StereoFrame::extractStereoFeatures_ORBSLAM after detectFeatures (stereoFrame.cpp:453-767) — the
 * point stereo the reference runs by default (stereoFrameHandler.cpp:91).  Points: for every left key point the right
 * key points of its row band (|octave difference| <= 1, uL - fx <= uR <= uL), best StereoFrame::descriptorDistance
 * (stereoFrame.h:185-201, < 100, first minimum), accepted below thOrbDist = 80; 11x11 SAD sliding window +-5 px with
 * parabola refinement on the pyramid level of the left key point (subPixelStereoRefine_ORBSLAM :340-404, which reads
 * BOTH windows from the RIGHT extractor's pyramid, :358/:376 — reproduced); survivors sorted by (distance, left index),
 * cut at 1.5*1.4*median distance; stereo_pt comes out in that order.  Lines: identical to GFPL_STEREO_BRUTE (:633-767).
 * Defined where the reference would throw or read out of range: a window that leaves its pyramid level (cv::Mat
 * rowRange / colRange assert) counts as "no refinement" (disparity -1); no accepted point at all (vDistIdx[0] of an
 * empty vector, :587) gives an empty stereo_pt. */
#define GFPL_STEREO_ORBSLAM 2 /* names the variant for callers; gfpl_stereo_match itself takes 0 or 1 only */
#define GFPL_MAX_PYR_LEVELS 8
typedef struct gfpl_pyramid {
  int32_t n_levels;                      /* == cfg.orb_nlevels */
  int32_t width[GFPL_MAX_PYR_LEVELS];    /* mvImagePyramid[l].cols */
  int32_t height[GFPL_MAX_PYR_LEVELS];   /* mvImagePyramid[l].rows */
  int32_t row_stride[GFPL_MAX_PYR_LEVELS]; /* bytes between rows of level l (Mat::step) */
  int64_t offset[GFPL_MAX_PYR_LEVELS];   /* byte offset of pixel (0,0) of level l inside one frame's block */
  int64_t frame_stride;                  /* bytes between the blocks of consecutive frames */
  const uint8_t* data;                   /* [B] blocks: orbExtractor_r->mvImagePyramid of every frame (CV_8UC1) */
} gfpl_pyramid;
int gfpl_stereo_match_orbslam(gfpl_ctx* ctx, const gfpl_raw_frames* raw, const gfpl_pyramid* right_pyramid,
                              gfpl_stereo_frames* out);

/* StereoFrame::estimateStereoUncertainty (stereoFrame.cpp:1448-1484): covSpt3D/covEpt3D for
 * every stereo line. out: [B][cap_ln]. */
int gfpl_stereo_uncertainty(gfpl_ctx* ctx, const gfpl_stereo_frames* frames, gfpl_line_cov* out);

/* ------------------------------------------------------------------------- */
/* M7/M8 — frame-to-frame matching                                            */
/* ------------------------------------------------------------------------- */

#define GFPL_F2F_HYBRID 0 /* StereoFrameHandler::crossFrameMatching_Hybrid, stereoFrameHandler.cpp:451-695 */
#define GFPL_F2F_BRUTE 1  /* StereoFrameHandler::crossFrameMatching_Brute,  stereoFrameHandler.cpp:698-860 */

/* prev/curr: stereo frames of B frame pairs; state[p].Tfw / .DT = prev_frame->Tfw / DT.  The
 * predicted curr->Tfw = prev->Tfw * prev->DT (predictFramePose, :153-157) is formed inside.
 * out: matched_pt / matched_ls as observation lists (all inlier flags = 1);
 * out->cap_pt >= max_point_match_num, out->cap_ln >= max_line_match_num. */
int gfpl_f2f_match(gfpl_ctx* ctx, const gfpl_stereo_frames* prev, const gfpl_stereo_frames* curr,
                   const gfpl_frame_state* state /*[B]*/, int32_t mode, gfpl_track_obs* out);

/* Key-frame to key-frame matching of MapHandler::lookForCommonMatches (mapHandler.cpp:211-490):
 * knn k=2 both ways (NORM_HAMMING), mutual check, ratio test (points) / MAD gap test (lines), and the chi-square
 * reprojection gate  error * sqrt(sigma2) < sqrt(7.815)  with DT = inverse_se3(kf1->T_kf_w) * kf0->T_kf_w
 * (:197, row-major [B][16]).  out: accepted pairs in ascending kf0 index (prev_idx = kf0 index, curr_idx = kf1
 * index; P/sP/eP/sigma2 from kf0, pl_obs/le_obs from kf1); out->cap_pt >= kf0->cap_pt, out->cap_ln >= kf0->cap_ln.
 * The map bookkeeping that follows each accepted pair (new MapPoint / MapLine, graphs) stays on the host. */
int gfpl_kf_match(gfpl_ctx* ctx, const gfpl_stereo_frames* kf0, const gfpl_stereo_frames* kf1,
                  const double* DT /*[B][16]*/, gfpl_track_obs* out);

/* Local map <-> key-frame half of MapHandler::lookForCommonMatches (mapHandler.cpp:509-770).
 * map: the landmarks the host walk of :516-521 / :636-643 keeps (local && last observer != kf1) with their
 * representative descriptor med_desc.row(0); kf1: the key frame's stereo features, `idx == -1` marks the ones the
 * KF<->KF half left unmatched (:538-546, :665-673).  On the device: the in-image test of :524-532 / :645-657 on
 * Twf * X, compaction of both sides, knn k=2 both ways (NORM_HAMMING), mutual check, ratio <= maxRatio12P (points)
 * or gap > lineDescriptorMAD * descThL (lines), depth > 0, and the gates of :596 (|proj - pl| < maxKFEpipP) and :743
 * (signed l.p < maxKFEpipL for both endpoints).  Twf = inverse_se3(kf1->T_kf_w), row-major [B][16].
 * out pairs come in the loop order of :579 / :709 (ascending selected-landmark order = ascending map_idx);
 * map_idx indexes the arrays of `map`, kf_idx the features of `kf1`.  addMap*Observation / full_graph stay on the host. */
typedef struct gfpl_local_map {
  int32_t B, cap_pt, cap_ln;
  const int32_t* n_pt;  /* [B] */
  const int32_t* n_ln;  /* [B] */
  const double* pt3d;   /* [B][cap_pt][3]  MapPoint::point3D (world) */
  const uint8_t* pdesc; /* [B][cap_pt][32] */
  const double* ln3d;   /* [B][cap_ln][6]  MapLine::line3D (world start, end) */
  const uint8_t* ldesc; /* [B][cap_ln][32] */
} gfpl_local_map;

typedef struct gfpl_map_pair { int32_t map_idx, kf_idx; } gfpl_map_pair;

typedef struct gfpl_map_matches {
  int32_t B, cap_pt, cap_ln;  /* cap_pt >= map->cap_pt, cap_ln >= map->cap_ln */
  int32_t* n_pt;              /* [B] */
  int32_t* n_ln;              /* [B] */
  gfpl_map_pair* pt;          /* [B][cap_pt] */
  gfpl_map_pair* ln;          /* [B][cap_ln] */
} gfpl_map_matches;

int gfpl_map_match(gfpl_ctx* ctx, const gfpl_local_map* map, const gfpl_stereo_frames* kf1,
                   const double* Twf /*[B][16]*/, gfpl_map_matches* out);

/* Descriptor part of MapPoint / MapLine::updateAverageDescDir (mapFeatures.cpp:51-82, 128-160):
 * for each of L landmarks, the observation whose ascending NORM_HAMMING distance row (self distance included) has the
 * smallest element int(1 + 0.5 (n-1)); ties -> lowest observation.  desc: all desc_list rows back to back,
 * offsets[L+1] (CSR).  out_idx[l] = position inside landmark l's list (0 when n < 2: the reference only runs it
 * from add*Observation, i.e. with n >= 2).  med_desc = desc_list[out_idx]. */
int gfpl_medoid_desc(gfpl_ctx* ctx, const uint8_t* desc, const int32_t* offsets, int32_t L, int32_t* out_idx);

/* ------------------------------------------------------------------------- */
/* C1-C5 — good line cutting                                                  */
/* ------------------------------------------------------------------------- */

/* StereoFrameHandler::estimateProjUncertainty_submodular (stereoFrameHandler.cpp:1618-1764) with
 * getPoseInfoOnLine (:1342-1411), getPoseInfoPoint (:1414-1447), logdet (linespec.h:43-56) or
 * getMinEigenValue (linespec.cpp:53-66), updateEndPointByRatio (:1451-1470).
 * DT_inv[p] = curr->Tfw^-1 * prev->Tfw (row-major 4x4).  prev_cov: [B][cov_cap] covariances of the
 * PREVIOUS frame's stereo lines, indexed by obs->ln[].prev_idx.  obs->ln[].sP/eP are updated in
 * place; cut: [B][obs->cap_ln]. */
int gfpl_linecut_submodular(gfpl_ctx* ctx, const double* DT_inv /*[B][16]*/, gfpl_track_obs* obs,
                            const gfpl_line_cov* prev_cov, int32_t cov_cap, gfpl_line_cut* cut);

/* C6: StereoFrameHandler::estimateProjUncertainty_descent (stereoFrameHandler.cpp:1476-1610) with getPoseCovOnLine
 * (:1249-1338): per matched line, independent of the other lines, a descent over the 8 neighbour cuts that takes every
 * strictly smaller volume at once (the cut moves inside the neighbour loop, :1545-1550).  The volume is the product of
 * the two largest eigenvalues of the rank-2 pose covariance; the reference takes the real parts of Eigen's general
 * EigenSolver sorted in descending order (getEigenValue, linespec.cpp:37-51) — here: the eigenvalues of the symmetric
 * matrix given by its lower triangle (cyclic Jacobi, as everywhere in this library), sorted, the two largest multiplied.
 * Step and range come from gfpl_config.cut_step / cut_rng (the reference's call site passes 0.05 and rngCutRatio,
 * :127).  Arguments as gfpl_linecut_submodular; obs->ln[].sP/eP are replaced by the cut end points (both taken from
 * the ORIGINAL end points, :1567-1568); out: [B][obs->cap_ln]. */
int gfpl_linecut_descent(gfpl_ctx* ctx, const double* DT_inv /*[B][16]*/, gfpl_track_obs* obs,
                         const gfpl_line_cov* prev_cov, int32_t cov_cap, gfpl_line_cut_descent* out);

/* ------------------------------------------------------------------------- */
/* P1-P4 — pose optimisation                                                  */
/* ------------------------------------------------------------------------- */

/* StereoFrameHandler::optimizeFunctions (stereoFrameHandler.cpp:2118-2245): H (6x6 row-major),
 * g (6), e for B problems at poses DT[p] (4x4 row-major, maps prev -> curr). */
int gfpl_pose_hg(gfpl_ctx* ctx, const double* DT /*[B][16]*/, const gfpl_track_obs* obs,
                 double* H /*[B][36]*/, double* g /*[B][6]*/, double* e /*[B]*/);

/* vector_stdv_mad (auxiliar.cpp:521-537; the threshold of removeOutliers, stereoFrameHandler.cpp:2058-2116) for B arrays:
 * out[p] = 1.4826 * sorted(fabsf(res - sorted(res)[n/2]))[n/2] of res[p][0..n[p]) (0 for n = 0), exact — the
 * order statistics are taken on the IEEE bit patterns.  res: [B][cap], non-negative values, cap < 16384. */
int gfpl_vector_stdv_mad(gfpl_ctx* ctx, const double* res, const int32_t* n /*[B]*/, int32_t B, int32_t cap, double* out /*[B]*/);

#define GFPL_POSE_SEED_PREV_DT 0 /* optimizePose()            stereoFrameHandler.cpp:1864-1937 */
#define GFPL_POSE_SEED_GIVEN 1   /* optimizePose(Matrix4d)    stereoFrameHandler.cpp:1939-2030 (adds the motion-step gate) */

/* optimizePose + gaussNewtonOptimization (:2032-2056) + removeOutliers (:2058-2116).
 * state[p]: prev_frame fields.  DT_ini: [B][16] seeds, used when overload == GFPL_POSE_SEED_GIVEN
 * (may be NULL otherwise).  obs->*_inlier are updated in place.  out: [B]. */
int gfpl_pose_optimize(gfpl_ctx* ctx, gfpl_track_obs* obs, const gfpl_frame_state* state,
                       const double* DT_ini, int32_t overload, gfpl_pose_result* out);

/* ------------------------------------------------------------------------- */
/* Whole tracking step for B independent frame pairs                          */
/* ------------------------------------------------------------------------- */

#define GFPL_TRACK_LINECUT 1       /* run estimateStereoUncertainty + estimateProjUncertainty_submodular */
#define GFPL_TRACK_F2F_BRUTE 2     /* crossFrameMatching_Brute instead of _Hybrid */
#define GFPL_TRACK_SEED_GIVEN 4    /* optimizePose(prev_frame->DT) as plslam_mod.cpp:408 does */
#define GFPL_TRACK_STEREO_INITIAL 8 /* curr stereo stage = extractInitialStereoFeatures */

/* insertStereoPair (stereoFrameHandler.cpp:83-151, with the brute-force stereo stage of :90) followed
 * by optimizePose (plslam_mod.cpp:387,408 / plstvo_dataset.cpp:299-330) for B frame pairs:
 * curr_raw -> curr stereo frame; f2f matching against `prev`; line cut; pose optimisation.
 * Outputs: curr (stereo frames of the current images), obs (matched lists + inlier masks),
 * cut ([B][obs->cap_ln], may be NULL when the cut is off), pose [B]. */
int gfpl_track(gfpl_ctx* ctx, const gfpl_stereo_frames* prev, const gfpl_frame_state* state,
               const gfpl_raw_frames* curr_raw, int32_t flags, gfpl_stereo_frames* curr,
               gfpl_track_obs* obs, gfpl_line_cut* cut, gfpl_pose_result* pose);

/* ---- resident replay: same step, inputs uploaded once, run many times ---- */

typedef struct gfpl_replay gfpl_replay;

int gfpl_replay_create(gfpl_ctx* ctx, int32_t B, int32_t cap_pt, int32_t cap_ln, gfpl_replay** out);
/* same, with separate (tighter) capacities for the previous frames' stereo features: a previous frame holds only the
 * stereo survivors, so its rows can be packed with a smaller stride than the raw detector output (less H2D traffic) */
int gfpl_replay_create_ex(gfpl_ctx* ctx, int32_t B, int32_t cap_pt, int32_t cap_ln, int32_t prev_cap_pt,
                          int32_t prev_cap_ln, gfpl_replay** out);
/* a replay holds a pointer to its context: destroy it BEFORE gfpl_ctx_destroy of that context */
void gfpl_replay_destroy(gfpl_replay* rp);
/* async H2D of the step inputs on the context stream (pinned host memory recommended) */
int gfpl_replay_upload(gfpl_replay* rp, const gfpl_stereo_frames* prev, const gfpl_frame_state* state,
                       const gfpl_raw_frames* curr_raw);
/* same with slim previous-frame records (capacities = the prev capacities of gfpl_replay_create_ex); they are
 * expanded to full feature records on the device */
int gfpl_replay_upload_slim(gfpl_replay* rp, const gfpl_prev_frames* prev, const gfpl_frame_state* state,
                            const gfpl_raw_frames* curr_raw);
/* as gfpl_replay_upload_slim, from the compact records above (expanded on the device) */
int gfpl_replay_upload_compact(gfpl_replay* rp, const gfpl_compact_input* in, const gfpl_frame_state* state /*[B]*/);
/* Sequence-chained replay: B sequences advance one frame per step, as StereoFrameHandler::updateFrame does
 * (stereoFrameHandler.cpp:908-912, `prev_frame = curr_frame`).  After a run,
 *   gfpl_replay_advance   turns the current stereo frames into the previous frames of the next step ON THE DEVICE and the
 *                         pose records into the frame state (Tfw, DT, DT_cov, Tfw_cov of optimizePose's result; time_stamp
 *                         <- time_stamp_curr; time_stamp_curr <- time_stamp_next[p], or one more equal step when NULL);
 *   gfpl_replay_upload_curr_compact   uploads only the NEXT current raw frames (the current-frame fields of
 *                         gfpl_compact_input; its prev_* fields are ignored and may be NULL),
 * so a step costs 35 KB of host-to-device traffic per EuRoC-shaped pair instead of 51 KB.  Needs a replay created with equal
 * capacities for both frame sets (gfpl_replay_create).  The first step of a sequence is a normal upload. */
int gfpl_replay_advance(gfpl_replay* rp, const double* time_stamp_next /*[B] or NULL*/);
int gfpl_replay_upload_curr_compact(gfpl_replay* rp, const gfpl_compact_input* in);
/* launch the whole step on the resident inputs (asynchronous; no host synchronisation inside) */
int gfpl_replay_run(gfpl_replay* rp, int32_t flags);
/* Software pipeline of gfpl_replay_run over n_sub contiguous sub-batches of the B pairs (1 = off, the default; at most
 * GFPL_MAX_SUBBATCHES).  The matching kernels (tensor / integer pipes, two large CTAs per SM) of sub-batch k+1 run on one
 * stream while the fp64 latency-bound kernels (line cut, pose optimisation: < 20 % of an SM's warp slots) of sub-batch k
 * run on high-priority streams of their own, so both kinds of work share the SMs; consecutive runs without an upload in
 * between overlap the same way (the matching of step n+1 starts while the last sub-batches of step n are optimised).
 * Pairs are independent, so the results are bit-identical to the unpipelined step.  Everything is ordered behind the
 * context stream: uploads before a run, downloads / gfpl_ctx_sync / timers after it see the same state as without the
 * pipeline.  gfpl_replay_stage_ms needs n_sub == 1. */
#define GFPL_MAX_SUBBATCHES 16
int gfpl_replay_set_pipeline(gfpl_replay* rp, int32_t n_sub);
/* async D2H of the per-pair pose records (pose: [B]); obs/cut/curr optional (NULL to skip) */
int gfpl_replay_download(gfpl_replay* rp, gfpl_pose_result* pose, gfpl_track_obs* obs, gfpl_line_cut* cut,
                         gfpl_stereo_frames* curr);
int64_t gfpl_replay_h2d_bytes(const gfpl_replay* rp);
int64_t gfpl_replay_d2h_bytes(const gfpl_replay* rp);
/* device pointer of the [B] gfpl_pose_result records (for an NCCL gather without a host bounce) */
void* gfpl_replay_pose_device_ptr(gfpl_replay* rp);

/* per-kernel timing of the last gfpl_replay_run, from CUDA events on the context stream (TimeLog analogue,
 * stereoFrame.h:66-86).  ms[k] for kernel k, ms[GFPL_NUM_STAGES-1] = whole step:
 *   0 stereo point knn (both directions)   1 stereo line knn (both directions)   2 stereo accept loops
 *   3 f2f point search (hybrid fused / knn) 4 f2f line knn (both directions)     5 f2f accept loops
 *   6 stereo uncertainty + line cut         7 pose optimisation                    8 whole step */
#define GFPL_NUM_STAGES 9
int gfpl_replay_stage_ms(gfpl_replay* rp, float* ms /*[GFPL_NUM_STAGES]*/);
/* algorithmic work of the last uploaded batch: DISTINCT descriptor pairs (256-bit distances) needed by kernels 0,1,3,4,
 * i.e. nq*nt per problem — one pass over the distance matrix serves both match directions (the reference's two knnMatch
 * calls evaluate every distance twice); valid after gfpl_replay_run + sync. pairs: [4] */
int gfpl_replay_pair_counts(gfpl_replay* rp, int64_t* pairs /*[4]*/);

/* CUDA-event stopwatch on the context stream: record slot a/b (0..7), then elapsed (synchronises on b) */
int gfpl_timer_record(gfpl_ctx* ctx, int32_t slot);
int gfpl_timer_elapsed_ms(gfpl_ctx* ctx, int32_t slot_a, int32_t slot_b, float* ms);

/* ------------------------------------------------------------------------- */
/* L2 — local bundle adjustment assembly                                      */
/* ------------------------------------------------------------------------- */

/* One LBA window, flat (what MapHandler::localBundleAdjustment gathers, mapHandler.cpp:1108-1209). */
typedef struct gfpl_lba_problem {
  int32_t Nkf, Npt, Nls;       /* local KFs, point landmarks, line landmarks */
  int32_t n_pt_obs, n_ls_obs;
  const double* kf_Tkw;        /* [Nkf][16]  T_kf_w of the local KFs (row-major) */
  const double* X_pt;          /* [Npt][3] */
  const double* X_ls;          /* [Nls][6] */
  /* per observation: landmark local index, local KF index (or -1), fixed KF pose index */
  const int32_t* pt_obs_lm;    /* [n_pt_obs] */
  const int32_t* pt_obs_kf;    /* [n_pt_obs]  kf_idx_loc, -1 if the observing KF is not local */
  const int32_t* pt_obs_fix;   /* [n_pt_obs]  index into fix_Tkw when pt_obs_kf == -1 */
  const double* pt_obs_uv;     /* [n_pt_obs][2] */
  const double* pt_obs_s2;     /* [n_pt_obs] sigma_list */
  const int32_t* ls_obs_lm;
  const int32_t* ls_obs_kf;
  const int32_t* ls_obs_fix;
  const double* ls_obs_le;     /* [n_ls_obs][3] */
  const double* ls_obs_s2;
  const double* fix_Tkw;       /* [n_fix][16] poses of non-local observing KFs */
  int32_t n_fix;
} gfpl_lba_problem;

/* Block-sparse normal equations of one assembly pass of levMarquardtOptimizationLBA
 * (mapHandler.cpp:1243-1426 / 1466-1649).  H = [Hkk ... ; per-observation off-diagonal blocks]. */
typedef struct gfpl_lba_blocks {
  double* Hkk;    /* [Nkf][36]      H.block(idx,idx,6,6) */
  double* Hpp;    /* [Npt][9]       H.block(jdx,jdx,3,3) */
  double* Hll;    /* [Nls][36]      H.block(jdx,jdx,6,6) */
  double* Hpk;    /* [n_pt_obs][18] H.block(jdx,idx,3,6) contribution of that observation (zero if kf == -1) */
  double* Hlk;    /* [n_ls_obs][36] H.block(jdx,idx,6,6) contribution of that observation */
  double* g;      /* [6*Nkf + 3*Npt + 6*Nls] */
  double* err;    /* [1] sum of w*r^2 (not normalised) */
} gfpl_lba_blocks;

/* Residual / Jacobian / block-H assembly of one pass; the damping, sparse solve and update loop stay
 * on the host (mapHandler.cpp:1428-1453,1652-1685). */
int gfpl_lba_assemble(gfpl_ctx* ctx, const gfpl_lba_problem* prob, gfpl_lba_blocks* out);

/* One Levenberg-Marquardt pass of levMarquardtOptimizationLBA on the device: the assembly above followed by the damped
 * solve  DX = (H + lambda_eff diag(H))^-1 g  (mapHandler.cpp:1436-1441 first pass, :1654-1660 later passes; SURVEY.md
 * §8f rank 3).  The reference factorises a dense->sparse copy with Eigen::SimplicialLDLT; here the block-diagonal landmark
 * part is eliminated (Schur complement) and the 6Nkf x 6Nkf remainder is factorised densely — the same solution.
 * scale_by_hmax != 0: lambda_eff = lambda * (int)max|H_ii| (the first pass, :1429-1437, `int Hmax`); else lambda_eff = lambda.
 * Observation lists must be sorted by landmark index (localBundleAdjustment emits them that way, :1146-1209).
 * DX: [6Nkf + 3Npt + 6Nls] in the reference's state order; err_sum: sum of w r^2 of the pass (not normalised).
 * Error accounting, the accept / reject of the step and the state update (:1443-1451, :1662-1685) stay on the host. */
int gfpl_lba_step(gfpl_ctx* ctx, const gfpl_lba_problem* prob, double lambda, int32_t scale_by_hmax, double* DX,
                  double* err_sum, double* lambda_eff);

/* The same pass for W independent windows in one call (a batched replay runs one local BA per inserted key frame:
 * mapHandler.cpp:1108-1838 is entered once per window, and the windows of different sequences share nothing).  Every
 * window's arrays travel in one H2D copy, each kernel is launched once for the whole batch (grid.y = window), and the
 * results come back in three D2H copies.  probs: [W] (host structs, host pointers; windows may differ in every size);
 * DX: the windows' updates concatenated — window w starts at sum over v < w of (6 Nkf_v + 3 Npt_v + 6 Nls_v);
 * err_sum, lambda_eff: [W].  The kernels' device time of the last call lies between timer slots 6 and 7
 * (gfpl_timer_elapsed_ms(ctx, 6, 7, &ms)). */
int gfpl_lba_step_batch(gfpl_ctx* ctx, int32_t W, const gfpl_lba_problem* probs, double lambda, int32_t scale_by_hmax,
                        double* DX, double* err_sum /*[W]*/, double* lambda_eff /*[W]*/);

/* ------------------------------------------------------------------------- */
/* Device micro-benchmarks used as roofline denominators          */
/* ------------------------------------------------------------------------- */

/* measured popc.b32 throughput (T popc/s) and fp64 FMA throughput (TFLOP/s, FMA = 2 flop) */
int gfpl_measure_peaks(gfpl_ctx* ctx, double* popc_tops, double* fp64_tflops);
/* measured dense tcgen05.mma kind::i8 throughput (TOP/s, 2 ops per int8 MAC): one CTA per SM, M = 128, N = 256, operands
 * resident in shared memory — the denominator of the tensor-core matching kernels' roofline */
int gfpl_measure_int8_peak(gfpl_ctx* ctx, double* int8_tops);
/* measured 3-input LOP3 throughput of the ALU pipe (T op/s): the second pipe of the matching loop */
int gfpl_measure_alu_peak(gfpl_ctx* ctx, double* lop3_tops);

#ifdef __cplusplus
}
#endif
#endif /* GFPL_C_H */
