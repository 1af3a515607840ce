/* TEST INFRASTRUCTURE — CPU oracle of the GF-PL-SLAM tracking hot path (plain C interface).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may use it.
 *
 * PARITY PIN: matching (M1-M3) is pinned against cv2.BFMatcher 4.13 (tests/golden/, made by
 * oracle/make_golden.py).  Everything else is "parity unpinned by the reference": the reference ships
 * no golden vectors for these functions and cannot be compiled here (Eigen3,
 * OpenCV C++, g2o, yaml-cpp, Boost absent), so the fp64 stages are a line-by-line restatement only.
 *
 * Same structs as include/gfpl_c.h; every orc_* function mirrors the gfpl_* function of the same
 * name, without a context (camera + config are passed explicitly).
 */
#ifndef GFPL_ORACLE_H
#define GFPL_ORACLE_H
#include "../include/gfpl_c.h"
#ifdef __cplusplus
extern "C" {
#endif

void orc_config_default(gfpl_config* cfg);

int orc_knn2_hamming256(const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t norm,
                        int32_t* out_idx, float* out_dist);
int orc_radius_hamming256(const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t norm, float radius,
                          int32_t* offsets, int32_t* out_idx, float* out_dist, int32_t cap);
/* lineDescriptorMAD's nn12_mad and the two budget thresholds, on knn k=2 results (d: [n][2]) */
double orc_line_descriptor_mad(const float* d, int32_t n, double mad_uninit_value);
double orc_budget_thres(const float* d, int32_t n, int32_t max_match_num);
double orc_vector_stdv_mad(const double* residues, int32_t n);

int orc_stereo_match(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_raw_frames* raw, int32_t variant,
                     gfpl_stereo_frames* out);
int orc_stereo_match_orbslam(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_raw_frames* raw,
                             const gfpl_pyramid* right_pyramid, gfpl_stereo_frames* out);
int orc_stereo_uncertainty(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_stereo_frames* frames,
                           gfpl_line_cov* out);
int orc_f2f_match(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_stereo_frames* prev,
                  const gfpl_stereo_frames* curr, const gfpl_frame_state* state, int32_t mode, gfpl_track_obs* out);
int orc_kf_match(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_stereo_frames* kf0,
                 const gfpl_stereo_frames* kf1, const double* DT, gfpl_track_obs* out);
int orc_map_match(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_local_map* map,
                  const gfpl_stereo_frames* kf1, const double* Twf, gfpl_map_matches* out);
int orc_medoid_desc(const uint8_t* desc, const int32_t* offsets, int32_t L, int32_t* out_idx);
int orc_linecut_descent(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv, gfpl_track_obs* obs,
                        const gfpl_line_cov* prev_cov, int32_t cov_cap, gfpl_line_cut_descent* cut);
void orc_pose_cov_on_line(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv, const gfpl_ln_obs* ln,
                          const gfpl_line_cov* cov, const double cut[2], double* cov36);
double orc_descent_volume(const double* cov36);
int orc_linecut_submodular(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv, gfpl_track_obs* obs,
                           const gfpl_line_cov* prev_cov, int32_t cov_cap, gfpl_line_cut* cut);
/* single evaluation of getPoseInfoOnLine / getPoseInfoPoint / logdet / min-eigenvalue (unit tests) */
void orc_pose_info_on_line(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv, const gfpl_ln_obs* ln,
                           const gfpl_line_cov* cov, const double cut[2], double* info36);
void orc_pose_info_point(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT_inv, const gfpl_pt_obs* pt,
                         double* info36);
double orc_logdet6(const double* M36);
double orc_min_eig6(const double* M36);
void orc_sym_eig6(const double* M36, double* ev6);
void orc_sym_eig3(const double* M9, double* ev3);

int orc_pose_hg(const gfpl_camera* cam, const gfpl_config* cfg, const double* DT, const gfpl_track_obs* obs,
                double* H, double* g, double* e);
int orc_pose_optimize(const gfpl_camera* cam, const gfpl_config* cfg, gfpl_track_obs* obs,
                      const gfpl_frame_state* state, const double* DT_ini, int32_t overload, gfpl_pose_result* out);
/* n_threads > 1 splits the B pairs over std::threads (CPU-baseline upper bound) */
int orc_track(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_stereo_frames* prev,
              const gfpl_frame_state* state, const gfpl_raw_frames* curr_raw, int32_t flags,
              gfpl_stereo_frames* curr, gfpl_track_obs* obs, gfpl_line_cut* cut, gfpl_pose_result* pose,
              int32_t n_threads);

int orc_lba_assemble(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_lba_problem* prob,
                     gfpl_lba_blocks* out);

int orc_lba_step(const gfpl_camera* cam, const gfpl_config* cfg, const gfpl_lba_problem* prob, double lambda,
                 int32_t scale_by_hmax, double* DX, double* err_sum, double* lambda_eff);

/* SE(3) helpers exposed for unit tests (auxiliar.cpp:154-238) */
void orc_expmap_se3(const double* x6, double* T16);
void orc_logmap_se3(const double* T16, double* x6);
void orc_inverse_se3(const double* T16, double* Tinv16);

#ifdef __cplusplus
}
#endif
#endif
