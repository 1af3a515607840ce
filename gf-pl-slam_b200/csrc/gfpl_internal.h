// gfpl_internal.h — host-side internals shared by the .cu translation units of libgfpl.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include <string>
#include <vector>

#include "../../include/gfpl_c.h"

namespace gfpl {

// device-side views: same structs as the C ABI, but every pointer is a DEVICE pointer
typedef gfpl_raw_frames DevRawFrames;
typedef gfpl_stereo_frames DevStereoFrames;
typedef gfpl_track_obs DevTrackObs;

struct Knn2Out {  // knn k=2 results of one direction, [B][cap_q]
  int2* idx;      // (i0, i1), -1 when missing
  float2* dist;   // (d0, d1), +inf when missing
};

// compacted sides of a local-map <-> KF matching problem (k_map.cu); all [B][cap] of the side they index
struct MapWork {
  int32_t *n_sel_pt, *n_un_pt, *n_sel_ln, *n_un_ln;  // [B]
  int32_t *sel_pt, *sel_ln;                          // [B][map.cap_*]  original landmark index
  int32_t *un_pt, *un_ln;                            // [B][kf1.cap_*]  original feature index
  uint8_t *dsel_pt, *dsel_ln, *dun_pt, *dun_ln;      // compacted descriptor rows
};

void set_last_error(gfpl_ctx* ctx, const char* fmt, ...);

}  // namespace gfpl

struct gfpl_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t stream = nullptr;
  gfpl_camera cam{};
  gfpl_config cfg{};
  int64_t launches = 0;
  std::string last_error;
  // pinned host staging of the batched local-BA entry point (grown on demand, freed with the ctx)
  void* staging = nullptr;
  size_t staging_bytes = 0;
  cudaEvent_t timer[8] = {};
  // first failure of an asynchronous copy enqueued by an entry point (checked when the call synchronises)
  cudaError_t pending_error = cudaSuccess;
  // private stream-ordered pool of this context (scratch of the host-pointer entry points)
  cudaMemPool_t pool = nullptr;
};

#define GFPL_CUDA_OK(ctx, expr)                                                              \
  do {                                                                                       \
    cudaError_t _e = (expr);                                                                 \
    if (_e != cudaSuccess) {                                                                 \
      gfpl::set_last_error(ctx, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return GFPL_E_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

#define GFPL_LAUNCH_OK(ctx)                                                                  \
  do {                                                                                       \
    (ctx)->launches++;                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess) {                                                                 \
      gfpl::set_last_error(ctx, "kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return GFPL_E_CUDA;                                                                    \
    }                                                                                        \
  } while (0)

namespace gfpl {

// ---- kernel launchers (all asynchronous on ctx->stream, device pointers only) --------------------

// K1: knn k=2 for B problems, optionally both directions in one launch.
// dir 0: queries = a (na rows), train = b (nb rows) -> out_ab ; dir 1 (if out_ba.idx): queries = b, train = a.
// reverse_best_only: the caller reads only out_ba.idx[].x (mutual checks) — both directions then come from one pass over
// the distance matrix (out_ba rows: idx = (best, -1), dist = (d0, +inf)).
int launch_knn2(gfpl_ctx* ctx, int B, const uint8_t* a, const int32_t* na, int cap_a, const uint8_t* b,
                const int32_t* nb, int cap_b, int norm, Knn2Out out_ab, Knn2Out out_ba, bool reverse_best_only = false);

// tensor-core form of the NORM_HAMMING both-directions call (k_hamming_tc.cu): forward k=2, reverse best
bool knn2_tc_fits(int cap_a, int cap_b);
bool knn2_tc_small_fits(int cap_a, int cap_b);
void tc_column_blocks(int cap_n, int* nblk, int* blk);   // column-blocked tensor-core forms: blocks of the train side
int launch_knn2_tc_small(gfpl_ctx* ctx, int B, const uint8_t* a, const int32_t* na, int cap_a, const uint8_t* b, const int32_t* nb,
                         int cap_b, int norm, Knn2Out out_ab, Knn2Out out_ba);
int launch_knn2_tc(gfpl_ctx* ctx, int B, const uint8_t* a, const int32_t* na, int cap_a, const uint8_t* b, const int32_t* nb,
                   int cap_b, Knn2Out out_ab, Knn2Out out_ba);

// K2: radius match for one problem (CSR), two passes
int launch_radius_count(gfpl_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int norm, float radius,
                        int32_t* counts);
int launch_radius_fill(gfpl_ctx* ctx, const uint8_t* q, int nq, const uint8_t* t, int nt, int norm, float radius,
                       const int32_t* offsets, int32_t* out_idx, float* out_dist);

// K3: stereo accept loops (points + lines) for B frames, one CTA per frame
int launch_stereo_accept(gfpl_ctx* ctx, const DevRawFrames& raw, int variant, Knn2Out pt_lr, Knn2Out pt_rl,
                         Knn2Out ln_lr, Knn2Out ln_rl, DevStereoFrames out, bool lines_only = false);
// point part of extractStereoFeatures_ORBSLAM (row-band search + SAD refinement); dpyr.data is a device pointer
int launch_stereo_orbslam_points(gfpl_ctx* ctx, const DevRawFrames& raw, const gfpl_pyramid& dpyr, DevStereoFrames out);
int launch_stereo_uncertainty(gfpl_ctx* ctx, const DevStereoFrames& fr, gfpl_line_cov* out);
int launch_expand_prev(gfpl_ctx* ctx, const gfpl_prev_frames& slim_dev, DevStereoFrames fr);
// compact step input (device staging copies) -> raw key points / key lines and previous-frame feature records
int launch_expand_compact(gfpl_ctx* ctx, const gfpl_compact_input& c, DevRawFrames raw, DevStereoFrames prev, bool with_prev = true);
// updateFrame: pose records -> previous-frame state (stamps: device [B] or NULL)
int launch_advance_state(gfpl_ctx* ctx, int B, const gfpl_pose_result* pose, const double* stamps, gfpl_frame_state* state);

// f2f: fused radius + projection-gate search (hybrid points), knn-candidate variant, and accept loops
int launch_f2f_hybrid_points(gfpl_ctx* ctx, const DevStereoFrames& prev, const DevStereoFrames& curr,
                             const gfpl_frame_state* state, double* T32 /*[B][32] scratch*/,
                             uint32_t* best_key /*[B][curr.cap_pt]*/);
int launch_f2f_hybrid_points_knn(gfpl_ctx* ctx, const DevStereoFrames& prev, const DevStereoFrames& curr,
                                 const gfpl_frame_state* state, Knn2Out pt_12, uint32_t* best_key);
int launch_f2f_accept(gfpl_ctx* ctx, const DevStereoFrames& prev, const DevStereoFrames& curr, int mode,
                      const uint32_t* best_key, Knn2Out pt_12, Knn2Out pt_21, Knn2Out ln_12, Knn2Out ln_21,
                      DevTrackObs out);

int launch_kf_accept(gfpl_ctx* ctx, const DevStereoFrames& kf0, const DevStereoFrames& kf1, const double* DT, Knn2Out pt_12,
                     Knn2Out pt_21, Knn2Out ln_12, Knn2Out ln_21, DevTrackObs out);

int launch_map_select(gfpl_ctx* ctx, const gfpl_local_map& map, const DevStereoFrames& kf1, const double* Twf, MapWork w);
int launch_map_accept(gfpl_ctx* ctx, const gfpl_local_map& map, const DevStereoFrames& kf1, const double* Twf, MapWork w,
                      Knn2Out pt_12, Knn2Out pt_21, Knn2Out ln_12, Knn2Out ln_21, gfpl_map_matches out);

int launch_medoid(gfpl_ctx* ctx, const uint8_t* desc, const int32_t* offsets, int L, int32_t* out);

// K5: line cut.  DT_inv may be NULL: then it is formed from `state` (curr->Tfw^-1 * prev->Tfw with the
// predicted curr->Tfw = prev->Tfw * prev->DT).
int launch_linecut(gfpl_ctx* ctx, const double* DT_inv, const gfpl_frame_state* state, DevTrackObs obs,
                   gfpl_line_cov* prev_cov, int cov_cap, gfpl_line_cut* cut, const gfpl_line_feature* prev_ln = nullptr,
                   int prev_cap_ln = 0);

int launch_linecut_descent(gfpl_ctx* ctx, const double* DT_inv, DevTrackObs obs, const gfpl_line_cov* prev_cov, int cov_cap,
                           gfpl_line_cut_descent* out);

// K4: pose
int launch_pose_hg(gfpl_ctx* ctx, const double* DT, const DevTrackObs& obs, double* H, double* g, double* e);
int launch_pose_optimize(gfpl_ctx* ctx, DevTrackObs obs, const gfpl_frame_state* state, const double* DT_ini,
                         int overload, gfpl_pose_result* out, void* raw_scratch = nullptr);
int launch_stdv_mad(gfpl_ctx* ctx, const double* res, const int32_t* n, int B, int cap, double* out);
size_t pose_raw_bytes(int B);   // size of the optimiser -> finalize hand-off buffer for B pairs

// K6 / §8f-3: local BA over a BATCH of independent windows.  One LbaWin per window, all pointers DEVICE pointers (into
// one buffer that holds every window's arrays); the kernels take the array and select their window by blockIdx.y.
struct LbaSolveWork {
  double* S;          // [6Nkf][6Nkf]
  double *Ypt, *Zpt;  // [n_pt_obs][18], [Npt][3]
  double *Yls, *Zls;  // [n_ls_obs][36], [Nls][6]
};
struct LbaWin {
  gfpl_lba_problem pb;
  gfpl_lba_blocks H;
  const int32_t *pt_off, *ls_off;   // CSR offsets of the landmark-sorted observation lists
  LbaSolveWork w;
  double* DX;                       // [6Nkf + 3Npt + 6Nls]
  double* lambda_eff;               // [1] lambda * (int)max|H_ii| if scale_by_hmax else lambda
};
// residual / Jacobian / block-H assembly of every window (block storage zeroed by the caller)
int launch_lba_assemble(gfpl_ctx* ctx, const LbaWin* wins, int W, int max_kf, int max_pt_obs, int max_ls_obs);
// DX = (H + lambda_eff diag H)^-1 g per window: landmark elimination (Schur complement), dense LDL^T, back substitution
int launch_lba_solve(gfpl_ctx* ctx, const LbaWin* wins, int W, int max_kf, int max_pt, int max_ls, double lambda, int scale_by_hmax);

int launch_measure_peaks(gfpl_ctx* ctx, double* popc_tops, double* fp64_tflops);
int launch_measure_alu_peak(gfpl_ctx* ctx, double* lop3_tops);
int launch_measure_int8_peak(gfpl_ctx* ctx, double* int8_tops);   // dense tcgen05.mma kind::i8, operands resident in smem

}  // namespace gfpl
