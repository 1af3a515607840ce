// k_stereo_orb.cu — point part of StereoFrame::extractStereoFeatures_ORBSLAM (stereoFrame.cpp:453-630) with
// subPixelStereoRefine_ORBSLAM (:340-404): row-band descriptor search, 11x11 SAD refinement on the right pyramid,
// median-distance cut, output sorted by (distance, left index).  One CTA per frame; float arithmetic replays the
// reference's float expressions operation by operation (this file is built with -fmad=false).
#include <climits>

#include "gfpl_hamming.cuh"
#include "gfpl_internal.h"
#include "gfpl_math.cuh"

namespace gfpl {

constexpr int ORB_THREADS = 256;
constexpr uint32_t ORB_NONE = 0xffffffffu;

struct ScaleTables {
  float sf[GFPL_MAX_PYR_LEVELS], isf[GFPL_MAX_PYR_LEVELS];
};

// ORBextractor.cc:415-431 (float arithmetic)
static ScaleTables make_scale_tables(const gfpl_config& cfg, int n) {
  ScaleTables t{};
  const float scaleFactor = (float)cfg.orb_scale_factor;
  t.sf[0] = 1.0f;
  for (int i = 1; i < n; i++) t.sf[i] = t.sf[i - 1] * scaleFactor;
  for (int i = 0; i < n; i++) t.isf[i] = 1.0f / t.sf[i];
  return t;
}

__device__ __forceinline__ double orb_level_sigma2(int level, double scale) {  // stereoFeatures.cpp:41-48
  double sigma2 = 1.0;
  for (int i = 0; i < level + 1; i++) sigma2 *= scale;
  return 1.0 / (sigma2 * sigma2);
}

// subPixelStereoRefine_ORBSLAM (:340-404).  Returns disparity (-1: none) and bestuR.
__device__ void subpixel_refine(const gfpl_pyramid& pyr, const uint8_t* __restrict__ frame, const ScaleTables& st,
                                const gfpl_keypoint& kpL, const gfpl_keypoint& kpR, float& disparity, float& bestuR) {
  disparity = -1;
  bestuR = kpR.x;
  if (kpL.octave < 0 || kpL.octave >= pyr.n_levels) return;
  const float uR0 = kpR.x;
  const float scaleFactor = st.isf[kpL.octave];
  const float scaleduL = roundf(kpL.x * scaleFactor);
  const float scaledvL = roundf(kpL.y * scaleFactor);
  const float scaleduR0 = roundf(uR0 * scaleFactor);
  const int w = 5, L = 5;
  const int W = pyr.width[kpL.octave], H = pyr.height[kpL.octave], step = pyr.row_stride[kpL.octave];
  const uint8_t* img = frame + pyr.offset[kpL.octave];  // both windows from the RIGHT pyramid (:358, :376)
  const int r0 = (int)(scaledvL - w), r1 = (int)(scaledvL + w + 1);
  const int cl0 = (int)(scaleduL - w), cl1 = (int)(scaleduL + w + 1);
  if (r0 < 0 || r1 > H || cl0 < 0 || cl1 > W) return;  // defined (cv::Mat range assert)
  const float iniu = scaleduR0 + L - w;
  const float endu = scaleduR0 + L + w + 1;
  if (iniu < 0 || endu >= W) return;  // :370-371
  const int cbase = (int)(scaleduR0 - L - w);
  if (cbase < 0) return;              // defined (first window's colRange)
  // integer SAD: every term of the reference's float expression is an exactly representable integer
  int sad[11];
#pragma unroll
  for (int k = 0; k < 11; k++) sad[k] = 0;
  const int cL = img[(size_t)(r0 + w) * step + cl0 + w];
  int cR[11];
#pragma unroll
  for (int k = 0; k < 11; k++) cR[k] = img[(size_t)(r0 + w) * step + cbase + k + w];
  for (int i = 0; i < 11; i++) {
    const uint8_t* rowL = img + (size_t)(r0 + i) * step + cl0;
    const uint8_t* rowR = img + (size_t)(r0 + i) * step + cbase;
    int rr[21];
#pragma unroll
    for (int j = 0; j < 21; j++) rr[j] = rowR[j];
#pragma unroll
    for (int j = 0; j < 11; j++) {
      const int il = (int)rowL[j] - cL;
#pragma unroll
      for (int k = 0; k < 11; k++) sad[k] += abs(il - (rr[j + k] - cR[k]));
    }
  }
  int bestDist = INT_MAX, bestincR = 0;
#pragma unroll
  for (int k = 0; k < 11; k++) {
    const float dist = (float)sad[k];
    if (dist < (float)bestDist) { bestDist = (int)dist; bestincR = k - L; }
  }
  if (bestincR == -L || bestincR == L) return;
  float dist1 = 0.f, dist2 = 0.f, dist3 = 0.f;
#pragma unroll
  for (int k = 1; k < 10; k++)
    if (k == L + bestincR) { dist1 = (float)sad[k - 1]; dist2 = (float)sad[k]; dist3 = (float)sad[k + 1]; }
  const float deltaR = (dist1 - dist3) / (2.0f * (dist1 + dist3 - 2.0f * dist2));
  if (deltaR < -1 || deltaR > 1) return;
  bestuR = st.sf[kpL.octave] * ((float)scaleduR0 + (float)bestincR + deltaR);
  disparity = (kpL.x - bestuR);
}

__global__ void __launch_bounds__(ORB_THREADS)
stereo_orbslam_points_kernel(DevRawFrames raw, gfpl_pyramid pyr, ScaleTables st, DevStereoFrames out, gfpl_camera cam,
                             gfpl_config cfg) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  uint32_t* s_key = reinterpret_cast<uint32_t*>(s_dyn);        // [cap_pt] (bestDist << 20) | iL, ORB_NONE if rejected
  float* s_depth = reinterpret_cast<float*>(s_key + raw.cap_pt);  // [cap_pt] mvDepth
  int* s_srcr = reinterpret_cast<int*>(s_depth + raw.cap_pt);  // [cap_pt] bestIdxR
  __shared__ int s_count[2];
  __shared__ int s_med;
  const int p = blockIdx.x;
  const size_t rp = (size_t)p * raw.cap_pt, op = (size_t)p * out.cap_pt;
  const int N = raw.n_pt_l[p], Nr = raw.n_pt_r[p];
  if (threadIdx.x < 2) s_count[threadIdx.x] = 0;
  if (threadIdx.x == 0) s_med = 0;
  __syncthreads();
  if (!cfg.has_points || N == 0 || Nr == 0) {  // :453
    if (threadIdx.x == 0) out.n_pt[p] = 0;
    return;
  }
  const uint8_t* frame = pyr.data + (size_t)p * pyr.frame_stride;
  const int nRows = pyr.height[0];
  const float minD = 0;
  const float maxD = (float)cam.fx;
  const float mbf = (float)(cam.fx * cam.b);
  const gfpl_keypoint* kl = raw.pt_l + rp;
  const gfpl_keypoint* kr = raw.pt_r + rp;
  int n_mine = 0;
  for (int iL = threadIdx.x; iL < N; iL += ORB_THREADS) {
    uint32_t key = ORB_NONE;
    const gfpl_keypoint kpL = kl[iL];
    const int levelL = kpL.octave;
    const float vL = kpL.y, uL = kpL.x;
    const float minU = uL - maxD;
    const float maxU = uL - minD;
    if (vL >= 0 && (size_t)vL < (size_t)nRows && !(maxU < 0)) {
      const int row = (int)vL;  // vRowIndices[vL]: float -> size_t truncation
      const Desc256 dL = load_desc_global(raw.pdesc_l + (rp + iL) * 32);
      uint32_t best = make_key(100u, 0u);  // bestDist = 100, strict <, candidates in ascending iR (:523-549)
      for (int iR = 0; iR < Nr; iR++) {
        const gfpl_keypoint kpR = kr[iR];
        if (kpR.octave < 0 || kpR.octave >= pyr.n_levels) continue;   // not a level of this pyramid: no scale table entry
        const float r = 2.0f * st.sf[kpR.octave];
        const int maxr = (int)ceilf(kpR.y + r);
        const int minr = (int)floorf(kpR.y - r);
        if (row < minr || row > maxr) continue;                                   // the row table (:476-486)
        if (kpR.octave < levelL - 1 || kpR.octave > levelL + 1) continue;
        if (kpR.x >= minU && kpR.x <= maxU) {
          const uint4* t = reinterpret_cast<const uint4*>(raw.pdesc_r + (rp + iR) * 32);
          const uint32_t dist = hamming256<GFPL_NORM_HAMMING>(dL, __ldg(t), __ldg(t + 1));
          if (dist < (best >> KEY_IDX_BITS)) best = make_key(dist, (uint32_t)iR);
        }
      }
      const int bestDist = (int)(best >> KEY_IDX_BITS);
      if (bestDist < 80) {  // thOrbDist
        const int bestIdxR = (int)(best & KEY_IDX_MASK);
        float disparity, bestuR;
        subpixel_refine(pyr, frame, st, kpL, kr[bestIdxR], disparity, bestuR);
        if (disparity >= minD && disparity < maxD) {
          if (disparity <= 0) disparity = (float)0.01;
          s_depth[iL] = mbf / disparity;
          s_srcr[iL] = bestIdxR;
          key = make_key((uint32_t)bestDist, (uint32_t)iL);
          n_mine++;
        }
      }
    }
    s_key[iL] = key;
  }
  if (n_mine) atomicAdd(&s_count[0], n_mine);
  __syncthreads();
  const int n_acc = s_count[0];
  if (n_acc == 0) {  // defined: empty vDistIdx (:587)
    if (threadIdx.x == 0) out.n_pt[p] = 0;
    return;
  }
  // rank = position in sort(vDistIdx) (:586); the keys are distinct
  int my_rank[8];
  int n_it = 0;
  for (int iL = threadIdx.x; iL < N; iL += ORB_THREADS, n_it++) {
    const uint32_t key = s_key[iL];
    int r = -1;
    if (key != ORB_NONE) {
      r = 0;
      for (int k = 0; k < N; k++) r += s_key[k] < key;
      if (r == n_acc / 2) s_med = (int)(key >> KEY_IDX_BITS);
    }
    if (n_it < 8) my_rank[n_it] = r;
  }
  __syncthreads();
  const float median = (float)s_med;
  const float thDist = 1.5f * 1.4f * median;
  n_it = 0;
  int kept = 0;
  for (int iL = threadIdx.x; iL < N; iL += ORB_THREADS, n_it++) {
    const uint32_t key = s_key[iL];
    if (key == ORB_NONE) continue;
    const int dist = (int)(key >> KEY_IDX_BITS);
    if ((float)dist >= thDist) continue;
    int r;
    if (n_it < 8) r = my_rank[n_it];
    else { r = 0; for (int k = 0; k < N; k++) r += s_key[k] < key; }
    const float disparity = mbf / s_depth[iL];
    // (`disparity < 0` of :606 cannot happen: disparity >= 0.01 was stored)
    const gfpl_keypoint kpL = kl[iL];
    gfpl_point_feature f;
    f.pl[0] = kpL.x; f.pl[1] = kpL.y;
    back_projection(cam, f.pl[0], f.pl[1], (double)disparity, f.P);
    f.disp = disparity;
    f.level = kpL.octave;
    f.sigma2 = orb_level_sigma2(f.level, cfg.orb_scale_factor);
    f.idx = r; f.src_l = iL; f.src_r = s_srcr[iL];
    out.pt[op + r] = f;
    const uint4* s = reinterpret_cast<const uint4*>(raw.pdesc_l + (rp + iL) * 32);
    uint4* d = reinterpret_cast<uint4*>(out.pdesc + (op + r) * 32);
    d[0] = s[0]; d[1] = s[1];
    kept++;
  }
  if (kept) atomicAdd(&s_count[1], kept);
  __syncthreads();
  if (threadIdx.x == 0) out.n_pt[p] = s_count[1];
}

int launch_stereo_orbslam_points(gfpl_ctx* ctx, const DevRawFrames& raw, const gfpl_pyramid& dpyr, DevStereoFrames out) {
  if (raw.B <= 0) return GFPL_OK;
  const size_t smem = (size_t)(raw.cap_pt > 0 ? raw.cap_pt : 1) * 12;
  if (smem > 200 * 1024) {
    set_last_error(ctx, "cap_pt %d too large for the ORB-SLAM stereo kernel", raw.cap_pt);
    return GFPL_E_INVALID;
  }
  if (smem > 48 * 1024)
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(stereo_orbslam_points_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const ScaleTables st = make_scale_tables(ctx->cfg, dpyr.n_levels);
  stereo_orbslam_points_kernel<<<raw.B, ORB_THREADS, smem, ctx->stream>>>(raw, dpyr, st, out, ctx->cam, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

}  // namespace gfpl
