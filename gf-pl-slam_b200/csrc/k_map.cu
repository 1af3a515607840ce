// k_map.cu — local map <-> key-frame half of MapHandler::lookForCommonMatches (mapHandler.cpp:509-770):
// (1) select_kernel: in-image test of the local landmarks on Twf*X and `idx == -1` test of the key frame's features,
//     both compacted (index list + 32-byte descriptor rows) so that the knn k=2 kernel of k_hamming.cu runs on them;
// (2) accept_kernel: mutual check, ratio / MAD-gap test and the reprojection gates, pairs in the reference's loop order.
// blockIdx.x = problem, blockIdx.y = 0 points / 1 lines.
#include "gfpl_internal.h"
#include "gfpl_hamming.cuh"
#include "gfpl_math.cuh"

namespace gfpl {

constexpr int MAP_THREADS = 256;

__device__ __forceinline__ int map_compact(bool pred, int& base, int* s_warp) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned bal = __ballot_sync(0xffffffffu, pred);
  if (lane == 0) s_warp[warp] = __popc(bal);
  __syncthreads();
  int before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < MAP_THREADS / 32; w++) {
    const int c = s_warp[w];
    before += (w < warp) ? c : 0;
    total += c;
  }
  const int slot = base + before + __popc(bal & ((1u << lane) - 1));
  base += total;
  __syncthreads();
  return slot;
}

// rank-th smallest of v[0..n) (ties by position), every thread returns it
__device__ float map_select_rank(const float* v, int n, int rank, float* s_out) {
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float x = v[j];
    int r = 0;
    for (int k = 0; k < n; k++) {
      const float y = v[k];
      r += (y < x) || (y == x && k < j);
    }
    if (r == rank) *s_out = x;
  }
  __syncthreads();
  const float out = *s_out;
  __syncthreads();
  return out;
}

__device__ __forceinline__ void copy_desc(uint8_t* dst, const uint8_t* src) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
  d[0] = s[0]; d[1] = s[1];
}

__device__ __forceinline__ bool in_image(const gfpl_camera& cam, const double* uv, double z) {  // :526, :649-651
  return uv[0] > 0 && uv[0] < cam.width && uv[1] > 0 && uv[1] < cam.height && z > 0.0;
}

__global__ void __launch_bounds__(MAP_THREADS)
map_select_kernel(gfpl_local_map map, DevStereoFrames kf1, const double* __restrict__ Twf_all, MapWork w, gfpl_camera cam) {
  __shared__ int s_warp[MAP_THREADS / 32 + 1];
  __shared__ double s_T[16];
  const int p = blockIdx.x;
  if (threadIdx.x < 16) s_T[threadIdx.x] = Twf_all[(size_t)p * 16 + threadIdx.x];
  __syncthreads();
  if (blockIdx.y == 0) {
    const size_t mp = (size_t)p * map.cap_pt, kp = (size_t)p * kf1.cap_pt;
    const int n = map.n_pt[p];
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += MAP_THREADS) {  // map_local_points (:513-533)
      const int i = i0 + threadIdx.x;
      bool keep = false;
      if (i < n) {
        double Pf[3], pf[2];
        se3_apply(s_T, map.pt3d + (mp + i) * 3, Pf);
        projection(cam, Pf, pf);
        keep = in_image(cam, pf, Pf[2]);
      }
      const int slot = map_compact(keep, base, s_warp);
      if (keep) {
        w.sel_pt[mp + slot] = i;
        copy_desc(w.dsel_pt + (mp + slot) * 32, map.pdesc + (mp + i) * 32);
      }
    }
    if (threadIdx.x == 0) w.n_sel_pt[p] = base;
    const int n1 = kf1.n_pt[p];
    base = 0;
    for (int i0 = 0; i0 < n1; i0 += MAP_THREADS) {  // unmatched_points (:536-547)
      const int i = i0 + threadIdx.x;
      const bool keep = i < n1 && kf1.pt[kp + i].idx == -1;
      const int slot = map_compact(keep, base, s_warp);
      if (keep) {
        w.un_pt[kp + slot] = i;
        copy_desc(w.dun_pt + (kp + slot) * 32, kf1.pdesc + (kp + i) * 32);
      }
    }
    if (threadIdx.x == 0) w.n_un_pt[p] = base;
  } else {
    const size_t ml = (size_t)p * map.cap_ln, kl = (size_t)p * kf1.cap_ln;
    const int n = map.n_ln[p];
    int base = 0;
    for (int i0 = 0; i0 < n; i0 += MAP_THREADS) {  // map_local_lines (:634-660)
      const int i = i0 + threadIdx.x;
      bool keep = false;
      if (i < n) {
        double sPf[3], ePf[3], spf[2], epf[2];
        se3_apply(s_T, map.ln3d + (ml + i) * 6, sPf);
        projection(cam, sPf, spf);
        se3_apply(s_T, map.ln3d + (ml + i) * 6 + 3, ePf);
        projection(cam, ePf, epf);
        keep = in_image(cam, spf, sPf[2]) && in_image(cam, epf, ePf[2]);
      }
      const int slot = map_compact(keep, base, s_warp);
      if (keep) {
        w.sel_ln[ml + slot] = i;
        copy_desc(w.dsel_ln + (ml + slot) * 32, map.ldesc + (ml + i) * 32);
      }
    }
    if (threadIdx.x == 0) w.n_sel_ln[p] = base;
    const int n1 = kf1.n_ln[p];
    base = 0;
    for (int i0 = 0; i0 < n1; i0 += MAP_THREADS) {  // unmatched_lines (:663-674)
      const int i = i0 + threadIdx.x;
      const bool keep = i < n1 && kf1.ln[kl + i].idx == -1;
      const int slot = map_compact(keep, base, s_warp);
      if (keep) {
        w.un_ln[kl + slot] = i;
        copy_desc(w.dun_ln + (kl + slot) * 32, kf1.ldesc + (kl + i) * 32);
      }
    }
    if (threadIdx.x == 0) w.n_un_ln[p] = base;
  }
}

__global__ void __launch_bounds__(MAP_THREADS)
map_accept_kernel(gfpl_local_map map, DevStereoFrames kf1, const double* __restrict__ Twf_all, MapWork w, Knn2Out pt_12,
                  Knn2Out pt_21, Knn2Out ln_12, Knn2Out ln_21, gfpl_map_matches out, gfpl_camera cam, gfpl_config cfg) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  float* s_v = reinterpret_cast<float*>(s_dyn);  // [map.cap_ln]
  __shared__ int s_warp[MAP_THREADS / 32 + 1];
  __shared__ float s_sel;
  __shared__ double s_T[16];
  const int p = blockIdx.x;
  if (threadIdx.x < 16) s_T[threadIdx.x] = Twf_all[(size_t)p * 16 + threadIdx.x];
  __syncthreads();
  if (blockIdx.y == 0) {
    const size_t mp = (size_t)p * map.cap_pt, kp = (size_t)p * kf1.cap_pt, op = (size_t)p * out.cap_pt;
    const int n0 = w.n_sel_pt[p], n1 = w.n_un_pt[p];
    int base = 0;
    if (cfg.has_points && n0 >= 2 && n1 >= 2) {  // D3
      for (int i0 = 0; i0 < n0; i0 += MAP_THREADS) {  // :579-628
        const int i = i0 + threadIdx.x;
        bool keep = false;
        gfpl_map_pair pr{0, 0};
        if (i < n0) {
          const int tdx = pt_12.idx[mp + i].x;
          const float2 md = pt_12.dist[mp + i];
          const int rl = cfg.best_lr_matches ? pt_21.idx[kp + tdx].x : i;
          const double dist_12 = (double)(md.x / md.y);
          if (i == rl && dist_12 <= cfg.max_ratio_12_p) {
            pr.map_idx = w.sel_pt[mp + i];
            pr.kf_idx = w.un_pt[kp + tdx];
            double Pf[3], pf[2];
            se3_apply(s_T, map.pt3d + (mp + pr.map_idx) * 3, Pf);
            if (Pf[2] > 0.0) {
              projection(cam, Pf, pf);
              const gfpl_point_feature& f = kf1.pt[kp + pr.kf_idx];
              const double dx = pf[0] - f.pl[0], dy = pf[1] - f.pl[1];
              keep = sqrt(dx * dx + dy * dy) < cfg.max_kf_epip_p;
            }
          }
        }
        const int slot = map_compact(keep, base, s_warp);
        if (keep && slot < out.cap_pt) out.pt[op + slot] = pr;
      }
      base = min(base, out.cap_pt);
    }
    if (threadIdx.x == 0) out.n_pt[p] = base;
  } else {
    const size_t ml = (size_t)p * map.cap_ln, kl = (size_t)p * kf1.cap_ln, ol = (size_t)p * out.cap_ln;
    const int n0 = w.n_sel_ln[p], n1 = w.n_un_ln[p];
    int base = 0;
    if (cfg.has_lines && n0 >= 2 && n1 >= 2) {  // D3
      for (int j = threadIdx.x; j < n0; j += MAP_THREADS) {  // lineDescriptorMAD (:702)
        const float2 md = ln_12.dist[ml + j];
        s_v[j] = fabsf((float)((double)(md.y - md.x) - cfg.mad_uninit_value));
      }
      __syncthreads();
      const float med = map_select_rank(s_v, n0, n0 / 2, &s_sel);
      const double nn12_dist_th = (1.4826 * (double)med) * cfg.desc_th_l;
      for (int i0 = 0; i0 < n0; i0 += MAP_THREADS) {  // :709-766
        const int i = i0 + threadIdx.x;
        bool keep = false;
        gfpl_map_pair pr{0, 0};
        if (i < n0) {
          const int tdx = ln_12.idx[ml + i].x;
          const float2 md = ln_12.dist[ml + i];
          const int rl = cfg.best_lr_matches ? ln_21.idx[kl + tdx].x : i;
          const double dist_12 = (double)(md.y - md.x);
          if (i == rl && dist_12 > nn12_dist_th) {
            pr.map_idx = w.sel_ln[ml + i];
            pr.kf_idx = w.un_ln[kl + tdx];
            double sP_[3], eP_[3], sp[2], ep[2];
            se3_apply(s_T, map.ln3d + (ml + pr.map_idx) * 6, sP_);
            projection(cam, sP_, sp);
            se3_apply(s_T, map.ln3d + (ml + pr.map_idx) * 6 + 3, eP_);
            projection(cam, eP_, ep);
            if (sP_[2] > 0.0 && eP_[2] > 0.0) {
              const gfpl_line_feature& f = kf1.ln[kl + pr.kf_idx];
              const double e0 = f.le[0] * sp[0] + f.le[1] * sp[1] + f.le[2];  // signed, as the reference (:741-743)
              const double e1 = f.le[0] * ep[0] + f.le[1] * ep[1] + f.le[2];
              keep = e0 < cfg.max_kf_epip_l && e1 < cfg.max_kf_epip_l;
            }
          }
        }
        const int slot = map_compact(keep, base, s_warp);
        if (keep && slot < out.cap_ln) out.ln[ol + slot] = pr;
      }
      base = min(base, out.cap_ln);
    }
    if (threadIdx.x == 0) out.n_ln[p] = base;
  }
}

// ---- MapPoint / MapLine::updateAverageDescDir, descriptor part (mapFeatures.cpp:51-82, 128-160) ------------------------
// One warp per landmark.  Row i's statistic is element int(1 + 0.5 (n-1)) of its ascending distance row (self distance 0
// included); it is found by bisection on the value (distances are integers in [0,256]) with the row recomputed from
// L1-resident descriptors, so no n x n matrix is stored.  Winner = strictly smallest statistic, ties -> lowest i.
constexpr int MEDOID_WARPS = 4;

__global__ void __launch_bounds__(MEDOID_WARPS * 32)
medoid_kernel(const uint8_t* __restrict__ desc, const int32_t* __restrict__ offsets, int L, int32_t* __restrict__ out) {
  const int lm = blockIdx.x * MEDOID_WARPS + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (lm >= L) return;
  const int o = offsets[lm], n = offsets[lm + 1] - o;
  const uint8_t* D = desc + (size_t)o * 32;
  const int r = (int)(1 + 0.5 * (n - 1));  // :76
  uint32_t best = 0xffffffffu;
  if (n >= 2) {
    for (int i = lane; i < n; i += 32) {
      const Desc256 a = load_desc_global(D + (size_t)i * 32);
      int lo = 0, hi = 256;  // smallest v with #{j : d_ij <= v} >= r + 1
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        int c = 0;
        for (int j = 0; j < n; j++) {
          const uint4* row = reinterpret_cast<const uint4*>(D + (size_t)j * 32);
          c += hamming256<GFPL_NORM_HAMMING>(a, __ldg(row), __ldg(row + 1)) <= (uint32_t)mid;
        }
        if (c >= r + 1) hi = mid; else lo = mid + 1;
      }
      const uint32_t key = ((uint32_t)lo << 20) | (uint32_t)i;
      best = min(best, key);
    }
#pragma unroll
    for (int s = 16; s > 0; s >>= 1) best = min(best, __shfl_xor_sync(0xffffffffu, best, s));
  }
  if (lane == 0) out[lm] = n >= 2 ? (int)(best & 0xfffffu) : 0;
}

int launch_medoid(gfpl_ctx* ctx, const uint8_t* desc, const int32_t* offsets, int L, int32_t* out) {
  if (L <= 0) return GFPL_OK;
  medoid_kernel<<<(L + MEDOID_WARPS - 1) / MEDOID_WARPS, MEDOID_WARPS * 32, 0, ctx->stream>>>(desc, offsets, L, out);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

int launch_map_select(gfpl_ctx* ctx, const gfpl_local_map& map, const DevStereoFrames& kf1, const double* Twf, MapWork w) {
  if (map.B <= 0) return GFPL_OK;
  map_select_kernel<<<dim3(map.B, 2), MAP_THREADS, 0, ctx->stream>>>(map, kf1, Twf, w, ctx->cam);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

int launch_map_accept(gfpl_ctx* ctx, const gfpl_local_map& map, const DevStereoFrames& kf1, const double* Twf, MapWork w,
                      Knn2Out pt_12, Knn2Out pt_21, Knn2Out ln_12, Knn2Out ln_21, gfpl_map_matches out) {
  if (map.B <= 0) return GFPL_OK;
  const size_t smem = (size_t)(map.cap_ln > 0 ? map.cap_ln : 1) * 4;
  if (smem > 48 * 1024)
    GFPL_CUDA_OK(ctx, cudaFuncSetAttribute(map_accept_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  map_accept_kernel<<<dim3(map.B, 2), MAP_THREADS, smem, ctx->stream>>>(map, kf1, Twf, w, pt_12, pt_21, ln_12, ln_21, out,
                                                                         ctx->cam, ctx->cfg);
  GFPL_LAUNCH_OK(ctx);
  return GFPL_OK;
}

}  // namespace gfpl
