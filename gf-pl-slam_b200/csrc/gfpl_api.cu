// gfpl_api.cu — the C ABI of include/gfpl_c.h: context, host-pointer entry points (upload -> kernels ->
// download on the context stream) and the resident replay pipeline.
//
// There is no CPU implementation behind any compute entry point: without a CUDA device
// gfpl_ctx_create fails with GFPL_E_NO_DEVICE and nothing else can be called.
#include <stdarg.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <thread>
#include <vector>
#include <new>

#include "gfpl_internal.h"

namespace gfpl {

static thread_local std::string tls_error;

void set_last_error(gfpl_ctx* ctx, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  if (ctx) ctx->last_error = buf;
  tls_error = buf;
}

// Stream-ordered device allocation that frees itself (cudaMallocAsync pool keeps the memory cached).
struct DevMem {
  gfpl_ctx* ctx;
  std::vector<void*> ptrs;
  explicit DevMem(gfpl_ctx* c) : ctx(c) {}
  ~DevMem() {
    for (void* p : ptrs) cudaFreeAsync(p, ctx->stream);
  }
  template <typename T>
  T* alloc(size_t n) {
    void* p = nullptr;
    if (n == 0) n = 1;
    const cudaError_t e = ctx->pool ? cudaMallocFromPoolAsync(&p, n * sizeof(T), ctx->pool, ctx->stream)
                                    : cudaMallocAsync(&p, n * sizeof(T), ctx->stream);
    if (e != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ptrs.push_back(p);
    return (T*)p;
  }
  template <typename T>
  T* upload(const T* host, size_t n) {
    T* d = alloc<T>(n);
    if (d && n && host) {
      const cudaError_t e = cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
      if (e != cudaSuccess) {   // reported as a failed mirror (REQUIRE_MEM) and again by finish()
        if (ctx->pending_error == cudaSuccess) ctx->pending_error = e;
        return nullptr;
      }
    }
    return d;
  }
};

template <typename T>
static inline void download(gfpl_ctx* ctx, T* host, const T* dev, size_t n) {
  if (host && n) {
    const cudaError_t e = cudaMemcpyAsync(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream);
    if (e != cudaSuccess && ctx->pending_error == cudaSuccess) ctx->pending_error = e;
  }
}

// end of a host-pointer entry point: results are on the host, and no copy of the call was dropped
static int finish(gfpl_ctx* ctx) {
  GFPL_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  if (ctx->pending_error != cudaSuccess) {
    const cudaError_t e = ctx->pending_error;
    ctx->pending_error = cudaSuccess;
    set_last_error(ctx, "an asynchronous copy of this call failed: %s", cudaGetErrorString(e));
    return GFPL_E_CUDA;
  }
  return GFPL_OK;
}

// per-problem counts of a batched view: every one within [0, cap] (the kernels size shared memory and output rows by
// the capacities and trust the counts)
static bool counts_ok(const int32_t* n, size_t B, int cap) {
  if (!n) return false;
  for (size_t p = 0; p < B; p++)
    if (n[p] < 0 || n[p] > cap) return false;
  return true;
}
static bool raw_counts_ok(const gfpl_raw_frames* h) {
  return h->cap_pt >= 0 && h->cap_ln >= 0 && counts_ok(h->n_pt_l, h->B, h->cap_pt) && counts_ok(h->n_pt_r, h->B, h->cap_pt) &&
         counts_ok(h->n_ln_l, h->B, h->cap_ln) && counts_ok(h->n_ln_r, h->B, h->cap_ln);
}
static bool stereo_counts_ok(const gfpl_stereo_frames* h) {
  return h->cap_pt >= 0 && h->cap_ln >= 0 && counts_ok(h->n_pt, h->B, h->cap_pt) && counts_ok(h->n_ln, h->B, h->cap_ln);
}
static bool obs_counts_ok(const gfpl_track_obs* h) {
  return h->cap_pt >= 0 && h->cap_ln >= 0 && counts_ok(h->n_pt, h->B, h->cap_pt) && counts_ok(h->n_ln, h->B, h->cap_ln);
}

// every matched line's prev_idx addresses a covariance record of its pair: 0 <= prev_idx < cov_cap
static bool line_cov_indices_ok(const gfpl_track_obs* h, int32_t cov_cap) {
  for (int p = 0; p < h->B; p++)
    for (int l = 0; l < h->n_ln[p]; l++) {
      const int32_t k = h->ln[(size_t)p * h->cap_ln + l].prev_idx;
      if (k < 0 || k >= cov_cap) return false;
    }
  return true;
}

// device mirrors of the batched views -------------------------------------------------------------
static bool mirror_raw(DevMem& m, const gfpl_raw_frames* h, DevRawFrames* d) {
  size_t B = h->B, np = B * h->cap_pt, nl = B * h->cap_ln;
  *d = *h;
  d->n_pt_l = m.upload(h->n_pt_l, B); d->n_pt_r = m.upload(h->n_pt_r, B);
  d->n_ln_l = m.upload(h->n_ln_l, B); d->n_ln_r = m.upload(h->n_ln_r, B);
  d->pt_l = m.upload(h->pt_l, np); d->pt_r = m.upload(h->pt_r, np);
  d->pdesc_l = m.upload(h->pdesc_l, np * 32); d->pdesc_r = m.upload(h->pdesc_r, np * 32);
  d->ln_l = m.upload(h->ln_l, nl); d->ln_r = m.upload(h->ln_r, nl);
  d->ldesc_l = m.upload(h->ldesc_l, nl * 32); d->ldesc_r = m.upload(h->ldesc_r, nl * 32);
  return d->n_pt_l && d->n_pt_r && d->n_ln_l && d->n_ln_r && d->pt_l && d->pt_r && d->pdesc_l && d->pdesc_r &&
         d->ln_l && d->ln_r && d->ldesc_l && d->ldesc_r;
}

static bool alloc_stereo(DevMem& m, int B, int cap_pt, int cap_ln, DevStereoFrames* d) {
  size_t np = (size_t)B * cap_pt, nl = (size_t)B * cap_ln;
  d->B = B; d->cap_pt = cap_pt; d->cap_ln = cap_ln;
  d->n_pt = m.alloc<int32_t>(B); d->n_ln = m.alloc<int32_t>(B);
  d->pt = m.alloc<gfpl_point_feature>(np); d->pdesc = m.alloc<uint8_t>(np * 32);
  d->ln = m.alloc<gfpl_line_feature>(nl); d->ldesc = m.alloc<uint8_t>(nl * 32);
  return d->n_pt && d->n_ln && d->pt && d->pdesc && d->ln && d->ldesc;
}

static bool mirror_stereo(DevMem& m, const gfpl_stereo_frames* h, DevStereoFrames* d) {
  size_t B = h->B, np = B * h->cap_pt, nl = B * h->cap_ln;
  *d = *h;
  d->n_pt = m.upload(h->n_pt, B); d->n_ln = m.upload(h->n_ln, B);
  d->pt = m.upload(h->pt, np); d->pdesc = m.upload(h->pdesc, np * 32);
  d->ln = m.upload(h->ln, nl); d->ldesc = m.upload(h->ldesc, nl * 32);
  return d->n_pt && d->n_ln && d->pt && d->pdesc && d->ln && d->ldesc;
}

static void download_stereo(gfpl_ctx* ctx, const gfpl_stereo_frames* h, const DevStereoFrames& d) {
  size_t B = h->B;
  download(ctx, h->n_pt, d.n_pt, B);
  download(ctx, h->n_ln, d.n_ln, B);
  if (h->cap_pt == d.cap_pt) {
    download(ctx, h->pt, d.pt, B * d.cap_pt);
    download(ctx, h->pdesc, d.pdesc, B * d.cap_pt * 32);
  } else {
    for (size_t p = 0; p < B; p++) {
      download(ctx, h->pt + p * h->cap_pt, d.pt + p * d.cap_pt, (size_t)d.cap_pt);
      download(ctx, h->pdesc + p * h->cap_pt * 32, d.pdesc + p * d.cap_pt * 32, (size_t)d.cap_pt * 32);
    }
  }
  if (h->cap_ln == d.cap_ln) {
    download(ctx, h->ln, d.ln, B * d.cap_ln);
    download(ctx, h->ldesc, d.ldesc, B * d.cap_ln * 32);
  } else {
    for (size_t p = 0; p < B; p++) {
      download(ctx, h->ln + p * h->cap_ln, d.ln + p * d.cap_ln, (size_t)d.cap_ln);
      download(ctx, h->ldesc + p * h->cap_ln * 32, d.ldesc + p * d.cap_ln * 32, (size_t)d.cap_ln * 32);
    }
  }
}

static bool alloc_obs(DevMem& m, int B, int cap_pt, int cap_ln, DevTrackObs* d) {
  size_t np = (size_t)B * cap_pt, nl = (size_t)B * cap_ln;
  d->B = B; d->cap_pt = cap_pt; d->cap_ln = cap_ln;
  d->n_pt = m.alloc<int32_t>(B); d->n_ln = m.alloc<int32_t>(B);
  d->pt = m.alloc<gfpl_pt_obs>(np); d->pt_inlier = m.alloc<uint8_t>(np);
  d->ln = m.alloc<gfpl_ln_obs>(nl); d->ln_inlier = m.alloc<uint8_t>(nl);
  return d->n_pt && d->n_ln && d->pt && d->pt_inlier && d->ln && d->ln_inlier;
}

static bool mirror_obs(DevMem& m, const gfpl_track_obs* h, DevTrackObs* d) {
  size_t B = h->B, np = B * h->cap_pt, nl = B * h->cap_ln;
  *d = *h;
  d->n_pt = m.upload(h->n_pt, B); d->n_ln = m.upload(h->n_ln, B);
  d->pt = m.upload(h->pt, np); d->pt_inlier = m.upload(h->pt_inlier, np);
  d->ln = m.upload(h->ln, nl); d->ln_inlier = m.upload(h->ln_inlier, nl);
  return d->n_pt && d->n_ln && d->pt && d->pt_inlier && d->ln && d->ln_inlier;
}

static void download_obs(gfpl_ctx* ctx, const gfpl_track_obs* h, const DevTrackObs& d) {
  size_t B = h->B;
  download(ctx, h->n_pt, d.n_pt, B);
  download(ctx, h->n_ln, d.n_ln, B);
  download(ctx, h->pt, d.pt, B * d.cap_pt);
  download(ctx, h->pt_inlier, d.pt_inlier, B * d.cap_pt);
  download(ctx, h->ln, d.ln, B * d.cap_ln);
  download(ctx, h->ln_inlier, d.ln_inlier, B * d.cap_ln);
}

static bool alloc_knn(DevMem& m, size_t n, Knn2Out* k) {
  k->idx = m.alloc<int2>(n);
  k->dist = m.alloc<float2>(n);
  return k->idx && k->dist;
}

}  // namespace gfpl

using namespace gfpl;

#define REQUIRE(ctx, cond)                                        \
  do {                                                            \
    if (!(cond)) {                                                \
      gfpl::set_last_error(ctx, "invalid argument: %s", #cond);   \
      return GFPL_E_INVALID;                                      \
    }                                                             \
  } while (0)
#define REQUIRE_MEM(ctx, cond)                                    \
  do {                                                            \
    if (!(cond)) {                                                \
      gfpl::set_last_error(ctx, "device allocation failed");      \
      return GFPL_E_NOMEM;                                        \
    }                                                             \
  } while (0)
#define CHECK(expr)              \
  do {                           \
    int _rc = (expr);            \
    if (_rc != GFPL_OK) return _rc; \
  } while (0)

extern "C" {

int gfpl_abi_version(void) { return GFPL_ABI_VERSION; }

void gfpl_config_default(gfpl_config* c) {
  memset(c, 0, sizeof(*c));
  c->has_points = 1; c->has_lines = 1; c->lr_in_parallel = 1; c->best_lr_matches = 1;
  c->use_line_conf_cut = 1; c->max_vol_line_cut = 1;
  c->max_line_match_num = 300; c->max_point_match_num = 500;
  c->min_features = 10; c->max_iters = 5; c->max_iters_ref = 10; c->max_iters_lba = 20;
  c->ratio_disp_std = 0.15; c->ratio_disp_std_hor = 0.9;
  c->max_dist_epip = 2.0; c->min_disp = 1.0; c->max_ratio_12_p = 0.9; c->point_match_radius = 50.0;
  c->stereo_overlap_th = 0.5; c->line_horiz_th = 0.1; c->desc_th_l = 0.1; c->line_cov_th = 10.0;
  c->line_match_radius = 80.0;
  c->homog_th = 0.0000001; c->min_error = 0.0000001; c->min_error_change = 0.0000001;
  c->inlier_k = 2.0; c->motion_step_th = 10; c->orb_scale_factor = 1.2; c->lsd_scale = 1;
  c->lambda_lba_lm = 0.001; c->lambda_lba_k = 10.0;
  c->f2f_point_window = 10.0; c->budget_slack = 1.2; c->cut_step = 0.05; c->cut_rng[0] = 0; c->cut_rng[1] = 1.0;
  c->mad_uninit_value = 0.0;
  c->max_kf_epip_p = 1.0; c->max_kf_epip_l = 1.0;
  c->orb_nlevels = 4;
}

void gfpl_camera_euroc(gfpl_camera* c) {
  c->fx = 458.654; c->fy = 458.654; c->cx = 367.215; c->cy = 248.375; c->b = 0.110078; c->width = 752; c->height = 480;
}
void gfpl_camera_kitti(gfpl_camera* c) {
  c->fx = 718.856; c->fy = 718.856; c->cx = 607.1928; c->cy = 185.2157; c->b = 0.537166; c->width = 1241; c->height = 376;
}

int gfpl_ctx_create(int device, const gfpl_camera* cam, const gfpl_config* cfg, gfpl_ctx** out) {
  if (!out) return GFPL_E_INVALID;
  *out = nullptr;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) {
    cudaGetLastError();
    set_last_error(nullptr, "no CUDA device %d (found %d): this library has no CPU fallback", device, ndev);
    return GFPL_E_NO_DEVICE;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) {
    set_last_error(nullptr, "device %d is sm_%d%d; libgfpl is built for sm_100a only", device, prop.major, prop.minor);
    return GFPL_E_NO_DEVICE;
  }
  gfpl_ctx* ctx = new (std::nothrow) gfpl_ctx();
  if (!ctx) return GFPL_E_NOMEM;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
    set_last_error(nullptr, "cannot create a stream on device %d", device);
    delete ctx;
    return GFPL_E_CUDA;
  }
  {  // scratch of the host-pointer entry points: a pool of this context's own (cached between calls, returned to the
     // driver by gfpl_ctx_destroy; the process-wide default pool is left alone)
    cudaMemPoolProps props{};
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = device;
    if (cudaMemPoolCreate(&ctx->pool, &props) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(ctx->pool, cudaMemPoolAttrReleaseThreshold, &keep);
    } else {
      cudaGetLastError();
      ctx->pool = nullptr;   // fall back to the default pool with its default release policy
    }
  }
  if (cam) ctx->cam = *cam; else gfpl_camera_euroc(&ctx->cam);
  if (cfg) ctx->cfg = *cfg; else gfpl_config_default(&ctx->cfg);
  *out = ctx;
  return GFPL_OK;
}

void gfpl_ctx_destroy(gfpl_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  for (auto& e : ctx->timer)
    if (e) cudaEventDestroy(e);
  cudaStreamDestroy(ctx->stream);
  if (ctx->pool) cudaMemPoolDestroy(ctx->pool);
  if (ctx->staging) cudaFreeHost(ctx->staging);
  delete ctx;
}

int gfpl_ctx_set_config(gfpl_ctx* ctx, const gfpl_config* cfg) {
  if (!ctx || !cfg) return GFPL_E_INVALID;
  ctx->cfg = *cfg;
  return GFPL_OK;
}
int gfpl_ctx_set_camera(gfpl_ctx* ctx, const gfpl_camera* cam) {
  if (!ctx || !cam) return GFPL_E_INVALID;
  ctx->cam = *cam;
  return GFPL_OK;
}
const char* gfpl_last_error(const gfpl_ctx* ctx) { return ctx ? ctx->last_error.c_str() : tls_error.c_str(); }
int gfpl_ctx_sync(gfpl_ctx* ctx) {
  if (!ctx) return GFPL_E_INVALID;
  return finish(ctx);
}
int64_t gfpl_ctx_launch_count(const gfpl_ctx* ctx) { return ctx ? ctx->launches : 0; }
void* gfpl_ctx_stream(gfpl_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

// ---- M1/M2 --------------------------------------------------------------------------------------------
int gfpl_knn2_hamming256_batch(gfpl_ctx* ctx, int32_t B, const uint8_t* q, const int32_t* nq, int32_t cap_q,
                               const uint8_t* t, const int32_t* nt, int32_t cap_t, int32_t norm, int32_t* out_idx,
                               float* out_dist) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, B >= 0 && cap_q >= 0 && cap_t >= 0);
  REQUIRE(ctx, norm == GFPL_NORM_HAMMING || norm == GFPL_NORM_HAMMING2);
  if (B == 0 || cap_q == 0) return GFPL_OK;
  REQUIRE(ctx, q && nq && t && nt && out_idx && out_dist && cap_t > 0);
  REQUIRE(ctx, cap_q < GFPL_MAX_KNN_ROWS && cap_t < GFPL_MAX_KNN_ROWS);   // the sort key packs a row index in 20 bits
  REQUIRE(ctx, counts_ok(nq, B, cap_q) && counts_ok(nt, B, cap_t));
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  size_t nqt = (size_t)B * cap_q, ntt = (size_t)B * cap_t;
  uint8_t* dq = m.upload(q, nqt * 32);
  uint8_t* dt = m.upload(t, ntt * 32);
  int32_t* dnq = m.upload(nq, B);
  int32_t* dnt = m.upload(nt, B);
  Knn2Out o{}, none{};
  REQUIRE_MEM(ctx, dq && dt && dnq && dnt && alloc_knn(m, nqt, &o));
  // rows past nq[p] are never written by the kernel: give them the "missing" value
  GFPL_CUDA_OK(ctx, cudaMemsetAsync(o.idx, 0xFF, nqt * sizeof(int2), ctx->stream));
  GFPL_CUDA_OK(ctx, cudaMemsetAsync(o.dist, 0, nqt * sizeof(float2), ctx->stream));
  CHECK(launch_knn2(ctx, B, dq, dnq, cap_q, dt, dnt, cap_t, norm, o, none));
  download(ctx, (int2*)out_idx, o.idx, nqt);
  download(ctx, (float2*)out_dist, o.dist, nqt);
  return finish(ctx);
}

int gfpl_knn2_mutual_batch(gfpl_ctx* ctx, int32_t B, const uint8_t* q, const int32_t* nq, int32_t cap_q, const uint8_t* t,
                           const int32_t* nt, int32_t cap_t, int32_t norm, int32_t* out_idx, float* out_dist,
                           int32_t* out_best_idx, float* out_best_dist) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, B >= 0 && cap_q > 0 && cap_t > 0);
  REQUIRE(ctx, norm == GFPL_NORM_HAMMING || norm == GFPL_NORM_HAMMING2);
  if (B == 0) return GFPL_OK;
  REQUIRE(ctx, q && nq && t && nt && out_idx && out_dist && out_best_idx && out_best_dist);
  REQUIRE(ctx, cap_q < GFPL_MAX_KNN_ROWS && cap_t < GFPL_MAX_KNN_ROWS);
  REQUIRE(ctx, counts_ok(nq, B, cap_q) && counts_ok(nt, B, cap_t));
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  size_t nqt = (size_t)B * cap_q, ntt = (size_t)B * cap_t;
  uint8_t* dq = m.upload(q, nqt * 32);
  uint8_t* dt = m.upload(t, ntt * 32);
  int32_t* dnq = m.upload(nq, B);
  int32_t* dnt = m.upload(nt, B);
  Knn2Out o{}, r{};
  REQUIRE_MEM(ctx, dq && dt && dnq && dnt && alloc_knn(m, nqt, &o) && alloc_knn(m, ntt, &r));
  GFPL_CUDA_OK(ctx, cudaMemsetAsync(o.idx, 0xFF, nqt * sizeof(int2), ctx->stream));
  GFPL_CUDA_OK(ctx, cudaMemsetAsync(o.dist, 0, nqt * sizeof(float2), ctx->stream));
  GFPL_CUDA_OK(ctx, cudaMemsetAsync(r.idx, 0xFF, ntt * sizeof(int2), ctx->stream));
  GFPL_CUDA_OK(ctx, cudaMemsetAsync(r.dist, 0, ntt * sizeof(float2), ctx->stream));
  CHECK(launch_knn2(ctx, B, dq, dnq, cap_q, dt, dnt, cap_t, norm, o, r, /*reverse_best_only=*/true));
  download(ctx, (int2*)out_idx, o.idx, nqt);
  download(ctx, (float2*)out_dist, o.dist, nqt);
  std::vector<int2> ri(ntt);
  std::vector<float2> rd(ntt);
  download(ctx, ri.data(), r.idx, ntt);
  download(ctx, rd.data(), r.dist, ntt);
  CHECK(finish(ctx));
  for (size_t k = 0; k < ntt; k++) { out_best_idx[k] = ri[k].x; out_best_dist[k] = rd[k].x; }
  return GFPL_OK;
}

int gfpl_knn2_hamming256(gfpl_ctx* ctx, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t norm,
                         int32_t* out_idx, float* out_dist) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, nq >= 0 && nt >= 2);  // the reference reads [i][1]
  if (nq == 0) return GFPL_OK;
  return gfpl_knn2_hamming256_batch(ctx, 1, q, &nq, nq, t, &nt, nt, norm, out_idx, out_dist);
}

// ---- M3 -----------------------------------------------------------------------------------------------
int gfpl_radius_hamming256(gfpl_ctx* ctx, const uint8_t* q, int32_t nq, const uint8_t* t, int32_t nt, int32_t norm,
                           float radius, int32_t* offsets, int32_t* out_idx, float* out_dist, int32_t cap) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, nq >= 0 && nt >= 0 && offsets && cap >= 0);
  REQUIRE(ctx, nq < GFPL_MAX_KNN_ROWS && nt < GFPL_MAX_KNN_ROWS);
  REQUIRE(ctx, norm == GFPL_NORM_HAMMING || norm == GFPL_NORM_HAMMING2);
  offsets[0] = 0;
  if (nq == 0) return GFPL_OK;
  REQUIRE(ctx, q && (nt == 0 || t));
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  uint8_t* dq = m.upload(q, (size_t)nq * 32);
  uint8_t* dt = m.upload(t, (size_t)nt * 32);
  int32_t* dcnt = m.alloc<int32_t>(nq);
  REQUIRE_MEM(ctx, dq && dt && dcnt);
  CHECK(launch_radius_count(ctx, dq, nq, dt, nt, norm, radius, dcnt));
  std::vector<int32_t> cnt(nq);
  download(ctx, cnt.data(), dcnt, nq);
  CHECK(finish(ctx));
  int64_t total = 0;
  for (int i = 0; i < nq; i++) { total += cnt[i]; offsets[i + 1] = (int32_t)total; }
  if (total > cap) {
    set_last_error(ctx, "radius match needs %lld entries, capacity is %d", (long long)total, cap);
    return GFPL_E_CAPACITY;
  }
  if (total == 0) return GFPL_OK;
  REQUIRE(ctx, out_idx && out_dist);
  int32_t* doff = m.upload(offsets, (size_t)nq + 1);
  int32_t* didx = m.alloc<int32_t>(total);
  float* ddist = m.alloc<float>(total);
  REQUIRE_MEM(ctx, doff && didx && ddist);
  CHECK(launch_radius_fill(ctx, dq, nq, dt, nt, norm, radius, doff, didx, ddist));
  download(ctx, out_idx, didx, total);
  download(ctx, out_dist, ddist, total);
  return finish(ctx);
}

int gfpl_measure_alu_peak(gfpl_ctx* ctx, double* lop3_tops) {
  if (!ctx || !lop3_tops) return GFPL_E_INVALID;
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return launch_measure_alu_peak(ctx, lop3_tops);
}

int gfpl_measure_int8_peak(gfpl_ctx* ctx, double* int8_tops) {
  if (!ctx || !int8_tops) return GFPL_E_INVALID;
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return launch_measure_int8_peak(ctx, int8_tops);
}

int gfpl_measure_peaks(gfpl_ctx* ctx, double* popc_tops, double* fp64_tflops) {
  if (!ctx || !popc_tops || !fp64_tflops) return GFPL_E_INVALID;
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return launch_measure_peaks(ctx, popc_tops, fp64_tflops);
}

}  // extern "C"

// =========================================================================================================
// Stage entry points, whole step, resident replay
// =========================================================================================================
namespace gfpl {

// scratch of one tracking step for B pairs
struct StepWork {
  Knn2Out pt_lr{}, pt_rl{}, ln_lr{}, ln_rl{};   // stereo knn (current frame)
  Knn2Out pt_12{}, pt_21{}, ln_12{}, ln_21{};   // f2f knn
  uint32_t* best_key = nullptr;                 // [B][cap_pt] hybrid point resolver
  double* f2f_T = nullptr;                      // [B][32] prev Tfw | (predicted curr Tfw)^-1
  gfpl_line_cov* prev_cov = nullptr;            // [B][cap_ln]
};

static bool alloc_work(DevMem& m, int B, int cap_pt, int cap_ln, StepWork* w) {
  size_t np = (size_t)B * cap_pt, nl = (size_t)B * cap_ln;
  bool ok = alloc_knn(m, np, &w->pt_lr) && alloc_knn(m, np, &w->pt_rl) && alloc_knn(m, nl, &w->ln_lr) &&
            alloc_knn(m, nl, &w->ln_rl) && alloc_knn(m, np, &w->pt_12) && alloc_knn(m, np, &w->pt_21) &&
            alloc_knn(m, nl, &w->ln_12) && alloc_knn(m, nl, &w->ln_21);
  w->best_key = m.alloc<uint32_t>(np);
  w->prev_cov = m.alloc<gfpl_line_cov>(nl);
  w->f2f_T = m.alloc<double>((size_t)B * 32);
  return ok && w->best_key && w->prev_cov && w->f2f_T;
}

// ---- views of the pairs [lo, lo + n) of a batch (every array is [B][cap]: a view is a pointer offset) ---------
static Knn2Out slice(const Knn2Out& k, size_t off) { return Knn2Out{k.idx ? k.idx + off : nullptr, k.dist ? k.dist + off : nullptr}; }
static DevRawFrames slice(const DevRawFrames& r, int lo, int n) {
  DevRawFrames v = r;
  const size_t op = (size_t)lo * r.cap_pt, ol = (size_t)lo * r.cap_ln;
  v.B = n;
  v.n_pt_l += lo; v.n_pt_r += lo; v.n_ln_l += lo; v.n_ln_r += lo;
  v.pt_l += op; v.pt_r += op; v.pdesc_l += op * 32; v.pdesc_r += op * 32;
  v.ln_l += ol; v.ln_r += ol; v.ldesc_l += ol * 32; v.ldesc_r += ol * 32;
  return v;
}
static DevStereoFrames slice(const DevStereoFrames& f, int lo, int n) {
  DevStereoFrames v = f;
  const size_t op = (size_t)lo * f.cap_pt, ol = (size_t)lo * f.cap_ln;
  v.B = n;
  v.n_pt += lo; v.n_ln += lo;
  v.pt += op; v.pdesc += op * 32; v.ln += ol; v.ldesc += ol * 32;
  return v;
}
static DevTrackObs slice(const DevTrackObs& o, int lo, int n) {
  DevTrackObs v = o;
  const size_t op = (size_t)lo * o.cap_pt, ol = (size_t)lo * o.cap_ln;
  v.B = n;
  v.n_pt += lo; v.n_ln += lo;
  v.pt += op; v.pt_inlier += op; v.ln += ol; v.ln_inlier += ol;
  return v;
}
// cap_pt / cap_ln: the strides the scratch was allocated with
static StepWork slice(const StepWork& w, int lo, int cap_pt, int cap_ln) {
  StepWork v = w;
  const size_t op = (size_t)lo * cap_pt, ol = (size_t)lo * cap_ln;
  v.pt_lr = slice(w.pt_lr, op); v.pt_rl = slice(w.pt_rl, op); v.pt_12 = slice(w.pt_12, op); v.pt_21 = slice(w.pt_21, op);
  v.ln_lr = slice(w.ln_lr, ol); v.ln_rl = slice(w.ln_rl, ol); v.ln_12 = slice(w.ln_12, ol); v.ln_21 = slice(w.ln_21, ol);
  v.best_key += op; v.prev_cov += ol; v.f2f_T += (size_t)lo * 32;
  return v;
}

static int run_stereo(gfpl_ctx* ctx, const DevRawFrames& raw, int variant, StepWork& w, DevStereoFrames curr,
                      cudaEvent_t mid = nullptr) {
  const int B = raw.B;
  if (ctx->cfg.has_points)
    CHECK(launch_knn2(ctx, B, raw.pdesc_l, raw.n_pt_l, raw.cap_pt, raw.pdesc_r, raw.n_pt_r, raw.cap_pt,
                      GFPL_NORM_HAMMING, w.pt_lr, ctx->cfg.best_lr_matches ? w.pt_rl : Knn2Out{}, true));
  if (mid) GFPL_CUDA_OK(ctx, cudaEventRecord(mid, ctx->stream));
  if (ctx->cfg.has_lines)
    CHECK(launch_knn2(ctx, B, raw.ldesc_l, raw.n_ln_l, raw.cap_ln, raw.ldesc_r, raw.n_ln_r, raw.cap_ln,
                      variant == GFPL_STEREO_INITIAL ? GFPL_NORM_HAMMING : GFPL_NORM_HAMMING2, w.ln_lr,
                      ctx->cfg.best_lr_matches ? w.ln_rl : Knn2Out{}, true));
  return GFPL_OK;
}

static int run_f2f_knn(gfpl_ctx* ctx, const DevStereoFrames& prev, const DevStereoFrames& curr,
                       const gfpl_frame_state* state, int mode, StepWork& w, cudaEvent_t mid = nullptr) {
  const int B = prev.B;
  const bool use_radius = ctx->cfg.best_lr_matches && ctx->cfg.lr_in_parallel;
  if (ctx->cfg.has_points) {
    if (mode == GFPL_F2F_BRUTE)
      CHECK(launch_knn2(ctx, B, prev.pdesc, prev.n_pt, prev.cap_pt, curr.pdesc, curr.n_pt, curr.cap_pt,
                        GFPL_NORM_HAMMING, w.pt_12, ctx->cfg.best_lr_matches ? w.pt_21 : Knn2Out{}, true));
    else if (use_radius)
      CHECK(launch_f2f_hybrid_points(ctx, prev, curr, state, w.f2f_T, w.best_key));
    else {
      CHECK(launch_knn2(ctx, B, prev.pdesc, prev.n_pt, prev.cap_pt, curr.pdesc, curr.n_pt, curr.cap_pt,
                        GFPL_NORM_HAMMING, w.pt_12, Knn2Out{}));
      CHECK(launch_f2f_hybrid_points_knn(ctx, prev, curr, state, w.pt_12, w.best_key));
    }
  }
  if (mid) GFPL_CUDA_OK(ctx, cudaEventRecord(mid, ctx->stream));
  if (ctx->cfg.has_lines)
    CHECK(launch_knn2(ctx, B, prev.ldesc, prev.n_ln, prev.cap_ln, curr.ldesc, curr.n_ln, curr.cap_ln,
                      GFPL_NORM_HAMMING, w.ln_12, ctx->cfg.best_lr_matches ? w.ln_21 : Knn2Out{}, true));
  return GFPL_OK;
}

}  // namespace gfpl

struct gfpl_replay {
  gfpl_ctx* ctx = nullptr;
  int B = 0, cap_pt = 0, cap_ln = 0;
  std::vector<void*> owned;
  DevRawFrames raw{};
  DevStereoFrames prev{}, curr{};
  DevTrackObs obs{};
  gfpl_frame_state* state = nullptr;
  gfpl_prev_frames slim{};   // staging for gfpl_replay_upload_slim (allocated on first use)
  gfpl_compact_input compact{};  // staging for gfpl_replay_upload_compact (device copies of the compact records)
  gfpl_line_cut* cut = nullptr;
  gfpl_pose_result* pose = nullptr;
  StepWork w{};
  cudaEvent_t ev[GFPL_NUM_STAGES] = {};
  int64_t h2d = 0, d2h = 0;
  bool timed = false;
  void* pose_raw = nullptr;   // optimiser -> finalize hand-off, [B]
  double* stamps = nullptr;   // time stamps of the next current frames (gfpl_replay_advance)
  // software pipeline over sub-batches (gfpl_replay_set_pipeline)
  int n_sub = 1;
  cudaStream_t s_match = nullptr;
  cudaStream_t s_solve[GFPL_MAX_SUBBATCHES] = {};
  cudaEvent_t ev_match[GFPL_MAX_SUBBATCHES] = {}, ev_solve[GFPL_MAX_SUBBATCHES] = {};
  cudaEvent_t ev_input = nullptr;
  bool solve_pending[GFPL_MAX_SUBBATCHES] = {};   // ev_solve[k] was recorded by an earlier run
  bool input_dirty = true;                        // the context stream touched the inputs since the last run
  template <typename T>
  T* alloc(size_t n) {
    void* p = nullptr;
    if (cudaMalloc(&p, (n ? n : 1) * sizeof(T)) != cudaSuccess) return nullptr;
    owned.push_back(p);
    return (T*)p;
  }
};

extern "C" {

int gfpl_stereo_match(gfpl_ctx* ctx, const gfpl_raw_frames* raw, int32_t variant, gfpl_stereo_frames* out) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, raw && out && raw->B >= 0 && raw->B == out->B);
  REQUIRE(ctx, out->cap_pt >= raw->cap_pt && out->cap_ln >= raw->cap_ln);
  REQUIRE(ctx, variant == GFPL_STEREO_INITIAL || variant == GFPL_STEREO_BRUTE);
  if (raw->B == 0) return GFPL_OK;
  REQUIRE(ctx, raw_counts_ok(raw) && raw->cap_pt < GFPL_MAX_KNN_ROWS && raw->cap_ln < GFPL_MAX_KNN_ROWS);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  DevRawFrames draw;
  DevStereoFrames dcur;
  StepWork w;
  REQUIRE_MEM(ctx, mirror_raw(m, raw, &draw) && alloc_stereo(m, raw->B, raw->cap_pt, raw->cap_ln, &dcur) &&
                       alloc_work(m, raw->B, raw->cap_pt, raw->cap_ln, &w));
  CHECK(run_stereo(ctx, draw, variant, w, dcur));
  CHECK(launch_stereo_accept(ctx, draw, variant, w.pt_lr, w.pt_rl, w.ln_lr, w.ln_rl, dcur));
  download_stereo(ctx, out, dcur);
  return finish(ctx);
}

int gfpl_stereo_match_orbslam(gfpl_ctx* ctx, const gfpl_raw_frames* raw, const gfpl_pyramid* pyr, gfpl_stereo_frames* out) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, raw && pyr && out && raw->B >= 0 && raw->B == out->B);
  REQUIRE(ctx, out->cap_pt >= raw->cap_pt && out->cap_ln >= raw->cap_ln);
  REQUIRE(ctx, pyr->data && pyr->n_levels == ctx->cfg.orb_nlevels && pyr->n_levels >= 1 && pyr->n_levels <= GFPL_MAX_PYR_LEVELS);
  for (int l = 0; l < pyr->n_levels; l++) {
    REQUIRE(ctx, pyr->width[l] > 0 && pyr->height[l] > 0 && pyr->row_stride[l] >= pyr->width[l] && pyr->offset[l] >= 0);
    REQUIRE(ctx, pyr->offset[l] + (int64_t)pyr->row_stride[l] * (pyr->height[l] - 1) + pyr->width[l] <= pyr->frame_stride);
  }
  if (raw->B == 0) return GFPL_OK;
  REQUIRE(ctx, raw_counts_ok(raw) && raw->cap_pt < GFPL_MAX_KNN_ROWS && raw->cap_ln < GFPL_MAX_KNN_ROWS);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  DevRawFrames draw;
  DevStereoFrames dcur;
  StepWork w;
  REQUIRE_MEM(ctx, mirror_raw(m, raw, &draw) && alloc_stereo(m, raw->B, raw->cap_pt, raw->cap_ln, &dcur) &&
                       alloc_work(m, raw->B, raw->cap_pt, raw->cap_ln, &w));
  gfpl_pyramid dpyr = *pyr;
  dpyr.data = m.upload(pyr->data, (size_t)raw->B * (size_t)pyr->frame_stride);
  REQUIRE_MEM(ctx, dpyr.data);
  if (ctx->cfg.has_lines)   // :633-767 == the line part of _Brute
    CHECK(launch_knn2(ctx, raw->B, draw.ldesc_l, draw.n_ln_l, draw.cap_ln, draw.ldesc_r, draw.n_ln_r, draw.cap_ln,
                      GFPL_NORM_HAMMING2, w.ln_lr, ctx->cfg.best_lr_matches ? w.ln_rl : Knn2Out{}, true));
  CHECK(launch_stereo_accept(ctx, draw, GFPL_STEREO_BRUTE, w.pt_lr, w.pt_rl, w.ln_lr, w.ln_rl, dcur, /*lines_only=*/true));
  CHECK(launch_stereo_orbslam_points(ctx, draw, dpyr, dcur));
  download_stereo(ctx, out, dcur);
  return finish(ctx);
}

int gfpl_stereo_uncertainty(gfpl_ctx* ctx, const gfpl_stereo_frames* frames, gfpl_line_cov* out) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, frames && out && frames->B >= 0);
  if (frames->B == 0 || frames->cap_ln == 0) return GFPL_OK;
  REQUIRE(ctx, stereo_counts_ok(frames));
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  DevStereoFrames d;
  REQUIRE_MEM(ctx, mirror_stereo(m, frames, &d));
  size_t n = (size_t)frames->B * frames->cap_ln;
  gfpl_line_cov* dc = m.alloc<gfpl_line_cov>(n);
  REQUIRE_MEM(ctx, dc);
  GFPL_CUDA_OK(ctx, cudaMemsetAsync(dc, 0, n * sizeof(gfpl_line_cov), ctx->stream));
  CHECK(launch_stereo_uncertainty(ctx, d, dc));
  download(ctx, out, dc, n);
  return finish(ctx);
}

int gfpl_f2f_match(gfpl_ctx* ctx, const gfpl_stereo_frames* prev, const gfpl_stereo_frames* curr,
                   const gfpl_frame_state* state, int32_t mode, gfpl_track_obs* out) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, prev && curr && state && out && prev->B == curr->B && prev->B == out->B && prev->B >= 0);
  REQUIRE(ctx, mode == GFPL_F2F_HYBRID || mode == GFPL_F2F_BRUTE);
  REQUIRE(ctx, out->cap_pt > 0 && out->cap_ln > 0);
  if (prev->B == 0) return GFPL_OK;
  REQUIRE(ctx, stereo_counts_ok(prev) && stereo_counts_ok(curr));
  REQUIRE(ctx, prev->cap_pt < GFPL_MAX_KNN_ROWS && curr->cap_pt < GFPL_MAX_KNN_ROWS && prev->cap_ln < GFPL_MAX_KNN_ROWS &&
                   curr->cap_ln < GFPL_MAX_KNN_ROWS);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  DevStereoFrames dp, dc;
  DevTrackObs dobs;
  StepWork w;
  const int cp = prev->cap_pt > curr->cap_pt ? prev->cap_pt : curr->cap_pt;
  const int cl = prev->cap_ln > curr->cap_ln ? prev->cap_ln : curr->cap_ln;
  REQUIRE_MEM(ctx, mirror_stereo(m, prev, &dp) && mirror_stereo(m, curr, &dc) &&
                       alloc_obs(m, prev->B, out->cap_pt, out->cap_ln, &dobs) && alloc_work(m, prev->B, cp, cl, &w));
  gfpl_frame_state* dstate = m.upload(state, prev->B);
  REQUIRE_MEM(ctx, dstate);
  CHECK(run_f2f_knn(ctx, dp, dc, dstate, mode, w));
  CHECK(launch_f2f_accept(ctx, dp, dc, mode, w.best_key, w.pt_12, w.pt_21, w.ln_12, w.ln_21, dobs));
  download_obs(ctx, out, dobs);
  return finish(ctx);
}

int gfpl_kf_match(gfpl_ctx* ctx, const gfpl_stereo_frames* kf0, const gfpl_stereo_frames* kf1, const double* DT,
                  gfpl_track_obs* out) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, kf0 && kf1 && DT && out && kf0->B == kf1->B && kf0->B == out->B && kf0->B >= 0);
  REQUIRE(ctx, out->cap_pt > 0 && out->cap_ln > 0);
  if (kf0->B == 0) return GFPL_OK;
  REQUIRE(ctx, stereo_counts_ok(kf0) && stereo_counts_ok(kf1));
  REQUIRE(ctx, kf0->cap_pt < GFPL_MAX_KNN_ROWS && kf1->cap_pt < GFPL_MAX_KNN_ROWS && kf0->cap_ln < GFPL_MAX_KNN_ROWS &&
                   kf1->cap_ln < GFPL_MAX_KNN_ROWS);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  DevStereoFrames d0, d1;
  DevTrackObs dobs;
  StepWork w;
  const int B = kf0->B;
  const int cp = kf0->cap_pt > kf1->cap_pt ? kf0->cap_pt : kf1->cap_pt;
  const int cl = kf0->cap_ln > kf1->cap_ln ? kf0->cap_ln : kf1->cap_ln;
  REQUIRE_MEM(ctx, mirror_stereo(m, kf0, &d0) && mirror_stereo(m, kf1, &d1) &&
                       alloc_obs(m, B, out->cap_pt, out->cap_ln, &dobs) && alloc_work(m, B, cp, cl, &w));
  double* dDT = m.upload(DT, (size_t)B * 16);
  REQUIRE_MEM(ctx, dDT);
  if (ctx->cfg.has_points)
    CHECK(launch_knn2(ctx, B, d0.pdesc, d0.n_pt, d0.cap_pt, d1.pdesc, d1.n_pt, d1.cap_pt, GFPL_NORM_HAMMING, w.pt_12,
                      ctx->cfg.best_lr_matches ? w.pt_21 : Knn2Out{}, true));
  if (ctx->cfg.has_lines)
    CHECK(launch_knn2(ctx, B, d0.ldesc, d0.n_ln, d0.cap_ln, d1.ldesc, d1.n_ln, d1.cap_ln, GFPL_NORM_HAMMING, w.ln_12,
                      ctx->cfg.best_lr_matches ? w.ln_21 : Knn2Out{}, true));
  CHECK(launch_kf_accept(ctx, d0, d1, dDT, w.pt_12, w.pt_21, w.ln_12, w.ln_21, dobs));
  download_obs(ctx, out, dobs);
  return finish(ctx);
}

int gfpl_map_match(gfpl_ctx* ctx, const gfpl_local_map* map, const gfpl_stereo_frames* kf1, const double* Twf,
                   gfpl_map_matches* out) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, map && kf1 && Twf && out && map->B == kf1->B && map->B == out->B && map->B >= 0);
  REQUIRE(ctx, out->cap_pt > 0 && out->cap_ln > 0 && map->cap_pt > 0 && map->cap_ln > 0 && kf1->cap_pt > 0 && kf1->cap_ln > 0);
  if (map->B == 0) return GFPL_OK;
  REQUIRE(ctx, counts_ok(map->n_pt, map->B, map->cap_pt) && counts_ok(map->n_ln, map->B, map->cap_ln) && stereo_counts_ok(kf1));
  REQUIRE(ctx, map->cap_pt < GFPL_MAX_KNN_ROWS && map->cap_ln < GFPL_MAX_KNN_ROWS && kf1->cap_pt < GFPL_MAX_KNN_ROWS &&
                   kf1->cap_ln < GFPL_MAX_KNN_ROWS);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  const size_t B = map->B, mp = B * map->cap_pt, ml = B * map->cap_ln, kp = B * kf1->cap_pt, kl = B * kf1->cap_ln;
  gfpl_local_map dm = *map;
  dm.n_pt = m.upload(map->n_pt, B); dm.n_ln = m.upload(map->n_ln, B);
  dm.pt3d = m.upload(map->pt3d, mp * 3); dm.pdesc = m.upload(map->pdesc, mp * 32);
  dm.ln3d = m.upload(map->ln3d, ml * 6); dm.ldesc = m.upload(map->ldesc, ml * 32);
  DevStereoFrames d1;
  REQUIRE_MEM(ctx, dm.n_pt && dm.n_ln && dm.pt3d && dm.pdesc && dm.ln3d && dm.ldesc && mirror_stereo(m, kf1, &d1));
  double* dT = m.upload(Twf, B * 16);
  MapWork w;
  w.n_sel_pt = m.alloc<int32_t>(B); w.n_un_pt = m.alloc<int32_t>(B);
  w.n_sel_ln = m.alloc<int32_t>(B); w.n_un_ln = m.alloc<int32_t>(B);
  w.sel_pt = m.alloc<int32_t>(mp); w.sel_ln = m.alloc<int32_t>(ml);
  w.un_pt = m.alloc<int32_t>(kp); w.un_ln = m.alloc<int32_t>(kl);
  w.dsel_pt = m.alloc<uint8_t>(mp * 32); w.dsel_ln = m.alloc<uint8_t>(ml * 32);
  w.dun_pt = m.alloc<uint8_t>(kp * 32); w.dun_ln = m.alloc<uint8_t>(kl * 32);
  Knn2Out pt_12, pt_21, ln_12, ln_21;
  gfpl_map_matches dout = *out;
  dout.n_pt = m.alloc<int32_t>(B); dout.n_ln = m.alloc<int32_t>(B);
  dout.pt = m.alloc<gfpl_map_pair>(B * out->cap_pt); dout.ln = m.alloc<gfpl_map_pair>(B * out->cap_ln);
  REQUIRE_MEM(ctx, dT && w.n_sel_pt && w.n_un_pt && w.n_sel_ln && w.n_un_ln && w.sel_pt && w.sel_ln && w.un_pt && w.un_ln &&
                       w.dsel_pt && w.dsel_ln && w.dun_pt && w.dun_ln && alloc_knn(m, mp, &pt_12) && alloc_knn(m, kp, &pt_21) &&
                       alloc_knn(m, ml, &ln_12) && alloc_knn(m, kl, &ln_21) && dout.n_pt && dout.n_ln && dout.pt && dout.ln);
  CHECK(launch_map_select(ctx, dm, d1, dT, w));
  if (ctx->cfg.has_points)
    CHECK(launch_knn2(ctx, map->B, w.dsel_pt, w.n_sel_pt, map->cap_pt, w.dun_pt, w.n_un_pt, kf1->cap_pt, GFPL_NORM_HAMMING,
                      pt_12, ctx->cfg.best_lr_matches ? pt_21 : Knn2Out{}, true));
  if (ctx->cfg.has_lines)
    CHECK(launch_knn2(ctx, map->B, w.dsel_ln, w.n_sel_ln, map->cap_ln, w.dun_ln, w.n_un_ln, kf1->cap_ln, GFPL_NORM_HAMMING,
                      ln_12, ctx->cfg.best_lr_matches ? ln_21 : Knn2Out{}, true));
  CHECK(launch_map_accept(ctx, dm, d1, dT, w, pt_12, pt_21, ln_12, ln_21, dout));
  download(ctx, out->n_pt, dout.n_pt, B);
  download(ctx, out->n_ln, dout.n_ln, B);
  download(ctx, out->pt, dout.pt, B * out->cap_pt);
  download(ctx, out->ln, dout.ln, B * out->cap_ln);
  return finish(ctx);
}

int gfpl_medoid_desc(gfpl_ctx* ctx, const uint8_t* desc, const int32_t* offsets, int32_t L, int32_t* out_idx) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, desc && offsets && out_idx && L >= 0);
  if (L == 0) return GFPL_OK;
  REQUIRE(ctx, offsets[0] == 0 && offsets[L] >= 0 && offsets[L] < (1 << 20));
  for (int l = 0; l < L; l++) REQUIRE(ctx, offsets[l + 1] >= offsets[l]);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  uint8_t* dd = m.upload(desc, (size_t)offsets[L] * 32);
  int32_t* doff = m.upload(offsets, (size_t)L + 1);
  int32_t* dout = m.alloc<int32_t>(L);
  REQUIRE_MEM(ctx, dd && doff && dout);
  CHECK(launch_medoid(ctx, dd, doff, L, dout));
  download(ctx, out_idx, dout, (size_t)L);
  return finish(ctx);
}

int gfpl_linecut_submodular(gfpl_ctx* ctx, const double* DT_inv, gfpl_track_obs* obs, const gfpl_line_cov* prev_cov,
                            int32_t cov_cap, gfpl_line_cut* cut) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, DT_inv && obs && prev_cov && cut && obs->B >= 0 && cov_cap > 0);
  if (obs->B == 0) return GFPL_OK;
  REQUIRE(ctx, obs_counts_ok(obs));
  REQUIRE(ctx, line_cov_indices_ok(obs, cov_cap));
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  DevTrackObs dobs;
  REQUIRE_MEM(ctx, mirror_obs(m, obs, &dobs));
  size_t B = obs->B;
  double* dDT = m.upload(DT_inv, B * 16);
  gfpl_line_cov* dcov = m.upload(prev_cov, B * cov_cap);
  gfpl_line_cut* dcut = m.alloc<gfpl_line_cut>(B * obs->cap_ln);
  REQUIRE_MEM(ctx, dDT && dcov && dcut);
  GFPL_CUDA_OK(ctx, cudaMemsetAsync(dcut, 0, B * obs->cap_ln * sizeof(gfpl_line_cut), ctx->stream));
  CHECK(launch_linecut(ctx, dDT, nullptr, dobs, dcov, cov_cap, dcut));
  download(ctx, cut, dcut, B * obs->cap_ln);
  download(ctx, obs->ln, dobs.ln, B * obs->cap_ln);
  return finish(ctx);
}

int gfpl_linecut_descent(gfpl_ctx* ctx, const double* DT_inv, gfpl_track_obs* obs, const gfpl_line_cov* prev_cov,
                         int32_t cov_cap, gfpl_line_cut_descent* out) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, DT_inv && obs && prev_cov && out && obs->B >= 0 && cov_cap > 0);
  if (obs->B == 0) return GFPL_OK;
  REQUIRE(ctx, obs_counts_ok(obs));
  REQUIRE(ctx, line_cov_indices_ok(obs, cov_cap));
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  DevTrackObs dobs;
  REQUIRE_MEM(ctx, mirror_obs(m, obs, &dobs));
  size_t B = obs->B;
  double* dDT = m.upload(DT_inv, B * 16);
  gfpl_line_cov* dcov = m.upload(prev_cov, B * cov_cap);
  gfpl_line_cut_descent* dout = m.alloc<gfpl_line_cut_descent>(B * obs->cap_ln);
  REQUIRE_MEM(ctx, dDT && dcov && dout);
  GFPL_CUDA_OK(ctx, cudaMemsetAsync(dout, 0, B * obs->cap_ln * sizeof(gfpl_line_cut_descent), ctx->stream));
  CHECK(launch_linecut_descent(ctx, dDT, dobs, dcov, cov_cap, dout));
  download(ctx, out, dout, B * obs->cap_ln);
  download(ctx, obs->ln, dobs.ln, B * obs->cap_ln);
  return finish(ctx);
}

int gfpl_pose_hg(gfpl_ctx* ctx, const double* DT, const gfpl_track_obs* obs, double* H, double* g, double* e) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, DT && obs && H && g && e && obs->B >= 0);
  if (obs->B == 0) return GFPL_OK;
  REQUIRE(ctx, obs_counts_ok(obs));
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  DevTrackObs dobs;
  REQUIRE_MEM(ctx, mirror_obs(m, obs, &dobs));
  size_t B = obs->B;
  double* dDT = m.upload(DT, B * 16);
  double* dH = m.alloc<double>(B * 36);
  double* dg = m.alloc<double>(B * 6);
  double* de = m.alloc<double>(B);
  REQUIRE_MEM(ctx, dDT && dH && dg && de);
  CHECK(launch_pose_hg(ctx, dDT, dobs, dH, dg, de));
  download(ctx, H, dH, B * 36);
  download(ctx, g, dg, B * 6);
  download(ctx, e, de, B);
  return finish(ctx);
}

int gfpl_vector_stdv_mad(gfpl_ctx* ctx, const double* res, const int32_t* n, int32_t B, int32_t cap, double* out) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, res && n && out && B >= 0 && cap > 0 && cap < (1 << 14));
  if (B == 0) return GFPL_OK;
  for (int p = 0; p < B; p++) REQUIRE(ctx, n[p] >= 0 && n[p] <= cap);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  double* dres = m.upload(res, (size_t)B * cap);
  int32_t* dn = m.upload(n, (size_t)B);
  double* dout = m.alloc<double>((size_t)B);
  REQUIRE_MEM(ctx, dres && dn && dout);
  CHECK(launch_stdv_mad(ctx, dres, dn, B, cap, dout));
  download(ctx, out, dout, (size_t)B);
  return finish(ctx);
}

int gfpl_pose_optimize(gfpl_ctx* ctx, gfpl_track_obs* obs, const gfpl_frame_state* state, const double* DT_ini,
                       int32_t overload, gfpl_pose_result* out) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, obs && state && out && obs->B >= 0);
  REQUIRE(ctx, overload == GFPL_POSE_SEED_PREV_DT || overload == GFPL_POSE_SEED_GIVEN);
  if (obs->B == 0) return GFPL_OK;
  REQUIRE(ctx, obs_counts_ok(obs));
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  DevTrackObs dobs;
  REQUIRE_MEM(ctx, mirror_obs(m, obs, &dobs));
  size_t B = obs->B;
  gfpl_frame_state* dstate = m.upload(state, B);
  double* dDT = DT_ini ? m.upload(DT_ini, B * 16) : nullptr;
  gfpl_pose_result* dout = m.alloc<gfpl_pose_result>(B);
  REQUIRE_MEM(ctx, dstate && dout && (!DT_ini || dDT));
  CHECK(launch_pose_optimize(ctx, dobs, dstate, dDT, overload, dout));
  download(ctx, out, dout, B);
  download(ctx, obs->pt_inlier, dobs.pt_inlier, B * obs->cap_pt);
  download(ctx, obs->ln_inlier, dobs.ln_inlier, B * obs->cap_ln);
  return finish(ctx);
}

// ---- whole step ------------------------------------------------------------------------------------------
// The step in two halves.  run_match: everything up to the matched observation lists (tensor / integer pipes);
// run_solve: stereo uncertainty + line cut + pose optimisation (fp64, latency-bound).  ev: 9 stage events or NULL.
static int run_match(gfpl_ctx* ctx, const DevStereoFrames& prev, const gfpl_frame_state* state, const DevRawFrames& raw,
                     int flags, DevStereoFrames curr, DevTrackObs obs, StepWork& w, cudaEvent_t* ev) {
  const int variant = (flags & GFPL_TRACK_STEREO_INITIAL) ? GFPL_STEREO_INITIAL : GFPL_STEREO_BRUTE;
  const int mode = (flags & GFPL_TRACK_F2F_BRUTE) ? GFPL_F2F_BRUTE : GFPL_F2F_HYBRID;
  if (ev) GFPL_CUDA_OK(ctx, cudaEventRecord(ev[0], ctx->stream));
  CHECK(run_stereo(ctx, raw, variant, w, curr, ev ? ev[1] : nullptr));
  if (ev) GFPL_CUDA_OK(ctx, cudaEventRecord(ev[2], ctx->stream));
  CHECK(launch_stereo_accept(ctx, raw, variant, w.pt_lr, w.pt_rl, w.ln_lr, w.ln_rl, curr));
  if (ev) GFPL_CUDA_OK(ctx, cudaEventRecord(ev[3], ctx->stream));
  CHECK(run_f2f_knn(ctx, prev, curr, state, mode, w, ev ? ev[4] : nullptr));
  if (ev) GFPL_CUDA_OK(ctx, cudaEventRecord(ev[5], ctx->stream));
  CHECK(launch_f2f_accept(ctx, prev, curr, mode, w.best_key, w.pt_12, w.pt_21, w.ln_12, w.ln_21, obs));
  if (ev) GFPL_CUDA_OK(ctx, cudaEventRecord(ev[6], ctx->stream));
  return GFPL_OK;
}

static int run_solve(gfpl_ctx* ctx, const DevStereoFrames& prev, const gfpl_frame_state* state, int flags, DevTrackObs obs,
                     gfpl_line_cut* cut, gfpl_pose_result* pose, StepWork& w, cudaEvent_t* ev, void* pose_raw) {
  if ((flags & GFPL_TRACK_LINECUT) && ctx->cfg.use_line_conf_cut && cut) {
    // (estimateStereoUncertainty runs inside the cut kernel, for the matched lines only)
    CHECK(launch_linecut(ctx, nullptr, state, obs, w.prev_cov, prev.cap_ln, cut, prev.ln, prev.cap_ln));
  }
  if (ev) GFPL_CUDA_OK(ctx, cudaEventRecord(ev[7], ctx->stream));
  CHECK(launch_pose_optimize(ctx, obs, state, nullptr,
                             (flags & GFPL_TRACK_SEED_GIVEN) ? GFPL_POSE_SEED_GIVEN : GFPL_POSE_SEED_PREV_DT, pose, pose_raw));
  if (ev) GFPL_CUDA_OK(ctx, cudaEventRecord(ev[8], ctx->stream));
  return GFPL_OK;
}

static int run_step(gfpl_ctx* ctx, const DevStereoFrames& prev, const gfpl_frame_state* state, const DevRawFrames& raw,
                    int flags, DevStereoFrames curr, DevTrackObs obs, gfpl_line_cut* cut, gfpl_pose_result* pose,
                    StepWork& w, cudaEvent_t* ev, void* pose_raw = nullptr) {
  CHECK(run_match(ctx, prev, state, raw, flags, curr, obs, w, ev));
  return run_solve(ctx, prev, state, flags, obs, cut, pose, w, ev, pose_raw);
}

int gfpl_track(gfpl_ctx* ctx, const gfpl_stereo_frames* prev, const gfpl_frame_state* state,
               const gfpl_raw_frames* curr_raw, int32_t flags, gfpl_stereo_frames* curr, gfpl_track_obs* obs,
               gfpl_line_cut* cut, gfpl_pose_result* pose) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, prev && state && curr_raw && curr && obs && pose);
  const int B = curr_raw->B;
  REQUIRE(ctx, B >= 0 && prev->B == B && curr->B == B && obs->B == B);
  REQUIRE(ctx, curr->cap_pt == curr_raw->cap_pt && curr->cap_ln == curr_raw->cap_ln);
  REQUIRE(ctx, obs->cap_pt > 0 && obs->cap_ln > 0);
  if (B == 0) return GFPL_OK;
  REQUIRE(ctx, raw_counts_ok(curr_raw) && stereo_counts_ok(prev));
  REQUIRE(ctx, curr_raw->cap_pt < GFPL_MAX_KNN_ROWS && curr_raw->cap_ln < GFPL_MAX_KNN_ROWS && prev->cap_pt < GFPL_MAX_KNN_ROWS &&
                   prev->cap_ln < GFPL_MAX_KNN_ROWS);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  DevRawFrames draw;
  DevStereoFrames dprev, dcurr;
  DevTrackObs dobs;
  StepWork w;
  const int cp = prev->cap_pt > curr->cap_pt ? prev->cap_pt : curr->cap_pt;
  const int cl = prev->cap_ln > curr->cap_ln ? prev->cap_ln : curr->cap_ln;
  REQUIRE_MEM(ctx, mirror_raw(m, curr_raw, &draw) && mirror_stereo(m, prev, &dprev) &&
                       alloc_stereo(m, B, curr->cap_pt, curr->cap_ln, &dcurr) &&
                       alloc_obs(m, B, obs->cap_pt, obs->cap_ln, &dobs) && alloc_work(m, B, cp, cl, &w));
  gfpl_frame_state* dstate = m.upload(state, (size_t)B);
  gfpl_pose_result* dpose = m.alloc<gfpl_pose_result>(B);
  gfpl_line_cut* dcut = cut ? m.alloc<gfpl_line_cut>((size_t)B * obs->cap_ln) : nullptr;
  REQUIRE_MEM(ctx, dstate && dpose && (!cut || dcut));
  if (dcut) GFPL_CUDA_OK(ctx, cudaMemsetAsync(dcut, 0, (size_t)B * obs->cap_ln * sizeof(gfpl_line_cut), ctx->stream));
  CHECK(run_step(ctx, dprev, dstate, draw, flags, dcurr, dobs, dcut, dpose, w, nullptr));
  download_stereo(ctx, curr, dcurr);
  download_obs(ctx, obs, dobs);
  if (cut) download(ctx, cut, dcut, (size_t)B * obs->cap_ln);
  download(ctx, pose, dpose, (size_t)B);
  return finish(ctx);
}

// ---- resident replay ---------------------------------------------------------------------------------------
int gfpl_replay_create(gfpl_ctx* ctx, int32_t B, int32_t cap_pt, int32_t cap_ln, gfpl_replay** out) {
  return gfpl_replay_create_ex(ctx, B, cap_pt, cap_ln, cap_pt, cap_ln, out);
}

int gfpl_replay_create_ex(gfpl_ctx* ctx, int32_t B, int32_t cap_pt, int32_t cap_ln, int32_t prev_cap_pt,
                          int32_t prev_cap_ln, gfpl_replay** out) {
  if (!ctx || !out) return GFPL_E_INVALID;
  REQUIRE(ctx, B > 0 && cap_pt > 0 && cap_ln > 0 && prev_cap_pt > 0 && prev_cap_ln > 0);
  REQUIRE(ctx, prev_cap_pt <= cap_pt && prev_cap_ln <= cap_ln);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  gfpl_replay* rp = new (std::nothrow) gfpl_replay();
  if (!rp) return GFPL_E_NOMEM;
  rp->ctx = ctx; rp->B = B; rp->cap_pt = cap_pt; rp->cap_ln = cap_ln;
  size_t np = (size_t)B * cap_pt, nl = (size_t)B * cap_ln;
  const int ocp = ctx->cfg.max_point_match_num, ocl = ctx->cfg.max_line_match_num;
  DevRawFrames& r = rp->raw;
  r.B = B; r.cap_pt = cap_pt; r.cap_ln = cap_ln;
  r.n_pt_l = rp->alloc<int32_t>(B); r.n_pt_r = rp->alloc<int32_t>(B);
  r.n_ln_l = rp->alloc<int32_t>(B); r.n_ln_r = rp->alloc<int32_t>(B);
  r.pt_l = rp->alloc<gfpl_keypoint>(np); r.pt_r = rp->alloc<gfpl_keypoint>(np);
  r.pdesc_l = rp->alloc<uint8_t>(np * 32); r.pdesc_r = rp->alloc<uint8_t>(np * 32);
  r.ln_l = rp->alloc<gfpl_keyline>(nl); r.ln_r = rp->alloc<gfpl_keyline>(nl);
  r.ldesc_l = rp->alloc<uint8_t>(nl * 32); r.ldesc_r = rp->alloc<uint8_t>(nl * 32);
  for (DevStereoFrames* f : {&rp->prev, &rp->curr}) {
    const bool is_prev = f == &rp->prev;
    f->B = B; f->cap_pt = is_prev ? prev_cap_pt : cap_pt; f->cap_ln = is_prev ? prev_cap_ln : cap_ln;
    const size_t fp = (size_t)B * f->cap_pt, fl = (size_t)B * f->cap_ln;
    f->n_pt = rp->alloc<int32_t>(B); f->n_ln = rp->alloc<int32_t>(B);
    f->pt = rp->alloc<gfpl_point_feature>(fp); f->pdesc = rp->alloc<uint8_t>(fp * 32);
    f->ln = rp->alloc<gfpl_line_feature>(fl); f->ldesc = rp->alloc<uint8_t>(fl * 32);
  }
  DevTrackObs& o = rp->obs;
  o.B = B; o.cap_pt = ocp; o.cap_ln = ocl;
  o.n_pt = rp->alloc<int32_t>(B); o.n_ln = rp->alloc<int32_t>(B);
  o.pt = rp->alloc<gfpl_pt_obs>((size_t)B * ocp); o.pt_inlier = rp->alloc<uint8_t>((size_t)B * ocp);
  o.ln = rp->alloc<gfpl_ln_obs>((size_t)B * ocl); o.ln_inlier = rp->alloc<uint8_t>((size_t)B * ocl);
  rp->state = rp->alloc<gfpl_frame_state>(B);
  rp->cut = rp->alloc<gfpl_line_cut>((size_t)B * ocl);
  rp->pose = rp->alloc<gfpl_pose_result>(B);
  StepWork& w = rp->w;
  for (Knn2Out* k : {&w.pt_lr, &w.pt_rl, &w.pt_12, &w.pt_21}) { k->idx = rp->alloc<int2>(np); k->dist = rp->alloc<float2>(np); }
  for (Knn2Out* k : {&w.ln_lr, &w.ln_rl, &w.ln_12, &w.ln_21}) { k->idx = rp->alloc<int2>(nl); k->dist = rp->alloc<float2>(nl); }
  w.best_key = rp->alloc<uint32_t>(np);
  w.prev_cov = rp->alloc<gfpl_line_cov>(nl);
  w.f2f_T = rp->alloc<double>((size_t)B * 32);
  rp->pose_raw = rp->alloc<uint8_t>(pose_raw_bytes(B));
  bool ok = true;
  for (void* p : rp->owned) ok = ok && p;
  ok = ok && rp->owned.size() == 12 + 12 + 6 + 3 + 16 + 3 + 1;
  if (!ok || !w.prev_cov) {
    gfpl_replay_destroy(rp);
    set_last_error(ctx, "device allocation failed for a replay of %d pairs", B);
    return GFPL_E_NOMEM;
  }
  cudaMemset(rp->cut, 0, (size_t)B * ocl * sizeof(gfpl_line_cut));
  for (auto& e : rp->ev) cudaEventCreate(&e);
  *out = rp;
  return GFPL_OK;
}

void gfpl_replay_destroy(gfpl_replay* rp) {
  if (!rp) return;
  cudaSetDevice(rp->ctx->device);
  cudaStreamSynchronize(rp->ctx->stream);
  if (rp->s_match) cudaStreamSynchronize(rp->s_match);
  for (auto& st : rp->s_solve)
    if (st) { cudaStreamSynchronize(st); cudaStreamDestroy(st); }
  if (rp->s_match) cudaStreamDestroy(rp->s_match);
  for (void* p : rp->owned) cudaFree(p);
  for (auto& e : rp->ev)
    if (e) cudaEventDestroy(e);
  for (auto& e : rp->ev_match)
    if (e) cudaEventDestroy(e);
  for (auto& e : rp->ev_solve)
    if (e) cudaEventDestroy(e);
  if (rp->ev_input) cudaEventDestroy(rp->ev_input);
  delete rp;
}

#define H2D(dst, src, n)                                                                                   \
  do {                                                                                                     \
    size_t _b = (size_t)(n) * sizeof(*(src));                                                              \
    GFPL_CUDA_OK(ctx, cudaMemcpyAsync((void*)(dst), (src), _b, cudaMemcpyHostToDevice, ctx->stream));      \
    rp->h2d += (int64_t)_b;                                                                                \
  } while (0)
#define D2H(dst, src, n)                                                                                   \
  do {                                                                                                     \
    size_t _b = (size_t)(n) * sizeof(*(src));                                                              \
    GFPL_CUDA_OK(ctx, cudaMemcpyAsync((dst), (src), _b, cudaMemcpyDeviceToHost, ctx->stream));             \
    rp->d2h += (int64_t)_b;                                                                                \
  } while (0)

int gfpl_replay_upload(gfpl_replay* rp, const gfpl_stereo_frames* prev, const gfpl_frame_state* state,
                       const gfpl_raw_frames* raw) {
  if (!rp) return GFPL_E_INVALID;
  gfpl_ctx* ctx = rp->ctx;
  REQUIRE(ctx, prev && state && raw);
  REQUIRE(ctx, prev->B == rp->B && raw->B == rp->B);
  REQUIRE(ctx, prev->cap_pt == rp->prev.cap_pt && prev->cap_ln == rp->prev.cap_ln);
  REQUIRE(ctx, raw->cap_pt == rp->cap_pt && raw->cap_ln == rp->cap_ln);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  rp->h2d = 0;
  rp->input_dirty = true;
  size_t B = rp->B, np = B * rp->cap_pt, nl = B * rp->cap_ln;
  H2D(rp->raw.n_pt_l, raw->n_pt_l, B); H2D(rp->raw.n_pt_r, raw->n_pt_r, B);
  H2D(rp->raw.n_ln_l, raw->n_ln_l, B); H2D(rp->raw.n_ln_r, raw->n_ln_r, B);
  H2D(rp->raw.pt_l, raw->pt_l, np); H2D(rp->raw.pt_r, raw->pt_r, np);
  H2D(rp->raw.pdesc_l, raw->pdesc_l, np * 32); H2D(rp->raw.pdesc_r, raw->pdesc_r, np * 32);
  H2D(rp->raw.ln_l, raw->ln_l, nl); H2D(rp->raw.ln_r, raw->ln_r, nl);
  H2D(rp->raw.ldesc_l, raw->ldesc_l, nl * 32); H2D(rp->raw.ldesc_r, raw->ldesc_r, nl * 32);
  const size_t pp = B * rp->prev.cap_pt, pl = B * rp->prev.cap_ln;
  H2D(rp->prev.n_pt, prev->n_pt, B); H2D(rp->prev.n_ln, prev->n_ln, B);
  H2D(rp->prev.pt, prev->pt, pp); H2D(rp->prev.pdesc, prev->pdesc, pp * 32);
  H2D(rp->prev.ln, prev->ln, pl); H2D(rp->prev.ldesc, prev->ldesc, pl * 32);
  H2D(rp->state, state, B);
  return GFPL_OK;
}

static int upload_raw_and_state(gfpl_replay* rp, const gfpl_frame_state* state, const gfpl_raw_frames* raw) {
  gfpl_ctx* ctx = rp->ctx;
  size_t B = rp->B, np = B * rp->cap_pt, nl = B * rp->cap_ln;
  H2D(rp->raw.n_pt_l, raw->n_pt_l, B); H2D(rp->raw.n_pt_r, raw->n_pt_r, B);
  H2D(rp->raw.n_ln_l, raw->n_ln_l, B); H2D(rp->raw.n_ln_r, raw->n_ln_r, B);
  H2D(rp->raw.pt_l, raw->pt_l, np); H2D(rp->raw.pt_r, raw->pt_r, np);
  H2D(rp->raw.pdesc_l, raw->pdesc_l, np * 32); H2D(rp->raw.pdesc_r, raw->pdesc_r, np * 32);
  H2D(rp->raw.ln_l, raw->ln_l, nl); H2D(rp->raw.ln_r, raw->ln_r, nl);
  H2D(rp->raw.ldesc_l, raw->ldesc_l, nl * 32); H2D(rp->raw.ldesc_r, raw->ldesc_r, nl * 32);
  H2D(rp->state, state, B);
  return GFPL_OK;
}

int gfpl_replay_upload_slim(gfpl_replay* rp, const gfpl_prev_frames* prev, const gfpl_frame_state* state,
                            const gfpl_raw_frames* raw) {
  if (!rp) return GFPL_E_INVALID;
  gfpl_ctx* ctx = rp->ctx;
  REQUIRE(ctx, prev && state && raw);
  REQUIRE(ctx, prev->B == rp->B && raw->B == rp->B);
  REQUIRE(ctx, prev->cap_pt == rp->prev.cap_pt && prev->cap_ln == rp->prev.cap_ln);
  REQUIRE(ctx, raw->cap_pt == rp->cap_pt && raw->cap_ln == rp->cap_ln);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  const size_t B = rp->B, pp = B * rp->prev.cap_pt, pl = B * rp->prev.cap_ln;
  if (!rp->slim.pt) {   // staging, allocated on first use; published only when every allocation succeeded
    gfpl_prev_frames st{};
    st.B = rp->B; st.cap_pt = rp->prev.cap_pt; st.cap_ln = rp->prev.cap_ln;
    st.n_pt = rp->alloc<int32_t>(B); st.n_ln = rp->alloc<int32_t>(B);
    st.pt = rp->alloc<gfpl_prev_point>(pp); st.ln = rp->alloc<gfpl_prev_line>(pl);
    REQUIRE_MEM(ctx, st.n_pt && st.n_ln && st.pt && st.ln);
    rp->slim = st;
  }
  rp->h2d = 0;
  rp->input_dirty = true;
  H2D(rp->slim.n_pt, prev->n_pt, B); H2D(rp->slim.n_ln, prev->n_ln, B);
  H2D(rp->slim.pt, prev->pt, pp); H2D(rp->slim.ln, prev->ln, pl);
  H2D(rp->prev.pdesc, prev->pdesc, pp * 32); H2D(rp->prev.ldesc, prev->ldesc, pl * 32);
  CHECK(upload_raw_and_state(rp, state, raw));
  return launch_expand_prev(ctx, rp->slim, rp->prev);
}

// device staging of the compact records (descriptors and counts go straight to their final buffers).  Pointers are
// published only when every allocation of a group succeeded, so a failed call can be retried.
static int ensure_compact_staging(gfpl_replay* rp, bool with_prev) {
  gfpl_ctx* ctx = rp->ctx;
  const size_t B = rp->B, np = B * rp->cap_pt, nl = B * rp->cap_ln, pp = B * rp->prev.cap_pt, pl = B * rp->prev.cap_ln;
  gfpl_compact_input& d = rp->compact;
  if (!d.pt_l) {
    gfpl_compact_input t{};
    t.pt_l = rp->alloc<gfpl_keypoint8>(np); t.pt_r = rp->alloc<gfpl_keypoint8>(np);
    t.pt_oct_l = rp->alloc<uint8_t>(np); t.pt_oct_r = rp->alloc<uint8_t>(np);
    t.ln_l = rp->alloc<gfpl_keyline20>(nl); t.ln_r = rp->alloc<gfpl_keyline20>(nl);
    t.ln_oct_l = rp->alloc<uint8_t>(nl); t.ln_oct_r = rp->alloc<uint8_t>(nl);
    REQUIRE_MEM(ctx, t.pt_l && t.pt_r && t.pt_oct_l && t.pt_oct_r && t.ln_l && t.ln_r && t.ln_oct_l && t.ln_oct_r);
    d.pt_l = t.pt_l; d.pt_r = t.pt_r; d.pt_oct_l = t.pt_oct_l; d.pt_oct_r = t.pt_oct_r;
    d.ln_l = t.ln_l; d.ln_r = t.ln_r; d.ln_oct_l = t.ln_oct_l; d.ln_oct_r = t.ln_oct_r;
  }
  if (with_prev && !d.prev_pt) {
    const gfpl_prev_point12* a = rp->alloc<gfpl_prev_point12>(pp);
    const uint8_t* b = rp->alloc<uint8_t>(pp);
    const gfpl_prev_line32* c = rp->alloc<gfpl_prev_line32>(pl);
    const uint8_t* e = rp->alloc<uint8_t>(pl);
    REQUIRE_MEM(ctx, a && b && c && e);
    d.prev_pt = a; d.prev_pt_level = b; d.prev_ln = c; d.prev_ln_level = e;
  }
  d.B = rp->B; d.cap_pt = rp->cap_pt; d.cap_ln = rp->cap_ln; d.prev_cap_pt = rp->prev.cap_pt; d.prev_cap_ln = rp->prev.cap_ln;
  d.n_pt_l = rp->raw.n_pt_l; d.n_pt_r = rp->raw.n_pt_r; d.n_ln_l = rp->raw.n_ln_l; d.n_ln_r = rp->raw.n_ln_r;
  d.prev_n_pt = rp->prev.n_pt; d.prev_n_ln = rp->prev.n_ln;
  d.pdesc_l = d.pdesc_r = d.ldesc_l = d.ldesc_r = d.prev_pdesc = d.prev_ldesc = nullptr;  // not staged
  return GFPL_OK;
}

// current-frame part of a compact step input -> the raw buffers (counts, compact key points / key lines, descriptors)
static int upload_compact_curr(gfpl_replay* rp, const gfpl_compact_input* in) {
  gfpl_ctx* ctx = rp->ctx;
  const size_t B = rp->B, np = B * rp->cap_pt, nl = B * rp->cap_ln;
  gfpl_compact_input& d = rp->compact;
  REQUIRE(ctx, counts_ok(in->n_pt_l, B, rp->cap_pt) && counts_ok(in->n_pt_r, B, rp->cap_pt) && counts_ok(in->n_ln_l, B, rp->cap_ln) &&
                   counts_ok(in->n_ln_r, B, rp->cap_ln));
  H2D(rp->raw.n_pt_l, in->n_pt_l, B); H2D(rp->raw.n_pt_r, in->n_pt_r, B);
  H2D(rp->raw.n_ln_l, in->n_ln_l, B); H2D(rp->raw.n_ln_r, in->n_ln_r, B);
  H2D(const_cast<gfpl_keypoint8*>(d.pt_l), in->pt_l, np); H2D(const_cast<gfpl_keypoint8*>(d.pt_r), in->pt_r, np);
  H2D(const_cast<uint8_t*>(d.pt_oct_l), in->pt_oct_l, np); H2D(const_cast<uint8_t*>(d.pt_oct_r), in->pt_oct_r, np);
  H2D(const_cast<gfpl_keyline20*>(d.ln_l), in->ln_l, nl); H2D(const_cast<gfpl_keyline20*>(d.ln_r), in->ln_r, nl);
  H2D(const_cast<uint8_t*>(d.ln_oct_l), in->ln_oct_l, nl); H2D(const_cast<uint8_t*>(d.ln_oct_r), in->ln_oct_r, nl);
  H2D(rp->raw.pdesc_l, in->pdesc_l, np * 32); H2D(rp->raw.pdesc_r, in->pdesc_r, np * 32);
  H2D(rp->raw.ldesc_l, in->ldesc_l, nl * 32); H2D(rp->raw.ldesc_r, in->ldesc_r, nl * 32);
  return GFPL_OK;
}

int gfpl_replay_upload_compact(gfpl_replay* rp, const gfpl_compact_input* in, const gfpl_frame_state* state) {
  if (!rp) return GFPL_E_INVALID;
  gfpl_ctx* ctx = rp->ctx;
  REQUIRE(ctx, in && state && in->B == rp->B);
  REQUIRE(ctx, in->cap_pt == rp->cap_pt && in->cap_ln == rp->cap_ln);
  REQUIRE(ctx, in->prev_cap_pt == rp->prev.cap_pt && in->prev_cap_ln == rp->prev.cap_ln);
  REQUIRE(ctx, counts_ok(in->prev_n_pt, rp->B, rp->prev.cap_pt) && counts_ok(in->prev_n_ln, rp->B, rp->prev.cap_ln));
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  const size_t B = rp->B, pp = B * rp->prev.cap_pt, pl = B * rp->prev.cap_ln;
  CHECK(ensure_compact_staging(rp, true));
  gfpl_compact_input& d = rp->compact;
  rp->h2d = 0;
  rp->input_dirty = true;
  CHECK(upload_compact_curr(rp, in));
  H2D(rp->prev.n_pt, in->prev_n_pt, B); H2D(rp->prev.n_ln, in->prev_n_ln, B);
  H2D(const_cast<gfpl_prev_point12*>(d.prev_pt), in->prev_pt, pp); H2D(const_cast<uint8_t*>(d.prev_pt_level), in->prev_pt_level, pp);
  H2D(const_cast<gfpl_prev_line32*>(d.prev_ln), in->prev_ln, pl); H2D(const_cast<uint8_t*>(d.prev_ln_level), in->prev_ln_level, pl);
  H2D(rp->prev.pdesc, in->prev_pdesc, pp * 32); H2D(rp->prev.ldesc, in->prev_ldesc, pl * 32);
  H2D(rp->state, state, B);
  return launch_expand_compact(ctx, d, rp->raw, rp->prev, /*with_prev=*/true);
}

// ---- sequence-chained replay: StereoFrameHandler::updateFrame (stereoFrameHandler.cpp:926-961) ----------------------------
int gfpl_replay_upload_curr_compact(gfpl_replay* rp, const gfpl_compact_input* in) {
  if (!rp) return GFPL_E_INVALID;
  gfpl_ctx* ctx = rp->ctx;
  REQUIRE(ctx, in && in->B == rp->B && in->cap_pt == rp->cap_pt && in->cap_ln == rp->cap_ln);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  CHECK(ensure_compact_staging(rp, false));
  rp->h2d = 0;
  rp->input_dirty = true;
  CHECK(upload_compact_curr(rp, in));
  return launch_expand_compact(ctx, rp->compact, rp->raw, rp->prev, /*with_prev=*/false);
}

int gfpl_replay_advance(gfpl_replay* rp, const double* time_stamp_next) {
  if (!rp) return GFPL_E_INVALID;
  gfpl_ctx* ctx = rp->ctx;
  // prev_frame = curr_frame: the two frame sets trade places, which needs equal strides
  REQUIRE(ctx, rp->prev.cap_pt == rp->curr.cap_pt && rp->prev.cap_ln == rp->curr.cap_ln);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  double* dts = nullptr;
  if (time_stamp_next) {
    if (!rp->stamps) rp->stamps = rp->alloc<double>(rp->B);
    REQUIRE_MEM(ctx, rp->stamps);
    dts = rp->stamps;
    H2D(dts, time_stamp_next, (size_t)rp->B);
  }
  rp->input_dirty = true;
  std::swap(rp->prev, rp->curr);
  return launch_advance_state(ctx, rp->B, rp->pose, dts, rp->state);
}

int gfpl_replay_set_pipeline(gfpl_replay* rp, int32_t n_sub) {
  if (!rp) return GFPL_E_INVALID;
  gfpl_ctx* ctx = rp->ctx;
  REQUIRE(ctx, n_sub >= 1 && n_sub <= GFPL_MAX_SUBBATCHES);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  if (n_sub > 1) {
    int lo_prio = 0, hi_prio = 0;   // numerically lower = higher priority
    GFPL_CUDA_OK(ctx, cudaDeviceGetStreamPriorityRange(&lo_prio, &hi_prio));
    if (!rp->s_match) GFPL_CUDA_OK(ctx, cudaStreamCreateWithPriority(&rp->s_match, cudaStreamNonBlocking, lo_prio));
    if (!rp->ev_input) GFPL_CUDA_OK(ctx, cudaEventCreateWithFlags(&rp->ev_input, cudaEventDisableTiming));
    for (int k = 0; k < n_sub; k++) {
      if (!rp->s_solve[k]) GFPL_CUDA_OK(ctx, cudaStreamCreateWithPriority(&rp->s_solve[k], cudaStreamNonBlocking, hi_prio));
      if (!rp->ev_match[k]) GFPL_CUDA_OK(ctx, cudaEventCreateWithFlags(&rp->ev_match[k], cudaEventDisableTiming));
      if (!rp->ev_solve[k]) GFPL_CUDA_OK(ctx, cudaEventCreateWithFlags(&rp->ev_solve[k], cudaEventDisableTiming));
    }
  }
  // changing the split: drain what is in flight so that no sub-batch of the old split is still being written
  if (rp->s_match) GFPL_CUDA_OK(ctx, cudaStreamSynchronize(rp->s_match));
  for (int k = 0; k < GFPL_MAX_SUBBATCHES; k++) {
    if (rp->s_solve[k]) GFPL_CUDA_OK(ctx, cudaStreamSynchronize(rp->s_solve[k]));
    rp->solve_pending[k] = false;
  }
  rp->n_sub = n_sub;
  rp->input_dirty = true;
  return GFPL_OK;
}

// RAII: the launchers enqueue on ctx->stream; the pipeline points it at the stream of the stage being enqueued
namespace {
struct StreamSwap {
  gfpl_ctx* ctx;
  cudaStream_t saved;
  explicit StreamSwap(gfpl_ctx* c) : ctx(c), saved(c->stream) {}
  ~StreamSwap() { ctx->stream = saved; }
};
}  // namespace

static int replay_run_pipelined(gfpl_replay* rp, int32_t flags) {
  gfpl_ctx* ctx = rp->ctx;
  const int S = rp->n_sub, B = rp->B;
  const int wcp = rp->cap_pt, wcl = rp->cap_ln;   // strides of the scratch arrays
  cudaStream_t main = ctx->stream;
  // What the caller ordered on the context stream before this call precedes the step: uploads and downloads (they
  // touch what the matching reads and writes) hold back the whole step; anything else — e.g. an NCCL gather of the
  // previous step's pose records enqueued behind gfpl_ctx_stream — only the optimiser stages, which own the records.
  GFPL_CUDA_OK(ctx, cudaEventRecord(rp->ev_input, main));
  if (rp->input_dirty) {
    GFPL_CUDA_OK(ctx, cudaStreamWaitEvent(rp->s_match, rp->ev_input, 0));
    rp->input_dirty = false;
  }
  StreamSwap swap(ctx);
  int last = -1;
  for (int k = 0; k < S; k++) {
    const int lo = (int)((int64_t)B * k / S), hi = (int)((int64_t)B * (k + 1) / S), n = hi - lo;
    if (n <= 0) continue;
    const DevRawFrames raw = slice(rp->raw, lo, n);
    const DevStereoFrames prev = slice(rp->prev, lo, n), curr = slice(rp->curr, lo, n);
    const DevTrackObs obs = slice(rp->obs, lo, n);
    StepWork w = slice(rp->w, lo, wcp, wcl);
    // the matching of this sub-batch rewrites obs / scratch that the previous run's optimiser of the same sub-batch reads
    if (rp->solve_pending[k]) GFPL_CUDA_OK(ctx, cudaStreamWaitEvent(rp->s_match, rp->ev_solve[k], 0));
    ctx->stream = rp->s_match;
    CHECK(run_match(ctx, prev, rp->state + lo, raw, flags, curr, obs, w, nullptr));
    GFPL_CUDA_OK(ctx, cudaEventRecord(rp->ev_match[k], rp->s_match));
    GFPL_CUDA_OK(ctx, cudaStreamWaitEvent(rp->s_solve[k], rp->ev_match[k], 0));
    GFPL_CUDA_OK(ctx, cudaStreamWaitEvent(rp->s_solve[k], rp->ev_input, 0));
    ctx->stream = rp->s_solve[k];
    CHECK(run_solve(ctx, prev, rp->state + lo, flags, obs, rp->cut + (size_t)lo * rp->obs.cap_ln, rp->pose + lo, w, nullptr,
                    (uint8_t*)rp->pose_raw + (size_t)lo * pose_raw_bytes(1)));
    GFPL_CUDA_OK(ctx, cudaEventRecord(rp->ev_solve[k], rp->s_solve[k]));
    rp->solve_pending[k] = true;
    last = k;
  }
  ctx->stream = main;
  // everything the caller enqueues on the context stream after this call sees the finished step
  for (int k = 0; k <= last; k++)
    if (rp->solve_pending[k]) GFPL_CUDA_OK(ctx, cudaStreamWaitEvent(main, rp->ev_solve[k], 0));
  return GFPL_OK;
}

int gfpl_replay_run(gfpl_replay* rp, int32_t flags) {
  if (!rp) return GFPL_E_INVALID;
  gfpl_ctx* ctx = rp->ctx;
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  if (rp->n_sub > 1) {
    rp->timed = false;
    return replay_run_pipelined(rp, flags);
  }
  rp->timed = true;
  return run_step(ctx, rp->prev, rp->state, rp->raw, flags, rp->curr, rp->obs, rp->cut, rp->pose, rp->w, rp->ev, rp->pose_raw);
}

int gfpl_replay_download(gfpl_replay* rp, gfpl_pose_result* pose, gfpl_track_obs* obs, gfpl_line_cut* cut,
                         gfpl_stereo_frames* curr) {
  if (!rp) return GFPL_E_INVALID;
  gfpl_ctx* ctx = rp->ctx;
  REQUIRE(ctx, pose);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  rp->d2h = 0;
  rp->input_dirty = true;   // the next pipelined run must not overwrite what these copies are still reading
  size_t B = rp->B;
  D2H(pose, rp->pose, B);
  if (obs) {
    REQUIRE(ctx, obs->B == rp->B && obs->cap_pt == rp->obs.cap_pt && obs->cap_ln == rp->obs.cap_ln);
    D2H(obs->n_pt, rp->obs.n_pt, B); D2H(obs->n_ln, rp->obs.n_ln, B);
    D2H(obs->pt, rp->obs.pt, B * rp->obs.cap_pt); D2H(obs->pt_inlier, rp->obs.pt_inlier, B * rp->obs.cap_pt);
    D2H(obs->ln, rp->obs.ln, B * rp->obs.cap_ln); D2H(obs->ln_inlier, rp->obs.ln_inlier, B * rp->obs.cap_ln);
  }
  if (cut) D2H(cut, rp->cut, B * rp->obs.cap_ln);
  if (curr) {
    REQUIRE(ctx, curr->B == rp->B && curr->cap_pt == rp->cap_pt && curr->cap_ln == rp->cap_ln);
    D2H(curr->n_pt, rp->curr.n_pt, B); D2H(curr->n_ln, rp->curr.n_ln, B);
    D2H(curr->pt, rp->curr.pt, B * rp->cap_pt); D2H(curr->pdesc, rp->curr.pdesc, B * rp->cap_pt * 32);
    D2H(curr->ln, rp->curr.ln, B * rp->cap_ln); D2H(curr->ldesc, rp->curr.ldesc, B * rp->cap_ln * 32);
  }
  return GFPL_OK;
}

int64_t gfpl_replay_h2d_bytes(const gfpl_replay* rp) { return rp ? rp->h2d : 0; }
int64_t gfpl_replay_d2h_bytes(const gfpl_replay* rp) { return rp ? rp->d2h : 0; }
void* gfpl_replay_pose_device_ptr(gfpl_replay* rp) { return rp ? (void*)rp->pose : nullptr; }

int gfpl_replay_stage_ms(gfpl_replay* rp, float* ms) {
  if (!rp || !ms) return GFPL_E_INVALID;
  gfpl_ctx* ctx = rp->ctx;
  REQUIRE(ctx, rp->timed);
  GFPL_CUDA_OK(ctx, cudaEventSynchronize(rp->ev[8]));
  for (int i = 0; i < 8; i++) GFPL_CUDA_OK(ctx, cudaEventElapsedTime(&ms[i], rp->ev[i], rp->ev[i + 1]));
  GFPL_CUDA_OK(ctx, cudaEventElapsedTime(&ms[8], rp->ev[0], rp->ev[8]));
  return GFPL_OK;
}

int gfpl_replay_pair_counts(gfpl_replay* rp, int64_t* pairs) {
  if (!rp || !pairs) return GFPL_E_INVALID;
  gfpl_ctx* ctx = rp->ctx;
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  const int B = rp->B;
  std::vector<int32_t> n(8 * (size_t)B);
  const int32_t* src[8] = {rp->raw.n_pt_l, rp->raw.n_pt_r, rp->raw.n_ln_l, rp->raw.n_ln_r,
                           rp->prev.n_pt, rp->curr.n_pt, rp->prev.n_ln, rp->curr.n_ln};
  for (int k = 0; k < 8; k++)
    GFPL_CUDA_OK(ctx, cudaMemcpyAsync(n.data() + (size_t)k * B, src[k], (size_t)B * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CHECK(finish(ctx));
  // distinct 256-bit distances: with best_lr_matches both directions are served by ONE pass over the nq x nt matrix
  // (knn2_sym_kernel), so a distance is evaluated once (the reference evaluates it twice, once per knnMatch call)
  pairs[0] = pairs[1] = pairs[2] = pairs[3] = 0;
  for (int p = 0; p < B; p++) {
    pairs[0] += (int64_t)n[p] * n[B + p];
    pairs[1] += (int64_t)n[2 * B + p] * n[3 * B + p];
    pairs[2] += (int64_t)n[4 * B + p] * n[5 * B + p];  // hybrid fused search: one direction
    pairs[3] += (int64_t)n[6 * B + p] * n[7 * B + p];
  }
  return GFPL_OK;
}

int gfpl_timer_record(gfpl_ctx* ctx, int32_t slot) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, slot >= 0 && slot < 8);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  if (!ctx->timer[slot]) GFPL_CUDA_OK(ctx, cudaEventCreate(&ctx->timer[slot]));
  GFPL_CUDA_OK(ctx, cudaEventRecord(ctx->timer[slot], ctx->stream));
  return GFPL_OK;
}

int gfpl_timer_elapsed_ms(gfpl_ctx* ctx, int32_t a, int32_t b, float* ms) {
  if (!ctx || !ms) return GFPL_E_INVALID;
  REQUIRE(ctx, a >= 0 && a < 8 && b >= 0 && b < 8 && ctx->timer[a] && ctx->timer[b]);
  GFPL_CUDA_OK(ctx, cudaEventSynchronize(ctx->timer[b]));
  GFPL_CUDA_OK(ctx, cudaEventElapsedTime(ms, ctx->timer[a], ctx->timer[b]));
  return GFPL_OK;
}

// ---- L2 / §8f-3: local bundle adjustment, batched over windows ------------------------------------------------------------
}  // extern "C"

namespace {

// CSR offsets of an observation list sorted by landmark (the order localBundleAdjustment produces, :1146-1209)
bool landmark_offsets(const int32_t* lm, int n_obs, int n_lm, int32_t* off) {
  for (int j = 0; j <= n_lm; j++) off[j] = 0;
  int prev = 0;
  for (int o = 0; o < n_obs; o++) {
    if (lm[o] < prev || lm[o] >= n_lm) return false;
    prev = lm[o];
    off[(size_t)lm[o] + 1]++;
  }
  for (int j = 0; j < n_lm; j++) off[(size_t)j + 1] += off[j];
  return true;
}

// All windows of a batch in ONE device buffer: [ inputs (one H2D copy) | block storage (one memset) | solve scratch | DX |
// lambda_eff | window descriptors ].  `host` is the packed image of the input region.
struct LbaBatch {
  int W = 0, max_kf = 0, max_pt = 0, max_ls = 0, max_pt_obs = 0, max_ls_obs = 0;
  uint8_t* host = nullptr;           // packed image of the input region (pinned staging of the context)
  std::vector<LbaWin> wins;          // device pointers
  std::vector<size_t> dx_off;        // [W + 1] offsets (in doubles) of the windows' DX in the concatenated output
  size_t in_bytes = 0, zero_off = 0, zero_bytes = 0, err_off = 0, dx_base = 0, lam_off = 0, wins_off = 0, total = 0;
  uint8_t* dev = nullptr;

  static size_t al(size_t x) { return (x + 255) & ~(size_t)255; }

  // pass 1 (dev == nullptr): sizes; pass 2: fill `host` and the device pointers
  template <typename T>
  const T* put(size_t& cur, const T* src, size_t n) {
    const size_t off = cur;
    cur = al(cur + n * sizeof(T));
    if (dev) {
      if (n) memcpy(host + off, src, n * sizeof(T));
      return (const T*)(dev + off);
    }
    return nullptr;
  }
  template <typename T>
  T* room(size_t& cur, size_t n) {
    const size_t off = cur;
    cur = al(cur + n * sizeof(T));
    return dev ? (T*)(dev + off) : nullptr;
  }

  // inputs of window w from offset `cur` on (pass 2: copied into the staging image)
  bool put_inputs(const gfpl_lba_problem* probs, bool with_solve, int w, size_t& cur) {
    {
      const gfpl_lba_problem& pb = probs[w];
      LbaWin L{};
      L.pb = pb;
      L.pb.kf_Tkw = put(cur, pb.kf_Tkw, (size_t)pb.Nkf * 16);
      L.pb.X_pt = put(cur, pb.X_pt, (size_t)pb.Npt * 3);
      L.pb.X_ls = put(cur, pb.X_ls, (size_t)pb.Nls * 6);
      L.pb.pt_obs_lm = put(cur, pb.pt_obs_lm, pb.n_pt_obs); L.pb.pt_obs_kf = put(cur, pb.pt_obs_kf, pb.n_pt_obs);
      L.pb.pt_obs_fix = put(cur, pb.pt_obs_fix, pb.n_pt_obs);
      L.pb.pt_obs_uv = put(cur, pb.pt_obs_uv, (size_t)pb.n_pt_obs * 2); L.pb.pt_obs_s2 = put(cur, pb.pt_obs_s2, pb.n_pt_obs);
      L.pb.ls_obs_lm = put(cur, pb.ls_obs_lm, pb.n_ls_obs); L.pb.ls_obs_kf = put(cur, pb.ls_obs_kf, pb.n_ls_obs);
      L.pb.ls_obs_fix = put(cur, pb.ls_obs_fix, pb.n_ls_obs);
      L.pb.ls_obs_le = put(cur, pb.ls_obs_le, (size_t)pb.n_ls_obs * 3); L.pb.ls_obs_s2 = put(cur, pb.ls_obs_s2, pb.n_ls_obs);
      L.pb.fix_Tkw = put(cur, pb.fix_Tkw, (size_t)pb.n_fix * 16);
      if (with_solve) {
        const size_t o_pt = cur;
        L.pt_off = room<int32_t>(cur, (size_t)pb.Npt + 1);
        const size_t o_ls = cur;
        L.ls_off = room<int32_t>(cur, (size_t)pb.Nls + 1);
        if (dev && (!landmark_offsets(pb.pt_obs_lm, pb.n_pt_obs, pb.Npt, (int32_t*)(host + o_pt)) ||
                    !landmark_offsets(pb.ls_obs_lm, pb.n_ls_obs, pb.Nls, (int32_t*)(host + o_ls))))
          return false;
      }
      if (dev) wins[w] = L;
    }
    return true;
  }

  std::vector<size_t> in_off;        // [W + 1] offsets of the windows' input regions (pass 1)

  bool layout(const gfpl_lba_problem* probs, bool with_solve) {
    size_t cur = 0;
    if (!dev) {                      // pass 1: sizes
      in_off.assign((size_t)W + 1, 0);
      for (int w = 0; w < W; w++) { in_off[w] = cur; put_inputs(probs, with_solve, w, cur); }
      in_off[W] = cur;
    } else {
      // pass 2: the packing is a host memcpy of every window's arrays (~230 KB per KITTI-shaped window) into the pinned
      // image — split over a few threads for large batches (the windows' regions are disjoint)
      const int nthr = W >= 64 ? (int)std::min<unsigned>(8u, std::max(1u, std::thread::hardware_concurrency())) : 1;
      std::vector<char> ok(nthr, 1);
      auto work = [&](int t) {
        for (int w = (int)((long long)W * t / nthr); w < (int)((long long)W * (t + 1) / nthr); w++) {
          size_t c = in_off[w];
          if (!put_inputs(probs, with_solve, w, c)) ok[t] = 0;
        }
      };
      if (nthr == 1) work(0);
      else {
        std::vector<std::thread> th;
        for (int t = 1; t < nthr; t++) th.emplace_back(work, t);
        work(0);
        for (auto& x : th) x.join();
      }
      for (char c : ok)
        if (!c) return false;
      cur = in_off[W];
    }
    in_bytes = cur;
    zero_off = cur;
    err_off = cur;
    double* err_all = room<double>(cur, W);   // ---- block storage (zeroed): err of every window first, contiguous
    for (int w = 0; w < W; w++) {
      const gfpl_lba_problem& pb = probs[w];
      const size_t N = 6 * (size_t)pb.Nkf + 3 * (size_t)pb.Npt + 6 * (size_t)pb.Nls;
      gfpl_lba_blocks H{};
      H.Hkk = room<double>(cur, (size_t)pb.Nkf * 36); H.Hpp = room<double>(cur, (size_t)pb.Npt * 9);
      H.Hll = room<double>(cur, (size_t)pb.Nls * 36); H.Hpk = room<double>(cur, (size_t)pb.n_pt_obs * 18);
      H.Hlk = room<double>(cur, (size_t)pb.n_ls_obs * 36); H.g = room<double>(cur, N);
      H.err = dev ? err_all + w : nullptr;
      if (dev) wins[w].H = H;
    }
    zero_bytes = cur - zero_off;
    if (with_solve) {
      for (int w = 0; w < W; w++) {   // ---- solve scratch
        const gfpl_lba_problem& pb = probs[w];
        const size_t nk = 6 * (size_t)pb.Nkf;
        LbaSolveWork k{};
        k.S = room<double>(cur, nk * nk);
        k.Ypt = room<double>(cur, (size_t)pb.n_pt_obs * 18); k.Zpt = room<double>(cur, (size_t)pb.Npt * 3);
        k.Yls = room<double>(cur, (size_t)pb.n_ls_obs * 36); k.Zls = room<double>(cur, (size_t)pb.Nls * 6);
        if (dev) wins[w].w = k;
      }
      dx_base = cur;               // ---- DX of every window, concatenated (one D2H copy)
      double* dx_all = room<double>(cur, dx_off[W]);
      lam_off = cur;
      double* lam_all = room<double>(cur, W);
      if (dev)
        for (int w = 0; w < W; w++) { wins[w].DX = dx_all + dx_off[w]; wins[w].lambda_eff = lam_all + w; }
    }
    wins_off = cur;
    room<LbaWin>(cur, W);
    total = cur;
    return true;
  }
};

// validate, lay out, upload and zero; on success b.dev holds everything and d_wins the device descriptor array
int lba_batch_begin(gfpl_ctx* ctx, DevMem& m, int W, const gfpl_lba_problem* probs, bool with_solve, LbaBatch& b, const LbaWin** d_wins) {
  b.W = W;
  b.dx_off.assign((size_t)W + 1, 0);
  for (int w = 0; w < W; w++) {
    const gfpl_lba_problem& pb = probs[w];
    REQUIRE(ctx, pb.Nkf >= 0 && pb.Npt >= 0 && pb.Nls >= 0 && pb.n_pt_obs >= 0 && pb.n_ls_obs >= 0 && pb.n_fix >= 0);
    b.dx_off[(size_t)w + 1] = b.dx_off[w] + 6 * (size_t)pb.Nkf + 3 * (size_t)pb.Npt + 6 * (size_t)pb.Nls;
    b.max_kf = pb.Nkf > b.max_kf ? pb.Nkf : b.max_kf;
    b.max_pt = pb.Npt > b.max_pt ? pb.Npt : b.max_pt;
    b.max_ls = pb.Nls > b.max_ls ? pb.Nls : b.max_ls;
    b.max_pt_obs = pb.n_pt_obs > b.max_pt_obs ? pb.n_pt_obs : b.max_pt_obs;
    b.max_ls_obs = pb.n_ls_obs > b.max_ls_obs ? pb.n_ls_obs : b.max_ls_obs;
  }
  b.layout(probs, with_solve);                       // sizes
  b.dev = m.alloc<uint8_t>(b.total);
  REQUIRE_MEM(ctx, b.dev);
  // the previous call's H2D copy out of the staging buffer has completed: every entry point ends with a synchronisation
  if (ctx->staging_bytes < b.in_bytes) {
    if (ctx->staging) cudaFreeHost(ctx->staging);
    ctx->staging = nullptr;
    ctx->staging_bytes = 0;
    const size_t want = b.in_bytes + b.in_bytes / 4;
    if (cudaHostAlloc(&ctx->staging, want, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      set_last_error(ctx, "cannot allocate %zu bytes of pinned staging memory", want);
      return GFPL_E_NOMEM;
    }
    ctx->staging_bytes = want;
  }
  b.host = (uint8_t*)ctx->staging;
  b.wins.resize(W);
  if (!b.layout(probs, with_solve)) {
    set_last_error(ctx, "local BA: observations must be sorted by landmark index and index existing landmarks");
    return GFPL_E_INVALID;
  }
  GFPL_CUDA_OK(ctx, cudaMemcpyAsync(b.dev, b.host, b.in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  GFPL_CUDA_OK(ctx, cudaMemcpyAsync(b.dev + b.wins_off, b.wins.data(), (size_t)W * sizeof(LbaWin), cudaMemcpyHostToDevice, ctx->stream));
  GFPL_CUDA_OK(ctx, cudaMemsetAsync(b.dev + b.zero_off, 0, b.zero_bytes, ctx->stream));
  *d_wins = (const LbaWin*)(b.dev + b.wins_off);
  return GFPL_OK;
}

}  // namespace

extern "C" {

int gfpl_lba_assemble(gfpl_ctx* ctx, const gfpl_lba_problem* pb, gfpl_lba_blocks* out) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, pb && out);
  REQUIRE(ctx, out->Hkk && out->Hpp && out->Hll && out->Hpk && out->Hlk && out->g && out->err);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  LbaBatch b;
  const LbaWin* dw = nullptr;
  CHECK(lba_batch_begin(ctx, m, 1, pb, false, b, &dw));
  const size_t N = 6 * (size_t)pb->Nkf + 3 * (size_t)pb->Npt + 6 * (size_t)pb->Nls;
  CHECK(launch_lba_assemble(ctx, dw, 1, b.max_kf, b.max_pt_obs, b.max_ls_obs));
  const gfpl_lba_blocks& o = b.wins[0].H;
  download(ctx, out->Hkk, o.Hkk, (size_t)pb->Nkf * 36); download(ctx, out->Hpp, o.Hpp, (size_t)pb->Npt * 9);
  download(ctx, out->Hll, o.Hll, (size_t)pb->Nls * 36); download(ctx, out->Hpk, o.Hpk, (size_t)pb->n_pt_obs * 18);
  download(ctx, out->Hlk, o.Hlk, (size_t)pb->n_ls_obs * 36); download(ctx, out->g, o.g, N);
  download(ctx, out->err, o.err, 1);
  return finish(ctx);
}

int gfpl_lba_step_batch(gfpl_ctx* ctx, int32_t W, const gfpl_lba_problem* probs, double lambda, int32_t scale_by_hmax, double* DX,
                        double* err_sum, double* lambda_eff) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, W >= 0);
  if (W == 0) return GFPL_OK;
  REQUIRE(ctx, probs && DX && err_sum && lambda_eff);
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  DevMem m(ctx);
  LbaBatch b;
  const LbaWin* dw = nullptr;
  CHECK(lba_batch_begin(ctx, m, W, probs, true, b, &dw));
  for (int k = 0; k < 2; k++)
    if (!ctx->timer[6 + k]) GFPL_CUDA_OK(ctx, cudaEventCreate(&ctx->timer[6 + k]));
  GFPL_CUDA_OK(ctx, cudaEventRecord(ctx->timer[6], ctx->stream));
  CHECK(launch_lba_assemble(ctx, dw, W, b.max_kf, b.max_pt_obs, b.max_ls_obs));
  CHECK(launch_lba_solve(ctx, dw, W, b.max_kf, b.max_pt, b.max_ls, lambda, scale_by_hmax));
  GFPL_CUDA_OK(ctx, cudaEventRecord(ctx->timer[7], ctx->stream));
  download(ctx, DX, (const double*)(b.dev + b.dx_base), b.dx_off[W]);
  download(ctx, err_sum, (const double*)(b.dev + b.err_off), (size_t)W);
  download(ctx, lambda_eff, (const double*)(b.dev + b.lam_off), (size_t)W);
  CHECK(finish(ctx));
  // a window without unknowns: the reference leaves lambda as it is and there is nothing to solve
  for (int w = 0; w < W; w++)
    if (b.dx_off[(size_t)w + 1] == b.dx_off[w]) { err_sum[w] = 0.0; lambda_eff[w] = lambda; }
  return GFPL_OK;
}

int gfpl_lba_step(gfpl_ctx* ctx, const gfpl_lba_problem* pb, double lambda, int32_t scale_by_hmax, double* DX,
                  double* err_sum, double* lambda_eff) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, pb && DX && err_sum && lambda_eff);
  return gfpl_lba_step_batch(ctx, 1, pb, lambda, scale_by_hmax, DX, err_sum, lambda_eff);
}

}  // extern "C"
