// gfpl_api.cu — the C ABI of include/gfpl_c.h: context, host-pointer entry points (upload -> kernels ->
// download on the context stream) and the resident replay pipeline.
//
// There is no CPU implementation behind any compute entry point: without a CUDA device
// gfpl_ctx_create fails with GFPL_E_NO_DEVICE and nothing else can be called.
#include <stdarg.h>
#include <string.h>

#include <mutex>
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
    if (cudaMallocAsync(&p, n * sizeof(T), ctx->stream) != cudaSuccess) return nullptr;
    ptrs.push_back(p);
    return (T*)p;
  }
  template <typename T>
  T* upload(const T* host, size_t n) {
    T* d = alloc<T>(n);
    if (d && n && host) cudaMemcpyAsync(d, host, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream);
    return d;
  }
};

template <typename T>
static inline void download(gfpl_ctx* ctx, T* host, const T* dev, size_t n) {
  if (host && n) cudaMemcpyAsync(host, dev, n * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream);
}

static int finish(gfpl_ctx* ctx) {
  GFPL_CUDA_OK(ctx, cudaStreamSynchronize(ctx->stream));
  return GFPL_OK;
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
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t keep = ~0ull;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
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
  cudaStreamDestroy(ctx->stream);
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

// ---- M1/M2 --------------------------------------------------------------------------------------------
int gfpl_knn2_hamming256_batch(gfpl_ctx* ctx, int32_t B, const uint8_t* q, const int32_t* nq, int32_t cap_q,
                               const uint8_t* t, const int32_t* nt, int32_t cap_t, int32_t norm, int32_t* out_idx,
                               float* out_dist) {
  if (!ctx) return GFPL_E_INVALID;
  REQUIRE(ctx, B >= 0 && cap_q >= 0 && cap_t >= 0);
  REQUIRE(ctx, norm == GFPL_NORM_HAMMING || norm == GFPL_NORM_HAMMING2);
  if (B == 0 || cap_q == 0) return GFPL_OK;
  REQUIRE(ctx, q && nq && t && nt && out_idx && out_dist && cap_t > 0);
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

int gfpl_measure_peaks(gfpl_ctx* ctx, double* popc_tops, double* fp64_tflops) {
  if (!ctx || !popc_tops || !fp64_tflops) return GFPL_E_INVALID;
  GFPL_CUDA_OK(ctx, cudaSetDevice(ctx->device));
  return launch_measure_peaks(ctx, popc_tops, fp64_tflops);
}

}  // extern "C"
