// C ABI of the B200-native ESLAM hot path (include/eslam_b200.h).  One translation unit: the constant
// decoder block is shared by every kernel.  Host side: argument checks, kernel-argument packing, launches.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <cmath>

#include "field.cuh"

#include "render.cuh"
#include "sample.cuh"
#include "optim.cuh"
#include "exchange.cuh"
#include "keyframes.cuh"
#include "ingest.cuh"
#include "qplane.cuh"
#include "qbwd.cuh"
#include "mcubes.cuh"

using namespace eslam;

static thread_local char g_err[512] = "";
static int g_debug = 0;

static int fail(int code, const char* what) {
  if (code > 0)
    snprintf(g_err, sizeof(g_err), "%s: %s", what, cudaGetErrorString((cudaError_t)code));
  else
    snprintf(g_err, sizeof(g_err), "%s: invalid or unsupported argument (%d)", what, code);
  return code;
}

#define CHECK_LAUNCH(what)                         \
  do {                                             \
    cudaError_t e_ = cudaGetLastError();           \
    if (e_ != cudaSuccess) return fail((int)e_, what); \
  } while (0)

#define REQUIRE(cond, what) \
  do {                      \
    if (!(cond)) return fail(ESLAM_EINVAL, what); \
  } while (0)

static inline cudaStream_t S_(eslam_stream_t s) { return (cudaStream_t)s; }

struct CfgK {
  int n_strat, n_imp;
  float tr, tr15, tr3, tr04;
  float w_fs, w_center, w_tail, w_depth;
  double w_color;
};

// Python evaluates `1.5 * truncation`, `3 * truncation`, `0.4 * truncation` in doubles before torch rounds
// the scalar to fp32 (Renderer.py:97, Tracker.py:135)
static CfgK make_cfg(const eslam_render_cfg_t* c) {
  CfgK k;
  k.n_strat = c->n_stratified;
  k.n_imp = c->n_importance;
  k.tr = (float)c->truncation;
  k.tr15 = (float)(1.5 * c->truncation);
  k.tr3 = (float)(3 * c->truncation);
  k.tr04 = (float)(0.4 * c->truncation);
  k.w_fs = (float)c->w_fs;
  k.w_center = (float)c->w_center;
  k.w_tail = (float)c->w_tail;
  k.w_depth = (float)c->w_depth;
  k.w_color = c->w_color;
  return k;
}

static int check_samples(int n_strat, int n_imp) {
  const int S = n_strat + n_imp;
  if (n_strat < 2 || n_imp < 1 || n_imp > 16 || S > ESLAM_MAX_SAMPLES || n_strat > NP) return ESLAM_EUNSUPPORTED;
  return 0;
}

template <typename K>
static int ensure_smem(K kernel, size_t bytes, bool* done) {
  if (*done) return 0;
  int rc = (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (rc == 0) *done = true;
  return rc;
}

template <typename K>
static int set_smem(K kernel, size_t bytes) {
  return (int)cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
}

// cudaFuncSetAttribute is per DEVICE: remember per device which kernels have been configured
constexpr int MAX_DEVICES = 64;
static int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return d >= 0 && d < MAX_DEVICES ? d : 0;
}

extern "C" {

const char* eslam_last_error(void) { return g_err; }
int eslam_abi_version(void) { return ESLAM_ABI_VERSION; }
void eslam_set_debug(int flags) { g_debug = flags; }
#ifdef ESLAM_PROFILE_PHASES
// profiling build only (tools/phase_profile.py): read and clear the per-phase clock sums of the backward kernel
int eslam_phase_counters(unsigned long long* out_host) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(out_host, g_phase, sizeof(unsigned long long) * 32);
  unsigned long long z[32] = {0};
  cudaMemcpyToSymbol(g_phase, z, sizeof(z));
  return 0;
}
#endif

int eslam_plane_import(const float* nchw, float* arena, const eslam_plane_t* pl, eslam_stream_t s) {
  REQUIRE(nchw && arena && pl && pl->H > 0 && pl->W > 0, "eslam_plane_import");
  const int HW = pl->H * pl->W;
  k_plane_layout<true><<<(HW + 31) / 32, 256, 0, S_(s)>>>(nchw, arena + pl->offset, HW);
  CHECK_LAUNCH("eslam_plane_import");
  return 0;
}

int eslam_plane_export(const float* arena, float* nchw, const eslam_plane_t* pl, eslam_stream_t s) {
  REQUIRE(nchw && arena && pl && pl->H > 0 && pl->W > 0, "eslam_plane_export");
  const int HW = pl->H * pl->W;
  k_plane_layout<false><<<(HW + 31) / 32, 256, 0, S_(s)>>>(arena + pl->offset, nchw, HW);
  CHECK_LAUNCH("eslam_plane_export");
  return 0;
}

// The forward-only kernels read the decoders as constant-bank operands (c_dec, field.cuh).  The block is refreshed
// with the documented device-to-device cudaMemcpyToSymbolAsync on the caller's stream: the runtime orders it with the
// kernels around it and takes care of the constant caches (a kernel storing through cudaGetSymbolAddress's pointer
// would rely on undocumented invalidation at kernel boundaries).  Inside a CUDA graph it is captured as a memcpy node.
int eslam_bind_decoders(const float* dec, eslam_stream_t s) {
  REQUIRE(dec, "eslam_bind_decoders");
  cudaError_t e = cudaMemcpyToSymbolAsync(c_dec, dec, sizeof(float) * DEC_N, 0, cudaMemcpyDeviceToDevice, S_(s));
  if (e != cudaSuccess) return fail((int)e, "eslam_bind_decoders");
  return 0;
}

int eslam_decode_points(const eslam_field_t* f, const float* arena, const float* pts, int64_t n, float* raw,
                        int flags, eslam_stream_t s) {
  REQUIRE(f && arena && pts && raw && n >= 0, "eslam_decode_points");
  if (n == 0) return 0;
  DecodeArgs a;
  memset(&a, 0, sizeof(a));
  int rc = make_field_k(f, &a.fk);
  if (rc) return fail(rc, "eslam_decode_points(field)");
  a.arena4 = reinterpret_cast<const float4*>(arena);
  a.pts = pts;
  a.n = n;
  a.raw = raw;
  a.flags = flags & 7;
  k_decode<<<(unsigned)((n + NP - 1) / NP), NP, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_decode_points");
  return 0;
}

int eslam_sample_plane_feature(const eslam_field_t* f, const float* arena, const float* p_nor, int64_t n, int which,
                               float* feat, eslam_stream_t s) {
  REQUIRE(f && arena && p_nor && feat && n >= 0 && (which == 0 || which == 1), "eslam_sample_plane_feature");
  if (n == 0) return 0;
  FeatArgs a;
  int rc = make_field_k(f, &a.fk);
  if (rc) return fail(rc, "eslam_sample_plane_feature(field)");
  a.arena4 = reinterpret_cast<const float4*>(arena);
  a.p_nor = p_nor;
  a.n = n;
  a.which = which;
  a.feat4 = reinterpret_cast<float4*>(feat);
  k_plane_feature<<<(unsigned)((n + NP - 1) / NP), NP, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_sample_plane_feature");
  return 0;
}

static int grid_sdf_impl(const eslam_field_t* f, const float* arena, const float* xs, const float* ys, const float* zs,
                         int nx, int ny, int nz, int64_t start, int64_t count, const float* hull, int n_hull,
                         const float* fxy, const float* fxz, const float* fyz, float* sdf, eslam_stream_t s) {
  REQUIRE(f && arena && xs && ys && zs && sdf && nx > 0 && ny > 0 && nz > 0 && start >= 0 && count >= 0 &&
              start + count <= (int64_t)nx * ny * nz && n_hull >= 0 && (n_hull == 0 || hull),
          "eslam_grid_sdf");
  if (count == 0) return 0;
  DecodeArgs a;
  memset(&a, 0, sizeof(a));
  int rc = make_field_k(f, &a.fk);
  if (rc) return fail(rc, "eslam_grid_sdf(field)");
  a.arena4 = reinterpret_cast<const float4*>(arena);
  a.n = count;
  a.sdf_out = sdf;
  a.flags = 1 | 2 | (n_hull > 0 ? 8 : 0) | (fxy ? 16 : 0);
  a.fxy = reinterpret_cast<const float4*>(fxy);
  a.fxz = reinterpret_cast<const float4*>(fxz);
  a.fyz = reinterpret_cast<const float4*>(fyz);
  a.hull = reinterpret_cast<const float4*>(hull);
  a.n_hull = n_hull;
  a.xs = xs;
  a.ys = ys;
  a.zs = zs;
  a.nx = nx;
  a.ny = ny;
  a.nz = nz;
  a.start = start;
  k_decode<<<(unsigned)((count + NP - 1) / NP), NP, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_grid_sdf");
  return 0;
}

int eslam_grid_sdf(const eslam_field_t* f, const float* arena, const float* xs, const float* ys, const float* zs,
                   int nx, int ny, int nz, int64_t start, int64_t count, float* sdf, eslam_stream_t s) {
  return grid_sdf_impl(f, arena, xs, ys, zs, nx, ny, nz, start, count, nullptr, 0, nullptr, nullptr, nullptr, sdf, s);
}

int eslam_grid_sdf_hull(const eslam_field_t* f, const float* arena, const float* xs, const float* ys, const float* zs,
                        int nx, int ny, int nz, int64_t start, int64_t count, const float* hull_planes, int n_planes,
                        float* sdf, eslam_stream_t s) {
  REQUIRE(hull_planes && n_planes > 0, "eslam_grid_sdf_hull");
  return grid_sdf_impl(f, arena, xs, ys, zs, nx, ny, nz, start, count, hull_planes, n_planes, nullptr, nullptr, nullptr,
                       sdf, s);
}

int eslam_grid_features(const eslam_field_t* f, const float* arena, const float* xs, const float* ys, const float* zs,
                        int nx, int ny, int nz, float* fxy, float* fxz, float* fyz, eslam_stream_t s) {
  REQUIRE(f && arena && xs && ys && zs && fxy && fxz && fyz && nx > 0 && ny > 0 && nz > 0, "eslam_grid_features");
  REQUIRE(nx <= 32767 && ny <= 32767 && nz <= 32767, "eslam_grid_features(lattice size)");
  GridFeatArgs a;
  memset(&a, 0, sizeof(a));
  int rc = make_field_k(f, &a.fk);
  if (rc) return fail(rc, "eslam_grid_features(field)");
  a.arena4 = reinterpret_cast<const float4*>(arena);
  const float* us[3] = {xs, xs, ys};
  const float* vs[3] = {ys, zs, zs};
  const int na[3] = {nx, nx, ny}, nb[3] = {ny, nz, nz}, ua[3] = {0, 0, 1}, va[3] = {1, 2, 2};
  float* out[3] = {fxy, fxz, fyz};
  for (int p = 0; p < 3; ++p) {
    a.us = us[p];
    a.vs = vs[p];
    a.na = na[p];
    a.nb = nb[p];
    a.plane = p;
    a.ua = ua[p];
    a.va = va[p];
    a.out = reinterpret_cast<float4*>(out[p]);
    const long long n = (long long)na[p] * nb[p];
    k_grid_features<<<(unsigned)((n + 31) / 32), 256, 0, S_(s)>>>(a);
    CHECK_LAUNCH("eslam_grid_features");
  }
  return 0;
}

int eslam_grid_sdf_separable(const eslam_field_t* f, const float* arena, const float* xs, const float* ys,
                             const float* zs, int nx, int ny, int nz, int64_t start, int64_t count, const float* fxy,
                             const float* fxz, const float* fyz, const float* hull_planes, int n_planes, float* sdf,
                             eslam_stream_t s) {
  REQUIRE(fxy && fxz && fyz && nx <= 32767 && ny <= 32767 && nz <= 32767, "eslam_grid_sdf_separable");
  return grid_sdf_impl(f, arena, xs, ys, zs, nx, ny, nz, start, count, hull_planes, n_planes, fxy, fxz, fyz, sdf, s);
}

int eslam_grid_preact(const eslam_field_t* f, const float* arena, const float* xs, const float* ys, const float* zs,
                      int nx, int ny, int nz, int iy0, int iy1, float* pxy, float* pxz, float* pyz, eslam_stream_t s) {
  REQUIRE(f && arena && xs && ys && zs && pxy && pxz && pyz && nx > 0 && ny > 0 && nz > 0 && iy0 >= 0 && iy1 > iy0 &&
              iy1 <= ny,
          "eslam_grid_preact");
  REQUIRE(nx <= 32767 && ny <= 32767 && nz <= 32767, "eslam_grid_preact(lattice size)");
  GridPreArgs a;
  memset(&a, 0, sizeof(a));
  int rc = make_field_k(f, &a.fk);
  if (rc) return fail(rc, "eslam_grid_preact(field)");
  a.arena4 = reinterpret_cast<const float4*>(arena);
  a.w1 = arena + f->dec_offset + S_W1;
  // rows of y in [iy0, iy1) only: the xy face out[iy][ix][4] and the yz face out[iy][4][iz] are both y-major, so a
  // row range is the same kernel on a shorter face (shifted coordinate and output pointers); the xz face is whole
  const int nyr = iy1 - iy0;
  const float* us[3] = {xs, xs, ys + iy0};
  const float* vs[3] = {ys + iy0, zs, zs};
  const int na[3] = {nx, nx, nyr}, nb[3] = {nyr, nz, nz}, ua[3] = {0, 0, 1}, va[3] = {1, 2, 2};
  float* out[3] = {pxy + (long long)iy0 * nx * 16, pxz, pyz + (long long)iy0 * nz * 16};
  for (int p = 0; p < 3; ++p) {
    a.us = us[p];
    a.vs = vs[p];
    a.na = na[p];
    a.nb = nb[p];
    a.plane = p;
    a.ua = ua[p];
    a.va = va[p];
    a.zmajor = p > 0;
    a.out = reinterpret_cast<float4*>(out[p]);
    const long long n = (long long)na[p] * nb[p];
    const long long want = (n + 127) / 128;  // >= 4 trips per CTA amortise the 4 KB copy of W1
    k_grid_preact<<<(unsigned)(want < 1 ? 1 : want), 256, 0, S_(s)>>>(a);
    CHECK_LAUNCH("eslam_grid_preact");
  }
  return 0;
}

int eslam_grid_sdf_factored(const eslam_field_t* f, const float* xs, const float* ys, const float* zs, int nx, int ny,
                            int nz, int64_t start, int64_t count, const float* pxy, const float* pxz, const float* pyz,
                            const float* hull_planes, int n_planes, float* sdf, eslam_stream_t s) {
  REQUIRE(f && xs && ys && zs && pxy && pxz && pyz && sdf && nx > 0 && ny > 0 && nz > 0 && start >= 0 && count >= 0 &&
              nx <= 32767 && ny <= 32767 && nz <= 32767 && start + count <= (int64_t)nx * ny * nz && n_planes >= 0 &&
              (n_planes == 0 || hull_planes),
          "eslam_grid_sdf_factored");
  if (count == 0) return 0;
  GridFacArgs a;
  memset(&a, 0, sizeof(a));
  for (int k = 0; k < 3; ++k) {
    a.lo[k] = f->bound[k][0];
    a.hi[k] = f->bound[k][1];
  }
  a.xs = xs;
  a.ys = ys;
  a.zs = zs;
  a.nx = nx;
  a.ny = ny;
  a.nz = nz;
  a.start = start;
  a.n = count;
  a.pxy = reinterpret_cast<const float4*>(pxy);
  a.pxz = reinterpret_cast<const float4*>(pxz);
  a.pyz = reinterpret_cast<const float4*>(pyz);
  a.hull = reinterpret_cast<const float4*>(hull_planes);
  a.n_hull = n_planes;
  a.sdf_out = sdf;
  const int64_t per_cta = FAC_THREADS * FAC_PT;
  k_grid_sdf_factored<<<(unsigned)((count + per_cta - 1) / per_cta), FAC_THREADS, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_grid_sdf_factored");
  return 0;
}

int eslam_grid_sdf_rows(const eslam_field_t* f, const float* xs, const float* ys, const float* zs, int nx, int ny, int nz,
                        int iy0, int iy1, int64_t out_base, const float* pxy, const float* pxz, const float* pyz,
                        const float* hull_planes, int n_planes, float* sdf, eslam_stream_t s) {
  REQUIRE(f && xs && ys && zs && pxy && pxz && pyz && sdf && nx > 0 && ny > 0 && nz > 0 && iy0 >= 0 && iy1 <= ny &&
              iy0 <= iy1 && nx <= 32767 && ny <= 32767 && nz <= 32767 && out_base >= 0 &&
              out_base <= (int64_t)iy0 * nx * nz && n_planes >= 0 && (n_planes == 0 || hull_planes),
          "eslam_grid_sdf_rows");
  if (iy0 == iy1) return 0;
  GridRowsArgs a;
  memset(&a, 0, sizeof(a));
  for (int k = 0; k < 3; ++k) {
    a.lo[k] = f->bound[k][0];
    a.hi[k] = f->bound[k][1];
  }
  a.xs = xs;
  a.ys = ys;
  a.zs = zs;
  a.nx = nx;
  a.ny = ny;
  a.nz = nz;
  a.iy0 = iy0;
  a.iy1 = iy1;
  a.out_base = out_base;
  a.pxy = reinterpret_cast<const float4*>(pxy);
  a.pxz = reinterpret_cast<const float4*>(pxz);
  a.pyz = reinterpret_cast<const float4*>(pyz);
  a.hull = reinterpret_cast<const float4*>(hull_planes);
  a.n_hull = n_planes;
  a.sdf_out = sdf;
  const dim3 grid((unsigned)((nx + GR_X - 1) / GR_X), (unsigned)((nz + GR_Z - 1) / GR_Z));
  k_grid_sdf_rows<<<grid, GR_THREADS, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_grid_sdf_rows");
  return 0;
}

static int sample_rays_impl(const eslam_field_t* f, const eslam_camera_t* cam, const eslam_render_cfg_t* cfg,
                            const int64_t* pix_idx, int n_img, int n_per_img, const float* c2w, const float* poses,
                            int pose_first, const float* depth, const double* color, bool frame_table,
                            const float* u_depth, const float* t_uni, const float* t_surf, int need_depth,
                            float* rays_o, float* rays_d, float* gt_depth, double* gt_color, int32_t* src, float* z,
                            int32_t* dl_list, int32_t* zord, uint8_t* band, int32_t* counters, float* c2w_out,
                            eslam_stream_t s) {
  REQUIRE(f && cam && cfg && pix_idx && depth && color && t_uni && t_surf && rays_o && rays_d && gt_depth &&
              gt_color && src && z && dl_list && zord && band && counters && n_img > 0 && n_per_img > 0 &&
              (c2w || poses),
          "eslam_sample_rays");
  REQUIRE(c2w || pose_first == 0, "eslam_sample_rays(c2w)");
  int rc = check_samples(cfg->n_stratified, cfg->n_importance);
  if (rc) return fail(rc, "eslam_sample_rays(samples)");
  SampleArgs a;
  memset(&a, 0, sizeof(a));
  rc = make_field_k(f, &a.fk);
  if (rc) return fail(rc, "eslam_sample_rays(field)");
  const CfgK k = make_cfg(cfg);
  a.H = cam->H;
  a.W = cam->W;
  a.H0 = cam->H0;
  a.W0 = cam->W0;
  a.Wc = cam->W1 - cam->W0;
  a.HWc = a.Wc * (cam->H1 - cam->H0);
  REQUIRE(a.Wc > 0 && cam->H1 > cam->H0 && cam->H1 <= cam->H && cam->W1 <= cam->W && cam->H0 >= 0 && cam->W0 >= 0,
          "eslam_sample_rays(crop)");
  a.fx = cam->fx;
  a.fy = cam->fy;
  a.cx = cam->cx;
  a.cy = cam->cy;
  a.n_strat = k.n_strat;
  a.n_imp = k.n_imp;
  a.tr = k.tr;
  a.tr15 = k.tr15;
  a.tr3 = k.tr3;
  a.tr04 = k.tr04;
  a.pix_idx = reinterpret_cast<const long long*>(pix_idx);
  a.n_img = n_img;
  a.n_per_img = n_per_img;
  a.c2w = c2w;
  a.poses = poses;
  a.pose_first = pose_first;
  if (frame_table) {
    a.depth_tab = reinterpret_cast<const float* const*>(depth);
    a.color_tab = reinterpret_cast<const double* const*>(color);
  } else {
    a.depth = depth;
    a.color = color;
  }
  a.u_depth = u_depth;
  a.t_uni = t_uni;
  a.t_surf = t_surf;
  a.need_depth = need_depth;
  a.rays_o = rays_o;
  a.rays_d = rays_d;
  a.gt_depth = gt_depth;
  a.gt_color = gt_color;
  a.src = src;
  a.z = z;
  a.dl_list = dl_list;
  a.zord = zord;
  a.band = band;
  a.counters = counters;
  a.c2w_out = c2w_out;
  const int N = n_img * n_per_img;
  const int n_blocks = (N + SB - 1) / SB;
  if (n_blocks > ESLAM_MAX_COMPACT_BLOCKS) return fail(ESLAM_EUNSUPPORTED, "eslam_sample_rays(too many rays per call)");
  a.agg = reinterpret_cast<unsigned long long*>(counters + ESLAM_N_COUNTERS);
  cudaError_t e = cudaMemsetAsync(counters, 0, sizeof(int32_t) * (ESLAM_N_COUNTERS + 2 * n_blocks), S_(s));
  if (e != cudaSuccess) return fail((int)e, "eslam_sample_rays(memset)");
  k_sample_rays<<<(N + SB - 1) / SB, SB, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_sample_rays(select)");
  RaySampleArgs b;
  b.n_strat = k.n_strat;
  b.n_imp = k.n_imp;
  b.tr = k.tr;
  b.tr15 = k.tr15;
  b.tr3 = k.tr3;
  b.tr04 = k.tr04;
  b.gt_depth = gt_depth;
  b.zord = zord;
  b.max_rays = N;
  b.u_depth = u_depth;
  b.t_uni = t_uni;
  b.t_surf = t_surf;
  b.z = z;
  b.band = band;
  b.counters = counters;
  k_ray_samples<<<(N + SB / 32 - 1) / (SB / 32), SB, 0, S_(s)>>>(b);
  CHECK_LAUNCH("eslam_sample_rays(samples)");
  return 0;
}


int eslam_sample_rays(const eslam_field_t* f, const eslam_camera_t* cam, const eslam_render_cfg_t* cfg,
                      const int64_t* pix_idx, int n_img, int n_per_img, const float* c2w, const float* poses,
                      int pose_first, const float* depth, const double* color, const float* u_depth,
                      const float* t_uni, const float* t_surf, int need_depth, float* rays_o, float* rays_d,
                      float* gt_depth, double* gt_color, int32_t* src, float* z, int32_t* dl_list, int32_t* zord,
                      uint8_t* band, int32_t* counters, float* c2w_out, eslam_stream_t s) {
  return sample_rays_impl(f, cam, cfg, pix_idx, n_img, n_per_img, c2w, poses, pose_first, depth, color, false, u_depth,
                          t_uni, t_surf, need_depth, rays_o, rays_d, gt_depth, gt_color, src, z, dl_list, zord, band,
                          counters, c2w_out, s);
}

int eslam_sample_rays_frames(const eslam_field_t* f, const eslam_camera_t* cam, const eslam_render_cfg_t* cfg,
                             const int64_t* pix_idx, int n_img, int n_per_img, const float* c2w, const float* poses,
                             int pose_first, const float* const* depth_frames, const double* const* color_frames,
                             const float* u_depth, const float* t_uni, const float* t_surf, int need_depth,
                             float* rays_o, float* rays_d, float* gt_depth, double* gt_color, int32_t* src, float* z,
                             int32_t* dl_list, int32_t* zord, uint8_t* band, int32_t* counters, float* c2w_out,
                             eslam_stream_t s) {
  return sample_rays_impl(f, cam, cfg, pix_idx, n_img, n_per_img, c2w, poses, pose_first,
                          reinterpret_cast<const float*>(depth_frames), reinterpret_cast<const double*>(color_frames),
                          true, u_depth, t_uni, t_surf, need_depth, rays_o, rays_d, gt_depth, gt_color, src, z, dl_list,
                          zord, band, counters, c2w_out, s);
}

int eslam_depth_samples(const eslam_render_cfg_t* cfg, const float* gt_depth, int n_rays, const float* u_depth,
                        const float* t_uni, const float* t_surf, float* z, int32_t* dl_list, int32_t* zord,
                        int32_t* counters, eslam_stream_t s) {
  REQUIRE(cfg && gt_depth && t_uni && t_surf && z && dl_list && zord && counters && n_rays >= 0,
          "eslam_depth_samples");
  int rc = check_samples(cfg->n_stratified, cfg->n_importance);
  if (rc) return fail(rc, "eslam_depth_samples(samples)");
  const int n_blocks = (n_rays + SB - 1) / SB;
  if (n_blocks > ESLAM_MAX_COMPACT_BLOCKS) return fail(ESLAM_EUNSUPPORTED, "eslam_depth_samples(too many rays per call)");
  cudaError_t e = cudaMemsetAsync(counters, 0, sizeof(int32_t) * (ESLAM_N_COUNTERS + 2 * n_blocks), S_(s));
  if (e != cudaSuccess) return fail((int)e, "eslam_depth_samples(memset)");
  if (n_rays == 0) return 0;
  const CfgK k = make_cfg(cfg);
  DepthOrdArgs o;
  o.gt_depth = gt_depth;
  o.n_rays = n_rays;
  o.zord = zord;
  o.dl_list = dl_list;
  o.counters = counters;
  o.agg = reinterpret_cast<unsigned long long*>(counters + ESLAM_N_COUNTERS);
  k_depth_ordinals<<<(n_rays + SB - 1) / SB, SB, 0, S_(s)>>>(o);
  CHECK_LAUNCH("eslam_depth_samples(ordinals)");
  RaySampleArgs b;
  b.n_strat = k.n_strat;
  b.n_imp = k.n_imp;
  b.tr = k.tr;
  b.tr15 = k.tr15;
  b.tr3 = k.tr3;
  b.tr04 = k.tr04;
  b.gt_depth = gt_depth;
  b.zord = zord;
  b.max_rays = n_rays;
  b.u_depth = u_depth;
  b.t_uni = t_uni;
  b.t_surf = t_surf;
  b.z = z;
  b.band = nullptr;
  b.counters = counters;
  k_ray_samples<<<(n_rays + SB / 32 - 1) / (SB / 32), SB, 0, S_(s)>>>(b);
  CHECK_LAUNCH("eslam_depth_samples(samples)");
  return 0;
}

int eslam_importance_samples(const eslam_field_t* f, const float* arena, const float* q_arena,
                             const eslam_render_cfg_t* cfg, const float* rays_o, const float* rays_d,
                             const int32_t* dl_list, const int32_t* counters, int max_rays, const float* u_coarse,
                             const float* u_fine, const float* t_uni, float* z, eslam_stream_t s) {
  REQUIRE(f && arena && q_arena && cfg && rays_o && rays_d && dl_list && counters && u_coarse && u_fine && t_uni && z &&
              max_rays >= 0,
          "eslam_importance_samples");
  if (max_rays == 0) return 0;
  int rc = check_samples(cfg->n_stratified, cfg->n_importance);
  if (rc) return fail(rc, "eslam_importance_samples(samples)");
  ImportanceArgs a;
  rc = make_field_k(f, &a.fk);
  if (rc) return fail(rc, "eslam_importance_samples(field)");
  a.arena4 = reinterpret_cast<const float4*>(arena);
  a.q4 = reinterpret_cast<const float4*>(q_arena);
  a.n_strat = cfg->n_stratified;
  a.n_imp = cfg->n_importance;
  a.rays_o = rays_o;
  a.rays_d = rays_d;
  a.dl_list = dl_list;
  a.counters = counters;
  a.u_coarse = u_coarse;
  a.u_fine = u_fine;
  a.t_uni = t_uni;
  a.z = z;
  const int rpb = NP / a.n_strat;
  k_importance<<<(max_rays + rpb - 1) / rpb, NP, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_importance_samples");
  return 0;
}

static int render_forward_impl(const eslam_field_t* f, const float* arena, const float* rays_o, const float* rays_d,
                               const float* z, int n_rays, int n_samples, const int32_t* counters, float* depth,
                               float* rgb, float* sdf, float* act4, uint32_t* actm, eslam_stream_t s) {
  REQUIRE(f && arena && rays_o && rays_d && z && depth && rgb && n_rays >= 0, "eslam_render_forward");
  REQUIRE((act4 == nullptr) == (actm == nullptr) && (!act4 || sdf), "eslam_render_forward(activations)");
  REQUIRE(n_samples >= 1 && n_samples <= ESLAM_MAX_SAMPLES, "eslam_render_forward(n_samples)");
  if (n_rays == 0) return 0;
  RenderFwdArgs a;
  int rc = make_field_k(f, &a.fk);
  if (rc) return fail(rc, "eslam_render_forward(field)");
  a.arena4 = reinterpret_cast<const float4*>(arena);
  a.rays_o = rays_o;
  a.rays_d = rays_d;
  a.z = z;
  a.n_rays = n_rays;
  a.S = n_samples;
  a.counters = counters;
  a.depth = depth;
  a.rgb = rgb;
  a.sdf = sdf;
  a.act4 = reinterpret_cast<float4*>(act4);
  a.actm = actm;
  const int rpb = NP / n_samples;
  k_render_fwd<<<(n_rays + rpb - 1) / rpb, NP, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_render_forward");
  return 0;
}

int eslam_render_forward(const eslam_field_t* f, const float* arena, const float* rays_o, const float* rays_d,
                         const float* z, int n_rays, int n_samples, const int32_t* counters, float* depth, float* rgb,
                         float* sdf, eslam_stream_t s) {
  return render_forward_impl(f, arena, rays_o, rays_d, z, n_rays, n_samples, counters, depth, rgb, sdf, nullptr, nullptr,
                             s);
}

int eslam_render_forward_act(const eslam_field_t* f, const float* arena, const float* rays_o, const float* rays_d,
                             const float* z, int n_rays, int n_samples, const int32_t* counters, float* depth,
                             float* rgb, float* sdf, float* act4, uint32_t* actm, eslam_stream_t s) {
  REQUIRE(act4 && actm && sdf, "eslam_render_forward_act");
  return render_forward_impl(f, arena, rays_o, rays_d, z, n_rays, n_samples, counters, depth, rgb, sdf, act4, actm, s);
}

}  // extern "C"

template <int MODE, bool GF, bool GR>
static int launch_bwd(const BwdArgs& a, int n_rays, cudaStream_t st) {
  static bool configured[MAX_DEVICES] = {false};
  const size_t bytes = sizeof(SmemBwd<GF>);
  static_assert(sizeof(SmemBwd<true>) <= 113 * 1024, "two CTAs of the backward kernel must fit one SM's shared memory");
  const int dev = current_device();
  if (!configured[dev]) {
    int rc = set_smem(k_render_bwd<MODE, GF, GR>, bytes);
    if (rc) return rc;
    configured[dev] = true;
  }
  const int rpb = MODE == 2 ? NP : ((NP / a.S) < 16 ? (NP / a.S) : 16);
  k_render_bwd<MODE, GF, GR><<<(n_rays + rpb - 1) / rpb, NT_BWD, bytes, st>>>(a);
  return (int)cudaGetLastError();
}

extern "C" {

int eslam_render_backward(const eslam_field_t* f, const float* arena, const float* rays_o, const float* rays_d,
                          const float* z, int n_rays, int n_samples, const float* g_depth, const float* g_rgb,
                          const float* g_sdf, float* grad_arena, float* g_rays_o, float* g_rays_d,
                          eslam_stream_t s) {
  REQUIRE(f && arena && rays_o && rays_d && z && g_depth && g_rgb && n_rays >= 0, "eslam_render_backward");
  REQUIRE(n_samples >= 1 && n_samples <= ESLAM_MAX_SAMPLES, "eslam_render_backward(n_samples)");
  REQUIRE((g_rays_o == nullptr) == (g_rays_d == nullptr), "eslam_render_backward(ray grads)");
  if (n_rays == 0 || (!grad_arena && !g_rays_o)) return 0;
  BwdArgs a;
  memset(&a, 0, sizeof(a));
  int rc = make_field_k(f, &a.fk);
  if (rc) return fail(rc, "eslam_render_backward(field)");
  a.arena4 = reinterpret_cast<const float4*>(arena);
  a.rays_o = rays_o;
  a.rays_d = rays_d;
  a.z = z;
  a.n_rays = n_rays;
  a.S = n_samples;
  a.g_depth = g_depth;
  a.g_rgb = g_rgb;
  a.g_sdf = g_sdf;
  a.grad_arena = grad_arena;
  a.g_rays_o = g_rays_o;
  a.g_rays_d = g_rays_d;
  if (grad_arena && g_rays_o)
    rc = launch_bwd<0, true, true>(a, n_rays, S_(s));
  else if (grad_arena)
    rc = launch_bwd<0, true, false>(a, n_rays, S_(s));
  else
    rc = launch_bwd<0, false, true>(a, n_rays, S_(s));
  if (rc) return fail(rc, "eslam_render_backward");
  return 0;
}

int eslam_decode_backward(const eslam_field_t* f, const float* arena, const float* pts, int64_t n, const float* g_raw,
                          float* grad_arena, float* g_pts, eslam_stream_t s) {
  REQUIRE(f && arena && pts && g_raw && n >= 0 && n < 0x7fffffff / 4, "eslam_decode_backward");
  if (n == 0 || (!grad_arena && !g_pts)) return 0;
  BwdArgs a;
  memset(&a, 0, sizeof(a));
  int rc = make_field_k(f, &a.fk);
  if (rc) return fail(rc, "eslam_decode_backward(field)");
  a.arena4 = reinterpret_cast<const float4*>(arena);
  a.rays_o = pts;
  a.n_rays = (int)n;
  a.S = 1;
  a.g_sdf = g_raw;
  a.grad_arena = grad_arena;
  a.g_rays_o = g_pts;
  if (grad_arena && g_pts)
    rc = launch_bwd<2, true, true>(a, (int)n, S_(s));
  else if (grad_arena)
    rc = launch_bwd<2, true, false>(a, (int)n, S_(s));
  else
    rc = launch_bwd<2, false, true>(a, (int)n, S_(s));
  if (rc) return fail(rc, "eslam_decode_backward");
  return 0;
}

int eslam_track_mask(const float* gt_depth, const float* depth, const uint8_t* band, int max_rays, int32_t* counters,
                     uint8_t* ray_mask, float* scratch, eslam_stream_t s) {
  REQUIRE(gt_depth && depth && band && counters && ray_mask && scratch && max_rays > 0, "eslam_track_mask");
  TrackMaskArgs a;
  a.gt_depth = gt_depth;
  a.depth = depth;
  a.band = band;
  a.max_rays = max_rays;
  a.counters = counters;
  a.ray_mask = ray_mask;
  a.scratch = scratch;
  k_track_mask<<<1, 1024, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_track_mask");
  return 0;
}

static int loss_backward_impl(const eslam_field_t* f, const float* arena, const eslam_camera_t* cam,
                              const eslam_render_cfg_t* cfg, const float* rays_o, const float* rays_d, const float* z,
                              const float* gt_depth, const double* gt_color, const int32_t* src,
                              const int64_t* pix_idx, int n_per_img, const uint8_t* ray_mask, const int32_t* counters,
                              const int32_t* norm_counters, int max_rays, float* grad_arena, float* pose_grad,
                              double* loss_acc, const float* sdf, const float* act4, const uint32_t* actm,
                              const float* q_arena, float* gq_arena, eslam_stream_t s, int part = 0,
                              const int32_t* dl_list = nullptr) {
  REQUIRE(f && arena && cam && cfg && rays_o && rays_d && z && gt_depth && gt_color && counters && max_rays >= 0,
          "eslam_loss_backward");
  REQUIRE(part >= 0 && part <= 2 && (part != 2 || dl_list), "eslam_loss_backward(part)");
  REQUIRE(!pose_grad || (src && pix_idx && n_per_img > 0), "eslam_loss_backward(pose)");
  if (max_rays == 0) return 0;
  int rc = check_samples(cfg->n_stratified, cfg->n_importance);
  if (rc) return fail(rc, "eslam_loss_backward(samples)");
  const CfgK k = make_cfg(cfg);
  BwdArgs a;
  memset(&a, 0, sizeof(a));
  rc = make_field_k(f, &a.fk);
  if (rc) return fail(rc, "eslam_loss_backward(field)");
  a.arena4 = reinterpret_cast<const float4*>(arena);
  a.rays_o = rays_o;
  a.rays_d = rays_d;
  a.z = z;
  a.n_rays = max_rays;
  a.S = k.n_strat + k.n_imp;
  a.counters = counters;
  a.dbg = g_debug;
  a.part = part;
  a.dl_list = dl_list;
  a.norm = norm_counters ? norm_counters : counters;
  a.gt_depth = gt_depth;
  a.gt_color = gt_color;
  a.src = src;
  a.pix_idx = reinterpret_cast<const long long*>(pix_idx);
  a.n_per_img = n_per_img;
  a.ray_mask = ray_mask;
  a.tr = k.tr;
  a.tr04 = k.tr04;
  a.w_fs = k.w_fs;
  a.w_center = k.w_center;
  a.w_tail = k.w_tail;
  a.w_depth = k.w_depth;
  a.w_color = k.w_color;
  a.fx = cam->fx;
  a.fy = cam->fy;
  a.cx = cam->cx;
  a.cy = cam->cy;
  a.W0 = cam->W0;
  a.H0 = cam->H0;
  a.Wc = cam->W1 - cam->W0;
  a.loss_acc = loss_acc;
  a.grad_arena = grad_arena;
  a.pose_grad = pose_grad;
  a.sdf_in = sdf;
  a.act4 = reinterpret_cast<const float4*>(act4);
  a.actm = actm;
  a.q4 = reinterpret_cast<const float4*>(q_arena);
  a.gq4 = reinterpret_cast<float4*>(gq_arena);
  if (q_arena && grad_arena) {  // mapping iteration in the Q form
    REQUIRE(gq_arena && !act4, "eslam_loss_backward_q");
    static bool configured[MAX_DEVICES][2] = {{false, false}};
    const size_t bytes = sizeof(SmemBwdQ<true>);
    static_assert(sizeof(SmemBwdQ<true>) <= 113 * 1024, "two CTAs of the Q backward must fit one SM's shared memory");
    const int gr = pose_grad ? 1 : 0;
    const int dev = current_device();
    if (!configured[dev][gr]) {
      rc = gr ? set_smem(k_map_bwd_q<true>, bytes) : set_smem(k_map_bwd_q<false>, bytes);
      if (rc) return fail(rc, "eslam_loss_backward_q(shared memory)");
      configured[dev][gr] = true;
    }
    const int S = a.S, rpb = (NP / S) < 16 ? (NP / S) : 16;
    const unsigned grid = (unsigned)((max_rays + rpb - 1) / rpb);
    if (part == 2) {
      static bool configured_dl[MAX_DEVICES][2] = {{false, false}};
      if (!configured_dl[dev][gr]) {
        rc = gr ? set_smem(k_map_bwd_q_dl<true>, bytes) : set_smem(k_map_bwd_q_dl<false>, bytes);
        if (rc) return fail(rc, "eslam_loss_backward_q_part(shared memory)");
        configured_dl[dev][gr] = true;
      }
      int sms = 148;
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
      const unsigned g2 = grid < (unsigned)(2 * sms) ? grid : (unsigned)(2 * sms);
      if (gr)
        k_map_bwd_q_dl<true><<<g2, NT_BWD, bytes, S_(s)>>>(a);
      else
        k_map_bwd_q_dl<false><<<g2, NT_BWD, bytes, S_(s)>>>(a);
    } else if (part == 1) {
      // part 1 reads nothing the importance pass writes: launched with programmatic stream serialization, it starts
      // when the kernel in front of it in the stream has let its dependents go (k_importance does so at once; any
      // other kernel never does, which leaves ordinary stream order).  Its own CTAs hold every register of an SM, so
      // started the ordinary way -- beside the importance kernel on another stream -- it would starve that kernel.
      cudaLaunchConfig_t lc;
      memset(&lc, 0, sizeof(lc));
      lc.gridDim = dim3(grid);
      lc.blockDim = dim3(NT_BWD);
      lc.dynamicSmemBytes = bytes;
      lc.stream = S_(s);
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      lc.attrs = at;
      lc.numAttrs = 1;
      rc = (int)(gr ? cudaLaunchKernelEx(&lc, k_map_bwd_q<true>, a) : cudaLaunchKernelEx(&lc, k_map_bwd_q<false>, a));
    } else if (gr)
      k_map_bwd_q<true><<<grid, NT_BWD, bytes, S_(s)>>>(a);
    else
      k_map_bwd_q<false><<<grid, NT_BWD, bytes, S_(s)>>>(a);
    if (!rc) rc = (int)cudaGetLastError();
  } else if (q_arena) {
    REQUIRE(!grad_arena && pose_grad && sdf && act4 && actm, "eslam_pose_backward_q");
    static bool configured[MAX_DEVICES] = {false};
    const size_t bytes = sizeof(SmemBwdQ<false>);
    const int dev = current_device();
    if (!configured[dev]) {
      rc = set_smem(k_pose_bwd_q, bytes);
      if (rc) return fail(rc, "eslam_pose_backward_q(shared memory)");
      configured[dev] = true;
    }
    const int S = a.S, rpb = (NP / S) < 16 ? (NP / S) : 16;
    k_pose_bwd_q<<<(max_rays + rpb - 1) / rpb, NT_BWD, bytes, S_(s)>>>(a);
    rc = (int)cudaGetLastError();
  } else if (grad_arena && pose_grad)
    rc = launch_bwd<1, true, true>(a, max_rays, S_(s));
  else if (grad_arena)
    rc = launch_bwd<1, true, false>(a, max_rays, S_(s));
  else if (pose_grad)
    rc = launch_bwd<1, false, true>(a, max_rays, S_(s));
  else
    return fail(ESLAM_EINVAL, "eslam_loss_backward(no gradient requested)");
  if (rc) return fail(rc, "eslam_loss_backward");
  return 0;
}

int eslam_loss_backward(const eslam_field_t* f, const float* arena, const eslam_camera_t* cam,
                        const eslam_render_cfg_t* cfg, const float* rays_o, const float* rays_d, const float* z,
                        const float* gt_depth, const double* gt_color, const int32_t* src, const int64_t* pix_idx,
                        int n_per_img, const uint8_t* ray_mask, const int32_t* counters,
                        const int32_t* norm_counters, int max_rays, float* grad_arena, float* pose_grad,
                        double* loss_acc, eslam_stream_t s) {
  return loss_backward_impl(f, arena, cam, cfg, rays_o, rays_d, z, gt_depth, gt_color, src, pix_idx, n_per_img, ray_mask,
                            counters, norm_counters, max_rays, grad_arena, pose_grad, loss_acc, nullptr, nullptr,
                            nullptr, nullptr, nullptr, s);
}

int eslam_pose_backward_act(const eslam_field_t* f, const float* arena, const eslam_camera_t* cam,
                            const eslam_render_cfg_t* cfg, const float* rays_o, const float* rays_d, const float* z,
                            const float* gt_depth, const double* gt_color, const int32_t* src, const int64_t* pix_idx,
                            int n_per_img, const uint8_t* ray_mask, const int32_t* counters, int max_rays,
                            const float* sdf, const float* act4, const uint32_t* actm, float* pose_grad,
                            double* loss_acc, eslam_stream_t s) {
  REQUIRE(sdf && act4 && actm && pose_grad, "eslam_pose_backward_act");
  return loss_backward_impl(f, arena, cam, cfg, rays_o, rays_d, z, gt_depth, gt_color, src, pix_idx, n_per_img, ray_mask,
                            counters, nullptr, max_rays, nullptr, pose_grad, loss_acc, sdf, act4, actm, nullptr, nullptr, s);
}

// Q form (qbwd.cuh): eslam_pose_backward_act on the pre-activated plane images; the cached activations must
// come from eslam_render_forward_q on the same q_arena
int eslam_pose_backward_q(const eslam_field_t* f, const float* arena, const float* q_arena, const eslam_camera_t* cam,
                          const eslam_render_cfg_t* cfg, const float* rays_o, const float* rays_d, const float* z,
                          const float* gt_depth, const double* gt_color, const int32_t* src, const int64_t* pix_idx,
                          int n_per_img, const uint8_t* ray_mask, const int32_t* counters, int max_rays,
                          const float* sdf, const float* act4, const uint32_t* actm, float* pose_grad,
                          double* loss_acc, eslam_stream_t s) {
  REQUIRE(q_arena && sdf && act4 && actm && pose_grad, "eslam_pose_backward_q");
  return loss_backward_impl(f, arena, cam, cfg, rays_o, rays_d, z, gt_depth, gt_color, src, pix_idx, n_per_img, ray_mask,
                            counters, nullptr, max_rays, nullptr, pose_grad, loss_acc, sdf, act4, actm, q_arena, nullptr,
                            s);
}

// Q form (qbwd.cuh): eslam_loss_backward with planes + decoders (+ poses) in the Q form.  The plane gradients
// arrive as 16-channel reductions in gq_arena (layout of q_arena); grad_arena receives the decoder gradients except
// dW1 (formed by eslam_q_adam_planes from gq_arena) and beta.
int eslam_loss_backward_q(const eslam_field_t* f, const float* arena, const float* q_arena, float* gq_arena,
                          const eslam_camera_t* cam, const eslam_render_cfg_t* cfg, const float* rays_o,
                          const float* rays_d, const float* z, const float* gt_depth, const double* gt_color,
                          const int32_t* src, const int64_t* pix_idx, int n_per_img, const uint8_t* ray_mask,
                          const int32_t* counters, const int32_t* norm_counters, int max_rays, float* grad_arena,
                          float* pose_grad, double* loss_acc, eslam_stream_t s) {
  REQUIRE(q_arena && gq_arena && grad_arena, "eslam_loss_backward_q");
  return loss_backward_impl(f, arena, cam, cfg, rays_o, rays_d, z, gt_depth, gt_color, src, pix_idx, n_per_img, ray_mask,
                            counters, norm_counters, max_rays, grad_arena, pose_grad, loss_acc, nullptr, nullptr,
                            nullptr, q_arena, gq_arena, s);
}

int eslam_loss_backward_q_part(const eslam_field_t* f, const float* arena, const float* q_arena, float* gq_arena,
                               const eslam_camera_t* cam, const eslam_render_cfg_t* cfg, const float* rays_o,
                               const float* rays_d, const float* z, const float* gt_depth, const double* gt_color,
                               const int32_t* src, const int64_t* pix_idx, int n_per_img, const int32_t* dl_list,
                               const int32_t* counters, const int32_t* norm_counters, int max_rays, float* grad_arena,
                               float* pose_grad, int part, eslam_stream_t s) {
  REQUIRE(q_arena && gq_arena && grad_arena, "eslam_loss_backward_q_part");
  return loss_backward_impl(f, arena, cam, cfg, rays_o, rays_d, z, gt_depth, gt_color, src, pix_idx, n_per_img, nullptr,
                            counters, norm_counters, max_rays, grad_arena, pose_grad, nullptr, nullptr, nullptr,
                            nullptr, q_arena, gq_arena, s, part, dl_list);
}

static int fill_adam(AdamArgs& a, int64_t n, const int64_t* seg_end, const double* seg_lr, int n_seg, int step,
                     double beta1, double beta2, double eps) {
  if (!(seg_end && seg_lr && n > 0 && (n % 4) == 0 && n_seg >= 1 && n_seg <= 4 && step >= 1)) return ESLAM_EINVAL;
  a.n = n;
  a.n_seg = n_seg;
  const double bc1 = 1.0 - pow(beta1, (double)step);
  for (int i = 0; i < 4; ++i) {
    a.seg_end[i] = i < n_seg ? seg_end[i] : n;
    a.seg_step[i] = i < n_seg ? (float)(seg_lr[i] / bc1) : 0.f;
    if (i < n_seg && seg_end[i] % 4 != 0) return ESLAM_EINVAL;
  }
  a.beta1 = (float)beta1;
  a.beta2 = (float)beta2;
  a.one_m_beta1 = (float)(1.0 - beta1);
  a.one_m_beta2 = (float)(1.0 - beta2);
  a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  a.eps = (float)eps;
  return 0;
}

static int adam_step_impl(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                          const int64_t* seg_end, const double* seg_lr, int n_seg, int step, double beta1, double beta2,
                          double eps, uint8_t* touched, eslam_stream_t s) {
  REQUIRE(param && grad && exp_avg && exp_avg_sq, "eslam_adam_step");
  AdamArgs a;
  a.touched = touched;
  a.p = param;
  a.g = grad;
  a.m = exp_avg;
  a.v = exp_avg_sq;
  if (fill_adam(a, n, seg_end, seg_lr, n_seg, step, beta1, beta2, eps)) return fail(ESLAM_EINVAL, "eslam_adam_step");
  const long long n4 = n / 4;
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_adam<<<(unsigned)blocks, 256, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_adam_step");
  return 0;
}

int eslam_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                    const int64_t* seg_end, const double* seg_lr, int n_seg, int step, double beta1, double beta2,
                    double eps, eslam_stream_t s) {
  return adam_step_impl(param, grad, exp_avg, exp_avg_sq, n, seg_end, seg_lr, n_seg, step, beta1, beta2, eps, nullptr, s);
}

int eslam_adam_step_sparse(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                           const int64_t* seg_end, const double* seg_lr, int n_seg, int step, double beta1,
                           double beta2, double eps, uint8_t* touched, eslam_stream_t s) {
  REQUIRE(touched, "eslam_adam_step_sparse");
  return adam_step_impl(param, grad, exp_avg, exp_avg_sq, n, seg_end, seg_lr, n_seg, step, beta1, beta2, eps, touched, s);
}

static void fill_q_adam(QAdamArgs& a, double lr_planes, double lr_cplanes, int step, double beta1, double beta2,
                        double eps) {
  const double bc1 = 1.0 - pow(beta1, (double)step);
  a.step_sdf = (float)(lr_planes / bc1);
  a.step_rgb = (float)(lr_cplanes / bc1);
  a.inv_bc2 = (float)(1.0 / sqrt(1.0 - pow(beta2, (double)step)));
  a.adam.beta1 = (float)beta1;
  a.adam.beta2 = (float)beta2;
  a.adam.one_m_beta1 = (float)(1.0 - beta1);
  a.adam.one_m_beta2 = (float)(1.0 - beta2);
  a.adam.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)step));
  a.adam.eps = (float)eps;
}

static int fill_peers(PeerSync& ps, const eslam_peers_t* p) {
  if (!p || p->world < 1 || p->world > MAX_PEERS || p->rank < 0 || p->rank >= p->world || !p->status || !p->local_sync)
    return ESLAM_EINVAL;
  ps.rank = p->rank;
  ps.world = p->world;
  ps.epoch = p->epoch;
  ps.status = p->status;
  ps.local = reinterpret_cast<unsigned long long*>(p->local_sync);
  ps.done_target = (unsigned long long)p->adam_seq * ESLAM_EXCH_CTAS;
  for (int r = 0; r < MAX_PEERS; ++r) {
    ps.flags[r] = r < p->world ? p->flags[r] : nullptr;
    if (r < p->world && !p->flags[r]) return ESLAM_EINVAL;
  }
  return 0;
}

int eslam_exchange_flag_words(void) { return N_SLOTS * MAX_PEERS; }

int eslam_exchange_counters(const eslam_peers_t* peers, const int32_t* counters, int32_t* const* pub, int n,
                            int32_t* norm, eslam_stream_t s) {
  REQUIRE(counters && pub && norm && n >= 1 && n <= 32, "eslam_exchange_counters");
  CounterExchArgs a;
  memset(&a, 0, sizeof(a));
  if (fill_peers(a.ps, peers)) return fail(ESLAM_EINVAL, "eslam_exchange_counters(peers)");
  for (int r = 0; r < a.ps.world; ++r) {
    REQUIRE(pub[r], "eslam_exchange_counters(pub)");
    a.pub[r] = pub[r];
  }
  a.local = counters;
  a.norm = norm;
  a.n = n;
  k_exchange_counters<<<1, 32, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_exchange_counters");
  return 0;
}

// tiles [lo, hi) of the optimiser tail owned by rank r, and the first texel of a tile
static void q_rank_units(const QGroups& qg, int world, int r, int* lo, int* hi) {
  const int total = qg.unit0[4], base = total / world, rem = total % world;
  *lo = r * base + std::min(r, rem);
  *hi = *lo + base + (r < rem ? 1 : 0);
}
static long long q_unit_first_texel(const QGroups& qg, int u) {
  if (u >= qg.unit0[4]) return (long long)qg.t0[3] + qg.n[3];
  int g = 0;
  while (g < 3 && u >= qg.unit0[g + 1]) ++g;
  return (long long)qg.t0[g] + (long long)(u - qg.unit0[g]) * QA_TILE;
}
// lo4[r] = first float4 of rank r's slice of the gradient images; returns the staging row pitch (float4)
static long long q_exchange_slices(const QGroups& qg, int world, long long* lo4) {
  long long smax = 0;
  for (int r = 0; r <= world; ++r) {
    int lo, hi;
    q_rank_units(qg, world, std::min(r, world - 1), &lo, &hi);
    lo4[r] = 4 * q_unit_first_texel(qg, r < world ? lo : hi);
  }
  for (int r = 0; r < world; ++r) smax = std::max(smax, lo4[r + 1] - lo4[r]);
  return (smax + 31) & ~31ll;
}

int eslam_exchange_aux(const eslam_peers_t* peers, float* aux_local, float* const* aux_pub, float* aux_sum, int n_aux,
                       double* auxd_local, double* const* auxd_pub, double* auxd_sum, int n_auxd, eslam_stream_t s) {
  REQUIRE(n_aux >= 0 && n_auxd >= 0 && (n_aux == 0 || (aux_local && aux_pub && aux_sum)) &&
              (n_auxd == 0 || (auxd_local && auxd_pub && auxd_sum)),
          "eslam_exchange_aux");
  AuxExchArgs a;
  memset(&a, 0, sizeof(a));
  if (fill_peers(a.ps, peers)) return fail(ESLAM_EINVAL, "eslam_exchange_aux(peers)");
  for (int r = 0; r < a.ps.world; ++r) {
    if (n_aux) a.aux_pub[r] = aux_pub[r];
    if (n_auxd) a.auxd_pub[r] = auxd_pub[r];
  }
  a.aux_local = aux_local;
  a.aux_sum = aux_sum;
  a.n_aux = n_aux;
  a.auxd_local = auxd_local;
  a.auxd_sum = auxd_sum;
  a.n_auxd = n_auxd;
  k_exchange_aux<<<1, 256, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_exchange_aux");
  return 0;
}

int64_t eslam_q_exchange_stage_floats(const eslam_field_t* f, int world) {
  if (!f || world < 1 || world > MAX_PEERS) return 0;
  QGroups qg;
  if (make_q_groups(f, QA_TILE, &qg)) return 0;
  long long lo4[MAX_PEERS + 1];
  return 4 * (int64_t)world * q_exchange_slices(qg, world, lo4);
}

int eslam_q_adam_exchange(const eslam_peers_t* peers, const eslam_field_t* f, float* const* param,
                          float* const* stage, float* gq_arena, float* grad_arena, float* mc_param, float* exp_avg,
                          float* exp_avg_sq, uint8_t* touched_q, double lr_planes, double lr_cplanes, double lr_dec,
                          int step, double beta1, double beta2, double eps, float* const* dec_pub, float* aux_local,
                          float* const* aux_pub, float* aux_sum, int n_aux, double* auxd_local,
                          double* const* auxd_pub, double* auxd_sum, int n_auxd, eslam_stream_t s) {
  REQUIRE(f && param && stage && gq_arena && grad_arena && exp_avg && exp_avg_sq && touched_q && dec_pub && step >= 1 &&
              n_aux >= 0 && n_auxd >= 0,
          "eslam_q_adam_exchange");
  REQUIRE(n_aux == 0 || (aux_local && aux_pub && aux_sum), "eslam_q_adam_exchange(aux)");
  REQUIRE(n_auxd == 0 || (auxd_local && auxd_pub && auxd_sum), "eslam_q_adam_exchange(auxd)");
  QExchArgs a;
  memset(&a, 0, sizeof(a));
  if (fill_peers(a.ps, peers)) return fail(ESLAM_EINVAL, "eslam_q_adam_exchange(peers)");
  REQUIRE(peers->adam_seq >= 1 && peers->world >= 2, "eslam_q_adam_exchange(adam_seq, world)");
  int rc = make_q_groups(f, QA_TILE, &a.q.qg);
  if (rc) return fail(rc, "eslam_q_adam_exchange(plane layout)");
  const int world = a.ps.world, rank = a.ps.rank;
  a.q.smax = q_exchange_slices(a.q.qg, world, a.lo4);
  int ulo, uhi;
  q_rank_units(a.q.qg, world, rank, &ulo, &uhi);
  REQUIRE(uhi > ulo, "eslam_q_adam_exchange(fewer tiles than ranks)");
  a.q.rank = rank;
  a.q.world = world;
  a.q.unit_lo = ulo;
  a.q.lo4 = a.lo4[rank];
  for (int r = 0; r < world; ++r) {
    REQUIRE(param[r] && stage[r] && dec_pub[r], "eslam_q_adam_exchange(arenas)");
    a.q.peer_p[r] = reinterpret_cast<float4*>(param[r]);
    a.stage[r] = reinterpret_cast<float4*>(stage[r]);
    a.dec_pub[r] = dec_pub[r];
    if (n_aux) a.aux_pub[r] = aux_pub[r];
    if (n_auxd) a.auxd_pub[r] = auxd_pub[r];
  }
  a.q.stage = a.stage[rank];
  a.q.mc_p = reinterpret_cast<float4*>(mc_param);
  a.q.arena4 = a.q.peer_p[rank];
  a.q.gq4 = reinterpret_cast<float4*>(gq_arena);
  a.q.m4 = reinterpret_cast<float4*>(exp_avg);
  a.q.v4 = reinterpret_cast<float4*>(exp_avg_sq);
  a.q.gdec = grad_arena + f->dec_offset;
  a.q.dec = param[rank] + f->dec_offset;
  a.q.touched = touched_q;
  fill_q_adam(a.q, lr_planes, lr_cplanes, step, beta1, beta2, eps);
  a.aux_local = aux_local;
  a.aux_sum = aux_sum;
  a.n_aux = n_aux;
  a.auxd_local = auxd_local;
  a.auxd_sum = auxd_sum;
  a.n_auxd = n_auxd;
  a.dbg = g_debug;
  a.ps.done_target = (unsigned long long)peers->adam_seq * (unsigned long long)(uhi - ulo);
  k_gq_push<<<ESLAM_EXCH_CTAS, EXCH_THREADS, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_q_adam_exchange(push)");
  const unsigned grid = (unsigned)(uhi - ulo);
  if (world <= 2)
    k_q_adam_exchange<2><<<grid, QA_THREADS, 0, S_(s)>>>(a);
  else if (world <= 4)
    k_q_adam_exchange<4><<<grid, QA_THREADS, 0, S_(s)>>>(a);
  else
    k_q_adam_exchange<8><<<grid, QA_THREADS, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_q_adam_exchange");
  DecPeersArgs d;
  memset(&d, 0, sizeof(d));
  d.world = world;
  for (int r = 0; r < world; ++r) d.dec_pub[r] = dec_pub[r];
  d.p = param[rank] + f->dec_offset;
  d.m = exp_avg + f->dec_offset;
  d.v = exp_avg_sq + f->dec_offset;
  const int64_t seg_end[1] = {4};
  const double seg_lr[1] = {lr_dec};
  if (fill_adam(d.adam, 4, seg_end, seg_lr, 1, step, beta1, beta2, eps))
    return fail(ESLAM_EINVAL, "eslam_q_adam_exchange(decoder adam)");
  k_dec_adam_peers<<<(DEC_N + 255) / 256, 256, 0, S_(s)>>>(d);
  CHECK_LAUNCH("eslam_q_adam_exchange(decoders)");
  return 0;
}

int eslam_pose_adam_step(float* poses, float* pose_grad, float* exp_avg, float* exp_avg_sq, int n, int first,
                         double lr_q, double lr_t, int step, double beta1, double beta2, double eps, float* grad7,
                         int apply, eslam_stream_t s) {
  REQUIRE(poses && pose_grad && n > 0 && first >= 0 && first <= n, "eslam_pose_adam_step");
  REQUIRE(!apply || (exp_avg && exp_avg_sq && step >= 1), "eslam_pose_adam_step(state)");
  if (first == n) return 0;
  PoseAdamArgs a;
  a.poses = poses;
  a.pose_grad = pose_grad;
  a.m = exp_avg;
  a.v = exp_avg_sq;
  a.n = n;
  a.first = first;
  const double bc1 = 1.0 - pow(beta1, (double)(step < 1 ? 1 : step));
  a.step_q = (float)(lr_q / bc1);
  a.step_t = (float)(lr_t / bc1);
  a.beta1 = (float)beta1;
  a.beta2 = (float)beta2;
  a.one_m_beta1 = (float)(1.0 - beta1);
  a.one_m_beta2 = (float)(1.0 - beta2);
  a.bc2_sqrt = (float)sqrt(1.0 - pow(beta2, (double)(step < 1 ? 1 : step)));
  a.eps = (float)eps;
  a.grad7 = grad7;
  a.apply = apply;
  const int cnt = n - first;
  k_pose_adam<<<(cnt + 31) / 32, 32, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_pose_adam_step");
  return 0;
}

int eslam_ingest_frame(const uint8_t* bgr, const uint16_t* depth_u16, int H, int W, int crop_edge,
                       double png_depth_scale, double scale, double* color, float* depth, eslam_stream_t s) {
  REQUIRE(bgr && depth_u16 && color && depth && H > 0 && W > 0 && crop_edge >= 0 && 2 * crop_edge < H && 2 * crop_edge < W,
          "eslam_ingest_frame");
  IngestArgs a;
  a.bgr = bgr;
  a.depth = depth_u16;
  a.H = H;
  a.W = W;
  a.edge = crop_edge;
  a.png_depth_scale = (float)png_depth_scale;
  a.scale = (float)scale;
  a.color = color;
  a.out_depth = depth;
  const long long n = (long long)(H - 2 * crop_edge) * (W - 2 * crop_edge);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_ingest_frame<<<(unsigned)blocks, 256, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_ingest_frame");
  return 0;
}

int eslam_ingest_frame_resized(const uint8_t* bgr, int Hs, int Ws, const uint16_t* depth_u16, int H, int W, int crop_edge,
                               double png_depth_scale, double scale, double* color, float* depth, eslam_stream_t s) {
  REQUIRE(bgr && depth_u16 && color && depth && H > 0 && W > 0 && Hs > 0 && Ws > 0 && crop_edge >= 0 &&
              2 * crop_edge < H && 2 * crop_edge < W,
          "eslam_ingest_frame_resized");
  IngestResizeArgs a;
  a.bgr = bgr;
  a.depth = depth_u16;
  a.Hs = Hs;
  a.Ws = Ws;
  a.H = H;
  a.W = W;
  a.edge = crop_edge;
  a.sx = (double)Ws / (double)W;
  a.sy = (double)Hs / (double)H;
  a.png_depth_scale = (float)png_depth_scale;
  a.scale = (float)scale;
  a.color = color;
  a.out_depth = depth;
  const long long n = (long long)(H - 2 * crop_edge) * (W - 2 * crop_edge);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_ingest_resize<<<(unsigned)blocks, 256, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_ingest_frame_resized");
  return 0;
}

int eslam_undistort_u8(const uint8_t* src, uint8_t* dst, int H, int W, double fx, double fy, double cx, double cy,
                       const double* distortion5_host, const double* inv_k9_host, eslam_stream_t s) {
  REQUIRE(src && dst && src != dst && H > 0 && W > 0 && distortion5_host && fx != 0.0 && fy != 0.0, "eslam_undistort_u8");
  UndistortArgs a;
  a.src = src;
  a.dst = dst;
  a.H = H;
  a.W = W;
  if (inv_k9_host) {
    for (int i = 0; i < 9; ++i) a.ir[i] = inv_k9_host[i];
  } else {  // inverse of [[fx 0 cx] [0 fy cy] [0 0 1]]
    const double ir[9] = {1.0 / fx, 0.0, -cx / fx, 0.0, 1.0 / fy, -cy / fy, 0.0, 0.0, 1.0};
    for (int i = 0; i < 9; ++i) a.ir[i] = ir[i];
  }
  a.fx = fx;
  a.fy = fy;
  a.cx = cx;
  a.cy = cy;
  a.k1 = distortion5_host[0];
  a.k2 = distortion5_host[1];
  a.p1 = distortion5_host[2];
  a.p2 = distortion5_host[3];
  a.k3 = distortion5_host[4];
  long long blocks = ((long long)H * W + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_undistort_u8<<<(unsigned)blocks, 256, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_undistort_u8");
  return 0;
}

int eslam_ingest_frame_crop(const uint8_t* bgr, const uint16_t* depth_u16, int H, int W, int Ho, int Wo, int crop_edge,
                            double png_depth_scale, double scale, double* color, float* depth, eslam_stream_t s) {
  REQUIRE(bgr && depth_u16 && color && depth && H > 1 && W > 1 && Ho > 1 && Wo > 1 && crop_edge >= 0 &&
              2 * crop_edge < Ho && 2 * crop_edge < Wo,
          "eslam_ingest_frame_crop");
  IngestCropArgs a;
  a.bgr = bgr;
  a.depth = depth_u16;
  a.H = H;
  a.W = W;
  a.Ho = Ho;
  a.Wo = Wo;
  a.edge = crop_edge;
  a.sy = (double)(H - 1) / (double)(Ho - 1);
  a.sx = (double)(W - 1) / (double)(Wo - 1);
  a.ny = (float)H / (float)Ho;
  a.nx = (float)W / (float)Wo;
  a.png_depth_scale = (float)png_depth_scale;
  a.scale = (float)scale;
  a.color = color;
  a.out_depth = depth;
  const long long n = (long long)(Ho - 2 * crop_edge) * (Wo - 2 * crop_edge);
  long long blocks = (n + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  k_ingest_crop<<<(unsigned)blocks, 256, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_ingest_frame_crop");
  return 0;
}

int eslam_matrix_to_pose(const float* c2w, float* poses, int n, eslam_stream_t s) {
  REQUIRE(c2w && poses && n >= 0, "eslam_matrix_to_pose");
  if (n == 0) return 0;
  k_matrix_to_pose<<<(n + 31) / 32, 32, 0, S_(s)>>>(c2w, poses, n);
  CHECK_LAUNCH("eslam_matrix_to_pose");
  return 0;
}

int eslam_pose_to_matrix(const float* poses, float* c2w, int n, eslam_stream_t s) {
  REQUIRE(c2w && poses && n >= 0, "eslam_pose_to_matrix");
  if (n == 0) return 0;
  k_pose_to_matrix<<<(n + 31) / 32, 32, 0, S_(s)>>>(poses, c2w, n);
  CHECK_LAUNCH("eslam_pose_to_matrix");
  return 0;
}

int eslam_keyframe_overlap(const eslam_camera_t* cam, const float* c2w, const float* depth, const int64_t* pix_idx,
                           int n_rays, const float* t_vals, int n_samples, const float* kf_c2w, int n_keyframes,
                           int32_t* inside, int32_t* n_pts, eslam_stream_t s) {
  REQUIRE(cam && c2w && depth && pix_idx && t_vals && kf_c2w && inside && n_pts && n_rays > 0 && n_samples > 0 &&
              n_keyframes >= 0,
          "eslam_keyframe_overlap");
  if (n_keyframes == 0) return 0;
  OverlapArgs a;
  a.c2w = c2w;
  a.depth = depth;
  a.pix_idx = reinterpret_cast<const long long*>(pix_idx);
  a.t_vals = t_vals;
  a.kf_c2w = kf_c2w;
  a.n_rays = n_rays;
  a.n_samples = n_samples;
  a.K = n_keyframes;
  a.H = cam->H;
  a.W = cam->W;
  a.fx = cam->fx;
  a.fy = cam->fy;
  a.cx = cam->cx;
  a.cy = cam->cy;
  a.inside = inside;
  a.n_pts = n_pts;
  k_keyframe_overlap<<<n_keyframes, OVERLAP_THREADS, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_keyframe_overlap");
  return 0;
}

// ---- the Q form: pre-activated planes (qplane.cuh) ---------------------------------------
int eslam_q_build(const eslam_field_t* f, const float* arena, float* q_arena, eslam_stream_t s) {
  REQUIRE(f && arena && q_arena, "eslam_q_build");
  QBuildArgs a;
  memset(&a, 0, sizeof(a));
  int rc = make_q_groups(f, QB_TPC, &a.qg);
  if (rc) return fail(rc, "eslam_q_build(plane layout)");
  a.arena4 = reinterpret_cast<const float4*>(arena);
  a.dec = arena + f->dec_offset;
  a.q2 = reinterpret_cast<float2*>(q_arena);
  k_q_build<<<a.qg.unit0[4], QB_TPC, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_q_build");
  return 0;
}

int eslam_q_touched_bytes(const eslam_field_t* f) {
  if (!f) return 0;
  long long total = 0;
  for (int i = 0; i < 12; ++i) total = std::max(total, f->plane[i].offset / 32 + (long long)f->plane[i].H * f->plane[i].W);
  return (int)total;  // one flag per texel, indexed by the texel's arena position
}

int eslam_q_adam_planes(const eslam_field_t* f, float* arena, float* gq_arena, float* exp_avg, float* exp_avg_sq,
                        float* grad_arena, uint8_t* touched_q, double lr_planes, double lr_cplanes, int step,
                        double beta1, double beta2, double eps, eslam_stream_t s) {
  REQUIRE(f && arena && gq_arena && exp_avg && exp_avg_sq && grad_arena && touched_q && step >= 1,
          "eslam_q_adam_planes");
  QAdamArgs a;
  memset(&a, 0, sizeof(a));
  int rc = make_q_groups(f, QA_TILE, &a.qg);
  if (rc) return fail(rc, "eslam_q_adam_planes(plane layout)");
  a.arena4 = reinterpret_cast<float4*>(arena);
  a.gq4 = reinterpret_cast<float4*>(gq_arena);
  a.m4 = reinterpret_cast<float4*>(exp_avg);
  a.v4 = reinterpret_cast<float4*>(exp_avg_sq);
  a.gdec = grad_arena + f->dec_offset;
  a.dec = arena + f->dec_offset;
  a.touched = touched_q;
  fill_q_adam(a, lr_planes, lr_cplanes, step, beta1, beta2, eps);
  k_q_adam_planes<<<a.qg.unit0[4], QA_THREADS, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_q_adam_planes");
  return 0;
}

int eslam_render_forward_q(const eslam_field_t* f, const float* q_arena, const float* rays_o, const float* rays_d,
                           const float* z, int n_rays, int n_samples, const int32_t* counters, float* depth,
                           float* rgb, float* sdf, float* act4, uint32_t* actm, eslam_stream_t s) {
  REQUIRE(f && q_arena && rays_o && rays_d && z && depth && rgb && n_rays >= 0, "eslam_render_forward_q");
  REQUIRE((act4 == nullptr) == (actm == nullptr) && (!act4 || sdf), "eslam_render_forward_q(activations)");
  REQUIRE(n_samples >= 1 && n_samples <= ESLAM_MAX_SAMPLES, "eslam_render_forward_q(n_samples)");
  if (n_rays == 0) return 0;
  RenderFwdQArgs a;
  int rc = make_field_k(f, &a.fk);
  if (rc) return fail(rc, "eslam_render_forward_q(field)");
  a.q4 = reinterpret_cast<const float4*>(q_arena);
  a.rays_o = rays_o;
  a.rays_d = rays_d;
  a.z = z;
  a.n_rays = n_rays;
  a.S = n_samples;
  a.counters = counters;
  a.depth = depth;
  a.rgb = rgb;
  a.sdf = sdf;
  a.act4 = reinterpret_cast<float4*>(act4);
  a.actm = actm;
  const int rpb = NP / n_samples;
  k_render_fwd_q<<<(n_rays + rpb - 1) / rpb, NP, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_render_forward_q");
  return 0;
}

// ---- marching cubes, mesh culling (mcubes.cuh) ------------------------------------------------------------------
static int fill_mc(McArgs& a, const float* sdf, int nx, int ny, int nz, double level, const uint8_t* n_tri,
                   const int8_t* tri) {
  if (!(sdf && n_tri && tri && nx >= 2 && ny >= 2 && nz >= 2)) return ESLAM_EINVAL;
  memset(&a, 0, sizeof(a));
  a.sdf = sdf;
  a.nx = nx;
  a.ny = ny;
  a.nz = nz;
  a.level = (float)level;
  a.n_tri = n_tri;
  a.tri = reinterpret_cast<const signed char*>(tri);
  a.n_cells = (long long)(nx - 1) * (ny - 1) * (nz - 1);
  if ((a.n_cells + MC_THREADS - 1) / MC_THREADS > 0x7fffffffLL) return ESLAM_EUNSUPPORTED;
  return 0;
}

int64_t eslam_mc_blocks(int nx, int ny, int nz) {
  if (nx < 2 || ny < 2 || nz < 2) return 0;
  return ((int64_t)(nx - 1) * (ny - 1) * (nz - 1) + MC_THREADS - 1) / MC_THREADS;
}

int eslam_mc_count(const float* sdf, int nx, int ny, int nz, double level, const uint8_t* n_tri, const int8_t* tri,
                   int32_t* block_count, eslam_stream_t s) {
  McArgs a;
  int rc = fill_mc(a, sdf, nx, ny, nz, level, n_tri, tri);
  if (rc || !block_count) return fail(rc ? rc : ESLAM_EINVAL, "eslam_mc_count");
  a.block_count = block_count;
  k_mc_count<<<(unsigned)eslam_mc_blocks(nx, ny, nz), MC_THREADS, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_mc_count");
  return 0;
}

int eslam_mc_emit(const float* sdf, const float* xs, const float* ys, const float* zs, int nx, int ny, int nz,
                  double level, const uint8_t* n_tri, const int8_t* tri, const int64_t* block_base, float* verts,
                  int64_t* keys, eslam_stream_t s) {
  McArgs a;
  int rc = fill_mc(a, sdf, nx, ny, nz, level, n_tri, tri);
  if (rc || !(xs && ys && zs && block_base && verts && keys)) return fail(rc ? rc : ESLAM_EINVAL, "eslam_mc_emit");
  a.xs = xs;
  a.ys = ys;
  a.zs = zs;
  a.block_base = reinterpret_cast<const long long*>(block_base);
  a.verts = verts;
  a.keys = reinterpret_cast<long long*>(keys);
  k_mc_emit<<<(unsigned)eslam_mc_blocks(nx, ny, nz), MC_THREADS, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_mc_emit");
  return 0;
}

int eslam_cull_frame(const float* verts, int64_t n, const float* w2c, const float* depth, const eslam_camera_t* cam,
                     double truncation, int eval_rec, uint8_t* seen, eslam_stream_t s) {
  REQUIRE(verts && w2c && depth && cam && seen && n >= 0, "eslam_cull_frame");
  if (n == 0) return 0;
  CullArgs a;
  a.verts = verts;
  a.n = n;
  a.w2c = w2c;
  a.depth = depth;
  a.H = cam->H;
  a.W = cam->W;
  a.fx = cam->fx;
  a.fy = cam->fy;
  a.cx = cam->cx;
  a.cy = cam->cy;
  a.truncation = (float)truncation;
  a.eval_rec = eval_rec;
  a.seen = seen;
  k_cull_frame<<<(unsigned)((n + 255) / 256), 256, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_cull_frame");
  return 0;
}

int eslam_finalize_loss(const eslam_render_cfg_t* cfg, const int32_t* counters, int tracker_rule, double* loss_acc,
                        float* loss_out, eslam_stream_t s) {
  // `counters` are the normalisers: pass the all-reduced ones (and all-reduced loss_acc) on several GPUs
  REQUIRE(cfg && counters && loss_acc, "eslam_finalize_loss");
  const CfgK k = make_cfg(cfg);
  FinalizeArgs a;
  a.counters = counters;
  a.loss_acc = loss_acc;
  a.loss_out = loss_out;
  a.tracker_rule = tracker_rule;
  a.w_fs = k.w_fs;
  a.w_center = k.w_center;
  a.w_tail = k.w_tail;
  a.w_depth = k.w_depth;
  a.w_color = k.w_color;
  k_finalize_loss<<<1, 32, 0, S_(s)>>>(a);
  CHECK_LAUNCH("eslam_finalize_loss");
  return 0;
}

}  // extern "C"
