/*
 * eslam_b200.h -- C ABI of the B200-native ESLAM render-and-optimise hot path.
 *
 * The reference (MohammadJohari/myslam) is pure Python and has NO FFI; this header is the boundary
 * a maintainer would bind from `src/` with ctypes.  Each entry point names the
 * reference code it replaces (paths relative to the reference root).
 *
 * Conventions
 *  - Every pointer is a DEVICE pointer unless its name ends in `_host`.  The library never
 *    allocates, frees or synchronises; the caller owns every buffer and passes the CUDA stream.
 *  - Return value: 0 on success, otherwise a cudaError_t (>0) or a negative ESLAM_E* code;
 *    eslam_last_error() gives a thread-local message.  No exceptions cross the ABI.
 *  - Parameters live in ONE fp32 arena per process: the 12 feature planes in channels-last
 *    [H][W][32] order (one 128-byte line per texel), then the packed decoders (ESLAM_DEC_FLOATS),
 *    described by eslam_field_t.  Gradients and Adam moments use arenas of the same layout.
 *  - Random numbers are INPUTS (pixel indices, uniforms): the caller draws them with torch so the
 *    stream is the reference's (common.py:108, Renderer.py:59, common.py:59).
 *  - The decoder weights are read through constant memory: eslam_bind_decoders() must be called on
 *    the same stream after every change of the packed decoder block and before any kernel below
 *    that decodes.  This is the only process-global state in the library.
 */
#ifndef ESLAM_B200_H
#define ESLAM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESLAM_ABI_VERSION 1
#define ESLAM_C_DIM 32        /* model.c_dim, configs/ESLAM.yaml:77 */
#define ESLAM_HIDDEN 16       /* decoders.py:39 hidden_size */
#define ESLAM_N_PLANES 12
#define ESLAM_DEC_FLOATS 2700 /* 1329 (sdf, padded to 1332) + 1363 (rgb, padded to 1364) + beta, padded */
#define ESLAM_MAX_SAMPLES 64  /* n_stratified + n_importance per ray */
#define ESLAM_MAX_PEERS 8    /* GPUs of one NVLink/NVSwitch box */
#define ESLAM_EXCH_CTAS 592   /* grid of the push kernel of eslam_q_adam_exchange (4 CTAs per SM) */

#define ESLAM_EINVAL (-1)
#define ESLAM_EUNSUPPORTED (-2)

typedef void* eslam_stream_t; /* cudaStream_t */

/* One feature plane inside the arena. */
typedef struct {
  int64_t offset; /* float offset of texel (0,0) channel 0; multiple of 4 */
  int32_t H, W;   /* rows, cols as in the reference's [1,32,H,W] tensors (ESLAM.py:196-210) */
} eslam_plane_t;

/* Plane order: sdf coarse xy,xz,yz | sdf fine xy,xz,yz | rgb coarse xy,xz,yz | rgb fine xy,xz,yz.
 * xy is [ny][nx], xz is [nz][nx], yz is [nz][ny] (ESLAM.py:200 swap). */
typedef struct {
  eslam_plane_t plane[ESLAM_N_PLANES];
  int64_t dec_offset;   /* float offset of the packed decoders in the arena */
  int64_t n_floats;     /* arena length */
  float bound[3][2];    /* scene bound after ESLAM.load_bound rounding (ESLAM.py:159-173) */
} eslam_field_t;

/* Packed decoder block, float offsets (row-major [out][in] like nn.Linear.weight; decoders.py:46-61);
 * every sub-block starts on a multiple of 4 floats, pads are zero:
 *   sdf:  W1[16][64] @0     b1[16] @1024  W2[16][16] @1040  b2[16] @1296  W3[1][16] @1312  b3[1] @1328
 *   rgb:  W1[16][64] @1332  b1[16] @2356  W2[16][16] @2372  b2[16] @2628  W3[3][16] @2644  b3[3] @2692
 *   beta[1] @2696                                                                                   */
#define ESLAM_DEC_BETA 2696

typedef struct {
  int32_t H, W;              /* image size */
  float fx, fy, cx, cy;
  int32_t H0, H1, W0, W1;    /* crop the pixels are drawn from (Tracker.py:170-173, Mapper.py:318-319) */
} eslam_camera_t;

/* Scalars stay DOUBLE across the ABI: the reference evaluates e.g. `1.5 * truncation` in Python
 * doubles and only then rounds to fp32 (Renderer.py:97), which differs from rounding first. */
typedef struct {
  int32_t n_stratified, n_importance;  /* rendering.* in the yaml */
  double truncation;                   /* model.truncation */
  double w_fs, w_center, w_tail, w_depth, w_color;
} eslam_render_cfg_t;

/* Device-side counters written by eslam_sample_rays / eslam_track_mask (int32[8]):
 * [0] R rays kept  [1] R0 depth-less rays among them  [2] rays in the loss mask
 * [3] front samples [4] center samples [5] tail samples (over masked rays)  [6] reserved
 * [7] tile dispatch ticket of the compaction kernel (its CTAs take their tile in arrival order, so the look-back over
 *     earlier tiles cannot wait for a CTA that is not running) */
#define ESLAM_N_COUNTERS 8
/* The counters BUFFER handed to eslam_sample_rays* / eslam_depth_samples must hold ESLAM_COUNTER_WORDS int32
 * (8-byte aligned): the 8 counters, then one 64-bit word per CTA of the compaction kernel (its totals, published
 * for the CTAs after it).  Everything else only reads the first ESLAM_N_COUNTERS. */
#define ESLAM_MAX_COMPACT_BLOCKS 4096 /* x 256 slots: up to 1 048 576 rays per call */
#define ESLAM_COUNTER_WORDS (ESLAM_N_COUNTERS + 2 * ESLAM_MAX_COMPACT_BLOCKS)
/* Device-side loss accumulators (double[8]): fs, center, tail, depth, colour sums; [5] = loss */
#define ESLAM_N_LOSS 8

const char* eslam_last_error(void);
int eslam_abi_version(void);
/* Profiling aid only (0 in production): bit0 skips the plane-gradient reductions and bit1 the decoder weight
 * gradients inside eslam_loss_backward, to time the phases of the fused kernel separately. */
void eslam_set_debug(int flags);

/* ---- layout: the reference's NCHW planes <-> the channels-last arena ------------------------ */
/* replaces nothing in the reference; it is the price of keeping ESLAM.py's [1,32,H,W] storage. */
int eslam_plane_import(const float* nchw, float* arena, const eslam_plane_t* plane_host, eslam_stream_t s);
int eslam_plane_export(const float* arena, float* nchw, const eslam_plane_t* plane_host, eslam_stream_t s);

/* ---- decoders --------------------------------------------------------------------------------- */
/* Copy the packed decoder block (device, ESLAM_DEC_FLOATS floats) into constant memory. */
int eslam_bind_decoders(const float* dec, eslam_stream_t s);

/* Decoders.forward / get_raw_sdf / get_raw_rgb (src/networks/decoders.py:87-146) on N points.
 * raw[N][4] = (r,g,b,sdf).  flags: bit0 = sdf only (rgb left untouched), bit1 = Mesher.eval_points
 * masking, sdf=-1 outside the OPEN bound box (src/utils/Mesher.py:143-153), bit2 = pts are already
 * normalised to [-1,1] (get_raw_sdf / get_raw_rgb take p_nor). */
int eslam_decode_points(const eslam_field_t* field_host, const float* arena, const float* pts, int64_t n,
                        float* raw, int flags, eslam_stream_t s);

/* Backward of eslam_decode_points for an upstream gradient g_raw[N][4]: grad_arena (+=, may be NULL) and
 * g_pts[N][3] (=, may be NULL).  This is autograd through Decoders.forward (decoders.py:127-146). */
int eslam_decode_backward(const eslam_field_t* field_host, const float* arena, const float* pts, int64_t n,
                          const float* g_raw, float* grad_arena, float* g_pts, eslam_stream_t s);

/* Decoders.sample_plane_feature (decoders.py:64-85): feat[N][64] for the sdf (which=0) or rgb (1) planes
 * from NORMALISED coordinates p_nor[N][3]. */
int eslam_sample_plane_feature(const eslam_field_t* field_host, const float* arena, const float* p_nor, int64_t n,
                               int which, float* feat, eslam_stream_t s);

/* Mesher.get_grid_uniform + eval_points sdf (src/utils/Mesher.py:130-186) for flat grid indices
 * [start, start+count): flat = (iy*nx + ix)*nz + iz, coordinates from the three axis arrays. */
int eslam_grid_sdf(const eslam_field_t* field_host, const float* arena, const float* xs, const float* ys,
                   const float* zs, int nx, int ny, int nz, int64_t start, int64_t count, float* sdf,
                   eslam_stream_t s);

/* eslam_grid_sdf with the mesh bound of Mesher.get_mesh (Mesher.py:206-217: points outside the convex hull of the
 * observed region get sdf = -1) applied in the same pass: hull_planes[n_planes][4] = (nx, ny, nz, d) per hull
 * face with the normal pointing outwards, a point is inside iff n.p + d <= 0 for every face (the geometric
 * predicate `mesh_bound.contains` evaluates for a convex mesh).  Tiles that lie outside entirely are not decoded. */
int eslam_grid_sdf_hull(const eslam_field_t* field_host, const float* arena, const float* xs, const float* ys,
                        const float* zs, int nx, int ny, int nz, int64_t start, int64_t count,
                        const float* hull_planes, int n_planes, float* sdf, eslam_stream_t s);

/* Separable form of the lattice query: on the regular marching-cubes lattice every bilinear tap depends on two
 * lattice indices only, feat(ix,iy,iz) = (Fxy[iy][ix] + Fxz[iz][ix]) + Fyz[iz][iy] per scale (decoders.py:82's
 * order).  eslam_grid_features resamples the sdf decoder's three plane pairs ONCE on the lattice's faces with the
 * arithmetic of the direct query (fxy[ny][nx][64], fxz[nz][nx][64], fyz[nz][ny][64] floats, caller-owned);
 * eslam_grid_sdf_separable then evaluates [start, start+count) from them (three 256-byte reads per voxel instead
 * of 24 corner fetches) and returns values bit-identical to eslam_grid_sdf / eslam_grid_sdf_hull (n_planes may
 * be 0).  The features must be recomputed after the parameters change. */
int eslam_grid_features(const eslam_field_t* field_host, const float* arena, const float* xs, const float* ys,
                        const float* zs, int nx, int ny, int nz, float* fxy, float* fxz, float* fyz, eslam_stream_t s);
int eslam_grid_sdf_separable(const eslam_field_t* field_host, const float* arena, const float* xs, const float* ys,
                             const float* zs, int nx, int ny, int nz, int64_t start, int64_t count, const float* fxy,
                             const float* fxz, const float* fyz, const float* hull_planes, int n_planes, float* sdf,
                             eslam_stream_t s);

/* Factored form of the lattice query (Mesher.py:159-186 with decoders.py:109-118's first layer pulled through the
 * sum of decoders.py:82): the first layer is linear in the summed feature, so
 *   W1 (Fxy + Fxz + Fyz) + b1 = (W1 Fxy + b1) + W1 Fxz + W1 Fyz.
 * eslam_grid_preact resamples the sdf decoder's plane pairs on the lattice's faces and applies W1 there
 * (caller-owned pxy[ny][nx][16], pxz[nx][4][nz][4], pyz[ny][4][nz][4] floats: 16 values per face texel, the z-major
 * faces split into four float4 components so a warp walking z reads coalesced rows); eslam_grid_sdf_factored then
 * evaluates [start, start+count): three 64-byte reads, 32 adds and the 16 -> 16 -> 1 tail per voxel.  Values agree with
 * eslam_grid_sdf / _hull to a few ulp of the first layer's pre-activation (re-associated sum: NOT bit-identical;
 * the tests hold 1e-5 against the direct form and 1e-4 against the reference).  n_planes may be 0.  The faces must
 * be recomputed after the parameters change; layers 2-3 read the decoders bound by eslam_bind_decoders.
 * [iy0, iy1): the lattice rows whose xy / yz face entries are computed (a rank that queries the flat range of a y-slab
 * needs only those; 0, ny for everything); the xz face is always computed whole. */
int eslam_grid_preact(const eslam_field_t* field_host, const float* arena, const float* xs, const float* ys,
                      const float* zs, int nx, int ny, int nz, int iy0, int iy1, float* pxy, float* pxz, float* pyz,
                      eslam_stream_t s);
int eslam_grid_sdf_factored(const eslam_field_t* field_host, const float* xs, const float* ys, const float* zs, int nx,
                            int ny, int nz, int64_t start, int64_t count, const float* pxy, const float* pxz,
                            const float* pyz, const float* hull_planes, int n_planes, float* sdf, eslam_stream_t s);

/* eslam_grid_sdf_factored for WHOLE lattice rows iy in [iy0, iy1) (all ix, iz): a warp keeps one x and 64 z and walks y,
 * so the xz face values stay in registers, the yz values are shared through L1, and the hidden 16 -> 16 layer runs as
 * packed FP32 (fma.rn.f32x2: two voxels per issue slot).  sdf[flat - out_base] for flat = (iy * nx + ix) * nz + iz;
 * out_base <= iy0 * nx * nz.  Same faces, bound and hull tests and the same operation order as
 * eslam_grid_sdf_factored: bit-identical values. */
int eslam_grid_sdf_rows(const eslam_field_t* field_host, const float* xs, const float* ys, const float* zs, int nx,
                        int ny, int nz, int iy0, int iy1, int64_t out_base, const float* pxy, const float* pxz,
                        const float* pyz, const float* hull_planes, int n_planes, float* sdf, eslam_stream_t s);

/* ---- the Q form: the first decoder layer applied to the planes ----------------------------
 * What both loops' iterations run on.  The first layer of decoders.py:87-125 is linear and commutes with the bilinear
 * fetch of decoders.py:64-85:  W1 (sum_planes bilinear(plane)) + b1 = sum_planes bilinear(W1_slice . plane) + b1,
 * so it is applied to the planes once per parameter change instead of to every sample.
 * eslam_q_build writes Q = W1_slice . plane for the 12 planes as 16-channel channels-last images (q_arena: half the
 * plane floats of the parameter arena, plane i at half its float offset; the planes of a (decoder, scale) group must
 * be contiguous in the arena; the decoders are read from the arena).  eslam_render_forward_q is
 * eslam_render_forward / _act on q_arena (layers 2-3 read the decoders bound by eslam_bind_decoders): 64 instead of
 * 128 bytes per corner and no 64 -> 16 layer.  Values differ from the parameter form by the re-association of the
 * first layer's sum (held to 1e-5 in tests/test_gpu_qform.py). */
int eslam_q_build(const eslam_field_t* field_host, const float* arena, float* q_arena, eslam_stream_t s);
int eslam_render_forward_q(const eslam_field_t* field_host, const float* q_arena, const float* rays_o,
                           const float* rays_d, const float* z, int n_rays, int n_samples, const int32_t* counters,
                           float* depth, float* rgb, float* sdf, float* act4, uint32_t* actm, eslam_stream_t s);
/* The tracker's loss + backward to the pose (Tracker.py:192-208) on q_arena: the cached activations must come from
 * eslam_render_forward_q on the same rays; `arena` supplies the decoders.  No forward MLPs, no first-layer backward;
 * the coordinate gradients fetch the Q corners once. */
int eslam_pose_backward_q(const eslam_field_t* field_host, const float* arena, const float* q_arena,
                          const eslam_camera_t* cam, const eslam_render_cfg_t* cfg, const float* rays_o,
                          const float* rays_d, const float* z, const float* gt_depth, const double* gt_color,
                          const int32_t* src, const int64_t* pix_idx, int n_per_img, const uint8_t* ray_mask,
                          const int32_t* counters, int max_rays, const float* sdf, const float* act4,
                          const uint32_t* actm, float* pose_grad, double* loss_acc, eslam_stream_t s);

/* The mapper's fused loss + backward (Mapper.py:110-144,337-349; arguments as eslam_loss_backward) on q_arena: the
 * plane gradients are reduced as 16-channel pre-activation gradients into gq_arena (layout of q_arena, +=),
 * grad_arena's decoder block receives the decoder gradients except dW1 (which eslam_q_adam_planes forms from
 * gq_arena) and beta; the coordinate (pose) gradients come from d pre-activation / d coordinate kept by the gather. */
int eslam_loss_backward_q(const eslam_field_t* field_host, const float* arena, const float* q_arena, float* gq_arena,
                          const eslam_camera_t* cam, const eslam_render_cfg_t* cfg, const float* rays_o,
                          const float* rays_d, const float* z, const float* gt_depth, const double* gt_color,
                          const int32_t* src, const int64_t* pix_idx, int n_per_img, const uint8_t* ray_mask,
                          const int32_t* counters, const int32_t* norm_counters, int max_rays, float* grad_arena,
                          float* pose_grad, double* loss_acc, eslam_stream_t s);

/* eslam_loss_backward_q as one launch of a PAIR (no ray mask, no loss sums): part 1 handles only the tiles (groups of
 * 128 / S consecutive kept rays) whose rays all carry a sensor depth, part 2 only the tiles holding a depth-less ray;
 * part 0 is every tile.  dl_list: the ascending list of depth-less kept rays eslam_sample_rays* left (counters[1]
 * entries), from which part 2 enumerates its tiles with a small persistent grid.  The depth-less rays (Renderer.py:108-134) are the only ones whose samples depend on the
 * current parameters.  Part 1 is launched with programmatic stream serialization: put right behind
 * eslam_importance_samples on the same stream it starts while that kernel runs (behind any other kernel: ordinary stream
 * order); part 2 goes behind the importance pass on another stream and fills part 1's last wave.  The two launches add
 * into the same gradient images / decoder / pose gradients. */
int eslam_loss_backward_q_part(const eslam_field_t* field_host, const float* arena, const float* q_arena,
                               float* gq_arena, const eslam_camera_t* cam, const eslam_render_cfg_t* cfg,
                               const float* rays_o, const float* rays_d, const float* z, const float* gt_depth,
                               const double* gt_color, const int32_t* src, const int64_t* pix_idx, int n_per_img,
                               const int32_t* dl_list, const int32_t* counters, const int32_t* norm_counters,
                               int max_rays, float* grad_arena, float* pose_grad, int part, eslam_stream_t s);

/* The plane half of `optimizer.step()` / `zero_grad()` (Mapper.py:288-306,348-350) from the gradient images.  Per
 * texel: d loss / d plane = W1_slice^T . GQ (consumed in registers by torch.optim.Adam's update, moments in
 * parameter-arena layout), d loss / d W1_slice += GQ (x) plane (added into grad_arena's decoder block, so the decoders
 * then take the ordinary eslam_adam_step), gq_arena zeroed where consumed.  touched_q: eslam_q_touched_bytes() flags
 * (one per texel), zeroed together with the moments: a texel whose gradient row has been zero since then is skipped
 * exactly.  Square root and reciprocal of the update are the approximate instructions (2 ulp). */
int eslam_q_touched_bytes(const eslam_field_t* field_host);
int eslam_q_adam_planes(const eslam_field_t* field_host, float* arena, float* gq_arena, float* exp_avg,
                        float* exp_avg_sq, float* grad_arena, uint8_t* touched_q, double lr_planes, double lr_cplanes,
                        int step, double beta1, double beta2, double eps, eslam_stream_t s);

/* ---- pixel pick, rays, bbox filter, depth-guided samples --------------------------------------- */
/* get_samples + the bbox pre-filter + the depth>0 half of render_batch_ray's sampling
 * (src/common.py:87-153, src/Tracker.py:175-187, src/Mapper.py:322-332, src/utils/Renderer.py:81-106).
 *   pix_idx[n_img*n_per_img]  int64 draws of torch.randint(Hc*Wc, ...) (common.py:108)
 *   c2w[n_img][16]; if poses != NULL, frames >= pose_first take (qw,qx,qy,qz,tx,ty,tz) from poses[n_img][7]
 *   depth[n_img][H][W] f32, color[n_img][H][W][3] f64
 *   u_depth[>=R1][S] uniforms for the perturbation of depth>0 rays, indexed by their ORDINAL among the
 *   depth>0 kept rays (the reference draws rand[R1,S], Renderer.py:59); NULL = no perturbation
 *   t_uni[n_stratified], t_surf[n_importance] = torch.linspace(0,1,n) tables (Renderer.py:85-86)
 * Outputs (capacity n_img*n_per_img rays, compacted in the reference's order):
 *   rays_o/rays_d[R][3], gt_depth[R], gt_color[R][3] f64, src[R] = original slot (frame = src / n_per_img),
 *   z[R][S] (filled for depth>0 rays), dl_list[R0] = compact index of each depth-less ray,
 *   zord[R] = ordinal of the ray among the depth>0 kept rays (its row of u_depth) or -1,
 *   band[R][4] u8 = (#front,#center,#tail,depth>0) per ray, counters (see above; [2..5] filled with the
 *   mapper's depth>0 mask), c2w_out[n_img][16] (may be NULL).  need_depth=1 drops depth<=0 rays (tracker). */
int eslam_sample_rays(const eslam_field_t* field_host, const eslam_camera_t* cam_host,
                      const eslam_render_cfg_t* cfg_host, const int64_t* pix_idx, int n_img, int n_per_img,
                      const float* c2w, const float* poses, int pose_first, const float* depth,
                      const double* color, const float* u_depth, const float* t_uni, const float* t_surf,
                      int need_depth, float* rays_o, float* rays_d, float* gt_depth, double* gt_color,
                      int32_t* src, float* z, int32_t* dl_list, int32_t* zord, uint8_t* band, int32_t* counters,
                      float* c2w_out, eslam_stream_t s);

/* eslam_sample_rays with the window's frames given as two DEVICE tables of n_img per-frame pointers
 * (depth_frames[k] -> [H][W] f32, color_frames[k] -> [H][W][3] f64) instead of one stacked tensor, so the
 * keyframes of a window are read where they live and Mapper.py:268-286's torch.stack of up to 20 full frames
 * (457 MB per call at Replica size) is not needed. */
int eslam_sample_rays_frames(const eslam_field_t* field_host, const eslam_camera_t* cam_host,
                             const eslam_render_cfg_t* cfg_host, const int64_t* pix_idx, int n_img, int n_per_img,
                             const float* c2w, const float* poses, int pose_first,
                             const float* const* depth_frames, const double* const* color_frames,
                             const float* u_depth, const float* t_uni, const float* t_surf, int need_depth,
                             float* rays_o, float* rays_d, float* gt_depth, double* gt_color, int32_t* src, float* z,
                             int32_t* dl_list, int32_t* zord, uint8_t* band, int32_t* counters, float* c2w_out,
                             eslam_stream_t s);

/* Depth-guided z_vals for an already compacted ray list with explicit gt_depth (the first half of
 * render_batch_ray when called through the reference's API, Renderer.py:88-106): fills the z rows of
 * depth>0 rays, lists the others in dl_list, counters[0]=n_rays, counters[1]=R0. */
int eslam_depth_samples(const eslam_render_cfg_t* cfg_host, const float* gt_depth, int n_rays, const float* u_depth,
                        const float* t_uni, const float* t_surf, float* z, int32_t* dl_list, int32_t* zord,
                        int32_t* counters, eslam_stream_t s);

/* The depth-less half of render_batch_ray's sampling (Renderer.py:108-134, common.py:41-77):
 * coarse SDF pass + inverse-cdf resampling for the rays in dl_list; fills their rows of z.
 * The SDF pass reads the Q images of the current parameters (q_arena, eslam_q_build) and the decoders from `arena`
 * (not from the constant bank: no eslam_bind_decoders needed).
 * u_coarse[>=R0][n_stratified], u_fine[>=R0][n_importance] indexed by depth-less ordinal.
 * max_rays bounds the launch (R0 is read on the device from counters[1]). */
int eslam_importance_samples(const eslam_field_t* field_host, const float* arena, const float* q_arena,
                             const eslam_render_cfg_t* cfg_host, const float* rays_o, const float* rays_d,
                             const int32_t* dl_list, const int32_t* counters, int max_rays, const float* u_coarse,
                             const float* u_fine, const float* t_uni, float* z, eslam_stream_t s);

/* ---- render --------------------------------------------------------------------------------- */
/* The second half of Renderer.render_batch_ray (Renderer.py:136-147): points -> decoders -> sdf2alpha ->
 * transmittance weights -> depth[R], rgb[R][3], sdf[R][S] (sdf may be NULL).  n_rays may be an upper bound
 * when counters != NULL (R = counters[0] is read on the device). */
int eslam_render_forward(const eslam_field_t* field_host, const float* arena, const float* rays_o,
                         const float* rays_d, const float* z, int n_rays, int n_samples, const int32_t* counters,
                         float* depth, float* rgb, float* sdf, eslam_stream_t s);

/* eslam_render_forward that also keeps what a pose-only backward pass needs of the forward: act4[R][S][4] =
 * (r, g, b, bit pattern of the sdf decoder's ReLU masks) and actm[R][S] = the rgb decoder's ReLU masks (bit j:
 * hidden-1 unit j active, bit 16+j: hidden-2 unit j active); sdf is required.  24 bytes per sample instead of a
 * second gather of 6 KB and two forward MLPs in eslam_pose_backward_act. */
int eslam_render_forward_act(const eslam_field_t* field_host, const float* arena, const float* rays_o,
                             const float* rays_d, const float* z, int n_rays, int n_samples,
                             const int32_t* counters, float* depth, float* rgb, float* sdf, float* act4,
                             uint32_t* actm, eslam_stream_t s);

/* Backward of the above for arbitrary upstream gradients (autograd through render_batch_ray):
 * g_depth[R], g_rgb[R][3], g_sdf[R][S] (may be NULL) -> grad_arena (+=, planes and decoders incl. beta; may
 * be NULL), g_rays_o/g_rays_d[R][3] (=, both or neither). */
int eslam_render_backward(const eslam_field_t* field_host, const float* arena, const float* rays_o,
                          const float* rays_d, const float* z, int n_rays, int n_samples, const float* g_depth,
                          const float* g_rgb, const float* g_sdf, float* grad_arena, float* g_rays_o,
                          float* g_rays_d, eslam_stream_t s);

/* ---- fused iteration pieces ----------------------------------------------------------------- */
/* Tracker.py:192-195: lower-median outlier mask over the kept rays and the masked band counts.
 * ray_mask[R] u8, counters[2..5]; scratch: max_rays+1 floats (scratch[max_rays] = the median). */
int eslam_track_mask(const float* gt_depth, const float* depth, const uint8_t* band, int max_rays,
                     int32_t* counters, uint8_t* ray_mask, float* scratch, eslam_stream_t s);

/* Forward recompute + the five losses + full backward in one pass
 * (Tracker.py:114-148,192-208; Mapper.py:110-144,337-349).
 *   ray_mask == NULL: mapper rule (sdf/depth terms over depth>0 rays, colour over all rays);
 *   else tracker rule (every term over masked rays).  counters[0] bounds the rays of THIS launch; the loss
 *   normalisers are norm_counters[0,2..5] (NULL = counters; on several GPUs the all-reduced counters, so every
 *   rank differentiates the same global-batch loss).
 *   grad_arena != NULL: plane + decoder + beta gradients are accumulated (+=).
 *   pose_grad != NULL: d loss / d c2w[frame][3][4] accumulated per frame (frame = src / n_per_img).
 *   loss_acc: double[ESLAM_N_LOSS] accumulators (+=), may be NULL. */
int eslam_loss_backward(const eslam_field_t* field_host, const float* arena, const eslam_camera_t* cam_host,
                        const eslam_render_cfg_t* cfg_host, const float* rays_o, const float* rays_d,
                        const float* z, const float* gt_depth, const double* gt_color, const int32_t* src,
                        const int64_t* pix_idx, int n_per_img, const uint8_t* ray_mask, const int32_t* counters,
                        const int32_t* norm_counters, int max_rays, float* grad_arena, float* pose_grad,
                        double* loss_acc, eslam_stream_t s);

/* The tracker's half of eslam_loss_backward (Tracker.py:192-208: losses over the outlier-masked rays, gradient to
 * the 7-dof pose only) on the activations eslam_render_forward_act left for the SAME rays and samples: no feature
 * gather and no forward MLPs, only compositing, the losses, the MLPs' backward-to-input pass and the bilinear
 * coordinate gradients. */
int eslam_pose_backward_act(const eslam_field_t* field_host, const float* arena, const eslam_camera_t* cam_host,
                            const eslam_render_cfg_t* cfg_host, const float* rays_o, const float* rays_d,
                            const float* z, const float* gt_depth, const double* gt_color, const int32_t* src,
                            const int64_t* pix_idx, int n_per_img, const uint8_t* ray_mask, const int32_t* counters,
                            int max_rays, const float* sdf, const float* act4, const uint32_t* actm,
                            float* pose_grad, double* loss_acc, eslam_stream_t s);

/* torch.optim.Adam single-tensor update (torch/optim/adam.py) over a flat arena with up to 4 lr segments
 * [seg_end[i-1], seg_end[i]) (multiples of 4 floats); zeroes the gradient afterwards (the zero_grad of the
 * next iteration).  Bias corrections are computed from `step` (1-based) on the host in doubles. */
int eslam_adam_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                    const int64_t* seg_end_host, const double* seg_lr_host, int n_seg, int step, double beta1,
                    double beta2, double eps, eslam_stream_t s);

/* eslam_adam_step that skips what torch's update leaves unchanged: touched[(n + 127) / 128] (zeroed by the caller
 * together with the moments whenever the optimiser state is re-created) holds one flag per group of 128
 * parameters, raised the first time any gradient of the group is non-zero.  Until then m = v = 0 and
 * p - lr * 0 / (0 + eps) = p, so parameters, moments and the zero gradient of the group are neither read (beyond
 * the gradient) nor written.  Bit-identical to eslam_adam_step. */
int eslam_adam_step_sparse(float* param, float* grad, float* exp_avg, float* exp_avg_sq, int64_t n,
                           const int64_t* seg_end_host, const double* seg_lr_host, int n_seg, int step,
                           double beta1, double beta2, double eps, uint8_t* touched, eslam_stream_t s);

/* cam_pose_to_matrix backward (common.py:169-181 with pytorch3d quaternion_to_matrix) for frames
 * [first, n): pose_grad[n][12] (d loss / d c2w[:3,:4]) -> grad7[n][7] = d loss / d (q,t) (may be NULL);
 * if apply: Adam step on poses[n][7] with lr_q / lr_t and state exp_avg/exp_avg_sq[n][7]; zeroes pose_grad. */
int eslam_pose_adam_step(float* poses, float* pose_grad, float* exp_avg, float* exp_avg_sq, int n, int first,
                         double lr_q, double lr_t, int step, double beta1, double beta2, double eps, float* grad7,
                         int apply, eslam_stream_t s);

/* loss_acc[5] = w_fs*fs/Nf + w_center*ce/Nc + w_tail*ta/Nt + w_color*col/(3*Ncol) + w_depth*dep/Nd (float64
 * like the reference's loss, whose colour term is float64), also written as float to loss_out (may be NULL);
 * loss_acc[0..4] are reset to 0. */
int eslam_finalize_loss(const eslam_render_cfg_t* cfg_host, const int32_t* counters, int tracker_rule,
                        double* loss_acc, float* loss_out, eslam_stream_t s);

/* The arithmetic of BaseDataset.__getitem__ (src/utils/datasets.py:88-95,108-112) after cv2 has decoded the two
 * files and when colour and depth have the same size (Replica; no undistortion, no resize): bgr[H][W][3] uint8 ->
 * color[H-2e][W-2e][3] float64 RGB / 255, depth_u16[H][W] -> depth[H-2e][W-2e] float32 = u16 / png_depth_scale *
 * scale (both in float32, as numpy and torch evaluate them), e = crop_edge.  Bit-exact with the reference's loader;
 * the host ships 8 bytes per pixel instead of 28. */
int eslam_ingest_frame(const uint8_t* bgr, const uint16_t* depth_u16, int H, int W, int crop_edge,
                       double png_depth_scale, double scale, double* color, float* depth, eslam_stream_t s);

/* eslam_ingest_frame for frames whose colour image is larger than the depth image (ScanNet: 1296x968 vs 640x480;
 * datasets.py:92-94 `cv2.resize(color_data, (W, H))` on the float64 image): bgr[Hs][Ws][3] -> colour resized to the
 * depth's [H][W] and cropped.  The resize restates what the opencv-python wheels compute (Intel IPP: float64 weights,
 * row pass then column pass, fma(w, b - a, a)); bit-exact with the reference's `ScanNet` loader run in this container
 * (tests/golden/ingest_scannet.npz).  No undistortion (TUM keeps cv2.undistort on the host, then this path). */
int eslam_ingest_frame_resized(const uint8_t* bgr, int Hs, int Ws, const uint16_t* depth_u16, int H, int W, int crop_edge,
                               double png_depth_scale, double scale, double* color, float* depth, eslam_stream_t s);

/* TUM-shaped frames (datasets.py:83-86): cv2.undistort(img, K, distortion) of the uint8 colour image with the new
 * camera matrix = K, on the device (OpenCV's initUndistortRectifyMap + remap(INTER_LINEAR, BORDER_CONSTANT): source
 * positions in float64 rounded to 1/32 pixel, integer blend).  distortion5_host = (k1, k2, p1, p2, k3);
 * inv_k9_host: the row-major inverse of K as the caller computed it (NULL: the closed form).  dst != src. */
int eslam_undistort_u8(const uint8_t* src, uint8_t* dst, int H, int W, double fx, double fy, double cx, double cy,
                       const double* distortion5_host, const double* inv_k9_host, eslam_stream_t s);

/* eslam_ingest_frame with the loader's `crop_size` step (datasets.py:98-106): the colour image / 255 resized to
 * [Ho][Wo] like F.interpolate(mode='bilinear', align_corners=True) on float64 (torch's CPU arithmetic, bit for bit),
 * the depth like F.interpolate(mode='nearest'), then crop_edge.  color [Ho-2e][Wo-2e][3], depth [Ho-2e][Wo-2e]. */
int eslam_ingest_frame_crop(const uint8_t* bgr, const uint16_t* depth_u16, int H, int W, int Ho, int Wo, int crop_edge,
                            double png_depth_scale, double scale, double* color, float* depth, eslam_stream_t s);

/* matrix_to_cam_pose / cam_pose_to_matrix (common.py:155-181 over pytorch3d 0.7.1 matrix_to_quaternion /
 * quaternion_to_matrix) for n cameras: c2w[n][16] row-major <-> poses[n][7] = (qw,qx,qy,qz,tx,ty,tz), evaluated in
 * torch's operation order.  Used once per optimize_mapping call for the window's poses (Mapper.py:289,352-362). */
int eslam_matrix_to_pose(const float* c2w, float* poses, int n, eslam_stream_t s);
int eslam_pose_to_matrix(const float* poses, float* c2w, int n, eslam_stream_t s);

/* Mapper.keyframe_selection_overlap (Mapper.py:146-203) up to `percent_inside`: the n_rays pixels pix_idx
 * (draws of randint(H*W), common.py:108) of the current frame that have depth > 0 are lifted to n_samples points
 * each between 0.8*d and d+0.5 (t_vals = linspace(0,1,n_samples)), projected into every keyframe kf_c2w[k]
 * (row-major 4x4 c2w, inverted here) and counted when they fall inside the image with a 20-pixel margin, in
 * front of the camera.  inside[k] = count for keyframe k, n_pts[0] = points tested; the reference's
 * percent_inside[k] = inside[k] / n_pts.  The caller keeps nonzero(inside) in randperm order (Mapper.py:205-209). */
int eslam_keyframe_overlap(const eslam_camera_t* cam_host, const float* c2w, const float* depth,
                           const int64_t* pix_idx, int n_rays, const float* t_vals, int n_samples,
                           const float* kf_c2w, int n_keyframes, int32_t* inside, int32_t* n_pts, eslam_stream_t s);

/* ---- mesh extraction around the grid query (Mesher.get_mesh, src/utils/Mesher.py:188-264) ---------------------------
 * Marching cubes over the SDF lattice sdf[(iy*nx + ix)*nz + iz] (what eslam_grid_sdf* writes) ON THE DEVICE, replacing
 * the D2H of the 1.3 GB volume + skimage.measure.marching_cubes (Mesher.py:219-243).  Case tables: n_tri[256] uint8 and
 * tri[256][15] int8 edge ids (myslam_b200/mc_tables.py: generated, conventions there; corner "inside" <=> value < level,
 * triangles wound with the normal towards increasing values).  Two passes without per-cell storage:
 *   eslam_mc_count  block_count[eslam_mc_blocks()] = triangles of each block of 256 consecutive cells
 *   eslam_mc_emit   block_base = exclusive scan of block_count (int64); writes the triangle soup verts[3T][3] (world
 *                   coordinates: lattice coordinate + linear interpolation along the crossed edge, Mesher.py:245) and
 *                   keys[3T] = 3 * flat(lower lattice corner) + axis, the lattice edge of each vertex (equal keys =
 *                   the same vertex: what welding needs). */
int64_t eslam_mc_blocks(int nx, int ny, int nz);
int eslam_mc_count(const float* sdf, int nx, int ny, int nz, double level, const uint8_t* n_tri, const int8_t* tri,
                   int32_t* block_count, eslam_stream_t s);
int eslam_mc_emit(const float* sdf, const float* xs, const float* ys, const float* zs, int nx, int ny, int nz,
                  double level, const uint8_t* n_tri, const int8_t* tri, const int64_t* block_base, float* verts,
                  int64_t* keys, eslam_stream_t s);

/* One frame of cull_mesh (src/tools/cull_mesh.py:58-100): seen[v] |= vertex v projects into the frame (in front of the
 * camera, inside the image; with eval_rec also not more than `truncation` behind the frame's depth, sampled
 * bilinearly like F.grid_sample(zeros, align_corners=True)).  w2c[16]: row-major inverse of the frame's c2w.  The
 * caller loops over the frames and drops the faces whose three vertices were never seen (cull_mesh.py:102-105). */
int eslam_cull_frame(const float* verts, int64_t n, const float* w2c, const float* depth,
                     const eslam_camera_t* cam_host, double truncation, int eval_rec, uint8_t* seen, eslam_stream_t s);

/* ---- multi-GPU mapping over peer memory (new in this build; the reference is single-GPU) ---------------------
 * One process per GPU; every rank owns ONE symmetric (NVLink peer-mapped) allocation holding, at identical
 * offsets: the parameter arena, a gradient-image staging block of eslam_q_exchange_stage_floats() floats, two
 * published copies of the aux blocks and counters, and a zero-initialised flag block of
 * eslam_exchange_flag_words() uint32.  The gradient arena stays in ordinary device memory.  Pointer tables
 * `x_host[r]` are rank r's copies as mapped into THIS process (index `rank` = local).
 * `epoch` must increase by one with every exchange call (either kind) and be the same on every rank;
 * `adam_seq` counts the eslam_q_adam_exchange calls (1-based); `local_sync` is a local zero-initialised
 * uint64[2]; `status` is a local int32 that becomes non-zero if a peer did not arrive within 4 s. */
typedef struct {
  int32_t rank, world;
  uint32_t epoch;
  uint32_t pad_;
  uint32_t* flags[ESLAM_MAX_PEERS];
  int32_t* status;
  uint64_t* local_sync;
  uint64_t adam_seq;
} eslam_peers_t;

int eslam_exchange_flag_words(void);

/* norm[i] = sum over ranks of counters[i], i < n <= 8: the all-reduce of the loss normalisers
 * (front/center/tail/depth/ray counts) that makes every rank differentiate the global-batch loss
 * (Mapper.py:110-144,337-346 evaluated over the union of the ranks' rays).  `counters` is this rank's block,
 * pub_host[r] rank r's published copy (alternate between two copies from call to call). */
int eslam_exchange_counters(const eslam_peers_t* peers_host, const int32_t* counters, int32_t* const* pub_host,
                            int n, int32_t* norm, eslam_stream_t s);

/* aux_sum[i] = sum over ranks of aux_local[i] (float, the [frames][12] pose-gradient block) and auxd_sum likewise
 * (double, the loss terms): published into *_pub_host[rank] (alternate between two copies from call to call), handshake,
 * summed in rank order on every rank, local blocks zeroed.  One CTA: meant for a side stream, so the pose step and the
 * next iteration's ray sampling overlap eslam_q_adam_exchange. */
int eslam_exchange_aux(const eslam_peers_t* peers_host, float* aux_local, float* const* aux_pub_host, float* aux_sum,
                       int n_aux, double* auxd_local, double* const* auxd_pub_host, double* auxd_sum, int n_auxd,
                       eslam_stream_t s);

/* Floats of the gradient-image staging block of one rank for `world` ranks. */
int64_t eslam_q_exchange_stage_floats(const eslam_field_t* field_host, int world);

/* `optimizer.step()` / `zero_grad()` of Mapper.py:348-350 plus the gradient all-reduce of a ray-sharded mapping, over
 * peer memory.  Ownership is by tiles of eslam_q_adam_planes: rank r owns a contiguous slice of the gradient images.
 * Three kernels: (1) every rank stores the peers' slices of its gradient image gq_arena into their staging rows (P2P
 * stores) and zeroes them; (2) after a handshake rank r runs eslam_q_adam_planes on its tiles with the gradient rows
 * summed over its own image and the staged rows in rank order (exp_avg / exp_avg_sq / touched_q are local, only the
 * owned slice is touched) and stores the new texels into every rank's arena (P2P stores, or one multimem.st when
 * mc_param != NULL); its decoder gradients (dW1 of the owned tiles + the backward's other decoder gradients in
 * grad_arena) are published into dec_pub_host[rank] (ESLAM_DEC_FLOATS floats, alternate between two copies from call
 * to call) and zeroed; a closing handshake makes everything visible; (3) every rank sums all ranks' published decoder
 * gradients in rank order and takes the decoders' Adam step (replicated).  The local aux (float, pose gradients) /
 * auxd (double, loss terms) blocks are published, zeroed, and summed over all ranks into aux_sum / auxd_sum.
 * The call is identical on every rank.  All-zero groups of 128 floats are not sent (the staging must start zeroed; the
 * owner clears what it consumes). */
int eslam_q_adam_exchange(const eslam_peers_t* peers_host, const eslam_field_t* field_host, float* const* param_host,
                          float* const* stage_host, float* gq_arena, float* grad_arena, float* mc_param,
                          float* exp_avg, float* exp_avg_sq, uint8_t* touched_q, double lr_planes, double lr_cplanes,
                          double lr_dec, int step, double beta1, double beta2, double eps,
                          float* const* dec_pub_host, float* aux_local, float* const* aux_pub_host, float* aux_sum,
                          int n_aux, double* auxd_local, double* const* auxd_pub_host, double* auxd_sum, int n_auxd,
                          eslam_stream_t s);

#ifdef __cplusplus
}
#endif
#endif /* ESLAM_B200_H */
