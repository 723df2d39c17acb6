// Pixel pick -> rays -> bbox filter -> ordered compaction -> depth-guided samples, the importance
// resampling of depth-less rays, and the tracker's median outlier mask.
// Reference: src/common.py:41-153, src/Tracker.py:175-195, src/Mapper.py:322-332, src/utils/Renderer.py:46-134.
// All arithmetic that decides which rays/samples exist uses explicit round-to-nearest intrinsics in the
// reference's operation order (no FMA contraction), so ray selection and z_vals are bit-exact.
#pragma once
#include "field.cuh"
#include "render.cuh"

namespace eslam {

constexpr int SB = 256;  // threads per block of the sampling kernels

struct SampleArgs {
  FieldK fk;
  int H, W, H0, W0, Wc, HWc;
  float fx, fy, cx, cy;
  int n_strat, n_imp;
  float tr, tr15, tr3, tr04;
  const long long* pix_idx;
  int n_img, n_per_img;
  const float* c2w;
  const float* poses;
  int pose_first;
  const float* depth;               // [n_img][H][W], or NULL when the per-frame tables below are used
  const double* color;              // [n_img][H][W][3]
  const float* const* depth_tab;    // [n_img] device pointers to [H][W] frames (keyframes stay where they are)
  const double* const* color_tab;   // [n_img] device pointers to [H][W][3] frames
  const float* u_depth;
  const float *t_uni, *t_surf;
  int need_depth;
  float *rays_o, *rays_d, *gt_depth;
  double* gt_color;
  int* src;
  float* z;
  int* dl_list;
  int* zord;
  unsigned char* band;
  int* counters;
  unsigned long long* agg;  // per-CTA totals for the look-back (counters + ESLAM_N_COUNTERS, zeroed per launch)
  float* c2w_out;
};

// torch.max / torch.min propagate NaN
__device__ __forceinline__ float max_nan(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fffffff) : fmaxf(a, b); }
__device__ __forceinline__ float min_nan(float a, float b) { return (a != a || b != b) ? __int_as_float(0x7fffffff) : fminf(a, b); }

// quaternion_to_matrix (pytorch3d 0.7.1) in the order torch evaluates it; row-major R[9]
__device__ __forceinline__ void quat_to_rot(const float* q, float* R) {
  const float r = q[0], i = q[1], j = q[2], k = q[3];
  const float n = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(r, r), __fmul_rn(i, i)), __fmul_rn(j, j)), __fmul_rn(k, k));
  const float s = __fdiv_rn(2.0f, n);
#define MUL __fmul_rn
#define ADD __fadd_rn
#define SUB __fsub_rn
  R[0] = SUB(1.0f, MUL(s, ADD(MUL(j, j), MUL(k, k))));
  R[1] = MUL(s, SUB(MUL(i, j), MUL(k, r)));
  R[2] = MUL(s, ADD(MUL(i, k), MUL(j, r)));
  R[3] = MUL(s, ADD(MUL(i, j), MUL(k, r)));
  R[4] = SUB(1.0f, MUL(s, ADD(MUL(i, i), MUL(k, k))));
  R[5] = MUL(s, SUB(MUL(j, k), MUL(i, r)));
  R[6] = MUL(s, SUB(MUL(i, k), MUL(j, r)));
  R[7] = MUL(s, ADD(MUL(j, k), MUL(i, r)));
  R[8] = SUB(1.0f, MUL(s, ADD(MUL(i, i), MUL(j, j))));
#undef MUL
#undef ADD
#undef SUB
}

struct RayEval {
  float o[3], d[3], depth;
  bool keep;
};

// bbox exit distance: min over axes of max over {lo,hi} of (bound-o)/d  (Tracker.py:177-180)
__device__ __forceinline__ float bbox_exit(const FieldK& fk, const float* o, const float* d) {
  float t = 0.f;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const float tl = __fdiv_rn(__fsub_rn(fk.lo[a], o[a]), d[a]);
    const float th = __fdiv_rn(__fsub_rn(fk.hi[a], o[a]), d[a]);
    const float m = max_nan(tl, th);
    t = (a == 0) ? m : min_nan(t, m);
  }
  return t;
}

__device__ __forceinline__ RayEval eval_ray(const SampleArgs& a, int slot) {
  RayEval r;
  const int frame = slot / a.n_per_img;
  const long long pix = a.pix_idx[slot];
  const int pr = (int)(pix / a.Wc), pc = (int)(pix - (long long)pr * a.Wc);
  const float pi = (float)(a.W0 + pc), pj = (float)(a.H0 + pr);
  const float* dframe = a.depth_tab ? a.depth_tab[frame] : a.depth + (long long)frame * a.H * a.W;
  r.depth = dframe[(long long)(a.H0 + pr) * a.W + (a.W0 + pc)];
  float Rm[9], t[3];
  if (a.poses && frame >= a.pose_first) {
    const float* p = a.poses + frame * 7;
    quat_to_rot(p, Rm);
    t[0] = p[4];
    t[1] = p[5];
    t[2] = p[6];
  } else {
    const float* m = a.c2w + frame * 16;
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      Rm[x * 3 + 0] = m[x * 4 + 0];
      Rm[x * 3 + 1] = m[x * 4 + 1];
      Rm[x * 3 + 2] = m[x * 4 + 2];
      t[x] = m[x * 4 + 3];
    }
  }
  const float d0 = __fdiv_rn(__fsub_rn(pi, a.cx), a.fx);
  const float d1 = -__fdiv_rn(__fsub_rn(pj, a.cy), a.fy);
  const float d2 = -1.0f;
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    r.d[x] = __fadd_rn(__fadd_rn(__fmul_rn(d0, Rm[x * 3 + 0]), __fmul_rn(d1, Rm[x * 3 + 1])), __fmul_rn(d2, Rm[x * 3 + 2]));
    r.o[x] = t[x];
  }
  const float te = bbox_exit(a.fk, r.o, r.d);
  r.keep = (te >= r.depth) && (!a.need_depth || r.depth > 0.f);
  return r;
}

// block-wide exclusive scan of two small counts packed as (lo 16 bits | hi 16 bits); returns exclusive
// prefix for this thread and the block total
__device__ __forceinline__ void block_scan2(int f0, int f1, int& ex0, int& ex1, int& tot0, int& tot1, int* sh /*[2*SB/32+2]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const unsigned b0 = __ballot_sync(0xffffffffu, f0), b1 = __ballot_sync(0xffffffffu, f1);
  const unsigned lt = (1u << lane) - 1u;
  if (lane == 0) {
    sh[warp] = __popc(b0);
    sh[SB / 32 + warp] = __popc(b1);
  }
  __syncthreads();
  int base0 = 0, base1 = 0, t0 = 0, t1 = 0;
  for (int w2 = 0; w2 < SB / 32; ++w2) {
    const int c0 = sh[w2], c1 = sh[SB / 32 + w2];
    if (w2 < warp) {
      base0 += c0;
      base1 += c1;
    }
    t0 += c0;
    t1 += c1;
  }
  ex0 = base0 + __popc(b0 & lt);
  ex1 = base1 + __popc(b1 & lt);
  tot0 = t0;
  tot1 = t1;
  __syncthreads();
}

// Ordered compaction across CTAs without re-evaluating the preceding slots: CTA b publishes its two totals in
// agg[b] (one 64-bit word, bit 0 = ready; zeroed by the host entry point before the launch) and sums the words of
// the CTAs before it, spinning on those not yet published.  b is the CTA's TILE, taken from a dispatch ticket
// (tile_ticket) rather than from blockIdx: a tile only ever waits for tiles whose CTAs already run, whatever order
// the hardware dispatches CTAs in, so the chain cannot deadlock even when the grid exceeds what is resident.
__device__ __forceinline__ int tile_ticket(int* ticket, int* sh_slot) {
  if (threadIdx.x == 0) *sh_slot = atomicAdd(ticket, 1);
  __syncthreads();
  return *sh_slot;
}

__device__ __forceinline__ void lookback2(unsigned long long* agg, int b, int tot0, int tot1, int* sh /*[2*SB/32+2]*/,
                                          int& base0, int& base1) {
  if (threadIdx.x == 0) {
    const unsigned long long w = ((unsigned long long)(unsigned)tot0 << 32) | ((unsigned long long)(unsigned)tot1 << 1) | 1ull;
    __threadfence();
    atomicExch(agg + b, w);
  }
  int p0 = 0, p1 = 0;
  for (int j = threadIdx.x; j < b; j += SB) {
    unsigned long long w;
    do {
      w = *reinterpret_cast<volatile unsigned long long*>(agg + j);
    } while (!(w & 1ull));
    p0 += (int)(w >> 32);
    p1 += (int)((w & 0xffffffffull) >> 1);
  }
  p0 = (int)warp_sum((float)p0);  // counts < 2^24: exact in fp32
  p1 = (int)warp_sum((float)p1);
  if ((threadIdx.x & 31) == 0) {
    sh[threadIdx.x >> 5] = p0;
    sh[SB / 32 + (threadIdx.x >> 5)] = p1;
  }
  __syncthreads();
  base0 = base1 = 0;
  for (int w2 = 0; w2 < SB / 32; ++w2) {
    base0 += sh[w2];
    base1 += sh[SB / 32 + w2];
  }
  __syncthreads();
}

// Depth-guided samples of one ray with depth>0 (Renderer.py:94-106 + perturbation :46-61), one WARP per ray:
// lane l owns elements l and l+32 of the concatenated [free(n_strat) | surface(n_imp)] list, finds their
// positions in the sorted order by counting (both lists are ascending, so torch.sort of the concatenation is a
// merge; ties: free before surface), then jitters inside the midpoints' intervals and classifies the sdf band.
// zs: this warp's 64-float shared scratch.  Returns band counts on every lane.
__device__ __forceinline__ void depth_guided_z_warp(float d, int n_strat, int n_imp, const float* __restrict__ t_uni,
                                                    const float* __restrict__ t_surf, float tr15, float tr3,
                                                    bool has_u, float u_lo, float u_hi, float* __restrict__ zout,
                                                    float tr, float tr04, float* zs, int& nf, int& nc, int& nt) {
  const int lane = threadIdx.x & 31;
  const int S = n_strat + n_imp;
  const float d12 = __fmul_rn(1.2f, d);
  const float dsurf = __fsub_rn(d, tr15);
  float val[2];
  bool is_free[2], live[2];
  int idx[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int e = lane + 32 * h;
    live[h] = e < S;
    is_free[h] = e < n_strat;
    idx[h] = is_free[h] ? e : e - n_strat;
    val[h] = 0.f;
    if (live[h])
      val[h] = is_free[h] ? __fmul_rn(d12, t_uni[idx[h]]) : __fadd_rn(dsurf, __fmul_rn(tr3, t_surf[idx[h]]));
  }
  int below[2] = {0, 0};  // free element: #surface < it ; surface element: filled from the ballots below
  for (int j = 0; j < n_imp; ++j) {
    const float sj = __fadd_rn(dsurf, __fmul_rn(tr3, t_surf[j]));
    const unsigned b0 = __ballot_sync(0xffffffffu, live[0] && is_free[0] && val[0] <= sj);
    const unsigned b1 = __ballot_sync(0xffffffffu, live[1] && is_free[1] && val[1] <= sj);
    const int n_le = __popc(b0) + __popc(b1);  // #free <= surface_j
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      if (is_free[h]) {
        below[h] += (sj < val[h]);
      } else if (idx[h] == j) {
        below[h] = n_le;
      }
    }
  }
#pragma unroll
  for (int h = 0; h < 2; ++h)
    if (live[h]) zs[idx[h] + below[h]] = val[h];
  __syncwarp();
  nf = nc = nt = 0;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int k = lane + 32 * h;
    bool f0 = false, f1 = false, f2 = false;
    if (k < S) {
      const float zc = zs[k];
      const float lower = k > 0 ? __fmul_rn(0.5f, __fadd_rn(zc, zs[k - 1])) : zc;
      const float upper = k < S - 1 ? __fmul_rn(0.5f, __fadd_rn(zs[k + 1], zc)) : zc;
      const float zp = has_u ? __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), h ? u_hi : u_lo)) : zc;
      zout[k] = zp;
      const int b = sdf_band(zp, d, tr, tr04);
      f0 = b == 0;
      f1 = b == 1;
      f2 = b == 2;
    }
    nf += __popc(__ballot_sync(0xffffffffu, f0));
    nc += __popc(__ballot_sync(0xffffffffu, f1));
    nt += __popc(__ballot_sync(0xffffffffu, f2));
  }
  __syncwarp();
}

__global__ void __launch_bounds__(SB) k_sample_rays(const __grid_constant__ SampleArgs a) {
  __shared__ int sh[2 * SB / 32 + 2];
  __shared__ int s_tile;
  const int N = a.n_img * a.n_per_img;
  const int tile = tile_ticket(a.counters + 7, &s_tile);  // counters[7]: zeroed with the counters before the launch
  const int start = tile * SB;
  const int slot = start + threadIdx.x;
  RayEval e;
  e.keep = false;
  e.depth = 0.f;
  if (slot < N) e = eval_ray(a, slot);
  const int fk_ = e.keep ? 1 : 0;
  const int fd_ = (e.keep && !(e.depth > 0.f)) ? 1 : 0;
  int ex0, ex1, tot0, tot1, base0, base1;
  block_scan2(fk_, fd_, ex0, ex1, tot0, tot1, sh);
  // kept / depth-less counts of all preceding slots: the order-preserving compaction of Tracker.py:184-187 /
  // Mapper.py:329-332 (boolean indexing) across CTAs
  lookback2(a.agg, tile, tot0, tot1, sh, base0, base1);
  if (e.keep) {
    const int r = base0 + ex0;
    const int r0 = base1 + ex1;  // ordinal among depth-less rays
    const int frame = slot / a.n_per_img;
    const long long pix = a.pix_idx[slot];
    const int pr = (int)(pix / a.Wc), pc = (int)(pix - (long long)pr * a.Wc);
#pragma unroll
    for (int x = 0; x < 3; ++x) {
      a.rays_o[r * 3 + x] = e.o[x];
      a.rays_d[r * 3 + x] = e.d[x];
    }
    a.gt_depth[r] = e.depth;
    const double* cframe = a.color_tab ? a.color_tab[frame] : a.color + (long long)frame * a.H * a.W * 3;
    const double* cp = cframe + ((long long)(a.H0 + pr) * a.W + (a.W0 + pc)) * 3;
    a.gt_color[(long long)r * 3 + 0] = cp[0];
    a.gt_color[(long long)r * 3 + 1] = cp[1];
    a.gt_color[(long long)r * 3 + 2] = cp[2];
    a.src[r] = slot;
    if (e.depth > 0.f) {
      a.zord[r] = r - r0;  // ordinal among depth>0 kept rays = row of u_depth
    } else {
      a.zord[r] = -1;
      a.dl_list[r0] = r;
      *reinterpret_cast<uchar4*>(a.band + r * 4) = make_uchar4(0, 0, 0, 0);
    }
  }
  if (threadIdx.x == 0) {
    if (tot0) atomicAdd(a.counters + 0, tot0);
    if (tot1) atomicAdd(a.counters + 1, tot1);
    if (tot0 - tot1) atomicAdd(a.counters + 2, tot0 - tot1);
  }
  if (a.c2w_out && slot < N && (slot % a.n_per_img) == 0) {
    const int frame = slot / a.n_per_img;
    float Rm[9], t[3];
    float* out = a.c2w_out + frame * 16;
    if (a.poses && frame >= a.pose_first) {
      const float* p = a.poses + frame * 7;
      quat_to_rot(p, Rm);
      t[0] = p[4];
      t[1] = p[5];
      t[2] = p[6];
      for (int x = 0; x < 3; ++x) {
        out[x * 4 + 0] = Rm[x * 3 + 0];
        out[x * 4 + 1] = Rm[x * 3 + 1];
        out[x * 4 + 2] = Rm[x * 3 + 2];
        out[x * 4 + 3] = t[x];
      }
      out[12] = out[13] = out[14] = 0.f;
      out[15] = 1.f;
    } else {
      for (int x = 0; x < 16; ++x) out[x] = a.c2w[frame * 16 + x];
    }
  }
}

// Depth-guided z_vals, one warp per compact ray (grid covers an upper bound of rays; R is read on the device).
// zord[r] = row of u_depth (ordinal among depth>0 rays) or -1 for a depth-less ray.  Fills z rows, band[r]
// and adds the band totals to counters[3..5].
struct RaySampleArgs {
  int n_strat, n_imp;
  float tr, tr15, tr3, tr04;
  const float* gt_depth;
  const int* zord;
  int max_rays;
  const float* u_depth;
  const float *t_uni, *t_surf;
  float* z;
  unsigned char* band;
  int* counters;
};

__global__ void __launch_bounds__(SB) k_ray_samples(const __grid_constant__ RaySampleArgs a) {
  __shared__ float s_zs[SB / 32][64];
  __shared__ float s_tu[ESLAM_MAX_SAMPLES], s_ts[16];
  __shared__ int s_cnt[3];
  const int R = min(a.counters[0], a.max_rays);
  if (blockIdx.x * (SB / 32) >= R) return;
  if (threadIdx.x < a.n_strat) s_tu[threadIdx.x] = a.t_uni[threadIdx.x];
  if (threadIdx.x < a.n_imp) s_ts[threadIdx.x] = a.t_surf[threadIdx.x];
  if (threadIdx.x < 3) s_cnt[threadIdx.x] = 0;
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r = blockIdx.x * (SB / 32) + warp;
  const int S = a.n_strat + a.n_imp;
  if (r < R) {
    const int ord = a.zord[r];
    if (ord >= 0) {
      float u_lo = 0.f, u_hi = 0.f;
      if (a.u_depth) {
        const float* up = a.u_depth + (long long)ord * S;
        if (lane < S) u_lo = up[lane];
        if (lane + 32 < S) u_hi = up[lane + 32];
      }
      int nf, nc, nt;
      depth_guided_z_warp(a.gt_depth[r], a.n_strat, a.n_imp, s_tu, s_ts, a.tr15, a.tr3, a.u_depth != nullptr, u_lo,
                          u_hi, a.z + (long long)r * S, a.tr, a.tr04, s_zs[warp], nf, nc, nt);
      if (lane == 0) {
        if (a.band) *reinterpret_cast<uchar4*>(a.band + r * 4) = make_uchar4(nf, nc, nt, 1);
        atomicAdd(&s_cnt[0], nf);
        atomicAdd(&s_cnt[1], nc);
        atomicAdd(&s_cnt[2], nt);
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < 3 && s_cnt[threadIdx.x]) atomicAdd(a.counters + 3 + threadIdx.x, s_cnt[threadIdx.x]);
}

// Ordinals for an explicit, already compacted ray list (render_batch_ray through the reference's API):
// zord / dl_list / counters[0..2] from gt_depth alone.
struct DepthOrdArgs {
  const float* gt_depth;
  int n_rays;
  int* zord;
  int* dl_list;
  int* counters;
  unsigned long long* agg;
};

__global__ void __launch_bounds__(SB) k_depth_ordinals(const __grid_constant__ DepthOrdArgs a) {
  __shared__ int sh[2 * SB / 32 + 2];
  __shared__ int s_tile;
  const int tile = tile_ticket(a.counters + 7, &s_tile);
  const int start = tile * SB;
  const int r = start + threadIdx.x;
  const bool in = r < a.n_rays;
  const float d = in ? a.gt_depth[r] : 1.f;
  const int fd_ = (in && !(d > 0.f)) ? 1 : 0;
  int ex0, ex1, tot0, tot1, base0, base1;
  block_scan2(fd_, 0, ex0, ex1, tot0, tot1, sh);
  lookback2(a.agg, tile, tot0, tot1, sh, base0, base1);
  if (in) {
    const int r0 = base0 + ex0;
    if (d > 0.f) {
      a.zord[r] = r - r0;
    } else {
      a.zord[r] = -1;
      a.dl_list[r0] = r;
    }
  }
  if (threadIdx.x == 0) {
    if (tot0) atomicAdd(a.counters + 1, tot0);
    const int here = min(SB, a.n_rays - start);
    if (here - tot0) atomicAdd(a.counters + 2, here - tot0);
    if (tile == 0) a.counters[0] = a.n_rays;
  }
}

// ---------------------------------------------------------------------------------------------------
// importance resampling of depth-less rays  (Renderer.py:108-134, common.py:41-77)
// ---------------------------------------------------------------------------------------------------
struct ImportanceArgs {
  FieldK fk;
  const float4* arena4;  // decoders (layers 2-3 of the sdf decoder, beta)
  const float4* q4;      // Q images of the current parameters
  int n_strat, n_imp;
  const float *rays_o, *rays_d;
  const int* dl_list;
  const int* counters;
  const float *u_coarse, *u_fine;
  const float* t_uni;
  float* z;
};

struct SmemImp {
  float4 P[NP * 4];  // first-layer pre-activations of the sdf decoder (p_slot layout)
  ax_t ax_i[6][NP];
  float ax_f[6][NP];
  float W[QW_TOTAL];
  float one[NP], w[NP], z[NP];
};

// The SDF-only forward of the coarse samples runs on the Q images with the decoder tail in shared memory, so the
// kernel depends neither on the constant bank (no eslam_bind_decoders in the mapping loop, whose decoders change
// every iteration) nor on the 64-channel gather.
__global__ void __launch_bounds__(NP) k_importance(const __grid_constant__ ImportanceArgs a) {
  __shared__ SmemImp sm;
  __shared__ float s_zi[NP];  // resampled depths, [rl*n_imp + i]
  // A kernel launched behind this one with programmatic stream serialization (eslam_loss_backward_q_part, part 1: the
  // tiles that do not read what this kernel writes) may start as soon as every CTA here is resident or gone; without
  // such a dependent the instruction does nothing.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int R0 = a.counters[1];
  const int NS = a.n_strat;
  const int RPB = NP / NS;
  const int r00 = blockIdx.x * RPB;
  if (r00 >= R0) return;
  const int rays_here = min(RPB, R0 - r00);
  const int n_valid = rays_here * NS;
  const int q = threadIdx.x;
  const bool valid = q < n_valid;
  const int rl = valid ? q / NS : 0;
  const int k = q - rl * NS;
  const int r0 = r00 + rl;
  const int ray = a.dl_list[r0];
  load_tail_weights(sm.W, reinterpret_cast<const float*>(a.arena4) + a.fk.dec_off, q, NP);
  float o[3], d[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    o[c] = a.rays_o[ray * 3 + c];
    d[c] = a.rays_d[ray * 3 + c];
  }
  const float far = __fadd_rn(bbox_exit(a.fk, o, d), 0.01f);
  float zp = 0.f, pn[3] = {0.f, 0.f, 0.f};
  if (valid) {
    // z_uni = 0*(1-t) + far*t == far*t; perturbation inside the midpoints' intervals
    const float zc = __fmul_rn(far, a.t_uni[k]);
    const float lower = k > 0 ? __fmul_rn(0.5f, __fadd_rn(zc, __fmul_rn(far, a.t_uni[k - 1]))) : zc;
    const float upper = k < NS - 1 ? __fmul_rn(0.5f, __fadd_rn(__fmul_rn(far, a.t_uni[k + 1]), zc)) : zc;
    zp = __fadd_rn(lower, __fmul_rn(__fsub_rn(upper, lower), a.u_coarse[(long long)r0 * NS + k]));
#pragma unroll
    for (int c = 0; c < 3; ++c) pn[c] = normalize_axis(__fadd_rn(o[c], __fmul_rn(d[c], zp)), a.fk.lo[c], a.fk.hi[c]);
  }
  write_axis_setups<2>(a.fk, 0, pn, sm.ax_i, sm.ax_f, q);
  __syncthreads();
  gather_preact_tile<0>(a.fk, a.q4, sm.ax_i, sm.ax_f, sm.P, n_valid, q);
  decoder_weights_wait();
  __syncthreads();
  float h1[16], h2[16], os[3];
  mlp_tail_s(sm.W - DW_B1, sm.P, q, h1, h2, os);
  const float sdf = tanhf(os[0]);
  float u, e, alpha;
  sdf_to_alpha(sdf, sm.W[2 * QW_STRIDE], u, e, alpha);
  sm.one[q] = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
  sm.z[q] = zp;
  __syncthreads();
  float T = 1.0f;
  for (int j = 0; j < k; ++j) T *= sm.one[rl * NS + j];
  sm.w[q] = alpha * T;
  __syncthreads();
  // inverse cdf: pdf = w[1..NS-2] (un-normalised), bins = midpoints (NS-1 of them), cdf has NS-1 entries
  if (valid && k < a.n_imp) {
    const float* w = sm.w + rl * NS;
    const float* zz = sm.z + rl * NS;
    const float uu = a.u_fine[(long long)r0 * a.n_imp + k];
    const int ncdf = NS - 1;
    // searchsorted(cdf, u, right=True): number of cdf entries <= u
    float c = 0.f;
    int inds = 0;
    float c_below = 0.f, c_above = 0.f;
    {
      float run = 0.f;
      // entry 0 is 0, entry m is sum_{x<m} pdf[x] = sum w[1..m]
      int cnt = 0;
      for (int m = 0; m < ncdf; ++m) {
        if (m > 0) run = __fadd_rn(run, w[m]);
        if (run <= uu) ++cnt;
      }
      inds = cnt;
      const int below = max(inds - 1, 0), above = min(inds, ncdf - 1);
      run = 0.f;
      for (int m = 0; m < ncdf; ++m) {
        if (m > 0) run = __fadd_rn(run, w[m]);
        if (m == below) c_below = run;
        if (m == above) c_above = run;
      }
      const float b0 = __fmul_rn(0.5f, __fadd_rn(zz[below + 1], zz[below]));
      const float b1 = __fmul_rn(0.5f, __fadd_rn(zz[above + 1], zz[above]));
      float denom = __fsub_rn(c_above, c_below);
      if (denom < 1e-5f) denom = 1.0f;
      const float t = __fdiv_rn(__fsub_rn(uu, c_below), denom);
      c = __fadd_rn(b0, __fmul_rn(t, __fsub_rn(b1, b0)));
    }
    s_zi[rl * a.n_imp + k] = c;
  }
  __syncthreads();
  if (valid && k == 0) {
    // sort the n_imp resampled depths, merge with the (ascending) coarse ones
    float zi[16];
    const int NI = a.n_imp;
    for (int i = 0; i < NI; ++i) {
      float v = s_zi[rl * NI + i];
      int j = i;
      while (j > 0 && zi[j - 1] > v) {
        zi[j] = zi[j - 1];
        --j;
      }
      zi[j] = v;
    }
    const float* zz = sm.z + rl * NS;
    float* out = a.z + (long long)ray * (NS + NI);
    int ia = 0, ib = 0;
    for (int m = 0; m < NS + NI; ++m) {
      const bool take_a = (ib >= NI) || (ia < NS && zz[ia] <= zi[ib]);
      out[m] = take_a ? zz[ia++] : zi[ib++];
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// tracker outlier mask: lower median of |gt - rendered| by radix select  (Tracker.py:192-195)
// ---------------------------------------------------------------------------------------------------
struct TrackMaskArgs {
  const float *gt_depth, *depth;
  const unsigned char* band;
  int max_rays;
  int* counters;
  unsigned char* ray_mask;
  float* scratch;
};

__global__ void __launch_bounds__(1024) k_track_mask(const __grid_constant__ TrackMaskArgs a) {
  __shared__ unsigned hist[256];
  __shared__ unsigned s_prefix, s_rank;
  __shared__ int s_tot[4];
  const int R = min(a.counters[0], a.max_rays);
  const int t = threadIdx.x;
  unsigned* key = reinterpret_cast<unsigned*>(a.scratch);
  for (int r = t; r < R; r += blockDim.x) key[r] = __float_as_uint(fabsf(__fsub_rn(a.gt_depth[r], a.depth[r])));
  if (t == 0) {
    s_prefix = 0u;
    s_rank = R > 0 ? (unsigned)((R - 1) / 2) : 0u;
  }
  if (t < 4) s_tot[t] = 0;
  __syncthreads();
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    if (t < 256) hist[t] = 0u;
    __syncthreads();
    const unsigned prefix = s_prefix;
    const unsigned himask = pass == 0 ? 0u : (0xffffffffu << (shift + 8));
    for (int r = t; r < R; r += blockDim.x) {
      const unsigned kk = key[r];
      if ((kk & himask) == prefix) atomicAdd(&hist[(kk >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (t < 32) {  // warp 0: lane l owns bins 8l..8l+7; a warp scan finds the bin that holds rank s_rank
      unsigned h[8], sum = 0u;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        h[i] = hist[t * 8 + i];
        sum += h[i];
      }
      unsigned incl = sum;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const unsigned n = __shfl_up_sync(0xffffffffu, incl, o);
        if (t >= o) incl += n;
      }
      const unsigned excl = incl - sum, rank = s_rank;
      const bool mine = rank >= excl && rank < incl;  // exactly one lane unless R == 0
      const unsigned who = __ballot_sync(0xffffffffu, mine);
      if (who == 0u) {
        if (t == 31) {
          s_prefix = prefix | (255u << shift);
          s_rank = 0u;
        }
      } else if (mine) {
        unsigned cum = excl;
        int b = 0;
#pragma unroll
        for (; b < 8; ++b) {
          if (cum + h[b] > rank) break;
          cum += h[b];
        }
        if (b > 7) b = 7;
        s_prefix = prefix | ((unsigned)(t * 8 + b) << shift);
        s_rank = rank - cum;
      }
    }
    __syncthreads();
  }
  const float median = __uint_as_float(s_prefix);
  const float thr = __fmul_rn(10.0f, median);
  int c0 = 0, c1 = 0, c2 = 0, c3 = 0;
  for (int r = t; r < R; r += blockDim.x) {
    const bool m = __uint_as_float(key[r]) < thr;
    a.ray_mask[r] = m ? 1 : 0;
    if (m) {
      c0 += 1;
      c1 += a.band[r * 4 + 0];
      c2 += a.band[r * 4 + 1];
      c3 += a.band[r * 4 + 2];
    }
  }
  c0 = (int)warp_sum((float)c0);
  c1 = (int)warp_sum((float)c1);
  c2 = (int)warp_sum((float)c2);
  c3 = (int)warp_sum((float)c3);
  if ((t & 31) == 0) {
    atomicAdd(&s_tot[0], c0);
    atomicAdd(&s_tot[1], c1);
    atomicAdd(&s_tot[2], c2);
    atomicAdd(&s_tot[3], c3);
  }
  __syncthreads();
  if (t < 4) a.counters[2 + t] = s_tot[t];
  if (t == 0) a.scratch[a.max_rays] = median;
}

}  // namespace eslam
