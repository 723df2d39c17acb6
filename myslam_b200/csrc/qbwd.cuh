// Fused loss + backward of both loops ON THE Q IMAGES: the first decoder layer lives in the
// 16-channel images Q = W1_slice . plane, so a sample gathers 64 instead of 128 bytes per corner, the 64 -> 16 layer
// is gone from forward, backward and the weight-gradient contraction, and the plane gradients leave the kernel as
// 16-channel reductions into gradient images (the optimiser tail, qplane.cuh, turns them into d loss / d plane and dW1).
// Reference semantics: src/networks/decoders.py:64-146, src/utils/Renderer.py:136-153, src/Tracker.py:114-148,192-208,
// src/Mapper.py:110-144,337-349.
//
// One CTA = NP points (whole rays), 2 * NP threads: threads [0, NP) own the sdf decoder of point tid, threads [NP, 2 NP)
// the rgb decoder of point tid - NP (render.cuh's split).  Differences from render_bwd_body beyond the Q form:
//   * the gather also forms J = d pre-activation / d normalised coordinate (3 x 16 per point and decoder: the corner
//     differences it has in registers anyway), parked in a row tile; the coordinate gradient is then J^T g in point
//     layout and the backward never fetches a corner twice (the parameter form re-gathers all 48 lines per sample)
//   * one staging round for the weight gradients: the row tile is overwritten in place by (ga2 | h1 | h2 | gout), the
//     pre-activation tile by ga1, one barrier, then dW2 / dW3 / db1 / db2 / db3 on the tensor cores
//   * everything a ray needs from global memory late in the kernel (gt depth / colour, outlier mask, loss normalisers,
//     source pixel of the pose gradient) is fetched at the top, so no L2 round trip is exposed behind a barrier
#pragma once
#include "render.cuh"

namespace eslam {

constexpr int RS = 52;  // floats per point of the row tile: J (3 x 16) or (ga2 | h1 | h2 | gout)

template <bool ROWS>
struct SmemBwdQ {
  float4 P[2][NP * 4];                 // pre-activations, later their gradient (p_slot layout)
  float R[2][ROWS ? NP * RS : 4];      // row tile per decoder
  ax_t ax_i[12][NP];
  float ax_f[12][NP];
  float W[QW_TOTAL];
  float one[NP], w[NP], z[NP], c[3][NP], gww[NP];
  float gp[2][3][NP];  // d loss / d normalised coordinate, per decoder half
  float rayv[4][16];   // per ray: rendered depth, r, g, b
  float rayg[4][16];   // per ray: upstream g_depth, g_rgb
  float rayd[16];      // per ray gt depth
  int raym[16];        // per ray loss-mask flag
  double rayc[3][16];  // per ray gt colour
  int norm[8];         // loss normalisers (counters layout)
  float raydir[2][16]; // per ray camera-frame direction (x, y) of its pixel; z is -1
  int rayframe[16];    // per ray frame of the window
  float red[NT_BWD / 32];
  double redd[(NT_BWD / 32) * 5];
};

// Sum over the 6 planes of decoder FIELD of the bilinear fetch from its Q images (this lane's 4 pre-activations) and,
// with WJ, its derivative with respect to the three normalised coordinates (grid_sample's grid gradient: corner
// differences times (size-1)/2, zero where the coordinate is clipped, GridSampler.h).
template <int FIELD, bool WJ>
__device__ __forceinline__ void gather_preact_j(const FieldK& fk, const float4* __restrict__ q4, const ax_t (*ax_i)[NP],
                                                const float (*ax_f)[NP], int qq, int sub, float4& pre,
                                                float4 (&J)[3]) {
  float4 v[6][4];
  float fu[6], fv[6], mu[6], mv[6];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const int t = s * 3 + p;
      const int au = FIELD * 6 + s * 3 + pair_u(p), av = FIELD * 6 + s * 3 + pair_v(p);
      const PlaneK& pl = fk.pl[FIELD * 6 + t];
      const int u0 = ax_i[au][qq], v0 = ax_i[av][qq];
      fu[t] = ax_f[au][qq];
      fv[t] = ax_f[av][qq];
      if (WJ) {
        mu[t] = axis_grad_mult(u0, fu[t], pl.W);
        mv[t] = axis_grad_mult(v0, fv[t], pl.H);
      }
      const int base = (pl.off4 >> 1) + (v0 * pl.W + u0) * 4 + sub;
      const int dx = (u0 + 1 < pl.W) ? 4 : 0, dy = (v0 + 1 < pl.H) ? pl.W * 4 : 0;
      v[t][0] = ldg4(q4 + base);
      v[t][1] = ldg4(q4 + base + dx);
      v[t][2] = ldg4(q4 + base + dy);
      v[t][3] = ldg4(q4 + base + dy + dx);
    }
  }
  if (WJ) J[0] = J[1] = J[2] = f4_zero();
  float4 acc[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    float4 sum = f4_zero();
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const int t = s * 3 + p;
      const float w00 = (1.f - fu[t]) * (1.f - fv[t]), w01 = fu[t] * (1.f - fv[t]), w10 = (1.f - fu[t]) * fv[t],
                  w11 = fu[t] * fv[t];
      float4 tap = f4_mul(w00, v[t][0]);
      tap = f4_fma(w01, v[t][1], tap);
      tap = f4_fma(w10, v[t][2], tap);
      tap = f4_fma(w11, v[t][3], tap);
      sum = (p == 0) ? tap : f4_add(sum, tap);  // (xy + xz) + yz, decoders.py:82
      if (WJ) {
        const float4 a = f4_sub(v[t][1], v[t][0]), b = f4_sub(v[t][3], v[t][2]);  // d/du on the lower / upper row
        const float4 c = f4_sub(v[t][2], v[t][0]), d = f4_sub(v[t][3], v[t][1]);  // d/dv on the left / right column
        const float4 du = f4_fma(fv[t], f4_sub(b, a), a);
        const float4 dv = f4_fma(fu[t], f4_sub(d, c), c);
        J[pair_u(p)] = f4_fma(mu[t], du, J[pair_u(p)]);
        J[pair_v(p)] = f4_fma(mv[t], dv, J[pair_v(p)]);
      }
    }
    acc[s] = sum;
  }
  pre = f4_add(acc[0], acc[1]);  // coarse + fine
}

// pre-activation tile and (WJ) the J rows of decoder FIELD for the NP slots of one half of the CTA
template <int FIELD, bool WJ>
__device__ __forceinline__ void gather_preact_rows(const FieldK& fk, const float4* __restrict__ q4,
                                                   const ax_t (*ax_i)[NP], const float (*ax_f)[NP], float4* P, float* R,
                                                   int n_valid, int tid) {
  const int warp = tid >> 5, lane = tid & 31, grp = lane >> 2, sub = lane & 3;
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    const int qq = warp * 32 + it * 8 + grp;
    float4 p = f4_zero(), J[3];
    J[0] = J[1] = J[2] = f4_zero();
    if (qq < n_valid) gather_preact_j<FIELD, WJ>(fk, q4, ax_i, ax_f, qq, sub, p, J);
    P[p_slot(qq, sub)] = p;
    if (WJ) {
      float4* row = reinterpret_cast<float4*>(R + qq * RS);
#pragma unroll
      for (int a = 0; a < 3; ++a) row[a * 4 + sub] = J[a];
    }
  }
}

// operand sources of the weight-gradient mma fragments (see render.cuh FromBuf / FromTile)
struct FromRow {  // row tile [NP][RS]
  const float* base;
  int o[4];
  __device__ __forceinline__ FromRow(const float* buf, int c, int t) : base(buf) {
    const int r[4] = {2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9};
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = r[i] * RS + c;
  }
  __device__ __forceinline__ float at(int q0, int i) const { return base[q0 * RS + o[i]]; }
};
struct FromP {  // pre-activation tile, p_slot layout (q0 is a multiple of 16, so the swizzle term depends on r only)
  const float* base;
  int o[4];
  __device__ __forceinline__ FromP(const float4* P, int c, int t) : base(reinterpret_cast<const float*>(P)) {
    const int r[4] = {2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9};
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = (r[i] * 4 + ((c >> 2) ^ ((r[i] >> 1) & 3))) * 4 + (c & 3);
  }
  __device__ __forceinline__ float at(int q0, int i) const { return base[q0 * 16 + o[i]]; }
};

// Weight gradients of one decoder's layers 2 and 3 and of all three biases (dW1 is formed per texel by the optimiser
// tail).  Every owner thread parks (ga2 | h1 | h2 | gout) in ITS OWN row of the row tile (whose J it has just consumed)
// and ga1 in its row of the pre-activation tile; after one barrier over the half: warps 0,1 dW2 (half of the points
// each), warp 2 dW3, warp 3 the biases.  Contains a barrier over the half: call from uniform control flow.
__device__ __forceinline__ void weight_grads_q(float* R, float4* P, float* gdec, int half, int q, const float (&h1)[16],
                                               const float (&h2)[16], const float (&ga1)[16], const float (&ga2)[16],
                                               const float (&gout)[3], bool with_wgrads) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wl = warp & 3;
  const int nout = half ? 3 : 1;
#pragma unroll
  for (int c = 0; c < 4; ++c) P[p_slot(q, c)] = make_float4(ga1[c * 4], ga1[c * 4 + 1], ga1[c * 4 + 2], ga1[c * 4 + 3]);
  if (with_wgrads) {
    float4* row = reinterpret_cast<float4*>(R + q * RS);
#pragma unroll
    for (int v = 0; v < 4; ++v) {
      row[v] = make_float4(ga2[v * 4], ga2[v * 4 + 1], ga2[v * 4 + 2], ga2[v * 4 + 3]);
      row[4 + v] = make_float4(h1[v * 4], h1[v * 4 + 1], h1[v * 4 + 2], h1[v * 4 + 3]);
      row[8 + v] = make_float4(h2[v * 4], h2[v * 4 + 1], h2[v * 4 + 2], h2[v * 4 + 3]);
    }
    row[12] = make_float4(gout[0], gout[1], gout[2], 0.f);
  }
  half_sync(half);
  if (!with_wgrads) return;
  float* gB1 = gdec + (half ? C_B1 : S_B1);
  float* gW2 = gdec + (half ? C_W2 : S_W2);
  float* gB2 = gdec + (half ? C_B2 : S_B2);
  float* gW3 = gdec + (half ? C_W3 : S_W3);
  float* gB3 = gdec + (half ? C_B3 : S_B3);
  const int g = lane >> 2, t = lane & 3;
  const FromRow g2_lo(R, g, t), g2_hi(R, 8 + g, t);  // ga2 channels g, g + 8
  const FromRow g3_lo(R, 48 + (g & 3), t);           // gout channel g (< 4; lanes with g >= 4 contribute zeros)
  if (wl < 2) {
    float acc[2][4];
    const FromRow b[2] = {FromRow(R, 16 + g, t), FromRow(R, 24 + g, t)};
    wgrad_tiles<16, 2>(g2_lo, g2_hi, b, wl * (NP / 2), (wl + 1) * (NP / 2), lane, acc);
    wgrad_store(gW2, 16, 0, 16, lane, acc[0]);
    wgrad_store(gW2, 16, 8, 16, lane, acc[1]);
  } else if (wl == 2) {
    float acc[2][4];
    const FromRow b[2] = {FromRow(R, 32 + g, t), FromRow(R, 40 + g, t)};
    wgrad_tiles<4, 2>(g3_lo, g3_lo, b, 0, NP, lane, acc);
    wgrad_store(gW3, 16, 0, nout, lane, acc[0]);
    wgrad_store(gW3, 16, 8, nout, lane, acc[1]);
  } else {
    float acc[4];
    const FromP g1_lo(P, g, t), g1_hi(P, 8 + g, t);
    wgrad_bias<16>(g1_lo, g1_hi, lane, acc);
    wgrad_store_bias(gB1, 16, lane, acc);
    wgrad_bias<16>(g2_lo, g2_hi, lane, acc);
    wgrad_store_bias(gB2, 16, lane, acc);
    wgrad_bias<4>(g3_lo, g3_lo, lane, acc);
    wgrad_store_bias(gB3, nout, lane, acc);
  }
}

// GF: gradients for the planes (gradient images a.gq4) and the decoders (a.grad_arena's decoder block, except dW1).
// GR: gradients for the poses.  CACHED (tracker, !GF): sdf, rgb and the ReLU masks of every sample come from
// k_render_fwd_q on the same rays, so neither the gather nor the forward MLPs run; the coordinate gradients then
// fetch the Q corners once (coord_grads_q).
template <bool GF, bool GR, bool CACHED>
__device__ __forceinline__ void render_bwd_q_body(const BwdArgs& a, const int tile) {
  static_assert(!(GF && CACHED), "the cached form has no activations for weight gradients");
  constexpr bool ROWS = !CACHED;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemBwdQ<ROWS>& sm = *reinterpret_cast<SmemBwdQ<ROWS>*>(smem_raw);
  const int R = a.counters ? min(a.counters[0], a.n_rays) : a.n_rays;
  const int S = a.S;
  const int RPB = min(NP / S, 16);
  const int ray0 = tile * RPB;
  if (ray0 >= R) return;
  const int rays_here = min(RPB, R - ray0);
  if (GF && a.part == 1) {  // CTA-uniform: a tile holding a depth-less ray belongs to the other launch of the pair
    bool depthless = false;
    for (int r = 0; r < rays_here; ++r) depthless = depthless || !(a.gt_depth[ray0 + r] > 0.f);
    if (depthless) return;
  }
  const int n_valid = rays_here * S;
  const int tid = threadIdx.x;
  const int half = tid >> 7;  // 0: sdf decoder, 1: rgb decoder (warp-uniform)
  const int q = tid & (NP - 1);
  const int warp = tid >> 5, lane = tid & 31;
  const bool valid = q < n_valid;
  const int rl = valid ? q / S : 0;
  const int k = q - rl * S;
  const int ray = ray0 + rl;

  PHASE_INIT();
  // ---- P0: points, normalised coordinates, axis set-ups of this half's two resolution groups; early fetches
  float pn[3] = {0.f, 0.f, 0.f}, zk = 0.f;
  if (valid) {
    zk = a.z[(long long)ray * S + k];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float p = __fadd_rn(a.rays_o[ray * 3 + c], __fmul_rn(a.rays_d[ray * 3 + c], zk));
      pn[c] = normalize_axis(p, a.fk.lo[c], a.fk.hi[c]);
    }
  }
  load_tail_weights(sm.W, reinterpret_cast<const float*>(a.arena4) + a.fk.dec_off, tid, NT_BWD);
  // per-ray inputs of the loss (owner: sample 0 of the ray, sdf half) and the normalisers travel global -> shared
  // asynchronously in the weights' copy group; the outlier mask (a byte) goes through a register of the warp that
  // will composite the ray (warp w: rays w, w + 8)
  const bool ray_owner = half == 0 && valid && k == 0;
  int pre_m[2] = {1, 1};
  if (ray_owner) {
    cp_async_small<4>(&sm.rayd[rl], a.gt_depth + ray);
#pragma unroll
    for (int c = 0; c < 3; ++c) cp_async_small<8>(&sm.rayc[c][rl], a.gt_color + (long long)ray * 3 + c);
  }
  if (a.ray_mask) {
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (warp + 8 * i < rays_here) pre_m[i] = (int)a.ray_mask[ray0 + warp + 8 * i];
  }
  if (tid < 2) cp_async16(&sm.norm[4 * tid], a.norm + 4 * tid);
  asm volatile("cp.async.commit_group;" ::: "memory");
  // source pixel of the ray (pose gradient): two dependent loads, parked in shared memory once the gather is out
  int pose_frame = 0;
  float pose_dx = 0.f, pose_dy = 0.f;
  if (GR && a.pose_grad && tid < rays_here) {
    const int slot = a.src[ray0 + tid];
    pose_frame = slot / a.n_per_img;
    const long long pix = a.pix_idx[slot];
    const float pi = (float)(a.W0 + (int)(pix % a.Wc)), pj = (float)(a.H0 + (int)(pix / a.Wc));
    pose_dx = __fdiv_rn(__fsub_rn(pi, a.cx), a.fx);
    pose_dy = -__fdiv_rn(__fsub_rn(pj, a.cy), a.fy);
  }
  if (!CACHED || GR) write_axis_setups<2>(a.fk, 2 * half, pn, sm.ax_i + 6 * half, sm.ax_f + 6 * half, q);
  __syncthreads();
  PHASE_MARK(0);
  float h1[16], h2[16], out[3] = {0.f, 0.f, 0.f};
  float sdf = 0.f, u = 0.f, e = 0.f, alpha = 0.f, one = 1.f, rgb[3] = {0.f, 0.f, 0.f};
  const float* Wh = sm.W + half * QW_STRIDE - DW_B1;  // field.cuh's DW_* offsets minus the absent W1
  float4* Ph = sm.P[half];
  float* Rh = sm.R[half];
  float beta;
  if (CACHED) {
    PHASE_MARK(1);
    decoder_weights_wait();
    __syncthreads();
    PHASE_MARK(2);
    beta = sm.W[2 * QW_STRIDE];
    unsigned m = 0u;
    if (valid) {
      const float4 c4 = a.act4[(long long)ray * S + k];
      if (half == 0) {
        m = __float_as_uint(c4.w);
        sdf = a.sdf_in[(long long)ray * S + k];
      } else {
        m = a.actm[(long long)ray * S + k];
        rgb[0] = c4.x;
        rgb[1] = c4.y;
        rgb[2] = c4.z;
      }
    }
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      h1[j] = (m >> j) & 1u ? 1.f : 0.f;
      h2[j] = (m >> (16 + j)) & 1u ? 1.f : 0.f;
    }
  } else {
    // ---- P2: first-layer pre-activations (without bias) from the Q images, and J
    if (half == 0)
      gather_preact_rows<0, GR>(a.fk, a.q4, sm.ax_i, sm.ax_f, Ph, Rh, n_valid, q);
    else
      gather_preact_rows<1, GR>(a.fk, a.q4, sm.ax_i, sm.ax_f, Ph, Rh, n_valid, q);
    PHASE_MARK(1);
    decoder_weights_wait();
    __syncthreads();
    PHASE_MARK(2);
    // ---- P3: layers 2 and 3 of this half's decoder
    beta = sm.W[2 * QW_STRIDE];
    mlp_tail_s(Wh, Ph, q, h1, h2, out);
    if (half == 0)
      sdf = tanhf(out[0]);
    else {
#pragma unroll
      for (int c = 0; c < 3; ++c) rgb[c] = sigmoidf_(out[c]);
    }
  }
  if (half == 0) {
    sdf_to_alpha(sdf, beta, u, e, alpha);
    one = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
    sm.one[q] = one;
    sm.w[q] = valid ? alpha : 0.f;  // replaced by the compositing weight below
    sm.z[q] = zk;
  } else {
#pragma unroll
    for (int c = 0; c < 3; ++c) sm.c[c][q] = rgb[c];
  }
  if (GR && a.pose_grad && tid < rays_here) {
    sm.rayframe[tid] = pose_frame;
    sm.raydir[0][tid] = pose_dx;
    sm.raydir[1][tid] = pose_dy;
  }
  PHASE_MARK(3);
  __syncthreads();
  PHASE_MARK(4);
  // ---- P4/P5: one WARP per ray (lane l: samples l and l + 32): transmittance as a prefix product, the ray's depth and
  //      colour, the loss gradients at them, and the backward of the compositing as a suffix sum -- shuffles only
  //      (Renderer.py:140-147; Tracker.py:192-204 / Mapper.py:337-346).  Leaves w and d loss / d alpha per sample.
  double ls[5] = {0.0, 0.0, 0.0, 0.0, 0.0};  // fs, center, tail, depth, colour
  const int n_all = sm.norm[0], n_mask = sm.norm[2];
  const float inv_f = a.w_fs / (float)sm.norm[3], inv_c = a.w_center / (float)sm.norm[4],
              inv_t = a.w_tail / (float)sm.norm[5];
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r2 = warp + 8 * i;
    if (r2 >= rays_here) break;
    const int j0 = r2 * S + lane, j1 = j0 + 32;
    const bool v0 = lane < S, v1 = lane + 32 < S;
    const float o0 = v0 ? sm.one[j0] : 1.f, o1 = v1 ? sm.one[j1] : 1.f;
    float inc0 = o0, inc1 = o1;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float t0 = __shfl_up_sync(0xffffffffu, inc0, d), t1 = __shfl_up_sync(0xffffffffu, inc1, d);
      if (lane >= d) {
        inc0 *= t0;
        inc1 *= t1;
      }
    }
    const float tot0 = __shfl_sync(0xffffffffu, inc0, 31);
    float T0 = __shfl_up_sync(0xffffffffu, inc0, 1), T1 = __shfl_up_sync(0xffffffffu, inc1, 1);
    if (lane == 0) T0 = T1 = 1.f;
    T1 *= tot0;
    const float w0 = v0 ? sm.w[j0] * T0 : 0.f, w1 = v1 ? sm.w[j1] * T1 : 0.f;
    const float z0 = v0 ? sm.z[j0] : 0.f, z1 = v1 ? sm.z[j1] : 0.f;
    float c0[3], c1[3], rv[4];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      c0[c] = v0 ? sm.c[c][j0] : 0.f;
      c1[c] = v1 ? sm.c[c][j1] : 0.f;
    }
    rv[0] = warp_sum(fmaf(w0, z0, w1 * z1));
#pragma unroll
    for (int c = 0; c < 3; ++c) rv[1 + c] = warp_sum(fmaf(w0, c0[c], w1 * c1[c]));
    // loss gradients at the rendered depth / colour (lane 0 does the float64 colour term and owns the loss sums)
    const float d = sm.rayd[r2];
    const int m = a.ray_mask ? pre_m[i] : (d > 0.f ? 1 : 0);
    float rg[4] = {0.f, 0.f, 0.f, 0.f};
    if (lane == 0) {
      if (m) {
        const float diff = d - rv[0];
        rg[0] = -2.0f * diff * (a.w_depth / (float)n_mask);
        ls[3] += (double)(diff * diff);
      }
      const bool col = a.ray_mask ? (m != 0) : true;
      const double ncol = 3.0 * (double)(a.ray_mask ? n_mask : n_all);
      if (col) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          const double diff = sm.rayc[c][r2] - (double)rv[1 + c];
          rg[1 + c] = (float)(-2.0 * diff * (a.w_color / ncol));
          ls[4] += diff * diff;
        }
      }
      sm.raym[r2] = m;
#pragma unroll
      for (int c = 0; c < 4; ++c) sm.rayg[c][r2] = rg[c];
    }
#pragma unroll
    for (int c = 0; c < 4; ++c) rg[c] = __shfl_sync(0xffffffffu, rg[c], 0);
    // backward of the compositing: gw = d loss / d w, B = sum over later samples of gw * w
    const float gw0 = rg[0] * z0 + rg[1] * c0[0] + rg[2] * c0[1] + rg[3] * c0[2];
    const float gw1 = rg[0] * z1 + rg[1] * c1[0] + rg[2] * c1[1] + rg[3] * c1[2];
    float s0 = gw0 * w0, s1 = gw1 * w1;
#pragma unroll
    for (int dd = 1; dd < 32; dd <<= 1) {
      const float t0 = __shfl_down_sync(0xffffffffu, s0, dd), t1 = __shfl_down_sync(0xffffffffu, s1, dd);
      if (lane + dd < 32) {
        s0 += t0;
        s1 += t1;
      }
    }
    const float tot1 = __shfl_sync(0xffffffffu, s1, 0);
    float B0 = __shfl_down_sync(0xffffffffu, s0, 1), B1 = __shfl_down_sync(0xffffffffu, s1, 1);
    if (lane == 31) B0 = B1 = 0.f;
    B0 += tot1;
    if (v0) {
      sm.w[j0] = w0;
      sm.gww[j0] = gw0 * T0 - B0 / o0;
    }
    if (v1) {
      sm.w[j1] = w1;
      sm.gww[j1] = gw1 * T1 - B1 / o1;
    }
  }
  __syncthreads();
  // ---- gradient at this half's decoder outputs
  float g_beta = 0.f, gout[3] = {0.f, 0.f, 0.f};
  if (half == 1) {
    if (valid) {
      const float ww = sm.w[q];
#pragma unroll
      for (int c = 0; c < 3; ++c) gout[c] = sm.rayg[1 + c][rl] * ww * rgb[c] * (1.0f - rgb[c]);
    }
  } else if (valid) {
    float gdir = 0.f;
    if (sm.raym[rl]) {
      const float d = sm.rayd[rl];
      const int band = sdf_band(zk, d, a.tr, a.tr04);
      if (band == 0) {
        const float r = sdf - 1.0f;
        gdir = 2.0f * r * inv_f;
        ls[0] = (double)(r * r);
      } else if (band < 3) {
        const float r = __fadd_rn(zk, __fmul_rn(sdf, a.tr)) - d;
        gdir = 2.0f * r * a.tr * (band == 1 ? inv_c : inv_t);
        if (band == 1)
          ls[1] = (double)(r * r);
        else
          ls[2] = (double)(r * r);
      }
    }
    const float g_alpha = sm.gww[q];
    const float du = u * (1.0f - u);
    const float g_sdf = gdir + g_alpha * (-beta * beta * e * du);
    g_beta = g_alpha * e * (u - beta * sdf * du);
    gout[0] = g_sdf * (1.0f - sdf * sdf);
  }
  PHASE_MARK(5);
  // ---- P6: backward through layers 3 and 2 -> gradient at the pre-activations
  float ga1[16], ga2[16];
  mlp_backward_hidden_s(Wh, gout, h1, h2, ga1, ga2);  // sdf: gout[1] = gout[2] = 0
  PHASE_MARK(6);
  if (GR && !CACHED) {  // coordinate gradient: J^T g (the thread's own row, written by the gather)
    const float* row = Rh + q * RS;
#pragma unroll
    for (int ax = 0; ax < 3; ++ax) {
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 j4 = lds4(row + ax * 16 + c * 4);
        acc = fmaf(j4.x, ga1[c * 4], acc);
        acc = fmaf(j4.y, ga1[c * 4 + 1], acc);
        acc = fmaf(j4.z, ga1[c * 4 + 2], acc);
        acc = fmaf(j4.w, ga1[c * 4 + 3], acc);
      }
      sm.gp[half][ax][q] = valid ? acc : 0.f;
    }
  }
  const bool wg = GF && !(a.dbg & 2);
  if (GF) {
    const float gb = warp_sum(g_beta);
    if (lane == 0) sm.red[warp] = gb;
  }
  weight_grads_q(Rh, Ph, GF ? a.grad_arena + a.fk.dec_off : nullptr, half, q, h1, h2, ga1, ga2, gout, wg);
  PHASE_MARK(7);
  if (GF && tid == 0) {
    float gb = 0.f;
    for (int i = 0; i < NP / 32; ++i) gb += sm.red[i];  // only the sdf half carries beta gradients
    atomicAdd(a.grad_arena + a.fk.dec_off + P_BETA, gb);
  }
  PHASE_MARK(8);
  // ---- P7: reductions into the gradient images (gather layout, each half its own decoder) / cached coordinate gradients
  {
    const int wl = (tid & (NP - 1)) >> 5;
    if (GF) {
      const int qb = wl * 32 + (lane >> 3) * 8;  // 8 consecutive points per 8-lane group
      if (a.dbg & 1) {
      } else if (half == 0)
        scatter_q<0>(a.fk, a.gq4, sm.ax_i, sm.ax_f, Ph, qb, n_valid, lane & 7);
      else
        scatter_q<1>(a.fk, a.gq4, sm.ax_i, sm.ax_f, Ph, qb, n_valid, lane & 7);
    } else if (CACHED) {
      if (half == 0)
        coord_grads_q<0>(a.fk, a.q4, sm.ax_i, sm.ax_f, Ph, wl, lane >> 2, lane & 3, n_valid, sm.gp[0]);
      else
        coord_grads_q<1>(a.fk, a.q4, sm.ax_i, sm.ax_f, Ph, wl, lane >> 2, lane & 3, n_valid, sm.gp[1]);
    }
  }
  PHASE_MARK(9);
  // ---- P8: ray / pose gradients: one warp per (ray, component): d loss / d o = sum_k g_k, d loss / d d = sum_k z_k g_k
  //      (g = coordinate gradient back in world units), then d loss / d t and d loss / d R = (d loss / d d) (x) dir_cam
  if (GR && a.pose_grad) {
    __syncthreads();
    for (int p = warp; p < rays_here * 6; p += NT_BWD / 32) {
      const int r2 = p / 6, comp = p - r2 * 6, ax = comp % 3;
      const bool is_d = comp >= 3;
      float acc = 0.f;
      for (int j = lane; j < S; j += 32) {
        const float g = sm.gp[0][ax][r2 * S + j] + sm.gp[1][ax][r2 * S + j];
        acc += is_d ? g * sm.z[r2 * S + j] : g;
      }
      acc = warp_sum(acc) * (2.0f / (a.fk.hi[ax] - a.fk.lo[ax]));
      float* dst = a.pose_grad + sm.rayframe[r2] * 12 + ax * 4;
      if (!is_d) {
        if (lane == 0) atomicAdd(dst + 3, acc);  // d loss / d t[ax]
      } else if (lane < 3) {
        const float dir = lane == 2 ? -1.0f : sm.raydir[lane][r2];
        atomicAdd(dst + lane, acc * dir);  // d loss / d R[ax][lane]
      }
    }
  }
  PHASE_MARK(10);
  // ---- loss sums
  if (a.loss_acc) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const double v = warp_sum_d(ls[i]);
      if (lane == 0) sm.redd[warp * 5 + i] = v;
    }
    __syncthreads();
    if (tid < 5) {
      double v = 0.0;
      for (int i = 0; i < NT_BWD / 32; ++i) v += sm.redd[i * 5 + tid];
      atomicAdd(a.loss_acc + tid, v);
    }
  }
}

// tracker: pose gradient on the activations k_render_fwd_q kept (eslam_pose_backward_q)
__global__ void __launch_bounds__(NT_BWD, 2) k_pose_bwd_q(const __grid_constant__ BwdArgs a) {
  render_bwd_q_body<false, true, true>(a, blockIdx.x);
}

// mapper: gradient images + decoder gradients (+ poses) (eslam_loss_backward_q)
template <bool GR>
__global__ void __launch_bounds__(NT_BWD, 2) k_map_bwd_q(const __grid_constant__ BwdArgs a) {
  render_bwd_q_body<true, GR, false>(a, blockIdx.x);
}

// mapper, part 2 of the split launch: only the tiles that hold a depth-less ray, enumerated from the (ascending) list
// of depth-less rays the sampling kernel left; a tile belongs to its first listed ray.  A small persistent grid: the
// list is a few percent of the batch, and a grid over all tiles would be CTAs that start only to leave.
template <bool GR>
__global__ void __launch_bounds__(NT_BWD, 2) k_map_bwd_q_dl(const __grid_constant__ BwdArgs a) {
  const int R0 = a.counters[1];
  const int RPB = min(NP / a.S, 16);
  for (int i = blockIdx.x; i < R0; i += gridDim.x) {
    const int tile = a.dl_list[i] / RPB;
    if (i > 0 && a.dl_list[i - 1] / RPB == tile) continue;
    __syncthreads();  // the previous tile's shared memory is done with
    render_bwd_q_body<true, GR, false>(a, tile);
  }
}

}  // namespace eslam
