// EXPERIMENTAL -- pre-activated planes ("apply the first decoder layer on the planes").
// Not wired into the product path yet: the Python mirror never calls these entry points; tests/test_gpu_experimental.py
// holds them to 1e-5 of the product's render forward.
//
// Bilinear interpolation is linear, so the first decoder layer of decoders.py:87-125 commutes with the plane fetch
// of decoders.py:64-85:
//     W1 . (sum over planes of bilinear(plane)) + b1  =  sum over planes of bilinear(W1_half . plane) + b1
// k_q_build keeps Q = W1_half . plane for each of the 12 planes as a 16-channel channels-last image (64 B per texel,
// plane i at HALF the float offset of plane i in the parameter arena); k_render_fwd_q is k_render_fwd on that image:
// 4 lanes per point instead of 8, 64 instead of 128 bytes per corner, both decoders gathered in one phase, and the
// 64 -> 16 layer (1 024 of the 1 300 FMA per sample and decoder) gone.  Values differ from k_render_fwd by the
// re-association of that one sum.
#pragma once
#include "optim.cuh"
#include "render.cuh"

namespace eslam {

// The three planes of a (decoder, scale) group are contiguous in the arena and share one 16 x 32 slice of W1, so the
// dense kernels below work on four flat texel ranges instead of twelve planes.
struct QGroups {
  int t0[4];     // first texel (arena float offset / 32) of group g = decoder * 2 + scale
  int n[4];      // texels of the group
  int unit0[5];  // first work unit (CTA or warp chunk) of the group; unit0[4] = total
};

inline int make_q_groups(const eslam_field_t* f, int texels_per_unit, QGroups* out) {
  int units = 0;
  for (int g = 0; g < 4; ++g) {
    long long n = 0;
    for (int p = 0; p < 3; ++p) {
      const eslam_plane_t& pl = f->plane[g * 3 + p];
      if (pl.offset % 32 != 0) return ESLAM_EINVAL;
      if (p > 0 && pl.offset != f->plane[g * 3 + p - 1].offset +
                                    (long long)f->plane[g * 3 + p - 1].H * f->plane[g * 3 + p - 1].W * 32)
        return ESLAM_EUNSUPPORTED;  // the planes of a group must be contiguous
      n += (long long)pl.H * pl.W;
    }
    if (n > 0x3fffffff) return ESLAM_EUNSUPPORTED;
    out->t0[g] = (int)(f->plane[g * 3].offset / 32);
    out->n[g] = (int)n;
    out->unit0[g] = units;
    units += (int)((n + texels_per_unit - 1) / texels_per_unit);
  }
  out->unit0[4] = units;
  return 0;
}

__device__ __forceinline__ int q_group_of(const QGroups& qg, int unit) {
  return (unit >= qg.unit0[2]) ? (unit >= qg.unit0[3] ? 3 : 2) : (unit >= qg.unit0[1] ? 1 : 0);
}

struct QBuildArgs {
  QGroups qg;
  const float4* arena4;
  const float* dec;  // packed decoder block (include/eslam_b200.h)
  float2* q2;
};

// Q = W1_slice . texel for every texel of the map.  W1 changes with every optimiser step, so this runs once per
// mapping iteration over all 212 k texels (room0): 27 MB read, 13.5 MB written, both L2 resident.  One CTA = QB_TPC
// texels of one group: the tile is staged into shared memory with fully coalesced 16-byte loads (every byte in
// flight is a distinct byte: 16 KB per CTA), then thread (texel, half) computes 8 of the texel's 16 outputs.
constexpr int QB_TPC = 128;

// swizzled texel tile: row t holds 8 float4; physical slot = c4 ^ (t & 7) (conflict-free for thread-per-row readers)
__device__ __forceinline__ int t_slot(int t, int c4) { return t * 8 + (c4 ^ (t & 7)); }

__global__ void __launch_bounds__(256) k_q_build(const __grid_constant__ QBuildArgs a) {
  __shared__ __align__(16) float sW[16 * 32];
  __shared__ float4 sT[QB_TPC * 8];
  const int g = q_group_of(a.qg, blockIdx.x);
  const int tl0 = (blockIdx.x - a.qg.unit0[g]) * QB_TPC;
  const int cnt = min(QB_TPC, a.qg.n[g] - tl0);
  const long long tbase = (long long)a.qg.t0[g] + tl0;
  const float4* src = a.arena4 + tbase * 8;
#pragma unroll
  for (int u = 0; u < QB_TPC * 8 / 256; ++u) {
    const int i = u * 256 + threadIdx.x, t = i >> 3;
    if (t < cnt) sT[t_slot(t, i & 7)] = ldg4(src + i);
  }
  const float* w1 = a.dec + ((g >> 1) ? C_W1 : S_W1) + (g & 1) * 32;  // columns of this scale (coarse | fine, decoders.py:84)
  for (int i = threadIdx.x; i < 16 * 32; i += 256) sW[i] = w1[(i >> 5) * 64 + (i & 31)];
  __syncthreads();
  const int t = threadIdx.x >> 1, half = threadIdx.x & 1;
  if (t >= cnt) return;
  float o[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[j] = 0.f;
  const float* w = sW + half * 8 * 32;
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) {
    const float4 x = sT[t_slot(t, c4)];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 wj = lds4(w + j * 32 + c4 * 4);
      o[j] = fmaf(wj.w, x.w, fmaf(wj.z, x.z, fmaf(wj.y, x.y, fmaf(wj.x, x.x, o[j]))));
    }
  }
  float4* dst = reinterpret_cast<float4*>(a.q2) + (tbase + t) * 4 + half * 2;
  dst[0] = make_float4(o[0], o[1], o[2], o[3]);
  dst[1] = make_float4(o[4], o[5], o[6], o[7]);
}

struct SmemFwdQ {
  float4 P[2][NP * 4];  // first-layer pre-activations (without bias) of the sdf / rgb decoder
  ax_t ax_i[12][NP];    // axis set-ups of the four resolution groups: [group*3 + axis]
  float ax_f[12][NP];
  float one[NP], w[NP], z[NP], c[3][NP];
};

// ---- dense tail of a mapping iteration in the Q form -------------------------------------------------------------
// The backward kernel of the Q form reduces the 16-channel gradient of the first layer's pre-activations into GQ
// images (layout of the Q arena).  Per texel the chain rule back to the parameters is dense and tiny:
//     d loss / d plane[texel][c]   = sum_j W1[j][scale*32 + c] * GQ[texel][j]
//     d loss / d W1[j][scale*32+c] = sum over the three planes of that scale and all texels of GQ[texel][j] * plane[texel][c]
// so it is fused with the plane half of Adam: the plane gradient lives in registers only.
// One CTA = QA_TILE consecutive texels of one group, ONE THREAD PER TEXEL: the thread scans its texel's 64-byte
// gradient row and `touched` byte; an active texel (non-zero row, or moments that are already non-zero) then takes
// the whole 32-channel update in that thread, so a tile costs two dependent memory round trips however many of its
// texels are active (an 8-lanes-per-texel form serialised up to 8 trips per warp and ran 40 us).  Exact skip as in
// k_adam: a texel whose GQ has been zero since the optimiser was created has m = v = 0 and a zero update; a tile
// without active texels ends after the scan.  dW1: rows and pre-update texels of the tile are parked in shared
// memory and thread (j, c4) sums its four outputs over the tile's non-zero rows, one 16-byte reduction per thread
// into the gradient arena's decoder block (the decoders then take the ordinary Adam step, k_q_build follows).
struct QAdamArgs {
  QGroups qg;      // units = tiles of QA_TILE texels
  float4* arena4;  // parameters; the planes are updated in place
  float4* gq4;     // gradient images, zeroed where consumed
  float4 *m4, *v4; // Adam moments in parameter-arena layout
  float* gdec;     // gradient arena's decoder block: dW1 is added here
  const float* dec;
  unsigned char* touched;  // one flag per texel, in texel order
  float step_sdf, step_rgb;  // lr / (1 - beta1^t) of the sdf / rgb planes
  AdamArgs adam;             // scalars only
};

constexpr int QA_TILE = 128;

__global__ void __launch_bounds__(QA_TILE, 4) k_q_adam_planes(const __grid_constant__ QAdamArgs a) {
  __shared__ __align__(16) float sW[16 * 32];
  __shared__ float4 sP[QA_TILE * 8];   // pre-update texels, t_slot layout
  __shared__ float4 sG[QA_TILE * 4];   // gradient rows, p_slot layout
  __shared__ int s_list[QA_TILE];
  __shared__ int s_n;
  const int g = q_group_of(a.qg, blockIdx.x);
  const int t = threadIdx.x;
  const int tl = (blockIdx.x - a.qg.unit0[g]) * QA_TILE + t;
  const bool valid = tl < a.qg.n[g];
  const long long texel = (long long)a.qg.t0[g] + tl;
  // ---- scan
  float4 r[4];
  bool nz = false, flag = false;
  if (valid) {
    const float4* row = a.gq4 + texel * 4;
#pragma unroll
    for (int c = 0; c < 4; ++c) r[c] = row[c];
    flag = a.touched[texel] != 0;
#pragma unroll
    for (int c = 0; c < 4; ++c) nz = nz || r[c].x != 0.f || r[c].y != 0.f || r[c].z != 0.f || r[c].w != 0.f;
    if (nz && !flag) a.touched[texel] = 1;
  }
  const bool active = nz || flag;
  if (t == 0) s_n = 0;
  if (!__syncthreads_or(active)) return;  // nothing to do in this tile
  const float* w1 = a.dec + ((g >> 1) ? C_W1 : S_W1) + (g & 1) * 32;
  for (int i = t; i < 16 * 32; i += QA_TILE) sW[i] = w1[(i >> 5) * 64 + (i & 31)];
  float4 pk[8];
  if (active) {
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) pk[c4] = a.arena4[texel * 8 + c4];
  }
  if (nz) {
    s_list[atomicAdd(&s_n, 1)] = t;
#pragma unroll
    for (int c = 0; c < 4; ++c) sG[p_slot(t, c)] = r[c];
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) sP[t_slot(t, c4)] = pk[c4];
  }
  __syncthreads();
  if (active) {
    const float ss = (g >> 1) ? a.step_rgb : a.step_sdf;
    const float gj[16] = {r[0].x, r[0].y, r[0].z, r[0].w, r[1].x, r[1].y, r[1].z, r[1].w,
                          r[2].x, r[2].y, r[2].z, r[2].w, r[3].x, r[3].y, r[3].z, r[3].w};
    const float4 z4 = f4_zero();
#pragma unroll
    for (int c4 = 0; c4 < 8; ++c4) {
      const long long at = texel * 8 + c4;
      float4 m = a.m4[at], v = a.v4[at], p = pk[c4];
      float4 gr = z4;
      if (nz) {
#pragma unroll
        for (int j = 0; j < 16; ++j) gr = f4_fma(gj[j], lds4(sW + j * 32 + c4 * 4), gr);
      }
      adam_one(p.x, gr.x, m.x, v.x, a.adam, ss);
      adam_one(p.y, gr.y, m.y, v.y, a.adam, ss);
      adam_one(p.z, gr.z, m.z, v.z, a.adam, ss);
      adam_one(p.w, gr.w, m.w, v.w, a.adam, ss);
      a.arena4[at] = p;
      a.m4[at] = m;
      a.v4[at] = v;
    }
    if (nz) {
      float4* row = a.gq4 + texel * 4;
#pragma unroll
      for (int c = 0; c < 4; ++c) row[c] = z4;
    }
  }
  // ---- dW1 of the tile: thread (j, c4) over the non-zero rows
  const int n = s_n;
  if (n == 0) return;
  const int j = t >> 3, c4 = t & 7;
  float4 acc = f4_zero();
  const float* sGf = reinterpret_cast<const float*>(sG);
  for (int i = 0; i < n; ++i) {
    const int tt = s_list[i];
    const float gv = sGf[p_slot(tt, j >> 2) * 4 + (j & 3)];
    acc = f4_fma(gv, sP[t_slot(tt, c4)], acc);
  }
  float4* dst = reinterpret_cast<float4*>(a.gdec + ((g >> 1) ? C_W1 : S_W1) + (g & 1) * 32);
  if (acc.x != 0.f || acc.y != 0.f || acc.z != 0.f || acc.w != 0.f) red_add_v4(dst + j * 16 + c4, acc);
}

// layers 2 and 3 on h1 = relu(pre + b1), weights as constant-memory operands (field.cuh mlp_forward, minus layer 1)
template <int B1, int W2, int B2, int W3, int B3, int NOUT>
__device__ __forceinline__ void mlp_tail(const float4* __restrict__ P, int q, float (&h1)[16], float (&h2)[16],
                                         float (&out)[NOUT]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 v = P[p_slot(q, c)];
    h1[c * 4 + 0] = fmaxf(v.x + c_dec[B1 + c * 4 + 0], 0.f);
    h1[c * 4 + 1] = fmaxf(v.y + c_dec[B1 + c * 4 + 1], 0.f);
    h1[c * 4 + 2] = fmaxf(v.z + c_dec[B1 + c * 4 + 2], 0.f);
    h1[c * 4 + 3] = fmaxf(v.w + c_dec[B1 + c * 4 + 3], 0.f);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float a = c_dec[B2 + j];
#pragma unroll
    for (int i = 0; i < 16; ++i) a = fmaf(c_dec[W2 + j * 16 + i], h1[i], a);
    h2[j] = fmaxf(a, 0.f);
  }
#pragma unroll
  for (int o = 0; o < NOUT; ++o) {
    float a = c_dec[B3 + o];
#pragma unroll
    for (int i = 0; i < 16; ++i) a = fmaf(c_dec[W3 + o * 16 + i], h2[i], a);
    out[o] = a;
  }
}

struct RenderFwdQArgs {
  FieldK fk;
  const float4* q4;
  const float *rays_o, *rays_d, *z;
  int n_rays, S;
  const int* counters;
  float *depth, *rgb, *sdf;
  float4* act4;    // optional, as in RenderFwdArgs
  unsigned* actm;
};

__global__ void __launch_bounds__(NP) k_render_fwd_q(const __grid_constant__ RenderFwdQArgs a) {
  __shared__ SmemFwdQ sm;
  const int R = a.counters ? min(a.counters[0], a.n_rays) : a.n_rays;
  const int S = a.S;
  const int RPB = NP / S;
  const int ray0 = blockIdx.x * RPB;
  if (ray0 >= R) return;
  const int rays_here = min(RPB, R - ray0);
  const int n_valid = rays_here * S;
  const int q = threadIdx.x;
  const bool valid = q < n_valid;
  const int rl = valid ? q / S : 0;
  const int k = q - rl * S;
  const int ray = ray0 + rl;
  float pn[3] = {0.f, 0.f, 0.f}, zk = 0.f;
  if (valid) {
    zk = a.z[(long long)ray * S + k];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float p = __fadd_rn(a.rays_o[ray * 3 + c], __fmul_rn(a.rays_d[ray * 3 + c], zk));
      pn[c] = normalize_axis(p, a.fk.lo[c], a.fk.hi[c]);
    }
  }
  write_axis_setups<4>(a.fk, 0, pn, sm.ax_i, sm.ax_f, q);
  __syncthreads();
  {  // gather layout: 4 lanes per point, 8 points per warp and trip, both decoders
    const int warp = q >> 5, lane = q & 31, grp = lane >> 2, sub = lane & 3;
#pragma unroll 1
    for (int it = 0; it < 4; ++it) {
      const int qq = warp * 32 + it * 8 + grp;
      float4 p0 = f4_zero(), p1 = f4_zero();
      if (qq < n_valid) {
        p0 = gather_preact<0>(a.fk, a.q4, sm.ax_i, sm.ax_f, qq, sub);
        p1 = gather_preact<1>(a.fk, a.q4, sm.ax_i, sm.ax_f, qq, sub);
      }
      sm.P[0][p_slot(qq, sub)] = p0;
      sm.P[1][p_slot(qq, sub)] = p1;
    }
  }
  __syncthreads();
  float h1[16], h2[16], os[1], oc[3];
  mlp_tail<S_B1, S_W2, S_B2, S_W3, S_B3, 1>(sm.P[0], q, h1, h2, os);
  const float sdf = tanhf(os[0]);
  unsigned mask_s = 0u, mask_c = 0u;
  if (a.act4) mask_s = relu_mask(h1, h2);
  mlp_tail<C_B1, C_W2, C_B2, C_W3, C_B3, 3>(sm.P[1], q, h1, h2, oc);
  if (a.act4) mask_c = relu_mask(h1, h2);
  const float beta = c_dec[P_BETA];
  float u, e, alpha;
  sdf_to_alpha(sdf, beta, u, e, alpha);
  sm.one[q] = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
  sm.z[q] = zk;
#pragma unroll
  for (int c = 0; c < 3; ++c) sm.c[c][q] = sigmoidf_(oc[c]);
  if (a.act4 && valid) {
    a.act4[(long long)ray * S + k] = make_float4(sm.c[0][q], sm.c[1][q], sm.c[2][q], __uint_as_float(mask_s));
    a.actm[(long long)ray * S + k] = mask_c;
  }
  __syncthreads();
  float T = 1.0f;
  for (int j = 0; j < k; ++j) T *= sm.one[rl * S + j];
  sm.w[q] = valid ? alpha * T : 0.f;
  if (valid && a.sdf) a.sdf[(long long)ray * S + k] = sdf;
  __syncthreads();
  if (valid && k < 4) {
    const float* v = (k == 0) ? sm.z : sm.c[k - 1];
    float acc = 0.f;
    for (int j = 0; j < S; ++j) acc = fmaf(sm.w[rl * S + j], v[rl * S + j], acc);
    if (k == 0)
      a.depth[ray] = acc;
    else
      a.rgb[ray * 3 + (k - 1)] = acc;
  }
}

}  // namespace eslam
