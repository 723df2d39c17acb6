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

struct QBuildArgs {
  FieldK fk;
  const float4* arena4;
  const float* dec;  // packed decoder block (include/eslam_b200.h)
  float4* q4;
};

// grid (x, 12): blockIdx.y = plane; 8 lanes per texel, lane `sub` holds input channels 4*sub..4*sub+3
__global__ void __launch_bounds__(256) k_q_build(const __grid_constant__ QBuildArgs a) {
  __shared__ __align__(16) float sW[16 * 32];
  const int pi = blockIdx.y;
  const int field = pi / 6, scale = (pi % 6) / 3;
  const int off4 = a.fk.pl[pi].off4;
  const long long n = (long long)a.fk.pl[pi].H * a.fk.pl[pi].W;
  const float* w1 = a.dec + (field ? C_W1 : S_W1) + scale * 32;  // columns of this scale (coarse | fine, decoders.py:84)
  for (int i = threadIdx.x; i < 16 * 32; i += 256) sW[i] = w1[(i >> 5) * 64 + (i & 31)];
  __syncthreads();
  const int sub = threadIdx.x & 7;
  const long long n4 = (n + 3) & ~3ll;  // whole warps (4 texels each) per trip: the shuffles below are convergent
  for (long long idx = (long long)blockIdx.x * 32 + (threadIdx.x >> 3); idx < n4; idx += (long long)gridDim.x * 32) {
    const bool valid = idx < n;
    const long long id = valid ? idx : n - 1;
    const float4 t = ldg4(a.arena4 + off4 + id * 8 + sub);
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 wj = lds4(sW + j * 32 + sub * 4);
      acc[j] = fmaf(wj.w, t.w, fmaf(wj.z, t.z, fmaf(wj.y, t.y, wj.x * t.x)));
    }
#pragma unroll
    for (int off = 1; off < 8; off <<= 1)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
    if (valid && sub < 4) {
      float4 o;
      o.x = sub == 0 ? acc[0] : sub == 1 ? acc[4] : sub == 2 ? acc[8] : acc[12];
      o.y = sub == 0 ? acc[1] : sub == 1 ? acc[5] : sub == 2 ? acc[9] : acc[13];
      o.z = sub == 0 ? acc[2] : sub == 1 ? acc[6] : sub == 2 ? acc[10] : acc[14];
      o.w = sub == 0 ? acc[3] : sub == 1 ? acc[7] : sub == 2 ? acc[11] : acc[15];
      a.q4[(long long)(off4 >> 1) + id * 4 + sub] = o;
    }
  }
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
// so it is fused with the plane half of Adam: the plane gradient lives in registers only.  grid (x, 12): blockIdx.y =
// plane; 8 lanes per texel, lane `sub` owns channels 4*sub..4*sub+3 of the texel (p, m, v) and the matching 16 x 4
// slice of dW1, accumulated in registers over the CTA's texels and reduced once per CTA into the gradient arena's
// decoder block (the decoders then take the ordinary Adam step, and k_q_build follows with the new W1).
// Exact skip as in k_adam: a group of 4 texels whose GQ has been zero since the optimiser was created has m = v = 0
// and a zero update; `touched` here is one flag per 4 texels of a plane (tq_base[plane] + texel / 4).
struct QAdamArgs {
  FieldK fk;
  float4* arena4;  // parameters; the planes are updated in place
  float4* gq4;     // gradient images, zeroed where consumed
  float4 *m4, *v4; // Adam moments in parameter-arena layout
  float* gdec;     // gradient arena's decoder block: dW1 is added here
  const float* dec;
  unsigned char* touched;
  int tq_base[12];
  float step_sdf, step_rgb;  // lr / (1 - beta1^t) of the sdf / rgb planes
  AdamArgs adam;             // scalars only
};

__global__ void __launch_bounds__(256) k_q_adam_planes(const __grid_constant__ QAdamArgs a) {
  __shared__ __align__(16) float sW[16 * 32];
  __shared__ float sdW[16 * 32];
  const int pi = blockIdx.y;
  const int field = pi / 6, scale = (pi % 6) / 3;
  const int off4 = a.fk.pl[pi].off4;
  const long long n = (long long)a.fk.pl[pi].H * a.fk.pl[pi].W;
  const float* w1 = a.dec + (field ? C_W1 : S_W1) + scale * 32;
  for (int i = threadIdx.x; i < 16 * 32; i += 256) {
    sW[i] = w1[(i >> 5) * 64 + (i & 31)];
    sdW[i] = 0.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, sub = threadIdx.x & 7;
  const float ss = field ? a.step_rgb : a.step_sdf;
  float acc[16][4];
#pragma unroll
  for (int j = 0; j < 16; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
  const float4 z4 = f4_zero();
  const long long n4 = (n + 3) & ~3ll;  // whole warps (4 texels = one touched group) per trip
  for (long long idx = (long long)blockIdx.x * 32 + (threadIdx.x >> 3); idx < n4; idx += (long long)gridDim.x * 32) {
    const bool in = idx < n;
    const long long id = in ? idx : n - 1;
    float4* gq = a.gq4 + (long long)(off4 >> 1) + id * 4;
    float4 g4[4];
    bool nz = false;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      g4[c] = in ? gq[c] : z4;
      nz = nz || g4[c].x != 0.f || g4[c].y != 0.f || g4[c].z != 0.f || g4[c].w != 0.f;
    }
    const bool any = __ballot_sync(0xffffffffu, nz) != 0u;
    unsigned char* flag = a.touched + a.tq_base[pi] + (idx >> 2);  // same byte for the whole warp
    const bool was = *flag != 0;
    if (!any && !was) continue;
    if (!was && lane == 0) *flag = 1;
    if (!in) continue;
    const long long at = (long long)off4 + id * 8 + sub;
    float4 p = a.arena4[at], m = a.m4[at], v = a.v4[at];
    const float gj[16] = {g4[0].x, g4[0].y, g4[0].z, g4[0].w, g4[1].x, g4[1].y, g4[1].z, g4[1].w,
                          g4[2].x, g4[2].y, g4[2].z, g4[2].w, g4[3].x, g4[3].y, g4[3].z, g4[3].w};
    float4 g = z4;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 wj = lds4(sW + j * 32 + sub * 4);
      g = f4_fma(gj[j], wj, g);
      acc[j][0] = fmaf(gj[j], p.x, acc[j][0]);
      acc[j][1] = fmaf(gj[j], p.y, acc[j][1]);
      acc[j][2] = fmaf(gj[j], p.z, acc[j][2]);
      acc[j][3] = fmaf(gj[j], p.w, acc[j][3]);
    }
    adam_one(p.x, g.x, m.x, v.x, a.adam, ss);
    adam_one(p.y, g.y, m.y, v.y, a.adam, ss);
    adam_one(p.z, g.z, m.z, v.z, a.adam, ss);
    adam_one(p.w, g.w, m.w, v.w, a.adam, ss);
    a.arena4[at] = p;
    a.m4[at] = m;
    a.v4[at] = v;
    if (any && sub < 4) gq[sub] = z4;
  }
  // dW1: lanes with the same `sub` (the 4 texels of a warp) first, then the CTA's warps, then one reduction per element
#pragma unroll
  for (int j = 0; j < 16; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      float v = acc[j][e];
      v += __shfl_xor_sync(0xffffffffu, v, 8);
      v += __shfl_xor_sync(0xffffffffu, v, 16);
      if (lane < 8 && v != 0.f) atomicAdd(&sdW[j * 32 + sub * 4 + e], v);
    }
  __syncthreads();
  float* dst = a.gdec + (field ? C_W1 : S_W1) + scale * 32;
  for (int i = threadIdx.x; i < 16 * 32; i += 256)
    if (sdW[i] != 0.f) atomicAdd(dst + (i >> 5) * 64 + (i & 31), sdW[i]);
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
