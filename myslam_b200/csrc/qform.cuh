// Device-side building blocks of the Q form: the first decoder layer applied to the planes
// (16-channel images Q = W1_slice . plane, 64 bytes per texel) instead of to every sample.  Included by render.cuh;
// the kernels that use them are in qplane.cuh (image build, forward, optimiser tail) and qbwd.cuh (backward).
// Gather layout of the Q form: 4 lanes per point, lane `sub` holds pre-activations 4*sub..4*sub+3.
#pragma once
#include "field.cuh"

namespace eslam {

// pre-activation tile: row q holds 4 float4; physical slot = c ^ ((q >> 1) & 3),
// conflict-free both for 4-lane writers (two consecutive rows per quarter warp) and point-layout readers
__device__ __forceinline__ int p_slot(int q, int c) { return q * 4 + (c ^ ((q >> 1) & 3)); }

// Sum over the 6 planes of decoder FIELD of the bilinear fetch from its Q images; this lane's 4 pre-activations.
template <int FIELD>
__device__ __forceinline__ float4 gather_preact(const FieldK& fk, const float4* __restrict__ q4, const ax_t (*ax_i)[NP],
                                                const float (*ax_f)[NP], int qq, int sub) {
  float4 v[6][4];
  float fu[6], fv[6];
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
      const int base = (pl.off4 >> 1) + (v0 * pl.W + u0) * 4 + sub;
      const int dx = (u0 + 1 < pl.W) ? 4 : 0, dy = (v0 + 1 < pl.H) ? pl.W * 4 : 0;
      v[t][0] = ldg4(q4 + base);
      v[t][1] = ldg4(q4 + base + dx);
      v[t][2] = ldg4(q4 + base + dy);
      v[t][3] = ldg4(q4 + base + dy + dx);
    }
  }
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
    }
    acc[s] = sum;
  }
  return f4_add(acc[0], acc[1]);  // coarse + fine
}

// fill the pre-activation tile of decoder FIELD for the NP slots of a CTA (or of one half of the backward kernel's CTA)
template <int FIELD>
__device__ __forceinline__ void gather_preact_tile(const FieldK& fk, const float4* __restrict__ q4,
                                                   const ax_t (*ax_i)[NP], const float (*ax_f)[NP], float4* P,
                                                   int n_valid, int tid) {
  const int warp = tid >> 5, lane = tid & 31, grp = lane >> 2, sub = lane & 3;
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    const int qq = warp * 32 + it * 8 + grp;
    float4 p = f4_zero();
    if (qq < n_valid) p = gather_preact<FIELD>(fk, q4, ax_i, ax_f, qq, sub);
    P[p_slot(qq, sub)] = p;
  }
}

// layers 2 and 3 on h1 = relu(P + b1) with the weights in shared memory (field.cuh mlp_forward_s minus its first layer)
__device__ __forceinline__ void mlp_tail_s(const float* __restrict__ W, const float4* __restrict__ P, int q,
                                           float (&h1)[16], float (&h2)[16], float (&out)[3]) {
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 v = P[p_slot(q, c)];
    const float4 b = lds4(W + DW_B1 + c * 4);
    h1[c * 4 + 0] = fmaxf(v.x + b.x, 0.f);
    h1[c * 4 + 1] = fmaxf(v.y + b.y, 0.f);
    h1[c * 4 + 2] = fmaxf(v.z + b.z, 0.f);
    h1[c * 4 + 3] = fmaxf(v.w + b.w, 0.f);
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    float a = W[DW_B2 + j];
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
      const float4 w = lds4(W + DW_W2 + j * 16 + i4 * 4);
      a = fmaf(w.x, h1[i4 * 4 + 0], a);
      a = fmaf(w.y, h1[i4 * 4 + 1], a);
      a = fmaf(w.z, h1[i4 * 4 + 2], a);
      a = fmaf(w.w, h1[i4 * 4 + 3], a);
    }
    h2[j] = fmaxf(a, 0.f);
  }
#pragma unroll
  for (int o = 0; o < 3; ++o) {
    float a = W[DW_B3 + o];
#pragma unroll
    for (int i4 = 0; i4 < 4; ++i4) {
      const float4 w = lds4(W + DW_W3 + o * 16 + i4 * 4);
      a = fmaf(w.x, h2[i4 * 4 + 0], a);
      a = fmaf(w.y, h2[i4 * 4 + 1], a);
      a = fmaf(w.z, h2[i4 * 4 + 2], a);
      a = fmaf(w.w, h2[i4 * 4 + 3], a);
    }
    out[o] = a;
  }
}

constexpr int QW_STRIDE = DW_STRIDE - DW_B1;  // b1 W2 b2 W3 b3 of one decoder (field.cuh's block minus W1): 340 floats
constexpr int QW_TOTAL = 2 * QW_STRIDE + 4;

template <int BYTES>
__device__ __forceinline__ void cp_async_small(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(d), "l"(gmem_src), "n"(BYTES) : "memory");
}

// b1 W2 b2 W3 b3 of both decoders -> shared memory (cp.async; decoder_weights_wait() + barrier before first use).
// The packed arena block (include/eslam_b200.h) holds them contiguously: sdf floats [1024, 1329), rgb [2356, 2695).
__device__ __forceinline__ void load_tail_weights(float* sW, const float* __restrict__ dec, int tid, int nthreads) {
  constexpr int SDF4 = 304 / 4;  // b1 W2 b2 W3[0]
  constexpr int RGB4 = 336 / 4;  // b1 W2 b2 W3[0..2]
  for (int i = tid; i < SDF4 + RGB4; i += nthreads) {
    if (i < SDF4)
      cp_async16(sW + 4 * i, dec + S_B1 + 4 * i);
    else
      cp_async16(sW + QW_STRIDE + 4 * (i - SDF4), dec + C_B1 + 4 * (i - SDF4));
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  if (tid < 32) sW[304 + tid] = 0.f;  // sdf W3 rows 1-2 (the sdf decoder has one output)
  if (tid == 32) sW[336] = dec[S_B3];
  if (tid >= 33 && tid < 36) sW[336 + (tid - 32)] = 0.f;
  if (tid >= 36 && tid < 39) sW[QW_STRIDE + 336 + (tid - 36)] = dec[C_B3 + (tid - 36)];
  if (tid == 39) sW[QW_STRIDE + 339] = 0.f;
  if (tid == 40) sW[2 * QW_STRIDE] = dec[P_BETA];
}

// Q form of the coordinate-gradient half of scatter_group (pose-only backward): P holds the
// gradient at the first layer's pre-activations (16 per point, p_slot layout); 4 lanes per point fetch the corners of
// the 16-channel Q images and the coordinate gradient is d/du of bilinear(Q) . g, with the clip rule of scatter_group.
template <int FIELD>
__device__ __forceinline__ void coord_grads_q(const FieldK& fk, const float4* __restrict__ q4, const ax_t (*ax_i)[NP],
                                              const float (*ax_f)[NP], const float4* P, int wl, int grp, int sub,
                                              int n_valid, float (*gp)[NP]) {
#pragma unroll 1
  for (int it = 0; it < 4; ++it) {
    const int q = wl * 32 + it * 8 + grp;
    float gpn[3] = {0.f, 0.f, 0.f};
    if (q < n_valid) {
      const float4 g4 = P[p_slot(q, sub)];
#pragma unroll
      for (int sc = 0; sc < 2; ++sc) {
        float4 v[3][4];
        int u0[3], v0[3];
        float fu[3], fv[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const int au = FIELD * 6 + sc * 3 + pair_u(p), av = FIELD * 6 + sc * 3 + pair_v(p);
          const PlaneK& pl = fk.pl[FIELD * 6 + sc * 3 + p];
          u0[p] = ax_i[au][q];
          v0[p] = ax_i[av][q];
          fu[p] = ax_f[au][q];
          fv[p] = ax_f[av][q];
          const int base = (pl.off4 >> 1) + (v0[p] * pl.W + u0[p]) * 4 + sub;
          const int dx = (u0[p] + 1 < pl.W) ? 4 : 0, dy = (v0[p] + 1 < pl.H) ? pl.W * 4 : 0;
          v[p][0] = ldg4(q4 + base);
          v[p][1] = ldg4(q4 + base + dx);
          v[p][2] = ldg4(q4 + base + dy);
          v[p][3] = ldg4(q4 + base + dy + dx);
        }
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const PlaneK& pl = fk.pl[FIELD * 6 + sc * 3 + p];
          const float d00 = f4_dot(g4, v[p][0]), d01 = f4_dot(g4, v[p][1]);
          const float d10 = f4_dot(g4, v[p][2]), d11 = f4_dot(g4, v[p][3]);
          const float du = (d01 - d00) * (1.f - fv[p]) + (d11 - d10) * fv[p];
          const float dv = (d10 - d00) * (1.f - fu[p]) + (d11 - d01) * fu[p];
          gpn[pair_u(p)] = fmaf(du, axis_grad_mult(u0[p], fu[p], pl.W), gpn[pair_u(p)]);
          gpn[pair_v(p)] = fmaf(dv, axis_grad_mult(v0[p], fv[p], pl.H), gpn[pair_v(p)]);
        }
      }
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      float v = gpn[c];
      v += __shfl_xor_sync(0xffffffffu, v, 1);
      v += __shfl_xor_sync(0xffffffffu, v, 2);
      if (sub == 0) gp[c][q] = v;
    }
  }
}

// Q form of scatter_group's reductions for the mapping backward: P holds the gradient at the first layer's
// pre-activations.  8-lane groups over 8 consecutive points of a ray as in scatter_group; within a group lanes 0-3 own
// the coarse scale and lanes 4-7 the fine scale of the SAME points (lane & 3 = float4 of the 16-channel texel), so one
// code path serves both scales and the run-length merge of the coarse cells keeps its run length: tap-major
// reductions into the GQ images, four corner sums kept in registers while the cell repeats.
// (The coordinate gradients need no corner fetch here: qbwd.cuh keeps J = d pre-activation / d coordinate.)
template <int FIELD>
__device__ __forceinline__ void scatter_q(const FieldK& fk, float4* __restrict__ gq4, const ax_t (*ax_i)[NP],
                                          const float (*ax_f)[NP], const float4* P, int qb, int n_valid, int sub8) {
  const int sc = sub8 >> 2, sub = sub8 & 3;
  const int axb = FIELD * 6 + sc * 3;
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const PlaneK pl = fk.pl[axb + p];
    const int qoff = pl.off4 >> 1;
    const int au = axb + pair_u(p), av = axb + pair_v(p);
    int cur = -1, cdx = 0, cdy = 0;
    float4 a00 = f4_zero(), a01 = f4_zero(), a10 = f4_zero(), a11 = f4_zero();
#pragma unroll 1
    for (int it = 0; it < 8; ++it) {
      const int q = qb + it;
      if (q < n_valid) {
        const int u0 = ax_i[au][q], v0 = ax_i[av][q];
        const float fu = ax_f[au][q], fv = ax_f[av][q];
        const int base = qoff + (v0 * pl.W + u0) * 4 + sub;
        const float4 g4 = P[p_slot(q, sub)];
        const float w00 = (1.f - fu) * (1.f - fv), w01 = fu * (1.f - fv), w10 = (1.f - fu) * fv, w11 = fu * fv;
        if (base != cur) {
          if (cur >= 0) {
            red_add_v4(gq4 + cur, a00);
            red_add_v4(gq4 + cur + cdx, a01);
            red_add_v4(gq4 + cur + cdy, a10);
            red_add_v4(gq4 + cur + cdy + cdx, a11);
          }
          cur = base;
          cdx = (u0 + 1 < pl.W) ? 4 : 0;
          cdy = (v0 + 1 < pl.H) ? pl.W * 4 : 0;
          a00 = f4_mul(w00, g4);
          a01 = f4_mul(w01, g4);
          a10 = f4_mul(w10, g4);
          a11 = f4_mul(w11, g4);
        } else {
          a00 = f4_fma(w00, g4, a00);
          a01 = f4_fma(w01, g4, a01);
          a10 = f4_fma(w10, g4, a10);
          a11 = f4_fma(w11, g4, a11);
        }
      }
    }
    if (cur >= 0) {
      red_add_v4(gq4 + cur, a00);
      red_add_v4(gq4 + cur + cdx, a01);
      red_add_v4(gq4 + cur + cdy, a10);
      red_add_v4(gq4 + cur + cdy + cdx, a11);
    }
  }
}

}  // namespace eslam
