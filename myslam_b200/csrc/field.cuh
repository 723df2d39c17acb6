// Device-side building blocks shared by every kernel of the hot path: the plane table, the
// bilinear tap set-up of grid_sample(border, align_corners=True), the 8-lanes-per-point gather,
// and the two register-resident decoder MLPs reading their weights from a shared-memory copy.
//
// Thread layouts used throughout (one CTA = NP points = whole rays):
//   "point layout":  thread q owns point q (MLPs, compositing, losses)
//   "gather layout": 8 consecutive lanes own one point, lane `sub` holds channels 4*sub..4*sub+3 of
//                    every texel (one float4 of the 128-byte channels-last texel line), so a warp
//                    instruction touches 4 full lines.
// Features cross between the two layouts through a swizzled shared-memory tile F[NP][64].
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/eslam_b200.h"

namespace eslam {

constexpr int NP = 128;      // points per CTA (= threads per CTA)
constexpr int NWARP = NP / 32;
typedef short ax_t;          // texel index along one axis (plane extents are far below 32768)

// ---- packed decoder offsets (include/eslam_b200.h) -------------------------------------------------
constexpr int S_W1 = 0, S_B1 = 1024, S_W2 = 1040, S_B2 = 1296, S_W3 = 1312, S_B3 = 1328;
constexpr int C_W1 = 1332, C_B1 = 2356, C_W2 = 2372, C_B2 = 2628, C_W3 = 2644, C_B3 = 2692;
constexpr int P_BETA = ESLAM_DEC_BETA;
constexpr int DEC_N = ESLAM_DEC_FLOATS;

struct PlaneK {
  int off4;  // float4 offset of texel (0,0) in the arena
  int H, W;
};

struct FieldK {
  PlaneK pl[12];
  float lo[3], hi[3];
  long long dec_off;
};

inline int make_field_k(const eslam_field_t* f, FieldK* out) {
  for (int i = 0; i < 12; ++i) {
    if (f->plane[i].offset % 4 != 0 || f->plane[i].H < 1 || f->plane[i].W < 1) return ESLAM_EINVAL;
    if (f->plane[i].H > 32767 || f->plane[i].W > 32767) return ESLAM_EUNSUPPORTED;
    if (f->plane[i].offset / 4 + (long long)f->plane[i].H * f->plane[i].W * 8 > 0x7fffffffLL) return ESLAM_EUNSUPPORTED;
    out->pl[i].off4 = (int)(f->plane[i].offset / 4);
    out->pl[i].H = f->plane[i].H;
    out->pl[i].W = f->plane[i].W;
  }
  for (int a = 0; a < 3; ++a) {
    out->lo[a] = f->bound[a][0];
    out->hi[a] = f->bound[a][1];
  }
  out->dec_off = f->dec_offset;
  return 0;
}

// sizes (nx, ny, nz) of the resolution group g (0 sdf coarse, 1 sdf fine, 2 rgb coarse, 3 rgb fine):
// xy is [ny][nx], xz is [nz][nx]
__device__ __forceinline__ int axis_size(const FieldK& fk, int g, int axis) {
  const PlaneK& xy = fk.pl[g * 3 + 0];
  const PlaneK& xz = fk.pl[g * 3 + 1];
  return axis == 0 ? xy.W : (axis == 1 ? xy.H : xz.H);
}

// normalize_3d_coordinate, common.py:204-218
__device__ __forceinline__ float normalize_axis(float p, float lo, float hi) {
  return __fsub_rn(__fmul_rn(__fdiv_rn(__fsub_rn(p, lo), __fsub_rn(hi, lo)), 2.0f), 1.0f);
}

// grid_sampler_compute_source_index(border, align_corners=True): ATen/native/GridSampler.h:27-70
__device__ __forceinline__ void axis_setup(float pn, int size, int& i0, float& fr) {
  float x = __fmul_rn(__fdiv_rn(__fadd_rn(pn, 1.0f), 2.0f), (float)(size - 1));
  x = fminf((float)(size - 1), fmaxf(x, 0.0f));
  float fl = floorf(x);
  i0 = (int)fl;
  fr = x - fl;
}

// d(source index)/d(pn): (size-1)/2 unless clipped (GridSampler.h clip_coordinates_set_grad uses <=0, >=max)
__device__ __forceinline__ float axis_grad_mult(int i0, float fr, int size) {
  bool clipped = (i0 >= size - 1) || (i0 == 0 && fr == 0.0f);
  return clipped ? 0.0f : 0.5f * (float)(size - 1);
}

__device__ __forceinline__ float4 ldg4(const float4* p) { return __ldg(p); }

__device__ __forceinline__ void red_add_v4(float4* addr, float4 v) {
  // no "memory" clobber: the gradient arena is never read by the kernels that reduce into it, and the clobber
  // would pin every later corner load behind the reductions (exposing one L2 round trip per tap)
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w));
}

__device__ __forceinline__ float4 f4_zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }
__device__ __forceinline__ float4 f4_fma(float s, float4 v, float4 a) {
  return make_float4(fmaf(s, v.x, a.x), fmaf(s, v.y, a.y), fmaf(s, v.z, a.z), fmaf(s, v.w, a.w));
}
__device__ __forceinline__ float4 f4_mul(float s, float4 v) { return make_float4(s * v.x, s * v.y, s * v.z, s * v.w); }
__device__ __forceinline__ float4 f4_add(float4 a, float4 b) {
  return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4_sub(float4 a, float4 b) {
  return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
__device__ __forceinline__ float f4_dot(float4 a, float4 b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }

// swizzled feature tile: row q holds 16 float4; physical slot = c4 ^ (q & 7)
__device__ __forceinline__ int f_slot(int q, int c4) { return q * 16 + (c4 ^ (q & 7)); }
__device__ __forceinline__ float f_scalar(const float4* F, int q, int c) {
  return reinterpret_cast<const float*>(F)[f_slot(q, c >> 2) * 4 + (c & 3)];
}

// One tap of one point in gather layout.
struct Tap {
  int base;    // float4 index of (v0,u0) texel + sub
  int dx, dy;  // float4 strides to the +1 neighbours (0 at the border)
  float fu, fv;
};

__device__ __forceinline__ Tap make_tap(const PlaneK& pl, int u0, float fu, int v0, float fv, int sub) {
  Tap t;
  t.base = pl.off4 + (v0 * pl.W + u0) * 8 + sub;
  t.dx = (u0 + 1 < pl.W) ? 8 : 0;
  t.dy = (v0 + 1 < pl.H) ? pl.W * 8 : 0;
  t.fu = fu;
  t.fv = fv;
  return t;
}

// axis pair of plane p (0 xy, 1 xz, 2 yz): first coordinate indexes W (grid_sample's x), second H
__device__ __forceinline__ int pair_u(int p) { return p == 2 ? 1 : 0; }
__device__ __forceinline__ int pair_v(int p) { return p == 0 ? 1 : 2; }

// Gather the 64 features of one decoder (FIELD 0 sdf, 1 rgb) for point q; this lane's 4 channels.
// ax_i/ax_f: shared arrays [g*3+axis][NP] for the resolution groups of this decoder, indexed by
// (scale*3+axis) + AXBASE.
template <int AXBASE>
__device__ __forceinline__ void gather_features(const FieldK& fk, int field, const float4* __restrict__ arena4,
                                                const ax_t (*ax_i)[NP], const float (*ax_f)[NP], int q, int sub,
                                                float4& out_coarse, float4& out_fine) {
  Tap tp[6];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const int au = AXBASE + s * 3 + pair_u(p), av = AXBASE + s * 3 + pair_v(p);
      tp[s * 3 + p] = make_tap(fk.pl[field * 6 + s * 3 + p], ax_i[au][q], ax_f[au][q], ax_i[av][q], ax_f[av][q], sub);
    }
  }
  float4 v[6][4];
#pragma unroll
  for (int t = 0; t < 6; ++t) {
    v[t][0] = ldg4(arena4 + tp[t].base);
    v[t][1] = ldg4(arena4 + tp[t].base + tp[t].dx);
    v[t][2] = ldg4(arena4 + tp[t].base + tp[t].dy);
    v[t][3] = ldg4(arena4 + tp[t].base + tp[t].dy + tp[t].dx);
  }
  float4 acc[2];
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    float4 sum = f4_zero();
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const int t = s * 3 + p;
      const float fu = tp[t].fu, fv = tp[t].fv;
      const float w00 = (1.f - fu) * (1.f - fv), w01 = fu * (1.f - fv), w10 = (1.f - fu) * fv, w11 = fu * fv;
      float4 tap = f4_mul(w00, v[t][0]);
      tap = f4_fma(w01, v[t][1], tap);
      tap = f4_fma(w10, v[t][2], tap);
      tap = f4_fma(w11, v[t][3], tap);
      sum = (p == 0) ? tap : f4_add(sum, tap);  // (xy + xz) + yz, decoders.py:82
    }
    acc[s] = sum;
  }
  out_coarse = acc[0];
  out_fine = acc[1];
}

// decoder weights for the forward-only kernels, filled by eslam_bind_decoders (one translation unit)
__constant__ __align__(16) float c_dec[DEC_N];

// ---- forward MLP, weights as constant-memory operands, fully unrolled -----------------------------------
// Used by the forward-only kernels (decode, render forward, importance): there the ~4 k straight-line
// instructions stay resident in the instruction cache and no shared-memory bandwidth is spent on weights
// (84 us vs 139 us for 4000 rays with the shared-memory form below).  The fused backward kernel, whose code is
// three times larger, uses the looped shared-memory form instead (441 -> 358 us).

template <int W1, int B1, int W2, int B2, int W3, int B3, int NOUT>
__device__ __forceinline__ void mlp_forward(const float4* __restrict__ F, int q, float (&h1)[16], float (&h2)[16],
                                            float (&out)[NOUT]) {
#pragma unroll
  for (int j = 0; j < 16; ++j) h1[j] = c_dec[B1 + j];
  const int swz = q & 7;
  const float4* row = F + q * 16;
#pragma unroll
  for (int c4 = 0; c4 < 16; ++c4) {
    const float4 f = row[c4 ^ swz];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      h1[j] = fmaf(c_dec[W1 + j * 64 + c4 * 4 + 0], f.x, h1[j]);
      h1[j] = fmaf(c_dec[W1 + j * 64 + c4 * 4 + 1], f.y, h1[j]);
      h1[j] = fmaf(c_dec[W1 + j * 64 + c4 * 4 + 2], f.z, h1[j]);
      h1[j] = fmaf(c_dec[W1 + j * 64 + c4 * 4 + 3], f.w, h1[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) h1[j] = fmaxf(h1[j], 0.f);
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

// ---- MLPs with the weights in shared memory (point layout) ------------------------------------------------
// One symmetric block per decoder, DW_STRIDE floats: W1[16][64] b1[16] W2[16][16] b2[16] W3[3][16] b3[3] (+pad);
// the sdf decoder's W3 rows 1-2 and b3[1..2] are zero, so both decoders run the SAME code with a different base
// pointer.  Loops over the 16 float4 feature chunks keep the code ~20x smaller than the fully unrolled
// constant-operand form (which stalled on instruction fetch, profiles/r01_bwd_full_summary.txt).
constexpr int DW_W1 = 0, DW_B1 = 1024, DW_W2 = 1040, DW_B2 = 1296, DW_W3 = 1312, DW_B3 = 1360, DW_STRIDE = 1364;
constexpr int DW_BETA = 2 * DW_STRIDE, DW_TOTAL = 2 * DW_STRIDE + 4;

// packed arena block (include/eslam_b200.h) -> symmetric shared block; all threads of the CTA cooperate.
// Asynchronous 16-byte copies (cp.async): issued at the top of the kernel, they land while the points are set up and
// the features gathered; decoder_weights_wait() + __syncthreads() must precede the first use.
// Both blocks start on 16-byte boundaries in the arena (offsets 0 and 1332 floats) and in shared memory (0, 1364).
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void load_decoder_weights(float* sW, const float* __restrict__ dec, int tid, int nthreads) {
  constexpr int SDF4 = (DW_W3 + 16) / 4;  // 332 float4: W1 b1 W2 b2 W3[0]
  constexpr int RGB4 = (DW_B3 + 3) / 4;   // 340 float4: W1 b1 W2 b2 W3[0..2] (b3 is the 3-float tail)
  for (int i = tid; i < SDF4 + RGB4; i += nthreads) {
    if (i < SDF4)
      cp_async16(sW + 4 * i, dec + 4 * i);
    else
      cp_async16(sW + DW_STRIDE + 4 * (i - SDF4), dec + C_W1 + 4 * (i - SDF4));
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  // the few scalars outside the 16-byte pattern
  if (tid < 32) sW[DW_W3 + 16 + tid] = 0.f;  // sdf W3 rows 1-2 (the sdf decoder has one output)
  if (tid == 32) sW[DW_B3] = dec[S_B3];
  if (tid == 33) sW[DW_B3 + 1] = 0.f;
  if (tid == 34) sW[DW_B3 + 2] = 0.f;
  if (tid == 35) sW[DW_B3 + 3] = 0.f;
  if (tid >= 36 && tid < 39) sW[DW_STRIDE + DW_B3 + (tid - 36)] = dec[C_B3 + (tid - 36)];
  if (tid == 39) sW[DW_STRIDE + DW_B3 + 3] = 0.f;
  if (tid == 40) sW[DW_BETA] = dec[P_BETA];
}
__device__ __forceinline__ void decoder_weights_wait() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

__device__ __forceinline__ float4 lds4(const float* p) { return *reinterpret_cast<const float4*>(p); }

__device__ __forceinline__ void mlp_forward_s(const float* __restrict__ W, const float4* __restrict__ F, int q,
                                              float (&h1)[16], float (&h2)[16], float (&out)[3]) {
#pragma unroll
  for (int j4 = 0; j4 < 4; ++j4) {
    const float4 b = lds4(W + DW_B1 + j4 * 4);
    h1[j4 * 4 + 0] = b.x;
    h1[j4 * 4 + 1] = b.y;
    h1[j4 * 4 + 2] = b.z;
    h1[j4 * 4 + 3] = b.w;
  }
  const int swz = q & 7;
  const float4* row = F + q * 16;
#pragma unroll 1
  for (int c4 = 0; c4 < 16; ++c4) {
    const float4 f = row[c4 ^ swz];
    const float* w = W + DW_W1 + c4 * 4;
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float4 wj = lds4(w + j * 64);
      h1[j] = fmaf(wj.x, f.x, h1[j]);
      h1[j] = fmaf(wj.y, f.y, h1[j]);
      h1[j] = fmaf(wj.z, f.z, h1[j]);
      h1[j] = fmaf(wj.w, f.w, h1[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 16; ++j) h1[j] = fmaxf(h1[j], 0.f);
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

__device__ __forceinline__ void mlp_backward_hidden_s(const float* __restrict__ W, const float (&gout)[3],
                                                      const float (&h1)[16], const float (&h2)[16], float (&ga1)[16],
                                                      float (&ga2)[16]) {
#pragma unroll
  for (int i4 = 0; i4 < 4; ++i4) {
    float4 g = f4_zero();
#pragma unroll
    for (int o = 0; o < 3; ++o) g = f4_fma(gout[o], lds4(W + DW_W3 + o * 16 + i4 * 4), g);
    ga2[i4 * 4 + 0] = h2[i4 * 4 + 0] > 0.f ? g.x : 0.f;
    ga2[i4 * 4 + 1] = h2[i4 * 4 + 1] > 0.f ? g.y : 0.f;
    ga2[i4 * 4 + 2] = h2[i4 * 4 + 2] > 0.f ? g.z : 0.f;
    ga2[i4 * 4 + 3] = h2[i4 * 4 + 3] > 0.f ? g.w : 0.f;
  }
#pragma unroll
  for (int i4 = 0; i4 < 4; ++i4) {
    float4 g = f4_zero();
#pragma unroll
    for (int j = 0; j < 16; ++j) g = f4_fma(ga2[j], lds4(W + DW_W2 + j * 16 + i4 * 4), g);
    ga1[i4 * 4 + 0] = h1[i4 * 4 + 0] > 0.f ? g.x : 0.f;
    ga1[i4 * 4 + 1] = h1[i4 * 4 + 1] > 0.f ? g.y : 0.f;
    ga1[i4 * 4 + 2] = h1[i4 * 4 + 2] > 0.f ? g.z : 0.f;
    ga1[i4 * 4 + 3] = h1[i4 * 4 + 3] > 0.f ? g.w : 0.f;
  }
}

__device__ __forceinline__ void mlp_backward_input_s(const float* __restrict__ W, const float (&ga1)[16],
                                                     float4* __restrict__ F, int q) {
  const int swz = q & 7;
  float4* row = F + q * 16;
#pragma unroll 1
  for (int c4 = 0; c4 < 16; ++c4) {
    float4 g = f4_zero();
    const float* w = W + DW_W1 + c4 * 4;
#pragma unroll
    for (int j = 0; j < 16; ++j) g = f4_fma(ga1[j], lds4(w + j * 64), g);
    row[c4 ^ swz] = g;
  }
}

// ---- small helpers -----------------------------------------------------------------------------------

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// sdf-band of a sample (Tracker.py:129-141 == Mapper.py:125-137): 0 front, 1 center, 2 tail, 3 none (back)
__device__ __forceinline__ int sdf_band(float z, float d, float tr, float tr04) {
  if (z < __fsub_rn(d, tr)) return 0;
  if (z > __fsub_rn(d, tr04) && z < __fadd_rn(d, tr04)) return 1;
  if (z > __fadd_rn(d, tr)) return 3;
  return 2;
}

}  // namespace eslam
