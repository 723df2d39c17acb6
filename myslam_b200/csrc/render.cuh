// Render family: point decode, ray forward (compositing) and the fused loss + backward.
// Reference semantics: src/networks/decoders.py:64-146, src/utils/Renderer.py:136-153,
// src/Tracker.py:114-148,192-208, src/Mapper.py:110-144,337-349.
#pragma once
#include "field.cuh"
#include "qform.cuh"

namespace eslam {

// ---------------------------------------------------------------------------------------------------
// shared-memory tiles
// ---------------------------------------------------------------------------------------------------
struct SmemFwd {
  float4 F[NP * 16];   // feature tile of the decoder being evaluated
  ax_t ax_i[6][NP];     // axis set-ups of that decoder's two resolution groups: [scale*3+axis]
  float ax_f[6][NP];
  float one[NP], w[NP], z[NP], c[3][NP];
};

#ifdef ESLAM_PROFILE_PHASES
// per-phase wall clocks of the backward kernel, summed over (CTA, warp 0 / warp 4) into g_phase[half][phase]
__device__ unsigned long long g_phase[2][16];
#define PHASE_MARK(i)                                                                         \
  do {                                                                                        \
    if ((threadIdx.x & 127) == 0) {                                                           \
      const long long now_ = clock64();                                                       \
      atomicAdd(&g_phase[threadIdx.x >> 7][i], (unsigned long long)(now_ - phase_t0_));       \
      phase_t0_ = now_;                                                                       \
    }                                                                                         \
  } while (0)
#define PHASE_INIT() long long phase_t0_ = clock64()
#else
#define PHASE_MARK(i) do {} while (0)
#define PHASE_INIT() do {} while (0)
#endif

constexpr int WG_STRIDE = 20;  // floats per point in the staging buffers: 16 values + the output-layer gradients
constexpr int NT_BWD = 256;  // threads of the backward kernel: threads 0-127 own the sdf decoder, 128-255 the rgb one

template <bool GF>
struct SmemBwd {
  float4 F0[NP * 16];  // sdf features, later d loss / d sdf features
  float4 F1[NP * 16];  // rgb
  ax_t ax_i[12][NP];
  float ax_f[12][NP];
  float W[DW_TOTAL];   // both decoders' weights, symmetric blocks (field.cuh)
  float act0[GF ? NP * WG_STRIDE : 4];  // gradient / activation staging for the weight-gradient products
  float act1[GF ? NP * WG_STRIDE : 4];
  float one[NP], w[NP], z[NP], c[3][NP], gww[NP];
  float gp[2][3][NP];  // d loss / d normalised coordinate, per decoder half
  float rayv[4][16];   // per ray: rendered depth, r, g, b
  float rayg[4][16];   // per ray: upstream g_depth, g_rgb
  float rayd[16];      // per ray gt depth
  int raym[16];        // per ray loss-mask flag
  float rayod[6][16];  // per ray d loss / d (o, d)
  float red[NT_BWD / 32];
  double redd[(NT_BWD / 32) * 5];
};

// ---------------------------------------------------------------------------------------------------
// common phases
// ---------------------------------------------------------------------------------------------------

// point layout: axis set-ups of resolution groups [g0, g0+NG) into ax rows [0, 3*NG)
template <int NG>
__device__ __forceinline__ void write_axis_setups(const FieldK& fk, int g0, const float (&pn)[3], ax_t (*ax_i)[NP],
                                                  float (*ax_f)[NP], int q) {
#pragma unroll
  for (int g = 0; g < NG; ++g) {
#pragma unroll
    for (int a = 0; a < 3; ++a) {
      int i0;
      float fr;
      axis_setup(pn[a], axis_size(fk, g0 + g, a), i0, fr);
      ax_i[g * 3 + a][q] = (ax_t)i0;
      ax_f[g * 3 + a][q] = fr;
    }
  }
}

// gather layout: fill the feature tile of decoder `field` for all NP slots of this CTA
template <int AXBASE>
__device__ __forceinline__ void gather_tile(const FieldK& fk, int field, const float4* __restrict__ arena4,
                                            const ax_t (*ax_i)[NP], const float (*ax_f)[NP], float4* F, int n_valid,
                                            int tid = threadIdx.x) {
  const int warp = tid >> 5, lane = tid & 31;
  const int grp = lane >> 3, sub = lane & 7;
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    const int qq = warp * 32 + it * 4 + grp;
    float4 fc = f4_zero(), ff = f4_zero();
    if (qq < n_valid) gather_features<AXBASE>(fk, field, arena4, ax_i, ax_f, qq, sub, fc, ff);
    F[f_slot(qq, sub)] = fc;
    F[f_slot(qq, 8 + sub)] = ff;
  }
}

// ---------------------------------------------------------------------------------------------------
// decode points / grid sdf  (Decoders.forward; Mesher.eval_points)
// ---------------------------------------------------------------------------------------------------
struct DecodeArgs {
  FieldK fk;
  const float4* arena4;
  const float* pts;  // [n][3] or NULL in grid mode
  long long n;
  float* raw;        // [n][4] (r,g,b,sdf) or NULL
  float* sdf_out;    // [n] or NULL
  int flags;         // bit0 sdf only, bit1 bound mask -> sdf=-1, bit2 pts are already normalised
  const float *xs, *ys, *zs;
  int nx, ny, nz;
  long long start;
  const float4* hull;  // flags bit3: n_hull half-spaces (nx, ny, nz, d), inside iff n.p + d <= 0 for all of them
  int n_hull;
  // flags bit4 (grid mode, sdf only): per-plane features resampled once on the lattice's three faces (k_grid_features)
  const float4 *fxy, *fxz, *fyz;  // [ny][nx][16], [nz][nx][16], [nz][ny][16] float4 (coarse | fine)
};

// ---------------------------------------------------------------------------------------------------
// separable lattice query ("mesh query separability")
// ---------------------------------------------------------------------------------------------------
// On the regular marching-cubes lattice every bilinear tap depends on two lattice indices only:
//   feat(ix, iy, iz) = (Fxy[iy][ix] + Fxz[iz][ix]) + Fyz[iz][iy]      per scale, in decoders.py:82's order,
// so the three planes are resampled ONCE on the lattice's faces (this kernel: the same tap arithmetic as
// gather_features, hence bit-identical features) and a voxel costs three 256-byte reads and two adds per channel
// instead of 24 corner fetches and their interpolation.
struct GridFeatArgs {
  FieldK fk;
  const float4* arena4;
  const float *us, *vs;  // lattice coordinates along the plane's first (W) and second (H) axis
  int na, nb, plane, ua, va;
  float4* out;           // [nb][na][16]
};

__global__ void __launch_bounds__(256) k_grid_features(const __grid_constant__ GridFeatArgs a) {
  const long long idx = (long long)blockIdx.x * 32 + (threadIdx.x >> 3);
  const int sub = threadIdx.x & 7;
  if (idx >= (long long)a.na * a.nb) return;
  const int ia = (int)(idx % a.na), ib = (int)(idx / a.na);
  const float pu = normalize_axis(a.us[ia], a.fk.lo[a.ua], a.fk.hi[a.ua]);
  const float pv = normalize_axis(a.vs[ib], a.fk.lo[a.va], a.fk.hi[a.va]);
#pragma unroll
  for (int s = 0; s < 2; ++s) {
    int u0, v0;
    float fu, fv;
    axis_setup(pu, axis_size(a.fk, s, a.ua), u0, fu);
    axis_setup(pv, axis_size(a.fk, s, a.va), v0, fv);
    const Tap t = make_tap(a.fk.pl[s * 3 + a.plane], u0, fu, v0, fv, sub);
    const float4 v00 = ldg4(a.arena4 + t.base), v01 = ldg4(a.arena4 + t.base + t.dx);
    const float4 v10 = ldg4(a.arena4 + t.base + t.dy), v11 = ldg4(a.arena4 + t.base + t.dy + t.dx);
    const float w00 = (1.f - fu) * (1.f - fv), w01 = fu * (1.f - fv), w10 = (1.f - fu) * fv, w11 = fu * fv;
    float4 tap = f4_mul(w00, v00);
    tap = f4_fma(w01, v01, tap);
    tap = f4_fma(w10, v10, tap);
    tap = f4_fma(w11, v11, tap);
    a.out[idx * 16 + s * 8 + sub] = tap;
  }
}


// ---------------------------------------------------------------------------------------------------
// factored lattice query
// ---------------------------------------------------------------------------------------------------
// The decoder's first layer is linear in the summed feature, so on the lattice it factors over the faces as well:
//   W1 (Fxy + Fxz + Fyz) + b1 = (W1 Fxy + b1) + W1 Fxz + W1 Fyz.
// k_grid_preact resamples each plane pair on its face (same taps as gather_features) and applies W1 there: 16 values
// per face texel instead of 64 channels.  A voxel then costs three 64-byte reads, 32 adds and the 16 -> 16 -> 1 tail of
// the MLP: 272 FMA instead of 1 296, 192 B instead of 768 B.  Unlike the separable form this is NOT bit-identical with
// decoders.py:109-125: the sum over the 64 inputs is re-associated into three partial sums (a few ulp of the
// pre-activation); the tests hold it to 1e-5 of the direct form and to the 1e-4 bar against the reference's values.
// Layout: pxy [ny][nx][4] float4 (one texel, broadcast to a z column; b1 folded in), pxz [nx][4][nz], pyz [ny][4][nz]:
// the lanes of a warp walk z, so each of the four float4 components is a coalesced 512-byte row.
struct GridPreArgs {
  FieldK fk;
  const float4* arena4;
  const float* w1;       // the sdf decoder's W1[16][64] with b1[16] right behind it (packed decoder block)
  const float *us, *vs;  // lattice coordinates along the plane's first (W) and second (H) axis
  int na, nb, plane, ua, va;
  int zmajor;            // 0: out[ib][ia][4] with b1 added; 1: out[ia][4][ib]
  float4* out;
};

__global__ void __launch_bounds__(256) k_grid_preact(const __grid_constant__ GridPreArgs a) {
  __shared__ __align__(16) float sW[16 * 64 + 16];
  for (int i = threadIdx.x; i < 16 * 64 + 16; i += 256) sW[i] = a.w1[i];
  __syncthreads();
  const int sub = threadIdx.x & 7;
  const long long n = (long long)a.na * a.nb;
  const long long n4 = (n + 3) & ~3ll;  // whole warps (4 texels each) per trip: the shuffles below are convergent
  for (long long idx = (long long)blockIdx.x * 32 + (threadIdx.x >> 3); idx < n4; idx += (long long)gridDim.x * 32) {
    const bool valid = idx < n;
    const long long id = valid ? idx : n - 1;
    int ia, ib;
    if (a.zmajor) {
      ia = (int)(id / a.nb);
      ib = (int)(id % a.nb);
    } else {
      ia = (int)(id % a.na);
      ib = (int)(id / a.na);
    }
    const float pu = normalize_axis(a.us[ia], a.fk.lo[a.ua], a.fk.hi[a.ua]);
    const float pv = normalize_axis(a.vs[ib], a.fk.lo[a.va], a.fk.hi[a.va]);
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = 0.f;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
      int u0, v0;
      float fu, fv;
      axis_setup(pu, axis_size(a.fk, s, a.ua), u0, fu);
      axis_setup(pv, axis_size(a.fk, s, a.va), v0, fv);
      const Tap t = make_tap(a.fk.pl[s * 3 + a.plane], u0, fu, v0, fv, sub);
      const float4 v00 = ldg4(a.arena4 + t.base), v01 = ldg4(a.arena4 + t.base + t.dx);
      const float4 v10 = ldg4(a.arena4 + t.base + t.dy), v11 = ldg4(a.arena4 + t.base + t.dy + t.dx);
      const float w00 = (1.f - fu) * (1.f - fv), w01 = fu * (1.f - fv), w10 = (1.f - fu) * fv, w11 = fu * fv;
      float4 tap = f4_mul(w00, v00);
      tap = f4_fma(w01, v01, tap);
      tap = f4_fma(w10, v10, tap);
      tap = f4_fma(w11, v11, tap);
      const float* w = sW + s * 32 + sub * 4;  // input channels s*32 + 4*sub .. +3 (coarse | fine, decoders.py:99-106)
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float4 wj = lds4(w + j * 64);
        acc[j] = fmaf(wj.x, tap.x, acc[j]);
        acc[j] = fmaf(wj.y, tap.y, acc[j]);
        acc[j] = fmaf(wj.z, tap.z, acc[j]);
        acc[j] = fmaf(wj.w, tap.w, acc[j]);
      }
    }
#pragma unroll
    for (int off = 1; off < 8; off <<= 1)
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
    if (!a.zmajor) {
#pragma unroll
      for (int j = 0; j < 16; ++j) acc[j] += sW[16 * 64 + j];
    }
    if (valid && sub < 4) {
      float4 o;
      o.x = sub == 0 ? acc[0] : sub == 1 ? acc[4] : sub == 2 ? acc[8] : acc[12];
      o.y = sub == 0 ? acc[1] : sub == 1 ? acc[5] : sub == 2 ? acc[9] : acc[13];
      o.z = sub == 0 ? acc[2] : sub == 1 ? acc[6] : sub == 2 ? acc[10] : acc[14];
      o.w = sub == 0 ? acc[3] : sub == 1 ? acc[7] : sub == 2 ? acc[11] : acc[15];
      const long long at = a.zmajor ? ((long long)ia * 4 + sub) * a.nb + ib : id * 4 + sub;
      a.out[at] = o;
    }
  }
}

struct GridFacArgs {
  float lo[3], hi[3];
  const float *xs, *ys, *zs;
  int nx, ny, nz;
  long long start, n;
  const float4 *pxy, *pxz, *pyz;
  const float4* hull;
  int n_hull;
  float* sdf_out;
};

constexpr int FAC_THREADS = 256;
constexpr int FAC_PT = 2;  // voxels per thread (FAC_THREADS apart, so every load stays a coalesced row): the uniform
                           // loads of the layer-2/3 weights and the CTA's index set-up are shared between them

__global__ void __launch_bounds__(FAC_THREADS) k_grid_sdf_factored(const __grid_constant__ GridFacArgs a) {
  __shared__ int s_idx[3];
  const long long base = (long long)blockIdx.x * (FAC_THREADS * FAC_PT);
  if (threadIdx.x == 0) {  // one 64-bit division per CTA; the per-thread index math below is 32-bit
    const long long f0 = a.start + base;
    const long long t0 = f0 / a.nz;
    const int iy0 = (int)(t0 / a.nx);
    s_idx[0] = (int)(t0 - (long long)iy0 * a.nx);
    s_idx[1] = iy0;
    s_idx[2] = (int)(f0 - t0 * a.nz);
  }
  __syncthreads();
  const unsigned nx = (unsigned)a.nx, nz = (unsigned)a.nz;
  long long gi[FAC_PT];
  bool valid[FAC_PT], inside[FAC_PT];
  const float4 *qxy[FAC_PT], *qxz[FAC_PT], *qyz[FAC_PT];
  bool any_inside = false;
#pragma unroll
  for (int k = 0; k < FAC_PT; ++k) {
    gi[k] = base + k * FAC_THREADS + threadIdx.x;
    valid[k] = gi[k] < a.n;
    unsigned x = (unsigned)s_idx[0], y = (unsigned)s_idx[1], z = (unsigned)s_idx[2] + k * FAC_THREADS + threadIdx.x;
    if (z >= nz) {  // the CTA's range runs over the end of a z column (and possibly of an x row)
      const unsigned w = z / nz;
      z -= w * nz;
      x += w;
      if (x >= nx) {
        const unsigned d = x / nx;
        x -= d * nx;
        y += d;
      }
    }
    if (!valid[k]) x = y = z = 0u;
    const float p[3] = {a.xs[x], a.ys[y], a.zs[z]};
    bool in = valid[k];
#pragma unroll
    for (int c = 0; c < 3; ++c) in = in && (p[c] < a.hi[c]) && (p[c] > a.lo[c]);  // Mesher.py:214-215
    for (int h = 0; h < a.n_hull && in; ++h) {  // Mesher.py:210-217: mesh_bound.contains -> sdf = -1
      const float4 hp = __ldg(a.hull + h);
      in = fmaf(hp.x, p[0], fmaf(hp.y, p[1], fmaf(hp.z, p[2], hp.w))) <= 0.f;
    }
    inside[k] = in;
    any_inside = any_inside || in;
    qxy[k] = a.pxy + ((long long)y * nx + x) * 4;
    qxz[k] = a.pxz + (long long)x * 4 * nz + z;
    qyz[k] = a.pyz + (long long)y * 4 * nz + z;
  }
  if (!__syncthreads_or(any_inside)) {  // the whole tile lies outside: nothing to decode
#pragma unroll
    for (int k = 0; k < FAC_PT; ++k)
      if (valid[k]) a.sdf_out[gi[k]] = -1.0f;
    return;
  }
  static_assert(FAC_PT == 2, "the hidden layer is packed FP32 over the thread's two voxels");
  float2 h1[16];  // channel i of voxel 0 / 1 in .x / .y
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 v0 = f4_add(f4_add(ldg4(qxy[0] + c), ldg4(qxz[0] + (long long)c * nz)), ldg4(qyz[0] + (long long)c * nz));
    const float4 v1 = f4_add(f4_add(ldg4(qxy[1] + c), ldg4(qxz[1] + (long long)c * nz)), ldg4(qyz[1] + (long long)c * nz));
    h1[c * 4 + 0] = make_float2(fmaxf(v0.x, 0.f), fmaxf(v1.x, 0.f));
    h1[c * 4 + 1] = make_float2(fmaxf(v0.y, 0.f), fmaxf(v1.y, 0.f));
    h1[c * 4 + 2] = make_float2(fmaxf(v0.z, 0.f), fmaxf(v1.z, 0.f));
    h1[c * 4 + 3] = make_float2(fmaxf(v0.w, 0.f), fmaxf(v1.w, 0.f));
  }
  // fma.rn.f32x2 (SASS FFMA2 R, R.F32x2, UR.F32, R): both voxels against one weight per issue slot; two exact FMAs, the
  // operation order of the scalar form
  float2 o2 = make_float2(c_dec[S_B3], c_dec[S_B3]);
#pragma unroll
  for (int j = 0; j < 16; ++j) {
    const float bj = c_dec[S_B2 + j];
    float2 acc = make_float2(bj, bj);
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float w = c_dec[S_W2 + j * 16 + i];
      acc = __ffma2_rn(make_float2(w, w), h1[i], acc);
    }
    const float w3 = c_dec[S_W3 + j];
    o2 = __ffma2_rn(make_float2(w3, w3), make_float2(fmaxf(acc.x, 0.f), fmaxf(acc.y, 0.f)), o2);
  }
  const float o[FAC_PT] = {o2.x, o2.y};
#pragma unroll
  for (int k = 0; k < FAC_PT; ++k) {
    float sdf = tanhf(o[k]);
    if (!inside[k]) sdf = -1.0f;
    if (valid[k]) a.sdf_out[gi[k]] = sdf;
  }
}

__global__ void __launch_bounds__(NP) k_decode(const __grid_constant__ DecodeArgs a) {
  __shared__ SmemFwd sm;
  const int q = threadIdx.x;
  const long long base = (long long)blockIdx.x * NP;
  const long long gi = base + q;
  const int n_valid = (int)min((long long)NP, a.n - base);
  const bool valid = q < n_valid;
  float p[3] = {0.f, 0.f, 0.f}, pn[3];
  if (valid) {
    if (a.pts) {
      p[0] = a.pts[gi * 3 + 0];
      p[1] = a.pts[gi * 3 + 1];
      p[2] = a.pts[gi * 3 + 2];
    } else {
      const long long f = a.start + gi;
      const int iz = (int)(f % a.nz);
      const long long t = f / a.nz;
      const int ix = (int)(t % a.nx), iy = (int)(t / a.nx);
      p[0] = a.xs[ix];
      p[1] = a.ys[iy];
      p[2] = a.zs[iz];
    }
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) pn[k] = (a.flags & 4) ? p[k] : normalize_axis(p[k], a.fk.lo[k], a.fk.hi[k]);
  bool inside = true;
  if (a.flags & 2) {
#pragma unroll
    for (int k = 0; k < 3; ++k) inside = inside && (p[k] < a.fk.hi[k]) && (p[k] > a.fk.lo[k]);
  }
  if (a.flags & 8) {  // convex mesh bound of the seen region (Mesher.py:210-217: mesh_bound.contains -> sdf = -1)
    for (int k = 0; k < a.n_hull && inside; ++k) {
      const float4 h = __ldg(a.hull + k);
      inside = fmaf(h.x, p[0], fmaf(h.y, p[1], fmaf(h.z, p[2], h.w))) <= 0.f;
    }
  }
  if ((a.flags & (2 | 8)) && (a.flags & 1) && !__syncthreads_or(valid && inside)) {
    if (valid && a.sdf_out) a.sdf_out[gi] = -1.0f;  // the whole tile lies outside: nothing to decode
    if (valid && a.raw) a.raw[gi * 4 + 3] = -1.0f;
    return;
  }
  // sdf decoder
  if (a.flags & 16) {  // separable lattice: sum the three resampled faces (k_grid_features)
    const long long f = a.start + gi;
    const long long t = f / a.nz;
    sm.ax_i[0][q] = (ax_t)(valid ? (int)(t % a.nx) : 0);
    sm.ax_i[1][q] = (ax_t)(valid ? (int)(t / a.nx) : 0);
    sm.ax_i[2][q] = (ax_t)(valid ? (int)(f % a.nz) : 0);
    __syncthreads();
    const int warp = q >> 5, lane = q & 31, grp = lane >> 3, sub = lane & 7;
#pragma unroll 2
    for (int it = 0; it < 8; ++it) {
      const int qq = warp * 32 + it * 4 + grp;
      float4 fc = f4_zero(), ff = f4_zero();
      if (qq < n_valid) {
        const int ix = sm.ax_i[0][qq], iy = sm.ax_i[1][qq], iz = sm.ax_i[2][qq];
        const float4* pxy = a.fxy + ((long long)iy * a.nx + ix) * 16 + sub;
        const float4* pxz = a.fxz + ((long long)iz * a.nx + ix) * 16 + sub;
        const float4* pyz = a.fyz + ((long long)iz * a.ny + iy) * 16 + sub;
        fc = f4_add(f4_add(ldg4(pxy), ldg4(pxz)), ldg4(pyz));  // (xy + xz) + yz, decoders.py:82
        ff = f4_add(f4_add(ldg4(pxy + 8), ldg4(pxz + 8)), ldg4(pyz + 8));
      }
      sm.F[f_slot(qq, sub)] = fc;
      sm.F[f_slot(qq, 8 + sub)] = ff;
    }
  } else {
    write_axis_setups<2>(a.fk, 0, pn, sm.ax_i, sm.ax_f, q);
    __syncthreads();
    gather_tile<0>(a.fk, 0, a.arena4, sm.ax_i, sm.ax_f, sm.F, n_valid);
  }
  __syncthreads();
  float h1[16], h2[16], os[1];
  mlp_forward<S_W1, S_B1, S_W2, S_B2, S_W3, S_B3, 1>(sm.F, q, h1, h2, os);
  float sdf = tanhf(os[0]);
  if (!inside) sdf = -1.0f;
  if (valid) {
    if (a.sdf_out) a.sdf_out[gi] = sdf;
    if (a.raw) a.raw[gi * 4 + 3] = sdf;
  }
  if (a.flags & 1) return;
  __syncthreads();
  write_axis_setups<2>(a.fk, 2, pn, sm.ax_i, sm.ax_f, q);
  __syncthreads();
  gather_tile<0>(a.fk, 1, a.arena4, sm.ax_i, sm.ax_f, sm.F, n_valid);
  __syncthreads();
  float oc[3];
  mlp_forward<C_W1, C_B1, C_W2, C_B2, C_W3, C_B3, 3>(sm.F, q, h1, h2, oc);
  if (valid && a.raw) {
    a.raw[gi * 4 + 0] = sigmoidf_(oc[0]);
    a.raw[gi * 4 + 1] = sigmoidf_(oc[1]);
    a.raw[gi * 4 + 2] = sigmoidf_(oc[2]);
  }
}

// Decoders.sample_plane_feature: feat[n][64] from normalised coordinates
struct FeatArgs {
  FieldK fk;
  const float4* arena4;
  const float* p_nor;
  long long n;
  int which;
  float4* feat4;
};

__global__ void __launch_bounds__(NP) k_plane_feature(const __grid_constant__ FeatArgs a) {
  __shared__ ax_t ax_i[6][NP];
  __shared__ float ax_f[6][NP];
  const int q = threadIdx.x;
  const long long base = (long long)blockIdx.x * NP;
  const int n_valid = (int)min((long long)NP, a.n - base);
  float pn[3] = {0.f, 0.f, 0.f};
  if (q < n_valid) {
    pn[0] = a.p_nor[(base + q) * 3 + 0];
    pn[1] = a.p_nor[(base + q) * 3 + 1];
    pn[2] = a.p_nor[(base + q) * 3 + 2];
  }
  write_axis_setups<2>(a.fk, a.which * 2, pn, ax_i, ax_f, q);
  __syncthreads();
  const int warp = q >> 5, lane = q & 31, grp = lane >> 3, sub = lane & 7;
  for (int it = 0; it < 8; ++it) {
    const int qq = warp * 32 + it * 4 + grp;
    if (qq < n_valid) {
      float4 fc, ff;
      gather_features<0>(a.fk, a.which, a.arena4, ax_i, ax_f, qq, sub, fc, ff);
      a.feat4[(base + qq) * 16 + sub] = fc;
      a.feat4[(base + qq) * 16 + 8 + sub] = ff;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// ray forward  (Renderer.py:136-147)
// ---------------------------------------------------------------------------------------------------
struct RenderFwdArgs {
  FieldK fk;
  const float4* arena4;
  const float *rays_o, *rays_d, *z;
  int n_rays, S;
  const int* counters;
  float *depth, *rgb, *sdf;
  float4* act4;    // optional [R][S]: (r, g, b, bits of the sdf decoder's ReLU masks) per sample, for the cached backward
  unsigned* actm;  // optional [R][S]: the rgb decoder's ReLU masks (bit j: h1[j] > 0, bit 16 + j: h2[j] > 0)
};

// bit j: h1[j] > 0, bit 16 + j: h2[j] > 0 -- all the backward pass needs of the hidden activations when no weight
// gradients are wanted
__device__ __forceinline__ unsigned relu_mask(const float (&h1)[16], const float (&h2)[16]) {
  unsigned m = 0u;
#pragma unroll
  for (int j = 0; j < 16; ++j) m |= (h1[j] > 0.f ? 1u << j : 0u) | (h2[j] > 0.f ? 1u << (16 + j) : 0u);
  return m;
}

__device__ __forceinline__ void sdf_to_alpha(float sdf, float beta, float& u, float& e, float& alpha) {
  u = sigmoidf_(-sdf * beta);  // Renderer.py:149-153
  e = expf(-beta * u);
  alpha = 1.0f - e;
}

__global__ void __launch_bounds__(NP) k_render_fwd(const __grid_constant__ RenderFwdArgs a) {
  __shared__ SmemFwd sm;
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
  write_axis_setups<2>(a.fk, 0, pn, sm.ax_i, sm.ax_f, q);
  __syncthreads();
  gather_tile<0>(a.fk, 0, a.arena4, sm.ax_i, sm.ax_f, sm.F, n_valid);
  __syncthreads();
  float h1[16], h2[16], os[1], oc[3];
  mlp_forward<S_W1, S_B1, S_W2, S_B2, S_W3, S_B3, 1>(sm.F, q, h1, h2, os);
  const float sdf = tanhf(os[0]);
  unsigned mask_s = 0u, mask_c = 0u;
  if (a.act4) mask_s = relu_mask(h1, h2);
  __syncthreads();
  write_axis_setups<2>(a.fk, 2, pn, sm.ax_i, sm.ax_f, q);
  __syncthreads();
  gather_tile<0>(a.fk, 1, a.arena4, sm.ax_i, sm.ax_f, sm.F, n_valid);
  __syncthreads();
  mlp_forward<C_W1, C_B1, C_W2, C_B2, C_W3, C_B3, 3>(sm.F, q, h1, h2, oc);
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

// ---------------------------------------------------------------------------------------------------
// fused loss + backward
// ---------------------------------------------------------------------------------------------------
// ---- the factored lattice query over whole lattice rows ---------------------------------------------------------------
// k_grid_sdf_factored is bound by issue slots: 590 instructions per voxel, 280 of them the FFMAs of the hidden layer,
// the rest loads, index arithmetic and the three-face sum.  Here a warp owns one x and 64 z (lane: z and z + 32) and
// walks y: the xz face values of its two voxels are the same eight lines every row (L1 hits; holding them in registers
// costs a resident CTA: 5.0 ms against 4.75), the yz values are shared by the CTA's warps (8 x) through L1, the xy
// texel is one line read by the whole warp; and the hidden layer is packed
// FP32 -- `fma.rn.f32x2` (SASS FFMA2 R, R.F32x2, UR.F32, R: the two voxels of a lane against one weight in a uniform
// register), two exact FMAs per issue slot, so the values are those of k_grid_sdf_factored bit for bit.
// (The same layer as mma.sync m16n8k8 TF32 with both operands split in two -- three products per tile, the precision
// the 1e-5 bar needs -- was measured first: 6.24 ms for the 990x680x490 lattice against 5.96 ms per voxel.  The legacy
// tensor path at 24 HMMA per 32 voxels is slower than 256 FFMA per voxel; tcgen05 would need the three-face sum staged
// through shared memory for a K = N = 16 tile.)
struct GridRowsArgs {
  float lo[3], hi[3];
  const float *xs, *ys, *zs;
  int nx, ny, nz;
  int iy0, iy1;        // lattice rows [iy0, iy1)
  long long out_base;  // flat lattice index of sdf_out[0]
  const float4 *pxy, *pxz, *pyz;
  const float4* hull;
  int n_hull;
  float* sdf_out;
};

constexpr int GR_THREADS = 256, GR_X = GR_THREADS / 32, GR_Z = 64;
#ifndef GR_RESIDENT
#define GR_RESIDENT 3  // 3: the xz values are re-read through L1 every row (<= 84 registers); 2: kept in registers
#endif

__device__ __forceinline__ float2 relu2(float2 v) { return make_float2(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f)); }

__global__ void __launch_bounds__(GR_THREADS, GR_RESIDENT) k_grid_sdf_rows(const __grid_constant__ GridRowsArgs a) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int x = blockIdx.x * GR_X + warp;
  if (x >= a.nx) return;  // warp-uniform; no CTA barrier below
  const int zA = blockIdx.y * GR_Z + lane, zB = zA + 32;
  const bool okA = zA < a.nz, okB = zB < a.nz;
  const int za = okA ? zA : a.nz - 1, zb = okB ? zB : a.nz - 1;
  // xz face values of the lane's two voxels: channel c of voxel A / B in .x / .y
#if GR_RESIDENT == 2
  float2 fxz[16];
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    const float4 va = ldg4(a.pxz + ((long long)x * 4 + c) * a.nz + za), vb = ldg4(a.pxz + ((long long)x * 4 + c) * a.nz + zb);
    fxz[c * 4 + 0] = make_float2(va.x, vb.x);
    fxz[c * 4 + 1] = make_float2(va.y, vb.y);
    fxz[c * 4 + 2] = make_float2(va.z, vb.z);
    fxz[c * 4 + 3] = make_float2(va.w, vb.w);
  }
#endif
  const float px = a.xs[x], pzA = a.zs[za], pzB = a.zs[zb];
  const bool in_x = px < a.hi[0] && px > a.lo[0];
  const bool inA0 = okA && in_x && pzA < a.hi[2] && pzA > a.lo[2], inB0 = okB && in_x && pzB < a.hi[2] && pzB > a.lo[2];
#pragma unroll 1
  for (int y = a.iy0; y < a.iy1; ++y) {
    float4 cxy[4], cyA[4], cyB[4];  // no software prefetch: the resident CTAs hide the loads
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      cxy[c] = ldg4(a.pxy + ((long long)y * a.nx + x) * 4 + c);
      cyA[c] = ldg4(a.pyz + ((long long)y * 4 + c) * a.nz + za);
      cyB[c] = ldg4(a.pyz + ((long long)y * 4 + c) * a.nz + zb);
    }
#if GR_RESIDENT != 2
    float2 fxz[16];  // the same 8 lines every row: L1 hits
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const float4 va = ldg4(a.pxz + ((long long)x * 4 + c) * a.nz + za), vb = ldg4(a.pxz + ((long long)x * 4 + c) * a.nz + zb);
      fxz[c * 4 + 0] = make_float2(va.x, vb.x);
      fxz[c * 4 + 1] = make_float2(va.y, vb.y);
      fxz[c * 4 + 2] = make_float2(va.z, vb.z);
      fxz[c * 4 + 3] = make_float2(va.w, vb.w);
    }
#endif
    const float py = a.ys[y];
    const bool in_y = py < a.hi[1] && py > a.lo[1];  // Mesher.py:214-215
    bool inA = inA0 && in_y, inB = inB0 && in_y;
    for (int h = 0; h < a.n_hull && (inA || inB); ++h) {  // Mesher.py:210-217: mesh_bound.contains -> sdf = -1
      const float4 hp = __ldg(a.hull + h);
      inA = inA && fmaf(hp.x, px, fmaf(hp.y, py, fmaf(hp.z, pzA, hp.w))) <= 0.f;
      inB = inB && fmaf(hp.x, px, fmaf(hp.y, py, fmaf(hp.z, pzB, hp.w))) <= 0.f;
    }
    float* dst = a.sdf_out + ((((long long)y * a.nx + x) * a.nz + zA) - a.out_base);
    if (!__any_sync(0xffffffffu, inA || inB)) {  // the warp's 64 voxels all lie outside: nothing to decode
      if (okA) dst[0] = -1.0f;
      if (okB) dst[32] = -1.0f;
      continue;
    }
    // h1 = relu((xy + xz) + yz), the sum order of k_grid_sdf_factored
    float2 h1[16];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      h1[c * 4 + 0] = relu2(__fadd2_rn(__fadd2_rn(make_float2(cxy[c].x, cxy[c].x), fxz[c * 4 + 0]), make_float2(cyA[c].x, cyB[c].x)));
      h1[c * 4 + 1] = relu2(__fadd2_rn(__fadd2_rn(make_float2(cxy[c].y, cxy[c].y), fxz[c * 4 + 1]), make_float2(cyA[c].y, cyB[c].y)));
      h1[c * 4 + 2] = relu2(__fadd2_rn(__fadd2_rn(make_float2(cxy[c].z, cxy[c].z), fxz[c * 4 + 2]), make_float2(cyA[c].z, cyB[c].z)));
      h1[c * 4 + 3] = relu2(__fadd2_rn(__fadd2_rn(make_float2(cxy[c].w, cxy[c].w), fxz[c * 4 + 3]), make_float2(cyA[c].w, cyB[c].w)));
    }
    float2 o = make_float2(c_dec[S_B3], c_dec[S_B3]);
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      const float bj = c_dec[S_B2 + j];
      float2 acc = make_float2(bj, bj);
#pragma unroll
      for (int i = 0; i < 16; ++i) {
        const float w = c_dec[S_W2 + j * 16 + i];
        acc = __ffma2_rn(make_float2(w, w), h1[i], acc);
      }
      const float w3 = c_dec[S_W3 + j];
      o = __ffma2_rn(make_float2(w3, w3), relu2(acc), o);
    }
    if (okA) dst[0] = inA ? tanhf(o.x) : -1.0f;
    if (okB) dst[32] = inB ? tanhf(o.y) : -1.0f;
  }
}

struct BwdArgs {
  FieldK fk;
  const float4* arena4;
  const float *rays_o, *rays_d, *z;
  int n_rays, S;
  const int* counters;
  int dbg;          // profiling aid (eslam_set_debug): bit0 skip plane reductions, bit1 skip weight gradients
  const int* norm;  // loss normalisers (counters layout); == counters on one GPU, all-reduced sums on several
  // upstream-gradient mode
  const float *g_depth, *g_rgb, *g_sdf;
  // fused-loss mode
  const float* gt_depth;
  const double* gt_color;
  const int* src;
  const long long* pix_idx;
  int n_per_img;
  const unsigned char* ray_mask;
  float tr, tr04, w_fs, w_center, w_tail, w_depth;
  double w_color;
  float fx, fy, cx, cy;
  int W0, H0, Wc;
  double* loss_acc;
  // cached forward (tracking): written by k_render_fwd for the same rays / samples; replaces gather + forward MLPs
  const float* sdf_in;
  const float4* act4;
  const unsigned* actm;
  // outputs
  float* grad_arena;
  float *g_rays_o, *g_rays_d;
  float* pose_grad;
  // Q form (qbwd.cuh): the Q images and, for the mapper, the gradient images the plane reductions go to
  const float4* q4;
  float4* gq4;
  // split launch of the mapper's backward (eslam_loss_backward_q_part): 0 = every tile, 1 = only the tiles whose rays
  // all carry a sensor depth; the tiles holding a depth-less ray (which wait for the importance samples) are
  // enumerated from dl_list by k_map_bwd_q_dl
  int part;
  const int* dl_list;
};

// ---- weight gradients on the tensor cores ----------------------------------------------------------------------
// dW[m][n] = sum over the tile's 128 points of grad[q][m] * act[q][n] is a 16 x N x 128 contraction per layer: dense,
// batched over points, and as scalar FMAs (two shared-memory loads each) it took 15 % of the kernel's issue slots.
// It runs as mma.sync m16n8k16 bf16 with every operand split into a bf16 head and middle (x = h + m + O(2^-16 x);
// h*h + m*h + h*m: ~2^-15 relative, inside the 1e-3 gradient bar where a single bf16 or TF32 product is not).
// A warp owns whole output tiles over all the points it is given; the bias gradients are the same contraction
// against a column of ones.
// x0, x1 -> packed bf16 pairs (low half = x0): head = bf16(x), middle = bf16(x - head)
__device__ __forceinline__ void split_bf16x2(float x0, float x1, uint32_t& hi, uint32_t& mid) {
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(x1), "f"(x0));
  const float h0 = __uint_as_float(hi << 16), h1 = __uint_as_float(hi & 0xffff0000u);
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(mid) : "f"(x1 - h1), "f"(x0 - h0));
}

__device__ __forceinline__ void mma_bf16(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// Operand sources for the mma fragments.  In a k-step of 16 points starting at q0 (a multiple of 16) a lane reads
// channel c of the point rows q0 + {2t, 2t+1, 2t+8, 2t+9}, so everything but q0 folds into four per-lane offsets
// computed once (for the swizzled feature tile the XOR term depends on the row's low 3 bits only).
struct FromBuf {  // staging buffer [NP][WG_STRIDE]
  const float* base;
  int o[4];
  __device__ __forceinline__ FromBuf(const float* buf, int c, int t) : base(buf) {
    const int r[4] = {2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9};
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = r[i] * WG_STRIDE + c;
  }
  __device__ __forceinline__ float at(int q0, int i) const { return base[q0 * WG_STRIDE + o[i]]; }
};
struct FromTile {  // swizzled tile F[NP][64]
  const float* base;
  int o[4];
  __device__ __forceinline__ FromTile(const float4* F, int c, int t) : base(reinterpret_cast<const float*>(F)) {
    const int r[4] = {2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9};
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = (r[i] * 16 + ((c >> 2) ^ (r[i] & 7))) * 4 + (c & 3);
  }
  __device__ __forceinline__ float at(int q0, int i) const { return base[q0 * 64 + o[i]]; }
};

// acc[i] (m16n8 fragments: rows g, g+8; columns 2t, 2t+1) = A^T B_i over the points [q_lo, q_hi) of the tile (multiples
// of 16), for NB B operands that share the A fragments.  a_lo / a_hi read gradient channels g and g + 8 (A_ROWS = 4:
// channels >= 4 are zero and a_hi is not read); b[i] reads activation channel n0_i + g.
// Every operand is split x = head + middle (+ a dropped 2^-16 tail) in bf16 and three products are accumulated:
// head*head, middle*head, head*middle.  The legacy mma.sync pipe, not instruction issue, bounds this phase on
// sm_100a (a TF32 m16n8k8 head/tail split, 6 mma per 16 points, measured 10.4 us per CTA; this form needs 3).
template <int A_ROWS, int NB, typename AFrag, typename BFrag>
__device__ __forceinline__ void wgrad_tiles(const AFrag& a_lo, const AFrag& a_hi, const BFrag (&b)[NB], int q_lo, int q_hi,
                                            int lane, float (&acc)[NB][4]) {
  const int g = lane >> 2;
  float x_acc[NB][4];  // the two cross terms (small) accumulate apart from the head product
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = x_acc[i][j] = 0.f;
#pragma unroll 2
  for (int q0 = q_lo; q0 < q_hi; q0 += 16) {
    uint32_t ah[4], am[4];
    if (g < A_ROWS) {
      split_bf16x2(a_lo.at(q0, 0), a_lo.at(q0, 1), ah[0], am[0]);
      split_bf16x2(a_lo.at(q0, 2), a_lo.at(q0, 3), ah[2], am[2]);
    } else {
      ah[0] = am[0] = ah[2] = am[2] = 0u;
    }
    if (A_ROWS > 8) {
      split_bf16x2(a_hi.at(q0, 0), a_hi.at(q0, 1), ah[1], am[1]);
      split_bf16x2(a_hi.at(q0, 2), a_hi.at(q0, 3), ah[3], am[3]);
    } else {
      ah[1] = am[1] = ah[3] = am[3] = 0u;
    }
#pragma unroll
    for (int i = 0; i < NB; ++i) {
      uint32_t bh[2], bm[2];
      split_bf16x2(b[i].at(q0, 0), b[i].at(q0, 1), bh[0], bm[0]);
      split_bf16x2(b[i].at(q0, 2), b[i].at(q0, 3), bh[1], bm[1]);
      mma_bf16(x_acc[i], am, bh);
      mma_bf16(x_acc[i], ah, bm);
      mma_bf16(acc[i], ah, bh);
    }
  }
#pragma unroll
  for (int i = 0; i < NB; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] += x_acc[i][j];
}

// the bias gradients: A^T 1 (column 0 of a tile whose B is one column of ones; 1.0 is exact in bf16)
template <int A_ROWS, typename AFrag>
__device__ __forceinline__ void wgrad_bias(const AFrag& a_lo, const AFrag& a_hi, int lane, float (&acc)[4]) {
  const int g = lane >> 2;
  float x_acc[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] = x_acc[j] = 0.f;
  uint32_t bh[2];
  bh[0] = bh[1] = (g == 0) ? 0x3f803f80u : 0u;
#pragma unroll 2
  for (int q0 = 0; q0 < NP; q0 += 16) {
    uint32_t ah[4], am[4];
    if (g < A_ROWS) {
      split_bf16x2(a_lo.at(q0, 0), a_lo.at(q0, 1), ah[0], am[0]);
      split_bf16x2(a_lo.at(q0, 2), a_lo.at(q0, 3), ah[2], am[2]);
    } else {
      ah[0] = am[0] = ah[2] = am[2] = 0u;
    }
    if (A_ROWS > 8) {
      split_bf16x2(a_hi.at(q0, 0), a_hi.at(q0, 1), ah[1], am[1]);
      split_bf16x2(a_hi.at(q0, 2), a_hi.at(q0, 3), ah[3], am[3]);
    } else {
      ah[1] = am[1] = ah[3] = am[3] = 0u;
    }
    mma_bf16(x_acc, am, bh);
    mma_bf16(acc, ah, bh);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) acc[j] += x_acc[j];
}

// fragment -> gradient arena: dst[m * ld + n0 + n] for the fragment's (m, n); rows m >= m_valid are dropped.
// Every CTA of the grid adds into the same 2700 floats, and same-line reductions serialise in L2, so lane pairs
// (t, t^1) swap halves and each lane issues one 16-byte red.global.add.v4 (even t: row g, odd t: row g+8, four
// consecutive columns) instead of four scalar atomics.
__device__ __forceinline__ void wgrad_store(float* dst, int ld, int n0, int m_valid, int lane, const float (&acc)[4]) {
  const int g = lane >> 2, t = lane & 3;
  const bool even = (t & 1) == 0;
  const float r0 = __shfl_xor_sync(0xffffffffu, even ? acc[2] : acc[0], 1);
  const float r1 = __shfl_xor_sync(0xffffffffu, even ? acc[3] : acc[1], 1);
  const int row = even ? g : g + 8;
  const int col = n0 + 2 * (t & 2);
  const float4 v = even ? make_float4(acc[0], acc[1], r0, r1) : make_float4(r0, r1, acc[2], acc[3]);
  if (row < m_valid) red_add_v4(reinterpret_cast<float4*>(dst + row * ld + col), v);
}
__device__ __forceinline__ void wgrad_store_bias(float* dst, int m_valid, int lane, const float (&acc)[4]) {
  const int g = lane >> 2, t = lane & 3;
  if (t == 0) {  // column 0 of the ones tile
    if (g < m_valid) atomicAdd(dst + g, acc[0]);
    if (g + 8 < m_valid) atomicAdd(dst + g + 8, acc[2]);
  }
}

// barrier over the 128 threads that own one decoder (hardware barriers 1 and 2; 0 is __syncthreads): from the MLP
// backward to the end of the scatter a half only touches its own staging buffer and feature tile, so the two
// halves need not wait for each other and their reduction-bound and FMA-bound phases can overlap
__device__ __forceinline__ void half_sync(int half) {
  asm volatile("bar.sync %0, %1;" ::"r"(1 + half), "r"(NP) : "memory");
}

// Weight gradients of BOTH decoders, accumulated into the gradient arena, in two rounds per half.  Warps 0-3 work
// on the sdf decoder, warps 4-7 on the rgb decoder (the same split as the thread halves, so `half` is also the
// decoder a warp computes for).  Contains barriers over each half: call from uniform control flow.
//   round 1  input layer  dW1 = ga1^T F : ga1 staged in act0 (sdf) / act1 (rgb); warp w of a half owns feature
//            columns 16w..16w+15 (two tiles sharing the A fragments); warp 0 of a half also takes db1
//   round 2  hidden and output layers: every owner thread parks (ga2 | h1 | h2 | gout) in ITS OWN row of the feature
//            tile, which is dead after round 1 and is only rewritten by mlp_backward_input afterwards; per half:
//            warps 0,1 dW2 (half of the points each, two tiles), warp 2 dW3 (two tiles), warp 3 db2 and db3
__device__ __forceinline__ void weight_grads(float* act0, float* act1, float4* F0, float4* F1, float* gdec, int half,
                                             int q, const float (&h1)[16], const float (&h2)[16], const float (&ga1)[16],
                                             const float (&ga2)[16], const float (&gout)[3]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wl = warp & 3;
  float* abuf = half ? act1 : act0;
  float4* F = half ? F1 : F0;
  const int nout = half ? 3 : 1;
  float* gW1 = gdec + (half ? C_W1 : S_W1);
  float* gB1 = gdec + (half ? C_B1 : S_B1);
  float* gW2 = gdec + (half ? C_W2 : S_W2);
  float* gB2 = gdec + (half ? C_B2 : S_B2);
  float* gW3 = gdec + (half ? C_W3 : S_W3);
  float* gB3 = gdec + (half ? C_B3 : S_B3);
  // ---- round 1
  {
    float4* a0 = reinterpret_cast<float4*>(abuf + q * WG_STRIDE);
#pragma unroll
    for (int v = 0; v < 4; ++v) a0[v] = make_float4(ga1[v * 4], ga1[v * 4 + 1], ga1[v * 4 + 2], ga1[v * 4 + 3]);
  }
  half_sync(half);
  const int g = lane >> 2, t = lane & 3;
  {
    float acc[2][4];
    const FromBuf a_lo(abuf, g, t), a_hi(abuf, g + 8, t);
    const FromTile b[2] = {FromTile(F, wl * 16 + g, t), FromTile(F, wl * 16 + 8 + g, t)};
    wgrad_tiles<16, 2>(a_lo, a_hi, b, 0, NP, lane, acc);
    wgrad_store(gW1, 64, wl * 16, 16, lane, acc[0]);
    wgrad_store(gW1, 64, wl * 16 + 8, 16, lane, acc[1]);
    if (wl == 0) {
      wgrad_bias<16>(a_lo, a_hi, lane, acc[0]);
      wgrad_store_bias(gB1, 16, lane, acc[0]);
    }
  }
  half_sync(half);
  // ---- round 2: columns 0-15 ga2, 16-31 h1, 32-47 h2, 48-51 gout of the thread's own feature-tile row
#pragma unroll
  for (int v = 0; v < 4; ++v) {
    F[f_slot(q, v)] = make_float4(ga2[v * 4], ga2[v * 4 + 1], ga2[v * 4 + 2], ga2[v * 4 + 3]);
    F[f_slot(q, 4 + v)] = make_float4(h1[v * 4], h1[v * 4 + 1], h1[v * 4 + 2], h1[v * 4 + 3]);
    F[f_slot(q, 8 + v)] = make_float4(h2[v * 4], h2[v * 4 + 1], h2[v * 4 + 2], h2[v * 4 + 3]);
  }
  F[f_slot(q, 12)] = make_float4(gout[0], gout[1], gout[2], 0.f);
  half_sync(half);
  const FromTile g2_lo(F, g, t), g2_hi(F, 8 + g, t);  // ga2 channels g, g + 8
  const FromTile g3_lo(F, 48 + (g & 3), t);           // gout channel g (< 4; lanes with g >= 4 contribute zeros)
  if (wl < 2) {
    float acc[2][4];
    const FromTile b[2] = {FromTile(F, 16 + g, t), FromTile(F, 24 + g, t)};
    wgrad_tiles<16, 2>(g2_lo, g2_hi, b, wl * (NP / 2), (wl + 1) * (NP / 2), lane, acc);
    wgrad_store(gW2, 16, 0, 16, lane, acc[0]);
    wgrad_store(gW2, 16, 8, 16, lane, acc[1]);
  } else if (wl == 2) {
    float acc[2][4];
    const FromTile b[2] = {FromTile(F, 32 + g, t), FromTile(F, 40 + g, t)};
    wgrad_tiles<4, 2>(g3_lo, g3_lo, b, 0, NP, lane, acc);
    wgrad_store(gW3, 16, 0, nout, lane, acc[0]);
    wgrad_store(gW3, 16, 8, nout, lane, acc[1]);
  } else {
    float acc[4];
    wgrad_bias<16>(g2_lo, g2_hi, lane, acc);
    wgrad_store_bias(gB2, 16, lane, acc);
    wgrad_bias<4>(g3_lo, g3_lo, lane, acc);
    wgrad_store_bias(gB3, nout, lane, acc);
  }
  half_sync(half);
}

// gather layout: scatter d loss/d features of one decoder into the plane gradients and/or accumulate the
// gradient with respect to the normalised coordinates.
// Scatter of one decoder's feature gradients (tile F, overwritten with d loss / d features) for the 8 consecutive
// points [qb, qb+8) owned by this 8-lane group, plus the gradient with respect to the normalised coordinates.
//
// The plane-gradient reductions are the one phase of the iteration that sits on a hardware limit
// (red.global.add.v4.f32 peaks at ~6.2 TB/s on B200, tools/microbench/l2_gather_red.cu), so the bytes are what
// must shrink: consecutive samples of a ray stay ~3 samples in the same 24 cm coarse cell, therefore the coarse
// taps run tap-major over the group's consecutive points and keep the four corner contributions in registers
// until the cell changes (run-length merging).  Fine taps (6 cm / 3 cm cells) rarely repeat and go point-major
// with all 12 corner loads of the scale in flight.
template <bool GF, bool GR, int AXBASE>
__device__ __forceinline__ void scatter_group(const FieldK& fk, int field, const float4* __restrict__ arena4,
                                              float4* __restrict__ garena4, const ax_t (*ax_i)[NP],
                                              const float (*ax_f)[NP], const float4* F, int qb, int n_valid, int sub,
                                              float (*gp)[NP], int dbg) {
  const bool do_red = GF && !(dbg & 1);
  // ---- pass 1, coarse scale, reductions only: tap-major over the consecutive points, merged per cell
  if (do_red) {
#pragma unroll
    for (int p = 0; p < 3; ++p) {
      const PlaneK& pl = fk.pl[field * 6 + p];
      const int au = AXBASE + pair_u(p), av = AXBASE + pair_v(p);
      int cur = -1, cdx = 0, cdy = 0;
      float4 a00 = f4_zero(), a01 = f4_zero(), a10 = f4_zero(), a11 = f4_zero();
#pragma unroll 1
      for (int it = 0; it < 8; ++it) {
        const int q = qb + it;
        if (q < n_valid) {
          const Tap t = make_tap(pl, ax_i[au][q], ax_f[au][q], ax_i[av][q], ax_f[av][q], sub);
          const float fu = t.fu, fv = t.fv;
          const float4 g4 = F[f_slot(q, sub)];
          const float w00 = (1.f - fu) * (1.f - fv), w01 = fu * (1.f - fv), w10 = (1.f - fu) * fv, w11 = fu * fv;
          if (t.base != cur) {
            if (cur >= 0) {
              red_add_v4(garena4 + cur, a00);
              red_add_v4(garena4 + cur + cdx, a01);
              red_add_v4(garena4 + cur + cdy, a10);
              red_add_v4(garena4 + cur + cdy + cdx, a11);
            }
            cur = t.base;
            cdx = t.dx;
            cdy = t.dy;
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
        red_add_v4(garena4 + cur, a00);
        red_add_v4(garena4 + cur + cdx, a01);
        red_add_v4(garena4 + cur + cdy, a10);
        red_add_v4(garena4 + cur + cdy + cdx, a11);
      }
    }
  }
  // ---- pass 2, point-major: coordinate gradients of both scales (12 corner loads in flight per scale) and
  //      the fine-scale reductions
#pragma unroll 1
  for (int it = 0; it < 8; ++it) {
    const int q = qb + it;
    float gpn[3] = {0.f, 0.f, 0.f};
    if (q < n_valid) {
#pragma unroll
      for (int sc = 0; sc < 2; ++sc) {
        if (!GR && sc == 0) continue;
        const float4 g4 = F[f_slot(q, sc * 8 + sub)];
        Tap tp[3];
        int u0[3], v0[3];
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const int au = AXBASE + sc * 3 + pair_u(p), av = AXBASE + sc * 3 + pair_v(p);
          u0[p] = ax_i[au][q];
          v0[p] = ax_i[av][q];
          tp[p] = make_tap(fk.pl[field * 6 + sc * 3 + p], u0[p], ax_f[au][q], v0[p], ax_f[av][q], sub);
        }
        if (GR) {
          float4 v[3][4];
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            v[p][0] = ldg4(arena4 + tp[p].base);
            v[p][1] = ldg4(arena4 + tp[p].base + tp[p].dx);
            v[p][2] = ldg4(arena4 + tp[p].base + tp[p].dy);
            v[p][3] = ldg4(arena4 + tp[p].base + tp[p].dy + tp[p].dx);
          }
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            const PlaneK& pl = fk.pl[field * 6 + sc * 3 + p];
            const float fu = tp[p].fu, fv = tp[p].fv;
            const float d00 = f4_dot(g4, v[p][0]), d01 = f4_dot(g4, v[p][1]);
            const float d10 = f4_dot(g4, v[p][2]), d11 = f4_dot(g4, v[p][3]);
            const float du = (d01 - d00) * (1.f - fv) + (d11 - d10) * fv;
            const float dv = (d10 - d00) * (1.f - fu) + (d11 - d01) * fu;
            gpn[pair_u(p)] = fmaf(du, axis_grad_mult(u0[p], fu, pl.W), gpn[pair_u(p)]);
            gpn[pair_v(p)] = fmaf(dv, axis_grad_mult(v0[p], fv, pl.H), gpn[pair_v(p)]);
          }
        }
        if (do_red && sc == 1) {
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            const float fu = tp[p].fu, fv = tp[p].fv;
            red_add_v4(garena4 + tp[p].base, f4_mul((1.f - fu) * (1.f - fv), g4));
            red_add_v4(garena4 + tp[p].base + tp[p].dx, f4_mul(fu * (1.f - fv), g4));
            red_add_v4(garena4 + tp[p].base + tp[p].dy, f4_mul((1.f - fu) * fv, g4));
            red_add_v4(garena4 + tp[p].base + tp[p].dy + tp[p].dx, f4_mul(fu * fv, g4));
          }
        }
      }
    }
    if (GR) {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float v = gpn[c];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        if (sub == 0) gp[c][q] = v;
      }
    }
  }
}

// MODE 0: upstream gradients of (depth, rgb, sdf) per ray are read (autograd through render_batch_ray).
// MODE 1 (FUSED): the five losses are evaluated in-kernel from gt data and device counters.
// MODE 2 (POINTS): S == 1, "rays" are plain points (rays_o = points, rays_d unused) and the upstream gradient
//         is g_raw[N][4] in g_sdf (autograd through Decoders.forward); g_rays_o receives d loss / d points.
// GF: gradients for planes + decoders (+beta) into grad_arena.  GR: gradients for rays / points / poses.
//
// One CTA = NP points (whole rays) and NT_BWD = 2*NP threads: threads [0,NP) own the sdf decoder of point
// q = tid, threads [NP,2NP) the rgb decoder of point q = tid-NP, so both decoders' gathers, MLPs and scatters
// run side by side and every thread carries one decoder's activations only.
// This is the PARAMETER form (64-channel features, plane gradients into the gradient arena): autograd through
// render_batch_ray / Decoders.forward and the cross-check of the Q form.  Both loops' iterations use qbwd.cuh.
template <int MODE, bool GF, bool GR>
__device__ __forceinline__ void render_bwd_body(const BwdArgs& a) {
  constexpr bool FUSED = MODE == 1;
  constexpr bool POINTS = MODE == 2;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  SmemBwd<GF>& sm = *reinterpret_cast<SmemBwd<GF>*>(smem_raw);
  const int R = a.counters ? min(a.counters[0], a.n_rays) : a.n_rays;
  const int S = a.S;
  const int RPB = POINTS ? NP : min(NP / S, 16);
  const int ray0 = blockIdx.x * RPB;
  if (ray0 >= R) return;
  const int rays_here = min(RPB, R - ray0);
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
  // ---- P0/P1: points, normalised coordinates, axis set-ups of this half's two resolution groups
  float pn[3] = {0.f, 0.f, 0.f}, zk = 0.f;
  if (valid) {
    zk = POINTS ? 0.f : a.z[(long long)ray * S + k];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float p = POINTS ? a.rays_o[(long long)ray * 3 + c]
                             : __fadd_rn(a.rays_o[ray * 3 + c], __fmul_rn(a.rays_d[ray * 3 + c], zk));
      pn[c] = normalize_axis(p, a.fk.lo[c], a.fk.hi[c]);
    }
  }
  write_axis_setups<2>(a.fk, 2 * half, pn, sm.ax_i + 6 * half, sm.ax_f + 6 * half, q);
  load_decoder_weights(sm.W, reinterpret_cast<const float*>(a.arena4) + a.fk.dec_off, tid, NT_BWD);
  __syncthreads();
  PHASE_MARK(0);
  // ---- P2/P3 from the cache: the tracker has just run k_render_fwd on these rays (it needs the rendered depth
  //      for its outlier mask before any gradient), which left sdf, rgb and the ReLU masks of every sample.
  //      Without weight gradients that is all the backward pass needs, so the gather and both forward MLPs go.
  const bool cached = !GF && FUSED && a.act4 != nullptr;
  float4* Fh = half ? sm.F1 : sm.F0;
  float h1[16], h2[16], out[3] = {0.f, 0.f, 0.f};
  float sdf = 0.f, u = 0.f, e = 0.f, alpha = 0.f, one = 1.f, rgb[3] = {0.f, 0.f, 0.f};
  const float* Wh = sm.W + half * DW_STRIDE;
  float beta;
  if (cached) {
    PHASE_MARK(1);
    decoder_weights_wait();
    __syncthreads();
    PHASE_MARK(2);
    beta = sm.W[DW_BETA];
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
    if (half == 0) {
      sdf_to_alpha(sdf, beta, u, e, alpha);
      one = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
      sm.one[q] = one;
      sm.z[q] = zk;
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) sm.c[c][q] = rgb[c];
    }
  } else {
  // ---- P2: gather this half's decoder features
  if (half == 0)
    gather_tile<0>(a.fk, 0, a.arena4, sm.ax_i, sm.ax_f, sm.F0, n_valid, q);
  else
    gather_tile<6>(a.fk, 1, a.arena4, sm.ax_i, sm.ax_f, sm.F1, n_valid, q);
  PHASE_MARK(1);
  decoder_weights_wait();
  __syncthreads();
  PHASE_MARK(2);
  // ---- P3: MLP forward of this half's decoder
  beta = sm.W[DW_BETA];
  if (a.dbg & 4) {  // profiling: skip the MLP arithmetic (results are meaningless)
#pragma unroll
    for (int j = 0; j < 16; ++j) h1[j] = h2[j] = sm.F0[q * 16].x * 0.f + (float)j;
    if (half == 0) {
      sdf = 0.1f;
      sdf_to_alpha(sdf, beta, u, e, alpha);
      one = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
      sm.one[q] = one;
      sm.z[q] = zk;
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        rgb[c] = 0.5f;
        sm.c[c][q] = rgb[c];
      }
    }
  } else {
    mlp_forward_s(Wh, Fh, q, h1, h2, out);
    if (half == 0) {
      sdf = tanhf(out[0]);
      sdf_to_alpha(sdf, beta, u, e, alpha);
      one = __fadd_rn(__fsub_rn(1.0f, alpha), 1e-10f);
      sm.one[q] = one;
      sm.z[q] = zk;
    } else {
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        rgb[c] = sigmoidf_(out[c]);
        sm.c[c][q] = rgb[c];
      }
    }
  }
  }
  PHASE_MARK(3);
  __syncthreads();
  PHASE_MARK(4);
  // ---- P4: compositing (sdf half)
  float T = 1.0f, w = 0.f;
  if (!POINTS) {
    if (half == 0) {
      for (int j = 0; j < k; ++j) T *= sm.one[rl * S + j];
      w = valid ? alpha * T : 0.f;
      sm.w[q] = w;
    }
    __syncthreads();
    if (half == 0 && valid && k < 4) {
      const float* v = (k == 0) ? sm.z : sm.c[k - 1];
      float acc = 0.f;
      for (int j = 0; j < S; ++j) acc = fmaf(sm.w[rl * S + j], v[rl * S + j], acc);
      sm.rayv[k][rl] = acc;
    }
    __syncthreads();
  }
  // ---- P5: upstream gradients of depth / rgb per ray, and the loss sums
  double ls[5] = {0.0, 0.0, 0.0, 0.0, 0.0};  // fs, center, tail, depth, colour
  float inv_f = 0.f, inv_c = 0.f, inv_t = 0.f;
  if (FUSED) {
    const int n_mask = a.norm[2];
    inv_f = a.w_fs / (float)a.norm[3];
    inv_c = a.w_center / (float)a.norm[4];
    inv_t = a.w_tail / (float)a.norm[5];
    if (half == 0 && valid && k == 0) {
      const float d = a.gt_depth[ray];
      const int m = a.ray_mask ? (int)a.ray_mask[ray] : (d > 0.f ? 1 : 0);
      const float dr = sm.rayv[0][rl];
      float gd = 0.f;
      if (m) {
        const float diff = d - dr;
        gd = -2.0f * diff * (a.w_depth / (float)n_mask);
        ls[3] = (double)(diff * diff);
      }
      sm.rayg[0][rl] = gd;
      const bool col = a.ray_mask ? (m != 0) : true;
      const double ncol = 3.0 * (double)(a.ray_mask ? n_mask : a.norm[0]);
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        float g = 0.f;
        if (col) {
          const double diff = a.gt_color[(long long)ray * 3 + c] - (double)sm.rayv[1 + c][rl];
          g = (float)(-2.0 * diff * (a.w_color / ncol));
          ls[4] += diff * diff;
        }
        sm.rayg[1 + c][rl] = g;
      }
      sm.rayd[rl] = d;
      sm.raym[rl] = m;
    }
  } else if (!POINTS) {
    if (half == 0 && valid && k < 4) sm.rayg[k][rl] = (k == 0) ? a.g_depth[ray] : a.g_rgb[ray * 3 + (k - 1)];
  }
  if (!POINTS) __syncthreads();
  // ---- compositing backward -> gradient at this half's decoder outputs
  float g_beta = 0.f, gout[3] = {0.f, 0.f, 0.f};
  if (POINTS) {
    if (valid) {
      const float* gr = a.g_sdf + (long long)ray * 4;
      if (half == 0) {
        gout[0] = gr[3] * (1.0f - sdf * sdf);
      } else {
#pragma unroll
        for (int c = 0; c < 3; ++c) gout[c] = gr[c] * rgb[c] * (1.0f - rgb[c]);
      }
    }
  } else {
    float gw = 0.f;
    if (half == 0) {
      gw = sm.rayg[0][rl] * zk + sm.rayg[1][rl] * sm.c[0][q] + sm.rayg[2][rl] * sm.c[1][q] + sm.rayg[3][rl] * sm.c[2][q];
      sm.gww[q] = valid ? gw * w : 0.f;
    } else if (valid) {
      const float ww = sm.w[q];
#pragma unroll
      for (int c = 0; c < 3; ++c) gout[c] = sm.rayg[1 + c][rl] * ww * rgb[c] * (1.0f - rgb[c]);
    }
    __syncthreads();
    if (half == 0 && valid) {
      float gdir = 0.f;
      if (FUSED) {
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
            ls[band] = (double)(r * r);
          }
        }
      } else {
        gdir = a.g_sdf ? a.g_sdf[(long long)ray * S + k] : 0.f;
      }
      float B = 0.f;
      for (int j = k + 1; j < S; ++j) B += sm.gww[rl * S + j];
      const float g_alpha = gw * T - B / one;
      const float du = u * (1.0f - u);
      const float g_sdf = gdir + g_alpha * (-beta * beta * e * du);
      g_beta = g_alpha * e * (u - beta * sdf * du);
      gout[0] = g_sdf * (1.0f - sdf * sdf);
    }
  }
  PHASE_MARK(5);
  // ---- P6: MLP backward of this half's decoder
  float ga1[16], ga2[16];
  if (a.dbg & 4) {
#pragma unroll
    for (int j = 0; j < 16; ++j) ga1[j] = ga2[j] = gout[0] + gout[1] * (float)j;
  } else {
    mlp_backward_hidden_s(Wh, gout, h1, h2, ga1, ga2);  // sdf: gout[1] = gout[2] = 0
  }
  PHASE_MARK(6);
  if (GF && !(a.dbg & 2)) {
    float* gdec = a.grad_arena + a.fk.dec_off;
    weight_grads(sm.act0, sm.act1, sm.F0, sm.F1, gdec, half, q, h1, h2, ga1, ga2, gout);
    const float gb = warp_sum(g_beta);
    if (lane == 0) sm.red[warp] = gb;
  }
  PHASE_MARK(7);
  if (a.dbg & 4) {
    Fh[q * 16] = make_float4(ga1[0], ga1[1], ga1[2], ga1[3]);
  } else {
    mlp_backward_input_s(Wh, ga1, Fh, q);
  }
  half_sync(half);
  if (GF && !(a.dbg & 2) && tid == 0) {
    float gb = 0.f;
    for (int i = 0; i < NP / 32; ++i) gb += sm.red[i];  // only the sdf half carries beta gradients
    atomicAdd(a.grad_arena + a.fk.dec_off + P_BETA, gb);
  }
  PHASE_MARK(8);
  // ---- P7: scatter to the planes / coordinate gradients (gather layout, each half its own decoder)
  {
    const int wl = (tid & (NP - 1)) >> 5, grp = lane >> 3, sub = lane & 7;
    const int qb = wl * 32 + grp * 8;  // 8 consecutive points (samples along a ray) per 8-lane group
    float4* garena4 = reinterpret_cast<float4*>(a.grad_arena);
    if (half == 0)
      scatter_group<GF, GR, 0>(a.fk, 0, a.arena4, garena4, sm.ax_i, sm.ax_f, sm.F0, qb, n_valid, sub, sm.gp[0], a.dbg);
    else
      scatter_group<GF, GR, 6>(a.fk, 1, a.arena4, garena4, sm.ax_i, sm.ax_f, sm.F1, qb, n_valid, sub, sm.gp[1], a.dbg);
  }
  PHASE_MARK(9);
  // ---- P8: ray / point / pose gradients
  if (GR) {
    __syncthreads();
    for (int t = tid; t < rays_here * 6; t += NT_BWD) {
      const int r2 = t / 6, comp = t - r2 * 6, ax = comp % 3;
      const bool is_d = comp >= 3;
      const float scale = 2.0f / (a.fk.hi[ax] - a.fk.lo[ax]);
      float acc = 0.f;
      for (int j = 0; j < S; ++j) {
        const float g = (sm.gp[0][ax][r2 * S + j] + sm.gp[1][ax][r2 * S + j]) * scale;
        acc += is_d ? g * sm.z[r2 * S + j] : g;
      }
      if (FUSED) {
        sm.rayod[comp][r2] = acc;
      } else if (!POINTS || !is_d) {
        (is_d ? a.g_rays_d : a.g_rays_o)[(long long)(ray0 + r2) * 3 + ax] = acc;
      }
    }
    if (FUSED && a.pose_grad) {
      __syncthreads();
      for (int t = tid; t < rays_here * 12; t += NT_BWD) {
        const int r2 = t / 12, el = t - r2 * 12, row = el >> 2, col = el & 3;
        const int slot = a.src[ray0 + r2];
        const int frame = slot / a.n_per_img;
        float val;
        if (col == 3) {
          val = sm.rayod[row][r2];  // d loss / d t
        } else {
          const long long pix = a.pix_idx[slot];
          const float pi = (float)(a.W0 + (int)(pix % a.Wc)), pj = (float)(a.H0 + (int)(pix / a.Wc));
          const float dir = col == 0 ? __fdiv_rn(__fsub_rn(pi, a.cx), a.fx)
                                     : (col == 1 ? -__fdiv_rn(__fsub_rn(pj, a.cy), a.fy) : -1.0f);
          val = sm.rayod[3 + row][r2] * dir;  // d loss / d R[row][col]
        }
        atomicAdd(a.pose_grad + frame * 12 + el, val);
      }
    }
  }
  PHASE_MARK(10);
  // ---- loss sums
  if (FUSED && a.loss_acc) {
#pragma unroll
    for (int i = 0; i < 5; ++i) {
      const double v = warp_sum_d(ls[i]);
      if (lane == 0) sm.redd[warp * 5 + i] = v;
    }
    __syncthreads();
    if (tid < 5) {
      double v = 0.0;
      for (int i = 0; i < NP / 32; ++i) v += sm.redd[i * 5 + tid];  // only the sdf half accumulates losses
      atomicAdd(a.loss_acc + tid, v);
    }
  }
  (void)Fh;
}

template <int MODE, bool GF, bool GR>
__global__ void __launch_bounds__(NT_BWD, 2) k_render_bwd(const __grid_constant__ BwdArgs a) {
  render_bwd_body<MODE, GF, GR>(a);
}

}  // namespace eslam
