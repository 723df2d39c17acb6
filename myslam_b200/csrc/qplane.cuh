// The Q form: pre-activated planes.  The default path of both loops; tests/test_gpu_qform.py
// holds it to the parameter form (forward 1e-5, gradients 1e-3) and tests/test_gpu_default_path.py to the oracle.
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

__global__ void __launch_bounds__(QB_TPC) k_q_build(const __grid_constant__ QBuildArgs a) {
  // rows 8..15 (the second half's outputs) start 4 floats later, so the two addresses of a warp's weight load (one per
  // half) fall into different banks
  __shared__ __align__(16) float sW[16 * 32 + 4];
  __shared__ float4 sT[QB_TPC * 8];
  const int g = q_group_of(a.qg, blockIdx.x);
  const int tl0 = (blockIdx.x - a.qg.unit0[g]) * QB_TPC;
  const int cnt = min(QB_TPC, a.qg.n[g] - tl0);
  const long long tbase = (long long)a.qg.t0[g] + tl0;
  const float4* src = a.arena4 + tbase * 8;
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int i = u * QB_TPC + threadIdx.x, t = i >> 3;
    sT[t_slot(t, i & 7)] = t < cnt ? ldg4(src + i) : f4_zero();
  }
  const float* w1 = a.dec + ((g >> 1) ? C_W1 : S_W1) + (g & 1) * 32;  // columns of this scale (coarse | fine, decoders.py:84)
  for (int i = threadIdx.x; i < 16 * 32; i += QB_TPC) sW[i + ((i >> 8) << 2)] = w1[(i >> 5) * 64 + (i & 31)];
  __syncthreads();
  // thread (texel pair, half): 8 of the 16 outputs of texels tp and tp + 64; each weight load feeds both texels
  const int tp = threadIdx.x >> 1, half = threadIdx.x & 1;
  float o[2][8];
#pragma unroll
  for (int j = 0; j < 8; ++j) o[0][j] = o[1][j] = 0.f;
  const float* w = sW + half * (8 * 32 + 4);
#pragma unroll
  for (int c4 = 0; c4 < 8; ++c4) {
    const float4 x0 = sT[t_slot(tp, c4)], x1 = sT[t_slot(tp + 64, c4)];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float4 wj = lds4(w + j * 32 + c4 * 4);
      o[0][j] = fmaf(wj.w, x0.w, fmaf(wj.z, x0.z, fmaf(wj.y, x0.y, fmaf(wj.x, x0.x, o[0][j]))));
      o[1][j] = fmaf(wj.w, x1.w, fmaf(wj.z, x1.z, fmaf(wj.y, x1.y, fmaf(wj.x, x1.x, o[1][j]))));
    }
  }
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    const int t = tp + 64 * k;
    if (t >= cnt) continue;
    float4* dst = reinterpret_cast<float4*>(a.q2) + (tbase + t) * 4 + half * 2;
    dst[0] = make_float4(o[k][0], o[k][1], o[k][2], o[k][3]);
    dst[1] = make_float4(o[k][4], o[k][5], o[k][6], o[k][7]);
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
// so it is fused with the plane half of Adam: the plane gradient lives in registers only.
// One CTA = one tile of QA_TILE consecutive texels of one group, three phases, every global access coalesced:
//   A  the tile's gradient rows (16 KB... 8 KB), its texels (16 KB) and `touched` bytes go to shared memory, all loads
//      in flight at once; on several GPUs the rows are the sum of the local image and the peers' staged slices
//   B  thread (texel, c4): plane gradient of its four channels (64 FMA against W1 in shared memory), Adam, stores --
//      the moments were requested in phase A already -- only for ACTIVE texels (non-zero row now, or non-zero moments from an earlier iteration: exact skip as in
//      k_adam); on several GPUs the new texel goes to every rank's arena (P2P stores)
//   C  dW1 of the tile = G^T P, a 16 x 32 x 128 contraction on the tensor cores (split-bf16 mma, render.cuh), one
//      16-byte reduction per lane into the gradient arena's decoder block
// (An 8-lanes-per-texel form with dW1 in registers ran 40 us, a thread-per-texel form 34 us: after the first
// iterations of a call nearly every 32-texel run holds an active texel, so per-texel divergence buys nothing, and
// the tail is bound by issue slots -- hence the approximate reciprocal / square root in phase B, 2 ulp, where the
// parameter-form k_adam keeps torch's exact operation order.)
struct QAdamArgs {
  QGroups qg;      // units = tiles of QA_TILE texels
  float4* arena4;  // parameters; the planes are updated in place
  float4* gq4;     // gradient images, zeroed where consumed
  float4 *m4, *v4; // Adam moments in parameter-arena layout
  float* gdec;     // gradient arena's decoder block: dW1 is added here
  const float* dec;
  unsigned char* touched;  // one flag per texel, in texel order
  float step_sdf, step_rgb;  // lr / (1 - beta1^t) of the sdf / rgb planes
  float inv_bc2;
  AdamArgs adam;             // scalars only
  // several GPUs (world > 1): this rank owns the units [unit_lo, unit_lo + gridDim.x)
  int rank, world, unit_lo;
  long long lo4, smax;              // first float4 of the owned slice of the gradient images; staging row pitch
  float4* stage;                    // local staging rows [world][smax]: the peers' slices of this rank's texels
  float4* peer_p[ESLAM_MAX_PEERS];  // every rank's parameter arena
  float4* mc_p;                     // multicast alias of the parameter arenas or NULL
};

constexpr int QA_TILE = 128, QA_THREADS = 256;

struct SmemQAdam {
  float4 sG[QA_TILE * 4];  // gradient rows, p_slot layout
  float4 sP[QA_TILE * 8];  // texels before the update, t_slot layout
  float sW[16 * 32];
  unsigned char nz[QA_TILE], act[QA_TILE];
};

struct FromT {  // t_slot tile: rows = texels, c = channel (q0 is a multiple of 16, so the swizzle term depends on r only)
  const float* base;
  int o[4];
  __device__ __forceinline__ FromT(const float4* T, int c, int t) : base(reinterpret_cast<const float*>(T)) {
    const int r[4] = {2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9};
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = (r[i] * 8 + ((c >> 2) ^ (r[i] & 7))) * 4 + (c & 3);
  }
  __device__ __forceinline__ float at(int q0, int i) const { return base[q0 * 32 + o[i]]; }
};
struct FromG {  // p_slot tile: rows = texels, c = channel
  const float* base;
  int o[4];
  __device__ __forceinline__ FromG(const float4* P, int c, int t) : base(reinterpret_cast<const float*>(P)) {
    const int r[4] = {2 * t, 2 * t + 1, 2 * t + 8, 2 * t + 9};
#pragma unroll
    for (int i = 0; i < 4; ++i) o[i] = (r[i] * 4 + ((c >> 2) ^ ((r[i] >> 1) & 3))) * 4 + (c & 3);
  }
  __device__ __forceinline__ float at(int q0, int i) const { return base[q0 * 16 + o[i]]; }
};

__device__ __forceinline__ float4 ld_sys_f4(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p)
               : "memory");
  return v;
}
__device__ __forceinline__ bool f4_nonzero(float4 v) { return v.x != 0.f || v.y != 0.f || v.z != 0.f || v.w != 0.f; }

// torch.optim.Adam's update with an approximate square root and reciprocal (2 ulp each)
__device__ __forceinline__ void adam_fast(float& p, float g, float& m, float& v, const QAdamArgs& a, float ss) {
  m = fmaf(a.adam.one_m_beta1, g - m, m);
  v = fmaf(a.adam.one_m_beta2 * g, g, a.adam.beta2 * v);
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(v));
  p = fmaf(-ss * m, __fdividef(1.0f, fmaf(r, a.inv_bc2, a.adam.eps)), p);
}

// WMAX: compile-time bound of the world size (1 = one GPU)
template <int WMAX>
__device__ __forceinline__ void q_adam_tile(const QAdamArgs& a, int unit, SmemQAdam& sm) {
  const int g = q_group_of(a.qg, unit);
  const int tl0 = (unit - a.qg.unit0[g]) * QA_TILE;
  const int cnt = min(QA_TILE, a.qg.n[g] - tl0);
  const long long tbase = (long long)a.qg.t0[g] + tl0;
  const int tid = threadIdx.x, lane = tid & 31;
  const float4 z4 = f4_zero();
  // ---- phase A.  The moments of the thread's four (texel, c4) items are requested here already, for the texels whose
  // `touched` flag is up: a texel that has never had a gradient has m = v = 0 and needs no load, so phase B's only
  // dependent round trip disappears behind the staging of the tile.
  float4 m[4], v[4];
  {
    bool fl[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int t = (tid >> 3) + k * 32;
      fl[k] = t < cnt && a.touched[tbase + t] != 0;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long long at = (tbase + (tid >> 3) + k * 32) * 8 + (tid & 7);
      m[k] = fl[k] ? a.m4[at] : z4;
      v[k] = fl[k] ? a.v4[at] : z4;
    }
  }
#pragma unroll
  for (int k = 0; k < QA_TILE * 4 / QA_THREADS; ++k) {
    const int i = tid + k * QA_THREADS, t = i >> 2, c = i & 3;
    float4 v = z4;
    bool nz = false;
    if (t < cnt) {
      const long long idx = tbase * 4 + i;
      if (WMAX == 1) {
        v = a.gq4[idx];
        nz = f4_nonzero(v);
        if (nz) a.gq4[idx] = z4;
      } else {
        // fixed rank order: the sum does not depend on who owns the slice
#pragma unroll
        for (int q = 0; q < WMAX; ++q)
          if (q < a.world) {
            float4* src = q == a.rank ? a.gq4 + idx : a.stage + q * a.smax + (idx - a.lo4);
            const float4 part = q == a.rank ? *src : ld_sys_f4(src);
            if (f4_nonzero(part)) {  // consumed: the image and the staging rows go back to zero
              nz = true;
              *src = z4;
            }
            v = q == 0 ? part : f4_add(v, part);
          }
      }
    }
    sm.sG[p_slot(t, c)] = v;
    const unsigned b = __ballot_sync(0xffffffffu, nz);
    if (c == 0) sm.nz[t] = ((b >> (lane & ~3)) & 0xfu) != 0u;
  }
#pragma unroll
  for (int k = 0; k < QA_TILE * 8 / QA_THREADS; ++k) {
    const int i = tid + k * QA_THREADS, t = i >> 3;
    sm.sP[t_slot(t, i & 7)] = t < cnt ? a.arena4[tbase * 8 + i] : z4;
  }
  bool flag = false;
  if (tid < cnt) flag = a.touched[tbase + tid] != 0;
  const float* w1 = a.dec + ((g >> 1) ? C_W1 : S_W1) + (g & 1) * 32;
  for (int i = tid; i < 16 * 32; i += QA_THREADS) sm.sW[i] = w1[(i >> 5) * 64 + (i & 31)];
  __syncthreads();
  bool mine_nz = false, mine_act = false;
  if (tid < QA_TILE) {
    mine_nz = sm.nz[tid] != 0;
    mine_act = mine_nz || flag;
    sm.act[tid] = mine_act;
    if (mine_nz && !flag) a.touched[tbase + tid] = 1;
  }
  const bool any_act = __syncthreads_or(mine_act);
  const bool any_nz = __syncthreads_or(mine_nz);
  if (!any_act) return;
  // ---- phase B: (texel, c4) items; the moments of the thread's four items are requested together
  {
    const int c4 = tid & 7;
    const float ss = (g >> 1) ? a.step_rgb : a.step_sdf;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int t = (tid >> 3) + k * 32;
      if (!(t < cnt && sm.act[t])) continue;
      const long long at = (tbase + t) * 8 + c4;
      float4 p = sm.sP[t_slot(t, c4)];
      float4 gr = z4;
      if (sm.nz[t]) {
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 gv = sm.sG[p_slot(t, c)];
          gr = f4_fma(gv.x, lds4(sm.sW + (c * 4 + 0) * 32 + c4 * 4), gr);
          gr = f4_fma(gv.y, lds4(sm.sW + (c * 4 + 1) * 32 + c4 * 4), gr);
          gr = f4_fma(gv.z, lds4(sm.sW + (c * 4 + 2) * 32 + c4 * 4), gr);
          gr = f4_fma(gv.w, lds4(sm.sW + (c * 4 + 3) * 32 + c4 * 4), gr);
        }
      }
      adam_fast(p.x, gr.x, m[k].x, v[k].x, a, ss);
      adam_fast(p.y, gr.y, m[k].y, v[k].y, a, ss);
      adam_fast(p.z, gr.z, m[k].z, v[k].z, a, ss);
      adam_fast(p.w, gr.w, m[k].w, v[k].w, a, ss);
      a.m4[at] = m[k];
      a.v4[at] = v[k];
      if (WMAX == 1) {
        a.arena4[at] = p;
      } else if (a.mc_p) {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(a.mc_p + at), "f"(p.x), "f"(p.y),
                     "f"(p.z), "f"(p.w)
                     : "memory");
      } else {
#pragma unroll
        for (int q = 0; q < WMAX; ++q)
          if (q < a.world) a.peer_p[q][at] = p;
      }
    }
  }
  // ---- phase C: dW1 += G^T P over the tile; warp w: channels 8 (w & 3) .., texels 64 (w >> 2) ..
  if (any_nz) {
    const int w = tid >> 5, nt = w & 3, kh = w >> 2;
    const int gg = lane >> 2, tt = lane & 3;
    float acc[1][4];
    const FromG a_lo(sm.sG, gg, tt), a_hi(sm.sG, gg + 8, tt);
    const FromT b[1] = {FromT(sm.sP, nt * 8 + gg, tt)};
    wgrad_tiles<16, 1>(a_lo, a_hi, b, kh * (QA_TILE / 2), (kh + 1) * (QA_TILE / 2), lane, acc);
    float* dst = a.gdec + ((g >> 1) ? C_W1 : S_W1) + (g & 1) * 32;
    wgrad_store(dst, 64, nt * 8, 16, lane, acc[0]);
  }
}

__global__ void __launch_bounds__(QA_THREADS, 3) k_q_adam_planes(const __grid_constant__ QAdamArgs a) {
  __shared__ SmemQAdam sm;
  q_adam_tile<1>(a, blockIdx.x, sm);
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
  sm.w[q] = valid ? alpha : 0.f;
  if (valid && a.sdf) a.sdf[(long long)ray * S + k] = sdf;
  __syncthreads();
  // compositing (Renderer.py:140-147), one WARP per ray: lane l holds samples l and l + 32, the transmittance is a
  // prefix product by shuffles
  const int warp = q >> 5, lane = q & 31;
  for (int r2 = warp; r2 < rays_here; r2 += NP / 32) {
    const int j0 = r2 * S + lane, j1 = j0 + 32;
    const bool v0 = lane < S, v1 = lane + 32 < S;
    float inc0 = v0 ? sm.one[j0] : 1.f, inc1 = v1 ? sm.one[j1] : 1.f;
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
    float rv[4];
    rv[0] = warp_sum(fmaf(w0, v0 ? sm.z[j0] : 0.f, w1 * (v1 ? sm.z[j1] : 0.f)));
#pragma unroll
    for (int c = 0; c < 3; ++c) rv[1 + c] = warp_sum(fmaf(w0, v0 ? sm.c[c][j0] : 0.f, w1 * (v1 ? sm.c[c][j1] : 0.f)));
    if (lane == 0) a.depth[ray0 + r2] = rv[0];
    if (lane >= 1 && lane < 4) a.rgb[(ray0 + r2) * 3 + (lane - 1)] = lane == 1 ? rv[1] : (lane == 2 ? rv[2] : rv[3]);
  }
}

}  // namespace eslam
