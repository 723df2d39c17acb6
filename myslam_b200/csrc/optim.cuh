// Layout conversion, fused Adam over the parameter arena, pose chain rule + Adam, loss finalisation.
// Reference: torch/optim/adam.py (_single_tensor_adam) as used at src/Tracker.py:291-296 and
// src/Mapper.py:288-306,348-350; src/common.py:169-181 for the pose chain.
#pragma once
#include "field.cuh"

namespace eslam {

// ---- NCHW [32][H*W]  <->  channels-last [H*W][32] ---------------------------------------------------
template <bool IMPORT>
__global__ void __launch_bounds__(256) k_plane_layout(const float* __restrict__ src, float* __restrict__ dst, int HW) {
  __shared__ float tile[32][33];
  const int t0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  if (IMPORT) {
#pragma unroll
    for (int c = ty; c < 32; c += 8)
      if (t0 + tx < HW) tile[c][tx] = src[(long long)c * HW + t0 + tx];
    __syncthreads();
#pragma unroll
    for (int r = ty; r < 32; r += 8)
      if (t0 + r < HW) dst[(long long)(t0 + r) * 32 + tx] = tile[tx][r];
  } else {
#pragma unroll
    for (int r = ty; r < 32; r += 8)
      if (t0 + r < HW) tile[tx][r] = src[(long long)(t0 + r) * 32 + tx];
    __syncthreads();
#pragma unroll
    for (int c = ty; c < 32; c += 8)
      if (t0 + tx < HW) dst[(long long)c * HW + t0 + tx] = tile[c][tx];
  }
}

// ---- Adam ------------------------------------------------------------------------------------------
struct AdamArgs {
  unsigned char* touched;  // optional: one flag per 128 floats (4 texels), see k_adam
  float *p, *g, *m, *v;
  long long n;
  long long seg_end[4];
  float seg_step[4];  // lr / (1 - beta1^t)
  int n_seg;
  float beta1, beta2, one_m_beta1, one_m_beta2, bc2_sqrt, eps;
};

// torch.lerp(a, b, w): a + w*(b-a) if w < 0.5 else b - (b-a)*(1-w)   (ATen/native/Lerp.h)
__device__ __forceinline__ float lerp_torch(float a, float b, float w) {
  const float diff = __fsub_rn(b, a);
  return (w < 0.5f) ? __fadd_rn(a, __fmul_rn(w, diff)) : __fsub_rn(b, __fmul_rn(diff, __fsub_rn(1.0f, w)));
}

__device__ __forceinline__ void adam_one(float& p, float g, float& m, float& v, const AdamArgs& a, float step_size) {
  m = lerp_torch(m, g, a.one_m_beta1);
  v = __fadd_rn(__fmul_rn(v, a.beta2), __fmul_rn(__fmul_rn(a.one_m_beta2, g), g));
  const float denom = __fadd_rn(__fdiv_rn(__fsqrt_rn(v), a.bc2_sqrt), a.eps);
  p = __fadd_rn(p, __fdiv_rn(__fmul_rn(-step_size, m), denom));
}

// With `touched` (zeroed together with the moments when the optimiser is re-created, Mapper.py:291-299): a group of
// 128 parameters whose gradient has been zero in every step so far still has m = v = 0, so torch's update is
// p - lr * 0 / (0 + eps) = p exactly and its moments, parameters and (already zero) gradient need not be touched.
// A warp reads the group's gradient (512 contiguous bytes) and skips the other 7/8 of the traffic for it; the
// flag is raised the first time any of the 128 gradients is non-zero.  Scenes whose bound is much larger than the
// observed region (ScanNet: ~70 % of the texels are never hit) save most of the optimiser stream; the result is
// bit-identical to the dense update.
__global__ void __launch_bounds__(256) k_adam(const __grid_constant__ AdamArgs a) {
  const long long n4 = a.n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  float4* p4 = reinterpret_cast<float4*>(a.p);
  float4* g4 = reinterpret_cast<float4*>(a.g);
  float4* m4 = reinterpret_cast<float4*>(a.m);
  float4* v4 = reinterpret_cast<float4*>(a.v);
  if (a.touched) {
    // whole warps per trip (n4 rounded up to 32) so the ballot below is convergent
    const long long n4r = (n4 + 31) & ~31ll;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4r; i += stride) {
      const bool in = i < n4;
      const float4 g = in ? g4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      const bool nz = g.x != 0.f || g.y != 0.f || g.z != 0.f || g.w != 0.f;
      const long long grp = i >> 5;
      const bool any = __ballot_sync(0xffffffffu, nz) != 0u;
      const bool was = a.touched[grp] != 0;  // same byte for the whole warp
      if (!any && !was) continue;
      if (!was && (threadIdx.x & 31) == 0) a.touched[grp] = 1;
      if (!in) continue;
      const long long e = i << 2;
      int seg = 0;
      while (seg < a.n_seg - 1 && e >= a.seg_end[seg]) ++seg;
      const float ss = a.seg_step[seg];
      float4 p = p4[i], m = m4[i], v = v4[i], gg = g;
      adam_one(p.x, gg.x, m.x, v.x, a, ss);
      adam_one(p.y, gg.y, m.y, v.y, a, ss);
      adam_one(p.z, gg.z, m.z, v.z, a, ss);
      adam_one(p.w, gg.w, m.w, v.w, a, ss);
      p4[i] = p;
      m4[i] = m;
      v4[i] = v;
      if (any) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    }
    return;
  }
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const long long e = i << 2;
    int seg = 0;
    while (seg < a.n_seg - 1 && e >= a.seg_end[seg]) ++seg;  // segments are multiples of 4 floats
    const float ss = a.seg_step[seg];
    float4 p = p4[i], g = g4[i], m = m4[i], v = v4[i];
    adam_one(p.x, g.x, m.x, v.x, a, ss);
    adam_one(p.y, g.y, m.y, v.y, a, ss);
    adam_one(p.z, g.z, m.z, v.z, a, ss);
    adam_one(p.w, g.w, m.w, v.w, a, ss);
    p4[i] = p;
    m4[i] = m;
    v4[i] = v;
    g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

// ---- pose chain rule + Adam --------------------------------------------------------------------------
struct PoseAdamArgs {
  float *poses, *pose_grad, *m, *v;
  int n, first;
  float step_q, step_t;  // lr / (1 - beta1^t)
  float beta1, beta2, one_m_beta1, one_m_beta2, bc2_sqrt, eps;
  float* grad7;
  int apply;
};

__global__ void k_pose_adam(const __grid_constant__ PoseAdamArgs a) {
  const int f = a.first + blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= a.n) return;
  float* q = a.poses + f * 7;
  float* G = a.pose_grad + f * 12;  // d loss / d c2w[:3,:4], row-major 3x4
  const float r = q[0], i = q[1], j = q[2], k = q[3];
  const float n = r * r + i * i + j * j + k * k;
  const float s = 2.0f / n;
  // R = I + s*A(q)
  const float A[9] = {-(j * j + k * k), i * j - k * r, i * k + j * r, i * j + k * r, -(i * i + k * k),
                      j * k - i * r,    i * k - j * r, j * k + i * r, -(i * i + j * j)};
  const float dA[4][9] = {
      {0.f, -k, j, k, 0.f, -i, -j, i, 0.f},
      {0.f, j, k, j, -2.f * i, -r, k, r, -2.f * i},
      {-2.f * j, i, r, i, 0.f, k, -r, k, -2.f * j},
      {-2.f * k, -r, i, r, -2.f * k, j, i, j, 0.f},
  };
  float gA = 0.f;
  float GR[9];
#pragma unroll
  for (int x = 0; x < 3; ++x)
#pragma unroll
    for (int y = 0; y < 3; ++y) {
      GR[x * 3 + y] = G[x * 4 + y];
      gA += GR[x * 3 + y] * A[x * 3 + y];
    }
  float g7[7];
#pragma unroll
  for (int mq = 0; mq < 4; ++mq) {
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < 9; ++e) acc += GR[e] * dA[mq][e];
    g7[mq] = s * acc - s * s * q[mq] * gA;  // ds/dq_m = -s^2 q_m
  }
  g7[4] = G[3];
  g7[5] = G[7];
  g7[6] = G[11];
  if (a.grad7) {
#pragma unroll
    for (int e = 0; e < 7; ++e) a.grad7[f * 7 + e] = g7[e];
  }
  if (a.apply) {
    AdamArgs aa;
    aa.beta2 = a.beta2;
    aa.one_m_beta1 = a.one_m_beta1;
    aa.one_m_beta2 = a.one_m_beta2;
    aa.bc2_sqrt = a.bc2_sqrt;
    aa.eps = a.eps;
#pragma unroll
    for (int e = 0; e < 7; ++e) {
      float p = q[e], m = a.m[f * 7 + e], v = a.v[f * 7 + e];
      adam_one(p, g7[e], m, v, aa, e < 4 ? a.step_q : a.step_t);
      q[e] = p;
      a.m[f * 7 + e] = m;
      a.v[f * 7 + e] = v;
    }
  }
#pragma unroll
  for (int e = 0; e < 12; ++e) G[e] = 0.f;
}

// ---- loss finalisation ---------------------------------------------------------------------------------
struct FinalizeArgs {
  const int* counters;
  double* loss_acc;
  float* loss_out;
  int tracker_rule;
  float w_fs, w_center, w_tail, w_depth;
  double w_color;
};

__global__ void k_finalize_loss(const __grid_constant__ FinalizeArgs a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const int R = a.counters[0], nm = a.counters[2];
  // torch.mean over an empty selection is NaN (0/0); kept on purpose
  const float fs = (float)a.loss_acc[0] / (float)a.counters[3];
  const float ce = (float)a.loss_acc[1] / (float)a.counters[4];
  const float ta = (float)a.loss_acc[2] / (float)a.counters[5];
  const float dep = (float)a.loss_acc[3] / (float)nm;
  const double ncol = 3.0 * (double)(a.tracker_rule ? nm : R);
  // gt colour is float64 in the reference (datasets.py:90), so the colour term and the total are float64
  const float sdf_part = a.w_fs * fs + a.w_center * ce + a.w_tail * ta;
  double loss = (double)sdf_part + a.w_color * (a.loss_acc[4] / ncol);
  loss = loss + (double)(a.w_depth * dep);
  a.loss_acc[5] = loss;
  a.loss_acc[6] = a.loss_acc[4] / ncol;
  if (a.loss_out) *a.loss_out = (float)loss;
  for (int e = 0; e < 5; ++e) a.loss_acc[e] = 0.0;
}

}  // namespace eslam
