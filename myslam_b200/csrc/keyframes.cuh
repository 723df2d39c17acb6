// Keyframe selection by view overlap: the projection count of Mapper.keyframe_selection_overlap
// (src/Mapper.py:146-203), one CTA per keyframe.  SURVEY.md section 8(f)-1.
#pragma once
#include "field.cuh"
#include "sample.cuh"

namespace eslam {

struct OverlapArgs {
  const float* c2w;          // [16] current camera
  const float* depth;        // [H][W] current frame
  const long long* pix_idx;  // [n_rays] draws of randint(H*W)
  const float* t_vals;       // [n_samples] torch.linspace(0, 1, n_samples)
  const float* kf_c2w;       // [K][16]
  int n_rays, n_samples, K;
  int H, W;
  float fx, fy, cx, cy;
  int* inside;               // [K] points of the current view that keyframe k sees
  int* n_pts;                // [1] points tested (depth > 0 rays x n_samples)
};

constexpr int OVERLAP_THREADS = 128;

__global__ void __launch_bounds__(OVERLAP_THREADS) k_keyframe_overlap(const __grid_constant__ OverlapArgs a) {
  __shared__ float w2c[12];
  __shared__ float cw[12];
  const int k = blockIdx.x;
  if (threadIdx.x < 12) cw[threadIdx.x] = a.c2w[threadIdx.x];
  if (threadIdx.x == 0) {
    // affine inverse of the keyframe's c2w (what torch.inverse returns for [R t; 0 1], Mapper.py:180)
    const float* m = a.kf_c2w + k * 16;
    const float r00 = m[0], r01 = m[1], r02 = m[2], r10 = m[4], r11 = m[5], r12 = m[6], r20 = m[8], r21 = m[9], r22 = m[10];
    const float c00 = r11 * r22 - r12 * r21, c01 = r12 * r20 - r10 * r22, c02 = r10 * r21 - r11 * r20;
    const float inv_det = 1.0f / (r00 * c00 + r01 * c01 + r02 * c02);
    float inv[9];
    inv[0] = c00 * inv_det;
    inv[1] = (r02 * r21 - r01 * r22) * inv_det;
    inv[2] = (r01 * r12 - r02 * r11) * inv_det;
    inv[3] = c01 * inv_det;
    inv[4] = (r00 * r22 - r02 * r20) * inv_det;
    inv[5] = (r02 * r10 - r00 * r12) * inv_det;
    inv[6] = c02 * inv_det;
    inv[7] = (r01 * r20 - r00 * r21) * inv_det;
    inv[8] = (r00 * r11 - r01 * r10) * inv_det;
    const float tx = m[3], ty = m[7], tz = m[11];
    for (int r = 0; r < 3; ++r) {
      w2c[r * 4 + 0] = inv[r * 3 + 0];
      w2c[r * 4 + 1] = inv[r * 3 + 1];
      w2c[r * 4 + 2] = inv[r * 3 + 2];
      w2c[r * 4 + 3] = -(inv[r * 3 + 0] * tx + inv[r * 3 + 1] * ty + inv[r * 3 + 2] * tz);
    }
  }
  __syncthreads();
  int cnt = 0, tested = 0;
  const int total = a.n_rays * a.n_samples;
  for (int p = threadIdx.x; p < total; p += OVERLAP_THREADS) {
    const int r = p / a.n_samples, s = p - r * a.n_samples;
    const long long pix = a.pix_idx[r];
    const float d = a.depth[pix];
    if (!(d > 0.f)) continue;  // Mapper.py:167-170
    ++tested;
    // get_rays_from_uv (common.py:87-99), full image: i = column, j = row
    const float pi = (float)(pix % a.W), pj = (float)(pix / a.W);
    const float dx = __fdiv_rn(__fsub_rn(pi, a.cx), a.fx), dy = -__fdiv_rn(__fsub_rn(pj, a.cy), a.fy), dz = -1.0f;
    const float t = a.t_vals[s];
    const float z = __fadd_rn(__fmul_rn(__fmul_rn(d, 0.8f), __fsub_rn(1.0f, t)), __fmul_rn(__fadd_rn(d, 0.5f), t));
    float pt[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      const float rd = __fadd_rn(__fadd_rn(__fmul_rn(dx, cw[c * 4 + 0]), __fmul_rn(dy, cw[c * 4 + 1])),
                                 __fmul_rn(dz, cw[c * 4 + 2]));
      pt[c] = __fadd_rn(cw[c * 4 + 3], __fmul_rn(rd, z));
    }
    float cam[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) cam[c] = w2c[c * 4 + 0] * pt[0] + w2c[c * 4 + 1] * pt[1] + w2c[c * 4 + 2] * pt[2] + w2c[c * 4 + 3];
    cam[0] = -cam[0];                              // Mapper.py:190
    const float zc = cam[2] + 1e-5f;               // K's last row is (0, 0, 1)
    const float u = (a.fx * cam[0] + a.cx * cam[2]) / zc;
    const float v = (a.fy * cam[1] + a.cy * cam[2]) / zc;
    const float edge = 20.f;
    if (u < (float)a.W - edge && u > edge && v < (float)a.H - edge && v > edge && zc < 0.f) ++cnt;
  }
  __shared__ int s_cnt, s_tested;
  if (threadIdx.x == 0) s_cnt = 0, s_tested = 0;
  __syncthreads();
  if (cnt) atomicAdd(&s_cnt, cnt);
  if (tested) atomicAdd(&s_tested, tested);
  __syncthreads();
  if (threadIdx.x == 0) {
    a.inside[k] = s_cnt;
    if (k == 0) a.n_pts[0] = s_tested;
  }
}

// ---- pose <-> matrix for a window of cameras (common.py:155-181 over pytorch3d 0.7.1's quaternion functions), one
//      thread per camera, in torch's evaluation order (bit-exact with the host mirror myslam_b200/common.py).
//      Replaces ~60 tiny ATen launches at the start and end of every optimize_mapping call.
__global__ void k_matrix_to_pose(const float* __restrict__ c2w, float* __restrict__ poses, int n) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  const float* m = c2w + f * 16;
  const float m00 = m[0], m01 = m[1], m02 = m[2], m10 = m[4], m11 = m[5], m12 = m[6], m20 = m[8], m21 = m[9], m22 = m[10];
  float sq[4], qa[4];
  sq[0] = __fadd_rn(__fadd_rn(__fadd_rn(1.0f, m00), m11), m22);
  sq[1] = __fsub_rn(__fsub_rn(__fadd_rn(1.0f, m00), m11), m22);
  sq[2] = __fsub_rn(__fadd_rn(__fsub_rn(1.0f, m00), m11), m22);
  sq[3] = __fadd_rn(__fsub_rn(__fsub_rn(1.0f, m00), m11), m22);
  int best = 0;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    qa[i] = sq[i] > 0.f ? __fsqrt_rn(sq[i]) : 0.f;
    if (qa[i] > qa[best]) best = i;  // torch.argmax: first maximum
  }
  float c[4];
  const float d = __fmul_rn(qa[best], qa[best]);
  switch (best) {
    case 0: c[0] = d; c[1] = __fsub_rn(m21, m12); c[2] = __fsub_rn(m02, m20); c[3] = __fsub_rn(m10, m01); break;
    case 1: c[0] = __fsub_rn(m21, m12); c[1] = d; c[2] = __fadd_rn(m10, m01); c[3] = __fadd_rn(m02, m20); break;
    case 2: c[0] = __fsub_rn(m02, m20); c[1] = __fadd_rn(m10, m01); c[2] = d; c[3] = __fadd_rn(m12, m21); break;
    default: c[0] = __fsub_rn(m10, m01); c[1] = __fadd_rn(m20, m02); c[2] = __fadd_rn(m21, m12); c[3] = d; break;
  }
  const float den = __fmul_rn(2.0f, fmaxf(qa[best], 0.1f));
#pragma unroll
  for (int i = 0; i < 4; ++i) poses[f * 7 + i] = __fdiv_rn(c[i], den);
  poses[f * 7 + 4] = m[3];
  poses[f * 7 + 5] = m[7];
  poses[f * 7 + 6] = m[11];
}

__global__ void k_pose_to_matrix(const float* __restrict__ poses, float* __restrict__ c2w, int n) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n) return;
  float R[9];
  quat_to_rot(poses + f * 7, R);
  float* m = c2w + f * 16;
#pragma unroll
  for (int x = 0; x < 3; ++x) {
    m[x * 4 + 0] = R[x * 3 + 0];
    m[x * 4 + 1] = R[x * 3 + 1];
    m[x * 4 + 2] = R[x * 3 + 2];
    m[x * 4 + 3] = poses[f * 7 + 4 + x];
  }
  m[12] = 0.f;
  m[13] = 0.f;
  m[14] = 0.f;
  m[15] = 1.f;
}

}  // namespace eslam
