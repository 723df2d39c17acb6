// Marching cubes over the SDF lattice ON THE DEVICE (Mesher.get_mesh, src/utils/Mesher.py:219-247: the reference copies
// the 1.3 GB volume to the host and calls skimage.measure.marching_cubes), frustum culling of the mesh vertices
// (src/tools/cull_mesh.py:58-100) and back-projection of keyframe depth pixels for the mesh bound (Mesher.py:63-128).
//
// Lattice: sdf[(iy * nx + ix) * nz + iz] (Mesher.py:179-184, meshgrid(indexing='xy')); cells are the (nx-1)(ny-1)(nz-1)
// cubes between neighbouring lattice points, cell id = (cy * (nx-1) + cx) * (nz-1) + cz, so consecutive threads walk z
// and every corner read is a coalesced row.  Case tables: myslam_b200/mc_tables.py (generated; conventions there).
// Two passes, no per-cell storage: k_mc_count leaves one triangle count per CTA, the host scans the ~1.3 M counts
// (torch.cumsum), k_mc_emit recomputes the configuration and writes the CTA's triangles at its base offset.  A vertex is
// identified by the lattice edge it lies on (key = 3 * flat(lower corner) + axis), which is what welding needs.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace eslam {

constexpr int MC_THREADS = 256;
constexpr int MC_MAX_TRI = 5;

struct McArgs {
  const float* sdf;
  const float *xs, *ys, *zs;
  int nx, ny, nz;
  float level;
  const unsigned char* n_tri;  // [256]
  const signed char* tri;      // [256][3 * MC_MAX_TRI] edge ids
  int* block_count;            // [n_blocks] (count pass)
  const long long* block_base; // [n_blocks] exclusive scan of block_count (emit pass)
  float* verts;                // [3 T][3] triangle soup
  long long* keys;             // [3 T] lattice-edge key of every vertex
  long long n_cells;
};

__device__ __forceinline__ int mc_config(const McArgs& a, long long cell, int& cx, int& cy, int& cz, float (&v)[8]) {
  const int nzc = a.nz - 1, nxc = a.nx - 1;
  cz = (int)(cell % nzc);
  const long long t = cell / nzc;
  cx = (int)(t % nxc);
  cy = (int)(t / nxc);
  int cfg = 0;
#pragma unroll
  for (int c = 0; c < 8; ++c) {
    const int ix = cx + (c & 1), iy = cy + ((c >> 1) & 1), iz = cz + ((c >> 2) & 1);
    v[c] = __ldg(a.sdf + ((long long)iy * a.nx + ix) * a.nz + iz);
    cfg |= (v[c] < a.level) ? (1 << c) : 0;
  }
  return cfg;
}

__global__ void __launch_bounds__(MC_THREADS) k_mc_count(const __grid_constant__ McArgs a) {
  __shared__ int s_warp[MC_THREADS / 32];
  const long long cell = (long long)blockIdx.x * MC_THREADS + threadIdx.x;
  int n = 0;
  if (cell < a.n_cells) {
    int cx, cy, cz;
    float v[8];
    n = a.n_tri[mc_config(a, cell, cx, cy, cz, v)];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) n += __shfl_xor_sync(0xffffffffu, n, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = n;
  __syncthreads();
  if (threadIdx.x == 0) {
    int tot = 0;
    for (int w = 0; w < MC_THREADS / 32; ++w) tot += s_warp[w];
    a.block_count[blockIdx.x] = tot;
  }
}

__global__ void __launch_bounds__(MC_THREADS) k_mc_emit(const __grid_constant__ McArgs a) {
  __shared__ int s_warp[MC_THREADS / 32];
  const long long cell = (long long)blockIdx.x * MC_THREADS + threadIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int cfg = 0, n = 0, cx = 0, cy = 0, cz = 0;
  float v[8];
  if (cell < a.n_cells) {
    cfg = mc_config(a, cell, cx, cy, cz, v);
    n = a.n_tri[cfg];
  }
  // exclusive scan of the triangle counts inside the CTA
  int inc = n;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  int before = 0;
  for (int w = 0; w < warp; ++w) before += s_warp[w];
  if (n == 0) return;
  long long at = a.block_base[blockIdx.x] + before + inc - n;  // first triangle of this cell
  const float* ax[3] = {a.xs, a.ys, a.zs};
  const int ci[3] = {cx, cy, cz};
  for (int t = 0; t < n; ++t, ++at) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const int e = a.tri[cfg * (3 * MC_MAX_TRI) + 3 * t + k];
      const int axis = e >> 2, b0 = e & 1, b1 = (e >> 1) & 1;
      const int o0 = axis == 0 ? 1 : 0, o1 = axis == 2 ? 1 : 2;  // the two other axes, increasing
      int off[3] = {0, 0, 0};
      off[o0] = b0;
      off[o1] = b1;
      const int c0 = off[0] | (off[1] << 1) | (off[2] << 2), c1 = c0 | (1 << axis);
      const float v0 = v[c0], v1 = v[c1];
      const float w = (a.level - v0) / (v1 - v0);
      float p[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) p[d] = ax[d][ci[d] + off[d]];
      const float q = ax[axis][ci[axis] + 1];
      p[axis] = fmaf(w, q - p[axis], p[axis]);
      const long long vi = at * 3 + k;
      a.verts[vi * 3 + 0] = p[0];
      a.verts[vi * 3 + 1] = p[1];
      a.verts[vi * 3 + 2] = p[2];
      const long long flat = ((long long)(cy + off[1]) * a.nx + (cx + off[0])) * a.nz + (cz + off[2]);
      a.keys[vi] = flat * 3 + axis;
    }
  }
}

// ---- frustum culling of mesh vertices (cull_mesh.py:58-100), one launch per frame ----------------------------------
// seen[v] |= the vertex projects into the frame (in front of the camera, inside the image) and, with eval_rec, is not
// more than `truncation` behind the frame's depth.  The reference's arithmetic: camera coordinates with x negated,
// uv = K cam, z = uv.z + 1e-5, uv /= z, depth sampled bilinearly (grid_sample, zeros padding, align_corners=True) at
// (uv.x / W, uv.y / H) mapped to [-1, 1].
struct CullArgs {
  const float* verts;  // [n][3] world
  long long n;
  const float* w2c;    // [16] row-major inverse of the frame's c2w
  const float* depth;  // [H][W]
  int H, W;
  float fx, fy, cx, cy, truncation;
  int eval_rec;
  unsigned char* seen;
};

__global__ void __launch_bounds__(256) k_cull_frame(const __grid_constant__ CullArgs a) {
  const long long i = (long long)blockIdx.x * 256 + threadIdx.x;
  if (i >= a.n || a.seen[i]) return;
  const float x = a.verts[i * 3], y = a.verts[i * 3 + 1], z = a.verts[i * 3 + 2];
  float cam[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) cam[r] = a.w2c[r * 4] * x + a.w2c[r * 4 + 1] * y + a.w2c[r * 4 + 2] * z + a.w2c[r * 4 + 3];
  cam[0] = -cam[0];
  const float zz = cam[2] + 1e-5f;
  const float u = (a.fx * cam[0] + a.cx * cam[2]) / zz, v = (a.fy * cam[1] + a.cy * cam[2]) / zz;
  bool vis = (0.f <= -zz) && (u < (float)a.W) && (u > 0.f) && (v < (float)a.H) && (v > 0.f);
  if (vis && a.eval_rec) {
    // grid_sample(align_corners=True): pixel = (g + 1) / 2 * (size - 1), g = 2 * (u / W) - 1
    const float px = (u / (float)a.W) * (float)(a.W - 1), py = (v / (float)a.H) * (float)(a.H - 1);
    const float fx0 = floorf(px), fy0 = floorf(py);
    const int x0 = (int)fx0, y0 = (int)fy0;
    const float wx = px - fx0, wy = py - fy0;
    auto at = [&](int yy, int xx) { return (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) ? a.depth[(long long)yy * a.W + xx] : 0.f; };
    const float d = at(y0, x0) * (1.f - wx) * (1.f - wy) + at(y0, x0 + 1) * wx * (1.f - wy) + at(y0 + 1, x0) * (1.f - wx) * wy +
                    at(y0 + 1, x0 + 1) * wx * wy;
    vis = (d + a.truncation >= -zz);
  }
  if (vis) a.seen[i] = 1;
}

}  // namespace eslam
