// Micro-benchmark behind DESIGN.md's L2-resident roofline: random 128-byte-line gathers (8 lanes x float4 per
// line, 4 lines per warp instruction -- the hot path's access shape) and red.global.add.v4.f32 scatters over a
// buffer the size of the Replica plane set (27 MB, L2-resident on B200).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather_red l2_gather_red.cu && ./l2_gather_red
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

template <int MODE>  // 0 gather, 1 red, 2 both
__global__ void __launch_bounds__(256) k(float4* buf, unsigned n_lines, int iters, float4* sink) {
  const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned grp = gtid >> 3, sub = gtid & 7;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int it = 0; it < iters; it += 4) {
    float4 v[4];
    unsigned line[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) line[j] = hash32(grp * 977u + (it + j) * 7919u) % n_lines;
    if (MODE != 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = __ldg(buf + (size_t)line[j] * 8 + sub);
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    }
    if (MODE != 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4* p = buf + (size_t)(n_lines + line[j]) * 8 + sub;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(acc.x));
      }
    }
  }
  if (acc.x == 123.456f) sink[gtid] = acc;
}

int main() {
  const unsigned n_lines = 27u * 1024 * 1024 / 128;  // 27 MB of 128 B lines (+ a second 27 MB region for the reds)
  float4* buf;
  cudaMalloc(&buf, (size_t)n_lines * 2 * 128);
  cudaMemset(buf, 0, (size_t)n_lines * 2 * 128);
  float4* sink;
  cudaMalloc(&sink, 1 << 20);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 256;
  for (int blocks_per_sm : {2, 4, 8}) {
    const int grid = 148 * blocks_per_sm;
    const double lines = (double)grid * 256 / 8 * iters;
    for (int mode = 0; mode < 3; ++mode) {
      float best = 1e9;
      for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        if (mode == 0) k<0><<<grid, 256>>>(buf, n_lines, iters, sink);
        if (mode == 1) k<1><<<grid, 256>>>(buf, n_lines, iters, sink);
        if (mode == 2) k<2><<<grid, 256>>>(buf, n_lines, iters, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      const double bytes = lines * 128 * (mode == 2 ? 2 : 1);
      printf("%s  CTAs/SM %d  %8.1f us  %7.1f GB/s  (%.1f M lines)\n",
             mode == 0 ? "gather      " : (mode == 1 ? "red.v4      " : "gather+red  "), blocks_per_sm, best * 1e3,
             bytes / best / 1e6, lines / 1e6);
    }
  }
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
