// Micro-benchmarks behind DESIGN.md's L2-resident roofline.  The hot path's access shape is random channels-last texel
// lines out of an L2-resident plane set (27 MB on Replica room0): 128-byte lines in the parameter form (8 lanes x
// float4), 64-byte lines in the Q form (4 lanes x float4); gathered with LDG.128 and scattered with
// red.global.add.v4.f32.  Measured here, per line size and CTAs per SM:
//   ldg      random line gathers, LDG.128 per lane (what the kernels do)
//   red      red.global.add.v4.f32 per lane (what the kernels do)
//   ldg+red  both interleaved
//   tma.g4   cp.async.bulk.tensor.2d.tile::gather4 -- four ARBITRARY rows of the arena viewed as [lines][line floats]
//            (= the four corner lines of one bilinear tap) into shared memory behind an mbarrier, one issuing lane
//            per warp, 4 stages in flight per warp; consumers read the rows back from shared memory
//   bulk.red cp.reduce.async.bulk.global.shared::cta.add.f32 of one line per instruction from shared memory, one
//            issuing lane per warp
// i.e. whether Blackwell's bulk-async data movement beats the per-lane LDG / RED form for this shape.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_gather_red l2_gather_red.cu && ./l2_gather_red [--json]
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cuda.h>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned hash32(unsigned x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}

// LPL: lanes per line (8 -> 128 B, 4 -> 64 B).  MODE 0 gather, 1 red, 2 both
template <int MODE, int LPL>
__global__ void __launch_bounds__(256) k_lane(float4* buf, unsigned n_lines, int iters, float4* sink) {
  const unsigned gtid = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned grp = gtid / LPL, sub = gtid % LPL;
  float4 acc = make_float4(0, 0, 0, 0);
  for (int it = 0; it < iters; it += 4) {
    float4 v[4];
    unsigned line[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) line[j] = hash32(grp * 977u + (it + j) * 7919u) % n_lines;
    if (MODE != 1) {
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = __ldg(buf + (size_t)line[j] * LPL + sub);
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc.x += v[j].x; acc.y += v[j].y; acc.z += v[j].z; acc.w += v[j].w; }
    }
    if (MODE != 0) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float4* p = buf + (size_t)(n_lines + line[j]) * LPL + sub;
        asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(1.f), "f"(2.f), "f"(3.f), "f"(acc.x));
      }
    }
  }
  if (acc.x == 123.456f) sink[gtid] = acc;
}

// ---- bulk-async variants ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, int bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// bounded wait: a transfer that never completes must not hang the GPU
__device__ __forceinline__ bool mbar_wait(unsigned long long* bar, unsigned parity) {
  for (int spin = 0; spin < (1 << 22); ++spin) {
    unsigned ok;
    asm volatile(
        "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

constexpr int STAGES = 4;

// LF: floats per line (32 or 16).  One warp = one issuing lane + 32 consumers; per stage 4 rows of LF floats.
template <int LF>
__global__ void __launch_bounds__(256) k_tma_gather4(const __grid_constant__ CUtensorMap tmap, unsigned n_lines, int iters,
                                                     float4* sink, int* fail) {
  __shared__ __align__(128) float stage[8][STAGES][4 * LF];
  __shared__ __align__(8) unsigned long long bar[8][STAGES];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned gw = blockIdx.x * 8 + warp;
  if (lane == 0)
    for (int s = 0; s < STAGES; ++s) mbar_init(&bar[warp][s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  float4 acc = make_float4(0, 0, 0, 0);
  const int n_groups = iters / 4;  // groups of 4 rows
  auto issue = [&](int gi) {
    const int s = gi % STAGES;
    int row[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) row[j] = (int)(hash32(gw * 977u + (gi * 4 + j) * 7919u) % n_lines);
    mbar_expect_tx(&bar[warp][s], 4 * LF * 4);
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cta.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, "
        "%5, %6}], [%7];" ::"r"(smem_u32(&stage[warp][s][0])),
        "l"(&tmap), "r"(0), "r"(row[0]), "r"(row[1]), "r"(row[2]), "r"(row[3]), "r"(smem_u32(&bar[warp][s]))
        : "memory");
  };
  if (lane == 0)
    for (int gi = 0; gi < STAGES && gi < n_groups; ++gi) issue(gi);
  for (int gi = 0; gi < n_groups; ++gi) {
    const int s = gi % STAGES;
    if (!mbar_wait(&bar[warp][s], (gi / STAGES) & 1)) {
      if (lane == 0) atomicExch(fail, 1);
      return;
    }
    // consumers: the 4 rows = 4 * LF floats = LF float4; lanes < LF read one float4 each
    if (lane < LF) {
      const float4 v = reinterpret_cast<const float4*>(&stage[warp][s][0])[lane];
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    __syncwarp();
    if (lane == 0 && gi + STAGES < n_groups) issue(gi + STAGES);
  }
  if (acc.x == 123.456f) sink[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int LF>
__global__ void __launch_bounds__(256) k_bulk_red(float* buf, unsigned n_lines, int iters, int* fail) {
  __shared__ __align__(128) float line[8][LF];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const unsigned gw = blockIdx.x * 8 + warp;
  if (lane < LF) line[warp][lane] = 1.0f;
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncwarp();
  if (lane != 0) return;
  for (int it = 0; it < iters; ++it) {
    const unsigned l = hash32(gw * 977u + it * 7919u) % n_lines;
    float* dst = buf + (size_t)(n_lines + l) * LF;
    asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;" ::"l"(dst),
                 "r"(smem_u32(&line[warp][0])), "r"(LF * 4)
                 : "memory");
    if ((it & 7) == 7) {
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 2;" ::: "memory");
    }
  }
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  (void)fail;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static bool make_map(CUtensorMap* m, void* base, unsigned n_lines, int lf) {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) != cudaSuccess || !fn) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)lf, (cuuint64_t)n_lines};
  const cuuint64_t strides[1] = {(cuuint64_t)lf * 4};
  const cuuint32_t box[2] = {(cuuint32_t)lf, 1};  // gather4: the box is one row; the instruction names four rows
  const cuuint32_t es[2] = {1, 1};
  return ((EncodeFn)fn)(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename F>
static float best_ms(F launch, cudaEvent_t e0, cudaEvent_t e1) {
  float best = 1e9f;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(e0);
    launch();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  const bool json = argc > 1 && !strcmp(argv[1], "--json");
  const size_t region = 27u * 1024 * 1024;  // the plane set (+ a second region of the same size for the reductions)
  float4* buf;
  cudaMalloc(&buf, region * 2);
  cudaMemset(buf, 0, region * 2);
  float4* sink;
  cudaMalloc(&sink, 8 << 20);
  int* fail;
  cudaMalloc(&fail, 4);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int iters = 256;
  double best_gbs[2][5] = {{0}};  // [line 128 / 64][ldg, red, ldg+red, tma.g4, bulk.red]
  for (int li = 0; li < 2; ++li) {
    const int lf = li == 0 ? 32 : 16, lb = lf * 4, lpl = lf / 4;
    const unsigned n_lines = (unsigned)(region / lb);
    CUtensorMap tmap;
    const bool have_map = make_map(&tmap, buf, n_lines, lf);
    for (int bps : {2, 4, 8}) {
      const int grid = 148 * bps;
      for (int mode = 0; mode < 5; ++mode) {
        double lines;
        float ms;
        cudaMemset(fail, 0, 4);
        if (mode < 3) {
          lines = (double)grid * 256 / lpl * iters * (mode == 2 ? 2 : 1);
          ms = best_ms([&] {
            if (li == 0) {
              if (mode == 0) k_lane<0, 8><<<grid, 256>>>(buf, n_lines, iters, sink);
              if (mode == 1) k_lane<1, 8><<<grid, 256>>>(buf, n_lines, iters, sink);
              if (mode == 2) k_lane<2, 8><<<grid, 256>>>(buf, n_lines, iters, sink);
            } else {
              if (mode == 0) k_lane<0, 4><<<grid, 256>>>(buf, n_lines, iters, sink);
              if (mode == 1) k_lane<1, 4><<<grid, 256>>>(buf, n_lines, iters, sink);
              if (mode == 2) k_lane<2, 4><<<grid, 256>>>(buf, n_lines, iters, sink);
            }
          }, e0, e1);
        } else if (mode == 3) {
          if (!have_map) { printf("tma.g4    line %3d B: cuTensorMapEncodeTiled failed\n", lb); continue; }
          lines = (double)grid * 8 * iters;
          ms = best_ms([&] {
            if (li == 0) k_tma_gather4<32><<<grid, 256>>>(tmap, n_lines, iters, sink, fail);
            else k_tma_gather4<16><<<grid, 256>>>(tmap, n_lines, iters, sink, fail);
          }, e0, e1);
        } else {
          lines = (double)grid * 8 * iters;
          ms = best_ms([&] {
            if (li == 0) k_bulk_red<32><<<grid, 256>>>((float*)buf, n_lines, iters, fail);
            else k_bulk_red<16><<<grid, 256>>>((float*)buf, n_lines, iters, fail);
          }, e0, e1);
        }
        int h_fail = 0;
        cudaError_t err = cudaDeviceSynchronize();
        cudaMemcpy(&h_fail, fail, 4, cudaMemcpyDeviceToHost);
        const char* names[5] = {"ldg      ", "red.v4   ", "ldg+red  ", "tma.g4   ", "bulk.red "};
        if (err != cudaSuccess || h_fail) {
          printf("%s line %3d B  CTAs/SM %d  FAILED (%s%s)\n", names[mode], lb, bps, cudaGetErrorString(err),
                 h_fail ? ", mbarrier timeout" : "");
          if (err != cudaSuccess) return 1;
          continue;
        }
        const double gbs = lines * lb / ms / 1e6;
        if (gbs > best_gbs[li][mode]) best_gbs[li][mode] = gbs;
        printf("%s line %3d B  CTAs/SM %d  %8.1f us  %8.1f GB/s  (%.2f M lines, %.2f G lines/s)\n", names[mode], lb, bps,
               ms * 1e3, gbs, lines / 1e6, lines / ms / 1e6);
      }
    }
  }
  if (json)
    printf("{\"line128\": {\"ldg\": %.1f, \"red\": %.1f, \"ldg_red\": %.1f, \"tma_gather4\": %.1f, \"bulk_red\": %.1f}, "
           "\"line64\": {\"ldg\": %.1f, \"red\": %.1f, \"ldg_red\": %.1f, \"tma_gather4\": %.1f, \"bulk_red\": %.1f}, "
           "\"unit\": \"GB/s\"}\n",
           best_gbs[0][0], best_gbs[0][1], best_gbs[0][2], best_gbs[0][3], best_gbs[0][4], best_gbs[1][0], best_gbs[1][1],
           best_gbs[1][2], best_gbs[1][3], best_gbs[1][4]);
  printf("cuda status: %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
