// Frame ingest on the device: the arithmetic half of BaseDataset.__getitem__
// (src/utils/datasets.py:79-115) after cv2 has decoded the files: BGR -> RGB, /255 in float64, uint16 depth ->
// float32 / png_depth_scale * scale, crop_edge.  The host ships 4 bytes per pixel + 2 instead of 24 + 4.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace eslam {

struct IngestArgs {
  const unsigned char* bgr;     // [H][W][3] as cv2.imread returns it
  const unsigned short* depth;  // [H][W] as cv2.imread(..., IMREAD_UNCHANGED) returns a 16-bit png
  int H, W, edge;
  float png_depth_scale, scale;
  double* color;  // [H-2e][W-2e][3] RGB in [0,1]
  float* out_depth;  // [H-2e][W-2e]
};

__global__ void __launch_bounds__(256) k_ingest_frame(const __grid_constant__ IngestArgs a) {
  const int Wc = a.W - 2 * a.edge, Hc = a.H - 2 * a.edge;
  const long long n = (long long)Wc * Hc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / Wc), c = (int)(i - (long long)r * Wc);
    const long long src = (long long)(r + a.edge) * a.W + (c + a.edge);
    const unsigned char* px = a.bgr + src * 3;
    // color_data / 255. on a uint8 array: float64 division (datasets.py:90)
    a.color[i * 3 + 0] = (double)px[2] / 255.0;
    a.color[i * 3 + 1] = (double)px[1] / 255.0;
    a.color[i * 3 + 2] = (double)px[0] / 255.0;
    // depth_data.astype(np.float32) / png_depth_scale, then * scale in float32 (datasets.py:91,95)
    a.out_depth[i] = __fmul_rn(__fdiv_rn((float)a.depth[src], a.png_depth_scale), a.scale);
  }
}

}  // namespace eslam
