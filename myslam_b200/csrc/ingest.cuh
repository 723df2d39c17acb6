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

// ScanNet-shaped frames (datasets.py:88-112): the colour image (1296x968) is larger than the depth image (640x480); the
// loader divides by 255 in float64, cv2.resize()s the float64 image to the depth's size (INTER_LINEAR) and crops the
// edge.  The resize is restated from the opencv-python wheels' behaviour (Intel IPP: float64 weights, row pass then
// column pass, each tap fma(w, b - a, a); oracle/eslam_oracle.py:ingest_frame_resized pins it bit for bit against the
// reference's loader).  One thread per output pixel; the four taps of a pixel are 12 bytes apart in rows 2.6 KB apart.
struct IngestResizeArgs {
  const unsigned char* bgr;     // [Hs][Ws][3]
  const unsigned short* depth;  // [H][W]
  int Hs, Ws, H, W, edge;
  double sx, sy;                // Ws / W, Hs / H
  float png_depth_scale, scale;
  double* color;     // [H-2e][W-2e][3]
  float* out_depth;  // [H-2e][W-2e]
};

__global__ void __launch_bounds__(256) k_ingest_resize(const __grid_constant__ IngestResizeArgs a) {
  const int Wc = a.W - 2 * a.edge, Hc = a.H - 2 * a.edge;
  const long long n = (long long)Wc * Hc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / Wc), c = (int)(i - (long long)r * Wc);
    const int dy = r + a.edge, dx = c + a.edge;
    const double fxx = __dsub_rn(__dmul_rn(__dadd_rn((double)dx, 0.5), a.sx), 0.5);
    const double fyy = __dsub_rn(__dmul_rn(__dadd_rn((double)dy, 0.5), a.sy), 0.5);
    const double flx = floor(fxx), fly = floor(fyy);
    const double wx = __dsub_rn(fxx, flx), wy = __dsub_rn(fyy, fly);
    const int x0 = min(max((int)flx, 0), a.Ws - 1), x1 = min(max((int)flx + 1, 0), a.Ws - 1);
    const int y0 = min(max((int)fly, 0), a.Hs - 1), y1 = min(max((int)fly + 1, 0), a.Hs - 1);
    const unsigned char* p00 = a.bgr + ((long long)y0 * a.Ws + x0) * 3;
    const unsigned char* p01 = a.bgr + ((long long)y0 * a.Ws + x1) * 3;
    const unsigned char* p10 = a.bgr + ((long long)y1 * a.Ws + x0) * 3;
    const unsigned char* p11 = a.bgr + ((long long)y1 * a.Ws + x1) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int s = 2 - ch;  // BGR -> RGB
      const double v00 = (double)p00[s] / 255.0, v01 = (double)p01[s] / 255.0;
      const double v10 = (double)p10[s] / 255.0, v11 = (double)p11[s] / 255.0;
      const double h0 = fma(wx, __dsub_rn(v01, v00), v00), h1 = fma(wx, __dsub_rn(v11, v10), v10);
      a.color[i * 3 + ch] = fma(wy, __dsub_rn(h1, h0), h0);
    }
    a.out_depth[i] = __fmul_rn(__fdiv_rn((float)a.depth[(long long)dy * a.W + dx], a.png_depth_scale), a.scale);
  }
}

}  // namespace eslam
