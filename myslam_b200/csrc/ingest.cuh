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

// TUM-shaped frames (datasets.py:83-86,98-106): cv2.undistort of the uint8 colour image, then "crop_size" -- a bilinear
// align_corners resize of the float64 colour and a nearest resize of the depth -- and crop_edge.
//
// k_undistort_u8 restates cv2.undistort (initUndistortRectifyMap + remap INTER_LINEAR / BORDER_CONSTANT, new camera
// matrix = K) as OpenCV publishes it: the source position of a pixel in float64 (operation order of the oracle's numpy
// restatement, oracle/eslam_oracle.py:undistort_u8, which is pinned against cv2 itself), rounded to 1/32 pixel, the four
// taps blended with integer weights that sum to 2^15.
struct UndistortArgs {
  const unsigned char* src;  // [H][W][3]
  unsigned char* dst;        // [H][W][3]
  int H, W;
  double ir[9];              // inverse of the camera matrix, row-major
  double fx, fy, cx, cy, k1, k2, p1, p2, k3;
};

__global__ void __launch_bounds__(256) k_undistort_u8(const __grid_constant__ UndistortArgs a) {
  const long long n = (long long)a.H * a.W;
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n; p += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(p / a.W), j = (int)(p - (long long)i * a.W);
    const double di = (double)i, dj = (double)j;
    const double _x = __dadd_rn(__dadd_rn(__dmul_rn(di, a.ir[1]), a.ir[2]), __dmul_rn(dj, a.ir[0]));
    const double _y = __dadd_rn(__dadd_rn(__dmul_rn(di, a.ir[4]), a.ir[5]), __dmul_rn(dj, a.ir[3]));
    const double _w = __dadd_rn(__dadd_rn(__dmul_rn(di, a.ir[7]), a.ir[8]), __dmul_rn(dj, a.ir[6]));
    const double w = __ddiv_rn(1.0, _w);
    const double x = __dmul_rn(_x, w), y = __dmul_rn(_y, w);
    const double x2 = __dmul_rn(x, x), y2 = __dmul_rn(y, y);
    const double r2 = __dadd_rn(x2, y2), xy2 = __dmul_rn(__dmul_rn(2.0, x), y);
    const double kr = __dadd_rn(
        1.0, __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(a.k3, r2), a.k2), r2), a.k1), r2));
    const double xd = __dadd_rn(__dadd_rn(__dmul_rn(x, kr), __dmul_rn(a.p1, xy2)),
                                __dmul_rn(a.p2, __dadd_rn(r2, __dmul_rn(2.0, x2))));
    const double yd = __dadd_rn(__dadd_rn(__dmul_rn(y, kr), __dmul_rn(a.p1, __dadd_rn(r2, __dmul_rn(2.0, y2)))),
                                __dmul_rn(a.p2, xy2));
    const long long iu = (long long)rint(__dmul_rn(__dadd_rn(__dmul_rn(a.fx, xd), a.cx), 32.0));
    const long long iv = (long long)rint(__dmul_rn(__dadd_rn(__dmul_rn(a.fy, yd), a.cy), 32.0));
    const long long sx = iu >> 5, sy = iv >> 5;
    const int ax = (int)(iu & 31), ay = (int)(iv & 31);
    const int w00 = (32 - ay) * (32 - ax) * 32, w01 = (32 - ay) * ax * 32, w10 = ay * (32 - ax) * 32, w11 = ay * ax * 32;
    auto tap = [&](long long yy, long long xx, int ch) -> int {
      return (yy >= 0 && yy < a.H && xx >= 0 && xx < a.W) ? (int)a.src[(yy * a.W + xx) * 3 + ch] : 0;
    };
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int acc = tap(sy, sx, ch) * w00 + tap(sy, sx + 1, ch) * w01 + tap(sy + 1, sx, ch) * w10 +
                      tap(sy + 1, sx + 1, ch) * w11;
      a.dst[p * 3 + ch] = (unsigned char)((acc + (1 << 14)) >> 15);
    }
  }
}

// crop_size + crop_edge.  The bilinear resize is torch's CPU kernel for float64 (F.interpolate, align_corners=True)
// bit for bit, as the oracle pins it: position = i (n_in - 1) / (n_out - 1) in float64, lower index from floorf() of
// the position rounded to float32, the four weight products, then (w01 b) -> fma(w00, a) -> fma(w10, c) -> fma(w11, d).
// The depth is F.interpolate(mode='nearest'): min(floorf(i * (float)(n_in / n_out)), n_in - 1).
struct IngestCropArgs {
  const unsigned char* bgr;     // [H][W][3]
  const unsigned short* depth;  // [H][W]
  int H, W, Ho, Wo, edge;
  double sy, sx;                // (H - 1) / (Ho - 1), (W - 1) / (Wo - 1)
  float ny, nx;                 // (float)H / Ho, (float)W / Wo
  float png_depth_scale, scale;
  double* color;     // [Ho-2e][Wo-2e][3]
  float* out_depth;  // [Ho-2e][Wo-2e]
};

__device__ __forceinline__ void align_corners_tap(double scale, int dst, int n_in, int& i0, int& i1, double& l0,
                                                  double& l1) {
  const double real = __dmul_rn(scale, (double)dst);
  i0 = min((int)floorf((float)real), n_in - 1);
  l1 = fmin(fmax(__dsub_rn(real, (double)i0), 0.0), 1.0);
  l0 = __dsub_rn(1.0, l1);
  i1 = i0 + (i0 < n_in - 1 ? 1 : 0);
}

__global__ void __launch_bounds__(256) k_ingest_crop(const __grid_constant__ IngestCropArgs a) {
  const int Wc = a.Wo - 2 * a.edge, Hc = a.Ho - 2 * a.edge;
  const long long n = (long long)Wc * Hc;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / Wc), c = (int)(i - (long long)r * Wc);
    const int dy = r + a.edge, dx = c + a.edge;
    int y0, y1, x0, x1;
    double hy, ly, hx, lx;
    align_corners_tap(a.sy, dy, a.H, y0, y1, hy, ly);
    align_corners_tap(a.sx, dx, a.W, x0, x1, hx, lx);
    const double w00 = __dmul_rn(hy, hx), w01 = __dmul_rn(hy, lx), w10 = __dmul_rn(ly, hx), w11 = __dmul_rn(ly, lx);
    const unsigned char* p00 = a.bgr + ((long long)y0 * a.W + x0) * 3;
    const unsigned char* p01 = a.bgr + ((long long)y0 * a.W + x1) * 3;
    const unsigned char* p10 = a.bgr + ((long long)y1 * a.W + x0) * 3;
    const unsigned char* p11 = a.bgr + ((long long)y1 * a.W + x1) * 3;
#pragma unroll
    for (int ch = 0; ch < 3; ++ch) {
      const int s = 2 - ch;  // BGR -> RGB
      double acc = __dmul_rn(w01, (double)p01[s] / 255.0);
      acc = __fma_rn(w00, (double)p00[s] / 255.0, acc);
      acc = __fma_rn(w10, (double)p10[s] / 255.0, acc);
      a.color[i * 3 + ch] = __fma_rn(w11, (double)p11[s] / 255.0, acc);
    }
    const int sy = min((int)floorf(__fmul_rn((float)dy, a.ny)), a.H - 1);
    const int sx = min((int)floorf(__fmul_rn((float)dx, a.nx)), a.W - 1);
    a.out_depth[i] = __fmul_rn(__fdiv_rn((float)a.depth[(long long)sy * a.W + sx], a.png_depth_scale), a.scale);
  }
}

}  // namespace eslam
