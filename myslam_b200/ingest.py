"""Frame ingest on the device: the arithmetic half of the reference's
`BaseDataset.__getitem__` (src/utils/datasets.py:79-115) after cv2 has decoded the colour and depth files.

    color_u8 = cv2.imread(color_path)                              # [H,W,3] uint8, BGR   (host, unchanged)
    depth_u16 = cv2.imread(depth_path, cv2.IMREAD_UNCHANGED)       # [H,W]   uint16       (host, unchanged)
    color, depth = ingest_frame(color_u8, depth_u16, png_depth_scale, crop_edge, device)

returns exactly what the reference's loader followed by `.to(device)` returns -- colour [H',W',3] float64 RGB in
[0,1], depth [H',W'] float32 -- but moves 8 bytes per pixel over PCIe instead of 28.  A colour image larger than the
depth image (ScanNet: 1296x968 vs 640x480) is resized to the depth's size on the device like the loader's
`cv2.resize(color_data, (W, H))`.  TUM-shaped frames: `distortion` (+ `cam`) runs the loader's
`cv2.undistort(color_data, K, distortion)` on the device, `crop_size` its bilinear (colour, align_corners) / nearest
(depth) resize (datasets.py:83-86,98-106).
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from ._lib import call, ptr, stream


def undistort(color_u8: torch.Tensor, cam, distortion) -> torch.Tensor:
    """cv2.undistort(img, K, distortion) with the new camera matrix = K (datasets.py:83-86) for an [H,W,3] uint8 CUDA
    tensor; cam = (fx, fy, cx, cy), distortion = (k1, k2, p1, p2, k3)."""
    if not color_u8.is_cuda or color_u8.dtype != torch.uint8 or color_u8.dim() != 3 or color_u8.shape[2] != 3:
        raise RuntimeError("undistort: [H,W,3] uint8 CUDA tensor expected; there is no CPU path")
    dist = [float(v) for v in np.asarray(distortion, dtype=np.float64).reshape(-1)]
    if len(dist) != 5:
        raise RuntimeError("undistort: distortion must be (k1, k2, p1, p2, k3) as in the reference's configs")
    fx, fy, cx, cy = (float(v) for v in cam)
    inv_k = np.linalg.inv(np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])).reshape(-1)
    src = color_u8.contiguous()
    dst = torch.empty_like(src)
    call("eslam_undistort_u8", ptr(src), ptr(dst), int(src.shape[0]), int(src.shape[1]), fx, fy, cx, cy,
         (C.c_double * 5)(*dist), (C.c_double * 9)(*inv_k), stream())
    return dst


def ingest_frame(color_u8, depth_u16, png_depth_scale: float, crop_edge: int = 0, device="cuda", scale: float = 1.0,
                 cam=None, distortion=None, crop_size=None):
    dev = torch.device(device)
    if dev.type != "cuda":
        raise RuntimeError("ingest_frame needs a CUDA device; myslam_b200 has no CPU path")
    if isinstance(color_u8, np.ndarray):
        color_u8 = torch.from_numpy(np.ascontiguousarray(color_u8))
    if isinstance(depth_u16, np.ndarray):
        if depth_u16.dtype != np.uint16:
            raise RuntimeError(f"depth must be the 16-bit png as decoded (uint16), got {depth_u16.dtype}")
        depth_u16 = torch.from_numpy(np.ascontiguousarray(depth_u16).view(np.int16))  # same bits; torch-friendly dtype
    if color_u8.dtype != torch.uint8 or color_u8.dim() != 3 or color_u8.shape[2] != 3:
        raise RuntimeError("colour must be [H,W,3] uint8 (BGR, as cv2.imread returns it)")
    if depth_u16.dtype not in (torch.int16, torch.uint16) or depth_u16.dim() != 2:
        raise RuntimeError("depth must be [H,W] 16-bit")
    H, W = int(depth_u16.shape[0]), int(depth_u16.shape[1])
    Hs, Ws = int(color_u8.shape[0]), int(color_u8.shape[1])
    e = int(crop_edge)
    c = color_u8.contiguous().to(dev, non_blocking=True)
    d = depth_u16.contiguous().to(dev, non_blocking=True)
    if distortion is not None:  # datasets.py:83-86: only the colour image is undistorted
        if cam is None:
            raise RuntimeError("ingest_frame: distortion needs cam = (fx, fy, cx, cy)")
        c = undistort(c, cam, distortion)
    if crop_size is not None:  # datasets.py:98-106
        if (Hs, Ws) != (H, W):
            raise RuntimeError("ingest_frame: crop_size with colour and depth images of different sizes is not a "
                               "configuration the reference ships; not implemented")
        Ho, Wo = int(crop_size[0]), int(crop_size[1])
        color = torch.empty(Ho - 2 * e, Wo - 2 * e, 3, dtype=torch.float64, device=dev)
        depth = torch.empty(Ho - 2 * e, Wo - 2 * e, dtype=torch.float32, device=dev)
        call("eslam_ingest_frame_crop", ptr(c), ptr(d), H, W, Ho, Wo, e, float(png_depth_scale), float(scale),
             ptr(color), ptr(depth), stream())
        return color, depth
    color = torch.empty(H - 2 * e, W - 2 * e, 3, dtype=torch.float64, device=dev)
    depth = torch.empty(H - 2 * e, W - 2 * e, dtype=torch.float32, device=dev)
    if (Hs, Ws) == (H, W):
        call("eslam_ingest_frame", ptr(c), ptr(d), H, W, e, float(png_depth_scale), float(scale), ptr(color),
             ptr(depth), stream())
    else:  # datasets.py:92-94: the colour image is resized to the depth image's size
        call("eslam_ingest_frame_resized", ptr(c), Hs, Ws, ptr(d), H, W, e, float(png_depth_scale), float(scale),
             ptr(color), ptr(depth), stream())
    return color, depth
