"""Golden fixture for the ScanNet-shaped frame ingest: a tiny ScanNet-layout dataset whose colour images
are about twice the depth images' size (as 1296x968 vs 640x480), loaded with the UNMODIFIED reference's `ScanNet`
dataset class (/root/reference/src/utils/datasets.py:79-167) on CPU -- cv2.resize of the float64 colour image to the
depth size, crop_edge -- pins oracle/eslam_oracle.py:ingest_frame_resized against it (bit for bit) and writes
tests/golden/ingest_scannet.npz.

    python tests/golden/make_golden_ingest_scannet.py
"""
import os
import sys
import tempfile
import types

import cv2
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import O, report, save  # noqa: E402  (also sets sys.path for the reference)


def main():
    from src.utils.datasets import ScanNet

    rng = np.random.default_rng(11)
    H, W, HS, WS, edge, pds = 48, 64, 97, 130, 2, 1000.0  # 130/64 = 2.03, 97/48 = 2.02: ScanNet's ratios
    with tempfile.TemporaryDirectory() as root:
        for sub in ("color", "depth", "pose"):
            os.makedirs(os.path.join(root, sub))
        for k in range(2):
            img = rng.integers(0, 256, size=(HS, WS, 3), dtype=np.uint8)
            img = cv2.GaussianBlur(img, (5, 5), 0)
            dep = (rng.random((H, W)) * 4.0 * pds).astype(np.uint16)
            dep[rng.random((H, W)) < 0.1] = 0
            cv2.imwrite(os.path.join(root, "color", f"{k}.jpg"), img)
            cv2.imwrite(os.path.join(root, "depth", f"{k}.png"), dep)
            with open(os.path.join(root, "pose", f"{k}.txt"), "w") as f:
                for row in np.eye(4):
                    f.write(" ".join(str(float(x)) for x in row) + "\n")
        cfg = {"dataset": "scannet", "cam": {"H": H, "W": W, "fx": 57.0, "fy": 57.0, "cx": 31.5, "cy": 23.5,
                                             "png_depth_scale": pds, "crop_edge": edge},
               "data": {"input_folder": root}}
        ds = ScanNet(cfg, types.SimpleNamespace(input_folder=None), 1, device="cpu")
        idx, color, depth, pose = ds[1]
        raw_c = cv2.imread(os.path.join(root, "color", "1.jpg"))
        raw_d = cv2.imread(os.path.join(root, "depth", "1.png"), cv2.IMREAD_UNCHANGED)
    assert color.dtype == torch.float64 and depth.dtype == torch.float32 and raw_d.dtype == np.uint16
    assert tuple(color.shape) == (H - 2 * edge, W - 2 * edge, 3)
    oc, od = O.ingest_frame_resized(raw_c, raw_d, pds, edge)
    report("scannet ingest colour vs oracle", oc, color, exact=True)
    report("scannet ingest depth vs oracle", od, depth, exact=True)
    save("ingest_scannet.npz", bgr=raw_c, depth_u16=raw_d, png_depth_scale=pds, crop_edge=edge, color=color, depth=depth)


if __name__ == "__main__":
    main()
