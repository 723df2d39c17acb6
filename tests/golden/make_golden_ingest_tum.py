"""Golden fixture for the TUM-shaped frame ingest: a tiny TUM-RGBD-layout dataset (rgb.txt, depth.txt,
groundtruth.txt; colour and depth of one size) loaded with the UNMODIFIED reference's `TUM_RGBD` dataset class
(/root/reference/src/utils/datasets.py:79-112,168-253) on CPU with the freiburg1 settings that matter -- cv2.undistort
of the colour image, crop_size (bilinear align_corners resize of the colour, nearest of the depth), crop_edge -- pins
oracle/eslam_oracle.py:ingest_frame_tum against it (bit for bit) and writes tests/golden/ingest_tum.npz.

    python tests/golden/make_golden_ingest_tum.py
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
    if not hasattr(np, "unicode_"):
        np.unicode_ = np.str_  # the reference's parse_list names the numpy < 2 alias (datasets.py:180)
    from src.utils.datasets import TUM_RGBD

    rng = np.random.default_rng(5)
    H, W, edge, pds = 60, 80, 2, 5000.0
    crop = [48, 64]  # 60/48 = 80/64 = 1.25, as 480/384 = 640/512
    cam = (64.7, 64.6, 39.8, 31.9)  # freiburg1's intrinsics scaled by 1/8
    dist = [0.2624, -0.9531, -0.0054, 0.0026, 1.1633]
    with tempfile.TemporaryDirectory() as root:
        for sub in ("rgb", "depth"):
            os.makedirs(os.path.join(root, sub))
        with open(os.path.join(root, "rgb.txt"), "w") as fr, open(os.path.join(root, "depth.txt"), "w") as fd, \
                open(os.path.join(root, "groundtruth.txt"), "w") as fg:
            fg.write("# timestamp tx ty tz qx qy qz qw\n")
            for k in range(2):
                img = cv2.GaussianBlur(rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8), (5, 5), 0)
                dep = (rng.random((H, W)) * 4.0 * pds).astype(np.uint16)
                dep[rng.random((H, W)) < 0.1] = 0
                cv2.imwrite(os.path.join(root, "rgb", f"{k}.png"), img)
                cv2.imwrite(os.path.join(root, "depth", f"{k}.png"), dep)
                fr.write(f"{k}.0 rgb/{k}.png\n")
                fd.write(f"{k}.0 depth/{k}.png\n")
                fg.write(f"{k}.0 0 0 0 0 0 0 1\n")
        cfg = {"dataset": "tumrgbd", "cam": {"H": H, "W": W, "fx": cam[0], "fy": cam[1], "cx": cam[2], "cy": cam[3],
                                              "png_depth_scale": pds, "crop_edge": edge, "crop_size": crop,
                                              "distortion": dist},
               "data": {"input_folder": root}}
        ds = TUM_RGBD(cfg, types.SimpleNamespace(input_folder=None), 1, device="cpu")
        idx, color, depth, pose = ds[1]
        raw_c = cv2.imread(os.path.join(root, "rgb", "1.png"))
        raw_d = cv2.imread(os.path.join(root, "depth", "1.png"), cv2.IMREAD_UNCHANGED)
    assert color.dtype == torch.float64 and depth.dtype == torch.float32 and raw_d.dtype == np.uint16
    assert tuple(color.shape) == (crop[0] - 2 * edge, crop[1] - 2 * edge, 3)
    und = cv2.undistort(raw_c, np.array([[cam[0], 0, cam[2]], [0, cam[1], cam[3]], [0, 0, 1.0]]), np.array(dist))
    assert np.array_equal(O.undistort_u8(raw_c, *cam, dist), und), "undistort restatement differs from cv2"
    assert int((und != raw_c).sum()) > 0
    oc, od = O.ingest_frame_tum(raw_c, raw_d, pds, cam, dist, crop, edge)
    report("tum ingest colour vs oracle", oc, color, exact=True)
    report("tum ingest depth vs oracle", od, depth, exact=True)
    save("ingest_tum.npz", bgr=raw_c, depth_u16=raw_d, png_depth_scale=pds, crop_edge=edge, crop_size=np.array(crop),
         cam=np.array(cam), distortion=np.array(dist), undistorted=und, color=color, depth=depth)


if __name__ == "__main__":
    main()
