"""Golden fixture for frame ingest: writes a tiny Replica-layout dataset to a temporary directory,
loads it with the UNMODIFIED reference's `Replica` dataset class (/root/reference/src/utils/datasets.py:79-138) on
CPU, pins oracle/eslam_oracle.py:ingest_frame against it and writes tests/golden/ingest.npz (the decoded uint8 /
uint16 arrays cv2.imread returns, and the loader's outputs).

    python tests/golden/make_golden_ingest.py
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
    from src.utils.datasets import Replica

    rng = np.random.default_rng(7)
    H, W, edge, pds = 40, 56, 3, 6553.5
    with tempfile.TemporaryDirectory() as root:
        os.makedirs(os.path.join(root, "results"))
        for k in range(2):
            img = rng.integers(0, 256, size=(H, W, 3), dtype=np.uint8)
            img = cv2.GaussianBlur(img, (5, 5), 0)  # jpeg-friendly
            dep = (rng.random((H, W)) * 4.0 * pds).astype(np.uint16)
            dep[rng.random((H, W)) < 0.05] = 0
            cv2.imwrite(os.path.join(root, "results", f"frame{k:06d}.jpg"), img)
            cv2.imwrite(os.path.join(root, "results", f"depth{k:06d}.png"), dep)
        with open(os.path.join(root, "traj.txt"), "w") as f:
            for k in range(2):
                f.write(" ".join(str(float(x)) for x in np.eye(4).reshape(-1)) + "\n")
        cfg = {"dataset": "replica", "cam": {"H": H, "W": W, "fx": 50.0, "fy": 50.0, "cx": 27.5, "cy": 19.5,
                                             "png_depth_scale": pds, "crop_edge": edge},
               "data": {"input_folder": root}}
        ds = Replica(cfg, types.SimpleNamespace(input_folder=None), 1, device="cpu")
        idx, color, depth, pose = ds[1]
        raw_c = cv2.imread(os.path.join(root, "results", "frame000001.jpg"))
        raw_d = cv2.imread(os.path.join(root, "results", "depth000001.png"), cv2.IMREAD_UNCHANGED)
    assert color.dtype == torch.float64 and depth.dtype == torch.float32 and raw_d.dtype == np.uint16
    oc, od = O.ingest_frame(raw_c, raw_d, pds, edge)
    report("ingest colour vs oracle", oc, color, exact=True)
    report("ingest depth vs oracle", od, depth, exact=True)
    save("ingest.npz", bgr=raw_c, depth_u16=raw_d, png_depth_scale=pds, crop_edge=edge, color=color, depth=depth)


if __name__ == "__main__":
    main()
