"""Golden fixture for keyframe selection: runs the UNMODIFIED reference's
Mapper.keyframe_selection_overlap (/root/reference/src/Mapper.py:146-209) on CPU in the build container, pins
oracle/eslam_oracle.py:keyframe_overlap against it and writes tests/golden/kfsel.npz.

    python tests/golden/make_golden_kfsel.py

`percent_inside` is captured by wrapping torch.nonzero; torch.randperm is replaced by the identity so the returned
list is nonzero(percent_inside) in order."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import O, Recorder, report, save  # noqa: E402  (also sets sys.path for the reference)


def scene(gen, K=14, H=120, W=160):
    cam = O.Camera(H=H, W=W, fx=130.0, fy=130.0, cx=79.5, cy=59.5)
    jj, ii = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
    depth = 1.6 + 0.5 * torch.sin(ii / 23.0) * torch.cos(jj / 17.0)
    depth[torch.rand(H, W, generator=gen) < 0.1] = 0.0
    color = torch.zeros(H, W, 3, dtype=torch.float64)  # not used by the selection (gathered and dropped)
    # current camera at the origin looking down -z (ESLAM convention); keyframes: small motions (overlap),
    # large rotations (no overlap), a camera behind the surface looking back, far translations
    def pose(q, t):
        return O.cam_pose_to_matrix(torch.tensor([list(q) + list(t)], dtype=torch.float32))[0]

    cur = pose((1.0, 0.01, -0.02, 0.005), (0.1, -0.05, 0.0))
    kfs = []
    for k in range(K):
        r = torch.rand(7, generator=gen) * 2 - 1
        if k % 4 == 0:    # near duplicate of the current view
            kfs.append(pose((1.0, 0.05 * r[0], 0.05 * r[1], 0.05 * r[2]), (0.2 * r[3], 0.2 * r[4], 0.2 * r[5])))
        elif k % 4 == 1:  # rotated away by a large angle about y
            kfs.append(pose((0.3 * r[0], 0.1 * r[1], 1.0, 0.1 * r[2]), (r[3], r[4], r[5])))
        elif k % 4 == 2:  # moderate rotation + translation: partial overlap
            kfs.append(pose((1.0, 0.25 * r[0], 0.25 * r[1], 0.1 * r[2]), (0.8 * r[3], 0.5 * r[4], 0.4 * r[5])))
        else:             # far away
            kfs.append(pose((1.0, 0.1 * r[0], 0.1 * r[1], 0.1 * r[2]), (6 * r[3], 6 * r[4], -3.0)))
    return cam, cur, depth, color, torch.stack(kfs, 0)


def main():
    from src.Mapper import Mapper

    gen = torch.Generator().manual_seed(4321)
    cam, cur, depth, color, kfs = scene(gen)
    K = kfs.shape[0]
    mp = object.__new__(Mapper)
    mp.device = "cpu"
    mp.H, mp.W, mp.fx, mp.fy, mp.cx, mp.cy = cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy
    mp.keyframe_list = list(range(K))
    mp.estimate_c2w_list = kfs.clone()
    captured = []
    _nonzero, _randperm = torch.nonzero, torch.randperm
    torch.nonzero = lambda x, *a, **k: (captured.append(x.clone()), _nonzero(x, *a, **k))[1]
    torch.randperm = lambda n, *a, **k: torch.arange(n)
    torch.manual_seed(17)
    try:
        with Recorder() as rec:
            sel = mp.keyframe_selection_overlap(color, depth, cur, K)
    finally:
        torch.nonzero, torch.randperm = _nonzero, _randperm
    percent = captured[0]
    frac, cnt, n_pts = O.keyframe_overlap(cam, cur, depth, color, kfs[:-2], O.ReplayDraws(rec.log))
    report("percent_inside vs oracle", frac, percent, exact=True)
    assert [int(s) for s in sel] == torch.nonzero(frac).squeeze(-1).tolist()
    assert 0 < len(sel) < K - 2, "fixture should have both overlapping and non-overlapping keyframes"
    print("  selected", [int(s) for s in sel], "counts", cnt.tolist(), "of", n_pts)
    save("kfsel.npz", H=cam.H, W=cam.W, fx=cam.fx, fy=cam.fy, cx=cam.cx, cy=cam.cy, cur_c2w=cur, depth=depth,
         color=color, kf_c2ws=kfs, idx=rec.log[0], percent_inside=percent, counts=cnt, n_pts=n_pts,
         selected=np.array([int(s) for s in sel]))


if __name__ == "__main__":
    main()
