"""Golden fixture for full-image inference: runs the UNMODIFIED reference's
Renderer.render_img (/root/reference/src/utils/Renderer.py:155-204) on CPU in the build container, pins
oracle/eslam_oracle.py:render_image against it and writes tests/golden/img.npz (with the recorded random draws).

    python tests/golden/make_golden_img.py
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
from make_golden import (CAM, CFG_BOUND, C_PLANES_RES, PLANES_RES, TRUNC, O, Recorder, look_pose, ref_decoders,  # noqa: E402
                         ref_planes, ref_renderer, report, save, synth_frame)


def main():
    gen = torch.Generator().manual_seed(1234)  # same scene as make_golden.py (field.npz)
    bound = O.rounded_bound(CFG_BOUND, 0.24)
    fld = O.make_field(bound, PLANES_RES, C_PLANES_RES, generator=gen, std=0.05, dec_scale=2.0)
    g2 = torch.Generator().manual_seed(99)
    c2w = O.cam_pose_to_matrix(look_pose(g2, 0.02))[0]
    _, depth = synth_frame(g2, None, fld.bound, hole_frac=0.06)
    rnd = ref_renderer(fld)
    rnd.ray_batch_size = 1000  # 48 x 64 = 3072 rays -> 4 chunks, the last one ragged
    dec = ref_decoders(fld)
    torch.manual_seed(5)
    with Recorder() as rec:
        d_ref, c_ref = rnd.render_img(ref_planes(fld), dec, c2w, TRUNC, "cpu", gt_depth=depth)
    d_o, c_o = O.render_image(fld, CAM, c2w, depth, TRUNC, 32, 8, O.ReplayDraws(rec.log), ray_batch_size=1000)
    assert d_ref.dtype == torch.float64 and d_o.dtype == torch.float64
    report("render_img depth vs oracle", d_o, d_ref, tol=1e-6)
    report("render_img colour vs oracle", c_o, c_ref, tol=1e-6)
    arrays = dict(c2w=c2w, gt_depth=depth, depth=d_ref, color=c_ref, ray_batch_size=1000, n_draws=len(rec.log))
    for k, t in enumerate(rec.log):
        arrays[f"draw.{k}"] = t
    save("img.npz", **arrays)


if __name__ == "__main__":
    main()
