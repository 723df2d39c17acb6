"""Host-side profile (cProfile) of the MapperStep.optimize_mapping drop-in call as bench.py's e2e leg issues it."""
import cProfile
import os
import pstats
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import myslam_b200 as M  # noqa: E402
from bench import build_inputs  # noqa: E402
from myslam_b200 import synthetic as S  # noqa: E402


def main():
    dev = "cuda:0"
    spec = S.REPLICA_ROOM0
    m = spec["mapping"]
    nf = m["mapping_window_size"]
    scene = S.make_scene(spec, dev, seed=0)
    cfg = S.run_cfg(spec)

    class E:
        pass

    e = E()
    e.bound, e.device = scene.bound, dev
    e.H, e.W, e.fx, e.fy, e.cx, e.cy = scene.cam
    rnd = M.Renderer(cfg, e)
    poses, cols, deps = build_inputs(spec, dev, nf, seed=1)
    poses = poses.to(dev)
    mp = M.MapperStep(cfg, rnd, scene.decoders, scene.all_planes, scene.bound, scene.cam, dev)
    kf = [{"gt_c2w": poses[k], "idx": torch.tensor(4 * k), "color": cols[k], "depth": deps[k],
           "est_c2w": poses[k].clone()} for k in range(nf - 1)]
    mp.keyframe_dict = kf
    mp.joint_opt = True
    mp.mapping_window_size = nf
    kf_list = list(range(0, 4 * (nf - 1), 4))
    gc, gd, cw = cols[-1], deps[-1], poses[-1]

    def step():
        out = mp.optimize_mapping(m["iters"], 1.0, torch.tensor(4 * nf), gc, gd, cw, kf, kf_list, cw)
        out.cpu()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    import time
    t0 = time.perf_counter()
    for _ in range(30):
        step()
    print("ms per call (wall):", (time.perf_counter() - t0) / 30 * 1e3)
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(30):
        step()
    pr.disable()
    st = pstats.Stats(pr)
    st.sort_stats("cumulative").print_stats(35)


if __name__ == "__main__":
    main()
