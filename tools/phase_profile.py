"""Per-phase wall clocks of the fused backward kernel (profiling build, -DESLAM_PROFILE_PHASES):

    nvcc ... -DESLAM_PROFILE_PHASES -o build_exp/libphase.so myslam_b200/csrc/eslam_b200.cu
    ESLAM_B200_LIB=build_exp/libphase.so python tools/phase_profile.py

Prints, for the sdf half (warp 0) and the rgb half (warp 4), the share of a CTA's lifetime spent in each phase."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import myslam_b200 as M  # noqa: E402
from bench import build_inputs  # noqa: E402
from myslam_b200 import synthetic as S  # noqa: E402
from myslam_b200._lib import load  # noqa: E402
from myslam_b200.decoders import synced_store  # noqa: E402
from myslam_b200.mapper import _mapper_state, map_window  # noqa: E402

NAMES = ["P0 points+axes+weights", "P2 gather (+J)", "  barrier", "P3 MLP layers 2-3", "  barrier", "P4/P5 composite+loss+grad",
         "P6 MLP backward (hidden)", "P6 J^T g + weight gradients", "P6 beta", "P7 reductions",
         "P8 ray/pose gradients"]


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
    st = _mapper_state(mp, m["pixels"], nf)
    store = synced_store(scene.all_planes, scene.decoders, scene.bound)
    lr = m["lr"]
    lib = load()
    lib.eslam_set_debug(int(os.environ.get("ESLAM_B200_DEBUG", "0")))
    buf = (C.c_ulonglong * 32)()
    run = lambda: map_window(store, st["ws"], st["sc"], poses, cols, deps, m["pixels"], m["iters"], lr["decoders_lr"],
                             lr["planes_lr"], lr["c_planes_lr"], True, m["joint_opt_cam_lr"])
    run()
    lib.eslam_phase_counters(buf)
    run()
    lib.eslam_phase_counters(buf)
    for half, tag in ((0, "sdf half"), (1, "rgb half")):
        v = [buf[half * 16 + i] for i in range(11)]
        tot = sum(v)
        print(f"# {tag}: {tot / (m['iters'] * 1334) / 1.965e3:.1f} us per CTA (clock64 / 1.965 GHz)")
        for n, x in zip(NAMES, v):
            print(f"  {n:34s} {100 * x / tot:5.1f} %   {x / (m['iters'] * 1334) / 1.965e3:6.2f} us")


if __name__ == "__main__":
    main()
