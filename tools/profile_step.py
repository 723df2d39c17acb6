"""Short, fixed launch sequence for ncu: one warm-up + one timed `optimize_mapping`-shaped call
(15 iterations x 4000 rays, 20-frame window) and one warm-up + one tracked frame (8 x 2000 rays)
at Replica room0 shapes.  Run plain first, then under ncu (B200_PROFILING.md)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import myslam_b200 as M  # noqa: E402
from bench import build_inputs  # noqa: E402
from myslam_b200 import synthetic as S  # noqa: E402
from myslam_b200.common import matrix_to_cam_pose  # noqa: E402
from myslam_b200.decoders import synced_store  # noqa: E402
from myslam_b200.hotpath import mapping_iteration  # noqa: E402
from myslam_b200.mapper import _mapper_state  # noqa: E402


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
    ws, sc = st["ws"], st["sc"]
    pix = m["pixels"] // nf
    lr = m["lr"]
    for rep in range(2):
        store.reset_adam()
        poses7 = torch.zeros(nf, 7, device=dev)
        poses7[1:] = matrix_to_cam_pose(poses[1:])
        for it in range(m["iters"]):
            mapping_iteration(ws, store, sc, poses, poses7, cols, deps, pix, it + 1, lr["decoders_lr"], lr["planes_lr"],
                              lr["c_planes_lr"], m["joint_opt_cam_lr"])
        torch.cuda.synchronize()
    trk = M.TrackerStep(cfg, rnd, scene.decoders, scene.all_planes, scene.bound, scene.cam, dev)
    pose0 = matrix_to_cam_pose(poses[:1])
    for rep in range(2):
        trk.track_frame(pose0, cols[:1].contiguous(), deps[:1].contiguous())
        torch.cuda.synchronize()
    # a 16 M-point slab of the 1 cm marching-cubes lattice (config 5) for the grid-query kernel
    from myslam_b200.mesher import grid_axes, query_grid_sdf

    axes = grid_axes(spec["bound"], 0.01)
    for rep in range(2):
        query_grid_sdf(scene.all_planes, scene.decoders, axes, scene.bound, start=100_000_000, count=1 << 24)
        torch.cuda.synchronize()
    print("profile_step done")


if __name__ == "__main__":
    main()
