"""CUDA-event time of the marching-cubes SDF query over the full 1 cm Replica room0 lattice (990 x 680 x 490) in its
default form (whole rows: eslam_grid_sdf_rows) and with ESLAM_B200_GRID_ROWS=0 (per-voxel kernel only).

    python tools/mesh_times.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from bench import time_region  # noqa: E402
from myslam_b200 import synthetic as S  # noqa: E402
from myslam_b200.mesher import grid_axes, query_grid_sdf  # noqa: E402


def main():
    dev = "cuda:0"
    spec = S.REPLICA_ROOM0
    scene = S.make_scene(spec, dev, seed=0)
    axes = grid_axes(spec["bound"], 0.01)
    total = len(axes[0]) * len(axes[1]) * len(axes[2])
    buf = torch.empty(total, dtype=torch.float32, device=dev)
    run = lambda: query_grid_sdf(scene.all_planes, scene.decoders, axes, scene.bound, out=buf)
    for name, env in (("whole rows (default)", "1"), ("per voxel", "0")):
        os.environ["ESLAM_B200_GRID_ROWS"] = env
        ms = time_region(run, 5, 2, False) / 5
        print(f"{name:24s} {ms:7.3f} ms   {total / ms / 1e6:7.1f} G points/s", flush=True)
    del os.environ["ESLAM_B200_GRID_ROWS"]


if __name__ == "__main__":
    main()
