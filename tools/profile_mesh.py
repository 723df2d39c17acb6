"""A 16 M-point slab of the 1 cm marching-cubes lattice (BASELINE config 5) through the default (factored) lattice
query and the two bit-exact forms, for ncu.  Run plain first, then under ncu (B200_PROFILING.md)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from myslam_b200 import synthetic as S  # noqa: E402
from myslam_b200.mesher import grid_axes, query_grid_sdf  # noqa: E402


def main():
    dev = "cuda:0"
    spec = S.REPLICA_ROOM0
    scene = S.make_scene(spec, dev, seed=0)
    axes = grid_axes(spec["bound"], 0.01)
    kw = dict(start=100_000_000, count=1 << 24)
    for rep in range(2):
        query_grid_sdf(scene.all_planes, scene.decoders, axes, scene.bound, **kw)
        torch.cuda.synchronize()
    if "--all" in sys.argv:
        query_grid_sdf(scene.all_planes, scene.decoders, axes, scene.bound, separable=True, **kw)
        query_grid_sdf(scene.all_planes, scene.decoders, axes, scene.bound, separable=False, **kw)
        torch.cuda.synchronize()
    print("profile_mesh done")


if __name__ == "__main__":
    main()
