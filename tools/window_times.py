"""CUDA-event time of one optimize_mapping-shaped call (15 iterations x 4000 rays, 20-frame window) under the launch
variants of the window loop: CUDA graph on/off (ESLAM_B200_GRAPH), split backward on/off (ESLAM_B200_SPLIT_BWD),
two-stream pipelining on/off (ESLAM_B200_PIPELINE).  Prints us per iteration for each.

    python tools/window_times.py [n_calls]
"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import myslam_b200 as M  # noqa: E402
from bench import build_inputs, time_region  # noqa: E402
from myslam_b200 import synthetic as S  # noqa: E402
from myslam_b200.decoders import synced_store  # noqa: E402
from myslam_b200.hotpath import FrameTable  # noqa: E402
from myslam_b200.mapper import _mapper_state, map_window  # noqa: E402


def main():
    n_calls = int(sys.argv[1]) if len(sys.argv) > 1 else 20
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
    arena0 = store.arena.clone()
    lr = m["lr"]
    window = FrameTable([cols[k] for k in range(nf)], [deps[k] for k in range(nf)], st["sc"].cam, dev)
    run = lambda: map_window(store, st["ws"], st["sc"], poses, window, window, m["pixels"], m["iters"],
                             lr["decoders_lr"], lr["planes_lr"], lr["c_planes_lr"], True, m["joint_opt_cam_lr"])
    variants = [("graph, split backward", {"ESLAM_B200_GRAPH": "1", "ESLAM_B200_SPLIT_BWD": "1"}),
                ("graph, one backward launch", {"ESLAM_B200_GRAPH": "1", "ESLAM_B200_SPLIT_BWD": "0"}),
                ("kernel by kernel, split backward", {"ESLAM_B200_GRAPH": "0", "ESLAM_B200_SPLIT_BWD": "1"}),
                ("kernel by kernel, one backward launch", {"ESLAM_B200_GRAPH": "0", "ESLAM_B200_SPLIT_BWD": "0"}),
                ("kernel by kernel, one stream", {"ESLAM_B200_GRAPH": "0", "ESLAM_B200_PIPELINE": "0"})]
    for name, env in variants:
        for k in ("ESLAM_B200_GRAPH", "ESLAM_B200_SPLIT_BWD", "ESLAM_B200_PIPELINE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        st["ws"].__dict__.pop("_window_graphs", None)  # a graph captured under another variant must not be replayed
        store.arena.copy_(arena0)
        store.gen += 1
        torch.manual_seed(7)
        ms = time_region(run, n_calls, 3, False) / n_calls
        torch.cuda.synchronize()
        t0 = time.perf_counter()  # host time to ENQUEUE a call (no sync inside): what a launch-bound loop is limited by
        for _ in range(n_calls):
            run()
        host = (time.perf_counter() - t0) / n_calls * 1e3
        torch.cuda.synchronize()
        print(f"{name:40s} {ms:7.3f} ms per call   {1e3 * ms / m['iters']:7.1f} us per iteration   "
              f"host enqueue {host:6.3f} ms per call", flush=True)


if __name__ == "__main__":
    main()
