"""Per-kernel CUDA-event timings of the multi-GPU exchange (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tools/exchange_times.py

counters exchange, fused reduce-scatter+Adam+all-gather (multimem and P2P), NCCL all-reduce + Adam, and the fused
backward with its gradient arena in ordinary vs symmetric memory."""
import ctypes as C
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import myslam_b200 as M  # noqa: E402
from bench import build_inputs, time_region  # noqa: E402
from myslam_b200 import synthetic as S  # noqa: E402
from myslam_b200._lib import call, ptr, stream  # noqa: E402
from myslam_b200.common import matrix_to_cam_pose  # noqa: E402
from myslam_b200.decoders import synced_store  # noqa: E402
from myslam_b200.dist import MappingExchange, PeerExchange  # noqa: E402
from myslam_b200.hotpath import mapping_iteration  # noqa: E402
from myslam_b200.mapper import _mapper_state  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = f"cuda:{local}"
    dist.init_process_group("nccl", device_id=torch.device(dev))
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
    N = pix * nf
    torch.manual_seed(77 + rank)
    store.reset_adam()
    poses7 = torch.zeros(nf, 7, device=dev)
    poses7[1:] = matrix_to_cam_pose(poses[1:])
    res = {}

    def bwd():
        call("eslam_loss_backward_q", store.ref(), ptr(store.arena), ptr(store.ensure_q()), ptr(store.ensure_q_grad()),
             C.byref(sc.cam), C.byref(sc.render), ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.z), ptr(ws.gt_depth),
             ptr(ws.gt_color), ptr(ws.src), ptr(idx), pix, None, ptr(ws.counters), None, N, ptr(store.ensure_grad()),
             ptr(ws.pose_grad), None, stream())

    # every timed region = the backward kernel (so the gradient images carry a real iteration's content) + one way of
    # turning them into the optimiser step; subtract the first line
    mapping_iteration(ws, store, sc, poses, poses7, cols, deps, pix, 1, 1e-3, 5e-3, 5e-3, 1e-3, apply_adam=False)
    store.reset_adam()
    idx = torch.randint(spec["H"] * spec["W"], (N,), device=dev)
    res["bwd alone"] = time_region(bwd, 50, 5, True) / 50
    store.reset_adam()

    def local_step():
        bwd()
        store.adam_step_q(1, 1e-3, 5e-3, 5e-3)

    res["bwd + 1-GPU tail (no exchange: replicas would diverge)"] = time_region(local_step, 50, 5, True) / 50
    nccl = MappingExchange()

    def nccl_step():
        bwd()
        nccl.reduce_grads([store.gq_arena, store.grad[store.dec_off:]], ws.pose_grad, None)
        store.adam_step_q(1, 1e-3, 5e-3, 5e-3)
        nccl.after_step(store)

    res["bwd + NCCL all-reduce(images, decoders, poses) + tail + decoder broadcast"] = time_region(nccl_step, 50, 5, True) / 50
    res["NCCL all-reduce of 8 counters"] = time_region(lambda: nccl.reduce_counters(ws.counters), 50, 5, True) / 50
    for mm in (False, True):
        ex = PeerExchange(store, ws, multimem=mm)
        if mm and not ex.multimem:
            continue
        tag = "multimem" if ex.multimem else "P2P"
        store.reset_adam()
        res[f"bwd alone, parameters in symmetric memory ({tag})"] = time_region(bwd, 50, 5, True) / 50
        store.reset_adam()

        def peer_step():
            bwd()
            ex.adam_exchange(1, 1e-3, 5e-3, 5e-3, ws.pose_grad, nf, None)

        res[f"peer counters exchange ({tag})"] = time_region(lambda: ex.reduce_counters(ws.counters), 50, 5, True) / 50
        res[f"bwd + peer exchange: push + tile Adam + all-gather + decoder step ({tag})"] = time_region(peer_step, 50, 5, True) / 50
        lib = M._lib.load()
        for bits, what in ((8, "barriers only"), (32, "no barriers"), (8 | 32, "empty launches")):
            lib.eslam_set_debug(bits)
            res[f"  [{what}] ({tag})"] = time_region(peer_step, 50, 5, True) / 50
        lib.eslam_set_debug(0)
        store.reset_adam()
        ex.check()
    if rank == 0:
        print(f"# world {world}, arena {store.n_floats * 4 / 1e6:.1f} MB")
        for k, v in res.items():
            print(f"{k:60s} {1e3 * v:9.2f} us")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
