"""CUDA-event timings of each kernel of the hot path at Replica room0 shapes (4000 mapping rays over a
20-frame window; 2000 tracking rays), launched alone back to back after warm-up."""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import myslam_b200 as M  # noqa: E402
from bench import build_inputs, time_region  # noqa: E402
from myslam_b200 import synthetic as S  # noqa: E402
from myslam_b200._lib import call, ptr, stream  # noqa: E402
from myslam_b200.common import matrix_to_cam_pose  # noqa: E402
from myslam_b200.decoders import synced_store  # noqa: E402
from myslam_b200.hotpath import mapping_iteration, tracking_iteration, _sample  # noqa: E402
from myslam_b200.mapper import _mapper_state  # noqa: E402
from myslam_b200.renderer import linspace_table  # noqa: E402
from myslam_b200.tracker import _tracker_state, _tracker_store  # noqa: E402


def main():
    dev = "cuda:0"
    spec = S.REPLICA_ROOM0
    m, t = spec["mapping"], spec["tracking"]
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
    store.reset_adam()
    poses7 = torch.zeros(nf, 7, device=dev)
    poses7[1:] = matrix_to_cam_pose(poses[1:])
    mapping_iteration(ws, store, sc, poses, poses7, cols, deps, pix, 1, 1e-3, 5e-3, 5e-3, 1e-3, apply_adam=False)
    idx = torch.randint(spec["H"] * spec["W"], (N,), device=dev)
    u = torch.rand(N, 40, device=dev)
    uc, uf = torch.rand(N, 32, device=dev), torch.rand(N, 8, device=dev)
    c2w = poses.reshape(nf, 16).contiguous()
    store.bind()
    res = {}

    def bwd(grad, pose):
        return lambda: call("eslam_loss_backward", store.ref(), ptr(store.arena), C.byref(sc.cam), C.byref(sc.render),
                            ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.z), ptr(ws.gt_depth), ptr(ws.gt_color), ptr(ws.src),
                            ptr(idx), pix, None, ptr(ws.counters), None, N, ptr(store.grad) if grad else None,
                            ptr(ws.pose_grad) if pose else None, None, stream())

    def timeit(name, fn, n=30):
        res[name] = 1e3 * time_region(fn, n, 5, False) / n

    timeit("map.loss_backward planes+poses", bwd(True, True))
    timeit("map.loss_backward planes only", bwd(True, False))
    timeit("map.loss_backward poses only", bwd(False, True))
    from myslam_b200 import _lib
    for flags, name in ((1, "no plane reductions"), (2, "no weight grads"), (3, "no reductions, no weight grads"),
                        (4, "no MLP arithmetic"), (6, "no MLP, no weight grads"), (7, "gather+scan+corner loads only")):
        _lib.load().eslam_set_debug(flags)
        timeit(f"map.loss_backward planes+poses [{name}]", bwd(True, True))
    _lib.load().eslam_set_debug(0)
    # the Q form (the default path): backward into gradient images, optimiser tail, Q rebuild
    if True:
        mq = torch.zeros(store.n_planes_end // 2, dtype=torch.float32, device=dev)
        mgq = torch.zeros_like(mq)
        tq = torch.zeros(_lib.load().eslam_q_touched_bytes(store.ref()), dtype=torch.uint8, device=dev)
        call("eslam_q_build", store.ref(), ptr(store.arena), ptr(mq), stream())
        timeit("map.loss_backward_q planes+poses", lambda: call(
            "eslam_loss_backward_q", store.ref(), ptr(store.arena), ptr(mq), ptr(mgq), C.byref(sc.cam),
            C.byref(sc.render), ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.z), ptr(ws.gt_depth), ptr(ws.gt_color),
            ptr(ws.src), ptr(idx), pix, None, ptr(ws.counters), None, N, ptr(store.grad), ptr(ws.pose_grad), None,
            stream()))
        timeit("map.q_adam_planes (tail)", lambda: call(
            "eslam_q_adam_planes", store.ref(), ptr(store.arena), ptr(mgq), ptr(store.exp_avg), ptr(store.exp_avg_sq),
            ptr(store.grad), ptr(tq), 5e-3, 5e-3, 1, 0.9, 0.999, 1e-8, stream()))
        timeit("map.q_build", lambda: call("eslam_q_build", store.ref(), ptr(store.arena), ptr(mq), stream()))
        store.grad.zero_()
    timeit("map.sample_rays", lambda: _sample(ws, store, sc, idx, nf, pix, c2w, poses7, 1, deps, cols, u, 0))
    timeit("map.importance", lambda: call(
        "eslam_importance_samples", store.ref(), ptr(store.arena), ptr(store.ensure_q()), C.byref(sc.render),
        ptr(ws.rays_o), ptr(ws.rays_d),
        ptr(ws.dl_list), ptr(ws.counters), N, ptr(uc), ptr(uf), ptr(linspace_table(32, dev)), ptr(ws.z), stream()))
    timeit("map.adam (6.79M params)", lambda: store.adam_step(1, 1e-3, 5e-3, 5e-3))
    timeit("map.render_forward 4000 rays", lambda: call(
        "eslam_render_forward", store.ref(), ptr(store.arena), ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.z), N, 40,
        ptr(ws.counters), ptr(ws.depth), ptr(ws.rgb), ptr(ws.sdf), stream()))
    timeit("torch.randint+3 rand (draws)", lambda: (torch.randint(816000, (N,), device=dev), torch.rand(N, 40, device=dev),
                                                    torch.rand(N, 32, device=dev), torch.rand(N, 8, device=dev)))
    timeit("bind_decoders", lambda: store.bind(force=True))
    res["map.R"] = int(ws.counters[0])
    res["map.R0"] = int(ws.counters[1])
    # tracking
    trk = M.TrackerStep(cfg, rnd, scene.decoders, scene.all_planes, scene.bound, scene.cam, dev)
    tst = _tracker_state(trk, t["pixels"])
    tstore = _tracker_store(trk, tst)
    tws, tsc = tst["ws"], tst["sc"]
    pose0 = matrix_to_cam_pose(poses[:1]).contiguous()
    col1, dep1 = cols[:1].contiguous(), deps[:1].contiguous()
    tracking_iteration(tws, tstore, tsc, pose0, col1, dep1, t["pixels"])
    tidx = torch.randint(530 * 1050, (2000,), device=dev)
    tu = torch.rand(2000, 40, device=dev)
    timeit("trk.sample_rays", lambda: _sample(tws, tstore, tsc, tidx, 1, 2000, None, pose0, 0, dep1, col1, tu, 1))
    timeit("trk.render_forward", lambda: call(
        "eslam_render_forward", tstore.ref(), ptr(tstore.arena), ptr(tws.rays_o), ptr(tws.rays_d), ptr(tws.z), 2000, 40,
        ptr(tws.counters), ptr(tws.depth), ptr(tws.rgb), ptr(tws.sdf), stream()))
    # experimental: the same forward on pre-activated plane images, and the image rebuild
    q_arena = torch.zeros(tstore.n_planes_end // 2, dtype=torch.float32, device=dev)
    timeit("q_build (12 planes)", lambda: call("eslam_q_build", tstore.ref(), ptr(tstore.arena), ptr(q_arena), stream()))
    timeit("trk.render_forward_q", lambda: call(
        "eslam_render_forward_q", tstore.ref(), ptr(q_arena), ptr(tws.rays_o), ptr(tws.rays_d), ptr(tws.z), 2000, 40,
        ptr(tws.counters), ptr(tws.depth), ptr(tws.rgb), ptr(tws.sdf), None, None, stream()))
    timeit("map.render_forward_q 4000 rays", lambda: call(
        "eslam_render_forward_q", store.ref(), ptr(q_arena), ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.z), N, 40,
        ptr(ws.counters), ptr(ws.depth), ptr(ws.rgb), ptr(ws.sdf), None, None, stream()))
    timeit("trk.track_mask", lambda: call("eslam_track_mask", ptr(tws.gt_depth), ptr(tws.depth), ptr(tws.band), 2000,
                                          ptr(tws.counters), ptr(tws.ray_mask), ptr(tws.scratch), stream()))
    timeit("trk.loss_backward poses", lambda: call(
        "eslam_loss_backward", tstore.ref(), ptr(tstore.arena), C.byref(tsc.cam), C.byref(tsc.render), ptr(tws.rays_o),
        ptr(tws.rays_d), ptr(tws.z), ptr(tws.gt_depth), ptr(tws.gt_color), ptr(tws.src), ptr(tidx), 2000,
        ptr(tws.ray_mask), ptr(tws.counters), None, 2000, None, ptr(tws.pose_grad), None, stream()))
    # the tracker's cached pose-only backward on the activations the forward left, parameter form and Q form
    call("eslam_render_forward_act", tstore.ref(), ptr(tstore.arena), ptr(tws.rays_o), ptr(tws.rays_d), ptr(tws.z), 2000,
         40, ptr(tws.counters), ptr(tws.depth), ptr(tws.rgb), ptr(tws.sdf), ptr(tws.act4), ptr(tws.actm), stream())
    timeit("trk.pose_backward_act (cached)", lambda: call(
        "eslam_pose_backward_act", tstore.ref(), ptr(tstore.arena), C.byref(tsc.cam), C.byref(tsc.render),
        ptr(tws.rays_o), ptr(tws.rays_d), ptr(tws.z), ptr(tws.gt_depth), ptr(tws.gt_color), ptr(tws.src), ptr(tidx), 2000,
        ptr(tws.ray_mask), ptr(tws.counters), 2000, ptr(tws.sdf), ptr(tws.act4), ptr(tws.actm), ptr(tws.pose_grad), None,
        stream()))
    call("eslam_render_forward_q", tstore.ref(), ptr(q_arena), ptr(tws.rays_o), ptr(tws.rays_d), ptr(tws.z), 2000, 40,
         ptr(tws.counters), ptr(tws.depth), ptr(tws.rgb), ptr(tws.sdf), ptr(tws.act4), ptr(tws.actm), stream())
    timeit("trk.pose_backward_q (cached)", lambda: call(
        "eslam_pose_backward_q", tstore.ref(), ptr(tstore.arena), ptr(q_arena), C.byref(tsc.cam), C.byref(tsc.render),
        ptr(tws.rays_o), ptr(tws.rays_d), ptr(tws.z), ptr(tws.gt_depth), ptr(tws.gt_color), ptr(tws.src), ptr(tidx), 2000,
        ptr(tws.ray_mask), ptr(tws.counters), 2000, ptr(tws.sdf), ptr(tws.act4), ptr(tws.actm), ptr(tws.pose_grad), None,
        stream()))
    timeit("trk.iteration (all launches)", lambda: tracking_iteration(tws, tstore, tsc, pose0, col1, dep1, 2000), 20)
    timeit("map.iteration (all launches)", lambda: mapping_iteration(ws, store, sc, poses, poses7, cols, deps, pix, 1,
                                                                    1e-3, 5e-3, 5e-3, 1e-3), 20)
    res["trk.R"] = int(tws.counters[0])
    for k, v in res.items():
        print(f"{k:40s} {v:10.2f}" + (" us" if isinstance(v, float) else ""))


if __name__ == "__main__":
    main()
