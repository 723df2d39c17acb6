"""The drop-ins under the reference's PROCESS structure (ESLAM.py:246-260): a tracker process and a mapper process that
share the 12 feature planes and the decoders through CUDA IPC (what torch.multiprocessing does with the tensors ESLAM
owns) and shake hands through shared counters exactly as Tracker.run / Mapper.run do:

  tracker (Tracker.py:254-313)   for every frame: when idx % every_frame == 1 wait until mapping_idx == idx - 1;
                                 update_params_from_mapping (decoders' state_dict, planes re-aliased); 8 x
                                 optimize_tracking with the caller's torch.optim.Adam; estimate_c2w_list[idx]; idx[0] = idx
  mapper  (Mapper.py:387-460)    wait until idx % every_frame == 0 and new; optimize_mapping (15 iterations, up to 20
                                 keyframes, joint pose optimisation after 5 keyframes); keyframe bookkeeping;
                                 mapping_first_frame, mapping_idx

Each process has its OWN parameter arena (FieldStore): the mapper writes the shared [1,32,H,W] tensors back at the end
of a call (myslam_b200.mapper.optimize_mapping), the tracker re-imports them when prev_mapping_idx changes
(myslam_b200.tracker._tracker_store) -- the path this tool exists to exercise, under the real race between the two.
Synthetic Replica-shaped frames are generated in both processes from the same seed (no dataset on the GPU box).

    python tools/e2e_two_process.py --frames 120 [--scale 0.5] [--out gpurun_out/e2e_two_process.json]
"""
import argparse
import copy
import json
import os
import sys
import time

import numpy as np
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tools")):
    if p not in sys.path:
        sys.path.insert(0, p)


def make_spec(scale):
    from myslam_b200 import synthetic as S

    spec = dict(S.REPLICA_ROOM0)
    if scale != 1.0:
        s = scale
        spec.update(H=int(680 * s), W=int(1200 * s), fx=600.0 * s, fy=600.0 * s, cx=(1200 * s - 1) / 2, cy=(680 * s - 1) / 2)
        spec["tracking"] = dict(spec["tracking"], ignore_edge_H=int(75 * s), ignore_edge_W=int(75 * s))
    return spec


def make_frames(spec, n, dev, seed):
    from myslam_b200 import synthetic as S

    gen = torch.Generator().manual_seed(seed)
    gt = S.trajectory(n, spec["room"], step_deg=0.5).to(dev)
    cam = (spec["H"], spec["W"], spec["fx"], spec["fy"], spec["cx"], spec["cy"])
    return gt, [S.render_box_room(gt[k], *cam, spec["room"], dev, hole_frac=0.02, generator=gen) for k in range(n)]


def holder(cfg, spec, planes, dec, bound, dev, cls):
    import myslam_b200 as M

    cam = (spec["H"], spec["W"], spec["fx"], spec["fy"], spec["cx"], spec["cy"])

    class E:
        pass

    e = E()
    e.bound, e.device = bound.clone(), dev
    e.H, e.W, e.fx, e.fy, e.cx, e.cy = cam
    rnd = M.Renderer(cfg, e)
    return cls(cfg, rnd, dec, planes, bound.clone(), cam, dev)


def tracker_proc(args, shared):
    import myslam_b200 as M
    from myslam_b200 import synthetic as S
    from myslam_b200.common import cam_pose_to_matrix, matrix_to_cam_pose

    dev = "cuda:0"
    torch.cuda.set_device(0)
    spec = make_spec(args.scale)
    cfg = S.run_cfg(spec)
    gt, frames = make_frames(spec, args.frames, dev, args.seed)
    planes = tuple([p for p in g] for g in shared["planes"])  # aliases of the mapper's storage
    dec = copy.deepcopy(shared["decoders"])                   # the tracker's own decoders (Tracker.py:101)
    dec.bound = shared["bound"].clone()
    trk = holder(cfg, spec, planes, dec, shared["bound"], dev, M.TrackerStep)
    t = cfg["tracking"]
    every = 4
    torch.manual_seed(args.seed)
    est, idx_s, map_idx, first = shared["est"], shared["idx"], shared["mapping_idx"], shared["first"]
    prev_mapping_idx, pre_c2w, n_updates = -1, None, 0
    while first[0] == 0:  # ESLAM.tracking waits for the first frame's mapping
        time.sleep(0.001)
    t0 = time.perf_counter()
    for idx in range(args.frames):
        if idx > 0 and (idx % every == 1 or every == 1):
            while map_idx[0] != idx - 1:
                time.sleep(0.001)
            pre_c2w = est[idx - 1].to(dev)[None]
        if int(map_idx[0]) != prev_mapping_idx:  # update_params_from_mapping
            dec.load_state_dict(shared["decoders"].state_dict())
            prev_mapping_idx = int(map_idx[0])
            trk.prev_mapping_idx = prev_mapping_idx
            n_updates += 1
        col, dep = frames[idx]
        if idx == 0:
            c2w = gt[0][None]
        else:
            if idx >= 2:
                pre = matrix_to_cam_pose(torch.stack([est[idx - 2].to(dev), pre_c2w[0]], 0))
                cam_pose = 2 * pre[1:] - pre[0:1]
            else:
                cam_pose = matrix_to_cam_pose(pre_c2w)
            T = torch.nn.Parameter(cam_pose[:, -3:].clone())
            R = torch.nn.Parameter(cam_pose[:, :4].clone())
            opt = torch.optim.Adam([{"params": [T], "lr": t["lr_T"], "betas": (0.5, 0.999)},
                                    {"params": [R], "lr": t["lr_R"], "betas": (0.5, 0.999)}])
            best, cand = float("inf"), None
            for _ in range(t["iters"]):
                pose = torch.cat([R, T], -1)
                loss = trk.optimize_tracking(pose, col[None].contiguous(), dep[None].contiguous(), t["pixels"], opt)
                if loss < best:
                    best, cand = loss, pose.clone().detach()
            c2w = cam_pose_to_matrix(cand)
        est[idx] = c2w[0].detach().cpu()
        pre_c2w = c2w.clone()
        idx_s[0] = idx
    shared["stats"][0] = time.perf_counter() - t0
    shared["stats"][1] = n_updates


def mapper_proc(args, shared):
    import myslam_b200 as M
    from myslam_b200 import synthetic as S

    dev = "cuda:0"
    torch.cuda.set_device(0)
    spec = make_spec(args.scale)
    cfg = S.run_cfg(spec)
    gt, frames = make_frames(spec, args.frames, dev, args.seed)
    planes = tuple([p for p in g] for g in shared["planes"])
    dec = shared["decoders"]  # the mapper trains the SHARED decoders (ESLAM.py:117-118)
    dec.bound = shared["bound"].clone()
    mpr = holder(cfg, spec, planes, dec, shared["bound"], dev, M.MapperStep)
    m = cfg["mapping"]
    every, kf_every = 4, 4
    torch.manual_seed(args.seed + 1)
    np.random.seed(args.seed)
    est, idx_s, map_idx, first = shared["est"], shared["idx"], shared["mapping_idx"], shared["first"]
    est[0] = gt[0].cpu()
    kf_dict, kf_list = [], []
    mpr.keyframe_dict, mpr.keyframe_list, mpr.estimate_c2w_list = kf_dict, kf_list, est
    init, prev_idx, n_calls = True, -1, 0
    while True:
        while True:
            idx = int(idx_s[0])
            if idx == args.frames - 1 or (idx % every == 0 and idx != prev_idx):
                break
            time.sleep(0.001)
        prev_idx = idx
        col, dep = frames[idx]
        cur_c2w = est[idx].to(dev)
        iters, lrf = (args.iters_first, 5.0) if init else (m["iters"], 1.0)
        mpr.joint_opt = (len(kf_list) > 4) and m["joint_opt"]
        cur_c2w = mpr.optimize_mapping(iters, lrf, torch.tensor(idx), col, dep, gt[idx], kf_dict, kf_list, cur_c2w)
        torch.cuda.synchronize()
        n_calls += 1
        if mpr.joint_opt:
            est[idx] = cur_c2w.detach().cpu()
        if idx % kf_every == 0:
            kf_list.append(idx)
            kf_dict.append({"gt_c2w": gt[idx], "idx": torch.tensor(idx), "color": col, "depth": dep,
                            "est_c2w": cur_c2w.clone()})
        init = False
        first[0] = 1
        map_idx[0] = idx
        if idx == args.frames - 1:
            break
    shared["stats"][2] = n_calls


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=120)
    ap.add_argument("--iters-first", type=int, default=300)
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "e2e_two_process.json"))
    args = ap.parse_args()
    mp.set_start_method("spawn", force=True)
    import eslam_oracle as O
    import myslam_b200 as M
    from e2e_slam import horn_ate

    dev = "cuda:0"
    spec = make_spec(args.scale)
    gen = torch.Generator().manual_seed(args.seed)
    bound = O.rounded_bound(spec["bound"], spec["bound_dividable"])
    fld = O.make_field(bound, spec["planes_res"], spec["c_planes_res"], generator=gen)
    planes = tuple([p.clone().to(dev) for p in g] for g in fld.planes)  # the storage ESLAM owns (ESLAM.py:175-218)
    dec = M.Decoders(c_dim=32, truncation=spec["truncation"], learnable_beta=True)
    dec.load_state_dict({**fld.dec, "beta": fld.beta})
    dec = dec.to(dev)
    dec.share_memory()
    shared = {"planes": planes, "decoders": dec, "bound": bound,
              "est": torch.zeros(args.frames, 4, 4).share_memory_(), "idx": torch.zeros(1, dtype=torch.int64).share_memory_(),
              "mapping_idx": torch.full((1,), -1, dtype=torch.int64).share_memory_(),
              "first": torch.zeros(1, dtype=torch.int64).share_memory_(), "stats": torch.zeros(4, dtype=torch.float64).share_memory_()}
    t0 = time.perf_counter()
    procs = [mp.Process(target=tracker_proc, args=(args, shared)), mp.Process(target=mapper_proc, args=(args, shared))]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=1500)
    wall = time.perf_counter() - t0
    ok = all(p.exitcode == 0 for p in procs)
    gt, _ = make_frames(spec, args.frames, dev, args.seed)
    est = shared["est"].to(dev)
    ate, raw = horn_ate(est, gt) if ok else (float("nan"), float("nan"))
    res = {"ok": ok, "exit_codes": [p.exitcode for p in procs], "frames": args.frames, "H": spec["H"], "W": spec["W"],
           "ate_rmse_m": ate, "mean_trans_err_m": raw, "wall_s": wall, "tracking_s": float(shared["stats"][0]),
           "tracker_parameter_updates": int(shared["stats"][1]), "mapping_calls": int(shared["stats"][2]),
           "what": "tracker and mapper PROCESSES sharing planes / decoders through CUDA IPC with the reference's handshake "
                   "(Tracker.py:264-266, Mapper.py:388-396); the tracker re-imports the planes the mapper process wrote "
                   "whenever prev_mapping_idx changes"}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res))
    if not ok:
        sys.exit(1)


if __name__ == "__main__":
    main()
