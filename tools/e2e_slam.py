"""End-to-end check of north_star's system-level criterion: ATE and depth-L1 of a full tracking+mapping run on a
synthetic Replica-shaped sequence, B200 kernels vs the reference's PyTorch path (oracle port, stock ATen kernels,
same GPU), same schedule, same initial map, same seeds.

The schedule is the reference's (ESLAM.yaml defaults: first frame mapped for `iters_first`, then every 4th frame
mapped for 15 iterations over a <=20-keyframe window with joint pose optimisation once 5 keyframes exist, every
frame tracked for 8 iterations from a constant-speed initial pose; Tracker.py:254-313, Mapper.py:387-460), run
sequentially in one process instead of two spin-waiting processes.

    python tools/e2e_slam.py --frames 60 --iters-first 300 [--scale 0.5]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

import eslam_oracle as O  # noqa: E402  (the reference arm of this comparison)
import myslam_b200 as M  # noqa: E402
from myslam_b200 import synthetic as S  # noqa: E402
from myslam_b200.common import cam_pose_to_matrix, matrix_to_cam_pose  # noqa: E402


def horn_ate(est, gt):
    """ATE RMSE of translations after a rigid (Horn / Umeyama without scale) alignment, as src/tools/eval_ate.py."""
    e, g = est[:, :3, 3].double().cpu().numpy().T, gt[:, :3, 3].double().cpu().numpy().T
    em, gm = e.mean(1, keepdims=True), g.mean(1, keepdims=True)
    W = (e - em) @ (g - gm).T
    U, _, Vt = np.linalg.svd(W.T)
    Sg = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        Sg[2, 2] = -1
    R = U @ Sg @ Vt
    t = gm - R @ em
    err = np.linalg.norm(R @ e + t - g, axis=0)
    return float(np.sqrt((err ** 2).mean())), float(np.linalg.norm(e - g, axis=0).mean())


class OursArm:
    def __init__(self, spec, cfg, fld, dev):
        self.dev = dev
        planes = tuple([p.clone().to(dev) for p in g] for g in fld.planes)
        dec = M.Decoders(c_dim=32, truncation=spec["truncation"], learnable_beta=True)
        dec.load_state_dict({**fld.dec, "beta": fld.beta})
        dec = dec.to(dev)
        dec.bound = fld.bound.clone()
        cam = (spec["H"], spec["W"], spec["fx"], spec["fy"], spec["cx"], spec["cy"])

        class E:
            pass

        e = E()
        e.bound, e.device = fld.bound.clone(), dev
        e.H, e.W, e.fx, e.fy, e.cx, e.cy = cam
        self.rnd = M.Renderer(cfg, e)
        self.trk = M.TrackerStep(cfg, self.rnd, dec, planes, fld.bound.clone(), cam, dev)
        self.mp = M.MapperStep(cfg, self.rnd, dec, planes, fld.bound.clone(), cam, dev)
        self.planes, self.dec = planes, dec
        self.t_track = self.t_map = 0.0

    def track(self, pose0, col, dep):
        t0 = time.perf_counter()
        self.trk.prev_mapping_idx = getattr(self, "_map_count", 0)  # planes changed in place: re-import
        best, losses, _ = self.trk.track_frame(pose0, col[None].contiguous(), dep[None].contiguous())
        torch.cuda.synchronize()
        self.t_track += time.perf_counter() - t0
        return best

    def map(self, iters, lr_factor, idx, col, dep, gt_c2w, kf_dict, kf_list, cur_c2w, joint):
        t0 = time.perf_counter()
        self.mp.keyframe_dict, self.mp.keyframe_list, self.mp.joint_opt = kf_dict, kf_list, joint
        out = self.mp.optimize_mapping(iters, lr_factor, idx, col, dep, gt_c2w, kf_dict, kf_list, cur_c2w)
        self._map_count = getattr(self, "_map_count", 0) + 1
        torch.cuda.synchronize()
        self.t_map += time.perf_counter() - t0
        return out

    def render_depth(self, c2w, dep):
        d, _ = self.rnd.render_img((self.trk.planes_xy, self.trk.planes_xz, self.trk.planes_yz, self.trk.c_planes_xy,
                                    self.trk.c_planes_xz, self.trk.c_planes_yz), self.dec, c2w, self.trk.truncation,
                                   self.dev, gt_depth=dep)
        return d


class OracleArm:
    """The reference's PyTorch path (oracle port) on the same device."""

    def __init__(self, spec, cfg, fld, dev):
        self.dev, self.spec, self.cfg = dev, spec, cfg
        self.fld = O.Field(tuple([p.clone().to(dev) for p in g] for g in fld.planes),
                           {k: v.clone().to(dev) for k, v in fld.dec.items()}, fld.beta.clone().to(dev),
                           fld.bound.clone().to(dev))
        self.cam = O.Camera(spec["H"], spec["W"], spec["fx"], spec["fy"], spec["cx"], spec["cy"])
        self.rc = O.RenderCfg(spec["n_stratified"], spec["n_importance"], spec["truncation"])
        self.draws = O.LiveDraws(None, dev)
        self.t_track = self.t_map = 0.0
        self.H, self.W, self.fx, self.fy, self.cx, self.cy = (spec["H"], spec["W"], spec["fx"], spec["fy"], spec["cx"],
                                                              spec["cy"])
        self.device = dev

    def track(self, pose0, col, dep):
        t = self.cfg["tracking"]
        t0 = time.perf_counter()
        best, _, _ = O.track_frame(self.fld, self.cam, self.rc, O.TRACK_W, pose0, col[None], dep[None], t["pixels"],
                                   t["ignore_edge_H"], t["ignore_edge_W"], t["iters"], t["lr_T"], t["lr_R"], self.draws)
        self.draws.log.clear()
        torch.cuda.synchronize()
        self.t_track += time.perf_counter() - t0
        return best

    def map(self, iters, lr_factor, idx, col, dep, gt_c2w, kf_dict, kf_list, cur_c2w, joint):
        from myslam_b200.mapper import keyframe_selection_overlap
        from myslam_b200.common import random_select

        m = self.cfg["mapping"]
        t0 = time.perf_counter()
        self.keyframe_list, self.estimate_c2w_list = kf_list, self.est_list
        if len(kf_dict) == 0:
            frames = []
        elif m["keyframe_selection_method"] == "global":
            frames = random_select(len(kf_dict) - 2, m["mapping_window_size"] - 1)
        else:
            frames = keyframe_selection_overlap(self, col, dep, cur_c2w, m["mapping_window_size"] - 1)
        if len(kf_list) > 1:
            frames = sorted(frames + [len(kf_list) - 1] + [len(kf_list) - 2])
        frames += [-1]
        deps = torch.stack([dep if f == -1 else kf_dict[f]["depth"] for f in frames], 0)
        cols = torch.stack([col if f == -1 else kf_dict[f]["color"] for f in frames], 0)
        c2ws = torch.stack([cur_c2w if f == -1 else kf_dict[f]["est_c2w"] for f in frames], 0)
        lr = m["lr"]
        c2ws_new, _ = O.map_window(self.fld, self.cam, self.rc, O.MAP_W, c2ws, cols, deps, m["pixels"], iters,
                                   lr["decoders_lr"] * lr_factor, lr["planes_lr"] * lr_factor,
                                   lr["c_planes_lr"] * lr_factor, joint, m["joint_opt_cam_lr"], self.draws)
        self.draws.log.clear()
        if joint:
            k = 0
            for f in frames[1:]:
                if f != -1:
                    kf_dict[f]["est_c2w"] = c2ws_new[1 + k].clone()
                    k += 1
                else:
                    cur_c2w = c2ws_new[-1].clone()
        torch.cuda.synchronize()
        self.t_map += time.perf_counter() - t0
        return cur_c2w


def run_slam(arm, spec, cfg, frames, gt_c2ws, dev, iters_first, seed):
    torch.manual_seed(seed)
    np.random.seed(seed)
    n = len(frames)
    est = torch.zeros(n, 4, 4, device=dev)
    arm.est_list = est
    if hasattr(arm, "mp"):
        arm.mp.estimate_c2w_list = est
    kf_dict, kf_list = [], []
    m = cfg["mapping"]
    every, kf_every = 4, 4
    for idx in range(n):
        col, dep = frames[idx]
        gt = gt_c2ws[idx]
        if idx == 0:
            c2w = gt.clone()
        else:
            if idx >= 2:
                pre = matrix_to_cam_pose(torch.stack([est[idx - 2], est[idx - 1]], 0))
                pose0 = 2 * pre[1:] - pre[0:1]
            else:
                pose0 = matrix_to_cam_pose(est[idx - 1][None])
            c2w = cam_pose_to_matrix(arm.track(pose0, col, dep))[0]
        est[idx] = c2w
        if idx % every == 0 or idx == n - 1:
            joint = len(kf_list) > 4 and m["joint_opt"]
            iters, lrf = (iters_first, 5.0) if idx == 0 else (m["iters"], 1.0)
            cur = arm.map(iters, lrf, torch.tensor(idx), col, dep, gt, kf_dict, kf_list, est[idx].clone(), joint)
            if joint:
                est[idx] = cur
            if idx % kf_every == 0:
                kf_list.append(idx)
                kf_dict.append({"gt_c2w": gt, "idx": torch.tensor(idx), "color": col, "depth": dep,
                                "est_c2w": cur.clone()})
    return est


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=60)
    ap.add_argument("--iters-first", type=int, default=300)
    ap.add_argument("--scale", type=float, default=1.0, help="image scale (1.0 = 1200x680)")
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--fast-rng", action="store_true", help="upper-bound draw shapes (no host sync); default: the "
                    "reference's draw shapes, so both arms consume the same random stream")
    ap.add_argument("--ours-only", action="store_true", help="skip the reference arm (its statistics over seeds are in "
                    "profiles/r02_e2e_slam.json): ATE / depth-L1 of this build alone")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "e2e_slam.json"))
    args = ap.parse_args()
    dev = "cuda:0"
    torch.backends.cuda.matmul.allow_tf32 = False
    spec = dict(S.REPLICA_ROOM0)
    if args.scale != 1.0:
        s = args.scale
        spec.update(H=int(680 * s), W=int(1200 * s), fx=600.0 * s, fy=600.0 * s, cx=(1200 * s - 1) / 2,
                    cy=(680 * s - 1) / 2)
        spec["tracking"] = dict(spec["tracking"], ignore_edge_H=int(75 * s), ignore_edge_W=int(75 * s))
    cfg = S.run_cfg(spec)
    gen = torch.Generator().manual_seed(args.seed)
    bound = O.rounded_bound(spec["bound"], spec["bound_dividable"])
    fld = O.make_field(bound, spec["planes_res"], spec["c_planes_res"], generator=gen)
    gt_c2ws = S.trajectory(args.frames, spec["room"], step_deg=0.5).to(dev)
    cam = (spec["H"], spec["W"], spec["fx"], spec["fy"], spec["cx"], spec["cy"])
    frames = [S.render_box_room(gt_c2ws[k], *cam, spec["room"], dev, hole_frac=0.02, generator=gen)
              for k in range(args.frames)]
    res = {}
    arms = (("b200", OursArm),) if args.ours_only else (("b200", OursArm), ("reference_torch_gpu", OracleArm))
    for name, cls in arms:
        arm = cls(spec, cfg, fld, dev)
        if isinstance(arm, OursArm):
            arm.trk.strict_rng = arm.mp.strict_rng = not args.fast_rng
        t0 = time.perf_counter()
        est = run_slam(arm, spec, cfg, frames, gt_c2ws, dev, args.iters_first, args.seed)
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        ate, raw = horn_ate(est, gt_c2ws)
        # depth L1 of the final map rendered from the GT poses of a few frames, by the SAME renderer on both arms
        ours = arm if isinstance(arm, OursArm) else None
        res[name] = {"ate_rmse_m": ate, "mean_trans_err_m": raw, "wall_s": wall, "track_s": arm.t_track,
                     "map_s": arm.t_map}
        res[name]["_arm"] = arm
        print(name, {k: v for k, v in res[name].items() if k != "_arm"}, flush=True)
    # depth L1 with our renderer on both final maps
    probe = OursArm(spec, cfg, fld, dev)
    for name, _ in arms:
        arm = res[name].pop("_arm")
        if isinstance(arm, OracleArm):
            f = arm.fld
            planes = tuple([p.detach().clone() for p in g] for g in f.planes)
            probe.dec.load_state_dict({**{k: v.detach() for k, v in f.dec.items()}, "beta": f.beta.detach()})
            (probe.trk.planes_xy, probe.trk.planes_xz, probe.trk.planes_yz, probe.trk.c_planes_xy,
             probe.trk.c_planes_xz, probe.trk.c_planes_yz) = planes
            r = probe
        else:
            r = arm
        l1 = []
        for k in range(0, args.frames, max(args.frames // 6, 1)):
            dep = frames[k][1]
            d = r.render_depth(gt_c2ws[k], dep)
            ok = dep > 0
            l1.append((d[ok].float() - dep[ok]).abs().mean().item())
        res[name]["depth_l1_m"] = float(np.mean(l1))
    if not args.ours_only:
        a, b = res["b200"], res["reference_torch_gpu"]
        res["ate_ratio"] = a["ate_rmse_m"] / max(b["ate_rmse_m"], 1e-12)
        res["depth_l1_ratio"] = a["depth_l1_m"] / max(b["depth_l1_m"], 1e-12)
        res["speedup_wall"] = b["wall_s"] / a["wall_s"]
    res["config"] = {"frames": args.frames, "iters_first": args.iters_first, "H": spec["H"], "W": spec["W"],
                     "seed": args.seed, "fast_rng": bool(args.fast_rng)}
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)
    print(json.dumps(res))


if __name__ == "__main__":
    main()
