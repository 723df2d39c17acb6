"""Synthetic Replica-/ScanNet-shaped inputs: an axis-aligned box room seen
from inside, rendered analytically, in the layout the reference's datasets hand to the hot path
(colour [H,W,3] float64 in [0,1], depth [H,W] float32 metres, c2w [4,4] float32 with the camera
looking down -z; src/utils/datasets.py:79-115).  Used by bench.py, the smoke test and the tests.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Tuple

import torch

REPLICA_ROOM0 = dict(
    bound=[[-1.9, 7.9], [-2.2, 4.5], [-2.5, 2.3]], H=680, W=1200, fx=600.0, fy=600.0, cx=599.5, cy=339.5,
    planes_res=(0.24, 0.06), c_planes_res=(0.24, 0.03), bound_dividable=0.24, truncation=0.06,
    n_stratified=32, n_importance=8, room=[[-1.5, 7.5], [-1.8, 4.1], [-2.1, 1.9]],
    tracking=dict(pixels=2000, iters=8, lr_T=0.002, lr_R=0.001, ignore_edge_W=75, ignore_edge_H=75,
                  w_sdf_fs=10, w_sdf_center=200, w_sdf_tail=50, w_depth=1, w_color=5),
    mapping=dict(pixels=4000, iters=15, mapping_window_size=20, keyframe_selection_method='global', joint_opt=True,
                 joint_opt_cam_lr=0.001, w_sdf_fs=5, w_sdf_center=200, w_sdf_tail=10, w_depth=0.1, w_color=5,
                 lr=dict(decoders_lr=0.001, planes_lr=0.005, c_planes_lr=0.005)))

SCANNET_0000 = dict(
    bound=[[-2.0, 11.0], [-2.0, 11.5], [-2.0, 5.5]], H=460, W=620,  # configs/ScanNet/scene0000.yaml:3
    fx=577.590698, fy=578.729797, cx=308.906342,
    cy=232.683609, planes_res=(0.24, 0.06), c_planes_res=(0.24, 0.03), bound_dividable=0.24, truncation=0.06,
    n_stratified=48, n_importance=8, room=[[0.2, 8.2], [0.2, 8.5], [0.1, 2.9]],
    tracking=dict(pixels=2000, iters=30, lr_T=0.0005, lr_R=0.0025, ignore_edge_W=75, ignore_edge_H=75,
                  w_sdf_fs=10, w_sdf_center=200, w_sdf_tail=50, w_depth=1, w_color=5),
    mapping=dict(pixels=4000, iters=30, mapping_window_size=20, keyframe_selection_method='global', joint_opt=True,
                 joint_opt_cam_lr=0.001, w_sdf_fs=5, w_sdf_center=200, w_sdf_tail=10, w_depth=0.1, w_color=5,
                 lr=dict(decoders_lr=0.001, planes_lr=0.005, c_planes_lr=0.005)))


def rounded_bound(bound, bound_dividable=0.24, scale=1.0) -> torch.Tensor:
    """ESLAM.load_bound (ESLAM.py:159-173): fp32 round-up of the upper bound."""
    b = (torch.tensor(bound, dtype=torch.float64) * scale).float()
    b[:, 1] = (((b[:, 1] - b[:, 0]) / bound_dividable).int() + 1) * bound_dividable + b[:, 0]
    return b


def plane_shapes(bound: torch.Tensor, res: float):
    """ESLAM.init_planes (ESLAM.py:196-203): (H,W) of the xy, xz, yz planes at one resolution."""
    nx, ny, nz = map(int, ((bound[:, 1] - bound[:, 0]) / res).tolist())
    return (ny, nx), (nz, nx), (nz, ny)


def init_planes(bound, planes_res, c_planes_res, device, generator=None, std=0.01):
    """The six plane lists [coarse, fine] of [1,32,H,W], N(0, std) like ESLAM.py:201-210."""
    groups: List[List[torch.Tensor]] = [[], [], [], [], [], []]
    for base, ress in ((0, planes_res), (3, c_planes_res)):
        for r in ress:
            for k, hw in enumerate(plane_shapes(bound, r)):
                p = torch.empty(1, 32, *hw).normal_(0, std, generator=generator)
                groups[base + k].append(p.to(device))
    return tuple(groups)


def look_at_pose(eye, target, up=(0.0, 0.0, 1.0)) -> torch.Tensor:
    """c2w [4,4] of a camera at `eye` looking at `target`, ESLAM convention (camera looks down -z, +y up)."""
    eye = torch.tensor(eye, dtype=torch.float64)
    fwd = torch.tensor(target, dtype=torch.float64) - eye
    fwd = fwd / fwd.norm()
    upv = torch.tensor(up, dtype=torch.float64)
    right = torch.linalg.cross(fwd, upv)
    right = right / right.norm()
    true_up = torch.linalg.cross(right, fwd)
    c2w = torch.eye(4, dtype=torch.float64)
    c2w[:3, 0], c2w[:3, 1], c2w[:3, 2], c2w[:3, 3] = right, true_up, -fwd, eye
    return c2w.float()


def trajectory(n: int, room, radius_frac=0.25, step_deg=0.5) -> torch.Tensor:
    """n poses on a smooth closed loop inside the room, looking outwards-ish (~0.5 deg / frame)."""
    lo = [r[0] for r in room]
    hi = [r[1] for r in room]
    c = [(a + b) / 2 for a, b in zip(lo, hi)]
    rx, ry = (hi[0] - lo[0]) * radius_frac, (hi[1] - lo[1]) * radius_frac
    out = []
    for k in range(n):
        a = math.radians(k * step_deg)
        eye = (c[0] + rx * math.cos(a), c[1] + ry * math.sin(a), c[2] + 0.1 * math.sin(3 * a))
        tgt = (c[0] + 3 * rx * math.cos(a + 0.8), c[1] + 3 * ry * math.sin(a + 0.8), c[2] - 0.3)
        out.append(look_at_pose(eye, tgt))
    return torch.stack(out, 0)


def render_box_room(c2w, H, W, fx, fy, cx, cy, room, device, hole_frac=0.03, generator=None):
    """Analytic RGB-D frame of the box room: depth is the camera-axis distance to the wall the pixel's ray
    hits (rays_d is un-normalised with dir_cam.z = -1, so the ray parameter IS that depth), colour a smooth
    function of the hit point; a fraction of pixels gets depth 0 (sensor holes -> depth-less rays)."""
    dev = torch.device(device)
    c2w = c2w.to(dev)
    jj, ii = torch.meshgrid(torch.arange(H, dtype=torch.float32, device=dev),
                            torch.arange(W, dtype=torch.float32, device=dev), indexing="ij")
    dirs = torch.stack([(ii - cx) / fx, -(jj - cy) / fy, -torch.ones_like(ii)], -1)
    d = torch.sum(dirs[..., None, :] * c2w[:3, :3], -1)
    o = c2w[:3, 3]
    lo = torch.tensor([r[0] for r in room], dtype=torch.float32, device=dev)
    hi = torch.tensor([r[1] for r in room], dtype=torch.float32, device=dev)
    t = torch.maximum((lo - o) / d, (hi - o) / d).min(-1)[0]
    hit = o + d * t[..., None]
    depth = t.clone()
    if hole_frac > 0:
        holes = torch.rand(H, W, generator=generator).to(dev) < hole_frac
        depth[holes] = 0.0
    color = torch.stack([0.5 + 0.4 * torch.sin(1.7 * hit[..., 0] + 0.5 * hit[..., 2]),
                         0.5 + 0.4 * torch.cos(1.3 * hit[..., 1] - 0.7 * hit[..., 0]),
                         0.5 + 0.3 * torch.sin(0.9 * (hit[..., 0] + hit[..., 1] + hit[..., 2]))], -1)
    return color.double().contiguous(), depth.float().contiguous()


@dataclass
class Scene:
    spec: dict
    bound: torch.Tensor
    all_planes: tuple
    decoders: object
    cam: Tuple


def make_scene(spec: dict, device, seed: int = 0, decoders_cls=None) -> Scene:
    """Random-init map with the reference's shapes (ESLAM.__init__ order: decoders first, then planes)."""
    from .decoders import Decoders

    gen = torch.Generator().manual_seed(seed)
    bound = rounded_bound(spec["bound"], spec["bound_dividable"])
    cls = decoders_cls or Decoders
    state = torch.random.get_rng_state()
    torch.manual_seed(seed)
    dec = cls(c_dim=32, truncation=spec["truncation"], learnable_beta=True)
    torch.random.set_rng_state(state)
    dec.bound = bound
    dec = dec.to(device)
    planes = init_planes(bound, spec["planes_res"], spec["c_planes_res"], device, gen)
    cam = (spec["H"], spec["W"], spec["fx"], spec["fy"], spec["cx"], spec["cy"])
    return Scene(spec, bound, planes, dec, cam)


def run_cfg(spec: dict) -> dict:
    """The slice of the reference's yaml the hot-path classes read."""
    return {"scale": 1,
            "rendering": {"perturb": True, "n_stratified": spec["n_stratified"],
                          "n_importance": spec["n_importance"], "learnable_beta": True},
            "model": {"c_dim": 32, "truncation": spec["truncation"]},
            "tracking": dict(spec["tracking"]), "mapping": dict(spec["mapping"])}
