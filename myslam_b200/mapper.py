"""Drop-in for the hot half of `src/Mapper.py` (reference lines 110-144, 211-364):
`optimize_mapping` with the reference's signature and side effects (planes and decoders updated in
place in the storage `ESLAM` owns, `keyframe_dict[*]['est_c2w']` rewritten, `cur_c2w` returned).
The per-iteration loop (sample, render, losses, backward, Adam) runs as fused kernels on the
parameter arena; the arena is written back to the reference's [1,32,H,W] tensors once per call.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

import torch

from .common import cam_pose_to_matrix, get_samples, matrix_to_cam_pose, random_select
from .decoders import decoder_tensors, synced_store
from .field import FieldStore
from ._lib import call, ptr, stream
from .hotpath import FrameTable, StepCfg, Workspace, make_camera, mapping_iteration, mapping_window_pipelined
from .renderer import TorchDraws, linspace_table
from .renderer import make_cfg


def _strict_default() -> bool:
    return os.environ.get("ESLAM_B200_STRICT_RNG", "0") == "1"


def _window_body(store: FieldStore, ws: Workspace, sc: StepCfg, c2ws, poses7, gt_colors, gt_depths, pix: int, iters: int,
                 lr_dec, lr_planes, lr_cplanes, lr_cam, out, pipelined: bool, draws=None, strict_rng=False, losses=None,
                 exchange=None):
    """Everything map_window puts on the stream, on buffers the caller owns (so that it can be captured):
    fresh Adam state, the poses of frames 1.. from their matrices, `iters` iterations, the matrices of the optimised
    poses into `out` [b,4,4]."""
    b = c2ws.shape[0]
    store.reset_adam()
    out.copy_(c2ws)
    if poses7 is not None:
        poses7.zero_()
        if b > 1:  # matrix_to_cam_pose(c2ws[1:]) (Mapper.py:289) as one launch
            call("eslam_matrix_to_pose", ptr(c2ws[1:]), ptr(poses7[1:]), b - 1, stream())
        ws.pose_m.zero_()
        ws.pose_v.zero_()
        ws.pose_grad.zero_()
    if pipelined:  # the default path: two streams, the next iteration's sampling under this one's optimiser step
        mapping_window_pipelined(ws, store, sc, c2ws, poses7, gt_colors, gt_depths, pix, iters, lr_dec, lr_planes,
                                 lr_cplanes, lr_cam, exchange=exchange)
    else:
        for it in range(iters):
            mapping_iteration(ws, store, sc, c2ws, poses7, gt_colors, gt_depths, pix, it + 1, lr_dec, lr_planes,
                              lr_cplanes, lr_cam, draws=draws, strict_rng=strict_rng, want_loss=losses is not None,
                              exchange=exchange)
            if losses is not None:
                losses.append(ws.loss_acc[5].clone())
    if poses7 is not None and b > 1:  # cam_pose_to_matrix of the optimised poses (Mapper.py:352-362) as one launch
        call("eslam_pose_to_matrix", ptr(poses7[1:]), ptr(out[1:]), b - 1, stream())


class _WindowGraph:
    """One `optimize_mapping` call's device work -- optimiser reset, `iters` pipelined iterations over the loop's three
    streams, pose conversion -- captured once as a CUDA graph and replayed (the loop is ~20 launches per iteration; a
    replay costs the host one call, so the mapper process is free while the window runs and the launch gaps between
    the kernels close).  Everything the kernels address is persistent: the parameter / optimiser arenas, the
    workspace and its draw buffers, and the three buffers below (window poses in, poses out, frame-pointer table),
    which are rewritten before every replay.  Adam restarts at step 1 in every call (Mapper.py:291-299), so the bias
    corrections baked into the launches are the same for every call of the same shape.  The random draws are
    torch's (its CUDA generator is graph safe)."""

    def __init__(self, store, ws, sc, frames: FrameTable, b, pix, iters, lrs, joint, lr_cam):
        from . import _lib

        dev = ws.device
        self.store = store  # kept alive: the graph addresses its arenas (and its id is part of the cache key)
        self.c2ws = torch.zeros(b, 4, 4, dtype=torch.float32, device=dev)
        self.out = torch.zeros(b, 4, 4, dtype=torch.float32, device=dev)
        self.poses7 = torch.zeros(b, 7, dtype=torch.float32, device=dev) if joint else None
        self.frames = FrameTable(frames.colors, frames.depths, sc.cam, dev)  # own table; pointers rewritten per replay
        self.iters = iters
        store.q_gen = -1  # the graph rebuilds the Q images itself: whatever changed the parameters between two calls
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        l0 = _lib.LAUNCHES
        with torch.cuda.graph(self.graph, capture_error_mode="thread_local"):
            _window_body(store, ws, sc, self.c2ws, self.poses7, self.frames, self.frames, pix, iters, *lrs, lr_cam,
                         self.out, True)
        self.n_launches = _lib.LAUNCHES - l0
        store.q_gen = -1  # nothing ran: the bookkeeping of the captured loop does not describe the arena

    def run(self, store, c2ws, frames: FrameTable):
        from . import _lib

        self.c2ws.copy_(c2ws)
        self.frames.colors, self.frames.depths = frames.colors, frames.depths  # keep the frames alive
        self.frames.table.copy_(frames.table)
        self.graph.replay()
        _lib.LAUNCHES += self.n_launches
        store.gen += self.iters  # one optimiser step per iteration; the Q images are those of the last but one
        store.q_gen = -1
        return self.out.clone()


def _graph_on() -> bool:
    return os.environ.get("ESLAM_B200_GRAPH", "1") == "1"


def map_window(store: FieldStore, ws: Workspace, sc: StepCfg, c2ws, gt_colors, gt_depths, n_pixels: int, iters: int,
               lr_dec: float, lr_planes: float, lr_cplanes: float, joint_opt: bool, lr_cam: float, draws=None,
               strict_rng: bool = False, losses: Optional[list] = None, exchange=None):
    """The loop of Mapper.optimize_mapping (Mapper.py:288-350) for an already chosen window:
    fresh Adam state, `iters` fused iterations on `store`.  Returns the window's c2ws [b,4,4] after the
    call (frame 0 is held fixed, Mapper.py:314).  `exchange` (myslam_b200.dist) makes the call one rank of a
    ray-sharded multi-GPU mapping: every rank passes the same window and draws its own `n_pixels` rays.

    On the default path (torch's generator, fixed-shape draws) a single-GPU call whose shape has been seen before
    replays a CUDA graph of the whole call (_WindowGraph; ESLAM_B200_GRAPH=0 launches kernel by kernel)."""
    b = c2ws.shape[0]
    pix = n_pixels // b
    c2ws = c2ws.float().contiguous()
    pipelined = (draws is None and not strict_rng and losses is None and iters > 0 and sc.perturb
                 and (exchange is None or hasattr(exchange, "adam_exchange"))
                 and os.environ.get("ESLAM_B200_PIPELINE", "1") == "1")
    if pipelined and exchange is None and isinstance(gt_depths, FrameTable) and gt_colors is gt_depths and _graph_on():
        graphs = ws.__dict__.setdefault("_window_graphs", {})
        key = (id(store), store.arena.data_ptr(), b, pix, iters, lr_dec, lr_planes, lr_cplanes, bool(joint_opt), lr_cam)
        ent = graphs.get(key)
        if ent is None:
            graphs[key] = "seen"  # the first call of a shape runs kernel by kernel (lazy allocations happen there)
        else:
            if ent == "seen":
                if len(graphs) > 64:
                    graphs.clear()
                try:
                    ent = _WindowGraph(store, ws, sc, gt_depths, b, pix, iters, (lr_dec, lr_planes, lr_cplanes),
                                       joint_opt, lr_cam)
                except Exception as exc:  # capture refused (driver / torch build): same kernels, launched one by one
                    import warnings

                    warnings.warn(f"myslam_b200: CUDA-graph capture of the mapping window failed ({exc}); "
                                  "launching kernel by kernel")
                    ent = "eager"
                graphs[key] = ent
            if ent != "eager":
                return ent.run(store, c2ws, gt_depths)
    poses7 = torch.zeros(b, 7, dtype=torch.float32, device=c2ws.device) if joint_opt else None
    out = torch.empty_like(c2ws)
    _window_body(store, ws, sc, c2ws, poses7, gt_colors, gt_depths, pix, iters, lr_dec, lr_planes, lr_cplanes, lr_cam,
                 out, pipelined, draws=draws, strict_rng=strict_rng, losses=losses, exchange=exchange)
    return out


def keyframe_selection_overlap(self, gt_color, gt_depth, c2w, num_keyframes, num_samples=8, num_rays=50):
    """Keyframes whose frusta see the current view's surface (reference Mapper.py:146-209): one kernel projects the
    50 x 8 sample points into every keyframe (eslam_keyframe_overlap); the pixel draw (randint, common.py:108),
    the CPU randperm pick (Mapper.py:205-209) and the returned list are the reference's."""
    device = self.device
    draws = getattr(self, "draws", None) or TorchDraws(device)
    idx = draws.randint(self.H * self.W, num_rays)
    kf_c2w = torch.stack([self.estimate_c2w_list[i] for i in self.keyframe_list], dim=0)[:-2]  # last two: always in
    K = kf_c2w.shape[0]
    if K == 0:
        return []
    kf_c2w = kf_c2w.to(device=device, dtype=torch.float32).reshape(K, 16).contiguous()
    depth = gt_depth.to(device=device, dtype=torch.float32).contiguous()
    cur = c2w.to(device=device, dtype=torch.float32).reshape(-1)[:16].contiguous()
    inside = torch.empty(K, dtype=torch.int32, device=device)
    n_pts = torch.zeros(1, dtype=torch.int32, device=device)
    cam = make_camera(self.H, self.W, self.fx, self.fy, self.cx, self.cy)
    call("eslam_keyframe_overlap", C.byref(cam), ptr(cur), ptr(depth), ptr(idx), num_rays,
         ptr(linspace_table(num_samples, device)), num_samples, ptr(kf_c2w), K, ptr(inside), ptr(n_pts), stream())
    self._last_overlap = (inside, n_pts)  # percent_inside = inside / n_pts (tests)
    # Mapper.py:205-209 (nonzero, CPU randperm, first num_keyframes, .cpu()) with ONE device->host copy: the K
    # per-keyframe counts (4 K bytes); the selection and the permutation (torch's CPU generator, as the reference) then
    # run on the host, where the list is needed anyway.  The reference syncs twice (nonzero, .cpu()).
    sel = torch.nonzero(inside.cpu()).squeeze(-1)
    sel = sel[torch.randperm(sel.shape[0])[:num_keyframes]]
    return list(sel.numpy())


def _device_frame(self, t: torch.Tensor, dtype) -> torch.Tensor:
    """A contiguous device copy of a keyframe image, cached across optimize_mapping calls (`keyframe_device: cpu`
    makes the reference re-upload every window frame on every call, Mapper.py:276-277)."""
    dev = torch.device(self.device)
    if t.device == dev and t.dtype == dtype and t.is_contiguous():
        return t
    cache = self.__dict__.setdefault("_kf_cache", {})
    ent = cache.get(id(t))
    if ent is not None and ent[0] is t and ent[2] == t._version:
        return ent[1]
    limit = int(os.environ.get("ESLAM_B200_KF_CACHE", "512"))
    while len(cache) >= limit:
        cache.pop(next(iter(cache)))
    out = t.to(device=dev, dtype=dtype).contiguous()
    cache[id(t)] = (t, out, t._version)
    return out


def _mapper_state(mp, n_rays, b):
    st = getattr(mp, "_b200", None)
    rnd = mp.renderer
    S = rnd.n_stratified + rnd.n_importance
    if st is None or not st["ws"].fits(n_rays, S, b):
        cam = make_camera(mp.H, mp.W, mp.fx, mp.fy, mp.cx, mp.cy)
        rc = make_cfg(rnd.n_stratified, rnd.n_importance, mp.truncation,
                      (mp.w_sdf_fs, mp.w_sdf_center, mp.w_sdf_tail, mp.w_depth, mp.w_color))
        st = {"ws": Workspace(mp.device, max(n_rays, mp.mapping_pixels), S, max(32, b)),
              "sc": StepCfg(cam, rc, bool(rnd.perturb))}
        mp._b200 = st
    return st


def optimize_mapping(self, iters, lr_factor, idx, cur_gt_color, cur_gt_depth, gt_cur_c2w, keyframe_dict,
                     keyframe_list, cur_c2w):
    """Mapping iterations over a window of keyframes (reference Mapper.optimize_mapping, Mapper.py:211-364).
    Returns the (possibly jointly optimised) cur_c2w."""
    all_planes = (self.planes_xy, self.planes_xz, self.planes_yz, self.c_planes_xy, self.c_planes_xz, self.c_planes_yz)
    cfg, device = self.cfg, self.device
    # ---- window selection, as the reference (Mapper.py:235-247)
    if len(keyframe_dict) == 0:
        optimize_frame: List[int] = []
    elif self.keyframe_selection_method == 'global':
        optimize_frame = random_select(len(self.keyframe_dict) - 2, self.mapping_window_size - 1)
    elif self.keyframe_selection_method == 'overlap':
        optimize_frame = self.keyframe_selection_overlap(cur_gt_color, cur_gt_depth, cur_c2w,
                                                         self.mapping_window_size - 1)
    if len(keyframe_list) > 1:
        optimize_frame = sorted(optimize_frame + [len(keyframe_list) - 1] + [len(keyframe_list) - 2])
    optimize_frame += [-1]
    b = len(optimize_frame)
    # ---- stage the window (Mapper.py:268-286)
    # (no torch.stack of the frames: the sampling kernel reads each frame where it lives)
    depths = [_device_frame(self, cur_gt_depth if f == -1 else keyframe_dict[f]['depth'], torch.float32)
              for f in optimize_frame]
    colors = [_device_frame(self, cur_gt_color if f == -1 else keyframe_dict[f]['color'], torch.float64)
              for f in optimize_frame]
    c2ws = torch.stack([cur_c2w if f == -1 else keyframe_dict[f]['est_c2w'] for f in optimize_frame], dim=0)
    st = _mapper_state(self, (self.mapping_pixels // b) * b, b)
    gt_colors = gt_depths = FrameTable(colors, depths, st["sc"].cam, device)
    store = synced_store(all_planes, self.decoders, self.bound)
    lr = cfg['mapping']['lr']
    c2ws_new = map_window(store, st["ws"], st["sc"], c2ws, gt_colors, gt_depths, self.mapping_pixels, iters,
                          lr['decoders_lr'] * lr_factor, lr['planes_lr'] * lr_factor, lr['c_planes_lr'] * lr_factor,
                          bool(self.joint_opt), self.joint_opt_cam_lr, draws=getattr(self, "draws", None),
                          strict_rng=getattr(self, "strict_rng", _strict_default()),
                          losses=getattr(self, "loss_log", None), exchange=getattr(self, "exchange", None))
    # ---- write the map back into the storage ESLAM owns (shared with the tracker process)
    store.push_planes(all_planes)
    store.push_decoders(self.decoders)
    store._sig = store.signature(all_planes, decoder_tensors(self.decoders))
    if self.joint_opt:
        k = 0
        for f in optimize_frame[1:]:
            if f != -1:
                keyframe_dict[f]['est_c2w'] = c2ws_new[1 + k]  # rows of one fresh [b,4,4] tensor nobody else holds
                k += 1
            else:
                cur_c2w = c2ws_new[-1]
    return cur_c2w


class MapperStep:
    """Standalone holder of exactly the attributes `optimize_mapping` reads (tests, bench)."""

    def __init__(self, cfg, renderer, decoders, all_planes, bound, cam, device, estimate_c2w_list=None):
        (self.planes_xy, self.planes_xz, self.planes_yz, self.c_planes_xy, self.c_planes_xz, self.c_planes_yz) = all_planes
        self.cfg, self.renderer, self.decoders, self.bound, self.device = cfg, renderer, decoders, bound, device
        self.H, self.W, self.fx, self.fy, self.cx, self.cy = cam
        m = cfg['mapping']
        self.truncation = cfg['model']['truncation']
        self.w_sdf_fs, self.w_sdf_center, self.w_sdf_tail = m['w_sdf_fs'], m['w_sdf_center'], m['w_sdf_tail']
        self.w_depth, self.w_color = m['w_depth'], m['w_color']
        self.mapping_pixels = m['pixels']
        self.mapping_window_size = m['mapping_window_size']
        self.keyframe_selection_method = m['keyframe_selection_method']
        self.joint_opt = False
        self.joint_opt_cam_lr = m['joint_opt_cam_lr']
        self.keyframe_dict: list = []
        self.keyframe_list: list = []
        self.estimate_c2w_list = estimate_c2w_list

    optimize_mapping = optimize_mapping
    keyframe_selection_overlap = keyframe_selection_overlap
