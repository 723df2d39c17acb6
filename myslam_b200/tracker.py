"""Drop-in for the hot half of `src/Tracker.py` (reference lines 114-210): `optimize_tracking`
with the reference's signature, bound onto the reference's `Tracker` (or used through
`TrackerStep` standalone), plus `track_frame`, the whole per-frame camera loop
(Tracker.py:291-309) with the Adam step fused on the device.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from .common import cam_pose_to_matrix
from .decoders import synced_store
from .hotpath import FrameTable, StepCfg, Workspace, make_camera, tracking_iteration
from .renderer import make_cfg


def _strict_default() -> bool:
    return os.environ.get("ESLAM_B200_STRICT_RNG", "0") == "1"


def _tracker_state(trk, batch_size):
    """Workspace + step configuration cached on the tracker object (built from the attributes the
    reference's Tracker.__init__ sets, Tracker.py:44-112)."""
    st = getattr(trk, "_b200", None)
    rnd = trk.renderer
    S = rnd.n_stratified + rnd.n_importance
    if st is None or not st["ws"].fits(batch_size, S, 1):
        cam = make_camera(trk.H, trk.W, trk.fx, trk.fy, trk.cx, trk.cy, trk.ignore_edge_H, trk.H - trk.ignore_edge_H,
                          trk.ignore_edge_W, trk.W - trk.ignore_edge_W)
        rc = make_cfg(rnd.n_stratified, rnd.n_importance, trk.truncation,
                      (trk.w_sdf_fs, trk.w_sdf_center, trk.w_sdf_tail, trk.w_depth, trk.w_color))
        st = {"ws": Workspace(trk.device, batch_size, S, 1), "sc": StepCfg(cam, rc, bool(rnd.perturb)),
              "pulled_idx": None}
        trk._b200 = st
    return st


def _tracker_store(trk, st):
    all_planes = (trk.planes_xy, trk.planes_xz, trk.planes_yz, trk.c_planes_xy, trk.c_planes_xz, trk.c_planes_yz)
    store = synced_store(all_planes, trk.decoders, trk.bound)
    # the planes alias memory another PROCESS (the mapper) writes (Tracker.py:222-232): version counters
    # cannot see that, so re-import whenever the tracker refreshed its parameters from the mapper
    pm = getattr(trk, "prev_mapping_idx", None)
    pm = int(pm) if pm is not None else None
    if st["pulled_idx"] != pm:
        store.pull_planes(all_planes)
        st["pulled_idx"] = pm
    store.ensure_q()  # outside any graph: a replayed frame graph reads the Q images, it does not rebuild them
    return store


class _IterGraphs:
    """One CUDA graph per Adam step index for the per-iteration drop-in (`optimize_tracking`): draws + ray sampling +
    forward + outlier mask + backward + loss + pose Adam, i.e. one replay instead of ~10 launches per iteration.
    Persistent buffers the graphs address: the [1,7] pose, the two [1,7] moment rows (lent to the caller's optimizer
    as its own state, see _fused_adam_plan) and a one-entry frame table."""

    def __init__(self, dev):
        self.pose = torch.zeros(1, 7, dtype=torch.float32, device=dev)
        self.m7 = torch.zeros(1, 7, dtype=torch.float32, device=dev)
        self.v7 = torch.zeros(1, 7, dtype=torch.float32, device=dev)
        self.frames = None
        self.frame_key = None
        self.graphs = {}
        self.owner = None  # weakref to the optimizer whose state currently lives in m7 / v7

    def lend(self, optimizer):
        """The moment rows for a NEW optimizer state, or None if another live optimizer still uses them."""
        import weakref

        cur = self.owner() if self.owner is not None else None
        if cur is not None and cur is not optimizer:
            return None
        self.owner = weakref.ref(optimizer)
        self.m7.zero_()
        self.v7.zero_()
        return self.m7, self.v7

    def set_frame(self, gt_color, gt_depth, cam, dev):
        c, d = gt_color[0], gt_depth[0]
        key = (c.data_ptr(), d.data_ptr())
        if self.frames is None:
            self.frames = FrameTable([c], [d], cam, dev)
        elif key != self.frame_key:
            self.frames.colors, self.frames.depths = [c], [d]
            self.frames.table.copy_(torch.tensor([[d.data_ptr()], [c.data_ptr()]], dtype=torch.int64))
        self.frame_key = key


def _fused_adam_plan(cam_pose, optimizer, ig=None):
    """If `cam_pose` is torch.cat([R, T], -1) of two leaf parameters R [1,4] and T [1,3] that are the only parameters
    of a plain torch.optim.Adam (what Tracker.run builds, Tracker.py:282-299), return what is needed to take the
    optimizer's step with the fused pose-Adam kernel ON THE OPTIMIZER'S OWN STATE; otherwise None (generic path:
    autograd + optimizer.step())."""
    fn = cam_pose.grad_fn
    if type(optimizer) is not torch.optim.Adam or fn is None or type(fn).__name__ != "CatBackward0":
        return None
    nxt = fn.next_functions
    if len(nxt) != 2 or any(n[0] is None or not hasattr(n[0], "variable") for n in nxt):
        return None
    Rq, T = nxt[0][0].variable, nxt[1][0].variable
    if tuple(Rq.shape) != (1, 4) or tuple(T.shape) != (1, 3) or tuple(cam_pose.shape) != (1, 7):
        return None
    groups = optimizer.param_groups
    if len(groups) != 2 or any(len(g["params"]) != 1 for g in groups):
        return None
    by_param = {id(g["params"][0]): g for g in groups}
    if id(Rq) not in by_param or id(T) not in by_param:
        return None
    gR, gT = by_param[id(Rq)], by_param[id(T)]
    for g in (gR, gT):
        if g.get("amsgrad") or g.get("maximize") or g.get("weight_decay", 0) != 0 or g.get("capturable") \
                or g.get("differentiable") or g.get("fused"):
            return None
    if gR["betas"] != gT["betas"] or gR["eps"] != gT["eps"]:
        return None
    sR, sT = optimizer.state[Rq], optimizer.state[T]
    if len(sR) == 0 and len(sT) == 0:
        # first step: create the optimizer's state the way torch does, with both parameters' moments in ONE [1,7]
        # row each so the kernel can address them as a pose
        lent = ig.lend(optimizer) if ig is not None else None
        if lent is not None:
            m7, v7 = lent
        else:
            m7 = torch.zeros(1, 7, dtype=torch.float32, device=Rq.device)
            v7 = torch.zeros(1, 7, dtype=torch.float32, device=Rq.device)
        for st_, sl in ((sR, slice(0, 4)), (sT, slice(4, 7))):
            st_["step"] = torch.tensor(0.0, dtype=torch.float32)
            st_["exp_avg"] = m7[:, sl]
            st_["exp_avg_sq"] = v7[:, sl]
    try:
        mR, mT, vR, vT = sR["exp_avg"], sT["exp_avg"], sR["exp_avg_sq"], sT["exp_avg_sq"]
    except KeyError:
        return None
    adjacent = (mT.data_ptr() == mR.data_ptr() + 16 and vT.data_ptr() == vR.data_ptr() + 16
                and mR.dtype == torch.float32 and mR.device == Rq.device and float(sR["step"]) == float(sT["step"]))
    if not adjacent:
        return None  # state created elsewhere: leave it to torch
    return Rq, T, gR, gT, sR, sT


def optimize_tracking(self, cam_pose, gt_color, gt_depth, batch_size, optimizer):
    """One iteration of camera tracking (reference Tracker.optimize_tracking, Tracker.py:150-210):
    sample pixels, render, losses, backward to the 7-dof pose, step the caller's optimizer.
    Returns the loss as a python float.

    When the caller's optimizer is the plain Adam over (R, T) that Tracker.run builds and `cam_pose` is their
    concatenation, the step is taken by the fused pose-Adam kernel directly on the optimizer's state (same update
    rule, `.grad` of both parameters set, step counters advanced) instead of through autograd and ~30 ATen launches;
    anything else goes through `cam_pose.backward` + `optimizer.step()`."""
    st = _tracker_state(self, batch_size)
    store = _tracker_store(self, st)
    ws = st["ws"]
    fused = os.environ.get("ESLAM_B200_FUSED_OPT", "1") == "1"
    graphs_on = fused and _use_graph(self) and st["sc"].perturb
    ig = st.get("iter_graphs")
    if graphs_on and ig is None:
        ig = st["iter_graphs"] = _IterGraphs(ws.device)
    plan = _fused_adam_plan(cam_pose, optimizer, ig if graphs_on else None) if fused else None
    draws, strict = getattr(self, "draws", None), getattr(self, "strict_rng", _strict_default())
    if plan is None:
        pose7 = cam_pose.detach().float().contiguous()
        tracking_iteration(ws, store, st["sc"], pose7, gt_color, gt_depth, batch_size, draws=draws, strict_rng=strict)
        optimizer.zero_grad()
        cam_pose.backward(ws.grad7[0:1].clone())
        optimizer.step()
        return ws.loss_acc[5].item()
    Rq, T, gR, gT, sR, sT = plan
    step = int(float(sR["step"])) + 1
    adam = {"step": step, "lr_q": float(gR["lr"]), "lr_t": float(gT["lr"]), "m": sR["exp_avg"], "v": sR["exp_avg_sq"],
            "betas": tuple(gR["betas"]), "eps": float(gR["eps"])}
    replay = (graphs_on and st.get("eager_done") and sR["exp_avg"].data_ptr() == ig.m7.data_ptr()
              and sR["exp_avg_sq"].data_ptr() == ig.v7.data_ptr())
    if replay:
        # one graph per Adam step index (the bias corrections are host-side constants of the launch)
        _check_track_frames(gt_color, gt_depth, st["sc"].cam)
        ig.set_frame(gt_color, gt_depth, st["sc"].cam, ws.device)
        ig.pose.copy_(cam_pose.detach().float().reshape(1, 7))
        key = (step, adam["lr_q"], adam["lr_t"], adam["betas"], adam["eps"], batch_size, store.arena.data_ptr(), id(ws))
        ent = ig.graphs.get(key)
        if ent is None:
            from . import _lib

            g = torch.cuda.CUDAGraph()
            l0 = _lib.LAUNCHES
            with torch.cuda.graph(g):
                tracking_iteration(ws, store, st["sc"], ig.pose, ig.frames, ig.frames, batch_size,
                                   apply_adam={**adam, "m": ig.m7, "v": ig.v7})
            ent = ig.graphs[key] = (g, _lib.LAUNCHES - l0)
        ent[0].replay()
        from . import _lib

        _lib.LAUNCHES += ent[1]
        pose7 = ig.pose
    else:
        pose7 = cam_pose.detach().float().clone().contiguous()
        tracking_iteration(ws, store, st["sc"], pose7, gt_color, gt_depth, batch_size, draws=draws, strict_rng=strict,
                           apply_adam=adam)
        st["eager_done"] = True
    g7 = ws.grad7[0:1].clone()
    with torch.no_grad():
        Rq.copy_(pose7[:, :4])
        T.copy_(pose7[:, 4:])
    Rq.grad, T.grad = g7[:, :4], g7[:, 4:]
    sR["step"] += 1
    sT["step"] += 1
    return ws.loss_acc[5].item()


def _check_track_frames(gt_color, gt_depth, cam):
    if gt_depth.dtype != torch.float32 or gt_color.dtype != torch.float64:
        raise RuntimeError("gt depth must be float32 and gt colour float64, as the reference's datasets produce "
                           "(src/utils/datasets.py:90-92)")
    if tuple(gt_depth.shape) != (1, cam.H, cam.W) or tuple(gt_color.shape) != (1, cam.H, cam.W, 3):
        raise RuntimeError(f"frame shapes {tuple(gt_depth.shape)} / {tuple(gt_color.shape)} do not match the camera")
    if not (gt_depth.is_cuda and gt_color.is_cuda and gt_depth.is_contiguous() and gt_color.is_contiguous()):
        raise RuntimeError("frames must be contiguous CUDA tensors; there is no CPU path")


class _FrameGraph:
    """The `iters` iterations of one tracked frame captured once as a CUDA graph (launch-bound inner loop: ~10
    small launches per iteration, each a ctypes call).  Everything the kernels address is persistent: the workspace,
    the parameter arena, a pose / loss / trace buffer, and a one-entry frame table whose two pointers are rewritten
    before every replay.  The random draws are torch's (its CUDA generator is graph safe)."""

    def __init__(self, trk, st, store, init_pose, gt_color, gt_depth, iters, batch_size, lr_T, lr_R):
        ws, sc = st["ws"], st["sc"]
        dev = ws.device
        self.key = (iters, batch_size, lr_T, lr_R, store.arena.data_ptr(), id(ws))
        self.pose = init_pose.detach().float().reshape(1, 7).clone().contiguous()  # a valid pose for the warm-up pass
        self.losses = torch.zeros(iters, dtype=torch.float64, device=dev)
        self.trace = torch.zeros(iters, 7, dtype=torch.float32, device=dev)
        self.frames = FrameTable([gt_color[0]], [gt_depth[0]], sc.cam, dev)

        def body():
            ws.pose_m.zero_()
            ws.pose_v.zero_()
            store.bind()
            for it in range(iters):
                self.trace[it].copy_(self.pose[0])
                tracking_iteration(ws, store, sc, self.pose, self.frames, self.frames, batch_size,
                                   apply_adam={"step": it + 1, "lr_q": lr_R, "lr_t": lr_T}, bind=False)
                self.losses[it].copy_(ws.loss_acc[5])

        side = torch.cuda.Stream(device=dev)  # one eager pass first: lazy initialisations must not be captured
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize(dev)
        from . import _lib

        self.graph = torch.cuda.CUDAGraph()
        l0 = _lib.LAUNCHES
        with torch.cuda.graph(self.graph):
            body()
        self.n_launches = _lib.LAUNCHES - l0  # kernels of this library inside the graph (bench.py's gpu_launches)

    def run(self, init_pose, gt_color, gt_depth):
        c, d = gt_color[0], gt_depth[0]
        self.frames.colors, self.frames.depths = [c], [d]  # keep the frame alive while the graph reads it
        self.frames.table.copy_(torch.tensor([[d.data_ptr()], [c.data_ptr()]], dtype=torch.int64))  # 16 bytes, staged
        self.pose.copy_(init_pose.detach().float().reshape(1, 7))
        self.graph.replay()
        from . import _lib

        _lib.LAUNCHES += self.n_launches
        return self.trace, self.losses, self.pose


def _use_graph(trk) -> bool:
    return (os.environ.get("ESLAM_B200_GRAPH", "1") == "1" and getattr(trk, "draws", None) is None
            and not getattr(trk, "strict_rng", _strict_default()))


def track_frame(self, init_pose, gt_color, gt_depth, iters=None, batch_size=None, lr_T=None, lr_R=None):
    """The camera loop of Tracker.run for one frame (Tracker.py:291-309) without leaving the device:
    `iters` iterations with the fused Adam(betas=(0.5,0.999)) on (R,T); returns
    (candidate_pose[1,7] = pose before the lowest-loss step, losses[iters] float64 tensor, final pose).
    One host sync at the end instead of one per iteration; the iterations replay as one CUDA graph
    (ESLAM_B200_GRAPH=0, injected draws or strict_rng fall back to launching them one by one)."""
    iters = self.num_cam_iters if iters is None else iters
    batch_size = self.tracking_pixels if batch_size is None else batch_size
    lr_T = self.cam_lr_T if lr_T is None else lr_T
    lr_R = self.cam_lr_R if lr_R is None else lr_R
    st = _tracker_state(self, batch_size)
    store = _tracker_store(self, st)
    ws = st["ws"]
    if _use_graph(self) and st["sc"].perturb:
        key = (iters, batch_size, lr_T, lr_R, store.arena.data_ptr(), id(ws))
        fg = st.get("graph")
        if fg is None or fg.key != key:
            _check_track_frames(gt_color, gt_depth, st["sc"].cam)
            fg = _FrameGraph(self, st, store, init_pose, gt_color, gt_depth, iters, batch_size, lr_T, lr_R)
            st["graph"] = fg
        _check_track_frames(gt_color, gt_depth, st["sc"].cam)
        trace, losses, pose = fg.run(init_pose, gt_color, gt_depth)
        best = torch.argmin(torch.nan_to_num(losses, nan=float("inf")))
        return trace[best][None].clone(), losses.clone(), pose.clone()
    pose = init_pose.detach().float().contiguous().clone()
    ws.pose_m.zero_()
    ws.pose_v.zero_()
    dev = ws.device
    losses = torch.empty(iters, dtype=torch.float64, device=dev)
    trace = torch.empty(iters, 7, dtype=torch.float32, device=dev)
    for it in range(iters):
        trace[it].copy_(pose[0])
        tracking_iteration(ws, store, st["sc"], pose, gt_color, gt_depth, batch_size,
                           draws=getattr(self, "draws", None),
                           strict_rng=getattr(self, "strict_rng", _strict_default()),
                           apply_adam={"step": it + 1, "lr_q": lr_R, "lr_t": lr_T})
        losses[it] = ws.loss_acc[5]
    # first minimum, NaN losses never win (`loss < current_min_loss`, Tracker.py:305)
    best = torch.argmin(torch.nan_to_num(losses, nan=float("inf")))
    return trace[best][None].clone(), losses, pose


class TrackerStep:
    """Standalone holder of exactly the attributes `optimize_tracking` reads, for use without the
    reference's Tracker (tests, bench)."""

    def __init__(self, cfg, renderer, decoders, all_planes, bound, cam, device):
        (self.planes_xy, self.planes_xz, self.planes_yz, self.c_planes_xy, self.c_planes_xz, self.c_planes_yz) = all_planes
        self.renderer, self.decoders, self.bound, self.device = renderer, decoders, bound, device
        self.H, self.W, self.fx, self.fy, self.cx, self.cy = cam
        t = cfg['tracking']
        self.truncation = cfg['model']['truncation']
        self.ignore_edge_H, self.ignore_edge_W = t['ignore_edge_H'], t['ignore_edge_W']
        self.w_sdf_fs, self.w_sdf_center, self.w_sdf_tail = t['w_sdf_fs'], t['w_sdf_center'], t['w_sdf_tail']
        self.w_depth, self.w_color = t['w_depth'], t['w_color']
        self.cam_lr_T, self.cam_lr_R = t['lr_T'], t['lr_R']
        self.num_cam_iters, self.tracking_pixels = t['iters'], t['pixels']
        self.prev_mapping_idx = -1

    optimize_tracking = optimize_tracking
    track_frame = track_frame
