"""The fused per-iteration render-and-optimise step as a short
sequence of kernel launches on the current stream, with no host synchronisation in the default
mode:

  tracking:  [q_build when the map changed] -> sample_rays -> render_forward_q -> track_mask -> pose_backward_q
             -> finalize -> pose Adam
  mapping:   q_build -> sample_rays -> importance_samples -> loss_backward_q(gradient images, decoders, poses)
             -> q_adam_planes -> decoder Adam -> pose Adam

Both loops run on the Q images of the map: the first (linear) decoder layer is applied to the planes once per
parameter change instead of to every sample (decoders.py:87-125 commutes with the bilinear fetch of :64-85).

`strict_rng=True` reproduces the reference's random-draw SHAPES ([R1,S], [R0,n_strat], [R0,n_imp]),
which costs one host sync per iteration (the reference has >= 12); it is what parity runs use.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Optional

import torch

from ._lib import COUNTER_WORDS, N_COUNTERS, N_LOSS, Camera, RenderCfg, call, on_stream, ptr, stream
from .field import FieldStore
from .renderer import TorchDraws, linspace_table, make_cfg


def make_camera(H, W, fx, fy, cx, cy, H0=0, H1=None, W0=0, W1=None) -> Camera:
    c = Camera()
    c.H, c.W = int(H), int(W)
    c.fx, c.fy, c.cx, c.cy = float(fx), float(fy), float(cx), float(cy)
    c.H0, c.H1 = int(H0), int(H if H1 is None else H1)
    c.W0, c.W1 = int(W0), int(W if W1 is None else W1)
    return c


class Workspace:
    """Per-process scratch for up to `max_rays` rays of `n_samples` samples and `max_frames` frames."""

    def __init__(self, device, max_rays: int, n_samples: int, max_frames: int = 32):
        dev = torch.device(device)
        N, S = int(max_rays), int(n_samples)
        f32, i32 = torch.float32, torch.int32
        self.device, self.max_rays, self.n_samples, self.max_frames = dev, N, S, max_frames
        self.rays_o = torch.empty(N, 3, dtype=f32, device=dev)
        self.rays_d = torch.empty(N, 3, dtype=f32, device=dev)
        self.gt_depth = torch.empty(N, dtype=f32, device=dev)
        self.gt_color = torch.empty(N, 3, dtype=torch.float64, device=dev)
        self.src = torch.empty(N, dtype=i32, device=dev)
        self.z = torch.empty(N, S, dtype=f32, device=dev)
        self.dl_list = torch.empty(N, dtype=i32, device=dev)
        self.zord = torch.empty(N, dtype=i32, device=dev)
        self.band = torch.empty(N, 4, dtype=torch.uint8, device=dev)
        self._counter_buf = torch.zeros(COUNTER_WORDS, dtype=i32, device=dev)  # counters + per-CTA look-back words
        self.counters = self._counter_buf[:N_COUNTERS]
        self.depth = torch.empty(N, dtype=f32, device=dev)
        self.rgb = torch.empty(N, 3, dtype=f32, device=dev)
        self.sdf = torch.empty(N, S, dtype=f32, device=dev)
        self.act4 = torch.empty(N, S, 4, dtype=f32, device=dev)   # forward activations kept for the tracker's
        self.actm = torch.empty(N, S, dtype=i32, device=dev)      # pose-only backward (eslam_render_forward_act)
        self.ray_mask = torch.empty(N, dtype=torch.uint8, device=dev)
        self.scratch = torch.empty(N + 1, dtype=f32, device=dev)
        self.loss_acc = torch.zeros(N_LOSS, dtype=torch.float64, device=dev)
        self.loss_out = torch.zeros(1, dtype=f32, device=dev)
        self.pose_grad = torch.zeros(max_frames, 12, dtype=f32, device=dev)
        self.grad7 = torch.zeros(max_frames, 7, dtype=f32, device=dev)
        self.pose_m = torch.zeros(max_frames, 7, dtype=f32, device=dev)
        self.pose_v = torch.zeros(max_frames, 7, dtype=f32, device=dev)
        self.c2w_out = torch.zeros(max_frames, 16, dtype=f32, device=dev)

    def fits(self, n_rays, n_samples, n_frames):
        return n_rays <= self.max_rays and n_samples == self.n_samples and n_frames <= self.max_frames


@dataclass
class StepCfg:
    """What one iteration needs besides the tensors: camera+crop, sampling and loss configuration."""
    cam: Camera
    render: RenderCfg
    perturb: bool = True


class FrameTable:
    """The frames of a mapping window, read by the sampling kernel WHERE THEY LIVE through a device table of
    per-frame pointers, instead of the reference's torch.stack of up to 20 full frames per call
    (Mapper.py:268-286: 457 MB of copies at Replica size).  colors[k]: [H,W,3] float64, depths[k]: [H,W] float32,
    contiguous CUDA tensors; the table keeps them alive."""

    def __init__(self, colors, depths, cam: Camera, device):
        if len(colors) != len(depths) or len(depths) == 0:
            raise RuntimeError("FrameTable: need the same (non-zero) number of colour and depth frames")
        dev = torch.device(device)
        for c, d in zip(colors, depths):
            if not (c.is_cuda and d.is_cuda) or c.device != d.device:
                raise RuntimeError("FrameTable: frames must be CUDA tensors of one device; there is no CPU path")
            if d.dtype != torch.float32 or c.dtype != torch.float64:
                raise RuntimeError("gt depth must be float32 and gt colour float64, as the reference's datasets "
                                   "produce (src/utils/datasets.py:90-92)")
            if tuple(d.shape) != (cam.H, cam.W) or tuple(c.shape) != (cam.H, cam.W, 3):
                raise RuntimeError(f"frame shapes {tuple(d.shape)} / {tuple(c.shape)} do not match the camera")
            if not (d.is_contiguous() and c.is_contiguous()):
                raise RuntimeError("frames must be contiguous")
        self.colors, self.depths, self.n = list(colors), list(depths), len(depths)
        host = torch.tensor([[d.data_ptr() for d in depths], [c.data_ptr() for c in colors]], dtype=torch.int64)
        self.table = host.to(dev)  # [2][n] device pointers


def _sample(ws: Workspace, store: FieldStore, sc: StepCfg, idx, n_img, n_per_img, c2w, poses, pose_first, depth,
            color, u_depth, need_depth):
    dev = ws.device
    t_uni = linspace_table(sc.render.n_stratified, dev)
    t_surf = linspace_table(sc.render.n_importance, dev)
    entry = "eslam_sample_rays"
    if isinstance(depth, FrameTable):
        entry, color, depth = "eslam_sample_rays_frames", depth.table[1], depth.table[0]
    call(entry, store.ref(), C.byref(sc.cam), C.byref(sc.render), ptr(idx), n_img, n_per_img, ptr(c2w),
         ptr(poses), pose_first, ptr(depth), ptr(color), ptr(u_depth), ptr(t_uni), ptr(t_surf), need_depth,
         ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.gt_depth), ptr(ws.gt_color), ptr(ws.src), ptr(ws.z),
         ptr(ws.dl_list), ptr(ws.zord), ptr(ws.band), ptr(ws.counters), ptr(ws.c2w_out), stream())


def _check_frames(depth, color, n_img, cam):
    if isinstance(depth, FrameTable):
        if color is not depth or depth.n != n_img:
            raise RuntimeError("pass the same FrameTable as gt_colors and gt_depths, one frame per camera")
        return
    if depth.dtype != torch.float32 or color.dtype != torch.float64:
        raise RuntimeError("gt depth must be float32 and gt colour float64, as the reference's datasets produce "
                           "(src/utils/datasets.py:90-92)")
    if tuple(depth.shape) != (n_img, cam.H, cam.W) or tuple(color.shape) != (n_img, cam.H, cam.W, 3):
        raise RuntimeError(f"frame stack shapes {tuple(depth.shape)} / {tuple(color.shape)} do not match the camera")
    if not (depth.is_contiguous() and color.is_contiguous()):
        raise RuntimeError("frame stacks must be contiguous")


def tracking_iteration(ws: Workspace, store: FieldStore, sc: StepCfg, pose7: torch.Tensor, gt_color, gt_depth,
                       n_pixels: int, draws=None, strict_rng: bool = False, apply_adam: Optional[dict] = None,
                       bind: bool = True):
    """One iteration of Tracker.optimize_tracking (Tracker.py:150-210) up to (and optionally including)
    the Adam step.  pose7: [1,7] contiguous fp32 device tensor (quaternion, translation).
    After the call: ws.loss_acc[5] = loss (float64), ws.grad7[0] = d loss / d pose.
    apply_adam: dict(step=, lr_q=, lr_t=[, m=, v=, betas=, eps=]) -> fused Adam (default betas .5,.999) on pose7 in
    place, on the workspace's moments or on the given [1,7] rows.
    The render kernels read the Q images of the map (FieldStore.ensure_q: rebuilt only when the parameters changed,
    i.e. when the mapper published new planes; the decoders are frozen while tracking, Tracker.py:111-112)."""
    dev = ws.device
    draws = draws or TorchDraws(dev)
    cam, rc = sc.cam, sc.render
    S = rc.n_stratified + rc.n_importance
    n_crop = (cam.H1 - cam.H0) * (cam.W1 - cam.W0)
    _check_frames(gt_depth, gt_color, 1, cam)
    idx = draws.randint(n_crop, n_pixels)
    if bind:  # the decoders are frozen while a frame is tracked: a caller looping over iterations binds once
        store.bind()
    q = store.ensure_q()
    if not sc.perturb:
        u = None
    elif strict_rng:
        _sample(ws, store, sc, idx, 1, n_pixels, None, pose7, 0, gt_depth, gt_color, None, 1)
        u = draws.rand(int(ws.counters[0].item()), S)
    else:
        u = draws.rand(n_pixels, S)
    _sample(ws, store, sc, idx, 1, n_pixels, None, pose7, 0, gt_depth, gt_color, u, 1)
    N = n_pixels
    # the forward pass (needed first: the outlier mask is a median over the rendered depth, Tracker.py:192-195)
    # keeps sdf, rgb and the ReLU masks of every sample, so the backward pass neither re-runs the forward MLPs nor
    # needs anything but the Q corners for the coordinate gradients
    call("eslam_render_forward_q", store.ref(), ptr(q), ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.z), N, S,
         ptr(ws.counters), ptr(ws.depth), ptr(ws.rgb), ptr(ws.sdf), ptr(ws.act4), ptr(ws.actm), stream())
    call("eslam_track_mask", ptr(ws.gt_depth), ptr(ws.depth), ptr(ws.band), N, ptr(ws.counters), ptr(ws.ray_mask),
         ptr(ws.scratch), stream())
    call("eslam_pose_backward_q", store.ref(), ptr(store.arena), ptr(q), C.byref(cam), C.byref(rc), ptr(ws.rays_o),
         ptr(ws.rays_d), ptr(ws.z), ptr(ws.gt_depth), ptr(ws.gt_color), ptr(ws.src), ptr(idx), n_pixels,
         ptr(ws.ray_mask), ptr(ws.counters), N, ptr(ws.sdf), ptr(ws.act4), ptr(ws.actm), ptr(ws.pose_grad),
         ptr(ws.loss_acc), stream())
    call("eslam_finalize_loss", C.byref(rc), ptr(ws.counters), 1, ptr(ws.loss_acc), ptr(ws.loss_out), stream())
    if apply_adam is None:
        call("eslam_pose_adam_step", ptr(pose7), ptr(ws.pose_grad), None, None, 1, 0, 0.0, 0.0, 1, 0.5, 0.999, 1e-8,
             ptr(ws.grad7), 0, stream())
    else:
        # moments: the workspace's, or the caller's (the rows of a torch.optim.Adam's own state, tracker.py)
        m, v = apply_adam.get("m", ws.pose_m), apply_adam.get("v", ws.pose_v)
        b1, b2 = apply_adam.get("betas", (0.5, 0.999))
        call("eslam_pose_adam_step", ptr(pose7), ptr(ws.pose_grad), ptr(m), ptr(v), 1, 0, apply_adam["lr_q"],
             apply_adam["lr_t"], apply_adam["step"], b1, b2, apply_adam.get("eps", 1e-8), ptr(ws.grad7), 1, stream())


def mapping_iteration(ws: Workspace, store: FieldStore, sc: StepCfg, c2ws, poses7, gt_colors, gt_depths,
                      pix_per_image: int, step: int, lr_dec: float, lr_planes: float, lr_cplanes: float,
                      lr_cam: float, draws=None, strict_rng: bool = False, want_loss: bool = False,
                      apply_adam: bool = True, reduce_counters=None, reduce_grads=None, exchange=None):
    """One iteration of the loop in Mapper.optimize_mapping (Mapper.py:308-350).
    c2ws [b,4,4] fp32; poses7 [b,7] or None (joint_opt: frames 1.. are taken from poses7 and updated).
    gt_colors / gt_depths: stacked [b,H,W,3] f64 / [b,H,W] f32 tensors, or one FrameTable passed as both.
    Planes/decoders live in `store` (Adam state in store.exp_avg*, reset by the caller per call).

    The backward kernel works on the Q images of the map and reduces the plane gradients as 16-channel GRADIENT
    IMAGES (store.gq_arena); the optimiser step (FieldStore.adam_step_q) turns them into plane gradients and dW1 on
    the fly.  With apply_adam=False (tests, gradient inspection; the caller has called store.reset_adam()) the whole
    gradient is left in store.grad in parameter form (FieldStore.parameter_grads) next to the gradient images.

    reduce_counters(counters)->norm and reduce_grads(tensors, pose_grad, loss_acc) are the two exchange
    points of the ray-sharded multi-GPU mapping (myslam_b200.dist); None on one GPU.  `exchange` is an object
    providing both; a dist.PeerExchange additionally replaces all-reduce + Adam by the fused peer-memory kernels."""
    dev = ws.device
    if exchange is not None:
        reduce_counters, reduce_grads = exchange.reduce_counters, exchange.reduce_grads
    fused_exchange = apply_adam and exchange is not None and hasattr(exchange, "adam_exchange")
    draws = draws or TorchDraws(dev)
    cam, rc = sc.cam, sc.render
    ns, ni = rc.n_stratified, rc.n_importance
    S = ns + ni
    b = c2ws.shape[0]
    N = pix_per_image * b
    _check_frames(gt_depths, gt_colors, b, cam)
    n_crop = (cam.H1 - cam.H0) * (cam.W1 - cam.W0)
    idx = draws.randint(n_crop, N)
    q = store.ensure_q()  # nothing in this loop reads the constant-bank decoders (they change every iteration)
    c2w_flat = c2ws.reshape(b, 16).float().contiguous()
    joint = poses7 is not None
    if strict_rng:
        _sample(ws, store, sc, idx, b, pix_per_image, c2w_flat, poses7, 1, gt_depths, gt_colors, None, 0)
        cnt = ws.counters[:2].tolist()
        r0 = cnt[1]
        u = draws.rand(cnt[0] - r0, S) if sc.perturb else None
    else:
        r0 = N
        u = u_c = u_f = None
        if sc.perturb and hasattr(draws, "rand_many"):  # one generator launch for the three uniform blocks
            u, u_c, u_f = draws.rand_many([(N, S), (N, ns), (N, ni)])
        elif sc.perturb:
            u = draws.rand(N, S)
    _sample(ws, store, sc, idx, b, pix_per_image, c2w_flat, poses7, 1, gt_depths, gt_colors, u, 0)
    norm, overlapped = None, exchange is not None and hasattr(exchange, "begin_counters")
    if overlapped:  # the counters are final here; their exchange overlaps the importance sampling below
        norm = exchange.begin_counters(ws.counters)
    if r0 > 0:
        if strict_rng or u_c is None:
            u_c = draws.rand(r0, ns)
            u_f = draws.rand(r0, ni)
        call("eslam_importance_samples", store.ref(), ptr(store.arena), ptr(q), C.byref(rc), ptr(ws.rays_o), ptr(ws.rays_d),
             ptr(ws.dl_list), ptr(ws.counters), r0, ptr(u_c), ptr(u_f), ptr(linspace_table(ns, dev)), ptr(ws.z),
             stream())
    grad, gq = store.ensure_grad(), store.ensure_q_grad()
    if overlapped:
        exchange.end_counters()
    elif reduce_counters is not None:
        norm = reduce_counters(ws.counters)
    call("eslam_loss_backward_q", store.ref(), ptr(store.arena), ptr(q), ptr(gq), C.byref(cam), C.byref(rc),
         ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.z), ptr(ws.gt_depth), ptr(ws.gt_color), ptr(ws.src), ptr(idx),
         pix_per_image, None, ptr(ws.counters), ptr(norm) if norm is not None else None, N, ptr(grad),
         ptr(ws.pose_grad) if joint else None, ptr(ws.loss_acc) if want_loss else None, stream())
    if fused_exchange:
        # reduce-scatter of the gradient images + plane Adam + all-gather, then the decoders' replicated step, as
        # kernels over peer memory (csrc/exchange.cuh)
        pose_sum, loss_sum = exchange.adam_exchange(step, lr_dec, lr_planes, lr_cplanes, ws.pose_grad if joint else None,
                                                    b, ws.loss_acc if want_loss else None)
        if want_loss:
            call("eslam_finalize_loss", C.byref(rc), ptr(norm), 0, ptr(loss_sum), ptr(ws.loss_out), stream())
            ws.loss_acc[5:7].copy_(loss_sum[5:7])
        if joint:
            call("eslam_pose_adam_step", ptr(poses7), ptr(pose_sum), ptr(ws.pose_m), ptr(ws.pose_v), b, 1, lr_cam,
                 lr_cam, step, 0.9, 0.999, 1e-8, ptr(ws.grad7), 1, stream())
        return
    if reduce_grads is not None:
        reduce_grads([gq, grad[store.dec_off:]], ws.pose_grad if joint else None, ws.loss_acc if want_loss else None)
    if want_loss:
        call("eslam_finalize_loss", C.byref(rc), ptr(norm if norm is not None else ws.counters), 0, ptr(ws.loss_acc),
             ptr(ws.loss_out), stream())
    if not apply_adam:
        store.grad.copy_(store.parameter_grads())  # inspection only: the whole gradient in parameter form
        if joint:
            call("eslam_pose_adam_step", ptr(poses7), ptr(ws.pose_grad), None, None, b, 1, 0.0, 0.0, 1, 0.9, 0.999,
                 1e-8, ptr(ws.grad7), 0, stream())
        return
    store.adam_step_q(step, lr_dec, lr_planes, lr_cplanes)
    if exchange is not None and hasattr(exchange, "after_step"):
        exchange.after_step(store)
    if joint:
        call("eslam_pose_adam_step", ptr(poses7), ptr(ws.pose_grad), ptr(ws.pose_m), ptr(ws.pose_v), b, 1, lr_cam,
             lr_cam, step, 0.9, 0.999, 1e-8, ptr(ws.grad7), 1, stream())


class _Pipe:
    """Side stream, events and the persistent random-draw buffers of the pipelined window loop: one per Workspace and
    batch shape, never replaced (a captured window graph addresses these buffers)."""

    def __init__(self, ws: Workspace, N: int, S: int, ns: int, ni: int):
        dev = ws.device
        self.key = (N, S, ns, ni)
        side = getattr(ws, "_side_stream", None)
        if side is None:
            side = ws._side_stream = torch.cuda.Stream(device=dev)
        self.side = side
        imp = getattr(ws, "_imp_stream", None)
        if imp is None:
            imp = ws._imp_stream = torch.cuda.Stream(device=dev)
        self.imp = imp
        self.ev_q, self.ev_imp = torch.cuda.Event(), torch.cuda.Event()
        self.ev_prep, self.ev_bwd, self.ev_side = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self.idx = torch.empty(N, dtype=torch.int64, device=dev)
        self.ubuf = torch.empty(N * (S + ns + ni), dtype=torch.float32, device=dev)
        self.u = self.ubuf[:N * S].view(N, S)
        self.u_c = self.ubuf[N * S:N * (S + ns)].view(N, ns)
        self.u_f = self.ubuf[N * (S + ns):].view(N, ni)


def mapping_window_pipelined(ws: Workspace, store: FieldStore, sc: StepCfg, c2ws, poses7, gt_colors, gt_depths,
                             pix_per_image: int, iters: int, lr_dec: float, lr_planes: float, lr_cplanes: float,
                             lr_cam: float, exchange=None):
    """`iters` iterations of Mapper.optimize_mapping's loop (Mapper.py:308-350) on the DEFAULT path (torch's generator,
    fixed-shape draws, no host sync), software-pipelined over two streams.  What an iteration's ray sampling needs of
    the previous one is only the poses, so per iteration

        main   importance(it) -> backward(it) ->  plane optimiser tail / peer exchange(it) -> decoder step -> Q(it+1)
        side                      [after backward]  pose sums + pose Adam(it) -> draws(it+1) -> ray sampling(it+1)
                                                    (-> normaliser exchange(it+1))

    and the main stream waits for the side stream's sampling only when it reaches importance(it+1).  The backward is a
    pair of launches (eslam_loss_backward_q_part; ESLAM_B200_SPLIT_BWD=0: one launch behind the importance pass): the
    tiles whose rays all carry a depth start while the importance kernel runs, the tiles with a depth-less ray follow
    that kernel on a third stream.  The draws live in
    persistent buffers (filled in place), so nothing is allocated on the side stream.  Same arithmetic as
    mapping_iteration; the uniforms are consumed in the same [N,S] | [N,n_strat] | [N,n_imp] blocks."""
    dev = ws.device
    cam, rc = sc.cam, sc.render
    ns, ni = rc.n_stratified, rc.n_importance
    S = ns + ni
    b = c2ws.shape[0]
    N = pix_per_image * b
    _check_frames(gt_depths, gt_colors, b, cam)
    n_crop = (cam.H1 - cam.H0) * (cam.W1 - cam.W0)
    pipes = ws.__dict__.setdefault("_pipes", {})
    pipe = pipes.get((N, S, ns, ni))
    if pipe is None:
        pipe = pipes[(N, S, ns, ni)] = _Pipe(ws, N, S, ns, ni)
    peer = exchange is not None and hasattr(exchange, "adam_exchange")
    if exchange is not None and not peer:
        raise RuntimeError("mapping_window_pipelined: only the peer-memory exchange (or none) is pipelined")
    joint = poses7 is not None
    split = os.environ.get("ESLAM_B200_SPLIT_BWD", "1") == "1"
    c2w_flat = c2ws.reshape(b, 16).float().contiguous()
    t_uni = linspace_table(ns, dev)
    main = torch.cuda.current_stream()
    side = pipe.side
    norm = [None]

    s_side, s_imp = side.cuda_stream, pipe.imp.cuda_stream  # raw handles: see _lib.on_stream

    def prep():  # draws + ray selection + depth-guided samples (+ the exchange of the loss normalisers)
        with torch.cuda.stream(side):  # torch's generator kernels: the one place torch's current stream matters
            torch.randint(n_crop, (N,), out=pipe.idx)
            if sc.perturb:
                pipe.ubuf.uniform_()
        _sample(ws, store, sc, pipe.idx, b, pix_per_image, c2w_flat, poses7, 1, gt_depths, gt_colors,
                pipe.u if sc.perturb else None, 0)
        if peer:
            norm[0] = exchange.reduce_counters(ws.counters)

    grad, gq = store.ensure_grad(), store.ensure_q_grad()
    side.wait_stream(main)
    with on_stream(s_side):
        prep()
        pipe.ev_prep.record(side)
    for it in range(iters):
        step = it + 1
        q = store.ensure_q()
        main.wait_event(pipe.ev_prep)

        def importance():
            call("eslam_importance_samples", store.ref(), ptr(store.arena), ptr(q), C.byref(rc), ptr(ws.rays_o),
                 ptr(ws.rays_d), ptr(ws.dl_list), ptr(ws.counters), N, ptr(pipe.u_c), ptr(pipe.u_f), ptr(t_uni), ptr(ws.z),
                 stream())

        def backward(part):
            call("eslam_loss_backward_q_part", store.ref(), ptr(store.arena), ptr(q), ptr(gq), C.byref(cam), C.byref(rc),
                 ptr(ws.rays_o), ptr(ws.rays_d), ptr(ws.z), ptr(ws.gt_depth), ptr(ws.gt_color), ptr(ws.src),
                 ptr(pipe.idx), pix_per_image, ptr(ws.dl_list), ptr(ws.counters),
                 ptr(norm[0]) if norm[0] is not None else None, N, ptr(grad), ptr(ws.pose_grad) if joint else None, part,
                 stream())

        if split:
            # Only the depth-less rays' samples depend on the current parameters (Renderer.py:108-134).  The tiles
            # without such a ray (part 1) are launched behind the importance kernel with programmatic stream
            # serialization and start while it runs; the few tiles that need its samples (part 2) follow it on another
            # stream and fill the last, partial wave of part 1.
            importance()
            pipe.ev_q.record(main)
            pipe.imp.wait_event(pipe.ev_q)
            with on_stream(s_imp):
                backward(2)
                pipe.ev_imp.record(pipe.imp)
            backward(1)
            main.wait_event(pipe.ev_imp)
        else:
            importance()
            backward(0)
        pipe.ev_bwd.record(main)
        side.wait_event(pipe.ev_bwd)
        with on_stream(s_side):
            if joint:
                pg = ws.pose_grad
                if peer:
                    pg, _ = exchange.reduce_small(ws.pose_grad, b, None)
                call("eslam_pose_adam_step", ptr(poses7), ptr(pg), ptr(ws.pose_m), ptr(ws.pose_v), b, 1, lr_cam, lr_cam,
                     step, 0.9, 0.999, 1e-8, ptr(ws.grad7), 1, stream())
            if step < iters:
                prep()
                pipe.ev_prep.record(side)
            else:
                pipe.ev_side.record(side)
        if peer:
            exchange.adam_exchange(step, lr_dec, lr_planes, lr_cplanes, None, 0, None)
        else:
            store.adam_step_q(step, lr_dec, lr_planes, lr_cplanes)
    main.wait_event(pipe.ev_side)
