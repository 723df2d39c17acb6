"""Drop-in for `src/utils/Renderer.py` (reference lines 26-204): same constructor and method
signatures; sampling, decoding and compositing run in the sm_100a kernels and
`render_batch_ray` stays autograd-connected to rays, planes and decoder parameters.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import torch

from . import _lib
from ._lib import COUNTER_WORDS, MAX_COMPACT_BLOCKS, N_COUNTERS, RenderCfg, call, ptr, stream
from .decoders import decoder_tensors, split_arena_grads, synced_store
from .field import flatten_planes


class TorchDraws:
    """Random draws taken from torch's CUDA generator with the reference's shapes and order
    (common.py:108, Renderer.py:59, common.py:59), so a seeded run consumes the same stream."""

    def __init__(self, device):
        self.device = device

    def randint(self, high, n):
        return torch.randint(high, (n,), device=self.device)

    def rand(self, rows, cols):
        return torch.rand(rows, cols, device=self.device)

    def rand_many(self, shapes):
        """Several uniform blocks from ONE generator call (fewer launches); only used where the reference's draw
        shapes are not reproduced anyway (strict_rng off)."""
        sizes = [r * c for r, c in shapes]
        flat = torch.rand(sum(sizes), device=self.device)
        out, o = [], 0
        for (r, c), n in zip(shapes, sizes):
            out.append(flat[o:o + n].view(r, c))
            o += n
        return out


class ReplayDraws:
    """Replays recorded draws (tests / parity runs); tensors are moved to the device."""

    def __init__(self, recorded, device):
        self.recorded = [t.to(device) for t in recorded]
        self.pos = 0

    def _next(self):
        t = self.recorded[self.pos]
        self.pos += 1
        return t

    def randint(self, high, n):
        t = self._next()
        assert t.dtype == torch.int64 and t.numel() == n, "replayed randint has the wrong shape"
        return t.contiguous()

    def rand(self, rows, cols):
        t = self._next()
        assert t.shape == (rows, cols), f"replayed rand {tuple(t.shape)} != {(rows, cols)}"
        return t.float().contiguous()


def make_cfg(n_stratified, n_importance, truncation, w=(0, 0, 0, 0, 0)) -> RenderCfg:
    c = RenderCfg()
    c.n_stratified, c.n_importance = int(n_stratified), int(n_importance)
    c.truncation = float(truncation)
    c.w_fs, c.w_center, c.w_tail, c.w_depth, c.w_color = (float(x) for x in w)
    return c


_LINSPACE = {}


def linspace_table(n, device):
    """torch.linspace(0,1,n) on the device, as Renderer.py:85-86 builds it every call."""
    key = (n, str(device))
    t = _LINSPACE.get(key)
    if t is None:
        t = torch.linspace(0.0, 1.0, steps=n, device=device)
        _LINSPACE[key] = t
    return t


class _RenderFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, rays_d, rays_o, z, store, dts, *leaves):
        R, S = z.shape
        dev = z.device
        depth = torch.empty(R, dtype=torch.float32, device=dev)
        rgb = torch.empty(R, 3, dtype=torch.float32, device=dev)
        sdf = torch.empty(R, S, dtype=torch.float32, device=dev)
        call("eslam_render_forward", store.ref(), ptr(store.arena), ptr(rays_o), ptr(rays_d), ptr(z), R, S, None,
             ptr(depth), ptr(rgb), ptr(sdf), stream())
        ctx.store, ctx.dts, ctx.gen = store, dts, store.gen
        ctx.save_for_backward(rays_d, rays_o, z)
        return depth, rgb, sdf

    @staticmethod
    def backward(ctx, g_depth, g_rgb, g_sdf):
        rays_d, rays_o, z = ctx.saved_tensors
        store = ctx.store
        if store.gen != ctx.gen:
            raise RuntimeError("render_batch_ray: the map's parameters changed between forward and backward (another "
                               "model or an optimiser step re-used this device's FieldStore); run backward first")
        R, S = z.shape
        needs = ctx.needs_input_grad
        need_rays = needs[0] or needs[1]
        need_leaves = needs[5:]
        want_field = any(need_leaves)
        dev = z.device
        garena = torch.zeros_like(store.arena) if want_field else None
        g_o = torch.empty(R, 3, dtype=torch.float32, device=dev) if need_rays else None
        g_d = torch.empty(R, 3, dtype=torch.float32, device=dev) if need_rays else None
        zero = lambda g, shape: (torch.zeros(shape, dtype=torch.float32, device=dev) if g is None
                                 else g.contiguous().float())
        gd_, gc_, gs_ = zero(g_depth, (R,)), zero(g_rgb, (R, 3)), zero(g_sdf, (R, S))
        store.bind()
        call("eslam_render_backward", store.ref(), ptr(store.arena), ptr(rays_o), ptr(rays_d), ptr(z), R, S, ptr(gd_),
             ptr(gc_), ptr(gs_), ptr(garena), ptr(g_o), ptr(g_d), stream())
        gp, gdec = [None] * 12, [None] * 13
        if want_field:
            gp, gdec = split_arena_grads(store, garena, need_leaves[:12], need_leaves[12:], ctx.dts)
        return (g_d if needs[0] else None, g_o if needs[1] else None, None, None, None, *gp, *gdec)


class Renderer(object):
    """Renderer class for rendering depth and color (reference Renderer.py:26-44)."""

    def __init__(self, cfg, eslam, ray_batch_size=10000):
        self.ray_batch_size = ray_batch_size
        self.perturb = cfg['rendering']['perturb']
        self.n_stratified = cfg['rendering']['n_stratified']
        self.n_importance = cfg['rendering']['n_importance']
        self.scale = cfg['scale']
        self.bound = eslam.bound.to(eslam.device, non_blocking=True)
        self.H, self.W, self.fx, self.fy, self.cx, self.cy = eslam.H, eslam.W, eslam.fx, eslam.fy, eslam.cx, eslam.cy
        self.draws = None  # optional injected draw source (tests); default: torch's generator

    # ---- reference helpers kept for API compatibility -------------------------------------------------
    def perturbation(self, z_vals):
        """Stratified jitter (Renderer.py:46-61); torch ops, used only by external callers."""
        mids = 0.5 * (z_vals[..., 1:] + z_vals[..., :-1])
        upper = torch.cat([mids, z_vals[..., -1:]], -1)
        lower = torch.cat([z_vals[..., :1], mids], -1)
        t_rand = torch.rand(z_vals.shape, device=z_vals.device)
        return lower + (upper - lower) * t_rand

    def sdf2alpha(self, sdf, beta=10):
        """1 - exp(-beta*sigmoid(-sdf*beta)) (Renderer.py:149-153)."""
        return 1. - torch.exp(-beta * torch.sigmoid(-sdf * beta))

    # ---- sampling (no grad) ------------------------------------------------------------------------------
    def sample_z(self, store, rays_o, rays_d, gt_depth, truncation, draws=None):
        """z_vals [R,S] exactly as Renderer.py:81-134 builds them; the uniforms are drawn with the
        reference's shapes ([R1,S], then [R0,n_strat] and [R0,n_imp] if depth-less rays exist)."""
        dev = rays_o.device
        R = rays_o.shape[0]
        ns, ni = self.n_stratified, self.n_importance
        S = ns + ni
        draws = draws or self.draws or TorchDraws(dev)
        cfg = make_cfg(ns, ni, truncation)
        d = gt_depth.reshape(-1).float().contiguous()
        r1 = int((d > 0).sum().item())  # the reference syncs here too (boolean indexing, Renderer.py:93)
        r0 = R - r1
        z = torch.empty(R, S, dtype=torch.float32, device=dev)
        dl = torch.empty(max(R, 1), dtype=torch.int32, device=dev)
        zord = torch.empty(max(R, 1), dtype=torch.int32, device=dev)
        cnt = torch.empty(COUNTER_WORDS, dtype=torch.int32, device=dev)
        u = draws.rand(r1, S) if self.perturb else None
        t_uni, t_surf = linspace_table(ns, dev), linspace_table(ni, dev)
        call("eslam_depth_samples", C.byref(cfg), ptr(d), R, ptr(u), ptr(t_uni), ptr(t_surf), ptr(z), ptr(dl),
             ptr(zord), ptr(cnt), stream())
        if r0 > 0:
            if not self.perturb:
                raise RuntimeError("rendering.perturb=False with depth-less rays is not supported by the kernels")
            u_c = draws.rand(r0, ns)
            u_f = draws.rand(r0, ni)
            call("eslam_importance_samples", store.ref(), ptr(store.arena), ptr(store.ensure_q()), C.byref(cfg),
                 ptr(rays_o), ptr(rays_d),
                 ptr(dl), ptr(cnt), r0, ptr(u_c), ptr(u_f), ptr(t_uni), ptr(z), stream())
        return z

    # ---- reference surface ----------------------------------------------------------------------------------
    def render_batch_ray(self, all_planes, decoders, rays_d, rays_o, device, truncation, gt_depth=None):
        """Render depth and colour for a batch of rays (Renderer.py:63-147).
        Returns (depth[R], rgb[R,3], sdf[R,S], z_vals[R,S])."""
        if gt_depth is None:
            raise RuntimeError("render_batch_ray needs gt_depth (the reference dereferences it unconditionally, "
                               "Renderer.py:91)")
        store = synced_store(all_planes, decoders, self.bound)
        ro = rays_o.reshape(-1, 3).float().contiguous()
        rd = rays_d.reshape(-1, 3).float().contiguous()
        if ro.shape[0] == 0:
            S = self.n_stratified + self.n_importance
            z0 = lambda *shape: torch.zeros(*shape, dtype=torch.float32, device=ro.device)
            return z0(0), z0(0, 3), z0(0, S), z0(0, S)
        with torch.no_grad():
            z = self.sample_z(store, ro.detach(), rd.detach(), gt_depth, truncation)
        dts = decoder_tensors(decoders)
        leaves = flatten_planes(all_planes) + [t if torch.is_tensor(t) else None for t in dts]
        depth, rgb, sdf = _RenderFn.apply(rd, ro, z, store, dts, *leaves)
        return depth, rgb, sdf, z

    def render_img(self, all_planes, decoders, c2w, truncation, device, gt_depth=None):
        """Full-image inference (Renderer.py:155-204): depth[H,W] float64, colour[H,W,3]; perturbation on.
        All H*W rays go through the kernels in ONE pass (one sampling launch, one render launch: SURVEY.md 8f-3);
        with `strict_rng` (ESLAM_B200_STRICT_RNG=1) the reference's chunks of ray_batch_size rays are kept so the
        random stream is consumed in the reference's shapes."""
        from .common import get_rays
        with torch.no_grad():
            H, W = self.H, self.W
            strict = getattr(self, "strict_rng", os.environ.get("ESLAM_B200_STRICT_RNG", "0") == "1")
            chunk = self.ray_batch_size if strict else min(H * W, MAX_COMPACT_BLOCKS * 256)
            rays_o, rays_d = get_rays(H, W, self.fx, self.fy, self.cx, self.cy, c2w, device)
            rays_o = rays_o.reshape(-1, 3).contiguous()
            rays_d = rays_d.reshape(-1, 3).contiguous()
            gt_depth = gt_depth.reshape(-1)
            depth_list, color_list = [], []
            for i in range(0, rays_d.shape[0], chunk):
                sl = slice(i, i + chunk)
                depth, color, _, _ = self.render_batch_ray(all_planes, decoders, rays_d[sl], rays_o[sl], device,
                                                           truncation, gt_depth=gt_depth[sl])
                depth_list.append(depth.double())
                color_list.append(color)
            depth = torch.cat(depth_list, dim=0).reshape(H, W)
            color = torch.cat(color_list, dim=0).reshape(H, W, 3)
            return depth, color
