"""Drop-in for `src/networks/decoders.py` (reference lines 28-146): same constructor, attributes,
`state_dict` keys and method signatures; the arithmetic runs in the sm_100a kernels.

`forward` is differentiable with respect to the points, the 12 planes and the decoder
parameters (autograd.Function over eslam_decode_points / eslam_decode_backward).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Tuple

import torch
import torch.nn as nn

from . import _lib
from ._lib import call, ptr, stream
from .field import DEC_LAYOUT, FieldStore, flatten_planes

_STORES: Dict[tuple, FieldStore] = {}


def decoder_tensors(decoders) -> list:
    """The 13 decoder leaves in DEC_LAYOUT order (beta may be a python number)."""
    named = dict(decoders.named_parameters())
    out = []
    for key, _, _ in DEC_LAYOUT:
        out.append(named[key] if key in named else getattr(decoders, "beta"))
    return out


def synced_store(all_planes, decoders, bound=None) -> FieldStore:
    """A FieldStore holding exactly the values of (all_planes, decoders), re-imported only when a
    tensor changed (data_ptr/_version signature); its decoders are bound for the kernels."""
    flat = flatten_planes(all_planes)
    dev = flat[0].device
    if dev.type != "cuda":
        raise RuntimeError("myslam_b200 needs CUDA tensors; there is no CPU fallback")
    b = decoders.bound if bound is None else bound
    key = (dev.index, tuple((p.shape[2], p.shape[3]) for p in flat), tuple(torch.as_tensor(b).flatten().tolist()))
    store = _STORES.get(key)
    if store is None:
        store = FieldStore.from_planes(all_planes, b, dev)
        _STORES[key] = store
    dts = decoder_tensors(decoders)
    sig = store.signature(all_planes, dts)
    if sig != store._sig:
        store.pull_planes(all_planes)
        sd = {k: t for (k, _, _), t in zip(DEC_LAYOUT, dts) if k != "beta"}
        store.pull_decoders(sd, dts[-1])
        store._sig = sig
    store.bind()
    return store


def split_arena_grads(store: FieldStore, garena: torch.Tensor, needs_planes, needs_dec, dts):
    """Gradient arena -> 12 NCHW plane grads + 13 decoder grads (None where not needed)."""
    gp = [store.export_plane(i, garena) if needs_planes[i] else None for i in range(12)]
    gd_all = store.dec_grad_dict(garena)
    gd = []
    for (key, _, _), t, need in zip(DEC_LAYOUT, dts, needs_dec):
        gd.append(gd_all[key].reshape(t.shape).clone() if (need and torch.is_tensor(t)) else None)
    return gp, gd


class _DecodeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pts, store, dts, *leaves):
        n = pts.shape[0]
        raw = torch.empty(n, 4, dtype=torch.float32, device=pts.device)
        call("eslam_decode_points", store.ref(), ptr(store.arena), ptr(pts), n, ptr(raw), 0, stream())
        ctx.store, ctx.dts = store, dts
        ctx.save_for_backward(pts)
        ctx.gen = store.gen
        return raw

    @staticmethod
    def backward(ctx, g_raw):
        (pts,) = ctx.saved_tensors
        store = ctx.store
        if store.gen != ctx.gen:
            raise RuntimeError("Decoders.forward: the map's parameters changed between forward and backward (another "
                               "model or an optimiser step re-used this device's FieldStore); run backward first")
        needs = ctx.needs_input_grad
        need_pts, need_leaves = needs[0], needs[3:]
        want_field = any(need_leaves)
        garena = torch.zeros_like(store.arena) if want_field else None
        g_pts = torch.empty_like(pts) if need_pts else None
        store.bind()
        call("eslam_decode_backward", store.ref(), ptr(store.arena), ptr(pts), pts.shape[0],
             ptr(g_raw.contiguous().float()), ptr(garena), ptr(g_pts), stream())
        gp, gd = ([None] * 12, [None] * 13)
        if want_field:
            gp, gd = split_arena_grads(store, garena, need_leaves[:12], need_leaves[12:], ctx.dts)
        return (g_pts, None, None, *gp, *gd)


class Decoders(nn.Module):
    """Decoders for SDF and RGB (reference decoders.py:28-62): two 64->16->16->{1,3} MLPs over summed
    tri-plane features, learnable `beta`.  `bound` is assigned externally (ESLAM.py:173)."""

    def __init__(self, c_dim=32, hidden_size=16, truncation=0.08, n_blocks=2, learnable_beta=True):
        super().__init__()
        if c_dim != 32 or hidden_size != 16 or n_blocks != 2:
            raise RuntimeError("the sm_100a kernels are specialised for c_dim=32, hidden_size=16, n_blocks=2 "
                               "(every config the reference ships); no generic fallback exists")
        self.c_dim = c_dim
        self.truncation = truncation
        self.n_blocks = n_blocks
        self.linears = nn.ModuleList([nn.Linear(2 * c_dim, hidden_size)] +
                                     [nn.Linear(hidden_size, hidden_size) for _ in range(n_blocks - 1)])
        self.c_linears = nn.ModuleList([nn.Linear(2 * c_dim, hidden_size)] +
                                       [nn.Linear(hidden_size, hidden_size) for _ in range(n_blocks - 1)])
        self.output_linear = nn.Linear(hidden_size, 1)
        self.c_output_linear = nn.Linear(hidden_size, 3)
        if learnable_beta:
            self.beta = nn.Parameter(10 * torch.ones(1))
        else:
            self.beta = 10

    # ---- reference surface -----------------------------------------------------------------------
    def sample_plane_feature(self, p_nor, planes_xy, planes_xz, planes_yz):
        """feat[N,64] = cat_s((xy_s+xz_s)+yz_s) (decoders.py:64-85).  Forward only."""
        if torch.is_grad_enabled() and (p_nor.requires_grad or any(p.requires_grad for p in planes_xy)):
            raise NotImplementedError("sample_plane_feature is forward-only here; differentiate through forward()")
        store = synced_store((planes_xy, planes_xz, planes_yz, planes_xy, planes_xz, planes_yz), self)
        pn = p_nor.detach().reshape(-1, 3).float().contiguous()
        feat = torch.empty(pn.shape[0], 64, dtype=torch.float32, device=pn.device)
        call("eslam_sample_plane_feature", store.ref(), ptr(store.arena), ptr(pn), pn.shape[0], 0, ptr(feat), stream())
        return feat

    def _decode_nor(self, p_nor, all_planes, flags):
        store = synced_store(all_planes, self)
        pn = p_nor.detach().reshape(-1, 3).float().contiguous()
        raw = torch.empty(pn.shape[0], 4, dtype=torch.float32, device=pn.device)
        call("eslam_decode_points", store.ref(), ptr(store.arena), ptr(pn), pn.shape[0], ptr(raw), flags | 4, stream())
        return raw

    def get_raw_sdf(self, p_nor, all_planes):
        """tanh SDF of normalised points (decoders.py:87-105).  Forward only."""
        return self._decode_nor(p_nor, all_planes, 1)[:, 3]

    def get_raw_rgb(self, p_nor, all_planes):
        """sigmoid RGB of normalised points (decoders.py:107-125).  Forward only."""
        return self._decode_nor(p_nor, all_planes, 0)[:, :3]

    def forward(self, p, all_planes):
        """raw[...,4] = (r,g,b,sdf) of world points p (decoders.py:127-146)."""
        p_shape = p.shape
        store = synced_store(all_planes, self)
        pts = p.reshape(-1, 3).float().contiguous()
        dts = decoder_tensors(self)
        leaves = flatten_planes(all_planes) + [t if torch.is_tensor(t) else None for t in dts]
        raw = _DecodeFn.apply(pts, store, dts, *leaves)
        return raw.reshape(*p_shape[:-1], -1)
