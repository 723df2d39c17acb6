"""The map as the kernels see it: ONE fp32 arena per process holding the 12 feature planes in
channels-last [H][W][32] order followed by the packed decoders, plus same-layout gradient and Adam
moment arenas.  The reference keeps the planes as 12 separate [1,32,H,W] tensors owned by
`ESLAM` (src/ESLAM.py:175-218) and the decoders as an nn.Module (src/networks/decoders.py:39-62);
`FieldStore.pull()` / `push()` convert between the two.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import DEC_BETA, DEC_FLOATS, FieldDesc, call, ptr, stream

# arena plane order: sdf coarse xy,xz,yz | sdf fine xy,xz,yz | rgb coarse ... | rgb fine ...
# all_planes order (Tracker.py:164): (planes_xy, planes_xz, planes_yz, c_planes_xy, c_planes_xz, c_planes_yz),
# each a list [coarse, fine]
def arena_slot(group: int, scale: int) -> int:
    """group = index into all_planes (0..5), scale 0/1 -> arena plane index."""
    fld, pair = divmod(group, 3)
    return fld * 6 + scale * 3 + pair


# (state_dict key, offset, n) in the packed decoder block (include/eslam_b200.h)
DEC_LAYOUT = (
    ("linears.0.weight", 0, 1024), ("linears.0.bias", 1024, 16), ("linears.1.weight", 1040, 256),
    ("linears.1.bias", 1296, 16), ("output_linear.weight", 1312, 16), ("output_linear.bias", 1328, 1),
    ("c_linears.0.weight", 1332, 1024), ("c_linears.0.bias", 2356, 16), ("c_linears.1.weight", 2372, 256),
    ("c_linears.1.bias", 2628, 16), ("c_output_linear.weight", 2644, 48), ("c_output_linear.bias", 2692, 3),
    ("beta", DEC_BETA, 1),
)


def flatten_planes(all_planes) -> List[torch.Tensor]:
    """The 12 plane tensors of an `all_planes` tuple in arena order."""
    out: List[Optional[torch.Tensor]] = [None] * 12
    for g, lst in enumerate(all_planes):
        if len(lst) != 2:
            raise RuntimeError("myslam_b200 supports exactly two plane scales (coarse, fine) per group")
        for s, p in enumerate(lst):
            out[arena_slot(g, s)] = p
    return out  # type: ignore[return-value]


class Signature:
    """Identity + version of the tensors a store mirrors.  Equal only if every tensor is the SAME live object
    (weak reference, so a new tensor that happens to reuse the address does not match), at the same address and
    autograd version."""

    def __init__(self, tensors):
        self.items = []
        for t in tensors:
            if torch.is_tensor(t):
                self.items.append((weakref.ref(t), t.data_ptr(), t._version))
            else:
                self.items.append((None, 0, float(t)))

    def __eq__(self, other):
        if not isinstance(other, Signature) or len(self.items) != len(other.items):
            return False
        for (ra, pa, va), (rb, pb, vb) in zip(self.items, other.items):
            if pa != pb or va != vb:
                return False
            if (ra is None) != (rb is None):
                return False
            if ra is not None and (ra() is None or ra() is not rb()):
                return False
        return True

    def __ne__(self, other):
        return not self.__eq__(other)


class FieldStore:
    """Arena + descriptor for one map.  Shapes are fixed at construction."""

    def __init__(self, plane_shapes: Sequence[Tuple[int, int]], bound, device):
        if len(plane_shapes) != 12:
            raise RuntimeError("need 12 plane shapes in arena order")
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise RuntimeError("FieldStore needs a CUDA device; myslam_b200 has no CPU path")
        self.shapes = [(int(h), int(w)) for h, w in plane_shapes]
        self.desc = FieldDesc()
        off = 0
        self.plane_off: List[int] = []
        for i, (h, w) in enumerate(self.shapes):
            self.desc.plane[i].offset = off
            self.desc.plane[i].H = h
            self.desc.plane[i].W = w
            self.plane_off.append(off)
            off += h * w * 32
        self.n_sdf_end = self.plane_off[6]
        self.n_planes_end = off
        self.dec_off = off
        off += DEC_FLOATS
        self.n_floats = off
        assert self.n_floats % 4 == 0
        self.desc.dec_offset = self.dec_off
        self.desc.n_floats = self.n_floats
        b = torch.as_tensor(bound, dtype=torch.float32).cpu()
        for a in range(3):
            self.desc.bound[a][0] = float(b[a, 0])
            self.desc.bound[a][1] = float(b[a, 1])
        self.bound = b
        self.arena = torch.zeros(self.n_floats, dtype=torch.float32, device=self.device)
        self.grad: Optional[torch.Tensor] = None
        self.exp_avg: Optional[torch.Tensor] = None
        self.exp_avg_sq: Optional[torch.Tensor] = None
        self.touched: Optional[torch.Tensor] = None  # one flag per 128 parameters: any non-zero gradient since reset
        self._sig = None  # (data_ptr, version) of what was last pulled
        # generation of the parameter arena: bumped by everything that writes it (imports, optimiser steps, exchanges).
        # The Q images and autograd contexts remember the generation they were made from.
        self.gen = 0
        # Q images: the first decoder layer applied to the planes, 16 channels per texel, plane i
        # at half the float offset of plane i in the parameter arena.  The render kernels of both loops read these.
        self.q_arena: Optional[torch.Tensor] = None
        self.q_gen = -1
        self.gq_arena: Optional[torch.Tensor] = None   # gradient images of the mapping backward (layout of q_arena)
        self.touched_q: Optional[torch.Tensor] = None  # exact-skip flags of eslam_q_adam_planes (one per texel)

    # ------------------------------------------------------------------ construction helpers
    @classmethod
    def from_planes(cls, all_planes, bound, device=None) -> "FieldStore":
        flat = flatten_planes(all_planes)
        for p in flat:
            if p.dim() != 4 or p.shape[0] != 1 or p.shape[1] != 32:
                raise RuntimeError(f"planes must be [1,32,H,W] (model.c_dim=32), got {tuple(p.shape)}")
        dev = device if device is not None else flat[0].device
        return cls([(p.shape[2], p.shape[3]) for p in flat], bound, dev)

    def matches(self, all_planes) -> bool:
        flat = flatten_planes(all_planes)
        return all((p.shape[2], p.shape[3]) == s for p, s in zip(flat, self.shapes))

    def ref(self):
        return C.byref(self.desc)

    @property
    def dec(self) -> torch.Tensor:
        return self.arena[self.dec_off:self.dec_off + DEC_FLOATS]

    def plane_view(self, i: int, which: Optional[torch.Tensor] = None) -> torch.Tensor:
        """[H,W,32] view of plane i in `which` arena (default: parameters)."""
        a = self.arena if which is None else which
        h, w = self.shapes[i]
        return a[self.plane_off[i]:self.plane_off[i] + h * w * 32].view(h, w, 32)

    # ------------------------------------------------------------------ reference layout -> arena
    def pull_planes(self, all_planes) -> None:
        for i, p in enumerate(flatten_planes(all_planes)):
            _lib.require_cuda(p, "plane")
            src = p.detach()
            if src.dtype != torch.float32 or not src.is_contiguous():
                src = src.float().contiguous()
            call("eslam_plane_import", ptr(src), ptr(self.arena), C.byref(self.desc.plane[i]), stream())
        self.gen += 1

    def pull_decoders(self, state: Dict[str, torch.Tensor], beta) -> None:
        dec = self.dec
        for key, off, n in DEC_LAYOUT:
            if key == "beta":
                val = beta if torch.is_tensor(beta) else torch.tensor([float(beta)])
                dec[off:off + 1].copy_(val.detach().reshape(1).to(self.device, torch.float32), non_blocking=True)
            else:
                dec[off:off + n].copy_(state[key].detach().reshape(-1), non_blocking=True)
        self.gen += 1

    def signature(self, all_planes, dec_tensors) -> "Signature":
        return Signature(flatten_planes(all_planes) + list(dec_tensors))

    # ------------------------------------------------------------------ arena -> reference layout
    def push_planes(self, all_planes, which: Optional[torch.Tensor] = None) -> None:
        a = self.arena if which is None else which
        for i, p in enumerate(flatten_planes(all_planes)):
            dst = p.detach()
            if not dst.is_contiguous() or dst.dtype != torch.float32:
                raise RuntimeError("planes must be contiguous float32 to be updated in place")
            call("eslam_plane_export", ptr(a), ptr(dst), C.byref(self.desc.plane[i]), stream())

    def export_plane(self, i: int, which: torch.Tensor) -> torch.Tensor:
        """NCHW copy of plane i of arena `which` (used for autograd gradients)."""
        h, w = self.shapes[i]
        out = torch.empty(1, 32, h, w, dtype=torch.float32, device=self.device)
        call("eslam_plane_export", ptr(which), ptr(out), C.byref(self.desc.plane[i]), stream())
        return out

    def push_decoders(self, module) -> None:
        dec = self.dec
        with torch.no_grad():
            sd = dict(module.named_parameters())
            for key, off, n in DEC_LAYOUT:
                if key in sd:
                    sd[key].copy_(dec[off:off + n].view_as(sd[key]))

    def dec_grad_dict(self, which: torch.Tensor) -> Dict[str, torch.Tensor]:
        g = which[self.dec_off:self.dec_off + DEC_FLOATS]
        return {key: g[off:off + n] for key, off, n in DEC_LAYOUT}

    def parameter_grads(self) -> torch.Tensor:
        """Parameter-form gradient arena (layout of `arena`) from what eslam_loss_backward_q left: the gradient images
        (d loss / d plane = GQ . W1_slice, dW1 = sum GQ (x) plane) and the decoder block of `grad`.  For callers that
        want to LOOK at gradients (tests, autograd bridges); the optimiser consumes the images directly."""
        out = torch.zeros_like(self.arena)
        out[self.dec_off:] = self.grad[self.dec_off:]
        for i in range(12):
            h, w = self.shapes[i]
            fld, sc = i // 6, (i % 6) // 3
            w1 = self.dec_off + (1332 if fld else 0)
            W = self.arena[w1:w1 + 1024].view(16, 64)[:, sc * 32:(sc + 1) * 32]
            G = self.gq_arena[self.plane_off[i] // 2: self.plane_off[i] // 2 + h * w * 16].view(h * w, 16)
            out[self.plane_off[i]: self.plane_off[i] + h * w * 32] = (G @ W).reshape(-1)
            out[w1:w1 + 1024].view(16, 64)[:, sc * 32:(sc + 1) * 32] += G.t() @ self.plane_view(i).reshape(h * w, 32)
        return out

    # ------------------------------------------------------------------ kernels' view
    def bind(self, force: bool = False) -> None:
        """Make this map's decoders the ones the forward-only kernels read (constant memory, process-global per
        device).  Skipped when this store at this generation is what the device's constant bank already holds."""
        key = (id(self), self.gen)
        if not force and _lib.BOUND_DECODERS.get(self.device.index) == key:
            return
        call("eslam_bind_decoders", ptr(self.dec), stream())
        _lib.BOUND_DECODERS[self.device.index] = key

    def ensure_q(self) -> torch.Tensor:
        """The Q images of the CURRENT parameters (rebuilt when the arena's generation moved on)."""
        if self.q_arena is None:
            self.q_arena = torch.zeros(self.n_planes_end // 2, dtype=torch.float32, device=self.device)
        if self.q_gen != self.gen:
            call("eslam_q_build", self.ref(), ptr(self.arena), ptr(self.q_arena), stream())
            self.q_gen = self.gen
        return self.q_arena

    def ensure_grad(self) -> torch.Tensor:
        if self.grad is None:
            self.grad = torch.zeros_like(self.arena)
        return self.grad

    def ensure_q_grad(self) -> torch.Tensor:
        if self.gq_arena is None:
            self.gq_arena = torch.zeros(self.n_planes_end // 2, dtype=torch.float32, device=self.device)
            n = _lib.load().eslam_q_touched_bytes(self.ref())
            self.touched_q = torch.zeros(n, dtype=torch.uint8, device=self.device)
        return self.gq_arena

    def reset_adam(self) -> None:
        """Fresh optimiser state, as Mapper.optimize_mapping builds a new Adam per call (Mapper.py:291-299)."""
        if self.exp_avg is None:
            self.exp_avg = torch.zeros_like(self.arena)
            self.exp_avg_sq = torch.zeros_like(self.arena)
            self.touched = torch.zeros((self.n_floats + 127) // 128, dtype=torch.uint8, device=self.device)
        else:
            self.exp_avg.zero_()
            self.exp_avg_sq.zero_()
            self.touched.zero_()
        self.ensure_grad().zero_()
        self.ensure_q_grad().zero_()
        self.touched_q.zero_()

    def adam_step_q(self, step: int, lr_dec: float, lr_planes: float, lr_cplanes: float, betas=(0.9, 0.999),
                    eps=1e-8) -> None:
        """One torch.optim.Adam step (Mapper.py:348-350) from what eslam_loss_backward_q left: the planes from the
        gradient images (eslam_q_adam_planes, which also completes dW1 in the gradient arena's decoder block), then
        the decoders; zeroes the gradient images and the decoder gradients."""
        if self.exp_avg is None or self.gq_arena is None:
            raise RuntimeError("FieldStore.adam_step_q before reset_adam(): there is no optimiser state")
        call("eslam_q_adam_planes", self.ref(), ptr(self.arena), ptr(self.gq_arena), ptr(self.exp_avg),
             ptr(self.exp_avg_sq), ptr(self.grad), ptr(self.touched_q), lr_planes, lr_cplanes, step, betas[0], betas[1],
             eps, stream())
        self.adam_step_decoders(step, lr_dec, betas, eps)

    def adam_step_decoders(self, step: int, lr_dec: float, betas=(0.9, 0.999), eps=1e-8) -> None:
        seg_end = (C.c_int64 * 1)(DEC_FLOATS)
        seg_lr = (C.c_double * 1)(lr_dec)
        o = self.dec_off
        call("eslam_adam_step", ptr(self.arena[o:]), ptr(self.grad[o:]), ptr(self.exp_avg[o:]), ptr(self.exp_avg_sq[o:]),
             DEC_FLOATS, seg_end, seg_lr, 1, step, betas[0], betas[1], eps, stream())
        self.gen += 1

    def adam_step(self, step: int, lr_dec: float, lr_planes: float, lr_cplanes: float, betas=(0.9, 0.999),
                  eps=1e-8) -> None:
        """One torch.optim.Adam step over planes (two lr groups) + decoders from the PARAMETER-form gradient arena
        (what eslam_render_backward / eslam_loss_backward leave); zeroes the gradient arena."""
        seg_end = (C.c_int64 * 3)(self.n_sdf_end, self.n_planes_end, self.n_floats)
        seg_lr = (C.c_double * 3)(lr_planes, lr_cplanes, lr_dec)
        if self.exp_avg is None:
            raise RuntimeError("FieldStore.adam_step before reset_adam(): there is no optimiser state")
        call("eslam_adam_step_sparse", ptr(self.arena), ptr(self.grad), ptr(self.exp_avg), ptr(self.exp_avg_sq),
             self.n_floats, seg_end, seg_lr, 3, step, betas[0], betas[1], eps, ptr(self.touched), stream())
        self.gen += 1
