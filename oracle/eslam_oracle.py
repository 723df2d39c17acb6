"""CPU oracle for the ESLAM render-and-optimise hot path.

THIS FILE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s CPU-baseline / `--impl reference`
legs may import it.  Nothing under `myslam_b200/` imports it, and the product
path has no CPU fallback: it raises when the CUDA library is missing.

What it is: a plain PyTorch (fp32, autograd, any device but meant for CPU)
restatement of the algorithm the reference implements in
`src/common.py`, `src/utils/Renderer.py`, `src/networks/decoders.py`,
`src/Tracker.py:114-210`, `src/Mapper.py:110-144,211-364` and
`src/utils/Mesher.py:130-186`.  Each function cites the reference lines it
follows.  Two deliberate differences from the reference's *shape* (not its
arithmetic):

* every random draw is INJECTED through a `Draws` object instead of being taken
  from the global generator, so the CUDA path and this oracle can consume the
  same numbers (`LiveDraws` reproduces the reference's call order and shapes on
  a torch generator; `ReplayDraws` replays a recorded list);
* state lives in a small `Field` record instead of in `ESLAM`/`Tracker`/`Mapper`
  attributes.

Parity pin: the reference ships no tests or golden
vectors, so this oracle is pinned against the reference ITSELF, imported
unmodified from /root/reference in the build container by
`tests/golden/make_golden.py` (bit-exact for pixel indices, z_vals, rays and
masks; <=1e-6 for rendered values, losses, gradients and post-Adam state).  The
resulting vectors are committed under `tests/golden/` and replayed by the
`-m "not gpu"` tests.  The one boundary the reference itself cannot pin is
`pytorch3d.transforms` (third-party, `pytorch3d==0.7.1`, not vendored, absent
here): `quaternion_to_matrix` / `matrix_to_quaternion` below restate its
published algorithm and are pinned against scipy's Rotation -- for that
boundary alone: PARITY UNPINNED against pytorch3d proper.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field as _dc_field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

# ----------------------------------------------------------------------------
# random draws (injected)
# ----------------------------------------------------------------------------


class LiveDraws:
    """Draw from a torch generator in the reference's order and shapes and keep a log.

    Order per iteration (SURVEY.md 7 'Bit-exact sampling'): randint(n*b) at
    common.py:108, rand[R1,S] at Renderer.py:59 (via :103), and only when
    depth-less rays exist rand[R0,n_strat] (Renderer.py:59 via :120) then
    rand[R0,n_imp] (common.py:59).
    """

    def __init__(self, generator: Optional[torch.Generator] = None, device="cpu"):
        self.gen = generator
        self.device = device
        self.log: List[torch.Tensor] = []

    def randint(self, high: int, n: int) -> torch.Tensor:
        t = torch.randint(high, (n,), generator=self.gen, device=self.device)
        self.log.append(t.clone())
        return t

    def rand(self, *shape: int) -> torch.Tensor:
        t = torch.rand(*shape, generator=self.gen, device=self.device)
        self.log.append(t.clone())
        return t


class ReplayDraws:
    """Replay a recorded list of draws; shapes are checked so a mismatch is loud."""

    def __init__(self, recorded: Sequence[torch.Tensor]):
        self.recorded = list(recorded)
        self.pos = 0

    def _next(self, shape) -> torch.Tensor:
        if self.pos >= len(self.recorded):
            raise RuntimeError("ReplayDraws exhausted")
        t = self.recorded[self.pos]
        self.pos += 1
        if tuple(t.shape) != tuple(shape):
            raise RuntimeError(f"draw {self.pos - 1}: recorded shape {tuple(t.shape)} != requested {tuple(shape)}")
        return t

    def randint(self, high: int, n: int) -> torch.Tensor:
        t = self._next((n,))
        assert int(t.max()) < high
        return t

    def rand(self, *shape: int) -> torch.Tensor:
        return self._next(shape)


# ----------------------------------------------------------------------------
# pose <-> matrix   (pytorch3d 0.7.1 published algorithm; common.py:155-181)
# ----------------------------------------------------------------------------


def quaternion_to_matrix(q: torch.Tensor) -> torch.Tensor:
    """Real-first quaternion [...,4] -> rotation [...,3,3], scaled by 2/|q|^2 so
    un-normalised (Adam-updated) quaternions stay valid.  pytorch3d 0.7.1
    `transforms.rotation_conversions.quaternion_to_matrix`; called at common.py:178."""
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    m = torch.stack(
        (
            1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
            two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
            two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return m.reshape(q.shape[:-1] + (3, 3))


def _sqrt_pos(x: torch.Tensor) -> torch.Tensor:
    out = torch.zeros_like(x)
    pos = x > 0
    out[pos] = torch.sqrt(x[pos])
    return out


def matrix_to_quaternion(M: torch.Tensor) -> torch.Tensor:
    """Rotation [...,3,3] -> real-first quaternion, best-conditioned of the four
    candidate rows (argmax of |q_i|, 0.1 floor), sign NOT standardised.
    pytorch3d 0.7.1 `matrix_to_quaternion`; called at common.py:165,167."""
    batch = M.shape[:-2]
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(M.reshape(batch + (9,)), -1)
    q_abs = _sqrt_pos(
        torch.stack(
            [1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22, 1.0 - m00 - m11 + m22], -1
        )
    )
    cand = torch.stack(
        [
            torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
            torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], -1),
            torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], -1),
            torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], -1),
        ],
        -2,
    )
    floor = torch.tensor(0.1, dtype=q_abs.dtype, device=q_abs.device)
    cand = cand / (2.0 * q_abs[..., None].max(floor))
    pick = F.one_hot(q_abs.argmax(-1), num_classes=4) > 0.5
    return cand[pick, :].reshape(batch + (4,))


def cam_pose_to_matrix(poses: torch.Tensor) -> torch.Tensor:
    """[B,7]=(qw,qx,qy,qz,tx,ty,tz) -> [B,4,4].  common.py:169-181."""
    B = poses.shape[0]
    c2w = torch.eye(4, device=poses.device).unsqueeze(0).repeat(B, 1, 1)
    c2w[:, :3, :3] = quaternion_to_matrix(poses[:, :4])
    c2w[:, :3, 3] = poses[:, 4:]
    return c2w


def matrix_to_cam_pose(mats: torch.Tensor) -> torch.Tensor:
    """[B,4,4] -> [B,7] rotation first.  common.py:155-167 (RT=True branch)."""
    return torch.cat([matrix_to_quaternion(mats[:, :3, :3]), mats[:, :3, 3]], -1)


# ----------------------------------------------------------------------------
# scene record
# ----------------------------------------------------------------------------

DECODER_KEYS = (
    "linears.0.weight", "linears.0.bias", "linears.1.weight", "linears.1.bias",
    "c_linears.0.weight", "c_linears.0.bias", "c_linears.1.weight", "c_linears.1.bias",
    "output_linear.weight", "output_linear.bias", "c_output_linear.weight", "c_output_linear.bias",
)


@dataclass
class Field:
    """The map: 6 plane lists [coarse, fine] of [1,C,H,W] (ESLAM.py:175-218 layout:
    xy=[ny,nx], xz=[nz,nx], yz=[nz,ny]), decoder weights keyed like
    `Decoders.state_dict()` (decoders.py:39-62), beta, bound[3,2]."""

    planes: Tuple[List[torch.Tensor], ...]  # (xy, xz, yz, c_xy, c_xz, c_yz)
    dec: Dict[str, torch.Tensor]
    beta: torch.Tensor  # shape [1]
    bound: torch.Tensor  # [3,2] fp32

    def leaves(self) -> List[torch.Tensor]:
        out = [p for group in self.planes for p in group]
        out += [self.dec[k] for k in DECODER_KEYS]
        out.append(self.beta)
        return out

    def clone(self, requires_grad: bool = False) -> "Field":
        cp = lambda t: t.detach().clone().requires_grad_(requires_grad)
        return Field(
            tuple([cp(p) for p in g] for g in self.planes),
            {k: cp(v) for k, v in self.dec.items()},
            cp(self.beta),
            self.bound.clone(),
        )


def rounded_bound(bound, bound_dividable: float, scale: float = 1.0) -> torch.Tensor:
    """fp32 round-up of the upper bound to a multiple of `bound_dividable`.  ESLAM.py:159-173."""
    b = torch.from_numpy(np.array(bound) * scale).float()
    b[:, 1] = (((b[:, 1] - b[:, 0]) / bound_dividable).int() + 1) * bound_dividable + b[:, 0]
    return b


def plane_shapes(bound: torch.Tensor, res: float) -> Tuple[Tuple[int, int], ...]:
    """(H,W) of the xy, xz, yz planes at one resolution.  ESLAM.py:196-203: fp32
    length / res truncated with int(), then the (x,z) swap."""
    nx, ny, nz = map(int, ((bound[:, 1] - bound[:, 0]) / res).tolist())
    return (ny, nx), (nz, nx), (nz, ny)


def make_field(bound: torch.Tensor, planes_res=(0.24, 0.06), c_planes_res=(0.24, 0.03), c_dim=32,
               hidden=16, generator: Optional[torch.Generator] = None, std=0.01, dec_scale=1.0) -> Field:
    """Random-init field with the reference's shapes and distributions
    (planes N(0,0.01) ESLAM.py:201-210; nn.Linear default init decoders.py:46-56; beta=10 :58-61)."""
    g = generator

    def plane(hw):
        return torch.empty(1, c_dim, *hw).normal_(0, std, generator=g)

    groups: List[List[torch.Tensor]] = [[], [], [], [], [], []]
    for r in planes_res:
        for k, hw in enumerate(plane_shapes(bound, r)):
            groups[k].append(plane(hw))
    for r in c_planes_res:
        for k, hw in enumerate(plane_shapes(bound, r)):
            groups[3 + k].append(plane(hw))

    def linear(o, i):
        lim = 1.0 / math.sqrt(i)
        w = (torch.rand(o, i, generator=g) * 2 - 1) * lim * dec_scale
        b = (torch.rand(o, generator=g) * 2 - 1) * lim * dec_scale
        return w, b

    dec: Dict[str, torch.Tensor] = {}
    for pre, n_out in (("", 1), ("c_", 3)):
        dec[f"{pre}linears.0.weight"], dec[f"{pre}linears.0.bias"] = linear(hidden, 2 * c_dim)
        dec[f"{pre}linears.1.weight"], dec[f"{pre}linears.1.bias"] = linear(hidden, hidden)
        dec[f"{pre}output_linear.weight"], dec[f"{pre}output_linear.bias"] = linear(n_out, hidden)
    return Field(tuple(groups), dec, 10.0 * torch.ones(1), bound.clone())


# ----------------------------------------------------------------------------
# A2/A3: pixel pick and rays   (common.py:87-153)
# ----------------------------------------------------------------------------


def pick_pixels(H0, H1, W0, W1, n, depths, colors, draws):
    """One `randint` for all b images, with replacement; row bi uses
    indices[bi*n:(bi+1)*n].  common.py:101-139.  Returns i,j [b,n] f32, depth [b,n],
    colour [b,n,3] and the flat indices [b*n] into the cropped image."""
    b = depths.shape[0]
    dev = depths.device
    d = depths[:, H0:H1, W0:W1]
    c = colors[:, H0:H1, W0:W1]
    # i[r,c] = W0+c, j[r,c] = H0+r on the (H1-H0, W1-W0) crop, flattened row-major
    gi, gj = torch.meshgrid(torch.linspace(W0, W1 - 1, W1 - W0, device=dev),
                            torch.linspace(H0, H1 - 1, H1 - H0, device=dev), indexing="ij")
    gi = gi.t().reshape(-1)
    gj = gj.t().reshape(-1)
    idx = draws.randint(gi.shape[0], n * b)
    idx = idx.clamp(0, gi.shape[0])  # a no-op, kept: common.py:109
    i = gi[idx].reshape(b, -1)
    j = gj[idx].reshape(b, -1)
    idx_b = idx.reshape(b, -1)
    dsel = torch.gather(d.reshape(b, -1), 1, idx_b)
    csel = torch.gather(c.reshape(b, -1, 3), 1, idx_b.unsqueeze(-1).expand(-1, -1, 3))
    return i, j, dsel, csel, idx


def rays_from_pixels(i, j, c2ws, fx, fy, cx, cy):
    """dir_cam=((i-cx)/fx, -(j-cy)/fy, -1); rays_d = R.dir (un-normalised);
    rays_o = t.  common.py:87-99."""
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1)
    rays_d = torch.sum(dirs.unsqueeze(-2) * c2ws[:, None, :3, :3], -1)
    rays_o = c2ws[:, None, :3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def sample_rays(H0, H1, W0, W1, n, fx, fy, cx, cy, c2ws, depths, colors, draws):
    """get_samples (common.py:141-153): -> rays_o[b*n,3], rays_d[b*n,3], depth[b*n],
    colour[b*n,3] (fp64 if `colors` is), plus the pixel indices."""
    i, j, d, c, idx = pick_pixels(H0, H1, W0, W1, n, depths, colors, draws)
    ro, rd = rays_from_pixels(i, j, c2ws, fx, fy, cx, cy)
    return ro.reshape(-1, 3), rd.reshape(-1, 3), d.reshape(-1), c.reshape(-1, 3), idx


# ----------------------------------------------------------------------------
# A4: bounding-box pre-filter   (Tracker.py:175-187, Mapper.py:322-332)
# ----------------------------------------------------------------------------


def bbox_exit(rays_o, rays_d, bound):
    """Distance (in units of |d|) at which the ray leaves the bound box:
    min over axes of max over {lo,hi} of (bound-o)/d."""
    t = (bound.unsqueeze(0) - rays_o.detach().unsqueeze(-1)) / rays_d.detach().unsqueeze(-1)
    return torch.min(torch.max(t, dim=2)[0], dim=1)[0]


def bbox_keep(rays_o, rays_d, depth, bound, need_depth: bool):
    keep = bbox_exit(rays_o, rays_d, bound) >= depth
    if need_depth:  # tracker only, Tracker.py:182
        keep = keep & (depth > 0)
    return keep


# ----------------------------------------------------------------------------
# A6-A8: field decode   (common.py:204-218, decoders.py:64-146)
# ----------------------------------------------------------------------------


def normalize_pts(p, bound):
    """((p-lo)/(hi-lo))*2-1 per axis.  common.py:204-218."""
    p = p.reshape(-1, 3)
    cols = [((p[:, a] - bound[a, 0]) / (bound[a, 1] - bound[a, 0])) * 2 - 1.0 for a in range(3)]
    return torch.stack(cols, -1)


def plane_features(p_nor, pxy, pxz, pyz):
    """Bilinear (border, align_corners) fetch from the 3 planes of each scale,
    (xy+xz)+yz per scale, concat coarse|fine.  decoders.py:64-85."""
    grid = p_nor[None, :, None]
    feats = []
    for s in range(len(pxy)):
        taps = []
        for plane, axes in ((pxy[s], [0, 1]), (pxz[s], [0, 2]), (pyz[s], [1, 2])):
            v = F.grid_sample(plane, grid[..., axes], padding_mode="border", align_corners=True, mode="bilinear")
            taps.append(v.squeeze(0).squeeze(-1).transpose(0, 1))
        feats.append(taps[0] + taps[1] + taps[2])
    return torch.cat(feats, -1)


def _trunk(h, dec, pre):
    h = F.relu(F.linear(h, dec[f"{pre}linears.0.weight"], dec[f"{pre}linears.0.bias"]))
    h = F.relu(F.linear(h, dec[f"{pre}linears.1.weight"], dec[f"{pre}linears.1.bias"]))
    return F.linear(h, dec[f"{pre}output_linear.weight"], dec[f"{pre}output_linear.bias"])


def raw_sdf(p_nor, fld: Field):
    """64->16->16->1, tanh.  decoders.py:87-105."""
    f = plane_features(p_nor, fld.planes[0], fld.planes[1], fld.planes[2])
    return torch.tanh(_trunk(f, fld.dec, "")).squeeze(-1)


def raw_rgb(p_nor, fld: Field):
    """64->16->16->3, sigmoid.  decoders.py:107-125."""
    f = plane_features(p_nor, fld.planes[3], fld.planes[4], fld.planes[5])
    return torch.sigmoid(_trunk(f, fld.dec, "c_"))


def decode(p, fld: Field):
    """raw[...,4] = (r,g,b,sdf).  decoders.py:127-146."""
    shp = p.shape
    pn = normalize_pts(p, fld.bound)
    raw = torch.cat([raw_rgb(pn, fld), raw_sdf(pn, fld).unsqueeze(-1)], -1)
    return raw.reshape(*shp[:-1], -1)


# ----------------------------------------------------------------------------
# A5/A9: sampling along rays and compositing   (Renderer.py:46-153, common.py:41-77)
# ----------------------------------------------------------------------------


def sdf2alpha(sdf, beta):
    """1-exp(-beta*sigmoid(-sdf*beta)).  Renderer.py:149-153."""
    return 1.0 - torch.exp(-beta * torch.sigmoid(-sdf * beta))


def transmittance_weights(alpha):
    """w_k = alpha_k * prod_{j<k}(1-alpha_j+1e-10).  Renderer.py:128-129,141-142."""
    ones = torch.ones((alpha.shape[0], 1), device=alpha.device)
    return alpha * torch.cumprod(torch.cat([ones, 1.0 - alpha + 1e-10], -1), -1)[:, :-1]


def perturb(z, u):
    """Stratified jitter inside the midpoints' intervals.  Renderer.py:46-61."""
    mids = 0.5 * (z[..., 1:] + z[..., :-1])
    upper = torch.cat([mids, z[..., -1:]], -1)
    lower = torch.cat([z[..., :1], mids], -1)
    return lower + (upper - lower) * u


def inverse_cdf(bins, weights, u):
    """sample_pdf (common.py:41-77) with its quirks: the pdf is the UN-normalised
    weights (:47-48), searchsorted right=True, `above` clamped to the last cdf
    entry, denominators < 1e-5 replaced by 1."""
    cdf = torch.cumsum(weights, -1)
    cdf = torch.cat([torch.zeros_like(cdf[..., :1]), cdf], -1)
    inds = torch.searchsorted(cdf, u.contiguous(), right=True)
    below = (inds - 1).clamp(min=0)
    above = inds.clamp(max=cdf.shape[-1] - 1)
    c0, c1 = torch.gather(cdf, 1, below), torch.gather(cdf, 1, above)
    b0, b1 = torch.gather(bins, 1, below), torch.gather(bins, 1, above)
    denom = c1 - c0
    denom = torch.where(denom < 1e-5, torch.ones_like(denom), denom)
    return b0 + ((u - c0) / denom) * (b1 - b0)


def ray_depths(fld: Field, rays_o, rays_d, gt_depth, truncation, n_strat, n_imp, draws, perturb_on=True):
    """z_vals [R, n_strat+n_imp] (no grad).  Renderer.py:81-134."""
    R = rays_o.shape[0]
    dev = rays_o.device
    z = torch.empty(R, n_strat + n_imp, device=dev)
    t_uni = torch.linspace(0.0, 1.0, steps=n_strat, device=dev)
    t_surf = torch.linspace(0.0, 1.0, steps=n_imp, device=dev)
    d = gt_depth.reshape(-1, 1)
    has = (d > 0).squeeze(-1)
    dn = d[has]
    # rays with depth: surface band +-1.5*trunc and free space up to 1.2*d  (Renderer.py:94-106)
    z_surf = dn.expand(-1, n_imp) - (1.5 * truncation) + (3 * truncation * t_surf)
    z_free = 0.0 + 1.2 * dn.expand(-1, n_strat) * t_uni
    zs, _ = torch.sort(torch.cat([z_free, z_surf], -1), -1)
    if perturb_on:
        zs = perturb(zs, draws.rand(*zs.shape))
    z[has] = zs
    # depth-less rays: coarse pass + inverse-cdf resampling  (Renderer.py:108-134)
    if not bool(has.all()):
        with torch.no_grad():
            o, dd = rays_o[~has].detach(), rays_d[~has].detach()
            tb = (fld.bound.unsqueeze(0) - o.unsqueeze(-1)) / dd.unsqueeze(-1)
            far = torch.min(torch.max(tb, dim=2)[0], dim=1)[0].unsqueeze(-1)
            far = far + 0.01
            zu = 0.0 * (1.0 - t_uni) + far * t_uni
            if perturb_on:
                zu = perturb(zu, draws.rand(*zu.shape))
            pts = o.unsqueeze(1) + dd.unsqueeze(1) * zu.unsqueeze(-1)
            s = raw_sdf(normalize_pts(pts.clone(), fld.bound), fld).reshape(pts.shape[0], pts.shape[1])
            w = transmittance_weights(sdf2alpha(s, fld.beta))
            mid = 0.5 * (zu[..., 1:] + zu[..., :-1])
            wmid = w[..., 1:-1]
            zi = inverse_cdf(mid, wmid, draws.rand(wmid.shape[0], n_imp))
            zu, _ = torch.sort(torch.cat([zu, zi], -1), -1)
            z[~has] = zu
    return z


def composite(fld: Field, rays_o, rays_d, z):
    """pts -> raw -> alpha -> weights -> depth, rgb.  Renderer.py:136-147."""
    pts = rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]
    raw = decode(pts, fld)
    w = transmittance_weights(sdf2alpha(raw[..., -1], fld.beta))
    rgb = torch.sum(w[..., None] * raw[..., :3], -2)
    depth = torch.sum(w * z, -1)
    return depth, rgb, raw[..., -1], w, raw


def render_rays(fld: Field, rays_o, rays_d, gt_depth, truncation, n_strat, n_imp, draws, perturb_on=True):
    """Renderer.render_batch_ray (Renderer.py:63-147) -> depth[R], rgb[R,3], sdf[R,S], z[R,S]."""
    z = ray_depths(fld, rays_o, rays_d, gt_depth, truncation, n_strat, n_imp, draws, perturb_on)
    depth, rgb, sdf, _, _ = composite(fld, rays_o, rays_d, z)
    return depth, rgb, sdf, z


# ----------------------------------------------------------------------------
# A10: losses   (Tracker.py:114-148,192-204; Mapper.py:110-144,337-346)
# ----------------------------------------------------------------------------


def sdf_band_masks(z, d, truncation):
    """front / center / tail boolean masks [R,S] from z vs gt depth d[R]."""
    dd = d[:, None]
    front = z < (dd - truncation)
    back = z > (dd + truncation)
    center = (z > (dd - 0.4 * truncation)) & (z < (dd + 0.4 * truncation))
    tail = (~front) & (~back) & (~center)
    return front, center, tail


def sdf_loss(sdf, z, d, truncation, w_fs, w_center, w_tail):
    front, center, tail = sdf_band_masks(z, d, truncation)
    fs = torch.mean(torch.square(sdf[front] - 1.0))
    resid = z + sdf * truncation - d[:, None]
    ce = torch.mean(torch.square(resid[center]))
    ta = torch.mean(torch.square(resid[tail]))
    return w_fs * fs + w_center * ce + w_tail * ta


@dataclass
class LossWeights:
    fs: float
    center: float
    tail: float
    depth: float
    color: float


TRACK_W = LossWeights(10, 200, 50, 1, 5)  # ESLAM.yaml:29-33
MAP_W = LossWeights(5, 200, 10, 0.1, 5)  # ESLAM.yaml:53-57


def tracking_loss(depth, rgb, sdf, z, gt_d, gt_c, truncation, w: LossWeights):
    """Tracker.py:192-204: lower-median outlier mask (10x) applied to every term."""
    err = (gt_d - depth.detach()).abs()
    mask = err < 10 * err.median()
    loss = sdf_loss(sdf[mask], z[mask], gt_d[mask], truncation, w.fs, w.center, w.tail)
    loss = loss + w.color * torch.square(gt_c - rgb)[mask].mean()
    loss = loss + w.depth * torch.square(gt_d[mask] - depth[mask]).mean()
    return loss, mask


def mapping_loss(depth, rgb, sdf, z, gt_d, gt_c, truncation, w: LossWeights):
    """Mapper.py:337-346: sdf and depth terms over depth>0 rays, colour over all rays."""
    mask = gt_d > 0
    loss = sdf_loss(sdf[mask], z[mask], gt_d[mask], truncation, w.fs, w.center, w.tail)
    loss = loss + w.color * torch.square(gt_c - rgb).mean()
    loss = loss + w.depth * torch.square(gt_d[mask] - depth[mask]).mean()
    return loss, mask


# ----------------------------------------------------------------------------
# A11: Adam   (torch/optim/adam.py single-tensor path; Tracker.py:291-296, Mapper.py:288-306)
# ----------------------------------------------------------------------------


def adam_update(p, g, m, v, step: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8):
    """In-place on p,m,v.  m.lerp_(g,1-b1); v=b2*v+(1-b2)g^2;
    p -= (lr/(1-b1^t)) * m / (sqrt(v)/sqrt(1-b2^t)+eps)."""
    m.lerp_(g, 1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


# ----------------------------------------------------------------------------
# whole iterations
# ----------------------------------------------------------------------------


@dataclass
class Camera:
    H: int
    W: int
    fx: float
    fy: float
    cx: float
    cy: float


@dataclass
class RenderCfg:
    n_stratified: int = 32
    n_importance: int = 8
    truncation: float = 0.06
    perturb: bool = True


@dataclass
class IterOut:
    loss: torch.Tensor
    idx: torch.Tensor
    keep: torch.Tensor
    rays_o: torch.Tensor
    rays_d: torch.Tensor
    gt_depth: torch.Tensor
    gt_color: torch.Tensor
    depth: torch.Tensor
    rgb: torch.Tensor
    sdf: torch.Tensor
    z: torch.Tensor
    mask: torch.Tensor
    extra: dict = _dc_field(default_factory=dict)


def tracking_forward(fld: Field, cam: Camera, rc: RenderCfg, w: LossWeights, cam_pose, gt_color, gt_depth,
                     n_pixels, edge_h, edge_w, draws) -> IterOut:
    """Tracker.optimize_tracking up to the loss (Tracker.py:150-204).  `cam_pose`
    [1,7] may require grad; planes/decoders are used as given."""
    c2w = cam_pose_to_matrix(cam_pose)
    ro, rd, d, c, idx = sample_rays(edge_h, cam.H - edge_h, edge_w, cam.W - edge_w, n_pixels,
                                    cam.fx, cam.fy, cam.cx, cam.cy, c2w, gt_depth, gt_color, draws)
    with torch.no_grad():
        keep = bbox_keep(ro, rd, d, fld.bound, need_depth=True)
    ro, rd, d, c = ro[keep], rd[keep], d[keep], c[keep]
    depth, rgb, sdf, z = render_rays(fld, ro, rd, d, rc.truncation, rc.n_stratified, rc.n_importance, draws, rc.perturb)
    loss, mask = tracking_loss(depth, rgb, sdf, z, d, c, rc.truncation, w)
    return IterOut(loss, idx, keep, ro, rd, d, c, depth, rgb, sdf, z, mask)


def mapping_forward(fld: Field, cam: Camera, rc: RenderCfg, w: LossWeights, c2ws, gt_colors, gt_depths,
                    pix_per_image, draws) -> IterOut:
    """One iteration of Mapper.optimize_mapping up to the loss (Mapper.py:318-346).
    `c2ws` [b,4,4] (may carry grad through cam_pose_to_matrix for joint_opt)."""
    ro, rd, d, c, idx = sample_rays(0, cam.H, 0, cam.W, pix_per_image, cam.fx, cam.fy, cam.cx, cam.cy,
                                    c2ws, gt_depths, gt_colors, draws)
    with torch.no_grad():
        keep = bbox_keep(ro, rd, d, fld.bound, need_depth=False)
    ro, rd, d, c = ro[keep], rd[keep], d[keep], c[keep]
    depth, rgb, sdf, z = render_rays(fld, ro, rd, d, rc.truncation, rc.n_stratified, rc.n_importance, draws, rc.perturb)
    loss, mask = mapping_loss(depth, rgb, sdf, z, d, c, rc.truncation, w)
    return IterOut(loss, idx, keep, ro, rd, d, c, depth, rgb, sdf, z, mask)


def track_frame(fld: Field, cam: Camera, rc: RenderCfg, w: LossWeights, init_pose, gt_color, gt_depth,
                n_pixels, edge_h, edge_w, iters, lr_T, lr_R, draws):
    """The camera-iteration loop of Tracker.run (Tracker.py:291-309): Adam(betas .5,.999)
    over T and R, best pose = the pose BEFORE the step with the lowest loss."""
    T = torch.nn.Parameter(init_pose[:, -3:].clone())
    Rq = torch.nn.Parameter(init_pose[:, :4].clone())
    opt = torch.optim.Adam([{"params": [T], "lr": lr_T, "betas": (0.5, 0.999)},
                            {"params": [Rq], "lr": lr_R, "betas": (0.5, 0.999)}])
    best, best_pose, losses = float("inf"), None, []
    for _ in range(iters):
        pose = torch.cat([Rq, T], -1)
        out = tracking_forward(fld, cam, rc, w, pose, gt_color, gt_depth, n_pixels, edge_h, edge_w, draws)
        opt.zero_grad()
        out.loss.backward()
        opt.step()
        lv = out.loss.item()
        losses.append(lv)
        if lv < best:
            best, best_pose = lv, pose.clone().detach()
    return best_pose, torch.cat([Rq, T], -1).detach(), losses


def map_window(fld: Field, cam: Camera, rc: RenderCfg, w: LossWeights, c2ws, gt_colors, gt_depths, n_pixels, iters,
               lr_dec, lr_planes, lr_cplanes, joint_opt, lr_cam, draws, state_out=None):
    """The per-call body of Mapper.optimize_mapping (Mapper.py:249-362) for an
    already-selected window: fresh Adam, per-group lrs, first pose fixed.
    Updates `fld` in place; returns (updated c2ws [b,4,4], losses).  state_out (dict): receives the optimiser's
    final exp_avg / exp_avg_sq of the 12 planes (leaf order of Field.leaves()) for the post-Adam parity tests."""
    b = c2ws.shape[0]
    pix = n_pixels // b
    dec_params = [fld.dec[k].requires_grad_(True) for k in DECODER_KEYS] + [fld.beta.requires_grad_(True)]
    pl = [p.requires_grad_(True) for g in fld.planes[:3] for p in g]
    cpl = [p.requires_grad_(True) for g in fld.planes[3:] for p in g]
    groups = [{"params": dec_params, "lr": lr_dec}, {"params": pl, "lr": lr_planes}, {"params": cpl, "lr": lr_cplanes}]
    if joint_opt:
        poses = torch.nn.Parameter(matrix_to_cam_pose(c2ws[1:]))
        groups.append({"params": [poses], "lr": lr_cam})
    opt = torch.optim.Adam(groups)
    losses = []
    for _ in range(iters):
        cw = torch.cat([c2ws[0:1], cam_pose_to_matrix(poses)], 0) if joint_opt else c2ws
        out = mapping_forward(fld, cam, rc, w, cw, gt_colors, gt_depths, pix, draws)
        opt.zero_grad()
        out.loss.backward()
        opt.step()
        losses.append(out.loss.item())
    if state_out is not None:
        leaves = [p for g in fld.planes for p in g]
        state_out["exp_avg"] = [opt.state[p]["exp_avg"].clone() for p in leaves]
        state_out["exp_avg_sq"] = [opt.state[p]["exp_avg_sq"].clone() for p in leaves]
    for t in dec_params + pl + cpl:
        t.requires_grad_(False)
    if joint_opt:
        c2ws = torch.cat([c2ws[0:1], cam_pose_to_matrix(poses.detach())], 0)
    return c2ws, losses


# ----------------------------------------------------------------------------
# A12: dense grid query for meshing   (Mesher.py:130-186)
# ----------------------------------------------------------------------------


def grid_axes(mc_bound, resolution, padding=0.05):
    """Per-axis sample positions (float64 numpy) of get_grid_uniform (Mesher.py:159-177)."""
    axes = []
    mc = torch.as_tensor(np.array(mc_bound))
    for a in range(3):
        n = int(((mc[a][1] - mc[a][0] + 2 * padding) / resolution).round().int().item())
        axes.append(np.linspace(float(mc[a][0]) - padding, float(mc[a][1]) + padding, n))
    return axes


def grid_points(axes):
    """Flat point list in the reference's order: meshgrid(indexing='xy'), so
    flat = (iy*nx + ix)*nz + iz.  Mesher.py:179-184."""
    xt, yt, zt = (torch.from_numpy(a).float() for a in axes)
    gx, gy, gz = torch.meshgrid(xt, yt, zt, indexing="xy")
    return torch.stack([gx.reshape(-1), gy.reshape(-1), gz.reshape(-1)], 1)


def query_points(fld: Field, p):
    """eval_points (Mesher.py:130-157): decode, then force sdf=-1 outside the OPEN bound box."""
    b = fld.bound
    inside = ((p[:, 0] < b[0][1]) & (p[:, 0] > b[0][0]) & (p[:, 1] < b[1][1]) & (p[:, 1] > b[1][0])
              & (p[:, 2] < b[2][1]) & (p[:, 2] > b[2][0]))
    with torch.no_grad():
        ret = decode(p, fld)
    ret[~inside, -1] = -1
    return ret


def query_lattice_factored(fld: Field, axes):
    """CPU restatement of the product's FACTORED lattice query (csrc/render.cuh: k_grid_preact + k_grid_sdf_factored),
    kept here so the algebra is pinned against the reference's golden lattice values without a GPU.
    On the regular lattice of Mesher.py:159-186 every tap of decoders.py:64-85 depends on two lattice indices, and the
    first layer of decoders.py:87-105 is linear in the summed feature, so
        W1 (Fxy + Fxz + Fyz) + b1 = (W1 Fxy + b1) + W1 Fxz + W1 Fyz
    with each term resampled once on a face of the lattice.  Returns sdf in the flat order of grid_points."""
    b = fld.bound
    t = [torch.from_numpy(np.asarray(a)).float() for a in axes]
    nor = [((t[a] - b[a, 0]) / (b[a, 1] - b[a, 0])) * 2 - 1.0 for a in range(3)]
    W1, b1 = fld.dec["linears.0.weight"], fld.dec["linears.0.bias"]

    def face(planes, u, v):  # [len(v), len(u), 16]; u runs along the plane's W axis, v along H
        gu, gv = torch.meshgrid(u, v, indexing="xy")
        grid = torch.stack([gu, gv], -1)[None]
        acc = 0
        for s, plane in enumerate(planes):
            f = F.grid_sample(plane, grid, padding_mode="border", align_corners=True, mode="bilinear")[0]
            acc = acc + torch.einsum("oc,chw->hwo", W1[:, s * 32:(s + 1) * 32], f)
        return acc

    with torch.no_grad():
        pxy = face(fld.planes[0], nor[0], nor[1]) + b1  # [ny, nx, 16]
        pxz = face(fld.planes[1], nor[0], nor[2])       # [nz, nx, 16]
        pyz = face(fld.planes[2], nor[1], nor[2])       # [nz, ny, 16]
        pre = (pxy[:, :, None] + pxz.permute(1, 0, 2)[None]) + pyz.permute(1, 0, 2)[:, None]  # [ny, nx, nz, 16]
        h = F.relu(pre)
        h = F.relu(F.linear(h, fld.dec["linears.1.weight"], fld.dec["linears.1.bias"]))
        sdf = torch.tanh(F.linear(h, fld.dec["output_linear.weight"], fld.dec["output_linear.bias"])).reshape(-1)
    p = grid_points(axes)
    inside = ((p[:, 0] < b[0][1]) & (p[:, 0] > b[0][0]) & (p[:, 1] < b[1][1]) & (p[:, 1] > b[1][0])
              & (p[:, 2] < b[2][1]) & (p[:, 2] > b[2][0]))
    sdf[~inside] = -1
    return sdf


# ---------------------------------------------------------------------------------------
# keyframe selection by view overlap
# ---------------------------------------------------------------------------------------
def keyframe_overlap(cam: Camera, c2w, gt_depth, gt_color, kf_c2ws, draws, num_samples=8, num_rays=50):
    """Mapper.keyframe_selection_overlap (Mapper.py:146-203) up to `percent_inside`: `num_rays` random pixels of
    the current frame with depth > 0, `num_samples` points per ray between 0.8*d and d+0.5, projected into every
    keyframe of kf_c2ws [K,4,4] (the caller drops the last two keyframes, Mapper.py:180); fraction of points that
    land inside the image with a 20-pixel margin and in front of the camera.  The reference then keeps
    nonzero(percent_inside) in a random (CPU randperm) order (Mapper.py:205-209)."""
    H, W, fx, fy, cx, cy = cam.H, cam.W, cam.fx, cam.fy, cam.cx, cam.cy
    rays_o, rays_d, d, _, _ = sample_rays(0, H, 0, W, num_rays, fx, fy, cx, cy, c2w.unsqueeze(0),
                                          gt_depth.unsqueeze(0), gt_color.unsqueeze(0), draws)
    d = d.reshape(-1, 1)
    ok = d[:, 0] > 0
    rays_o, rays_d, d = rays_o[ok], rays_d[ok], d[ok].repeat(1, num_samples)
    t = torch.linspace(0., 1., steps=num_samples).to(d.device)
    z = (d * 0.8) * (1. - t) + (d + 0.5) * t
    pts = (rays_o[..., None, :] + rays_d[..., None, :] * z[..., :, None]).reshape(1, -1, 3)
    w2cs = torch.inverse(kf_c2ws)
    ones = torch.ones_like(pts[..., 0]).reshape(1, -1, 1)
    homo = torch.cat([pts, ones], dim=-1).reshape(1, -1, 4, 1).expand(w2cs.shape[0], -1, -1, -1)
    cam_pts = (w2cs.unsqueeze(1).expand(-1, homo.shape[1], -1, -1) @ homo)[:, :, :3]
    K = torch.tensor([[fx, .0, cx], [.0, fy, cy], [.0, .0, 1.0]], device=d.device).reshape(3, 3)
    cam_pts[:, :, 0] *= -1
    uv = K @ cam_pts
    zc = uv[:, :, -1:] + 1e-5
    uv = uv[:, :, :2] / zc
    edge = 20
    mask = (uv[:, :, 0] < W - edge) * (uv[:, :, 0] > edge) * (uv[:, :, 1] < H - edge) * (uv[:, :, 1] > edge)
    mask = (mask & (zc[:, :, 0] < 0)).squeeze(-1)
    return mask.sum(dim=1) / uv.shape[1], mask.sum(dim=1), uv.shape[1]


# ---------------------------------------------------------------------------------------
# full-image inference
# ---------------------------------------------------------------------------------------
def image_rays(cam: Camera, c2w):
    """get_rays (common.py:183-201): rays of every pixel, row-major, [H*W,3] each."""
    i, j = torch.meshgrid(torch.linspace(0, cam.W - 1, cam.W), torch.linspace(0, cam.H - 1, cam.H), indexing="ij")
    i, j = i.t(), j.t()
    dirs = torch.stack([(i - cam.cx) / cam.fx, -(j - cam.cy) / cam.fy, -torch.ones_like(i)], -1).to(c2w.device)
    dirs = dirs.reshape(cam.H, cam.W, 1, 3)
    rays_d = torch.sum(dirs * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o.reshape(-1, 3), rays_d.reshape(-1, 3)


def render_image(fld: Field, cam: Camera, c2w, gt_depth, truncation, n_strat, n_imp, draws, ray_batch_size=10000):
    """Renderer.render_img (Renderer.py:155-204): chunks of ray_batch_size rays through render_batch_ray with the
    perturbation ON (ESLAM.yaml:74); depth [H,W] float64, colour [H,W,3] float32."""
    with torch.no_grad():
        ro, rd = image_rays(cam, c2w)
        d = gt_depth.reshape(-1)
        depths, colors = [], []
        for i in range(0, rd.shape[0], ray_batch_size):
            dep, col, _, _ = render_rays(fld, ro[i:i + ray_batch_size], rd[i:i + ray_batch_size],
                                         d[i:i + ray_batch_size], truncation, n_strat, n_imp, draws)
            depths.append(dep.double())
            colors.append(col)
        return torch.cat(depths, 0).reshape(cam.H, cam.W), torch.cat(colors, 0).reshape(cam.H, cam.W, 3)


# ---------------------------------------------------------------------------------------
# frame ingest
# ---------------------------------------------------------------------------------------
def ingest_frame(color_u8_bgr, depth_u16, png_depth_scale, crop_edge=0, scale=1.0):
    """The arithmetic of BaseDataset.__getitem__ (datasets.py:88-112) after cv2.imread, for same-size colour and
    depth (no undistortion / resize): numpy arrays in, (colour [H',W',3] float64 RGB, depth [H',W'] float32) out."""
    import numpy as np

    color = color_u8_bgr[:, :, ::-1] / 255.                       # cvtColor(BGR2RGB); uint8 / float -> float64
    depth = depth_u16.astype(np.float32) / png_depth_scale
    color = torch.from_numpy(np.ascontiguousarray(color))
    depth = torch.from_numpy(depth) * scale
    if crop_edge > 0:
        color = color[crop_edge:-crop_edge, crop_edge:-crop_edge]
        depth = depth[crop_edge:-crop_edge, crop_edge:-crop_edge]
    return color.contiguous(), depth.contiguous()


def _fma(a, b, c):
    """round(a * b + c) with ONE rounding, exactly (python has no math.fma before 3.13)."""
    from fractions import Fraction

    return float(Fraction(a) * Fraction(b) + Fraction(c))


def _linear_taps(n_src, n_dst):
    """Source index pair and float64 weight of every destination index of cv2.resize(INTER_LINEAR) along one axis."""
    import numpy as np

    f = (np.arange(n_dst) + 0.5) * (n_src / n_dst) - 0.5
    s = np.floor(f).astype(np.int64)
    return np.clip(s, 0, n_src - 1), np.clip(s + 1, 0, n_src - 1), f - s


def ingest_frame_resized(color_u8_bgr, depth_u16, png_depth_scale, crop_edge=0, scale=1.0):
    """BaseDataset.__getitem__ (datasets.py:88-112) when the colour image is larger than the depth image (ScanNet:
    1296x968 colour, 640x480 depth): colour / 255 in float64, cv2.resize to the depth's size, crop_edge.
    cv2.resize(INTER_LINEAR) of a float64 image is third-party arithmetic: the opencv-python wheels (x86-64, this
    container's 4.13 and the reference's environment alike) route it to Intel IPP, whose result this restates bit for
    bit (found by comparing against cv2 here): weights in float64, a row pass then a column pass, each tap
    fma(w, b - a, a).  An OpenCV built WITHOUT IPP uses float32 weights (differences up to ~4e-7)."""
    import numpy as np

    color = color_u8_bgr[:, :, ::-1] / 255.
    H, W = depth_u16.shape
    hs, ws = color.shape[:2]
    x0, x1, fx = _linear_taps(ws, W)
    y0, y1, fy = _linear_taps(hs, H)
    rows = sorted(set(y0.tolist()) | set(y1.tolist()))
    hor = {}
    for r in rows:
        hor[r] = np.array([[_fma(fx[c], color[r, x1[c], ch] - color[r, x0[c], ch], color[r, x0[c], ch]) for ch in range(3)]
                           for c in range(W)])
    out = np.empty((H, W, 3))
    for r in range(H):
        a, b = hor[int(y0[r])], hor[int(y1[r])]
        for c in range(W):
            for ch in range(3):
                out[r, c, ch] = _fma(fy[r], b[c, ch] - a[c, ch], a[c, ch])
    depth = depth_u16.astype(np.float32) / png_depth_scale
    color = torch.from_numpy(np.ascontiguousarray(out))
    depth = torch.from_numpy(depth) * scale
    if crop_edge > 0:
        color = color[crop_edge:-crop_edge, crop_edge:-crop_edge]
        depth = depth[crop_edge:-crop_edge, crop_edge:-crop_edge]
    return color.contiguous(), depth.contiguous()


def undistort_u8(bgr_u8, fx, fy, cx, cy, distortion):
    """cv2.undistort(img, K, distortion) of an 8-bit image with the new camera matrix = K (datasets.py:83-86; third
    party: OpenCV's initUndistortRectifyMap + remap(INTER_LINEAR, BORDER_CONSTANT), restated from their published
    algorithm and pinned bit for bit against cv2 4.13 here): the source position of every pixel in float64, rounded
    to 1/32 pixel, the four taps blended with integer weights 32 (32 - ax)(32 - ay) ... that sum to 2^15, rounded
    by + 2^14 >> 15; taps outside the image are 0.  distortion = (k1, k2, p1, p2, k3)."""
    import numpy as np

    k1, k2, p1, p2, k3 = [float(v) for v in distortion]
    H, W = bgr_u8.shape[:2]
    K = np.array([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]])
    ir = np.linalg.inv(K)
    i = np.arange(H, dtype=np.float64)[:, None]
    j = np.arange(W, dtype=np.float64)[None, :]
    _x = i * ir[0, 1] + ir[0, 2] + j * ir[0, 0]
    _y = i * ir[1, 1] + ir[1, 2] + j * ir[1, 0]
    _w = i * ir[2, 1] + ir[2, 2] + j * ir[2, 0]
    w = 1.0 / _w
    x, y = _x * w, _y * w
    x2, y2 = x * x, y * y
    r2, _2xy = x2 + y2, 2 * x * y
    kr = 1 + ((k3 * r2 + k2) * r2 + k1) * r2
    xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2)
    yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy
    iu = np.rint((fx * xd + cx) * 32).astype(np.int64)
    iv = np.rint((fy * yd + cy) * 32).astype(np.int64)
    sx, sy, ax, ay = iu >> 5, iv >> 5, iu & 31, iv & 31

    def tap(yy, xx):
        ok = (yy >= 0) & (yy < H) & (xx >= 0) & (xx < W)
        v = bgr_u8[np.clip(yy, 0, H - 1), np.clip(xx, 0, W - 1)].astype(np.int64)
        v[~ok] = 0
        return v

    acc = (tap(sy, sx) * ((32 - ay) * (32 - ax) * 32)[..., None] + tap(sy, sx + 1) * ((32 - ay) * ax * 32)[..., None] +
           tap(sy + 1, sx) * (ay * (32 - ax) * 32)[..., None] + tap(sy + 1, sx + 1) * (ay * ax * 32)[..., None])
    return ((acc + (1 << 14)) >> 15).astype(np.uint8)


def _align_corners_taps(n_in, n_out):
    """Index pair and weights of F.interpolate(mode='bilinear', align_corners=True) along one axis, as ATen's CPU kernel
    computes them: position = i (n_in - 1) / (n_out - 1) in float64, the lower index from floorf() of the position
    ROUNDED TO FLOAT32 (so a position a hair below an integer already belongs to it), weights from the float64 rest."""
    import numpy as np

    scale = (n_in - 1) / (n_out - 1) if n_out > 1 else 0.0
    real = scale * np.arange(n_out, dtype=np.float64)
    i0 = np.minimum(np.floor(real.astype(np.float32)).astype(np.int64), n_in - 1)
    l1 = np.clip(real - i0, 0.0, 1.0)
    return i0, i0 + (i0 < n_in - 1), 1.0 - l1, l1


def resize_bilinear_align_corners(img, size):
    """F.interpolate(img.permute(2,0,1)[None], size, mode='bilinear', align_corners=True) for a float64 [H,W,C] image
    (datasets.py:100-104), bit for bit as torch's CPU kernel evaluates it here: the four weight products, then
    (w01 b) -> fma(w00, a, .) -> fma(w10, c, .) -> fma(w11, d, .)."""
    import numpy as np

    Hi, Wi, C = img.shape
    Ho, Wo = size
    y0, y1, hy, ly = _align_corners_taps(Hi, Ho)
    x0, x1, hx, lx = _align_corners_taps(Wi, Wo)
    out = np.empty((Ho, Wo, C))
    for r in range(Ho):
        for c in range(Wo):
            w00, w01, w10, w11 = hy[r] * hx[c], hy[r] * lx[c], ly[r] * hx[c], ly[r] * lx[c]
            for ch in range(C):
                acc = w01 * img[y0[r], x1[c], ch]
                acc = _fma(w00, img[y0[r], x0[c], ch], acc)
                acc = _fma(w10, img[y1[r], x0[c], ch], acc)
                out[r, c, ch] = _fma(w11, img[y1[r], x1[c], ch], acc)
    return out


def _nearest_index(n_in, n_out):
    """F.interpolate(mode='nearest'): min(floorf(i * (float32)(n_in / n_out)), n_in - 1)."""
    import numpy as np

    scale = np.float32(n_in) / np.float32(n_out)
    return np.minimum(np.floor(np.arange(n_out, dtype=np.float32) * scale).astype(np.int64), n_in - 1)


def ingest_frame_tum(color_u8_bgr, depth_u16, png_depth_scale, cam=None, distortion=None, crop_size=None, crop_edge=0,
                     scale=1.0):
    """BaseDataset.__getitem__ (datasets.py:79-112) for TUM-shaped frames (colour and depth of one size): optional
    cv2.undistort of the uint8 colour image (cam = (fx, fy, cx, cy)), / 255 in float64, optional crop_size (bilinear
    align_corners resize of the colour, nearest of the depth), crop_edge."""
    import numpy as np

    bgr = color_u8_bgr
    if distortion is not None:
        bgr = undistort_u8(bgr, *cam, distortion)
    color = bgr[:, :, ::-1] / 255.
    depth = torch.from_numpy(depth_u16.astype(np.float32) / png_depth_scale) * scale
    if crop_size is not None:
        color = resize_bilinear_align_corners(np.ascontiguousarray(color), crop_size)
        iy, ix = _nearest_index(depth.shape[0], crop_size[0]), _nearest_index(depth.shape[1], crop_size[1])
        depth = depth[torch.from_numpy(iy)][:, torch.from_numpy(ix)]
    color = torch.from_numpy(np.ascontiguousarray(color))
    if crop_edge > 0:
        color = color[crop_edge:-crop_edge, crop_edge:-crop_edge]
        depth = depth[crop_edge:-crop_edge, crop_edge:-crop_edge]
    return color.contiguous(), depth.contiguous()


# ----------------------------------------------------------------------------
# (f)-2: mesh extraction around the query   (Mesher.py:219-247, cull_mesh.py:58-105)
# ----------------------------------------------------------------------------
# skimage.measure.marching_cubes is third party (Lewiner's variant, environment.yaml: scikit-image 0.19.2) and absent
# here: PARITY UNPINNED against it.  What the reference's mesh and this restatement share by construction is the vertex
# set: one vertex on every lattice edge whose end values straddle the level, at the linear interpolation point.  The
# restatement below re-derives a cell's polygons from first principles PER CELL (contour segments on the six faces,
# chained into loops); it does not read the product's generated case tables, so the two can be compared.


def _mc_face_quads():
    quads = []
    for a in range(3):
        u, v = (a + 1) % 3, (a + 2) % 3
        for side in (0, 1):
            quad = [(0, 0), (1, 0), (1, 1), (0, 1)]
            if side == 0:
                quad = quad[::-1]
            quads.append([tuple((side if d == a else (bu if d == u else bv)) for d in range(3)) for bu, bv in quad])
    return quads


def marching_cubes_vertices(vol, level, xs, ys, zs):
    """{lattice edge: world position} of every level crossing; vol[ix, iy, iz].  The vertex set of Mesher.py:219-247."""
    out = {}
    P = (np.asarray(xs, np.float64), np.asarray(ys, np.float64), np.asarray(zs, np.float64))
    for axis in range(3):
        lo = [slice(None)] * 3
        hi = [slice(None)] * 3
        lo[axis], hi[axis] = slice(0, -1), slice(1, None)
        v0, v1 = vol[tuple(lo)], vol[tuple(hi)]
        for idx in np.argwhere((v0 < level) != (v1 < level)):
            i = tuple(int(x) for x in idx)
            a, b = float(v0[i]), float(v1[i])
            t = (level - a) / (b - a)
            p = [float(P[d][i[d]]) for d in range(3)]
            p[axis] = p[axis] + t * (float(P[axis][i[axis] + 1]) - p[axis])
            out[(i, axis)] = np.array(p)
    return out


def marching_cubes_loops(vol, level):
    """Per cell, the closed polygons of the iso-surface as lists of lattice edges ((ix,iy,iz), axis), oriented with the
    normal towards values >= level; ambiguous faces separate the inside (< level) corners."""
    quads = _mc_face_quads()
    nx, ny, nz = vol.shape
    inside = vol < level
    loops = {}
    for cx in range(nx - 1):
        for cy in range(ny - 1):
            for cz in range(nz - 1):
                blk = inside[cx:cx + 2, cy:cy + 2, cz:cz + 2]
                if blk.all() or not blk.any():
                    continue
                nxt = {}
                for quad in quads:
                    flags = [bool(blk[c]) for c in quad]
                    for x in range(4):
                        if not (flags[x] and not flags[(x + 1) % 4]):
                            continue
                        i = x
                        while not (not flags[(i - 1) % 4] and flags[i]):
                            i = (i - 1) % 4
                        i = (i - 1) % 4  # the boundary step that entered this inside arc

                        def edge(c0, c1):
                            axis = [d for d in range(3) if c0[d] != c1[d]][0]
                            lo = tuple(min(c0[d], c1[d]) for d in range(3))
                            return ((cx + lo[0], cy + lo[1], cz + lo[2]), axis)

                        nxt[edge(quad[x], quad[(x + 1) % 4])] = edge(quad[i], quad[(i + 1) % 4])
                cell, seen = [], set()
                for s in nxt:
                    if s in seen:
                        continue
                    loop, e = [], s
                    while e not in seen:
                        seen.add(e)
                        loop.append(e)
                        e = nxt[e]
                    cell.append(loop)
                loops[(cx, cy, cz)] = cell
    return loops


def cull_mask(points, depth, c2w, K, H, W, truncation, eval_rec):
    """Visibility of mesh vertices in one frame, cull_mesh.py:58-100 (torch ops as there)."""
    fx, fy, cx, cy = K
    w2c = torch.inverse(c2w)
    Km = torch.tensor([[fx, 0.0, cx], [0.0, fy, cy], [0.0, 0.0, 1.0]], dtype=torch.float32)
    ones = torch.ones_like(points[:, 0]).reshape(-1, 1)
    homo = torch.cat([points, ones], dim=1).reshape(-1, 4, 1).float()
    cam = (w2c @ homo)[:, :3]
    cam[:, 0] *= -1
    uv = Km @ cam.float()
    z = uv[:, -1:] + 1e-5
    uv = (uv[:, :2] / z).squeeze(-1)
    grid = uv[None, None].clone()
    grid[..., 0] = grid[..., 0] / W
    grid[..., 1] = grid[..., 1] / H
    grid = 2 * grid - 1
    ds = F.grid_sample(depth[None, None], grid, padding_mode="zeros", align_corners=True).squeeze()
    front = (0 <= -z[:, 0, 0]) & (uv[:, 0] < W) & (uv[:, 0] > 0) & (uv[:, 1] < H) & (uv[:, 1] > 0)
    if eval_rec:
        return (ds + truncation >= -z[:, 0, 0]) & front
    return front
