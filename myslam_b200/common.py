"""Host-side mirror of the parts of `src/common.py` the hot path's callers use (reference
lines 41-218).  Everything per-iteration lives in the kernels; these torch versions exist for the
once-per-frame / once-per-call callers (pose initialisation, keyframe selection, render_img) and so
the package carries no dependency on pytorch3d.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


# ---- pytorch3d.transforms stand-ins (real-first quaternions) -----------------------
def quaternion_to_matrix(q: torch.Tensor) -> torch.Tensor:
    """[...,4] -> [...,3,3] with the 2/|q|^2 scale (pytorch3d 0.7.1), so un-normalised q is fine."""
    r, i, j, k = torch.unbind(q, -1)
    s = 2.0 / (q * q).sum(-1)
    rows = (1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r),
            s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
            s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j))
    return torch.stack(rows, -1).reshape(q.shape[:-1] + (3, 3))


def matrix_to_quaternion(M: torch.Tensor) -> torch.Tensor:
    """[...,3,3] -> [...,4]: best-conditioned candidate (argmax |q_i|, 0.1 floor), sign not standardised."""
    lead = M.shape[:-2]
    m = M.reshape(lead + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = torch.unbind(m, -1)
    sq = torch.stack([1.0 + m00 + m11 + m22, 1.0 + m00 - m11 - m22, 1.0 - m00 + m11 - m22,
                      1.0 - m00 - m11 + m22], -1)
    qa = torch.where(sq > 0, torch.sqrt(sq.clamp_min(0)), torch.zeros_like(sq))
    cands = torch.stack([
        torch.stack([qa[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], -1),
        torch.stack([m21 - m12, qa[..., 1] ** 2, m10 + m01, m02 + m20], -1),
        torch.stack([m02 - m20, m10 + m01, qa[..., 2] ** 2, m12 + m21], -1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, qa[..., 3] ** 2], -1)], -2)
    cands = cands / (2.0 * qa[..., None].clamp_min(0.1))
    best = qa.argmax(-1)
    return torch.gather(cands, -2, best[..., None, None].expand(lead + (1, 4))).squeeze(-2)


def cam_pose_to_matrix(batch_poses: torch.Tensor) -> torch.Tensor:
    """[B,7] (quaternion, translation) -> [B,4,4] (common.py:169-181)."""
    B = batch_poses.shape[0]
    c2w = torch.eye(4, device=batch_poses.device).unsqueeze(0).repeat(B, 1, 1)
    c2w[:, :3, :3] = quaternion_to_matrix(batch_poses[:, :4])
    c2w[:, :3, 3] = batch_poses[:, 4:]
    return c2w


def matrix_to_cam_pose(batch_matrices: torch.Tensor, RT: bool = True) -> torch.Tensor:
    """[B,4,4] -> [B,7] (common.py:155-167)."""
    q = matrix_to_quaternion(batch_matrices[:, :3, :3])
    t = batch_matrices[:, :3, 3]
    return torch.cat([q, t], -1) if RT else torch.cat([t, q], -1)


# ---- rays ----------------------------------------------------------------------------------------------
def get_rays(H, W, fx, fy, cx, cy, c2w, device):
    """Rays of a whole image (common.py:183-201)."""
    if isinstance(c2w, np.ndarray):
        c2w = torch.from_numpy(c2w)
    c2w = c2w.to(device)
    jj, ii = torch.meshgrid(torch.linspace(0, H - 1, H, device=device), torch.linspace(0, W - 1, W, device=device),
                            indexing="ij")
    dirs = torch.stack([(ii - cx) / fx, -(jj - cy) / fy, -torch.ones_like(ii)], -1)
    rays_d = torch.sum(dirs.reshape(H, W, 1, 3) * c2w[:3, :3], -1)
    rays_o = c2w[:3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def get_rays_from_uv(i, j, c2ws, H, W, fx, fy, cx, cy, device):
    """Rays of chosen pixels (common.py:87-99)."""
    dirs = torch.stack([(i - cx) / fx, -(j - cy) / fy, -torch.ones_like(i)], -1).unsqueeze(-2)
    rays_d = torch.sum(dirs * c2ws[:, None, :3, :3], -1)
    rays_o = c2ws[:, None, :3, -1].expand(rays_d.shape)
    return rays_o, rays_d


def get_samples(H0, H1, W0, W1, n, H, W, fx, fy, cx, cy, c2ws, depths, colors, device):
    """n random pixels per image of the crop -> rays, depth, colour (common.py:101-153).  Torch version
    for the once-per-call callers (keyframe selection); the per-iteration path is eslam_sample_rays."""
    b = c2ws.shape[0]
    Wc, Hc = W1 - W0, H1 - H0
    idx = torch.randint(Hc * Wc, (n * b,), device=device)
    row = torch.div(idx, Wc, rounding_mode="floor")
    col = idx - row * Wc
    i = (W0 + col).float().reshape(b, -1)
    j = (H0 + row).float().reshape(b, -1)
    flat = ((H0 + row) * W + (W0 + col)).reshape(b, -1)
    d = torch.gather(depths.reshape(b, -1), 1, flat)
    c = torch.gather(colors.reshape(b, -1, 3), 1, flat.unsqueeze(-1).expand(-1, -1, 3))
    rays_o, rays_d = get_rays_from_uv(i, j, c2ws, H, W, fx, fy, cx, cy, device)
    return rays_o.reshape(-1, 3), rays_d.reshape(-1, 3), d.reshape(-1), c.reshape(-1, 3)


def normalize_3d_coordinate(p, bound):
    """((p-lo)/(hi-lo))*2-1 (common.py:204-218); out of place."""
    p = p.reshape(-1, 3)
    bound = bound.to(p.device)
    return ((p - bound[:, 0]) / (bound[:, 1] - bound[:, 0])) * 2 - 1.0


def random_select(l, k):
    """k random indices of range(l) (common.py:79-85)."""
    return list(np.random.permutation(np.array(range(l)))[:min(l, k)])
