"""Drop-in for the decoder-query half of `src/utils/Mesher.py` (reference lines 130-186):
`eval_points` on explicit points and the dense grid query that feeds marching cubes, with the
coordinates generated on the device and only the SDF head evaluated for the volume pass.
"""
from __future__ import annotations

import os

import numpy as np
import torch

from ._lib import call, ptr, stream
from .decoders import synced_store


def grid_axes(marching_cubes_bound, resolution, padding=0.05):
    """Per-axis sample positions of Mesher.get_grid_uniform (Mesher.py:159-177), float64 numpy."""
    mc = np.asarray(marching_cubes_bound, dtype=np.float64)
    axes = []
    for a in range(3):
        n = int(np.round((mc[a][1] - mc[a][0] + 2 * padding) / resolution))
        axes.append(np.linspace(mc[a][0] - padding, mc[a][1] + padding, n))
    return axes


def eval_points(p, all_planes, decoders, bound=None):
    """Mesher.eval_points (Mesher.py:130-157): raw[N,4] with sdf forced to -1 outside the OPEN bound box."""
    store = synced_store(all_planes, decoders, bound)
    pts = p.detach().reshape(-1, 3).float().contiguous()
    raw = torch.empty(pts.shape[0], 4, dtype=torch.float32, device=pts.device)
    call("eslam_decode_points", store.ref(), ptr(store.arena), ptr(pts), pts.shape[0], ptr(raw), 2, stream())
    return raw


def hull_planes(vertices, faces) -> torch.Tensor:
    """Outward half-spaces [F,4] = (n, d) with n.p + d <= 0 inside, of a CONVEX triangle mesh (the convex hull
    Mesher.get_bound_from_frames builds with Open3D, Mesher.py:63-128).  float64 numpy in, float32 tensor out."""
    v = np.asarray(vertices, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    n = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    n /= np.linalg.norm(n, axis=1, keepdims=True).clip(1e-30)
    d = -(n * v[f[:, 0]]).sum(1)
    centre = v.mean(0)
    flip = (n @ centre + d) > 0  # orient every face away from the centroid
    n[flip] *= -1
    d[flip] *= -1
    return torch.from_numpy(np.concatenate([n, d[:, None]], 1).astype(np.float32))


def query_grid_sdf(all_planes, decoders, axes, bound=None, start=0, count=None, chunk=1 << 24, out=None, hull=None,
                   separable=None):
    """SDF on the flat index range [start, start+count) of the marching-cubes lattice
    (flat = (iy*nx + ix)*nz + iz, Mesher.py:179-184), coordinates generated in-kernel.
    Shard over GPUs by giving each rank its own [start, count).  hull: optional [F,4] half-spaces (hull_planes) of
    the mesh bound; points outside get sdf = -1 in the same pass (Mesher.py:210-217).
    separable (default: on when the three resampled faces fit 4 GB): the plane features are resampled once on the
    lattice's faces and each voxel sums three of them (bit-identical values, eslam_grid_sdf_separable)."""
    store = synced_store(all_planes, decoders, bound)
    dev = store.device
    xs, ys, zs = (torch.from_numpy(np.asarray(a)).float().to(dev) for a in axes)
    nx, ny, nz = xs.numel(), ys.numel(), zs.numel()
    total = nx * ny * nz
    count = total - start if count is None else count
    if out is None:
        out = torch.empty(count, dtype=torch.float32, device=dev)
    if hull is not None:
        hull = hull.to(device=dev, dtype=torch.float32).contiguous()
    face_bytes = 256 * (nx * ny + nx * nz + ny * nz)
    if separable is None:
        separable = face_bytes <= (4 << 30) and os.environ.get("ESLAM_B200_SEPARABLE", "1") == "1"
    faces = None
    if separable and max(nx, ny, nz) <= 32767:
        faces = [torch.empty(b, a, 64, dtype=torch.float32, device=dev) for a, b in ((nx, ny), (nx, nz), (ny, nz))]
        call("eslam_grid_features", store.ref(), ptr(store.arena), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz, ptr(faces[0]),
             ptr(faces[1]), ptr(faces[2]), stream())
    done = 0
    while done < count:
        n = min(chunk, count - done)
        if faces is not None:
            call("eslam_grid_sdf_separable", store.ref(), ptr(store.arena), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz,
                 start + done, n, ptr(faces[0]), ptr(faces[1]), ptr(faces[2]), ptr(hull) if hull is not None else None,
                 hull.shape[0] if hull is not None else 0, out[done:done + n].data_ptr(), stream())
        elif hull is None:
            call("eslam_grid_sdf", store.ref(), ptr(store.arena), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz, start + done,
                 n, out[done:done + n].data_ptr(), stream())
        else:
            call("eslam_grid_sdf_hull", store.ref(), ptr(store.arena), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz,
                 start + done, n, ptr(hull), hull.shape[0], out[done:done + n].data_ptr(), stream())
        done += n
    return out
