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
                   separable=None, factored=None):
    """SDF on the flat index range [start, start+count) of the marching-cubes lattice
    (flat = (iy*nx + ix)*nz + iz, Mesher.py:179-184), coordinates generated in-kernel.
    Shard over GPUs by giving each rank its own [start, count).  hull: optional [F,4] half-spaces (hull_planes) of
    the mesh bound; points outside get sdf = -1 in the same pass (Mesher.py:210-217).

    Three forms of the same query:
      direct     every voxel gathers its 24 plane corners (eslam_grid_sdf / _hull);
      separable  the plane features are resampled once on the lattice's faces and a voxel sums three of them:
                 bit-identical to the direct form (eslam_grid_features + eslam_grid_sdf_separable);
      factored   the decoder's (linear) first layer is applied on the faces as well, a voxel adds three 16-vectors and
                 runs the 16->16->1 tail: equal to the direct form up to the re-association of the first layer's sum
                 (tests: 1e-5; eslam_grid_preact + eslam_grid_sdf_factored).
    Default: factored (ESLAM_B200_FACTORED=0 selects separable, ESLAM_B200_SEPARABLE=0 as well selects direct);
    an explicit separable=True/False asks for that bit-exact form."""
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
    hull_p = ptr(hull) if hull is not None else None
    hull_n = hull.shape[0] if hull is not None else 0
    fits = max(nx, ny, nz) <= 32767 and 256 * (nx * ny + nx * nz + ny * nz) <= (4 << 30)
    if factored is None:
        factored = separable is None and os.environ.get("ESLAM_B200_FACTORED", "1") == "1"
    if separable is None:
        separable = os.environ.get("ESLAM_B200_SEPARABLE", "1") == "1"
    faces = None
    if factored and fits:
        mode = "factored"
        faces = [torch.empty(n, dtype=torch.float32, device=dev) for n in (ny * nx * 16, nx * nz * 16, ny * nz * 16)]
        call("eslam_grid_preact", store.ref(), ptr(store.arena), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz, ptr(faces[0]),
             ptr(faces[1]), ptr(faces[2]), stream())
    elif separable and fits:
        mode = "separable"
        faces = [torch.empty(b, a, 64, dtype=torch.float32, device=dev) for a, b in ((nx, ny), (nx, nz), (ny, nz))]
        call("eslam_grid_features", store.ref(), ptr(store.arena), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz, ptr(faces[0]),
             ptr(faces[1]), ptr(faces[2]), stream())
    else:
        mode = "direct" if hull is None else "hull"
    done = 0
    while done < count:
        n = min(chunk, count - done)
        dst = out[done:done + n].data_ptr()
        if mode == "factored":
            call("eslam_grid_sdf_factored", store.ref(), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz, start + done, n,
                 ptr(faces[0]), ptr(faces[1]), ptr(faces[2]), hull_p, hull_n, dst, stream())
        elif mode == "separable":
            call("eslam_grid_sdf_separable", store.ref(), ptr(store.arena), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz,
                 start + done, n, ptr(faces[0]), ptr(faces[1]), ptr(faces[2]), hull_p, hull_n, dst, stream())
        elif mode == "direct":
            call("eslam_grid_sdf", store.ref(), ptr(store.arena), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz, start + done,
                 n, dst, stream())
        else:
            call("eslam_grid_sdf_hull", store.ref(), ptr(store.arena), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz,
                 start + done, n, hull_p, hull_n, dst, stream())
        done += n
    return out
