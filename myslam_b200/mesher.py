"""Drop-in for the decoder-query half of `src/utils/Mesher.py` (reference lines 130-186):
`eval_points` on explicit points and the dense grid query that feeds marching cubes, with the
coordinates generated on the device and only the SDF head evaluated for the volume pass.
"""
from __future__ import annotations

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


def query_grid_sdf(all_planes, decoders, axes, bound=None, start=0, count=None, chunk=1 << 24, out=None):
    """SDF on the flat index range [start, start+count) of the marching-cubes lattice
    (flat = (iy*nx + ix)*nz + iz, Mesher.py:179-184), coordinates generated in-kernel.
    Shard over GPUs by giving each rank its own [start, count)."""
    store = synced_store(all_planes, decoders, bound)
    dev = store.device
    xs, ys, zs = (torch.from_numpy(np.asarray(a)).float().to(dev) for a in axes)
    nx, ny, nz = xs.numel(), ys.numel(), zs.numel()
    total = nx * ny * nz
    count = total - start if count is None else count
    if out is None:
        out = torch.empty(count, dtype=torch.float32, device=dev)
    done = 0
    while done < count:
        n = min(chunk, count - done)
        call("eslam_grid_sdf", store.ref(), ptr(store.arena), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz, start + done, n,
             out[done:done + n].data_ptr(), stream())
        done += n
    return out
