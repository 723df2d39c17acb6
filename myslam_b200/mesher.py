"""Drop-in for `src/utils/Mesher.py` (reference lines 63-264) and `src/tools/cull_mesh.py:36-114` on the device:
`eval_points` on explicit points, the dense grid query that feeds marching cubes (coordinates generated in-kernel,
SDF head only, convex mesh bound as half-space tests in the same pass), marching cubes over the lattice, vertex colours,
frustum culling.  Only the hull of the keyframe points (Qhull via scipy, a few hundred thousand points once per mesh)
and the PLY writer run on the host.
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
        # only the lattice rows this flat range touches (flat = (iy*nx + ix)*nz + iz): a rank of a sharded query
        # resamples 1/world of the xy and yz faces instead of all of them
        iy0 = (start // nz) // nx
        iy1 = ((start + max(count, 1) - 1) // nz) // nx + 1
        call("eslam_grid_preact", store.ref(), ptr(store.arena), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz, iy0, iy1,
             ptr(faces[0]), ptr(faces[1]), ptr(faces[2]), stream())
    elif separable and fits:
        mode = "separable"
        faces = [torch.empty(b, a, 64, dtype=torch.float32, device=dev) for a, b in ((nx, ny), (nx, nz), (ny, nz))]
        call("eslam_grid_features", store.ref(), ptr(store.arena), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz, ptr(faces[0]),
             ptr(faces[1]), ptr(faces[2]), stream())
    else:
        mode = "direct" if hull is None else "hull"
    done = 0
    if mode == "factored" and os.environ.get("ESLAM_B200_GRID_ROWS", "1") == "1":
        # whole lattice rows go to the tensor-core form (eslam_grid_sdf_rows); what is left of a range that does not
        # start / end on a row boundary (a rank's shard) to the per-voxel form below
        row = nx * nz
        r0, r1 = -(-start // row), (start + count) // row
        if r1 > r0:
            def flat(lo, hi):
                if hi > lo:
                    call("eslam_grid_sdf_factored", store.ref(), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz, lo, hi - lo,
                         ptr(faces[0]), ptr(faces[1]), ptr(faces[2]), hull_p, hull_n, out[lo - start:hi - start].data_ptr(),
                         stream())

            store.bind()
            flat(start, r0 * row)
            call("eslam_grid_sdf_rows", store.ref(), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz, r0, r1, start,
                 ptr(faces[0]), ptr(faces[1]), ptr(faces[2]), hull_p, hull_n, ptr(out), stream())
            flat(r1 * row, start + count)
            return out
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


# ---------------------------------------------------------------------------------------------------------------------
# the rest of Mesher.get_mesh around the query (Mesher.py:188-264, cull_mesh.py:36-114) on the device
# ---------------------------------------------------------------------------------------------------------------------
_MC_TABLES = {}


def _mc_tables(dev):
    from . import mc_tables as T

    key = str(dev)
    if key not in _MC_TABLES:
        assert T.MAX_TRI == 5
        _MC_TABLES[key] = (torch.from_numpy(T.N_TRI.copy()).to(dev), torch.from_numpy(T.TRI_TABLE.copy()).to(dev).contiguous())
    return _MC_TABLES[key]


def marching_cubes(sdf, axes, level=0.0, weld=True):
    """Iso-surface of the lattice values sdf[(iy*nx + ix)*nz + iz] (the output of query_grid_sdf) at `level`, on the
    device (replaces the D2H copy of the volume + skimage.measure.marching_cubes, Mesher.py:219-243).
    Returns (vertices [V,3] float32 in WORLD coordinates, faces [F,3] int64), both on the device; normals point towards
    increasing values.  weld=True merges the vertices that lie on the same lattice edge (as skimage's output is);
    weld=False returns the triangle soup (faces = arange)."""
    from . import _lib

    dev = sdf.device
    xs, ys, zs = (torch.from_numpy(np.asarray(a)).float().to(dev) for a in axes)
    nx, ny, nz = xs.numel(), ys.numel(), zs.numel()
    if sdf.numel() != nx * ny * nz or sdf.dtype != torch.float32 or not sdf.is_contiguous():
        raise RuntimeError("marching_cubes: sdf must be the contiguous float32 lattice of the given axes")
    n_tri, tri = _mc_tables(dev)
    nb = int(_lib.load().eslam_mc_blocks(nx, ny, nz))
    counts = torch.empty(nb, dtype=torch.int32, device=dev)
    call("eslam_mc_count", ptr(sdf), nx, ny, nz, float(level), ptr(n_tri), ptr(tri), ptr(counts), stream())
    incl = torch.cumsum(counts, 0, dtype=torch.int64)
    total = int(incl[-1].item()) if nb else 0  # the one host sync: the output has to be allocated
    if total == 0:
        return torch.zeros(0, 3, device=dev), torch.zeros(0, 3, dtype=torch.int64, device=dev)
    base = (incl - counts).contiguous()
    verts = torch.empty(3 * total, 3, dtype=torch.float32, device=dev)
    keys = torch.empty(3 * total, dtype=torch.int64, device=dev)
    call("eslam_mc_emit", ptr(sdf), ptr(xs), ptr(ys), ptr(zs), nx, ny, nz, float(level), ptr(n_tri), ptr(tri), ptr(base),
         ptr(verts), ptr(keys), stream())
    if not weld:
        return verts, torch.arange(3 * total, device=dev).view(-1, 3)
    uniq, inv = torch.unique(keys, return_inverse=True)
    welded = torch.empty(uniq.numel(), 3, dtype=torch.float32, device=dev)
    welded[inv] = verts  # all copies of a vertex carry identical bits
    return welded, inv.view(-1, 3)


def backproject_depth(depth, c2w, cam, stride=1):
    """World points of the depth > 0 pixels of one frame (every `stride`-th pixel), rays as common.py:87-99."""
    H, W, fx, fy, cx, cy = cam
    dev = depth.device
    jj, ii = torch.meshgrid(torch.arange(0, H, stride, device=dev), torch.arange(0, W, stride, device=dev), indexing="ij")
    d = depth[::stride, ::stride]
    ok = d > 0
    ii, jj, d = ii[ok].float(), jj[ok].float(), d[ok].float()
    cam_pts = torch.stack([(ii - cx) / fx * d, -(jj - cy) / fy * d, -d], -1)
    c2w = c2w.to(dev).float()
    return cam_pts @ c2w[:3, :3].T + c2w[:3, 3]


def bound_from_frames(keyframe_dict, cam, mesh_bound_scale=1.02, max_points=400000):
    """The convex mesh bound of Mesher.get_bound_from_frames (Mesher.py:63-128) as (hull vertices, hull faces, [F,4]
    half-spaces for query_grid_sdf).  The reference fuses the keyframes into an Open3D TSDF volume, extracts its mesh
    and takes the convex hull of the mesh vertices plus the camera centres, scaled about its centre.  Open3D is not
    available here: the surface points are the back-projected keyframe depths themselves (the TSDF mesh lies within a
    voxel, 4/512 m, of them), the hull is Qhull's (scipy.spatial.ConvexHull) on the host -- a few hundred thousand
    points, once per mesh -- and the point-in-hull test runs on the device as half-space tests (eslam_grid_sdf*)."""
    from scipy.spatial import ConvexHull

    n_kf = max(len(keyframe_dict), 1)
    H, W = cam[0], cam[1]
    stride = max(1, int(np.ceil(np.sqrt(n_kf * H * W / max_points))))
    pts = []
    for kf in keyframe_dict:
        pts.append(backproject_depth(kf["depth"], kf["est_c2w"], cam, stride))
        pts.append(kf["est_c2w"].to(pts[-1].device).float()[:3, 3][None])
    pts = torch.cat(pts, 0).double().cpu().numpy()
    hull = ConvexHull(pts)
    used = np.unique(hull.simplices)
    remap = -np.ones(pts.shape[0], dtype=np.int64)
    remap[used] = np.arange(used.size)
    v, f = pts[used], remap[hull.simplices]
    centre = v.mean(0)  # open3d get_center(): mean of the hull's vertices
    v = centre + mesh_bound_scale * (v - centre)
    return v, f, hull_planes(v, f)


def cull_mesh(vertices, faces, depths, c2ws, cam, truncation, eval_rec=True):
    """cull_mesh (src/tools/cull_mesh.py:36-114) on the device: drop the faces whose three vertices are ALL unseen in
    every frame, then the vertices nothing references.  vertices [V,3] / faces [F,3] device tensors; depths: iterable of
    [H,W] float32 frames, c2ws: matching [4,4] poses.  Returns (vertices', faces', kept-vertex index)."""
    from .hotpath import make_camera
    import ctypes as C

    dev = vertices.device
    verts = vertices.float().contiguous()
    seen = torch.zeros(verts.shape[0], dtype=torch.uint8, device=dev)
    camera = make_camera(*cam)
    for depth, c2w in zip(depths, c2ws):
        w2c = torch.inverse(c2w.to(dev).float()).contiguous()
        d = depth.to(dev).float().contiguous()
        call("eslam_cull_frame", ptr(verts), verts.shape[0], ptr(w2c), ptr(d), C.byref(camera), float(truncation),
             1 if eval_rec else 0, ptr(seen), stream())
    unseen = seen == 0
    drop = unseen[faces].all(dim=1)          # cull_mesh.py:102: faces whose vertices were never seen
    faces = faces[~drop]
    used = torch.zeros(verts.shape[0], dtype=torch.bool, device=dev)
    used[faces.reshape(-1)] = True
    idx = torch.nonzero(used).squeeze(-1)    # remove_unreferenced_vertices
    remap = torch.full((verts.shape[0],), -1, dtype=torch.int64, device=dev)
    remap[idx] = torch.arange(idx.numel(), device=dev)
    return verts[idx], remap[faces], idx


def write_ply(path, vertices, faces, colors=None):
    """Binary little-endian PLY (what the reference exports through trimesh, Mesher.py:262-263)."""
    v = np.asarray(vertices.detach().cpu() if torch.is_tensor(vertices) else vertices, dtype=np.float32)
    f = np.asarray(faces.detach().cpu() if torch.is_tensor(faces) else faces, dtype=np.int32)
    has_c = colors is not None
    if has_c:
        c = np.asarray(colors.detach().cpu() if torch.is_tensor(colors) else colors, dtype=np.float64)
        c = np.clip(np.round(c * 255.0), 0, 255).astype(np.uint8)
    with open(path, "wb") as fh:
        hdr = ["ply", "format binary_little_endian 1.0", f"element vertex {v.shape[0]}", "property float x",
               "property float y", "property float z"]
        if has_c:
            hdr += ["property uchar red", "property uchar green", "property uchar blue"]
        hdr += [f"element face {f.shape[0]}", "property list uchar int vertex_indices", "end_header"]
        fh.write(("\n".join(hdr) + "\n").encode())
        if has_c:
            rec = np.empty(v.shape[0], dtype=[("p", "<f4", 3), ("c", "u1", 3)])
            rec["p"], rec["c"] = v, c
        else:
            rec = np.empty(v.shape[0], dtype=[("p", "<f4", 3)])
            rec["p"] = v
        fh.write(rec.tobytes())
        frec = np.empty(f.shape[0], dtype=[("n", "u1"), ("i", "<i4", 3)])
        frec["n"], frec["i"] = 3, f
        fh.write(frec.tobytes())


def get_mesh(all_planes, decoders, keyframe_dict, cam, marching_cubes_bound, resolution, level_set=0.0,
             mesh_bound_scale=1.02, scale=1.0, bound=None, color=True, mesh_out_file=None):
    """Mesher.get_mesh (Mesher.py:188-264) without leaving the device: mesh bound from the keyframes -> SDF on the
    lattice with the bound applied in the same pass -> marching cubes -> vertex colours through the decoders.
    Returns dict(vertices [V,3] (divided by `scale` like Mesher.py:260), faces [F,3], vertex_colors [V,3] or None)."""
    axes = grid_axes(marching_cubes_bound, resolution)
    hull = None
    if keyframe_dict:
        _, _, hull = bound_from_frames(keyframe_dict, cam, mesh_bound_scale)
    sdf = query_grid_sdf(all_planes, decoders, axes, bound, hull=hull)
    verts, faces = marching_cubes(sdf, axes, level_set)
    colors = None
    if color and verts.shape[0]:
        colors = eval_points(verts, all_planes, decoders, bound)[:, :3]
    verts = verts / scale
    if mesh_out_file is not None:
        write_ply(mesh_out_file, verts, faces, colors)
    return {"vertices": verts, "faces": faces, "vertex_colors": colors}
