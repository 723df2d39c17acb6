"""GPU: the device half of Mesher.get_mesh around the query (Mesher.py:188-264, cull_mesh.py:36-114): marching cubes over
the SDF lattice against the oracle's vertex set and per-cell polygons, welded mesh closed and oriented, vertex colours,
frustum culling against the reference's torch arithmetic, the convex mesh bound."""
from collections import Counter

import numpy as np
import pytest
import torch

import eslam_oracle as O
from conftest import golden_field, load_npz, rel_err, to_device_scene

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _volume(seed, shape):
    """A smooth random field with many ambiguous cells, as lattice values in the reference's flat order."""
    rng = np.random.default_rng(seed)
    nx, ny, nz = shape
    vol = rng.standard_normal(shape)
    for _ in range(2):  # a little smoothing keeps the surface from being pure noise, ambiguous faces remain
        vol = (vol + np.roll(vol, 1, 0) + np.roll(vol, 1, 1) + np.roll(vol, 1, 2)) / 4.0
    xs, ys, zs = np.linspace(-1.0, 1.3, nx), np.linspace(0.2, 2.0, ny), np.linspace(-0.5, 0.4, nz)
    flat = torch.from_numpy(np.ascontiguousarray(vol.transpose(1, 0, 2)).reshape(-1)).float().to(DEV)  # (iy*nx+ix)*nz+iz
    return vol.astype(np.float32).astype(np.float64), (xs, ys, zs), flat


@pytest.mark.parametrize("shape,level", [((17, 13, 19), 0.05), ((40, 33, 9), -0.1), ((3, 2, 2), 0.0)])
def test_marching_cubes_matches_the_oracle(shape, level):
    from myslam_b200.mesher import marching_cubes

    vol, axes, flat = _volume(sum(shape), shape)
    verts, faces = marching_cubes(flat, axes, level, weld=False)
    verts_w, faces_w = marching_cubes(flat, axes, level, weld=True)
    pos = O.marching_cubes_vertices(vol, level, *[a.astype(np.float32).astype(np.float64) for a in axes])
    loops = O.marching_cubes_loops(vol, level)
    n_tri = sum(len(l) - 2 for cell in loops.values() for l in cell)
    assert verts.shape[0] == 3 * n_tri and faces.shape[0] == n_tri and faces_w.shape[0] == n_tri
    # welded vertices == the oracle's level crossings (one per straddling lattice edge), positions to float32 rounding
    assert verts_w.shape[0] == len(pos)
    want = np.array(sorted(tuple(np.round(p, 5)) for p in pos.values()))
    got = np.array(sorted(tuple(np.round(p, 5)) for p in verts_w.double().cpu().numpy()))
    assert np.abs(want - got).max() < 2e-5
    # every interior edge of the welded mesh is shared by exactly two triangles with opposite directions
    f = faces_w.cpu().numpy()
    edges = Counter()
    for tri in f:
        for a, b in ((0, 1), (1, 2), (2, 0)):
            edges[(int(tri[a]), int(tri[b]))] += 1
    v = verts_w.double().cpu().numpy()
    lo = np.array([a[0] for a in axes])
    hi = np.array([a[-1] for a in axes])

    def on_border(i):
        return bool(((np.abs(v[i] - lo) < 1e-6) | (np.abs(v[i] - hi) < 1e-6)).any())

    for (a, b), c in edges.items():
        assert c == 1
        if not (on_border(a) and on_border(b)):
            assert edges.get((b, a), 0) == 1


def test_marching_cubes_sphere_is_closed_and_outward():
    from myslam_b200.mesher import marching_cubes

    n = 48
    ax = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(ax, ax, ax, indexing="ij")
    vol = np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.62
    flat = torch.from_numpy(np.ascontiguousarray(vol.transpose(1, 0, 2)).reshape(-1)).float().to(DEV)
    v, f = marching_cubes(flat, (ax, ax, ax), 0.0)
    p = v[f]  # [F,3,3]
    nrm = torch.cross(p[:, 1] - p[:, 0], p[:, 2] - p[:, 0], dim=-1)
    assert bool(((nrm * p.mean(1)).sum(-1) > 0).all()), "normals must point towards increasing values"
    area = 0.5 * nrm.norm(dim=-1).sum().item()
    assert abs(area - 4 * np.pi * 0.62 ** 2) / (4 * np.pi * 0.62 ** 2) < 0.01
    vol6 = (p[:, 0] * torch.cross(p[:, 1], p[:, 2], dim=-1)).sum().item() / 6.0  # closed + oriented => enclosed volume
    assert abs(vol6 - 4 / 3 * np.pi * 0.62 ** 3) / (4 / 3 * np.pi * 0.62 ** 3) < 0.01
    assert (v.norm(dim=-1) - 0.62).abs().max().item() < 1e-3
    # Euler characteristic of a sphere
    e = torch.cat([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]]).sort(dim=1).values
    n_e = torch.unique(e, dim=0).shape[0]
    assert v.shape[0] - n_e + f.shape[0] == 2


def test_get_mesh_on_the_golden_scene_and_vertex_colours():
    """Mesher.get_mesh end to end on the fixture map: lattice query (golden mesh.npz pins the values) -> marching cubes ->
    colours; the mesh is what the oracle's vertex set / polygons give on the SAME lattice values, colours are
    eval_points at the vertices."""
    from myslam_b200.mesher import eval_points, get_mesh, grid_axes, marching_cubes, query_grid_sdf

    fld, d = golden_field(), load_npz("mesh.npz")
    planes, dec = to_device_scene(fld)
    res = float(d["resolution"])
    axes = grid_axes(d["mc_bound"], res)
    sdf = query_grid_sdf(planes, dec, axes, fld.bound)
    # the fixture map is untrained (its sdf does not cross 0): extract the level set at the median in-bound value
    level = float(sdf[sdf > -1].median())
    out = get_mesh(planes, dec, [], None, d["mc_bound"], res, level_set=level, bound=fld.bound)
    nx, ny, nz = (len(a) for a in axes)
    vol = sdf.cpu().numpy().reshape(ny, nx, nz).transpose(1, 0, 2).astype(np.float64)
    pos = O.marching_cubes_vertices(vol, level, *[a.astype(np.float32).astype(np.float64) for a in axes])
    assert out["vertices"].shape[0] == len(pos) > 100
    loops = O.marching_cubes_loops(vol, level)
    assert out["faces"].shape[0] == sum(len(l) - 2 for cell in loops.values() for l in cell)
    raw = O.query_points(fld, out["vertices"].cpu())
    assert rel_err(out["vertex_colors"], raw[:, :3]) < 1e-4
    assert torch.equal(out["vertex_colors"], eval_points(out["vertices"], planes, dec, fld.bound)[:, :3])


def test_cull_mesh_matches_the_references_arithmetic(tmp_path):
    from myslam_b200.mesher import cull_mesh, write_ply
    from myslam_b200 import synthetic as S

    g = torch.Generator().manual_seed(4)
    H, W, fx, fy, cx, cy = 60, 80, 70.0, 70.0, 39.5, 29.5
    room = [[-1.0, 1.2], [-0.8, 0.9], [-0.7, 0.8]]
    poses = S.trajectory(5, room, step_deg=25.0)
    frames = [S.render_box_room(p, H, W, fx, fy, cx, cy, room, "cpu", hole_frac=0.05, generator=g) for p in poses]
    V = 5000
    verts = torch.stack([torch.empty(V).uniform_(lo - 0.3, hi + 0.3, generator=g) for lo, hi in room], -1)
    faces = torch.randint(0, V, (9000, 3), generator=g)
    for eval_rec in (True, False):
        seen = torch.zeros(V, dtype=torch.bool)
        for (col, dep), c2w in zip(frames, poses):
            seen |= O.cull_mask(verts.clone(), dep, c2w, (fx, fy, cx, cy), H, W, 0.06, eval_rec)
        v2, f2, idx = cull_mesh(verts.to(DEV), faces.to(DEV), [f[1] for f in frames], list(poses), (H, W, fx, fy, cx, cy),
                                0.06, eval_rec)
        keep_face = ~((~seen)[faces].all(dim=1))
        # a vertex exactly at a visibility threshold may fall on either side in float32: allow a handful
        ref_faces = faces[keep_face]
        assert abs(f2.shape[0] - ref_faces.shape[0]) <= 3
        if f2.shape[0] == ref_faces.shape[0]:
            assert torch.equal(idx.cpu()[f2.cpu()], ref_faces)
        assert 0 < f2.shape[0] < faces.shape[0]
    p = tmp_path / "m.ply"
    write_ply(str(p), v2, f2, torch.rand(v2.shape[0], 3))
    head = open(p, "rb").read(200).decode("ascii", "ignore")
    assert head.startswith("ply") and f"element vertex {v2.shape[0]}" in head and f"element face {f2.shape[0]}" in head


def test_bound_from_frames_contains_what_the_keyframes_saw():
    from myslam_b200.mesher import backproject_depth, bound_from_frames
    from myslam_b200 import synthetic as S

    g = torch.Generator().manual_seed(2)
    cam = (60, 80, 70.0, 70.0, 39.5, 29.5)
    room = [[-1.0, 1.2], [-0.8, 0.9], [-0.7, 0.8]]
    poses = S.trajectory(4, room, step_deg=30.0)
    kfs = []
    for p in poses:
        col, dep = S.render_box_room(p, *cam, room, DEV, hole_frac=0.0, generator=g)
        kfs.append({"depth": dep, "color": col, "est_c2w": p.to(DEV)})
    v, f, planes = bound_from_frames(kfs, cam, mesh_bound_scale=1.02)
    assert planes.shape[1] == 4 and planes.shape[0] == f.shape[0] >= 4
    pts = torch.cat([backproject_depth(k["depth"], k["est_c2w"], cam) for k in kfs]).cpu()
    inside = (pts @ planes[:, :3].T + planes[:, 3] <= 1e-5).all(dim=1)
    assert bool(inside.all()), "every back-projected keyframe pixel lies inside the (slightly enlarged) hull"
    far = pts.mean(0) + 10.0
    assert not bool((far @ planes[:, :3].T + planes[:, 3] <= 0).all())
