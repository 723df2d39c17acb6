"""CPU: the generated marching-cubes case tables (myslam_b200/mc_tables.py) against the oracle's per-cell restatement
(contours traced on the faces of every cell of a random volume): same polygons, same orientation, manifold surface."""
from collections import Counter

import numpy as np

import eslam_oracle as O
from myslam_b200 import mc_tables as T


def table_triangles(vol, level):
    """{cell: [triangle as 3 lattice edges]} from the tables."""
    n = vol.shape
    cfg = np.zeros((n[0] - 1, n[1] - 1, n[2] - 1), dtype=np.int32)
    for k, (ox, oy, oz) in enumerate(T.CORNER_OFFSETS):
        cfg |= ((vol[ox:n[0] - 1 + ox, oy:n[1] - 1 + oy, oz:n[2] - 1 + oz] < level).astype(np.int32) << k)
    out = {}
    for cell in np.argwhere(T.N_TRI[cfg] > 0):
        c = cfg[tuple(cell)]
        tris = []
        for t in range(T.N_TRI[c]):
            tri = []
            for e in T.TRI_TABLE[c, 3 * t:3 * t + 3]:
                p0 = tuple(int(x) for x in (cell + T.CORNER_OFFSETS[T.EDGE_CORNERS[e][0]]))
                tri.append((p0, int(T.EDGE_AXIS[e])))
            tris.append(tri)
        out[tuple(int(x) for x in cell)] = tris
    return out


def test_tables_shape():
    assert T.MAX_TRI == 5 and T.TRI_TABLE.shape == (256, 15) and T.N_TRI[0] == 0 and T.N_TRI[255] == 0
    for cfg in range(256):
        used = T.TRI_TABLE[cfg][:3 * T.N_TRI[cfg]]
        assert (used >= 0).all() and (used < 12).all() and (T.TRI_TABLE[cfg][3 * T.N_TRI[cfg]:] == -1).all()


def test_tables_reproduce_the_oracles_polygons_on_a_random_volume():
    rng = np.random.default_rng(1)
    vol = rng.standard_normal((9, 10, 11))
    tris = table_triangles(vol, 0.1)
    loops = O.marching_cubes_loops(vol, 0.1)
    assert set(tris) == set(loops)
    for cell, tl in tris.items():
        edges = Counter()
        for tri in tl:
            for a, b in ((0, 1), (1, 2), (2, 0)):
                edges[(tri[a], tri[b])] += 1
        boundary = {e for e, c in edges.items() if edges.get((e[1], e[0]), 0) == 0}
        assert all(c == 1 for c in edges.values())
        want = set()
        for loop in loops[cell]:
            for i in range(len(loop)):
                want.add((loop[i], loop[(i + 1) % len(loop)]))
        # the triangulated polygons have exactly the oracle's contour segments as their (directed) boundary
        assert boundary == want or boundary == {(b, a) for a, b in want}, cell
        assert len(tl) == sum(len(l) - 2 for l in loops[cell])


def test_surface_is_closed_oriented_and_on_the_level_set():
    n = 20
    xs = ys = zs = np.linspace(-1, 1, n)
    X, Y, Z = np.meshgrid(xs, ys, zs, indexing="ij")
    vol = np.sqrt(X ** 2 + Y ** 2 + Z ** 2) - 0.7  # negative inside the ball
    tris = table_triangles(vol, 0.0)
    pos = O.marching_cubes_vertices(vol, 0.0, xs, ys, zs)
    edges = Counter()
    out_ok = 0
    n_t = 0
    for tl in tris.values():
        for tri in tl:
            p = np.array([pos[v] for v in tri])
            nrm = np.cross(p[1] - p[0], p[2] - p[0])
            out_ok += float(nrm @ p.mean(0)) > 0  # normal towards increasing values = away from the centre
            n_t += 1
            for a, b in ((0, 1), (1, 2), (2, 0)):
                edges[(tri[a], tri[b])] += 1
    assert out_ok == n_t
    assert all(c == 1 and edges.get((b, a), 0) == 1 for (a, b), c in edges.items())  # closed 2-manifold
    r = np.array([np.linalg.norm(p) for p in pos.values()])
    assert abs(r - 0.7).max() < 2e-3
    used = {v for tl in tris.values() for tri in tl for v in tri}
    assert used == set(pos)  # the triangles use exactly the level crossings
