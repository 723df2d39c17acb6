"""Marching-cubes case tables, GENERATED (no table is copied from anywhere): for each of the 256 corner configurations
the iso-contour is traced face by face and chained into closed loops, which are fanned into triangles.

Conventions (shared by csrc/mcubes.cuh and the oracle's restatement):
  corner c in 0..7     offset (c & 1, (c >> 1) & 1, (c >> 2) & 1) along (x, y, z)
  edge e in 0..11      e = 4 * axis + (b0 + 2 * b1), (b0, b1) = the corner bits of the two OTHER axes in increasing
                       axis order; it joins the corners whose `axis` bit is 0 and 1
  configuration        bit c set  <=>  value[c] < level  ("inside": the solid side of a signed distance)
  triangles            wound so that the normal points towards the OUTSIDE (increasing value, free space)
  ambiguous faces      (two diagonal inside corners) are cut so that each inside corner is separated: the rule only
                       looks at the face's own four corners, so two cells sharing a face agree and the surface has no
                       cracks

The reference calls skimage.measure.marching_cubes (Mesher.py:219-243; third party, Lewiner's variant, not present in
this image).  Every vertex of either triangulation lies on a lattice edge whose end values straddle the level, at the
linear interpolation point, so the VERTEX SETS are the same; the triangulations differ where a cell's polygon can be
fanned or an ambiguous case resolved in more than one way.
"""
from __future__ import annotations

import numpy as np

CORNER_OFFSETS = np.array([[c & 1, (c >> 1) & 1, (c >> 2) & 1] for c in range(8)], dtype=np.int32)


def edge_id(axis: int, c: int) -> int:
    """Edge along `axis` through corner c (either end)."""
    others = [a for a in range(3) if a != axis]
    b0, b1 = (c >> others[0]) & 1, (c >> others[1]) & 1
    return 4 * axis + b0 + 2 * b1


def _edge_table():
    ends = np.zeros((12, 2), dtype=np.int32)
    for axis in range(3):
        for c in range(8):
            if (c >> axis) & 1:
                continue
            ends[edge_id(axis, c)] = (c, c | (1 << axis))
    return ends


EDGE_CORNERS = _edge_table()           # [12][2] corner ids (axis bit 0, axis bit 1)
EDGE_AXIS = np.arange(12, dtype=np.int32) // 4


def _faces():
    """Six faces as 4 corner ids in counter-clockwise order seen from OUTSIDE the cube."""
    faces = []
    for a in range(3):
        u, v = (a + 1) % 3, (a + 2) % 3  # e_u x e_v = e_a
        for side in (0, 1):
            quad = [(0, 0), (1, 0), (1, 1), (0, 1)]  # CCW seen from +a
            if side == 0:
                quad = quad[::-1]                     # outward normal is -a
            faces.append([(side << a) | (bu << u) | (bv << v) for bu, bv in quad])
    return faces


def _edge_between(c0: int, c1: int) -> int:
    d = c0 ^ c1
    axis = {1: 0, 2: 1, 4: 2}[d]
    return edge_id(axis, c0)


def _edge_faces(e: int):
    """The two cube faces (axis, side) an edge lies in."""
    axis = e // 4
    others = [a for a in range(3) if a != axis]
    b = (e % 4) & 1, (e % 4) >> 1
    return {(others[0], b[0]), (others[1], b[1])}


_MID = None


def _triangulate(loop):
    """Triangles (same orientation as the loop) of the closed polygon `loop` (crossing edges of one cell).  Among all
    triangulations: the fewest triangles / diagonals lying INSIDE a cube face (a loop can cross an ambiguous face twice;
    a triangle spanned by three vertices of one face would be a zero-volume membrane the neighbouring cell duplicates,
    a diagonal in a face a line both cells' sheets touch), then the smallest area with the vertices at the edge
    midpoints."""
    global _MID
    if _MID is None:
        _MID = (CORNER_OFFSETS[EDGE_CORNERS[:, 0]] + CORNER_OFFSETS[EDGE_CORNERS[:, 1]]) / 2.0
    n = len(loop)

    def all_tri(i, j):  # triangulations of the sub-polygon i..j (indices into loop), as lists of index triples
        if j - i < 2:
            return [[]]
        out = []
        for k in range(i + 1, j):
            for left in all_tri(i, k):
                for right in all_tri(k, j):
                    out.append(left + [(i, k, j)] + right)
        return out

    best, best_key = None, None
    for cand in all_tri(0, n - 1):
        in_face, area = 0, 0.0
        for (i, k, j) in cand:
            a, b, c = loop[i], loop[k], loop[j]
            if _edge_faces(a) & _edge_faces(b) & _edge_faces(c):
                in_face += 4  # a whole triangle inside a face
            for (p, q) in ((i, k), (k, j), (i, j)):
                # a DIAGONAL of the polygon (not one of its sides) inside a cube face: the neighbouring cell may draw
                # the same one, pinching the surface along that line
                if (q - p) % n not in (1, n - 1) and _edge_faces(loop[p]) & _edge_faces(loop[q]):
                    in_face += 1
            area += 0.5 * float(np.linalg.norm(np.cross(_MID[b] - _MID[a], _MID[c] - _MID[a])))
        key = (in_face, round(area, 9), cand)
        if best_key is None or key < best_key:
            best, best_key = cand, key
    return [(loop[i], loop[k], loop[j]) for (i, k, j) in best]


def _build():
    faces = _faces()
    tri_lists = []
    for cfg in range(256):
        inside = [(cfg >> c) & 1 for c in range(8)]
        nxt = {}  # crossing edge -> next crossing edge along the oriented contour (inside on the left, seen from outside)
        for quad in faces:
            exits, entries = [], []  # positions i of the boundary step quad[i] -> quad[i+1]
            for i in range(4):
                a, b = quad[i], quad[(i + 1) % 4]
                if inside[a] and not inside[b]:
                    exits.append(i)
                elif not inside[a] and inside[b]:
                    entries.append(i)
            for x in exits:
                # the entry that closes THIS inside arc: walk backwards from the exit to where the arc was entered
                i = x
                while True:
                    i = (i - 1) % 4
                    if i in entries:
                        break
                e_from = _edge_between(quad[x], quad[(x + 1) % 4])
                e_to = _edge_between(quad[i], quad[(i + 1) % 4])
                assert e_from not in nxt
                nxt[e_from] = e_to
        tris, seen = [], set()
        for start in sorted(nxt):
            if start in seen:
                continue
            loop, e = [], start
            while e not in seen:
                seen.add(e)
                loop.append(e)
                e = nxt[e]
            assert e == start and len(loop) >= 3
            tris.extend(_triangulate(loop))
        tri_lists.append(tris)
    # orientation: one inside corner at the origin must give a normal pointing away from it
    mid = (CORNER_OFFSETS[EDGE_CORNERS[:, 0]] + CORNER_OFFSETS[EDGE_CORNERS[:, 1]]) / 2.0
    t = tri_lists[1][0]
    n = np.cross(mid[t[1]] - mid[t[0]], mid[t[2]] - mid[t[0]])
    flip = float(n @ (mid[list(t)].mean(0) - CORNER_OFFSETS[0])) < 0
    max_t = max(len(t) for t in tri_lists)
    n_tri = np.zeros(256, dtype=np.uint8)
    table = -np.ones((256, 3 * max_t), dtype=np.int8)
    for cfg, tris in enumerate(tri_lists):
        n_tri[cfg] = len(tris)
        for k, tr in enumerate(tris):
            table[cfg, 3 * k: 3 * k + 3] = tr[::-1] if flip else tr
    return n_tri, table


N_TRI, TRI_TABLE = _build()   # [256] uint8, [256][3 * MAX_TRI] int8 edge ids (-1 padded)
MAX_TRI = TRI_TABLE.shape[1] // 3
