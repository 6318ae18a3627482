"""TEST INFRASTRUCTURE -- CPU restatement (numpy, float32) of iso-surface extraction from a TSDF lattice by
marching cubes, the step the reference hands its fused volume to:
  bounded path    /root/reference/gssr/utils/mesh_utils.py:178   volume.extract_triangle_mesh()   (Open3D 0.18.0,
                  ScalableTSDFVolume; requirements.txt:6, absent here)
  unbounded path  /root/reference/gssr/utils/mcube_utils.py:71-80  skimage.measure.marching_cubes(level=0)  (absent here)
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this; the product path
(gs-sr_b200/csrc/mcubes.cu) never does.

PARITY UNPINNED against Open3D / skimage: both are third-party dependencies that are not vendored, not installed
and have no vectors in the reference.  What is restated is the published algorithm (Lorensen & Cline 1987) with the
conventions Open3D's extractor uses -- a corner is inside when f < level, a cell produces triangles only when all
eight corners are observed, a vertex on a lattice edge is placed at f0 / (f0 - f1) of the edge and shared by the
cells around that edge, colours are interpolated with the same weight -- and the face-ambiguity rule below, which
makes the case table consistent across shared faces (no holes).  Pinned properties (tests/test_mcubes_*.py): every
interior mesh edge is used by exactly two triangles in opposite directions, Euler characteristic of closed surfaces,
vertices lie on the trilinear zero set along their edge, outward orientation (towards f > level), and the CUDA
extractor reproduces vertices, colours and faces of this restatement bit for bit.

Conventions: lattice arrays are (nz, ny, nx) with x fastest (gsr_b200.tsdf.BoundedTSDFVolume); corner k of a cell is
offset (k & 1, k >> 1 & 1, k >> 2 & 1) in (x, y, z); lattice edge 4 a + j runs along axis a from the corner whose other
two axes (in increasing order) have bits (j & 1, j >> 1).  Vertices are ordered by (owner voxel linear index, axis),
faces by (cell linear index, table order).
"""
from __future__ import annotations

import numpy as np

F = np.float32
_UV = ((1, 2), (2, 0), (0, 1))           # (u, v) with e_u x e_v = e_a


def _edge_id(c0, c1):
    d = c0 ^ c1
    a = d.bit_length() - 1
    base = c0 & c1
    others = [ax for ax in range(3) if ax != a]
    return 4 * a + ((base >> others[0]) & 1) + 2 * ((base >> others[1]) & 1)


def edge_corners(e):
    """The two corners (lower, upper along the edge axis) of lattice edge e of a cell."""
    a, j = divmod(e, 4)
    others = [ax for ax in range(3) if ax != a]
    base = ((j & 1) << others[0]) | ((j >> 1) << others[1])
    return base, base | (1 << a)


def case_triangles(case):
    """Triangles (as triples of cell edge ids) of one sign configuration; bit k of `case` = corner k inside.

    On every cell face the crossing points are joined by directed segments that keep the inside corners on their
    left (seen from outside the cell); a face with four crossings cuts each inside corner off separately -- a rule
    that depends on the four corner signs only, so the two cells sharing a face agree.  The segments close into loops,
    each loop is triangulated (see _triangulate) and wound so that normals point to the outside corners."""
    inside = [(case >> k) & 1 for k in range(8)]
    nxt = {}
    for a in range(3):
        u, v = _UV[a]
        for s in (0, 1):
            cyc = [(0, 0), (1, 0), (1, 1), (0, 1)]
            if s == 0:
                cyc.reverse()
            corners = [(s << a) | (bu << u) | (bv << v) for bu, bv in cyc]
            kind = []
            for i in range(4):
                c0, c1 = corners[i], corners[(i + 1) % 4]
                kind.append("X" if inside[c0] and not inside[c1] else "Y" if inside[c1] and not inside[c0] else "-")
            for i in range(4):
                if kind[i] != "X":
                    continue
                j = (i - 1) % 4
                while kind[j] != "Y":
                    j = (j - 1) % 4
                src = _edge_id(corners[i], corners[(i + 1) % 4])
                dst = _edge_id(corners[j], corners[(j + 1) % 4])
                assert src not in nxt
                nxt[src] = dst
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
        loop.reverse()                                       # normals towards the outside corners
        tris.extend(_triangulate(loop))
    return tris


def _share_face(e0, e1):
    """Do two cell edges lie on a common cell face?"""
    c = [edge_corners(e0), edge_corners(e1)]
    for a in range(3):
        for s in (0, 1):
            if all(((k >> a) & 1) == s for pair in c for k in pair):
                return True
    return False


def _all_triangulations(poly):
    n = len(poly)
    if n < 3:
        return [[]]
    out = []
    for k in range(1, n - 1):
        for left in _all_triangulations(poly[:k + 1]):
            for right in _all_triangulations(poly[k:]):
                out.append(left + [(poly[0], poly[k], poly[n - 1])] + right)
    return out


def _triangulate(loop):
    """Triangulation of one loop without a diagonal between two crossings of the same cell face: such a diagonal lies
    in a face with four crossings, where the neighbouring cell could pick the same one and the two sheets would touch
    along it (an edge with four triangles).  First admissible triangulation in enumeration order."""
    n = len(loop)
    for tri in _all_triangulations(loop):
        ok = True
        for t in tri:
            for k in range(3):
                i, j = loop.index(t[k]), loop.index(t[(k + 1) % 3])
                if (i - j) % n not in (1, n - 1) and _share_face(t[k], t[(k + 1) % 3]):
                    ok = False
        if ok:
            # rotate every triangle so that it starts at its smallest edge id (a canonical, reproducible table)
            res = []
            for t in tri:
                k = t.index(min(t))
                res.append((t[k], t[(k + 1) % 3], t[(k + 2) % 3]))
            return res
    raise AssertionError("no admissible triangulation")


def build_table():
    """(256, 3 * MAX_TRIS) int8 table padded with -1, and the (256,) triangle counts."""
    rows = [case_triangles(c) for c in range(256)]
    mt = max(len(r) for r in rows)
    tab = -np.ones((256, 3 * mt), dtype=np.int8)
    for c, r in enumerate(rows):
        flat = [e for t in r for e in t]
        tab[c, :len(flat)] = flat
    return tab, np.array([len(r) for r in rows], dtype=np.int32)


TABLE, NTRI = build_table()
MAX_TRIS = TABLE.shape[1] // 3


def cell_cases(tsdf, weight=None, min_weight=None, level=0.0):
    """(nz, ny, nx) uint8: sign configuration of the cell whose corner 0 is that voxel; 0 where the cell leaves the
    lattice or (with min_weight) one of its corners has weight <= min_weight."""
    f = np.asarray(tsdf, dtype=F)
    nz, ny, nx = f.shape
    inside = f < F(level)
    ok = np.ones_like(inside) if min_weight is None else (np.asarray(weight, dtype=F) > F(min_weight))
    case = np.zeros((nz, ny, nx), dtype=np.uint8)
    valid = np.zeros((nz, ny, nx), dtype=bool)
    valid[:nz - 1, :ny - 1, :nx - 1] = True
    for k in range(8):
        dx, dy, dz = k & 1, (k >> 1) & 1, (k >> 2) & 1
        sl = (slice(dz, nz - 1 + dz), slice(dy, ny - 1 + dy), slice(dx, nx - 1 + dx))
        case[:nz - 1, :ny - 1, :nx - 1] |= (inside[sl].astype(np.uint8) << k)
        valid[:nz - 1, :ny - 1, :nx - 1] &= ok[sl]
    case[~valid] = 0
    return case


def extract(tsdf, weight=None, min_weight=None, level=0.0, origin=(0.0, 0.0, 0.0), voxel_size=1.0, rgb=None):
    """Marching cubes of the lattice.  Returns verts (V,3) f32, faces (F,3) i32, colors (V,3) f32 or None."""
    f = np.asarray(tsdf, dtype=F)
    nz, ny, nx = f.shape
    case = cell_cases(f, weight, min_weight, level)
    nvox = nz * ny * nx
    # an owned edge (voxel, axis) carries a vertex when one of the (up to four) cells around it is active and sees
    # different signs at its two ends
    pad = np.zeros((nz + 1, ny + 1, nx + 1), dtype=np.uint8)
    pad[1:, 1:, 1:] = case
    emask = np.zeros((nz, ny, nx), dtype=np.uint8)
    for a in range(3):
        u, v = [ax for ax in range(3) if ax != a]
        used = np.zeros((nz, ny, nx), dtype=bool)
        for bu in (0, 1):
            for bv in (0, 1):
                off = [0, 0, 0]
                off[u], off[v] = -bu, -bv                    # the cell in which this edge has (u, v) bits (bu, bv)
                c = pad[1 + off[2]:nz + 1 + off[2], 1 + off[1]:ny + 1 + off[1], 1 + off[0]:nx + 1 + off[0]]
                k0 = (bu << u) | (bv << v)
                k1 = k0 | (1 << a)
                used |= ((c >> k0) & 1) != ((c >> k1) & 1)
        emask |= used.astype(np.uint8) << a
    cnt = ((emask & 1) + ((emask >> 1) & 1) + ((emask >> 2) & 1)).astype(np.int64).ravel()
    vbase = np.concatenate([[0], np.cumsum(cnt)[:-1]])
    V = int(cnt.sum())
    verts = np.zeros((V, 3), dtype=F)
    colors = None if rgb is None else np.zeros((V, 3), dtype=F)
    lin = np.arange(nvox)
    ix, iy, iz = lin % nx, (lin // nx) % ny, lin // (nx * ny)
    idx3 = (ix, iy, iz)
    em = emask.ravel()
    ff = (f - F(level)).astype(F).ravel() if level != 0.0 else f.ravel()
    org = np.asarray(origin, dtype=F)
    vs = F(voxel_size)
    stride = (1, nx, nx * ny)
    col = None if rgb is None else np.asarray(rgb, dtype=F).reshape(nvox, 3)
    for a in range(3):
        sel = np.nonzero((em >> a) & 1)[0]
        rank = np.zeros(sel.shape, dtype=np.int64)
        for b in range(a):
            rank += (em[sel] >> b) & 1
        dst = vbase[sel] + rank
        f0, f1 = ff[sel], ff[sel + stride[a]]
        t = (f0 / (f0 - f1)).astype(F)
        for ax in range(3):
            g = idx3[ax][sel].astype(F)
            if ax == a:
                g = (g + t).astype(F)
            verts[dst, ax] = (org[ax] + (g * vs).astype(F)).astype(F)
        if col is not None:
            c0, c1 = col[sel], col[sel + stride[a]]
            colors[dst] = (c0 + (t[:, None] * (c1 - c0).astype(F)).astype(F)).astype(F)
    cs = case.ravel()
    ntri = NTRI[cs].astype(np.int64)
    tbase = np.concatenate([[0], np.cumsum(ntri)[:-1]])
    faces = np.zeros((int(ntri.sum()), 3), dtype=np.int32)
    act = np.nonzero(ntri)[0]
    for t in range(MAX_TRIS):
        cells = act[ntri[act] > t]
        if cells.size == 0:
            break
        for k in range(3):
            e = TABLE[cs[cells], 3 * t + k].astype(np.int64)
            a, j = e // 4, e % 4
            u = np.where(a == 0, 1, 0)
            v = np.where(a == 2, 1, 2)
            st = np.array(stride)
            owner = cells + (j & 1) * st[u] + (j >> 1) * st[v]
            m = em[owner].astype(np.int64)
            rank = np.where(a > 0, m & 1, 0) + np.where(a > 1, (m >> 1) & 1, 0)
            faces[tbase[cells] + t, k] = (vbase[owner] + rank).astype(np.int32)
    return verts, faces, colors


def edge_use_counts(faces):
    """Directed-edge bookkeeping of a triangle list: dict (lo, hi) -> [uses lo->hi, uses hi->lo]."""
    d = {}
    for tri in np.asarray(faces):
        for k in range(3):
            a, b = int(tri[k]), int(tri[(k + 1) % 3])
            ent = d.setdefault((min(a, b), max(a, b)), [0, 0])
            ent[0 if a < b else 1] += 1
    return d
