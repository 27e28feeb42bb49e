"""Deterministic conforming triangulator for a union of axis-aligned material rectangles.

Replaces the gmsh call of the reference (reference: mesh_and_materials/mesh.py:81-149,
``geo`` kernel rectangles + ``Box``/``Min`` size fields + ``generate(2)``).  gmsh is not
available and its Delaunay/frontal output cannot be reproduced, so this is a different
algorithm that honours the same inputs and the same contract:

* one triangulated surface per material rectangle, conforming across shared edges,
* target edge length inside a material = its ``mesh_size`` (the reference's ``VIn``),
  never larger than the largest material size (the reference's ``VOut``),
* cell tag = 1 + index of the material in the list (reference: mesh.py:113-126).

Both methods use one gradation-limited size field

    h(z, r) = min_m ( size_m + (growth-1) * dist((z, r), rect_m) )     capped at max_m size_m

so fine layers (0.02 um couplers) blend into coarse ones (10 um diamonds) isotropically.

``method="quadtree"`` (default) - graded point cloud + Delaunay.  Every rectangle edge (split where another
rectangle's corner lies on it) gets a graded 1-D subdivision by h; the interior of every rectangle gets the
corners of a quadtree refined until a cell is no larger than ~h at its centre and corners; interior points closer
than 0.55 h to the rectangle's own boundary are dropped, so that every boundary sub-edge satisfies the Gabriel
condition and appears in the Delaunay triangulation (``scipy.spatial.Delaunay``, Qhull) of all points - which is
then conforming to the material interfaces by construction (checked; a missing sub-edge gets its midpoint
inserted).  Uniform regions come out as right isosceles triangles, graded regions keep every angle above ~14
degrees (tests/test_mesher.py asserts 12 degrees and a 6:1 longest-edge/height bound at the cfgs' own sizes).
Nodes are numbered z-major (z outer, r inner).

``method="rows"`` - the round-1 "row zipper": constant-z rows over the whole r range, consecutive rows stitched by
merging their sorted r-sequences.  Fast and banded, but rows that are fine because of a thin layer stay fine at
radii where the r-spacing has grown to micrometres: slivers below 0.1 degrees in the gasket and diamond regions
of the with-diamond layouts.  Kept for comparison.

Everything is numpy / scipy; the only Python loops are over rectangles, edges and refinement levels.
"""
from __future__ import annotations

import numpy as np
from scipy.spatial import Delaunay

__all__ = ["triangulate_rectangles", "size_field", "MeshArrays"]


class MeshArrays:
    """Plain container: ``nodes [N,2] f64`` (z, r), ``tris [E,3] i32`` (CCW), ``cell_tag [E] i32``."""

    def __init__(self, nodes, tris, cell_tag, row_ptr=None):
        self.nodes = np.ascontiguousarray(nodes, dtype=np.float64)
        self.tris = np.ascontiguousarray(tris, dtype=np.int32)
        self.cell_tag = np.ascontiguousarray(cell_tag, dtype=np.int32)
        self.row_ptr = row_ptr  # node offset of every z-row (None when read from disk)

    @property
    def num_nodes(self):
        return self.nodes.shape[0]

    @property
    def num_cells(self):
        return self.tris.shape[0]


def _breakpoints(values, span):
    """Sorted unique coordinates, merging values closer than 1e-9 of the span."""
    v = np.sort(np.asarray(values, dtype=np.float64))
    keep = [v[0]]
    tol = 1e-9 * span
    for x in v[1:]:
        if x - keep[-1] > tol:
            keep.append(x)
    return np.array(keep)


def size_field(z, r, rects, sizes, slope, hmax):
    """Gradation-limited target edge length at points (z, r) (broadcast numpy arrays)."""
    z = np.asarray(z, dtype=np.float64)
    r = np.asarray(r, dtype=np.float64)
    h = np.full(np.broadcast(z, r).shape, hmax, dtype=np.float64)
    for (z0, z1, r0, r1), s in zip(rects, sizes):
        dz = np.maximum(np.maximum(z0 - z, z - z1), 0.0)
        dr = np.maximum(np.maximum(r0 - r, r - r1), 0.0)
        np.minimum(h, s + slope * np.hypot(dz, dr), out=h)
    return h


def _subdivide(a, b, hfun):
    """Graded subdivision of [a, b]: node density 1/h, end points exact.

    ``hfun(t)`` returns the target spacing at coordinates ``t`` (vectorised).  The number
    of intervals is the rounded-up integral of 1/h; nodes sit at equal increments of that
    integral.  A constant field gives ``np.linspace`` (bit-identical for identical input).
    """
    length = b - a
    probe = hfun(np.array([a, 0.5 * (a + b), b]))
    hmin_guess = float(probe.min())
    m = int(min(max(64, 8 * np.ceil(length / hmin_guess)), 400000))
    t = np.linspace(a, b, m + 1)
    h = hfun(t)
    if h.max() - h.min() <= 1e-12 * h.max():
        n = max(1, int(np.ceil(length / h[0] - 1e-9)))
        return np.linspace(a, b, n + 1)
    dens = 1.0 / h
    cum = np.concatenate(([0.0], np.cumsum(0.5 * (dens[1:] + dens[:-1]) * np.diff(t))))
    n = max(1, int(np.ceil(cum[-1] - 1e-9)))
    pts = np.interp(np.arange(1, n) * (cum[-1] / n), cum, t)
    return np.concatenate(([a], pts, [b]))


def _zip_strip(ida, ca, idb, cb):
    """Triangulate the strip between a lower row (ids ``ida`` at coords ``ca``) and an upper
    row (``idb``, ``cb``); both coordinate arrays are strictly increasing and share end points.
    Returns CCW triangles in (z, r) with z = row direction."""
    m, n = len(ca) - 1, len(cb) - 1
    if m == n and np.array_equal(ca, cb):
        lo = np.stack([ida[:-1], idb[:-1], ida[1:]], axis=1)
        up = np.stack([ida[1:], idb[:-1], idb[1:]], axis=1)
        out = np.empty((2 * m, 3), dtype=np.int64)
        out[0::2] = lo
        out[1::2] = up
        return out
    # merge the "advance" events of both chains; ties advance the lower chain first
    keys = np.concatenate((ca[1:], cb[1:]))
    kind = np.concatenate((np.zeros(m, dtype=np.int64), np.ones(n, dtype=np.int64)))
    order = np.lexsort((kind, keys))
    kind = kind[order]
    i = np.cumsum(kind == 0) - (kind == 0)  # lower advances before this event
    j = np.cumsum(kind == 1) - (kind == 1)  # upper advances before this event
    out = np.empty((m + n, 3), dtype=np.int64)
    low = kind == 0
    out[low, 0] = ida[i[low]]
    out[low, 1] = idb[j[low]]
    out[low, 2] = ida[i[low] + 1]
    upm = ~low
    out[upm, 0] = ida[i[upm]]
    out[upm, 1] = idb[j[upm]]
    out[upm, 2] = idb[j[upm] + 1]
    return out


def triangulate_rectangles(rects, sizes, bounds=None, growth=1.3, size_scale=1.0, method="quadtree"):
    """Mesh the union of rectangles with ``method`` 'quadtree' (default) or 'rows' (see the module docstring).

    Parameters
    ----------
    rects : sequence of [zmin, zmax, rmin, rmax]
    sizes : target edge length per rectangle
    bounds : ignored - like gmsh in the reference, the mesh is the union of the material
        rectangles (run_no_diamond.py passes domain bounds wider than its materials)
    growth : maximum ratio between neighbouring edge lengths (size-field slope = growth-1)
    size_scale : multiplies every size (``0.5`` gives ~4x the nodes) - the refinement knob
        for the synthetic >= 1 M-DOF benchmark meshes

    Returns
    -------
    MeshArrays with ``cell_tag = 1 + index`` of the first rectangle containing the cell.
    """
    rects = [tuple(float(v) for v in rc) for rc in rects]
    sizes = [float(s) * float(size_scale) for s in sizes]
    if len(rects) == 0:
        raise ValueError("no materials to mesh")
    if any(s <= 0 for s in sizes):
        raise ValueError("mesh sizes must be positive")
    if method == "quadtree":
        return _triangulate_quadtree(rects, sizes, float(growth))
    if method != "rows":
        raise ValueError("method must be 'quadtree' or 'rows'")
    zlo = min(rc[0] for rc in rects)
    zhi = max(rc[1] for rc in rects)
    rlo = min(rc[2] for rc in rects)
    rhi = max(rc[3] for rc in rects)
    zb = _breakpoints([v for rc in rects for v in rc[:2]], zhi - zlo)
    rb = _breakpoints([v for rc in rects for v in rc[2:]], rhi - rlo)
    slope = float(growth) - 1.0
    hmax = max(sizes)

    def snap(v, grid):
        return grid[np.argmin(np.abs(grid - v))]

    rects = [(snap(a, zb), snap(b, zb), snap(c, rb), snap(d, rb)) for a, b, c, d in rects]

    # ---- row positions: 1-D field = min over r of h(z, r) --------------------------------
    def hz(t):
        h = np.full(t.shape, hmax)
        for (z0, z1, _, _), s in zip(rects, sizes):
            np.minimum(h, s + slope * np.maximum(np.maximum(z0 - t, t - z1), 0.0), out=h)
        return h

    zrows = [zb[:1]]
    row_zint = []  # z-interval index of the strip above each row
    for k in range(len(zb) - 1):
        pts = _subdivide(zb[k], zb[k + 1], hz)
        zrows.append(pts[1:])
        row_zint.extend([k] * (len(pts) - 1))
    zrows = np.concatenate(zrows)
    nrows = len(zrows)

    # ---- which material owns each (z-interval, r-interval) box ---------------------------
    box_tag = np.zeros((len(zb) - 1, len(rb) - 1), dtype=np.int32)
    for kz in range(len(zb) - 1):
        zc = 0.5 * (zb[kz] + zb[kz + 1])
        for kr in range(len(rb) - 1):
            rc_ = 0.5 * (rb[kr] + rb[kr + 1])
            for idx, (z0, z1, r0, r1) in enumerate(rects):
                if z0 <= zc <= z1 and r0 <= rc_ <= r1:
                    box_tag[kz, kr] = idx + 1
                    break

    # ---- nodes of every row, per r-interval ----------------------------------------------
    row_coords = []   # per row: list over r-intervals of coordinate arrays (end points shared)
    row_ptr = np.zeros(nrows + 1, dtype=np.int64)
    for i, zr in enumerate(zrows):
        segs = []
        hr = lambda t, zr=zr: size_field(zr, t, rects, sizes, slope, hmax)
        for kr in range(len(rb) - 1):
            segs.append(_subdivide(rb[kr], rb[kr + 1], hr))
        row_coords.append(segs)
        row_ptr[i + 1] = row_ptr[i] + 1 + sum(len(s) - 1 for s in segs)
    nn = int(row_ptr[-1])
    nodes = np.empty((nn, 2), dtype=np.float64)
    seg_off = []  # per row: offset (within the row) of the first node of every r-interval
    for i, segs in enumerate(row_coords):
        flat = np.concatenate([segs[0]] + [s[1:] for s in segs[1:]])
        nodes[row_ptr[i]:row_ptr[i + 1], 0] = zrows[i]
        nodes[row_ptr[i]:row_ptr[i + 1], 1] = flat
        off = np.zeros(len(segs), dtype=np.int64)
        off[1:] = np.cumsum([len(s) - 1 for s in segs[:-1]])
        seg_off.append(off)

    # ---- stitch consecutive rows ----------------------------------------------------------
    tri_blocks, tag_blocks = [], []
    for i in range(nrows - 1):
        kz = row_zint[i]
        for kr in range(len(rb) - 1):
            tag = box_tag[kz, kr]
            if tag == 0:
                continue  # hole in the rectangle layout
            ca, cb = row_coords[i][kr], row_coords[i + 1][kr]
            ida = row_ptr[i] + seg_off[i][kr] + np.arange(len(ca))
            idb = row_ptr[i + 1] + seg_off[i + 1][kr] + np.arange(len(cb))
            t = _zip_strip(ida, ca, idb, cb)
            tri_blocks.append(t)
            tag_blocks.append(np.full(len(t), tag, dtype=np.int32))
    tris = np.concatenate(tri_blocks)
    cell_tag = np.concatenate(tag_blocks)

    # ---- drop nodes that no triangle uses (only when the layout has holes) ---------------
    used = np.zeros(nn, dtype=bool)
    used[tris.ravel()] = True
    if not used.all():
        remap = np.cumsum(used) - 1
        tris = remap[tris]
        nodes = nodes[used]
        row_ptr = None
    return MeshArrays(nodes, tris.astype(np.int32), cell_tag, row_ptr)


# ----------------------------------------------------------------------------------------------------
# quadtree points + Delaunay
# ----------------------------------------------------------------------------------------------------
_SPLIT = 1.3        # a quadtree cell is split while its longer side exceeds _SPLIT * h: spacing in (0.65 h, 1.3 h]
_KEEP_OFF = 0.55    # interior points keep this many local sizes away from their rectangle's boundary


def _quadtree_interior_points(rect, hfun, H):
    """Corners of the leaves of a quadtree on ``rect`` whose root cells are ~H wide, strictly inside the rectangle."""
    z0, z1, r0, r1 = rect
    Lz, Lr = z1 - z0, r1 - r0
    nz = max(1, int(np.ceil(Lz / H - 1e-9)))                  # root cells no larger than H: where h is constant the legs
    nr = max(1, int(np.ceil(Lr / H - 1e-9)))                  # of the triangles are the target size (as np.linspace would)
    iz, ir = np.meshgrid(np.arange(nz, dtype=np.int64), np.arange(nr, dtype=np.int64), indexing="ij")
    level, i, j = 0, iz.ravel(), ir.ravel()
    leaves = []
    while True:
        dz, dr = Lz / (nz << level), Lr / (nr << level)
        h = hfun(z0 + (i + 0.5) * dz, r0 + (j + 0.5) * dr)
        for a in (0.0, 1.0):
            for b in (0.0, 1.0):
                h = np.minimum(h, hfun(z0 + (i + a) * dz, r0 + (j + b) * dr))
        split = max(dz, dr) > _SPLIT * h
        leaves.append((level, i[~split], j[~split]))
        if not split.any() or level >= 30:
            break
        i, j = i[split], j[split]
        i, j = np.concatenate([2 * i, 2 * i, 2 * i + 1, 2 * i + 1]), np.concatenate([2 * j, 2 * j + 1, 2 * j, 2 * j + 1])
        level += 1
    lmax = level
    pts = []
    for l, i, j in leaves:
        s = 1 << (lmax - l)
        for a in (0, 1):
            for b in (0, 1):
                pts.append(((i + a) * s << 32) | ((j + b) * s))          # one int64 per lattice point
    packed = np.unique(np.concatenate(pts))
    ci, cj = packed >> 32, packed & 0xffffffff
    NZ, NR = nz << lmax, nr << lmax
    inside = (ci > 0) & (ci < NZ) & (cj > 0) & (cj < NR)
    return np.stack([z0 + ci[inside] * (Lz / NZ), r0 + cj[inside] * (Lr / NR)], axis=1)


def _triangulate_quadtree(rects, sizes, growth):
    zlo, zhi = min(rc[0] for rc in rects), max(rc[1] for rc in rects)
    rlo, rhi = min(rc[2] for rc in rects), max(rc[3] for rc in rects)
    zb = _breakpoints([v for rc in rects for v in rc[:2]], zhi - zlo)
    rb = _breakpoints([v for rc in rects for v in rc[2:]], rhi - rlo)
    snap = lambda v, grid: float(grid[np.argmin(np.abs(grid - v))])
    rects = [(snap(a, zb), snap(b, zb), snap(c, rb), snap(d, rb)) for a, b, c, d in rects]
    slope, hmax = growth - 1.0, max(sizes)
    hfun = lambda z, r: size_field(z, r, rects, sizes, slope, hmax)

    # ---- boundary segments: every rectangle edge, cut where a corner of any rectangle lies on it
    corners = {(z, r) for z0, z1, r0, r1 in rects for z in (z0, z1) for r in (r0, r1)}
    segs = set()
    for z0, z1, r0, r1 in rects:
        for z in (z0, z1):
            cut = sorted({r0, r1} | {r for (zc, r) in corners if zc == z and r0 < r < r1})
            segs.update(("h", z, a, b) for a, b in zip(cut[:-1], cut[1:]))
        for r in (r0, r1):
            cut = sorted({z0, z1} | {z for (z, rc) in corners if rc == r and z0 < z < z1})
            segs.update(("v", r, a, b) for a, b in zip(cut[:-1], cut[1:]))
    # overlapping collinear pieces (an edge shared by rectangles whose corners differ) are cut at all their end points
    lines = {}
    for kind, c, a, b in segs:
        lines.setdefault((kind, c), []).append((a, b))
    bpts, bedges = [], []
    for (kind, c), spans in sorted(lines.items()):
        ends = np.array(sorted({v for ab in spans for v in ab}))
        for a, b in zip(ends[:-1], ends[1:]):
            mid = 0.5 * (a + b)
            if not any(lo <= mid <= hi for lo, hi in spans):
                continue
            if kind == "h":
                t = _subdivide(a, b, lambda t, c=c: hfun(np.full_like(t, c), t))
                bpts.append(np.stack([np.full_like(t, c), t], axis=1))
            else:
                t = _subdivide(a, b, lambda t, c=c: hfun(t, np.full_like(t, c)))
                bpts.append(np.stack([t, np.full_like(t, c)], axis=1))
            bedges.append(len(t))

    # ---- interior points per rectangle
    ipts = []
    for rc, sz in zip(rects, sizes):
        probe_z = np.array([rc[0], rc[1], rc[0], rc[1], 0.5 * (rc[0] + rc[1])])
        probe_r = np.array([rc[2], rc[2], rc[3], rc[3], 0.5 * (rc[2] + rc[3])])
        H = max(float(hfun(probe_z, probe_r).max()), min(sz, hmax))
        p = _quadtree_interior_points(rc, hfun, H)
        d = np.minimum.reduce([p[:, 0] - rc[0], rc[1] - p[:, 0], p[:, 1] - rc[2], rc[3] - p[:, 1]])
        ipts.append(p[d >= _KEEP_OFF * hfun(p[:, 0], p[:, 1])])

    span = max(zhi - zlo, rhi - rlo)
    res = 1e-9 * span                                           # coordinates closer than this are the same node

    def keys(points):
        """One sortable int64 per point (z-major): 30 bits per coordinate at resolution ``res``."""
        kz = np.round((points[:, 0] - zlo) / res).astype(np.int64)
        kr = np.round((points[:, 1] - rlo) / res).astype(np.int64)
        return (kz << 31) | kr

    def unique_sorted(points):
        k = keys(points)
        k, first = np.unique(k, return_index=True)               # sorted by key = z-major numbering
        return points[first], k

    def boundary_edges(point_keys):
        """(index pairs) of all boundary sub-edges in the numbering of the sorted points."""
        out = []
        for seg in bpts:
            idx = np.searchsorted(point_keys, keys(seg))
            out.append(np.stack([idx[:-1], idx[1:]], axis=1))
        return np.concatenate(out)

    extra = np.zeros((0, 2))
    for _attempt in range(6):
        pts, pkeys = unique_sorted(np.concatenate(bpts + ipts + [extra]))
        tri = Delaunay(pts).simplices.astype(np.int64)
        e = np.sort(np.concatenate([tri[:, [0, 1]], tri[:, [1, 2]], tri[:, [2, 0]]]), axis=1)
        have = np.unique(e[:, 0] * len(pts) + e[:, 1])
        be = np.sort(boundary_edges(pkeys), axis=1)
        bk = be[:, 0] * len(pts) + be[:, 1]
        pos = np.minimum(np.searchsorted(have, bk), len(have) - 1)
        missing = be[have[pos] != bk].tolist()
        if not missing:
            break
        # recover a missing interface edge by splitting it (the segment lists are refined in place)
        mids = np.array([0.5 * (pts[a] + pts[b]) for a, b in missing])
        for n, seg in enumerate(bpts):
            add = []
            for m in mids:
                on = (abs(seg[0, 0] - seg[-1, 0]) < 1e-30 and abs(m[0] - seg[0, 0]) <= 1e-12 * span and seg[0, 1] < m[1] < seg[-1, 1]) or \
                     (abs(seg[0, 1] - seg[-1, 1]) < 1e-30 and abs(m[1] - seg[0, 1]) <= 1e-12 * span and seg[0, 0] < m[0] < seg[-1, 0])
                if on:
                    add.append(m)
            if add:
                allp = np.concatenate([seg, np.array(add)])
                order = np.lexsort((allp[:, 1], allp[:, 0]))
                bpts[n] = allp[order]
    else:
        raise RuntimeError("mesher: could not recover all material interfaces in the Delaunay triangulation")

    p = pts[tri]
    area2 = (p[:, 1, 0] - p[:, 0, 0]) * (p[:, 2, 1] - p[:, 0, 1]) - (p[:, 2, 0] - p[:, 0, 0]) * (p[:, 1, 1] - p[:, 0, 1])
    cen = p.mean(axis=1)
    tag = np.zeros(len(tri), np.int32)
    for idx in range(len(rects) - 1, -1, -1):                             # the first rectangle containing the cell wins
        z0, z1, r0, r1 = rects[idx]
        tag[(cen[:, 0] > z0) & (cen[:, 0] < z1) & (cen[:, 1] > r0) & (cen[:, 1] < r1)] = idx + 1
    keep = (tag > 0) & (np.abs(area2) > 1e-14 * span * span * 1e-6)
    tri, tag, area2 = tri[keep], tag[keep], area2[keep]
    flip = area2 < 0
    tri[flip] = tri[flip][:, [0, 2, 1]]                                   # CCW
    used = np.zeros(len(pts), dtype=bool)
    used[tri.ravel()] = True
    if not used.all():
        remap = np.cumsum(used) - 1
        tri, pts = remap[tri], pts[used]
    return MeshArrays(pts, tri.astype(np.int32), tag, None)
