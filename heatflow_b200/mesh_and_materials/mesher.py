"""Deterministic conforming triangulator for a union of axis-aligned material rectangles.

Replaces the gmsh call of the reference (reference: mesh_and_materials/mesh.py:81-149,
``geo`` kernel rectangles + ``Box``/``Min`` size fields + ``generate(2)``).  gmsh is not
available and its Delaunay/frontal output cannot be reproduced, so this is a different
algorithm that honours the same inputs and the same contract:

* one triangulated surface per material rectangle, conforming across shared edges,
* target edge length inside a material = its ``mesh_size`` (the reference's ``VIn``),
  never larger than the largest material size (the reference's ``VOut``),
* cell tag = 1 + index of the material in the list (reference: mesh.py:113-126).

Algorithm ("row zipper").  The domain is cut into constant-z rows.  Row positions are a
graded 1-D subdivision of every interval between material z-breakpoints; along every row
the r-nodes are a graded 1-D subdivision of every interval between material
r-breakpoints.  Both use one gradation-limited size field

    h(z, r) = min_m ( size_m + (growth-1) * dist((z, r), rect_m) )     capped at max_m size_m

so fine layers (0.02 um couplers) blend into coarse ones (10 um diamonds) isotropically.
Consecutive rows are stitched by merging their two sorted r-sequences ("zipper"), which
yields a conforming triangulation for any pair of node counts; identical rows give the
regular two-triangles-per-quad pattern.  Nodes are numbered row-major (z outer, r inner),
which keeps the P1 operator banded (good gather locality for the GPU SpMV).

Everything is numpy; the only Python loops are over rows and r-intervals.
"""
from __future__ import annotations

import numpy as np

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


def triangulate_rectangles(rects, sizes, bounds=None, growth=1.3, size_scale=1.0):
    """Mesh the union of rectangles.

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
