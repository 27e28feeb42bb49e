"""MSH 4.1 ASCII reader / writer for triangle meshes with physical surface groups.

The reference persists its mesh with ``gmsh.write`` (reference: mesh_and_materials/mesh.py:191-195;
gmsh default = MSH 4.1 ASCII; only elements of physical groups are saved, i.e. triangles)
and reads it back through ``gmsh.open`` + ``gmshio.model_to_mesh`` (reference:
run_with_diamond.py:240-245), which turns the *physical* tag of each triangle's surface into
``cell_tags``.  This module keeps that on-disk contract without gmsh:

* ``write_msh`` emits ``$MeshFormat/$PhysicalNames/$Entities/$Nodes/$Elements`` with one
  surface entity + one physical group per material (names = material names),
* ``read_msh`` parses the same sections (also from gmsh-written files: point/curve
  entities and non-triangle element blocks are skipped, node tags may be sparse) and
  returns ``(nodes[N,2], tris[E,3], cell_tag[E], physical_names)`` with the physical tag per cell.
"""
from __future__ import annotations

import io

import numpy as np

TRI3 = 2  # gmsh element type of the 3-node triangle


def _rows_to_text(out, arr):
    """Rows of a 2-D array as space-separated lines: integers verbatim, floats as the shortest repr that
    round-trips (plain Python formatting of `tolist()` rows is ~2 x faster than np.savetxt here)."""
    arr = np.asarray(arr)
    if arr.ndim == 1:
        arr = arr[:, None]
    if arr.dtype.kind in "iu":
        fmt = " ".join(["%d"] * arr.shape[1])
        out.write("\n".join(fmt % tuple(r) for r in arr.tolist()))
    else:
        out.write("\n".join(" ".join(map(repr, r)) for r in arr.tolist()))
    out.write("\n")


def write_msh(path, nodes, tris, cell_tag, names):
    """``names``: dict physical tag -> name (tag == surface entity tag, as in the reference
    where both are numbered 1..n in material order, mesh.py:113-126)."""
    nodes = np.asarray(nodes, dtype=np.float64)
    tris = np.asarray(tris, dtype=np.int64)
    cell_tag = np.asarray(cell_tag, dtype=np.int64)
    tags = sorted(int(t) for t in np.unique(cell_tag))
    n_nodes, n_tris = nodes.shape[0], tris.shape[0]
    # classify every node on the lowest-tag surface that uses it
    owner = np.full(n_nodes, np.iinfo(np.int64).max, dtype=np.int64)
    np.minimum.at(owner, tris.ravel(), np.repeat(cell_tag, 3))
    out = io.StringIO()
    out.write("$MeshFormat\n4.1 0 8\n$EndMeshFormat\n")
    out.write(f"$PhysicalNames\n{len(tags)}\n")
    for t in tags:
        out.write(f'2 {t} "{names.get(t, f"surface_{t}")}"\n')
    out.write("$EndPhysicalNames\n")
    out.write(f"$Entities\n0 0 {len(tags)} 0\n")
    for t in tags:
        used = nodes[np.unique(tris[cell_tag == t])]
        lo, hi = used.min(axis=0), used.max(axis=0)
        out.write(f"{t} {lo[0]:.17g} {lo[1]:.17g} 0 {hi[0]:.17g} {hi[1]:.17g} 0 1 {t} 0\n")
    out.write("$EndEntities\n")
    blocks = [(t, np.flatnonzero(owner == t)) for t in tags]
    blocks = [(t, idx) for t, idx in blocks if idx.size]
    out.write(f"$Nodes\n{len(blocks)} {n_nodes} 1 {n_nodes}\n")
    for t, idx in blocks:
        out.write(f"2 {t} 0 {idx.size}\n")
        _rows_to_text(out, idx + 1)
        xyz = np.zeros((idx.size, 3))
        xyz[:, :2] = nodes[idx]
        _rows_to_text(out, xyz)
    out.write("$EndNodes\n")
    eblocks = [(t, np.flatnonzero(cell_tag == t)) for t in tags]
    out.write(f"$Elements\n{len(eblocks)} {n_tris} 1 {n_tris}\n")
    for t, idx in eblocks:
        out.write(f"2 {t} {TRI3} {idx.size}\n")
        rows = np.column_stack((idx + 1, tris[idx] + 1))
        _rows_to_text(out, rows)
    out.write("$EndElements\n")
    with open(path, "w") as f:
        f.write(out.getvalue())


def _sections(text):
    """Split an MSH file into {section name: list of lines}."""
    secs, name, buf = {}, None, []
    for line in text.splitlines():
        s = line.strip()
        if not s:
            continue
        if s.startswith("$End"):
            secs[name] = buf
            name, buf = None, []
        elif s.startswith("$"):
            name, buf = s[1:], []
        elif name is not None:
            buf.append(s)
    return secs


def read_msh(path):
    with open(path, "r") as f:
        secs = _sections(f.read())
    if "MeshFormat" not in secs:
        raise ValueError(f"{path}: not an MSH file ($MeshFormat missing)")
    version = float(secs["MeshFormat"][0].split()[0])
    if not (4.0 <= version < 5.0) or int(secs["MeshFormat"][0].split()[1]) != 0:
        raise ValueError(f"{path}: only MSH 4.x ASCII is supported (found {secs['MeshFormat'][0]!r})")
    names = {}
    for line in secs.get("PhysicalNames", [])[1:]:
        dim, tag, nm = line.split(None, 2)
        if int(dim) == 2:
            names[int(tag)] = nm.strip().strip('"')
    # surface entity -> first physical tag
    surf_phys = {}
    ent = secs.get("Entities")
    if ent:
        npnt, ncur, nsur, _ = (int(v) for v in ent[0].split())
        base = 1 + npnt + ncur
        for line in ent[base:base + nsur]:
            tok = line.split()
            nphys = int(tok[7])
            if nphys:
                surf_phys[int(tok[0])] = int(tok[8])
    # nodes
    lines = secs["Nodes"]
    nblocks, ntot = int(lines[0].split()[0]), int(lines[0].split()[1])
    tags = np.empty(ntot, dtype=np.int64)
    xyz = np.empty((ntot, 3), dtype=np.float64)
    pos, got = 1, 0
    for _ in range(nblocks):
        _, _, parametric, nb = (int(v) for v in lines[pos].split())
        pos += 1
        if nb:
            tags[got:got + nb] = np.array(lines[pos:pos + nb], dtype=np.int64)
            pos += nb
            blk = np.array(" ".join(lines[pos:pos + nb]).split(), dtype=np.float64).reshape(nb, -1)
            xyz[got:got + nb] = blk[:, :3]
            pos += nb
            got += nb
    order = np.argsort(tags, kind="stable")
    tags, xyz = tags[order], xyz[order]
    # triangles
    lines = secs["Elements"]
    nblocks = int(lines[0].split()[0])
    pos = 1
    conn, ctag, etags = [], [], []
    for _ in range(nblocks):
        dim, etag, etype, nb = (int(v) for v in lines[pos].split())
        pos += 1
        if dim == 2 and etype == TRI3 and nb:
            blk = np.array(" ".join(lines[pos:pos + nb]).split(), dtype=np.int64).reshape(nb, 4)
            conn.append(blk[:, 1:])
            etags.append(blk[:, 0])
            ctag.append(np.full(nb, surf_phys.get(etag, etag), dtype=np.int64))
        pos += nb
    if not conn:
        raise ValueError(f"{path}: no 3-node triangles found")
    eorder = np.argsort(np.concatenate(etags), kind="stable")  # cells in element-tag order
    conn = np.concatenate(conn)[eorder]
    ctag = np.concatenate(ctag)[eorder]
    # node tags -> 0-based contiguous indices; drop nodes no triangle references
    idx = np.searchsorted(tags, conn)
    if np.any(tags[np.minimum(idx, len(tags) - 1)] != conn):
        raise ValueError(f"{path}: element references an unknown node tag")
    used = np.zeros(len(tags), dtype=bool)
    used[idx.ravel()] = True
    remap = np.cumsum(used) - 1
    return xyz[used, :2].copy(), remap[idx].astype(np.int32), ctag.astype(np.int32), names
