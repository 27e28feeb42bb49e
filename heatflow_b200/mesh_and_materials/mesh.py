"""Mesh container with the reference's ``Mesh`` interface, backed by the in-repo mesher.

Reference interface kept (reference: mesh_and_materials/mesh.py:36-195): ``Mesh(name,
boundaries, materials)``, ``build_mesh()``, ``write(filename)``, ``to_dolfinx()``,
``Mesh.msh_to_dolfinx(filename)``, module-level ``COMM`` and ``SCALE``.  ``build_mesh``
sets ``mat._tag`` / ``mat.tag`` / ``self.material_tags`` to 1..n in material order exactly
as the gmsh path does (mesh.py:113-126).  The "dolfinx" conversion returns light-weight
stand-ins (``Domain``, ``MeshTags``) that expose what the runners read:
``domain.geometry.x`` ([N,3], z, r, 0), ``cell_tags.values`` / ``.indices``.
"""
from __future__ import annotations

import os

import numpy as np

from .materials import Material  # noqa: F401  (re-export, like the reference's star imports)
from .mesher import MeshArrays, triangulate_rectangles
from .msh_io import read_msh, write_msh


class _SerialComm:
    """Single-process communicator (the reference's optional-MPI fallback, mesh.py:3-13)."""
    rank = 0
    size = 1

    def Barrier(self):
        pass


COMM = _SerialComm()
SCALE = 1e6  # 1 model unit = 1 um (kept for import compatibility; unused, as in the reference)


class _Geometry:
    def __init__(self, x):
        self.x = x
        self.dim = 2


class Domain:
    """What the runners need from ``dolfinx.mesh.Mesh``: node coordinates + cell connectivity."""

    def __init__(self, arrays: MeshArrays):
        self.arrays = arrays
        x = np.zeros((arrays.num_nodes, 3), dtype=np.float64)
        x[:, :2] = arrays.nodes
        self.geometry = _Geometry(x)
        self.comm = COMM

    @property
    def cells(self):
        return self.arrays.tris


class MeshTags:
    def __init__(self, values, dim=2):
        self.values = np.asarray(values)
        self.indices = np.arange(len(self.values), dtype=np.int32)
        self.dim = dim


class Mesh:
    # extra knobs of the in-repo mesher (not in the reference); defaults reproduce cfg sizes
    growth = 1.3
    size_scale = 1.0
    method = "quadtree"       # 'quadtree' (graded points + Delaunay, quality bound) or 'rows' (round-1 row zipper)

    def __init__(self, name, boundaries, materials):
        if not isinstance(name, str):
            raise TypeError("name must be a string")
        if len(boundaries) != 4:
            raise ValueError("boundaries must be 4 floats")
        self.name = name
        self.boundaries = [float(b) for b in boundaries]
        self.materials = list(materials)
        self.material_tags = {}
        self.mesh = None

    def _check_mesh(self, base_bounds):
        """Same three checks as the reference (mesh.py:46-77)."""
        seen = {tuple(round(x, 12) for x in base_bounds): "BASE"}
        for m in self.materials:
            key = tuple(round(x, 12) for x in m.boundaries)
            if key in seen:
                raise RuntimeError(
                    f"Duplicate rectangle:\n    {m.name} has boundaries {key}\n    already used by {seen[key]}")
            seen[key] = m.name
        for m in self.materials:
            bx, BX, by, BY = m.boundaries
            if BX - bx <= 0 or BY - by <= 0:
                raise ValueError(f"{m.name}: invalid rectangle (bx,BX,by,BY) = {m.boundaries}")
        print('no mesh errors found')

    def build_mesh(self):
        self._check_mesh(self.boundaries)
        for i, mat in enumerate(self.materials):
            if mat.mesh_size is None:
                raise ValueError(f"{mat.name}: mesh_size required")
            mat._tag = mat.tag = i + 1
            self.material_tags[mat.name] = i + 1
        self.mesh = triangulate_rectangles(
            [m.boundaries for m in self.materials], [m.mesh_size for m in self.materials],
            bounds=self.boundaries, growth=self.growth, size_scale=self.size_scale, method=self.method)
        return self.mesh

    def to_dolfinx(self, *, comm=COMM, gdim: int = 2, rank: int = 0):
        if self.mesh is None:
            raise RuntimeError("Mesh not built – call build_mesh() first.")
        return Domain(self.mesh), MeshTags(self.mesh.cell_tag), MeshTags(np.zeros(0, np.int32), dim=1)

    _msh_cache = {}          # (path, mtime, size) -> parsed arrays: a sweep loads the same file for several contexts

    @staticmethod
    def _load_msh(filename):
        """Arrays of an MSH file.  Parsing the ASCII file takes 0.8 s at 1.4e5 nodes, so the arrays are kept in memory for
        the process and in a binary side file NEXT TO the mesh folder (`<folder>.msh_cache.npz`, keyed by the file's
        size and modification time; the folder itself keeps the reference's two files only) for the other ranks and
        the next run."""
        st = os.stat(filename)
        key = (os.path.abspath(filename), st.st_mtime_ns, st.st_size)
        hit = Mesh._msh_cache.get(key)
        if hit is not None:
            return hit
        side = os.path.normpath(os.path.dirname(os.path.abspath(filename))) + ".msh_cache.npz"
        try:
            with np.load(side) as z:
                if (str(z["name"]), int(z["mtime_ns"]), int(z["size"])) == (os.path.basename(filename), key[1], key[2]):
                    hit = (z["nodes"], z["tris"], z["tag"])
        except (OSError, KeyError, ValueError):
            hit = None
        if hit is None:
            nodes, tris, tag, _ = read_msh(filename)
            hit = (nodes, tris, tag)
            try:                                                  # best effort: a read-only location just skips the cache
                tmp = f"{side}.{os.getpid()}.tmp.npz"
                np.savez(tmp, nodes=nodes, tris=tris, tag=tag, name=os.path.basename(filename), mtime_ns=key[1], size=key[2])
                os.replace(tmp, side)
            except OSError:
                pass
        Mesh._msh_cache.clear()                                   # one mesh at a time is enough (width groups come in turn)
        Mesh._msh_cache[key] = hit
        return hit

    @staticmethod
    def msh_to_dolfinx(filename: str, *, comm=COMM, gdim: int = 2, rank: int = 0):
        nodes, tris, tag = (a.copy() for a in Mesh._load_msh(filename))
        arrays = MeshArrays(nodes, tris, tag)
        return Domain(arrays), MeshTags(arrays.cell_tag), MeshTags(np.zeros(0, np.int32), dim=1)

    def write(self, filename: str):
        if self.mesh is None:
            raise RuntimeError("Mesh not built – call build_mesh() first.")
        names = {i + 1: m.name for i, m in enumerate(self.materials)}
        write_msh(filename, self.mesh.nodes, self.mesh.tris, self.mesh.cell_tag, names)
