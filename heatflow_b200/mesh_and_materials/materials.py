"""Material record: a named axis-aligned rectangle with physical properties.

Mirrors the reference data class (reference: mesh_and_materials/materials.py:16-34) so
runner code written against the reference keeps working: same constructor arguments,
same attribute names (``name``, ``boundaries``, ``properties``, ``mesh_size``) and the
same validation errors.  ``_tag``/``tag`` are attached by the mesher exactly as the
reference mesher does (reference: mesh_and_materials/mesh.py:113-126).

Coordinates are ``[zmin, zmax, rmin, rmax]`` in metres (the reference calls them
xmin/xmax/ymin/ymax; x is the axial z direction, y the radial r direction).
"""
from __future__ import annotations

from numbers import Real


class Material:
    def __init__(self, name, boundaries, properties=None, mesh_size=None, material_tag=None):
        if not isinstance(name, str):
            raise TypeError(f"name must be a string, got {type(name)}")
        try:
            n = len(boundaries)
        except TypeError:
            n = -1
        if n != 4:
            raise ValueError("boundaries must be [xmin,xmax,ymin,ymax]")
        lo_x, hi_x, lo_y, hi_y = (float(v) for v in boundaries)
        if not (lo_x < hi_x and lo_y < hi_y):
            raise ValueError(f"Invalid boundaries {boundaries}")
        if mesh_size is not None and (isinstance(mesh_size, bool) or not isinstance(mesh_size, Real)):
            raise TypeError(f"mesh_size must be a number, got {type(mesh_size)}")
        self.name = name
        self.boundaries = [lo_x, hi_x, lo_y, hi_y]
        self.mesh_size = None if mesh_size is None else float(mesh_size)
        self.properties = dict(properties) if properties else {}
        if material_tag is not None:
            self._tag = self.tag = int(material_tag)

    def contains(self, x, y):
        """True when (x, y) lies in the closed rectangle."""
        lo_x, hi_x, lo_y, hi_y = self.boundaries
        return lo_x <= x <= hi_x and lo_y <= y <= hi_y

    def __repr__(self):
        return f"Material({self.name!r}, bounds={self.boundaries}, size={self.mesh_size})"
