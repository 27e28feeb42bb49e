"""Point time-series extraction from the XDMF files written by ``xdmf_utils.XDMFFile``.

Same call signature and return value as the reference helper (reference:
io_utilities/xdmf_extract.py:6-60), which went through ``meshio.xdmf.TimeSeriesReader``;
meshio is not available here, so the XML is parsed directly (``Format="Binary"`` and
``Format="XML"`` DataItems, paths relative to the .xdmf file).
"""
from __future__ import annotations

import os
import xml.etree.ElementTree as ET

import numpy as np
from scipy.interpolate import griddata
from scipy.spatial import cKDTree

_DTYPES = {("Float", "8"): "<f8", ("Float", "4"): "<f4", ("Int", "8"): "<i8", ("Int", "4"): "<i4"}


def _read_item(item, folder):
    dims = [int(d) for d in item.get("Dimensions").split()]
    kind = item.get("NumberType") or item.get("DataType") or "Float"
    dtype = _DTYPES[(kind, item.get("Precision", "8"))]
    fmt = item.get("Format", "XML")
    if fmt == "XML":
        return np.array(item.text.split(), dtype=dtype).reshape(dims)
    if fmt == "Binary":
        data = np.fromfile(os.path.join(folder, item.text.strip()), dtype=dtype, offset=int(item.get("Seek", 0)),
                           count=int(np.prod(dims)))
        return data.reshape(dims)
    raise ValueError(f"unsupported DataItem format {fmt!r} (HDF5 is not available in this build)")


class TimeSeriesReader:
    """Minimal stand-in for meshio's reader: ``read_points_cells``, ``num_steps``, ``read_data``."""

    def __init__(self, path):
        self.folder = os.path.dirname(os.path.abspath(path))
        domain = ET.parse(path).getroot().find("Domain")
        self._mesh = None
        self._steps = []
        for grid in domain.findall("Grid"):
            if grid.get("GridType") == "Uniform" and self._mesh is None:
                self._mesh = grid
            elif grid.get("GridType") == "Collection":
                for g in grid.findall("Grid"):
                    self._steps.append(g)
        if self._mesh is None:
            raise ValueError(f"{path}: no uniform mesh grid")
        self.num_steps = len(self._steps)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        return False

    def read_points_cells(self):
        pts = _read_item(self._mesh.find("Geometry").find("DataItem"), self.folder)
        cells = _read_item(self._mesh.find("Topology").find("DataItem"), self.folder)
        xyz = np.zeros((pts.shape[0], 3))
        xyz[:, : pts.shape[1]] = pts
        return xyz, cells

    def read_data(self, i):
        g = self._steps[i]
        t = float(g.find("Time").get("Value"))
        point_data = {}
        for att in g.findall("Attribute"):
            vals = _read_item(att.find("DataItem"), self.folder)
            point_data[att.get("Name")] = vals.reshape(vals.shape[0], -1)[:, 0] if vals.ndim > 1 and vals.shape[1] == 1 else vals
        return t, point_data, {}


def extract_point_timeseries_xdmf(xdmf_path, function_name, query_points, method="nearest"):
    """Returns ``(times [S], data [n_points, S])`` sampled at ``query_points`` [(x, y), ...]."""
    with TimeSeriesReader(xdmf_path) as reader:
        points, _ = reader.read_points_cells()
        pts2d = points[:, :2]
        tree = cKDTree(pts2d)
        n_steps = reader.num_steps
        times = np.empty(n_steps)
        data = np.empty((len(query_points), n_steps))
        nearest = [tree.query(qp)[1] for qp in query_points]
        for i in range(n_steps):
            t, point_data, _ = reader.read_data(i)
            times[i] = t
            vals = point_data[function_name]
            for j, qp in enumerate(query_points):
                data[j, i] = vals[nearest[j]] if method == "nearest" else griddata(pts2d, vals, qp, method="linear")
    order = np.argsort(times)
    return times[order], data[:, order]
