"""XDMF3 time-series writer (mesh once + one nodal scalar per write), no HDF5 needed.

Keeps the layout dolfinx's ``XDMFFile.write_mesh`` / ``write_function`` produce for the
reference (reference: run_with_diamond.py:415-424, :483-484; io_utilities/xdmf_utils.py:5-26):

    <Xdmf Version="3.0"><Domain>
      <Grid Name="mesh" GridType="Uniform">  Topology(Triangle|Polyline) + Geometry(XY)
      <Grid Name="<function name>" GridType="Collection" CollectionType="Temporal">
         one <Grid> per write: xi:include of the mesh grid's Topology/Geometry,
         <Time Value="t"/>, <Attribute Name="<function name>" AttributeType="Scalar" Center="Node">

HDF5 is not available in this image, so heavy data goes to raw little-endian files
(``Format="Binary"``, one file per DataItem under ``<stem>_data/``; paths are relative to the
.xdmf file).  ``io_utilities.xdmf_extract`` reads the same files back.
"""
from __future__ import annotations

import os
from xml.sax.saxutils import quoteattr

import numpy as np


def _safe(name):
    return "".join(ch if ch.isalnum() else "_" for ch in name)


class XDMFFile:
    def __init__(self, comm, path, mode="w"):
        if mode != "w":
            raise NotImplementedError("only mode 'w' is supported")
        self.path = str(path)
        self.folder = os.path.dirname(os.path.abspath(self.path))
        self.stem = os.path.splitext(os.path.basename(self.path))[0]
        self.data_dir = f"{self.stem}_data"
        os.makedirs(os.path.join(self.folder, self.data_dir), exist_ok=True)
        self._mesh_xml = None
        self._series = {}       # function name -> [relative heavy-data file, list of per-step grid XML, bytes written]
        self._n_nodes = 0
        self._closed = False
        self._unflushed = 0

    def _dump(self, rel, array):
        array.tofile(os.path.join(self.folder, rel))

    def write_mesh(self, domain):
        x = np.ascontiguousarray(domain.geometry.x[:, :2], dtype="<f8")
        cells = np.ascontiguousarray(domain.cells, dtype="<i8")
        self._n_nodes = x.shape[0]
        topo_rel = f"{self.data_dir}/mesh_topology.bin"
        geom_rel = f"{self.data_dir}/mesh_geometry.bin"
        self._dump(topo_rel, cells)
        self._dump(geom_rel, x)
        kind = "Triangle" if cells.shape[1] == 3 else "PolyLine"
        self._mesh_xml = (
            f'    <Grid Name="mesh" GridType="Uniform">\n'
            f'      <Topology TopologyType="{kind}" NumberOfElements="{cells.shape[0]}" NodesPerElement="{cells.shape[1]}">\n'
            f'        <DataItem Dimensions="{cells.shape[0]} {cells.shape[1]}" NumberType="Int" Precision="8" '
            f'Format="Binary" Endian="Little">{topo_rel}</DataItem>\n'
            f'      </Topology>\n'
            f'      <Geometry GeometryType="XY">\n'
            f'        <DataItem Dimensions="{x.shape[0]} 2" NumberType="Float" Precision="8" '
            f'Format="Binary" Endian="Little">{geom_rel}</DataItem>\n'
            f'      </Geometry>\n'
            f'    </Grid>\n')
        self._flush()

    FLUSH_EVERY = 32            # the light-data XML is rewritten every this many steps and on close()

    def write_function(self, u, t=0.0):
        """``u``: object with ``.name`` and ``.x.array`` (nodal values), like a dolfinx Function.  The values are
        appended to ONE heavy-data file per function (``Seek`` offsets in the DataItems) and the per-step grid text
        is built once, so a run of S steps costs O(S) host work; the XML is complete after ``close()`` (and after
        every ``FLUSH_EVERY`` steps, so that an interrupted run leaves a readable file)."""
        if self._mesh_xml is None:
            raise RuntimeError("write_mesh must be called before write_function")
        values = np.ascontiguousarray(u.x.array[: self._n_nodes], dtype="<f8")
        name = getattr(u, "name", "f")
        if name not in self._series:
            rel = f"{self.data_dir}/{_safe(name)}.bin"
            open(os.path.join(self.folder, rel), "wb").close()
            self._series[name] = [rel, [], 0]
        entry = self._series[name]
        rel, grids, offset = entry
        with open(os.path.join(self.folder, rel), "ab") as f:
            values.tofile(f)
        q = quoteattr(name)
        grids.append(
            f'      <Grid Name={q} GridType="Uniform">\n'
            f'        <xi:include xpointer="xpointer(/Xdmf/Domain/Grid[@GridType=\'Uniform\'][1]/*[self::Topology or self::Geometry])" />\n'
            f'        <Time Value="{float(t)!r}" />\n'
            f'        <Attribute Name={q} AttributeType="Scalar" Center="Node">\n'
            f'          <DataItem Dimensions="{self._n_nodes} 1" NumberType="Float" Precision="8" '
            f'Format="Binary" Endian="Little" Seek="{offset}">{rel}</DataItem>\n'
            f'        </Attribute>\n'
            f'      </Grid>\n')
        entry[2] = offset + values.nbytes
        self._unflushed += 1
        if self._unflushed >= self.FLUSH_EVERY or sum(len(e[1]) for e in self._series.values()) == 1:
            self._flush()

    def _flush(self):
        out = ['<?xml version="1.0"?>\n<!DOCTYPE Xdmf SYSTEM "Xdmf.dtd" []>\n'
               '<Xdmf Version="3.0" xmlns:xi="https://www.w3.org/2001/XInclude">\n  <Domain>\n']
        out.append(self._mesh_xml or "")
        for name, (_rel, grids, _n) in self._series.items():
            out.append(f'    <Grid Name={quoteattr(name)} GridType="Collection" CollectionType="Temporal">\n')
            out.extend(grids)
            out.append('    </Grid>\n')
        out.append('  </Domain>\n</Xdmf>\n')
        tmp = self.path + ".tmp"
        with open(tmp, "w") as f:
            f.write("".join(out))
        os.replace(tmp, self.path)
        self._unflushed = 0

    def close(self):
        if not self._closed:
            self._flush()
            self._closed = True

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


def init_xdmf(domain, sim_folder, output_name):
    """Open ``<sim_folder>/<output_name>.xdmf`` and write the mesh (reference: xdmf_utils.py:5-26)."""
    xdmf = XDMFFile(getattr(domain, "comm", None), os.path.join(sim_folder, f"{output_name}.xdmf"), "w")
    xdmf.write_mesh(domain)
    return xdmf


def save_params(sim_folder, params_dict):
    """``key = value`` lines in ``params.txt`` (reference: xdmf_utils.py:29-44)."""
    with open(os.path.join(sim_folder, "params.txt"), "w") as f:
        for key, val in params_dict.items():
            f.write(f"{key} = {val}\n")
