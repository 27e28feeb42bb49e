"""Shared builders for the tests: one cfg -> (mesh arrays, coefficient tables, BC sets, heating).

The product-side objects (mesher, RowDirichletBC, problem.*) and the oracle-side objects
(oracle.heat_oracle) are built from the same cfg so the tests can compare them.
"""
import copy
import os

import numpy as np
import yaml

from heatflow_b200 import fem, problem
from heatflow_b200.mesh_and_materials import Domain, Mesh
from oracle import heat_oracle as ho

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def load_cfg(name):
    with open(os.path.join(ROOT, "cfgs", f"{name}.yaml")) as f:
        cfg = yaml.safe_load(f)
    cfg = copy.deepcopy(cfg)
    cfg["heating"]["file"] = os.path.join(ROOT, cfg["heating"]["file"])
    return cfg


class Case:
    pass


def build_case(cfg_name="geballe_no_diamond", size_scale=8.0, growth=1.3, method=None, arrays=None):
    """Mesh + problem data for a cfg at a coarsened mesh size (size_scale > 1 = coarser).  `arrays`
    (nodes, tris, cell_tag): use this mesh instead of running the mesher (reference golden files)."""
    cfg = load_cfg(cfg_name)
    with_diamond = "p_diam" in cfg["mats"]
    mats, bounds, info = (problem.stack_with_diamond if with_diamond else problem.stack_no_diamond)(cfg)
    mesh = Mesh("mesh.msh", bounds, mats)
    mesh.size_scale = size_scale
    mesh.growth = growth
    if method is not None:
        mesh.method = method
    import contextlib, io
    if arrays is not None:
        from heatflow_b200.mesh_and_materials.mesher import MeshArrays
        for tag, m in enumerate(mats, start=1):                # tag = material list index + 1 (mesh.py:113-126)
            m._tag = m.tag = tag
        arrays = MeshArrays(*arrays)
    else:
        with contextlib.redirect_stdout(io.StringIO()):
            arrays = mesh.build_mesh()
    c = Case()
    c.name = cfg_name
    c.cfg, c.mats, c.bounds, c.arrays = cfg, mats, bounds, arrays
    c.nodes, c.tris, c.cell_tag = arrays.nodes, arrays.tris, arrays.cell_tag
    c.tags = np.array([m.tag for m in mats], dtype=np.int32)
    c.kappa_t = np.array([m.properties["k"] for m in mats])
    c.rhoc_t = np.array([m.properties["rho_cv"] for m in mats])
    c.kappa_c = c.kappa_t[c.cell_tag - 1]
    c.rhoc_c = c.rhoc_t[c.cell_tag - 1]
    c.num_steps = int(cfg["timing"]["num_steps"])
    c.dt = float(cfg["timing"]["t_final"]) / c.num_steps
    c.ic = float(cfg["heating"]["ic_temp"])
    c.fwhm = float(cfg["heating"]["fwhm"])
    c.coeff = problem.gaussian_coeff(c.fwhm)
    c.heat_t, c.heat_T = problem.read_heating_curve(cfg["heating"]["file"])
    c.amps = problem.heating_amplitudes((np.arange(c.num_steps) + 1) * c.dt, c.heat_t, c.heat_T, c.ic)
    pc = next(m for m in mats if m.name == "p_coupler")
    c.heating_z = pc.boundaries[0]
    c.r_sample = info["r_sample"]
    # product-side BCs
    c.domain = Domain(arrays)
    c.V = fem.functionspace(c.domain, ("Lagrange", 1))
    gaussian = lambda x, y, t: (ho.heating_amplitude(t, c.heat_t, c.heat_T, c.ic) - c.ic) * np.exp(c.coeff * y * y) + c.ic
    c.bcs = problem.standard_bcs(c.V, c.heating_z, c.r_sample, c.ic, gaussian)
    c.bc_dofs, c.bc_value, c.gauss_slot, c.gauss_r = problem.device_bc_arrays(len(c.nodes), c.bcs, c.bcs[3], c.nodes)
    # oracle-side BCs (independent restatement of the dof location)
    c.oracle_bcs = [
        (ho.locate_row_dofs(c.nodes, "left"), "const"),
        (ho.locate_row_dofs(c.nodes, "right"), "const"),
        (ho.locate_row_dofs(c.nodes, "top"), "const"),
        (ho.locate_row_dofs(c.nodes, "x", coord=c.heating_z, length=2 * abs(c.r_sample), center=0.0), "gauss"),
    ]
    return c


def make_oracle(c):
    return ho.Oracle2D(c.nodes, c.tris, c.rhoc_c, c.kappa_c, c.dt, c.oracle_bcs, c.ic, c.fwhm, c.heat_t, c.heat_T)


def make_solver(c, rtol=1e-14, warm=0.0, mode=0, ordering="auto", recycle=0):
    from heatflow_b200.solver import HeatSolver
    s = HeatSolver(0)
    s.set_ordering(ordering)
    s.set_mesh(c.nodes, c.tris, c.cell_tag)
    s.set_materials(c.tags, c.kappa_t, c.rhoc_t)
    s.set_bcs(c.bc_dofs, c.bc_value, c.gauss_slot, c.gauss_r)
    s.build_operator(c.dt, True)
    s.set_solver(rtol=rtol, warm=warm, mode=mode)
    s.set_recycle(recycle)
    s.set_state(np.full(len(c.nodes), c.ic))
    return s
