"""Minimal P1 function-space stand-ins for the dolfinx objects the reference runners touch.

Only what the hot path reads is provided (reference: run_with_diamond.py:279-280, bc.py:38-41,
:104-112): ``functionspace(domain, ("Lagrange", 1))`` with ``.mesh`` and
``tabulate_dof_coordinates()``, ``locate_dofs_geometrical`` and a ``Function`` whose
``.x.array`` is a numpy vector.  The P1 dof map is the identity (dof i = mesh node i), which is
also what the reference assumes when it indexes ``u_n.x.array`` with geometry-node indices
(run_with_diamond.py:443-449, :489).
"""
from __future__ import annotations

import numpy as np


class FunctionSpace:
    def __init__(self, mesh, element=("Lagrange", 1)):
        family, degree = element[0], element[1]
        if (family, degree) not in (("Lagrange", 1), ("DG", 0), ("P", 1), ("CG", 1)):
            raise NotImplementedError(f"only P1 / DG0 spaces are supported, got {element}")
        self.mesh = mesh
        self.element = (family, degree)
        self.is_cellwise = family == "DG"

    @property
    def num_dofs(self):
        return self.mesh.cells.shape[0] if self.is_cellwise else self.mesh.geometry.x.shape[0]

    def tabulate_dof_coordinates(self):
        if self.is_cellwise:
            return self.mesh.geometry.x[self.mesh.cells].mean(axis=1)
        return self.mesh.geometry.x


def functionspace(mesh, element):
    return FunctionSpace(mesh, element)


class _Vector:
    def __init__(self, n):
        self.array = np.zeros(n, dtype=np.float64)

    def scatter_forward(self):
        pass


class Function:
    def __init__(self, V, name="f"):
        self.function_space = V
        self.x = _Vector(V.num_dofs)
        self.name = name


def locate_dofs_geometrical(V, marker):
    """Indices of dofs whose coordinates satisfy ``marker(x)`` with ``x`` of shape (3, N)."""
    x = V.tabulate_dof_coordinates().T
    return np.flatnonzero(np.asarray(marker(x), dtype=bool)).astype(np.int32)
