"""Regenerates tests/golden/*.npz from the oracle (run from the repo root: python tests/golden/make_golden.py).

The reference itself cannot be imported here (dolfinx / petsc4py / gmsh are not installed and
the reference ships no meshes, outputs or golden numbers - SURVEY.md section 8c), so these fixtures
pin the *oracle* (and through it the GPU path) against regressions; the oracle in turn is
pinned by the closed-form / quadrature known-answer tests in tests/test_oracle.py.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import build_case, make_oracle  # noqa: E402
from oracle import heat_oracle as ho  # noqa: E402


def make(cfg_name, scale, out):
    c = build_case(cfg_name, scale)
    O = make_oracle(c)
    watch = ho.nearest_nodes(c.nodes, [(c.heating_z + 0.5 * 6.2e-8, 0.0), (0.951e-6, 0.0), (0.0, 5e-6)])
    hist, fields = O.run(c.num_steps, watch, keep_fields=True)
    keep = [0, c.num_steps // 4, c.num_steps // 2, c.num_steps - 1]
    np.savez_compressed(
        out, nodes=c.nodes, tris=c.tris, cell_tag=c.cell_tag, bc_dofs=c.bc_dofs, gauss_slot=c.gauss_slot,
        watch=watch, hist=hist, field_steps=np.array(keep), fields=np.array([fields[k] for k in keep]),
        rowptr=O.rowptr, col=O.col, diag_A=O.A.diagonal(), amps=c.amps, dt=c.dt, size_scale=scale)
    print(out, "N =", len(c.nodes), "steps =", c.num_steps, "hist[-1] =", hist[-1])


if __name__ == "__main__":
    here = os.path.dirname(os.path.abspath(__file__))
    make("geballe_no_diamond", 16.0, os.path.join(here, "no_diamond_s16.npz"))
    make("geballe_with_diamond", 16.0, os.path.join(here, "with_diamond_s16.npz"))
