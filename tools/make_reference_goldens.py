#!/usr/bin/env python
"""Pins the oracle (and the CUDA path) against the UNMODIFIED reference: writes tests/golden/ref_*.npz.

Needs a machine with the reference's own stack (dolfinx 0.7-0.9 + petsc4py with MUMPS + gmsh)
and a checkout of cebarker1000/heatflow.  This container has neither, so the script has never been run here:
`tests/test_reference_goldens.py` skips while no `tests/golden/ref_*.npz` exists, and checks the oracle (CPU) and the
CUDA path (`-m gpu`) against every such file once one is committed.  The packing half (`pack_golden`) is exercised
by `tests/test_reference_goldens.py::test_pack_golden_roundtrip` with oracle data under a random dof permutation.

    python tools/make_reference_goldens.py --reference /path/to/heatflow \
        [--cfg geballe_with_diamond geballe_no_diamond] [--scale 16] [--out tests/golden]

How the reference is driven (nothing of it is copied or edited):
  phase "mesh"  (this repo's code)   our mesher at `--scale` -> <tmp>/mesh.msh (MSH 4.1, physical surfaces named after
                the materials, what gmsh.write leaves behind: mesh_and_materials/mesh.py:191-195) + mesh_cfg.yaml
                with `material_tags` (run_with_diamond.py:199-216) + the mesh arrays as .npy.
  phase "run"   (reference on sys.path, this repo NOT) imports run_with_diamond / run_no_diamond from `--reference` and
                calls `run_simulation(cfg, mesh_folder, rebuild_mesh=False, output_folder=..., watcher_points=...,
                write_xdmf=True)` as with_diamond.py:40-49 does.  Three dolfinx entry points are wrapped *around* the
                call so that what the runner computes can be read back without touching its source:
                  dolfinx.fem.functionspace          -> the P1 space V (dof coordinates = the dof map)
                  dolfinx.fem.petsc.assemble_matrix  -> the assembled operator A (run_with_diamond.py:381-382)
                  dolfinx.io.XDMFFile.write_function -> u_n after every step (run_with_diamond.py:483-484)
                watcher_points.csv (and run_no_diamond's radial_gradient_raw.csv) are read from the output folder.
  pack          dolfinx renumbers the nodes on import (gmshio.model_to_mesh); dof coordinates are matched to our
                nodes exactly (distance 0 up to 1e-9 of the mesh size) and everything is stored in OUR numbering:
                CSR of A with sorted columns, fields at every step, watcher histories, the permutation itself.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = {"pside": None, "oside": (0.951e-6, 0.0), "offaxis": (0.0, 5e-6)}     # pside: half a coupler width past the heated face


def pack_golden(nodes, tris, cell_tag, dof_xy, a_indptr, a_indices, a_data, times, fields_dof, watch_names, watch_hist,
                extra=None):
    """Everything in the caller's (= this repo's) node numbering.  `dof_xy[d]` = coordinates of reference dof d;
    `a_*` = CSR of the reference operator in dof numbering; `fields_dof[k]` = u after `times[k]` in dof numbering."""
    from scipy.sparse import csr_matrix
    from scipy.spatial import cKDTree
    nodes = np.asarray(nodes, dtype=np.float64)
    n = len(nodes)
    dof_xy = np.asarray(dof_xy, dtype=np.float64)[:, :2]
    if dof_xy.shape[0] != n:
        raise RuntimeError(f"reference space has {dof_xy.shape[0]} dofs, mesh has {n} nodes")
    dist, node_of_dof = cKDTree(nodes).query(dof_xy)
    scale = float(np.abs(nodes).max())
    if dist.max() > 1e-9 * scale or len(np.unique(node_of_dof)) != n:
        raise RuntimeError(f"dof coordinates do not match the mesh nodes one to one (max distance {dist.max():.3e})")
    A = csr_matrix((np.asarray(a_data, dtype=np.float64), np.asarray(a_indices), np.asarray(a_indptr)), shape=(n, n))
    P = csr_matrix((np.ones(n), (node_of_dof, np.arange(n))), shape=(n, n))          # ours <- dof
    A_ours = (P @ A @ P.T).tocsr()
    # P A P^T keeps explicitly stored zeros of A (BC rows/columns) only if scipy does not prune them: rebuild the
    # pattern from the index arrays instead of trusting the product
    pat = csr_matrix((np.ones(len(a_indices)), np.asarray(a_indices), np.asarray(a_indptr)), shape=(n, n))
    pat = (P @ pat @ P.T).tocsr()
    pat.sort_indices()
    A_ours.sort_indices()
    rows = np.repeat(np.arange(n), np.diff(pat.indptr))
    val = np.asarray(A_ours[rows, pat.indices]).ravel()
    fields = np.empty((len(fields_dof), n))
    for k, f in enumerate(fields_dof):
        fields[k, node_of_dof] = np.asarray(f, dtype=np.float64)
    out = dict(nodes=nodes, tris=np.asarray(tris, dtype=np.int32), cell_tag=np.asarray(cell_tag, dtype=np.int32),
               node_of_dof=node_of_dof.astype(np.int32), rowptr=pat.indptr.astype(np.int32), col=pat.indices.astype(np.int32),
               val=val, times=np.asarray(times, dtype=np.float64), fields=fields,
               watch_names=np.array(list(watch_names)), watch_hist=np.asarray(watch_hist, dtype=np.float64))
    out.update(extra or {})
    return out


# --------------------------------------------------------------------------------------------------------------
def phase_mesh(args):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import yaml
    from helpers import build_case
    from heatflow_b200.mesh_and_materials.msh_io import write_msh
    c = build_case(args.cfg[0], args.scale)
    os.makedirs(args.work, exist_ok=True)
    write_msh(os.path.join(args.work, "mesh.msh"), c.nodes, c.tris, c.cell_tag, {int(m.tag): m.name for m in c.mats})
    mesh_cfg = dict(c.cfg)
    mesh_cfg["material_tags"] = {m.name: int(m.tag) for m in c.mats}
    with open(os.path.join(args.work, "mesh_cfg.yaml"), "w") as f:
        yaml.safe_dump(mesh_cfg, f)
    with open(os.path.join(args.work, "cfg.yaml"), "w") as f:
        yaml.safe_dump(c.cfg, f)                                  # heating file as an absolute path
    np.save(os.path.join(args.work, "nodes.npy"), c.nodes)
    np.save(os.path.join(args.work, "tris.npy"), c.tris)
    np.save(os.path.join(args.work, "cell_tag.npy"), c.cell_tag)
    watch = dict(WATCH)
    watch["pside"] = (float(c.heating_z + 0.5 * 6.2e-8), 0.0)
    with open(os.path.join(args.work, "watch.json"), "w") as f:
        json.dump(watch, f)


def phase_run(args):
    """Runs inside a process that has ONLY the reference on sys.path (the repo root holds shims of the same names)."""
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ROOT]
    sys.path.insert(0, os.path.abspath(args.reference))
    os.chdir(os.path.abspath(args.reference))                      # the runners resolve cfg-relative paths from here
    import yaml
    import dolfinx
    import dolfinx.fem
    import dolfinx.fem.petsc
    import dolfinx.io
    cap = {"V": None, "A": None, "fields": [], "times": []}

    real_fs = dolfinx.fem.functionspace
    def functionspace(mesh, element, *a, **k):
        V = real_fs(mesh, element, *a, **k)
        fam = element[0] if isinstance(element, (tuple, list)) else None
        if cap["V"] is None and fam in ("Lagrange", "CG", "P") and V.dofmap.index_map_bs == 1 and V.dofmap.bs == 1:
            cap["V"] = V
        return V
    dolfinx.fem.functionspace = functionspace

    real_am = dolfinx.fem.petsc.assemble_matrix
    def assemble_matrix(*a, **k):
        A = real_am(*a, **k)
        if cap["A"] is None:
            cap["A"] = A                                            # the runner calls A.assemble() right after
        return A
    dolfinx.fem.petsc.assemble_matrix = assemble_matrix

    real_xdmf = dolfinx.io.XDMFFile
    class RecordingXDMF(real_xdmf):
        def write_function(self, u, t=0.0, *a, **k):
            if getattr(u, "name", "") == "Temperature (K)":
                cap["fields"].append(np.array(u.x.array, dtype=np.float64, copy=True))
                cap["times"].append(float(t))
            return super().write_function(u, t, *a, **k)
    dolfinx.io.XDMFFile = RecordingXDMF

    with open(os.path.join(args.work, "cfg.yaml")) as f:
        cfg = yaml.safe_load(f)
    with open(os.path.join(args.work, "watch.json")) as f:
        watch = {k: tuple(v) for k, v in json.load(f).items()}
    module = __import__("run_with_diamond" if "p_diam" in cfg["mats"] else "run_no_diamond")
    out = os.path.join(args.work, "out")
    module.run_simulation(cfg, args.work, rebuild_mesh=False, visualize_mesh=False, output_folder=out,
                          watcher_points=watch, write_xdmf=True, suppress_print=False)
    V, A = cap["V"], cap["A"]
    if V is None or A is None or len(cap["fields"]) < 2:
        raise RuntimeError("the wrapped dolfinx entry points were not reached - has the reference changed?")
    n = V.dofmap.index_map.size_local
    indptr, indices, data = A.getValuesCSR()
    import pandas as pd
    w = pd.read_csv(os.path.join(out, "watcher_points.csv"))
    names = [c for c in w.columns if c != "time"]
    extra = {}
    raw = os.path.join(out, "radial_gradient_raw.csv")
    if os.path.isfile(raw):                                         # run_no_diamond.py:603-617
        g = pd.read_csv(raw, index_col=0)
        extra["grad_raw_z"] = np.array([float(c) for c in g.columns])
        extra["grad_raw"] = g.to_numpy(dtype=np.float64)
    import petsc4py
    versions = {"dolfinx": dolfinx.__version__, "petsc4py": petsc4py.__version__}
    try:
        import gmsh
        versions["gmsh"] = gmsh.__version__
    except Exception:
        pass
    np.savez(os.path.join(args.work, "captured.npz"), dof_xy=V.tabulate_dof_coordinates()[:n, :2], indptr=indptr,
             indices=indices, data=data, times=np.array(cap["times"]), fields=np.array(cap["fields"])[:, :n],
             watch_names=np.array(names), watch_hist=w[names].to_numpy(dtype=np.float64),
             watch_time=w["time"].to_numpy(dtype=np.float64), versions=json.dumps(versions), **extra)


def main():
    ap = argparse.ArgumentParser(description=__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("--reference", required=True, help="checkout of cebarker1000/heatflow (unmodified)")
    ap.add_argument("--cfg", nargs="+", default=["geballe_with_diamond", "geballe_no_diamond"])
    ap.add_argument("--scale", type=float, default=16.0, help="mesh size factor of tests/helpers.build_case (1 = cfg sizes)")
    ap.add_argument("--out", default=os.path.join(ROOT, "tests", "golden"))
    ap.add_argument("--phase", choices=["all", "mesh", "run"], default="all")
    ap.add_argument("--work", default=None)
    args = ap.parse_args()
    if args.phase == "mesh":
        return phase_mesh(args)
    if args.phase == "run":
        return phase_run(args)
    for name in args.cfg:
        work = tempfile.mkdtemp(prefix=f"hf_ref_{name}_")
        base = [sys.executable, os.path.abspath(__file__), "--reference", args.reference, "--scale", str(args.scale),
                "--cfg", name, "--work", work]
        subprocess.run(base + ["--phase", "mesh"], check=True)
        env = {k: v for k, v in os.environ.items() if k != "PYTHONPATH"}
        subprocess.run(base + ["--phase", "run"], check=True, env=env, cwd=os.path.abspath(args.reference))
        cap = np.load(os.path.join(work, "captured.npz"), allow_pickle=False)
        extra = {"cfg_name": np.array(name), "size_scale": np.array(args.scale), "versions": cap["versions"],
                 "watch_time": cap["watch_time"]}
        for k in ("grad_raw_z", "grad_raw"):
            if k in cap.files:
                extra[k] = cap[k]
        gold = pack_golden(np.load(os.path.join(work, "nodes.npy")), np.load(os.path.join(work, "tris.npy")),
                           np.load(os.path.join(work, "cell_tag.npy")), cap["dof_xy"], cap["indptr"], cap["indices"],
                           cap["data"], cap["times"], cap["fields"], cap["watch_names"], cap["watch_hist"], extra)
        scale = int(args.scale) if float(args.scale).is_integer() else args.scale
        path = os.path.join(args.out, f"ref_{name}_s{scale}.npz")
        np.savez_compressed(path, **gold)
        print(f"{path}: N = {len(gold['nodes'])}, {len(gold['times'])} fields, nnz = {len(gold['col'])}, {cap['versions']}")


if __name__ == "__main__":
    main()
