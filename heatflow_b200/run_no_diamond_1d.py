"""1-D (axis) runner with optional radial-heat-loss correction (reference: run_no_diamond_1d.py).

``run_1d`` keeps the reference signature and return value ``(domain_1d, cell_tags_1d, maps)``
(run_no_diamond_1d.py:166, :823).  The interval mesh is the set of 2-D mesh edges lying on
r = 0 (:30-164); the un-weighted P1 forms (:537-546), the three Dirichlet sets (:578-591) and
the per-step source ``2 kappa_node grad(t, z) / delta_r`` (:718-747) are restated here, and the
tridiagonal solve runs through the same GPU PCG kernels as the 2-D path (interval cells, nv = 2).
"""
from __future__ import annotations

import os
import time

import numpy as np
import yaml

from . import fem, problem
from .dirichlet_bc.bc import RowDirichletBC
from .io_utilities.xdmf_utils import XDMFFile
from .mesh_and_materials.materials import Material
from .mesh_and_materials.mesh import Domain, Mesh, MeshTags
from .mesh_and_materials.mesher import MeshArrays
from .runners import DEFAULT_MAX_ITERS, DEFAULT_RTOL, _NamedField, _watchers, suppress_output
from .solver import HeatSolver


def extract_1d_submesh_from_2d(domain_2d, cell_tags_2d, tolerance=1e-10):
    """Interval mesh of the 2-D edges with both vertices on |r| <= tolerance.

    Tag of a 1-D cell = tag of the lowest-index 2-D cell containing the edge (the reference
    takes ``facet_to_cells[facet][0]`` built in ascending cell order, :112-134).  1-D vertices
    and cells are ordered by z.  ``maps = (entity_map, vertex_map, geom_map)``: parent 2-D cell
    of every 1-D cell, and parent 2-D node of every 1-D vertex (twice, as geometry = vertices).
    """
    x = domain_2d.geometry.x
    tris = np.asarray(domain_2d.cells)
    on_axis = np.abs(x[:, 1]) <= tolerance
    a = tris[:, [0, 1, 2]].ravel()
    b = tris[:, [1, 2, 0]].ravel()
    cell = np.repeat(np.arange(len(tris)), 3)
    keep = on_axis[a] & on_axis[b]
    if not keep.any():
        raise ValueError("No facets found on the r=0 axis. Check tolerance or mesh.")
    lo = np.minimum(a[keep], b[keep]).astype(np.int64)
    hi = np.maximum(a[keep], b[keep]).astype(np.int64)
    key = lo * x.shape[0] + hi
    _, first = np.unique(key, return_index=True)          # first occurrence = lowest 2-D cell index
    lo, hi, parent = lo[first], hi[first], cell[keep][first]
    print(f"Found {len(lo)} facets on the r=0 axis")
    verts = np.unique(np.concatenate((lo, hi)))
    verts = verts[np.argsort(x[verts, 0], kind="stable")]
    new_id = np.full(x.shape[0], -1, dtype=np.int64)
    new_id[verts] = np.arange(len(verts))
    e0, e1 = new_id[lo], new_id[hi]
    swap = x[lo, 0] > x[hi, 0]
    e0[swap], e1[swap] = e1[swap].copy(), e0[swap].copy()
    order = np.argsort(x[verts[e0], 0], kind="stable")
    cells_1d = np.column_stack((e0, e1))[order].astype(np.int32)
    parent = parent[order]
    tags_1d = np.asarray(cell_tags_2d.values)[parent]
    z = x[verts, 0]
    print(f"Created 1D submesh with {len(verts)} vertices")
    arrays = MeshArrays(np.column_stack((z, np.zeros_like(z))), np.zeros((0, 3), np.int32), np.zeros(0, np.int32))
    domain_1d = Domain(arrays)
    domain_1d.arrays.tris = cells_1d                       # interval connectivity [E,2]
    cell_tags_1d = MeshTags(tags_1d.astype(np.asarray(cell_tags_2d.values).dtype), dim=1)
    print(f"1D mesh z-range: [{z.min():.6e}, {z.max():.6e}]")
    print(f"1D mesh has {len(cells_1d)} cells")
    print("Material tag distribution:")
    for tag, count in zip(*np.unique(tags_1d, return_counts=True)):
        print(f"  Tag {tag}: {count} cells")
    return domain_1d, cell_tags_1d, (parent.astype(np.int32), verts.astype(np.int32), verts.astype(np.int32))


def _find_gradient_file(mesh_folder_2d):
    roots = [os.path.join(mesh_folder_2d, '..', 'outputs', 'geballe_no_diamond_read_flux'),
             os.path.join(mesh_folder_2d, '..', '..', 'outputs', 'geballe_no_diamond_read_flux'),
             os.path.join(os.getcwd(), 'outputs', 'geballe_no_diamond_read_flux'),
             os.path.join(os.getcwd(), 'sim_outputs', 'geballe_no_diamond_read_flux')]
    for fname, label in (('radial_gradient.csv', 'smoothed'), ('radial_gradient_raw.csv', 'raw')):
        for root in roots:
            path = os.path.join(root, fname)
            if os.path.exists(path):
                print(f"Found {label} radial gradient file: {path}")
                return path
    return None


def run_1d(cfg, mesh_folder_2d, mesh_folder_1d=None, rebuild_mesh=False, visualize_mesh=False, output_folder=None,
           watcher_points=None, write_xdmf=True, suppress_print=False, use_radial_correction=True,
           radial_gradient_path=None, device=0):
    with suppress_output(suppress_print):
        program_start_time = time.time()
        if mesh_folder_1d is None:
            mesh_folder_1d = mesh_folder_2d
        mesh_cfg_path = os.path.join(mesh_folder_2d, 'mesh_cfg.yaml')
        mesh_file_path = os.path.join(mesh_folder_2d, 'mesh.msh')
        missing = [n for n, p in (('mesh.msh', mesh_file_path), ('mesh_cfg.yaml', mesh_cfg_path)) if not os.path.isfile(p)]
        if missing:
            raise FileNotFoundError(f"Missing required file(s) in {mesh_folder_2d}: {', '.join(missing)}")
        with open(mesh_cfg_path, 'r') as f:
            mesh_cfg = yaml.safe_load(f)
        mat_tag_map = mesh_cfg.get('material_tags', {})
        domain_2d, cell_tags_2d, _ = Mesh.msh_to_dolfinx(mesh_file_path)
        print("Loaded 2D mesh successfully")
        print(f"Radial heating correction: {'ENABLED' if use_radial_correction else 'DISABLED'} (user choice)")

        domain_1d, cell_tags_1d, maps = extract_1d_submesh_from_2d(domain_2d, cell_tags_2d)
        z_nodes = domain_1d.geometry.x[:, 0]
        cells_1d = domain_1d.cells
        if visualize_mesh:
            print(f"1D mesh nodes: {z_nodes}")

        materials_1d = []
        for mat_name, mat_tag in mat_tag_map.items():
            if mat_name in cfg['mats']:
                props = cfg['mats'][mat_name]
                m = Material(mat_name, boundaries=[0.0, 1.0, 0.0, 1.0],
                             properties={"rho_cv": float(props['rho']) * float(props['cv']), "k": float(props['k'])},
                             mesh_size=float(props['mesh']))
                m.tag = mat_tag
                materials_1d.append(m)
        heat_t, heat_T = problem.read_heating_curve(cfg['heating']['file'])

        print('Assigning material properties...')
        tag_to_k = {m.tag: m.properties["k"] for m in materials_1d}
        kappa_per_cell = np.array([tag_to_k[tag] for tag in cell_tags_1d.values])
        print('Material properties assigned.')

        # ---- radial gradient table of a previous 2-D run (:316-378) ---------------------
        grad_interp = None
        radial_grad_file = None
        if use_radial_correction:
            radial_grad_file = radial_gradient_path if radial_gradient_path is not None else _find_gradient_file(mesh_folder_2d)
            if radial_grad_file is None:
                print("Warning: Could not find radial gradient file. Disabling radial heating correction.")
                use_radial_correction = False
            else:
                import pandas as pd
                from scipy.interpolate import RegularGridInterpolator
                grad_df = pd.read_csv(radial_grad_file, index_col=0)
                grad_times = grad_df.index.values.astype(float)
                grad_z = grad_df.columns.values.astype(float)
                grad_interp = RegularGridInterpolator((grad_times, grad_z), grad_df.values, method='linear')
                print(f"Loaded gradient data: {grad_df.shape[0]} timesteps, {grad_df.shape[1]} z-positions")
        delta_r = 0.0
        if use_radial_correction:
            delta_r = 0.1e-6 if 'radial_gradient.csv' in radial_grad_file else 0.07e-6    # (:466-480)

        t_final = float(cfg['timing']['t_final'])
        num_steps = int(cfg['timing']['num_steps'])
        dt = t_final / num_steps
        ic_temp = float(cfg['heating']['ic_temp'])
        offset = heat_T[0] - ic_temp

        def heating_offset(t):
            return float(np.interp(t, heat_t, heat_T, left=heat_T[0], right=heat_T[-1])) - offset

        z_sample = float(cfg['mats']['p_sample']['z'])
        z_ins_pside = float(cfg['mats']['p_ins']['z'])
        z_coupler = float(cfg['mats']['p_coupler']['z'])
        mesh_zmin = -(z_sample / 2) - z_ins_pside - z_coupler
        heating_z = mesh_zmin + z_ins_pside

        def heating_1d(x, y, t):
            return (heating_offset(t) - ic_temp) + ic_temp

        V = fem.functionspace(domain_1d, ("Lagrange", 1))
        left_bc = RowDirichletBC(V, 'left', value=ic_temp)
        right_bc = RowDirichletBC(V, 'right', value=ic_temp)
        heating_bc = RowDirichletBC(V, 'x', coord=heating_z, value=heating_1d)
        obj_bcs = [left_bc, right_bc, heating_bc]

        solver = HeatSolver(device)
        solver.set_mesh(z_nodes, cells_1d, cell_tags_1d.values)
        solver.set_materials([m.tag for m in materials_1d], [m.properties["k"] for m in materials_1d],
                             [m.properties["rho_cv"] for m in materials_1d])
        # the heating dofs use the device Gaussian with r = 0: (amp - ic) * exp(0) + ic == heating_1d
        dofs, value, gslot, gr = problem.device_bc_arrays(len(z_nodes), obj_bcs, heating_bc, domain_1d.geometry.x)
        solver.set_bcs(dofs, value, gslot, np.zeros(len(gslot)))
        solver.build_operator(dt, axisymmetric=False)
        solver.set_solver(rtol=DEFAULT_RTOL, max_iters=DEFAULT_MAX_ITERS)
        solver.set_state(np.full(len(z_nodes), ic_temp))

        if output_folder is not None:
            save_folder = output_folder
            os.makedirs(save_folder, exist_ok=True)
            with open(os.path.join(save_folder, 'used_config.yaml'), 'w') as f:
                yaml.safe_dump(cfg, f)
        else:
            save_folder = os.path.join(os.getcwd(), 'sim_outputs', '1d_simulation')
            os.makedirs(save_folder, exist_ok=True)
        xdmf_path = os.path.join(save_folder, "output.xdmf")
        watcher_csv_path = os.path.join(save_folder, "watcher_points.csv")
        u_n = _NamedField('Temperature (K)', len(z_nodes))
        u_n.x.array[:] = ic_temp
        xdmf = None
        if write_xdmf:
            xdmf = XDMFFile(domain_1d.comm, xdmf_path, "w")
            xdmf.write_mesh(domain_1d)
            xdmf.write_function(u_n, 0.0)

        watcher_names, watcher_coords = _watchers(watcher_points)
        watcher_z = [c[0] for c in watcher_coords]
        watcher_nodes = [int(np.argmin(np.abs(z_nodes - zc))) for zc in watcher_z]
        watcher_data = {name: [] for name in watcher_names}
        watcher_time = []

        node_kappas = None
        if use_radial_correction:
            # first cell (in cell order) whose span contains the node; kappa looked up with the
            # *tag value* as an index into the per-cell array - reference quirk kept (:689-696)
            zlo = np.minimum(z_nodes[cells_1d[:, 0]], z_nodes[cells_1d[:, 1]])
            zhi = np.maximum(z_nodes[cells_1d[:, 0]], z_nodes[cells_1d[:, 1]])
            node_kappas = np.empty(len(z_nodes))
            for i, zc in enumerate(z_nodes):
                hit = np.flatnonzero((zlo <= zc) & (zc <= zhi))
                cell_idx = int(hit[0]) if hit.size else 0
                node_kappas[i] = kappa_per_cell[cell_tags_1d.values[cell_idx]] if cell_idx < len(cell_tags_1d.values) else kappa_per_cell[0]

        progress_interval = max(1, num_steps // 10)
        step_times = []
        loop_start_time = time.time()
        print('Beginning 1D simulation loop...')
        startup_time = time.time() - program_start_time
        for step in range(num_steps):
            step_start = time.time()
            t = (step + 1) * dt
            if use_radial_correction:
                t_c = np.clip(t, grad_times.min(), grad_times.max())
                z_c = np.clip(z_nodes, grad_z.min(), grad_z.max())
                grad_vals = grad_interp(np.column_stack([np.full_like(z_c, t_c), z_c]))
                clamped = z_nodes != z_c
                grad_vals[clamped] *= 0.1
                solver.set_source(2.0 * node_kappas * grad_vals / delta_r)
            solver.step(amp=heating_offset(t), t_ic=ic_temp, coeff=0.0)
            if write_xdmf or watcher_points is not None:
                state = solver.get_state()
                if write_xdmf:
                    u_n.x.array[:] = state
                    xdmf.write_function(u_n, t)
                if watcher_points is not None:
                    watcher_time.append(t)
                    for name, node in zip(watcher_names, watcher_nodes):
                        watcher_data[name].append(state[node])
            step_times.append(time.time() - step_start)
            if (step + 1) % progress_interval == 0 or (step + 1) == num_steps:
                recent = step_times[max(0, len(step_times) - progress_interval):]
                print(f"1D Simulation progress: {int((step + 1) / num_steps * 100)}% (step {step + 1}/{num_steps}) | "
                      f"Avg time/step (interval): {sum(recent) / len(recent):.4f} s")
        if write_xdmf:
            xdmf.close()
        if watcher_points is not None:
            import pandas as pd
            df = pd.DataFrame({'time': watcher_time})
            for name in watcher_names:
                df[name] = watcher_data[name]
            df.to_csv(watcher_csv_path, index=False)
        solver.close()
        total_time = time.time() - program_start_time
        print("\n--- 1D Simulation Timing Summary ---")
        print(f"Total time: {total_time:.2f} s")
        print(f"Startup time: {startup_time:.2f} s")
        print(f"Loop time: {time.time() - loop_start_time:.2f} s")
        print(f"Average time per step: {(sum(step_times) / len(step_times) if step_times else 0.0):.4f} s")
        print(f"Radial heating correction: {'ENABLED' if use_radial_correction else 'DISABLED'}")
        print("------------------------------------\n")
        return domain_1d, cell_tags_1d, maps
