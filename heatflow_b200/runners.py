"""Shared body of the 2-D simulation runners (with / without diamond anvils).

Follows the structure of the reference runners step for step (reference:
run_with_diamond.py:27-528, run_no_diamond.py:29-630) - geometry, mesh (re)build or load,
heating curve, material fields, BCs, operator, outputs, time loop, CSVs, timing summary - but
every numerical stage runs on the GPU through ``HeatSolver`` (C-ABI, sm_100a kernels):

    reference (dolfinx/PETSc, per step)               here
    ------------------------------------------------  --------------------------------------
    inner_bc.update(t)            (Python per dof)    device Gaussian kernel, amplitude per step
    assemble_vector/apply_lifting/set_bc              fused RHS + initial-residual kernel
    KSP PREONLY + LU (MUMPS)                          Jacobi-PCG (fused SpMV/dot/axpy kernels)
    u_n.x.array[node] watchers                        device gather into a [steps, n_watch] buffer
    L2 gradient projection + LU  (no-diamond only)    r-weighted mass PCG on the same kernels
"""
from __future__ import annotations

import contextlib
import copy
import os
import time

import numpy as np
import yaml
from scipy.spatial import cKDTree

from . import fem, problem
from .io_utilities.xdmf_utils import XDMFFile
from .mesh_and_materials.mesh import COMM, Mesh
from .solver import HeatSolver

# solver defaults: rtol on the Jacobi-scaled residual relative to ||b_free|| (1e-13
# already gives ~2e-11 agreement with the sparse-LU path; 1e-14 leaves two orders of margin)
DEFAULT_RTOL = 1e-14
DEFAULT_MAX_ITERS = 50000
# initial guess of every solve: u_n + DEFAULT_WARM * (u_n - u_{n-1}); the stopping test is unchanged, the
# extrapolated start saves ~15 % of the PCG iterations on the reference configs
DEFAULT_WARM = 1.0
# the corrections of the last DEFAULT_RECYCLE solves (A-orthogonalised) seed the next solve by Galerkin
# projection (hf_set_recycle): same solver, same tolerance, ~5 x fewer PCG iterations over a 100-step run
DEFAULT_RECYCLE = 128
RECYCLE_BYTES = 8 << 30       # cap on the memory of the recycled basis (vectors * N doubles; ensembles keep two arrays)


def recycle_vectors(num_dofs, want=DEFAULT_RECYCLE):
    return int(max(0, min(want, RECYCLE_BYTES // (16 * max(1, num_dofs)))))


@contextlib.contextmanager
def suppress_output(enabled):
    if not enabled:
        yield
    else:
        with open(os.devnull, 'w') as fnull:
            with contextlib.redirect_stdout(fnull), contextlib.redirect_stderr(fnull):
                yield


class _NamedField:
    """What XDMFFile.write_function needs from a dolfinx Function."""

    def __init__(self, name, n):
        self.name = name
        self.x = fem._Vector(n)


def _watchers(watcher_points):
    if watcher_points is None:
        return [], []
    if isinstance(watcher_points, dict):
        return list(watcher_points.keys()), list(watcher_points.values())
    if isinstance(watcher_points, list):
        return [pt['name'] for pt in watcher_points], [pt['coords'] for pt in watcher_points]
    raise ValueError("watcher_points must be a dict or list of dicts")


def prepare_mesh(cfg, stack, mesh_folder, rebuild_mesh):
    """Build+write or check+load ``mesh.msh`` / ``mesh_cfg.yaml`` (run_with_diamond.py:183-245).
    Returns (materials, info, domain, cell_tags, mat_tag_map)."""
    materials, bounds, info = stack(cfg)
    gmsh_domain = Mesh(name='mesh.msh', boundaries=bounds, materials=materials)
    mesh_cfg_path = os.path.join(mesh_folder, 'mesh_cfg.yaml')
    mesh_file_path = os.path.join(mesh_folder, 'mesh.msh')
    if rebuild_mesh:
        gmsh_domain.build_mesh()
        mat_tag_map = {mat.name: (getattr(mat, '_tag', None) if getattr(mat, '_tag', None) is not None else -1)
                       for mat in materials}
        os.makedirs(mesh_folder, exist_ok=True)
        mesh_cfg = copy.deepcopy(cfg)
        mesh_cfg['material_tags'] = mat_tag_map
        with open(mesh_cfg_path, 'w') as f:
            yaml.safe_dump(mesh_cfg, f)
        gmsh_domain.write(mesh_file_path)
    else:
        missing = [n for n, p in (('mesh.msh', mesh_file_path), ('mesh_cfg.yaml', mesh_cfg_path)) if not os.path.isfile(p)]
        if missing:
            raise FileNotFoundError(f"Missing required file(s) in {mesh_folder}: {', '.join(missing)}")
        with open(mesh_cfg_path, 'r') as f:
            mesh_cfg = yaml.safe_load(f)
        mat_tag_map = mesh_cfg.get('material_tags', {})
    domain, cell_tags, _ = Mesh.msh_to_dolfinx(mesh_file_path, comm=COMM)
    return materials, info, domain, cell_tags, mesh_cfg['material_tags']


def configure_solver(domain, cell_tags, materials, mat_tag_map, bcs, gaussian_bc, dt, device=0,
                     rtol=DEFAULT_RTOL, max_iters=DEFAULT_MAX_ITERS, ic_temp=None, warm=DEFAULT_WARM,
                     recycle=DEFAULT_RECYCLE, sharing=1):
    """HeatSolver with mesh, DG0 tables, Dirichlet sets and the assembled operator."""
    solver = HeatSolver(device)
    solver.set_sharing(sharing)
    nodes = domain.geometry.x[:, :2]
    solver.set_mesh(nodes, domain.cells, cell_tags.values)
    tags = [mat_tag_map[m.name] for m in materials]
    solver.set_materials(tags, [m.properties["k"] for m in materials], [m.properties["rho_cv"] for m in materials])
    dofs, value, gslot, gr = problem.device_bc_arrays(nodes.shape[0], bcs, gaussian_bc, nodes)
    solver.set_bcs(dofs, value, gslot, gr)
    solver.build_operator(dt, axisymmetric=True)
    solver.set_solver(rtol=rtol, max_iters=max_iters, warm=warm)
    solver.set_recycle(recycle_vectors(nodes.shape[0], recycle))
    if ic_temp is not None:
        solver.set_state(np.full(nodes.shape[0], float(ic_temp)))
    return solver


class Simulation2D:
    """cfg + mesh folder -> configured ``HeatSolver`` plus everything the time loop needs
    (set-up half of the reference runners, run_with_diamond.py:183-394)."""

    def __init__(self, cfg, stack, mesh_folder, rebuild_mesh=False, visualize_mesh=False, device=0,
                 rtol=DEFAULT_RTOL, max_iters=DEFAULT_MAX_ITERS, sharing=1):
        self.cfg = cfg
        self.materials, self.info, self.domain, self.cell_tags, self.mat_tag_map = prepare_mesh(
            cfg, stack, mesh_folder, rebuild_mesh)
        if visualize_mesh:
            print("visualize_mesh: the gmsh GUI is not available in this build - skipped")
        materials, domain, cell_tags, mat_tag_map = self.materials, self.domain, self.cell_tags, self.mat_tag_map
        r_sample = self.info["r_sample"]
        p_coupler = next(m for m in materials if m.name == "p_coupler")

        self.heat_t, self.heat_T = heat_t, heat_T = problem.read_heating_curve(cfg['heating']['file'])

        self.V = V = fem.functionspace(domain, ("Lagrange", 1))
        print('Assigning material properties...')
        unknown = set(np.unique(cell_tags.values)) - {mat_tag_map[m.name] for m in materials}
        if unknown:
            raise KeyError(f"cell tags {sorted(unknown)} have no material in mesh_cfg.yaml")
        print('Material properties assigned.')

        t_final = float(cfg['timing']['t_final'])
        self.num_steps = num_steps = int(cfg['timing']['num_steps'])
        self.dt = dt = t_final / num_steps
        self.ic_temp = ic_temp = float(cfg['heating']['ic_temp'])
        self.coeff = coeff = problem.gaussian_coeff(float(cfg['heating']['fwhm']))
        offset = heat_T[0] - ic_temp

        def heating_offset(t):
            return float(np.interp(t, heat_t, heat_T, left=heat_T[0], right=heat_T[-1])) - offset

        def gaussian(x, y, t):
            return (heating_offset(t) - ic_temp) * np.exp(coeff * (y - 0.0) ** 2) + ic_temp

        self.obj_bcs = problem.standard_bcs(V, p_coupler.boundaries[0], r_sample, ic_temp, gaussian)
        self.inner_bc = self.obj_bcs[3]
        self.solver = configure_solver(domain, cell_tags, materials, mat_tag_map, self.obj_bcs, self.inner_bc, dt,
                                       device=device, rtol=rtol, max_iters=max_iters, ic_temp=ic_temp, sharing=sharing)
        self.n_dofs = domain.geometry.x.shape[0]
        self.mesh_coords = domain.geometry.x[:, :2]
        self.step_t = (np.arange(num_steps) + 1) * dt
        self.amps = problem.heating_amplitudes(self.step_t, heat_t, heat_T, ic_temp)

    def watcher_nodes(self, watcher_coords):
        """Nearest mesh node of every watcher point (run_with_diamond.py:443-449)."""
        if not len(watcher_coords):
            return []
        tree = cKDTree(self.mesh_coords)
        return [int(tree.query(coords)[1]) for coords in watcher_coords]

    def sample_tag(self, name="p_sample"):
        return int(self.mat_tag_map[name])

    def set_sharing(self, n_concurrent):
        """Re-plan the on-chip kernel for ``n_concurrent`` simulations sharing the GPU (sweep engine)."""
        self.solver.set_sharing(n_concurrent)
        self.solver.build_operator(self.dt, axisymmetric=True)

    def set_conductivity(self, name, k):
        """Re-assemble the operator with material ``name`` at conductivity ``k`` (sweep variants)."""
        tags = [self.mat_tag_map[m.name] for m in self.materials]
        kappa = [float(k) if m.name == name else m.properties["k"] for m in self.materials]
        self.solver.set_materials(tags, kappa, [m.properties["rho_cv"] for m in self.materials])
        self.solver.build_operator(self.dt, axisymmetric=True)

    def close(self):
        self.solver.close()


def run_2d(cfg, stack, mesh_folder, rebuild_mesh=False, visualize_mesh=False, output_folder=None,
           watcher_points=None, write_xdmf=True, suppress_print=False, radial_outputs=False,
           progress_splits=5, device=0):
    with suppress_output(suppress_print):
        program_start_time = time.time()
        sim = Simulation2D(cfg, stack, mesh_folder, rebuild_mesh, visualize_mesh, device)
        domain, solver = sim.domain, sim.solver
        num_steps, dt, ic_temp, coeff = sim.num_steps, sim.dt, sim.ic_temp, sim.coeff
        n_dofs = sim.n_dofs

        if output_folder is not None:
            save_folder = output_folder
            os.makedirs(save_folder, exist_ok=True)
            with open(os.path.join(save_folder, 'used_config.yaml'), 'w') as f:
                yaml.safe_dump(cfg, f)
        else:
            save_folder = os.path.join(os.getcwd(), 'sim_outputs', 'refactor_test')
            os.makedirs(save_folder, exist_ok=True)
        xdmf_path = os.path.join(save_folder, "output.xdmf")
        watcher_csv_path = os.path.join(save_folder, "watcher_points.csv")

        u_n = _NamedField('Temperature (K)', n_dofs)
        u_n.x.array[:] = ic_temp
        xdmf = None
        if write_xdmf:
            xdmf = XDMFFile(domain.comm, xdmf_path, "w")
            xdmf.write_mesh(domain)
            xdmf.write_function(u_n, 0.0)

        mesh_coords = sim.mesh_coords
        watcher_names, watcher_coords = _watchers(watcher_points)
        watcher_nodes = sim.watcher_nodes(watcher_coords) if watcher_points is not None else []
        watcher_data = {name: [] for name in watcher_names}
        watcher_time = []

        grad = None
        if radial_outputs:
            grad = _RadialGradientSampler(mesh_coords)

        progress_interval = max(1, num_steps // progress_splits)
        step_times = []
        loop_start_time = time.time()
        print('Beginning loop...')
        startup_time = time.time() - program_start_time
        step_t, amps = sim.step_t, sim.amps
        # whole progress intervals run on the device without host round trips.  XDMF fields come back through
        # hf_run's pinned field buffer (asynchronous copies that overlap the next steps) and are written per
        # interval; only the gradient CSV rows need a projection solve and a host round trip after every step.
        max_chunk = max(1, (1 << 27) // max(1, n_dofs))          # <= 1 GB of fields on the host at a time
        step = 0
        while step < num_steps:
            chunk = 1 if radial_outputs else min(progress_interval - (step % progress_interval), num_steps - step, max_chunk)
            t0 = time.time()
            hist, iters, fields = solver.run(amps[step:step + chunk], ic_temp, coeff, watcher_nodes, keep_fields=write_xdmf)
            if radial_outputs:
                grad.record(step_t[step], solver.project_gradient())
            if write_xdmf:
                for k in range(chunk):
                    u_n.x.array[:] = fields[k]
                    xdmf.write_function(u_n, step_t[step + k])
            elapsed = time.time() - t0
            for k in range(chunk):
                if watcher_points is not None:
                    watcher_time.append(step_t[step + k])
                    for name, val in zip(watcher_names, hist[k]):
                        watcher_data[name].append(val)
                step_times.append(elapsed / chunk)
            step += chunk
            if step % progress_interval == 0 or step == num_steps:
                percent = int(step / num_steps * 100)
                recent = step_times[max(0, len(step_times) - progress_interval):]
                print(f"Simulation progress: {percent}% (step {step}/{num_steps}) | "
                      f"Avg time/step (interval): {sum(recent) / len(recent):.4f} s")

        if write_xdmf:
            xdmf.close()
        import pandas as pd
        if watcher_points is not None:
            df = pd.DataFrame({'time': watcher_time})
            for name in watcher_names:
                df[name] = watcher_data[name]
            df.to_csv(watcher_csv_path, index=False)
        if radial_outputs:
            grad.write(save_folder)
        solver.close()

        total_time = time.time() - program_start_time
        loop_time = time.time() - loop_start_time
        avg_step_time = sum(step_times) / len(step_times) if step_times else 0.0
        print("\n--- Timing Summary ---")
        print(f"Total time: {total_time:.2f} s")
        print(f"Startup time: {startup_time:.2f} s")
        print(f"Loop time: {loop_time:.2f} s")
        print(f"Average time per step: {avg_step_time:.4f} s")
        print("----------------------\n")


class _RadialGradientSampler:
    """Band-averaged and raw r=0 samples of the projected radial gradient, written as
    ``radial_gradient.csv`` / ``radial_gradient_raw.csv`` (run_no_diamond.py:457-465, :494-513,
    :553-566, :603-617)."""

    def __init__(self, mesh_coords, dz_bin=0.2e-6, band=0.25e-6, r_tol=1e-12):
        z_min, z_max = mesh_coords[:, 0].min(), mesh_coords[:, 0].max()
        edges = np.arange(z_min, z_max + dz_bin, dz_bin)
        in_band = np.flatnonzero((mesh_coords[:, 1] > 0.0) & (mesh_coords[:, 1] <= band))
        which = np.searchsorted(edges, mesh_coords[in_band, 0]) - 1
        ok = (which >= 0) & (which < len(edges) - 1)
        self.z_centres, self.groups = [], []
        for k in np.unique(which[ok]):
            self.z_centres.append(0.5 * (edges[k] + edges[k + 1]))
            self.groups.append(in_band[ok][which[ok] == k])
        axis = np.flatnonzero(np.abs(mesh_coords[:, 1]) <= r_tol)
        order = np.argsort(mesh_coords[axis, 0])
        self.axis_nodes = axis[order]
        self.axis_z = mesh_coords[self.axis_nodes, 0]
        print(f"Found {len(self.axis_nodes)} nodes exactly on r=0 axis")
        self.rows, self.raw_rows, self.times = [], [], []

    def record(self, t, grad):
        gr = grad[:, 1]
        self.rows.append([float(np.mean(gr[g])) for g in self.groups])
        self.raw_rows.append([float(v) for v in gr[self.axis_nodes]])
        self.times.append(t)

    def write(self, folder):
        import pandas as pd
        if self.rows:
            df = pd.DataFrame(self.rows, columns=self.z_centres)
            df.index = self.times
            df.index.name = 'time'
            df.to_csv(os.path.join(folder, "radial_gradient.csv"))
        if self.raw_rows:
            df = pd.DataFrame(self.raw_rows, columns=self.axis_z)
            df.index = self.times
            df.index.name = 'time'
            path = os.path.join(folder, "radial_gradient_raw.csv")
            df.to_csv(path)
            print(f"Saved raw gradient data at r=0 nodes to {path}")
