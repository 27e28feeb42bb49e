"""``run_simulation`` for the five-layer stack without diamonds (reference: run_no_diamond.py:29).

Same signature and outputs as the reference, including the per-step r-weighted L2 projection
of grad(T) and the ``radial_gradient.csv`` / ``radial_gradient_raw.csv`` files
(run_no_diamond.py:471-491, :544-566, :603-617).  ``parameter_sweep`` imports this runner.
"""
from . import problem
from .run_with_diamond import _cli
from .runners import run_2d, suppress_output  # noqa: F401


def run_simulation(cfg, mesh_folder, rebuild_mesh=False, visualize_mesh=False, output_folder=None,
                   watcher_points=None, write_xdmf=True, suppress_print=False):
    return run_2d(cfg, problem.stack_no_diamond, mesh_folder, rebuild_mesh, visualize_mesh, output_folder,
                  watcher_points, write_xdmf, suppress_print, radial_outputs=True, progress_splits=10)


if __name__ == '__main__':
    _cli(run_simulation)
