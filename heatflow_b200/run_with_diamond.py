"""``run_simulation`` for the stack with diamond anvils, gasket and gasket insulator.

Drop-in for the reference entry point (reference: run_with_diamond.py:27): same signature,
same files read (``mesh.msh``, ``mesh_cfg.yaml``, heating CSV) and written (``used_config.yaml``,
``watcher_points.csv``, ``output.xdmf``), same exceptions; returns ``None``.  The solve itself
runs on the GPU (see ``heatflow_b200.runners``).
"""
import argparse
import json

import yaml

from . import problem
from .runners import run_2d, suppress_output  # noqa: F401


def run_simulation(cfg, mesh_folder, rebuild_mesh=False, visualize_mesh=False, output_folder=None,
                   watcher_points=None, write_xdmf=True, suppress_print=False):
    return run_2d(cfg, problem.stack_with_diamond, mesh_folder, rebuild_mesh, visualize_mesh, output_folder,
                  watcher_points, write_xdmf, suppress_print, radial_outputs=False, progress_splits=5)


def _cli(run):
    parser = argparse.ArgumentParser(description='Heatflow simulation runner')
    parser.add_argument('--config', type=str, default='simulation_template.yaml')
    parser.add_argument('--mesh-folder', type=str, default='meshes')
    parser.add_argument('--rebuild-mesh', action='store_true')
    parser.add_argument('--visualize-mesh', action='store_true')
    parser.add_argument('--output-folder', type=str)
    # the reference declares type='dict' here, which argparse rejects; JSON is accepted instead
    parser.add_argument('--watcher-points', type=json.loads, help='JSON, e.g. {"pside": [z, r]}')
    parser.add_argument('--write-xdmf', action='store_true')
    parser.add_argument('--suppress-print', action='store_true')
    args = parser.parse_args()
    with open(args.config, 'r') as f:
        cfg = yaml.safe_load(f)
    run(cfg, args.mesh_folder, args.rebuild_mesh, args.visualize_mesh, args.output_folder, args.watcher_points,
        args.write_xdmf, args.suppress_print)


if __name__ == '__main__':
    _cli(run_simulation)
