"""Import shim: `import parameter_sweep` / `python parameter_sweep.py ...` keep working as in the reference layout."""
from heatflow_b200.parameter_sweep import *  # noqa: F401,F403
from heatflow_b200.parameter_sweep import (create_parameter_grid, get_mesh_folder_for_width, get_watcher_points,  # noqa: F401
                                           initialize_worker, main, modify_config_for_parameters, run_parameter_sweep,
                                           run_single_simulation, set_single_thread)

if __name__ == '__main__':
    main()
