"""Import shim: `import run_no_diamond_1d` keeps working as in the reference layout."""
from heatflow_b200.run_no_diamond_1d import *  # noqa: F401,F403
from heatflow_b200.run_no_diamond_1d import extract_1d_submesh_from_2d, run_1d  # noqa: F401
