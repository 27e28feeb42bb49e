"""Import shim for the reference package name (`from mesh_and_materials.mesh import *`)."""
from heatflow_b200.mesh_and_materials import *  # noqa: F401,F403
