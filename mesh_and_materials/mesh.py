from heatflow_b200.mesh_and_materials.mesh import *  # noqa: F401,F403
from heatflow_b200.mesh_and_materials.mesh import COMM, SCALE, Mesh  # noqa: F401
