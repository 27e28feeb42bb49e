from heatflow_b200.mesh_and_materials.materials import Material  # noqa: F401
