from .materials import Material
from .mesh import COMM, Domain, Mesh, MeshTags
from .mesher import MeshArrays, triangulate_rectangles
from .msh_io import read_msh, write_msh

__all__ = ["Material", "Mesh", "COMM", "Domain", "MeshTags", "MeshArrays",
           "triangulate_rectangles", "read_msh", "write_msh"]
