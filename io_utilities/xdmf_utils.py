from heatflow_b200.io_utilities.xdmf_utils import XDMFFile, init_xdmf, save_params  # noqa: F401
