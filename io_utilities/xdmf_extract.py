from heatflow_b200.io_utilities.xdmf_extract import extract_point_timeseries_xdmf  # noqa: F401
