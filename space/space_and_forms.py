from heatflow_b200.space.space_and_forms import Space  # noqa: F401
