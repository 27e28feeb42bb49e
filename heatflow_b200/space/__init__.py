from .space_and_forms import Space  # noqa: F401
