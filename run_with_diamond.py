"""Import shim: `import run_with_diamond` keeps working as in the reference layout."""
from heatflow_b200.run_with_diamond import *  # noqa: F401,F403
from heatflow_b200.run_with_diamond import run_simulation, suppress_output  # noqa: F401

if __name__ == '__main__':
    from heatflow_b200.run_with_diamond import _cli
    _cli(run_simulation)
