"""Post-processing helpers the reference's driver scripts import as ``analysis_utils``
(analysis_utils.py:6-95).  ``calculate_rmse`` is the number the drivers print after a run
(no_diamond.py:90, with_diamond.py, sweep_test.py:87); plotting is outside the hot path and is
only provided when matplotlib is importable."""
import numpy as np


def calculate_rmse(exp_time, exp_data, sim_time, sim_data):
    """RMSE between experiment and simulation at the experimental time points: the simulation is
    linearly interpolated (``np.interp``, clamped at both ends) onto ``exp_time``
    (analysis_utils.py:66-93)."""
    sim_at_exp = np.interp(np.asarray(exp_time, dtype=float), np.asarray(sim_time, dtype=float),
                           np.asarray(sim_data, dtype=float))
    return float(np.sqrt(np.mean((sim_at_exp - np.asarray(exp_data, dtype=float)) ** 2)))


def plot_temperature_curves(*args, **kwargs):
    """Plot helper of the reference drivers (analysis_utils.py:6-63); not part of the GPU path."""
    try:
        import matplotlib  # noqa: F401
    except ImportError as exc:
        raise NotImplementedError("plot_temperature_curves needs matplotlib, which is not installed here") from exc
    raise NotImplementedError("plotting is out of scope of heatflow_b200 (see DESIGN.md section 7)")
