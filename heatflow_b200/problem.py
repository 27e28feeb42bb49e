"""Problem definition shared by the runners: cfg -> material stack, heating curve, BC sets.

Restates the set-up blocks the reference runners carry inline:
  * geometry + materials  run_with_diamond.py:60-181, run_no_diamond.py:62-131
  * heating curve         run_with_diamond.py:254-274, :343-351
  * Dirichlet lines       run_with_diamond.py:361-374 (list order left, right, 'top', inner)
Every cfg number goes through ``float()`` because PyYAML loads dot-less mantissas such as
``5e-6`` as strings.
"""
from __future__ import annotations

import numpy as np

from .dirichlet_bc.bc import RowDirichletBC, resolve_last_wins
from .mesh_and_materials.materials import Material

WITH_DIAMOND_ORDER = ["p_diam", "p_ins", "p_coupler", "p_sample", "o_coupler", "o_ins", "o_diam", "gasket", "g_ins"]
NO_DIAMOND_ORDER = ["p_ins", "p_coupler", "p_sample", "o_coupler", "o_ins"]


def _num(cfg, mat, key):
    return float(cfg["mats"][mat][key])


def _material(cfg, name, box):
    return Material(
        name,
        boundaries=box,
        properties={"rho_cv": _num(cfg, name, "rho") * _num(cfg, name, "cv"), "k": _num(cfg, name, "k")},
        mesh_size=_num(cfg, name, "mesh"),
    )


def stack_with_diamond(cfg):
    """Nine rectangles [zmin, zmax, rmin, rmax]; tags follow list order (run_with_diamond.py:181)."""
    r_s = _num(cfg, "p_sample", "r")
    r_tot = r_s + _num(cfg, "gasket", "r") + _num(cfg, "g_ins", "r")
    z_oi, z_pi = _num(cfg, "o_ins", "z"), _num(cfg, "p_ins", "z")
    z_s, z_c, z_d = _num(cfg, "p_sample", "z"), _num(cfg, "p_coupler", "z"), _num(cfg, "p_diam", "z")
    zmin = -(z_s / 2) - z_pi - z_c - z_d
    zmax = (z_s / 2) + z_oi + z_c + z_d
    box = {}
    box["p_diam"] = [zmin, zmin + z_d, 0.0, r_tot]
    box["o_diam"] = [zmax - z_d, zmax, 0.0, r_tot]
    box["p_ins"] = [box["p_diam"][1], box["p_diam"][1] + z_pi, 0.0, 0.0 + r_s]
    box["o_ins"] = [box["o_diam"][0] - z_oi, box["o_diam"][0], 0.0, 0.0 + r_s]
    box["p_coupler"] = [box["p_ins"][1], box["p_ins"][1] + z_c, 0.0, 0.0 + r_s]
    box["o_coupler"] = [box["o_ins"][0] - z_c, box["o_ins"][0], 0.0, 0.0 + r_s]
    box["p_sample"] = [box["p_coupler"][1], box["p_coupler"][1] + z_s, 0.0, 0.0 + r_s]
    box["g_ins"] = [box["p_diam"][1], box["o_diam"][0], 0.0 + r_s, 0.0 + r_s + _num(cfg, "g_ins", "r")]
    box["gasket"] = [box["p_diam"][1], box["o_diam"][0], box["g_ins"][3], r_tot]
    mats = [_material(cfg, n, box[n]) for n in WITH_DIAMOND_ORDER]
    return mats, [zmin, zmax, 0.0, r_tot], {"r_sample": r_s}


def stack_no_diamond(cfg):
    """Five stacked layers, no diamonds/gasket (run_no_diamond.py:62-131)."""
    r_s = _num(cfg, "p_sample", "r")
    r_oi, r_c, r_pi = _num(cfg, "o_ins", "r"), _num(cfg, "p_coupler", "r"), _num(cfg, "p_ins", "r")
    z_oi, z_pi = _num(cfg, "o_ins", "z"), _num(cfg, "p_ins", "z")
    z_s, z_c = _num(cfg, "p_sample", "z"), _num(cfg, "p_coupler", "z")
    zmin = -(z_s / 2) - z_pi - z_c
    zmax = (z_s / 2) + z_oi + z_c
    rmax = r_s + r_oi
    box = {}
    box["p_ins"] = [zmin, zmin + z_pi, 0.0, 0.0 + r_pi]
    box["p_coupler"] = [box["p_ins"][1], box["p_ins"][1] + z_c, 0.0, 0.0 + r_c]
    box["p_sample"] = [box["p_coupler"][1], box["p_coupler"][1] + z_s, 0.0, 0.0 + r_s]
    box["o_coupler"] = [box["p_sample"][1], box["p_sample"][1] + z_c, 0.0, 0.0 + r_c]
    box["o_ins"] = [box["o_coupler"][1], box["o_coupler"][1] + z_oi, 0.0, 0.0 + r_oi]
    mats = [_material(cfg, n, box[n]) for n in NO_DIAMOND_ORDER]
    return mats, [zmin, zmax, 0.0, rmax], {"r_sample": r_s}


def read_heating_curve(path):
    """(time, temp) arrays of the experimental heating CSV, sorted and cleaned as in the reference."""
    import pandas as pd

    raw = pd.read_csv(path)
    for column in ("temp", "time"):
        if column not in raw.columns:
            raise ValueError(f"Heating CSV file {path} must contain a '{column}' column")
    df = raw.sort_values("time")
    df = df.assign(time=pd.to_numeric(raw["time"], errors="coerce"), temp=pd.to_numeric(raw["temp"], errors="coerce"))
    df = df.dropna(subset=["time", "temp"]).reset_index(drop=True)
    return df["time"].to_numpy(dtype=np.float64), df["temp"].to_numpy(dtype=np.float64)


def heating_amplitudes(step_times, curve_t, curve_T, ic_temp):
    """heating_offset(t) at every step time: clamped linear interpolation of the curve, shifted to
    start from ``ic_temp`` (run_with_diamond.py:343-351)."""
    shift = curve_T[0] - ic_temp
    return np.interp(np.asarray(step_times, dtype=np.float64), curve_t, curve_T, left=curve_T[0], right=curve_T[-1]) - shift


def gaussian_coeff(fwhm):
    return -4.0 * np.log(2.0) / float(fwhm) ** 2


def standard_bcs(V, heating_z, r_sample, ic_temp, gaussian):
    """[left, right, 'top', inner heating line] (run_with_diamond.py:361-374)."""
    left = RowDirichletBC(V, "left", value=ic_temp)
    right = RowDirichletBC(V, "right", value=ic_temp)
    outer = RowDirichletBC(V, "top", value=ic_temp)
    inner = RowDirichletBC(V, "x", coord=heating_z, length=abs(r_sample) * 2, center=0.0, value=gaussian)
    return [left, right, outer, inner]


def device_bc_arrays(num_dofs, bcs, gaussian_bc, coords):
    """Flatten a reference-ordered BC list for ``hf_set_bcs``: sorted unique dofs with last-wins
    owners, their constant values, and the slots / radii of the dofs owned by ``gaussian_bc``."""
    owner = resolve_last_wins(num_dofs, bcs)
    dofs = np.flatnonzero(owner >= 0).astype(np.int32)
    value = np.zeros(len(dofs))
    gauss_slot = []
    for slot, d in enumerate(dofs):
        bc = bcs[owner[d]]
        if bc is gaussian_bc:
            gauss_slot.append(slot)
        elif bc.is_constant:
            value[slot] = bc.constant_value
        else:
            raise ValueError("only constant and Gaussian Dirichlet values can be evaluated on the device")
    gauss_slot = np.array(gauss_slot, dtype=np.int32)
    gauss_r = coords[dofs[gauss_slot], 1].astype(np.float64) if len(gauss_slot) else np.zeros(0)
    if len(gauss_slot):   # g at t = 0, as the reference's `for x in obj_bcs: x.update(0.0)` leaves it
        at0 = dict(zip(gaussian_bc.row_dofs.tolist(), gaussian_bc.values(0.0)))
        value[gauss_slot] = [at0[int(d)] for d in dofs[gauss_slot]]
    return dofs, value, gauss_slot, gauss_r
