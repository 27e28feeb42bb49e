"""Heating curve for cfgs/konopkova.yaml from the reference's raw digitised trace.

`experimental_data/konopkova_pside.csv` (shipped by the reference) is headerless, two columns:
time in microseconds (0.589 .. 8.46) and the p-side temperature in units of 1000 K (1.844 .. 2.131).
The runners read a CSV with `time` [s] and `temp` [K] columns (run_with_diamond.py:254-274), so this
script converts units, sorts by time, drops duplicate time stamps and writes
`experimental_data/konopkova_heat_data.csv`.  Deterministic; committed next to its output.
"""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = np.loadtxt(os.path.join(ROOT, "experimental_data", "konopkova_pside.csv"), delimiter=",")
order = np.argsort(raw[:, 0], kind="stable")
t, T = raw[order, 0] * 1e-6, raw[order, 1] * 1e3
keep = np.concatenate(([True], np.diff(t) > 0))
with open(os.path.join(ROOT, "experimental_data", "konopkova_heat_data.csv"), "w") as f:
    f.write("time,temp\n")
    for a, b in zip(t[keep], T[keep]):
        f.write(f"{a:.9e},{b:.6f}\n")
print(f"{keep.sum()} rows, t = {t[keep][0]:.3e} .. {t[keep][-1]:.3e} s, T = {T.min():.1f} .. {T.max():.1f} K")
