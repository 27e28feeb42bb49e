"""Developer probe (not part of the product): effect of hf_set_recycle on iterations, time and parity."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
if os.environ.get("HF_DEV_LIB"):
    from heatflow_b200 import _lib as _l
    _l.LIB_PATH = os.path.abspath(os.environ["HF_DEV_LIB"])
    import ctypes as _C                                   # older builds: drop the entry points they lack
    _raw = _C.CDLL(_l.LIB_PATH)
    for _name in [n for n in _l.SIGNATURES if not hasattr(_raw, n)]:
        del _l.SIGNATURES[_name]
    from heatflow_b200 import solver as _s
    if "hf_set_sharing" not in _l.SIGNATURES:
        _s.HeatSolver.set_sharing = lambda self, n: None
from helpers import build_case, make_solver, make_oracle

def probe(name, scale, steps, caps, mode=0, check=True):
    c = build_case(name, scale)
    steps = min(steps, c.num_steps)
    ofields = None
    if check:
        O = make_oracle(c)
        _, ofields = O.run(steps, [0], keep_fields=True)
    for cap in caps:
        s = make_solver(c, warm=1.0, mode=mode, ordering=os.environ.get("HF_ORD", "auto"))
        n, nnz = s.sizes()
        s.set_recycle(cap)
        s.run(c.amps[20:23], c.ic, c.coeff, [0])
        s.set_state(np.full(n, c.ic))
        hist, iters, fields = s.run(c.amps[:steps], c.ic, c.coeff, [0, n // 2], keep_fields=check)
        ms = s.stats()["run_ms"]
        err = max(np.abs(f / of - 1).max() for f, of in zip(fields, ofields)) if check else float("nan")
        print(f"{name} scale={scale} N={n} mode={mode} recycle={cap}: steps={steps} iters total={int(iters.sum())} "
              f"last={iters[-3:]} dev={ms:.1f} ms -> {n*steps/ms/1e3:.2f} MDOF-steps/s  max rel err vs oracle {err:.2e}", flush=True)
        s.close()

if __name__ == "__main__":
    a = sys.argv[1:]
    probe(a[0], float(a[1]), int(a[2]), [int(v) for v in a[3].split(",")], mode=int(a[4]) if len(a) > 4 else 0,
          check=(a[5] != "nocheck") if len(a) > 5 else True)
