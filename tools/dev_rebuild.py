"""Developer probe: cost of re-assembling the operator with other material values (sweep over k)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import build_case, make_solver, make_oracle
c = build_case("geballe_with_diamond", float(sys.argv[1]) if len(sys.argv) > 1 else 1.0)
s = make_solver(c, warm=1.0, recycle=128)
n = len(c.nodes)
for k in range(4):
    t0 = time.time()
    kap = c.kappa_t.copy(); kap[[m.name for m in c.mats].index("p_sample")] *= (1 + 0.25 * k)
    s.set_materials(c.tags, kap, c.rhoc_t); s.build_operator(c.dt, True)
    t1 = time.time()
    s.set_state(np.full(n, c.ic)); hist, iters, _ = s.run(c.amps, c.ic, c.coeff, [0, n // 2])
    t2 = time.time()
    print(f"rebuild {1e3*(t1-t0):.1f} ms, run wall {1e3*(t2-t1):.1f} ms (device {s.stats()['run_ms']:.1f} ms), iters {int(iters.sum())}")
# parity of the last variant against the oracle
c.kappa_c = kap[c.cell_tag - 1]
O = make_oracle(c); oh, _ = O.run(c.num_steps, [0, n // 2])
print("max rel err vs oracle", np.abs(hist / oh - 1).max())
