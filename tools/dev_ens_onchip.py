"""Developer probe: time one sweep tile through the ensemble engines (argv: scale B [recycle] [steps])."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import build_case, make_solver
from heatflow_b200 import problem
scale, B = float(sys.argv[1]), int(sys.argv[2])
recycle = int(sys.argv[3]) if len(sys.argv) > 3 else 128
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 100
c = build_case("geballe_with_diamond", scale)
tag = int(c.tags[[m.name for m in c.mats].index("p_sample")])
n = len(c.nodes)
fws = np.logspace(-6, -4, 64)
for uniform in (True, False):
    ks = [10.0] * B if uniform else list(np.logspace(0, 2, 64)[20:20 + B])
    cf = [problem.gaussian_coeff(f) for f in fws[10:10 + B]]
    for env in ({"HF_ENS_NH": "2"}, {"HF_ENS_NH": "1"}, {"HF_ENS_STREAM": "1"}):
        for k in ("HF_ENS_NH", "HF_ENS_STREAM"):
            os.environ.pop(k, None)
        os.environ.update(env)
        s = make_solver(c, warm=1.0, ordering="hilbert", recycle=recycle)
        t0 = time.time(); s.ens_create(ks, cf, tag); t_create = time.time() - t0
        path = s.ens_path()
        s.ens_run(c.amps[:3], c.ic, [0])                      # warm-up
        t0 = time.time(); s.ens_destroy(); t_destroy = time.time() - t0
        s.set_state(np.full(n, c.ic))
        t0 = time.time(); s.ens_create(ks, cf, tag); t_create2 = time.time() - t0
        s.set_profile(True)
        t0 = time.time()
        hist, iters = s.ens_run(c.amps[:steps], c.ic, [0, n // 2])
        wall = time.time() - t0
        ms = s.stats()["run_ms"]
        solve_ms, _ = s.solve_profile()
        print(f"N={n} B={B} uniform_k={uniform} {env} path={path}: create {t_create*1e3:.1f}/{t_create2*1e3:.1f} ms destroy {t_destroy*1e3:.1f} ms; "
              f"{steps} steps dev {ms:.1f} ms wall {wall*1e3:.1f} ms, iterations {int(iters.sum())} -> {ms/B:.2f} ms/sim, "
              f"{B/(wall+t_create2+t_destroy):.1f} sims/s incl. create/destroy, {ms*1e3/max(1,iters.sum()):.2f} us/iteration (all incl.), solves {solve_ms:.1f} ms = {solve_ms*1e3/max(1,iters.sum()):.2f} us/iteration", flush=True)
        s.close()
