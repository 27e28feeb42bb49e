"""Developer timing probe (not part of the product): full-size configs on the GPU."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import build_case, make_solver

def probe(name, scale, warm=0.0, mode=0, steps=None):
    c = build_case(name, scale)
    t0 = time.time(); s = make_solver(c, warm=warm, mode=mode); setup = time.time() - t0
    n, nnz = s.sizes()
    amps = c.amps if steps is None else c.amps[:steps]
    s.run(amps[:3], c.ic, c.coeff, [0])            # warm-up
    s.set_state(np.full(n, c.ic))
    t0 = time.time(); hist, iters, _ = s.run(amps, c.ic, c.coeff, [0, n // 2]); dt = time.time() - t0
    ms_f = s.bench_kernels(20, True); ms_w = s.bench_kernels(20, False)
    tot = int(iters.sum())
    print(f"{name} scale={scale} warm={warm} mode={mode}: N={n} nnz={nnz} setup={setup:.2f}s steps={len(amps)} "
          f"iters total={tot} max={iters.max()} time={dt:.3f}s -> {n*len(amps)/dt/1e6:.2f} MDOF-steps/s, "
          f"{dt/max(tot,1)*1e6:.2f} us/iter | kernels flushed spmv={ms_f[0]*1e3:.1f}us upd={ms_f[1]*1e3:.1f}us "
          f"warm spmv={ms_w[0]*1e3:.1f}us upd={ms_w[1]*1e3:.1f}us", flush=True)
    bytes_iter = 10 * nnz + 4 * n / 32 + 64 * n
    print(f"   k_pcg_iter algorithmic GB/s: flushed {bytes_iter/ms_f[0]/1e6:.0f}  warm {bytes_iter/ms_w[0]/1e6:.0f}", flush=True)
    s.close()

if __name__ == "__main__":
    args = sys.argv[1:]
    if args:
        probe(args[0], float(args[1]), mode=int(args[2]), steps=int(args[3]) if len(args) > 3 else None)
    else:
        for mode in (1, 2):
            probe("geballe_with_diamond", 1.0, mode=mode)
        probe("geballe_with_diamond", 0.35, mode=0, steps=10)
