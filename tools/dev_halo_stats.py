"""CPU-only: halo sizes of the patch decomposition under the Hilbert node order (argv: cfg scale R)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from helpers import build_case
from oracle import heat_oracle as ho

def hilbert(x, y):
    n = 65536
    d = np.zeros(len(x), dtype=np.uint64)
    x = x.astype(np.int64).copy(); y = y.astype(np.int64).copy()
    s = n // 2
    while s > 0:
        rx = (x & s) > 0; ry = (y & s) > 0
        d += np.uint64(s) * np.uint64(s) * ((3 * rx.astype(np.uint64)) ^ ry.astype(np.uint64))
        m = ~ry
        fl = m & rx
        x[fl] = n - 1 - x[fl]; y[fl] = n - 1 - y[fl]
        x[m], y[m] = y[m].copy(), x[m].copy()
        s //= 2
    return d

name, scale = sys.argv[1], float(sys.argv[2])
c = build_case(name, scale)
N = len(c.nodes)
lo = c.nodes.min(axis=0); span = (c.nodes.max(axis=0) - lo).max()
q = np.clip((c.nodes - lo) / span * 65535.0, 0, 65535).astype(np.int64)
key = hilbert(q[:, 0], q[:, 1])
order = np.lexsort((np.arange(N), key))
rank = np.empty(N, dtype=np.int64); rank[order] = np.arange(N)
rowptr, col = ho.csr_pattern(N, c.tris)
row = np.repeat(np.arange(N), np.diff(rowptr))
for label, rk in (("given", np.arange(N)), ("hilbert", rank)):
    ri, ci = rk[row], rk[col]
    for R in [int(a) for a in sys.argv[3:]] or [128, 256, 1024]:
        ch_r, ch_c = ri // R, ci // R
        out = ch_r != ch_c
        pairs = np.unique(ch_r[out] * N + ci[out])
        cnt = np.bincount(pairs // N, minlength=(N + R - 1) // R)
        print(f"{label:8s} R={R:5d}: chunks={len(cnt)} halo mean={cnt.mean():.1f} max={cnt.max()} p99={np.percentile(cnt, 99):.0f} total/N={cnt.sum() / N:.3f}")
