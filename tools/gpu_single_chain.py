"""Development aid: wall time of the reference's own use case — one default `gpslc()` call (1 chain, 24 outer iterations of 10 MH
sweeps + 5 elliptical-slice passes) followed by sampleITE, on synthetic data of a given size."""
import sys, os, time
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from bench import synthetic
ctx = g.Context(0)
for n in [int(a) for a in sys.argv[1:]] or [1024]:
    counts, X, T, Y = synthetic(n, n // 64, 10)
    g.gpslc(counts, X[:, :1], T, Y, ctx=ctx, hyperparams=g.getHyperParameters().__class__(**{**g.getHyperParameters().__dict__, "nOuter": 1, "nBurnIn": 1}))  # warm-up (allocations)
    t = time.perf_counter(); gobj = g.gpslc(counts, X, T, Y, seed=3, ctx=ctx); t1 = time.perf_counter() - t
    t = time.perf_counter(); ite = g.sampleITE(gobj, 0.0, ctx=ctx); t2 = time.perf_counter() - t
    print(f"[team={os.environ.get('GPSLC_TEAM', 'auto')}] n={n}, nX=10, 1 chain: gpslc() {t1:.2f} s, sampleITE (15 x 10 draws) {t2:.3f} s, finite {bool(np.isfinite(ite).all())}")
