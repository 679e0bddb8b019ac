"""Development aid: large-n counterfactual sweep (BASELINE c5 shape, reduced doT count) — memory logic and timing."""
import sys, os, time
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200 import estimation as ge
from bench import synthetic
ctx = g.Context(0)
cases = ((1024, 16, 15), (2048, 32, 8), (4096, 64, 8), (8192, 128, 32))
if len(sys.argv) > 1:
    cases = [c for c in cases if str(c[0]) in sys.argv[1:]]
for n, n_obj, ndot in cases:
    counts, X, T, Y = synthetic(n, n_obj, 10)
    nX, nU = 10, 1
    n_params = 6 + 4 * nX + 2 * nU + nU * nX
    rng = np.random.default_rng(0)
    rec = np.ones(n_params + n)
    rec[:n_params] = 0.8 + 0.4 * rng.random(n_params)
    rec[2] = 0.3                      # yNoise
    rec[n_params:] = np.repeat(rng.standard_normal(n_obj), n // n_obj)
    doT = np.linspace(T.min(), T.max(), ndot)
    t = time.perf_counter()
    o = ge.ite(rec[None, None, :], X, T, Y, 1, doT, np.array([0], dtype=np.int32), 1e-10, 10, ctx=ctx)
    dt = time.perf_counter() - t
    fl = ndot * 8 * n ** 3 / 3
    print(f"n={n}: {ndot} doT x 1 sample x 10 draws: {dt:.2f} s ({fl/dt/1e12:.2f} TFLOP/s), info {o['info'].ravel()}, mean|ITE| {np.abs(o['mean']).mean():.3f}, finite {np.isfinite(o['samples']).all()}")
    # consistency: SATE fast path vs mean of MeanITE
    s = ge.sate(rec[None, None, :], X, T, Y, 1, doT[:2], np.array([0], dtype=np.int32), 1e-10, 2, ctx=ctx)
    print("   SATE mean vs mean(MeanITE):", s["mean"].ravel(), o["mean"][:2, 0, 0].mean(axis=1))
