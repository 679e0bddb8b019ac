"""Development aid: MH sweep rate at small n (usage: [n n_obj nX chains]); compares builds through GPSLC_LIB_SUFFIX."""
import os, sys, time
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200.inference import ChainSampler
from bench import synthetic, default_priors
n, n_obj, nX, C = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (256, 4, 2, 1024)
counts, X, T, Y = synthetic(n, n_obj, nX)
ctx = g.Context(0)
s = ChainSampler(default_priors(), X, T, Y, 1, counts, 24, 10, 5, n_chains=C, seed=1234, ctx=ctx)
s.mh_sweeps(3); ctx.synchronize()
t = time.perf_counter(); s.mh_sweeps(10); ctx.synchronize(); dt = (time.perf_counter() - t) / 10
S = s.n_sites
print(f"lib{os.environ.get('GPSLC_LIB_SUFFIX', '')}: n={n} nX={nX} C={C}: {C / dt:.0f} sweeps/s, {C * (S - 1) * (n ** 3 / 3 + 2 * n * n) / dt / 1e12:.2f} TFLOP/s")
st = s.state(); print("finite", bool(np.isfinite(np.nan_to_num(st)).all()))
