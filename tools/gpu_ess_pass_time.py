"""Development aid: time of one elliptical-slice pass over U (all chains) at a given shape, after a few MH sweeps
(usage: [n n_obj nX chains])."""
import os, sys, time
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200.inference import ChainSampler
from bench import synthetic, default_priors
n, n_obj, nX, C = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (256, 4, 5, 1024)
counts, X, T, Y = synthetic(n, n_obj, nX)
ctx = g.Context(0)
s = ChainSampler(default_priors(), X, T, Y, 1, counts, 24, 10, 5, n_chains=C, seed=1234, ctx=ctx)
s.mh_sweeps(2); s.ess_pass(0); ctx.synchronize()
t = time.perf_counter()
for j in range(1, 6):
    s.ess_pass(j)
ctx.synchronize(); dt = (time.perf_counter() - t) / 5
print(f"n={n} nX={nX} C={C}: slice pass {dt * 1e3:.2f} ms")
