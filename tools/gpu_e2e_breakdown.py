"""Development aid: where the wall time of the e2e `Posterior` call of bench.py goes."""
import sys, os, time
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200.inference import ChainSampler
from bench import synthetic
counts, X, T, Y = synthetic(1024, 16, 10)
ctx = g.Context(0)
pri = g.getPriorParameters()
for rep in range(3):
    t0 = time.perf_counter()
    s = ChainSampler(pri, X, T, Y, 1, counts, 1, 10, 0, n_chains=512, seed=rep, ctx=ctx)
    ctx.synchronize(); t1 = time.perf_counter()
    s.run(1)
    ctx.synchronize(); t2 = time.perf_counter()
    out = s.samples()
    t3 = time.perf_counter()
    s.close()
    t4 = time.perf_counter()
    print(f"rep {rep}: create+generate+initial factors {1e3*(t1-t0):.1f} ms, run(1 outer = 10 sweeps) {1e3*(t2-t1):.1f} ms, samples D2H {1e3*(t3-t2):.1f} ms, close {1e3*(t4-t3):.1f} ms, total {1e3*(t4-t0):.1f} ms")
