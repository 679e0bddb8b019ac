"""Development aid: what do the draws `MeanITE + L22 xi` cost inside gpslc_ite? The same call (c3 shape, 512 tasks, device-resident
inputs and outputs so that no copy is timed) with samplesPerPosterior = 0, 1, 10, 40."""
import sys, os, time, ctypes
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200 import estimation as ge
from gpslc_b200.inference import ChainSampler
from bench import synthetic
C = 512
counts, X, T, Y = synthetic(1024, 16, 10)
ctx = g.Context(0)
s = ChainSampler(g.getPriorParameters(), X, T, Y, 1, counts, 24, 10, 5, n_chains=C, seed=1234, ctx=ctx)
packed = s.state()[None]
ret0 = np.zeros(1, dtype=np.int32)
for spp in (0, 1, 10, 40, 0, 10):
    best = 1e9
    for rep in range(3):
        ctx.synchronize(); t = time.perf_counter()
        o = ge.ite(packed, X, T, Y, 1, 0.0, ret0, 1e-10, spp, seed=rep, ctx=ctx, want_samples=spp > 0)
        ctx.synchronize(); best = min(best, time.perf_counter() - t)
    print(f"spp {spp}: best of 3 {best*1e3:.1f} ms (host-timed call incl. copies: {C*spp*1024*8/1e6:.0f} MB of draws D2H)")
