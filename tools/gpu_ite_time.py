"""Development aid: the ITE leg of bench.py in isolation (c3 shape, one posterior sample per chain), for `ncu --metrics gpu__time_duration.sum`."""
import sys, os, time
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200 import estimation as ge
from gpslc_b200.inference import ChainSampler
from bench import synthetic
C = int(sys.argv[1]) if len(sys.argv) > 1 else 512
counts, X, T, Y = synthetic(1024, 16, 10)
ctx = g.Context(0)
s = ChainSampler(g.getPriorParameters(), X, T, Y, 1, counts, 24, 10, 5, n_chains=C, seed=1234, ctx=ctx)
packed = s.state()[None]
ret0 = np.zeros(1, dtype=np.int32)
for rep in range(3):
    t = time.perf_counter()
    o = ge.ite(packed, X, T, Y, 1, 0.0, ret0, 1e-10, 10, seed=rep, ctx=ctx)
    dt = time.perf_counter() - t
    print(f"call {rep}: {dt*1e3:.1f} ms -> {C*10/dt:.0f} ITE samples/s, info max {o['info'].max()}")
