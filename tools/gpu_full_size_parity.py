"""One-off full-size parity run (BASELINE c3 shape: n=1024, 16 objects, nX=10, nU=1): the CUDA chain against the oracle chain driven by
the same Philox streams, one outer iteration (10 MH sweeps of 58 sites + 5 elliptical-slice passes), two chains."""
import sys, os, time
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200.inference import ChainSampler
from oracle import data as od, inference as oi
n, n_obj, nX = (int(a) for a in sys.argv[1:4]) if len(sys.argv) > 3 else (1024, 16, 10)
nOuter, nMH, nES, C = 1, 10, 5, 2
counts, X, T, Y = od.synthetic(n, n_obj, nX, seed=1234)
md = od.model_data_from_arrays(counts, X, T, Y, nU=1)
ctx = g.Context(0)
t = time.perf_counter()
s = ChainSampler(md.prior, X, T, Y, 1, counts, nOuter, nMH, nES, n_chains=C, seed=77, ctx=ctx)
s.run(nOuter); got = s.samples(); acc, ev = s.stats(); s.close()
tg = time.perf_counter() - t
lines = [f"# Full-size chain parity, n={n}, {n_obj} objects, nX={nX}: CUDA vs oracle, {nOuter} outer iteration ({nMH} MH sweeps + {nES} slice passes), {C} chains", ""]
for c in range(C):
    t = time.perf_counter()
    want, st = oi.posterior(md, nOuter, nMH, nES, seed=77, chain=c)
    to = time.perf_counter() - t
    err = np.nanmax(np.abs(want - got[:, c, :]) / (1e-9 + np.abs(want)))
    lines.append(f"* chain {c}: max relative deviation of the packed sample (58 hyperparameters + U) {err:.3e}; accepted MH updates GPU {int(acc[c].sum())}; "
                 f"slice evaluations GPU {int(ev[c])}; oracle {to:.1f} s on the host, GPU {tg:.2f} s for both chains")
    print(lines[-1])
    assert err < 1e-7, err
open(os.path.join(root, "gpurun_out", "parity_full_size_r01.md"), "w").write("\n".join(lines) + "\n")
