"""Does the uNoise rank non-uniformity that shows up at 32768 SBC trials (textbook slice rule) come from autocorrelation / burn-in of
the thinned chains or from the sampler's target? Same machinery as tests/test_gpu_sbc.py with longer burn-in and thinning, plus the
rank histogram of uNoise. usage: gpu_sbc_mixing.py TRIALS BURN THIN [binary]"""
import sys, os, time
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import scipy.stats as sst
import gpslc_b200 as g
import test_gpu_sbc as T
from oracle import model as om
ctx = g.Context(0)
T.TRIALS, T.BURN, T.THIN = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
binary = len(sys.argv) > 4 and sys.argv[4] == "binary"
t = time.perf_counter()
pv = T.run_sbc(ctx, 1, 31, binary=binary, nES=1 if binary else 2)
print(f"trials {T.TRIALS} burn {T.BURN} thin {T.THIN} binary {binary}: {time.perf_counter() - t:.0f} s")
for k, v in sorted(pv.items(), key=lambda kv: kv[1])[:6]:
    print(f"  {k}: p = {v:.3g}")
