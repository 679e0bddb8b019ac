"""Development aid: MH sweep time at a given shape. usage: python tools/gpu_sweep_time.py n n_obj nX chains [reps]"""
import sys, os, time
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200.inference import ChainSampler
from bench import synthetic
n, n_obj, nX, C = (int(x) for x in sys.argv[1:5])
reps = int(sys.argv[5]) if len(sys.argv) > 5 else 2
ctx = g.Context(0)
counts, X, T, Y = synthetic(n, n_obj, nX)
s = ChainSampler(g.getPriorParameters(), X, T, Y, 1, counts, 24, 10, 5, n_chains=C, seed=1, ctx=ctx)
s.mh_sweeps(1); ctx.synchronize()
t = time.perf_counter(); s.mh_sweeps(reps); ctx.synchronize(); dt = (time.perf_counter() - t) / reps
fl = C * (s.n_sites - 1) * (n ** 3 / 3 + 2 * n * n)
print(f"[{os.environ.get('GPSLC_LIB_SUFFIX', '')}] n={n} nX={nX} C={C}: sweep {dt*1e3:.1f} ms = {C/dt:.1f} sweeps/s, {fl/dt/1e12:.2f} TFLOP/s ({fl/dt/35.89e12:.3f} of peak)")
