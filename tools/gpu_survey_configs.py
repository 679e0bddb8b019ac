"""Development aid: timing of the BASELINE.json configurations other than the bench workload."""
import sys, os, time, json
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200.inference import ChainSampler
from gpslc_b200 import estimation as ge
sys.path.insert(0, root)
from bench import synthetic

ctx = g.Context(0)
pri = g.getPriorParameters()
out = {}
def sweep_time(name, n, n_obj, nX, C, reps=2):
    counts, X, T, Y = synthetic(n, n_obj, nX)
    s = ChainSampler(pri, X, T, Y, 1, counts, 24, 10, 5, n_chains=C, seed=1, ctx=ctx)
    s.mh_sweeps(1); ctx.synchronize()
    t = time.perf_counter(); s.mh_sweeps(reps); ctx.synchronize(); dt = (time.perf_counter() - t) / reps
    S = s.n_sites
    fl = C * (S - 1) * (n ** 3 / 3 + 2 * n * n)
    t2 = time.perf_counter(); s.ess_pass(0); ctx.synchronize(); de = time.perf_counter() - t2
    acc, ev = s.stats()
    print(f"{name}: n={n} nX={nX} C={C}: sweep {dt*1e3:.1f} ms = {C/dt:.1f} sweeps/s, {fl/dt/1e12:.2f} TFLOP/s; one ESS pass {de*1e3:.1f} ms, mean evals {ev.mean():.2f} max {ev.max()}")
    out[name] = dict(n=n, nX=nX, chains=C, sweep_ms=dt*1e3, sweeps_per_s=C/dt, tflops=fl/dt/1e12, ess_pass_ms=de*1e3, ess_evals_mean=float(ev.mean()), ess_evals_max=int(ev.max()))
    return s, (counts, X, T, Y)

s, _ = sweep_time("c2", 256, 4, 5, 1024); s.close()
s, _ = sweep_time("c1shape", 150, 6, 0, 2048); s.close()
s, d = sweep_time("c3", 1024, 16, 10, 512)
# full outer iteration + SATE at c3
t = time.perf_counter(); s.run(1); ctx.synchronize(); dt = time.perf_counter() - t
print(f"c3: one full outer iteration (10 sweeps + 5 ESS passes) {dt:.2f} s")
out["c3"]["outer_iteration_s"] = dt
smp = s.samples()
counts, X, T, Y = d
t = time.perf_counter(); so = ge.sate(smp, X, T, Y, 1, [0.0], np.array([0], dtype=np.int32), 1e-10, 10, ctx=ctx); dt = time.perf_counter() - t
print(f"c3: sampleSATE over 512 chains x 1 retained sample: {dt:.3f} s -> {512*10/dt:.0f} SATE samples/s; info max {so['info'].max()}")
out["c3"]["sate_512x1_s"] = dt
t = time.perf_counter(); io = ge.ite(smp[:, :64], X, T, Y, 1, [0.0], np.array([0], dtype=np.int32), 1e-10, 10, ctx=ctx); dt = time.perf_counter() - t
print(f"c3: sampleITE over 64 chains x 1 retained sample x 10 draws: {dt:.3f} s -> {64*10/dt:.0f} ITE samples/s ({64*8*1024**3/3/dt/1e12:.2f} TFLOP/s); info max {io['info'].max()}")
out["c3"]["ite_64x1_s"] = dt
s.close()
s, _ = sweep_time("c4", 4096, 64, 10, 64, reps=1); s.close()
json.dump(out, open(os.path.join(root, "gpurun_out", "configs_r01c.json"), "w"), indent=1)
