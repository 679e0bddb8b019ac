"""Simulation-based calibration at a larger scale than tests/test_gpu_sbc.py (same machinery): more trials for more power, both
elliptical-slice rules, continuous and binary treatment. Writes gpurun_out/sbc_r02.md."""
import sys, os, time
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root); sys.path.insert(0, os.path.join(root, "tests"))
import gpslc_b200 as g
import test_gpu_sbc as T
ctx = g.Context(0)
T.TRIALS = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
lines = [f"# Simulation-based calibration of the CUDA sampler, {T.TRIALS} trials (tools/gpu_sbc_large.py)", "",
         f"n={T.N}, objects {T.COUNTS}, nX={T.NX}, nU={T.NU}; every chain has its own dataset drawn from the model prior; burn-in {T.BURN}, thinning {T.THIN} outer "
         f"iterations, {T.DRAWS} draws per chain; chi-square test of rank uniformity ({T.DRAWS + 1} bins). Bonferroni threshold at alpha = 0.01 over 20 statistics: 5e-4.", ""]
for rule, seed, binary, nES in ((1, 31, False, 2), (0, 32, False, 2), (1, 33, True, 1), (0, 34, True, 1)):
    t = time.perf_counter()
    pv = T.run_sbc(ctx, rule, seed, binary=binary, nES=nES)
    dt = time.perf_counter() - t
    lines += [f"## {'binary' if binary else 'continuous'} T, nESInner={nES}, ess_rule={rule} ({'textbook likelihood-only slice test' if rule else 'Gen `elliptical_slice` as recollected (full update weight), the default'}) — {dt:.1f} s", "",
              "| statistic | p-value |", "|---|---|"]
    for k, v in sorted(pv.items(), key=lambda kv: kv[1]):
        lines.append(f"| {k[0]}{'' if k[0] in ('uNoise','tNoise','yNoise','tyLS','tScale','yScale') else '[' + str(k[1]) + (',' + str(k[2]) if k[0] == 'uxLS' else '') + ']'} | {v:.4g} |")
    lines.append("")
    print(f"binary {binary} rule {rule}: min p {min(pv.values()):.3g} ({min(pv, key=pv.get)}), {sum(v < 5e-4 for v in pv.values())} statistics below 5e-4, {dt:.1f} s")
os.makedirs(os.path.join(root, "gpurun_out"), exist_ok=True)
open(os.path.join(root, "gpurun_out", "sbc_r02.md"), "w").write("\n".join(lines) + "\n")
