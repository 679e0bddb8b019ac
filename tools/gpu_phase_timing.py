"""Development aid: where does a CTA's time go inside factor_run? Needs the phase-timing build
   GPSLC_EXTRA_FLAGS=-DGPSLC_PHASE_TIMING GPSLC_LIB_SUFFIX=_prof bash causalgpslc.jl_b200/build.sh
usage: GPSLC_LIB_SUFFIX=_prof python tools/gpu_phase_timing.py [n n_obj nX chains]"""
import ctypes, os, sys
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200.inference import ChainSampler
from bench import synthetic, default_priors
n, n_obj, nX, C = (int(x) for x in sys.argv[1:5]) if len(sys.argv) >= 5 else (1024, 16, 10, 512)
counts, X, T, Y = synthetic(n, n_obj, nX)
ctx = g.Context(0)
smp = ChainSampler(default_priors(), X, T, Y, 1, counts, nOuter=24, nMHInner=10, nESInner=5, n_chains=C, seed=1234, ctx=ctx)
smp.mh_sweeps(1); ctx.synchronize()
out = (ctypes.c_ulonglong * 64)()
ctx.lib.gpslc_debug_phase_cycles(out, 1)
smp.mh_sweeps(2); ctx.synchronize()
ctx.lib.gpslc_debug_phase_cycles(out, 1)
v = np.array(list(out), dtype=np.float64)
names = {0: "diag k-loop (+stage_cols)", 1: "diag gen -> Cs, w", 5: "P2 (total)", 8: "  P2: 16x16 potf2+inverse (warp 0)", 9: "  P2: TRSM",
         10: "  P2: SYRK", 11: "  P2: 64x64 inverse assembly", 6: "store L_jj, z, gram", 2: "row k-loops",
         3: "row epilogue (gen, Linv mult, store)", 4: "end-of-panel sync"}
tot = v[7]
print(f"n={n} nX={nX} chains={C}: thread-0 cycles per phase, fraction of the CTA's kernel time")
for k in (0, 1, 5, 8, 9, 10, 11, 6, 2, 3, 4):
    print(f"  {names[k]:40s} {100 * v[k] / tot:6.2f} %")
print(f"  {'outside factor_run':40s} {100 * (tot - v[[0, 1, 5, 6, 2, 3, 4]].sum()) / tot:6.2f} %")
print(f"  row k-loop, per-warp average: waiting for operands (full barrier) {100 * v[12] / v[13]:.2f} % of the loop, "
      f"elected producer (empty barrier + TMA issue) {100 * v[14] / v[13]:.2f} %")
print(f"  8x8 potf2 (warp 0): loads {100 * v[16] / tot:.2f} %, factor loop {100 * v[17] / tot:.2f} %, inverse loop {100 * v[18] / tot:.2f} %, stores {100 * v[19] / tot:.2f} % of the CTA's time")
print("  per warp, fraction of the CTA's kernel time: end-of-panel barrier wait / row k-loop total / operand (mbarrier) wait; stage refills issued")
for w in range(8):
    print(f"    warp {w}: barrier {100 * v[24 + w] / tot:5.2f} %   k-loop {100 * v[32 + w] / tot:5.2f} %   operand wait {100 * v[40 + w] / tot:5.2f} %   refills {int(v[48 + w])}")
