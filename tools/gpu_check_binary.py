"""Development aid: binary-treatment chains, CUDA vs oracle."""
import sys, os, time
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200.inference import ChainSampler
from oracle import data as od, inference as oi, model as om

def run_case(n, n_obj, nX, nU, with_u, nOuter=3, nMH=2, nES=2, seed=7, C=2, **opts):
    counts, X, T, Y = od.synthetic(n, n_obj, max(nX, 1), seed=5)
    if nX == 0: X = None
    Tb = T > np.median(T)
    md = od.model_data_from_arrays(counts if with_u else None, X, Tb, Y, nU=nU)
    s = ChainSampler(md.prior, X, Tb, Y, md.spec.nU, counts if with_u else None, nOuter, nMH, nES, n_chains=C, seed=seed, **opts)
    st0 = s.state()
    s.run(nOuter); smp = s.samples(); acc, ev = s.stats(); evl = s.ess_evals_logit
    for c in range(C):
        st = oi.generate_initial_state(md, seed, c, observe_x=bool(opts.get("observe_x", 0)))
        packed = oi.pack_sample(md.spec, st)
        e0 = np.nanmax(np.abs(packed - st0[c]) / (1e-12 + np.abs(packed)))
        stats = {}
        so, _ = oi.posterior(md, nOuter, nMH, nES, seed=seed, chain=c, stats=stats, observe_x=bool(opts.get("observe_x", 0)))
        err = np.nanmax(np.abs(so - smp[:, c, :]) / (1e-9 + np.abs(so)), axis=1)
        print(f"n={n} nX={nX} nU={md.spec.nU} chain {c}: init {e0:.1e} per-outer {np.array2string(err, precision=1)} accepts {np.array_equal(stats['accepts'], acc[c].astype(np.int64))} essU {stats.get('ess_evals',0)}/{ev[c]} essL {stats.get('ess_evals_logit',0)}/{evl[c]}")
    s.close()

run_case(48, 4, 3, 1, True)
run_case(100, 5, 2, 2, True)
run_case(150, 6, 0, 1, True)
run_case(64, 4, 3, 1, False)
run_case(64, 4, 3, 1, False, observe_x=1)
run_case(72, 4, 0, 1, False)
