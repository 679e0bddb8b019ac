"""Development aid: CUDA sampler vs the oracle chain driven by the same Philox streams."""
import sys, os, time
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200.inference import ChainSampler, Posterior
from oracle import data as od, inference as oi, model as om

def run_case(n, n_obj, nX, nU, nOuter, nMH, nES, seed, C=3, counts_on=True):
    counts, X, T, Y = od.synthetic(n, n_obj, nX, seed=5)
    if not counts_on: counts = None
    md = od.model_data_from_arrays(counts, X, T, Y, nU=nU)
    pri = md.prior
    s = ChainSampler(pri, X, T, Y, md.spec.nU, counts, nOuter, nMH, nES, n_chains=C, seed=seed)
    st0 = s.state(); lp0, q0 = s.terms()
    # oracle initial state
    worst = 0
    for c in range(C):
        st = oi.generate_initial_state(md, seed, c)
        packed = oi.pack_sample(md.spec, st)
        e0 = np.nanmax(np.abs(packed - st0[c]) / (1e-300 + np.abs(packed)))
        lpo = np.array([om.factor_logpdf(md, st, f) if om.factor_exists(md.spec, f) else 0.0 for f in range(md.spec.nX + 2)])
        e1 = np.max(np.abs(lpo - lp0[c]) / (1 + np.abs(lpo)))
        worst = max(worst, e0, e1)
    print(f"n={n} nX={nX} nU={md.spec.nU} init state/logpdf rel err {worst:.2e}")
    t = time.time(); s.run(nOuter); smp = s.samples(); dt = time.time() - t
    acc, ev = s.stats()
    for c in range(C):
        stats = {}
        so, _ = oi.posterior(md, nOuter, nMH, nES, seed=seed, chain=c, stats=stats)
        err = np.nanmax(np.abs(so - smp[:, c, :]) / (1e-12 + np.abs(so)), axis=1)
        print(f"   chain {c}: per-outer max rel err {np.array2string(err, precision=1)} accepts match {np.array_equal(stats['accepts'], acc[c].astype(np.int64))} ess evals {stats.get('ess_evals')} vs {ev[c]}")
    print(f"   gpu time {dt:.3f}s")
    s.close()

run_case(48, 4, 3, 1, 3, 2, 2, seed=7)
run_case(100, 5, 2, 2, 2, 2, 2, seed=11)
run_case(150, 6, 0, 1, 3, 3, 2, seed=3)
run_case(64, 4, 3, 1, 2, 2, 2, seed=9, counts_on=False)
run_case(70, 4, 0, 1, 3, 2, 2, seed=9, counts_on=False)
run_case(256, 4, 5, 1, 2, 2, 1, seed=21, C=2)
