"""Development aid: how often does a default-length chain of ours land where the reference's golden ITE summaries are?
For every dataset with golden files: C chains, both elliptical-slice acceptance rules; per chain the fraction of individuals whose
mean ITE falls inside the golden 90 % interval (the reference's own gate is >= 0.5 on one chain, test/driver.jl:46-52)."""
import sys, os
import numpy as np, pandas as pd
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
GOLD = os.path.join(root, "tests", "golden")
ctx = g.Context(0)
C = int(sys.argv[1]) if len(sys.argv) > 1 else 64
cases = [("additive_linear", (0, 1)), ("additive_nonlinear", (0, 1)), ("multiplicative_linear", (0, 1)), ("multiplicative_nonlinear", (0, 1)),
         ("NEEC_sampled", (0, 0.6, 1))]
for name, dots in cases:
    for rule in (0,):
        gobj = g.gpslc(os.path.join(GOLD, "data", name + ".csv"), seed=100, n_chains=C, ctx=ctx, ess_rule=rule)
        for doT in dots:
            ite = g.sampleITE(gobj, float(doT), all_chains=True, ctx=ctx)     # [C, n, R*spp]
            exp = pd.read_csv(os.path.join(GOLD, "results", f"{name}_{doT}.csv"))
            means = ite.mean(axis=2)                                            # [C, n]
            inside = ((exp["LowerBound"].values[None] <= means) & (means <= exp["UpperBound"].values[None])).mean(axis=1)
            corr = np.array([np.corrcoef(m, exp["Mean"].values)[0, 1] for m in means])
            pooled = means.mean(axis=0)
            pin = ((exp["LowerBound"].values <= pooled) & (pooled <= exp["UpperBound"].values)).mean()
            print(f"{name:26s} doT={doT:<4} ess_rule={rule} ls_unsquared={os.environ.get("GPSLC_LS_UNSQUARED", "0")}: chains with inside>=0.5: {100*(inside>=0.5).mean():5.1f} %, median inside {np.median(inside):.2f}, "
                  f"max {inside.max():.2f}; corr with golden means: median {np.median(corr):+.2f}, frac>0.5: {100*(corr>0.5).mean():5.1f} %; pooled-over-chains inside {pin:.2f}")
