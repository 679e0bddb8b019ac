"""Development aid: NEEC posterior hyperparameters and ITE means at several doT, over many chains."""
import sys, os
import numpy as np, pandas as pd
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
GOLD = os.path.join(root, "tests", "golden")
ctx = g.Context(0)
C = 64
for nOuter in (24, 200):
    h = g.getHyperParameters(); h.nOuter = nOuter
    if nOuter > 24: h.nBurnIn = nOuter // 2
    gobj = g.gpslc(os.path.join(GOLD, "data", "NEEC_sampled.csv"), seed=100, n_chains=C, hyperparams=h, ctx=ctx)
    P = gobj.posteriorPacked            # [nOuter, C, stride]; nX = 0, nU = 1: params = uNoise,tNoise,yNoise,tyLS,tScale,yScale,utLS,uyLS
    names = ["uNoise", "tNoise", "yNoise", "tyLS", "tScale", "yScale", "utLS", "uyLS"]
    last = P[h.nBurnIn - 1:, :, :8].reshape(-1, 8)
    print(f"nOuter={nOuter}: retained-sample medians [q10, q90] over {C} chains")
    for k, nm in enumerate(names):
        print(f"   {nm:7s} {np.median(last[:, k]):8.3f} [{np.quantile(last[:, k], .1):7.3f}, {np.quantile(last[:, k], .9):7.3f}]")
    for doT in (0.0, 0.6, 1.0):
        ite = g.sampleITE(gobj, doT, all_chains=True, ctx=ctx)
        means = ite.mean(axis=2)
        exp = pd.read_csv(os.path.join(GOLD, "results", f"NEEC_sampled_{doT if doT != 0 else 0}.csv".replace("0.0", "0").replace("_1.0", "_1")))
        inside = ((exp["LowerBound"].values[None] <= means) & (means <= exp["UpperBound"].values[None])).mean(axis=1)
        print(f"   doT={doT}: our mean of means per chain: median {np.median(means.mean(axis=1)):+.3f} [{np.quantile(means.mean(axis=1), .1):+.3f}, {np.quantile(means.mean(axis=1), .9):+.3f}]; golden {exp['Mean'].mean():+.3f}; inside median {np.median(inside):.2f}")
