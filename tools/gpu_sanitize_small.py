"""Small invocations of the code paths added in round 2 (compute-sanitizer is closed on this pool, so this is a plain smoke run with
finiteness / symmetry checks; written so that it can run under `compute-sanitizer --tool memcheck` where that is allowed): shared-Kp ITE (one CTA per task, cluster teams, grid-wide base factor), fused ITE summary, dense SigmaU sampler, radix-select
summaries, 256-bit covariance stores with ragged n, many feature dimensions."""
import os, sys
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200 import estimation as ge
from gpslc_b200.inference import ChainSampler
from bench import synthetic

ctx = g.Context(0)
rng = np.random.default_rng(0)
for n, n_obj, nX, team in ((100, 4, 2, None), (200, 4, 1, "2"), (580, 4, 1, None)):
    counts, X, T, Y = synthetic(n, n_obj, nX, seed=3)
    npar = 6 + 4 * nX + 2 + nX
    smp = np.ones((2, 1, npar + n)); smp[:, :, :npar] = 0.8 + 0.4 * rng.random((2, 1, npar)); smp[:, :, 2] = 0.3
    smp[:, :, npar:] = rng.standard_normal((2, 1, n))
    if team: os.environ["GPSLC_TEAM"] = team
    o = ge.ite(smp, X, T, Y, 1, (0.1, 0.7, -0.3), np.arange(2, dtype=np.int32), 1e-10, 2, want_cov=(n < 300), ctx=ctx)
    s, info = ge.ite_summary(smp, X, T, Y, 1, (0.1, 0.7), np.arange(2, dtype=np.int32), 1e-10, 3, ctx=ctx)
    os.environ.pop("GPSLC_TEAM", None)
    print("ite shared", n, team, int(o["info"].max()), bool(np.isfinite(o["samples"]).all()), s.shape, int(info.max()))
n = 70
counts, X, T, Y = synthetic(n, 2, 2, seed=5)
i = np.arange(n); S = 0.5 * np.exp(-np.abs(i[:, None] - i[None, :]) / 5.0) + 0.6 * np.eye(n)
sm = ChainSampler({**g.getPriorParameters(), "SigmaU": S}, X, T > np.median(T), Y, 1, None, 2, 1, 1, n_chains=2, seed=1, ctx=ctx)
sm.run(2); print("dense SigmaU binary", bool(np.isfinite(sm.samples()).all())); sm.close()
print("summarize select", ge.summarize(rng.standard_normal((1, 9000, 5)), 0.9, ctx=ctx).shape)
for nn, D in ((100, 3), (129, 30), (64, 1)):
    F = rng.standard_normal((nn, D))
    K = g.cov_build(F, F, np.ones((2, D)), [1.0, 2.0], [0.1, 0.2], ctx=ctx)
    print("cov_build", nn, D, bool(np.allclose(K[0], K[0].T)))
print("sanitize-small done")
