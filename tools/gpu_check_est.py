"""Development aid: CUDA ITE/SATE vs the oracle's literal restatement of likelihood.jl / estimation.jl."""
import sys, os, time
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import gpslc_b200 as g
from gpslc_b200 import estimation as ge
from oracle import data as od, inference as oi, estimation as oe

def case(n, n_obj, nX, nU, with_u=True, doTs=(0.3, -0.5)):
    counts, X, T, Y = od.synthetic(n, n_obj, max(nX, 1), seed=8)
    if nX == 0: X = None
    md = od.model_data_from_arrays(counts if with_u else None, X, T, Y, nU=nU)
    smp = np.stack([oi.posterior(md, 4, 1, 1, seed=5, chain=c, observe_x=True)[0] for c in range(2)], axis=1)  # [4, 2, stride]
    ret = np.array([1, 3], dtype=np.int32)
    jit = 1e-10
    t = time.time()
    out = ge.ite(smp, X, T, Y, md.spec.nU, doTs, ret, jit, 3, seed=9, want_cov=True)
    so = ge.sate(smp, X, T, Y, md.spec.nU, doTs, ret, jit, 3, seed=9)
    dt = time.time() - t
    em = ec = es = esm = esv = ess = 0.0
    for d, doT in enumerate(doTs):
        for c in range(2):
            M, Cv = oe.ite_distributions(md.spec, smp[:, c, :], X, T, Y, doT, 2, 2, jit)
            S = oe.ite_samples(M, Cv, 3, seed=9, chain=c, dot_index=d)
            em = max(em, np.abs(M - out["mean"][d, c]).max() / np.abs(M).max())
            ec = max(ec, np.abs(Cv - out["cov"][d, c]).max() / np.abs(Cv).max())
            es = max(es, np.abs(S.T - out["samples"][d, c]).max() / np.abs(S).max())
            ms, vs = oe.sate_distributions(M, Cv)
            ss = oe.sate_samples(ms, vs, 3, seed=9, chain=c, dot_index=d)
            esm = max(esm, np.abs(ms - so["mean"][d, c]).max() / (1e-300 + np.abs(ms).max()))
            esv = max(esv, np.abs(vs - so["var"][d, c]).max() / np.abs(vs).max())
            ess = max(ess, np.abs(ss - so["samples"][d, c]).max() / np.abs(ss).max())
    print(f"n={n} nX={nX} nU={md.spec.nU}: mean {em:.1e} cov {ec:.1e} draws {es:.1e} | sate mean {esm:.1e} var {esv:.1e} draws {ess:.1e}  info {out['info'].max()} {so['info'].max()}  ({dt:.2f}s)")

case(40, 4, 3, 1)
case(100, 5, 2, 2)
case(150, 6, 0, 1)
case(64, 4, 3, 1, with_u=False)
case(70, 5, 0, 1, with_u=False)
case(300, 6, 4, 1)
# zero-effect KAT through the C ABI (test/estimation.jl:6-67): n = 1, doT == T
smp = np.zeros((1, 1, 6 + 4 + 2 + 1 + 1)); smp[0, 0, :6] = [1, 1, 1.0, 1.0, 1, 1.0]; smp[0, 0, 6:13] = 1.0; smp[0, 0, 13] = 1.0
o = ge.ite(smp, np.ones((1, 1)), np.array([1.0]), np.array([0.37]), 1, [1.0], np.array([0], dtype=np.int32), 1e-10, 5, want_cov=True)
print("zero-effect", o["mean"].ravel(), o["cov"].ravel(), o["samples"].ravel())
