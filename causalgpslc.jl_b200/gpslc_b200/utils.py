"""Host mirror of the layout-defining helpers of /root/reference/src/utils.jl (trivial host code, no compute)."""
import numpy as np


def generateSigmaU(nIndividualsArray, eps=1e-13, cov=1.0):
    """src/utils.jl:17-33"""
    counts = [int(c) for c in nIndividualsArray]
    n = sum(counts)
    S = np.eye(n)
    i = 0
    for m in counts:
        S[i:i + m, i:i + m] = np.ones((m, m)) * cov
        i += m
    S[np.diag_indices(n)] = 1 + eps
    return S


def removeAdjacent(v):
    """src/utils.jl:39-52"""
    out = []
    for e in v:
        if not out or e != out[-1]:
            out.append(e)
    return out


def toMatrix(X, n, m):
    """src/utils.jl:60-64 — reshape(permutedims(hcat(X...)), (n, m)), column-major (interleaves unless m == 1 or X is
    already a Matrix)."""
    if isinstance(X, np.ndarray) and X.ndim == 2:
        return X.reshape(-1, order="F").reshape((n, m), order="F")
    h = np.stack([np.asarray(v, dtype=np.float64) for v in X], axis=1)
    return h.T.reshape(-1, order="F").reshape((n, m), order="F")


def getN(g):
    """src/utils.jl:130-132"""
    return g.Y.shape[0]


def getNX(g):
    """src/utils.jl:138-140"""
    return g.X.shape[1]


def getNU(g):
    """src/utils.jl:146-148"""
    return g.hyperparams.nU


def getNumPosteriorSamples(g):
    """src/utils.jl:156-161 — length(nBurnIn:stepSize:nOuter)"""
    h = g.hyperparams
    return len(range(h.nBurnIn, h.nOuter + 1, h.stepSize))


def extractParameters(g, posteriorSampleIdx):
    """src/utils.jl:92-124 (1-based index): (uyLS, xyLS, tyLS, yNoise, yScale, U n×nU)."""
    s = g.posteriorSamples[posteriorSampleIdx - 1]
    nU = getNU(g)
    if nU is None:
        uyLS = U = None
    else:
        uyLS = np.array([s[("uyLS", u, "LS")] for u in range(1, nU + 1)])
        U = np.stack([s[("U", u, "U")] for u in range(1, nU + 1)], axis=1)
    xyLS = None if g.X is None else np.array([s[("xyLS", k, "LS")] for k in range(1, getNX(g) + 1)])
    return uyLS, xyLS, s["tyLS"], s["yNoise"], s["yScale"], U
