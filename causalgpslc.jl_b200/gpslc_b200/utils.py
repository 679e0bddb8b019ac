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
