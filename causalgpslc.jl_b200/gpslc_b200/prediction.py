"""Host mirror of /root/reference/src/prediction.jl: `predictCounterfactualEffects`. The reference calls sampleITE once
per doT (re-extracting parameters and re-factorising every time); here all doT values go to the GPU in
one gpslc_ite call."""
import numpy as np

from .driver import _ret
from .estimation import ite as _ite
from .utils import getN, getNumPosteriorSamples


def predictCounterfactualEffects(g, nSamplesPerMixture, fidelity=100, minDoT=None, maxDoT=None, ctx=None):
    """src/prediction.jl:23-36 -> (ite [d, n, R*nSamplesPerMixture], doTrange)."""
    minDoT = float(np.min(g.T)) if minDoT is None else float(minDoT)
    maxDoT = float(np.max(g.T)) if maxDoT is None else float(maxDoT)
    doTrange = np.linspace(minDoT, maxDoT, fidelity + 1)      # minDoT:step:maxDoT, step = |max-min|/fidelity
    o = _ite(g.posteriorPacked[:, :1], g.X, g.T, g.Y, g.hyperparams.nU, doTrange, _ret(g),
             g.hyperparams.predictionCovarianceNoise, nSamplesPerMixture, seed=g.seed, ctx=ctx)
    ite = np.swapaxes(o["samples"][:, 0], 1, 2)               # [d, n, R*spp]
    assert ite.shape == (len(doTrange), getN(g), getNumPosteriorSamples(g) * nSamplesPerMixture)
    return ite, doTrange
