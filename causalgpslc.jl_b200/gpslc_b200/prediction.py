"""Host mirror of /root/reference/src/prediction.jl: `predictCounterfactualEffects`. The reference calls sampleITE once
per doT (re-extracting parameters and re-factorising every time); here all doT values go to the GPU in
one gpslc_ite call."""
import numpy as np

from ._lib import check_info
from .driver import _ret
from .estimation import ite as _ite
from .utils import getN, getNumPosteriorSamples


def predictCounterfactualEffects(g, nSamplesPerMixture, fidelity=100, minDoT=None, maxDoT=None, ctx=None, world_size=1, rank=0):
    """src/prediction.jl:23-36 -> (ite [d, n, R*nSamplesPerMixture], doTrange).
    world_size/rank (not in the reference): the doT values are the independent units of the sweep; rank r
    computes the contiguous block `shard_chains(len(doTrange), world_size, r)` and returns only those rows of `ite` (plus the
    full range). The draws are keyed by the global doT index, so the shards concatenate to the unsharded result."""
    from .parallel import shard_chains
    minDoT = float(np.min(g.T)) if minDoT is None else float(minDoT)
    maxDoT = float(np.max(g.T)) if maxDoT is None else float(maxDoT)
    doTrange = np.linspace(minDoT, maxDoT, fidelity + 1)      # minDoT:step:maxDoT, step = |max-min|/fidelity
    off, cnt = shard_chains(len(doTrange), world_size, rank)
    o = _ite(g.posteriorPacked[:, :1], g.X, g.T, g.Y, g.hyperparams.nU, doTrange[off:off + cnt], _ret(g),
             g.hyperparams.predictionCovarianceNoise, nSamplesPerMixture, seed=g.seed, ctx=ctx, dot_offset=off)
    check_info(o["info"], "predictCounterfactualEffects", doTrange[off:off + cnt])
    ite = np.swapaxes(o["samples"][:, 0], 1, 2)               # [d_local, n, R*spp]
    assert ite.shape == (cnt, getN(g), getNumPosteriorSamples(g) * nSamplesPerMixture)
    return ite, doTrange
