"""Host mirror of /root/reference/src/prediction.jl: `predictCounterfactualEffects`. The reference calls sampleITE once
per doT (re-extracting parameters and re-factorising every time); here all doT values go to the GPU in
one gpslc_ite call."""
import numpy as np

from ._lib import check_info
from .driver import _ret
from .estimation import ite as _ite, ite_subset_summary as _ite_subset_summary
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


def subgroupEffectCurve(g, idx, nSamplesPerMixture, fidelity=100, minDoT=None, maxDoT=None, credible_interval=0.90, ctx=None,
                        world_size=1, rank=0):
    """The reference's documented subgroup workflow (docs/src/index.md:101-114) as one device call:
        ite, doT = predictCounterfactualEffects(g, nSamples); sate = mean(ite[:, idx, :], dims=2)[:, 1, :]
        interval = summarizeEstimates(sate)
    `idx`: boolean mask over the individuals (the example's `vec(g.obj .== "MA")`). Returns (interval, sate, doTrange) with
    interval = dict(Mean, LowerBound, UpperBound), one entry per doT of this rank's block, and sate [d_local, R*nSamplesPerMixture].
    Neither the n x (R*nSamples) draws of each doT nor the subgroup averages cross PCIe before they are reduced."""
    from .parallel import shard_chains
    minDoT = float(np.min(g.T)) if minDoT is None else float(minDoT)
    maxDoT = float(np.max(g.T)) if maxDoT is None else float(maxDoT)
    doTrange = np.linspace(minDoT, maxDoT, fidelity + 1)
    off, cnt = shard_chains(len(doTrange), world_size, rank)
    summ, sub, info = _ite_subset_summary(g.posteriorPacked[:, :1], g.X, g.T, g.Y, g.hyperparams.nU, doTrange[off:off + cnt], _ret(g),
                                          g.hyperparams.predictionCovarianceNoise, nSamplesPerMixture, idx, seed=g.seed,
                                          credible_interval=credible_interval, ctx=ctx, dot_offset=off)
    check_info(info, "subgroupEffectCurve", doTrange[off:off + cnt])
    interval = {"Mean": summ[0, :, 0].copy(), "LowerBound": summ[0, :, 1].copy(), "UpperBound": summ[0, :, 2].copy()}
    return interval, sub[0].T.copy(), doTrange
