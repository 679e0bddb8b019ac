"""Host mirror of /root/reference/src/driver.jl: `gpslc`, `samplePosterior`, `sampleITE`, `sampleSATE`,
`summarizeEstimates` with the reference's signatures; the bodies call the CUDA library."""
import copy

import numpy as np

from ._lib import check_info
from .data import prepareData, objectCounts
from .estimation import ite as _ite, sate as _sate, retained_indices
from .hyperparameters import getHyperParameters, getPriorParameters
from .inference import Posterior
from .types import GPSLCObject, PosteriorSample
from .utils import generateSigmaU


def samplePosterior(hyperparams, priorparams, SigmaU, X, T, Y, n_chains=1, seed=0, ctx=None, _counts=None, **opts):
    """src/driver.jl:59-69. Mutates priorparams["SigmaU"] like the reference (App. B11). Returns the packed samples
    [nOuter, n_chains, stride] (and stats)."""
    priorparams["SigmaU"] = SigmaU
    pp = priorparams if _counts is None else {**priorparams, "_obj_counts": _counts}
    return Posterior(pp, X, T, Y, hyperparams.nU, hyperparams.nOuter, hyperparams.nMHInner, hyperparams.nESInner,
                     n_chains=n_chains, seed=seed, ctx=ctx, return_stats=True, **opts)


def _make_object(hyperparams, priorparams, SigmaU, obj, X, T, Y, counts, n_chains, seed, ctx, opts):
    """The three GPSLCObject constructors (src/types.jl:271-290), including their overwrites of nU/nMHInner/nESInner
    with `nothing` (App. B12)."""
    T = np.asarray(T)
    if SigmaU is None:
        hyperparams.nU = None
        if X is None:
            hyperparams.nMHInner = None
            hyperparams.nESInner = None
    packed, stats = samplePosterior(hyperparams, priorparams, SigmaU, X, T, Y, n_chains=n_chains, seed=seed, ctx=ctx,
                                    _counts=counts, **opts)
    n = T.shape[0]
    nU = hyperparams.nU or 0
    nX = 0 if X is None else X.shape[1]
    views = [PosteriorSample(packed[i, 0], n, nU, nX, T.dtype == np.bool_) for i in range(packed.shape[0])]
    return GPSLCObject(hyperparams, priorparams, SigmaU, obj, None if X is None else np.asarray(X, dtype=np.float64), T,
                       np.asarray(Y, dtype=np.float64), views, packed, seed, stats)


def gpslc(*args, hyperparams=None, priorparams=None, n_chains=1, seed=0, ctx=None, **opts):
    """src/driver.jl:27-44.
        gpslc("file.csv" | DataFrame; hyperparams, priorparams)
        gpslc(obj_counts | None, X | None, T, Y; hyperparams, priorparams)     (obj is passed to generateSigmaU as counts, App. B4)
    Extra keywords (not in the reference): n_chains, seed, u_layout_mode, ess_rule, observe_x."""
    hyperparams = copy.copy(hyperparams) if hyperparams is not None else getHyperParameters()
    priorparams = priorparams if priorparams is not None else getPriorParameters()
    if len(args) == 1:
        SigmaU, obj, X, T, Y = prepareData(args[0])     # default eps / cov, whatever priorparams says (src/driver.jl:31)
        counts = None                                   # the structure (counts, eps, cov) is read off SigmaU itself
    elif len(args) == 4:
        obj, X, T, Y = args
        counts = None
        if obj is not None:
            SigmaU = generateSigmaU([int(c) for c in obj], priorparams["sigmaUNoise"], priorparams["sigmaUCov"])
        else:
            SigmaU = None
    else:
        raise TypeError("gpslc(data) or gpslc(obj, X, T, Y)")
    return _make_object(hyperparams, priorparams, SigmaU, obj, X, T, Y, counts, n_chains, seed, ctx, opts)


def _ret(g):
    h = g.hyperparams
    return retained_indices(h.nBurnIn, h.stepSize, h.nOuter)


def ITEDistributions(g, doT, ctx=None):
    """src/estimation.jl:66-86 -> (MeanITEs [R, n], CovITEs [R, n, n]) for chain 0."""
    o = _ite(g.posteriorPacked[:, :1], g.X, g.T, g.Y, g.hyperparams.nU, float(doT), _ret(g),
             g.hyperparams.predictionCovarianceNoise, 0, want_cov=True, want_samples=False, ctx=ctx)
    check_info(o["info"], "ITEDistributions", doT)
    return o["mean"][0, 0], o["cov"][0, 0]


def sampleITE(g, doT, samplesPerPosterior=10, all_chains=False, ctx=None):
    """src/driver.jl:86-89 -> n × (R*samplesPerPosterior) (chain 0; all_chains=True returns [n_chains, n, R*spp])."""
    packed = g.posteriorPacked if all_chains else g.posteriorPacked[:, :1]
    o = _ite(packed, g.X, g.T, g.Y, g.hyperparams.nU, float(doT), _ret(g), g.hyperparams.predictionCovarianceNoise,
             samplesPerPosterior, seed=g.seed, ctx=ctx)
    check_info(o["info"], "sampleITE", doT)
    s = np.swapaxes(o["samples"][0], 1, 2)          # [C, n, R*spp]
    return s if all_chains else s[0]


def SATEDistributions(g, doT, ctx=None):
    """src/estimation.jl:127-140 -> (MeanSATEs [R], VarSATEs [R])."""
    o = _sate(g.posteriorPacked[:, :1], g.X, g.T, g.Y, g.hyperparams.nU, float(doT), _ret(g),
              g.hyperparams.predictionCovarianceNoise, 0, ctx=ctx)
    check_info(o["info"], "SATEDistributions", doT)
    return o["mean"][0, 0], o["var"][0, 0]


def sampleSATE(g, doT, samplesPerPosterior=10, all_chains=False, var_as_std=True, ctx=None):
    """src/driver.jl:108-111 -> vector of length R*samplesPerPosterior."""
    packed = g.posteriorPacked if all_chains else g.posteriorPacked[:, :1]
    o = _sate(packed, g.X, g.T, g.Y, g.hyperparams.nU, float(doT), _ret(g), g.hyperparams.predictionCovarianceNoise,
              samplesPerPosterior, seed=g.seed, var_as_std=var_as_std, ctx=ctx)
    check_info(o["info"], "sampleSATE", doT)
    return o["samples"][0] if all_chains else o["samples"][0, 0]


def summarizeEstimates(samples, savetofile="", credible_interval=0.90, ctx=None):
    """src/driver.jl:129-149: DataFrame Individual/Mean/LowerBound/UpperBound from the n x m matrix sampleITE returns. The
    row statistics (mean, two type-7 quantiles) are computed by the CUDA library (gpslc_summarize)."""
    import pandas as pd
    from .estimation import summarize
    samples = np.asarray(samples, dtype=np.float64)
    st = summarize(np.ascontiguousarray(samples.T), credible_interval, ctx=ctx)      # [n, 3]
    df = pd.DataFrame({"Individual": np.arange(1, samples.shape[0] + 1), "Mean": st[:, 0], "LowerBound": st[:, 1],
                       "UpperBound": st[:, 2]})
    if savetofile != "":
        df.to_csv(savetofile, index=False)
        print("Saved mean and 90% credible intervals to " + savetofile)
    return df
