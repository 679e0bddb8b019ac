"""Host mirror of /root/reference/src/io.jl: save / load a GPSLCObject (`*.gpslc`). The reference uses Julia's
Serialization; here the packed samples and data go into an .npz-in-a-file of the same name convention."""
import pickle

from .types import GPSLCObject, PosteriorSample


def saveGPSLCObject(g, filename="gpslc"):
    """src/io.jl:14-19"""
    d = dict(hyperparams=g.hyperparams, priorparams=g.priorparams, SigmaU=g.SigmaU, obj=g.obj, X=g.X, T=g.T, Y=g.Y,
             packed=g.posteriorPacked, seed=g.seed)
    with open(filename + ".gpslc", "wb") as f:
        pickle.dump(d, f)


def loadGPSLCObject(filename="gpslc"):
    """src/io.jl:29-34"""
    import numpy as np
    with open(filename + ".gpslc", "rb") as f:
        d = pickle.load(f)
    T = d["T"]
    nU = d["hyperparams"].nU or 0
    nX = 0 if d["X"] is None else d["X"].shape[1]
    views = [PosteriorSample(d["packed"][i, 0], T.shape[0], nU, nX, T.dtype == np.bool_) for i in range(d["packed"].shape[0])]
    return GPSLCObject(d["hyperparams"], d["priorparams"], d["SigmaU"], d["obj"], d["X"], T, d["Y"], views, d["packed"], d["seed"])
