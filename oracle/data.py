"""Oracle restatement of src/data.jl (`prepareData`) plus the synthetic workload generator of SURVEY.md §8(d).
Test infrastructure only (see oracle/__init__.py)."""
import numpy as np

from .model import ModelSpec, ModelData, get_prior_parameters


def remove_adjacent(v):
    """src/utils.jl:39-52"""
    out = []
    for e in v:
        if not out or e != out[-1]:
            out.append(e)
    return out


def prepare_data(path_or_df, eps=1.0e-13, cov=1.0):
    """src/data.jl:20-70. Returns (counts or None, obj or None, X or None, T, Y); rows are SORTED by `obj`
    (data.jl:25); counts follow data.jl:29-39 (total count of each label, listed per run of
    adjacent labels — identical to run lengths once sorted)."""
    import pandas as pd
    df = pd.read_csv(path_or_df) if isinstance(path_or_df, str) else path_or_df.copy()
    counts = obj = None
    if "obj" in df.columns:
        df = df.sort_values("obj", kind="stable").reset_index(drop=True)
        labels = df["obj"].tolist()
        tot = {}
        for o in labels:
            tot[o] = tot.get(o, 0) + 1
        counts = [tot[o] for o in remove_adjacent(labels)]
        obj = np.array(labels)
    T = df["T"].to_numpy()
    Y = df["Y"].to_numpy(dtype=np.float64)
    cols = [c for c in df.columns if c not in ("T", "Y", "obj")]
    X = df[cols].to_numpy(dtype=np.float64) if cols else None
    return counts, obj, X, T, Y


def model_data_from_arrays(counts, X, T, Y, nU=1, prior=None, u_layout_reference=True, sigma_u_dense=None):
    """Assemble ModelData the way the GPSLCObject constructors dispatch (src/types.jl:271-290): no SigmaU => nU=0."""
    T = np.asarray(T)
    binary = T.dtype == np.bool_
    n = T.shape[0]
    prior = dict(prior) if prior is not None else get_prior_parameters()
    spec = ModelSpec(n=n, nU=(nU if (counts is not None or sigma_u_dense is not None) else 0), nX=(0 if X is None else X.shape[1]), binary=binary,
                     u_layout_reference=u_layout_reference)
    return ModelData(spec=spec, X=None if X is None else np.asarray(X, dtype=np.float64),
                     T=T.astype(np.float64), Y=np.asarray(Y, dtype=np.float64),
                     counts=list(counts) if counts is not None else [],
                     eps=prior["sigmaUNoise"], cov=prior["sigmaUCov"], prior=prior,
                     sigma_u_dense=None if sigma_u_dense is None else np.asarray(sigma_u_dense, dtype=np.float64))


def synthetic(n, n_obj, nX, seed=1234):
    """equal-size objects in sorted order; u_obj ~ N(0,1); X ~ N(0,1); w, v ~ N(0,1)/sqrt(nX);
    T = 0.5 u + 0.3 Xw + 0.5 e_T; Y = sin(T) + u + 0.3 Xv + 0.3 e_Y. Returns (counts, X, T, Y)."""
    assert n % n_obj == 0
    rng = np.random.default_rng(seed)
    m = n // n_obj
    counts = [m] * n_obj
    u_obj = rng.standard_normal(n_obj)
    obj = np.repeat(np.arange(n_obj), m)
    X = rng.standard_normal((n, nX)) if nX > 0 else None
    if nX > 0:
        w = rng.standard_normal(nX) / np.sqrt(nX)
        v = rng.standard_normal(nX) / np.sqrt(nX)
        xw, xv = X @ w, X @ v
    else:
        xw = xv = np.zeros(n)
    T = 0.5 * u_obj[obj] + 0.3 * xw + 0.5 * rng.standard_normal(n)
    Y = np.sin(T) + u_obj[obj] + 0.3 * xv + 0.3 * rng.standard_normal(n)
    return counts, X, T, Y
