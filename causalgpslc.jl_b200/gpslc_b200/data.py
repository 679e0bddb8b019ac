"""Host mirror of /root/reference/src/data.jl (`prepareData`): CSV/DataFrame -> (SigmaU, obj, X, T, Y). Host I/O that
runs once; outside the GPU hot path."""
import numpy as np

from .utils import generateSigmaU, removeAdjacent


def objectCounts(obj):
    """[a, a, a, b, c, c] -> [3, 1, 2] (src/data.jl:27-38): the sizes of the blocks of SigmaU, in row order. `obj` must be
    sorted the way prepareData leaves it; this is what the CUDA path consumes instead of the dense SigmaU."""
    if obj is None:
        return None
    labels = list(obj)
    tot = {}
    for o in labels:
        tot[o] = tot.get(o, 0) + 1
    return [tot[o] for o in removeAdjacent(labels)]


def prepareData(df, confounderEps=1.0e-13, confounderCov=1.0):
    """src/data.jl:20-70. Rows are sorted by `obj` (data.jl:25); returns (SigmaU, obj, X, T, Y) like the reference."""
    import pandas as pd
    if isinstance(df, str):
        df = pd.read_csv(df)
    else:
        df = df.copy()
    if "obj" in df.columns:
        df = df.sort_values("obj", kind="stable").reset_index(drop=True)
        obj = np.array(df["obj"].tolist())
        SigmaU = generateSigmaU(objectCounts(obj), confounderEps, confounderCov)
    else:
        print("No object labels to assign latent confounders to (column must be titled `obj`)")
        obj = None
        SigmaU = None
        print("Assuming no latent confounding")
    T = df["T"].to_numpy()
    Y = df["Y"].to_numpy(dtype=np.float64)
    cols = [c for c in df.columns if c not in ("T", "Y", "obj")]
    if not cols:
        print("No observed covariates found in data")
        X = None
    else:
        X = df[cols].to_numpy(dtype=np.float64)
    return SigmaU, obj, X, T, Y
