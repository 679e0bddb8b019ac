"""Host mirror of /root/reference/src/io.jl: save / load a GPSLCObject (`*.gpslc`).

The reference dumps the Julia object with `Serialization` (src/io.jl:14-19), a format only Julia can read and which is not
stable across Julia versions. The B200 path keeps posterior samples in the packed layout of SURVEY.md App. A7, so the file is a
small language-neutral container that both this module and the Julia glue (julia/CausalGPSLCB200.jl: savePacked / loadPacked)
read and write:

    bytes 0..7    magic  b"GPSLCB2\\0"
    bytes 8..15   little-endian uint64 H = length of the header
    bytes 16..    H bytes of UTF-8 JSON: {"version", "hyperparams", "priorparams" (scalars only; SigmaU is rebuilt from
                  "obj_counts"), "seed", "arrays": [{"name", "dtype", "shape"}, ...]}
    then          the arrays in header order, raw little-endian, C order (row-major), each padded to a multiple of 8 bytes
                  ("packed" [nOuter][n_chains][stride] float64, "T" float64 or bool, "Y", "X" [n][nX], "obj" labels, "obj_counts")
"""
import json
import struct

import numpy as np

from .data import objectCounts
from .hyperparameters import HyperParameters
from .types import GPSLCObject, PosteriorSample
from .utils import generateSigmaU

MAGIC = b"GPSLCB2\0"


def _strip(filename):
    """src/io.jl:15-17, 30-32: the extension `.gpslc` is optional"""
    if len(filename) > 6 and filename[-6:] == ".gpslc":
        filename = filename[:-6]
    return filename


def saveGPSLCObject(g, filename="gpslc"):
    """src/io.jl:14-19"""
    arrays = [("packed", np.ascontiguousarray(g.posteriorPacked, dtype=np.float64)),
              ("T", np.ascontiguousarray(g.T)), ("Y", np.ascontiguousarray(g.Y, dtype=np.float64))]
    if g.X is not None:
        arrays.append(("X", np.ascontiguousarray(g.X, dtype=np.float64)))
    if g.obj is not None:
        obj = np.asarray(g.obj)
        if obj.dtype.kind not in "iufb":
            obj = np.unique(obj, return_inverse=True)[1].astype(np.int64)    # labels only matter up to equality
        arrays.append(("obj", np.ascontiguousarray(obj)))
    if g.SigmaU is not None:
        from .inference import sigma_u_to_counts
        try:
            counts = sigma_u_to_counts(g.SigmaU, g.priorparams["sigmaUNoise"], g.priorparams["sigmaUCov"])
            arrays.append(("obj_counts", np.asarray(counts, dtype=np.int64)))
        except ValueError:
            arrays.append(("SigmaU", np.ascontiguousarray(g.SigmaU, dtype=np.float64)))     # unstructured: stored densely
    h = g.hyperparams
    header = {"version": 1,
              "hyperparams": {k: getattr(h, k) for k in ("nU", "nOuter", "nMHInner", "nESInner", "nBurnIn", "stepSize",
                                                           "predictionCovarianceNoise")},
              "priorparams": {k: v for k, v in g.priorparams.items() if np.isscalar(v)},
              "seed": int(g.seed),
              "arrays": [{"name": nm, "dtype": a.dtype.str, "shape": list(a.shape)} for nm, a in arrays]}
    hb = json.dumps(header).encode()
    with open(_strip(filename) + ".gpslc", "wb") as f:
        f.write(MAGIC)
        f.write(struct.pack("<Q", len(hb)))
        f.write(hb)
        for _, a in arrays:
            b = a.tobytes()
            f.write(b)
            f.write(b"\0" * (-len(b) % 8))


def loadGPSLCObject(filename="gpslc"):
    """src/io.jl:29-34"""
    with open(_strip(filename) + ".gpslc", "rb") as f:
        if f.read(8) != MAGIC:
            raise ValueError("not a .gpslc file written by saveGPSLCObject of the B200 path")
        (hl,) = struct.unpack("<Q", f.read(8))
        header = json.loads(f.read(hl).decode())
        arr = {}
        for d in header["arrays"]:
            dt = np.dtype(d["dtype"])
            nbytes = int(np.prod(d["shape"], dtype=np.int64)) * dt.itemsize
            arr[d["name"]] = np.frombuffer(f.read(nbytes), dtype=dt).reshape(d["shape"]).copy()
            f.read(-nbytes % 8)
    hp = HyperParameters(**header["hyperparams"])
    pp = dict(header["priorparams"])
    SigmaU = None
    if "obj_counts" in arr:
        SigmaU = generateSigmaU(arr["obj_counts"].tolist(), pp["sigmaUNoise"], pp["sigmaUCov"])
    elif "SigmaU" in arr:
        SigmaU = arr["SigmaU"]
    if SigmaU is not None:
        pp["SigmaU"] = SigmaU              # samplePosterior left it in the caller's dict (src/driver.jl:61; test/io.jl:20)
    T, packed, X = arr["T"], arr["packed"], arr.get("X")
    nU = hp.nU or 0
    nX = 0 if X is None else X.shape[1]
    views = [PosteriorSample(packed[i, 0], T.shape[0], nU, nX, T.dtype == np.bool_) for i in range(packed.shape[0])]
    return GPSLCObject(hp, pp, SigmaU, arr.get("obj"), X, T, arr["Y"], views, packed, header["seed"])
