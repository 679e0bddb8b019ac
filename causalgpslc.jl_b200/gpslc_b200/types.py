"""Host mirror of /root/reference/src/types.jl: `GPSLCObject` and the posterior-sample view.

The reference stores `posteriorSamples::Vector{Any}` of Gen choicemaps (src/types.jl:257). Here the samples stay in the
packed layout the CUDA library produces and `PosteriorSample` exposes the same addresses:
sample[:tyLS] -> s["tyLS"]; sample[:uyLS=>u=>:LS] -> s[("uyLS", u, "LS")]; sample[:U=>u=>:U] -> s[("U", u, "U")]
(1-based indices, like the reference)."""
from dataclasses import dataclass, field
from typing import Any, Optional

import numpy as np

from .hyperparameters import HyperParameters

_SCALARS = {"uNoise": 0, "tNoise": 1, "yNoise": 2, "tyLS": 3, "tScale": 4, "yScale": 5}


class PosteriorSample:
    """One element of `g.posteriorSamples`: a read-only view of a packed record addressed like a Gen choicemap
    (address list: src/proposal.jl:8-22, src/utils.jl:92-124)."""

    def __init__(self, rec, n, nU, nX, binary):
        self.rec, self.n, self.nU, self.nX, self.binary = rec, n, nU, nX, binary
        self.n_params = 6 + 4 * nX + 2 * nU + nU * nX

    def __getitem__(self, addr):
        nX, nU, n = self.nX, self.nU, self.n
        if isinstance(addr, str):
            if addr in _SCALARS:
                return float(self.rec[_SCALARS[addr]])
            if addr == "logitT" and self.binary:
                o = self.n_params + nU * n
                return self.rec[o:o + n]
            raise KeyError(addr)
        name, i = addr[0], addr[1]
        if name == "xNoise":
            return float(self.rec[6 + i - 1])
        if name == "xScale":
            return float(self.rec[6 + nX + i - 1])
        if name == "xtLS":
            return float(self.rec[6 + 2 * nX + i - 1])
        if name == "xyLS":
            return float(self.rec[6 + 3 * nX + i - 1])
        if name == "utLS":
            return float(self.rec[6 + 4 * nX + i - 1])
        if name == "uyLS":
            return float(self.rec[6 + 4 * nX + nU + i - 1])
        if name == "uxLS":
            return float(self.rec[6 + 4 * nX + 2 * nU + (i - 1) * nX + addr[2] - 1])
        if name == "U":
            o = self.n_params + (i - 1) * n
            return self.rec[o:o + n]
        raise KeyError(addr)

    def addresses(self):
        out = [k for k in _SCALARS if not np.isnan(self.rec[_SCALARS[k]])]
        for k in range(1, self.nX + 1):
            for nm, leaf in (("xNoise", "Noise"), ("xScale", "Scale"), ("xtLS", "LS"), ("xyLS", "LS")):
                if not np.isnan(self[(nm, k, leaf)]):
                    out.append((nm, k, leaf))
        for i in range(1, self.nU + 1):
            out += [("utLS", i, "LS"), ("uyLS", i, "LS"), ("U", i, "U")]
            out += [("uxLS", i, j, "LS") for j in range(1, self.nX + 1)]
        if self.binary:
            out.append("logitT")
        return out


@dataclass
class GPSLCObject:
    """src/types.jl:249-258. `posteriorPacked` [nOuter, n_chains, stride] is the library's buffer; `posteriorSamples`
    lists chain 0's samples as choicemap-like views (the reference runs one chain)."""
    hyperparams: HyperParameters
    priorparams: dict
    SigmaU: Optional[np.ndarray]
    obj: Any
    X: Optional[np.ndarray]
    T: np.ndarray
    Y: np.ndarray
    posteriorSamples: list = field(default_factory=list)
    posteriorPacked: Optional[np.ndarray] = None
    seed: int = 0
    stats: Any = None

    @property
    def n_chains(self):
        return self.posteriorPacked.shape[1]
