"""Host mirror of /root/reference/src/hyperparameters.jl and the `HyperParameters` struct (src/types.jl:22-30)."""
from dataclasses import dataclass
from typing import Optional


def getPriorParameters():
    """src/hyperparameters.jl:38-70 — InvGamma(shape=4, scale=4) on every noise/scale/lengthscale."""
    d = {}
    for fam in ["uNoise", "xNoise", "tNoise", "yNoise", "xScale", "tScale", "yScale", "uxLS", "utLS", "xtLS", "uyLS",
                "xyLS", "tyLS"]:
        d[fam + "Shape"] = 4.0
        d[fam + "Scale"] = 4.0
    d["sigmaUNoise"] = 1.0e-13
    d["sigmaUCov"] = 1.0
    d["drift"] = 0.5
    return d


@dataclass
class HyperParameters:
    """src/types.jl:22-30 (mutable)."""
    nU: Optional[int]
    nOuter: int
    nMHInner: Optional[int]
    nESInner: Optional[int]
    nBurnIn: int
    stepSize: int
    predictionCovarianceNoise: float


def getHyperParameters():
    """src/hyperparameters.jl:85-102 (the code's values, not its stale docstring — SURVEY.md App. B7)."""
    return HyperParameters(1, 24, 10, 5, 10, 1, 1e-10)
