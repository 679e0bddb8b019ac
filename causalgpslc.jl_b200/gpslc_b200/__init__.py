"""Python host side of the B200 GP-SLC hot path: mirrors the reference's Julia API (names, argument meaning, error
behaviour) over the C ABI of libgpslc_b200.so. Julia is not available in the build image; julia/CausalGPSLCB200.jl holds
the `ccall` glue a Julia maintainer would use."""
from ._lib import Context, GpslcError, PosDefException, load, LIB_PATH, HOST, DEVICE  # noqa: F401
from .kernel import rbfKernelLog, processCov, cov_build, chol_logpdf, rbf_logpdf  # noqa: F401
from .hyperparameters import getPriorParameters, getHyperParameters, HyperParameters  # noqa: F401
from .utils import (generateSigmaU, removeAdjacent, toMatrix, getN, getNX, getNU, getNumPosteriorSamples,  # noqa: F401
                    extractParameters)
from .types import GPSLCObject, PosteriorSample  # noqa: F401
from .data import prepareData, objectCounts  # noqa: F401
from .inference import Posterior, ChainSampler  # noqa: F401
from .driver import (gpslc, samplePosterior, sampleITE, sampleSATE, summarizeEstimates, ITEDistributions,  # noqa: F401
                     SATEDistributions)
from .prediction import predictCounterfactualEffects, subgroupEffectCurve  # noqa: F401
from .io import saveGPSLCObject, loadGPSLCObject  # noqa: F401
