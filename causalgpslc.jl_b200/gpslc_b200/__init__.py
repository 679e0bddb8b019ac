"""Python host side of the B200 GP-SLC hot path: mirrors the reference's Julia API over the C ABI of
libgpslc_b200.so. (Julia is not available in the build image; julia/CausalGPSLCB200.jl is the `ccall` glue.)"""
from ._lib import Context, GpslcError, load, LIB_PATH, HOST, DEVICE  # noqa: F401
from .kernel import rbfKernelLog, processCov, cov_build, chol_logpdf, rbf_logpdf  # noqa: F401
