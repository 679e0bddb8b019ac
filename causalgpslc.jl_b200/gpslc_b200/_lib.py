"""ctypes binding of libgpslc_b200.so (include/gpslc.h). The library is the product; this module only marshals
NumPy arrays into the C ABI. There is deliberately no fallback: if the shared library or a B200 is missing, calls fail."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GPSLC_LIB_SUFFIX selects a development build of the same sources (build.sh), e.g. "_prof" for the phase-timing build
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libgpslc_b200" + os.environ.get("GPSLC_LIB_SUFFIX", "") + ".so")

HOST, DEVICE = 0, 1

c_double_p = ctypes.POINTER(ctypes.c_double)
c_int_p = ctypes.POINTER(ctypes.c_int)

_lib = None


class GpslcError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"gpslc error {code}: {msg}")
        self.code = code


class PosDefException(GpslcError, ArithmeticError):
    """The analogue of Julia's LinearAlgebra.PosDefException(info): a covariance the reference would hand to `cholesky`
    (Kp in src/likelihood.jl:42-43, CovITE in the `mvnormal` draw of src/estimation.jl:105) is not positive definite."""

    def __init__(self, info, where):
        GpslcError.__init__(self, 3, f"matrix is not positive definite; Cholesky factorization failed (leading minor {info}) in {where}")
        self.info = int(info)


def check_info(info, what, doT=None):
    """Raise PosDefException for the first non-zero entry of a [doT][chain][sample] info array (the reference aborts the whole
    sampleITE / sampleSATE call with a PosDefException in that case; the library reports it per task and carries on)."""
    bad = np.argwhere(np.asarray(info) != 0)
    if bad.size:
        d, c, r = (int(x) for x in bad[0])
        where = f"{what}: doT index {d}" + (f" (doT = {np.atleast_1d(doT)[d]})" if doT is not None else "") + f", chain {c}, retained sample {r}"
        raise PosDefException(int(np.asarray(info)[d, c, r]), where + f"; {len(bad)} of {np.asarray(info).size} tasks failed")


def load():
    """Load the shared library (no GPU needed for this step)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(f"{LIB_PATH} not found: run causalgpslc.jl_b200/build.sh (there is no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)
    vp, i, d_p, i_p, sz, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint64
    sigs = {
        "gpslc_version": (i, []),
        "gpslc_create": (i, [i, ctypes.POINTER(vp)]),
        "gpslc_destroy": (None, [vp]),
        "gpslc_last_error": (ctypes.c_char_p, [vp]),
        "gpslc_synchronize": (i, [vp]),
        "gpslc_stream": (vp, [vp]),
        "gpslc_launch_count": (ctypes.c_ulonglong, [vp]),
        "gpslc_malloc": (i, [vp, sz, ctypes.POINTER(vp)]),
        "gpslc_free": (i, [vp, vp]),
        "gpslc_memcpy_h2d": (i, [vp, vp, vp, sz]),
        "gpslc_memcpy_d2h": (i, [vp, vp, vp, sz]),
        "gpslc_cov_build": (i, [vp, i, i, i, i, d_p, d_p, i, d_p, d_p, d_p, d_p]),
        "gpslc_chol_logpdf": (i, [vp, i, i, i, d_p, i, d_p, i, d_p, d_p, d_p, i_p]),
        "gpslc_rbf_logpdf": (i, [vp, i, i, i, i, d_p, i, d_p, d_p, d_p, d_p, i, d_p, d_p, d_p, i_p]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def declared_symbols():
    """Every entry point include/gpslc.h declares (used by the CPU-side export test)."""
    import re
    hdr = os.path.join(os.path.dirname(os.path.dirname(_HERE)), "include", "gpslc.h")
    txt = open(hdr).read()
    return sorted(set(re.findall(r"\b(gpslc_[a-z0-9_]+)\s*\(", txt)))


def ptr(a):
    """void* of a NumPy array (or None)."""
    if a is None:
        return None
    return ctypes.c_void_p(a.ctypes.data)


def f64(a, order="F"):
    return np.ascontiguousarray(a, dtype=np.float64) if order == "C" else np.asfortranarray(a, dtype=np.float64)


class Context:
    """One per GPU (gpslc_ctx)."""

    def __init__(self, device=0):
        self.lib = load()
        h = ctypes.c_void_p()
        rc = self.lib.gpslc_create(device, ctypes.byref(h))
        if rc != 0:
            raise GpslcError(rc, "gpslc_create failed (no sm_100 GPU visible? there is no CPU fallback)")
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.lib.gpslc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def check(self, rc):
        if rc != 0:
            raise GpslcError(rc, self.lib.gpslc_last_error(self.h).decode())

    @property
    def stream(self):
        return self.lib.gpslc_stream(self.h)

    @property
    def launches(self):
        return int(self.lib.gpslc_launch_count(self.h))

    def synchronize(self):
        self.check(self.lib.gpslc_synchronize(self.h))
