"""Host mirror of /root/reference/src/estimation.jl (+ likelihood.jl): ITE / SATE distributions and samples computed by
the CUDA library (csrc/estimation.cu) from the sampler's packed posterior samples."""
import ctypes

import numpy as np

from ._lib import ptr, HOST
from .inference import GpslcData, _bind


def retained_indices(nBurnIn, stepSize, nOuter):
    """``nBurnIn:stepSize:nOuter`` (1-based, includes nBurnIn; src/estimation.jl:72,78) as 0-based indices."""
    return np.arange(nBurnIn, nOuter + 1, stepSize, dtype=np.int32) - 1


def _bind_est(lib):
    if getattr(lib, "_est_bound", False):
        return
    vp, i, dbl, u64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_double, ctypes.c_uint64
    P = ctypes.POINTER
    lib.gpslc_ite.restype = i
    lib.gpslc_ite.argtypes = [vp, i, P(GpslcData), vp, i, i, i, vp, i, vp, i, dbl, i, u64, i, vp, vp, vp, vp]
    lib.gpslc_sate.restype = i
    lib.gpslc_sate.argtypes = [vp, i, P(GpslcData), vp, i, i, i, vp, i, vp, i, dbl, i, u64, i, i, vp, vp, vp, vp]
    lib.gpslc_ite_slice.restype = i
    lib.gpslc_ite_slice.argtypes = [vp, i, P(GpslcData), vp, i, i, i, vp, i, vp, i, i, dbl, i, u64, i, vp, vp, vp, vp]
    lib.gpslc_sate_slice.restype = i
    lib.gpslc_sate_slice.argtypes = [vp, i, P(GpslcData), vp, i, i, i, vp, i, vp, i, i, dbl, i, u64, i, i, vp, vp, vp, vp]
    lib.gpslc_ite_summary.restype = i
    lib.gpslc_ite_summary.argtypes = [vp, i, P(GpslcData), vp, i, i, i, vp, i, vp, i, i, dbl, i, u64, i, dbl, vp, vp]
    lib.gpslc_summarize.restype = i
    lib.gpslc_summarize.argtypes = [vp, i, vp, i, i, i, dbl, vp]
    lib.gpslc_ite_subset_summary.restype = i
    lib.gpslc_ite_subset_summary.argtypes = [vp, i, P(GpslcData), vp, i, i, i, vp, i, vp, i, i, dbl, i, u64, i, vp, dbl, vp, vp, vp]
    lib.gpslc_subset_mean.restype = i
    lib.gpslc_subset_mean.argtypes = [vp, i, vp, i, i, i, vp, vp]
    lib._est_bound = True


def _data_struct(X, T, Y, nU):
    keep = {"T": np.ascontiguousarray(np.asarray(T), dtype=np.float64), "Y": np.ascontiguousarray(Y, dtype=np.float64)}
    d = GpslcData()
    d.n = keep["T"].shape[0]
    d.nU = int(nU or 0)
    d.binary = int(np.asarray(T).dtype == np.bool_)
    if X is not None:
        keep["X"] = np.asfortranarray(np.asarray(X, dtype=np.float64))
        d.nX = keep["X"].shape[1]
        d.X = keep["X"].ctypes.data
    d.T = keep["T"].ctypes.data
    d.Y = keep["Y"].ctypes.data
    return d, keep


def ite(samples, X, T, Y, nU, doT, ret_idx, jitter, spp, seed=0, chain_offset=0, want_cov=False, want_samples=True, ctx=None,
        dot_offset=0, out=None):
    """samples [n_outer, n_chains, stride]; doT scalar or array. Returns dict(mean [D,C,R,n], cov [D,C,R,n,n] or None,
    samples [D,C,R*spp,n] or None, info [D,C,R]). dot_offset: global index of doT[0] when doT is a slice of a sharded sweep.
    out: optional dict of preallocated C-contiguous float64 arrays "mean" / "samples" of exactly those shapes (e.g. views of
    pinned host memory that a serving loop reuses); the library writes into them instead of fresh arrays."""
    from .kernel import default_context
    ctx = ctx or default_context()
    _bind(ctx.lib); _bind_est(ctx.lib)
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    n_outer, C, stride = samples.shape
    d, keep = _data_struct(X, T, Y, nU)
    doT = np.ascontiguousarray(np.atleast_1d(np.asarray(doT, dtype=np.float64)))
    ret = np.ascontiguousarray(ret_idx, dtype=np.int32)
    D, R, n = doT.shape[0], ret.shape[0], d.n
    def _buf(key, shape):
        if out is not None and key in out:
            b = out[key]
            if b.shape != shape or b.dtype != np.float64 or not b.flags["C_CONTIGUOUS"]:
                raise ValueError(f"out[{key!r}] must be a C-contiguous float64 array of shape {shape}")
            return b
        return np.empty(shape)
    mean = _buf("mean", (D, C, R, n))
    cov = np.empty((D, C, R, n, n)) if want_cov else None
    smp = _buf("samples", (D, C, R * spp, n)) if (want_samples and spp > 0) else None
    info = np.empty((D, C, R), dtype=np.int32)
    if dot_offset:
        ctx.check(ctx.lib.gpslc_ite_slice(ctx.h, HOST, ctypes.byref(d), ptr(samples), n_outer, C, stride, ptr(ret), R, ptr(doT), D,
                                          int(dot_offset), float(jitter), int(spp), int(seed), int(chain_offset), ptr(mean), ptr(cov),
                                          ptr(smp), ptr(info)))
    else:
        ctx.check(ctx.lib.gpslc_ite(ctx.h, HOST, ctypes.byref(d), ptr(samples), n_outer, C, stride, ptr(ret), R, ptr(doT), D,
                                    float(jitter), int(spp), int(seed), int(chain_offset), ptr(mean), ptr(cov), ptr(smp), ptr(info)))
    return {"mean": mean, "cov": cov, "samples": smp, "info": info}


def ite_summary(samples, X, T, Y, nU, doT, ret_idx, jitter, spp, seed=0, chain_offset=0, credible_interval=0.90, ctx=None, dot_offset=0):
    """gpslc_ite_summary: the ITE draws of every (doT, chain) are summarised on the device (Mean / LowerBound / UpperBound per
    individual over the R*spp draws) and never shipped to the host. Returns (summary [D, C, n, 3], info [D, C, R])."""
    from .kernel import default_context
    ctx = ctx or default_context()
    _bind(ctx.lib); _bind_est(ctx.lib)
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    n_outer, C, stride = samples.shape
    d, keep = _data_struct(X, T, Y, nU)
    doT = np.ascontiguousarray(np.atleast_1d(np.asarray(doT, dtype=np.float64)))
    ret = np.ascontiguousarray(ret_idx, dtype=np.int32)
    D, R, n = doT.shape[0], ret.shape[0], d.n
    out = np.empty((D, C, n, 3))
    info = np.empty((D, C, R), dtype=np.int32)
    ctx.check(ctx.lib.gpslc_ite_summary(ctx.h, HOST, ctypes.byref(d), ptr(samples), n_outer, C, stride, ptr(ret), R, ptr(doT), D,
                                        int(dot_offset), float(jitter), int(spp), int(seed), int(chain_offset),
                                        float(credible_interval), ptr(out), ptr(info)))
    return out, info


def ite_subset_summary(samples, X, T, Y, nU, doT, ret_idx, jitter, spp, mask, seed=0, chain_offset=0, credible_interval=0.90, ctx=None,
                       dot_offset=0):
    """gpslc_ite_subset_summary — the subgroup workflow of docs/src/index.md:101-114 fused on the device: the ITE draws of every doT
    are averaged over the individuals selected by `mask` and summarised over the draws. Returns (summary [C, D, 3] = Mean / LowerBound /
    UpperBound per doT, sate [C, R*spp, D], info [D, C, R])."""
    from .kernel import default_context
    ctx = ctx or default_context()
    _bind(ctx.lib); _bind_est(ctx.lib)
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    n_outer, C, stride = samples.shape
    d, keep = _data_struct(X, T, Y, nU)
    doT = np.ascontiguousarray(np.atleast_1d(np.asarray(doT, dtype=np.float64)))
    ret = np.ascontiguousarray(ret_idx, dtype=np.int32)
    D, R, n = doT.shape[0], ret.shape[0], d.n
    mask = np.ascontiguousarray(np.asarray(mask).astype(bool), dtype=np.uint8)
    if mask.shape != (n,):
        raise ValueError("mask must have one entry per individual")
    out = np.empty((C, D, 3)); sub = np.empty((C, R * spp, D))
    info = np.empty((D, C, R), dtype=np.int32)
    ctx.check(ctx.lib.gpslc_ite_subset_summary(ctx.h, HOST, ctypes.byref(d), ptr(samples), n_outer, C, stride, ptr(ret), R, ptr(doT), D,
                                               int(dot_offset), float(jitter), int(spp), int(seed), int(chain_offset), ptr(mask),
                                               float(credible_interval), ptr(sub), ptr(out), ptr(info)))
    return out, sub, info


def subset_mean(samples, mask, ctx=None):
    """gpslc_subset_mean: `mean(ite[:, idx, :], dims=2)` for samples [batch, m, n] (the layout gpslc_ite writes) -> [batch, m]."""
    from .kernel import default_context
    ctx = ctx or default_context()
    _bind(ctx.lib); _bind_est(ctx.lib)
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    B, m, n = samples.shape
    mask = np.ascontiguousarray(np.asarray(mask).astype(bool), dtype=np.uint8)
    if mask.shape != (n,):
        raise ValueError("mask must have one entry per individual")
    out = np.empty((B, m))
    ctx.check(ctx.lib.gpslc_subset_mean(ctx.h, HOST, ptr(samples), B, m, n, ptr(mask), ptr(out)))
    return out


def sate(samples, X, T, Y, nU, doT, ret_idx, jitter, spp, seed=0, chain_offset=0, var_as_std=True, ctx=None, dot_offset=0):
    """Returns dict(mean [D,C,R], var [D,C,R], samples [D,C,R*spp], info)."""
    from .kernel import default_context
    ctx = ctx or default_context()
    _bind(ctx.lib); _bind_est(ctx.lib)
    samples = np.ascontiguousarray(samples, dtype=np.float64)
    n_outer, C, stride = samples.shape
    d, keep = _data_struct(X, T, Y, nU)
    doT = np.ascontiguousarray(np.atleast_1d(np.asarray(doT, dtype=np.float64)))
    ret = np.ascontiguousarray(ret_idx, dtype=np.int32)
    D, R = doT.shape[0], ret.shape[0]
    mean = np.empty((D, C, R)); var = np.empty((D, C, R)); smp = np.empty((D, C, R * spp)); info = np.empty((D, C, R), dtype=np.int32)
    if dot_offset:
        ctx.check(ctx.lib.gpslc_sate_slice(ctx.h, HOST, ctypes.byref(d), ptr(samples), n_outer, C, stride, ptr(ret), R, ptr(doT), D,
                                           int(dot_offset), float(jitter), int(spp), int(seed), int(chain_offset),
                                           int(bool(var_as_std)), ptr(mean), ptr(var), ptr(smp), ptr(info)))
    else:
        ctx.check(ctx.lib.gpslc_sate(ctx.h, HOST, ctypes.byref(d), ptr(samples), n_outer, C, stride, ptr(ret), R, ptr(doT), D,
                                     float(jitter), int(spp), int(seed), int(chain_offset), int(bool(var_as_std)), ptr(mean), ptr(var),
                                     ptr(smp), ptr(info)))
    return {"mean": mean, "var": var, "samples": smp, "info": info}


def summarize(samples, credible_interval=0.90, ctx=None):
    """gpslc_summarize: samples [batch, m, n] (or [m, n]) -> [batch, n, 3] (or [n, 3]) = Mean, LowerBound, UpperBound per
    individual; the device-side body of summarizeEstimates (src/driver.jl:129-149)."""
    from .kernel import default_context
    ctx = ctx or default_context()
    _bind(ctx.lib); _bind_est(ctx.lib)
    s = np.ascontiguousarray(samples, dtype=np.float64)
    squeeze = s.ndim == 2
    if squeeze:
        s = s[None]
    batch, m, n = s.shape
    out = np.empty((batch, n, 3))
    ctx.check(ctx.lib.gpslc_summarize(ctx.h, HOST, ptr(s), batch, m, n, float(credible_interval), ptr(out)))
    return out[0] if squeeze else out
