"""Host mirror of /root/reference/src/kernel.jl (`rbfKernelLog`, `processCov`) and of the MvNormal log-density the
reference reaches through Gen `mvnormal`, all computed by the CUDA library."""
import numpy as np

from . import _lib
from ._lib import f64, ptr, HOST

_default_ctx = None


def default_context():
    global _default_ctx
    if _default_ctx is None:
        _default_ctx = _lib.Context(0)
    return _default_ctx


def _features(x, n=None):
    """Reference inputs are n×D matrices, length-n vectors (D=1), or vectors of vectors (src/kernel.jl:24-42)."""
    a = np.asarray(x, dtype=np.float64)
    if a.ndim == 1:
        a = a[:, None]
    return np.asfortranarray(a)


def cov_build(f1, f2, ls, scale, noise=None, ctx=None):
    """Batched `processCov(rbfKernelLog(f1, f2, ls), scale[, noise])` -> array [batch, n, n].
    f1/f2: [n, D] shared or [batch, n, D]; ls: [batch, D] (or [D]); scale/noise: [batch] (or scalars)."""
    ctx = ctx or default_context()
    f1 = np.asarray(f1, dtype=np.float64)
    f2 = np.asarray(f2, dtype=np.float64)
    ls = np.atleast_2d(np.asarray(ls, dtype=np.float64))
    batch, D = ls.shape
    shared = f1.ndim == 2
    n = f1.shape[-2]
    assert f1.shape == f2.shape, "X1 and X2 are different sizes!"  # src/kernel.jl:25
    assert f1.shape[-1] == D
    # library layout: [batch][D][n]
    g1 = np.ascontiguousarray(np.swapaxes(f1, -1, -2))
    g2 = g1 if f2 is f1 else np.ascontiguousarray(np.swapaxes(f2, -1, -2))
    scale = np.ascontiguousarray(np.broadcast_to(np.asarray(scale, dtype=np.float64), (batch,)))
    nz = None if noise is None else np.ascontiguousarray(np.broadcast_to(np.asarray(noise, dtype=np.float64), (batch,)))
    K = np.empty((batch, n, n), dtype=np.float64)  # each n×n block column-major == its transpose row-major (symmetric use)
    ls = np.ascontiguousarray(ls)
    ctx.check(ctx.lib.gpslc_cov_build(ctx.h, HOST, n, batch, D, ptr(g1), ptr(g2), int(shared), ptr(ls), ptr(scale), ptr(nz), ptr(K)))
    # K[b] is column-major: K[b].T is the matrix in NumPy's row-major view
    return np.swapaxes(K, 1, 2)


def rbfKernelLog(X1, X2, LS, ctx=None):
    """src/kernel.jl:24-42. Returns the n×n log-kernel ``-sum_d (X1[i,d]-X2[j,d])^2/LS[d]^2``.
    Computed as log of the device-built exp (the device path never materialises the log form); to stay exact the
    library is asked for scale=1, and the log is taken of a value in (0,1] — for parity tests prefer `cov_build`."""
    f1, f2 = _features(X1), _features(X2)
    D = f1.shape[1]
    ls = np.broadcast_to(np.asarray(LS, dtype=np.float64), (D,))
    K = cov_build(f1, f2, ls[None, :], 1.0, None, ctx)[0]
    with np.errstate(divide="ignore"):
        return np.log(K)


def processCov(logCov, scale, noise=None):
    """src/kernel.jl:53-59 on a host log-kernel (elementwise; kept for API completeness — the hot path fuses it)."""
    k = np.exp(np.asarray(logCov, dtype=np.float64)) * scale
    if noise is not None:
        k = k + np.eye(k.shape[0]) * noise
    return k


def chol_logpdf(K, y, ctx=None):
    """Batched log N(y; 0, K): K [batch, n, n] symmetric, y [batch, n] or [n]. Returns (logpdf, logdet, quad, info)."""
    ctx = ctx or default_context()
    K = np.ascontiguousarray(np.asarray(K, dtype=np.float64))
    if K.ndim == 2:
        K = K[None]
    batch, n, _ = K.shape
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
    y_shared = int(y.ndim == 1)
    lp = np.empty(batch); ld = np.empty(batch); q = np.empty(batch); info = np.empty(batch, dtype=np.int32)
    ctx.check(ctx.lib.gpslc_chol_logpdf(ctx.h, HOST, n, batch, ptr(K), n, ptr(y), y_shared, ptr(lp), ptr(ld), ptr(q), ptr(info)))
    return lp, ld, q, info


def rbf_logpdf(feat, ls, scale, noise, y, ctx=None):
    """Fused build + Cholesky log-density. feat [n, D] shared or [batch, n, D]; ls [batch, D]; y [batch, n] or [n]."""
    ctx = ctx or default_context()
    feat = np.asarray(feat, dtype=np.float64)
    ls = np.ascontiguousarray(np.atleast_2d(np.asarray(ls, dtype=np.float64)))
    batch, D = ls.shape
    shared = feat.ndim == 2
    n = feat.shape[-2]
    g = np.ascontiguousarray(np.swapaxes(feat, -1, -2))
    scale = np.ascontiguousarray(np.broadcast_to(np.asarray(scale, dtype=np.float64), (batch,)))
    noise = np.ascontiguousarray(np.broadcast_to(np.asarray(noise, dtype=np.float64), (batch,)))
    y = np.ascontiguousarray(np.asarray(y, dtype=np.float64))
    y_shared = int(y.ndim == 1)
    lp = np.empty(batch); ld = np.empty(batch); q = np.empty(batch); info = np.empty(batch, dtype=np.int32)
    ctx.check(ctx.lib.gpslc_rbf_logpdf(ctx.h, HOST, n, batch, D, ptr(g), int(shared), ptr(ls), ptr(scale), ptr(noise), ptr(y),
                                       y_shared, ptr(lp), ptr(ld), ptr(q), ptr(info)))
    return lp, ld, q, info
