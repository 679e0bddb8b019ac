"""Oracle restatement of /root/reference/src/kernel.jl. Test infrastructure only (see oracle/__init__.py)."""
import math

import numpy as np


def rbf_kernel_log_scalar(xi, xip, ls):
    """src/kernel.jl:13-19 — ``-sum((Xi .- Xiprime).^2 ./ LS.^2)``; no 1/2 factor, lengthscale squared."""
    xi = np.atleast_1d(np.asarray(xi, dtype=np.float64))
    xip = np.atleast_1d(np.asarray(xip, dtype=np.float64))
    ls = np.asarray(ls, dtype=np.float64)
    assert ls.shape == () or ls.shape[0] == xi.shape[0], "vector lengthscale doesn't match individual"
    return -float(np.sum((xi - xip) ** 2 / ls ** 2))


def rbf_kernel_log_loops(x1, x2, ls):
    """src/kernel.jl:24-42 — the literal scalar double loop (use for small n only)."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    assert x1.shape == x2.shape, "X1 and X2 are different sizes!"
    n = x1.shape[0]
    cov = np.zeros((n, n))
    for i in range(n):
        for ip in range(n):
            cov[i, ip] = rbf_kernel_log_scalar(x1[i], x2[ip], ls)
    return cov


def rbf_kernel_log(x1, x2, ls):
    """Vectorised form of src/kernel.jl:24-42. The per-dimension terms are summed in dimension order, like
    Julia's ``sum`` over a short vector, so it agrees with the literal loop to the last bit for D <= 8 and to
    a few ulp beyond (Julia switches to pairwise summation at 16 elements)."""
    x1 = np.asarray(x1, dtype=np.float64)
    x2 = np.asarray(x2, dtype=np.float64)
    assert x1.shape == x2.shape, "X1 and X2 are different sizes!"
    if x1.ndim == 1:
        x1 = x1[:, None]
        x2 = x2[:, None]
    n, d = x1.shape
    ls = np.broadcast_to(np.asarray(ls, dtype=np.float64), (d,))
    acc = np.zeros((n, n))
    for k in range(d):
        diff = x1[:, k][:, None] - x2[:, k][None, :]
        acc += diff * diff / (ls[k] * ls[k])
    return -acc


def process_cov(log_cov, scale, noise=None):
    """src/kernel.jl:53-59 — ``exp.(logCov) * scale + 1I * noise`` (two-argument form: no noise)."""
    k = np.exp(np.asarray(log_cov, dtype=np.float64)) * scale
    if noise is not None:
        k = k + np.eye(k.shape[0]) * noise
    return k


def logit(p):
    """src/kernel.jl:46"""
    return math.log(p / (1 - p))


def expit(x):
    """src/kernel.jl:49 — ``exp(x) / (1 + exp(x))`` (overflows to NaN for x > ~709 exactly like the reference)."""
    x = np.asarray(x, dtype=np.float64)
    with np.errstate(over="ignore", invalid="ignore"):
        e = np.exp(x)
        return e / (1.0 + e)
