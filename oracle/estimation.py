"""Oracle restatement of src/likelihood.jl, src/estimation.jl, src/prediction.jl and `summarizeEstimates`
(src/driver.jl:129-149) of the reference. Test infrastructure only (see oracle/__init__.py).

`likelihood_distribution` follows the reference line by line (including its three separate factorisations of the same
matrix: LU, LU, Bunch-Kaufman — SURVEY.md App. B8) so that the restructured CUDA path (one augmented Cholesky, App. A6)
is checked against the reference's own arithmetic, not against itself.
"""
import math

import numpy as np
import scipy.linalg as sla

from . import philox as px
from .kernel import rbf_kernel_log, process_cov


def likelihood_distribution(uyLS, xyLS, tyLS, yNoise, yScale, U, X, T, Y, doT):
    """All four methods of src/likelihood.jl:8-52, 55-94, 97-136, 139-174 (U and/or X may be None)."""
    n = Y.shape[0]
    T = np.asarray(T, dtype=np.float64)
    logk = np.zeros((n, n))
    if U is not None:
        assert U.shape[0] == n
        logk = logk + rbf_kernel_log(U, U, uyLS)            # likelihood.jl:24
    if X is not None:
        assert X.shape[0] == n
        logk = logk + rbf_kernel_log(X, X, xyLS)            # :25
    tyCovLog = rbf_kernel_log(T, T, tyLS)                   # :26
    doTv = np.full(n, float(doT))
    tyCovLogS = rbf_kernel_log(T, doTv, tyLS)               # :27
    tyCovLogSS = rbf_kernel_log(doTv, doTv, tyLS)           # :28

    CovWW = process_cov(logk + tyCovLog, yScale, 0.0)       # :30
    CovWW = np.triu(CovWW) + np.triu(CovWW, 1).T            # Symmetric(): upper triangle (:31)
    CovWWp = CovWW + yNoise * np.eye(n)                     # :32
    CovWWs = process_cov(logk + tyCovLogS, yScale, 0.0)     # :35
    CovWsWs = process_cov(logk + tyCovLogSS, yScale, 0.0)   # :38
    CovWsWs = np.triu(CovWsWs) + np.triu(CovWsWs, 1).T      # :39

    CovWWpInvCovWW = np.linalg.solve(CovWWp, CovWW)         # :42  Symmetric \ Matrix => LU
    CovWWpInvCovWWs = np.linalg.solve(CovWWp, CovWWs)       # :43

    CovC11 = CovWW - CovWW @ CovWWpInvCovWW                 # :46
    CovC12 = CovWWs - CovWW @ CovWWpInvCovWWs               # :47
    CovC21 = CovWWs.T - CovWWs.T @ CovWWpInvCovWW           # :48
    CovC22 = CovWsWs - CovWWs.T @ CovWWpInvCovWWs           # :49
    return Y, CovWW, CovWWs, CovWWp, CovC11, CovC12, CovC21, CovC22


def conditional_ite(uyLS, xyLS, tyLS, yNoise, yScale, U, X, T, Y, doT):
    """src/estimation.jl:36-50"""
    Y, CovWW, CovWWs, CovWWp, C11, C12, C21, C22 = likelihood_distribution(
        uyLS, xyLS, tyLS, yNoise, yScale, U, X, T, Y, doT)
    MeanITE = (CovWWs.T - CovWW) @ sla.solve(CovWWp, Y, assume_a="sym")   # :46  Symmetric \ Vector => Bunch-Kaufman
    CovITE = C11 - C12 - C21 + C22                                         # :47
    return MeanITE, CovITE


def conditional_sate(MeanITE, CovITE):
    """src/estimation.jl:116-121"""
    n = MeanITE.shape[0]
    return float(np.sum(MeanITE)) / n, float(np.sum(CovITE)) / n ** 2


def retained_indices(nBurnIn, stepSize, nOuter):
    """``nBurnIn:stepSize:nOuter`` 1-based, INCLUDING index nBurnIn (src/estimation.jl:72,78; utils.jl:156-161)."""
    return list(range(nBurnIn, nOuter + 1, stepSize))


def extract_parameters(spec, sample):
    """src/utils.jl:92-124 on a packed sample record: (uyLS, xyLS, tyLS, yNoise, yScale, U n×nU).
    U is assembled column-wise (no interleave here — App. B1)."""
    uyLS = np.array([sample[spec.idx("uyLS", i)] for i in range(spec.nU)]) if spec.nU else None
    xyLS = np.array([sample[spec.idx("xyLS", k)] for k in range(spec.nX)]) if spec.nX else None
    U = None
    if spec.nU:
        U = sample[spec.n_params:spec.n_params + spec.nU * spec.n].reshape(spec.nU, spec.n).T.copy()
    return uyLS, xyLS, sample[spec.idx("tyLS")], sample[spec.idx("yNoise")], sample[spec.idx("yScale")], U


def ite_distributions(spec, samples, X, T, Y, doT, nBurnIn, stepSize, jitter):
    """src/estimation.jl:66-86. `samples` is the [nOuter, stride] packed posterior of one chain."""
    idx = retained_indices(nBurnIn, stepSize, samples.shape[0])
    n = Y.shape[0]
    MeanITEs = np.zeros((len(idx), n))
    CovITEs = np.zeros((len(idx), n, n))
    for r, i in enumerate(idx):
        uyLS, xyLS, tyLS, yNoise, yScale, U = extract_parameters(spec, samples[i - 1])
        m, c = conditional_ite(uyLS, xyLS, tyLS, yNoise, yScale, U, X, T, Y, doT)
        MeanITEs[r] = m
        c = np.triu(c) + np.triu(c, 1).T                     # LinearAlgebra.Symmetric(CovITE)  (:82)
        CovITEs[r] = c + np.eye(n) * jitter
    return MeanITEs, CovITEs


def ite_samples(MeanITEs, CovITEs, spp, seed=0, chain=0, dot_index=0):
    """src/estimation.jl:95-109 — `mvnormal(mean, cov)` per draw == mean + chol(cov) z; n × (R*spp), sample-major
    within each mixture component."""
    R, n = MeanITEs.shape
    out = np.zeros((n, R * spp))
    i = 0
    for j in range(R):
        L = sla.cholesky(CovITEs[j], lower=True)
        for _ in range(spp):
            z = px.Stream(seed, chain, i, px.stream_b(px.TAG_ITE, dot_index)).normal_vector(n)
            out[:, i] = MeanITEs[j] + L @ z
            i += 1
    return out


def sate_distributions(MeanITEs, CovITEs):
    """src/estimation.jl:127-140"""
    R = MeanITEs.shape[0]
    ms, vs = np.zeros(R), np.zeros(R)
    for i in range(R):
        ms[i], vs[i] = conditional_sate(MeanITEs[i], CovITEs[i])
    return ms, vs


def sate_samples(MeanSATEs, VarSATEs, spp, seed=0, chain=0, dot_index=0, var_as_std=True):
    """src/estimation.jl:148-163 — `normal(mean, var)`: Gen's second argument is a standard deviation, so the
    reference scales by the VARIANCE (App. B5); `var_as_std=False` gives the statistically intended sqrt."""
    R = MeanSATEs.shape[0]
    out = np.zeros(R * spp)
    i = 0
    for j in range(R):
        sd = VarSATEs[j] if var_as_std else math.sqrt(max(VarSATEs[j], 0.0))
        for _ in range(spp):
            z = px.Stream(seed, chain, i, px.stream_b(px.TAG_SATE, dot_index)).normal()
            out[i] = MeanSATEs[j] + sd * z
            i += 1
    return out


def dot_range(minDoT, maxDoT, fidelity):
    """src/prediction.jl:24-28 — ``minDoT:step:maxDoT`` with step = |max-min|/fidelity (fidelity+1 points)."""
    return np.linspace(minDoT, maxDoT, fidelity + 1)


def summarize_estimates(samples, credible_interval=0.90):
    """src/driver.jl:129-149 — row mean and the two quantiles (Julia `quantile` default == NumPy 'linear')."""
    lo = (1 - credible_interval) / 2
    hi = 1 - lo
    return (np.mean(samples, axis=1), np.quantile(samples, lo, axis=1), np.quantile(samples, hi, axis=1))


# ---- the restructured algebra the CUDA path uses, kept here so tests can check the two agree ----

def conditional_ite_restructured(uyLS, xyLS, tyLS, yNoise, yScale, U, X, T, Y, doT, jitter):
    """One Cholesky of the augmented matrix [[Kp, D], [D', P + jitter I]]; returns (MeanITE, chol(CovITE+jitter I))."""
    n = Y.shape[0]
    T = np.asarray(T, dtype=np.float64)
    logk = np.zeros((n, n))
    if U is not None:
        logk = logk + rbf_kernel_log(U, U, uyLS)
    if X is not None:
        logk = logk + rbf_kernel_log(X, X, xyLS)
    E = yScale * np.exp(logk)
    dt = (T[:, None] - T[None, :]) ** 2 / tyLS ** 2
    Ett = np.exp(-dt)
    a = np.exp(-((T - float(doT)) ** 2) / tyLS ** 2)
    Kp = E * Ett + yNoise * np.eye(n)
    D = E * (a[:, None] - Ett)                      # Kws - Kww
    P = E * (Ett - a[:, None] - a[None, :] + 1.0)   # Kww - Kws - Kws' + Kss
    L = sla.cholesky(Kp, lower=True)
    W = sla.solve_triangular(L, D, lower=True)
    zy = sla.solve_triangular(L, Y, lower=True)
    mean = W.T @ zy
    cov = P - W.T @ W + jitter * np.eye(n)
    return mean, cov
