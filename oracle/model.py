"""Oracle restatement of the GP-SLC generative model: src/model.jl, src/model_likelihood.jl, src/model_prior.jl,
src/utils.jl:17-33,60-64 and src/hyperparameters.jl of the reference. Test infrastructure only.

The eight Gen models (src/model.jl:11-130) differ only in which parents exist (U, X) and in whether T is real or
binary, so one ``ModelSpec`` with three flags restates all of them. Densities that the reference obtains from
Gen/Distributions (`inv_gamma`, `mvnormal`, `bernoulli`) are written from their definitions.
"""
import math
from dataclasses import dataclass, field

import numpy as np
import scipy.linalg as sla

from .kernel import rbf_kernel_log, process_cov, expit

LOG_2PI = math.log(2.0 * math.pi)

# src/hyperparameters.jl:39-69
PRIOR_KEYS = [
    "uNoiseShape", "uNoiseScale", "xNoiseShape", "xNoiseScale", "tNoiseShape", "tNoiseScale",
    "yNoiseShape", "yNoiseScale", "xScaleShape", "xScaleScale", "tScaleShape", "tScaleScale",
    "yScaleShape", "yScaleScale", "uxLSShape", "uxLSScale", "utLSShape", "utLSScale",
    "xtLSShape", "xtLSScale", "uyLSShape", "uyLSScale", "xyLSShape", "xyLSScale",
    "tyLSShape", "tyLSScale", "sigmaUNoise", "sigmaUCov", "drift",
]


def get_prior_parameters():
    """src/hyperparameters.jl:38-70"""
    d = {k: 4.0 for k in PRIOR_KEYS[:26]}
    d["sigmaUNoise"] = 1.0e-13
    d["sigmaUCov"] = 1.0
    d["drift"] = 0.5
    return d


def generate_sigma_u(counts, eps=1e-13, cov=1.0):
    """src/utils.jl:17-33 — identity, each object's block set to `cov`, then the whole diagonal to 1+eps."""
    counts = [int(c) for c in counts]
    n = sum(counts)
    s = np.eye(n)
    i = 0
    for m in counts:
        s[i:i + m, i:i + m] = np.ones((m, m)) * cov
        i += m
    s[np.diag_indices(n)] = 1 + eps
    return s


def to_matrix(vectors, n, m):
    """src/utils.jl:60-64 — ``reshape(permutedims(hcat(X...)), (n, m))`` with Julia's column-major reshape.
    For a list of m length-n vectors this INTERLEAVES individuals and dimensions unless m == 1
    (SURVEY.md App. B1); for an n×m Matrix input it is the identity."""
    if isinstance(vectors, np.ndarray) and vectors.ndim == 2:
        # hcat(M...) splats the scalars into a 1 x (n*m) row in column-major order
        flat = vectors.reshape(-1, order="F")
        return flat.reshape((n, m), order="F")
    h = np.stack([np.asarray(v, dtype=np.float64) for v in vectors], axis=1)  # hcat: len x count
    p = h.T  # permutedims: count x len
    return p.reshape(-1, order="F").reshape((n, m), order="F")


@dataclass
class ModelSpec:
    """Which of the eight models of src/model.jl:11-130 is meant, plus its sizes."""
    n: int
    nU: int  # 0 == `nothing`
    nX: int  # 0 == `nothing`
    binary: bool
    u_layout_reference: bool = True  # App. B1: reproduce toMatrix interleave (identity when nU == 1)

    @property
    def has_u(self):
        return self.nU > 0

    @property
    def has_x(self):
        return self.nX > 0

    # ---- packed parameter layout ----
    @property
    def n_params(self):
        return 6 + 4 * self.nX + 2 * self.nU + self.nU * self.nX

    def idx(self, name, i=0, j=0):
        nX, nU = self.nX, self.nU
        base = {"uNoise": 0, "tNoise": 1, "yNoise": 2, "tyLS": 3, "tScale": 4, "yScale": 5}
        if name in base:
            return base[name]
        if name == "xNoise":
            return 6 + i
        if name == "xScale":
            return 6 + nX + i
        if name == "xtLS":
            return 6 + 2 * nX + i
        if name == "xyLS":
            return 6 + 3 * nX + i
        if name == "utLS":
            return 6 + 4 * nX + i
        if name == "uyLS":
            return 6 + 4 * nX + nU + i
        if name == "uxLS":
            return 6 + 4 * nX + 2 * nU + i * nX + j
        raise KeyError(name)

    def active_params(self):
        """(name, i, j) of every latent hyperparameter the model actually traces, by variant
        (src/model.jl:12-21, 31-36, 46-52, 62-64)."""
        out = []
        if self.has_u:
            out.append(("uNoise", 0, 0))
        if self.has_u or self.has_x:
            out.append(("tNoise", 0, 0))
        out.append(("yNoise", 0, 0))
        out.append(("tyLS", 0, 0))
        if self.has_u or self.has_x:
            out.append(("tScale", 0, 0))
        out.append(("yScale", 0, 0))
        if self.has_u and self.has_x:
            out += [("xNoise", k, 0) for k in range(self.nX)]
            out += [("xScale", k, 0) for k in range(self.nX)]
        if self.has_x:
            out += [("xtLS", k, 0) for k in range(self.nX)]
            out += [("xyLS", k, 0) for k in range(self.nX)]
        if self.has_u:
            out += [("utLS", i, 0) for i in range(self.nU)]
            out += [("uyLS", i, 0) for i in range(self.nU)]
            if self.has_x:
                out += [("uxLS", i, j) for i in range(self.nU) for j in range(self.nX)]
        return out

    def mh_sites(self):
        """One MH sweep's site order — src/inference.jl:23-44 (full), 76-87 (no X), 127-137 (no U),
        158-160 (neither; those three are visited once per OUTER iteration, not per nMHInner)."""
        s = []
        if self.has_u:
            s.append(("uNoise", 0, 0))
        if self.has_u or self.has_x:
            s.append(("tNoise", 0, 0))
        s.append(("yNoise", 0, 0))
        s.append(("tyLS", 0, 0))
        if self.has_u:
            for k in range(self.nU):
                s.append(("utLS", k, 0))
                s.append(("uyLS", k, 0))
                for l in range(self.nX):
                    s.append(("uxLS", k, l))
        if self.has_x:
            for k in range(self.nX):
                if self.has_u:
                    s.append(("xNoise", k, 0))
                s.append(("xtLS", k, 0))
                s.append(("xyLS", k, 0))
                if self.has_u:
                    s.append(("xScale", k, 0))
        if self.has_u or self.has_x:
            s.append(("tScale", 0, 0))
        s.append(("yScale", 0, 0))
        return s

    def prior_key(self, name):
        return name + "Shape", name + "Scale"


# ---------------------------------------------------------------------------------------------- densities

def ig_logpdf(x, shape, scale):
    """log InvGamma(x; shape a, scale b) = a log b - lgamma(a) - (a+1) log x - b/x; -inf for x <= 0."""
    if not (x > 0.0):
        return -math.inf
    return shape * math.log(scale) - math.lgamma(shape) - (shape + 1.0) * math.log(x) - scale / x


def mvn_logpdf_chol(y, K):
    """log N(y; 0, K) through a Cholesky factor, as Distributions.MvNormal/PDMats does
    (SURVEY.md §2.1 (ii)). Raises numpy.linalg.LinAlgError where the reference raises PosDefException."""
    L = sla.cholesky(K, lower=True, check_finite=False)
    z = sla.solve_triangular(L, y, lower=True, check_finite=False)
    n = y.shape[0]
    return -0.5 * (n * LOG_2PI + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(z @ z))


def sigma_u_blocks(counts):
    return [int(c) for c in counts]


def u_prior_quad_logdet(u, counts, eps, cov):
    """Closed form of u' SigmaU^-1 u and log det SigmaU for the block matrix of src/utils.jl:17-33.
    Each block is c*11' + d*I with d = fl(1+eps) - c ("numerically singular by construction")."""
    d = (1.0 + eps) - cov
    quad = 0.0
    logdet = 0.0
    i = 0
    for m in counts:
        blk = u[i:i + m]
        mean = float(np.sum(blk)) / m
        dev = blk - mean
        quad += float(np.sum(dev * dev)) / d + m * mean * mean / (d + m * cov)
        logdet += (m - 1) * math.log(d) + math.log(d + m * cov)
        i += m
    return quad, logdet


def u_prior_logpdf(u, u_noise, counts, eps, cov):
    """log N(u; 0, uNoise * SigmaU) (src/model_likelihood.jl:4-10, src/model_prior.jl:27-30), closed form."""
    n = u.shape[0]
    quad, logdet = u_prior_quad_logdet(u, counts, eps, cov)
    return -0.5 * (n * LOG_2PI + n * math.log(u_noise) + logdet + quad / u_noise)


def u_prior_logpdf_data(data, u, u_noise):
    """log N(u; 0, uNoise * SigmaU) for the SigmaU of `data`: closed form for the block matrix, Cholesky for a dense one
    (L_S^-1 u with L_S = chol(SigmaU); chol(uNoise*SigmaU) = sqrt(uNoise) L_S)."""
    if data.sigma_u_dense is None:
        return u_prior_logpdf(u, u_noise, data.counts, data.eps, data.cov)
    L = data.sigma_u_chol()
    z = sla.solve_triangular(L, u, lower=True, check_finite=False)
    n = u.shape[0]
    return -0.5 * (n * LOG_2PI + n * math.log(u_noise) + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(z @ z) / u_noise)


def bernoulli_logpmf(t, logit_t):
    """Σ_i log Bernoulli(T_i; expit(logitT_i)) (src/model_prior.jl:22-24), in the overflow-safe form."""
    x = np.asarray(logit_t, dtype=np.float64)
    t = np.asarray(t, dtype=np.float64)
    # log p = -softplus(-x), log(1-p) = -softplus(x)
    sp_pos = np.logaddexp(0.0, x)
    sp_neg = np.logaddexp(0.0, -x)
    return float(np.sum(np.where(t > 0.5, -sp_neg, -sp_pos)))


# ---------------------------------------------------------------------------------------------- model state

@dataclass
class ModelData:
    """Observed data + confounder structure. T is float (0/1 for binary)."""
    spec: ModelSpec
    X: np.ndarray  # n x nX (or None)
    T: np.ndarray  # n
    Y: np.ndarray  # n
    counts: list = field(default_factory=list)  # object sizes (has_u only)
    eps: float = 1e-13
    cov: float = 1.0
    prior: dict = field(default_factory=get_prior_parameters)
    sigma_u_dense: np.ndarray = None  # an unstructured SigmaU (src/driver.jl:59-69 accepts any matrix): Cholesky path

    def sigma_u_chol(self):
        """lower Cholesky factor of the dense SigmaU (what generateU's mvnormal computes, src/model_prior.jl:27-30)"""
        if getattr(self, "_ls", None) is None:
            self._ls = sla.cholesky(self.sigma_u_dense, lower=True, check_finite=False)
        return self._ls


@dataclass
class State:
    """One chain's latent state: packed hyperparameters, U vectors (nU x n), logitT (binary), and the model's
    own X when it is not observed (App. B3)."""
    theta: np.ndarray
    U: np.ndarray  # nU x n
    logitT: np.ndarray = None
    Xmodel: np.ndarray = None  # n x nX, only for the no-U variants in reference-faithful mode

    def copy(self):
        return State(self.theta.copy(), self.U.copy(),
                     None if self.logitT is None else self.logitT.copy(),
                     None if self.Xmodel is None else self.Xmodel.copy())


def effective_u(spec, U):
    """The n x nU matrix the model's kernels see (src/model_likelihood.jl:7)."""
    if spec.nU == 0:
        return None
    if spec.u_layout_reference:
        return to_matrix([U[i] for i in range(spec.nU)], spec.n, spec.nU)
    return U.T.copy()


def effective_uxls(spec, theta):
    """nX x nU lengthscale matrix (src/model_prior.jl:110: ``toMatrix(uxLS, nX, nU)``), row k feeds X_k."""
    nU, nX = spec.nU, spec.nX
    vecs = [np.array([theta[spec.idx("uxLS", i, j)] for j in range(nX)]) for i in range(nU)]
    if spec.u_layout_reference:
        return to_matrix(vecs, nX, nU)
    return np.stack(vecs, axis=1)


def model_x(data, st):
    """The X the T and Y kernels use: the data, except in the no-U variants where the reference never observes X
    (src/inference.jl:116-118, 310-316) and Xmodel holds the `generate`-time N(0,I) draw."""
    if st.Xmodel is not None:
        return st.Xmodel
    return data.X


def factor_cov(data, st, f):
    """Covariance of factor f: 0..nX-1 -> X_k (model_likelihood.jl:13-22), nX -> T / logitT (25-80),
    nX+1 -> Y (83-120)."""
    spec = data.spec
    th = st.theta
    nX, nU = spec.nX, spec.nU
    Ue = effective_u(spec, st.U)
    if f < nX:
        ls = effective_uxls(spec, th)[f, :]
        logk = rbf_kernel_log(Ue, Ue, ls)
        return process_cov(logk, th[spec.idx("xScale", f)], th[spec.idx("xNoise", f)])
    Xm = model_x(data, st)
    if f == nX:
        logk = np.zeros((spec.n, spec.n))
        if spec.has_u:
            logk = logk + rbf_kernel_log(Ue, Ue, np.array([th[spec.idx("utLS", i)] for i in range(nU)]))
        if spec.has_x:
            logk = logk + rbf_kernel_log(Xm, Xm, np.array([th[spec.idx("xtLS", k)] for k in range(nX)]))
        return process_cov(logk, th[spec.idx("tScale")], th[spec.idx("tNoise")])
    logk = np.zeros((spec.n, spec.n))
    if spec.has_u:
        logk = logk + rbf_kernel_log(Ue, Ue, np.array([th[spec.idx("uyLS", i)] for i in range(nU)]))
    if spec.has_x:
        logk = logk + rbf_kernel_log(Xm, Xm, np.array([th[spec.idx("xyLS", k)] for k in range(nX)]))
    logk = logk + rbf_kernel_log(data.T, data.T, th[spec.idx("tyLS")])
    return process_cov(logk, th[spec.idx("yScale")], th[spec.idx("yNoise")])


def factor_target(data, st, f):
    spec = data.spec
    if f < spec.nX:
        return data.X[:, f]
    if f == spec.nX:
        return st.logitT if spec.binary else data.T
    return data.Y


def factor_exists(spec, f):
    """X_k factors exist only with U (otherwise X ~ N(0,I), model_prior.jl:175-181); the T factor only with a
    parent (otherwise T/logitT ~ N(0,I), model_prior.jl:187-200)."""
    if f < spec.nX:
        return spec.has_u
    if f == spec.nX:
        return spec.has_u or spec.has_x
    return True


def std_normal_logpdf(v):
    v = np.asarray(v, dtype=np.float64)
    return -0.5 * (v.shape[0] * LOG_2PI + float(v @ v))


def factor_logpdf(data, st, f):
    if factor_exists(data.spec, f):
        return mvn_logpdf_chol(factor_target(data, st, f), factor_cov(data, st, f))
    # parentless variable: identity covariance
    if f < data.spec.nX:
        x = st.Xmodel[:, f] if st.Xmodel is not None else data.X[:, f]
        return std_normal_logpdf(x)
    return std_normal_logpdf(factor_target(data, st, f))


def log_joint_terms(data, st):
    """Every term of the log joint in model order (src/model.jl): dict name -> value."""
    spec = data.spec
    pr = data.prior
    terms = {}
    for (name, i, j) in spec.active_params():
        a, b = spec.prior_key(name)
        terms[(name, i, j)] = ig_logpdf(st.theta[spec.idx(name, i, j)], pr[a], pr[b])
    if spec.has_u:
        for i in range(spec.nU):
            terms[("U", i)] = u_prior_logpdf_data(data, st.U[i], st.theta[spec.idx("uNoise")])
    for f in range(spec.nX + 2):
        terms[("factor", f)] = factor_logpdf(data, st, f)
    if spec.binary:
        terms[("bernoulli",)] = bernoulli_logpmf(data.T, st.logitT)
    return terms


def log_joint(data, st):
    return float(sum(log_joint_terms(data, st).values()))
