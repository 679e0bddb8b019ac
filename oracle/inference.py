"""Oracle restatement of src/proposal.jl and src/inference.jl (all eight `Posterior` methods) on top of the
recollected Gen 0.4.4 semantics of `generate`, `mh` and `elliptical_slice` (SURVEY.md App. C — parity unpinned).
Test infrastructure only (see oracle/__init__.py).

Two cost models compute the SAME chain (same Philox streams, same accept decisions up to round-off):
  * ``mode="faithful"``    — every `mh`/`elliptical_slice` evaluation re-scores the whole model like Gen's dynamic
                             DSL `update` does (nX+5 kernel builds, nU+nX+2 Choleskys). This is the
                             reference's cost and is what the CPU baseline times.
  * ``mode="incremental"`` — only the terms that depend on the changed address are recomputed (App. A4).
"""
import math

import numpy as np
import scipy.linalg as sla

from . import philox as px
from .kernel import rbf_kernel_log, process_cov
from .model import (ModelSpec, ModelData, State, ig_logpdf, u_prior_logpdf, u_prior_logpdf_data, u_prior_quad_logdet, factor_logpdf,
                    factor_cov, factor_exists, bernoulli_logpmf, log_joint, effective_uxls, LOG_2PI)


# ------------------------------------------------------------------------------------------------ proposal

def proposal_params(cur, variance):
    """src/proposal.jl:32-41 — InvGamma with mean `cur` and variance `variance`."""
    shape = (cur * cur / variance) + 2
    scale = cur * (shape - 1)
    return shape, scale


def param_factors(spec, name, i, j):
    """Factors whose covariance depends on a hyperparameter. Empty for uNoise."""
    nX = spec.nX
    if name == "uNoise":
        return []
    if name in ("tNoise", "tScale", "utLS", "xtLS"):
        return [nX]
    if name in ("yNoise", "yScale", "tyLS", "uyLS", "xyLS"):
        return [nX + 1]
    if name in ("xNoise", "xScale"):
        return [i]
    if name == "uxLS":
        if spec.u_layout_reference:
            return [(i + j * spec.nU) % nX]  # row of toMatrix(uxLS, nX, nU) holding element (i, j) — App. B1
        return [j]
    raise KeyError(name)


# ------------------------------------------------------------------------------------------------ scoring

class Scorer:
    """Caches the per-term log densities of one chain so both cost models share the control flow."""

    def __init__(self, data, st, mode):
        self.data = data
        self.mode = mode
        self.n_chol = 0  # factorizations performed (cost accounting)
        self.n_build = 0
        spec = data.spec
        self.nF = spec.nX + 2
        self.f_lp = [self._factor(st, f) for f in range(self.nF)]
        self.u_lp = [self._uprior(st, i) for i in range(spec.nU)]

    def _factor(self, st, f):
        spec = self.data.spec
        if factor_exists(spec, f):
            self.n_chol += 1
            self.n_build += (1 if f < spec.nX else (spec.has_u + spec.has_x + (f == spec.nX + 1)))
        return factor_logpdf(self.data, st, f)

    def _uprior(self, st, i):
        d = self.data
        if self.mode == "faithful":
            # the reference factorises uNoise*SigmaU on every update (model_prior.jl:27-30): the work is executed here so that the
            # CPU baseline pays for it; the VALUE comes from the closed form (the dense route is only 1e-6 accurate)
            self.n_chol += 1
            if d.sigma_u_dense is None:
                if getattr(self, "_sigma_u", None) is None:
                    from .model import generate_sigma_u
                    self._sigma_u = generate_sigma_u(d.counts, d.eps, d.cov)
                try:
                    sla.cholesky(st.theta[d.spec.idx("uNoise")] * self._sigma_u, lower=True, check_finite=False)
                except np.linalg.LinAlgError:
                    pass
        return u_prior_logpdf_data(d, st.U[i], st.theta[d.spec.idx("uNoise")])

    def rescore(self, st_new, factors, u_terms):
        """Return (delta_loglik_terms, new_cache) for a proposed state. In faithful mode everything is recomputed
        (and must agree with the cache for unchanged terms — asserted loosely)."""
        spec = self.data.spec
        if self.mode == "faithful":
            new_f = [self._factor(st_new, f) for f in range(self.nF)]
            new_u = [self._uprior(st_new, i) for i in range(spec.nU)]
        else:
            new_f = list(self.f_lp)
            new_u = list(self.u_lp)
            for f in factors:
                new_f[f] = self._factor(st_new, f)
            for i in u_terms:
                new_u[i] = self._uprior(st_new, i)
        delta = (sum(new_f) - sum(self.f_lp)) + (sum(new_u) - sum(self.u_lp))
        return delta, (new_f, new_u)

    def commit(self, cache):
        self.f_lp, self.u_lp = cache


# ------------------------------------------------------------------------------------------------ init

def sample_u_prior(stream, n, counts, u_noise, eps, cov):
    """U ~ N(0, uNoise*SigmaU) exploiting the block structure: sqrt(uNoise)*(sqrt(cov) z_obj + sqrt(d) z_i)."""
    d = (1.0 + eps) - cov
    z = stream.normal_vector(n + len(counts))
    zi, zo = z[:n], z[n:]
    obj = np.repeat(np.arange(len(counts)), counts)
    return math.sqrt(u_noise) * (math.sqrt(cov) * zo[obj] + math.sqrt(d) * zi)


def sample_u_prior_data(data, stream, u_noise):
    """U ~ N(0, uNoise*SigmaU): block form, or sqrt(uNoise) * L_S z for a dense SigmaU (z = the stream's first n normals)."""
    if data.sigma_u_dense is None:
        return sample_u_prior(stream, data.spec.n, data.counts, u_noise, data.eps, data.cov)
    return math.sqrt(u_noise) * (data.sigma_u_chol() @ stream.normal_vector(data.spec.n))


def generate_initial_state(data, seed, chain, observe_x=False):
    """Gen `generate(model, args, obs)` (src/inference.jl:20,73,123,156,189,261,321,370): unobserved addresses are
    drawn from their priors in model order."""
    spec = data.spec
    pr = data.prior
    theta = np.full(spec.n_params, np.nan)
    for (name, i, j) in spec.active_params():
        p = spec.idx(name, i, j)
        a, b = spec.prior_key(name)
        s = px.Stream(seed, chain, p, px.stream_b(px.TAG_INIT_PARAM, 0))
        theta[p] = s.inv_gamma(pr[a], pr[b])
    U = np.zeros((spec.nU, spec.n))
    for k in range(spec.nU):
        s = px.Stream(seed, chain, k, px.stream_b(px.TAG_INIT_VEC, 0))
        U[k] = sample_u_prior_data(data, s, theta[spec.idx("uNoise")])
    st = State(theta, U)
    if (not spec.has_u) and spec.has_x and not observe_x:
        st.Xmodel = np.zeros((spec.n, spec.nX))
        for k in range(spec.nX):
            s = px.Stream(seed, chain, k, px.stream_b(px.TAG_INIT_XMODEL, 0))
            st.Xmodel[:, k] = s.normal_vector(spec.n)
    if spec.binary:
        s = px.Stream(seed, chain, spec.nU, px.stream_b(px.TAG_INIT_VEC, 0))
        z = s.normal_vector(spec.n)
        if factor_exists(spec, spec.nX):
            K = factor_cov(data, st, spec.nX)
            L = sla.cholesky(K, lower=True)
            st.logitT = L @ z
        else:
            st.logitT = z
    return st


# ------------------------------------------------------------------------------------------------ MH / ESS

# Safety cap shared with the CUDA path (csrc/sampler.cu): after this many evaluations of one slice the current proposal is
# taken (the bracket has collapsed onto the current state long before; Gen itself has no cap).
ESS_MAX_EVALS = 200


def mh_site(data, st, sc, site_index, name, i, j, seed, chain, it):
    """One `mh(trace, paramProposal, (drift, addr))`. Returns accepted flag."""
    spec = data.spec
    pr = data.prior
    p = spec.idx(name, i, j)
    cur = st.theta[p]
    shape_f, scale_f = proposal_params(cur, pr["drift"])
    new = px.Stream(seed, chain, site_index, px.stream_b(px.TAG_MH_PROP, it)).inv_gamma(shape_f, scale_f)
    fwd = ig_logpdf(new, shape_f, scale_f)
    shape_b, scale_b = proposal_params(new, pr["drift"])
    bwd = ig_logpdf(cur, shape_b, scale_b)
    a, b = spec.prior_key(name)
    dprior = ig_logpdf(new, pr[a], pr[b]) - ig_logpdf(cur, pr[a], pr[b])
    st_new = st.copy()
    st_new.theta[p] = new
    factors = [f for f in param_factors(spec, name, i, j) if factor_exists(spec, f)]
    u_terms = list(range(spec.nU)) if name == "uNoise" else []
    try:
        dlik, cache = sc.rescore(st_new, factors, u_terms)
    except np.linalg.LinAlgError:
        dlik, cache = -math.inf, None  # documented relaxation: a non-PD proposal is rejected
    alpha = dprior + dlik - fwd + bwd
    u = px.Stream(seed, chain, site_index, px.stream_b(px.TAG_MH_ACC, it)).uniform()
    if math.log(u) < alpha:
        st.theta[p] = new
        sc.commit(cache)
        return True
    return False


def ess_u(data, st, sc, k, seed, chain, it, stats=None, ess_rule="gen_joint_weight"):
    """`elliptical_slice(trace, :U=>k=>:U, zeros(n), uCov)` (src/inference.jl:50-54)."""
    spec = data.spec
    nu = sample_u_prior_data(data, px.Stream(seed, chain, k, px.stream_b(px.TAG_ESS_NU, it)), st.theta[spec.idx("uNoise")])
    sca = px.Stream(seed, chain, k, px.stream_b(px.TAG_ESS_SCALAR, it))
    u, v = sca.uniform_pair()
    logu = math.log(u)
    theta = 2.0 * math.pi * v
    tmin, tmax = theta - 2.0 * math.pi, theta
    f = st.U[k].copy()
    factors = [g for g in range(spec.nX + 2) if factor_exists(spec, g)]
    evals = 0
    while True:
        st_new = st.copy()
        st_new.U[k] = f * math.cos(theta) + nu * math.sin(theta)
        try:
            w, cache = sc.rescore(st_new, factors, [k])
            if ess_rule == "likelihood_only":
                w -= (cache[1][k] - sc.u_lp[k])
        except np.linalg.LinAlgError:
            w, cache = -math.inf, sc.rescore(st, [], [])[1]
            st_new = st.copy()
        evals += 1
        if not (w <= logu) or evals >= ESS_MAX_EVALS:   # Gen: `while weight <= log(u)` (a NaN weight leaves the loop)
            break
        if theta < 0:
            tmin = theta
        else:
            tmax = theta
        theta = tmin + (tmax - tmin) * sca.uniform()
    st.U[k] = st_new.U[k]
    sc.commit(cache)
    if stats is not None:
        stats["ess_evals"] = stats.get("ess_evals", 0) + evals
    return evals


def logit_t_cov(data, st):
    """src/inference.jl:216-227 — built from per-dimension U vectors (no toMatrix interleave) and the model X."""
    spec = data.spec
    th = st.theta
    logk = np.zeros((spec.n, spec.n))
    for i in range(spec.nU):
        logk = logk + rbf_kernel_log(st.U[i], st.U[i], th[spec.idx("utLS", i)])
    if spec.has_x:
        # both binary methods that reach here with X observe X (full model), or use data X (no-U, :338)
        logk = logk + rbf_kernel_log(data.X, data.X, np.array([th[spec.idx("xtLS", k)] for k in range(spec.nX)]))
    return process_cov(logk, th[spec.idx("tScale")], th[spec.idx("tNoise")])


def ess_logit_t(data, st, sc, L_stale, seed, chain, it, stats=None, ess_rule="gen_joint_weight"):
    """`elliptical_slice(trace, :logitT, zeros(n), logitTCov)` (src/inference.jl:233, 293, 347); the prior draw ν uses
    the covariance computed once per outer iteration (App. B6) while the weight uses the model's current one."""
    spec = data.spec
    a = spec.nU
    z = px.Stream(seed, chain, a, px.stream_b(px.TAG_ESS_NU, it)).normal_vector(spec.n)
    nu = L_stale @ z
    sca = px.Stream(seed, chain, a, px.stream_b(px.TAG_ESS_SCALAR, it))
    u, v = sca.uniform_pair()
    logu = math.log(u)
    theta = 2.0 * math.pi * v
    tmin, tmax = theta - 2.0 * math.pi, theta
    f = st.logitT.copy()
    fT = spec.nX
    bern_old = bernoulli_logpmf(data.T, st.logitT)
    evals = 0
    while True:
        st_new = st.copy()
        st_new.logitT = f * math.cos(theta) + nu * math.sin(theta)
        w, cache = sc.rescore(st_new, [fT], [])
        if ess_rule == "likelihood_only":
            w = 0.0     # textbook rule: only the Bernoulli likelihood of the sliced address enters the test
        w += bernoulli_logpmf(data.T, st_new.logitT) - bern_old
        evals += 1
        if not (w <= logu) or evals >= ESS_MAX_EVALS:
            break
        if theta < 0:
            tmin = theta
        else:
            tmax = theta
        theta = tmin + (tmax - tmin) * sca.uniform()
    st.logitT = st_new.logitT
    sc.commit(cache)
    if stats is not None:
        stats["ess_evals_logit"] = stats.get("ess_evals_logit", 0) + evals
    return evals


# ------------------------------------------------------------------------------------------------ schedule

def pack_sample(spec, st):
    """Packed posterior-sample record."""
    parts = [st.theta, st.U.reshape(-1)]
    if spec.binary:
        parts.append(st.logitT)
    if st.Xmodel is not None:
        parts.append(st.Xmodel.reshape(-1, order="F"))
    return np.concatenate(parts)


def sample_stride(spec, observe_x=False):
    s = spec.n_params + spec.nU * spec.n + (spec.n if spec.binary else 0)
    if (not spec.has_u) and spec.has_x and not observe_x:
        s += spec.n * spec.nX
    return s


def posterior(data, nOuter, nMHInner, nESInner, seed=0, chain=0, mode="incremental", observe_x=False,
              ess_rule="gen_joint_weight", stats=None, init_state=None):
    """`Posterior(priorparams, X, T, Y, nU, nOuter, nMHInner, nESInner)` — all eight methods
    (src/inference.jl:4-59, 62-102, 112-143, 146-165, 169-242, 245-302, 305-353, 356-379).
    Returns (samples[nOuter, stride], final State)."""
    spec = data.spec
    st = init_state.copy() if init_state is not None else generate_initial_state(data, seed, chain, observe_x)
    sc = Scorer(data, st, mode)
    sites = spec.mh_sites()
    if not spec.has_u and not spec.has_x:
        nMHInner = 1  # inference.jl:157-160, 371-374: the three sites are visited once per outer iteration
    out = np.zeros((nOuter, sample_stride(spec, observe_x)))
    acc = np.zeros(len(sites), dtype=np.int64)
    for i in range(nOuter):
        for j in range(nMHInner):
            for s, (name, a, b) in enumerate(sites):
                acc[s] += mh_site(data, st, sc, s, name, a, b, seed, chain, i * nMHInner + j)
        if spec.has_u or spec.binary:
            do_logit = spec.binary and (spec.has_u or spec.has_x)
            if do_logit:
                L_stale = sla.cholesky(logit_t_cov(data, st), lower=True)
            for j in range(nESInner if (spec.has_u or spec.has_x) else 0):
                it = i * nESInner + j
                if do_logit:
                    ess_logit_t(data, st, sc, L_stale, seed, chain, it, stats, ess_rule)
                for k in range(spec.nU):
                    ess_u(data, st, sc, k, seed, chain, it, stats, ess_rule)
        out[i] = pack_sample(spec, st)
    if stats is not None:
        stats["accepts"] = acc
        stats["n_chol"] = sc.n_chol
        stats["n_build"] = sc.n_build
    return out, st
