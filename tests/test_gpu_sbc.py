"""Simulation-based calibration (Talts et al. 2018) of the CUDA sampler — the correct version of what the reference's
disabled test/sbc.jl sets out to do (its pass/fail logic is vacuous: ranks are jittered by 1e20 and the decision is
inverted, test/sbc.jl:50,62-68; only the shapes are taken from it). Each chain gets ITS OWN dataset simulated from the
model prior with a known ground truth theta*; the rank of theta* among thinned posterior draws must be uniform.

The slice-sampling rule matters: with the textbook likelihood-only test (ess_rule=1) the sampler targets the posterior and
the ranks are uniform; the default rule reproduces Gen's `elliptical_slice` as recollected in SURVEY.md App. C (the full
`update` weight, which counts the Gaussian prior of the sliced address again) and is reported, not asserted."""
import numpy as np
import pytest
import scipy.stats as sst

import gpslc_b200 as g
from gpslc_b200.inference import ChainSampler
from oracle import kernel as ok, model as om

pytestmark = pytest.mark.gpu

N, COUNTS, NX, NU = 12, [4, 4, 4], 2, 1
TRIALS, BURN, THIN, DRAWS = 640, 60, 20, 15   # thinning 20 outer iterations: at 6 the xScale ranks are visibly autocorrelated


def simulate(rng, prior, binary=False):
    spec = om.ModelSpec(N, NU, NX, binary)
    theta = np.full(spec.n_params, np.nan)
    for (name, i, j) in spec.active_params():
        theta[spec.idx(name, i, j)] = sst.invgamma.rvs(prior[name + "Shape"], scale=prior[name + "Scale"], random_state=rng)
    obj = np.repeat(np.arange(len(COUNTS)), COUNTS)
    d = (1.0 + prior["sigmaUNoise"]) - prior["sigmaUCov"]
    U = np.sqrt(theta[0]) * (np.sqrt(prior["sigmaUCov"]) * rng.standard_normal(len(COUNTS))[obj] + np.sqrt(d) * rng.standard_normal(N))
    Uc = U[:, None]
    def draw(K):
        return np.linalg.cholesky(K) @ rng.standard_normal(N)
    X = np.zeros((N, NX))
    for k in range(NX):
        X[:, k] = draw(ok.process_cov(ok.rbf_kernel_log(Uc, Uc, theta[spec.idx("uxLS", 0, k)]), theta[spec.idx("xScale", k)], theta[spec.idx("xNoise", k)]))
    xt = np.array([theta[spec.idx("xtLS", k)] for k in range(NX)]); xy = np.array([theta[spec.idx("xyLS", k)] for k in range(NX)])
    T = draw(ok.process_cov(ok.rbf_kernel_log(Uc, Uc, theta[spec.idx("utLS", 0)]) + ok.rbf_kernel_log(X, X, xt), theta[spec.idx("tScale")], theta[spec.idx("tNoise")]))
    logitT = None
    if binary:      # src/model.jl:73-89: the GP draw is logitT, T_i ~ Bernoulli(expit(logitT_i)), and the Y kernel sees the Boolean T
        logitT = T
        T = (rng.random(N) < ok.expit(logitT)).astype(np.float64)
    Y = draw(ok.process_cov(ok.rbf_kernel_log(Uc, Uc, theta[spec.idx("uyLS", 0)]) + ok.rbf_kernel_log(X, X, xy) + ok.rbf_kernel_log(T, T, theta[spec.idx("tyLS")]),
                            theta[spec.idx("yScale")], theta[spec.idx("yNoise")]))
    return theta, U, X, (T > 0.5) if binary else T, Y, logitT


def run_sbc(ctx, ess_rule, seed, binary=False, nES=2):
    prior = g.getPriorParameters()
    rng = np.random.default_rng(seed)
    sims = [simulate(rng, prior, binary) for _ in range(TRIALS)]
    truth = np.stack([s[0] for s in sims])
    X = np.stack([s[2] for s in sims]); T = np.stack([s[3] for s in sims]); Y = np.stack([s[4] for s in sims])
    nOuter = BURN + THIN * DRAWS
    smp = ChainSampler(prior, X, T, Y, NU, COUNTS, nOuter, 2, nES, n_chains=TRIALS, seed=seed, ess_rule=ess_rule, ctx=ctx,
                       per_chain_data=True)
    smp.run(nOuter)
    out = smp.samples()
    smp.close()
    draws = out[BURN + THIN - 1::THIN][:DRAWS]           # [DRAWS, TRIALS, stride]
    spec = om.ModelSpec(N, NU, NX, binary)
    pvals = {}
    for (name, i, j) in spec.active_params():
        p = spec.idx(name, i, j)
        ranks = (draws[:, :, p] < truth[None, :, p]).sum(axis=0)
        counts = np.bincount(ranks, minlength=DRAWS + 1)
        pvals[(name, i, j)] = sst.chisquare(counts).pvalue
    # a function of U: mean of the object-level confounder of the first object
    uhat = draws[:, :, spec.n_params:spec.n_params + COUNTS[0]].mean(axis=2)
    utrue = np.stack([s[1][:COUNTS[0]].mean() for s in sims])
    ranks = (uhat < utrue[None, :]).sum(axis=0)
    pvals[("U_obj1", 0, 0)] = sst.chisquare(np.bincount(ranks, minlength=DRAWS + 1)).pvalue
    if binary:      # a function of the sliced logitT vector
        o = spec.n_params + NU * N
        lhat = draws[:, :, o:o + N].mean(axis=2)
        ltrue = np.stack([s[5].mean() for s in sims])
        ranks = (lhat < ltrue[None, :]).sum(axis=0)
        pvals[("logitT_mean", 0, 0)] = sst.chisquare(np.bincount(ranks, minlength=DRAWS + 1)).pvalue
    return pvals


def test_sbc_rank_uniformity_textbook_slice_rule(ctx):
    pvals = run_sbc(ctx, ess_rule=1, seed=2024)
    for k, v in sorted(pvals.items(), key=lambda kv: kv[1]):
        print(f"SBC ess_rule=1 {k}: p = {v:.4f}")
    alpha = 0.01 / len(pvals)                      # Bonferroni
    bad = {k: v for k, v in pvals.items() if v < alpha}
    assert not bad, bad


def test_sbc_binary_treatment_textbook_slice_rule(ctx):
    """Binary T: logitT ~ N(0, K_T), T_i ~ Bernoulli(expit(logitT_i)); ess_rule=1 must apply to the logitT slice as well (Bernoulli
    terms only). One slice pass per outer iteration, so that the covariance the direction nu is drawn with (computed once per
    outer iteration) is the current one and the update is an exact elliptical slice step."""
    pvals = run_sbc(ctx, ess_rule=1, seed=2026, binary=True, nES=1)
    for k, v in sorted(pvals.items(), key=lambda kv: kv[1]):
        print(f"SBC binary ess_rule=1 {k}: p = {v:.4f}")
    alpha = 0.01 / len(pvals)
    bad = {k: v for k, v in pvals.items() if v < alpha}
    assert not bad, bad


def test_sbc_default_gen_rule_is_not_calibrated_for_u(ctx):
    """What the product default (ess_rule=0: Gen's `elliptical_slice` as recollected — the test uses the full `update` weight,
    which counts the Gaussian prior of the sliced address a second time) does, asserted as what it is:
    the chain does NOT target the model's posterior in the U direction. uNoise and the object-level confounder fail rank
    uniformity decisively (p < 1e-6; measured 0 and 1e-65 at 4096 trials, profiles/sbc_r01.md), while every hyperparameter that does
    not touch U stays calibrated. It remains the default because the target is the reference's behaviour; whoever can run Gen
    should confirm inference/elliptical_slice.jl and, if the recollection is wrong, flip the default to 1."""
    pvals = run_sbc(ctx, ess_rule=0, seed=2025)
    for k, v in sorted(pvals.items(), key=lambda kv: kv[1]):
        print(f"SBC ess_rule=0 {k}: p = {v:.4g}")
    assert pvals[("uNoise", 0, 0)] < 1e-6 and pvals[("U_obj1", 0, 0)] < 1e-6
    alpha = 0.01 / len(pvals)
    independent_of_u = [k for k in pvals if k[0] in ("tNoise", "yNoise", "tyLS", "tScale", "yScale", "xNoise", "xScale", "xtLS", "xyLS")]
    assert len(independent_of_u) == 13
    bad = {k: pvals[k] for k in independent_of_u if pvals[k] < alpha}
    assert not bad, bad
