"""Parity tests proper (run on the B200 box with -m gpu): every call goes through the C ABI of libgpslc_b200.so and is
compared with the oracle on the same seeded inputs, with the reference's golden vectors, or through size-independent
properties at BASELINE.json's full sizes. Tolerances: covariance elements abs <= 4 ulp * scale;
log-densities rel <= 1e-10; MeanITE / CovITE rel <= 1e-8."""
import os

import numpy as np
import pytest

import gpslc_b200 as g
from gpslc_b200 import estimation as ge
from gpslc_b200.inference import ChainSampler
from oracle import kernel as ok, model as om, inference as oi, estimation as oe, data as od

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


# ------------------------------------------------------------------------------------------------ covariance build
def test_cov_build_reference_kats(ctx, kats):
    k = kats["rbfKernelLog_magic"]
    X = np.array(k["X"], dtype=float)
    assert np.array_equal(g.rbfKernelLog(X, X, k["LS"], ctx=ctx), np.array(k["expected"], dtype=float))
    ones = np.ones((10, 5))
    assert np.array_equal(g.rbfKernelLog(ones, ones, 0.1, ctx=ctx), np.zeros((10, 10)))
    assert np.array_equal(g.cov_build(np.zeros((1, 1)), np.zeros((1, 1)), [[1.0]], 2.0, None, ctx=ctx)[0], [[2.0]])
    assert np.array_equal(g.cov_build(np.zeros((1, 1)), np.zeros((1, 1)), [[1.0]], 0.0, 1e-5, ctx=ctx)[0], [[1e-5]])
    assert np.array_equal(g.cov_build(np.zeros((1, 1)), np.zeros((1, 1)), [[1.0]], 2.0, 1e-5, ctx=ctx)[0], [[2.0 + 1e-5]])


@pytest.mark.parametrize("n,D,batch", [(1, 1, 1), (7, 2, 3), (63, 3, 2), (64, 1, 2), (65, 12, 2), (129, 5, 3), (150, 1, 4), (272, 7, 2)])
def test_cov_build_vs_oracle(ctx, n, D, batch):
    rng = np.random.default_rng(n * 31 + D)
    F = rng.standard_normal((n, D)); ls = 0.3 + 2 * rng.random((batch, D)); sc = 0.2 + rng.random(batch); nz = 0.1 + rng.random(batch)
    K = g.cov_build(F, F, ls, sc, nz, ctx=ctx)
    for b in range(batch):
        want = ok.process_cov(ok.rbf_kernel_log(F, F, ls[b]), sc[b], nz[b])
        assert np.max(np.abs(K[b] - want)) <= 4 * np.finfo(float).eps * (sc[b] + nz[b])
    # per-batch features and X1 != X2 (src/likelihood.jl:27: K(T, doT))
    F1 = rng.standard_normal((batch, n, D)); F2 = rng.standard_normal((batch, n, D))
    K = g.cov_build(F1, F2, ls, sc, None, ctx=ctx)
    for b in range(batch):
        want = ok.process_cov(ok.rbf_kernel_log(F1[b], F2[b], ls[b]), sc[b])
        assert np.max(np.abs(K[b] - want)) <= 4 * np.finfo(float).eps * sc[b]


def test_cov_build_bool_features(ctx):
    t = np.array([True, False, True, True, False])
    K = g.cov_build(t.astype(float)[:, None], t.astype(float)[:, None], [[0.7]], 1.0, None, ctx=ctx)[0]
    assert np.allclose(K, ok.process_cov(ok.rbf_kernel_log(t, t, 0.7), 1.0), rtol=0, atol=4e-16)


# ------------------------------------------------------------------------------------------------ Cholesky log-density
@pytest.mark.parametrize("n", [1, 2, 8, 63, 64, 65, 127, 128, 150, 200, 256, 272, 513])
def test_chol_logpdf_vs_oracle(ctx, n):
    rng = np.random.default_rng(n)
    batch = 3
    F = rng.standard_normal((n, 4)); ls = 0.5 + 2 * rng.random((batch, 4)); sc = 0.5 + rng.random(batch); nz = 0.05 + rng.random(batch)
    y = rng.standard_normal((batch, n))
    Ks = np.stack([ok.process_cov(ok.rbf_kernel_log(F, F, ls[b]), sc[b], nz[b]) for b in range(batch)])
    want = np.array([om.mvn_logpdf_chol(y[b], Ks[b]) for b in range(batch)])
    lp, ld, q, info = g.chol_logpdf(Ks, y, ctx=ctx)
    assert np.all(info == 0)
    assert np.max(np.abs(lp - want) / np.abs(want)) <= 1e-10
    sign, ldo = np.linalg.slogdet(Ks)
    assert np.max(np.abs(ld - ldo) / (1 + np.abs(ldo))) <= 1e-11
    lp2, ld2, q2, info2 = g.rbf_logpdf(F, ls, sc, nz, y, ctx=ctx)      # fused build + factor
    assert np.all(info2 == 0) and np.max(np.abs(lp2 - want) / np.abs(want)) <= 1e-10
    lp3, *_ = g.chol_logpdf(Ks, y[0], ctx=ctx)                          # shared y
    assert abs(lp3[0] - want[0]) <= 1e-10 * abs(want[0])


def test_many_feature_dimensions(ctx):
    """More than 24 feature dimensions: the panel's column features no longer fit the shared-memory staging area and are read from
    global memory (same values); up to 64 dimensions (nU + nX + 1) are supported, more are refused loudly."""
    rng = np.random.default_rng(11)
    n, batch = 150, 2
    for D in (25, 40, 64):
        F = rng.standard_normal((n, D)); ls = 2.0 + 3 * rng.random((batch, D)); sc = 0.5 + rng.random(batch); nz = 0.05 + rng.random(batch)
        y = rng.standard_normal((batch, n))
        want = np.array([om.mvn_logpdf_chol(y[b], ok.process_cov(ok.rbf_kernel_log(F, F, ls[b]), sc[b], nz[b])) for b in range(batch)])
        lp, ld, q, info = g.rbf_logpdf(F, ls, sc, nz, y, ctx=ctx)
        assert np.all(info == 0) and np.max(np.abs(lp - want) / np.abs(want)) <= 1e-10
        K = g.cov_build(F, F, ls, sc, nz, ctx=ctx)
        assert np.max(np.abs(K[0] - ok.process_cov(ok.rbf_kernel_log(F, F, ls[0]), sc[0], nz[0]))) <= 4 * np.finfo(float).eps * (sc[0] + nz[0])
    counts, X, T, Y = od.synthetic(60, 3, 30, seed=4)          # Y factor: 32 dimensions
    md = od.model_data_from_arrays(counts, X, T, Y, nU=1)
    *_, got, acc, ev = _run_pair(md, X, T, Y, counts, 2, 1, 1, 3, 1)
    want, _ = oi.posterior(md, 2, 1, 1, seed=3, chain=0)
    assert np.allclose(np.nan_to_num(want), np.nan_to_num(got[:, 0, :]), rtol=1e-8, atol=1e-11)
    with pytest.raises(g.GpslcError) as ei:
        g.rbf_logpdf(rng.standard_normal((n, 65)), np.ones((1, 65)), [1.0], [0.1], rng.standard_normal((1, n)), ctx=ctx)
    assert ei.value.code == 2


def test_chol_not_positive_definite_reports_lapack_info(ctx):
    K = np.eye(70); K[40, 40] = -1.0
    lp, ld, q, info = g.chol_logpdf(K[None], np.ones(70), ctx=ctx)
    assert info[0] == 41 and lp[0] == -np.inf                            # leading minor 41 (PosDefException(41) in Julia)
    K = np.ones((5, 5))                                                  # singular
    assert g.chol_logpdf(K[None], np.ones(5), ctx=ctx)[3][0] == 2
    Ks = np.stack([np.eye(9), -np.eye(9), 2 * np.eye(9)])                # failure of one batch element does not leak
    lp, ld, q, info = g.chol_logpdf(Ks, np.ones(9), ctx=ctx)
    assert list(info) == [0, 1, 0] and np.isclose(ld[2], 9 * np.log(2.0)) and np.isclose(q[0], 9.0)


def test_full_size_properties_n1024(ctx):
    """BASELINE c3 size: closed-form checks that need no oracle factorisation."""
    n, batch = 1024, 5
    rng = np.random.default_rng(5)
    y = rng.standard_normal(n)
    # (i) huge lengthscale => K = s*11' + z*I : det and quadratic form in closed form
    F = rng.standard_normal((n, 3)); sc = np.array([0.7, 1.3, 2.0, 0.4, 1.0]); nz = np.array([0.3, 0.5, 1.0, 2.0, 0.1])
    lp, ld, q, info = g.rbf_logpdf(F, np.full((batch, 3), 1e9), sc, nz, y, ctx=ctx)
    for b in range(batch):
        ld_want = (n - 1) * np.log(nz[b]) + np.log(nz[b] + n * sc[b])
        q_want = (y @ y) / nz[b] - sc[b] * y.sum() ** 2 / (nz[b] * (nz[b] + n * sc[b]))
        assert abs(ld[b] - ld_want) <= 1e-9 * abs(ld_want) and abs(q[b] - q_want) <= 1e-9 * abs(q_want)
    # (ii) fused path == dense path on the same matrices, typical hyperparameters
    ls = 0.8 + 2 * rng.random((batch, 3))
    K = g.cov_build(F, F, ls, sc, nz, ctx=ctx)
    a = g.chol_logpdf(K, y, ctx=ctx); b_ = g.rbf_logpdf(F, ls, sc, nz, y, ctx=ctx)
    assert np.all(a[3] == 0) and np.max(np.abs(a[0] - b_[0]) / np.abs(a[0])) <= 1e-12
    # (iii) linearity of the solve: quad(2y) = 4 quad(y)
    c = g.rbf_logpdf(F, ls, sc, nz, 2 * y, ctx=ctx)
    assert np.max(np.abs(c[2] - 4 * b_[2]) / b_[2]) <= 1e-12 and np.allclose(c[1], b_[1], rtol=1e-14)


def test_full_size_properties_n4096(ctx):
    """BASELINE c4 size (64 panels per matrix): closed forms and linearity through the fused build + Cholesky."""
    n, batch = 4096, 2
    rng = np.random.default_rng(6)
    y = rng.standard_normal(n)
    F = rng.standard_normal((n, 11)); sc = np.array([0.7, 1.3]); nz = np.array([0.3, 0.5])
    lp, ld, q, info = g.rbf_logpdf(F, np.full((batch, 11), 1e9), sc, nz, y, ctx=ctx)
    assert np.all(info == 0)
    for b in range(batch):
        ld_want = (n - 1) * np.log(nz[b]) + np.log(nz[b] + n * sc[b])
        q_want = (y @ y) / nz[b] - sc[b] * y.sum() ** 2 / (nz[b] * (nz[b] + n * sc[b]))
        assert abs(ld[b] - ld_want) <= 1e-9 * abs(ld_want) and abs(q[b] - q_want) <= 1e-8 * abs(q_want)
    ls = 2.0 + 2 * rng.random((batch, 11))
    b1 = g.rbf_logpdf(F, ls, sc, nz, y, ctx=ctx); b2 = g.rbf_logpdf(F, ls, sc, nz, -3 * y, ctx=ctx)
    assert np.all(b1[3] == 0) and np.max(np.abs(b2[2] - 9 * b1[2]) / b1[2]) <= 1e-12 and np.allclose(b2[1], b1[1], rtol=1e-14)
    # permuting the individuals permutes K symmetrically: same determinant and quadratic form (different panel contents)
    perm = rng.permutation(n)
    b3 = g.rbf_logpdf(F[perm], ls, sc, nz, y[perm], ctx=ctx)
    assert np.max(np.abs(b3[1] - b1[1]) / np.abs(b1[1])) <= 1e-10 and np.max(np.abs(b3[2] - b1[2]) / b1[2]) <= 1e-9


def test_large_ite_in_team_mode_is_consistent_with_sate(ctx):
    """n = 2048 (4096 x 4096 augmented matrices, few tasks => cluster teams picked automatically): the SATE fast path (one
    n x n Cholesky with two right-hand sides, never forms CovITE) must agree with the mean / grand sum of what the ITE path
    produces, and the zero-effect structure must hold at doT where T == doT for nobody (finite, PD, symmetric CovITE)."""
    n, n_obj, nX = 2048, 32, 4
    counts, X, T, Y = od.synthetic(n, n_obj, nX, seed=21)
    spec = om.ModelSpec(n, 1, nX, False)
    rng = np.random.default_rng(3)
    rec = np.ones(spec.n_params + n)
    rec[:spec.n_params] = 0.8 + 0.4 * rng.random(spec.n_params)
    rec[2] = 0.3
    rec[spec.n_params:] = np.repeat(rng.standard_normal(n_obj), n // n_obj)
    doT = np.array([-0.5, 0.1, 0.9])
    ret = np.array([0], dtype=np.int32)
    o = ge.ite(rec[None, None, :], X, T, Y, 1, doT, ret, 1e-10, 2, want_cov=True, ctx=ctx)
    so = ge.sate(rec[None, None, :], X, T, Y, 1, doT, ret, 1e-10, 2, ctx=ctx)
    assert o["info"].max() == 0 and so["info"].max() == 0 and np.all(np.isfinite(o["samples"]))
    for d in range(3):
        M, Cv = o["mean"][d, 0, 0], o["cov"][d, 0, 0]
        assert np.array_equal(Cv, Cv.T)
        assert abs(M.mean() - so["mean"][d, 0, 0]) <= 1e-9 * max(1.0, abs(M).max())
        assert abs(Cv.sum() / n ** 2 - so["var"][d, 0, 0]) <= 1e-7 * abs(so["var"][d, 0, 0]) + 1e-12


# ------------------------------------------------------------------------------------------------ sampler vs oracle chain
def _run_pair(md, X, T, Y, counts, nOuter, nMH, nES, seed, C, **opts):
    s = ChainSampler(md.prior, X, T, Y, md.spec.nU, counts, nOuter, nMH, nES, n_chains=C, seed=seed, **opts)
    st0 = s.state(); lp0, q0 = s.terms()
    s.run(nOuter)
    got = s.samples(); acc, ev = s.stats()
    _run_pair.logit_evals = s.ess_evals_logit
    s.close()
    return st0, lp0, q0, got, acc, ev


@pytest.mark.parametrize("binary", [False, True])
@pytest.mark.parametrize("n,n_obj,nX,nU,with_u", [(48, 4, 3, 1, True), (100, 5, 2, 2, True), (150, 6, 0, 1, True),
                                                  (64, 4, 3, 1, False), (72, 4, 0, 1, False), (130, 2, 6, 1, True)])
def test_sampler_reproduces_oracle_chain(ctx, n, n_obj, nX, nU, with_u, binary):
    """Same Philox streams => the CUDA chains and the oracle chain coincide for all eight `Posterior` methods (full, no-X,
    no-U, neither x real/binary T; nU=2 exercises the reference toMatrix interleave, App. B1). Accept decisions and
    slice-evaluation counts (U_k and logitT) must be identical."""
    counts, X, T, Y = od.synthetic(n, n_obj, max(nX, 1), seed=5)
    if nX == 0:
        X = None
    if binary:
        T = T > np.median(T)
    md = od.model_data_from_arrays(counts if with_u else None, X, T, Y, nU=nU)
    C, nOuter, nMH, nES, seed = 3, 3, 2, 2, 17
    st0, lp0, q0, got, acc, ev = _run_pair(md, X, T, Y, counts if with_u else None, nOuter, nMH, nES, seed, C)
    for c in range(C):
        st = oi.generate_initial_state(md, seed, c)
        packed = oi.pack_sample(md.spec, st)
        assert np.allclose(np.nan_to_num(packed), np.nan_to_num(st0[c]), rtol=1e-10, atol=1e-12)
        for f in range(md.spec.nX + 2):
            if om.factor_exists(md.spec, f):
                want = om.factor_logpdf(md, st, f)
                assert abs(lp0[c, f] - want) <= 1e-10 * abs(want)
        for k in range(md.spec.nU):
            qw, _ = om.u_prior_quad_logdet(st.U[k], counts, md.eps, md.cov)
            assert abs(q0[c, k] - qw) <= 1e-9 * abs(qw)        # closed form on both sides; the (u - mean)^2 / 1e-13 term cancels ~7 digits
                                                               # (a dense LAPACK evaluation of the same quantity is only 1e-6 accurate)
        stats = {}
        want, _ = oi.posterior(md, nOuter, nMH, nES, seed=seed, chain=c, stats=stats)
        assert np.array_equal(stats["accepts"], acc[c].astype(np.int64))
        assert stats.get("ess_evals", 0) == int(ev[c])
        assert stats.get("ess_evals_logit", 0) == int(_run_pair.logit_evals[c])
        assert np.allclose(np.nan_to_num(want), np.nan_to_num(got[:, c, :]), rtol=1e-8, atol=1e-11)


@pytest.mark.parametrize("binary", [False, True])
def test_sampler_cluster_teams_match_single_cta(ctx, monkeypatch, binary):
    """Few chains at large n run one thread-block cluster per (chain, lane) task (mh_lanes_kernel<1>, eval_factors_kernel<1>): every
    CTA of a team replays the same site loop, the factorisations are shared. The chains must be bit-identical to the
    one-CTA-per-task kernels for every team size (accept decisions included), here forced on small ragged problems."""
    counts, X, T, Y = od.synthetic(150, 6, 3, seed=4)
    if binary:
        T = T > np.median(T)
    md = od.model_data_from_arrays(counts, X, T, Y, nU=1)
    outs = {}
    for team in (1, 2, 4, 8):
        monkeypatch.setenv("GPSLC_TEAM", str(team))
        outs[team] = _run_pair(md, X, T, Y, counts, 3, 2, 2, 21, 3)
    monkeypatch.delenv("GPSLC_TEAM")
    for team in (2, 4, 8):
        for a, b in zip(outs[1], outs[team]):
            assert np.array_equal(a, b, equal_nan=True), team
    want, _ = oi.posterior(md, 3, 2, 2, seed=21, chain=1)
    got = outs[4][3][:, 1, :]
    assert np.nanmax(np.abs(want - got) / (1e-9 + np.abs(want))) < 1e-8


def test_sampler_option_switches(ctx):
    counts, X, T, Y = od.synthetic(60, 3, 2, seed=2)
    # column-wise U layout instead of the reference interleave; textbook ESS rule
    md = od.model_data_from_arrays(counts, X, T, Y, nU=2, u_layout_reference=False)
    *_, got, acc, ev = _run_pair(md, X, T, Y, counts, 2, 2, 2, 4, 2, u_layout_mode=1, ess_rule=1)
    for c in range(2):
        want, _ = oi.posterior(md, 2, 2, 2, seed=4, chain=c, ess_rule="likelihood_only")
        assert np.allclose(want, got[:, c, :], rtol=1e-9, atol=1e-12)
    # no-U model conditioned on the observed X (observe_x=1) vs the reference-faithful random X (App. B3)
    md = od.model_data_from_arrays(None, X, T, Y, nU=1)
    *_, got, acc, ev = _run_pair(md, X, T, Y, None, 2, 2, 2, 4, 2, observe_x=1)
    for c in range(2):
        want, _ = oi.posterior(md, 2, 2, 2, seed=4, chain=c, observe_x=True)
        assert want.shape[1] == got.shape[2] and np.allclose(np.nan_to_num(want), np.nan_to_num(got[:, c, :]), rtol=1e-9, atol=1e-12)


def test_sampler_dense_sigma_u_matches_oracle(ctx):
    """`samplePosterior(hyperparams, priorparams, SigmaU, X, T, Y)` accepts ANY SigmaU (src/driver.jl:59-69; generateU factors
    uNoise*SigmaU, src/model_prior.jl:27-30). A matrix that is not a generateSigmaU block matrix goes through
    gpslc_data.sigma_u_dense: factored once on the device, U prior density by forward substitution, draws L_S z. Chains must
    reproduce the oracle's Cholesky-based chain; real and binary T, ragged n (two 64-blocks), nU = 2 as well."""
    for n, nX, nU, binary in ((100, 2, 1, False), (70, 1, 2, True)):
        counts, X, T, Y = od.synthetic(n, 2, nX, seed=13)
        if binary:
            T = T > np.median(T)
        i = np.arange(n)
        S = 0.5 * np.exp(-np.abs(i[:, None] - i[None, :]) / 5.0) + 0.6 * np.eye(n)
        pri = {**g.getPriorParameters(), "SigmaU": S}
        md = od.model_data_from_arrays(None, X, T, Y, nU=nU, sigma_u_dense=S)
        s = ChainSampler(pri, X, T, Y, nU, None, 3, 2, 2, n_chains=3, seed=31, ctx=ctx)
        st0 = s.state(); lp0, q0 = s.terms()
        s.run(3)
        got = s.samples(); acc, ev = s.stats()
        s.close()
        L = np.linalg.cholesky(S)
        for c in range(3):
            st = oi.generate_initial_state(md, 31, c)
            assert np.allclose(np.nan_to_num(oi.pack_sample(md.spec, st)), np.nan_to_num(st0[c]), rtol=1e-10, atol=1e-12)
            for k in range(nU):
                z = np.linalg.solve(L, st.U[k])
                assert abs(q0[c, k] - z @ z) <= 1e-11 * (z @ z)
            stats = {}
            want, _ = oi.posterior(md, 3, 2, 2, seed=31, chain=c, stats=stats)
            assert np.array_equal(stats["accepts"], acc[c].astype(np.int64)) and stats["ess_evals"] == int(ev[c])
            assert np.allclose(np.nan_to_num(want), np.nan_to_num(got[:, c, :]), rtol=1e-8, atol=1e-11)
    # a SigmaU that is not positive definite is the PosDefException generateU would throw
    bad = np.eye(20); bad[5, 5] = -1.0
    counts, X, T, Y = od.synthetic(20, 2, 1, seed=1)
    with pytest.raises(g.GpslcError) as ei:
        ChainSampler({**g.getPriorParameters(), "SigmaU": bad}, X, T, Y, 1, None, 1, 1, 1, ctx=ctx)
    assert ei.value.code == 3


def test_binary_chain_with_textbook_slice_rule_and_small_prior_shapes(ctx):
    """ess_rule = 1 must apply to BOTH sliced addresses (U_k and logitT: the logitT test then uses the Bernoulli terms only), and
    InvGamma priors with shape < 1 go through the boosted gamma sampler; both against the oracle chain."""
    counts, X, T, Y = od.synthetic(90, 3, 2, seed=19)
    Tb = T > np.median(T)
    md = od.model_data_from_arrays(counts, X, Tb, Y, nU=1)
    *_, got, acc, ev = _run_pair(md, X, Tb, Y, counts, 3, 2, 2, 8, 2, ess_rule=1)
    logit_evals = _run_pair.logit_evals.copy()
    *_, got0, _, _ = _run_pair(md, X, Tb, Y, counts, 3, 2, 2, 8, 2, ess_rule=0)
    assert not np.array_equal(got, got0)
    for c in range(2):
        stats = {}
        want, _ = oi.posterior(md, 3, 2, 2, seed=8, chain=c, ess_rule="likelihood_only", stats=stats)
        assert stats["ess_evals_logit"] == int(logit_evals[c])
        assert np.allclose(np.nan_to_num(want), np.nan_to_num(got[:, c, :]), rtol=1e-8, atol=1e-11)
    pri = {**g.getPriorParameters(), "tyLSShape": 0.5, "yNoiseShape": 0.9, "uNoiseShape": 0.35}
    md = od.model_data_from_arrays(counts, X, T, Y, nU=1, prior=pri)
    *_, got, acc, ev = _run_pair(md, X, T, Y, counts, 2, 2, 1, 5, 2)
    for c in range(2):
        want, _ = oi.posterior(md, 2, 2, 1, seed=5, chain=c)
        assert np.allclose(np.nan_to_num(want), np.nan_to_num(got[:, c, :]), rtol=1e-8, atol=1e-11)
    for bad in ({"tyLSShape": 0.0}, {"yScaleScale": -1.0}, {"drift": 0.0}, {"xtLSShape": float("nan")}):
        with pytest.raises(g.GpslcError) as ei:
            ChainSampler({**g.getPriorParameters(), **bad}, X, T, Y, 1, counts, 1, 1, 1, ctx=ctx)
        assert ei.value.code == 2


def test_non_pd_prediction_raises_like_the_reference(ctx):
    """A posterior sample whose Kp = Kww + yNoise*I is not positive definite makes `\\` / `mvnormal` throw PosDefException in the
    reference (src/likelihood.jl:42-43, src/estimation.jl:105); the drivers must not hand back garbage draws."""
    counts, X, T, Y = od.synthetic(40, 4, 2, seed=6)
    h = g.getHyperParameters(); h.nOuter, h.nBurnIn = 4, 2
    gobj = g.gpslc(counts, X, T, Y, hyperparams=h, ctx=ctx)
    assert g.sampleITE(gobj, 0.3, ctx=ctx).shape == (40, 30)
    gobj.posteriorPacked[2, 0, 2] = -50.0           # yNoise of the second retained sample
    for fn in (lambda: g.sampleITE(gobj, 0.3, ctx=ctx), lambda: g.sampleSATE(gobj, 0.3, ctx=ctx), lambda: g.ITEDistributions(gobj, 0.3, ctx=ctx),
               lambda: g.SATEDistributions(gobj, 0.3, ctx=ctx), lambda: g.predictCounterfactualEffects(gobj, 2, fidelity=3, ctx=ctx)):
        with pytest.raises(g.PosDefException) as ei:
            fn()
        assert ei.value.info >= 1 and "retained sample 1" in str(ei.value)


def test_chain_sharding_is_invariant(ctx):
    """Chains 2,3 of a 4-chain run == a 2-chain run with chain_offset=2 (multi-GPU sharding)."""
    counts, X, T, Y = od.synthetic(80, 4, 2, seed=3)
    pri = g.getPriorParameters()
    a = g.Posterior({**pri, "_obj_counts": counts}, X, T, Y, 1, 2, 2, 2, n_chains=4, seed=9, ctx=ctx)
    b = g.Posterior({**pri, "_obj_counts": counts}, X, T, Y, 1, 2, 2, 2, n_chains=2, seed=9, chain_offset=2, ctx=ctx)
    assert np.array_equal(a[:, 2:], b)


def test_incremental_caches_stay_consistent_at_c3_shape(ctx):
    """n=1024, nX=10: after MH sweeps + an ESS pass, the cached factor log-densities equal a from-scratch re-evaluation
    of the final state (the invariant that lets a site re-score one factor instead of the reference's thirteen)."""
    counts, X, T, Y = od.synthetic(1024, 16, 10)
    s = ChainSampler(g.getPriorParameters(), X, T, Y, 1, counts, 1, 1, 1, n_chains=4, seed=1234, ctx=ctx)
    s.mh_sweeps(1)
    s.ess_pass(0)
    lp, q = s.terms()
    st = s.state()
    s.set_state(st)
    lp2, q2 = s.terms()
    acc, ev = s.stats()
    s.close()
    assert np.max(np.abs(lp - lp2) / np.abs(lp2)) <= 1e-12 and np.allclose(q, q2, rtol=1e-12)
    assert acc.sum() > 0 and np.all(ev >= 1)
    # spot-check one factor of one chain against the oracle at full size (Y factor: 12-dim kernel)
    md = od.model_data_from_arrays(counts, X, T, Y, nU=1)
    state = om.State(st[0, :md.spec.n_params].copy(), st[0, md.spec.n_params:].reshape(1, 1024).copy())
    want = om.factor_logpdf(md, state, 11)
    assert abs(lp[0, 11] - want) <= 1e-10 * abs(want)


@pytest.mark.timeout(600)
def test_chain_reproduces_oracle_at_c3_size(ctx):
    """BASELINE c3 shape (n=1024, 16 objects, nX=10, nU=1): one full default outer iteration — 10 MH sweeps of 58 sites with
    identical accept decisions + 5 elliptical-slice passes — of the CUDA chain against the oracle chain driven by the same Philox
    streams (the oracle needs ~25 s per chain on the host)."""
    n, n_obj, nX = 1024, 16, 10
    counts, X, T, Y = od.synthetic(n, n_obj, nX, seed=1234)
    md = od.model_data_from_arrays(counts, X, T, Y, nU=1)
    s = ChainSampler(md.prior, X, T, Y, 1, counts, 1, 10, 5, n_chains=2, seed=77, ctx=ctx)
    s.run(1); got = s.samples(); acc, ev = s.stats(); s.close()
    stats = {}
    want, _ = oi.posterior(md, 1, 10, 5, seed=77, chain=1, stats=stats)
    assert np.array_equal(stats["accepts"], acc[1].astype(np.int64)) and stats["ess_evals"] == int(ev[1])
    assert 150 < acc[1].sum() < 450
    err = np.nanmax(np.abs(want - got[:, 1, :]) / (1e-9 + np.abs(want)))
    print("c3-size chain: max relative deviation", err)
    assert err < 1e-9


@pytest.mark.timeout(600)
def test_factor_logpdf_at_c4_size_vs_scipy_cholesky(ctx):
    """BASELINE c4 size (n=4096, 64 objects, nX=10): the T and Y factor log-densities of a `generate`d state, computed by the
    fused build + 64-panel Cholesky (here in cluster-team mode: one chain), against SciPy's LAPACK Cholesky of the oracle-built
    covariance, rel 1e-10 (north_star's deterministic tolerance)."""
    n, n_obj, nX = 4096, 64, 10
    counts, X, T, Y = od.synthetic(n, n_obj, nX, seed=1234)
    md = od.model_data_from_arrays(counts, X, T, Y, nU=1)
    s = ChainSampler(md.prior, X, T, Y, 1, counts, 1, 1, 1, n_chains=1, seed=3, ctx=ctx)
    st = s.state(); lp, q = s.terms(); s.close()
    state = om.State(st[0, :md.spec.n_params].copy(), st[0, md.spec.n_params:].reshape(1, n).copy())
    for f in (nX, nX + 1, 3):
        want = om.factor_logpdf(md, state, f)
        print("c4-size factor", f, lp[0, f], want)
        assert abs(lp[0, f] - want) <= 1e-10 * abs(want)


@pytest.mark.timeout(1500)
def test_ite_at_c5_size_vs_reference_algebra(ctx):
    """BASELINE c5 size (n=8192): MeanITE and SATE of one (posterior sample, doT) from the 16384 x 16384 augmented Cholesky (cluster
    teams) against the oracle's line-by-line restatement of likelihood.jl / estimation.jl (two LU solves with n x n right-hand sides,
    Bunch-Kaufman, four n x n products: ~1e13 flops on the host). The posterior sample is a SAMPLED one: a chain's state after one
    outer iteration (1 MH sweep + 1 slice pass) at n=8192."""
    n, n_obj, nX = 8192, 128, 10
    counts, X, T, Y = od.synthetic(n, n_obj, nX, seed=1234)
    s = ChainSampler(g.getPriorParameters(), X, T, Y, 1, counts, 1, 1, 1, n_chains=1, seed=5, ctx=ctx)
    s.run(1); smp = s.samples(); acc, ev = s.stats(); s.close()
    assert acc.sum() > 5 and ev[0] >= 1
    ret = np.zeros(1, dtype=np.int32)
    doT = 0.3
    out = ge.ite(smp, X, T, Y, 1, [doT], ret, 1e-10, 2, seed=4, ctx=ctx)
    so = ge.sate(smp, X, T, Y, 1, [doT], ret, 1e-10, 2, seed=4, ctx=ctx)
    assert out["info"].max() == 0 and so["info"].max() == 0 and np.all(np.isfinite(out["samples"]))
    spec = om.ModelSpec(n, 1, nX, False)
    uyLS, xyLS, tyLS, yNoise, yScale, U = oe.extract_parameters(spec, smp[0, 0])
    M, Cv = oe.conditional_ite(uyLS, xyLS, tyLS, yNoise, yScale, U, X, T, Y, doT)
    assert np.max(np.abs(M - out["mean"][0, 0, 0])) <= 1e-8 * np.max(np.abs(M))
    ms, vs = oe.conditional_sate(M, Cv)
    vs += 1e-10 / n                                  # the jitter ITEDistributions adds to the diagonal (src/estimation.jl:82)
    print("c5-size: MeanSATE", ms, so["mean"][0, 0, 0], "VarSATE", vs, so["var"][0, 0, 0])
    assert abs(ms - so["mean"][0, 0, 0]) <= 1e-8 * max(abs(ms), np.max(np.abs(M)))
    assert abs(vs - so["var"][0, 0, 0]) <= 1e-7 * abs(vs)
    # the draws scatter around MeanITE with the posterior covariance: standardised by the oracle's diagonal they are O(1)
    zs = (out["samples"][0, 0] - M[None, :]) / np.sqrt(np.maximum(np.diag(Cv), 1e-300))[None, :]
    assert np.abs(zs).max() < 8.0 and 0.5 < zs.std() < 1.5


def test_error_behaviour(ctx):
    counts, X, T, Y = od.synthetic(20, 2, 1, seed=1)
    pri = g.getPriorParameters()
    with pytest.raises(g.GpslcError):                     # object counts must sum to n
        ChainSampler(pri, X, T, Y, 1, [5, 5], 1, 1, 1, ctx=ctx)
    with pytest.raises(g.GpslcError):                     # SigmaU needs cov < 1 + eps
        ChainSampler({**pri, "sigmaUCov": 2.0}, X, T, Y, 1, counts, 1, 1, 1, ctx=ctx)
    # duplicated individuals + vanishing noise prior => the initial covariance is singular: the reference would throw
    # PosDefException out of `generate`; the library returns GPSLC_ERR_NOT_PD (code 3)
    tiny = {**pri, "yNoiseScale": 1e-300, "tNoiseScale": 1e-300, "xNoiseScale": 1e-300}
    with pytest.raises(g.GpslcError) as ei:
        ChainSampler(tiny, np.ones((20, 1)), np.ones(20), np.ones(20), 1, [20], 1, 1, 1, ctx=ctx)
    assert ei.value.code == 3
    # non-finite data does not hang the slice sampler (Gen's `while weight <= log(u)` exits on a NaN weight)
    Yb = Y.copy(); Yb[3] = np.nan
    s = ChainSampler(pri, X, T, Yb, 1, counts, 1, 1, 1, ctx=ctx)
    s.run(1)
    assert s.stats()[1][0] >= 1
    s.close()


def test_empty_inputs_are_no_ops(ctx):
    """Zero retained samples, zero interventions, zero batch elements: success and empty outputs, no launch, no fault."""
    counts, X, T, Y = od.synthetic(24, 2, 2, seed=2)
    md = od.model_data_from_arrays(counts, X, T, Y, nU=1)
    smp = oi.posterior(md, 2, 1, 1, seed=1, chain=0, observe_x=True)[0][:, None, :]
    o = ge.ite(smp, X, T, Y, 1, np.zeros(0), np.array([0], dtype=np.int32), 1e-10, 3, ctx=ctx)
    assert o["mean"].shape == (0, 1, 1, 24) and o["samples"].shape == (0, 1, 3, 24)
    o = ge.ite(smp, X, T, Y, 1, [0.5], np.zeros(0, dtype=np.int32), 1e-10, 3, ctx=ctx)
    assert o["mean"].shape == (1, 1, 0, 24) and o["info"].size == 0
    so = ge.sate(smp, X, T, Y, 1, np.zeros(0), np.array([1], dtype=np.int32), 1e-10, 2, ctx=ctx)
    assert so["samples"].shape == (0, 1, 2)
    assert ge.summarize(np.zeros((0, 5, 7)), 0.9, ctx=ctx).shape == (0, 7, 3)
    # spp = 0: distributions only
    o = ge.ite(smp, X, T, Y, 1, [0.5], np.array([1], dtype=np.int32), 1e-10, 0, ctx=ctx)
    assert o["samples"] is None and np.all(np.isfinite(o["mean"])) and o["info"].max() == 0


# ------------------------------------------------------------------------------------------------ ITE / SATE
@pytest.mark.parametrize("n,n_obj,nX,nU,with_u", [(40, 4, 3, 1, True), (100, 5, 2, 2, True), (150, 6, 0, 1, True),
                                                  (64, 4, 3, 1, False), (70, 5, 0, 1, False), (300, 6, 4, 1, True)])
def test_ite_sate_vs_reference_algebra(ctx, n, n_obj, nX, nU, with_u):
    """One augmented Cholesky (CUDA) vs the oracle's line-by-line restatement of likelihood.jl / estimation.jl
    (LU, LU, Bunch-Kaufman + per-draw Cholesky)."""
    counts, X, T, Y = od.synthetic(n, n_obj, max(nX, 1), seed=8)
    if nX == 0:
        X = None
    md = od.model_data_from_arrays(counts if with_u else None, X, T, Y, nU=nU)
    smp = np.stack([oi.posterior(md, 4, 1, 1, seed=5, chain=c, observe_x=True)[0] for c in range(2)], axis=1)
    ret = np.array([1, 3], dtype=np.int32)
    jit, spp, doTs = 1e-10, 3, (0.3, -0.5)
    out = ge.ite(smp, X, T, Y, md.spec.nU, doTs, ret, jit, spp, seed=9, want_cov=True, ctx=ctx)
    so = ge.sate(smp, X, T, Y, md.spec.nU, doTs, ret, jit, spp, seed=9, ctx=ctx)
    assert out["info"].max() == 0 and so["info"].max() == 0
    for d, doT in enumerate(doTs):
        for c in range(2):
            M, Cv = oe.ite_distributions(md.spec, smp[:, c, :], X, T, Y, doT, 2, 2, jit)
            assert np.max(np.abs(M - out["mean"][d, c])) <= 1e-8 * np.max(np.abs(M))
            assert np.max(np.abs(Cv - out["cov"][d, c])) <= 1e-8 * np.max(np.abs(Cv))
            # draws share the normal stream; CovITE + 1e-10 I is nearly singular so the factors agree only to
            # ~eps*cond: compare on the scale of the covariance
            S = oe.ite_samples(M, Cv, spp, seed=9, chain=c, dot_index=d)
            assert np.max(np.abs(S.T - out["samples"][d, c])) <= 1e-4 * np.sqrt(np.max(np.abs(Cv)))
            ms, vs = oe.sate_distributions(M, Cv)
            assert np.allclose(ms, so["mean"][d, c], rtol=1e-8, atol=1e-12) and np.allclose(vs, so["var"][d, c], rtol=1e-8)
            ss = oe.sate_samples(ms, vs, spp, seed=9, chain=c, dot_index=d)
            assert np.allclose(ss, so["samples"][d, c], rtol=1e-7, atol=1e-12)


def test_ite_full_size_vs_reference_algebra_n1024(ctx):
    """BASELINE c3 size: MeanITE / CovITE / SATE of the fused 2048 x 2048 Cholesky against the oracle's line-by-line restatement
    of likelihood.jl / estimation.jl (two LU solves, Bunch-Kaufman) on one posterior-like sample."""
    n, n_obj, nX = 1024, 16, 10
    counts, X, T, Y = od.synthetic(n, n_obj, nX, seed=1234)
    spec = om.ModelSpec(n, 1, nX, False)
    rng = np.random.default_rng(9)
    rec = np.ones(spec.n_params + n)
    rec[:spec.n_params] = 0.8 + 0.6 * rng.random(spec.n_params)
    rec[2] = 0.2
    rec[spec.n_params:] = np.repeat(rng.standard_normal(n_obj), n // n_obj)
    ret = np.array([0], dtype=np.int32)
    out = ge.ite(rec[None, None, :], X, T, Y, 1, [0.25], ret, 1e-10, 2, want_cov=True, ctx=ctx)
    so = ge.sate(rec[None, None, :], X, T, Y, 1, [0.25], ret, 1e-10, 2, ctx=ctx)
    M, Cv = oe.ite_distributions(spec, rec[None, :], X, T, Y, 0.25, 1, 1, 1e-10)
    assert out["info"].max() == 0
    assert np.max(np.abs(M - out["mean"][0, 0])) <= 1e-8 * np.max(np.abs(M))
    assert np.max(np.abs(Cv - out["cov"][0, 0])) <= 1e-8 * np.max(np.abs(Cv))
    ms, vs = oe.sate_distributions(M, Cv)
    assert np.allclose(ms, so["mean"][0, 0], rtol=1e-8, atol=1e-12) and np.allclose(vs, so["var"][0, 0], rtol=1e-7)


def test_ite_cluster_teams_match_single_cta(ctx, monkeypatch):
    """Team mode (one thread-block cluster per augmented Cholesky, csrc/factor.cuh) only re-partitions the row blocks of a
    panel, so MeanITE, CovITE, info and the draws must be bit-identical to the one-CTA-per-task kernel for every team
    size; ragged n (not a multiple of 64), more tasks than resident clusters is covered by the R x doT x chain product."""
    for n, n_obj, nX in ((300, 6, 4), (521, 1, 2)):
        counts, X, T, Y = od.synthetic(n, n_obj, nX, seed=3)
        md = od.model_data_from_arrays(counts, X, T, Y, nU=1)
        smp = np.stack([oi.posterior(md, 2, 1, 1, seed=11, chain=c, observe_x=True)[0] for c in range(2)], axis=1)
        ret = np.array([0, 1], dtype=np.int32)
        doTs = (0.2, -0.4, 1.0)
        outs = {}
        for team in (1, 2, 4, 8):
            monkeypatch.setenv("GPSLC_TEAM", str(team))
            outs[team] = ge.ite(smp, X, T, Y, 1, doTs, ret, 1e-10, 5, seed=2, want_cov=True, ctx=ctx)
        monkeypatch.delenv("GPSLC_TEAM")
        ref = outs[1]
        assert ref["info"].max() == 0
        M, Cv = oe.ite_distributions(md.spec, smp[:, 1, :], X, T, Y, doTs[2], 1, 1, 1e-10)
        assert np.max(np.abs(M - ref["mean"][2, 1])) <= 1e-8 * np.max(np.abs(M))
        for team in (2, 4, 8):
            for key in ("mean", "cov", "samples", "info"):
                assert np.array_equal(ref[key], outs[team][key]), (n, team, key)


def test_ite_shared_kp_factor_matches_one_cholesky_per_task(ctx, monkeypatch):
    """Several doT values per posterior sample (predictCounterfactualEffects): chol(Kp), L11^-1 Y and the panels' P2 outputs are computed
    once per (chain, sample) and every doT continues from them (factor.cuh PRE, ite_base_kernel) instead of re-factoring Kp inside
    every augmented matrix. Same arithmetic in the same order, so MeanITE, CovITE, draws and info must be BIT-identical to the fused
    path (GPSLC_ITE_SHARE=0), for one-CTA tasks, cluster teams, and the grid-wide factorisation of a single large sample."""
    cases = [(40, 4, 2, 2, 2, None), (300, 6, 3, 2, 2, "2"), (521, 1, 2, 1, 1, "4"), (1100, 4, 3, 1, 1, None)]
    for n, n_obj, nX, C, R, team in cases:
        counts, X, T, Y = od.synthetic(n - n % n_obj, n_obj, nX, seed=23)
        nn = len(T)
        spec = om.ModelSpec(nn, 1, nX, False)
        rng = np.random.default_rng(n)
        smp = np.ones((R, C, spec.n_params + nn))
        smp[:, :, :spec.n_params] = 0.7 + 0.6 * rng.random((R, C, spec.n_params))
        smp[:, :, 2] = 0.25
        smp[:, :, spec.n_params:] = np.repeat(rng.standard_normal((R, C, n_obj)), nn // n_obj, axis=2)
        ret = np.arange(R, dtype=np.int32)
        doTs = (0.3, -0.4, 1.1)
        if team:
            monkeypatch.setenv("GPSLC_TEAM", team)
        monkeypatch.setenv("GPSLC_ITE_SHARE", "0")
        ref = ge.ite(smp, X, T, Y, 1, doTs, ret, 1e-10, 3, seed=6, want_cov=(nn < 600), ctx=ctx)
        monkeypatch.setenv("GPSLC_ITE_SHARE", "1")
        got = ge.ite(smp, X, T, Y, 1, doTs, ret, 1e-10, 3, seed=6, want_cov=(nn < 600), ctx=ctx)
        monkeypatch.delenv("GPSLC_ITE_SHARE")
        if team:
            monkeypatch.delenv("GPSLC_TEAM")
        assert ref["info"].max() == 0 and got["info"].max() == 0
        for key in ("mean", "cov", "samples", "info"):
            if ref[key] is not None:
                assert np.array_equal(ref[key], got[key]), (n, key, np.abs(ref[key] - got[key]).max())
    # model variants: no U / no X / neither, Boolean treatment (doT in {0, 1})
    _, X2, T2, Y2 = od.synthetic(150, 6, 2, seed=29)
    for nU, Xv, Tv, dv in ((0, X2, T2, (0.2, 0.9)), (1, None, T2, (0.2, 0.9)), (0, None, T2, (0.2, 0.9, -1.0)), (1, X2, T2 > np.median(T2), (1.0, 0.0))):
        nX = 0 if Xv is None else Xv.shape[1]
        spec = om.ModelSpec(150, nU, nX, False)
        rec = np.ones((2, 1, spec.n_params + nU * 150))
        rec[:, :, :spec.n_params] = 0.7 + 0.6 * np.random.default_rng(nU + 3 * nX).random((2, 1, spec.n_params))
        rec[:, :, 2] = 0.25
        ret2 = np.arange(2, dtype=np.int32)
        monkeypatch.setenv("GPSLC_ITE_SHARE", "0")
        ref = ge.ite(rec, Xv, Tv, Y2, nU, dv, ret2, 1e-10, 2, seed=3, want_cov=True, ctx=ctx)
        monkeypatch.setenv("GPSLC_ITE_SHARE", "1")
        got = ge.ite(rec, Xv, Tv, Y2, nU, dv, ret2, 1e-10, 2, seed=3, want_cov=True, ctx=ctx)
        monkeypatch.delenv("GPSLC_ITE_SHARE")
        assert ref["info"].max() == 0
        for key in ("mean", "cov", "samples", "info"):
            assert np.array_equal(ref[key], got[key]), (nU, nX, key)
    # a posterior sample whose Kp is not positive definite fails in the shared factor: every doT of that sample reports it
    smp[0, 0, 2] = -30.0
    bad = ge.ite(smp, X, T, Y, 1, doTs, ret, 1e-10, 2, ctx=ctx)
    assert np.all(bad["info"][:, 0, 0] >= 1)


def test_zero_effect_identity_through_c_abi(ctx, kats):
    """doT == T => MeanITE == 0 and CovITE == jitter exactly, for U/X present or absent (test/estimation.jl:6-247)."""
    k = kats["conditionalITE_zero_effect"]
    jit = kats["ITEDistributions_jitter"]["predictionCovarianceNoise"]
    for nU, X in ((1, np.array(k["X"])), (1, None), (0, np.array(k["X"])), (0, None)):
        nX = 0 if X is None else 1
        spec = om.ModelSpec(1, nU, nX, False)
        rec = np.ones(spec.n_params + nU)          # every hyperparameter 1.0, U = 1.0
        o = ge.ite(rec[None, None, :], X, np.array(k["realT"]), np.array([0.37]), nU, [k["doT_real"]],
                   np.array([0], dtype=np.int32), jit, 5, want_cov=True, ctx=ctx)
        assert np.all(o["mean"] == 0.0) and np.all(o["cov"] == jit)
        assert abs(o["samples"].mean()) <= 3 * np.sqrt(jit) and o["samples"].var() <= 10 * jit
        s = ge.sate(rec[None, None, :], X, np.array(k["realT"]), np.array([0.37]), nU, [k["doT_real"]],
                    np.array([0], dtype=np.int32), jit, 5, ctx=ctx)
        assert np.all(s["mean"] == 0.0) and np.allclose(s["var"], jit, rtol=1e-12)


def test_summarize_estimates_on_device(ctx, kats):
    """gpslc_summarize vs the reference's quantile KAT (test/driver.jl:54-71) and vs the oracle (NumPy type-7 quantiles) on
    ragged shapes: one sample, non-power-of-two sample counts, a single individual, several batch elements."""
    k = kats["summarizeEstimates_quantiles"]
    samples = np.array(k["samples"], dtype=float)[None, :]
    for ci, (lo, hi) in k["intervals"].items():
        df = g.summarizeEstimates(samples, credible_interval=float(ci), ctx=ctx)
        assert np.isclose(df["LowerBound"][0], lo, rtol=1e-13) and np.isclose(df["UpperBound"][0], hi, rtol=1e-13)
        assert np.isclose(df["Mean"][0], samples.mean(), rtol=1e-13)
    rng = np.random.default_rng(5)
    for batch, m, n in ((1, 1, 1), (1, 2, 5), (3, 10, 33), (2, 150, 272), (1, 1000, 70), (1, 4097, 3), (4, 10, 2048)):
        x = rng.standard_normal((batch, m, n)) * 3 + 1
        got = ge.summarize(x, 0.9, ctx=ctx)
        for b in range(batch):
            mean, lb, ub = oe.summarize_estimates(x[b].T, 0.9)
            assert np.allclose(got[b, :, 0], mean, rtol=1e-12, atol=1e-13)
            assert np.allclose(got[b, :, 1], lb, rtol=1e-12, atol=1e-13) and np.allclose(got[b, :, 2], ub, rtol=1e-12, atol=1e-13)
    # more than 8192 samples per individual (pooled chains): radix-selection path; ties, negative values, ragged n, both quantiles
    for batch, m, n in ((1, 9000, 2), (2, 20000, 7), (1, 76800, 9)):
        x = rng.standard_normal((batch, m, n)) * 3 + 1
        x[:, ::7, :] = np.round(x[:, ::7, :], 1)                # many exact ties
        for ci in (0.9, 0.5):
            got = ge.summarize(x, ci, ctx=ctx)
            for b in range(batch):
                mean, lb, ub = oe.summarize_estimates(x[b].T, ci)
                assert np.allclose(got[b, :, 0], mean, rtol=1e-11, atol=1e-12)
                assert np.array_equal(got[b, :, 1], lb) or np.allclose(got[b, :, 1], lb, rtol=1e-13, atol=1e-14)
                assert np.allclose(got[b, :, 2], ub, rtol=1e-13, atol=1e-14)


# ------------------------------------------------------------------------------------------------ public API end to end
def test_public_api_shapes_and_golden_gate(ctx, kats):
    """gpslc(csv) -> sampleITE(g, 0.6) -> summarizeEstimates: >= 50% of the 150 individuals' mean ITE inside the
    reference's golden 90% interval (test/driver.jl:46-52, test/test_utils.jl:3-12)."""
    import pandas as pd
    k = kats["NEEC_gate"]
    gobj = g.gpslc(os.path.join(GOLD, k["data"]), seed=1234, ctx=ctx)
    assert len(gobj.posteriorSamples) == 24 and g.getN(gobj) == 150 and g.getNU(gobj) == 1 and gobj.X is None
    ite = g.sampleITE(gobj, k["doT"], ctx=ctx)
    assert ite.shape == (150, 15 * 10)
    actual = g.summarizeEstimates(ite)
    expected = pd.read_csv(os.path.join(GOLD, k["golden"]))
    inside = ((expected["LowerBound"] <= actual["Mean"]) & (actual["Mean"] <= expected["UpperBound"])).mean()
    print("NEEC fraction of mean ITEs inside the golden 90% interval:", inside)
    assert inside >= k["min_fraction_inside"], inside
    sate = g.sampleSATE(gobj, k["doT"], ctx=ctx)
    assert sate.shape == (150,) and np.all(np.isfinite(sate))
    M, Cv = g.ITEDistributions(gobj, k["doT"], ctx=ctx)
    assert M.shape == (15, 150) and Cv.shape == (15, 150, 150)
    ms, vs = g.SATEDistributions(gobj, k["doT"], ctx=ctx)
    assert np.allclose(ms, M.mean(axis=1), rtol=1e-8, atol=1e-12) and np.allclose(vs, Cv.sum(axis=(1, 2)) / 150 ** 2, rtol=1e-7)
    uyLS, xyLS, tyLS, yNoise, yScale, U = g.extractParameters(gobj, 10)
    assert xyLS is None and U.shape == (150, 1) and tyLS > 0


# ---- the reference's 14 golden ITE summaries (test/test_results/*.csv). Only NEEC_sampled_0.6.csv is used by a reference test
# (test/driver.jl:46-52); every file gets an asserted status here. Product configuration (ess_rule 0, current kernel convention,
# default budget), 64 chains, seed 100 - the first row of each block of profiles/golden_table_r02.md, which also shows that no
# combination of slice rule / lengthscale convention / 4x budget changes any verdict.
#   gate   fraction of chains that pass the reference's own gate (>= 50 % of the individuals' mean ITE inside the golden 90 % interval)
#   corr   median correlation of a chain's mean ITEs with the golden Mean column (None where the effects are inside the noise)
#   mean / sd / width   chain-pooled summary statistics of BASELINE.md section 2: mean and sd over individuals of the mean ITE, mean CI width
_GOLDEN_FITS = {}


def _golden_stats(ctx, name, tag, chains=64, seed=100):
    import pandas as pd
    if name not in _GOLDEN_FITS:
        _GOLDEN_FITS[name] = g.gpslc(os.path.join(GOLD, "data", name + ".csv"), seed=seed, n_chains=chains, ctx=ctx)
    gobj = _GOLDEN_FITS[name]
    assert len(gobj.posteriorSamples) == 24
    doT = {"true": True, "false": False}.get(tag, None) if tag in ("true", "false") else float(tag)
    ite = g.sampleITE(gobj, doT, all_chains=True, ctx=ctx)                    # [C, n, R*spp]
    assert ite.shape == (chains, g.getN(gobj), 150) and np.all(np.isfinite(ite))
    exp = pd.read_csv(os.path.join(GOLD, "results", f"{name}_{tag}.csv"))
    lo, hi, gm = exp["LowerBound"].values, exp["UpperBound"].values, exp["Mean"].values
    means = ite.mean(axis=2)
    inside = ((lo[None] <= means) & (means <= hi[None])).mean(axis=1)
    corr = np.array([np.corrcoef(m, gm)[0, 1] for m in means])
    q = np.quantile(ite, [0.05, 0.95], axis=2)
    pooled = means.mean(axis=0)
    return {"gate": float((inside >= 0.5).mean()), "corr": float(np.median(corr)), "mean": float(pooled.mean()), "sd": float(pooled.std(ddof=1)),
            "width": float((q[1] - q[0]).mean()), "g_mean": float(gm.mean()), "g_sd": float(gm.std(ddof=1)), "g_width": float((hi - lo).mean())}


# files the CUDA path reproduces: (name, tag, min gate, min corr, |mean - golden| tolerance, sd ratio range, width ratio range)
_GOLDEN_PASS = [
    ("NEEC_sampled", "0.6", 0.95, None, 0.06, (1.6, 2.6), (0.40, 0.70)),          # the reference's own test; measured gate 1.00, mean -0.218 vs -0.209
    ("NEEC_sampled", "0", 0.85, None, 0.60, (2.0, 3.3), (1.8, 2.8)),              # gate 1.00, mean +0.68 vs +0.29 (wide intervals on both sides)
    ("multiplicative_linear", "0", 0.95, 0.95, 0.05, (0.95, 1.10), (0.68, 0.86)),   # gate 1.00, corr +0.99, mean -0.848 vs -0.849, sd 2.10 vs 2.04
    ("multiplicative_linear", "1", 0.95, 0.95, 0.25, (0.90, 1.03), (0.74, 0.92)),   # gate 1.00, corr +0.98, mean -0.20 vs -0.40
    ("IHDP_sampled", "true", 0.95, 0.97, 0.15, (1.00, 1.14), (0.62, 0.80)),         # binary T: gate 1.00, corr +0.99, mean +2.47 vs +2.39
    ("IHDP_sampled", "false", 0.95, 0.97, 0.05, (0.98, 1.10), (0.60, 0.77)),        # gate 1.00, corr +0.99, mean -1.329 vs -1.322
]
# files it does NOT reproduce, with the measured status (gate, corr) - unexplained without a run of the reference
_GOLDEN_GAP = [
    ("NEEC_sampled", "1", 0.00, -0.21), ("NEEC_sampled", "1.0", 0.00, -0.21),
    ("additive_linear", "0", 0.06, -0.75), ("additive_linear", "1", 0.08, -0.72),
    ("additive_nonlinear", "0", 0.12, +0.83), ("additive_nonlinear", "1", 0.11, +0.91),
    ("multiplicative_nonlinear", "0", 0.48, +0.88), ("multiplicative_nonlinear", "1", 0.20, +0.71),
]


@pytest.mark.parametrize("name,tag,min_gate,min_corr,mean_tol,sd_ratio,width_ratio", _GOLDEN_PASS, ids=[f"{a}_{b}" for a, b, *_ in _GOLDEN_PASS])
def test_reference_golden_file(ctx, name, tag, min_gate, min_corr, mean_tol, sd_ratio, width_ratio):
    st = _golden_stats(ctx, name, tag)
    print(name, tag, st)
    assert st["gate"] >= min_gate, st
    if min_corr is not None:
        assert st["corr"] >= min_corr, st
    assert abs(st["mean"] - st["g_mean"]) <= mean_tol, st
    assert sd_ratio[0] <= st["sd"] / st["g_sd"] <= sd_ratio[1], st
    assert width_ratio[0] <= st["width"] / st["g_width"] <= width_ratio[1], st


@pytest.mark.parametrize("name,tag,gate,corr", _GOLDEN_GAP, ids=[f"{a}_{b}" for a, b, *_ in _GOLDEN_GAP])
def test_reference_golden_file_known_gap(ctx, request, name, tag, gate, corr):
    """Golden files no reference test uses and the CUDA path does not reproduce under ANY of the eight convention combinations of
    profiles/golden_table_r02.md. Two assertions: (i) the mismatch is the recorded one (a silent change in either direction fails the
    test); (ii) the reference's gate itself, as a strict expected failure carrying the measured numbers."""
    st = _golden_stats(ctx, name, tag)
    print(name, tag, st)
    assert abs(st["gate"] - gate) <= 0.25 and abs(st["corr"] - corr) <= 0.15, (st, gate, corr)
    request.applymarker(pytest.mark.xfail(strict=True, reason=f"open parity gap: {100 * gate:.0f} % of 64 chains pass the reference's gate on {name}_{tag}.csv "
                                                               f"(median correlation with the golden means {corr:+.2f}); same under ess_rule 0/1, l / l^2 and 4x "
                                                               "budget (profiles/golden_table_r02.md)"))
    assert st["gate"] >= 0.9, st


def test_data_prep_and_persistence_round_trip_through_the_gpu_path(ctx, tmp_path):
    """SURVEY section 8 row f4, the formats either side of the hot path: prepareData (rows sorted by obj, SigmaU from the object
    counts, src/data.jl:20-70) feeds the sampler, and a saved / re-loaded GPSLCObject (src/io.jl; test/io.jl) drives the estimation
    kernels to exactly the same draws as the object it was saved from."""
    path = os.path.join(GOLD, "data", "minimal.csv")
    SigmaU, obj, X, T, Y = g.prepareData(path)
    assert list(obj) == sorted(obj) and SigmaU.shape == (24, 24) and X.shape == (24, 2)
    h = g.getHyperParameters(); h.nOuter, h.nBurnIn = 6, 3
    a = g.gpslc(path, hyperparams=h, seed=11, ctx=ctx)
    # the same data handed over as (SigmaU-structure, X, T, Y): identical chain
    b = g.gpslc(g.objectCounts(obj), X, T, Y, hyperparams=g.HyperParameters(1, 6, 10, 5, 3, 1, 1e-10), seed=11, ctx=ctx)
    assert np.array_equal(a.posteriorPacked, b.posteriorPacked)
    g.saveGPSLCObject(a, str(tmp_path / "run.gpslc"))
    back = g.loadGPSLCObject(str(tmp_path / "run"))
    assert np.array_equal(back.posteriorPacked, a.posteriorPacked) and back.hyperparams == a.hyperparams
    assert np.array_equal(back.priorparams["SigmaU"], SigmaU)
    assert np.array_equal(g.sampleITE(back, 0.4, ctx=ctx), g.sampleITE(a, 0.4, ctx=ctx))
    assert np.array_equal(g.sampleSATE(back, 0.4, ctx=ctx), g.sampleSATE(a, 0.4, ctx=ctx))
    u1 = g.extractParameters(back, 4); u2 = g.extractParameters(a, 4)
    assert all(np.array_equal(x, y) for x, y in zip(u1, u2))


def test_gpslc_accepts_the_four_csv_shapes(ctx):
    """test/gpslc.jl: full / no covariates / no objects / neither, with nOuter=5, nMHInner=1, nESInner=1."""
    for f, has_u, has_x in (("minimal.csv", True, True), ("no_cov.csv", True, False), ("no_objects.csv", False, True),
                            ("no_objects_no_cov.csv", False, False)):
        h = g.getHyperParameters(); h.nOuter, h.nMHInner, h.nESInner, h.nBurnIn = 5, 1, 1, 2
        gobj = g.gpslc(os.path.join(GOLD, "data", f), hyperparams=h, ctx=ctx)
        assert len(gobj.posteriorSamples) == 5
        assert (gobj.hyperparams.nU is not None) == has_u and (gobj.X is not None) == has_x
        ite = g.sampleITE(gobj, 0.5, samplesPerPosterior=2, ctx=ctx)
        assert ite.shape == (g.getN(gobj), 4 * 2) and np.all(np.isfinite(ite))


def test_predict_counterfactual_effects(ctx):
    """test/prediction.jl on n=1, plus shape/consistency on a small dataset: row d of the sweep == sampleITE at doT_d."""
    gobj = g.gpslc([1], np.ones((1, 1)), np.array([1.0]), np.array([0.42]), ctx=ctx)
    ite, rng_ = g.predictCounterfactualEffects(gobj, 15, minDoT=0.0, maxDoT=1.0, ctx=ctx)
    assert ite.shape == (101, 1, 15 * 15) and -1.0 <= ite.mean() <= 1.0 and len(rng_) == 101
    counts, X, T, Y = od.synthetic(40, 4, 2, seed=6)
    h = g.getHyperParameters(); h.nOuter, h.nBurnIn = 6, 3
    gobj = g.gpslc(counts, X, T, Y, hyperparams=h, ctx=ctx)
    ite, rng_ = g.predictCounterfactualEffects(gobj, 3, fidelity=4, ctx=ctx)
    assert ite.shape == (5, 40, 4 * 3) and rng_[0] == T.min() and rng_[-1] == T.max()
    # row d of the sweep is sampleITE at doT_d: same posterior samples, so the per-individual average of the draws must sit at the
    # mixture mean of the MeanITEs within a few standard errors of the CovITE diagonal (12 draws per individual)
    M0, C0 = g.ITEDistributions(gobj, rng_[2], ctx=ctx)
    se = np.sqrt(np.einsum("rii->ri", C0).mean(axis=0) / 12 + M0.var(axis=0) / 12)
    assert np.all(np.isfinite(ite)) and np.all(np.abs(ite[2].mean(axis=1) - M0.mean(axis=0)) < 6 * se + 1e-6)
    one = g.sampleITE(gobj, rng_[2], samplesPerPosterior=3, ctx=ctx)
    assert one.shape == ite[2].shape and np.allclose(one.mean(axis=1), ite[2].mean(axis=1), atol=8 * se.max())


def test_counterfactual_sweep_shards_concatenate(ctx):
    """The doT values of predictCounterfactualEffects are the units sharded over GPUs (BASELINE config c5): rank-local slices
    (gpslc_ite_slice, global doT index in the RNG key) must concatenate to exactly the unsharded sweep, draws included."""
    counts, X, T, Y = od.synthetic(60, 4, 2, seed=12)
    h = g.getHyperParameters(); h.nOuter, h.nBurnIn = 6, 3
    gobj = g.gpslc(counts, X, T, Y, hyperparams=h, ctx=ctx)
    whole, rng_ = g.predictCounterfactualEffects(gobj, 3, fidelity=6, ctx=ctx)
    for world in (2, 3):
        parts = [g.predictCounterfactualEffects(gobj, 3, fidelity=6, ctx=ctx, world_size=world, rank=r)[0] for r in range(world)]
        assert sum(p.shape[0] for p in parts) == 7
        assert np.array_equal(np.concatenate(parts, axis=0), whole)


def test_subgroup_effect_curve_matches_the_documented_workflow(ctx):
    """docs/src/index.md:101-114 on the example data: `ite, doT = predictCounterfactualEffects(g, nSamples)`, `maITE = ite[:, maIdx, :]`,
    `sate = mean(maITE, dims=2)[:, 1, :]`, `summarizeEstimates(sate)`. The fused device path (gpslc_ite_subset_summary) must equal the
    unfused host arithmetic on the SAME draws (the sweep's draws are keyed by the global doT index): subgroup averages to 1e-12,
    the summary to the quantile KAT's tolerance; shards concatenate; gpslc_subset_mean alone agrees as well."""
    path = os.path.join(GOLD, "data", "NEEC_sampled.csv")
    h = g.getHyperParameters(); h.nOuter, h.nBurnIn = 8, 3
    gobj = g.gpslc(path, hyperparams=h, seed=5, ctx=ctx)
    idx = np.asarray(gobj.obj) == "MA"                                          # the example's `vec(g.obj .== "MA")`
    assert 0 < idx.sum() < len(idx)
    ite, rng_ = g.predictCounterfactualEffects(gobj, 7, fidelity=9, ctx=ctx)    # [d, n, R*7]
    want_sate = ite[:, idx, :].mean(axis=1)                                     # [d, m]
    interval, sate, rng2 = g.subgroupEffectCurve(gobj, idx, 7, fidelity=9, ctx=ctx)
    assert np.array_equal(rng_, rng2) and sate.shape == want_sate.shape
    assert np.max(np.abs(sate - want_sate)) <= 1e-12 * max(1.0, np.abs(want_sate).max())
    lo, hi = np.quantile(want_sate, [0.05, 0.95], axis=1)
    assert np.allclose(interval["Mean"], want_sate.mean(axis=1), rtol=1e-12, atol=1e-13)
    assert np.allclose(interval["LowerBound"], lo, rtol=1e-12, atol=1e-13) and np.allclose(interval["UpperBound"], hi, rtol=1e-12, atol=1e-13)
    # the same table through the unfused public calls
    tab = g.summarizeEstimates(want_sate, ctx=ctx)
    assert np.allclose(tab["Mean"], interval["Mean"], rtol=1e-12, atol=1e-13) and np.allclose(tab["UpperBound"], interval["UpperBound"], rtol=1e-12, atol=1e-13)
    # doT shards (BASELINE c5 layout) concatenate to the whole curve
    parts = [g.subgroupEffectCurve(gobj, idx, 7, fidelity=9, ctx=ctx, world_size=3, rank=r) for r in range(3)]
    assert np.array_equal(np.concatenate([p[1] for p in parts], axis=0), sate)
    assert np.array_equal(np.concatenate([p[0]["LowerBound"] for p in parts]), interval["LowerBound"])
    # stand-alone subgroup mean on the gpslc_ite layout [batch][m][n]
    from gpslc_b200 import estimation as ge_
    lay = np.ascontiguousarray(np.swapaxes(ite, 1, 2))
    assert np.max(np.abs(ge_.subset_mean(lay, idx, ctx=ctx) - want_sate)) <= 1e-12 * max(1.0, np.abs(want_sate).max())
    with pytest.raises(Exception):
        ge_.subset_mean(lay, np.zeros(len(idx), bool), ctx=ctx)                 # empty subgroup is an argument error


def test_device_resident_buffers_through_the_c_abi(ctx):
    """GPSLC_DEVICE (include/gpslc.h): the estimation entry points take and return DEVICE pointers, so a counterfactual sweep is
    summarised where gpslc_ite left it in HBM. Device-mode results must equal host-mode results bit for bit: gpslc_ite (mean, draws,
    info) -> gpslc_summarize and gpslc_subset_mean on the device draws; gpslc_ite_subset_summary with device outputs."""
    import ctypes
    import torch
    from gpslc_b200 import DEVICE
    from gpslc_b200.estimation import _bind_est, _data_struct
    from gpslc_b200.inference import _bind
    _bind(ctx.lib); _bind_est(ctx.lib)
    n, nX = 90, 2
    counts, X, T, Y = od.synthetic(n, 3, nX, seed=21)
    h = g.getHyperParameters(); h.nOuter, h.nBurnIn = 5, 2
    gobj = g.gpslc(counts, X, T, Y, hyperparams=h, seed=4, ctx=ctx)
    packed = np.ascontiguousarray(gobj.posteriorPacked[:, :1])
    n_outer, C, stride = packed.shape
    ret = np.arange(1, 5, dtype=np.int32); R, spp = len(ret), 3
    doT = np.array([0.2, -0.4]); D = len(doT)
    host = ge.ite(packed, X, T, Y, 1, doT, ret, 1e-10, spp, seed=9, ctx=ctx)
    dev = torch.device("cuda", 0)
    dt = lambda a, dtype=torch.float64: torch.as_tensor(np.ascontiguousarray(a), dtype=dtype, device=dev)
    Xd, Td, Yd, Sd = dt(np.asfortranarray(X).T.copy()), dt(T), dt(Y), dt(packed)     # X column-major n x nX
    d, _keep = _data_struct(X, T, Y, 1)
    d.X, d.T, d.Y = Xd.data_ptr(), Td.data_ptr(), Yd.data_ptr()
    mean_d = torch.empty((D, C, R, n), dtype=torch.float64, device=dev)
    draws_d = torch.empty((D, C, R * spp, n), dtype=torch.float64, device=dev)
    info_d = torch.empty((D, C, R), dtype=torch.int32, device=dev)
    ctx.check(ctx.lib.gpslc_ite(ctx.h, DEVICE, ctypes.byref(d), Sd.data_ptr(), n_outer, C, stride, ret.ctypes.data, R, doT.ctypes.data, D,
                                1e-10, spp, 9, 0, mean_d.data_ptr(), None, draws_d.data_ptr(), info_d.data_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(mean_d.cpu().numpy(), host["mean"]) and np.array_equal(draws_d.cpu().numpy(), host["samples"])
    assert int(info_d.abs().max()) == 0
    # summarise / subgroup-average the device draws in place
    summ_d = torch.empty((D * C, n, 3), dtype=torch.float64, device=dev)
    ctx.check(ctx.lib.gpslc_summarize(ctx.h, DEVICE, draws_d.data_ptr(), D * C, R * spp, n, 0.9, summ_d.data_ptr()))
    want = ge.summarize(host["samples"].reshape(D * C, R * spp, n), 0.9, ctx=ctx)
    assert np.array_equal(summ_d.cpu().numpy(), want)
    mask = np.zeros(n, dtype=np.uint8); mask[5:40:3] = 1
    sub_d = torch.empty((D * C, R * spp), dtype=torch.float64, device=dev)
    ctx.check(ctx.lib.gpslc_subset_mean(ctx.h, DEVICE, draws_d.data_ptr(), D * C, R * spp, n, mask.ctypes.data, sub_d.data_ptr()))
    assert np.array_equal(sub_d.cpu().numpy(), ge.subset_mean(host["samples"].reshape(D * C, R * spp, n), mask, ctx=ctx))
    assert np.allclose(sub_d.cpu().numpy(), host["samples"].reshape(D * C, R * spp, n)[:, :, mask.astype(bool)].mean(axis=2), rtol=1e-13, atol=1e-14)
    # fused subgroup summary with device outputs == host-mode call
    hs, hsub, hinfo = ge.ite_subset_summary(packed, X, T, Y, 1, doT, ret, 1e-10, spp, mask, seed=9, ctx=ctx)
    fs_d = torch.empty((C, D, 3), dtype=torch.float64, device=dev); fsub_d = torch.empty((C, R * spp, D), dtype=torch.float64, device=dev)
    ctx.check(ctx.lib.gpslc_ite_subset_summary(ctx.h, DEVICE, ctypes.byref(d), Sd.data_ptr(), n_outer, C, stride, ret.ctypes.data, R,
                                               doT.ctypes.data, D, 0, 1e-10, spp, 9, 0, mask.ctypes.data, 0.9, fsub_d.data_ptr(),
                                               fs_d.data_ptr(), info_d.data_ptr()))
    torch.cuda.synchronize()
    assert np.array_equal(fs_d.cpu().numpy(), hs) and np.array_equal(fsub_d.cpu().numpy(), hsub)
    assert np.allclose(hsub[0].T, host["samples"][:, 0][:, :, mask.astype(bool)].mean(axis=2), rtol=1e-13, atol=1e-14)


def test_binary_treatment_end_to_end_ihdp(ctx):
    """IHDP_sampled.csv (n=272, 6 covariates, 200 objects, Bool T): gpslc -> sampleITE(true/false) -> summarizeEstimates through the
    public API runs the binary-T sampler (logitT slices); ONE default chain must pass the reference's own gate against the golden
    files test/test_results/IHDP_sampled_{true,false}.csv (>= 50 % of the mean ITEs inside the golden 90 % interval; measured 0.99)."""
    import pandas as pd
    gobj = g.gpslc(os.path.join(GOLD, "data", "IHDP_sampled.csv"), seed=7, ctx=ctx)
    assert gobj.T.dtype == np.bool_ and len(gobj.posteriorSamples) == 24
    lt = gobj.posteriorSamples[-1]["logitT"]
    assert lt.shape == (272,) and np.all(np.isfinite(lt))
    for doT, name in ((True, "true"), (False, "false")):
        ite = g.sampleITE(gobj, doT, ctx=ctx)
        assert ite.shape == (272, 150) and np.all(np.isfinite(ite))
        exp = pd.read_csv(os.path.join(GOLD, "results", f"IHDP_sampled_{name}.csv"))
        act = g.summarizeEstimates(ite, ctx=ctx)
        inside = ((exp["LowerBound"] <= act["Mean"]) & (act["Mean"] <= exp["UpperBound"])).mean()
        print("IHDP", name, "fraction inside the golden interval:", inside)
        assert inside >= 0.9
        assert np.corrcoef(act["Mean"], exp["Mean"])[0, 1] >= 0.97
