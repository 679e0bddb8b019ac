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
            assert abs(q0[c, k] - qw) <= 1e-6 * abs(qw)        # limited by the 1e-13-scale deviations inside an object
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
    with pytest.raises(Exception):
        ge.summarize(np.zeros((1, 9000, 2)), 0.9, ctx=ctx)      # more than 8192 samples per individual: refused loudly


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


def _golden_match(ctx, name, doT, chains=16, seed=100):
    """Per-chain (fraction of mean ITEs inside the golden 90 % interval, correlation of the mean ITEs with the golden means)."""
    import pandas as pd
    gobj = g.gpslc(os.path.join(GOLD, "data", name + ".csv"), seed=seed, n_chains=chains, ctx=ctx)
    assert g.getN(gobj) == 200 and gobj.X.shape == (200, 3) and len(gobj.posteriorSamples) == 24
    ite = g.sampleITE(gobj, float(doT), all_chains=True, ctx=ctx)             # [C, n, R*spp]
    assert ite.shape == (chains, 200, 150) and np.all(np.isfinite(ite))
    exp = pd.read_csv(os.path.join(GOLD, "results", f"{name}_{doT}.csv"))
    means = ite.mean(axis=2)
    inside = ((exp["LowerBound"].values[None] <= means) & (means <= exp["UpperBound"].values[None])).mean(axis=1)
    corr = np.array([np.corrcoef(m, exp["Mean"].values)[0, 1] for m in means])
    return inside, corr


@pytest.mark.parametrize("name,min_corr,all_pass_gate", [("multiplicative_linear", 0.9, True), ("multiplicative_nonlinear", 0.6, False),
                                                         ("additive_nonlinear", 0.7, False)])
def test_synthetic_golden_fixtures(ctx, name, min_corr, all_pass_gate):
    """The reference ships four synthetic n=200 datasets (T, Y, X1..X3, obj) with golden ITE summaries at doT = 0 and 1
    (test/test_results/<name>_{0,1}.csv) but no test that uses them. Default hyperparameters, 16 chains, whole
    path gpslc -> sampleITE. The chains differ from the reference's (Philox vs Julia's Xoshiro) and a default run is only 24
    outer iterations, so the check is distributional: the individuals' mean ITEs must be strongly correlated with the golden
    means in the typical chain, and for multiplicative_linear every chain must also pass the reference's own gate (at least
    half of the mean ITEs inside the golden 90 % interval, test/driver.jl:46-52). Measured match rates for all fixtures, and the
    open question about additive_linear and NEEC at doT = 1, are in DESIGN.md §5 (tools/gpu_golden_probe.py)."""
    for doT in (0, 1):
        inside, corr = _golden_match(ctx, name, doT)
        print(name, doT, "inside: median %.2f min %.2f; corr: median %.2f" % (np.median(inside), inside.min(), np.median(corr)))
        assert np.median(corr) >= min_corr, (name, doT, np.median(corr))
        if all_pass_gate:
            assert inside.min() >= 0.5, (name, doT, inside.min())


@pytest.mark.xfail(reason="open parity question (DESIGN.md §5): with the kernel convention of the reference's current source the mean "
                          "ITEs of additive_linear come out anti-correlated with the untested golden file", strict=False)
def test_synthetic_golden_additive_linear(ctx):
    inside, corr = _golden_match(ctx, "additive_linear", 0)
    assert np.median(corr) >= 0.5


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
    M0 = g.ITEDistributions(gobj, rng_[2], ctx=ctx)[0]
    assert np.all(np.isfinite(ite)) and np.abs(ite[2].mean(axis=1) - M0.mean(axis=0)).max() < 5.0


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


def test_binary_treatment_end_to_end_ihdp(ctx):
    """IHDP_sampled.csv (n=272, 6 covariates, 200 objects, Bool T): gpslc -> sampleITE(true/false) runs the binary-T
    sampler; the golden files test/test_results/IHDP_sampled_{true,false}.csv have no test in the reference,
    so only a loose sanity gate is applied: a majority of mean ITEs inside the golden 90% interval widened by its own width."""
    import pandas as pd
    gobj = g.gpslc(os.path.join(GOLD, "data", "IHDP_sampled.csv"), seed=7, ctx=ctx)
    assert gobj.T.dtype == np.bool_ and len(gobj.posteriorSamples) == 24
    lt = gobj.posteriorSamples[-1]["logitT"]
    assert lt.shape == (272,) and np.all(np.isfinite(lt))
    for doT, name in ((True, "true"), (False, "false")):
        ite = g.sampleITE(gobj, doT, ctx=ctx)
        assert ite.shape == (272, 150) and np.all(np.isfinite(ite))
        exp = pd.read_csv(os.path.join(GOLD, "results", f"IHDP_sampled_{name}.csv"))
        act = g.summarizeEstimates(ite)
        wdt = (exp["UpperBound"] - exp["LowerBound"])
        inside = ((exp["LowerBound"] - wdt <= act["Mean"]) & (act["Mean"] <= exp["UpperBound"] + wdt)).mean()
        print("IHDP", name, "fraction inside widened golden interval:", inside)
        assert inside >= 0.5
