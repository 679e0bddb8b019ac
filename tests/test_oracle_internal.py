"""Internal consistency of the oracle: the pieces restated from mathematical definitions agree
with independent evaluations, and the two cost models compute the same chain."""
import math

import numpy as np
import pytest
import scipy.stats as sst

from oracle import philox as px, kernel as ok, model as om, inference as oi, estimation as oe, data as od


def test_philox_random123_kat():
    # Random123 kat_vectors, philox4x32-10
    assert px.philox4x32(0, 0, 0, 0, 0, 0) == (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)
    assert px.philox4x32(*([0xffffffff] * 6)) == (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)
    assert px.philox4x32(0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344, 0xa4093822, 0x299f31d0) == \
        (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1)
    blocks = np.arange(5, dtype=np.uint64)
    v = px.philox4x32_vec(blocks, 7, 9, 11, 13, 17)
    for b in range(5):
        assert tuple(int(x[b]) for x in v) == px.philox4x32(b, 7, 9, 11, 13, 17)


def test_samplers_distribution():
    s = px.Stream(99, 0, 1, 2)
    z = s.normal_vector(20001)
    assert sst.kstest(z, "norm").pvalue > 1e-3
    g = np.array([px.Stream(5, c, 0, 0).gamma(4.0) for c in range(4000)])
    assert sst.kstest(g, "gamma", args=(4.0,)).pvalue > 1e-3
    ig = np.array([px.Stream(6, c, 0, 0).inv_gamma(4.0, 4.0) for c in range(4000)])
    assert sst.kstest(ig, "invgamma", args=(4.0, 0, 4.0)).pvalue > 1e-3
    # element i of normal_vector is independent of the vector length
    assert np.array_equal(px.Stream(1, 2, 3, 4).normal_vector(7)[:5], px.Stream(1, 2, 3, 4).normal_vector(5))


def test_ig_logpdf_matches_scipy():
    for x, a, b in [(0.7, 4.0, 4.0), (3.1, 2.5, 0.9), (1e-3, 4.0, 4.0)]:
        assert np.isclose(om.ig_logpdf(x, a, b), sst.invgamma.logpdf(x, a, scale=b), rtol=1e-13)
    assert om.ig_logpdf(-1.0, 4.0, 4.0) == -math.inf


def test_mvn_logpdf_matches_scipy():
    rng = np.random.default_rng(0)
    A = rng.standard_normal((30, 30)); K = A @ A.T + 30 * np.eye(30); y = rng.standard_normal(30)
    assert np.isclose(om.mvn_logpdf_chol(y, K), sst.multivariate_normal.logpdf(y, np.zeros(30), K), rtol=1e-12)
    with pytest.raises(np.linalg.LinAlgError):
        om.mvn_logpdf_chol(y, -K)


def test_kernel_vectorised_equals_literal_loops():
    rng = np.random.default_rng(1)
    X = rng.standard_normal((17, 4)); ls = 0.5 + rng.random(4)
    assert np.allclose(ok.rbf_kernel_log(X, X, ls), ok.rbf_kernel_log_loops(X, X, ls), rtol=0, atol=1e-14)
    t = rng.standard_normal(17)
    assert np.allclose(ok.rbf_kernel_log(t, np.full(17, 0.3), 1.7), ok.rbf_kernel_log_loops(t, np.full(17, 0.3), 1.7), atol=1e-15)
    b = (rng.random(9) < 0.5).astype(float)     # Bool treatments are promoted by subtraction (src/kernel.jl:17)
    assert set(np.unique(ok.rbf_kernel_log(b, b, 1.0))) <= {0.0, -1.0}


def test_u_prior_closed_form_vs_extended_precision():
    """the closed form is accurate to FP64 round-off; a dense Cholesky of the nearly singular
    uNoise*SigmaU (what the reference does) is only ~1e-6 relative."""
    import mpmath as mp
    mp.mp.dps = 60
    counts = [5, 7, 4]
    n = sum(counts)
    rng = np.random.default_rng(3)
    un = 1.3
    u = oi.sample_u_prior(px.Stream(1, 0, 0, 0), n, counts, un, 1e-13, 1.0)
    got = om.u_prior_logpdf(u, un, counts, 1e-13, 1.0)
    S = om.generate_sigma_u(counts)
    Sm = mp.matrix(n, n)
    for i in range(n):
        for j in range(n):
            Sm[i, j] = mp.mpf(float(S[i, j])) * mp.mpf(un)
    um = mp.matrix([mp.mpf(float(x)) for x in u])
    quad = (um.T * mp.lu_solve(Sm, um))[0]
    logdet = mp.log(mp.det(Sm))
    want = float(-(n * mp.log(2 * mp.pi) + logdet + quad) / 2)
    assert abs(got - want) <= 1e-11 * abs(want)
    dense = om.mvn_logpdf_chol(u, un * S)
    assert abs(dense - want) <= 1e-3 * abs(want)   # the reference's own evaluation: loose by construction


def test_to_matrix_interleave_quirk():
    """App. B1: reshape(permutedims(hcat(U...)), (n, m)) interleaves unless m == 1."""
    got = om.to_matrix([np.arange(11, 17.0), np.arange(21, 27.0)], 6, 2)
    want = np.array([[11, 14], [21, 24], [12, 15], [22, 25], [13, 16], [23, 26]], dtype=float)
    assert np.array_equal(got, want)
    v = np.arange(5.0)
    assert np.array_equal(om.to_matrix([v], 5, 1)[:, 0], v)
    M = np.arange(12.0).reshape(6, 2)
    assert np.array_equal(om.to_matrix(M, 6, 2), M)     # Matrix input (extractParameters, src/utils.jl:107) is unchanged


@pytest.mark.parametrize("with_u,with_x,binary", [(u, x, b) for u in (1, 0) for x in (1, 0) for b in (0, 1)])
def test_all_eight_variants_run_and_modes_agree(with_u, with_x, binary):
    counts, X, T, Y = od.synthetic(24, 3, 2, seed=4)
    md = od.model_data_from_arrays(counts if with_u else None, X if with_x else None, (T > 0) if binary else T, Y, nU=2)
    a, fa = oi.posterior(md, 3, 2, 2, seed=5, mode="incremental")
    b, fb = oi.posterior(md, 3, 2, 2, seed=5, mode="faithful")
    assert a.shape == (3, oi.sample_stride(md.spec)) and np.array_equal(np.isnan(a), np.isnan(b))
    assert np.allclose(np.nan_to_num(a), np.nan_to_num(b), rtol=1e-9, atol=1e-12)
    assert np.isfinite(om.log_joint(md, fa))
    # the chain moves: every traced hyperparameter takes >1 value (test/inference.jl:9-28)
    long, _ = oi.posterior(md, 10, 7, 6, seed=6)
    for (name, i, j) in md.spec.active_params():
        assert len(np.unique(long[:, md.spec.idx(name, i, j)])) > 1, name


def test_site_counts_match_reference_schedule():
    """S = 6 + nU(2+nX) + 4nX for the full model: 58 at c3, 33 at nX=5, 8 with no X."""
    assert len(om.ModelSpec(10, 1, 10, False).mh_sites()) == 58
    assert len(om.ModelSpec(10, 1, 5, False).mh_sites()) == 33
    assert len(om.ModelSpec(10, 1, 0, False).mh_sites()) == 8
    assert [s[0] for s in om.ModelSpec(10, 0, 0, False).mh_sites()] == ["yNoise", "tyLS", "yScale"]
    names = [s[0] for s in om.ModelSpec(10, 1, 1, False).mh_sites()]
    assert names == ["uNoise", "tNoise", "yNoise", "tyLS", "utLS", "uyLS", "uxLS", "xNoise", "xtLS", "xyLS", "xScale", "tScale", "yScale"]


def test_faithful_cost_model_counts():
    """Reference cost per MH update: nU+nX+2 Choleskys and nX+5 builds."""
    counts, X, T, Y = od.synthetic(20, 2, 3, seed=1)
    md = od.model_data_from_arrays(counts, X, T, Y, nU=1)
    st = oi.generate_initial_state(md, 1, 0)
    sc = oi.Scorer(md, st, "faithful")
    c0, b0 = sc.n_chol, sc.n_build
    name, i, j = md.spec.mh_sites()[1]
    oi.mh_site(md, st, sc, 1, name, i, j, 1, 0, 0)
    assert sc.n_chol - c0 == 1 + 3 + 2 and sc.n_build - b0 == 3 + 5


def test_restructured_gp_conditional_equals_reference_algebra():
    """App. A6 (one augmented Cholesky) == src/likelihood.jl:24-49 + src/estimation.jl:46-47 (LU, LU, Bunch-Kaufman)."""
    counts, X, T, Y = od.synthetic(40, 4, 3, seed=2)
    md = od.model_data_from_arrays(counts, X, T, Y, nU=1)
    smp, _ = oi.posterior(md, 2, 1, 1, seed=2)
    p = oe.extract_parameters(md.spec, smp[-1])
    for doT in (0.0, 0.6, float(T[3])):
        m1, c1 = oe.conditional_ite(*p, X, T, Y, doT)
        m2, c2 = oe.conditional_ite_restructured(*p, X, T, Y, doT, 0.0)
        assert np.allclose(m1, m2, rtol=1e-9, atol=1e-11) and np.allclose(c1, c2, rtol=1e-8, atol=1e-11)


def test_sate_var_as_std_quirk():
    ms, vs = np.array([0.5]), np.array([0.04])
    a = oe.sate_samples(ms, vs, 2000, seed=1, var_as_std=True)
    b = oe.sate_samples(ms, vs, 2000, seed=1, var_as_std=False)
    assert abs(a.std() - 0.04) < 0.004 and abs(b.std() - 0.2) < 0.02   # App. B5


def test_prepare_data_sorts_by_obj(tmp_path):
    import pandas as pd
    df = pd.DataFrame({"T": [0.1, 0.2, 0.3, 0.4], "Y": [1.0, 2.0, 3.0, 4.0], "X1": [5.0, 6.0, 7.0, 8.0], "obj": ["b", "a", "b", "a"]})
    counts, obj, X, T, Y = od.prepare_data(df)
    assert counts == [2, 2] and list(obj) == ["a", "a", "b", "b"] and list(T) == [0.2, 0.4, 0.1, 0.3]   # src/data.jl:25
    counts, obj, X, T, Y = od.prepare_data(df.drop(columns=["obj", "X1"]))
    assert counts is None and obj is None and X is None


def test_gamma_sampler_small_shapes_and_dense_sigma_u():
    """The Philox gamma sampler (shared specification with csrc/rng.cuh) covers shape < 1 through the boost
    Gamma(a) = Gamma(a+1) U^(1/a); a dense SigmaU scores and samples like the closed form of the block matrix."""
    import scipy.stats as sst
    from oracle import philox as px, model as om, inference as oi, data as od
    for shape in (0.3, 0.9, 1.0, 4.0):
        x = np.array([px.Stream(5, c, 1, 2).gamma(shape) for c in range(4000)])
        assert sst.kstest(x, sst.gamma(shape).cdf).pvalue > 1e-3, shape
    counts, X, T, Y = od.synthetic(24, 3, 2, seed=3)
    S = om.generate_sigma_u(counts, 0.25, 0.5)             # well conditioned, so the Cholesky route is accurate
    pri = {**om.get_prior_parameters(), "sigmaUNoise": 0.25, "sigmaUCov": 0.5}
    blk = od.model_data_from_arrays(counts, X, T, Y, nU=1, prior=pri)
    dns = od.model_data_from_arrays(None, X, T, Y, nU=1, prior=pri, sigma_u_dense=S)
    u = np.random.default_rng(0).standard_normal(24)
    assert abs(om.u_prior_logpdf_data(blk, u, 1.3) - om.u_prior_logpdf_data(dns, u, 1.3)) < 1e-10
    assert dns.spec.nU == 1
    a, _ = oi.posterior(dns, 2, 1, 1, seed=3, chain=0)
    assert a.shape == (2, dns.spec.n_params + 24) and np.all(np.isfinite(a[:, dns.spec.n_params:]))
