"""The oracle against every known-answer test the reference's own test-suite holds for the hot path
(tests/golden/reference_kats.json, transcribed from /root/reference/test/*.jl with file:line)."""
import numpy as np
import pytest

from oracle import kernel as ok, model as om, estimation as oe, data as od


def test_rbf_kernel_log_magic(kats):
    k = kats["rbfKernelLog_magic"]
    X = np.array(k["X"], dtype=float)
    assert np.array_equal(ok.rbf_kernel_log(X, X, k["LS"]), np.array(k["expected"], dtype=float))
    assert np.array_equal(ok.rbf_kernel_log_loops(X, X, k["LS"]), np.array(k["expected"], dtype=float))
    # Vector-of-Vectors method (src/kernel.jl:34-42) is the same arithmetic on rows
    assert np.array_equal(ok.rbf_kernel_log_loops([[1, 2], [3, 4], [5, 6]], [[1, 2], [3, 4], [5, 6]], 1), np.array(k["expected"]))


def test_rbf_kernel_log_ones(kats):
    k = kats["rbfKernelLog_ones"]
    X = np.full(k["X_shape"], k["X_fill"])
    assert np.array_equal(ok.rbf_kernel_log(X, X, k["LS"]), np.zeros((10, 10)))


def test_rbf_kernel_log_scalar(kats):
    k = kats["rbfKernelLogScalar_same_point"]
    for v in k["values"]:
        assert ok.rbf_kernel_log_scalar([v], [v], k["LS"]) == k["expected"]
    x = np.random.default_rng(0).random(10)
    assert ok.rbf_kernel_log_scalar(x, x, 0.3) == 0.0  # test/kernel.jl:41-47


def test_rbf_no_half_and_squared_lengthscale():
    # pins "no 1/2 factor, lengthscale squared" (src/kernel.jl:17)
    assert ok.rbf_kernel_log_scalar([3.0], [1.0], 2.0) == -1.0
    assert ok.rbf_kernel_log_scalar([3.0, 0.0], [1.0, 1.0], [2.0, 0.5]) == -(1.0 + 4.0)
    with pytest.raises(AssertionError):
        ok.rbf_kernel_log_scalar([1.0, 2.0], [1.0, 2.0], [1.0, 2.0, 3.0])  # src/kernel.jl:14-16


def test_process_cov(kats):
    for name in ("processCov_scale", "processCov_noise", "processCov_scale_noise"):
        k = kats[name]
        got = ok.process_cov(np.array(k["logCov"]), k["scale"], k.get("noise"))
        assert np.array_equal(got, np.array(k["expected"]))


def test_logit_expit(kats):
    assert ok.logit(kats["logit_half"]["p"]) == 0.0
    assert float(ok.expit(kats["expit_zero"]["x"])) == 0.5


def test_generate_sigma_u(kats):
    k = kats["generateSigmaU"]
    assert np.array_equal(om.generate_sigma_u(k["counts"], k["eps"], k["cov"]), np.array(k["expected"]))


def test_remove_adjacent_and_to_matrix_shape(kats):
    k = kats["removeAdjacent"]
    assert od.remove_adjacent(k["input"]) == k["expected"]
    t = kats["toMatrix_shape"]
    U = [np.random.default_rng(i).random(t["len"]) for i in range(t["n_vectors"])]
    assert om.to_matrix(U, t["n"], t["m"]).shape == (t["n"], t["m"])


def _variants(k):
    U = np.array(k["U"]); X = np.array(k["X"])
    for T, doT in ((np.array(k["realT"]), k["doT_real"]), (np.array(k["binaryT"], dtype=float), float(k["doT_binary"]))):
        yield (None, None, None, None, T, doT)
        yield (None, np.array(k["xyLS"]), None, X, T, doT)
        yield (np.array(k["uyLS"]), None, U, None, T, doT)
        yield (np.array(k["uyLS"]), np.array(k["xyLS"]), U, X, T, doT)


def test_conditional_ite_zero_effect(kats):
    """doT == T  =>  MeanITE == 0 and CovITE == 0 exactly, all 8 variants (test/estimation.jl:6-67)."""
    k = kats["conditionalITE_zero_effect"]
    Y = np.array([0.37])
    for (uyLS, xyLS, U, X, T, doT) in _variants(k):
        m, c = oe.conditional_ite(uyLS, xyLS, k["tyLS"], k["yNoise"], k["yScale"], U, X, T, Y, doT)
        assert np.all(m == 0.0) and np.all(c == 0.0)
        ms, vs = oe.conditional_sate(m, c)     # test/estimation.jl:69-137
        assert ms == 0.0 and vs == 0.0


def test_ite_distributions_jitter(kats):
    """Through a posterior run on n=1 with doT == T the ITE mean is 0 and the covariance is exactly the jitter
    (test/estimation.jl:139-247)."""
    from oracle import inference as oi
    jit = kats["ITEDistributions_jitter"]["predictionCovarianceNoise"]
    md = od.model_data_from_arrays([1], np.ones((1, 1)), np.array([1.0]), np.array([0.42]), nU=1)
    smp, _ = oi.posterior(md, 24, 2, 2, seed=1)
    M, C = oe.ite_distributions(md.spec, smp, md.X, md.T, md.Y, 1.0, 10, 1, jit)
    assert M.shape == (15, 1) and np.all(M == 0.0) and np.all(C == jit)
    s = oe.ite_samples(M, C, 10, seed=3)
    assert s.shape == (1, 150) and abs(s.mean()) <= np.sqrt(jit) and s.var() <= 2 * jit  # test/estimation.jl:251-393


def test_summarize_estimates_quantiles(kats):
    k = kats["summarizeEstimates_quantiles"]
    samples = np.array(k["samples"], dtype=float)[None, :]
    for ci, (lo, hi) in k["intervals"].items():
        mean, lb, ub = oe.summarize_estimates(samples, float(ci))
        assert np.isclose(lb[0], lo) and np.isclose(ub[0], hi)


def test_num_posterior_samples(kats):
    k = kats["numPosteriorSamples"]
    assert len(oe.retained_indices(k["nBurnIn"], k["stepSize"], k["nOuter"])) == k["expected"]
    assert oe.retained_indices(10, 1, 24)[0] == 10  # index nBurnIn itself is retained (src/estimation.jl:78)


def test_default_parameters(kats):
    pr = om.get_prior_parameters()
    assert len(pr) == 29 and pr["sigmaUNoise"] == 1e-13 and pr["sigmaUCov"] == 1.0 and pr["drift"] == 0.5
    assert all(pr[k] == 4.0 for k in pr if k.endswith("Shape") or k.endswith("Scale"))
