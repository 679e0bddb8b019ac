"""CPU-side checks of the host layer: the C-ABI library loads and exports every symbol include/gpslc.h declares, fails
loudly without a GPU (no fallback), and the host mirror's layout logic agrees with the oracle."""
import ctypes
import os

import numpy as np
import pytest

import gpslc_b200 as g
from gpslc_b200 import _lib, inference as gi, estimation as ge
from oracle import model as om, estimation as oe, data as od


def test_library_exports_every_declared_symbol(lib_built):
    syms = _lib.declared_symbols()
    assert len(syms) >= 25
    for s in syms:
        assert hasattr(lib_built, s), s
    assert lib_built.gpslc_version() >= 100


def test_header_cites_reference_for_each_entry_point():
    hdr = open(os.path.join(os.path.dirname(_lib.LIB_PATH), "..", "..", "include", "gpslc.h")).read()
    for needle in ("src/kernel.jl:24-42", "src/model_likelihood.jl", "src/inference.jl", "src/estimation.jl:66-109",
                   "src/estimation.jl:116-163", "src/driver.jl:59-69"):
        assert needle in hdr


def test_no_gpu_means_loud_failure_not_fallback(lib_built):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is visible here")
    h = ctypes.c_void_p()
    assert lib_built.gpslc_create(0, ctypes.byref(h)) == 5     # GPSLC_ERR_NO_DEVICE
    with pytest.raises(g.GpslcError):
        g.Context(0)
    with pytest.raises(g.GpslcError):
        g.rbfKernelLog(np.ones((3, 2)), np.ones((3, 2)), 1.0)   # the product path never computes on the CPU


def test_product_package_does_not_import_oracle():
    pkg = os.path.dirname(os.path.abspath(g.__file__))
    for root, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                txt = open(os.path.join(root, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt, f


def test_sigma_u_round_trip_and_rejection():
    S = g.generateSigmaU([2, 3, 1], 1e-13, 1.0)
    assert gi.sigma_u_to_counts(S, 1e-13, 1.0) == [2, 3, 1]
    assert np.array_equal(S, om.generate_sigma_u([2, 3, 1]))
    bad = S.copy(); bad[0, 4] = bad[4, 0] = 0.3
    with pytest.raises(ValueError):
        gi.sigma_u_to_counts(bad, 1e-13, 1.0)


def test_struct_packing_matches_header_order():
    pri = g.getPriorParameters()
    pri["tyLSShape"] = 7.0
    pri["uNoiseScale"] = 3.0
    X = np.arange(12.0).reshape(4, 3)
    d, p, o, keep = gi.make_structs(pri, X, np.array([True, False, True, True]), np.ones(4), 2, [2, 2], 24, 10, 5, 8, 99, 16, 1, 0, 1)
    assert (d.n, d.nX, d.nU, d.binary, d.n_obj) == (4, 3, 2, 1, 2)
    assert p.shape[12] == 7.0 and p.scale[0] == 3.0 and p.drift == 0.5
    assert (o.nOuter, o.nMHInner, o.nESInner, o.n_chains, o.seed, o.chain_offset, o.u_layout_mode, o.observe_x) == (24, 10, 5, 8, 99, 16, 1, 1)
    assert keep["X"].flags["F_CONTIGUOUS"] and list(keep["T"]) == [1.0, 0.0, 1.0, 1.0]
    assert gi.PRIOR_FAMILIES == ["uNoise", "xNoise", "tNoise", "yNoise", "xScale", "tScale", "yScale", "uxLS", "utLS", "xtLS",
                                 "uyLS", "xyLS", "tyLS"]


def test_posterior_sample_view_addresses_match_oracle_layout():
    spec = om.ModelSpec(n=5, nU=2, nX=3, binary=False)
    rec = np.arange(spec.n_params + 2 * 5, dtype=float)
    s = g.PosteriorSample(rec, 5, 2, 3, False)
    assert s["tyLS"] == rec[spec.idx("tyLS")] and s["yScale"] == rec[spec.idx("yScale")]
    for k in range(3):
        assert s[("xyLS", k + 1, "LS")] == rec[spec.idx("xyLS", k)]
        assert s[("xNoise", k + 1, "Noise")] == rec[spec.idx("xNoise", k)]
    for i in range(2):
        assert s[("uyLS", i + 1, "LS")] == rec[spec.idx("uyLS", i)]
        assert np.array_equal(s[("U", i + 1, "U")], rec[spec.n_params + i * 5: spec.n_params + (i + 1) * 5])
        for j in range(3):
            assert s[("uxLS", i + 1, j + 1, "LS")] == rec[spec.idx("uxLS", i, j)]


def test_retained_indices_and_defaults(kats):
    k = kats["default_hyperparameters"]
    h = g.getHyperParameters()
    assert (h.nU, h.nOuter, h.nMHInner, h.nESInner, h.nBurnIn, h.stepSize, h.predictionCovarianceNoise) == \
        (k["nU"], k["nOuter"], k["nMHInner"], k["nESInner"], k["nBurnIn"], k["stepSize"], k["predictionCovarianceNoise"])
    r = ge.retained_indices(h.nBurnIn, h.stepSize, h.nOuter)
    assert len(r) == 15 and r[0] == 9 and r[-1] == 23
    assert list(r + 1) == oe.retained_indices(h.nBurnIn, h.stepSize, h.nOuter)
    assert g.getPriorParameters() == om.get_prior_parameters()


def test_prepare_data_matches_oracle(kats):
    path = os.path.join(os.path.dirname(__file__), "golden", "data", "NEEC_sampled.csv")
    SigmaU, obj, X, T, Y, counts = g.prepareData(path)
    c2, o2, X2, T2, Y2 = od.prepare_data(path)
    assert counts == c2 == [25] * 6 and X is None and X2 is None
    assert np.array_equal(T, T2) and np.array_equal(Y, Y2) and list(obj) == list(o2)
    assert list(obj) == sorted(obj)                     # rows sorted by obj (src/data.jl:25)
    assert SigmaU.shape == (150, 150) and SigmaU[0, 24] == 1.0 and SigmaU[0, 25] == 0.0 and SigmaU[3, 3] == 1 + 1e-13
    p2 = os.path.join(os.path.dirname(__file__), "golden", "data", "IHDP_sampled.csv")
    _, _, X, T, Y, counts = g.prepareData(p2)
    assert T.dtype == np.bool_ and X.shape == (272, 6) and sum(counts) == 272


def test_to_matrix_mirror_matches_oracle():
    U = [np.arange(11, 17.0), np.arange(21, 27.0)]
    assert np.array_equal(g.toMatrix(U, 6, 2), om.to_matrix(U, 6, 2))
