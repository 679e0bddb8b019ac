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
    SigmaU, obj, X, T, Y = g.prepareData(path)          # five values, like src/data.jl:69
    counts = g.objectCounts(obj)
    c2, o2, X2, T2, Y2 = od.prepare_data(path)
    assert counts == c2 == [25] * 6 and X is None and X2 is None
    assert np.array_equal(T, T2) and np.array_equal(Y, Y2) and list(obj) == list(o2)
    assert list(obj) == sorted(obj)                     # rows sorted by obj (src/data.jl:25)
    assert SigmaU.shape == (150, 150) and SigmaU[0, 24] == 1.0 and SigmaU[0, 25] == 0.0 and SigmaU[3, 3] == 1 + 1e-13
    p2 = os.path.join(os.path.dirname(__file__), "golden", "data", "IHDP_sampled.csv")
    _, obj2, X, T, Y = g.prepareData(p2)
    counts = g.objectCounts(obj2)
    assert T.dtype == np.bool_ and X.shape == (272, 6) and sum(counts) == 272


def test_to_matrix_mirror_matches_oracle():
    U = [np.arange(11, 17.0), np.arange(21, 27.0)]
    assert np.array_equal(g.toMatrix(U, 6, 2), om.to_matrix(U, 6, 2))


# ------------------------------------------------------------------------------------------------ Julia glue vs the header
def _c_prototypes():
    """{name: (return type, [parameter types])} of every function include/gpslc.h declares, comments stripped."""
    import re
    txt = open(os.path.join(os.path.dirname(__file__), "..", "include", "gpslc.h")).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z0-9_ \*]*?)\b(gpslc_[a-z0-9_]+)\s*\(([^;{}]*?)\)\s*;", txt, flags=re.S):
        ret, name, params = m.group(1).strip(), m.group(2), m.group(3)
        plist = []
        for prm in [x.strip() for x in params.split(",")]:
            if prm in ("void", ""):
                continue
            prm = re.sub(r"\s+", " ", prm)
            ty = re.sub(r"\s*[A-Za-z_][A-Za-z0-9_]*$", "", prm) if not prm.endswith("*") else prm    # drop the parameter name
            plist.append(ty.replace(" *", "*").strip())
        protos[name] = (ret, plist)
    return protos


_C2JL = {"int": {"Cint"}, "double": {"Cdouble"}, "uint64_t": {"UInt64"}, "size_t": {"Csize_t"},
         "const double*": {"Ptr{Cdouble}"}, "double*": {"Ptr{Cdouble}"}, "const int*": {"Ptr{Cint}"}, "int*": {"Ptr{Cint}"},
         "unsigned long long*": {"Ptr{Culonglong}"}, "gpslc_ctx*": {"Ptr{Cvoid}"}, "const gpslc_ctx*": {"Ptr{Cvoid}"},
         "gpslc_ctx**": {"Ref{Ptr{Cvoid}}"}, "const gpslc_data*": {"Ref{GpslcData}"}, "const gpslc_prior*": {"Ref{GpslcPrior}"},
         "const gpslc_opts*": {"Ref{GpslcOpts}"}, "const char*": {"Cstring"}, "void": {"Cvoid"},
         "const unsigned char*": {"Ptr{UInt8}"}}


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_glue_ccalls_match_the_header():
    """The Julia glue cannot be executed in this image, so it is checked statically: every `ccall` names a function the header
    declares, with the same number of arguments, matching C types position by position, the right return type, and as many
    actual arguments as declared types. The three structs must list the header's fields in order with matching types."""
    import re
    jl = open(os.path.join(os.path.dirname(__file__), "..", "causalgpslc.jl_b200", "julia", "CausalGPSLCB200.jl")).read()
    protos = _c_prototypes()
    assert "gpslc_posterior" in protos and "gpslc_ite_slice" in protos and len(protos) >= 30
    n_calls = 0
    for m in re.finditer(r"ccall\(\(:(gpslc_[a-z0-9_]+), LIB\[\]\)", jl):
        name = m.group(1)
        # the balanced argument list of this ccall
        i = m.start() + len("ccall")
        depth, j = 0, i
        while True:
            depth += jl[j] == "("
            depth -= jl[j] == ")"
            j += 1
            if depth == 0:
                break
        args = _split_top(jl[i + 1:j - 1])
        ret, tys, actual = args[1], _split_top(args[2].strip()[1:-1]), args[3:]
        assert name in protos, name
        cret, cparams = protos[name]
        assert ret in _C2JL[cret], (name, ret, cret)
        assert len(tys) == len(cparams), (name, len(tys), len(cparams))
        for k, (jt, ct) in enumerate(zip(tys, cparams)):
            assert jt in _C2JL[ct], (name, k, jt, ct)
        assert len(actual) == len(tys), (name, len(actual), len(tys))
        n_calls += 1
    assert n_calls >= 9
    # struct layouts
    hdr = re.sub(r"/\*.*?\*/", " ", open(os.path.join(os.path.dirname(__file__), "..", "include", "gpslc.h")).read(), flags=re.S)
    jmap = {"int": "Cint", "double": "Cdouble", "const double*": "Ptr{Cdouble}", "const int*": "Ptr{Cint}", "uint64_t": "UInt64"}
    for cname, jname in (("gpslc_data", "GpslcData"), ("gpslc_prior", "GpslcPrior"), ("gpslc_opts", "GpslcOpts")):
        body = re.search(r"typedef struct \{([^{}]*)\}\s*" + cname + ";", hdr, flags=re.S).group(1)
        cfields = []
        for decl in [d.strip() for d in body.split(";") if d.strip()]:
            ty, names = re.match(r"^((?:const )?[a-z0-9_]+\s*\*?)\s*(.*)$", re.sub(r"\s+", " ", decl)).groups()
            ty = ty.strip().replace(" *", "*")
            for nm in [x.strip() for x in names.split(",")]:
                arr = re.match(r"(\w+)\[(\d+)\]", nm)
                cfields.append((arr.group(1), f"NTuple{{{arr.group(2)},{jmap[ty]}}}") if arr else (nm, jmap[ty]))
        jbody = re.search(r"struct " + jname + r"\b.*?\n(.*?)\nend", jl, flags=re.S).group(1)
        jfields = re.findall(r"(\w+)::([A-Za-z0-9{},]+)", jbody)
        assert jfields == cfields, (cname, jfields, cfields)
    # the Python ctypes mirror lists the same fields in the same order
    for cname, cls in (("gpslc_data", gi.GpslcData), ("gpslc_prior", gi.GpslcPrior), ("gpslc_opts", gi.GpslcOpts)):
        body = re.search(r"typedef struct \{([^{}]*)\}\s*" + cname + ";", hdr, flags=re.S).group(1)
        names = re.findall(r"(\w+)(?:\[\d+\])?\s*[;,]", body)
        assert [f[0] for f in cls._fields_] == names, (cname, names)
    assert "using LinearAlgebra" in jl and "seed=rand(UInt64)" in jl


def test_sigma_u_structure_detection():
    """The block structure is read off the matrix itself (the model uses the matrix, whatever priorparams says); anything that is
    not a generateSigmaU matrix goes to the dense path."""
    S = g.generateSigmaU([2, 3, 1], 1e-13, 1.0)
    counts, eps, cov = gi.sigma_u_structure(S)
    assert counts == [2, 3, 1] and cov == 1.0 and 1.0 + eps == S[0, 0]
    counts, eps, cov = gi.sigma_u_structure(g.generateSigmaU([4, 4], 0.25, 0.5))
    assert counts == [4, 4] and cov == 0.5 and 1.0 + eps == 1.25
    assert gi.sigma_u_structure(g.generateSigmaU([1, 1, 1], 1e-13, 1.0))[0] == [1, 1, 1]
    bad = S.copy(); bad[0, 4] = bad[4, 0] = 0.3
    assert gi.sigma_u_structure(bad) is None
    rng = np.random.default_rng(0)
    A = rng.standard_normal((6, 6))
    assert gi.sigma_u_structure(A @ A.T + 6 * np.eye(6)) is None
    # make_structs routes a dense SigmaU through gpslc_data.sigma_u_dense and a block one through the counts
    pri = g.getPriorParameters()
    d, *_ , keep = gi.make_structs({**pri, "SigmaU": A @ A.T + 6 * np.eye(6)}, None, np.zeros(6), np.zeros(6), 1, None, 1, 1, 1, 1, 0, 0, 0, 0, 0)
    assert d.sigma_u_dense and d.n_obj == 0
    d, *_ , keep = gi.make_structs({**pri, "SigmaU": g.generateSigmaU([4, 2], 0.25, 0.5)}, None, np.zeros(6), np.zeros(6), 1, None, 1, 1, 1, 1, 0, 0, 0, 0, 0)
    assert not d.sigma_u_dense and d.n_obj == 2 and d.sigma_u_cov == 0.5 and d.sigma_u_eps == 0.25


def test_gpslc_file_round_trip(tmp_path):
    """test/io.jl: save / load equality including the posterior samples; the `.gpslc` extension is optional (src/io.jl:15-17)."""
    from gpslc_b200.types import GPSLCObject, PosteriorSample
    rng = np.random.default_rng(1)
    n, nU, nX, nOuter = 6, 1, 2, 3
    stride = 6 + 4 * nX + 2 * nU + nU * nX + nU * n
    packed = rng.random((nOuter, 2, stride))
    T = rng.random(n) > 0.5
    pri = g.getPriorParameters()
    S = g.generateSigmaU([2, 4], pri["sigmaUNoise"], pri["sigmaUCov"])
    pri["SigmaU"] = S
    views = [PosteriorSample(packed[i, 0], n, nU, nX, True) for i in range(nOuter)]
    obj = np.array(["a", "a", "b", "b", "b", "b"])
    go = GPSLCObject(g.getHyperParameters(), pri, S, obj, rng.random((n, nX)), T, rng.random(n), views, packed, 77)
    for name in ("one", "two.gpslc"):
        g.saveGPSLCObject(go, str(tmp_path / name))
    assert sorted(p.name for p in tmp_path.iterdir()) == ["one.gpslc", "two.gpslc"]
    for name in ("one.gpslc", "two"):
        back = g.loadGPSLCObject(str(tmp_path / name))
        assert back.hyperparams == go.hyperparams and back.seed == 77
        assert np.array_equal(back.posteriorPacked, packed) and np.array_equal(back.SigmaU, S) and np.array_equal(back.X, go.X)
        assert back.T.dtype == np.bool_ and np.array_equal(back.T, T) and np.array_equal(back.Y, go.Y)
        assert {k: v for k, v in back.priorparams.items() if k != "SigmaU"} == {k: v for k, v in pri.items() if k != "SigmaU"}
        assert np.array_equal(back.priorparams["SigmaU"], S)
        assert len(back.posteriorSamples) == nOuter and back.posteriorSamples[1]["tyLS"] == packed[1, 0, 3]
        assert np.array_equal(g.objectCounts(back.obj), [2, 4])
    with pytest.raises(ValueError):
        (tmp_path / "junk.gpslc").write_bytes(b"not a gpslc file at all")
        g.loadGPSLCObject(str(tmp_path / "junk"))


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the reference's algorithm on the host cores, oracle port): ONE JSON line on stdout with the GPU
    arm's metric / unit / config, `impl`, a `cpu_baseline` describing the run and an `e2e` equal to the value with zero copy bytes."""
    import json, subprocess, sys
    root = os.path.join(os.path.dirname(__file__), "..")
    sys.path.insert(0, root)
    import bench
    env = dict(os.environ); env.pop("RANK", None)
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                       capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "mh_sweeps_per_sec" and d["unit"] == "sweeps/s" and d["higher_is_better"] is True
    assert d["config"] == bench.bench_config(58) and d["dtype"] == "f64" and d["vs_baseline"] is None
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] == os.cpu_count() and cb["value"] == d["value"] > 0 and "chains" in cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # ranks other than 0 of a torchrun launch do no work and print nothing
    env["RANK"] = "1"
    r1 = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0"],
                        capture_output=True, text=True, timeout=120, env=env)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
