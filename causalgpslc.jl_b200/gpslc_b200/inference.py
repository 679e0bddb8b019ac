"""Host mirror of /root/reference/src/inference.jl: `Posterior` for all eight model variants, executed by the CUDA
many-chain sampler (csrc/sampler.cu) through the C ABI."""
import ctypes

import numpy as np

from . import _lib
from ._lib import ptr, HOST

PRIOR_FAMILIES = ["uNoise", "xNoise", "tNoise", "yNoise", "xScale", "tScale", "yScale", "uxLS", "utLS", "xtLS", "uyLS",
                  "xyLS", "tyLS"]


class GpslcData(ctypes.Structure):
    _fields_ = [("n", ctypes.c_int), ("nX", ctypes.c_int), ("nU", ctypes.c_int), ("binary", ctypes.c_int),
                ("X", ctypes.c_void_p), ("T", ctypes.c_void_p), ("Y", ctypes.c_void_p),
                ("n_obj", ctypes.c_int), ("obj_counts", ctypes.c_void_p),
                ("sigma_u_eps", ctypes.c_double), ("sigma_u_cov", ctypes.c_double), ("per_chain_data", ctypes.c_int),
                ("sigma_u_dense", ctypes.c_void_p)]


class GpslcPrior(ctypes.Structure):
    _fields_ = [("shape", ctypes.c_double * 13), ("scale", ctypes.c_double * 13), ("drift", ctypes.c_double)]


class GpslcOpts(ctypes.Structure):
    _fields_ = [("nOuter", ctypes.c_int), ("nMHInner", ctypes.c_int), ("nESInner", ctypes.c_int),
                ("n_chains", ctypes.c_int), ("seed", ctypes.c_uint64), ("chain_offset", ctypes.c_int),
                ("u_layout_mode", ctypes.c_int), ("ess_rule", ctypes.c_int), ("observe_x", ctypes.c_int)]


def _bind(lib):
    if getattr(lib, "_sampler_bound", False):
        return
    vp, i = ctypes.c_void_p, ctypes.c_int
    P = ctypes.POINTER
    lib.gpslc_sampler_create.restype = i
    lib.gpslc_sampler_create.argtypes = [vp, i, P(GpslcData), P(GpslcPrior), P(GpslcOpts), P(vp)]
    lib.gpslc_sampler_destroy.restype = None
    lib.gpslc_sampler_destroy.argtypes = [vp]
    lib.gpslc_sampler_layout.restype = i
    lib.gpslc_sampler_layout.argtypes = [vp, P(i), P(i), P(i), P(i)]
    for name in ("gpslc_sampler_run", "gpslc_sampler_mh_sweeps", "gpslc_sampler_ess_pass"):
        getattr(lib, name).restype = i
        getattr(lib, name).argtypes = [vp, i]
    lib.gpslc_sampler_get_samples.restype = i
    lib.gpslc_sampler_get_samples.argtypes = [vp, i, vp, P(i)]
    lib.gpslc_sampler_samples_device.restype = vp
    lib.gpslc_sampler_samples_device.argtypes = [vp]
    lib.gpslc_sampler_get_state.restype = i
    lib.gpslc_sampler_get_state.argtypes = [vp, i, vp]
    lib.gpslc_sampler_set_state.restype = i
    lib.gpslc_sampler_set_state.argtypes = [vp, i, vp]
    lib.gpslc_sampler_get_terms.restype = i
    lib.gpslc_sampler_get_terms.argtypes = [vp, vp, vp]
    lib.gpslc_sampler_get_stats.restype = i
    lib.gpslc_sampler_get_stats.argtypes = [vp, vp, vp, vp]
    lib.gpslc_posterior.restype = i
    lib.gpslc_posterior.argtypes = [vp, P(GpslcData), P(GpslcPrior), P(GpslcOpts), vp, vp, vp]
    lib._sampler_bound = True


def sigma_u_structure(SigmaU):
    """Detect the block structure generateSigmaU produces (src/utils.jl:17-33) in a dense SigmaU: identity with every object's
    block set to one common `cov` and the whole diagonal to one common value 1 + eps. Returns (counts, eps, cov) with
    fl(1 + eps) == the diagonal value exactly, or None when the matrix is anything else (the library then factors it densely,
    like the reference's generateU does, src/model_prior.jl:27-30)."""
    S = np.asarray(SigmaU, dtype=np.float64)
    n = S.shape[0]
    if S.ndim != 2 or S.shape != (n, n):
        raise ValueError("SigmaU must be a square matrix")
    diag = S[0, 0]
    if not np.all(np.diagonal(S) == diag):
        return None
    counts, cov, i = [], None, 0
    while i < n:
        j = i + 1
        while j < n and S[i, j] != 0.0:
            j += 1
        if j - i > 1:
            if cov is None:
                cov = S[i, i + 1]
            elif S[i, i + 1] != cov:
                return None
        counts.append(j - i)
        i = j
    cov = 0.0 if cov is None else float(cov)
    R = np.zeros((n, n))
    i = 0
    for m in counts:
        R[i:i + m, i:i + m] = cov
        i += m
    R[np.diag_indices(n)] = diag
    if not np.array_equal(R, S) or not (diag - cov > 0.0) or not (cov >= 0.0):
        return None
    return counts, float(diag - 1.0), cov


def sigma_u_to_counts(SigmaU, eps, cov):
    """Object counts of a SigmaU built by generateSigmaU(counts, eps, cov); ValueError for anything else."""
    st = sigma_u_structure(SigmaU)
    from .utils import generateSigmaU
    if st is None or not np.array_equal(generateSigmaU(st[0], eps, cov), np.asarray(SigmaU, dtype=np.float64)):
        raise ValueError("SigmaU is not the block matrix generateSigmaU(counts, sigmaUNoise, sigmaUCov) produces")
    return st[0]


def make_structs(priorparams, X, T, Y, nU, counts, nOuter, nMHInner, nESInner, n_chains, seed, chain_offset,
                 u_layout_mode, ess_rule, observe_x, per_chain_data=False):
    """per_chain_data: T, Y are [n_chains, n] and X is [n_chains, n, nX] (one dataset per chain).
    counts: object counts of a block-structured SigmaU (with priorparams["sigmaUNoise"], ["sigmaUCov"] as its eps / cov), or
    None: priorparams["SigmaU"] is analysed — block matrices go to the closed-form path with the eps / cov found IN the matrix,
    anything else is handed over densely (gpslc_data.sigma_u_dense)."""
    T = np.asarray(T)
    binary = T.dtype == np.bool_
    keep = {}
    keep["T"] = np.ascontiguousarray(T, dtype=np.float64)
    keep["Y"] = np.ascontiguousarray(Y, dtype=np.float64)
    n = keep["T"].shape[-1]
    nX = 0
    if X is not None:
        X = np.asarray(X, dtype=np.float64)
        nX = X.shape[-1]
        # each dataset column-major n x nX
        keep["X"] = np.ascontiguousarray(np.swapaxes(X, -1, -2)) if per_chain_data else np.asfortranarray(X)
    d = GpslcData()
    d.n, d.nX, d.nU, d.binary = n, nX, int(nU or 0), int(binary)
    d.X = keep["X"].ctypes.data if nX else None
    d.T = keep["T"].ctypes.data
    d.Y = keep["Y"].ctypes.data
    d.sigma_u_eps = float(priorparams["sigmaUNoise"])
    d.sigma_u_cov = float(priorparams["sigmaUCov"])
    if nU:
        if counts is None:
            st = sigma_u_structure(priorparams["SigmaU"])
            if st is not None:
                counts, d.sigma_u_eps, d.sigma_u_cov = st
            else:
                keep["SigmaU"] = np.asfortranarray(priorparams["SigmaU"], dtype=np.float64)
                if keep["SigmaU"].shape != (n, n):
                    raise ValueError("SigmaU must be n x n")
                d.sigma_u_dense = keep["SigmaU"].ctypes.data
        if counts is not None:
            keep["counts"] = np.ascontiguousarray(counts, dtype=np.int32)
            d.n_obj = len(counts)
            d.obj_counts = keep["counts"].ctypes.data
    d.per_chain_data = int(bool(per_chain_data))
    p = GpslcPrior()
    for k, fam in enumerate(PRIOR_FAMILIES):
        p.shape[k] = float(priorparams[fam + "Shape"])
        p.scale[k] = float(priorparams[fam + "Scale"])
    p.drift = float(priorparams["drift"])
    o = GpslcOpts()
    o.nOuter, o.nMHInner, o.nESInner = int(nOuter), int(nMHInner or 0), int(nESInner or 0)
    o.n_chains, o.seed, o.chain_offset = int(n_chains), int(seed), int(chain_offset)
    o.u_layout_mode, o.ess_rule, o.observe_x = int(u_layout_mode), int(ess_rule), int(observe_x)
    return d, p, o, keep


class ChainSampler:
    """gpslc_sampler: n_chains independent chains resident on one GPU."""

    def __init__(self, priorparams, X, T, Y, nU, counts, nOuter, nMHInner, nESInner, n_chains=1, seed=0, chain_offset=0,
                 u_layout_mode=0, ess_rule=0, observe_x=0, ctx=None, per_chain_data=False):
        from .kernel import default_context
        self.ctx = ctx or default_context()
        lib = self.ctx.lib
        _bind(lib)
        self.d, self.p, self.o, self._keep = make_structs(priorparams, X, T, Y, nU, counts, nOuter, nMHInner, nESInner,
                                                          n_chains, seed, chain_offset, u_layout_mode, ess_rule, observe_x,
                                                          per_chain_data)
        h = ctypes.c_void_p()
        self.ctx.check(lib.gpslc_sampler_create(self.ctx.h, HOST, ctypes.byref(self.d), ctypes.byref(self.p),
                                                ctypes.byref(self.o), ctypes.byref(h)))
        self.h = h
        a, b, c, e = ctypes.c_int(), ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        lib.gpslc_sampler_layout(self.h, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c), ctypes.byref(e))
        self.n_params, self.stride, self.n_sites, self.n_factors = a.value, b.value, c.value, e.value
        self.n_chains = n_chains
        self.nU = int(nU or 0)

    def close(self):
        if getattr(self, "h", None):
            self.ctx.lib.gpslc_sampler_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, n_outer):
        self.ctx.check(self.ctx.lib.gpslc_sampler_run(self.h, n_outer))

    def mh_sweeps(self, count):
        self.ctx.check(self.ctx.lib.gpslc_sampler_mh_sweeps(self.h, count))

    def ess_pass(self, index):
        self.ctx.check(self.ctx.lib.gpslc_sampler_ess_pass(self.h, index))

    def samples(self):
        done = ctypes.c_int()
        self.ctx.check(self.ctx.lib.gpslc_sampler_get_samples(self.h, HOST, None, ctypes.byref(done)))
        out = np.empty((done.value, self.n_chains, self.stride))
        self.ctx.check(self.ctx.lib.gpslc_sampler_get_samples(self.h, HOST, ptr(out), ctypes.byref(done)))
        return out

    def state(self):
        out = np.empty((self.n_chains, self.stride))
        self.ctx.check(self.ctx.lib.gpslc_sampler_get_state(self.h, HOST, ptr(out)))
        return out

    def set_state(self, packed):
        packed = np.ascontiguousarray(packed, dtype=np.float64)
        assert packed.shape == (self.n_chains, self.stride)
        self.ctx.check(self.ctx.lib.gpslc_sampler_set_state(self.h, HOST, ptr(packed)))

    def terms(self):
        lp = np.empty((self.n_chains, self.n_factors))
        q = np.empty((self.n_chains, max(self.nU, 1)))
        self.ctx.check(self.ctx.lib.gpslc_sampler_get_terms(self.h, ptr(lp), ptr(q)))
        return lp, q[:, :self.nU]

    def stats(self):
        acc = np.empty((self.n_chains, self.n_sites), dtype=np.uint64)
        ev = np.empty(self.n_chains, dtype=np.uint64)
        self.ess_evals_logit = np.empty(self.n_chains, dtype=np.uint64)
        self.ctx.check(self.ctx.lib.gpslc_sampler_get_stats(self.h, ptr(acc), ptr(ev), ptr(self.ess_evals_logit)))
        return acc, ev


def Posterior(priorparams, X, T, Y, nU, nOuter, nMHInner, nESInner, n_chains=1, seed=0, chain_offset=0, u_layout_mode=0,
              ess_rule=0, observe_x=0, ctx=None, return_stats=False):
    """`Posterior(priorparams, X, T, Y, nU, nOuter, nMHInner, nESInner)` (src/inference.jl:4-379; dispatch on
    `X === nothing`, `nU === nothing` and the element type of T picks one of the eight methods). priorparams must carry
    "SigmaU" when nU is not None (src/driver.jl:61). Returns the packed samples [nOuter, n_chains, stride]."""
    counts = priorparams.get("_obj_counts") if nU else None      # None: make_structs analyses priorparams["SigmaU"]
    s = ChainSampler(priorparams, X, T, Y, nU, counts, nOuter, nMHInner, nESInner, n_chains, seed, chain_offset,
                     u_layout_mode, ess_rule, observe_x, ctx)
    try:
        s.run(nOuter)
        out = s.samples()
        if return_stats:
            return out, s.stats()
        return out
    finally:
        s.close()
