"""Quick GPU check of the primitives against the oracle (development aid; the real tests live in tests/)."""
import sys, os, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "causalgpslc.jl_b200"))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import gpslc_b200 as g
from oracle import kernel as ok, model as om

rng = np.random.default_rng(0)
X = np.array([[1, 2], [3, 4], [5, 6.]])
print("KAT", g.rbfKernelLog(X, X, 1.0))
for n, D, batch in [(5, 2, 3), (64, 3, 4), (100, 12, 5), (150, 1, 7), (256, 6, 9), (1024, 12, 6)]:
    F = rng.standard_normal((n, D)); ls = 0.5 + rng.random((batch, D)) * 2; sc = 0.5 + rng.random(batch); nz = 0.2 + rng.random(batch)
    y = rng.standard_normal((batch, n))
    K = g.cov_build(F, F, ls, sc, nz)
    Ko = np.stack([ok.process_cov(ok.rbf_kernel_log(F, F, ls[b]), sc[b], nz[b]) for b in range(batch)])
    print(n, D, batch, "cov maxabs", np.abs(K - Ko).max())
    lp, ld, q, info = g.chol_logpdf(Ko, y)
    lpo = np.array([om.mvn_logpdf_chol(y[b], Ko[b]) for b in range(batch)])
    print("   dense logpdf rel", np.abs((lp - lpo) / lpo).max(), info)
    lp2, ld2, q2, info2 = g.rbf_logpdf(F, ls, sc, nz, y)
    print("   fused logpdf rel", np.abs((lp2 - lpo) / lpo).max(), info2, "logdet rel", np.abs(ld2-ld).max())
# non-PD
K = np.eye(70); K[40, 40] = -1.0
print(g.chol_logpdf(K[None], np.ones(70)))
# timing of fused factor at n=1024
ctx = g.kernel.default_context()
n, D, batch = 1024, 12, 592
F = rng.standard_normal((n, D)); ls = 1.0 + rng.random((batch, D)) * 2; sc = 0.5 + rng.random(batch); nz = 0.2 + rng.random(batch)
y = rng.standard_normal(n)
for rep in range(3):
    t = time.time(); lp, *_ = g.rbf_logpdf(F, ls, sc, nz, y); dt = time.time() - t
    print("fused n=1024 batch", batch, "time", dt, "TFLOP/s", batch * (n**3 / 3) / dt / 1e12)
