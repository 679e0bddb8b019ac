"""CPU oracle for the GP-SLC hot path — TEST INFRASTRUCTURE ONLY.

A NumPy/SciPy FP64 restatement of the reference algorithm (KDL-umass/CausalGPSLC.jl v1.0.1, mounted
read-only at /root/reference while building). Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it. The product path
(``causalgpslc.jl_b200/``) never does and has no CPU fallback.

Pinning status
--------------
* Deterministic pieces written purely in the reference's own Julia (covariance build, SigmaU, the GP
  conditional, the quantile summary) are PINNED against the reference's known-answer tests
  (test/kernel.jl:56-90, test/utils.jl:2-16, test/estimation.jl:6-137, test/driver.jl:54-71); see
  tests/test_oracle_kat.py and tests/golden/reference_kats.json.
* Everything that the reference delegates to Gen.jl 0.4.4 / Distributions 0.25.58 / PDMats 0.11.10
  (MvNormal and InvGamma densities, `mh`, `elliptical_slice`, `generate`) is restated from the published
  mathematical definitions: **parity unpinned** for those pieces — Julia is not available in the build
  container and the reference holds no golden vector for them. The only reference fixture
  that crosses them is the loose NEEC gate (test/driver.jl:46-52), which tests/ reproduce.
"""
