"""Achieved HBM bandwidth of the materialised covariance build (gpslc_cov_build, parity layer; SURVEY.md §8d: HBM-bound, 8 n^2 bytes
written per matrix) with device-resident buffers, timed with CUDA events on the library's stream. Writes gpurun_out/cov_build_r02.json."""
import sys, os, json, ctypes
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import torch
import gpslc_b200 as g
from gpslc_b200._lib import DEVICE
ctx = g.Context(0)
dev = torch.device("cuda", 0)
stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
out = {}
# the write-only roofline of this GPU for the same amount of data: a plain fill of a 2.15 GB buffer (torch's vectorised fill kernel).
# MEASURED_PEAKS.json's 6545.9 GB/s is a COPY (read + write bytes); a pure store stream is measured here beside it.
buf = torch.empty(256, 1024, 1024, dtype=torch.float64, device=dev)
for _ in range(3):
    buf.fill_(1.0)
torch.cuda.synchronize()
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record()
for _ in range(10):
    buf.fill_(1.0)
f1.record()
torch.cuda.synchronize()
fill_ms = f0.elapsed_time(f1) / 10
out["write_only_fill_2.15GB"] = {"ms": fill_ms, "GBps": buf.numel() * 8 / fill_ms / 1e6}
print(f"write-only fill of {buf.numel() * 8 / 1e9:.3f} GB: {fill_ms:.3f} ms -> {buf.numel() * 8 / fill_ms / 1e6:.0f} GB/s")
del buf
cases = ((1024, 12, 256), (4096, 12, 16), (256, 6, 4096), (1024, 6, 256), (1024, 3, 256), (1024, 1, 256), (4096, 1, 16))
if len(sys.argv) > 1:
    cases = [tuple(int(v) for v in a.split(",")) for a in sys.argv[1:]]
for n, D, batch in cases:
    f = torch.randn(D, n, dtype=torch.float64, device=dev)                 # shared features, [D][n]
    ls = (0.8 + torch.rand(batch, D, dtype=torch.float64, device=dev))
    sc = torch.ones(batch, dtype=torch.float64, device=dev); nz = torch.full((batch,), 0.1, dtype=torch.float64, device=dev)
    K = torch.empty(batch, n, n, dtype=torch.float64, device=dev)
    torch.cuda.synchronize()
    call = lambda: ctx.check(ctx.lib.gpslc_cov_build(ctx.h, DEVICE, n, batch, D, f.data_ptr(), f.data_ptr(), 1, ls.data_ptr(), sc.data_ptr(),
                                                     nz.data_ptr(), K.data_ptr()))
    for _ in range(3):
        call()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record(stream)
    for _ in range(reps):
        call()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_ = batch * (8.0 * n * n + 8.0 * n * D)
    out[f"n{n}_D{D}_b{batch}"] = {"ms": ms, "algorithmic_GB": bytes_ / 1e9, "achieved_GBps": bytes_ / ms / 1e6, "frac_of_6545.9": bytes_ / ms / 1e6 / 6545.9,
                                  "frac_of_write_only_fill": bytes_ / ms / 1e6 / out["write_only_fill_2.15GB"]["GBps"]}
    print(f"n={n} D={D} batch={batch}: {ms:.3f} ms per call, {bytes_/1e9:.3f} GB -> {bytes_/ms/1e6:.0f} GB/s ({bytes_/ms/1e6/6545.9:.3f} of the measured HBM copy peak)")
out["note"] = "CUDA-event timing on the library stream, 10 calls after 3 warm-ups, no profiler attached; bytes = 8 n^2 written + 8 n D read per matrix; peak = MEASURED_PEAKS.json hbm_gbs 6545.9 (copy)"
json.dump(out, open(os.path.join(root, "gpurun_out", "cov_build_r02.json"), "w"), indent=1)
