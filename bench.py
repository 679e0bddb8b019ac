#!/usr/bin/env python
"""Benchmark of the GP-SLC hot path on B200 (BASELINE.json metric: MH sweeps/sec summed over chains at n=1024).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (oracle port)

One "step" = one Metropolis-Hastings sweep (the body of `for j = 1:nMHInner`, /root/reference/src/inference.jl:22-45:
S = 6 + nU(2+nX) + 4nX = 58 single-site updates) of EVERY chain on the rank: config c3 of BASELINE.json — synthetic
n=1024, 16 confounder objects, 10-dim X, 512 chains per GPU (weak scaling: chains are independent, no data-path
collective). Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(ROOT, "causalgpslc.jl_b200"))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

WORKLOAD = dict(name="c3", n=1024, n_obj=16, nX=10, nU=1, chains_per_gpu=512)
FP64_PEAK_FILE = os.path.join(ROOT, "profiles", "fp64_peak_r01.json")
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")   # dram bytes per mh_lanes launch from `ncu --set full`


def synthetic(n, n_obj, nX, seed=1234):
    """SURVEY.md §8(d) generator (same as oracle.data.synthetic; restated so the product arm does not import oracle/)."""
    rng = np.random.default_rng(seed)
    m = n // n_obj
    counts = [m] * n_obj
    u_obj = rng.standard_normal(n_obj)
    obj = np.repeat(np.arange(n_obj), m)
    X = rng.standard_normal((n, nX))
    w = rng.standard_normal(nX) / np.sqrt(nX)
    v = rng.standard_normal(nX) / np.sqrt(nX)
    T = 0.5 * u_obj[obj] + 0.3 * (X @ w) + 0.5 * rng.standard_normal(n)
    Y = np.sin(T) + u_obj[obj] + 0.3 * (X @ v) + 0.3 * rng.standard_normal(n)
    return counts, X, T, Y


def default_priors():
    from gpslc_b200.hyperparameters import getPriorParameters
    return getPriorParameters()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for (t, line) in self.rows:
            if t < t0 or t > t1 + 0.1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except Exception:
                continue
            names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def fp64_peak():
    try:
        d = json.load(open(FP64_PEAK_FILE))
        return float(d["dgemm8192_tflops_sustained"]), "profiles/fp64_peak_r01.json (cuBLAS DGEMM 8192^3 sustained on this pool's B200; MEASURED_PEAKS.json has no FP64 entry)"
    except Exception:
        return 37.0, "nominal FP64 fallback (no measured file)"


def ncu_traffic():
    try:
        return float(json.load(open(TRAFFIC_FILE))["dram_bytes_per_launch"])
    except Exception:
        return None


def cpu_sample(steps_sites=6, threads=None):
    """Bounded sample of the SAME workload on the host cores: the first `steps_sites` single-site MH updates of one
    sweep of one chain, with the reference's cost model (every update re-scores the whole model: nX+5 kernel builds and
    nU+nX+2 Choleskys) executed by the NumPy/SciPy oracle port."""
    from oracle import data as od, inference as oi
    w = WORKLOAD
    counts, X, T, Y = od.synthetic(w["n"], w["n_obj"], w["nX"])
    md = od.model_data_from_arrays(counts, X, T, Y, nU=w["nU"])
    st = oi.generate_initial_state(md, 1234, 0)
    sc = oi.Scorer(md, st, "faithful")
    sites = md.spec.mh_sites()

    def run(first, count):
        t = time.perf_counter()
        for s in range(first, first + count):
            name, a, b = sites[s % len(sites)]
            oi.mh_site(md, st, sc, s % len(sites), name, a, b, 1234, 0, s // len(sites))
        return time.perf_counter() - t
    return run, len(sites)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    run, S = cpu_sample()
    sites_per_step = 6
    pos = 0
    for _ in range(args.warmup):
        run(pos, sites_per_step); pos += sites_per_step
    t = 0.0
    for _ in range(args.steps):
        t += run(pos, sites_per_step); pos += sites_per_step
    sweeps = args.steps * sites_per_step / S
    value = sweeps / t
    cores = os.cpu_count()
    sample = (f"{sites_per_step} of the {S} single-site MH updates of one sweep of ONE chain per step at the c3 shape "
              f"(n=1024, nX=10, nU=1), reference cost model (full model re-score per update: 15 kernel builds + 13 Choleskys), "
              f"NumPy/SciPy oracle port, BLAS threads = all {cores} host cores; scaled to sweeps/s")
    line = {"impl": "reference", "metric": "mh_sweeps_per_sec", "value": value, "unit": "sweeps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * t / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "c3: n=1024, 16 objects, nX=10, nU=1; one chain on the host (the reference is single-chain, single-process)"},
            "cpu_baseline": {"value": value, "unit": "sweeps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


def run_ours(args):
    import torch
    import gpslc_b200 as g
    from gpslc_b200.inference import ChainSampler, Posterior

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    w = WORKLOAD
    C = w["chains_per_gpu"]
    counts, X, T, Y = synthetic(w["n"], w["n_obj"], w["nX"])
    pri = default_priors()
    ctx = g.Context(local)
    smp = ChainSampler(pri, X, T, Y, w["nU"], counts, nOuter=24, nMHInner=10, nESInner=5, n_chains=C, seed=1234,
                       chain_offset=rank * C, ctx=ctx)
    stream = torch.cuda.ExternalStream(ctx.stream, device=torch.device("cuda", local))
    S = smp.n_sites

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    for _ in range(args.warmup):
        smp.mh_sweeps(1)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    launches0 = ctx.launches
    barrier()
    t0 = time.time()
    ev[0].record(stream)
    for k in range(args.steps):
        smp.mh_sweeps(1)
        ev[k + 1].record(stream)
    barrier()
    t1 = time.time()
    launches = ctx.launches - launches0
    ms = ev[0].elapsed_time(ev[-1])
    per_launch_ms = [ev[k].elapsed_time(ev[k + 1]) for k in range(args.steps)]
    clk = clocks.stop(t0, t1) if rank == 0 else None
    acc, _ = smp.stats()
    if world > 1:
        tt = torch.tensor([ms], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    value = world * C * args.steps / (ms * 1e-3)

    # ---- e2e: the public API call with HOST buffers (Posterior == gpslc_posterior): data H2D, `generate`, 10 MH sweeps
    # (default nMHInner) of every chain, packed samples D2H — all inside the timed region.
    e2e_sweeps = 10
    def e2e_call(seed):
        return Posterior({**pri, "_obj_counts": counts}, X, T, Y, w["nU"], 1, e2e_sweeps, 0, n_chains=C, seed=seed,
                         chain_offset=rank * C, ctx=ctx)
    e2e_reps = 0 if args.no_e2e else 2
    out = e2e_call(1) if not args.no_e2e else np.zeros(1)
    barrier()
    te = time.perf_counter()
    for r in range(e2e_reps):
        tc = time.perf_counter()
        out = e2e_call(2 + r)
        if rank == 0:
            print(f"e2e call {r}: {time.perf_counter() - tc:.3f} s", file=sys.stderr)
    barrier()
    te = time.perf_counter() - te
    if world > 1:
        tt = torch.tensor([te], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt.item())
    e2e_value = world * C * e2e_sweeps * e2e_reps / te if e2e_reps else None
    h2d = (X.size + T.size + Y.size) * 8 + 4 * len(counts) + 27 * 8
    d2h = out.size * 8

    # ---- second half of BASELINE.json's metric: ITE samples/sec. One ITE sample = one length-n draw from N(MeanITE, CovITE)
    # (src/estimation.jl:105). The current state of every chain serves as one retained posterior sample: C tasks, each one
    # augmented 2n x 2n Cholesky + spp draws, through the host-buffer C ABI (H2D of the samples, D2H of the draws included).
    ite = None
    if not args.no_e2e:
        from gpslc_b200 import estimation as ge
        spp = 10
        packed = smp.state()[None, :, :]
        ret0 = np.zeros(1, dtype=np.int32)
        ite_reps = 2
        # pinned host buffers for the posterior samples going in and the means / draws coming out, reused across calls as a
        # serving loop would; the copies in both directions are inside the timed region
        pin = lambda shape: torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
        packed_pin = pin(packed.shape); packed_pin[...] = packed
        obuf = {"mean": pin((1, C, 1, w["n"])), "samples": pin((1, C, spp, w["n"]))}
        ge.ite(packed_pin, X, T, Y, w["nU"], 0.0, ret0, 1e-10, spp, seed=1, chain_offset=rank * C, ctx=ctx, out=obuf)   # warm-up (sizes the workspace)
        barrier()
        ti = time.perf_counter()
        for r in range(ite_reps):
            o = ge.ite(packed_pin, X, T, Y, w["nU"], 0.0, ret0, 1e-10, spp, seed=2 + r, chain_offset=rank * C, ctx=ctx, out=obuf)
        barrier()
        ti = (time.perf_counter() - ti) / ite_reps
        if world > 1:
            tt = torch.tensor([ti], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ti = float(tt.item())
        nI = w["n"]
        ite = {"value": world * C * spp / ti, "unit": "ITE samples/s", "seconds": ti,
               "workload": f"sampleITE(doT=0) for {C} posterior samples per GPU (one per chain) x {spp} draws at n={nI}: "
                           "one fused 2n x 2n Cholesky per sample; pinned host buffers in, draws out to pinned host memory",
               "h2d_bytes_per_call": int(packed.nbytes + (X.size + T.size + Y.size) * 8), "d2h_bytes_per_call": int(obuf["mean"].nbytes + obuf["samples"].nbytes),
               "tflops": world * C * (8.0 * nI ** 3 / 3.0) / ti / 1e12, "all_pd": bool(o["info"].max() == 0)}

    if rank == 0:
        n = w["n"]
        flops_per_factor = n ** 3 / 3.0 + 2.0 * n * n
        flops_per_launch = C * (S - 1) * flops_per_factor   # the uNoise site needs no factorisation (App. A4)
        avg_launch_s = float(np.mean(per_launch_ms)) * 1e-3
        peak, peak_src = fp64_peak()
        achieved = flops_per_launch / avg_launch_s / 1e12
        line = {"metric": "mh_sweeps_per_sec", "value": value, "unit": "sweeps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": "c3: synthetic n=1024, 16 objects x 64, nX=10, nU=1, 512 chains per GPU; step = 1 MH sweep "
                                       "(58 single-site updates) of every chain; continuous T, default InvGamma(4,4) priors",
                           "chains_per_gpu": C, "sites_per_sweep": S,
                           "l2": "working set (296 resident factor scratch slots x 4.25 MiB = 1.26 GB) exceeds the 126 MB L2; no explicit flush",
                           "seed": 1234},
                "clocks": clk, "gpu_launches": int(launches),
                "e2e": {"value": e2e_value, "unit": "sweeps/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "call": f"Posterior(host X,T,Y; nOuter=1, nMHInner={e2e_sweeps}, nESInner=0, {C} chains) incl. generate + H2D + D2H"},
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                             "traffic": ncu_traffic(), "kernel": "mh_lanes_kernel (fused RBF build + blocked Cholesky + solve, DMMA)",
                             "algorithmic": f"{S - 1} factors x (n^3/3 + 2n^2) flops x {C} chains per launch",
                             "peak_source": peak_src},
                "mh_accept_rate": float(acc.sum() / max(1, C * S * (args.warmup + args.steps))), "ite": ite}
        # ---- CPU baseline on the box's host cores (bounded sample, rank 0 at N=1 only)
        if world == 1 and not args.no_cpu:
            run, Ssites = cpu_sample()
            run(0, 2)
            nsite = Ssites                      # one full sweep of one chain (about 10-15 s of CPU work)
            tc = run(2, nsite)
            line["cpu_baseline"] = {"value": (nsite / Ssites) / tc, "unit": "sweeps/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": f"{nsite} of {Ssites} single-site MH updates of one sweep of one chain at the c3 shape, "
                                              "reference cost model (full model re-score per update), NumPy/SciPy oracle port with "
                                              f"BLAS threads = all {os.cpu_count()} host cores; {tc:.1f} s of CPU work"}
        emit(line)
    smp.close()
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def claim_stdout():
    """Everything libraries print to fd 1 (e.g. NCCL's version banner) goes to stderr; the JSON line alone goes to stdout."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="development: skip the e2e leg")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
