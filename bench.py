#!/usr/bin/env python
"""Benchmark of the GP-SLC hot path on B200 (BASELINE.json metric: MH sweeps/sec summed over chains at n=1024).

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --steps K --warmup W    # the reference algorithm on the host cores (oracle port): one chain of the
                                                           # workload per core, one BLAS thread each, sweeps/s summed over them

One "step" = one Metropolis-Hastings sweep (the body of `for j = 1:nMHInner`, /root/reference/src/inference.jl:22-45:
S = 6 + nU(2+nX) + 4nX = 58 single-site updates) of EVERY chain on the rank: config c3 of BASELINE.json — synthetic
n=1024, 16 confounder objects, 10-dim X, 512 chains per GPU (weak scaling: chains are independent, no data-path
collective). Prints ONE JSON line (rank 0).

Keys beyond the contract:
  e2e      the reference's call sequence through the host-buffer API: Posterior(...) with the DEFAULT inner budget
           (nMHInner=10, nESInner=5; only nOuter is shortened) followed by sampleSATE(doT=0, samplesPerPosterior=10) for every
           chain; at N > 1 the packed samples are gathered over NCCL (device buffers) inside the timed region.
  c2, c4   the other single-GPU BASELINE configurations (sweeps/s and fraction of the FP64 peak); c5 when N == 8.
  strong   (N > 1) the same kernel with 512 chains TOTAL, i.e. 512/N per GPU.
  cpu_baseline           the oracle port in the reference's cost model, as many chains of the workload as host cores (one process and
                         one BLAS thread per chain); .variants: ONE chain with all cores / one BLAS thread, and the incremental cost model.
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
TRAFFIC_FILE = os.path.join(ROOT, "profiles", "ncu_traffic_r02.json")   # dram bytes per mh_lanes launch from `ncu --set full`
TRAFFIC_FILE_OLD = os.path.join(ROOT, "profiles", "ncu_traffic_r01.json")


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
    for f in (TRAFFIC_FILE, TRAFFIC_FILE_OLD):
        try:
            return float(json.load(open(f))["dram_bytes_per_launch"])
        except Exception:
            continue
    return None


def bench_config(S=58):
    """The `config` object of both arms (ours and --impl reference): the workload is the same, what each arm executes of it per step is
    said in `cpu_baseline.sample` on the reference side."""
    return {"workload": "c3: synthetic n=1024, 16 objects x 64, nX=10, nU=1, 512 chains per GPU; step = 1 MH sweep "
                        "(58 single-site updates) of every chain; continuous T, default InvGamma(4,4) priors",
            "chains_per_gpu": WORKLOAD["chains_per_gpu"], "sites_per_sweep": S,
            "l2": "working set (296 resident factor scratch slots x 4.25 MiB = 1.26 GB) exceeds the 126 MB L2; no explicit flush",
            "seed": 1234}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle port)
def cpu_sample(mode="faithful", chain=0):
    """Bounded sample of the SAME workload on the host cores: single-site MH updates of one sweep of one chain executed by the
    NumPy/SciPy oracle port. mode "faithful" = the reference's cost model (every update re-scores the whole model: nX+5 kernel
    builds and nU+nX+2 Choleskys, the U-prior one included — SURVEY.md §3.2); "incremental" = one build + one Cholesky per update
    (the algorithm the CUDA path runs)."""
    from oracle import data as od, inference as oi
    w = WORKLOAD
    counts, X, T, Y = od.synthetic(w["n"], w["n_obj"], w["nX"])
    md = od.model_data_from_arrays(counts, X, T, Y, nU=w["nU"])
    st = oi.generate_initial_state(md, 1234, chain)
    sc = oi.Scorer(md, st, mode)
    sites = md.spec.mh_sites()

    def run(first, count):
        t = time.perf_counter()
        for s in range(first, first + count):
            name, a, b = sites[s % len(sites)]
            oi.mh_site(md, st, sc, s % len(sites), name, a, b, 1234, chain, s // len(sites))
        return time.perf_counter() - t
    return run, len(sites)


def _chain_worker(chain, mode, sites_per_step, warmup, steps, gate, q):
    """One host process = one chain of the workload with ONE BLAS thread (the reference is single-threaded Julia per chain): `warmup`
    untimed and `steps` timed groups of `sites_per_step` single-site updates, all workers released together."""
    try:
        from threadpoolctl import threadpool_limits
        with threadpool_limits(limits=1):
            run, S = cpu_sample(mode, chain)
            pos = 0
            for _ in range(warmup):
                run(pos, sites_per_step); pos += sites_per_step
            gate.wait()
            t0 = time.perf_counter()
            for _ in range(steps):
                run(pos, sites_per_step); pos += sites_per_step
            q.put((chain, time.perf_counter() - t0, S))
    except Exception as e:      # a dead worker must not leave the others at the gate
        try:
            gate.abort()
        except Exception:
            pass
        q.put((chain, None, repr(e)))


def cpu_many_chains(mode, sites_per_step, warmup, steps, workers=None):
    """The many-chain workload on all host cores: `workers` (default: every core) independent chains of it, one process and one BLAS
    thread each. Returns (aggregate sweeps/s, wall seconds of the slowest worker, sites per sweep, workers)."""
    import multiprocessing as mp
    P = workers or os.cpu_count() or 1
    ctxm = mp.get_context("spawn")
    gate, q = ctxm.Barrier(P), ctxm.Queue()
    saved = {k: os.environ.get(k) for k in ("OMP_NUM_THREADS", "OPENBLAS_NUM_THREADS", "MKL_NUM_THREADS")}
    for k in saved:
        os.environ[k] = "1"       # inherited by the spawned interpreters before they load a BLAS
    try:
        procs = [ctxm.Process(target=_chain_worker, args=(c, mode, sites_per_step, warmup, steps, gate, q), daemon=True) for c in range(P)]
        for pr in procs:
            pr.start()
        res = [q.get() for _ in range(P)]
        for pr in procs:
            pr.join(30)
    finally:
        for k, v in saved.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    bad = [r for r in res if r[1] is None]
    if bad:
        raise RuntimeError(f"CPU baseline worker failed: {bad[0][2]}")
    T = max(r[1] for r in res)
    S = res[0][2]
    return P * steps * sites_per_step / S / T, T, S, P


def cpu_variants(budget_s=22.0):
    """The three CPU figures BASELINE.md §4.3 asks for, each a bounded sample scaled to sweeps/s of one chain."""
    from threadpoolctl import threadpool_limits
    cores = os.cpu_count()
    out = {}
    plan = [("faithful_all_cores", "faithful", None, 24), ("incremental_all_cores", "incremental", None, 58),
            ("faithful_1_thread", "faithful", 1, 5), ("incremental_1_thread", "incremental", 1, 20)]
    for key, mode, thr, nsite in plan:
        run, S = cpu_sample(mode)
        if thr is None:
            run(0, 2)
            t = run(2, nsite)
        else:
            with threadpool_limits(limits=thr):
                run(0, 1)
                t = run(1, nsite)
        out[key] = {"value": (nsite / S) / t, "unit": "sweeps/s (one chain)", "blas_threads": cores if thr is None else thr,
                    "sites_timed": nsite, "seconds": t}
    out["note"] = ("faithful = the reference's cost (full model re-score per single-site update: 15 kernel builds + 13 Choleskys incl. the "
                   "U prior); incremental = 1 build + 1 Cholesky per update (what the CUDA path computes). The reference is one chain on "
                   f"one thread: 1-thread x {cores} cores = an ideal many-chain projection of {cores} independent chains")
    return out


def cpu_c1():
    """BASELINE configs[0] on the host: the oracle port's Posterior (24 / 10 / 5, one chain, incremental cost model — the reference's
    own full re-score costs 3x that at this shape) on NEEC_sampled.csv plus the line-by-line ITE distributions at doT = 0.6."""
    from oracle import data as od, inference as oi, estimation as oe
    path = os.path.join(ROOT, "tests", "golden", "data", "NEEC_sampled.csv")
    if not os.path.exists(path):
        return None
    t = time.perf_counter()
    counts, obj, X, T, Y = od.prepare_data(path)
    md = od.model_data_from_arrays(counts, X, T, Y, nU=1)
    smp, _ = oi.posterior(md, 24, 10, 5, seed=2, chain=0, mode="incremental")
    M, Cv = oe.ite_distributions(md.spec, smp, X, T, Y, 0.6, 10, 1, 1e-10)
    oe.ite_samples(M, Cv, 10, seed=2)
    return time.perf_counter() - t


def run_reference(args):
    """--impl reference: the reference's algorithm for this path (oracle port with the reference's cost model; Julia / Gen.jl cannot run
    here) on the box's host cores, on the same workload: as many of its independent chains as there are cores, one process and one
    BLAS thread per chain - the reference itself is single-threaded per chain -, a bounded number of single-site updates per step."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sites_per_step = 3
    value, T, S, P = cpu_many_chains("faithful", sites_per_step, args.warmup, args.steps)
    sample = (f"{P} of the 512 chains (one per host core, one BLAS thread each), {sites_per_step} of the {S} single-site MH updates of a "
              f"sweep per step at the c3 shape (n=1024, nX=10, nU=1), reference cost model (full model re-score per update: 15 kernel "
              f"builds + 13 Choleskys), NumPy/SciPy oracle port; sweeps/s summed over the {P} chains")
    line = {"impl": "reference", "metric": "mh_sweeps_per_sec", "value": value, "unit": "sweeps/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * T / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": bench_config(S),
            "cpu_baseline": {"value": value, "unit": "sweeps/s", "cores": P, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": "sweeps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_ours(args):
    import torch
    import gpslc_b200 as g
    from gpslc_b200 import estimation as ge
    from gpslc_b200.inference import ChainSampler
    from gpslc_b200.parallel import all_gather_samples_device

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    w = WORKLOAD
    C = w["chains_per_gpu"]
    counts, X, T, Y = synthetic(w["n"], w["n_obj"], w["nX"])
    pri = default_priors()
    ctx = g.Context(local)
    stream = torch.cuda.ExternalStream(ctx.stream, device=dev)
    peak, peak_src = fp64_peak()

    def barrier():
        ctx.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return float(x)
        tt = torch.tensor([x], device="cuda", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return float(tt.item())

    def time_sweeps(smp, steps, warmup):
        """`steps` single-sweep launches timed with CUDA events on the library's stream; returns (total ms max over ranks,
        per-launch ms list of this rank)."""
        for _ in range(warmup):
            smp.mh_sweeps(1)
        barrier()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
        ev[0].record(stream)
        for k in range(steps):
            smp.mh_sweeps(1)
            ev[k + 1].record(stream)
        barrier()
        return max_over_ranks(ev[0].elapsed_time(ev[-1])), [ev[k].elapsed_time(ev[k + 1]) for k in range(steps)]

    # ---- headline: c3, one MH sweep of every chain per step
    smp = ChainSampler(pri, X, T, Y, w["nU"], counts, nOuter=24, nMHInner=10, nESInner=5, n_chains=C, seed=1234,
                       chain_offset=rank * C, ctx=ctx)
    S = smp.n_sites
    for _ in range(args.warmup):
        smp.mh_sweeps(1)
    barrier()
    clocks = ClockSampler(local)
    if rank == 0:
        clocks.start()
        time.sleep(0.3)
    launches0 = ctx.launches
    t0 = time.time()
    ms, per_launch_ms = time_sweeps(smp, args.steps, 0)
    t1 = time.time()
    launches = ctx.launches - launches0
    clk = clocks.stop(t0, t1) if rank == 0 else None
    acc, _ = smp.stats()
    value = world * C * args.steps / (ms * 1e-3)
    state_c3 = smp.state() if not args.no_e2e else None
    smp.close()

    def frac_of(n, nX, chains, sec_per_sweep, S_):
        return chains * (S_ - 1) * (n ** 3 / 3.0 + 2.0 * n * n) / sec_per_sweep / 1e12

    # ---- strong-scaling point: the same 512 chains spread over the N GPUs (64 per GPU at N = 8)
    strong = None
    if world > 1 and not args.no_extra:
        Cs = C // world
        s2 = ChainSampler(pri, X, T, Y, w["nU"], counts, 24, 10, 5, n_chains=Cs, seed=1234, chain_offset=rank * Cs, ctx=ctx)
        ms2, _ = time_sweeps(s2, 5, 2)
        s2.close()
        tf = frac_of(w["n"], w["nX"], Cs, ms2 * 1e-3 / 5, S)
        strong = {"chains_total": Cs * world, "chains_per_gpu": Cs, "value": Cs * world * 5 / (ms2 * 1e-3), "unit": "sweeps/s",
                  "ms_per_step": ms2 / 5, "tflops_per_gpu": tf, "frac_per_gpu": tf / peak}

    # ---- e2e: the reference's call sequence with HOST buffers: Posterior with the default inner budget
    # (10 MH sweeps + 5 elliptical-slice passes per outer iteration; nOuter shortened to E_OUTER) + sampleSATE(doT=0, 10 draws per
    # retained sample) for every chain; at N > 1 the packed samples are all-gathered over NCCL from the device buffers.
    e2e = None
    if not args.no_e2e:
        E_OUTER, nMH, nES, spp = 2, 10, 5, 10
        ret = np.arange(E_OUTER, dtype=np.int32)
        pin = lambda shape: torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
        # untimed warm-up of the same call sequence at the smallest budget (sizes the library's workspaces and staging arena, first use
        # of the cluster-team kernels of the slice sampler's tail rounds)
        sw = ChainSampler(pri, X, T, Y, w["nU"], counts, 1, 1, 1, n_chains=C, seed=98, chain_offset=rank * C, ctx=ctx)
        sw.run(1)
        if world > 1:
            all_gather_samples_device(sw, world)     # first use of the all_gather channels
            torch.cuda.synchronize()
        pw = sw.samples()
        sw.close()
        ge.sate(pw, X, T, Y, w["nU"], 0.0, np.zeros(1, dtype=np.int32), 1e-10, spp, seed=98, chain_offset=rank * C, ctx=ctx)
        barrier()
        te = time.perf_counter()
        s3 = ChainSampler(pri, X, T, Y, w["nU"], counts, E_OUTER, nMH, nES, n_chains=C, seed=99, chain_offset=rank * C, ctx=ctx)   # H2D + generate
        s3.run(E_OUTER)
        gather_ms = 0.0
        gathered_bytes = 0
        if world > 1:
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ctx.synchronize()
            g0.record()
            allsamp = all_gather_samples_device(s3, world)          # [world, nOuter, C, stride] on the device, NCCL over NVLink
            g1.record()
            torch.cuda.synchronize()
            gather_ms = g0.elapsed_time(g1)
            gathered_bytes = allsamp.numel() * 8
            if rank == 0:
                host_all = pin(tuple(allsamp.shape))
                torch.from_numpy(host_all).copy_(allsamp)            # the gathered posterior lands on rank 0's host
        packed = s3.samples()                                        # D2H of this rank's [nOuter, C, stride]
        _, ess_ev = s3.stats()
        s3.close()
        so = ge.sate(packed, X, T, Y, w["nU"], 0.0, ret, 1e-10, spp, seed=99, chain_offset=rank * C, ctx=ctx)   # H2D samples, D2H draws
        barrier()
        te = max_over_ranks(time.perf_counter() - te)
        gather_ms = max_over_ranks(gather_ms)
        h2d = (X.size + T.size + Y.size) * 8 * 2 + 4 * len(counts) + 27 * 8 + packed.size * 8
        d2h = packed.size * 8 + so["samples"].size * 8 + so["mean"].size * 16 + so["info"].size * 4 + (gathered_bytes if rank == 0 else 0)
        e2e = {"value": world * C * E_OUTER * nMH / te, "unit": "sweeps/s", "h2d_bytes_per_step": int(h2d / (E_OUTER * nMH)),
               "d2h_bytes_per_step": int(d2h / (E_OUTER * nMH)), "seconds": te,
               "outer_iterations_per_s": world * C * E_OUTER / te,
               "sate_samples": int(so["samples"].size), "sate_all_pd": bool(so["info"].max() == 0),
               "ess_evals_per_slice_mean": float(ess_ev.mean() / (E_OUTER * nES)),
               "gather_ms": gather_ms, "gathered_bytes": int(gathered_bytes),
               "call": f"Posterior(host X,T,Y; nOuter={E_OUTER}, nMHInner={nMH}, nESInner={nES} (the default inner budget), {C} chains per GPU) "
                       f"+ sampleSATE(doT=0, samplesPerPosterior={spp}) for every chain; incl. generate, H2D, D2H"
                       + (", NCCL all_gather of the device-resident samples" if world > 1 else "")}

    # ---- second half of BASELINE.json's metric: ITE samples/sec. One ITE sample = one length-n draw from N(MeanITE, CovITE)
    # (src/estimation.jl:105). The current state of every chain serves as one retained posterior sample: C tasks, each one
    # augmented 2n x 2n Cholesky + spp draws, through the host-buffer C ABI (H2D of the samples, D2H of the draws included).
    ite = None
    if not args.no_e2e:
        spp = 10
        packed = state_c3[None, :, :]
        ret0 = np.zeros(1, dtype=np.int32)
        ite_reps = 2
        pin = lambda shape: torch.empty(shape, dtype=torch.float64, pin_memory=True).numpy()
        packed_pin = pin(packed.shape); packed_pin[...] = packed
        obuf = {"mean": pin((1, C, 1, w["n"])), "samples": pin((1, C, spp, w["n"]))}
        ge.ite(packed_pin, X, T, Y, w["nU"], 0.0, ret0, 1e-10, spp, seed=1, chain_offset=rank * C, ctx=ctx, out=obuf)   # warm-up (sizes the workspace)
        barrier()
        ti = time.perf_counter()
        for r in range(ite_reps):
            o = ge.ite(packed_pin, X, T, Y, w["nU"], 0.0, ret0, 1e-10, spp, seed=2 + r, chain_offset=rank * C, ctx=ctx, out=obuf)
        barrier()
        ti = max_over_ranks((time.perf_counter() - ti) / ite_reps)
        nI = w["n"]
        ite = {"value": world * C * spp / ti, "unit": "ITE samples/s", "seconds": ti,
               "workload": f"sampleITE(doT=0) for {C} posterior samples per GPU (one per chain) x {spp} draws at n={nI}: "
                           "one fused 2n x 2n Cholesky per sample; pinned host buffers in, draws out to pinned host memory",
               "h2d_bytes_per_call": int(packed.nbytes + (X.size + T.size + Y.size) * 8), "d2h_bytes_per_call": int(obuf["mean"].nbytes + obuf["samples"].nbytes),
               "tflops": world * C * (8.0 * nI ** 3 / 3.0) / ti / 1e12, "all_pd": bool(o["info"].max() == 0)}

    # ---- the other BASELINE configurations, device-timed like the headline (rank-local, weak scaling)
    extra = {}
    if not args.no_extra:
        for name, n2, nobj2, nX2, C2, st2, wu2 in (("c2", 256, 4, 5, 1024, 5, 3), ("c4", 4096, 64, 10, 64, 2, 1)):
            c2_, X2, T2, Y2 = synthetic(n2, nobj2, nX2)
            sx = ChainSampler(pri, X2, T2, Y2, 1, c2_, 24, 10, 5, n_chains=C2, seed=1234, chain_offset=rank * C2, ctx=ctx)
            msx, _ = time_sweeps(sx, st2, wu2)
            Sx = sx.n_sites
            sx.close()
            tf = frac_of(n2, nX2, C2, msx * 1e-3 / st2, Sx)
            extra[name] = {"value": world * C2 * st2 / (msx * 1e-3), "unit": "sweeps/s", "ms_per_step": msx / st2, "steps": st2, "warmup": wu2,
                           "tflops_per_gpu": tf, "frac": tf / peak,
                           "workload": f"n={n2}, {nobj2} objects, nX={nX2}, nU=1, {C2} chains per GPU, {Sx} sites per sweep"}
        if rank == 0:
            # c1, the reference's own use case (BASELINE configs[0]): gpslc(NEEC_sampled.csv) with default HyperParameters, ONE chain,
            # then sampleITE(g, 0.6) and summarizeEstimates, end to end through the host mirror of the public API (CSV parsing included)
            neec = os.path.join(ROOT, "tests", "golden", "data", "NEEC_sampled.csv")
            if os.path.exists(neec):
                g.gpslc(neec, seed=1, ctx=ctx)                      # warm-up (workspace sizes of this shape)
                ctx.synchronize()
                tc1 = time.perf_counter()
                gobj = g.gpslc(neec, seed=2, ctx=ctx)
                tfit = time.perf_counter() - tc1
                itec1 = g.sampleITE(gobj, 0.6, ctx=ctx)
                g.summarizeEstimates(itec1, ctx=ctx)
                tc1 = time.perf_counter() - tc1
                extra["c1"] = {"seconds": tc1, "gpslc_seconds": tfit, "mh_sweeps_per_s": 240.0 / tfit, "ite_samples": int(itec1.shape[1]),
                               "workload": "gpslc(NEEC_sampled.csv: n=150, 6 objects, no covariates) with default HyperParameters (24 outer x (10 MH "
                                           "sweeps + 5 slice passes)), 1 chain, + sampleITE(g, 0.6) + summarizeEstimates"}
        if world == 8 or os.environ.get("GPSLC_BENCH_C5"):       # the env switch runs a proportional share (32 doT per GPU) at any N
            extra["c5"] = run_c5(g, ge, ctx, rank, world, pri, peak, barrier, max_over_ranks, n_dot=32 * world)

    if rank == 0:
        n = w["n"]
        flops_per_factor = n ** 3 / 3.0 + 2.0 * n * n
        flops_per_launch = C * (S - 1) * flops_per_factor   # the uNoise site needs no factorisation (App. A4)
        avg_launch_s = float(np.mean(per_launch_ms)) * 1e-3
        achieved = flops_per_launch / avg_launch_s / 1e12
        line = {"metric": "mh_sweeps_per_sec", "value": value, "unit": "sweeps/s", "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": bench_config(S),
                "clocks": clk, "gpu_launches": int(launches), "e2e": e2e,
                "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                             "traffic": ncu_traffic(), "kernel": "mh_lanes_kernel (fused RBF build + blocked Cholesky + solve, DMMA)",
                             "algorithmic": f"{S - 1} factors x (n^3/3 + 2n^2) flops x {C} chains per launch",
                             "peak_source": peak_src},
                "mh_accept_rate": float(acc.sum() / max(1, C * S * (args.warmup + args.steps))), "ite": ite}
        line.update(extra)
        if strong is not None:
            line["strong"] = strong
        # ---- CPU baseline on the box's host cores (bounded samples, rank 0 at N=1 only)
        if world == 1 and not args.no_cpu:
            v = cpu_variants()
            v["c1_seconds"] = cpu_c1()
            mv, mT, mS, mP = cpu_many_chains("faithful", 6, 1, 3)
            v["faithful_many_chains"] = {"value": mv, "unit": "sweeps/s (summed over the chains)", "chains": mP, "blas_threads": 1,
                                         "sites_timed": 18, "seconds": mT}
            line["cpu_baseline"] = {"value": mv, "unit": "sweeps/s", "cores": mP, "kind": "port",
                                    "sample": f"{mP} of the 512 chains, one process with one BLAS thread per host core, 18 of {mS} single-site MH "
                                              "updates of one sweep each at the c3 shape, reference cost model (full model re-score per update, "
                                              "U-prior Cholesky executed), NumPy/SciPy oracle port; sweeps/s summed over the chains; "
                                              f"{mT:.1f} s wall ({mP * mT:.0f} s of CPU work)",
                                    "variants": v}
        emit(line)
    ctx.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_c5(g, ge, ctx, rank, world, pri, peak, barrier, max_over_ranks, n_dot=256, n=8192):
    """BASELINE c5: predictCounterfactualEffects over 256 doT at n=8192 for ONE posterior sample, 32 doT per GPU. The posterior
    sample is a sampled one: a single chain (same seed on every rank, hence the same sample) after one outer iteration."""
    from gpslc_b200.inference import ChainSampler
    from gpslc_b200.parallel import shard_chains
    spp = 10
    counts, X, T, Y = synthetic(n, n // 64, 10)
    t0 = time.perf_counter()
    s = ChainSampler(pri, X, T, Y, 1, counts, 1, 1, 1, n_chains=1, seed=5, ctx=ctx)
    s.run(1)
    smp = s.samples()
    s.close()
    t_fit = time.perf_counter() - t0
    doT = np.linspace(T.min(), T.max(), n_dot)
    off, cnt = shard_chains(n_dot, world, rank)
    ret = np.zeros(1, dtype=np.int32)
    # untimed first call of the same shape: sizes the factor workspace (34 GB of scratch at 32 doT per GPU) and the staging arena
    ge.ite_summary(smp, X, T, Y, 1, doT[off:off + cnt], ret, 1e-10, spp, seed=4, ctx=ctx, dot_offset=off)
    barrier()
    t0 = time.perf_counter()
    summ, info = ge.ite_summary(smp, X, T, Y, 1, doT[off:off + cnt], ret, 1e-10, spp, seed=5, ctx=ctx, dot_offset=off)   # draws stay in HBM
    barrier()
    t = max_over_ranks(time.perf_counter() - t0)
    fl = n_dot * (7.0 * n ** 3 / 3.0) + world * n ** 3 / 3.0
    return {"seconds": t, "tflops_aggregate": fl / t / 1e12, "frac_per_gpu": fl / t / 1e12 / world / peak, "n": n, "n_doT": n_dot,
            "doT_per_gpu": cnt, "spp": spp, "all_pd": bool(info.max() == 0), "posterior_sample_seconds": t_fit,
            "flops": f"{n_dot} x 7 n^3/3 (TRSM + SYRK + CovITE factor per doT) + n^3/3 (Kp) once per GPU",
            "summary_finite": bool(np.isfinite(summ).all())}


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
    ap.add_argument("--no-e2e", action="store_true", help="development: skip the e2e and ITE legs")
    ap.add_argument("--no-extra", action="store_true", help="development: skip the c2 / c4 / c5 / strong-scaling legs")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
