"""BASELINE config c5: n=8192 single-posterior counterfactual sweep, doT values sharded over the ranks (one process per GPU,
`python -m torch.distributed.run --nproc-per-node N tools/gpu_c5_sweep.py [n_doT_total] [n]`), summaries gathered over NCCL.
Each rank: gpslc_ite_slice on its contiguous block of doT values (cluster-team kernel), gpslc_summarize of its draws, then an
all_gather of the [doT, n, 3] summaries (the 168 MB of raw draws never leave the GPUs' hosts)."""
import os, sys, time
import numpy as np
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import torch
import torch.distributed as dist
import gpslc_b200 as g
from gpslc_b200 import estimation as ge
from gpslc_b200.parallel import shard_chains, gather_chain_axis
from bench import synthetic

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
n_dot = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = g.Context(local)
n_obj, nX, nU, spp = n // 64, 10, 1, 10
counts, X, T, Y = synthetic(n, n_obj, nX)
n_params = 6 + 4 * nX + 2 * nU + nU * nX
rng = np.random.default_rng(0)
rec = np.ones(n_params + n)
rec[:n_params] = 0.8 + 0.4 * rng.random(n_params)
rec[2] = 0.3
rec[n_params:] = np.repeat(rng.standard_normal(n_obj), n // n_obj)
doT = np.linspace(T.min(), T.max(), n_dot)
off, cnt = shard_chains(n_dot, world, rank)
ret = np.array([0], dtype=np.int32)
warm = "--cold" not in sys.argv
if warm:    # first call sizes the workspace (34 GB of factor scratch at 32 doT per GPU); the timed call below is the steady state
    ge.ite(rec[None, None, :], X, T, Y, nU, doT[off:off + cnt], ret, 1e-10, spp, ctx=ctx, dot_offset=off)
if world > 1:
    dist.barrier()
t0 = time.perf_counter()
o = ge.ite(rec[None, None, :], X, T, Y, nU, doT[off:off + cnt], ret, 1e-10, spp, seed=5, ctx=ctx, dot_offset=off)
t1 = time.perf_counter()
summ = ge.summarize(o["samples"][:, 0], 0.9, ctx=ctx)                 # [cnt, n, 3]
t2 = time.perf_counter()
allsum = gather_chain_axis(summ, n_dot, axis=0, device=torch.device("cuda", local)) if world > 1 else summ
t3 = time.perf_counter()
tt = torch.tensor([t1 - t0, t2 - t1, t3 - t2], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
if rank == 0:
    fl = n_dot * 8 * n ** 3 / 3
    a, b, c = (float(x) for x in tt)
    print(f"c5 sweep: n={n}, {n_dot} doT over {world} GPU(s) ({cnt} on rank 0), spp={spp}: ITE {a:.2f} s ({fl / a / 1e12:.1f} TFLOP/s aggregate), "
          f"summaries {b:.3f} s, gather {c:.3f} s; info max {int(o['info'].max())}; gathered {allsum.shape}, finite {bool(np.isfinite(allsum).all())}; "
          f"mean ITE at first/last doT {allsum[0, :, 0].mean():+.4f} / {allsum[-1, :, 0].mean():+.4f}")
ctx.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
