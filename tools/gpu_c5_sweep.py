"""BASELINE config c5 (or a share of it): n=8192 single-posterior counterfactual sweep, doT values sharded over the ranks (one process
per GPU: `python -m torch.distributed.run --nproc-per-node N tools/gpu_c5_sweep.py [n_doT_total] [n]`; N = 1 works too). The posterior
sample is a SAMPLED one (one chain, one outer iteration at full size); each rank calls gpslc_ite_summary on its block of doT values, so the
draws are summarised where they are produced and only [doT, n, 3] summaries exist on the host (bench.run_c5). Development knobs
GPSLC_TEAM / GPSLC_CTAS_PER_SM select the cluster size and the resident CTAs per SM."""
import json, os, sys
root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, os.path.join(root, "causalgpslc.jl_b200")); sys.path.insert(0, root)
import torch
import torch.distributed as dist
import gpslc_b200 as g
from gpslc_b200 import estimation as ge
import bench

rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
n_dot = int(sys.argv[1]) if len(sys.argv) > 1 else 256
n = int(sys.argv[2]) if len(sys.argv) > 2 else 8192
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = g.Context(local)


def barrier():
    ctx.synchronize(); torch.cuda.synchronize()
    if world > 1:
        dist.barrier()


def max_over_ranks(x):
    if world == 1:
        return float(x)
    t = torch.tensor([x], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


peak, _ = bench.fp64_peak()
out = bench.run_c5(g, ge, ctx, rank, world, bench.default_priors(), peak, barrier, max_over_ranks, n_dot=n_dot, n=n)
if rank == 0:
    out["knobs"] = {k: os.environ.get(k) for k in ("GPSLC_TEAM", "GPSLC_CTAS_PER_SM", "GPSLC_LIB_SUFFIX")}
    print(json.dumps(out))
ctx.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()
