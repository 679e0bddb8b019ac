"""World-size-2 gloo test of the multi-GPU host logic: chains are sharded by a contiguous block partition keyed by global
chain id, and the only exchange is the final gather along the chain axis."""
import os
import sys

import numpy as np
import pytest

from gpslc_b200.parallel import shard_chains


def test_shard_chains_partition():
    for total in (1, 7, 512, 1000):
        for world in (1, 2, 3, 8):
            parts = [shard_chains(total, world, r) for r in range(world)]
            assert parts[0][0] == 0 and sum(p[1] for p in parts) == total
            for (o1, n1), (o2, _) in zip(parts, parts[1:]):
                assert o1 + n1 == o2
            assert max(p[1] for p in parts) - min(p[1] for p in parts) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from gpslc_b200.parallel import shard_chains, gather_chain_axis
    dist.init_process_group("gloo", rank=rank, world_size=world)
    total = 5
    off, nloc = shard_chains(total, world, rank)
    # stand-in for the packed samples of this rank's chains: value encodes (outer, GLOBAL chain, slot)
    local = np.zeros((3, nloc, 4))
    for i in range(3):
        for c in range(nloc):
            local[i, c] = 100 * i + 10 * (off + c) + np.arange(4)
    full = gather_chain_axis(local, total, axis=1)
    # counterfactual sweep: the doT values are the sharded units, the [doT, n, 3] summaries are gathered along axis 0
    n_dot = 7
    doff, dcnt = shard_chains(n_dot, world, rank)
    summ = np.stack([np.full((6, 3), float(doff + d)) for d in range(dcnt)]) if dcnt else np.zeros((0, 6, 3))
    allsum = gather_chain_axis(summ, n_dot, axis=0)
    assert allsum.shape == (n_dot, 6, 3) and np.array_equal(allsum[:, 0, 0], np.arange(n_dot, dtype=float))
    q.put((rank, full))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_two_ranks_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.zeros((3, 5, 4))
    for i in range(3):
        for c in range(5):
            want[i, c] = 100 * i + 10 * c + np.arange(4)
    for rank, full in got:
        assert np.array_equal(full, want)
