"""Multi-GPU plumbing: independent chains are sharded over ranks (one process per GPU) with no data-path collective;
the only exchange is a final gather of packed samples / summaries. Works with NCCL (GPU) and gloo (CPU
tests)."""
import numpy as np


def shard_chains(total_chains, world_size, rank):
    """Contiguous block partition -> (chain_offset, n_local). Philox streams are keyed by the GLOBAL chain id, so the
    union over ranks is independent of world_size."""
    base, rem = divmod(total_chains, world_size)
    n_local = base + (1 if rank < rem else 0)
    offset = rank * base + min(rank, rem)
    return offset, n_local


def gather_chain_axis(local, total_chains, axis=1, device=None):
    """all_gather of per-rank arrays that differ only along the chain axis; returns the array for all chains on every
    rank. `local` is a NumPy array; communication goes through torch.distributed (NCCL needs `device`)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local
    rank = dist.get_rank()
    sizes = [shard_chains(total_chains, world, r)[1] for r in range(world)]
    mx = max(sizes)
    moved = np.moveaxis(local, axis, 0)
    pad_shape = (mx,) + moved.shape[1:]
    buf = np.zeros(pad_shape, dtype=local.dtype)
    buf[:sizes[rank]] = moved
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device)
    outs = [torch.empty_like(t) for _ in range(world)]
    dist.all_gather(outs, t)
    parts = [o.cpu().numpy()[:sizes[r]] for r, o in enumerate(outs)]
    return np.moveaxis(np.concatenate(parts, axis=0), 0, axis)


class _DevView:
    """Zero-copy view of a library-owned device buffer for torch (`torch.as_tensor` understands __cuda_array_interface__)."""

    def __init__(self, ptr, shape):
        self.__cuda_array_interface__ = {"shape": tuple(int(x) for x in shape), "typestr": "<f8", "data": (int(ptr), False),
                                         "version": 2, "strides": None}


def samples_device_tensor(sampler, device=None):
    """The sampler's packed samples [outer_done, n_chains, stride] as a torch CUDA tensor that ALIASES the library's buffer
    (gpslc_sampler_samples_device): valid until the next run() / close()."""
    import ctypes
    import torch
    done = ctypes.c_int()
    sampler.ctx.check(sampler.ctx.lib.gpslc_sampler_get_samples(sampler.h, 0, None, ctypes.byref(done)))   # also synchronises the stream
    p = sampler.ctx.lib.gpslc_sampler_samples_device(sampler.h)
    dev = device if device is not None else torch.device("cuda", sampler.ctx.device)
    return torch.as_tensor(_DevView(p, (done.value, sampler.n_chains, sampler.stride)), device=dev)


def all_gather_samples_device(sampler, world_size=None, device=None):
    """The one collective of the path: NCCL all_gather of every rank's device-resident packed samples, no host
    hop. All ranks must hold the same number of chains (weak scaling; pad otherwise). Returns [world, outer, chains, stride]
    on the device."""
    import torch
    import torch.distributed as dist
    t = samples_device_tensor(sampler, device)
    world = world_size or (dist.get_world_size() if dist.is_initialized() else 1)
    if world == 1:
        return t[None]
    out = torch.empty((world,) + tuple(t.shape), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(out, t.contiguous())
    return out
