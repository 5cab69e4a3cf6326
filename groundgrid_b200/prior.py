"""Multi-GPU plumbing of the path.

Scans of different streams are independent, so they shard across ranks with no data-path
collective (`shard`).  The one exchange the path admits is the broadcast of the rolling terrain
prior -- the layers "ground" and "groundpatch" plus the map position -- when several clouds are
evaluated against the same ego frame on different GPUs.  The two layers are contiguous in the
device arena (2 * N * N floats starting at "ground"), so the broadcast is ONE NCCL call on the
handle's own memory (NVLink 5 / NVSwitch, no staging copy).
"""
import numpy as np


def shard(n_items, rank, world):
    """Contiguous block partition: item indices owned by `rank` (sizes differ by at most one)."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return list(range(start, start + base + (1 if rank < extra else 0)))


class _DevicePtr:
    """Exposes raw device memory to torch through __cuda_array_interface__ (no copy)."""

    def __init__(self, ptr, n_floats):
        self.__cuda_array_interface__ = {"shape": (n_floats,), "typestr": "<f4", "data": (int(ptr), False), "version": 3, "strides": None}


def prior_tensor(handle, slot=0):
    """torch view (float32, 2*N*N) of ground||groundpatch of one slot, aliasing the handle's memory."""
    import torch

    n2 = handle.n * handle.n
    ptr = handle.layer_device_ptr("ground", slot)
    ptr_c = handle.layer_device_ptr("groundpatch", slot)
    assert ptr_c == ptr + 4 * n2, "ground and groundpatch must be contiguous"
    # the tensor lives on the HANDLE's device (not on whatever device happens to be current in this process)
    dev = getattr(handle, "device", None)
    if dev is None:
        dev = torch.cuda.current_device()
    with torch.cuda.device(dev):
        return torch.as_tensor(_DevicePtr(ptr, 2 * n2), device=f"cuda:{dev}")


def broadcast_prior_tensors(prior, position, src, group=None):
    """Collective: after the call every rank holds src's prior (flat float32 tensor) and map position (2 float64)."""
    import torch.distributed as dist

    dist.broadcast(prior, src=src, group=group)
    dist.broadcast(position, src=src, group=group)
    return prior, position


def broadcast_prior(handle, src, slot=0, group=None):
    """NCCL broadcast of the rolling terrain prior of `slot` from rank `src` to every rank, in place
    on the handles' device memory; the map position travels with it."""
    import torch

    handle.synchronize()
    prior = prior_tensor(handle, slot)
    if torch.distributed.get_backend(group) == "nccl":
        assert prior.device.index == torch.cuda.current_device(), "the process group's device and the handle's device differ"
    pos = torch.tensor(handle.position(slot), dtype=torch.float64, device=prior.device)
    broadcast_prior_tensors(prior, pos, src, group)
    torch.cuda.synchronize()
    xy = pos.cpu().numpy()
    handle.set_position(float(xy[0]), float(xy[1]), slot)
    return xy
