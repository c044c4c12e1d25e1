"""Multi-GPU plumbing: environment batches shard independently (one process per GPU, no data-path collective in the
physics); torch.distributed is used for exactly two things:
  1. a one-time broadcast of the compiled model constants from rank 0,
  2. an optional per-step all-gather of the observation rows.
Works with backend "nccl" on GPUs and "gloo" on CPU (tests)."""
import io

import numpy as np


def shard_range(n_total, rank, world):
    """contiguous env index range [lo, hi) owned by `rank` (remainder spread over the first ranks)"""
    base, rem = divmod(n_total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def model_to_bytes(model):
    from .mjcf.compiler import save_model

    buf = io.BytesIO()
    save_model(model, buf)
    return buf.getvalue()


def model_from_bytes(data):
    from .mjcf.compiler import load_model

    return load_model(io.BytesIO(data))


def broadcast_model(model, src=0, device=None):
    """rank `src` passes a compiled Model; every rank returns an identical Model"""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return model
    rank = dist.get_rank()
    dev = device if device is not None else torch.device("cpu")
    if rank == src:
        payload = np.frombuffer(model_to_bytes(model), dtype=np.uint8).copy()
        n = torch.tensor([len(payload)], dtype=torch.int64, device=dev)
    else:
        n = torch.zeros(1, dtype=torch.int64, device=dev)
    dist.broadcast(n, src)
    buf = torch.from_numpy(payload).to(dev) if rank == src else torch.empty(int(n.item()), dtype=torch.uint8, device=dev)
    dist.broadcast(buf, src)
    return model if rank == src else model_from_bytes(buf.cpu().numpy().tobytes())


def allgather_obs(local_obs, out=None):
    """[n_local, d] on every rank -> [world * n_local, d] on every rank (rank-major order)"""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local_obs
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * local_obs.shape[0], local_obs.shape[1]), dtype=local_obs.dtype, device=local_obs.device)
    dist.all_gather_into_tensor(out, local_obs.contiguous())
    return out
