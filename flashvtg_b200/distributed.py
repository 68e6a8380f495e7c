"""Multi-GPU plumbing of the hot path: videos shard by rank, results gather once.

The eval forward has no cross-video operation, so every rank runs the whole kernel
sequence on its own contiguous slice of the batch with replicated weights and NO data-path
collective; the only exchange is one all-gather of the ranked-span records at the end (NCCL over
NVLink on GPUs; the same code runs on gloo for the CPU tests of the host logic).
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.distributed as dist


def bind_to_gpu_numa_node(device_index: int) -> bool:
    """Pin the calling process to the CPUs NVML reports as local to GPU `device_index`, so that the pinned
    host buffers it allocates afterwards (first touch) sit on the socket that GPU's PCIe link hangs off.
    With one rank per GPU and 8 ranks streaming features over PCIe at once, buffers on the wrong socket
    share the inter-socket link.  Returns False (and changes nothing) when NVML is unavailable."""
    try:
        import pynvml
        pynvml.nvmlInit()
        # NVML enumerates by PCI bus id; CUDA_VISIBLE_DEVICES remapping is resolved through the UUID
        props = torch.cuda.get_device_properties(device_index)
        uuid = getattr(props, "uuid", None)
        handle = None
        if uuid is not None:
            try:
                handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + str(uuid)).encode())
            except Exception:  # noqa: BLE001
                handle = None
        if handle is None:
            handle = pynvml.nvmlDeviceGetHandleByIndex(device_index)
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:  # noqa: BLE001 - best effort: affinity is an optimisation, never a requirement
        return False


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) slice of n items owned by `rank`: the first n % world ranks get one
    extra item, so shard sizes differ by at most one."""
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int) -> Dict[str, torch.Tensor]:
    n = next(iter(batch.values())).shape[0]
    s, e = shard_range(n, rank, world)
    return {k: v[s:e] for k, v in batch.items()}


def gather_records(local: Dict[str, torch.Tensor], n_total: int, group=None) -> Dict[str, torch.Tensor]:
    """All-gather per-video result tensors (leading dim = local shard) into global order.

    Shards are padded to the largest shard so ONE all_gather_into_tensor per field suffices; the
    padding rows are dropped afterwards.  Returns tensors with leading dim n_total on every rank."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    max_shard = -(-n_total // world)
    out = {}
    for k, t in local.items():
        s, e = shard_range(n_total, rank, world)
        assert t.shape[0] == e - s, f"{k}: local shard has {t.shape[0]} rows, expected {e - s}"
        if t.shape[0] < max_shard:
            pad = torch.zeros((max_shard - t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            t = torch.cat([t, pad], 0)
        buf = torch.empty((world * max_shard,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t.contiguous(), group=group)
        parts = []
        for r in range(world):
            rs, re_ = shard_range(n_total, r, world)
            parts.append(buf[r * max_shard: r * max_shard + (re_ - rs)])
        out[k] = torch.cat(parts, 0)
    return out
