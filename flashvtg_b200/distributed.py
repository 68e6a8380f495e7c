"""Multi-GPU plumbing of the hot path: videos shard by rank, results gather once.

The eval forward has no cross-video operation, so every rank runs the whole kernel
sequence on its own contiguous slice of the batch with replicated weights and NO data-path
collective; the only exchange is one all-gather of the ranked-span records at the end (NCCL over
NVLink on GPUs; the same code runs on gloo for the CPU tests of the host logic).
"""
from __future__ import annotations

from typing import Optional, Dict, Tuple

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


def shard_plan(vid_len, world: int, mode: str = "contiguous"):
    """Video indices each rank processes -> list of `world` int64 index tensors (a partition of range(n)).

    contiguous : rank r gets shard_range(n, r, world) in input order (what a DistributedSampler-less split does).
    balanced   : SURVEY section 8(e): sort by true clip count, then deal the sorted list out in snake order
                 (r, 2w-1-r, ...), so every rank receives the same number of videos (+-1) AND the same length
                 distribution - per-rank FLOPs (proportional to clips processed) differ by a fraction of one video.
    bucketed   : sort by clip count, cut the sorted list into `world` contiguous runs whose padded cost
                 (videos x longest video of the run) is equal: each rank pads to ITS longest video only, which
                 removes most of the padding work of a ragged batch at the price of unequal shard sizes.
    Lengths are host metadata (the dataset index knows every video's duration), so every rank computes the same
    plan without communicating."""
    ln = torch.as_tensor(vid_len).to(torch.int64).cpu()
    n = int(ln.numel())
    if mode == "contiguous":
        return [torch.arange(*shard_range(n, r, world)) for r in range(world)]
    order = torch.argsort(ln, descending=True, stable=True)
    if mode == "balanced":
        parts = [[] for _ in range(world)]
        for i, idx in enumerate(order.tolist()):
            rnd, pos = divmod(i, world)
            parts[pos if rnd % 2 == 0 else world - 1 - pos].append(idx)
        return [torch.tensor(p, dtype=torch.int64) for p in parts]
    if mode == "bucketed":
        sl = ln[order].tolist()
        # equalise count x max over `world` contiguous runs of the descending list: binary search on the cost cap
        lo, hi = 1, max(1, n * (sl[0] if sl else 1))

        def cuts(cap):
            out, start = [], 0
            while start < n and len(out) < world:
                cnt = max(1, min(n - start, cap // max(sl[start], 1)))
                out.append((start, start + cnt))
                start += cnt
            return out, start
        while lo < hi:
            mid = (lo + hi) // 2
            _, covered = cuts(mid)
            if covered >= n:
                hi = mid
            else:
                lo = mid + 1
        runs, _ = cuts(lo)
        runs += [(n, n)] * (world - len(runs))
        return [order[a:b] for a, b in runs]
    raise ValueError(f"unknown shard mode {mode!r}")


def shard_batch(batch: Dict[str, torch.Tensor], rank: int, world: int, mode: str = "contiguous",
                crop: bool = True) -> Dict[str, torch.Tensor]:
    """This rank's rows of a padded batch.  Non-contiguous plans also crop the clip / token axes to the
    shard's own longest video / query (`crop`), so a rank never projects padding it does not need."""
    n = next(iter(batch.values())).shape[0]
    if mode == "contiguous":
        s, e = shard_range(n, rank, world)
        return {k: v[s:e] for k, v in batch.items()}
    idx = shard_plan(batch["vid_len"], world, mode)[rank]
    out = {k: v[idx] for k, v in batch.items()}
    if crop and idx.numel() > 0:
        lv = int(out["vid_len"].max())
        lt = int(out["txt_len"].max()) if "txt_len" in out else None
        for k in ("src_vid", "src_vid_mask"):
            if k in out:
                out[k] = out[k][:, :lv].contiguous()
        for k in ("src_txt", "src_txt_mask"):
            if k in out and lt is not None:
                out[k] = out[k][:, :lt].contiguous()
    return out


def gather_records(local: Dict[str, torch.Tensor], n_total: int, group=None, plan=None,
                   pad_last: Optional[Dict[str, int]] = None) -> Dict[str, torch.Tensor]:
    """All-gather per-video result tensors (leading dim = local shard) into global order with ONE collective.

    Every field's rows are viewed as bytes and laid side by side in one [max_shard][row_bytes] record buffer
    (shards padded to the largest one), a single all_gather_into_tensor moves it, and the fields are cut back
    out.  `plan` (shard_plan(...)) restores the input order of a non-contiguous split; default: contiguous.
    Returns tensors with leading dim n_total on every rank.

    Ranks that cropped their shard to their own longest video (shard_batch(crop=True), the bucketed / balanced
    plans) hold per-clip fields of different widths (saliency [n, Lv_rank]); an all-gather needs equal records, so
    the last dim of every field is zero-padded to a common size first: `pad_last[name]` when the caller knows it
    (no extra traffic), otherwise the maximum over the ranks (one small all_reduce)."""
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = [shard_range(n_total, r, world) for r in range(world)] if plan is None else None
    counts = [e - s for s, e in sizes] if plan is None else [int(p.numel()) for p in plan]
    max_shard = max(max(counts), 1)
    names = list(local)
    dev = local[names[0]].device
    last = torch.tensor([(local[k].shape[-1] if local[k].dim() > 1 else 1) for k in names], dtype=torch.int64)
    if pad_last is not None and all(k in pad_last for k in names if local[k].dim() > 1):
        want = torch.tensor([(pad_last[k] if local[k].dim() > 1 else 1) for k in names], dtype=torch.int64)
        assert bool((want >= last).all()), "pad_last smaller than a local field"
    else:
        want = last.to(dev)
        dist.all_reduce(want, op=dist.ReduceOp.MAX, group=group)
        want = want.cpu()
    padded = {}
    for i, k in enumerate(names):
        t = local[k]
        if t.dim() > 1 and int(want[i]) > t.shape[-1]:
            t = torch.nn.functional.pad(t, (0, int(want[i]) - t.shape[-1]))
        padded[k] = t
    local = padded
    cols, metas = [], []
    for k in names:
        t = local[k].contiguous()
        assert t.shape[0] == counts[rank], f"{k}: local shard has {t.shape[0]} rows, expected {counts[rank]}"
        row = t.reshape(t.shape[0], -1).view(torch.uint8) if t.shape[0] else \
            torch.empty(0, max(1, int(torch.tensor(t.shape[1:]).prod())) * t.element_size(), dtype=torch.uint8, device=dev)
        cols.append(row)
        metas.append((k, t.dtype, tuple(t.shape[1:]), row.shape[1]))
    rec = torch.zeros(max_shard, sum(m[3] for m in metas), dtype=torch.uint8, device=dev)
    if counts[rank]:
        rec[:counts[rank]] = torch.cat(cols, 1)
    buf = torch.empty(world * max_shard, rec.shape[1], dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(buf, rec, group=group)
    buf = buf.view(world, max_shard, -1)
    valid = torch.cat([buf[r, :counts[r]] for r in range(world)], 0)   # rank-major order
    if plan is not None:
        inv = torch.empty(n_total, dtype=torch.int64)
        inv[torch.cat([p.to(torch.int64) for p in plan])] = torch.arange(n_total)
        valid = valid[inv.to(dev)]
    out, off = {}, 0
    for k, dt, shape, nb in metas:
        out[k] = valid[:, off:off + nb].contiguous().view(dt).view((n_total,) + shape)
        off += nb
    return out


class PackedGather:
    """The bench / serving fast path for EQUAL shards: every rank's records already sit in one contiguous
    buffer (FvtgResult.packed, written directly by the kernels), so the exchange is a single
    all_gather_into_tensor into a preallocated [world][n_packed] buffer on a side stream - no packing kernels, no
    reassembly copies, and the collective of step k overlaps the kernels of step k+1.  Fields come back as views
    shaped [world][B_local][...] (global video index = rank * B_local + local index)."""

    def __init__(self, model, B_local: int, Lv: int, device, group=None, depth: int = 2):
        self.group = group
        self.world = dist.get_world_size(group)
        self.lay, self.n = model.packed_layout(B_local, Lv)
        self.B, self.Lv, self.topk = B_local, Lv, model.cfg.max_num_moment
        self.bufs = [torch.empty(self.world, self.n, dtype=torch.float32, device=device) for _ in range(depth)]
        self.is_cuda = torch.device(device).type == "cuda"
        self.stream = torch.cuda.Stream(device=device) if self.is_cuda else None
        self.events = [None] * depth
        self.i = 0

    def gather(self, packed: torch.Tensor):
        """Enqueue the collective for `packed` (a FvtgResult.packed of the current stream); returns the slot
        index to pass to wait() / views()."""
        slot = self.i % len(self.bufs)
        self.i += 1
        if self.is_cuda:
            cur = torch.cuda.current_stream()
            self.stream.wait_stream(cur)           # the records are complete
            with torch.cuda.stream(self.stream):
                dist.all_gather_into_tensor(self.bufs[slot].view(-1), packed, group=self.group)
                ev = torch.cuda.Event()
                ev.record(self.stream)
            packed.record_stream(self.stream)
            self.events[slot] = ev
        else:
            dist.all_gather_into_tensor(self.bufs[slot].view(-1), packed, group=self.group)
        return slot

    def wait(self, slot: int):
        if self.is_cuda and self.events[slot] is not None:
            torch.cuda.current_stream().wait_event(self.events[slot])

    def views(self, slot: int):
        buf = self.bufs[slot]

        def f(name, *shape, as_int=False):
            o, n = self.lay[name]
            t = buf[:, o:o + n]
            return (t.view(torch.int32) if as_int else t).view(self.world, *shape)
        return {"nms_windows": f("nms_windows", self.B, self.topk, 3), "saliency": f("saliency", self.B, self.Lv),
                "count": f("count", self.B, as_int=True), "nms_count": f("nms_count", self.B, as_int=True)}
