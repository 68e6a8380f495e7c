"""CPU, world_size 2 over gloo: the N > 1 host logic (video sharding + the single all-gather of
ranked-span records).  The kernels themselves are exercised by the -m gpu tests; the collective and
the shard arithmetic are what differs at N > 1."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flashvtg_b200.distributed import gather_records, shard_batch, shard_range
        full = {"windows": torch.arange(n_total * 50 * 3, dtype=torch.float32).view(n_total, 50, 3),
                "count": torch.arange(n_total, dtype=torch.int32) % 51,
                "saliency": torch.arange(n_total * 75, dtype=torch.float32).view(n_total, 75) * 0.5}
        local = shard_batch(full, rank, world)
        s, e = shard_range(n_total, rank, world)
        assert local["count"].shape[0] == e - s
        # stand-in for the per-rank kernel sequence: a rank-independent function of the shard
        local = {k: v + 1 for k, v in local.items()}
        got = gather_records(local, n_total)
        ok = all(torch.equal(got[k], full[k] + 1) for k in full)
        q.put((rank, ok, (s, e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [8, 7, 1])
def test_shard_and_gather_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok, _ in res)
    spans = sorted(se for _, _, se in res)
    assert spans[0][0] == 0 and spans[-1][1] == n_total and spans[0][1] == spans[1][0]


def test_shard_range_properties():
    from flashvtg_b200.distributed import shard_range
    for n in (0, 1, 5, 1024, 1027):
        for world in (1, 2, 4, 8):
            parts = [shard_range(n, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
            sizes = [e - s for s, e in parts]
            assert max(sizes) - min(sizes) <= 1
