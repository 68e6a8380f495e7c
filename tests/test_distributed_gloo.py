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


class _StubModel:
    """packed_layout / cfg of FlashVTGB200 without the CUDA library (host logic only)."""
    class cfg:  # noqa: N801
        max_num_moment = 50

    def packed_layout(self, B, Lv):
        from flashvtg_b200.model import FlashVTGB200
        return FlashVTGB200.packed_layout(self, B, Lv)


def _worker_plan(rank, world, port, n_total, mode, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from flashvtg_b200.distributed import PackedGather, gather_records, shard_batch, shard_plan
        g = torch.Generator().manual_seed(7)
        vlen = torch.randint(3, 76, (n_total,), generator=g, dtype=torch.int32)
        vlen[0] = 75
        full = {"vid_len": vlen, "txt_len": torch.randint(1, 33, (n_total,), generator=g, dtype=torch.int32),
                "src_vid": torch.randn(n_total, 75, 6, generator=g), "src_txt": torch.randn(n_total, 32, 5, generator=g)}
        plan = shard_plan(vlen, world, mode)
        local = shard_batch(full, rank, world, mode)
        ok = local["vid_len"].shape[0] == plan[rank].numel()
        if plan[rank].numel() and mode != "contiguous":
            ok = ok and local["src_vid"].shape[1] == int(local["vid_len"].max())   # cropped to the shard's longest
        # a per-video function of the shard rows (what the kernels compute), gathered back to input order
        res = {"score": local["vid_len"].float() * 2 + 1, "tag": local["txt_len"].to(torch.int32) + 5}
        got = gather_records(res, n_total, plan=plan)
        ok = ok and torch.equal(got["score"], vlen.float() * 2 + 1) and torch.equal(got["tag"], full["txt_len"] + 5)
        # per-clip field whose width is the shard's own longest video (saliency [n, Lv_rank]): ranks of a cropped plan
        # hold DIFFERENT widths - the records must be padded to a common one before the collective (an unequal
        # all_gather_into_tensor hangs NCCL and gloo alike; it did, on the 8-GPU bucketed bench), with and without
        # the caller's hint
        sal_full = full["src_vid"][:, :, 0] * (torch.arange(75)[None, :] < vlen[:, None])
        res2 = {"saliency": local["src_vid"][:, :, 0] * (torch.arange(local["src_vid"].shape[1])[None, :] <
                                                            local["vid_len"][:, None]),
                "count": local["vid_len"]}
        for hint in (None, {"saliency": 75}):
            got2 = gather_records(res2, n_total, plan=plan, pad_last=hint)
            w = got2["saliency"].shape[1]
            ok = ok and w <= 75 and torch.equal(got2["saliency"], sal_full[:, :w]) and \
                bool((sal_full[:, w:] == 0).all()) and torch.equal(got2["count"], vlen)
        # packed fast path: equal shards, one collective, views [world][B_local][...]
        m = _StubModel()
        B, Lv = 4, 9
        lay, n = m.packed_layout(B, Lv)
        packed = torch.zeros(n)
        o, cnt = lay["saliency"]
        packed[o:o + cnt] = torch.arange(cnt, dtype=torch.float32) + 100 * rank
        o, cnt = lay["count"]
        packed[o:o + cnt].view(torch.int32).copy_(torch.arange(cnt, dtype=torch.int32) + 7 * rank)
        pg = PackedGather(m, B, Lv, "cpu")
        slot = pg.gather(packed)
        pg.wait(slot)
        v = pg.views(slot)
        for r in range(world):
            ok = ok and torch.equal(v["saliency"][r].reshape(-1), torch.arange(B * Lv, dtype=torch.float32) + 100 * r)
            ok = ok and torch.equal(v["count"][r], torch.arange(B, dtype=torch.int32) + 7 * r)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["contiguous", "balanced", "bucketed"])
def test_shard_plans_and_single_collective_gather_world2(mode):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_plan, args=(r, 2, port, 11, mode, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(ok for _, ok in res)


def test_shard_plan_properties():
    from flashvtg_b200.distributed import shard_plan
    g = torch.Generator().manual_seed(3)
    for n in (1, 5, 64, 1000):
        ln = torch.randint(1, 200, (n,), generator=g)
        for world in (1, 2, 8):
            for mode in ("contiguous", "balanced", "bucketed"):
                plan = shard_plan(ln, world, mode)
                assert len(plan) == world
                assert torch.equal(torch.cat(plan).sort().values, torch.arange(n))   # a partition
            bal = shard_plan(ln, world, "balanced")
            sizes = [int(p.numel()) for p in bal]
            assert max(sizes) - min(sizes) <= 1
            if n >= 8 * world:
                clips = [int(ln[p].sum()) for p in bal]
                assert max(clips) - min(clips) <= int(ln.max())        # within one video of each other
                cost = [int(p.numel()) * int(ln[p].max()) for p in shard_plan(ln, world, "bucketed") if p.numel()]
                naive = -(-n // world) * int(ln.max())
                assert max(cost) <= naive                                  # never worse than padding to the global max


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
