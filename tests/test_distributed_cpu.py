"""Host-side logic of the multi-GPU paths on CPU: the unit / row partitions and the log-sum-exp all-reduce, with two
gloo ranks.  (The kernels themselves need a GPU: tests/test_rowshard_gpu.py.)"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import diffreg_b200
from diffreg_b200 import distributed as D


def test_shard_rows_partition():
    for n in (1, 7, 16, 16384, 4801):
        for world in (1, 2, 3, 4, 8):
            blocks = [D.shard_rows(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[r][1] == blocks[r + 1][0] for r in range(world - 1))
            sizes = [b - a for a, b in blocks]
            assert max(sizes) - min(sizes) <= 1


def test_shard_units_partition():
    for n in (0, 1, 8, 13):
        for world in (1, 2, 4, 8):
            got = sorted(sum((D.shard_units(n, world, r) for r in range(world)), []))
            assert got == list(range(n))


def test_lse_combine_matches_direct():
    g = torch.Generator().manual_seed(0)
    x = torch.randn(5, 300, generator=g, dtype=torch.float64) * 20           # 5 shards of log2-domain values
    parts = []
    for r in range(5):
        m = x[r].max()
        parts.append(torch.stack((m, torch.exp2(x[r] - m).sum()))[None])
    red = D.lse_combine(parts)
    got = red[..., 0] + torch.log2(red[..., 1])
    want = torch.log2(torch.exp2(x.reshape(-1) - x.max()).sum()) + x.max()
    assert abs(got.item() - want.item()) < 1e-9


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(1)
        full = torch.randn(2, 64, 33, generator=g) * 5          # [B, N, M+1] log2-domain entries; rows are sharded
        a, b = D.shard_rows(64, world, rank)
        loc = full[:, a:b]
        m = loc.max(dim=1)[0]
        partial = torch.stack((m, torch.exp2(loc - m[:, None]).sum(dim=1)), dim=-1)
        exact = D.lse_allreduce(partial.clone())
        lse = exact[..., 0] + torch.log2(exact[..., 1])
        want = torch.log2(torch.exp2(full.double() - full.max()).sum(dim=1)) + full.max()
        err1 = (lse.double() - want).abs().max().item()
        # single all-reduce with a reference that is off by a few units
        one = D.lse_allreduce(partial.clone(), ref=lse + 3.0)
        lse2 = one[..., 0] + torch.log2(one[..., 1])
        err2 = (lse2.double() - want).abs().max().item()
        q.put((rank, err1, err2))
    finally:
        dist.destroy_process_group()


def test_lse_allreduce_two_gloo_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for _, err1, err2 in res:
        assert err1 < 1e-5 and err2 < 1e-5


def _p2p_worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # No GPU here: drg_p2p_create cannot allocate.  The point of the test is the protocol -- every rank must learn
        # that the exchange is unavailable and raise TOGETHER (a rank that went on alone would hang its peers).
        try:
            D.P2PComm(128, 8, device="cpu")
            q.put((rank, "created"))
        except D.P2PUnavailable as e:
            q.put((rank, "unavailable:" + str(e)[:40]))
        # the sharded driver then falls back to the NCCL / gloo all-reduce path instead of raising ...
        op = D.RowShardedSinkhorn(exchange=None)
        op._p2p_failed = False
        q.put((rank, "fallback" if (dist.get_backend() != "nccl" and op._p2p(1, 31) is None) else "p2p"))
    finally:
        dist.destroy_process_group()


def test_p2p_comm_fails_on_all_ranks_together_without_gpu():
    if torch.cuda.is_available():
        pytest.skip("needs a box without a GPU (the working exchange is covered by tools/bench_rowshard.py on 2 GPUs)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_p2p_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(4)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    states = sorted(res)
    assert all(s.startswith("unavailable") or s == "fallback" for _, s in states), states
    assert sum(s == "fallback" for _, s in states) == 2
