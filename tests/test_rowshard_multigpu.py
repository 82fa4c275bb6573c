"""BASELINE.json configs[4] on REAL ranks: the row-sharded Sinkhorn with two NCCL ranks on two GPUs, both exchange paths
(in-kernel peer-to-peer over NVLink, and NCCL all-reduces), against the oracle and against the unsharded kernel.
Needs two GPUs: skipped on the single-GPU box (there: tests/test_rowshard_gpu.py, emulated shards); run with
    gpurun --gpus 2 -- python -m pytest tests/test_rowshard_multigpu.py -q        (log kept under profiles/)."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import diffreg_b200
    from diffreg_b200 import distributed as D
    from oracle import diffreg_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        res = []
        for (B, N, M, iters, kind) in [(1, 1000, 1536, 20, "arbitrary"), (2, 257, 4096, 5, "prefix"), (1, 2048, 8192, 100, "full")]:
            gen = torch.Generator().manual_seed(N + M)
            s = torch.randn(B, N, M, generator=gen) * 2.0
            sm = torch.ones(B, N, dtype=torch.bool)
            tm = torch.ones(B, M, dtype=torch.bool)
            if kind == "prefix":
                sm[:, N - 9:] = False
                tm[:, M - 17:] = False
            elif kind == "arbitrary":
                sm = torch.rand(B, N, generator=gen) > 0.1
                tm = torch.rand(B, M, generator=gen) > 0.1
            alpha = torch.tensor(1.0)
            a, b = D.shard_rows(N, world, rank)
            loc = [s[:, a:b].contiguous().to(dev), alpha.to(dev), iters, sm[:, a:b].contiguous().to(dev), tm.to(dev)]
            outs = {}
            for exchange in ("p2p", "nccl"):
                op = D.RowShardedSinkhorn(exchange=exchange)
                outs[exchange] = op(*loc, out_mode="conf", apply_mask=True)
                if op.comm is not None:
                    assert op.comm.status() == 0
                    op.comm.close()
            whole = diffreg_b200.ops.sinkhorn(s.to(dev), alpha.to(dev), iters, sm.to(dev), tm.to(dev), out_mode="conf", apply_mask=True)
            err_paths = (outs["p2p"] - outs["nccl"]).abs().max().item()
            err_whole = (outs["p2p"] - whole[:, a:b]).abs().max().item()
            err_oracle = None
            if N * M <= 2_000_000:      # the CPU oracle in seconds
                filled = s.masked_fill(~O.pair_mask(sm, tm), float("-inf"))
                ref = O.log_optimal_transport(filled, alpha, iters, sm, tm).exp()[:, :-1, :-1]
                err_oracle = (outs["p2p"].cpu() - ref[:, a:b]).abs().max().item()
            res.append((B, N, M, iters, err_paths, err_whole, err_oracle))
        q.put((rank, res))
    finally:
        dist.destroy_process_group()


def test_two_ranks_match_oracle_and_unsharded():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = [q.get(timeout=300) for _ in procs]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    for rank, res in got:
        for (B, N, M, iters, err_paths, err_whole, err_oracle) in res:
            print(f"rank {rank}: B={B} N={N} M={M} iters={iters}: p2p vs nccl {err_paths:.2e}, vs unsharded {err_whole:.2e}, vs oracle {err_oracle}")
            assert err_paths <= 1e-6
            assert err_whole <= 5e-6
            assert err_oracle is None or err_oracle <= 1e-4
