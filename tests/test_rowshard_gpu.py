"""Row-sharded Sinkhorn kernels (drg_sinkhorn_shard_*) on one GPU: P emulated shards must reproduce the unsharded
result and the oracle (1e-4 abs, fp32)."""
import pytest
import torch

from oracle import diffreg_oracle as O
from helpers import TOL_LOG, finite_close

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("B,N,M,P,kind", [(1, 64, 48, 2, "full"), (2, 130, 260, 3, "prefix"), (1, 1000, 1024, 4, "arbitrary"),
                                          (1, 257, 4096, 8, "full"), (1, 96, 8192, 2, "full"), (1, 33, 1531, 2, "arbitrary")])
def test_emulated_shards_match_oracle(B, N, M, P, kind):
    import diffreg_b200
    gen = torch.Generator().manual_seed(N * 7 + M + P)
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
    filled = s.masked_fill(~O.pair_mask(sm, tm), float("-inf"))
    ref = O.log_optimal_transport(filled, alpha, 4, sm, tm)
    emu = diffreg_b200.EmulatedRowShards(P)
    out = emu(s.cuda(), alpha.cuda(), 4, sm.cuda(), tm.cuda(), out_mode="conf", apply_mask=True)
    assert out.shape == (B, N, M)
    assert (out.cpu() - ref.exp()[:, :-1, :-1]).abs().max() <= TOL_LOG
    whole = diffreg_b200.ops.sinkhorn(s.cuda(), alpha.cuda(), 4, sm.cuda(), tm.cuda(), out_mode="conf", apply_mask=True)
    assert (out - whole).abs().max() <= 2e-6


def test_single_rank_process_group_matches_unsharded():
    """RowShardedSinkhorn through torch.distributed (NCCL, world size 1 on this box)."""
    import os
    import torch.distributed as dist
    import diffreg_b200
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29533")
    created = False
    if not dist.is_initialized():
        dist.init_process_group("nccl", rank=0, world_size=1)
        created = True
    try:
        gen = torch.Generator().manual_seed(3)
        s = torch.randn(1, 300, 2048, generator=gen).cuda()
        ones_s = torch.ones(1, 300, dtype=torch.bool).cuda()
        ones_t = torch.ones(1, 2048, dtype=torch.bool).cuda()
        alpha = torch.tensor(1.0).cuda()
        out = diffreg_b200.RowShardedSinkhorn()(s, alpha, 5, ones_s, ones_t, out_mode="conf")
        whole = diffreg_b200.ops.sinkhorn(s, alpha, 5, ones_s, ones_t, out_mode="conf")
        assert (out - whole).abs().max() <= 2e-6
    finally:
        if created:
            dist.destroy_process_group()
