"""BASELINE.json configs[4]: N = M = 16384 log-domain Sinkhorn, 100 iterations, rows sharded over the ranks, the
per-column log-sum-exp partials all-reduced (NCCL) every iteration.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node P --master-addr 127.0.0.1 --master-port 29511 \
        tools/bench_rowshard.py [--size 16384] [--iters 100] [--reps 3]
    python tools/bench_rowshard.py            # P = 1 (the unsharded kernels through the same driver)
Prints one JSON line on rank 0: ms per Sinkhorn call (max over ranks, CUDA events), achieved algorithmic GB/s
((2I+2) * 4 (N+1)(M+1) bytes per call, whole job) and a checksum-level parity check of the result against the
unsharded kernel on rank 0 when the matrix fits (n <= 8192)."""
import argparse, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import diffreg_b200
from diffreg_b200 import distributed as D

ap = argparse.ArgumentParser()
ap.add_argument("--size", type=int, default=16384)
ap.add_argument("--iters", type=int, default=100)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"])
ap.add_argument("--profile", action="store_true", help="rank 0 also prints the average device time of every kernel of one call (CUPTI)")
args = ap.parse_args()
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
os.environ.setdefault("MASTER_ADDR", "127.0.0.1"); os.environ.setdefault("MASTER_PORT", "29511")
dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
N = M = args.size
a, b = D.shard_rows(N, world, rank)
g = torch.Generator(device=dev).manual_seed(5000)            # same stream on every rank: generate the full rows lazily
# scores ~ N(0,1): rank r draws only its block, seeded per row block so that any world size sees the same matrix
scores = torch.empty(1, b - a, M, device=dev)
blk = 1024
for r0 in range(a, b, blk):
    r1 = min(b, r0 + blk)
    gg = torch.Generator(device=dev).manual_seed(5000 + r0 // blk)
    full_blk = torch.randn(blk, M, generator=gg, device=dev)
    scores[0, r0 - a:r1 - a] = full_blk[r0 % blk: r0 % blk + (r1 - r0)] if (r0 % blk) else full_blk[: r1 - r0]
src_mask = torch.ones(1, b - a, dtype=torch.bool, device=dev)
tgt_mask = torch.ones(1, M, dtype=torch.bool, device=dev)
alpha = torch.tensor(1.0, device=dev)
op = D.RowShardedSinkhorn(exchange=args.exchange)
out = op(scores, alpha, args.iters, src_mask, tgt_mask, out_mode="conf")
torch.cuda.synchronize(); dist.barrier()
times = []
for _ in range(args.reps):
    dist.barrier(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    out = op(scores, alpha, args.iters, src_mask, tgt_mask, out_mode="conf")
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    times.append(float(t.item()))
ms = sorted(times)[len(times) // 2]
if args.profile:
    from collections import defaultdict
    from torch.profiler import profile, ProfilerActivity
    dist.barrier(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        op(scores, alpha, args.iters, src_mask, tgt_mask, out_mode="conf")
        torch.cuda.synchronize()
    if rank == 0:
        tot, cnt = defaultdict(float), defaultdict(int)
        for ev in prof.events():
            if ev.device_type == torch.autograd.DeviceType.CUDA:
                nm = ev.name.split("(")[0][:60]
                tot[nm] += ev.device_time; cnt[nm] += 1
        for nm in sorted(tot, key=lambda k: -tot[k]):
            print(f"[profile] {tot[nm]:10.1f} us  {cnt[nm]:5d} launches  {tot[nm] / cnt[nm]:8.1f} us each  {nm}", file=sys.stderr, flush=True)
# the two exchange paths must agree (same arithmetic up to the order of the P-way combine)
other = D.RowShardedSinkhorn(exchange="nccl" if args.exchange == "p2p" else "p2p")
ref = other(scores, alpha, min(args.iters, 10), src_mask, tgt_mask, out_mode="conf")
mine = op(scores, alpha, min(args.iters, 10), src_mask, tgt_mask, out_mode="conf")
diff = (ref - mine).abs().max()
dist.all_reduce(diff, op=dist.ReduceOp.MAX)
if other.comm is not None:
    other.comm.close()
# size-independent property: every real column of the full plan sums to 1 minus its dustbin-row entry; check the global column sums
col = out.double().sum(dim=1)
dist.all_reduce(col, op=dist.ReduceOp.SUM)
E = 4.0 * (N + 1) * (M + 1)
if rank == 0:
    print(json.dumps({"workload": f"row-sharded log-Sinkhorn N=M={N}, iters={args.iters}", "n_gpus": world, "ms_per_call": ms,
                      "algorithmic_GBps_whole_job": (2 * args.iters + 2) * E / (ms * 1e-3) / 1e9,
                      "ms_per_iteration": ms / args.iters, "col_sum_min": float(col.min()), "col_sum_max": float(col.max()),
                      "exchange": args.exchange, "max_abs_diff_vs_other_exchange_10_iters": float(diff), "p2p_status": (op.comm.status() if op.comm is not None else None)}), flush=True)
if op.comm is not None:
    op.comm.close()
dist.destroy_process_group()
