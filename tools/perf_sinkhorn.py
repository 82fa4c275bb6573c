"""Scratch timing of the Sinkhorn kernels (not the bench): CUDA events, L2 flushed between runs."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffreg_b200
from diffreg_b200 import ops

dev = "cuda"
flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)
def timeit(fn, iters=10, flush_l2=True):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        if flush_l2: flush.zero_()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts)//2], ts[0]

for (B, N, M, I) in [(1,1024,1024,3),(1,2048,2048,3),(1,4096,4096,3),(1,4800,2048,3),(1,4096,4096,1),(1,16384,16384,3),(1,16384,16384,10),(16,2048,2048,3)]:
    s = torch.randn(B, N, M, device=dev)
    sm = torch.ones(B, N, dtype=torch.bool, device=dev); tm = torch.ones(B, M, dtype=torch.bool, device=dev)
    alpha = torch.tensor(1.0, device=dev)
    E = 4.0 * B * (N + 1) * (M + 1)
    for mode in ("conf", "none"):
        med, best = timeit(lambda: ops.sinkhorn(s, alpha, I, sm, tm, out_mode=mode))
        passes = 2 * I + (2 if mode == "conf" else 0)
        print(json.dumps(dict(B=B, N=N, M=M, iters=I, mode=mode, us_median=round(med,1), us_best=round(best,1),
                              algo_GBps=round(passes * E / med / 1e3, 1))), flush=True)
    if M <= 4096:
        med, best = timeit(lambda: ops.dual_softmax(s, sm, tm, 0.1))
        print(json.dumps(dict(B=B, N=N, M=M, op="dual_softmax", us_median=round(med,1), algo_GBps=round(4 * E / med / 1e3, 1))), flush=True)
    del s
