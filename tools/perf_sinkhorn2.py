"""Sinkhorn timing at the headline shape for kernel-variant experiments (env: DRG_SKH_STAGE_FLOATS, DRG_SKH_L2_KEEP)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffreg_b200 import ops, _lib
dev = "cuda"
shapes = [(1, 4096, 4096, 3)]
if len(sys.argv) > 1 and sys.argv[1] == "all":
    shapes = [(1, 1024, 1024, 3), (1, 2048, 2048, 3), (1, 4096, 4096, 3), (1, 4800, 2048, 3), (1, 8192, 8192, 3), (1, 16384, 16384, 3), (16, 2048, 2048, 3), (1, 4800, 1532, 3)]
for (B, N, M, I) in shapes:
    s = torch.randn(B, N, M, device=dev)
    sm = torch.ones(B, N, dtype=torch.bool, device=dev); tm = torch.ones(B, M, dtype=torch.bool, device=dev)
    alpha = torch.tensor(1.0, device=dev)
    for _ in range(3):
        ops.sinkhorn(s, alpha, I, sm, tm, out_mode="conf")
    torch.cuda.synchronize()
    _lib.profile_enable(True)
    reps = 20
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        ops.sinkhorn(s, alpha, I, sm, tm, out_mode="conf")
    e1.record(); torch.cuda.synchronize()
    prof = _lib.profile_read(); _lib.profile_enable(False)
    E = 4.0 * B * (N + 1) * (M + 1)
    us = e0.elapsed_time(e1) * 1e3 / reps
    per = lambda k: round(prof[k][0] * 1e3 / prof[k][1], 2) if prof[k][1] else None
    print(json.dumps(dict(B=B, N=N, M=M, us_call=round(us, 1), algo_GBps=round((2 * I + 2) * E / us / 1e3), iter_us=per("skh_iter"),
                          col_us=per("skh_col"), final_us=per("skh_final"), prep_us=per("skh_prep"))), flush=True)
    del s
