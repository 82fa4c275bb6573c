"""Fixed cost per launch of the wide-row Sinkhorn iteration (tuning tool): N x 16384 matrices with few rows per CTA."""
import os, sys, json
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from diffreg_b200 import ops

M = 16384
alpha = torch.tensor(1.0, device="cuda")
for N in (1024, 2048, 4096, 8192):
    s = torch.randn(1, N, M, device="cuda")
    sm = torch.ones(1, N, dtype=torch.bool, device="cuda"); tm = torch.ones(1, M, dtype=torch.bool, device="cuda")
    fn = lambda: ops.sinkhorn(s, alpha, 50, sm, tm, out_mode="none")
    fn(); torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    tot, cnt = defaultdict(float), defaultdict(int)
    for ev in prof.events():
        if ev.device_type == torch.autograd.DeviceType.CUDA:
            nm = ev.name.split("(")[0].split("<")[0][-28:]
            tot[nm] += ev.device_time; cnt[nm] += 1
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); fn(); e1.record(); torch.cuda.synchronize()
    print(json.dumps({"N": N, "rows_per_cta": round(N / 148, 1), "us_per_iteration": round(e0.elapsed_time(e1) * 1e3 / 50, 1),
                      "kernels_us_each": {k: round(tot[k] / cnt[k], 1) for k in tot if cnt[k] >= 50}}), flush=True)
