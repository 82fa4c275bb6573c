"""Per-kernel device times of the bench step from the CUPTI activity records (torch.profiler), eager launches of the
20 sampler steps after warm-up: the in-stream durations, kernels back to back as in the timed region (tuning tool;
the judged numbers are bench.py's CUDA-event times and the ncu launch list under profiles/)."""
import os
import sys
from collections import defaultdict

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
from torch.profiler import profile, ProfilerActivity
import diffreg_b200
import bench

dev = torch.device("cuda", 0)
n, c = 4096, 256
host = bench.make_inputs(3000, n, c)
d = {k: v.to(dev) for k, v in host.items()}
head = diffreg_b200.Matching(bench.MATCH_CFG).to(dev).eval()
with torch.no_grad():
    head.src_proj.weight.copy_(d["W"])
proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
smp = diffreg_b200.DenoisingSampler("4d", head, proc, bench.SAMPLER_STEPS, noise_seed=1234)
feats = [d[k] for k in ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")]
bufs = [d["x_T"].clone(), torch.empty_like(d["x_T"])]
counter = torch.zeros(1, dtype=torch.int64, device=dev)
step = lambda i: smp.step(i % 20, bufs[i % 2], None, *feats, x_out=bufs[(i + 1) % 2], noise_counter=counter)
for i in range(5):
    step(i)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for i in range(20):
        step(5 + i)
    torch.cuda.synchronize()
tot, cnt = defaultdict(float), defaultdict(int)
order = []
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0][:70]
        if name not in tot:
            order.append(name)
        tot[name] += ev.device_time if hasattr(ev, "device_time") else ev.cuda_time
        cnt[name] += 1
total = 0.0
for name in order:
    per_step = tot[name] / 20
    total += per_step
    print(f"{per_step:8.1f} us/step  {cnt[name] / 20:4.1f} launches/step  {tot[name] / cnt[name]:7.1f} us each  {name}")
print(f"{total:8.1f} us/step in kernels + memsets")
