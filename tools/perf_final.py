"""Final-pass (DDIM + noise + arg-max) timing at 4096^2 inside the eager step (tuning: DRG_FT_ROWS)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
import diffreg_b200
from diffreg_b200 import _lib
import bench
dev = torch.device("cuda", 0)
n, c = 4096, 256
host = bench.make_inputs(3000, n, c)
d = {k: v.to(dev) for k, v in host.items()}
head = diffreg_b200.Matching(bench.MATCH_CFG).to(dev).eval()
with torch.no_grad():
    head.src_proj.weight.copy_(d["W"])
proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
smp = diffreg_b200.DenoisingSampler("4d", head, proc, 20, noise_seed=1234)
feats = [d[k] for k in ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")]
bufs = [d["x_T"].clone(), torch.empty_like(d["x_T"])]
counter = torch.zeros(1, dtype=torch.int64, device=dev)
for i in range(5):
    smp.step(i % 20, bufs[i % 2], None, *feats, x_out=bufs[(i + 1) % 2], noise_counter=counter)
torch.cuda.synchronize()
_lib.profile_enable(True)
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(40):
    smp.step(i % 20, bufs[i % 2], None, *feats, x_out=bufs[(i + 1) % 2], noise_counter=counter)
e1.record(); torch.cuda.synchronize()
prof = _lib.profile_read()
print("DRG_FT_ROWS", os.environ.get("DRG_FT_ROWS", "auto"), "final us", round(1e3 * prof["skh_final"][0] / prof["skh_final"][1], 2),
      "iter us", round(1e3 * prof["skh_iter"][0] / prof["skh_iter"][1], 2), "step us (eager, profiled)", round(1e3 * e0.elapsed_time(e1) / 40, 1))
