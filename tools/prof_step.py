"""A few eager denoising steps at the headline shape, for ncu launch lists / captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
import diffreg_b200
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
host = bench.make_inputs(3000, n, bench.FEAT_DIM)
d = {k: v.to(dev) for k, v in host.items()}
head = diffreg_b200.Matching(bench.MATCH_CFG).to(dev).eval()
with torch.no_grad():
    head.src_proj.weight.copy_(d["W"])
proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
smp = diffreg_b200.DenoisingSampler("4d", head, proc, bench.SAMPLER_STEPS, noise_seed=1)
feats = [d[k] for k in ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")]
x = d["x_T"]
for i in range(steps):
    x, _, aux = smp.step(i, x, None, *feats)
torch.cuda.synchronize()
print("cond", aux["pose"]["condition"].item(), "matches", int(aux["match"][3].item()))
