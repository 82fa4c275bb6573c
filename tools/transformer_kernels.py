"""Per-kernel device times of one forward of the denoising transformer drop-in (torch.profiler / CUPTI; tuning tool)."""
import os, sys
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import diffreg_b200


class Cfg(dict):
    __getattr__ = dict.__getitem__


n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
C = int(sys.argv[2]) if len(sys.argv) > 2 else 528
bnds = [[-3.6, -2.4, 1.14], [1.093, 0.78, 2.92]]
cfg = Cfg(feature_dim=C, n_head=4, layer_types=['self', 'cross'] * 3, positioning_type="procrustes", pe_type="rotary", entangled=False,
          vol_bnds=bnds, voxel_size=0.04)
g = torch.Generator().manual_seed(1)
lo, hi = torch.tensor(bnds[0]), torch.tensor(bnds[1])
s_pcd = (lo + (hi - lo) * torch.rand(1, n, 3, generator=g)).cuda()
t_pcd = (lo + (hi - lo) * torch.rand(1, n, 3, generator=g)).cuda()
sf, tf = torch.randn(1, n, C, generator=g).cuda(), torch.randn(1, n, C, generator=g).cuda()
sm = torch.ones(1, n, dtype=torch.bool).cuda()
net = diffreg_b200.RepositioningTransformer(cfg).cuda().eval()
net.graph_replay = False      # per-kernel times of the eager launches
for _ in range(2):
    net(sf, tf, s_pcd, t_pcd, sm, sm, {})
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    net(sf, tf, s_pcd, t_pcd, sm, sm, {})
    torch.cuda.synchronize()
tot, cnt = defaultdict(float), defaultdict(int)
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0][:80]
        tot[name] += ev.device_time
        cnt[name] += 1
total = sum(tot.values())
for name in sorted(tot, key=lambda k: -tot[k]):
    print(f"{tot[name]:9.1f} us  {cnt[name]:4d} launches  {tot[name] / cnt[name]:8.1f} us each  {name}")
print(f"{total:9.1f} us in kernels + copies per forward (n = {n}, C = {C}, 6 layers x 2 directions)")
