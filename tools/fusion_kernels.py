"""Per-kernel device times of one forward of the 2D-3D fusion module drop-in (torch.profiler / CUPTI; tuning tool)."""
import os, sys
from collections import defaultdict
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import diffreg_b200

n_img = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
n_pcd = int(sys.argv[2]) if len(sys.argv) > 2 else 4800
g = torch.Generator().manual_seed(1)
x = (torch.randn(1, n_img, 512, generator=g).cuda(), torch.randn(1, n_img, 1024, generator=g).cuda(),
     (torch.rand(1, n_img, 2, generator=g) * 2.0 - 1.0).cuda(), torch.randn(1, n_pcd, 512, generator=g).cuda(),
     (torch.randn(1, n_pcd, 3, generator=g) * 0.8).cuda())
net = diffreg_b200.CrossModalFusionModule(512, 512, 256, 256, 4, ["self", "cross"] * 3).cuda().eval()
net.graph_replay = False      # per-kernel times of the eager launches
for _ in range(2):
    net(*x)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    net(*x)
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
print(f"{total:9.1f} us in kernels + copies per forward ({n_img} image patches x {n_pcd} points, 6 blocks)")
