import os, sys, time
sys.path.insert(0, "/root/repo")
from types import SimpleNamespace
import torch
import diffreg_b200
from diffreg_b200 import ops, _lib
import bench
dev = torch.device("cuda", 0)
keys = ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")
pb = bench.make_batch(4000, 1, 4800, 2048, 256, invalid_fraction=0.05)
dd = [pb[k].to(dev) for k in keys]
cfg = dict(bench.MATCH_CFG)
h = diffreg_b200.Matching2D3D(cfg).to(dev).eval()
with torch.no_grad():
    h.src_proj.weight.copy_(pb["W"].to(dev))
proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
smp = diffreg_b200.DenoisingSampler("2d3d", h, proc, 10)
x = torch.randn(1, 4800, 2048, device=dev)
shift = None
_lib.profile_enable(True)
for k in range(10):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    x, shift, aux = smp.step(k, x, shift, *dd)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ws = [v for kk, v in ops._workspaces.items() if kk[2] == "procrustes"][0]
    st = ws[:48].cpu().view(torch.int32).tolist()
    pad = st[9] & 0xFFFFFFFF
    prof = _lib.profile_read()
    print(f"step {k}: {dt*1e3:.2f} ms  Kb {st[0]} n_cand {st[1] & 0xFFFFFFFF} path {'GENERAL' if pad >> 31 else 'fast'} bin {pad & 0x7FFFFFFF} segG {st[10]} broken {st[11]}",
          {a: round(1e3*b[0]/max(b[1],1),1) for a, b in prof.items() if b[1]}, flush=True)
    _lib.profile_enable(False); _lib.profile_enable(True)
