"""Diagnostics of the pose step over the 20 sampler steps of the bench workload: per step the number of candidates, the
select kernel's path (fast short list / general select) and the size of the crossing histogram bin, plus CUDA-event times of
the three procrustes stages (eager).    python tools/procr_state.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
import diffreg_b200
from diffreg_b200 import ops, _lib
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
for i in range(3):
    smp.step(i % 20, bufs[i % 2], None, *feats, x_out=bufs[(i + 1) % 2], noise_counter=counter)
torch.cuda.synchronize()
bufs[0].copy_(d["x_T"])
_lib.profile_enable(True)
for i in range(20):
    smp.step(i, bufs[i % 2], None, *feats, x_out=bufs[(i + 1) % 2], noise_counter=counter)
    torch.cuda.synchronize()
    ws = [v for k, v in ops._workspaces.items() if k[2] == "procrustes"][0]
    st = ws[:40].cpu().view(torch.int32).tolist()
    pad = st[9] & 0xFFFFFFFF
    prof = _lib.profile_read()
    print(f"step {i:2d} Kb {st[0]} n_cand {st[1] & 0xFFFFFFFF} sh {st[7]} path {'GENERAL' if pad >> 31 else 'fast'} bin {pad & 0x7FFFFFFF}  "
          + " ".join(f"{k} {1e3 * v[0] / max(v[1], 1):.1f}us" for k, v in prof.items() if k in ("topk_threshold", "topk_collect", "procr_select", "procr_solve")), flush=True)
    _lib.profile_enable(False); _lib.profile_enable(True)
