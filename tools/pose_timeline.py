"""Phase times of the fused noisy-matching -> pose call (drg_sinkhorn_soft_procrustes) at the headline shape: clock64
stamps of CTA 0 of the persistent Sinkhorn (iterations + candidate-search tail), CUDA-event times of the two kernels and
the candidate statistics the pose kernel saw (tuning tool)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffreg_b200
from diffreg_b200 import ops, _lib
import bench

lib = diffreg_b200.load_library()
lib.drg_tuning_set_stamp_buffer.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda", 0)
n = 4096
FLUSH = "--flush" in sys.argv       # stand-alone calls, L2 flushed (256 MB memset) before each: cold code and data
IN_STEP = "--step" in sys.argv      # stamps of the LAST of 8 real sampler steps instead of stand-alone calls on x_T
host = bench.make_inputs(3000, n, 256)
d = {k: v.to(dev) for k, v in host.items()}
alpha = torch.tensor(1.0, device=dev)
x = d["x_T"].clone()
stamps = torch.zeros(1024, dtype=torch.int64, device=dev)
if IN_STEP:
    from types import SimpleNamespace
    head = diffreg_b200.Matching(bench.MATCH_CFG).to(dev).eval()
    with torch.no_grad():
        head.src_proj.weight.copy_(d["W"])
    proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
    smp = diffreg_b200.DenoisingSampler("4d", head, proc, bench.SAMPLER_STEPS, noise_seed=1234)
    feats = [d[k] for k in ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")]
    bufs = [x, torch.empty_like(x)]
    counter = torch.zeros(1, dtype=torch.int64, device=dev)
    state = {"i": 0}

    def call():
        i = state["i"]
        smp.step(i % 20, bufs[i % 2], None, *feats, x_out=bufs[(i + 1) % 2], noise_counter=counter)
        state["i"] += 1
else:
    call = lambda: ops.sinkhorn_soft_procrustes(x, alpha, 3, d["src_mask"], d["tgt_mask"], d["s_pcd"], d["t_pcd"], 1.0, 40.0)
if FLUSH:
    junk = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    inner = call

    def call():
        junk.zero_()
        inner()
for _ in range(3):
    call()
torch.cuda.synchronize()
lib.drg_tuning_set_stamp_buffer(stamps.data_ptr())
_lib.profile_enable(True)
for _ in range(5):
    call()
torch.cuda.synchronize()
prof = _lib.profile_read()
_lib.profile_enable(False)
lib.drg_tuning_set_stamp_buffer(None)
t = stamps.tolist()
MHZ = 1965.0
us = lambda a, b: (t[b] - t[a]) / MHZ if t[a] and t[b] else float("nan")
print("kernel times (CUDA events, eager):", {k: round(1e3 * v[0] / max(v[1], 1), 1) for k, v in prof.items() if v[1]})
print(f"persist2: setup {us(0, 1):.1f} us")
for it in range(3):
    b0 = 10 + it * 100
    print(f"  it {it}: prologue {us(b0, b0 + 1):.1f}  pass {us(b0 + 1, b0 + 2):.1f}  partials {us(b0 + 2, b0 + 3):.1f}  barrier1 {us(b0 + 3, b0 + 4):.1f}"
          f"  merge {us(b0 + 4, b0 + 5):.1f}  barrier2 {us(b0 + 5, b0 + 6):.1f}")
print(f"  tail: bound {us(700, 701):.1f}  pass {us(701, 702):.1f}  flush {us(702, 703):.1f}   total kernel {us(704, 703):.1f} us")
ns = lambda a, b: (t[b] - t[a]) / 1e3 if t[a] and t[b] else float("nan")
print(f"pose kernel (CTA 0): hist+state+walk {ns(800, 801):.1f}  scan {ns(801, 802):.1f}  T {ns(802, 803):.1f}  sort {ns(803, 804):.1f}  "
      f"gather+moments {ns(804, 805):.1f}  reduce+ticket {ns(805, 806):.1f}")
print(f"pose kernel (finisher warp of CTA 0): start->all sums arrived {ns(800, 807):.1f}  combine {ns(807, 808):.1f}  svd {ns(808, 809):.1f}  "
      f"publish + warp of CTA 0's rows {ns(809, 810):.1f}  total {ns(800, 810):.1f} us")
if t[908] and t[909]:
    print(f"SM clock during the single-thread SVD: {(t[909] - t[908]) / max(t[809] - t[808], 1) * 1e3:.0f} MHz (clock64 / globaltimer)")
ws = [v for k, v in ops._workspaces.items() if k[2] == "procrustes"][0]
st = ws[:40].cpu().view(torch.int32).tolist()
pad = st[9] & 0xFFFFFFFF
print(f"state: Kb {st[0]} n_cand {st[1] & 0xFFFFFFFF} kmin {st[6] & 0xFFFFFFFF:#x} sh {st[7]} path {'GENERAL' if pad >> 31 else 'fast'} crossing-bin size {pad & 0x7FFFFFFF}")
