"""In-kernel timeline of the fused attention kernel (clock64 stamps of CTA (0, 0): the MMA warp and softmax thread 0, per key tile)
through the library's tuning hook drg_tuning_set_stamp_buffer: python tools/fa_timeline.py L S d H"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, ctypes
from diffreg_b200 import ops, load_library
L, S, d, H = (int(a) for a in sys.argv[1:5])
g = torch.Generator().manual_seed(1)
q, k, v = (torch.randn(1, n, H * d, generator=g).cuda() for n in (L, S, S))
q16, k16 = ops.prep_heads(q, H, 0), ops.prep_heads(k, H, 1)
lib = load_library()
tl = torch.zeros(64 * 16, dtype=torch.int64, device="cuda")
for _ in range(2):
    ops.attention(q16, k16, v, H, None, None, 1.0 / math.sqrt(d), d)
lib.drg_tuning_set_stamp_buffer.argtypes = [ctypes.c_void_p]
lib.drg_tuning_set_stamp_buffer(tl.data_ptr())
ops.attention(q16, k16, v, H, None, None, 1.0 / math.sqrt(d), d)
torch.cuda.synchronize()
t = tl.cpu().view(64, 16)
names = ["mma:iter start", "mma:S(t+1) issued", "mma:p(t) seen", "mma:PV issued", "sm:tile start", "sm:S ready", "sm:ld done", "sm:max written",
         "sm:barrier passed", "sm:exp done", "sm:pv(t-1) seen", "sm:P stored", "sm:arrived", "mma:kfull seen", "mma:S mmas issued", "sm:loop top"]
base = int(t[8, 4])
for tt in range(10, 13):
    print("tile", tt, " ".join(f"{names[i]}={int(t[tt, i]) - base}" for i in range(16)))

ts = t[:, 4] - t[0, 4]
print("softmax tile starts (cycles since tile 0):", [int(x) for x in ts[:min(64, (S + 63) // 64)].tolist()])
