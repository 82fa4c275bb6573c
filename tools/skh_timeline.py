"""Timeline of CTA 0 of the persistent Sinkhorn kernel from clock64 stamps (DRG_SKH_TIMES=1)."""
import sys, os, ctypes
os.environ["DRG_SKH_TIMES"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffreg_b200
from diffreg_b200 import ops
lib = diffreg_b200.load_library()
N = M = 4096
s = torch.randn(1, N, M, device="cuda")
sm = torch.ones(1, N, dtype=torch.bool, device="cuda"); tm = torch.ones(1, M, dtype=torch.bool, device="cuda")
alpha = torch.tensor(1.0, device="cuda")
for _ in range(3):
    ops.sinkhorn(s, alpha, 3, sm, tm, out_mode="none")
torch.cuda.synchronize()
buf = (ctypes.c_longlong * 512)()
lib.drg_debug_read_times.argtypes = [ctypes.c_void_p, ctypes.c_int]
assert lib.drg_debug_read_times(buf, 512) == 0
t = list(buf)
t0 = t[0]
rel = lambda k: (t[k] - t0) if t[k] else None
print("start->setup done (mask counts):", rel(1))
for it in range(3):
    base = 10 + it * 100
    print(f"--- iteration {it}")
    print(" iter start", rel(base + 0), " prologue done", rel(base + 1), " row warp0 done", rel(base + 2), " phaseA sync", rel(base + 3),
          " grid barrier 1", rel(base + 4), " merge done", rel(base + 5), " grid barrier 2", rel(base + 6))
    print(" row warp0: full-wait passed per slab:", [rel(base + 20 + k) for k in range(8)])
    print(" col warp16: u_ready passed per slab: ", [rel(base + 40 + k) for k in range(8)])
    print(" col warp16: slab done:               ", [rel(base + 60 + k) for k in range(8)])

import struct
print("fast flags:", t[400:403], "dv bits:", [hex(x & 0xffffffff) for x in t[410:413]], [struct.unpack("f", struct.pack("I", x & 0xffffffff))[0] if x >= 0 else None for x in t[410:413]])
