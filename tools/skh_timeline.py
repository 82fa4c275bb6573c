"""Timeline of CTA 0 of the persistent Sinkhorn kernel from its clock64 stamps (tuning tool).
The stamp buffer is caller memory handed to the library through drg_tuning_set_stamp_buffer."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffreg_b200
from diffreg_b200 import ops

lib = diffreg_b200.load_library()
lib.drg_tuning_set_stamp_buffer.argtypes = [ctypes.c_void_p]
N = M = 4096
s = torch.randn(1, N, M, device="cuda")
sm = torch.ones(1, N, dtype=torch.bool, device="cuda")
tm = torch.ones(1, M, dtype=torch.bool, device="cuda")
alpha = torch.tensor(1.0, device="cuda")
stamps = torch.zeros(512, dtype=torch.int64, device="cuda")
lib.drg_tuning_set_stamp_buffer(stamps.data_ptr())
for _ in range(3):
    ops.sinkhorn(s, alpha, 3, sm, tm, out_mode="none")
torch.cuda.synchronize()
lib.drg_tuning_set_stamp_buffer(None)
t = stamps.tolist()
t0 = t[0]
rel = lambda k: (t[k] - t0) if t[k] else None
print("start->setup done (mask counts):", rel(1))
for it in range(3):
    base = 10 + it * 100
    print(f"--- iteration {it} (mode {t[400 + it]}: 1 scaled, 2 first-iteration scaled, 0 log-domain)")
    print(" iter start", rel(base + 0), " prologue done", rel(base + 1), " pass done", rel(base + 2), " partials written", rel(base + 3),
          " grid barrier 1", rel(base + 4), " merge done", rel(base + 5), " grid barrier 2", rel(base + 6))
