"""Error of the CUDA log-domain Sinkhorn against an fp64 evaluation of the reference's formula, as a function of the
iteration count, on the 48 x 64 golden fixture (lot_iters100) and on random problems (tuning / analysis tool)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from helpers import load
from oracle import diffreg_oracle as O
from diffreg_b200 import ops

g = load("lot_iters100")
s, sm, tm = g["scores"], g["src_mask"], g["tgt_mask"]
al = torch.tensor(float(g["alpha"]))
for it in (1, 2, 3, 5, 10, 20, 30, 50, 100):
    ref = O.log_optimal_transport(s.double(), al.double(), it, sm, tm)
    out, u, v = ops.sinkhorn(s.cuda(), al.cuda(), it, sm.cuda(), tm.cuda(), out_mode="log_full", return_potentials=True)
    err = (out.cpu().double() - ref).abs()
    i, j = divmod(int(err[0].argmax()), err.shape[2])
    print(f"fixture 48x64 iters {it:3d}: max err {err.max().item():.3e} at ({i},{j}) value {ref[0, i, j].item():.4f}  "
          f"err over real block {err[0, :-1, :-1].max().item():.2e}  dustbin row {err[0, -1, :].max().item():.2e}  dustbin col {err[0, :, -1].max().item():.2e}")
gen = torch.Generator().manual_seed(0)
for (N, M, scale) in [(48, 64, 1.0), (48, 64, 3.0), (512, 512, 1.0), (2048, 2048, 1.0)]:
    s = torch.randn(1, N, M, generator=gen) * scale
    sm = torch.ones(1, N, dtype=torch.bool)
    tm = torch.ones(1, M, dtype=torch.bool)
    for it in (3, 100):
        ref = O.log_optimal_transport(s.double().cuda(), al.double().cuda(), it, sm.cuda(), tm.cuda())
        out = ops.sinkhorn(s.cuda(), al.cuda(), it, sm.cuda(), tm.cuda(), out_mode="log_full")
        err = (out.double() - ref).abs()
        print(f"random {N}x{M} scale {scale} iters {it:3d}: max err {err.max().item():.3e}  real block {err[0, :-1, :-1].max().item():.2e}  "
              f"dustbin row {err[0, -1, :].max().item():.2e} col {err[0, :, -1].max().item():.2e}")
