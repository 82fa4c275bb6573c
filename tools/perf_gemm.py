"""Similarity GEMM timing at the headline shape."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffreg_b200 import ops
n, k = 4096, 256
A = torch.randn(1, n, k, device="cuda"); B = torch.randn(1, n, k, device="cuda")
A3 = ops.prep_operand(A, 1.0 / 16, True, 0); B3 = ops.prep_operand(B, 1.0 / 16, True, 1)
out = torch.empty(1, n, n, device="cuda")
for split3 in (False, True):
    for _ in range(3):
        ops.gemm_nt(A3, B3, out=out, split3=True) if split3 else ops.gemm_nt(A / 16, B / 16, out=out)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.gemm_nt(A3, B3, out=out, split3=True) if split3 else ops.gemm_nt(A / 16, B / 16, out=out)
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / 20
    print(json.dumps({"split3": split3, "us": round(us, 1),
                      "useful_TFLOPs": round(2 * k * n * n / us / 1e6, 1)}), flush=True)
