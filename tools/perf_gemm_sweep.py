"""Scaling of the split GEMM with K and with the output size (tuning tool): time = a (per output tile) + b * K (per k-step)?"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffreg_b200 import ops

torch.manual_seed(0)


def run(n, m, k, split=True):
    A = torch.randn(1, n, k, device="cuda") / 16; B = torch.randn(1, m, k, device="cuda") / 16
    if split:
        a, b = ops.prep_operand(A, 1.0, True, 0), ops.prep_operand(B, 1.0, True, 1)
    out = torch.empty(1, n, m, device="cuda")
    fn = (lambda: ops.gemm_nt(a, b, out=out, split3=True, K=k)) if split else (lambda: ops.gemm_nt(A, B, out=out))
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / 20


for (n, m, k) in [(4096, 4096, 64), (4096, 4096, 128), (4096, 4096, 256), (4096, 4096, 512), (4096, 4096, 1024), (2048, 2048, 256),
                  (8192, 8192, 256), (4096, 2048, 256), (4736, 4096, 256), (4096, 4096, 192)]:
    print(json.dumps({"N": n, "M": m, "K": k, "split_us": round(run(n, m, k), 1), "tf32_us": round(run(n, m, k, False), 1)}), flush=True)
