"""Debug/verification driver for the tcgen05 GEMM (run under `timeout` on the GPU box)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffreg_b200 import ops

dev = "cuda"
torch.manual_seed(0)
shapes = [(1, 128, 64, 32), (1, 128, 256, 32), (1, 128, 256, 256), (1, 256, 512, 64), (1, 100, 70, 36), (2, 300, 1000, 528),
          (1, 1024, 1024, 256), (1, 4096, 4096, 256), (1, 4096, 4096, 768), (3, 130, 1530, 256), (1, 333, 1531, 40),
          (1, 8192, 256, 768), (16, 2048, 2048, 256)]
if len(sys.argv) > 1:
    shapes = shapes[: int(sys.argv[1])]
for (b, n, m, k) in shapes:
    A = torch.randn(b, n, k, device=dev)
    B = torch.randn(b, m, k, device=dev)
    C = ops.gemm_nt(A, B, alpha=0.5)
    torch.cuda.synchronize()
    ref = 0.5 * torch.einsum("bnk,bmk->bnm", A.double(), B.double())
    err = (C.double() - ref).abs().max().item()
    scale = ref.abs().max().item()
    # 3xTF32
    A3 = ops.prep_operand(A, 1.0, True, 0)
    B3 = ops.prep_operand(B, 1.0, True, 1)
    C3 = ops.gemm_nt(A3, B3, alpha=0.5, split3=True, K=k)
    torch.cuda.synchronize()
    err3 = (C3.double() - ref).abs().max().item()
    ref32 = 0.5 * torch.einsum("bnk,bmk->bnm", A, B)
    err32 = (ref32.double() - ref).abs().max().item()
    print(json.dumps(dict(shape=[b, n, m, k], max_abs_err_tf32=err, max_abs_err_3xtf32=err3, torch_fp32_err=err32, ref_max=scale)), flush=True)
