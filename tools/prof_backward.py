"""One forward + backward of log_optimal_transport at the sampler's shape (for ncu launch lists): python tools/prof_backward.py [n iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffreg_b200
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator().manual_seed(1)
s0 = torch.randn(1, n, n, generator=g).cuda()
W = torch.randn(1, n + 1, n + 1, generator=g).cuda()
ones = torch.ones(1, n, dtype=torch.bool).cuda()
for _ in range(2):
    s = s0.clone().requires_grad_()
    a = torch.tensor(1.0, device="cuda", requires_grad=True)
    (diffreg_b200.log_optimal_transport(s, a, iters, ones, ones) * W).sum().backward()
torch.cuda.synchronize()
