"""Single Sinkhorn call repeated a few times, for ncu launch lists / captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffreg_b200 import ops
N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
M = int(sys.argv[2]) if len(sys.argv) > 2 else N
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = "cuda"
s = torch.randn(1, N, M, device=dev)
sm = torch.ones(1, N, dtype=torch.bool, device=dev); tm = torch.ones(1, M, dtype=torch.bool, device=dev)
alpha = torch.tensor(1.0, device=dev)
for _ in range(reps):
    ops.sinkhorn(s, alpha, 3, sm, tm, out_mode="conf")
torch.cuda.synchronize()
