"""One fused-attention call at a transformer shape (for ncu captures): python tools/prof_attention.py [L S d H]"""
import os, sys, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffreg_b200 import ops
L, S, d, H = (int(a) for a in (sys.argv[1:5] + ["4096", "4096", "132", "4"][len(sys.argv) - 1:]))
g = torch.Generator().manual_seed(1)
q, k, v = (torch.randn(1, n, H * d, generator=g).cuda() for n in (L, S, S))
q16, k16 = ops.prep_heads(q, H, 0), ops.prep_heads(k, H, 1)
for _ in range(3):
    ops.attention(q16, k16, v, H, None, None, 1.0 / math.sqrt(d), d)
torch.cuda.synchronize()
