"""Fused attention kernel against the three-kernel path at the denoising transformer's shapes (tuning tool; CUDA events)."""
import os, sys, json, math
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffreg_b200 import ops


def timed(fn, reps=7):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


for (H, L, S, d) in ((4, 4096, 4096, 132), (4, 4096, 4096, 64), (4, 2048, 4800, 64), (4, 4800, 2048, 64)):
    g = torch.Generator().manual_seed(1)
    q, k, v = (torch.randn(1, n, H * d, generator=g).cuda() for n in (L, S, S))
    scale = 1.0 / math.sqrt(d)
    q16, k16 = ops.prep_heads(q, H, 0), ops.prep_heads(k, H, 1)
    vt = v.view(1, S, H, d).permute(0, 2, 3, 1).contiguous().view(H, d, S)
    vt16 = ops.prep_operand(vt, 1.0, True, 1)

    def three():
        logits = ops.gemm_nt(q16, k16, split3=True, K=d)
        p16 = ops.attn_softmax(logits, H, None, None, scale)
        return ops.gemm_nt(p16, vt16, split3=True, K=S)

    fused = lambda: ops.attention(q16, k16, v, H, None, None, scale, d)
    a, b = fused(), three().view(1, H, L, d).permute(0, 2, 1, 3).reshape(1, L, H * d)
    flops = 2 * 2.0 * H * L * S * d
    def graphed(fn):          # GPU time without the eager launch gaps: one CUDA-graph replay of the same launches
        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            fn(); fn()
            g_ = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g_, stream=st):
                fn()
        torch.cuda.synchronize()
        return timed(g_.replay)
    tf, t3 = graphed(fused), graphed(three)
    print(json.dumps({"H": H, "L": L, "S": S, "d": d, "fused_us": round(tf, 1), "three_kernel_us": round(t3, 1),
                      "useful_TFLOPs": round(flops / tf / 1e6, 1), "max_abs_diff": (a - b).abs().max().item()}), flush=True)
