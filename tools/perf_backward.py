"""Forward + backward of log_optimal_transport at the sampler's shape: this library's autograd.Function (CUDA forward and backward
kernels) against torch's autograd through the reference function on the same GPU (tuning tool; CUDA events)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffreg_b200
from oracle import ref_loader

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
g = torch.Generator().manual_seed(1)
scores0 = torch.randn(1, n, n, generator=g).cuda()
W = torch.randn(1, n + 1, n + 1, generator=g).cuda()
ones = torch.ones(1, n, dtype=torch.bool).cuda()


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def run(fn_lot):
    s = scores0.clone().requires_grad_()
    a = torch.tensor(1.0, device="cuda", requires_grad=True)
    out = fn_lot(s, a, iters, ones, ones)
    (out * W).sum().backward()
    return s.grad, a.grad


out = {"n": n, "iters": iters}
gs, ga = run(diffreg_b200.log_optimal_transport)
out["dropin_fwd_bwd_ms"] = timed(lambda: run(diffreg_b200.log_optimal_transport))
if ref_loader.available():
    ref = ref_loader.load_flavour("4d")
    rs, ra = run(ref.matching.log_optimal_transport)
    out["reference_autograd_fwd_bwd_ms"] = timed(lambda: run(ref.matching.log_optimal_transport))
    out["max_rel_diff_grad_scores"] = ((gs - rs).abs().max() / rs.abs().max()).item()
    out["rel_diff_grad_alpha"] = (abs(ga.item() - ra.item()) / max(1.0, abs(ra.item())))
from diffreg_b200 import ops
al = torch.tensor(1.0).cuda()
pots = ops.sinkhorn_potentials_per_iteration(scores0, al, iters, ones, ones)
out["backward_only_ms"] = timed(lambda: ops.sinkhorn_backward(scores0, al, iters, ones, ones, W, potentials=pots))
out["potentials_per_iteration_ms"] = timed(lambda: ops.sinkhorn_potentials_per_iteration(scores0, al, iters, ones, ones))
print(json.dumps(out), flush=True)
