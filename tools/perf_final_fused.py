"""Where the time of the fused Sinkhorn + final-pass launch goes (tuning tool): the same 4096^2 call with the parts of the
final phase switched on one by one, CUDA-event time per call (L2 flushed between calls by streaming other matrices)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffreg_b200 import ops

n = 4096
g = torch.Generator(device="cuda").manual_seed(1)
sim = torch.randn(1, n, n, generator=g, device="cuda")
x_t = torch.randn(1, n, n, generator=g, device="cuda")
noise = torch.randn(1, n, n, generator=g, device="cuda")
other = [torch.randn(1, n, n, generator=g, device="cuda") for _ in range(2)]
out = torch.empty(1, n, n, device="cuda")
alpha = torch.tensor(1.0, device="cuda")
m = torch.ones(1, n, dtype=torch.bool, device="cuda")


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        other[0].add_(other[1])                      # stream 192 MB through L2: the next call starts cold, as in a step
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    return sorted(ts)[len(ts) // 2]


cases = {
    "potentials only (3 iterations)": lambda: ops.sinkhorn(sim, alpha, 3, m, m, out_mode="none"),
    "+ conf": lambda: ops.sinkhorn(sim, alpha, 3, m, m, out_mode="conf", out=out),
    "+ ddim, no noise": lambda: ops.sinkhorn(sim, alpha, 3, m, m, out_mode="ddim", x_t=x_t, k_x0=0.7, k_xt=0.1, sigma=0.0, out=out),
    "+ ddim, noise tensor": lambda: ops.sinkhorn(sim, alpha, 3, m, m, out_mode="ddim", x_t=x_t, noise=noise, k_x0=0.7, k_xt=0.1, sigma=0.3, out=out),
    "+ ddim, philox": lambda: ops.sinkhorn(sim, alpha, 3, m, m, out_mode="ddim", x_t=x_t, k_x0=0.7, k_xt=0.1, sigma=0.3, noise_seed=7, out=out),
    "+ ddim, philox, keys (floor 0.2)": lambda: ops.sinkhorn(sim, alpha, 3, m, m, out_mode="ddim", x_t=x_t, k_x0=0.7, k_xt=0.1, sigma=0.3,
                                                            noise_seed=7, out=out, want_best=True, best_floor=0.2),
}
for name, fn in cases.items():
    print(json.dumps({"case": name, "us": round(timed(fn), 1)}), flush=True)
