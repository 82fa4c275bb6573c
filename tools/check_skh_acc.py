import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from diffreg_b200 import ops
dev = "cuda"
N = M = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
g = torch.Generator(device=dev).manual_seed(7)
s = torch.randn(1, N, M, generator=g, device=dev)
sm = torch.ones(1, N, dtype=torch.bool, device=dev); tm = torch.ones(1, M, dtype=torch.bool, device=dev)
alpha = torch.tensor(1.0, device=dev)
for iters in (1, 3):
    out, u, v = ops.sinkhorn(s, alpha, iters, sm, tm, out_mode="log_full", return_potentials=True)
    # fp64 reference
    Z = torch.full((1, N + 1, M + 1), 1.0, dtype=torch.float64, device=dev); Z[:, :N, :M] = s.double()
    norm = -torch.log(torch.tensor(float(N + M), dtype=torch.float64, device=dev))
    log_mu = torch.cat([norm.expand(N), (torch.log(torch.tensor(float(M), dtype=torch.float64, device=dev)) + norm)[None]])[None]
    log_nu = torch.cat([norm.expand(M), (torch.log(torch.tensor(float(N), dtype=torch.float64, device=dev)) + norm)[None]])[None]
    ur = torch.zeros_like(log_mu); vr = torch.zeros_like(log_nu)
    for _ in range(iters):
        ur = log_mu - torch.logsumexp(Z + vr[:, None, :], dim=2)
        vr = log_nu - torch.logsumexp(Z + ur[:, :, None], dim=1)
    du = (u.double() - ur).abs(); dv = (v.double() - vr).abs()
    print(json.dumps(dict(ws=os.environ.get("DRG_SKH_WS"), iters=iters, du_max=du.max().item(), du_argmax=int(du.argmax()), dv_max=dv.max().item(),
                          dv_argmax=int(dv.argmax()), du_mean=du.mean().item(), dv_mean=dv.mean().item(),
                          dv_big=int((dv > 2e-5).sum()), du_big=int((du > 2e-5).sum()))), flush=True)
