"""Forward + backward of the matching head and the pose layer in training mode (Matching.forward -> conf -> SoftProcrustes -> a loss on
conf, R, t -> backward) at the sampler's shape: this library's modules against the reference modules on the same GPU (tuning tool)."""
import os, sys, json
from types import SimpleNamespace
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import diffreg_b200
from oracle import ref_loader, diffreg_oracle as O

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
C = 256
pb = O.make_problem(7, 1, n, n, C)
cfg = dict(match_type="sinkhorn", confidence_threshold=0.2, feature_dim=C, entangled=True, dsmax_temperature=0.1,
           skh_init_bin_score=1.0, skh_iters=3, skh_prefilter=False)
pcfg = SimpleNamespace(sample_rate=1.0, max_condition_num=1e9)
dev = "cuda"
t = {k: pb[k].to(dev) for k in ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")}
g = torch.Generator().manual_seed(1)
Wc = torch.rand(1, n, n, generator=g).to(dev)
gR, gt = torch.randn(1, 3, 3, generator=g).to(dev), torch.randn(1, 3, 1, generator=g).to(dev)


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def step(head, proc, with_pose=True):
    src, tgt = t["src_feats"].clone().requires_grad_(), t["tgt_feats"].clone().requires_grad_()
    head.zero_grad(set_to_none=True)
    conf, _ = head(src, tgt, None, None, t["src_mask"], t["tgt_mask"], {})
    loss = (conf * Wc).sum()
    if with_pose:
        R, tt, _, _, _, _ = proc(conf, t["s_pcd"], t["t_pcd"], t["src_mask"], t["tgt_mask"])
        loss = loss + (R * gR).sum() + (tt * gt).sum()
    loss.backward()
    return src.grad, head.src_proj.weight.grad, head.bin_score.grad


out = {"n": n, "C": C}
head = diffreg_b200.Matching(cfg).to(dev).train()
proc = diffreg_b200.SoftProcrustesLayer(pcfg)
a = step(head, proc)
out["dropin_matching_fwd_bwd_ms"] = timed(lambda: step(head, proc, False))
out["dropin_matching_pose_fwd_bwd_ms"] = timed(lambda: step(head, proc, True))
if ref_loader.available():
    ref = ref_loader.load_flavour("4d")
    torch.backends.cuda.matmul.allow_tf32 = False
    rhead = ref.matching.Matching(cfg).to(dev).train()
    rhead.load_state_dict(head.state_dict())
    rproc = ref.procrustes.SoftProcrustesLayer(pcfg)
    b = step(rhead, rproc)
    out["reference_matching_fwd_bwd_ms"] = timed(lambda: step(rhead, rproc, False))
    out["reference_matching_pose_fwd_bwd_ms"] = timed(lambda: step(rhead, rproc, True), reps=3)
    out["rel_diff_grad_src_feats"] = ((a[0] - b[0]).abs().max() / b[0].abs().max()).item()
    out["rel_diff_grad_weight"] = ((a[1] - b[1]).abs().max() / b[1].abs().max()).item()
    out["rel_diff_grad_bin_score"] = (abs(a[2].item() - b[2].item()) / max(1.0, abs(b[2].item())))
print(json.dumps(out), flush=True)
