"""Golden gradients of log_optimal_transport and of Matching.forward (Sinkhorn branch, training mode) from the UNMODIFIED reference's
autograd.  Run in the build container only (needs /root/reference):  python tests/golden/make_golden_lotb.py"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402

torch.set_num_threads(1)
ref = ref_loader.load_flavour("4d")


def save(tag, out):
    np.savez_compressed(os.path.join(HERE, tag + ".npz"),
                        **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()})
    print("wrote", tag)


def run_lot(tag, B, N, M, iters, kind, seed):
    g = torch.Generator().manual_seed(seed)
    scores = (torch.randn(B, N, M, generator=g) * 2.0)
    sm, tm = torch.ones(B, N, dtype=torch.bool), torch.ones(B, M, dtype=torch.bool)
    if kind == "prefix":
        sm[0, N - 5:] = False
        tm[B - 1, M - 7:] = False
    elif kind == "arbitrary":
        sm, tm = torch.rand(B, N, generator=g) > 0.1, torch.rand(B, M, generator=g) > 0.1
    scores = scores.masked_fill(~(sm[..., None] * tm[:, None]).bool(), float("-inf")).requires_grad_()   # as matching.py:163-165 leaves them
    alpha = torch.tensor(1.0 + 0.25 * seed % 1, requires_grad=True)
    out = ref.matching.log_optimal_transport(scores, alpha, iters, sm, tm)
    W = torch.randn(out.shape, generator=g)
    (out * W).sum().backward()
    save(tag, dict(scores=scores.detach(), alpha=alpha.detach(), iters=iters, src_mask=sm, tgt_mask=tm, grad_out=W, out=out.detach(),
                   grad_scores=scores.grad, grad_alpha=alpha.grad))


def run_matching(tag, B, N, M, C, entangled, seed, match_type="sinkhorn"):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    cfg = dict(match_type=match_type, confidence_threshold=0.2, feature_dim=C, entangled=entangled, dsmax_temperature=0.1,
               skh_init_bin_score=1.0, skh_iters=3, skh_prefilter=False)
    head = ref.matching.Matching(cfg).train()
    src = torch.randn(B, N, C, generator=g).requires_grad_()
    tgt = torch.randn(B, M, C, generator=g).requires_grad_()
    sm, tm = torch.ones(B, N, dtype=torch.bool), torch.ones(B, M, dtype=torch.bool)
    sm[0, N - 4:] = False
    tm[B - 1, M - 6:] = False
    pe = {}
    if not entangled:
        a_s, a_t = torch.rand(B, N, C // 2, generator=g) * 6.28, torch.rand(B, M, C // 2, generator=g) * 6.28
        dup = lambda a: torch.stack([a, a], -1).reshape(a.shape[0], a.shape[1], C)
        pe = dict(src_pe=torch.stack([dup(a_s).cos(), dup(a_s).sin()], -1), tgt_pe=torch.stack([dup(a_t).cos(), dup(a_t).sin()], -1))
    conf, match = head(src, tgt, pe.get("src_pe"), pe.get("tgt_pe"), sm, tm, {}, pe_type="rotary")
    W = torch.rand(conf.shape, generator=g)
    (conf * W).sum().backward()
    rec = dict(src_feats=src.detach(), tgt_feats=tgt.detach(), src_mask=sm, tgt_mask=tm, entangled=int(entangled), W=W, match_type=match_type,
               conf=conf.detach(), match=match, grad_src=src.grad, grad_tgt=tgt.grad, grad_weight=head.src_proj.weight.grad,
               weight=head.src_proj.weight.detach())
    if match_type == "sinkhorn":
        rec.update(grad_bin_score=head.bin_score.grad, bin_score=head.bin_score.detach())
    rec.update(pe)
    save(tag, rec)


run_lot("lotb_full_i3", 2, 23, 31, 3, "none", 81)
run_lot("lotb_prefix_i3", 2, 40, 33, 3, "prefix", 82)
run_lot("lotb_arbitrary_i5", 1, 37, 52, 5, "arbitrary", 83)
run_lot("lotb_wide_i1", 1, 9, 300, 1, "prefix", 84)
run_matching("lotb_matching_train_entangled", 2, 30, 26, 32, True, 85)
run_matching("lotb_matching_train_rotary", 1, 28, 35, 24, False, 86)
run_matching("lotb_matching_train_dualsoftmax", 2, 27, 33, 32, True, 87, match_type="dual_softmax")


class cpu_cuda:
    def __enter__(self):
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._orig


def run_matching_2d3d(tag, N, M, C, seed):
    """The 2D-3D head in training mode (experiments/<exp>/matching.py:91-147): conf_matrix and the gathered weights carry gradients."""
    r2 = ref_loader.load_flavour("2d3d")
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    cfg = dict(match_type="sinkhorn", confidence_threshold=0.2, feature_dim=C, entangled=True, dsmax_temperature=0.1,
               skh_init_bin_score=1.0, skh_iters=3, skh_prefilter=False)
    head = r2.matching.Matching(cfg).train()
    src = torch.randn(1, N, C, generator=g).requires_grad_()
    tgt = torch.randn(1, M, C, generator=g).requires_grad_()
    sm, tm = torch.rand(1, N, generator=g) > 0.1, torch.rand(1, M, generator=g) > 0.1
    with cpu_cuda():
        conf, si, ti, w = head(src, tgt, sm, tm, True)
    W = torch.rand(conf.shape, generator=g)
    ((conf * W).sum() + (w * torch.arange(1, w.numel() + 1)).sum()).backward()
    save(tag, dict(src_feats=src.detach(), tgt_feats=tgt.detach(), src_mask=sm, tgt_mask=tm, W=W, conf=conf.detach(), src_indices=si,
                   tgt_indices=ti, weights=w.detach(), grad_src=src.grad, grad_tgt=tgt.grad, grad_weight=head.src_proj.weight.grad,
                   grad_bin_score=head.bin_score.grad, weight=head.src_proj.weight.detach(), bin_score=head.bin_score.detach()))


run_matching_2d3d("lotb_matching2d3d_train", 31, 26, 32, 88)
