"""Golden gradients of batch_weighted_procrustes and of SoftProcrustesLayer.forward from the UNMODIFIED reference's autograd (through its
host SVD).  Run in the build container only (needs /root/reference):  python tests/golden/make_golden_procrb.py"""
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402
from oracle import diffreg_oracle as O  # noqa: E402

torch.set_num_threads(1)
ref = ref_loader.load_flavour("4d")


def save(tag, out):
    np.savez_compressed(os.path.join(HERE, tag + ".npz"),
                        **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()})
    print("wrote", tag)


def problem(g, B, K, noise):
    X = torch.randn(B, K, 3, generator=g)
    Y = torch.stack([(O.random_rotation(g) @ X[b].t()).t() for b in range(B)]) + noise * torch.randn(B, K, 3, generator=g) + torch.randn(B, 1, 3, generator=g)
    return X, Y


def run_kabsch(tag, B, K, noise, seed):
    g = torch.Generator().manual_seed(seed)
    X, Y = problem(g, B, K, noise)
    w = torch.rand(B, K, 1, generator=g)
    w[:, K - 3:] = 0.0                                     # zeroed tail entries, as forward() leaves them
    w.requires_grad_()
    R, t, cond = ref.procrustes.SoftProcrustesLayer.batch_weighted_procrustes(X, Y, w)
    gR, gt = torch.randn(B, 3, 3, generator=g), torch.randn(B, 3, 1, generator=g)
    ((R * gR).sum() + (t * gt).sum()).backward()
    save(tag, dict(X=X, Y=Y, w=w.detach(), R=R.detach(), t=t.detach(), grad_R=gR, grad_t=gt, grad_w=w.grad))


def run_layer(tag, B, N, M, seed):
    g = torch.Generator().manual_seed(seed)
    src = torch.randn(B, N, 3, generator=g)
    Rs = [O.random_rotation(g) for _ in range(B)]
    perm = torch.randperm(N, generator=g)[:M]
    tgt = torch.stack([(Rs[b] @ src[b, perm].t()).t() for b in range(B)]) + 0.02 * torch.randn(B, M, 3, generator=g)
    conf = torch.rand(B, N, M, generator=g) * 0.05
    for b in range(B):
        conf[b, perm, torch.arange(M)] += 0.5 + 0.4 * torch.rand(M, generator=g)
    sm, tm = torch.ones(B, N, dtype=torch.bool), torch.ones(B, M, dtype=torch.bool)
    sm[0, N - 6:] = False
    conf.requires_grad_()
    layer = ref.procrustes.SoftProcrustesLayer(SimpleNamespace(sample_rate=0.5, max_condition_num=1e6))
    R, t, Rf, tf, cond, mask = layer(conf, src, tgt, sm, tm)
    gR, gt = torch.randn(B, 3, 3, generator=g), torch.randn(B, 3, 1, generator=g)
    ((Rf * gR).sum() + (tf * gt).sum()).backward()
    save(tag, dict(conf=conf.detach(), src_pcd=src, tgt_pcd=tgt, src_mask=sm, tgt_mask=tm, sample_rate=0.5, max_condition_num=1e6,
                   R=R.detach(), t=t.detach(), grad_R=gR, grad_t=gt, grad_conf=conf.grad))


run_kabsch("procrb_kabsch_clean", 2, 60, 0.01, 91)
run_kabsch("procrb_kabsch_noisy", 3, 200, 0.3, 92)
run_layer("procrb_layer", 2, 40, 30, 93)
