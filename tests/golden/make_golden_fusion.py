"""Golden vectors for the 2D-3D flavour's fusion / denoising transformer from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_fusion.py
Imports Diff-Reg-2d3d/experiments/<exp>/fusion_module.py and the two vision3d layer files it is built from as they lie
(oracle/ref_loader.load_fusion), runs CrossModalFusionModule / TransformerLayer / FourierEmbedding on seeded inputs on the CPU
(``FourierEmbedding.forward`` hard-codes ``.cuda()``: ``torch.Tensor.cuda`` is a no-op while it runs) and stores inputs, weights
(state_dict) and outputs as tests/golden/fusion_*.npz."""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import ref_loader  # noqa: E402

torch.set_num_threads(1)
ns = ref_loader.load_fusion()


class cpu_cuda:
    def __enter__(self):
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self

    def __exit__(self, *exc):
        torch.Tensor.cuda = self._orig


def save(tag, out):
    np.savez_compressed(os.path.join(HERE, tag + ".npz"),
                        **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()})
    print("wrote", tag)


def randomise(mod, g):
    for prm in mod.parameters():              # biases and LayerNorm affine away from (0 / 1, 0), weights at a realistic scale
        prm.copy_(torch.randn(prm.shape, generator=g) * (0.3 if prm.dim() == 1 else 1.0 / prm.shape[-1] ** 0.5))
    for name, prm in mod.named_parameters():
        if name.endswith("norm.weight"):
            prm.add_(1.0)


@torch.no_grad()
def run_layer(tag, B, N, M, C, H, masked, seed):
    g = torch.Generator().manual_seed(seed)
    layer = ns.transformer.TransformerLayer(C, H, dropout=None, act_cfg="ReLU").eval()
    randomise(layer, g)
    q, k = torch.randn(B, N, C, generator=g), torch.randn(B, M, C, generator=g)
    km = torch.rand(B, M, generator=g) < 0.2 if masked else None          # True = key ignored
    out, scores = layer(q, k, k, k_masks=km, return_attention_score=True)
    rec = dict(q=q, k=k, has_mask=int(masked), k_masks=km if masked else torch.zeros(B, M, dtype=torch.bool), n_head=H, out=out,
               scores=scores)
    rec.update({"w." + n: v for n, v in layer.state_dict().items()})
    save(tag, rec)


@torch.no_grad()
def run_embedding(tag, rows, n, L, k0, use_pi, use_input, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(2, rows, n, generator=g) * 1.5
    with cpu_cuda():
        out = ns.embedding.FourierEmbedding(L, k0=k0, use_pi=use_pi, use_input=use_input)(x)
    save(tag, dict(x=x, length=L, k0=k0, use_pi=int(use_pi), use_input=int(use_input), out=out))


@torch.no_grad()
def run_fusion(tag, n_img, n_pcd, img_dim, hidden, out_dim, H, blocks, use_embedding, masked, seed):
    g = torch.Generator().manual_seed(seed)
    net = ns.fusion.CrossModalFusionModule(img_dim, img_dim, out_dim, hidden, H, blocks, use_embedding=use_embedding).eval()
    randomise(net, g)
    img_feats, img_dino = torch.randn(1, n_img, img_dim, generator=g), torch.randn(1, n_img, 2 * img_dim, generator=g)
    pcd_feats = torch.randn(1, n_pcd, img_dim, generator=g)
    img_pixels = torch.rand(1, n_img, 2, generator=g) * 2.0 - 1.0                      # normalised pixel coordinates
    pcd_points = torch.randn(1, n_pcd, 3, generator=g) * 0.8 + torch.tensor([0.3, -0.2, 2.0])
    im = torch.rand(1, n_img, generator=g) < 0.1 if masked else None
    pm = torch.rand(1, n_pcd, generator=g) < 0.1 if masked else None
    with cpu_cuda():
        io, po = net(img_feats, img_dino, img_pixels, pcd_feats, pcd_points, im, pm)
    rec = dict(img_feats=img_feats, img_feats_dino=img_dino, img_pixels=img_pixels, pcd_feats=pcd_feats, pcd_points=pcd_points,
               has_mask=int(masked), img_masks=im if masked else torch.zeros(1, n_img, dtype=torch.bool),
               pcd_masks=pm if masked else torch.zeros(1, n_pcd, dtype=torch.bool), n_head=H, blocks=np.asarray(blocks),
               use_embedding=int(use_embedding), hidden=hidden, out_dim=out_dim, img_out=io, pcd_out=po)
    rec.update({"w." + n: v for n, v in net.state_dict().items()})
    save(tag, rec)


run_embedding("fusion_embed_pixels", 37, 2, 10, 0.0, False, True, 51)
run_embedding("fusion_embed_points_pi", 29, 3, 6, -2.0, True, False, 52)
run_layer("fusion_layer_self", 1, 45, 45, 64, 4, False, 53)
run_layer("fusion_layer_cross_masked", 2, 33, 51, 48, 4, True, 54)
# img_in_proj_all is nn.Linear(img_input_dim, hidden) fed with 2 * hidden columns: img_input_dim = 2 * hidden, as shipped (512 / 256)
run_fusion("fusion_module_embed", 40, 53, 64, 32, 24, 4, ["self", "cross", "self", "cross"], True, False, 55)
run_fusion("fusion_module_masked_noembed", 36, 30, 48, 24, 24, 2, ["self", "cross"], False, True, 56)
