"""Golden vectors for the denoising transformer from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_tf.py
Imports Diff-Reg-4dmatch/models/transformer.py as it lies (GeometryAttentionLayer, RepositioningTransformer), runs them on
seeded inputs on the CPU and stores inputs, weights (the modules' state_dict) and outputs as tests/golden/tf_*.npz."""
import importlib
import os
import sys

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, f"{REF}/Diff-Reg-4dmatch")
tf_mod = importlib.import_module("models.transformer")
torch.set_num_threads(1)

VOL_BNDS = [[-3.6, -2.4, 1.14], [1.093, 0.78, 2.92]]     # configs/test/4dmatch.yaml:49-50


class Cfg(dict):
    __getattr__ = dict.__getitem__


def save(tag, out):
    np.savez_compressed(os.path.join(HERE, tag + ".npz"),
                        **{k: (v.detach().numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()})
    print("wrote", tag)


def masks(g, B, L, S, kind):
    xm, sm = torch.ones(B, L, dtype=torch.bool), torch.ones(B, S, dtype=torch.bool)
    if kind == "prefix":
        xm[:, L - 5:] = False
        sm[:, S - 9:] = False
    elif kind == "arbitrary":
        xm = torch.rand(B, L, generator=g) > 0.15
        sm = torch.rand(B, S, generator=g) > 0.15
    return xm, sm


@torch.no_grad()
def run_layer(tag, B, L, S, C, H, pe_type, with_pe, kind, seed):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    layer = tf_mod.GeometryAttentionLayer(Cfg(feature_dim=C, n_head=H, pe_type=pe_type)).eval()
    for prm in layer.parameters():            # LayerNorm affine away from (1, 0), weights at a realistic scale
        prm.copy_(torch.randn(prm.shape, generator=g) * (0.3 if prm.dim() == 1 else 1.0 / prm.shape[-1] ** 0.5) + (1.0 if prm.dim() == 1 else 0.0))
    x = torch.randn(B, L, C, generator=g)
    src = torch.randn(B, S, C, generator=g)
    xm, sm = masks(g, B, L, S, kind)
    x_pe = s_pe = None
    if with_pe:
        if pe_type == "rotary":
            ax, asr = torch.rand(B, L, C // 2, generator=g) * 6.28, torch.rand(B, S, C // 2, generator=g) * 6.28
            dup = lambda a: torch.stack([a, a], -1).reshape(a.shape[0], a.shape[1], C)
            x_pe = torch.stack([dup(ax).cos(), dup(ax).sin()], -1)
            s_pe = torch.stack([dup(asr).cos(), dup(asr).sin()], -1)
        else:
            x_pe, s_pe = torch.randn(B, L, C, generator=g), torch.randn(B, S, C, generator=g)
    out = layer(x, src, x_pe, s_pe, xm if kind != "none" else None, sm if kind != "none" else None)
    rec = dict(x=x, source=src, x_mask=xm, source_mask=sm, has_mask=int(kind != "none"), has_pe=int(with_pe), pe_type=pe_type,
               n_head=H, out=out)
    if with_pe:
        rec.update(x_pe=x_pe, source_pe=s_pe)
    rec.update({"w." + k: v for k, v in layer.state_dict().items()})
    save(tag, rec)


@torch.no_grad()
def run_transformer(tag, B, N, M, C, H, pe_type, entangled, layer_types, kind, seed):
    g = torch.Generator().manual_seed(seed)
    torch.manual_seed(seed)
    cfg = Cfg(feature_dim=C, n_head=H, layer_types=layer_types, positioning_type="procrustes", pe_type=pe_type, entangled=entangled,
              vol_bnds=VOL_BNDS, voxel_size=0.04)
    net = tf_mod.RepositioningTransformer(cfg).eval()
    for prm in net.parameters():
        if prm.dim() == 1:
            prm.copy_(torch.randn(prm.shape, generator=g) * 0.3 + 1.0)
    lo, hi = torch.tensor(VOL_BNDS[0]), torch.tensor(VOL_BNDS[1])
    s_pcd = lo + (hi - lo) * torch.rand(B, N, 3, generator=g)
    t_pcd = lo + (hi - lo) * torch.rand(B, M, 3, generator=g)
    sf, tf = torch.randn(B, N, C, generator=g), torch.randn(B, M, C, generator=g)
    sm, tm = masks(g, B, N, M, kind)
    data = {}
    so, to, spe, tpe = net(sf, tf, s_pcd, t_pcd, sm, tm, data)
    rec = dict(src_feat=sf, tgt_feat=tf, s_pcd=s_pcd, t_pcd=t_pcd, src_mask=sm, tgt_mask=tm, pe_type=pe_type, entangled=int(entangled),
               n_head=H, layer_types=np.asarray(layer_types), voxel_size=0.04, vol_bnds=np.asarray(VOL_BNDS, dtype=np.float32),
               src_out=so, tgt_out=to, src_pe=spe, tgt_pe=tpe)
    rec.update({"w." + k: v for k, v in net.state_dict().items()})
    save(tag, rec)


run_layer("tf_layer_rotary_cross", 2, 37, 45, 48, 4, "rotary", True, "prefix", 11)
run_layer("tf_layer_rotary_self_nomask", 1, 50, 50, 24, 2, "rotary", True, "none", 12)
run_layer("tf_layer_sinusoidal_arbitrary", 2, 33, 29, 48, 4, "sinusoidal", True, "arbitrary", 13)
run_layer("tf_layer_nope_prefix", 1, 64, 70, 64, 4, "rotary", False, "prefix", 14)
run_transformer("tf_stack_rotary", 1, 41, 38, 48, 4, "rotary", False, ["self", "cross", "self", "cross"], "prefix", 21)
run_transformer("tf_stack_sinusoidal_entangled", 2, 30, 34, 48, 4, "sinusoidal", True, ["self", "cross"], "arbitrary", 22)
