"""Denoising transformer through the C ABI: the geometry attention layer and the self / cross stack
against the reference's golden outputs (tests/golden/tf_*.npz, made by make_golden_tf.py from the unmodified reference),
against the oracle on seeded inputs at larger sizes, and against the reference module itself (oracle/_ref) at the real
4DMatch width (C = 528, 4 heads of 132).  Tolerance: 1e-4 abs on the layer outputs (LayerNorm-ed, O(1))."""
import math

import pytest
import torch

from oracle import diffreg_oracle as O
from oracle import ref_loader
from helpers import load, names

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _ops():
    from diffreg_b200 import ops
    return ops


class Cfg(dict):
    __getattr__ = dict.__getitem__


def _weights(g, prefix="w."):
    return {k[len(prefix):]: v for k, v in g.items() if k.startswith(prefix)}


def test_attn_softmax_masks_scale_and_operand():
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    B, H, L, S = 2, 3, 37, 53
    a = torch.randn(B * H, L, S, generator=g) * 4.0
    qm = torch.rand(B, L, generator=g) > 0.2
    km = torch.rand(B, S, generator=g) > 0.2
    km[1] = False                                   # batch 1: no valid key at all -> NaN rows for its valid queries
    scale = 1.0 / math.sqrt(12.0)
    ref = a.view(B, H, L, S).clone()
    ref.masked_fill_(qm[:, None, :, None] & ~km[:, None, None, :], float("-inf"))
    ref = torch.softmax(ref * scale, dim=-1).view(B * H, L, S)
    P16, P = ops.attn_softmax(a.cuda(), H, qm.cuda(), km.cuda(), scale, want_operand=True, want_probs=True)
    P = P.cpu()
    assert torch.equal(torch.isnan(P), torch.isnan(ref))
    ok = ~torch.isnan(ref)
    assert (P[ok] - ref[ok]).abs().max().item() <= 1e-6
    # the split operand: [lo | hi | tail] fp16 halves of the row-scaled probabilities
    kc = ops.split_cols(S)
    P16 = P16.cpu()
    inv = P16[..., 2 * kc:].contiguous().view(torch.float32)[..., 0:1]
    rec = (P16[..., :S].view(torch.float16).float() + P16[..., kc:kc + S].view(torch.float16).float()) * inv
    assert (rec[ok] - ref[ok]).abs().max().item() <= 2.0 ** -21
    assert (P16[..., S:kc] == 0).all() and (P16[..., kc + S:2 * kc] == 0).all()
    # no masks at all
    P2 = ops.attn_softmax(a.cuda(), H, None, None, scale, want_operand=False, want_probs=True).cpu()
    assert (P2 - torch.softmax(a * scale, dim=-1)).abs().max().item() <= 1e-6


def test_layernorm_and_residual():
    ops = _ops()
    g = torch.Generator().manual_seed(4)
    for C in (24, 528, 1056):
        x = torch.randn(3, 41, C, generator=g) * 3.0 + 1.5
        w, b, r = torch.randn(C, generator=g), torch.randn(C, generator=g), torch.randn(3, 41, C, generator=g)
        ref = torch.nn.functional.layer_norm(x.double(), (C,), w.double(), b.double(), 1e-5)
        out = ops.layernorm(x.cuda(), w.cuda(), b.cuda(), 1e-5).cpu()
        assert (out.double() - ref).abs().max().item() <= 2e-5
        out = ops.layernorm(x.cuda(), w.cuda(), b.cuda(), 1e-5, residual=r.cuda()).cpu()
        assert (out.double() - (ref + r.double())).abs().max().item() <= 2e-5


def test_prep_heads_is_head_major_with_rotary():
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    B, L, H, d = 2, 19, 3, 8
    C = H * d
    x = torch.randn(B, L, C, generator=g)
    ang = torch.rand(B, L, C // 2, generator=g) * 6.28
    dup = torch.stack([ang, ang], -1).reshape(B, L, C)
    pe = torch.stack([dup.cos(), dup.sin()], -1)
    want = O.embed_rotary(x, dup.cos(), dup.sin()).view(B, L, H, d).permute(0, 2, 1, 3).reshape(B * H, L, d)
    out = ops.prep_heads(x.cuda(), H, 1, pe=pe.cuda(), pe_type="rotary").cpu()       # pattern 1 = [hi | lo | tail]
    kc = ops.split_cols(d)
    assert out.shape == (B * H, L, ops.split_pitch(d))
    inv = out[..., 2 * kc:].contiguous().view(torch.float32)[..., 0:1]
    rec = (out[..., :d].view(torch.float16).float() + out[..., kc:kc + d].view(torch.float16).float()) * inv
    assert (rec - want).abs().max().item() <= 2.0 ** -20 * want.abs().max().item()


def _run_layer(g, dev="cuda"):
    import diffreg_b200
    C, H = g["x"].shape[-1], int(g["n_head"])
    layer = diffreg_b200.GeometryAttentionLayer(Cfg(feature_dim=C, n_head=H, pe_type=str(g["pe_type"]))).to(dev).eval()
    layer.load_state_dict(_weights(g), strict=True)
    has_mask, has_pe = bool(int(g["has_mask"])), bool(int(g["has_pe"]))
    c = lambda t: t.to(dev)
    return layer(c(g["x"]), c(g["source"]), c(g["x_pe"]) if has_pe else None, c(g["source_pe"]) if has_pe else None,
                 c(g["x_mask"]) if has_mask else None, c(g["source_mask"]) if has_mask else None)


@pytest.mark.parametrize("name", names("tf_layer_"))
def test_layer_against_the_reference_golden(name):
    g = load(name)
    out = _run_layer(g).cpu()
    assert out.shape == g["out"].shape
    assert (out - g["out"]).abs().max().item() <= TOL


def test_self_attention_shares_the_staged_input():
    """x is source (the 'self' layers): same result as passing an equal copy."""
    g = load("tf_layer_rotary_self_nomask")
    import diffreg_b200
    C, H = g["x"].shape[-1], int(g["n_head"])
    layer = diffreg_b200.GeometryAttentionLayer(Cfg(feature_dim=C, n_head=H, pe_type="rotary")).cuda().eval()
    layer.load_state_dict(_weights(g), strict=True)
    x, pe = g["x"].cuda(), g["x_pe"].cuda()
    a = layer(x, x, pe, pe, None, None)
    b = layer(x, x.clone(), pe, pe.clone(), None, None)
    assert torch.equal(a, b)


@pytest.mark.parametrize("name", names("tf_stack_"))
def test_transformer_against_the_reference_golden(name):
    import diffreg_b200
    g = load(name)
    C, H = g["src_feat"].shape[-1], int(g["n_head"])
    cfg = Cfg(feature_dim=C, n_head=H, layer_types=[str(t) for t in g["layer_types"]], positioning_type="procrustes",
              pe_type=str(g["pe_type"]), entangled=bool(int(g["entangled"])), vol_bnds=g["vol_bnds"].tolist(),
              voxel_size=float(g["voxel_size"]))
    net = diffreg_b200.RepositioningTransformer(cfg).cuda().eval()
    net.load_state_dict(_weights(g), strict=True)
    c = lambda k: g[k].cuda()
    data = {}
    so, to, spe, tpe = net(c("src_feat"), c("tgt_feat"), c("s_pcd"), c("t_pcd"), c("src_mask"), c("tgt_mask"), data)
    assert (so.cpu() - g["src_out"]).abs().max().item() <= TOL and (to.cpu() - g["tgt_out"]).abs().max().item() <= TOL
    assert (spe.cpu() - g["src_pe"]).abs().max().item() <= 1e-6 and (tpe.cpu() - g["tgt_pe"]).abs().max().item() <= 1e-6
    assert data["position_layers"] == {}


def test_layer_at_1024_tokens_against_the_oracle():
    """L = 1024 queries x S = 1100 keys, C = 256, 4 heads of 64, prefix masks, rotary code: the fp32 drop-in against the
    oracle evaluated in fp64; its error must stay within a small multiple of the fp32 oracle's own."""
    import diffreg_b200
    g = torch.Generator().manual_seed(31)
    B, L, S, C, H = 1, 1024, 1100, 256, 4
    layer = diffreg_b200.GeometryAttentionLayer(Cfg(feature_dim=C, n_head=H, pe_type="rotary")).eval()
    for prm in layer.parameters():
        with torch.no_grad():
            prm.copy_(torch.randn(prm.shape, generator=g) * (0.3 if prm.dim() == 1 else 1.0 / prm.shape[-1] ** 0.5) + (1.0 if prm.dim() == 1 else 0.0))
    w = {k: v.clone() for k, v in layer.state_dict().items()}
    x, src = torch.randn(B, L, C, generator=g), torch.randn(B, S, C, generator=g)
    xm, sm = torch.ones(B, L, dtype=torch.bool), torch.ones(B, S, dtype=torch.bool)
    xm[:, L - 40:] = False
    sm[:, S - 100:] = False
    ax, asr = torch.rand(B, L, C // 2, generator=g) * 6.28, torch.rand(B, S, C // 2, generator=g) * 6.28
    dup = lambda a: torch.stack([a, a], -1).reshape(a.shape[0], a.shape[1], C)
    x_pe, s_pe = torch.stack([dup(ax).cos(), dup(ax).sin()], -1), torch.stack([dup(asr).cos(), dup(asr).sin()], -1)
    ref64 = O.geometry_attention_layer(w, x.double(), src.double(), x_pe.double(), s_pe.double(), xm, sm, "rotary", H)
    ref32 = O.geometry_attention_layer(w, x, src, x_pe, s_pe, xm, sm, "rotary", H)
    layer = layer.cuda()
    out = layer(x.cuda(), src.cuda(), x_pe.cuda(), s_pe.cuda(), xm.cuda(), sm.cuda()).cpu()
    err = (out.double() - ref64).abs().max().item()
    err32 = (ref32.double() - ref64).abs().max().item()
    assert err <= TOL and err <= max(8 * err32, 2e-5), (err, err32)


@pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present (/root/reference or oracle/_ref)")
def test_denoising_transformer_against_the_reference_module_at_528():
    """The denoising transformer of pipeline.py:84-85 (six layers, C = 528, 4 heads of 132, rotary, not entangled) with the
    reference's own module run on the same GPU (TF32 off) as the yardstick, same weights."""
    import diffreg_b200
    ref = ref_loader.load_flavour("4d")
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        g = torch.Generator().manual_seed(41)
        torch.manual_seed(41)
        B, N, M, C, H = 1, 300, 340, 528, 4
        bnds = [[-3.6, -2.4, 1.14], [1.093, 0.78, 2.92]]
        cfg = Cfg(feature_dim=C, n_head=H, layer_types=['self', 'cross', 'self', 'cross', 'self', 'cross'], positioning_type="procrustes",
                  pe_type="rotary", entangled=False, vol_bnds=bnds, voxel_size=0.04)
        rnet = ref.transformer.RepositioningTransformer(cfg).cuda().eval()
        onet = diffreg_b200.RepositioningTransformer(cfg).cuda().eval()
        onet.load_state_dict(rnet.state_dict(), strict=True)
        lo, hi = torch.tensor(bnds[0]), torch.tensor(bnds[1])
        s_pcd = (lo + (hi - lo) * torch.rand(B, N, 3, generator=g)).cuda()
        t_pcd = (lo + (hi - lo) * torch.rand(B, M, 3, generator=g)).cuda()
        sf, tf = torch.randn(B, N, C, generator=g).cuda(), torch.randn(B, M, C, generator=g).cuda()
        sm, tm = torch.ones(B, N, dtype=torch.bool).cuda(), torch.ones(B, M, dtype=torch.bool).cuda()
        sm[:, N - 11:] = False
        tm[:, M - 30:] = False
        with torch.no_grad():
            rs, rt, _, _ = rnet(sf, tf, s_pcd, t_pcd, sm, tm, {})
        os_, ot, _, _ = onet(sf, tf, s_pcd, t_pcd, sm, tm, {})
        assert (os_ - rs).abs().max().item() <= TOL and (ot - rt).abs().max().item() <= TOL
    finally:
        ref_loader.unload()


def test_denoising_transformer_graph_replay_matches_eager():
    """The self / cross stack replays as one CUDA graph from the second call with the same shapes on (graphs.py): identical
    results on fresh inputs; stacks with positioning layers never take that path."""
    import diffreg_b200
    g = torch.Generator().manual_seed(51)
    torch.manual_seed(51)
    B, N, M, C, H = 1, 150, 170, 48, 4
    bnds = [[-3.6, -2.4, 1.14], [1.093, 0.78, 2.92]]
    cfg = Cfg(feature_dim=C, n_head=H, layer_types=['self', 'cross'], positioning_type="procrustes", pe_type="rotary", entangled=False,
              vol_bnds=bnds, voxel_size=0.04)
    net = diffreg_b200.RepositioningTransformer(cfg).cuda().eval()
    lo, hi = torch.tensor(bnds[0]), torch.tensor(bnds[1])
    sm, tm = torch.ones(B, N, dtype=torch.bool).cuda(), torch.ones(B, M, dtype=torch.bool).cuda()
    sm[:, N - 7:] = False
    for call in range(3):
        s_pcd = (lo + (hi - lo) * torch.rand(B, N, 3, generator=g)).cuda()
        t_pcd = (lo + (hi - lo) * torch.rand(B, M, 3, generator=g)).cuda()
        sf, tf = torch.randn(B, N, C, generator=g).cuda(), torch.randn(B, M, C, generator=g).cuda()
        net.graph_replay = True
        d1 = {}
        a = net(sf, tf, s_pcd, t_pcd, sm, tm, d1)
        net.graph_replay = False
        b = net(sf, tf, s_pcd, t_pcd, sm, tm, {})
        assert d1["position_layers"] == {}
        assert all(torch.equal(x, y) for x, y in zip(a, b)), call
    assert net._graphs.replays == 2


def test_layernorm_stages_the_next_linear_operand():
    """layernorm(stage=True) also writes its result as the left split operand; prep_operand hands that out instead of staging again,
    and it is bit-identical to a separate staging of the fp32 result."""
    ops = _ops()
    g = torch.Generator().manual_seed(6)
    for C in (24, 256, 528):
        x, r = torch.randn(2, 37, C, generator=g).cuda() * 3.0, torch.randn(2, 37, C, generator=g).cuda()
        w, b = torch.randn(C, generator=g).cuda(), torch.randn(C, generator=g).cuda()
        for pre in (False, True):
            out = ops.layernorm(x, w, b, 1e-5, residual=r, pre_add=pre, stage=True)
            plain = ops.layernorm(x, w, b, 1e-5, residual=r, pre_add=pre)
            assert (out - plain).abs().max().item() <= 2e-6            # (the two instantiations may contract y * w + b differently)
            fused = ops.prep_operand(out, 1.0, True, 0)
            assert fused is out._drg_a16
            again = ops.prep_operand(out.clone(), 1.0, True, 0)          # a separate staging of the same fp32 values
            kc = ops.split_cols(C)
            assert torch.equal(fused[..., :2 * kc + 2], again[..., :2 * kc + 2])     # hi / lo halves, padding and 1 / scale: bit for bit
            nf = fused[..., 2 * kc:].contiguous().view(torch.float32)[..., 1]
            na = again[..., 2 * kc:].contiguous().view(torch.float32)[..., 1]
            assert (nf - na).abs().max().item() <= 1e-5 * na.abs().max().item()      # the row norm: another summation order
