"""Fused attention kernel (drg_attention_split16: logits in TMEM, online softmax, P.V accumulated in TMEM) through the C ABI
against an fp64 evaluation of the reference's expression (4d transformer.py:79-85 / vision3d transformer.py:127-154) and against
this library's three-kernel path (Q.K^T GEMM, drg_attn_softmax, P.V GEMM).  Tolerance 2e-5 abs on outputs of O(1) values."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref64(q, k, v, H, qm, km, scale):
    """q [B, L, C], k / v [B, S, C] -> [B, L, C] in fp64; keys masked for valid queries only."""
    B, L, C = q.shape
    S = k.shape[1]
    d = C // H
    qh, kh, vh = (t.double().view(B, -1, H, d) for t in (q, k, v))
    a = torch.einsum("nlhd,nshd->nlsh", qh, kh)
    if km is not None:
        qv = qm if qm is not None else torch.ones(B, L, dtype=torch.bool)
        a = a.masked_fill(qv[:, :, None, None] & ~km[:, None, :, None], float("-inf"))
    a = torch.softmax(a * scale, dim=2)
    return torch.einsum("nlsh,nshd->nlhd", a, vh).reshape(B, L, C)


def _run(q, k, v, H, qm, km, scale, nsplit=0):
    from diffreg_b200 import ops
    d = q.shape[-1] // H
    c = lambda t: None if t is None else t.cuda()
    q16 = ops.prep_heads(q.cuda(), H, 0)
    k16 = ops.prep_heads(k.cuda(), H, 1)
    out = ops.attention(q16, k16, v.cuda(), H, c(qm), c(km), scale, d, nsplit=nsplit)
    torch.cuda.synchronize()
    return out.cpu()


CASES = [  # B, H, L, S, d, masks
    (1, 1, 128, 64, 64, "none"),
    (1, 2, 128, 128, 64, "none"),
    (2, 4, 200, 333, 64, "prefix"),
    (1, 4, 257, 190, 132, "arbitrary"),
    (2, 3, 70, 1030, 24, "keys_only"),
    (1, 4, 1024, 1100, 132, "prefix"),
    (1, 2, 300, 700, 176, "none"),
]


@pytest.mark.parametrize("B,H,L,S,d,kind", CASES)
def test_attention_against_fp64(B, H, L, S, d, kind):
    g = torch.Generator().manual_seed(1000 + L + S + d)
    C = H * d
    q, k, v = (torch.randn(B, n, C, generator=g) * sc for n, sc in ((L, 1.3), (S, 1.1), (S, 2.0)))
    v = v + 0.5
    qm = km = None
    if kind == "prefix":
        qm, km = torch.ones(B, L, dtype=torch.bool), torch.ones(B, S, dtype=torch.bool)
        qm[:, L - 7:] = False
        km[:, S - 40:] = False
    elif kind == "arbitrary":
        qm, km = torch.rand(B, L, generator=g) > 0.2, torch.rand(B, S, generator=g) > 0.2
    elif kind == "keys_only":
        km = torch.rand(B, S, generator=g) > 0.3
    scale = 1.0 / math.sqrt(d)
    ref = _ref64(q, k, v, H, qm, km, scale)
    out = _run(q, k, v, H, qm, km, scale)
    assert out.shape == ref.shape
    assert not torch.isnan(out).any()
    err = (out.double() - ref).abs().max().item()
    assert err <= 2e-5, err


def test_attention_sharp_rows_move_the_reference():
    """Logits spread over +-60 (log2 units well beyond the lazy threshold) with the row maxima at the END of the key range: the
    running reference moves several times and the accumulator is rescaled in TMEM."""
    g = torch.Generator().manual_seed(5)
    B, H, L, S, d = 1, 2, 256, 1536, 64                         # 24 key tiles: the reference also moves after chunks of O were drained
    q, k, v = torch.randn(B, L, H * d, generator=g), torch.randn(B, S, H * d, generator=g), torch.randn(B, S, H * d, generator=g)
    k = k * torch.linspace(0.2, 6.0, S).view(1, S, 1)          # later keys have larger norms: the row maximum keeps growing
    scale = 1.0
    ref = _ref64(q, k, v, H, None, None, scale)
    out = _run(q, k, v, H, None, None, scale)
    err = (out.double() - ref).abs().max().item()
    assert err <= 5e-5, err


def test_attention_nan_for_a_valid_query_without_valid_keys():
    g = torch.Generator().manual_seed(6)
    B, H, L, S, d = 2, 2, 130, 150, 64
    q, k, v = torch.randn(B, L, H * d, generator=g), torch.randn(B, S, H * d, generator=g), torch.randn(B, S, H * d, generator=g)
    qm, km = torch.rand(B, L, generator=g) > 0.3, torch.rand(B, S, generator=g) > 0.3
    km[1] = False                                   # batch 1: no valid key -> NaN rows for its valid queries only
    ref = _ref64(q, k, v, H, qm, km, 0.125)
    out = _run(q, k, v, H, qm, km, 0.125)
    assert torch.equal(torch.isnan(out), torch.isnan(ref))
    ok = ~torch.isnan(ref)
    assert (out.double()[ok] - ref[ok]).abs().max().item() <= 2e-5


def test_attention_matches_the_three_kernel_path():
    from diffreg_b200 import ops
    g = torch.Generator().manual_seed(7)
    B, H, L, S, d = 1, 4, 384, 520, 132
    q, k, v = (torch.randn(B, n, H * d, generator=g).cuda() for n in (L, S, S))
    km = (torch.rand(B, S, generator=g) > 0.1).cuda()
    qm = (torch.rand(B, L, generator=g) > 0.1).cuda()
    scale = 1.0 / math.sqrt(d)
    q16, k16 = ops.prep_heads(q, H, 0), ops.prep_heads(k, H, 1)
    fused = ops.attention(q16, k16, v, H, qm, km, scale, d)
    logits = ops.gemm_nt(q16, k16, split3=True, K=d)
    p16 = ops.attn_softmax(logits, H, qm, km, scale)
    vt = v.view(B, S, H, d).permute(0, 2, 3, 1).contiguous().view(B * H, d, S)
    o = ops.gemm_nt(p16, ops.prep_operand(vt, 1.0, True, 1), split3=True, K=S)
    three = o.view(B, H, L, d).permute(0, 2, 1, 3).reshape(B, L, H * d)
    assert (fused - three).abs().max().item() <= 2e-5


def test_attention_refuses_heads_it_cannot_hold():
    from diffreg_b200 import ops
    from diffreg_b200._lib import DiffRegLibraryError
    q = torch.randn(1, 64, 200).cuda()
    q16, k16 = ops.prep_heads(q, 1, 0), ops.prep_heads(q, 1, 1)
    with pytest.raises(DiffRegLibraryError):
        ops.attention(q16, k16, q, 1, None, None, 1.0, 200)


def test_attention_long_rows_accumulate_in_chunks():
    """4096 keys of comparable weight per row (a flat attention: every key contributes) -- the case where the tensor core's
    truncating accumulator would bias O by ~2^-17 relative: chunks of 8 key tiles are summed with round-to-nearest adds instead."""
    g = torch.Generator().manual_seed(8)
    B, H, L, S, d = 1, 2, 256, 4096, 132
    q, k = torch.randn(B, L, H * d, generator=g) * 0.05, torch.randn(B, S, H * d, generator=g)
    v = torch.rand(B, S, H * d, generator=g) + 1.0            # all positive: truncation errors cannot cancel
    scale = 1.0 / math.sqrt(d)
    ref = _ref64(q, k, v, H, None, None, scale)
    out = _run(q, k, v, H, None, None, scale)
    err = (out.double() - ref).abs().max().item()
    assert err <= 1e-5, err                                    # measured 6.9e-6 on values ~1.5 (one chunk = 96 truncating MMAs); unchunked ~5e-5


@pytest.mark.parametrize("nsplit", [1, 2, 3, 8])
def test_attention_with_the_keys_split_over_several_ctas(nsplit):
    """Small grids share the keys of a (query tile, head) among several CTAs (grid.z) and combine the partial results: same
    result whatever the split, masks and NaN rows included."""
    g = torch.Generator().manual_seed(9)
    B, H, L, S, d = 2, 2, 200, 1500, 64
    q, k, v = torch.randn(B, L, H * d, generator=g), torch.randn(B, S, H * d, generator=g) * 1.5, torch.randn(B, S, H * d, generator=g)
    qm, km = torch.rand(B, L, generator=g) > 0.2, torch.rand(B, S, generator=g) > 0.2
    km[1, :1100] = False                     # batch 1: the first splits see masked keys only
    scale = 1.0 / math.sqrt(d)
    ref = _ref64(q, k, v, H, qm, km, scale)
    out = _run(q, k, v, H, qm, km, scale, nsplit=nsplit)
    assert torch.equal(torch.isnan(out), torch.isnan(ref))
    ok = ~torch.isnan(ref)
    assert (out.double()[ok] - ref[ok]).abs().max().item() <= 2e-5
    km[0] = False                            # batch 0: no valid key at all -> NaN for its valid queries in every split
    ref = _ref64(q, k, v, H, qm, km, scale)
    out = _run(q, k, v, H, qm, km, scale, nsplit=nsplit)
    assert torch.equal(torch.isnan(out), torch.isnan(ref))
