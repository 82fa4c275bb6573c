"""tcgen05 similarity GEMM (drg_gemm_nt_tf32, drg_gemm_nt_split16) and operand preparation (drg_prep_operand) through
the C ABI against fp64 references.  Tolerances: plain tf32 keeps 10 mantissa bits per operand
(|err| <= 2^-10 * sum|a||b| worst case; we assert 8e-3 * sqrt(K) * scale); the split-operand product
(row-scaled x = fp16 hi + fp16 lo; lo.hi + hi.lo + hi.hi with fp32 accumulation) must be fp32-accurate."""
import math

import pytest
import torch

from oracle import diffreg_oracle as O

pytestmark = pytest.mark.gpu


def _ops():
    import diffreg_b200
    return diffreg_b200.ops


SHAPES = [(1, 128, 64, 32), (1, 128, 256, 256), (1, 100, 70, 36), (2, 300, 1000, 528), (1, 1024, 1024, 256),
          (3, 130, 1530, 256), (1, 333, 1531, 40), (1, 4100, 260, 768), (1, 1, 1, 4)]


@pytest.mark.parametrize("b,n,m,k", SHAPES)
def test_gemm_tf32_and_3xtf32(b, n, m, k):
    g = torch.Generator().manual_seed(n * 31 + m * 7 + k)
    A = torch.randn(b, n, k, generator=g)
    B = torch.randn(b, m, k, generator=g)
    ref = 0.25 * torch.einsum("bnk,bmk->bnm", A.double(), B.double())
    ops = _ops()
    C = ops.gemm_nt(A.cuda(), B.cuda(), alpha=0.25).cpu()
    assert C.shape == ref.shape
    assert (C.double() - ref).abs().max().item() <= 8e-3 * math.sqrt(k) * 0.25
    A3 = ops.prep_operand(A.cuda(), 1.0, True, 0)
    B3 = ops.prep_operand(B.cuda(), 1.0, True, 1)
    assert A3.shape == (b, n, ops.split_pitch(k)) and A3.dtype == torch.int16
    C3 = ops.gemm_nt(A3, B3, alpha=0.25, split3=True, K=k).cpu()
    # fp32-accurate: the dropped lo.lo term and the rounding of lo are ~2^-22 relative per product; the fp32 accumulator
    # is rounded once per MMA step (K = 16): |err| <= (steps + 8) * 2^-24 * max|C| + 2^-21 * sqrt(K) * (rms |a||b|)
    bound = (3 * k / 16 + 8) * 2.0 ** -24 * ref.abs().max().item() + 2.0 ** -21 * math.sqrt(k) * 0.25 + 1e-7
    assert (C3.double() - ref).abs().max().item() <= bound
    # against torch's own fp32 product (what the reference computes): no worse than a small multiple of its error
    ref32 = 0.25 * torch.einsum("bnk,bmk->bnm", A, B)
    err32 = (ref32.double() - ref).abs().max().item()
    assert (C3.double() - ref).abs().max().item() <= max(8 * err32, 1e-6)


def test_gemm_full_size_property():
    """4096 x 4096 x 256 (headline shape): linearity in alpha and agreement with torch fp32 on sampled rows."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(1, 4096, 256, generator=g, device="cuda") / 16
    B = torch.randn(1, 4096, 256, generator=g, device="cuda") / 16
    A3 = ops.prep_operand(A, 1.0, True, 0)
    B3 = ops.prep_operand(B, 1.0, True, 1)
    C = ops.gemm_nt(A3, B3, split3=True)
    C2 = ops.gemm_nt(A3, B3, alpha=2.0, split3=True)
    assert torch.equal(C2, 2.0 * C)
    rows = torch.tensor([0, 1, 127, 128, 2047, 4095], device="cuda")
    ref = A[0, rows].double() @ B[0].double().t()
    assert (C[0, rows].double() - ref).abs().max().item() <= 1e-6


def test_prep_operand_rotary_and_scale():
    ops = _ops()
    g = torch.Generator().manual_seed(9)
    x = torch.randn(2, 50, 24, generator=g)
    ang = torch.rand(2, 50, 12, generator=g) * 6.28
    cos = torch.stack([ang.cos(), ang.cos()], -1).reshape(2, 50, 24)
    sin = torch.stack([ang.sin(), ang.sin()], -1).reshape(2, 50, 24)
    pe = torch.stack([cos, sin], -1)
    ref_emb = O.embed_rotary(x, cos, sin)
    scale = 1.0 / 24 ** 0.5
    out, emb = ops.prep_operand(x.cuda(), scale, False, 0, pe=pe.cuda(), pe_type="rotary", want_embedded=True)
    assert torch.equal(emb.cpu(), ref_emb)
    assert (out.cpu() - ref_emb * scale).abs().max() <= 1e-7
    out3 = ops.prep_operand(x.cuda(), scale, True, 1, pe=pe.cuda(), pe_type="rotary").cpu()
    kc = ops.split_cols(24)
    assert out3.shape == (2, 50, ops.split_pitch(24)) and out3.dtype == torch.int16
    # pattern 1 = [hi | lo | tail]: fp16 halves of the ROW-SCALED values, each segment padded to kc; tail = (1 / scale, norm, 0, 0)
    tail = out3[..., 2 * kc:].contiguous().view(torch.float32)
    inv = tail[..., 0:1]
    hi = out3[..., :24].view(torch.float16).float()
    lo = out3[..., kc:kc + 24].view(torch.float16).float()
    assert (out3[..., 24:kc] == 0).all() and (out3[..., kc + 24:2 * kc] == 0).all()          # zero padding
    want = out.cpu()
    amax = want.abs().amax(-1, keepdim=True)
    assert torch.equal(inv, torch.exp2(torch.floor(torch.log2(amax)) - 14))              # row maximum scaled into [2^14, 2^15)
    assert (tail[..., 1] - want.norm(dim=-1)).abs().max() <= 1e-6 * want.norm(dim=-1).max()
    assert (tail[..., 2:] == 0).all()
    out3a = ops.prep_operand(x.cuda(), scale, True, 0, pe=pe.cuda(), pe_type="rotary").cpu()          # pattern 0 = [lo | hi | tail]
    assert torch.equal(out3a[..., :24], out3[..., kc:kc + 24]) and torch.equal(out3a[..., kc:kc + 24], out3[..., :24])
    assert torch.equal(out3a[..., 2 * kc:], out3[..., 2 * kc:])
    assert torch.equal(hi, (want / inv).half().float())                   # hi = fp16(x 2^e), round to nearest
    assert ((hi + lo) * inv - want).abs().max() <= 2.0 ** -21 * want.abs().max().item()   # hi + lo carries ~22 significant bits
    # sinusoidal = additive
    pe2 = torch.randn(2, 50, 24, generator=g)
    out_s, emb_s = ops.prep_operand(x.cuda(), 1.0, False, 0, pe=pe2.cuda(), pe_type="sinusoidal", want_embedded=True)
    assert torch.equal(emb_s.cpu(), x + pe2)


def test_gemm_rejects_bad_k():
    import diffreg_b200
    with pytest.raises(diffreg_b200._lib.DiffRegLibraryError):
        _ops().gemm_nt(torch.zeros(1, 8, 6, device="cuda"), torch.zeros(1, 8, 6, device="cuda"))


def test_split_gemm_dynamic_range():
    """Rows of wildly different magnitude (1e-20 .. 1e+20, far outside fp16's range) and an all-zero row: the per-row
    power-of-two scales of the split operands keep every entry of the product accurate relative to its row / column norms."""
    ops = _ops()
    g = torch.Generator().manual_seed(77)
    n, m, k = 200, 136, 96
    A = torch.randn(1, n, k, generator=g)
    B = torch.randn(1, m, k, generator=g)
    sa = 10.0 ** torch.linspace(-20, 20, n).view(1, n, 1)
    sb = 10.0 ** torch.linspace(15, -15, m).view(1, m, 1)
    A = A * sa
    B = B * sb
    A[0, 7] = 0.0
    C = ops.gemm_nt(ops.prep_operand(A.cuda(), 1.0, True, 0), ops.prep_operand(B.cuda(), 1.0, True, 1), split3=True, K=k).cpu()
    ref = A.double() @ B.double().transpose(1, 2)
    scale = A.double().norm(dim=-1).unsqueeze(-1) * B.double().norm(dim=-1).unsqueeze(-2)      # |a_i| |b_j|
    rel = ((C.double() - ref).abs() / scale.clamp_min(1e-300))
    rel[0, 7] = 0.0
    assert torch.isfinite(C).all()
    assert (C[0, 7] == 0).all()
    assert rel.max().item() <= 2e-6          # fp32 sgemm's own bound is ~ sqrt(K) 2^-24 ~ 6e-7 of |a||b|


def test_project_pair_split_matches_fp64():
    """Projection with the split epilogue (drg_project_split16) feeding the similarity GEMM: sim = (xs W^T)(xt W^T)^T / C."""
    ops = _ops()
    g = torch.Generator().manual_seed(78)
    for (n, m, c) in ((256, 192, 64), (160, 224, 132), (1024, 1024, 256)):
        xs = torch.randn(1, n, c, generator=g) * 3.0
        xt = torch.randn(1, m, c, generator=g) * 0.2
        W = torch.randn(c, c, generator=g) / c ** 0.5
        w16 = ops.prep_operand(W.cuda(), 1.0, True, 1)
        scale = 1.0 / c ** 0.5
        a, b, plain = ops.project_pair_split(xs.cuda(), xt.cuda(), w16, c, scale, want_plain=True)
        assert a.shape == (1, n, ops.split_pitch(c)) and b.shape == (1, m, ops.split_pitch(c))
        ys, yt = xs.double() @ W.double().t(), xt.double() @ W.double().t()
        assert (plain.cpu().double() - torch.cat([ys[0], yt[0]], 0)).abs().max().item() <= 2e-6 * ys.abs().max().item() * 4
        sim = ops.gemm_nt(a, b, split3=True, K=c).cpu()
        ref = (ys * scale) @ (yt * scale).transpose(1, 2)
        ref32 = ((xs @ W.t()) * scale) @ ((xt @ W.t()) * scale).transpose(1, 2)
        err32 = (ref32.double() - ref).abs().max().item()
        assert (sim.double() - ref).abs().max().item() <= max(8 * err32, 1e-6)


@pytest.mark.parametrize("b,n,m,k", [(1, 512, 512, 64), (1, 256, 768, 200), (2, 384, 1100, 132), (1, 1000, 2052, 96), (1, 640, 640, 256),
                                     (3, 256, 512, 72)])
def test_split_gemm_cluster_modes(b, n, m, k):
    """Shapes that take the two-CTA cluster paths (cta_group::2 pairs for 256-wide tiles, multicast B tiles for 128-wide ones:
    an even number of 128-row blocks, M >= 2 tile widths), with ragged last row / column blocks, several batches and K that is not
    a multiple of the 64-column k-step: same result as the fp64 product within the split GEMM's bound."""
    ops = _ops()
    g = torch.Generator().manual_seed(1000 * n + m + k)
    A = torch.randn(b, n, k, generator=g)
    B = torch.randn(b, m, k, generator=g)
    ref = torch.einsum("bnk,bmk->bnm", A.double(), B.double())
    C = ops.gemm_nt(ops.prep_operand(A.cuda(), 1.0, True, 0), ops.prep_operand(B.cuda(), 1.0, True, 1), split3=True, K=k).cpu()
    ref32 = torch.einsum("bnk,bmk->bnm", A, B)
    err32 = (ref32.double() - ref).abs().max().item()
    assert (C.double() - ref).abs().max().item() <= max(8 * err32, 1e-6)
