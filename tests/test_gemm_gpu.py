"""tcgen05 similarity GEMM (drg_gemm_nt_tf32) and operand preparation (drg_prep_operand) through
the C ABI against fp64 references.  Tolerances: plain tf32 keeps 10 mantissa bits per operand
(|err| <= 2^-10 * sum|a||b| worst case; we assert 8e-3 * sqrt(K) * scale); the 3xTF32 operand
split must be fp32-accurate: 2e-6 * sqrt(K) * scale."""
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
    assert A3.shape == (b, n, 3 * k)
    C3 = ops.gemm_nt(A3, B3, alpha=0.25).cpu()
    # the tensor core truncates its fp32 accumulator once per K=8 step: |err| <= steps * 2^-24 * max|C|
    assert (C3.double() - ref).abs().max().item() <= (3 * k / 8 + 8) * 2.0 ** -24 * ref.abs().max().item() + 1e-7
    # shared-tile kernel on the same operands (falls back to the generic one when k is not a multiple of 32)
    C3s = ops.gemm_nt(A3, B3, alpha=0.25, split3=True).cpu()
    assert (C3s.double() - ref).abs().max().item() <= (3 * k / 8 + 8) * 2.0 ** -24 * ref.abs().max().item() + 1e-7
    assert (C3s - C3).abs().max().item() <= 4 * (3 * k / 8 + 8) * 2.0 ** -24 * ref.abs().max().item() + 1e-7


def test_gemm_full_size_property():
    """4096 x 4096 x 256 (headline shape): linearity in alpha and agreement with torch fp32 on sampled rows."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    A = torch.randn(1, 4096, 256, generator=g, device="cuda") / 16
    B = torch.randn(1, 4096, 256, generator=g, device="cuda") / 16
    A3 = ops.prep_operand(A, 1.0, True, 0)
    B3 = ops.prep_operand(B, 1.0, True, 1)
    C = ops.gemm_nt(A3, B3)
    C2 = ops.gemm_nt(A3, B3, alpha=2.0)
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
    hi, lo, hi2 = out3[..., :24], out3[..., 24:48], out3[..., 48:]        # pattern 1 = [hi | lo | hi]
    out3a = ops.prep_operand(x.cuda(), scale, True, 0, pe=pe.cuda(), pe_type="rotary").cpu()          # pattern 0 = [lo | hi | hi]
    assert torch.equal(out3a[..., :24], lo) and torch.equal(out3a[..., 24:48], hi) and torch.equal(out3a[..., 48:], hi)
    assert torch.equal(hi, hi2)
    assert (hi + lo - out.cpu()).abs().max() <= 2.0 ** -21 * out.abs().max().item()   # hi + lo carries ~21 mantissa bits
    assert ((lo.view(torch.int32) & 0x1FFF) == 0).all()
    assert ((hi.view(torch.int32) & 0x1FFF) == 0).all()          # hi is a tf32 value
    # sinusoidal = additive
    pe2 = torch.randn(2, 50, 24, generator=g)
    out_s, emb_s = ops.prep_operand(x.cuda(), 1.0, False, 0, pe=pe2.cuda(), pe_type="sinusoidal", want_embedded=True)
    assert torch.equal(emb_s.cpu(), x + pe2)


def test_gemm_rejects_bad_k():
    import diffreg_b200
    with pytest.raises(diffreg_b200._lib.DiffRegLibraryError):
        _ops().gemm_nt(torch.zeros(1, 8, 6, device="cuda"), torch.zeros(1, 8, 6, device="cuda"))
