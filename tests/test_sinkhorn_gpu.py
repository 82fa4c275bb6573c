"""CUDA Sinkhorn / dual-softmax (drg_sinkhorn, drg_dual_softmax through the C ABI) against the
oracle and the reference's golden vectors.  Tolerance: 1e-4 abs on the log matrix and on the
confidences (BASELINE.json north_star), fp32."""
import pytest
import torch

from oracle import diffreg_oracle as O
from helpers import TOL_LOG, finite_close, load, names

pytestmark = pytest.mark.gpu


def _ops():
    import diffreg_b200
    return diffreg_b200.ops


@pytest.mark.parametrize("name", [n for n in names("lot_") if n != "lot_fp64_state"])
def test_log_full_against_reference_golden(name):
    g = load(name)
    dev = "cuda"
    out = _ops().sinkhorn(g["scores"].to(dev), torch.tensor(float(g["alpha"]), device=dev), int(g["iters"]),
                          g["src_mask"].to(dev), g["tgt_mask"].to(dev), out_mode="log_full")
    tol = TOL_LOG          # (also for the 100-iteration fixture: measured 1.6e-6 against an fp64 evaluation, tools/lot_error_probe.py)
    ok, err = finite_close(out.cpu(), g["out"], tol)
    assert ok, err


def _masks(B, N, M, kind, gen):
    sm = torch.ones(B, N, dtype=torch.bool)
    tm = torch.ones(B, M, dtype=torch.bool)
    if kind == "prefix":
        for b in range(B):
            sm[b, int(torch.randint(max(1, N // 2), N + 1, (1,), generator=gen)):] = False
            tm[b, int(torch.randint(max(1, M // 2), M + 1, (1,), generator=gen)):] = False
    elif kind == "arbitrary":
        sm = torch.rand(B, N, generator=gen) > 0.1
        tm = torch.rand(B, M, generator=gen) > 0.1
        sm[:, 0] = True
        tm[:, 0] = True
    return sm, tm


SHAPES = [
    (1, 1, 1, "full"), (1, 5, 3, "full"), (2, 17, 1023, "prefix"), (1, 1024, 1024, "full"), (3, 200, 1025, "arbitrary"),
    (1, 333, 1530, "arbitrary"), (1, 257, 2048, "prefix"), (2, 100, 2050, "prefix"), (1, 130, 4096, "full"),
    (1, 64, 4100, "arbitrary"), (1, 40, 8192, "full"), (1, 20, 8195, "prefix"), (1, 9, 16384, "full"), (1, 600, 37, "arbitrary"),
    (16, 96, 160, "prefix"),
]


@pytest.mark.parametrize("B,N,M,kind", SHAPES)
def test_sinkhorn_vs_oracle(B, N, M, kind):
    gen = torch.Generator().manual_seed(B * 1000003 + N * 1009 + M)
    sm, tm = _masks(B, N, M, kind, gen)
    s = torch.randn(B, N, M, generator=gen) * 3.0
    filled = s.masked_fill(~O.pair_mask(sm, tm), float("-inf"))
    alpha = torch.tensor(1.0)
    ref = O.log_optimal_transport(filled, alpha, 3, sm, tm)
    ops = _ops()
    dev = "cuda"
    # (1) API-parity mode: -inf already stored in the scores
    out, u, v = ops.sinkhorn(filled.to(dev), alpha.to(dev), 3, sm.to(dev), tm.to(dev), out_mode="log_full", return_potentials=True)
    ok, err = finite_close(out.cpu(), ref, TOL_LOG)
    assert ok, err
    ur, vr, _ = O.sinkhorn_potentials(filled, alpha, 3, sm, tm)
    assert (u.cpu() - ur).abs().max() <= TOL_LOG and (v.cpu() - vr).abs().max() <= TOL_LOG
    # (2) fused mask mode on UNFILLED scores must give the same matrix
    out2 = ops.sinkhorn(s.to(dev), alpha.to(dev), 3, sm.to(dev), tm.to(dev), out_mode="log_full", apply_mask=True)
    ok, err = finite_close(out2.cpu(), ref, TOL_LOG)
    assert ok, err
    # (3) confidence output = exp()[:, :-1, :-1]
    conf = ops.sinkhorn(s.to(dev), alpha.to(dev), 3, sm.to(dev), tm.to(dev), out_mode="conf", apply_mask=True)
    assert (conf.cpu() - ref.exp()[:, :-1, :-1]).abs().max() <= TOL_LOG


def test_sinkhorn_properties_full_size():
    """4096 x 4096 (BASELINE.json headline shape): size-independent properties.
    After the last column update the column marginals of exp(out) are exact: every real column
    sums to exp(norm)*... = 1/(ms+ns) * (ms+ns) -> check sum_i P_ij = nu_j / exp(norm)."""
    ops = _ops()
    dev = "cuda"
    N = M = 4096
    gen = torch.Generator(device=dev).manual_seed(7)
    s = torch.randn(1, N, M, generator=gen, device=dev)
    sm = torch.ones(1, N, dtype=torch.bool, device=dev)
    tm = torch.ones(1, M, dtype=torch.bool, device=dev)
    alpha = torch.tensor(1.0, device=dev)
    out = ops.sinkhorn(s, alpha, 3, sm, tm, out_mode="log_full")
    P = out.double().exp()
    col = P.sum(dim=1)                                  # [1, M+1]
    # log_nu - norm = 0 for real columns, log(ms) for the dustbin column
    assert (col[0, :M] - 1.0).abs().max() < 1e-4
    assert abs(col[0, M].item() - N) / N < 1e-4
    # shift invariance of the real block: adding c to the scores and to alpha shifts nothing in P
    out_shift = ops.sinkhorn(s + 2.5, alpha + 2.5, 3, sm, tm, out_mode="log_full")
    assert (out_shift - out).abs().max() < 2e-4
    # against torch on the GPU with the oracle's formula (same device, fp32)
    ref = O.log_optimal_transport(s, alpha, 3, sm, tm)
    assert (ref - out).abs().max() < TOL_LOG


def test_shift_argument():
    ops = _ops()
    dev = "cuda"
    gen = torch.Generator().manual_seed(3)
    s = torch.randn(1, 70, 90, generator=gen)
    sm = torch.ones(1, 70, dtype=torch.bool)
    tm = torch.ones(1, 90, dtype=torch.bool)
    alpha = torch.tensor(1.0)
    shift = s.min()
    ref = O.log_optimal_transport(s - shift, alpha, 3, sm, tm)
    out = ops.sinkhorn(s.to(dev), alpha.to(dev), 3, sm.to(dev), tm.to(dev), shift=shift.reshape(1).to(dev))
    assert (out.cpu() - ref).abs().max() <= TOL_LOG


@pytest.mark.parametrize("B,N,M", [(1, 8, 8), (3, 28, 22), (16, 200, 260), (2, 1000, 2048), (1, 50, 4100)])
def test_dual_softmax(B, N, M):
    gen = torch.Generator().manual_seed(N * 7 + M)
    sm, tm = _masks(B, N, M, "prefix", gen)
    sim = torch.randn(B, N, M, generator=gen) * 0.3
    p = O.MatchingParams(src_proj_weight=None, match_type="dual_softmax", temperature=0.1)
    ref = O.confidence_from_similarity(p, sim, sm, tm)
    out = _ops().dual_softmax(sim.cuda(), sm.cuda(), tm.cuda(), 0.1)
    assert (out.cpu() - ref).abs().max() <= TOL_LOG


def test_ddim_output_mode():
    ops = _ops()
    dev = "cuda"
    gen = torch.Generator().manual_seed(11)
    N, M = 60, 44
    sm, tm = _masks(1, N, M, "arbitrary", gen)
    sim = torch.randn(1, N, M, generator=gen)
    x_t = torch.randn(1, N, M, generator=gen).masked_fill(~O.pair_mask(sm, tm), float("-inf"))
    noise = torch.randn(1, N, M, generator=gen)
    alpha = torch.tensor(1.0)
    ac = O.alphas_cumprod()
    t, tn = 999, 949
    r, rm1, san, c, sigma = O.ddim_coefficients(ac, t, tn)
    x0 = O.log_optimal_transport(sim.masked_fill(~O.pair_mask(sm, tm), float("-inf")), alpha, 3, sm, tm).exp()[:, :-1, :-1]
    ref = O.ddim_update(x_t, x0, ac, t, tn, noise)
    xmin = torch.full((1,), float("inf"), device=dev)
    xn, conf = ops.sinkhorn(sim.to(dev), alpha.to(dev), 3, sm.to(dev), tm.to(dev), out_mode="ddim", apply_mask=True,
                            x_t=x_t.to(dev), noise=noise.to(dev), k_x0=san - c / rm1, k_xt=c * r / rm1, sigma=sigma,
                            want_conf=True, x_min=xmin)
    ok, err = finite_close(xn.cpu(), ref, TOL_LOG)
    assert ok, err
    assert (conf.cpu() - x0).abs().max() <= TOL_LOG
    valid = torch.isfinite(ref)
    assert abs(xmin.item() - ref[valid].min().item()) <= TOL_LOG


def test_unsupported_and_errors():
    import diffreg_b200
    ops = _ops()
    with pytest.raises(diffreg_b200._lib.DiffRegLibraryError):
        ops.sinkhorn(torch.zeros(1, 4, 4), torch.tensor(1.0), 3, torch.ones(1, 4, dtype=torch.bool), torch.ones(1, 4, dtype=torch.bool))
    with pytest.raises(diffreg_b200._lib.DiffRegLibraryError):
        z = torch.zeros(1, 2, 20000, device="cuda")
        ops.sinkhorn(z, torch.tensor(1.0, device="cuda"), 3, torch.ones(1, 2, dtype=torch.bool, device="cuda"),
                     torch.ones(1, 20000, dtype=torch.bool, device="cuda"))


@pytest.mark.parametrize("scale,alpha_v,iters", [(60.0, 1.0, 4), (3.0, -30.0, 3), (25.0, 2.0, 20), (3.0, 1.0, 40)])
def test_persistent_kernel_fallback_paths(scale, alpha_v, iters):
    """The register-slab kernel leaves its scaled arithmetic when the potentials move by more than 2^50 between
    iterations (huge dynamic range of the scores) or when the dustbin score is very negative, and runs the classical
    log-domain pass instead; many iterations exercise the grid-barrier counters.  Same tolerance as everywhere."""
    B, N, M = 2, 300, 1024
    gen = torch.Generator().manual_seed(int(scale * 10) + iters)
    sm, tm = _masks(B, N, M, "arbitrary", gen)
    s = torch.randn(B, N, M, generator=gen) * scale
    filled = s.masked_fill(~O.pair_mask(sm, tm), float("-inf"))
    alpha = torch.tensor(alpha_v)
    ref = O.log_optimal_transport(filled, alpha, iters, sm, tm)
    out = _ops().sinkhorn(s.cuda(), alpha.cuda(), iters, sm.cuda(), tm.cuda(), out_mode="log_full", apply_mask=True)
    # the log matrix reaches magnitudes of several hundred here: 1e-4 absolute below 100, relative 2e-6 above
    a, b = out.cpu().double(), ref.double()
    fin = torch.isfinite(b)
    assert torch.equal(torch.isfinite(a), fin)
    err = ((a[fin] - b[fin]).abs() / b[fin].abs().clamp_min(50.0)).max().item()
    assert err <= 2e-6, err
