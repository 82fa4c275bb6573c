"""Correspondence extraction (drg_match_count / drg_match_write) against the reference's golden
vectors and the oracle.  Index work: bit-exact."""
import pytest
import torch

from oracle import diffreg_oracle as O
from helpers import load, names

pytestmark = pytest.mark.gpu


def _ops():
    import diffreg_b200
    return diffreg_b200.ops


@pytest.mark.parametrize("name", names("getmatch_"))
def test_get_match_golden(name):
    g = load(name)
    idx, mconf, mask = _ops().get_match(g["conf"].cuda(), float(g["thr"]), bool(g["mutual"]))
    assert torch.equal(idx.cpu(), g["index"]) and torch.equal(mconf.cpu(), g["mconf"]) and torch.equal(mask.cpu(), g["mask"])


@pytest.mark.parametrize("name", names("mts_"))
def test_top1_select_golden(name):
    g = load(name)
    thr = float(g["threshold"]) if bool(g["has_threshold"]) else None
    r, c, s = _ops().top1_select(g["score"].cuda(), True, thr, bool(g["mutual"]))
    assert torch.equal(r.cpu(), g["rows"]) and torch.equal(c.cpu(), g["cols"]) and torch.equal(s.cpu(), g["scores"])


@pytest.mark.parametrize("name", names("mtsk_"))
def test_mutual_topk_select_k_gt_1_golden(name):
    """mutual_topk_select with k = 2, 3 against the reference's outputs (largest with a threshold; smallest as a bool matrix)."""
    import diffreg_b200
    g = load(name)
    k, mutual = int(g["k"]), bool(g["mutual"])
    r, c, s = diffreg_b200.mutual_topk_select(g["score"].cuda(), k, largest=True, threshold=float(g["threshold"]), mutual=mutual)
    assert torch.equal(r.cpu(), g["rows"]) and torch.equal(c.cpu(), g["cols"]) and torch.equal(s.cpu(), g["scores"])
    cm = diffreg_b200.mutual_topk_select(g["score"].cuda(), k, largest=False, threshold=None, mutual=mutual, reduce_result=False)
    assert cm.dtype == torch.bool and torch.equal(cm.cpu(), g["corr_smallest"])


@pytest.mark.parametrize("name", names("bmts_"))
def test_batch_mutual_topk_select_golden(name):
    """batch_mutual_topk_select (k = 2, masks, threshold: the 2D-3D fine matching's call) against the reference's outputs."""
    import diffreg_b200
    g = load(name)
    b, r, c, s = diffreg_b200.batch_mutual_topk_select(g["score"].cuda(), int(g["k"]), g["row_masks"].cuda(), g["col_masks"].cuda(),
                                                       largest=True, threshold=float(g["threshold"]), mutual=bool(g["mutual"]))
    assert torch.equal(b.cpu(), g["batch"]) and torch.equal(r.cpu(), g["rows"]) and torch.equal(c.cpu(), g["cols"])
    assert torch.equal(s.cpu(), g["scores"])


@pytest.mark.parametrize("B,N,M,k", [(1, 64, 64, 2), (7, 33, 65, 3), (2, 300, 17, 8), (1, 5, 900, 4)])
@pytest.mark.parametrize("mutual", [True, False])
def test_batch_topk_select_vs_oracle(B, N, M, k, mutual):
    import diffreg_b200
    g = torch.Generator().manual_seed(B * 100 + N + M + k)
    score = torch.rand(B, N, M, generator=g)
    rm = torch.rand(B, N, generator=g) > 0.2
    cm = torch.rand(B, M, generator=g) > 0.2
    want = O.batch_mutual_topk_select(score, k, rm, cm, True, 0.4, mutual)
    got = diffreg_b200.batch_mutual_topk_select(score.cuda(), k, rm.cuda(), cm.cuda(), True, 0.4, mutual)
    for a, w_ in zip(got, want):
        assert torch.equal(a.cpu(), w_)


@pytest.mark.parametrize("B,N,M", [(1, 1, 1), (2, 33, 65), (1, 257, 1030), (3, 128, 1024), (1, 1000, 37), (1, 2048, 2050)])
@pytest.mark.parametrize("mutual", [True, False])
def test_get_match_vs_oracle(B, N, M, mutual):
    g = torch.Generator().manual_seed(N + M)
    conf = torch.rand(B, N, M, generator=g)
    conf[:, N // 2] = 0.0                      # an all-zero row (padded rows look like this)
    if N > 2 and M > 2:
        conf[0, 1, 1] = conf[0, 1, 2] = 2.0    # a tie for the row maximum
    thr = 0.2 if mutual else 0.97
    ref_idx, ref_conf, ref_mask = O.get_match(conf, thr, mutual)
    idx, mconf, mask = _ops().get_match(conf.cuda(), thr, mutual)
    assert torch.equal(idx.cpu(), ref_idx) and torch.equal(mconf.cpu(), ref_conf) and torch.equal(mask.cpu(), ref_mask)


@pytest.mark.parametrize("N,M", [(1, 1), (40, 70), (513, 1025), (1200, 300), (2048, 2048)])
@pytest.mark.parametrize("mutual", [True, False])
@pytest.mark.parametrize("thr", [None, 0.999])
def test_top1_select_vs_oracle(N, M, mutual, thr):
    g = torch.Generator().manual_seed(N * 3 + M)
    s = torch.rand(N, M, generator=g)
    r0, c0, v0 = O.mutual_topk_select(s, 1, True, thr, mutual)
    r, c, v = _ops().top1_select(s.cuda(), True, thr, mutual)
    assert torch.equal(r.cpu(), r0) and torch.equal(c.cpu(), c0) and torch.equal(v.cpu(), v0)


def test_top1_select_smallest():
    g = torch.Generator().manual_seed(4)
    s = torch.rand(90, 50, generator=g)
    r0, c0, v0 = O.mutual_topk_select(s, 1, False, 0.5, True)
    r, c, v = _ops().top1_select(s.cuda(), False, 0.5, True)
    assert torch.equal(r.cpu(), r0) and torch.equal(c.cpu(), c0) and torch.equal(v.cpu(), v0)


def test_full_size_mutual_property():
    """4096 x 4096: every mutual match is the arg-max of its row and of its column; count matches torch."""
    conf = torch.rand(1, 4096, 4096, generator=torch.Generator(device="cuda").manual_seed(1), device="cuda")
    idx, mconf, _ = _ops().get_match(conf, 0.2, True, want_mask=False)
    rows, cols = idx[:, 1], idx[:, 2]
    assert torch.equal(conf[0].max(dim=1)[0][rows], mconf) and torch.equal(conf[0].max(dim=0)[0][cols], mconf)
    assert torch.equal(conf[0, rows, cols], mconf) and bool((mconf > 0.2).all())
    ref = O.get_match(conf, 0.2, True)[0]
    assert torch.equal(idx, ref)


@pytest.mark.parametrize("B,N,M,kind", [(1, 64, 48, "full"), (2, 130, 260, "prefix"), (1, 1000, 1024, "arbitrary"), (1, 300, 4096, "full")])
def test_fused_bests_of_the_final_pass(B, N, M, kind):
    """The row / column bests written by the Sinkhorn final pass (DDIM and CONF modes) give the same mutual top-1
    matches as the stand-alone extraction on the stored confidence matrix."""
    ops = _ops()
    gen = torch.Generator().manual_seed(N + 3 * M)
    s = torch.randn(B, N, M, generator=gen) * 3.0
    sm = torch.ones(B, N, dtype=torch.bool)
    tm = torch.ones(B, M, dtype=torch.bool)
    if kind == "prefix":
        sm[:, N - 7:] = False
        tm[:, M - 20:] = False
    elif kind == "arbitrary":
        sm = torch.rand(B, N, generator=gen) > 0.1
        tm = torch.rand(B, M, generator=gen) > 0.1
    alpha = torch.tensor(1.0)
    x_t = torch.randn(B, N, M, generator=gen)
    args = (s.cuda(), alpha.cuda(), 3, sm.cuda(), tm.cuda())
    conf, rb, cb = ops.sinkhorn(*args, out_mode="conf", apply_mask=True, want_best=True)
    xn, conf2, rb2, cb2 = ops.sinkhorn(*args, out_mode="ddim", apply_mask=True, x_t=x_t.cuda(), k_x0=0.7, k_xt=0.1, sigma=0.0,
                                       want_conf=True, want_best=True)
    assert torch.equal(conf, conf2) and torch.equal(rb, rb2) and torch.equal(cb, cb2)
    for thr in (None, 0.05):
        idx, vals = ops.match_from_best(rb, cb, M, thr)
        ref_idx, ref_vals, _ = ops._match(conf, 1, True, thr, True, False)
        assert torch.equal(idx, ref_idx) and torch.equal(vals, ref_vals)
    # with the matcher's threshold as the tracking floor the pass skips the arg-max bookkeeping almost everywhere and the
    # matches at that (or any higher) threshold are unchanged; rows / columns without an entry above the floor keep key 0
    for floor in (0.05, 0.2):
        _, rbf, cbf = ops.sinkhorn(*args, out_mode="conf", apply_mask=True, want_best=True, best_floor=floor)
        for thr in (floor, 0.5):
            idx_f, vals_f = ops.match_from_best(rbf, cbf, M, thr)
            idx_a, vals_a = ops.match_from_best(rb, cb, M, thr)
            assert torch.equal(idx_f, idx_a) and torch.equal(vals_f, vals_a)
        row_has = (conf > floor).any(dim=2)
        col_has = (conf > floor).any(dim=1)
        assert torch.equal(rbf != 0, row_has) and torch.equal(cbf != 0, col_has)
        assert torch.equal(rbf[row_has], rb[row_has]) and torch.equal(cbf[col_has], cb[col_has])
    # and against the oracle's get_match on the oracle's own confidence matrix (no exact ties in random data)
    filled = s.masked_fill(~O.pair_mask(sm, tm), float("-inf"))
    ref_conf = O.log_optimal_transport(filled, alpha, 3, sm, tm).exp()[:, :-1, :-1]
    o_idx, _, _ = O.get_match(ref_conf, 0.05, True)
    idx, _ = ops.match_from_best(rb, cb, M, 0.05)
    assert torch.equal(idx.cpu(), o_idx)
