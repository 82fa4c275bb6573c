"""Training path: the CUDA backward of log_optimal_transport (drg_sinkhorn_backward) and the differentiable
Sinkhorn branch of Matching.forward, through the C ABI, against the reference's autograd (golden gradients, tests/golden/lotb_*.npz,
made by make_golden_lotb.py from the unmodified reference) and against the oracle's written-out backward in fp64.
Tolerance: 2e-5 relative to the largest gradient entry (fp32 exps of sums of three potentials)."""
import pytest
import torch

from oracle import diffreg_oracle as O
from helpers import load, names

pytestmark = pytest.mark.gpu
REL = 2e-5


@pytest.mark.parametrize("name", [n for n in names("lotb_") if "matching" not in n])
def test_sinkhorn_backward_against_the_reference_autograd(name):
    import diffreg_b200
    g = load(name)
    scores = g["scores"].cuda().requires_grad_()
    alpha = torch.tensor(float(g["alpha"]), device="cuda", requires_grad=True)
    out = diffreg_b200.log_optimal_transport(scores, alpha, int(g["iters"]), g["src_mask"].cuda(), g["tgt_mask"].cuda())
    assert out.requires_grad
    fin_o = torch.isfinite(g["out"])
    assert (out.detach().cpu()[fin_o] - g["out"][fin_o]).abs().max().item() <= 1e-4
    (out * g["grad_out"].cuda()).sum().backward()
    fin = torch.isfinite(g["scores"])
    scale = g["grad_scores"][fin].abs().max().item()
    assert (scores.grad.cpu()[fin] - g["grad_scores"][fin]).abs().max().item() <= REL * scale
    assert abs(alpha.grad.item() - float(g["grad_alpha"])) <= 1e-4 * max(1.0, abs(float(g["grad_alpha"])))


@pytest.mark.parametrize("B,N,M,iters", [(1, 700, 900, 3), (3, 257, 130, 3), (1, 64, 1500, 2), (2, 100, 100, 7)])
def test_sinkhorn_backward_against_the_oracle_in_fp64(B, N, M, iters):
    from diffreg_b200 import ops
    g = torch.Generator().manual_seed(N + M + iters)
    scores = torch.randn(B, N, M, generator=g) * 3.0
    sm, tm = torch.rand(B, N, generator=g) > 0.1, torch.rand(B, M, generator=g) > 0.1
    scores = scores.masked_fill(~(sm[..., None] * tm[:, None]).bool(), float("-inf"))
    alpha = torch.tensor(0.7)
    G = torch.randn(B, N + 1, M + 1, generator=g)
    gs64, ga64 = O.log_optimal_transport_backward(scores.double(), alpha.double(), iters, sm, tm, G.double())
    gs, ga = ops.sinkhorn_backward(scores.cuda(), alpha.cuda(), iters, sm.cuda(), tm.cuda(), G.cuda())
    fin = torch.isfinite(scores)
    scale = gs64[fin].abs().max().item()
    assert (gs.cpu().double()[fin] - gs64[fin]).abs().max().item() <= REL * scale
    assert abs(ga.item() - ga64.item()) <= 1e-4 * max(1.0, abs(ga64.item()))


def test_sinkhorn_backward_is_deterministic():
    from diffreg_b200 import ops
    g = torch.Generator().manual_seed(3)
    scores, G = torch.randn(1, 300, 400, generator=g).cuda(), torch.randn(1, 301, 401, generator=g).cuda()
    ones = lambda n: torch.ones(1, n, dtype=torch.bool, device="cuda")
    a = ops.sinkhorn_backward(scores, torch.tensor(1.0).cuda(), 3, ones(300), ones(400), G)
    b = ops.sinkhorn_backward(scores, torch.tensor(1.0).cuda(), 3, ones(300), ones(400), G)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


@pytest.mark.parametrize("B,N,M", [(1, 300, 400), (2, 130, 70), (1, 65, 1025)])
def test_dual_softmax_backward_against_the_oracle_in_fp64(B, N, M):
    from diffreg_b200 import ops
    g = torch.Generator().manual_seed(N + M)
    sim = torch.randn(B, N, M, generator=g) * 0.3
    sm, tm = torch.rand(B, N, generator=g) > 0.15, torch.rand(B, M, generator=g) > 0.15
    G = torch.randn(B, N, M, generator=g)
    want = O.dual_softmax_backward(sim.double(), sm, tm, 0.1, G.double())
    got = ops.dual_softmax_backward(sim.cuda(), sm.cuda(), tm.cuda(), 0.1, G.cuda()).cpu().double()
    assert (got - want).abs().max().item() <= REL * want.abs().max().item()
    again = ops.dual_softmax_backward(sim.cuda(), sm.cuda(), tm.cuda(), 0.1, G.cuda()).cpu().double()
    assert torch.equal(got, again)


def test_matching_2d3d_head_in_training_mode_against_the_reference_autograd():
    """The 2D-3D head (matching.py:91-147): conf_matrix and the gathered weights are differentiable."""
    import diffreg_b200
    g = load("lotb_matching2d3d_train")
    C = g["src_feats"].shape[-1]
    cfg = dict(match_type="sinkhorn", confidence_threshold=0.2, feature_dim=C, entangled=True, dsmax_temperature=0.1,
               skh_init_bin_score=1.0, skh_iters=3, skh_prefilter=False)
    head = diffreg_b200.Matching2D3D(cfg).cuda().train()
    with torch.no_grad():
        head.src_proj.weight.copy_(g["weight"].cuda())
        head.bin_score.copy_(torch.tensor(float(g["bin_score"])))
    src, tgt = g["src_feats"].cuda().requires_grad_(), g["tgt_feats"].cuda().requires_grad_()
    conf, si, ti, w = head(src, tgt, g["src_mask"].cuda(), g["tgt_mask"].cuda(), True)
    assert torch.equal(si.cpu(), g["src_indices"]) and torch.equal(ti.cpu(), g["tgt_indices"])
    assert (conf.detach().cpu() - g["conf"]).abs().max().item() <= 1e-5 and (w.detach().cpu() - g["weights"]).abs().max().item() <= 1e-5
    ((conf * g["W"].cuda()).sum() + (w * torch.arange(1, w.numel() + 1, device="cuda")).sum()).backward()
    for got, key in ((src.grad, "grad_src"), (tgt.grad, "grad_tgt"), (head.src_proj.weight.grad, "grad_weight")):
        want = g[key]
        assert (got.cpu() - want).abs().max().item() <= 1e-4 * max(1e-3, want.abs().max().item()), key
    assert abs(head.bin_score.grad.item() - float(g["grad_bin_score"])) <= 1e-4 * max(1.0, abs(float(g["grad_bin_score"])))


@pytest.mark.parametrize("name", names("lotb_matching_"))
def test_matching_forward_in_training_mode_against_the_reference_autograd(name):
    """Matching.forward with autograd recording (module in train(), features that require grad): conf_matrix and the gradients of
    the features, the projection weight and bin_score against the reference module's."""
    import diffreg_b200
    g = load(name)
    C = g["src_feats"].shape[-1]
    mt = str(g["match_type"])
    cfg = dict(match_type=mt, confidence_threshold=0.2, feature_dim=C, entangled=bool(int(g["entangled"])), dsmax_temperature=0.1,
               skh_init_bin_score=1.0, skh_iters=3, skh_prefilter=False)
    head = diffreg_b200.Matching(cfg).cuda().train()
    with torch.no_grad():
        head.src_proj.weight.copy_(g["weight"].cuda())
        if mt == "sinkhorn":
            head.bin_score.copy_(torch.tensor(float(g["bin_score"])))
    src, tgt = g["src_feats"].cuda().requires_grad_(), g["tgt_feats"].cuda().requires_grad_()
    spe = g["src_pe"].cuda() if "src_pe" in g else None
    tpe = g["tgt_pe"].cuda() if "tgt_pe" in g else None
    data = {}
    conf, match = head(src, tgt, spe, tpe, g["src_mask"].cuda(), g["tgt_mask"].cuda(), data, pe_type="rotary")
    assert conf.requires_grad and set(data) == {"src_feats_nopos", "tgt_feats_nopos", "src_feats", "tgt_feats"}
    assert (conf.detach().cpu() - g["conf"]).abs().max().item() <= 1e-5
    assert torch.equal(match.cpu(), g["match"])
    (conf * g["W"].cuda()).sum().backward()
    for got, key in ((src.grad, "grad_src"), (tgt.grad, "grad_tgt"), (head.src_proj.weight.grad, "grad_weight")):
        want = g[key]
        assert (got.cpu() - want).abs().max().item() <= 1e-4 * max(1e-3, want.abs().max().item()), key
    if mt == "sinkhorn":
        assert abs(head.bin_score.grad.item() - float(g["grad_bin_score"])) <= 1e-4 * max(1.0, abs(float(g["grad_bin_score"])))
    # eval mode + no_grad: the forward-only kernels, same confidences
    head.eval()
    with torch.no_grad():
        conf2, _ = head(src.detach(), tgt.detach(), spe, tpe, g["src_mask"].cuda(), g["tgt_mask"].cuda(), {}, pe_type="rotary")
    assert (conf2 - conf.detach()).abs().max().item() <= 1e-5


@pytest.mark.parametrize("name", names("procrb_kabsch_"))
def test_weighted_procrustes_backward_against_the_reference_autograd(name):
    from diffreg_b200.procrustes import SoftProcrustesLayer
    g = load(name)
    w = g["w"].cuda().requires_grad_()
    R, t, cond = SoftProcrustesLayer.batch_weighted_procrustes(g["X"].cuda(), g["Y"].cuda(), w)
    assert R.requires_grad and t.requires_grad and not cond.requires_grad
    assert (R.detach().cpu() - g["R"]).abs().max().item() <= 1e-5 and (t.detach().cpu() - g["t"]).abs().max().item() <= 1e-5
    ((R * g["grad_R"].cuda()).sum() + (t * g["grad_t"].cuda()).sum()).backward()
    want = g["grad_w"]
    assert (w.grad.cpu() - want).abs().max().item() <= 5e-5 * want.abs().max().item()
    # deterministic
    w2 = g["w"].cuda().requires_grad_()
    R2, t2, _ = SoftProcrustesLayer.batch_weighted_procrustes(g["X"].cuda(), g["Y"].cuda(), w2)
    ((R2 * g["grad_R"].cuda()).sum() + (t2 * g["grad_t"].cuda()).sum()).backward()
    assert torch.equal(w.grad, w2.grad)


def test_weighted_procrustes_backward_against_the_oracle_in_fp64():
    from diffreg_b200 import ops
    g = torch.Generator().manual_seed(17)
    B, K = 4, 4096
    X = torch.randn(B, K, 3, generator=g)
    Y = torch.stack([(O.random_rotation(g) @ X[b].t()).t() for b in range(B)]) + 0.05 * torch.randn(B, K, 3, generator=g) + 1.5
    w = torch.rand(B, K, 1, generator=g)
    gR, gt = torch.randn(B, 3, 3, generator=g), torch.randn(B, 3, 1, generator=g)
    R, t, _ = ops.weighted_procrustes(X.cuda(), Y.cuda(), w.cuda())
    R32, _, _ = O.batch_weighted_procrustes(X, Y, w)                      # (the reference's solve returns fp32: quirk Q5)
    assert (R.cpu() - R32).abs().max().item() <= 1e-5
    want = O.weighted_procrustes_backward(X.double(), Y.double(), w.double(), R32.double(), gR.double(), gt.double())
    got = ops.weighted_procrustes_backward(X.cuda(), Y.cuda(), w.cuda(), R, gR.cuda(), gt.cuda()).cpu().double()
    assert (got - want).abs().max().item() <= 5e-5 * want.abs().max().item()


def test_soft_procrustes_layer_in_training_mode_against_the_reference_autograd():
    """SoftProcrustesLayer.forward with a tracked confidence matrix: pose and dL/d conf (non-zero at the selected entries only)."""
    from types import SimpleNamespace
    import diffreg_b200
    g = load("procrb_layer")
    layer = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=float(g["sample_rate"]), max_condition_num=float(g["max_condition_num"])))
    conf = g["conf"].cuda().requires_grad_()
    R, t, Rf, tf, cond, mask = layer(conf, g["src_pcd"].cuda(), g["tgt_pcd"].cuda(), g["src_mask"].cuda(), g["tgt_mask"].cuda())
    assert (R.detach().cpu() - g["R"]).abs().max().item() <= 1e-5 and (t.detach().cpu() - g["t"]).abs().max().item() <= 1e-5
    assert bool(mask.all())
    ((Rf * g["grad_R"].cuda()).sum() + (tf * g["grad_t"].cuda()).sum()).backward()
    want = g["grad_conf"]
    assert torch.equal(conf.grad.cpu() != 0, want != 0)
    assert (conf.grad.cpu() - want).abs().max().item() <= 5e-5 * want.abs().max().item()
    # the forward-only call gives the same pose
    with torch.no_grad():
        R0 = layer(conf.detach(), g["src_pcd"].cuda(), g["tgt_pcd"].cuda(), g["src_mask"].cuda(), g["tgt_mask"].cuda())[0]
    assert (R0 - R.detach()).abs().max().item() <= 1e-5


def test_training_step_through_matching_and_pose_against_the_reference_modules():
    """Matching.forward -> conf -> SoftProcrustesLayer -> a loss on conf, R, t -> backward, with the reference's own modules on the
    same GPU as the yardstick (same weights): gradients of the features, the projection weight and bin_score."""
    from types import SimpleNamespace
    from oracle import ref_loader
    if not ref_loader.available():
        pytest.skip("reference sources not present (/root/reference or oracle/_ref)")
    import diffreg_b200
    ref = ref_loader.load_flavour("4d")
    try:
        torch.backends.cuda.matmul.allow_tf32 = False
        n, C = 320, 64
        pb = O.make_problem(11, 1, n, n - 40, C, prefix_valid=[(300, 260)])
        cfg = dict(match_type="sinkhorn", confidence_threshold=0.2, feature_dim=C, entangled=True, dsmax_temperature=0.1,
                   skh_init_bin_score=1.0, skh_iters=3, skh_prefilter=False)
        pcfg = SimpleNamespace(sample_rate=1.0, max_condition_num=1e9)
        t = {k: pb[k].cuda() for k in ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")}
        g = torch.Generator().manual_seed(1)
        Wc = torch.rand(1, n, n - 40, generator=g).cuda()
        gR, gt = torch.randn(1, 3, 3, generator=g).cuda(), torch.randn(1, 3, 1, generator=g).cuda()

        def step(head, proc):
            src, tgt = t["src_feats"].clone().requires_grad_(), t["tgt_feats"].clone().requires_grad_()
            conf, _ = head(src, tgt, None, None, t["src_mask"], t["tgt_mask"], {})
            R, tt, _, _, _, _ = proc(conf, t["s_pcd"], t["t_pcd"], t["src_mask"], t["tgt_mask"])
            ((conf * Wc).sum() + (R * gR).sum() + (tt * gt).sum()).backward()
            return src.grad, tgt.grad, head.src_proj.weight.grad, head.bin_score.grad
        head = diffreg_b200.Matching(cfg).cuda().train()
        rhead = ref.matching.Matching(cfg).cuda().train()
        rhead.load_state_dict(head.state_dict())
        a = step(head, diffreg_b200.SoftProcrustesLayer(pcfg))
        b = step(rhead, ref.procrustes.SoftProcrustesLayer(pcfg))
        for x, y in zip(a[:3], b[:3]):
            assert (x - y).abs().max().item() <= 1e-4 * y.abs().max().item()
        assert abs(a[3].item() - b[3].item()) <= 1e-4 * max(1.0, abs(b[3].item()))
    finally:
        ref_loader.unload()
