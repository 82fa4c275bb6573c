"""SoftProcrustes (drg_soft_procrustes, drg_weighted_procrustes) against the reference's golden
vectors and the oracle.  Tolerances from BASELINE.json: rotation 1e-5 rad, translation 1e-5 m."""
import pytest
import torch

from oracle import diffreg_oracle as O
from helpers import TOL_ROT, TOL_TRANS, load, names, rot_angle

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _ops():
    import diffreg_b200
    return diffreg_b200.ops


def test_kabsch_golden():
    g = load("kabsch_b4")
    R, t, cond = _ops().weighted_procrustes(g["X"].cuda(), g["Y"].cuda(), g["w"].cuda())
    assert rot_angle(R.cpu(), g["R"]).max() <= TOL_ROT
    assert (t.cpu() - g["t"]).abs().max() <= TOL_TRANS
    assert torch.allclose(cond.cpu(), g["condition"], rtol=1e-5)


@pytest.mark.parametrize("name", names("procrustes"))
def test_soft_procrustes_golden(name):
    g = load(name)
    out = _ops().soft_procrustes(g["conf"].cuda(), g["s_pcd"].cuda(), g["t_pcd"].cuda(), g["src_mask"].cuda(), g["tgt_mask"].cuda(),
                                 float(g["sample_rate"]), float(g["max_condition_num"]),
                                 padded_lengths=name.startswith("procrustes3d"), want_warped=True)
    ok_ref = g["solution_mask"]
    assert torch.equal(out["solution_mask"].cpu(), ok_ref)
    well = torch.isfinite(g["condition"]) & (g["condition"] < 1e4)
    if well.any():
        assert rot_angle(out["R"].cpu()[well], g["R"][well]).max() <= TOL_ROT
        assert (out["t"].cpu()[well] - g["t"][well]).abs().max() <= TOL_TRANS
        assert torch.allclose(out["condition"].cpu()[well], g["condition"][well], rtol=1e-4)
    assert rot_angle(out["R_forwd"].cpu(), g["R_forwd"]).max() <= TOL_ROT
    assert (out["t_forwd"].cpu() - g["t_forwd"]).abs().max() <= TOL_TRANS
    warped = O.warp_points(g["R_forwd"], g["t_forwd"], g["s_pcd"])
    assert (out["src_warped"].cpu() - warped).abs().max() <= 2e-5


def _conf_like(B, N, M, gen, sharp=6.0):
    """A Sinkhorn-looking confidence matrix: softmax rows of scaled noise."""
    return torch.softmax(torch.randn(B, N, M, generator=gen) * sharp, dim=2)


@pytest.mark.parametrize("B,N,M,rate,prefix", [(1, 64, 48, 1.0, False), (3, 200, 260, 1.0, True), (2, 700, 300, 0.5, True),
                                               (1, 1024, 1024, 1.0, False), (1, 37, 1500, 0.3, False)])
def test_soft_procrustes_vs_oracle(B, N, M, rate, prefix):
    gen = torch.Generator().manual_seed(B * 100 + N + M)
    pv = [(int(N * 0.8) - b, int(M * 0.9) - 2 * b) for b in range(B)] if prefix else None
    pb = O.make_problem(17 + N, B, N, M, C=8, prefix_valid=pv)
    conf = _conf_like(B, N, M, gen)
    conf = conf * O.pair_mask(pb["src_mask"], pb["tgt_mask"])
    ref = O.soft_procrustes(conf, pb["s_pcd"], pb["t_pcd"], pb["src_mask"], pb["tgt_mask"], rate, 1e9)
    out = _ops().soft_procrustes(conf.cuda(), pb["s_pcd"].cuda(), pb["t_pcd"].cuda(), pb["src_mask"].cuda(), pb["tgt_mask"].cuda(),
                                 rate, 1e9, want_selection=True)
    # the selected set is exactly the top-K set of the reference's sort (values are distinct)
    for b in range(B):
        cap = int(max(int(pb["src_mask"][b].sum()), int(pb["tgt_mask"][b].sum())) * rate)
        K = int(torch.tensor([int(max(int(pb["src_mask"][i].sum()), int(pb["tgt_mask"][i].sum())) * rate) for i in range(B)],
                             dtype=torch.float32).mean().int())
        kb = min(K, cap)
        w = out["sel_w"][b].cpu()
        assert int((w > 0).sum()) <= kb
        top = conf[b].reshape(-1).sort(descending=True)[0][:kb]
        got = w.sort(descending=True)[0][:kb]
        assert torch.equal(got, top)
    assert rot_angle(out["R"].cpu(), ref[0]).max() <= TOL_ROT
    assert (out["t"].cpu() - ref[1]).abs().max() <= TOL_TRANS


@pytest.mark.parametrize("kind", ["tied", "sparse"])
def test_soft_procrustes_ties_take_the_general_select(kind):
    """Heavily tied confidences put every candidate into one histogram bin: the select kernel must leave its short-list
    path for the general radix select and still return exactly K entries whose values are the K largest (ties are
    ordered by lowest flat index on the GPU; the reference's sort order among equal values is unspecified, so the pose is
    only checked for being a finite rotation)."""
    B, N, M = 1, 700, 600
    gen = torch.Generator().manual_seed(77)
    pb = O.make_problem(99, B, N, M, C=8)
    if kind == "tied":
        conf = torch.floor(torch.rand(B, N, M, generator=gen) * 8.0) / 8.0        # 8 distinct values, ~52 k entries each
    else:
        conf = torch.zeros(B, N, M)
        idx = torch.randperm(N * M, generator=gen)[:300]                            # fewer positive entries than K
        conf.view(-1)[idx] = 0.2 + 0.7 * torch.rand(300, generator=gen)
    out = _ops().soft_procrustes(conf.cuda(), pb["s_pcd"].cuda(), pb["t_pcd"].cuda(), pb["src_mask"].cuda(), pb["tgt_mask"].cuda(),
                                 1.0, 1e9, want_selection=True)
    kb = max(N, M)
    w = out["sel_w"][0].cpu()
    src, tgt = out["sel_src"][0].cpu().long(), out["sel_tgt"][0].cpu().long()
    top = conf[0].reshape(-1).sort(descending=True)[0][:kb]
    assert torch.equal(w.sort(descending=True)[0][:kb], top)
    flat = (src * M + tgt)[:kb]
    assert flat.unique().numel() == kb                                              # K distinct entries
    assert torch.equal(conf[0].reshape(-1)[flat], w[:kb])                           # the weights are the entries' values
    if kind == "tied":
        # among the entries tied with the K-th value the lowest flat indices win
        kth = top[-1]
        tied_sel = flat[w[:kb] == kth].sort()[0]
        tied_all = (conf[0].reshape(-1) == kth).nonzero().flatten()
        assert torch.equal(tied_sel, tied_all[: tied_sel.numel()])
    R = out["R"].cpu()[0]
    assert torch.isfinite(R).all() and (R @ R.t() - torch.eye(3)).abs().max() <= 1e-4


def test_soft_procrustes_full_size_recovers_motion():
    """4096 x 4096 with a planted permutation: the recovered pose is the planted rigid motion."""
    N = M = 4096
    gen = torch.Generator().manual_seed(3)
    s = torch.randn(1, N, 3, generator=gen)
    Rgt = O.random_rotation(gen)
    tgt = torch.randn(3, generator=gen)
    perm = torch.randperm(N, generator=gen)
    t_pcd = (s[0] @ Rgt.t() + tgt)[perm][None]
    conf = torch.rand(1, N, M, generator=gen) * 1e-3
    conf[0, perm, torch.arange(M)] = 0.5 + 0.4 * torch.rand(M, generator=gen)     # src perm[j] <-> tgt j
    ones = torch.ones(1, N, dtype=torch.bool)
    out = _ops().soft_procrustes(conf.cuda(), s.cuda(), t_pcd.cuda(), ones.cuda(), ones.cuda(), 1.0, 40.0, want_warped=True)
    assert bool(out["solution_mask"][0])
    assert rot_angle(out["R"].cpu()[0], Rgt) <= 1e-4
    assert (out["t"].cpu()[0, :, 0] - tgt).abs().max() <= 1e-4
    assert (out["src_warped"].cpu()[0] - (s[0] @ Rgt.t() + tgt)).abs().max() <= 1e-3


def test_degenerate_inputs_gate_to_identity():
    ops = _ops()
    N, M = 50, 60
    conf = torch.zeros(1, N, M)
    pts = torch.zeros(1, N, 3)
    ones_s = torch.ones(1, N, dtype=torch.bool)
    ones_t = torch.ones(1, M, dtype=torch.bool)
    out = ops.soft_procrustes(conf.cuda(), pts.cuda(), torch.zeros(1, M, 3).cuda(), ones_s.cuda(), ones_t.cuda(), 1.0, 40.0)
    assert not bool(out["solution_mask"][0])
    assert torch.equal(out["R_forwd"].cpu()[0], torch.eye(3)) and torch.equal(out["t_forwd"].cpu()[0], torch.zeros(3, 1))


@pytest.mark.parametrize("B,N,M,kind", [(1, 64, 48, "full"), (2, 200, 260, "prefix"), (1, 1024, 1024, "arbitrary"), (1, 37, 1530, "arbitrary")])
def test_fused_sinkhorn_procrustes_matches_two_calls(B, N, M, kind):
    """drg_sinkhorn_soft_procrustes (no confidence matrix in memory) == drg_sinkhorn(conf) + drg_soft_procrustes."""
    ops = _ops()
    gen = torch.Generator().manual_seed(B + N + M)
    pv = [(N - 5 - b, M - 9 - b) for b in range(B)] if kind == "prefix" else None
    pb = O.make_problem(5 + N, B, N, M, C=8, prefix_valid=pv, arbitrary_invalid=0.1 if kind == "arbitrary" else 0.0)
    x = torch.randn(B, N, M, generator=gen) * 2.0
    alpha = torch.tensor(1.0).cuda()
    args = (x.cuda(), alpha, 3, pb["src_mask"].cuda(), pb["tgt_mask"].cuda())
    conf = ops.sinkhorn(*args, out_mode="conf", apply_mask=True)
    two = ops.soft_procrustes(conf, pb["s_pcd"].cuda(), pb["t_pcd"].cuda(), pb["src_mask"].cuda(), pb["tgt_mask"].cuda(), 1.0, 1e9,
                              want_warped=True)
    one = ops.sinkhorn_soft_procrustes(*args, pb["s_pcd"].cuda(), pb["t_pcd"].cuda(), 1.0, 1e9, apply_mask=True, want_warped=True)
    assert rot_angle(one["R"].cpu(), two["R"].cpu()).max() <= TOL_ROT
    assert (one["t"] - two["t"]).abs().max() <= TOL_TRANS
    assert (one["src_warped"] - two["src_warped"]).abs().max() <= 5e-5
    assert torch.equal(one["solution_mask"], two["solution_mask"])


@pytest.mark.parametrize("fused", [True, False])
def test_pose_is_bit_reproducible(fused):
    """the candidate list is appended with atomics, so its order changes from run to run; the pose
    kernel sorts each CTA's selected candidates by flat index and adds lanes / warps / CTAs in fixed order, so R, t, the
    condition number and the warped points must be IDENTICAL bit for bit over repeated calls -- on the fused
    Sinkhorn -> pose path (candidate search inside the persistent Sinkhorn) and on the stand-alone path (stored matrix)."""
    from diffreg_b200 import ops
    N, M = 1536, 1280
    pb = O.make_problem(99, 1, N, M, 8, prefix_valid=[(1500, 1250)])
    g = torch.Generator().manual_seed(5)
    x = torch.randn(1, N, M, generator=g).to(DEV) * 3.0
    alpha = torch.tensor(1.0, device=DEV)
    args = [pb[k].to(DEV) for k in ("src_mask", "tgt_mask", "s_pcd", "t_pcd")]
    junk = torch.empty(64 << 20, dtype=torch.uint8, device=DEV)

    def run():
        if fused:
            return ops.sinkhorn_soft_procrustes(x, alpha, 3, args[0], args[1], args[2], args[3], 1.0, 1e9)
        conf = ops.sinkhorn(x, alpha, 3, args[0], args[1], out_mode="conf", apply_mask=True)
        return ops.soft_procrustes(conf, args[2], args[3], args[0], args[1], 1.0, 1e9, want_warped=True)

    first = {k: v.clone() for k, v in run().items()}
    for rep in range(6):
        if rep % 2:
            junk.random_(0, 255)           # disturb the timing of the next call (cold L2)
        out = run()
        for k in ("R", "t", "R_forwd", "t_forwd", "condition", "src_warped"):
            assert torch.equal(out[k], first[k]), f"{k} changed on repetition {rep}"
