"""BASELINE.json's five configurations at their FULL sizes, CUDA path against the oracle on the same seeded inputs.

The oracle is plain torch code; at these sizes it runs on the GPU box's CUDA device as eager PyTorch (the
"reference's own PyTorch implementation of this path on the same inputs" of the north_star -- fp32, allow_tf32 off)
so that a 16384^2 x 100-iteration Sinkhorn or a 16 x 2048^2 full sort finishes in seconds.  Nothing here reads
/root/reference.  Tolerances: 1e-4 abs on confidences / log matrix, 1e-5 rad and 1e-5 m on the pose, indices
bit-exact above a 1e-5 top-1 margin."""
from types import SimpleNamespace

import pytest
import torch

from oracle import diffreg_oracle as O
from helpers import TOL_LOG, TOL_ROT, TOL_TRANS, check_top1_pairs, finite_close, rot_angle

pytestmark = pytest.mark.gpu
DEV = "cuda"
KEYS = ("src_feats", "tgt_feats", "s_pcd", "t_pcd", "src_mask", "tgt_mask")


@pytest.fixture(autouse=True)
def _fp32_reference():
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cuda.matmul.allow_tf32 = old
    torch.cuda.empty_cache()


def _cfg(C, match_type="sinkhorn"):
    return dict(match_type=match_type, confidence_threshold=0.2, feature_dim=C, entangled=True, dsmax_temperature=0.1,
                skh_init_bin_score=1.0, skh_iters=3, skh_prefilter=False)


def _head(cls, pb, match_type="sinkhorn", **kw):
    import diffreg_b200
    C = pb["W"].shape[0]
    m = getattr(diffreg_b200, cls)(_cfg(C, match_type), **kw).to(DEV).eval()
    with torch.no_grad():
        m.src_proj.weight.copy_(pb["W"].to(DEV))
    return m


def _params(pb, dev):
    return O.MatchingParams(src_proj_weight=pb["W"].to(dev), bin_score=torch.tensor(1.0, device=dev))


def test_config0_4dmatch_single_pair_1024():
    """configs[0]: 4DMatch-shaped single pair, N=M=1024, d=256, Sinkhorn + SoftProcrustes, one 4d sampler step
    (noise supplied), against the oracle on the CPU -- the reference's own CPU-runnable case."""
    import diffreg_b200
    N = M = 1024
    pb = O.make_problem(1000, 1, N, M, 256)
    g = torch.Generator().manual_seed(1001)
    x_T = torch.randn(1, N, M, generator=g)
    noise = torch.randn(1, N, M, generator=g)
    trace_ref = []
    ref = O.sampler("4d", _params(pb, "cpu"), *[pb[k] for k in KEYS], x_T, 1, noises=[noise], state_dtype=torch.float32,
                    trace=trace_ref)
    head = _head("Matching", pb)
    proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
    smp = diffreg_b200.DenoisingSampler("4d", head, proc, 1)
    trace = []
    out = smp.sample(x_T.to(DEV), *[pb[k].to(DEV) for k in KEYS], noises=[noise.to(DEV)], trace=trace)
    t0, r0 = trace[0], trace_ref[0]
    assert (t0["x0"].cpu() - r0["x0"]).abs().max() <= TOL_LOG
    assert (t0["conf_d"].cpu() - r0["conf_d"]).abs().max() <= TOL_LOG
    assert rot_angle(t0["pose"]["R"].cpu(), r0["pose"][0]).max() <= TOL_ROT
    assert (t0["pose"]["t"].cpu() - r0["pose"][1]).abs().max() <= TOL_TRANS
    assert (t0["pose"]["src_warped"].cpu() - r0["warped"]).abs().max() <= 5e-5
    ok, err = finite_close(t0["x_out"].cpu(), r0["x_out"].float(), TOL_LOG)
    assert ok, err
    ok, err = finite_close(out["conf_matrix_pred"].cpu(), ref["conf_matrix_pred"].float(), TOL_LOG)
    assert ok, err


def test_config1_3dmatch_batch16_dual_softmax_2048():
    """configs[1]: 16 pairs, valid counts in [1792, 2048] padded to the batch maximum (prefix masks, padded features
    and points zero), dual-softmax matching + SoftProcrustes, one step."""
    import diffreg_b200
    B, L = 16, 2048
    g = torch.Generator().manual_seed(2000)
    valid = [(int(torch.randint(1792, L + 1, (1,), generator=g)), int(torch.randint(1792, L + 1, (1,), generator=g))) for _ in range(B)]
    valid[0] = (L, L)
    pb = O.make_problem(2001, B, L, L, 256, prefix_valid=valid)
    d = {k: pb[k].to(DEV) for k in KEYS}
    p = _params(pb, DEV)
    p.match_type = "dual_softmax"
    ref_conf, ref_match = O.matching_forward_3d(p, d["src_feats"], d["tgt_feats"], None, None, d["src_mask"], d["tgt_mask"])[:2]
    head = _head("Matching", pb, "dual_softmax")
    data = {}
    conf, match = head(d["src_feats"], d["tgt_feats"], None, None, d["src_mask"], d["tgt_mask"], data, None)
    assert conf.shape == (B, L, L)
    assert (conf - ref_conf).abs().max().item() <= TOL_LOG
    # matches at thr 0.2: identical unless an entry sits within 1e-5 of the threshold / of a competing maximum
    got = set(map(tuple, match.cpu().tolist()))
    want = set(map(tuple, ref_match.cpu().tolist()))
    for (b, i, j) in got ^ want:
        c = ref_conf[b, i, j].item()
        near_thr = abs(c - 0.2) <= 1e-5
        near_tie = (ref_conf[b, i].max().item() - c <= 1e-5) and (ref_conf[b, :, j].max().item() - c <= 1e-5)
        assert near_thr or near_tie, (b, i, j, c)
    # pose from the reference confidences (fp32) through both implementations
    proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
    R, t, Rf, tf, cond, ok_mask = proc(ref_conf, d["s_pcd"], d["t_pcd"], d["src_mask"], d["tgt_mask"])
    rR, rt, rRf, rtf, rcond, rok = O.soft_procrustes(ref_conf, d["s_pcd"], d["t_pcd"], d["src_mask"], d["tgt_mask"], 1.0, 40.0)
    assert rot_angle(R.cpu(), rR.cpu()).max() <= TOL_ROT
    assert (t.cpu() - rt.cpu()).abs().max() <= TOL_TRANS
    assert torch.equal(ok_mask.cpu().bool(), rok.cpu().bool())
    assert torch.allclose(cond.cpu().double(), rcond.cpu().double(), rtol=1e-4)


def test_config2_4dmatch_sampler_4096_steps_and_properties():
    """configs[2] (the bench workload): N=M=4096, d=256, 4d sampler.  Two steps against the oracle trace with the noise
    supplied, then the throughput mode (in-kernel Philox noise) checked through properties of the drawn noise."""
    import diffreg_b200
    from diffreg_b200.sampler import ddim_coefficients
    N = M = 4096
    pb = O.make_problem(3000, 1, N, M, 256)
    d = {k: pb[k].to(DEV) for k in KEYS}
    g = torch.Generator(device=DEV).manual_seed(3001)
    x_T = torch.randn(1, N, M, generator=g, device=DEV)
    noises = [torch.randn(1, N, M, generator=g, device=DEV) for _ in range(2)]
    # (the oracle's sampler() runs all 20 steps; its loop body is driven here for two steps only)
    ac = O.alphas_cumprod()
    pairs = O.time_pairs(20)
    p = _params(pb, DEV)
    x = x_T.clone()
    head = _head("Matching", pb)
    proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
    smp = diffreg_b200.DenoisingSampler("4d", head, proc, 20)
    xg = x_T.clone()
    for k in range(2):
        warped, conf_d, pose = O.noisy_matching_to_pose(x, p.bin_score, p.skh_iters, d["s_pcd"], d["t_pcd"], d["src_mask"],
                                                        d["tgt_mask"], 1.0, 40.0)
        sim, *_ = O.similarity(p, d["src_feats"], d["tgt_feats"])
        x0 = O.confidence_from_similarity(p, sim, d["src_mask"], d["tgt_mask"])
        x = O.ddim_update(x, x0, ac, pairs[k][0], pairs[k][1], noises[k]).float()
        xg, _, aux = smp.step(k, xg, None, *[d[kk] for kk in KEYS], noise=noises[k], want_x0=True)
        assert (aux["x0"] - x0).abs().max().item() <= TOL_LOG
        assert (aux["conf_d"] - conf_d).abs().max().item() <= TOL_LOG
        assert rot_angle(aux["pose"]["R"].cpu(), pose[0].cpu()).max() <= TOL_ROT
        assert (aux["pose"]["t"].cpu() - pose[1].cpu()).abs().max() <= TOL_TRANS
        assert (aux["pose"]["src_warped"] - warped).abs().max().item() <= 5e-5
        assert (xg - x).abs().max().item() <= TOL_LOG
    # throughput mode: x_next - (deterministic part) = sigma * eps with eps ~ N(0, 1), fresh per step
    counter = torch.zeros(1, dtype=torch.int64, device=DEV)
    smp2 = diffreg_b200.DenoisingSampler("4d", head, proc, 20, noise_seed=5)
    zero = torch.zeros_like(noises[0])
    eps = []
    for k in range(2):
        a, _, _ = smp2.step(0, x_T, None, *[d[kk] for kk in KEYS], noise_counter=counter)
        b, _, _ = smp.step(0, x_T, None, *[d[kk] for kk in KEYS], noise=zero)
        _, _, sigma = ddim_coefficients(smp.ac, *smp.pairs[0])
        eps.append(((a - b) / sigma).flatten())
    for e in eps:
        assert abs(e.mean().item()) < 2e-3 and abs(e.std().item() - 1.0) < 2e-3
        assert abs((e ** 3).mean().item()) < 1e-2 and abs((e ** 4).mean().item() - 3.0) < 3e-2
    assert abs((eps[0] * eps[1]).mean().item()) < 2e-3          # the device counter advanced: independent draws
    assert int(counter.item()) == 2


def test_config3_2d3d_sampler_4800x2048_10_steps():
    """configs[3]: 2D-3D flavour, rectangular N=4800 x M=2048, ~5 % invalid entries on each side (arbitrary masks),
    10 steps without noise, final Sinkhorn + mutual_topk_select(k=1, mutual=False)."""
    import diffreg_b200
    N, M, STEPS = 4800, 2048, 10
    pb = O.make_problem(4000, 1, N, M, 256, arbitrary_invalid=0.05)
    d = {k: pb[k].to(DEV) for k in KEYS}
    x_T = torch.randn(1, N, M, generator=torch.Generator(device=DEV).manual_seed(4001), device=DEV)
    ref = O.sampler("2d3d", _params(pb, DEV), *[d[k] for k in KEYS], x_T, STEPS, state_dtype=torch.float32)
    head = _head("Matching2D3D", pb)
    proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
    smp = diffreg_b200.DenoisingSampler("2d3d", head, proc, STEPS)
    out = smp.sample(x_T.clone(), *[d[k] for k in KEYS])
    ok, err = finite_close(out["x_final"].cpu(), ref["x_final"].float().cpu(), TOL_LOG)
    assert ok, err
    assert (out["conf_matrix_pred"] - ref["conf_matrix_pred"]).abs().max().item() <= TOL_LOG
    mp = out["match_pred"].cpu()
    ok, msg = check_top1_pairs(ref["conf_matrix_pred"][0].cpu(), mp[:, 1], mp[:, 2], False)
    assert ok, msg
    assert rot_angle(out["pose"]["R"].cpu(), ref["pose"][0].cpu()).max() <= TOL_ROT
    assert (out["pose"]["t"].cpu() - ref["pose"][1].cpu()).abs().max() <= TOL_TRANS


def test_config4_sinkhorn_16384_100_iterations_and_row_shards():
    """configs[4]: bare log_optimal_transport, N=M=16384, I=100, masks all true, scores ~ N(0,1), alpha = 1; also the
    row-sharded kernels with 8 emulated shards on this one GPU (the NCCL reduction itself is covered by the gloo tests)."""
    import diffreg_b200
    N = M = 16384
    s = torch.randn(1, N, M, generator=torch.Generator(device=DEV).manual_seed(5000), device=DEV)
    sm = torch.ones(1, N, dtype=torch.bool, device=DEV)
    tm = torch.ones(1, M, dtype=torch.bool, device=DEV)
    alpha = torch.tensor(1.0, device=DEV)
    ref = O.log_optimal_transport(s, alpha, 100, sm, tm)
    out = diffreg_b200.log_optimal_transport(s, alpha, 100, sm, tm)
    assert out.shape == (1, N + 1, M + 1)
    err = (out - ref).abs().max().item()
    assert err <= TOL_LOG, err
    # converged plan: both marginals (fp64 sums of the fp32 log matrix)
    # (log_optimal_transport returns Z + u + v - norm: every real row and column then carries mass 1)
    P = out[0].double().exp()
    assert P[:N].sum(dim=1).log().abs().max().item() <= 1e-4
    assert P[:, :M].sum(dim=0).log().abs().max().item() <= 1e-4
    del P, ref
    whole = diffreg_b200.ops.sinkhorn(s, alpha, 100, sm, tm, out_mode="conf")
    shards = diffreg_b200.EmulatedRowShards(8)(s, alpha, 100, sm, tm, out_mode="conf")
    assert (shards - whole).abs().max().item() <= 2e-6
    assert (whole - out[:, :N, :M].exp()).abs().max().item() <= 1e-6


def test_config2_all_20_steps_against_the_fp64_creep_of_the_reference():
    """the bench runs 20 sampler steps with an fp32 state; the reference's state drifts to fp64 after the
    first step (its fp64 schedule buffers promote it).  All 20 steps at N = M = 4096 with the noise supplied,
    against the oracle with state_dtype=None (the reference's dtype behaviour, its Sinkhorn on the state in fp64): the final
    sigmoid(x) must agree to 1e-4, and the per-step drift of the state is written to gpurun_out/ (copied to profiles/)."""
    import json
    import os
    import diffreg_b200
    N = M = 4096
    STEPS = 20
    pb = O.make_problem(3000, 1, N, M, 256)
    d = {k: pb[k].to(DEV) for k in KEYS}
    g = torch.Generator(device=DEV).manual_seed(3003)
    x_T = torch.randn(1, N, M, generator=g, device=DEV)
    noises = [torch.randn(1, N, M, generator=g, device=DEV) for _ in range(STEPS)]
    p = _params(pb, DEV)
    ac = O.alphas_cumprod()
    pairs = O.time_pairs(STEPS)
    head = _head("Matching", pb)
    proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
    smp = diffreg_b200.DenoisingSampler("4d", head, proc, STEPS)
    sim, *_ = O.similarity(p, d["src_feats"], d["tgt_feats"])                    # fixed features: x0 is the same every step
    x0_ref = O.confidence_from_similarity(p, sim, d["src_mask"], d["tgt_mask"])
    del sim
    x = x_T.clone()        # reference state: fp32 at the start, fp64 from the first update on
    xg = x_T.clone()       # ours: fp32 throughout
    drift = []
    for k in range(STEPS):
        warped, _, pose = O.noisy_matching_to_pose(x, p.bin_score, p.skh_iters, d["s_pcd"], d["t_pcd"], d["src_mask"],
                                                   d["tgt_mask"], 1.0, 40.0)
        x = O.ddim_update(x, x0_ref, ac, pairs[k][0], pairs[k][1], noises[k])      # state_dtype=None: stays fp64
        assert x.dtype == torch.float64
        xg, _, aux = smp.step(k, xg, None, *[d[kk] for kk in KEYS], noise=noises[k])
        err = (xg.double() - x).abs().max().item()
        scale = x.abs().max().item()
        ang = rot_angle(aux["pose"]["R"].cpu(), pose[0].cpu()).max().item()
        drift.append({"step": k, "t": pairs[k][0], "max_abs_state_err": err, "state_abs_max": scale,
                      "pose_angle_err_rad": ang, "pose_t_err": (aux["pose"]["t"].cpu() - pose[1].cpu()).abs().max().item(),
                      "gate_agrees": bool(torch.equal(aux["pose"]["solution_mask"].cpu().bool(), pose[5].cpu().bool()))})
        assert err <= TOL_LOG * max(1.0, scale), (k, err, scale)
        del warped
    conf_ref = torch.sigmoid(x)
    conf = diffreg_b200.ops.sigmoid(xg)
    final_err = (conf.double() - conf_ref).abs().max().item()
    out_dir = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
    os.makedirs(out_dir, exist_ok=True)
    with open(os.path.join(out_dir, "r2_drift_20steps_4096.json"), "w") as f:
        json.dump({"workload": "BASELINE configs[2], 20 steps, N=M=4096, noise supplied; ours (fp32 state) vs oracle with the "
                               "reference's fp64 state creep", "final_sigmoid_max_abs_err": final_err, "steps": drift}, f, indent=1)
    assert final_err <= TOL_LOG, final_err


def test_config1_3dmatch_variant_and_3d_sampler_at_2048():
    """configs[1] through the 3DMatch SoftProcrustes variant (padded lengths, 3d procrustes.py:61-62),
    and the '3d' sampler flavour (x -= x.min() first, no noise, max_condition_num = 0 as configs/test/3dmatch.yaml ships it,
    final Sinkhorn + top-1 union) at 2048 x 2048 against the oracle."""
    import diffreg_b200
    from diffreg_b200.procrustes import SoftProcrustesLayer3DMatch
    B, L = 4, 2048
    g = torch.Generator().manual_seed(2100)
    valid = [(int(torch.randint(1792, L + 1, (1,), generator=g)), int(torch.randint(1792, L + 1, (1,), generator=g))) for _ in range(B)]
    pb = O.make_problem(2101, B, L, L, 256, prefix_valid=valid)
    d = {k: pb[k].to(DEV) for k in KEYS}
    p = _params(pb, DEV)
    p.match_type = "dual_softmax"
    ref_conf = O.matching_forward_3d(p, d["src_feats"], d["tgt_feats"], None, None, d["src_mask"], d["tgt_mask"])[0]
    proc3 = SoftProcrustesLayer3DMatch(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
    R, t, Rf, tf, cond, ok_mask = proc3(ref_conf, d["s_pcd"], d["t_pcd"], d["src_mask"], d["tgt_mask"])
    rR, rt, rRf, rtf, rcond, rok = O.soft_procrustes(ref_conf, d["s_pcd"], d["t_pcd"], d["src_mask"], d["tgt_mask"], 1.0, 40.0,
                                                     padded_lengths=True)
    assert rot_angle(R.cpu(), rR.cpu()).max() <= TOL_ROT
    assert (t.cpu() - rt.cpu()).abs().max() <= TOL_TRANS
    assert torch.equal(ok_mask.cpu().bool(), rok.cpu().bool())
    del ref_conf
    # the 3d sampler flavour, one pair, three steps + the final Sinkhorn / selection
    # (all-true masks: with padding the reference's own x - x.min() turns the state into inf / nan after the first step,
    #  because get_warped_from_noising_matching leaves -inf in the caller's x; 3DMatch tests run unpadded)
    pb1 = O.make_problem(2102, 1, L, L, 256)
    d1 = [pb1[k].to(DEV) for k in KEYS]
    x_T = torch.randn(1, L, L, generator=torch.Generator().manual_seed(2103)).to(DEV)
    p1 = _params(pb1, DEV)
    ref = O.sampler("3d", p1, *d1, x_T, 3, max_condition_num=0.0, state_dtype=torch.float32)
    head = _head("Matching", pb1)
    proc = SoftProcrustesLayer3DMatch(SimpleNamespace(sample_rate=1.0, max_condition_num=0.0))
    out = diffreg_b200.DenoisingSampler("3d", head, proc, 3).sample(x_T, *d1)
    ok, err = finite_close(out["conf_matrix_pred"].cpu(), ref["conf_matrix_pred"].cpu().float(), TOL_LOG)
    assert ok, err
    okp, msg = check_top1_pairs(ref["conf_matrix_pred"][0].cpu(), out["match_pred"][:, 1].cpu(), out["match_pred"][:, 2].cpu(), mutual=False)
    assert okp, msg
