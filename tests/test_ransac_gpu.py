"""Correspondence RANSAC (csrc/ransac.cu) against the oracle's restatement of open3d's
registration_ransac_based_on_correspondence (Diff-Reg-4dmatch/models/loss.py:13-24, 366-398).  PARITY UNPINNED: open3d is not
in this image and its draws are not reproducible; the oracle evaluates the SAME trials (shared counter-based draws), so the
comparison is trial by trial: inlier counts equal except where a distance sits on the threshold to fp32 rounding."""
import numpy as np
import pytest
import torch

from helpers import rot_angle
from oracle import diffreg_oracle as orc

pytestmark = pytest.mark.gpu


def _problem(seed, B, N, M, corr_per_b, inlier_ratio, noise=0.005):
    """Planted rigid motions: tgt[j] = R src[i] + t (+ noise) for the inlier correspondences, random partners for the rest."""
    g = torch.Generator().manual_seed(seed)
    src = torch.rand(B, N, 3, generator=g) * 2 - 1
    tgt = torch.rand(B, M, 3, generator=g) * 2 - 1
    Rs, ts, rows = [], [], []
    for b in range(B):
        R = orc.random_rotation(g).float()
        t = torch.rand(3, generator=g) - 0.5
        C = corr_per_b[b]
        i = torch.randperm(N, generator=g)[:C]
        j = torch.randperm(M, generator=g)[:C]
        n_in = int(round(C * inlier_ratio))
        tgt[b, j[:n_in]] = src[b, i[:n_in]] @ R.T + t + noise * torch.randn(n_in, 3, generator=g)
        Rs.append(R)
        ts.append(t)
        rows.append(torch.stack([torch.full((C,), b, dtype=torch.int64), i, j], 1))
    return src, tgt, torch.cat(rows) if rows else torch.zeros(0, 3, dtype=torch.int64), torch.stack(Rs), torch.stack(ts)


@pytest.mark.parametrize("ransac_n", [3, 4])
def test_trials_match_the_oracle(ransac_n):
    from diffreg_b200 import ops
    B, T = 3, 4096
    src, tgt, match, Rgt, tgt_t = _problem(11 + ransac_n, B, 2000, 2100, [200, 1500, 57], 0.4)
    got = ops.ransac_correspondence(src.cuda(), tgt.cuda(), match.cuda(), 0.05, ransac_n, T, seed=1234, want_trials=True)
    ref = orc.ransac_correspondence(src.numpy(), tgt.numpy(), match.numpy(), 0.05, ransac_n, T, seed=1234)
    cnt, err = got["trial_count"].cpu().numpy(), got["trial_err2"].cpu().numpy()
    # same draws -> the same trials are degenerate (repeated correspondence) ...
    assert ((cnt < 0) == (ref["trial_count"] < 0)).all()
    # ... and the same correspondences are inliers, up to distances on the threshold (fp32 FMA order of the two evaluations)
    diff = np.abs(cnt - ref["trial_count"])
    assert diff.max() <= 2 and (diff > 0).mean() < 0.01, (diff.max(), (diff > 0).mean())
    same = diff == 0
    assert np.allclose(err[same], ref["trial_err2"][same], rtol=2e-3, atol=1e-6)
    for b in range(B):
        # the device's arg-max is exactly open3d's sequential "is better" scan over the device's own trial records
        assert int(got["best_trial"][b]) == orc.ransac_best_trial(cnt[b], err[b])
        h = int(got["best_trial"][b])
        assert int(got["inlier_count"][b]) == cnt[b, h]
        C = int((match[:, 0] == b).sum())
        assert abs(float(got["fitness"][b]) - cnt[b, h] / C) < 1e-6
        assert abs(float(got["inlier_rmse"][b]) - np.sqrt(err[b, h] / cnt[b, h])) < 1e-6
        # the pose of that trial is the oracle's fit of the same sample (numpy SVD in fp64)
        pick = orc.ransac_draws(1234, b, T, ransac_n, C)[h]
        rows = match[match[:, 0] == b]
        R, t, valid = orc.rigid_fit(src[b][rows[pick, 1]].numpy()[None], tgt[b][rows[pick, 2]].numpy()[None])
        assert valid[0]
        P = got["pose"][b].cpu()
        assert rot_angle(P[:3, :3], torch.from_numpy(R[0])).item() < 1e-5
        assert (P[:3, 3] - torch.from_numpy(t[0])).abs().max().item() < 1e-5
        assert torch.equal(P[3], torch.tensor([0.0, 0.0, 0.0, 1.0]))
        # and it recovers the planted motion to the noise level
        assert rot_angle(P[:3, :3], Rgt[b]).item() < 0.05
        assert (P[:3, 3] - tgt_t[b]).abs().max().item() < 0.05


def test_full_size_recovers_the_planted_pose_and_is_reproducible():
    from diffreg_b200 import ops
    src, tgt, match, Rgt, tgt_t = _problem(5, 2, 4096, 4096, [3000, 2500], 0.25)
    a = ops.ransac_correspondence(src.cuda(), tgt.cuda(), match.cuda(), 0.05, 3, 50000, seed=7)
    b = ops.ransac_correspondence(src.cuda(), tgt.cuda(), match.cuda(), 0.05, 3, 50000, seed=7)
    for k in a:
        assert torch.equal(a[k], b[k]), k
    c = ops.ransac_correspondence(src.cuda(), tgt.cuda(), match.cuda(), 0.05, 3, 50000, seed=8)
    assert not torch.equal(a["best_trial"], c["best_trial"])
    for out in (a, c):
        P = out["pose"].cpu()
        assert (out["fitness"].cpu() > 0.2).all()
        for e in range(2):
            assert rot_angle(P[e, :3, :3], Rgt[e]).item() < 0.03
            assert (P[e, :3, 3] - tgt_t[e]).abs().max().item() < 0.03
        # proper rotations
        assert torch.allclose(torch.linalg.det(P[:, :3, :3].double()), torch.ones(2, dtype=torch.float64), atol=1e-5)


def test_few_or_no_correspondences_give_the_identity():
    """loss.py:384-387: fewer than 3 matches -> eye / zeros; also a batch element with no rows at all and an empty match list."""
    from diffreg_b200 import ops, registration
    src, tgt, match, Rgt, tgt_t = _problem(3, 4, 64, 80, [2, 40, 0, 3], 1.0, noise=0.0)
    out = ops.ransac_correspondence(src.cuda(), tgt.cuda(), match.cuda(), 0.05, 3, 1024, seed=0)
    eye = torch.eye(4)
    assert torch.equal(out["pose"][0].cpu(), eye) and torch.equal(out["pose"][2].cpu(), eye)
    assert out["best_trial"].cpu().tolist()[0] == -1 and out["best_trial"].cpu().tolist()[2] == -1
    assert out["inlier_count"].cpu().tolist()[1] == 40 and out["inlier_count"].cpu().tolist()[3] == 3
    assert rot_angle(out["pose"][1, :3, :3].cpu(), Rgt[1]).item() < 1e-5
    rot, trn = registration.ransac_regist_coarse(src.cuda(), tgt.cuda(), None, None, match.cuda(), max_iteration=1024)
    assert rot.shape == (4, 3, 3) and trn.shape == (4, 3, 1)
    assert torch.equal(rot.cpu(), out["pose"][:, :3, :3].cpu())
    empty = ops.ransac_correspondence(src.cuda(), tgt.cuda(), match[:0].cuda(), 0.05, 3, 256)
    assert torch.equal(empty["pose"].cpu(), eye.expand(4, 4, 4))
    # reference-named single-problem entry point: numpy in, 4 x 4 float64 numpy out
    rows = match[match[:, 0] == 1]
    T = registration.ransac_pose_estimation(src[1].numpy(), tgt[1].numpy(), [rows[:, 1].numpy(), rows[:, 2].numpy()], max_iteration=1024)
    assert T.shape == (4, 4) and T.dtype == np.float64
    assert rot_angle(torch.from_numpy(T[:3, :3]), Rgt[1]).item() < 1e-5


def test_argument_errors():
    from diffreg_b200 import ops
    from diffreg_b200._lib import DiffRegLibraryError
    src, tgt, match, _, _ = _problem(1, 1, 16, 16, [8], 1.0)
    with pytest.raises(DiffRegLibraryError):
        ops.ransac_correspondence(src.cuda(), tgt.cuda(), match.cuda(), 0.05, 2, 64)
    with pytest.raises(DiffRegLibraryError):
        ops.ransac_correspondence(src.cuda(), tgt.cuda(), match.cuda(), 0.0, 3, 64)
    with pytest.raises(Exception):
        ops.ransac_correspondence(src, tgt, match)


def test_explicit_offsets_through_the_c_abi():
    """The C ABI takes the batch offsets from the caller as well (the wrapper lets the library find them): same result."""
    from diffreg_b200 import ops
    from diffreg_b200._lib import check, load_library
    src, tgt, match, _, _ = _problem(21, 3, 500, 500, [120, 0, 300], 0.5)
    src, tgt, match = src.cuda(), tgt.cuda(), match.cuda()
    want = ops.ransac_correspondence(src, tgt, match, 0.05, 3, 2048, seed=9)
    lib = load_library()
    offsets = torch.tensor([0, 120, 120, 420], dtype=torch.int32, device="cuda")
    pose = torch.empty(3, 4, 4, device="cuda")
    fit, rmse = torch.empty(3, device="cuda"), torch.empty(3, device="cuda")
    best, cnt = torch.empty(3, dtype=torch.int32, device="cuda"), torch.empty(3, dtype=torch.int32, device="cuda")
    nbytes = lib.drg_ransac_workspace_bytes(3, 2048)
    ws = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    args = [src.data_ptr(), tgt.data_ptr(), 3, 500, 500, match.data_ptr(), match.shape[0], offsets.data_ptr(), 0.05, 3, 2048, 9,
            pose.data_ptr(), fit.data_ptr(), rmse.data_ptr(), best.data_ptr(), cnt.data_ptr(), None, None, ws.data_ptr(), nbytes,
            torch.cuda.current_stream().cuda_stream]
    check(lib.drg_ransac_correspondence(*args))
    check(lib.drg_ransac_correspondence(*args))  # the tickets reset themselves: a second call on the same workspace
    torch.cuda.synchronize()
    assert torch.equal(pose, want["pose"]) and torch.equal(best, want["best_trial"]) and torch.equal(cnt, want["inlier_count"])
    assert torch.equal(fit, want["fitness"]) and torch.equal(rmse, want["inlier_rmse"])
    assert best.cpu().tolist()[1] == -1
    args[-2] = nbytes - 1
    with pytest.raises(Exception):
        check(lib.drg_ransac_correspondence(*args))


def test_matching_head_to_ransac_pose():
    """The consumer chain of the evaluation (Diff-Reg-3dmatch/lib/tester.py:79-85): Matching.forward -> match_pred ->
    ransac_regist_coarse, all on the device.  Features of corresponding points agree up to noise, a third of the target
    points have no partner: the RANSAC pose is the planted motion to the point noise."""
    import diffreg_b200
    from diffreg_b200 import registration
    B, N, M, C = 2, 1024, 1100, 64
    g = torch.Generator().manual_seed(77)
    src_feats = torch.randn(B, N, C, generator=g)
    s_pcd = torch.rand(B, N, 3, generator=g) * 2 - 1
    tgt_feats = torch.randn(B, M, C, generator=g)
    t_pcd = torch.rand(B, M, 3, generator=g) * 2 - 1
    Rs, ts = [], []
    for b in range(B):
        R, t = orc.random_rotation(g).float(), torch.rand(3, generator=g) - 0.5
        part = torch.randperm(N, generator=g)[: (2 * M) // 3]          # distinct partners for two thirds of the target points
        where = torch.randperm(M, generator=g)[: (2 * M) // 3]
        tgt_feats[b, where] = src_feats[b, part] + 0.2 * torch.randn(len(part), C, generator=g)
        t_pcd[b, where] = s_pcd[b, part] @ R.T + t + 0.004 * torch.randn(len(part), 3, generator=g)
        Rs.append(R)
        ts.append(t)
    cfg = dict(match_type="sinkhorn", confidence_threshold=0.2, feature_dim=C, entangled=True, dsmax_temperature=0.1,
               skh_init_bin_score=1.0, skh_iters=3, skh_prefilter=False)
    head = diffreg_b200.Matching(cfg).cuda().eval()
    with torch.no_grad():
        head.src_proj.weight.copy_(torch.eye(C))                        # identity projection: the similarity is the features' own
    src_mask, tgt_mask = torch.ones(B, N, dtype=torch.bool).cuda(), torch.ones(B, M, dtype=torch.bool).cuda()
    # the head divides both sides by sqrt(C) (matching.py:141): a scale of 3.5 puts partners at a similarity of ~12, strangers at ~N(0, 1.5)
    conf, match_pred = head(3.5 * src_feats.cuda(), 3.5 * tgt_feats.cuda(), None, None, src_mask, tgt_mask, {})
    assert match_pred.shape[1] == 3 and match_pred.shape[0] > 200
    rot, trn = registration.ransac_regist_coarse(s_pcd.cuda(), t_pcd.cuda(), src_mask, tgt_mask, match_pred, seed=3)
    for b in range(B):
        assert rot_angle(rot[b].cpu(), Rs[b]).item() < 0.02
        assert (trn[b, :, 0].cpu() - ts[b]).abs().max().item() < 0.02


def test_out_of_range_point_numbers_do_not_read_outside_the_clouds():
    """A match row pointing outside the point clouds (an IndexError on the host in the reference) is clamped on the device: the
    call completes and the valid correspondences still decide the pose (run under compute-sanitizer memcheck: 0 errors)."""
    from diffreg_b200 import ops
    src, tgt, match, Rgt, _ = _problem(31, 1, 200, 220, [100], 0.8, noise=0.0)
    bad = match.clone()
    bad[3, 1], bad[5, 2], bad[7, 1] = 10 ** 9, -4, 200
    out = ops.ransac_correspondence(src.cuda(), tgt.cuda(), bad.cuda(), 0.05, 3, 2048, seed=2)
    torch.cuda.synchronize()
    assert int(out["inlier_count"][0]) >= 75
    assert rot_angle(out["pose"][0, :3, :3].cpu(), Rgt[0]).item() < 1e-4
