"""Correspondence RANSAC behind the reference's names, computing through libdiffreg_b200.so.

Mirrors Diff-Reg-4dmatch/models/loss.py:13-24 (`ransac_pose_estimation`, an open3d call in the reference) and
loss.py:366-398 (`MatchMotionLoss.ransac_regist_coarse`).  All 50 000 trials of every batch element run in one launch
(`drg_ransac_correspondence`, csrc/ransac.cu) and nothing goes through the host.

Differences a caller can see:
  * the result is a function of (inputs, seed): open3d seeds from random_device, so the reference's own result changes from
    run to run (its tester averages repetitions for that reason, Diff-Reg-3dmatch/lib/tester.py:25);
  * every trial is evaluated (open3d 0.13 may stop early on its confidence estimate; with the reference's criteria
    (50000, 1000) the confidence is clamped to 1 and it does not);
  * `ransac_regist_coarse` returns fp32 tensors on the inputs' device (the reference: CPU, fp64 where open3d produced the
    pose, fp32 identity otherwise) -- callers only compare them with the ground truth.
"""
import numpy as np
import torch

from . import ops


def ransac_pose_estimation(src_pcd, tgt_pcd, corrs, distance_threshold=0.05, ransac_n=3, max_iteration=50000, seed=0,
                           device=None):
    """src_pcd [n,3], tgt_pcd [m,3] (numpy or tensor), corrs = [src_indices, tgt_indices] -> 4 x 4 transformation
    (numpy float64, as open3d's `result.transformation`)."""
    if device is None:
        device = src_pcd.device if torch.is_tensor(src_pcd) and src_pcd.is_cuda else torch.device("cuda", torch.cuda.current_device())
    src = torch.as_tensor(np.asarray(src_pcd) if not torch.is_tensor(src_pcd) else src_pcd, dtype=torch.float32).to(device)[None]
    tgt = torch.as_tensor(np.asarray(tgt_pcd) if not torch.is_tensor(tgt_pcd) else tgt_pcd, dtype=torch.float32).to(device)[None]
    s_ind = torch.as_tensor(np.asarray(corrs[0]) if not torch.is_tensor(corrs[0]) else corrs[0]).to(device=device, dtype=torch.int64)
    t_ind = torch.as_tensor(np.asarray(corrs[1]) if not torch.is_tensor(corrs[1]) else corrs[1]).to(device=device, dtype=torch.int64)
    match = torch.stack([torch.zeros_like(s_ind), s_ind, t_ind], dim=1)
    out = ops.ransac_correspondence(src, tgt, match, distance_threshold, ransac_n, max_iteration, seed)
    return out["pose"][0].double().cpu().numpy()


def ransac_regist_coarse(batched_src_pcd, batched_tgt_pcd, src_mask, tgt_mask, match_pred, distance_threshold=0.05, ransac_n=3,
                         max_iteration=50000, seed=0):
    """loss.py:366-398: (rot [B,3,3], trn [B,3,1]) from match_pred [C,3] rows (b, i, j); fewer than 3 matches -> identity.
    The masks only trim the padded tails in the reference (match_pred never points into them); they are accepted and unused."""
    del src_mask, tgt_mask
    out = ops.ransac_correspondence(batched_src_pcd, batched_tgt_pcd, match_pred, distance_threshold, ransac_n, max_iteration, seed)
    pose = out["pose"]
    return pose[:, :3, :3].contiguous(), pose[:, :3, 3:].contiguous()
