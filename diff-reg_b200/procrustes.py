"""Drop-in replacement for the reference's SoftProcrustesLayer, computing through libdiffreg_b200.so.

Mirrors Diff-Reg-4dmatch/models/procrustes.py:10-93 (identical copy under Diff-Reg-2d3d/experiments/<exp>/);
`SoftProcrustesLayer3DMatch` is the Diff-Reg-3dmatch variant that uses the padded lengths
(Diff-Reg-3dmatch/models/procrustes.py:61-62).

Differences a caller can see, both documented in DESIGN.md:
  * `condition` and `solution_mask` live on the device (the reference leaves them on the CPU after its
    host SVD); callers only index with them.
  * nothing synchronises with the host: no `.cpu()` SVD, no tensor-valued slice bound.
"""
import torch
import torch.nn as nn

from . import ops
from .matching import _no_grad_inputs


class SoftProcrustesLayer(nn.Module):
    padded_lengths = False

    def __init__(self, config):
        super().__init__()
        self.sample_rate = config.sample_rate
        self.max_condition_num = config.max_condition_num

    @staticmethod
    def batch_weighted_procrustes(X, Y, w, eps=0.0001):
        """X, Y [B,K,3], w [B,K,1] -> (R [B,3,3], t [B,3,1], condition [B] fp64)"""
        _no_grad_inputs(X, Y, w)
        with torch.no_grad():
            return ops.weighted_procrustes(X, Y, w, eps)

    def forward(self, conf_matrix, src_pcd, tgt_pcd, src_mask, tgt_mask):
        """-> (R, t, R_forwd, t_forwd, condition, solution_mask)"""
        _no_grad_inputs(conf_matrix, src_pcd, tgt_pcd)
        with torch.no_grad():
            o = ops.soft_procrustes(conf_matrix, src_pcd, tgt_pcd, src_mask, tgt_mask, self.sample_rate, self.max_condition_num,
                                    padded_lengths=self.padded_lengths)
        return o["R"], o["t"], o["R_forwd"], o["t_forwd"], o["condition"], o["solution_mask"]

    def forward_warp(self, conf_matrix, src_pcd, tgt_pcd, src_mask, tgt_mask):
        """forward() plus the source points moved by the gated pose (pipeline.py:218-220) in the same kernel.
        -> (src_pcd_wrapped [B,N,3], pose 6-tuple)"""
        _no_grad_inputs(conf_matrix, src_pcd, tgt_pcd)
        with torch.no_grad():
            o = ops.soft_procrustes(conf_matrix, src_pcd, tgt_pcd, src_mask, tgt_mask, self.sample_rate, self.max_condition_num,
                                    padded_lengths=self.padded_lengths, want_warped=True)
        return o["src_warped"], (o["R"], o["t"], o["R_forwd"], o["t_forwd"], o["condition"], o["solution_mask"])


class SoftProcrustesLayer3DMatch(SoftProcrustesLayer):
    padded_lengths = True
