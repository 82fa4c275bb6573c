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


class _WeightedProcrustesFn(torch.autograd.Function):
    """batch_weighted_procrustes, differentiable with respect to the weights: forward = the Kabsch kernel,
    backward = drg_weighted_procrustes_backward (no SVD backwards: a 3 x 3 linear solve on the symmetric factor R^T M)."""

    @staticmethod
    def forward(ctx, X, Y, w, eps):
        R, t, cond = ops.weighted_procrustes(X, Y, w, eps)
        ctx.save_for_backward(X, Y, w, R)
        ctx.eps = float(eps)
        ctx.mark_non_differentiable(cond)
        return R, t, cond

    @staticmethod
    def backward(ctx, grad_R, grad_t, _grad_cond):
        X, Y, w, R = ctx.saved_tensors
        gw = ops.weighted_procrustes_backward(X, Y, w, R, grad_R, grad_t, ctx.eps)
        return None, None, gw.to(w.dtype), None


class SoftProcrustesLayer(nn.Module):
    padded_lengths = False

    def __init__(self, config):
        super().__init__()
        self.sample_rate = config.sample_rate
        self.max_condition_num = config.max_condition_num

    @staticmethod
    def batch_weighted_procrustes(X, Y, w, eps=0.0001):
        """X, Y [B,K,3], w [B,K,1] -> (R [B,3,3], t [B,3,1], condition [B] fp64)"""
        if torch.is_grad_enabled() and w.requires_grad and w.is_cuda and not (X.requires_grad or Y.requires_grad):
            return _WeightedProcrustesFn.apply(X, Y, w, eps)       # differentiable in the weights (the points are data)
        _no_grad_inputs(X, Y, w)
        with torch.no_grad():
            return ops.weighted_procrustes(X, Y, w, eps)

    def _forward_train(self, conf_matrix, src_pcd, tgt_pcd, src_mask, tgt_mask):
        """forward() with autograd recording through the weights (procrustes.py:48-93): the top-K selection is index work on the
        detached matrix (torch.topk; the selected entries are then gathered from the tracked matrix), the solve is
        _WeightedProcrustesFn."""
        B, N, M = conf_matrix.shape
        if self.padded_lengths:                                                 # 3d/models/procrustes.py:61-62
            entry_max = torch.full((B,), int(max(N, M) * self.sample_rate), dtype=torch.int32, device=conf_matrix.device)
        else:
            entry_max = (torch.stack([src_mask.sum(dim=1), tgt_mask.sum(dim=1)], dim=0).max(dim=0)[0] * self.sample_rate).int()
        K = int(entry_max.float().mean().int())
        flat = conf_matrix.view(B, -1)
        idx = torch.topk(flat.detach(), K, dim=1, largest=True, sorted=True)[1]
        w = flat.gather(1, idx)
        bi = torch.arange(B, device=idx.device).view(-1, 1).expand(-1, K)
        X, Y = src_pcd[bi, idx // M], tgt_pcd[bi, idx % M]
        w = w * (torch.arange(K, device=w.device).view(1, -1) < entry_max[:, None]).to(w.dtype)      # w[~w_mask] = 0 (:74-76)
        R, t, condition = self.batch_weighted_procrustes(X.detach(), Y.detach(), w[..., None])
        solution_mask = condition < self.max_condition_num
        eye, zero = torch.eye(3, dtype=R.dtype, device=R.device), torch.zeros(3, 1, dtype=R.dtype, device=R.device)
        R_forwd = torch.where(solution_mask[:, None, None], R, eye)
        t_forwd = torch.where(solution_mask[:, None, None], t, zero)
        return R, t, R_forwd, t_forwd, condition, solution_mask

    def forward(self, conf_matrix, src_pcd, tgt_pcd, src_mask, tgt_mask):
        """-> (R, t, R_forwd, t_forwd, condition, solution_mask); differentiable with respect to conf_matrix when autograd is
        recording (the points are data: tracked points raise)."""
        if torch.is_grad_enabled() and conf_matrix.requires_grad and conf_matrix.is_cuda and not (src_pcd.requires_grad or tgt_pcd.requires_grad):
            return self._forward_train(conf_matrix, src_pcd, tgt_pcd, src_mask, tgt_mask)
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
