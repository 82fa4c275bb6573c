"""Drop-in replacements for the reference's matching modules, computing through libdiffreg_b200.so.

Mirrors (same names, arguments, return structure, state_dict keys):
    log_optimal_transport          Diff-Reg-4dmatch/models/matching.py:6-38
    mutual_topk_select             Diff-Reg-2d3d/vision3d/ops/mutual_topk_select.py:7-60
    Matching (3D flavour)          Diff-Reg-4dmatch/models/matching.py:41-173, Diff-Reg-3dmatch/models/matching.py:96-283
    Matching2D3D (2D-3D flavour)   Diff-Reg-2d3d/experiments/<exp>/matching.py:41-147

Forward only: the kernels have no backward, so calls with autograd-tracked inputs raise instead of
silently detaching.  There is no CPU path: CPU tensors raise.
"""
import torch
import torch.nn as nn

from . import ops
from ._lib import DiffRegLibraryError


_FORWARD_ONLY = ("diffreg_b200 kernels are forward-only: call under torch.no_grad() with the module in eval() mode "
                 "(training keeps the reference modules)")


def _no_grad_inputs(*tensors, module=None):
    """The autograd rule of SURVEY.md 8b: never detach silently.  Must run BEFORE grad mode is switched off, so the
    public entry points call it first and only then enter torch.no_grad() (a @torch.no_grad() decorator would make
    this check dead code).  Raises when autograd is recording and (a) an input requires grad, or (b) the module is in
    training mode with trainable parameters -- the reference would return a differentiable result in both cases."""
    if not torch.is_grad_enabled():
        return
    for t in tensors:
        if torch.is_tensor(t) and t.requires_grad:
            raise DiffRegLibraryError(_FORWARD_ONLY)
    if module is not None and module.training and any(q.requires_grad for q in module.parameters()):
        raise DiffRegLibraryError(_FORWARD_ONLY)


class _LogOptimalTransportFn(torch.autograd.Function):
    """log_optimal_transport with a CUDA backward: forward = the Sinkhorn kernels, backward =
    drg_sinkhorn_backward (2 I + 1 passes over the scores instead of the ~60 torch's autograd makes through the unrolled
    logsumexp recursion of matching.py:30-32)."""

    @staticmethod
    def forward(ctx, scores, alpha, iters, src_mask, tgt_mask):
        u_all, v_all, out = ops.sinkhorn_potentials_per_iteration(scores.detach(), alpha.detach().reshape(()).float(), iters, src_mask,
                                                                  tgt_mask, want_out=True)
        ctx.save_for_backward(scores, alpha, src_mask, tgt_mask, u_all, v_all)     # the potentials after every iteration: O(I (N + M))
        ctx.iters = int(iters)
        return out

    @staticmethod
    def backward(ctx, grad_out):
        scores, alpha, src_mask, tgt_mask, u_all, v_all = ctx.saved_tensors
        gs, ga = ops.sinkhorn_backward(scores, alpha, ctx.iters, src_mask, tgt_mask, grad_out.contiguous(), potentials=(u_all, v_all))
        return (gs.to(scores.dtype) if ctx.needs_input_grad[0] else None,
                ga.to(alpha.dtype).reshape(alpha.shape) if ctx.needs_input_grad[1] else None, None, None, None)


class _DualSoftmaxFn(torch.autograd.Function):
    """The dual-softmax confidence (matching.py:147-157) with a CUDA backward (drg_dual_softmax_backward)."""

    @staticmethod
    def forward(ctx, sim, src_mask, tgt_mask, temperature):
        ctx.save_for_backward(sim, src_mask, tgt_mask)
        ctx.temperature = float(temperature)
        return ops.dual_softmax(sim, src_mask, tgt_mask, temperature)

    @staticmethod
    def backward(ctx, grad_conf):
        sim, src_mask, tgt_mask = ctx.saved_tensors
        return ops.dual_softmax_backward(sim, src_mask, tgt_mask, ctx.temperature, grad_conf.contiguous()), None, None, None


def log_optimal_transport(scores, alpha, iters, src_mask, tgt_mask):
    """[B,N,M] scores -> [B,N+1,M+1] log-assignment (Z + u + v - norm).

    Computed in fp32; an fp64 `scores` (the reference's fp64 sampler state) gives an
    fp64 result holding the fp32-accurate values.  Differentiable with respect to `scores` and `alpha` (fp32, CUDA): when
    autograd is recording and either requires grad, the result carries a CUDA backward (_LogOptimalTransportFn)."""
    if torch.is_grad_enabled() and (scores.requires_grad or (torch.is_tensor(alpha) and alpha.requires_grad)):
        if scores.dtype != torch.float32:
            raise DiffRegLibraryError("log_optimal_transport: the differentiable path takes fp32 scores")
        return _LogOptimalTransportFn.apply(scores, alpha, iters, src_mask, tgt_mask)
    with torch.no_grad():
        out = ops.sinkhorn(scores, alpha, iters, src_mask, tgt_mask, out_mode="log_full")
        return out.to(scores.dtype) if scores.dtype == torch.float64 else out


def mutual_topk_select(score_mat, k, largest=True, threshold=None, mutual=True, reduce_result=True):
    """Mutual top-k selection on a 2-D score matrix (vision3d/ops/mutual_topk_select.py:7-60; 3d models/matching.py:6-59).
    -> (row_indices [K], col_indices [K], scores [K]) in row-major order, or the [N,M] bool matrix if not reduce_result.
    k = 1 (the sampler's final selection: 3d pipeline.py:275-277, 2d3d model.py:692-694, matching.py:134-136) runs the
    single-pass arg-max kernels; k > 1 (k <= 8) the general top-k kernels."""
    # the reference's gathered scores are differentiable: with a tracked score_mat the selection runs on the detached matrix and
    # the scores are gathered from the tracked one
    tracked = torch.is_grad_enabled() and score_mat.requires_grad and score_mat.is_cuda
    if not tracked:
        _no_grad_inputs(score_mat)
    with torch.no_grad():
        sm = score_mat.detach()
        if k == 1 and reduce_result:
            rows, cols, vals = ops.top1_select(sm, largest, threshold, mutual)
        else:
            index, vals, mask = ops.topk_select(sm.unsqueeze(0), k, largest, threshold, mutual, want_mask=not reduce_result)
            if not reduce_result:
                return mask.squeeze(0)
            rows, cols = index[:, 1].contiguous(), index[:, 2].contiguous()
    return rows, cols, (score_mat[rows, cols] if tracked else vals)


def batch_mutual_topk_select(score_mat, k, row_masks=None, col_masks=None, largest=True, threshold=None, mutual=True,
                             reduce_result=True):
    """Batched mutual top-k selection (vision3d/ops/mutual_topk_select.py:63-133; the 2D-3D fine matching, model.py:738-746).
    score_mat [B,N,M] -> (batch_indices, row_indices, col_indices, scores), or the [B,N,M] bool matrix."""
    _no_grad_inputs(score_mat)
    with torch.no_grad():
        index, vals, mask = ops.topk_select(score_mat, k, largest, threshold, mutual, row_masks, col_masks,
                                            want_mask=not reduce_result)
        if reduce_result:
            return index[:, 0].contiguous(), index[:, 1].contiguous(), index[:, 2].contiguous(), vals
        return mask


class Matching(nn.Module):
    """3D flavour.  `precision`: '3xtf32' (default; the fp32-accurate tensor-core GEMM: three-term hi/lo split products --
    since round 2 on row-scaled fp16 split operands instead of tf32 ones, same accuracy) or 'tf32' (one kind::tf32 pass)."""

    def __init__(self, config, precision="3xtf32"):
        super().__init__()
        self.match_type = config['match_type']
        self.confidence_threshold = config['confidence_threshold']
        d_model = config['feature_dim']
        self.src_proj = nn.Linear(d_model, d_model, bias=False)
        self.tgt_proj = nn.Linear(d_model, d_model, bias=False)   # unused by forward, kept for checkpoints (matching.py:53)
        self.entangled = config['entangled']
        if self.match_type == "dual_softmax":
            self.temperature = config['dsmax_temperature']
        elif self.match_type == 'sinkhorn':
            self.skh_init_bin_score = config['skh_init_bin_score']
            self.skh_iters = config['skh_iters']
            self.skh_prefilter = config['skh_prefilter']
            self.bin_score = nn.Parameter(torch.tensor(self.skh_init_bin_score, requires_grad=True))
        else:
            raise NotImplementedError()
        if precision not in ("3xtf32", "tf32"):
            raise ValueError(precision)
        self.precision = precision
        self._w_cache = None

    # ---- correspondence extraction (static, like the reference) ----
    @staticmethod
    def get_match(conf_matrix, thr=0.0, mutual=True):
        """(index [K,3], mconf [K], mask).  With a tracked conf_matrix the gathered mconf = conf[index] is differentiable, as
        the reference's is (matching.py:86-87): the selection runs on the detached matrix, the gather on the tracked one."""
        tracked = torch.is_grad_enabled() and conf_matrix.requires_grad and conf_matrix.is_cuda
        if not tracked:
            _no_grad_inputs(conf_matrix)
        with torch.no_grad():
            index, mconf, mask = ops.get_match(conf_matrix.detach(), thr, mutual, want_mask=True)
        if tracked:
            mconf = conf_matrix[index[:, 0], index[:, 1], index[:, 2]]
        return index, mconf, mask

    @staticmethod
    def get_topk_match(conf_matrix, thr, mutual=True):
        return Matching.get_match(conf_matrix, thr, mutual)

    # ---- similarity ----
    def _weight_operand(self):
        w = self.src_proj.weight
        key = (w.data_ptr(), w._version, self.precision, w.device)
        if self._w_cache is None or self._w_cache[0] != key:
            split = self.precision == "3xtf32"
            self._w_cache = (key, ops.prep_operand(w.detach(), 1.0, split, 1))
        return self._w_cache[1]

    def project(self, feats):
        """src_proj applied through the tensor-core GEMM: [B,L,C] -> [B,L,C]."""
        B, L, C = feats.shape
        split = self.precision == "3xtf32"
        a = ops.prep_operand(feats, 1.0, split, 0)
        out = ops.gemm_nt(a.reshape(B * L, a.shape[-1]), self._weight_operand(), split3=split, K=feats.shape[-1])
        return out.view(B, L, self.src_proj.weight.shape[0])

    def similarity(self, src_feats, tgt_feats, src_pe=None, tgt_pe=None, pe_type="rotary", data=None):
        """Projection (same weight on both sides, matching.py:127-128), optional positional embedding,
        1/sqrt(C) scaling and the N x M contraction.  Returns sim [B,N,M]."""
        C = self.src_proj.weight.shape[0]
        split = self.precision == "3xtf32"
        scale = 1.0 / (C ** .5)
        use_pe = (not self.entangled) and src_pe is not None
        if split and not use_pe:
            # fast path: one split launch for both feature sets, one projection GEMM whose epilogue already writes the
            # scaled, split operands of the similarity GEMM (3 launches in all)
            B, N, _ = src_feats.shape
            a, b, plain = ops.project_pair_split(src_feats, tgt_feats, self._weight_operand(), C, scale, want_plain=data is not None)
            if data is not None:
                fs, ft = plain[:B * N].view(B, N, C), plain[B * N:].view(B, tgt_feats.shape[1], C)
                data["src_feats_nopos"] = fs
                data["tgt_feats_nopos"] = ft
                data["src_feats"] = fs
                data["tgt_feats"] = ft
            return ops.gemm_nt(a, b, split3=True, K=C)
        fs = self.project(src_feats)
        ft = self.project(tgt_feats)
        want = data is not None and use_pe
        a = ops.prep_operand(fs, scale, split, 0, pe=src_pe if use_pe else None, pe_type=pe_type if use_pe else None,
                             want_embedded=want)
        b = ops.prep_operand(ft, scale, split, 1, pe=tgt_pe if use_pe else None, pe_type=pe_type if use_pe else None,
                             want_embedded=want)
        if data is not None:
            data["src_feats_nopos"] = fs
            data["tgt_feats_nopos"] = ft
            data["src_feats"] = a[1] if want else fs
            data["tgt_feats"] = b[1] if want else ft
        if want:
            a, b = a[0], b[0]
        return ops.gemm_nt(a, b, split3=split, K=C)

    def confidence(self, sim, src_mask, tgt_mask):
        B, N, M = sim.shape
        if src_mask is None:
            src_mask = torch.ones(B, N, dtype=torch.bool, device=sim.device)
            tgt_mask = torch.ones(B, M, dtype=torch.bool, device=sim.device)
        if self.match_type == "dual_softmax":
            return ops.dual_softmax(sim, src_mask, tgt_mask, self.temperature)
        return ops.sinkhorn(sim, self.bin_score, self.skh_iters, src_mask, tgt_mask, out_mode="conf", apply_mask=True)

    def _records_grad(self, *tensors):
        return torch.is_grad_enabled() and (any(torch.is_tensor(t) and t.requires_grad for t in tensors) or
                                            (self.training and any(q.requires_grad for q in self.parameters())))

    def _conf_train(self, src_feats, tgt_feats, src_pe, tgt_pe, src_mask, tgt_mask, data, pe_type):
        """conf_matrix of either branch with autograd recording (matching.py:118-173).  The
        projections, the position code and the similarity contraction are torch ops (library GEMMs and elementwise kernels,
        differentiable as they stand); the Sinkhorn -- where autograd would make ~60 passes over the N x M matrix forward and
        backward -- runs on this library's kernels in both directions (_LogOptimalTransportFn)."""
        fs = torch.nn.functional.linear(src_feats, self.src_proj.weight)
        ft = torch.nn.functional.linear(tgt_feats, self.src_proj.weight)            # the same weight on both sides (:127-128)
        data.update({"src_feats_nopos": fs, "tgt_feats_nopos": ft})
        if not self.entangled and src_pe is not None:                              # :135-137
            def embed(x, pe):
                if pe_type == "rotary":
                    x2 = torch.stack([-x[..., 1::2], x[..., ::2]], dim=-1).reshape_as(x)
                    return x * pe[..., 0] + x2 * pe[..., 1]
                if pe_type == "sinusoidal":
                    return x + pe
                raise KeyError(pe_type)
            fs, ft = embed(fs, src_pe), embed(ft, tgt_pe)
        data.update({"src_feats": fs, "tgt_feats": ft})
        scale = fs.shape[-1] ** .5
        sim = torch.einsum("bsc,btc->bst", fs / scale, ft / scale)                   # :144-145, :161
        if self.match_type == "dual_softmax":                                      # :147-157 (temperature and masks inside the kernels)
            if src_mask is None:
                src_mask = torch.ones(sim.shape[:2], dtype=torch.bool, device=sim.device)
                tgt_mask = torch.ones(sim.shape[0], sim.shape[2], dtype=torch.bool, device=sim.device)
            return _DualSoftmaxFn.apply(sim, src_mask, tgt_mask, self.temperature)
        if src_mask is not None:
            sim = sim.masked_fill(~(src_mask[..., None] * tgt_mask[:, None]).bool(), float("-inf"))   # :163-165
        else:
            src_mask = torch.ones(sim.shape[:2], dtype=torch.bool, device=sim.device)
            tgt_mask = torch.ones(sim.shape[0], sim.shape[2], dtype=torch.bool, device=sim.device)
        log_assign = log_optimal_transport(sim, self.bin_score, self.skh_iters, src_mask, tgt_mask)
        return log_assign.exp()[:, :-1, :-1].contiguous()                          # :169-170

    def forward(self, src_feats, tgt_feats, src_pe, tgt_pe, src_mask, tgt_mask, data, pe_type="rotary"):
        """-> (conf_matrix [B,N,M], coarse_match [K,3] int64); writes the four feature tensors into `data`.  When autograd is
        recording (training mode / inputs that require grad) the Sinkhorn branch returns a differentiable conf_matrix
        (_conf_train: both branches; likewise forward1 and the 2D-3D head's Sinkhorn branch)."""
        if src_feats.is_cuda and self._records_grad(src_feats, tgt_feats, src_pe, tgt_pe):
            conf_matrix = self._conf_train(src_feats, tgt_feats, src_pe, tgt_pe, src_mask, tgt_mask, data, pe_type)
            with torch.no_grad():
                coarse_match, _, _ = ops.get_match(conf_matrix.detach(), self.confidence_threshold, True, want_mask=False)
            return conf_matrix, coarse_match
        _no_grad_inputs(src_feats, tgt_feats, src_pe, tgt_pe, module=self)
        with torch.no_grad():
            sim = self.similarity(src_feats, tgt_feats, src_pe, tgt_pe, pe_type, data)
            conf_matrix = self.confidence(sim, src_mask, tgt_mask)
            coarse_match, _, _ = ops.get_match(conf_matrix, self.confidence_threshold, True, want_mask=False)
        return conf_matrix, coarse_match

    def forward1(self, src_feats, tgt_feats, src_pe, tgt_pe, src_mask, tgt_mask, data, pe_type="rotary", mutual=False):
        """3DMatch variant (3d matching.py:221-283): top-1 row/column matches as [K,3] with a zero batch column."""
        if src_feats.is_cuda and self._records_grad(src_feats, tgt_feats, src_pe, tgt_pe):
            conf_matrix = self._conf_train(src_feats, tgt_feats, src_pe, tgt_pe, src_mask, tgt_mask, data, pe_type)
            with torch.no_grad():
                r, c, _ = ops.top1_select(conf_matrix.detach().squeeze(0), True, None, mutual)
                coarse_match = torch.cat([torch.zeros_like(r).unsqueeze(-1), r.unsqueeze(-1), c.unsqueeze(-1)], dim=-1)
            return conf_matrix, coarse_match
        _no_grad_inputs(src_feats, tgt_feats, src_pe, tgt_pe, module=self)
        with torch.no_grad():
            sim = self.similarity(src_feats, tgt_feats, src_pe, tgt_pe, pe_type, data)
            conf_matrix = self.confidence(sim, src_mask, tgt_mask)
            r, c, _ = ops.top1_select(conf_matrix.squeeze(0), True, None, mutual)
            coarse_match = torch.cat([torch.zeros_like(r).unsqueeze(-1), r.unsqueeze(-1), c.unsqueeze(-1)], dim=-1)
        return conf_matrix, coarse_match


class Matching2D3D(Matching):
    """2D-3D flavour: no positional arguments, no `data`; returns top-1 selected correspondences."""

    def __init__(self, config, mutual=True, precision="3xtf32"):
        super().__init__(config, precision)
        self.mutual = mutual

    def forward(self, src_feats, tgt_feats, src_mask, tgt_mask, mutual=True):
        """-> (conf_matrix [1,N,M], src_indices [K], tgt_indices [K], weights [K]); differentiable (conf_matrix and the gathered
        weights, as in the reference: matching.py:122-136) when autograd is recording and match_type is sinkhorn."""
        if self.match_type == "sinkhorn" and src_feats.is_cuda and self._records_grad(src_feats, tgt_feats):
            conf_matrix = self._conf_train(src_feats, tgt_feats, None, None, src_mask, tgt_mask, {}, None)
            with torch.no_grad():
                src_indices, tgt_indices, _ = ops.top1_select(conf_matrix.detach().squeeze(0), True, None, mutual)
            return conf_matrix, src_indices, tgt_indices, conf_matrix[0][src_indices, tgt_indices]
        _no_grad_inputs(src_feats, tgt_feats, module=self)
        with torch.no_grad():
            sim = self.similarity(src_feats, tgt_feats)
            conf_matrix = self.confidence(sim, src_mask, tgt_mask)
            if self.match_type != "sinkhorn":
                # the reference only defines the selection inside its sinkhorn branch (matching.py:134-136)
                raise NotImplementedError("the 2D-3D head selects correspondences in the sinkhorn branch only")
            src_indices, tgt_indices, weights = ops.top1_select(conf_matrix.squeeze(0), True, None, mutual)
        return conf_matrix, src_indices, tgt_indices, weights
