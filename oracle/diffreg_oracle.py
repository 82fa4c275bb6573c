"""CPU oracle for the Diff-Reg denoising hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (torch CPU ops, fp32 unless the reference itself
promotes to fp64) of the reference's coarse matching-matrix update.  It exists
so that the CUDA path in ``diff-reg_b200/`` can be checked on a box that has no
copy of the reference.  Only ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it; the
product path never does (it raises when the CUDA library is missing).

Parity pinning.  The reference ships NO golden vectors, known-answer tests or
fixtures for this path, so the reference's own tests pin
nothing: "parity unpinned" by the reference's tests.  What pins this oracle
instead is the reference code itself, imported unmodified from /root/reference
in the build container by ``tests/golden/make_golden.py``; the inputs and the
reference's outputs are committed as ``tests/golden/*.npz`` and
``tests/test_oracle_golden.py`` checks every function here against them.

Reference files restated (paths relative to the reference checkout):
  4d  = Diff-Reg-4dmatch/   3d = Diff-Reg-3dmatch/
  2d3d = Diff-Reg-2d3d/experiments/2d3dmatr.rgbdv2.stage4.level3.stage1/

The arithmetic deliberately follows the reference's operation ORDER (cat-built
score matrix, ``logsumexp`` of a materialised sum, full descending sort, fp64
SVD on the host) because the oracle doubles as the CPU timing baseline ("port")
and must pay for the same passes the reference pays for.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import torch

NEG_INF = float("-inf")


# --------------------------------------------------------------------------------------
# Sinkhorn (a4)
# --------------------------------------------------------------------------------------
def log_optimal_transport(scores, alpha, iters, src_mask, tgt_mask):
    """Log-domain Sinkhorn with one dustbin row and one dustbin column.

    Follows 4d/models/matching.py:6-38 (byte-identical copies at
    3d/models/matching.py:61-93 and 2d3d/matching.py:6-38).

    scores : [B, N, M]   may contain -inf (padding) and may be fp64 (quirk Q4)
    alpha  : 0-dim tensor, the shared dustbin score
    masks  : [B, N] / [B, M] bool; only their COUNTS are used here (Q1: padded
             rows/columns keep marginal mass ``norm`` and drain into the dustbin)
    returns: [B, N+1, M+1] log-assignment  (Z + u + v - norm)
    """
    B, N, M = scores.shape
    n_src = src_mask.sum(dim=1, keepdim=True)          # :14  int64 [B,1]
    n_tgt = tgt_mask.sum(dim=1, keepdim=True)          # :15
    a = alpha                                          # cat() promotes fp32 alpha with fp64 scores (Q4)

    # :17-22  augmented score matrix, dustbins hold alpha
    top = torch.cat((scores, a.expand(B, N, 1)), dim=2)
    bottom = a.expand(B, 1, M + 1)
    Z = torch.cat((top, bottom), dim=1)

    norm = -(n_src + n_tgt).log()                      # :24  fp32 [B,1]
    log_mu = torch.cat((norm.expand(B, N), n_tgt.log() + norm), dim=1)   # :26
    log_nu = torch.cat((norm.expand(B, M), n_src.log() + norm), dim=1)   # :27

    u = torch.zeros_like(log_mu)
    v = torch.zeros_like(log_nu)
    for _ in range(int(iters)):                        # :30-32
        u = log_mu - torch.logsumexp(Z + v[:, None, :], dim=2)
        v = log_nu - torch.logsumexp(Z + u[:, :, None], dim=1)
    out = Z + u[:, :, None] + v[:, None, :]            # :34
    return out - norm.view(-1, 1, 1)                   # :36


def sinkhorn_potentials(scores, alpha, iters, src_mask, tgt_mask):
    """Same recursion as ``log_optimal_transport`` but returns (u, v, norm).

    Not a reference function: it exposes the dual potentials the CUDA library
    also returns, so that they can be checked directly.
    """
    B, N, M = scores.shape
    n_src = src_mask.sum(dim=1, keepdim=True)
    n_tgt = tgt_mask.sum(dim=1, keepdim=True)
    Z = torch.cat((torch.cat((scores, alpha.expand(B, N, 1)), dim=2), alpha.expand(B, 1, M + 1)), dim=1)
    norm = -(n_src + n_tgt).log()
    log_mu = torch.cat((norm.expand(B, N), n_tgt.log() + norm), dim=1)
    log_nu = torch.cat((norm.expand(B, M), n_src.log() + norm), dim=1)
    u = torch.zeros_like(log_mu)
    v = torch.zeros_like(log_nu)
    for _ in range(int(iters)):
        u = log_mu - torch.logsumexp(Z + v[:, None, :], dim=2)
        v = log_nu - torch.logsumexp(Z + u[:, :, None], dim=1)
    return u, v, norm


def log_optimal_transport_backward(scores, alpha, iters, src_mask, tgt_mask, grad_out):
    """(dL/d scores, dL/d alpha) of ``log_optimal_transport`` for a given dL/d out -- what torch's autograd returns for
    ``loss.backward()`` through 4d/models/matching.py:6-38, written out so that the CUDA backward can be
    checked in fp64 without autograd.  With Pc_t = exp(Z + u_t + v_t - log_nu) (column softmax of the v_t step) and
    Pr_t = exp(Z + v_{t-1} + u_t - log_mu) (row softmax of the u_t step):
        gv = colsum(G); gu = rowsum(G)
        for t = I..1:  gZ -= gv Pc_t;  gu -= sum_j gv Pc_t;   gZ -= gu Pr_t;  gv = -sum_i gu Pr_t;  gu = 0
    Checked against the autograd of the unmodified reference (tests/test_oracle_golden.py, golden ``lotb_*``)."""
    B, N, M = scores.shape
    dt = scores.dtype
    n_src = src_mask.sum(dim=1, keepdim=True)
    n_tgt = tgt_mask.sum(dim=1, keepdim=True)
    Z = torch.cat((torch.cat((scores, alpha.expand(B, N, 1)), dim=2), alpha.expand(B, 1, M + 1)), dim=1)
    norm = -(n_src + n_tgt).log().to(dt)
    log_mu = torch.cat((norm.expand(B, N), n_tgt.log().to(dt) + norm), dim=1)
    log_nu = torch.cat((norm.expand(B, M), n_src.log().to(dt) + norm), dim=1)
    us, vs = [], [torch.zeros_like(log_nu)]
    u, v = torch.zeros_like(log_mu), vs[0]
    for _ in range(int(iters)):
        u = log_mu - torch.logsumexp(Z + v[:, None, :], dim=2)
        v = log_nu - torch.logsumexp(Z + u[:, :, None], dim=1)
        us.append(u)
        vs.append(v)
    gZ = grad_out.clone()
    gu, gv = grad_out.sum(dim=2), grad_out.sum(dim=1)
    for t in range(int(iters), 0, -1):
        u_t, v_t, v_p = us[t - 1], vs[t], vs[t - 1]
        Pc = torch.exp(Z + u_t[:, :, None] + v_t[:, None, :] - log_nu[:, None, :])
        gZ = gZ - gv[:, None, :] * Pc
        gu = gu - (gv[:, None, :] * Pc).sum(dim=2)
        Pr = torch.exp(Z + v_p[:, None, :] + u_t[:, :, None] - log_mu[:, :, None])
        gZ = gZ - gu[:, :, None] * Pr
        gv = -(gu[:, :, None] * Pr).sum(dim=1)
        gu = torch.zeros_like(gu)
    return gZ[:, :N, :M], gZ[:, N, :].sum() + gZ[:, :N, M].sum()


def dual_softmax_backward(sim, src_mask, tgt_mask, temperature, grad_conf):
    """dL/d sim of the dual-softmax confidence (4d/models/matching.py:147-157) for a given dL/d conf: with A = softmax over the src
    axis (invalid src rows at -inf), B = softmax over the tgt axis (invalid tgt columns at -inf), P = A B, c_j = sum_i G P,
    r_i = sum_j G P:  (2 P G - A c_j - B r_i) / T.  Checked against the reference's autograd (golden ``lotb_matching_train_dual*``)."""
    s1 = (sim / temperature).masked_fill(~src_mask[:, :, None], float("-inf"))
    s2 = (sim / temperature).masked_fill(~tgt_mask[:, None, :], float("-inf"))
    A, Bm = torch.softmax(s1, dim=1), torch.softmax(s2, dim=2)
    P = A * Bm
    c = (grad_conf * P).sum(dim=1, keepdim=True)
    r = (grad_conf * P).sum(dim=2, keepdim=True)
    return (2.0 * P * grad_conf - A * c - Bm * r) / temperature


def pair_mask(src_mask, tgt_mask):
    """[B,N,M] validity mask, as built at 4d/models/matching.py:163-165."""
    return (src_mask[..., None] * tgt_mask[:, None]).bool()


# --------------------------------------------------------------------------------------
# rotary embedding (a11)
# --------------------------------------------------------------------------------------
def embed_rotary(x, cos, sin):
    """4d/models/position_encoding.py:26-35: rotate feature pairs (x0,x1)->(-x1,x0)."""
    rot = torch.empty_like(x)
    rot[..., 0::2] = -x[..., 1::2]
    rot[..., 1::2] = x[..., 0::2]
    return x * cos + rot * sin


def volumetric_position_code(xyz, feature_dim, vol_origin, voxel_size, pe_type="rotary"):
    """``VolumetricPositionEncoding.forward`` 4d/models/position_encoding.py:49-87.
    xyz [B,N,3] -> rotary: [B,N,d,2] = (cos, sin) with every angle duplicated over a feature pair;
    sinusoidal: [B,N,d] = cat(sinx, cosx, siny, cosy, sinz, cosz).  d must be a multiple of 6."""
    B, N, _ = xyz.shape
    origin = torch.as_tensor(vol_origin, dtype=torch.float32, device=xyz.device).view(1, 1, -1)
    vox = (xyz - origin) / voxel_size                                                   # voxelize :16-24
    d3 = feature_dim // 3
    div = torch.exp(torch.arange(0, d3, 2, dtype=torch.float, device=xyz.device) * (-math.log(10000.0) / d3)).view(1, 1, -1)
    parts = []
    for a in range(3):
        ang = vox[..., a:a + 1] * div
        parts.append((torch.sin(ang), torch.cos(ang)))
    if pe_type == "sinusoidal":
        return torch.cat([t for sc in parts for t in sc], dim=-1)
    if pe_type != "rotary":
        raise KeyError(pe_type)
    dup = lambda f: torch.stack([f, f], dim=-1).view(B, N, -1)                          # noqa: E731  theta_k, theta_k
    sin_pos = torch.cat([dup(s_) for s_, _ in parts], dim=-1)
    cos_pos = torch.cat([dup(c_) for _, c_ in parts], dim=-1)
    return torch.stack([cos_pos, sin_pos], dim=-1)


def embed_pos(pe_type, x, pe):
    """``VolumetricPositionEncoding.embed_pos`` 4d/models/position_encoding.py:37-46."""
    if pe_type == "rotary":
        return embed_rotary(x, pe[..., 0], pe[..., 1])
    if pe_type == "sinusoidal":
        return x + pe
    raise KeyError(pe_type)


# --------------------------------------------------------------------------------------
# correspondence extraction (a5, a6)
# --------------------------------------------------------------------------------------
def get_match(conf, thr=0.0, mutual=True):
    """4d/models/matching.py:71-88.  Returns (index [K,3] int64, mconf [K], mask)."""
    keep = conf > thr
    if mutual:
        row_best = conf.max(dim=2, keepdim=True)[0]
        col_best = conf.max(dim=1, keepdim=True)[0]
        keep = keep & (conf == row_best) & (conf == col_best)
    index = keep.nonzero()
    mconf = conf[index[:, 0], index[:, 1], index[:, 2]]
    return index, mconf, keep


def mutual_topk_select(score_mat, k, largest=True, threshold=None, mutual=True, reduce_result=True):
    """2d3d vision3d/ops/mutual_topk_select.py:7-60 (copies 3d/models/matching.py:6-59,
    3d/models/pipeline.py:12-65), with the hard-coded ``.cuda()`` of :34,:39 replaced by
    the score matrix's own device so it runs on CPU."""
    n_rows, n_cols = score_mat.shape
    dev = score_mat.device
    row_pick = score_mat.topk(k=k, largest=largest, dim=1)[1]            # [N,k]
    row_hit = torch.zeros_like(score_mat, dtype=torch.bool)
    row_hit[torch.arange(n_rows, device=dev)[:, None].expand(-1, k), row_pick] = True
    col_pick = score_mat.topk(k=k, largest=largest, dim=0)[1]            # [k,M]
    col_hit = torch.zeros_like(score_mat, dtype=torch.bool)
    col_hit[col_pick, torch.arange(n_cols, device=dev)[None, :].expand(k, -1)] = True
    corr = (row_hit & col_hit) if mutual else (row_hit | col_hit)
    if threshold is not None:
        corr = corr & ((score_mat > threshold) if largest else (score_mat < threshold))
    if not reduce_result:
        return corr
    r, c = torch.nonzero(corr, as_tuple=True)
    return r, c, score_mat[r, c]


def batch_mutual_topk_select(score_mat, k, row_masks=None, col_masks=None, largest=True, threshold=None, mutual=True,
                             reduce_result=True):
    """2d3d vision3d/ops/mutual_topk_select.py:63-133 (the fine matching of model.py:738-746), ``.cuda()`` replaced by the
    score tensor's device.  The masks are applied AFTER the top-k selection, as in the reference (:122-125)."""
    B, n_rows, n_cols = score_mat.shape
    dev = score_mat.device
    bi = torch.arange(B, device=dev)
    row_pick = score_mat.topk(k=k, largest=largest, dim=2)[1]            # [B,N,k]
    row_hit = torch.zeros_like(score_mat, dtype=torch.bool)
    row_hit[bi.view(B, 1, 1).expand(-1, n_rows, k), torch.arange(n_rows, device=dev).view(1, n_rows, 1).expand(B, -1, k), row_pick] = True
    col_pick = score_mat.topk(k=k, largest=largest, dim=1)[1]            # [B,k,M]
    col_hit = torch.zeros_like(score_mat, dtype=torch.bool)
    col_hit[bi.view(B, 1, 1).expand(-1, k, n_cols), col_pick, torch.arange(n_cols, device=dev).view(1, 1, n_cols).expand(B, k, -1)] = True
    corr = (row_hit & col_hit) if mutual else (row_hit | col_hit)
    if threshold is not None:
        corr = corr & ((score_mat > threshold) if largest else (score_mat < threshold))
    if row_masks is not None:
        corr = corr & row_masks.unsqueeze(2)
    if col_masks is not None:
        corr = corr & col_masks.unsqueeze(1)
    if not reduce_result:
        return corr
    b, r, c = torch.nonzero(corr, as_tuple=True)
    return b, r, c, score_mat[b, r, c]


# --------------------------------------------------------------------------------------
# matching head (a1, a1', a1'', a2, a3)
# --------------------------------------------------------------------------------------
@dataclass
class MatchingParams:
    """The learnable state of the reference ``Matching`` module (matching.py:43-68)."""
    src_proj_weight: torch.Tensor            # [C,C]; applied to BOTH sides (:127-128)
    bin_score: Optional[torch.Tensor] = None  # 0-dim, sinkhorn only
    match_type: str = "sinkhorn"
    skh_iters: int = 3
    temperature: float = 0.1
    confidence_threshold: float = 0.2
    entangled: bool = True


def similarity(p: MatchingParams, src_feats, tgt_feats, src_pe=None, tgt_pe=None):
    """Projection, optional rotary PE, 1/sqrt(C) scaling and the contraction.

    4d/models/matching.py:127-128,135-137,144-145,161.  Returns (sim, fs, ft, fs_nopos,
    ft_nopos) where fs/ft are what the reference stores in ``data`` (:131-141).
    """
    W = p.src_proj_weight
    fs_nopos = src_feats @ W.t()
    ft_nopos = tgt_feats @ W.t()
    fs, ft = fs_nopos, ft_nopos
    if not p.entangled:
        fs = embed_rotary(fs, src_pe[..., 0], src_pe[..., 1])
        ft = embed_rotary(ft, tgt_pe[..., 0], tgt_pe[..., 1])
    C = fs.shape[-1]
    a = fs / C ** 0.5
    b = ft / C ** 0.5
    sim = torch.einsum("bsc,btc->bst", a, b)
    return sim, fs, ft, fs_nopos, ft_nopos


def confidence_from_similarity(p: MatchingParams, sim, src_mask, tgt_mask):
    """Branches :147-157 (dual softmax) and :159-170 (sinkhorn) of 4d/models/matching.py."""
    if p.match_type == "dual_softmax":
        s = sim / p.temperature
        s_rows = s.masked_fill(~src_mask[:, :, None], NEG_INF)   # invalid src rows
        s_cols = s.masked_fill(~tgt_mask[:, None, :], NEG_INF)   # invalid tgt cols
        return torch.softmax(s_rows, dim=1) * torch.softmax(s_cols, dim=2)
    s = sim.masked_fill(~pair_mask(src_mask, tgt_mask), NEG_INF)
    log_assign = log_optimal_transport(s, p.bin_score, p.skh_iters, src_mask, tgt_mask)
    return log_assign.exp()[:, :-1, :-1].contiguous()


def matching_forward_3d(p: MatchingParams, src_feats, tgt_feats, src_pe, tgt_pe, src_mask, tgt_mask):
    """``Matching.forward`` of the 3D flavour, 4d/models/matching.py:118-173.
    Returns (conf [B,N,M], coarse_match [K,3] int64)."""
    sim, *_ = similarity(p, src_feats, tgt_feats, src_pe, tgt_pe)
    conf = confidence_from_similarity(p, sim, src_mask, tgt_mask)
    idx, _, _ = get_match(conf, p.confidence_threshold)
    return conf, idx


def matching_forward_2d3d(p: MatchingParams, src_feats, tgt_feats, src_mask, tgt_mask, mutual=True):
    """``Matching.forward`` of the 2D-3D flavour, 2d3d/matching.py:91-147.
    Returns (conf [1,N,M], src_idx [K], tgt_idx [K], weights [K])."""
    sim, *_ = similarity(p, src_feats, tgt_feats)
    conf = confidence_from_similarity(p, sim, src_mask, tgt_mask)
    r, c, w = mutual_topk_select(conf.squeeze(0), 1, largest=True, threshold=None, mutual=mutual)
    return conf, r, c, w


def matching_forward1_3d(p: MatchingParams, src_feats, tgt_feats, src_pe, tgt_pe, src_mask, tgt_mask, mutual=False):
    """``Matching.forward1``, 3d/models/matching.py:221-283: top-1 row/col union as [K,3]."""
    sim, *_ = similarity(p, src_feats, tgt_feats, src_pe, tgt_pe)
    conf = confidence_from_similarity(p, sim, src_mask, tgt_mask)
    r, c, _ = mutual_topk_select(conf.squeeze(0), 1, largest=True, threshold=None, mutual=mutual)
    return conf, torch.stack((torch.zeros_like(r), r, c), dim=-1)


# --------------------------------------------------------------------------------------
# SoftProcrustes (a7, a8)
# --------------------------------------------------------------------------------------
def batch_weighted_procrustes(X, Y, w, eps=1e-4):
    """Weighted Kabsch, 4d/models/procrustes.py:18-44.

    X, Y: [B,K,3]; w: [B,K,1].  Returns (R [B,3,3] fp32, t [B,3,1], condition [B] fp64).
    Note w is normalised by (sum|w| + eps), so the normalised weights sum to slightly
    less than one and the centroids are slightly shrunk (:27-30)."""
    B = X.shape[0]
    total = w.abs().sum(dim=1, keepdim=True)
    wn = w / (total + eps)
    cx = (wn * X).sum(dim=1, keepdim=True)
    cy = (wn * Y).sum(dim=1, keepdim=True)
    S = (Y - cy).transpose(1, 2) @ (wn * (X - cx))                      # :34
    U, D, V = torch.linalg.svd(S.double().cpu(), full_matrices=False)  # :35-36 (Vh here)
    V = V.transpose(1, 2)
    cond = D.max(dim=1)[0] / D.min(dim=1)[0]
    fix = torch.eye(3, dtype=torch.float64).repeat(B, 1, 1)
    fix[:, 2, 2] = torch.linalg.det(U) * torch.linalg.det(V)           # reflection fix :38-40
    R = (U @ fix @ V.transpose(1, 2)).float().to(X.device)
    t = cy.transpose(1, 2) - R @ cx.transpose(1, 2)
    return R, t, cond                                                   # cond stays on the CPU, as in the reference


def weighted_procrustes_backward(X, Y, w, R, grad_R, grad_t, eps=1e-4):
    """dL/d w [B,K,1] of ``batch_weighted_procrustes`` for given dL/d R, dL/d t -- what torch's autograd returns through
    4d/models/procrustes.py:18-44 (incl. its host SVD), written without an SVD: H = R^T M is symmetric and a perturbation dM turns
    R by R [omega]x with (tr(H) I - H) omega = vee(R^T dM - dM^T R).  Checked against the reference's autograd (golden ``procrb_*``)."""
    def vee(A):
        return torch.stack([A[..., 2, 1], A[..., 0, 2], A[..., 1, 0]], -1)

    def hat(e):
        z = torch.zeros_like(e[..., 0])
        return torch.stack([torch.stack([z, -e[..., 2], e[..., 1]], -1), torch.stack([e[..., 2], z, -e[..., 0]], -1),
                            torch.stack([-e[..., 1], e[..., 0], z], -1)], -2)
    W1 = w.abs().sum(1, keepdim=True)
    wt = w / (W1 + eps)
    sw = wt.sum(1, keepdim=True)
    muX, muY = (wt * X).sum(1, keepdim=True), (wt * Y).sum(1, keepdim=True)
    Xc, Yc = X - muX, Y - muY
    M = Yc.transpose(1, 2) @ (wt * Xc)
    gRp = grad_R - grad_t @ muX
    gmuY = grad_t.transpose(1, 2)
    gmuX = -(R.transpose(1, 2) @ grad_t).transpose(1, 2)
    H = R.transpose(1, 2) @ M
    Kmat = torch.diag_embed(H.diagonal(dim1=1, dim2=2).sum(-1, keepdim=True).expand(-1, 3)) - H
    C = R.transpose(1, 2) @ gRp
    e = torch.linalg.solve(Kmat, vee(0.5 * (C - C.transpose(1, 2))).unsqueeze(-1)).squeeze(-1)
    gM = 2.0 * R @ hat(e)
    gmuX = gmuX - (gM.transpose(1, 2) @ ((1 - sw) * muY).transpose(1, 2)).transpose(1, 2)
    gmuY = gmuY - (gM @ ((1 - sw) * muX).transpose(1, 2)).transpose(1, 2)
    gwt = ((Yc @ gM) * Xc).sum(-1, keepdim=True) + (X * gmuX).sum(-1, keepdim=True) + (Y * gmuY).sum(-1, keepdim=True)
    return gwt / (W1 + eps) - torch.sign(w) * (gwt * w).sum(1, keepdim=True) / (W1 + eps) ** 2


def soft_procrustes(conf, src_pcd, tgt_pcd, src_mask, tgt_mask, sample_rate=1.0,
                    max_condition_num=40.0, padded_lengths=False):
    """``SoftProcrustesLayer.forward``, 4d/models/procrustes.py:48-93.

    ``padded_lengths=True`` selects the 3d variant (3d/models/procrustes.py:61-62) that
    uses the padded sizes instead of the mask sums.  Returns the reference's 6-tuple
    (R, t, R_forwd, t_forwd, condition, solution_mask)."""
    B, N, M = conf.shape
    if padded_lengths:
        src_len = torch.tensor([float(N)], device=conf.device)
        tgt_len = torch.tensor([float(M)], device=conf.device)
    else:
        src_len = src_mask.sum(dim=1)
        tgt_len = tgt_mask.sum(dim=1)
    cap = (torch.maximum(src_len, tgt_len) * sample_rate).int()          # :63-64
    K = int(cap.float().mean().int())                                    # :65
    vals, flat = conf.reshape(B, -1).sort(descending=True, dim=1)        # :66 full sort
    w = vals[:, :K].clone()
    flat = flat[:, :K]
    i_src = flat // M
    i_tgt = flat % M
    bidx = torch.arange(B, device=conf.device)[:, None].expand(B, K)
    P = src_pcd[bidx, i_src]
    Q = tgt_pcd[bidx, i_tgt]
    w[torch.arange(K, device=conf.device)[None, :].expand(B, K) >= cap[:, None]] = 0.0       # :74-76
    R, t, cond = batch_weighted_procrustes(P, Q, w[..., None])
    ok = cond < max_condition_num                                        # :87 (CPU tensors, like cond)
    R_f, t_f = R.clone(), t.clone()
    okd = ok.to(R.device)
    R_f[~okd] = torch.eye(3, dtype=R.dtype, device=R.device)
    t_f[~okd] = torch.zeros(3, 1, dtype=R.dtype, device=R.device)
    return R, t, R_f, t_f, cond, ok


def warp_points(R, t, pts):
    """(R p + t) for row-vector point sets, 4d/models/pipeline.py:220."""
    return (R.float() @ pts.transpose(1, 2) + t.float()).transpose(1, 2)


# --------------------------------------------------------------------------------------
# diffusion schedule and DDIM update (a10)
# --------------------------------------------------------------------------------------
def alphas_cumprod(timesteps=1000, s=0.008):
    """Cosine schedule, 4d/models/pipeline.py:24-34,97-99 (fp64 throughout)."""
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
    f = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    f = f / f[0]
    betas = torch.clip(1 - (f[1:] / f[:-1]), 0, 0.999)
    return torch.cumprod(1.0 - betas, dim=0)


def time_pairs(sampling_steps, timesteps=1000):
    """4d/models/pipeline.py:166-169: [(999,949),...,(49,0)] for 20 steps."""
    ts = torch.linspace(0, timesteps - 1, steps=sampling_steps + 1)
    ts = list(reversed(ts.int().tolist()))
    return list(zip(ts[:-1], ts[1:]))


def ddim_coefficients(ac, t, t_next, eta=1.0):
    """Scalars of one reverse step as fp64 Python floats.

    Returns (sqrt_recip, sqrt_recipm1, sqrt_alpha_next, c, sigma), see
    4d/models/pipeline.py:101-104,182-186."""
    a, an = ac[t], ac[t_next]
    sigma = eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
    c = (1 - an - sigma ** 2).sqrt()
    return (float(torch.sqrt(1.0 / a)), float(torch.sqrt(1.0 / a - 1)), float(an.sqrt()),
            float(c), float(sigma))


def ddim_update(x_t, x0, ac, t, t_next, noise=None, eta=1.0):
    """pred_noise (:201-205) and the x update (:190 with noise; 3d :256 / 2d3d :678 without).

    Reproduces the fp64 promotion (Q4): the schedule buffers are fp64 so the result is
    fp64 whatever the input dtype."""
    ac = ac.to(x_t.device)
    r = torch.sqrt(1.0 / ac[t]).reshape(1, 1, 1)
    rm1 = torch.sqrt(1.0 / ac[t] - 1).reshape(1, 1, 1)
    pred = (r * x_t - x0) / rm1
    a, an = ac[t], ac[t_next]
    sigma = eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
    c = (1 - an - sigma ** 2).sqrt()
    out = x0 * an.sqrt() + c * pred
    if noise is not None:
        out = out + sigma * noise
    return out


def noisy_matching_to_pose(x, alpha, iters, s_pcd, t_pcd, src_mask, tgt_mask,
                           sample_rate=1.0, max_condition_num=40.0, padded_lengths=False):
    """``get_warped_from_noising_matching`` 4d/models/pipeline.py:207-223 (3d :293-309,
    2d3d model.py:830-846).  Mutates ``x`` in place (Q7).  Returns (src_warped, conf, pose6)."""
    x.masked_fill_(~pair_mask(src_mask, tgt_mask), NEG_INF)
    conf = log_optimal_transport(x, alpha, iters, src_mask, tgt_mask).exp()[:, :-1, :-1].contiguous().float()
    pose = soft_procrustes(conf, s_pcd, t_pcd, src_mask, tgt_mask, sample_rate, max_condition_num, padded_lengths)
    return warp_points(pose[2], pose[3], s_pcd), conf, pose


def sampler(flavour, p: MatchingParams, src_feats, tgt_feats, s_pcd, t_pcd, src_mask, tgt_mask,
            x_T, steps, noises=None, sample_rate=1.0, max_condition_num=40.0,
            tgt_mask_pose=None, t_pcd_pose=None, state_dtype=None, trace=None, pe_fn=None):
    """Reverse-diffusion loop with the denoising transformer replaced by FIXED features.

    flavour '4d'  : 4d/models/pipeline.py:171-192   (sigma*noise term, final sigmoid)
    flavour '3d'  : 3d/models/pipeline.py:235-278   (x -= x.min() first, no noise, final
                                                    Sinkhorn + top-1 row/col union)
    flavour '2d3d': 2d3d/model.py:651-694           (no noise, final Sinkhorn + top-1 union;
                                                    pose step matches against t_pcd_pose with
                                                    tgt_mask_pose)
    ``state_dtype=torch.float32`` casts the state back to fp32 after every update (what the
    CUDA path does); ``None`` keeps the reference's fp64 creep (Q4).
    ``pe_fn(src_warped, t_pcd) -> (src_pe, tgt_pe)``: the position codes the denoising transformer
    hands to the matching head every step (pipeline.py:177-178; used when ``p.entangled`` is False).
    Returns a dict with the final matrix, matches and the last pose.
    """
    ac = alphas_cumprod()
    x = x_T.clone()
    out = {}
    pose_mask = tgt_mask if tgt_mask_pose is None else tgt_mask_pose
    pose_pcd = t_pcd if t_pcd_pose is None else t_pcd_pose
    for k, (t, t_next) in enumerate(time_pairs(steps)):
        if flavour == "3d":
            x = x - x.min()
        warped, conf_d, pose = noisy_matching_to_pose(
            x, p.bin_score, p.skh_iters, s_pcd, pose_pcd, src_mask, pose_mask,
            sample_rate, max_condition_num, padded_lengths=(flavour == "3d"))
        src_pe, tgt_pe = pe_fn(warped, t_pcd) if pe_fn is not None else (None, None)
        sim, *_ = similarity(p, src_feats, tgt_feats, src_pe, tgt_pe)
        x0 = confidence_from_similarity(p, sim, src_mask, tgt_mask)
        if flavour != "2d3d":
            get_match(x0, p.confidence_threshold)       # computed and discarded (:172 / :178)
        else:
            mutual_topk_select(x0.squeeze(0), 1, True, None, True)
        noise = noises[k] if (flavour == "4d" and noises is not None) else None
        x_prev = x
        x = ddim_update(x, x0, ac, t, t_next, noise)
        if state_dtype is not None:
            x = x.to(state_dtype)
        if trace is not None:
            trace.append({"x_in": x_prev, "conf_d": conf_d, "pose": pose, "warped": warped,
                          "x0": x0, "x_out": x})
        out["pose"] = pose
    if flavour == "4d":
        out["conf_matrix_pred"] = torch.sigmoid(x)
    else:
        s = x - x.min() if flavour == "3d" else x
        s = s.masked_fill(~pair_mask(src_mask, tgt_mask), NEG_INF)
        conf = log_optimal_transport(s, p.bin_score, p.skh_iters, src_mask, tgt_mask).exp()[:, :-1, :-1].contiguous()
        r, c, w = mutual_topk_select(conf.squeeze(0), 1, True, None, False)
        out["conf_matrix_pred"] = conf
        out["match_pred"] = torch.stack((torch.zeros_like(r), r, c), dim=-1)
        out["match_weights"] = w
    out["x_final"] = x
    return out


# --------------------------------------------------------------------------------------
# synthetic inputs shared by tests, smoke() and bench.py
# --------------------------------------------------------------------------------------
# ---------------------------------------------------------------------------------------------------------------
# Denoising transformer: one geometry attention layer and the self / cross stack
# ---------------------------------------------------------------------------------------------------------------
def geometry_attention_layer(w, x, source, x_pe, source_pe, x_mask, source_mask, pe_type, nhead, eps=1e-5):
    """``GeometryAttentionLayer.forward`` 4d/models/transformer.py:43-96.  `w`: dict with the module's state_dict keys
    (q_proj.weight, k_proj.weight, v_proj.weight, merge.weight, mlp.0.weight, mlp.2.weight, norm1.weight/bias,
    norm2.weight/bias).  Computed in the dtype of x (fp32 like the reference, or fp64 for an error yardstick)."""
    dt = x.dtype
    g = lambda k: w[k].to(dt)
    bs = x.shape[0]
    q, k, v = x, source, source
    if pe_type == "sinusoidal":
        if x_pe is not None:                      # :52-55
            q = q + x_pe
            k = k + source_pe
        qw, kw, vw = q @ g("q_proj.weight").t(), k @ g("k_proj.weight").t(), v @ g("v_proj.weight").t()
    elif pe_type == "rotary":
        qw, kw, vw = q @ g("q_proj.weight").t(), k @ g("k_proj.weight").t(), v @ g("v_proj.weight").t()
        if x_pe is not None:                      # :66-70
            qw = embed_rotary(qw, x_pe[..., 0].to(dt), x_pe[..., 1].to(dt))
            kw = embed_rotary(kw, source_pe[..., 0].to(dt), source_pe[..., 1].to(dt))
    else:
        raise KeyError()
    dim = qw.shape[-1] // nhead
    qw = qw.view(bs, -1, nhead, dim)
    kw = kw.view(bs, -1, nhead, dim)
    vw = vw.view(bs, -1, nhead, dim)
    a = torch.einsum("nlhd,nshd->nlsh", qw, kw)   # :79
    if source_mask is not None:                   # :80-81: keys are masked for VALID queries only
        a = a.masked_fill(x_mask[:, :, None, None] * (~source_mask[:, None, :, None]), float("-inf"))
    a = a / qw.size(3) ** 0.5
    a = torch.softmax(a, dim=2)
    o = torch.einsum("nlsh,nshd->nlhd", a, vw).contiguous()
    message = o.view(bs, -1, nhead * dim) @ g("merge.weight").t()
    message = torch.nn.functional.layer_norm(message, (message.shape[-1],), g("norm1.weight"), g("norm1.bias"), eps)
    h = torch.cat([x, message], dim=2) @ g("mlp.0.weight").t()
    message = torch.relu(h) @ g("mlp.2.weight").t()
    message = torch.nn.functional.layer_norm(message, (message.shape[-1],), g("norm2.weight"), g("norm2.bias"), eps)
    return x + message


def attention_stack(layer_weights, layer_types, src_feat, tgt_feat, src_pe, tgt_pe, src_mask, tgt_mask, pe_type, nhead,
                    entangled=False):
    """The self / cross layers of ``RepositioningTransformer.forward`` 4d/models/transformer.py:169-233 (no positioning
    layer: the denoising transformer of pipeline.py:84-85 has none).  `layer_weights[i]` is the state dict of layer i."""
    if entangled:                                 # :217-219
        src_feat = embed_pos(pe_type, src_feat, src_pe)
        tgt_feat = embed_pos(pe_type, tgt_feat, tgt_pe)
        sp = tp = None
    else:
        sp, tp = src_pe, tgt_pe
    for w, name in zip(layer_weights, layer_types):
        if name == "self":
            src_feat = geometry_attention_layer(w, src_feat, src_feat, sp, sp, src_mask, src_mask, pe_type, nhead)
            tgt_feat = geometry_attention_layer(w, tgt_feat, tgt_feat, tp, tp, tgt_mask, tgt_mask, pe_type, nhead)
        elif name == "cross":
            src_feat = geometry_attention_layer(w, src_feat, tgt_feat, sp, tp, src_mask, tgt_mask, pe_type, nhead)
            tgt_feat = geometry_attention_layer(w, tgt_feat, src_feat, tp, sp, tgt_mask, src_mask, pe_type, nhead)
        else:
            raise KeyError(name)
    return src_feat, tgt_feat


# ---------------------------------------------------------------------------------------------------------------
# 2D-3D flavour of the same row: CrossModalFusionModule (vision3d TransformerLayer blocks + Fourier embedding)
# ---------------------------------------------------------------------------------------------------------------
def fourier_embedding(x, length, k0=0.0, use_pi=True, use_input=False):
    """``FourierEmbedding.forward`` 2d3d/vision3d/layers/embedding.py:75-99: [x | sin(f_0 x), cos(f_0 x), sin(f_1 x) ...] with
    f_l = 2^(k0 + l) [* pi]; each (sin | cos) block is as wide as x."""
    n = x.shape[-1]
    f = (2.0 ** torch.arange(k0, k0 + length).float()).to(x.dtype).view(1, -1, 1)
    if use_pi:
        f = f * math.pi
    th = f * x.reshape(-1, 1, n)
    emb = torch.cat([torch.sin(th), torch.cos(th)], dim=-1).view(*x.shape[:-1], 2 * length * n)
    return torch.cat([x, emb], dim=-1) if use_input else emb


def vision3d_transformer_layer(w, q_tokens, k_tokens, v_tokens, k_masks, nhead, eps=1e-5):
    """``TransformerLayer.forward`` 2d3d/vision3d/layers/transformer.py:8-301 as the fusion module calls it (tokens and key masks
    only; k_masks True = key ignored, :76,:131).  `w`: the layer's state_dict (attention.attention.{q,k,v}_token_layer.*,
    attention.linear.*, attention.norm.*, output.expand.*, output.squeeze.*, output.norm.*).  Returns (tokens, scores)."""
    dt = q_tokens.dtype
    lin = lambda name, x: x @ w[name + ".weight"].to(dt).t() + w[name + ".bias"].to(dt)
    B, N, C = q_tokens.shape
    M = k_tokens.shape[1]
    d = C // nhead
    q = lin("attention.attention.q_token_layer", q_tokens).view(B, N, nhead, d).permute(0, 2, 1, 3)
    k = lin("attention.attention.k_token_layer", k_tokens).view(B, M, nhead, d).permute(0, 2, 1, 3)
    v = lin("attention.attention.v_token_layer", v_tokens).view(B, M, nhead, d).permute(0, 2, 1, 3)
    a = torch.einsum("bhnc,bhmc->bhnm", q, k) / d ** 0.5                                   # :127-128
    if k_masks is not None:
        a = a.masked_fill(k_masks[:, None, None, :], float("-inf"))                        # :133-134
    a = torch.softmax(a, dim=-1)
    hidden = torch.matmul(a, v).permute(0, 2, 1, 3).reshape(B, N, C)                       # :154-156
    hidden = lin("attention.linear", hidden)
    ln = lambda name, x: torch.nn.functional.layer_norm(x, (C,), w[name + ".weight"].to(dt), w[name + ".bias"].to(dt), eps)
    tokens = ln("attention.norm", hidden + q_tokens)                                       # :214
    hidden = lin("output.squeeze", torch.relu(lin("output.expand", tokens)))               # :232-234
    return ln("output.norm", tokens + hidden), a                                           # :236


def cross_modal_fusion(w, blocks, nhead, img_feats, img_feats_dino, img_pixels, pcd_feats, pcd_points, img_masks=None,
                       pcd_masks=None, use_embedding=True, embedding_dim=10):
    """``CrossModalFusionModule.forward`` 2d3d/experiments/<exp>/fusion_module.py:61-107.  `w`: the module's state_dict."""
    dt = img_feats.dtype
    lin = lambda name, x: x @ w[name + ".weight"].to(dt).t() + w[name + ".bias"].to(dt)
    img = torch.relu(torch.cat([lin("img_in_proj", img_feats), lin("img_in_proj_dino", img_feats_dino)], dim=-1))   # :83
    img = lin("img_in_proj_all", img)
    pcd = lin("pcd_in_proj", pcd_feats)
    if use_embedding:
        img = img + lin("img_emb_proj", fourier_embedding(img_pixels, embedding_dim, use_pi=False, use_input=True))   # :51-54
        pts = pcd_points - pcd_points.mean(dim=1)                                                                      # :57
        pcd = pcd + lin("pcd_emb_proj", fourier_embedding(pts, embedding_dim, use_pi=False, use_input=True))
    for i, block in enumerate(blocks):
        lw = {k[len(f"transformer.{i}."):]: v for k, v in w.items() if k.startswith(f"transformer.{i}.")}
        if block == "self":                                                                                            # :95-101
            img = vision3d_transformer_layer(lw, img, img, img, img_masks, nhead)[0]
            pcd = vision3d_transformer_layer(lw, pcd, pcd, pcd, pcd_masks, nhead)[0]
        elif block == "cross":
            img = vision3d_transformer_layer(lw, img, pcd, pcd, pcd_masks, nhead)[0]
            pcd = vision3d_transformer_layer(lw, pcd, img, img, img_masks, nhead)[0]
        else:
            raise KeyError(block)
    return lin("out_proj", img), lin("out_proj", pcd)


def random_rotation(gen):
    q = torch.randn(4, generator=gen)
    q = q / q.norm()
    w, x, y, z = q.tolist()
    return torch.tensor([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                         [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                         [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def make_problem(seed, B, N, M, C=256, prefix_valid=None, arbitrary_invalid=0.0):
    """Seeded synthetic pair(s): unit-normal features, nn.Linear-style projection weight,
    points related by a random rigid motion plus 1 % noise, masks (all true / prefix /
    arbitrary)."""
    g = torch.Generator().manual_seed(seed)
    src_feats = torch.randn(B, N, C, generator=g)
    tgt_feats = torch.randn(B, M, C, generator=g)
    bound = 1.0 / math.sqrt(C)
    W = (torch.rand(C, C, generator=g) * 2 - 1) * bound
    s_pcd = torch.randn(B, N, 3, generator=g)
    t_pcd = torch.empty(B, M, 3)
    for b in range(B):
        R = random_rotation(g)
        t = torch.randn(3, generator=g)
        perm = torch.randint(0, N, (M,), generator=g)
        t_pcd[b] = s_pcd[b, perm] @ R.t() + t + 0.01 * torch.randn(M, 3, generator=g)
    src_mask = torch.ones(B, N, dtype=torch.bool)
    tgt_mask = torch.ones(B, M, dtype=torch.bool)
    if prefix_valid is not None:
        for b, (ns, nt) in enumerate(prefix_valid):
            src_mask[b, ns:] = False
            tgt_mask[b, nt:] = False
            src_feats[b, ns:] = 0
            tgt_feats[b, nt:] = 0
            s_pcd[b, ns:] = 0
            t_pcd[b, nt:] = 0
    if arbitrary_invalid > 0:
        src_mask &= torch.rand(B, N, generator=g) >= arbitrary_invalid
        tgt_mask &= torch.rand(B, M, generator=g) >= arbitrary_invalid
    return dict(src_feats=src_feats, tgt_feats=tgt_feats, W=W, s_pcd=s_pcd, t_pcd=t_pcd,
                src_mask=src_mask, tgt_mask=tgt_mask)


# --------------------------------------------------------------------------------------
# Correspondence RANSAC.  PARITY UNPINNED: the algorithm lives in open3d (pinned 0.13.0 in
# Diff-Reg-4dmatch/eccv24_4d_env.yml, absent from this image and not importable here), called at
# Diff-Reg-4dmatch/models/loss.py:13-24 (ransac_pose_estimation) from loss.py:366-398 (ransac_regist_coarse).  This is
# a restatement of open3d's published registration_ransac_based_on_correspondence (trial = draw ransac_n correspondences
# with replacement, Umeyama-without-scale fit, inliers = correspondences closer than the threshold under the fit, best =
# more inliers, then lower rmse) with the CUDA path's counter-based draws, so that both evaluate the SAME trials; open3d
# itself seeds from random_device and is not reproducible run to run.  No golden vectors exist for it.
# --------------------------------------------------------------------------------------
def _splitmix64(x):
    import numpy as np
    with np.errstate(over="ignore"):
        x = x + np.uint64(0x9E3779B97F4A7C15)
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))


def ransac_draws(seed, b, trials, ransac_n, num_corr):
    """[trials, ransac_n] correspondence numbers of batch element b (the generator of csrc/ransac.cu: draw_index)."""
    import numpy as np
    h = np.arange(trials, dtype=np.uint64)[:, None]
    j = np.arange(ransac_n, dtype=np.uint64)[None, :]
    ctr = (np.uint64(b) << np.uint64(40)) ^ (h << np.uint64(4)) ^ j
    r = _splitmix64(np.uint64(seed) ^ _splitmix64(ctr))
    return (((r >> np.uint64(32)) * np.uint64(num_corr)) >> np.uint64(32)).astype(np.int64)


def rigid_fit(xs, ys):
    """Kabsch / Umeyama without scale on [T,n,3] fp32 point pairs in fp64 (numpy SVD, det fix on the last singular pair)
    -> R [T,3,3] fp32, t [T,3] fp32, valid [T] (second singular value above 1e-7 of the first)."""
    import numpy as np
    xs = xs.astype(np.float64)
    ys = ys.astype(np.float64)
    mx, my = xs.mean(1), ys.mean(1)
    H = np.einsum("tka,tkc->tac", ys - my[:, None], xs - mx[:, None])
    U, s, Vt = np.linalg.svd(H)
    valid = (s[:, 0] > 0) & (s[:, 1] > 1e-7 * s[:, 0])
    d = np.sign(np.linalg.det(U) * np.linalg.det(Vt))
    D = np.tile(np.eye(3), (len(H), 1, 1))
    D[:, 2, 2] = d
    R = U @ D @ Vt
    t = my - np.einsum("tac,tc->ta", R, mx)
    return R.astype(np.float32), t.astype(np.float32), valid


def ransac_correspondence(src_pcd, tgt_pcd, match_pred, distance_threshold=0.05, ransac_n=3, max_iteration=50000, seed=0,
                          chunk=2048):
    """numpy restatement; src_pcd [B,N,3], tgt_pcd [B,M,3], match_pred [C,3] (b, i, j) grouped by b.
    -> dict of numpy arrays: pose [B,4,4], fitness, inlier_rmse, best_trial, inlier_count, trial_count [B,T], trial_err2 [B,T]."""
    import numpy as np
    src = np.asarray(src_pcd, dtype=np.float32)
    tgt = np.asarray(tgt_pcd, dtype=np.float32)
    match = np.asarray(match_pred, dtype=np.int64).reshape(-1, 3)
    B, T = src.shape[0], int(max_iteration)
    thr2 = np.float32(distance_threshold) * np.float32(distance_threshold)
    out = {"pose": np.tile(np.eye(4, dtype=np.float32), (B, 1, 1)), "fitness": np.zeros(B, np.float32),
           "inlier_rmse": np.zeros(B, np.float32), "best_trial": np.full(B, -1, np.int32), "inlier_count": np.zeros(B, np.int32),
           "trial_count": np.zeros((B, T), np.int32), "trial_err2": np.zeros((B, T), np.float32)}
    for b in range(B):
        rows = match[match[:, 0] == b]
        C = len(rows)
        if C < 3:  # loss.py:384-387
            continue
        S, G = src[b][rows[:, 1]], tgt[b][rows[:, 2]]
        pick = ransac_draws(seed, b, T, ransac_n, C)
        R, t, valid = rigid_fit(S[pick], G[pick])
        cnt = np.zeros(T, np.int32)
        err = np.zeros(T, np.float32)
        for a in range(0, T, chunk):
            P = np.einsum("tac,kc->tka", R[a:a + chunk], S) + t[a:a + chunk, None, :]
            d2 = ((P - G[None]) ** 2).sum(-1).astype(np.float32)
            inl = d2 < thr2
            cnt[a:a + chunk] = inl.sum(1)
            err[a:a + chunk] = np.where(inl, d2, np.float32(0)).sum(1, dtype=np.float32)
        cnt = np.where(valid, cnt, -1).astype(np.int32)
        err = np.where(valid, err, np.float32(0)).astype(np.float32)
        out["trial_count"][b], out["trial_err2"][b] = cnt, err
        best = ransac_best_trial(cnt, err)
        if best < 0:
            continue
        out["best_trial"][b], out["inlier_count"][b] = best, cnt[best]
        out["fitness"][b] = np.float32(cnt[best]) / np.float32(C)
        out["inlier_rmse"][b] = np.sqrt(err[best] / np.float32(cnt[best]))
        out["pose"][b, :3, :3], out["pose"][b, :3, 3] = R[best], t[best]
    return out


def ransac_best_trial(trial_count, trial_err2):
    """open3d's sequential 'is better' scan as one arg-max: most inliers, then the smaller error sum, then the first trial."""
    import numpy as np
    cnt = np.asarray(trial_count)
    if not (cnt > 0).any():
        return -1
    top = cnt.max()
    cand = np.flatnonzero(cnt == top)
    e = np.asarray(trial_err2)[cand]
    return int(cand[np.flatnonzero(e == e.min())[0]])
