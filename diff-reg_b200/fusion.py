"""Drop-in for the 2D-3D flavour's denoising / coarse transformer:
``CrossModalFusionModule`` of Diff-Reg-2d3d/experiments/<exp>/fusion_module.py:10-107 and the pieces of ``vision3d.layers`` it is built
from -- ``TransformerLayer`` (vision3d/layers/transformer.py:241-301: multi-head attention with biased linears, post-norm residual
blocks) and ``FourierEmbedding`` (vision3d/layers/embedding.py:52-99).  Same constructor arguments, argument order, return values
and ``state_dict`` keys (``transformer.<i>.attention.attention.q_token_layer.weight`` ..., ``img_in_proj.weight`` ...), so a
reference checkpoint loads strictly.  Shadow the reference module with

    # experiments/<exp>/fusion_module.py
    from diffreg_b200.fusion import CrossModalFusionModule      # noqa: F401

Every linear layer and both attention products run on the tcgen05 split GEMM (the bias rides in its epilogue,
``drg_gemm_nt_split16_bias``), the head split happens in the operand staging, the masked softmax writes the P.V operand, the
post-norm ``LayerNorm(x + f(x))`` is one kernel (``drg_layernorm`` with pre_add), the Fourier embedding is ``drg_fourier_embed``.
What the fusion module never uses of ``TransformerLayer`` -- the five optional embedding inputs, key / pair weights, pair masks,
dropout > 0, activations other than ReLU -- raises NotImplementedError instead of being computed differently.  CUDA tensors
only; forward-only (training keeps the reference module)."""
from typing import List, Optional

import torch
from torch import nn

from . import ops
from .graphs import ForwardGraphCache
from .matching import _no_grad_inputs


def _act_name(act_cfg):
    name = act_cfg if isinstance(act_cfg, str) else act_cfg.get("type", None)
    if name != "ReLU":
        raise NotImplementedError(f"diffreg_b200.fusion: activation {act_cfg!r} stays with the reference module (ReLU only)")
    return name


def _no_dropout(dropout):
    if dropout is not None and dropout > 0:
        raise NotImplementedError("diffreg_b200.fusion: dropout > 0 stays with the reference module (inference path)")


class _WeightCache:
    """Right-hand split operands of nn.Linear weights, re-staged when a weight changes."""
    def __init__(self):
        self._c = {}

    def get(self, lin):
        w = lin.weight
        key = (w.data_ptr(), w._version, w.device)
        hit = self._c.get(id(lin))
        if hit is None or hit[0] != key:
            hit = (key, ops.prep_operand(w.detach(), 1.0, True, 1))
            self._c[id(lin)] = hit
        return hit[1]


def _linear(cache, lin, x, relu=False, staged=None):
    """nn.Linear (with bias) of x [..., K] -> [..., C_out] through the split GEMM; relu: applied to x in the staging; staged: the
    left split operand of x when the caller has it already (q / k / v of a self-attention block share one staging)."""
    K = x.shape[-1]
    rows = x.numel() // K
    a16 = staged if staged is not None else (ops.prep_relu(x, 0) if relu else ops.prep_operand(x, 1.0, True, 0))
    out = ops.gemm_nt(a16.reshape(rows, a16.shape[-1]), cache.get(lin), split3=True, K=K, bias=lin.bias)
    return out.view(*x.shape[:-1], lin.weight.shape[0])


class FourierEmbedding(nn.Module):
    """vision3d/layers/embedding.py:52-99."""
    def __init__(self, length: int, k0: float = 0.0, use_pi: bool = True, use_input: bool = False) -> None:
        super().__init__()
        self.length = length
        self.k0 = k0
        self.use_pi = use_pi
        self.use_input = use_input

    def forward(self, inputs, center=None):
        return ops.fourier_embed(inputs, self.length, self.k0, self.use_pi, self.use_input, center=center)


class MultiHeadAttention(nn.Module):
    """vision3d/layers/transformer.py:8-144 (token inputs and key masks)."""
    def __init__(self, d_model, num_heads, q_embed_proj=False, k_embed_proj=False, v_embed_proj=False, qk_embed_proj=False,
                 qv_embed_proj=False, dropout=None):
        super().__init__()
        assert d_model % num_heads == 0, f"'d_model={d_model}' is not divisible by 'num_heads={num_heads}'."
        if q_embed_proj or k_embed_proj or v_embed_proj or qk_embed_proj or qv_embed_proj:
            raise NotImplementedError("diffreg_b200.fusion: embedding projections stay with the reference module")
        _no_dropout(dropout)
        if (d_model // num_heads) % 4:
            raise ValueError("MultiHeadAttention: d_model / num_heads must be a multiple of 4 (operand rows of the head GEMMs)")
        self.d_model = d_model
        self.num_heads = num_heads
        self.d_model_per_head = d_model // num_heads
        self.q_token_layer = nn.Linear(d_model, d_model)
        self.k_token_layer = nn.Linear(d_model, d_model)
        self.v_token_layer = nn.Linear(d_model, d_model)
        self.dropout = nn.Identity()
        self._cache = _WeightCache()

    def forward(self, q_tokens, k_tokens, v_tokens, q_embeds=None, k_embeds=None, v_embeds=None, qk_embeds=None, qv_embeds=None,
                k_weights=None, k_masks=None, qk_weights=None, qk_masks=None, want_scores=False):
        if any(t is not None for t in (q_embeds, k_embeds, v_embeds, qk_embeds, qv_embeds, k_weights, qk_weights, qk_masks)):
            raise NotImplementedError("diffreg_b200.fusion: embeddings / weights / pair masks stay with the reference module")
        B, N, C = q_tokens.shape
        M = k_tokens.shape[1]
        H, d = self.num_heads, self.d_model_per_head
        # one staging per distinct input: self blocks pass the same tensor three times, cross blocks the same keys and values
        sq = ops.prep_operand(q_tokens, 1.0, True, 0)
        sk = sq if k_tokens is q_tokens else ops.prep_operand(k_tokens, 1.0, True, 0)
        sv = sk if v_tokens is k_tokens else (sq if v_tokens is q_tokens else ops.prep_operand(v_tokens, 1.0, True, 0))
        q = _linear(self._cache, self.q_token_layer, q_tokens, staged=sq)
        k = _linear(self._cache, self.k_token_layer, k_tokens, staged=sk)
        v = _linear(self._cache, self.v_token_layer, v_tokens, staged=sv)
        q16 = ops.prep_heads(q, H, 0)
        k16 = ops.prep_heads(k, H, 1)
        keep = None if k_masks is None else ~k_masks                                   # k_masks: True = ignored (transformer.py:76)
        if not want_scores and d <= ops.FLASH_MAX_HEAD:                                # one fused kernel; the scores never reach HBM
            return ops.attention(q16, k16, v, H, None, keep, 1.0 / d ** 0.5, d), None
        logits = ops.gemm_nt(q16, k16, split3=True, K=d)                               # [B*H, N, M]
        res = ops.attn_softmax(logits, H, None, keep, 1.0 / d ** 0.5, want_operand=True, want_probs=want_scores)
        p16, scores = res if want_scores else (res, None)
        vt = v.view(B, M, H, d).permute(0, 2, 3, 1).contiguous().view(B * H, d, M)
        if M % 4:
            vt = torch.nn.functional.pad(vt, (0, 4 - M % 4))
        o = ops.gemm_nt(p16, ops.prep_operand(vt, 1.0, True, 1), split3=True, K=M)     # [B*H, N, d]
        hidden = o.view(B, H, N, d).permute(0, 2, 1, 3).contiguous().view(B, N, C)
        return hidden, (scores.view(B, H, N, M) if scores is not None else None)


class AttentionLayer(nn.Module):
    """vision3d/layers/transformer.py:147-215."""
    def __init__(self, d_model, num_heads, q_embed_proj=False, k_embed_proj=False, v_embed_proj=False, qk_embed_proj=False,
                 qv_embed_proj=False, dropout=None):
        super().__init__()
        self.attention = MultiHeadAttention(d_model, num_heads, q_embed_proj, k_embed_proj, v_embed_proj, qk_embed_proj, qv_embed_proj,
                                            dropout)
        self.linear = nn.Linear(d_model, d_model)
        self.dropout = nn.Identity()
        self.norm = nn.LayerNorm(d_model)
        self._cache = _WeightCache()

    def forward(self, q_tokens, k_tokens, v_tokens, want_scores=False, **kw):
        hidden, scores = self.attention(q_tokens, k_tokens, v_tokens, want_scores=want_scores, **kw)
        hidden = _linear(self._cache, self.linear, hidden)
        out = ops.layernorm(hidden, self.norm.weight, self.norm.bias, self.norm.eps, residual=q_tokens, pre_add=True, stage=True)
        return out, scores


class AttentionOutput(nn.Module):
    """vision3d/layers/transformer.py:218-237."""
    def __init__(self, d_model, dropout=None, act_cfg="ReLU"):
        super().__init__()
        _no_dropout(dropout)
        _act_name(act_cfg)
        self.expand = nn.Linear(d_model, d_model * 2)
        self.activation = nn.ReLU()
        self.squeeze = nn.Linear(d_model * 2, d_model)
        self.dropout = nn.Identity()
        self.norm = nn.LayerNorm(d_model)
        self._cache = _WeightCache()

    def forward(self, input_tokens):
        hidden = _linear(self._cache, self.expand, input_tokens)
        hidden = _linear(self._cache, self.squeeze, hidden, relu=True)
        return ops.layernorm(hidden, self.norm.weight, self.norm.bias, self.norm.eps, residual=input_tokens, pre_add=True, stage=True)


class TransformerLayer(nn.Module):
    """vision3d/layers/transformer.py:241-301."""
    def __init__(self, d_model, num_heads, q_embed_proj=False, k_embed_proj=False, v_embed_proj=False, qk_embed_proj=False,
                 qv_embed_proj=False, dropout=None, act_cfg="ReLU"):
        super().__init__()
        self.attention = AttentionLayer(d_model, num_heads, q_embed_proj, k_embed_proj, v_embed_proj, qk_embed_proj, qv_embed_proj, dropout)
        self.output = AttentionOutput(d_model, dropout=dropout, act_cfg=act_cfg)

    def forward(self, q_tokens, k_tokens, v_tokens, q_embeds=None, k_embeds=None, v_embeds=None, qk_embeds=None, qv_embeds=None,
                k_weights=None, k_masks=None, qk_weights=None, qk_masks=None, return_attention_score=False):
        _no_grad_inputs(q_tokens, k_tokens, v_tokens, module=self)
        with torch.no_grad():
            hidden, scores = self.attention(q_tokens, k_tokens, v_tokens, want_scores=return_attention_score, q_embeds=q_embeds,
                                            k_embeds=k_embeds, v_embeds=v_embeds, qk_embeds=qk_embeds, qv_embeds=qv_embeds,
                                            k_weights=k_weights, k_masks=k_masks, qk_weights=qk_weights, qk_masks=qk_masks)
            out = self.output(hidden)
        if return_attention_score:
            return out, scores
        return out


class CrossModalFusionModule(nn.Module):
    """experiments/<exp>/fusion_module.py:10-107."""
    def __init__(self, img_input_dim: int, pcd_input_dim: int, output_dim: int, hidden_dim: int, num_heads: int, blocks: List[str],
                 dropout: Optional[float] = None, activation_fn: str = "ReLU", use_embedding: bool = True, embedding_dim: int = 10):
        super().__init__()
        self.use_embedding = use_embedding
        if self.use_embedding:
            self.embedding = FourierEmbedding(embedding_dim, use_pi=False, use_input=True)
            self.img_emb_proj = nn.Linear(embedding_dim * 4 + 2, hidden_dim)
            self.pcd_emb_proj = nn.Linear(embedding_dim * 6 + 3, hidden_dim)
        else:
            self.embedding = None
            self.img_emb_proj = None
            self.pcd_emb_proj = None
        self.img_in_proj = nn.Linear(img_input_dim, hidden_dim)
        self.img_in_proj_dino = nn.Linear(img_input_dim * 2, hidden_dim)
        self.img_in_proj_all = nn.Linear(img_input_dim, hidden_dim)
        self.pcd_in_proj = nn.Linear(pcd_input_dim, hidden_dim)
        self.out_proj = nn.Linear(hidden_dim, output_dim)
        self.blocks = blocks
        layers = []
        for block in self.blocks:
            assert block in ["self", "cross"]
            layers.append(TransformerLayer(hidden_dim, num_heads, dropout=dropout, act_cfg=activation_fn))
        self.transformer = nn.ModuleList(layers)
        self._cache = _WeightCache()
        self.graph_replay = True          # (set False to run every call eagerly)
        self._graphs = ForwardGraphCache()

    def _padded_linear(self, lin, x):
        """A linear whose input width is not a multiple of 4 (the 42 / 63 wide Fourier embeddings): zero columns are appended to
        the input and to the weight operand (operand rows are staged 16 bytes at a time)."""
        K = x.shape[-1]
        pad = (-K) % 4
        if pad == 0:
            return _linear(self._cache, lin, x)
        xp = torch.nn.functional.pad(x, (0, pad))
        key = ("pad", id(lin))
        w = lin.weight
        ver = (w.data_ptr(), w._version, w.device)
        hit = self._cache._c.get(key)
        if hit is None or hit[0] != ver:
            hit = (ver, ops.prep_operand(torch.nn.functional.pad(w.detach(), (0, pad)), 1.0, True, 1))
            self._cache._c[key] = hit
        rows = xp.numel() // (K + pad)
        a16 = ops.prep_operand(xp, 1.0, True, 0)
        out = ops.gemm_nt(a16.reshape(rows, a16.shape[-1]), hit[1], split3=True, K=K + pad, bias=lin.bias)
        return out.view(*x.shape[:-1], w.shape[0])

    def create_2d_embedding(self, pixels):
        return self._padded_linear(self.img_emb_proj, self.embedding(pixels))

    def create_3d_embedding(self, points):
        # points - points.mean(dim=1) (fusion_module.py:57; the broadcast there only works for one cloud per call)
        if points.shape[0] != 1:
            raise ValueError("CrossModalFusionModule.create_3d_embedding: one point cloud per call, as in the reference")
        # the mean in fp64, rounded once: the 2^9 x frequencies amplify every ulp of the centre 512-fold
        center = points.double().mean(dim=1).float().reshape(-1)
        return self._padded_linear(self.pcd_emb_proj, self.embedding(points, center=center))

    def forward(self, img_feats, img_feats_dino, img_pixels, pcd_feats, pcd_points, img_masks=None, pcd_masks=None):
        _no_grad_inputs(img_feats, img_feats_dino, img_pixels, pcd_feats, pcd_points, module=self)
        with torch.no_grad():
            if self.graph_replay and img_feats.is_cuda:      # one CUDA-graph replay per call once a signature was seen twice
                return self._graphs.run(self, self._forward, (img_feats, img_feats_dino, img_pixels, pcd_feats, pcd_points, img_masks,
                                                              pcd_masks))
            return self._forward(img_feats, img_feats_dino, img_pixels, pcd_feats, pcd_points, img_masks, pcd_masks)

    def _forward(self, img_feats, img_feats_dino, img_pixels, pcd_feats, pcd_points, img_masks, pcd_masks):
        a = _linear(self._cache, self.img_in_proj, img_feats)
        b = _linear(self._cache, self.img_in_proj_dino, img_feats_dino)
        img_tokens = _linear(self._cache, self.img_in_proj_all, torch.cat([a, b], dim=-1), relu=True)   # relu(cat(...)) in the staging
        pcd_tokens = _linear(self._cache, self.pcd_in_proj, pcd_feats)
        if self.use_embedding:
            img_tokens = img_tokens + self.create_2d_embedding(img_pixels)
            pcd_tokens = pcd_tokens + self.create_3d_embedding(pcd_points)
        for i, block in enumerate(self.blocks):
            if block == "self":
                img_tokens = self.transformer[i](img_tokens, img_tokens, img_tokens, k_masks=img_masks)
                pcd_tokens = self.transformer[i](pcd_tokens, pcd_tokens, pcd_tokens, k_masks=pcd_masks)
            else:
                img_tokens = self.transformer[i](img_tokens, pcd_tokens, pcd_tokens, k_masks=pcd_masks)
                pcd_tokens = self.transformer[i](pcd_tokens, img_tokens, img_tokens, k_masks=img_masks)
        return _linear(self._cache, self.out_proj, img_tokens), _linear(self._cache, self.out_proj, pcd_tokens)
