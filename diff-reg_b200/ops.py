"""Tensor-level wrappers: torch CUDA tensors in, C-ABI calls out.  torch is used for device
memory, streams and nothing else."""
import functools

import torch

from . import _lib
from ._lib import SinkhornArgs, check, load_library

_workspaces = {}


def _require_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.DiffRegLibraryError("diffreg_b200 operates on CUDA tensors only (no CPU fallback)")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _on_device(fn):
    """Run `fn` with the device of its first CUDA tensor argument current: the launch, the workspace and the stream
    handed to the C ABI then belong to the tensors' device even when another device is current in the caller."""
    @functools.wraps(fn)
    def wrapped(*args, **kwargs):
        for a in list(args) + list(kwargs.values()):
            if torch.is_tensor(a) and a.is_cuda:
                if a.device.index != torch.cuda.current_device():
                    with torch.cuda.device(a.device):
                        return fn(*args, **kwargs)
                break
        return fn(*args, **kwargs)
    return wrapped


def workspace(nbytes, device, tag="default"):
    """Grow-only per-(device, stream, tag) scratch buffer handed to the C ABI."""
    key = (device.index, _stream(), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _as_mask(m, batch=None, length=None, device=None):
    if m is None:      # the reference accepts src_mask = tgt_mask = None (matching.py:10-12): every entry is valid
        return torch.ones(batch, length, dtype=torch.bool, device=device)
    if m.dtype != torch.bool:
        m = m != 0
    return m.contiguous()


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


@_on_device
def sinkhorn(scores, alpha, iters, src_mask, tgt_mask, out_mode="log_full", apply_mask=False, shift=None,
             x_t=None, noise=None, k_x0=0.0, k_xt=0.0, sigma=0.0, want_conf=False, x_min=None, return_potentials=False,
             xt_shift=None, noise_seed=None, noise_offset=0, noise_offset_dev=None, out=None, want_best=False, best_floor=None):
    """Log-domain Sinkhorn with dustbins (drg_sinkhorn).
    want_best: also return the packed row / column bests for match_from_best; best_floor: only confidences above it are
    tracked (pass the matcher's threshold: exact for match_from_best with a threshold >= best_floor, and much cheaper).

    out_mode: 'log_full' -> [B,N+1,M+1] log-assignment; 'conf' -> [B,N,M] exp()[:, :-1, :-1];
              'ddim' -> x_next [B,N,M] (and conf if want_conf); 'none' -> potentials only.
    """
    _require_cuda(scores, alpha, src_mask, tgt_mask, shift, x_t, noise, x_min, xt_shift)
    lib = load_library()
    scores = _f32c(scores)
    B, N, M = scores.shape
    dev = scores.device
    src_mask = _as_mask(src_mask, B, N, dev)
    tgt_mask = _as_mask(tgt_mask, B, M, dev)
    alpha = _f32c(alpha.detach().reshape(()))
    mode = {"log_full": _lib.DRG_OUT_LOG_FULL, "conf": _lib.DRG_OUT_CONF, "ddim": _lib.DRG_OUT_DDIM,
            "none": _lib.DRG_OUT_NONE}[out_mode]
    want_shape = {_lib.DRG_OUT_LOG_FULL: (B, N + 1, M + 1), _lib.DRG_OUT_CONF: (B, N, M), _lib.DRG_OUT_DDIM: (B, N, M)}.get(mode)
    if want_shape is None:
        out = None
    elif out is None:
        out = torch.empty(*want_shape, dtype=torch.float32, device=dev)
    elif tuple(out.shape) != want_shape or out.dtype != torch.float32 or not out.is_contiguous() or not out.is_cuda:
        raise ValueError(f"sinkhorn: out must be a contiguous CUDA fp32 tensor of shape {want_shape}")
    conf = torch.empty(B, N, M, dtype=torch.float32, device=dev) if (mode == _lib.DRG_OUT_DDIM and want_conf) else None
    u = torch.empty(B, N + 1, dtype=torch.float32, device=dev) if return_potentials else None
    v = torch.empty(B, M + 1, dtype=torch.float32, device=dev) if return_potentials else None
    if x_t is not None:
        x_t = _f32c(x_t)
    if noise is not None:
        noise = _f32c(noise)
    rowbest = torch.empty(B, N, dtype=torch.int64, device=dev) if want_best else None
    colbest = torch.empty(B, M, dtype=torch.int64, device=dev) if want_best else None
    nbytes = lib.drg_sinkhorn_workspace_bytes(B, N, M)
    if nbytes == 0:
        raise _lib.DiffRegLibraryError(f"sinkhorn: unsupported shape B={B} N={N} M={M}")
    ws = workspace(nbytes, dev, "sinkhorn")
    a = SinkhornArgs(scores=_ptr(scores), src_mask=_ptr(src_mask), tgt_mask=_ptr(tgt_mask), alpha=_ptr(alpha),
                     shift=_ptr(shift), B=B, N=N, M=M, iters=int(iters), apply_mask=int(bool(apply_mask)),
                     out_mode=mode, out=_ptr(out), u=_ptr(u), v=_ptr(v), x_t=_ptr(x_t), xt_shift=_ptr(xt_shift), noise=_ptr(noise),
                     conf=_ptr(conf), k_x0=float(k_x0), k_xt=float(k_xt), sigma=float(sigma), x_min=_ptr(x_min),
                     gen_noise=int(noise_seed is not None and noise is None), noise_seed=int(noise_seed or 0),
                     noise_offset=int(noise_offset), noise_offset_dev=_ptr(noise_offset_dev), rowbest=_ptr(rowbest),
                     colbest=_ptr(colbest), has_best_floor=int(best_floor is not None),
                     best_floor=float(best_floor) if best_floor is not None else 0.0)
    check(lib.drg_sinkhorn(a, ws.data_ptr(), ws.numel(), _stream()))
    res = [out]
    if conf is not None:
        res.append(conf)
    if return_potentials:
        res += [u, v]
    if want_best:
        res += [rowbest, colbest]
    return res[0] if len(res) == 1 else tuple(res)


@_on_device
def sinkhorn_potentials_per_iteration(scores, alpha, iters, src_mask, tgt_mask, want_out=False):
    """u_t [I,B,N+1], v_t [I,B,M+1] for t = 1..I (what the backward needs), by running the forward with iters = 1..I -- the
    persistent kernel keeps only the final pair.  want_out: also the log-assignment of the last run ([B,N+1,M+1])."""
    B, N, M = scores.shape
    dev = scores.device
    iters = int(iters)
    u_all = torch.empty(iters, B, N + 1, dtype=torch.float32, device=dev)
    v_all = torch.empty(iters, B, M + 1, dtype=torch.float32, device=dev)
    out = None
    for t in range(1, iters + 1):
        last = want_out and t == iters
        res = sinkhorn(scores, alpha, t, src_mask, tgt_mask, out_mode="log_full" if last else "none", return_potentials=True)
        if last:
            out = res[0]
        u_all[t - 1].copy_(res[1])
        v_all[t - 1].copy_(res[2])
    return (u_all, v_all, out) if want_out else (u_all, v_all)


@_on_device
def sinkhorn_backward(scores, alpha, iters, src_mask, tgt_mask, grad_out, potentials=None):
    """(dL/d scores [B,N,M], dL/d alpha 0-dim) of log_optimal_transport given dL/d out [B,N+1,M+1] (drg_sinkhorn_backward).
    potentials: (u_all, v_all) of sinkhorn_potentials_per_iteration when the forward kept them, else they are recomputed."""
    _require_cuda(scores, alpha, src_mask, tgt_mask, grad_out)
    lib = load_library()
    scores = _f32c(scores.detach())
    B, N, M = scores.shape
    dev = scores.device
    sm, tm = _as_mask(src_mask, B, N, dev), _as_mask(tgt_mask, B, M, dev)
    a = _f32c(alpha.detach().reshape(()))
    G = _f32c(grad_out)
    if tuple(G.shape) != (B, N + 1, M + 1):
        raise ValueError(f"sinkhorn_backward: grad_out of shape {tuple(G.shape)} for scores {tuple(scores.shape)}")
    iters = int(iters)
    u_all, v_all = potentials if potentials is not None else sinkhorn_potentials_per_iteration(scores, a, iters, sm, tm)
    gs = torch.empty(B, N, M, dtype=torch.float32, device=dev)
    ga = torch.empty(B, dtype=torch.float32, device=dev)
    nbytes = lib.drg_sinkhorn_backward_workspace_bytes(B, N, M, iters)
    ws = workspace(nbytes, dev, "sinkhorn_backward")
    check(lib.drg_sinkhorn_backward(scores.data_ptr(), a.data_ptr(), sm.data_ptr(), tm.data_ptr(), B, N, M, iters, u_all.data_ptr(),
                                    v_all.data_ptr(), G.data_ptr(), gs.data_ptr(), ga.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return gs, ga.sum()


@_on_device
def dual_softmax(sim, src_mask, tgt_mask, temperature):
    """conf = softmax over src (masked) * softmax over tgt (masked) of sim / temperature."""
    _require_cuda(sim, src_mask, tgt_mask)
    lib = load_library()
    sim = _f32c(sim)
    B, N, M = sim.shape
    src_mask = _as_mask(src_mask)
    tgt_mask = _as_mask(tgt_mask)
    out = torch.empty_like(sim)
    nbytes = lib.drg_sinkhorn_workspace_bytes(B, N, M)
    if nbytes == 0:
        raise _lib.DiffRegLibraryError(f"dual_softmax: unsupported shape B={B} N={N} M={M}")
    ws = workspace(nbytes, sim.device, "sinkhorn")
    check(lib.drg_dual_softmax(sim.data_ptr(), src_mask.data_ptr(), tgt_mask.data_ptr(), B, N, M, float(temperature),
                               out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return out


def split_cols(K):
    """16-bit columns of ONE segment of a split operand of a [.., K] tensor (K rounded up to whole 64-column k-steps)."""
    return (int(K) + 63) // 64 * 64


SPLIT_TAIL = 8   # 16-bit columns behind the two segments of a split operand row: four floats (1 / row scale, row norm, 0, 0)


def split_pitch(K):
    """Row length (16-bit columns) of a split operand of a [.., K] tensor: [seg0 | seg1 | tail]."""
    return 2 * split_cols(K) + SPLIT_TAIL


@_on_device
def gemm_nt(A, B, alpha=1.0, out=None, split3=False, K=None, bias=None):
    """C[b] = alpha * A[b] @ B[b]^T on the tensor cores.  A [batch,N,K] or [N,K]; B likewise.
    split3=False: fp32 operands, one tcgen05 kind::tf32 pass (drg_gemm_nt_tf32; 10 mantissa bits of the operands).
    split3=True: A, B are the 16-bit split operands of prep_operand(split=True) (patterns 0 / 1, torch.int16 [.., split_pitch(K)]);
    the product is fp32-accurate (drg_gemm_nt_split16: lo.hi + hi.lo + hi.hi in kind::f16 with fp32 accumulation, the operands'
    power-of-two row scales undone in the epilogue).  K: the column count of the original operands (default: the padded segment
    width)."""
    _require_cuda(A, B)
    lib = load_library()
    if split3:
        if A.dtype != torch.int16 or B.dtype != torch.int16:
            raise ValueError("gemm_nt(split3=True) takes the int16 split operands of prep_operand(split=True)")
        A, B = A.contiguous(), B.contiguous()
    else:
        A, B = _f32c(A), _f32c(B)
    squeeze = A.dim() == 2
    if squeeze:
        A = A.unsqueeze(0)
        B = B.unsqueeze(0)
    batch, N, KA = A.shape
    M = B.shape[1]
    if B.shape[0] != batch or B.shape[2] != KA:
        raise ValueError(f"gemm_nt: incompatible shapes {tuple(A.shape)} x {tuple(B.shape)}")
    if out is None:
        out = torch.empty(batch, N, M, dtype=torch.float32, device=A.device)
    if split3:
        kc = (KA - SPLIT_TAIL) // 2
        K = kc if K is None else int(K)
        if split_pitch(K) != KA:
            raise ValueError(f"gemm_nt: split operands of width {KA} do not belong to K = {K}")
        if bias is not None:
            bias = _f32c(bias)
            if bias.numel() != M:
                raise ValueError(f"gemm_nt: bias of {bias.numel()} entries for {M} output columns")
            check(lib.drg_gemm_nt_split16_bias(A.data_ptr(), B.data_ptr(), bias.data_ptr(), out.data_ptr(), batch, N, M, K, float(alpha),
                                               _stream()))
        else:
            check(lib.drg_gemm_nt_split16(A.data_ptr(), B.data_ptr(), out.data_ptr(), batch, N, M, K, float(alpha), _stream()))
    else:
        if bias is not None:
            raise ValueError("gemm_nt: bias needs the split operands (split3=True)")
        check(lib.drg_gemm_nt_tf32(A.data_ptr(), B.data_ptr(), out.data_ptr(), batch, N, M, KA, float(alpha), _stream()))
    return out.squeeze(0) if squeeze else out


class LazyPositionCode:
    """A volumetric position code that is NOT materialised: the points and the encoder's constants.  Handed to
    prep_operand / Matching.similarity in place of the [B,N,d,2] (rotary) or [B,N,d] (sinusoidal) tensor, it makes the
    operand staging compute cos / sin from xyz itself (drg_prep_operand_xyz) -- bit-identical, no code tensor in HBM."""

    def __init__(self, xyz, div_term, origin, voxel_size, pe_type, feature_dim):
        self.xyz = xyz
        self.div_term = div_term
        self.origin = tuple(float(v) for v in origin)
        self.voxel_size = float(voxel_size)
        self.pe_type = pe_type
        self.feature_dim = int(feature_dim)


@_on_device
def prep_operand(x, scale=1.0, split=True, pattern=0, pe=None, pe_type=None, want_embedded=False):
    """Positional embedding + scaling + (split=True) the 16-bit hi/lo split of a [..., K] feature tensor (drg_prep_operand).
    pe: the position code tensor, or a LazyPositionCode (then the code is computed inside the kernel from the points).
    Returns out (split: torch.int16 [..., split_pitch(K)] holding the fp16 bit patterns of the row-scaled values, [lo | hi | tail]
    for pattern 0 and [hi | lo | tail] for pattern 1, see include/diffreg_b200.h; else fp32 [..., K]) and, if want_embedded, the
    embedded features."""
    lazy = isinstance(pe, LazyPositionCode)
    _require_cuda(x, None if lazy else pe)
    if split and pattern == 0 and pe is None and scale == 1.0 and not want_embedded:
        staged = getattr(x, "_drg_a16", None)          # written by the kernel that produced x (layernorm(stage=True))
        if staged is not None and staged.shape[:-1] == x.shape[:-1]:
            return staged
    lib = load_library()
    x = _f32c(x)
    K = x.shape[-1]
    rows = x.numel() // K
    code = 0
    if lazy:
        if pe_type is not None and pe_type != pe.pe_type:
            raise ValueError(f"prep_operand: pe_type {pe_type!r} does not match the lazy code's {pe.pe_type!r}")
        code = {"rotary": 1, "sinusoidal": 2}[pe.pe_type]
        xyz = _f32c(pe.xyz)
        if xyz.numel() != rows * 3 or pe.feature_dim != K:
            raise ValueError(f"prep_operand: lazy position code of {xyz.numel() // 3} points x {pe.feature_dim} does not fit {rows} x {K}")
    elif pe is not None:
        code = {"rotary": 1, "sinusoidal": 2}[pe_type]
        pe = _f32c(pe)
        want = (*x.shape, 2) if code == 1 else tuple(x.shape)
        if tuple(pe.shape) != want:
            raise ValueError(f"prep_operand: position code shape {tuple(pe.shape)} != {want}")
    if split:
        out = torch.empty(*x.shape[:-1], split_pitch(K), dtype=torch.int16, device=x.device)
    else:
        out = torch.empty(*x.shape[:-1], K, dtype=torch.float32, device=x.device)
    emb = torch.empty_like(x) if want_embedded else None
    if lazy:
        import ctypes
        origin = (ctypes.c_float * 3)(*pe.origin)
        div = pe.div_term.to(x.device)
        check(lib.drg_prep_operand_xyz(x.data_ptr(), xyz.data_ptr(), div.data_ptr(), origin, pe.voxel_size, code, rows, K, float(scale),
                                       int(bool(split)), int(pattern), _ptr(emb), out.data_ptr(), _stream()))
    else:
        check(lib.drg_prep_operand(x.data_ptr(), _ptr(pe), code, rows, K, float(scale), int(bool(split)), int(pattern),
                                   _ptr(emb), out.data_ptr(), _stream()))
    return (out, emb) if want_embedded else out


@_on_device
def prep_heads(x, heads, pattern, pe=None, pe_type=None, scale=1.0):
    """Per-head split operands of the attention GEMMs: x [B, L, C] (C = heads * d) -> torch.int16 [B * heads, L, split_pitch(d)],
    head-major, with the rotary / additive position code (pe [B, L, C, 2] / [B, L, C]) applied on the way (drg_prep_operand_ext;
    Diff-Reg-4dmatch/models/transformer.py:60-79)."""
    _require_cuda(x, pe)
    lib = load_library()
    x = _f32c(x)
    B, L, C = x.shape
    if C % heads or (C // heads) % 4:
        raise ValueError(f"prep_heads: {C} channels do not split into {heads} heads of a multiple of 4 channels")
    d = C // heads
    code = 0
    if pe is not None:
        code = {"rotary": 1, "sinusoidal": 2}[pe_type]
        pe = _f32c(pe)
        want = (B, L, C, 2) if code == 1 else (B, L, C)
        if tuple(pe.shape) != want:
            raise ValueError(f"prep_heads: position code shape {tuple(pe.shape)} != {want}")
    out = torch.empty(B * heads, L, split_pitch(d), dtype=torch.int16, device=x.device)
    check(lib.drg_prep_operand_ext(x.data_ptr(), _ptr(pe), code, B * L * heads, d, float(scale), 1, int(pattern), 0, int(heads), int(L),
                                   None, out.data_ptr(), _stream()))
    return out


@_on_device
def prep_relu(x, pattern=0):
    """Split operand of relu(x) (the nn.ReLU between the MLP's two linears, transformer.py:33-37)."""
    _require_cuda(x)
    lib = load_library()
    x = _f32c(x)
    K = x.shape[-1]
    out = torch.empty(*x.shape[:-1], split_pitch(K), dtype=torch.int16, device=x.device)
    check(lib.drg_prep_operand_ext(x.data_ptr(), None, 0, x.numel() // K, K, 1.0, 1, int(pattern), 1, 1, 1, None, out.data_ptr(), _stream()))
    return out


@_on_device
def attn_softmax(logits, heads, q_mask, kv_mask, scale, want_operand=True, want_probs=False):
    """Masked, scaled softmax over the keys of logits [B * heads, L, S] (drg_attn_softmax; transformer.py:80-84).
    Returns the probabilities as the left split operand of the P.V GEMM (torch.int16 [B * heads, L, split_pitch(S)]) and / or fp32."""
    _require_cuda(logits, q_mask, kv_mask)
    lib = load_library()
    logits = _f32c(logits)
    BH, L, S = logits.shape
    B = BH // heads
    qm = _as_mask(q_mask) if q_mask is not None and kv_mask is not None else None
    km = _as_mask(kv_mask) if kv_mask is not None else None
    P = torch.empty_like(logits) if want_probs else None
    P16 = torch.empty(BH, L, split_pitch(S), dtype=torch.int16, device=logits.device) if want_operand else None
    check(lib.drg_attn_softmax(logits.data_ptr(), _ptr(qm), _ptr(km), B, heads, L, S, float(scale), _ptr(P), _ptr(P16), _stream()))
    if want_operand and want_probs:
        return P16, P
    return P16 if want_operand else P


@_on_device
def dual_softmax_backward(sim, src_mask, tgt_mask, temperature, grad_conf):
    """dL/d sim [B,N,M] of dual_softmax given dL/d conf (drg_dual_softmax_backward)."""
    _require_cuda(sim, src_mask, tgt_mask, grad_conf)
    lib = load_library()
    sim = _f32c(sim.detach())
    B, N, M = sim.shape
    sm, tm = _as_mask(src_mask, B, N, sim.device), _as_mask(tgt_mask, B, M, sim.device)
    G = _f32c(grad_conf)
    out = torch.empty_like(sim)
    ws = workspace(lib.drg_dual_softmax_backward_workspace_bytes(B, N, M), sim.device, "dual_softmax_backward")
    check(lib.drg_dual_softmax_backward(sim.data_ptr(), sm.data_ptr(), tm.data_ptr(), B, N, M, float(temperature), G.data_ptr(),
                                        out.data_ptr(), ws.data_ptr(), ws.numel(), _stream()))
    return out


FLASH_MAX_HEAD = 176     # widest head drg_attention_split16 takes (shared memory: Q, K, V^T and P tiles of one CTA)


@_on_device
def attention(q16, k16, v, heads, q_mask, kv_mask, scale, d, nsplit=0):
    """softmax(Q K^T * scale + mask) V per head in one kernel (drg_attention_split16).  q16 [B*H, L, split_pitch(d)] / k16
    [B*H, S, split_pitch(d)]: the per-head split operands of prep_heads (patterns 0 / 1); v [B, S, H*d] fp32; masks [B, L] /
    [B, S] bool (True = valid) or None, keys masked for valid queries only.  nsplit: CTAs sharing the keys of one (query tile,
    head) -- 0 = chosen by the library (small grids are split), 1 = never.  Returns [B, L, H*d] fp32."""
    _require_cuda(q16, k16, v, q_mask, kv_mask)
    lib = load_library()
    BH, L, _ = q16.shape
    S = k16.shape[1]
    B = BH // heads
    v = _f32c(v)
    vt16 = torch.empty(BH, d, split_pitch(S), dtype=torch.int16, device=q16.device)  # V^T per head, the keys along the row
    ws = workspace(BH * d * 4, q16.device, tag="attention_vmax")
    check(lib.drg_prep_vt_split16(v.data_ptr(), B, heads, S, int(d), vt16.data_ptr(), ws.data_ptr(), _stream()))
    qm = _as_mask(q_mask, B, L, q16.device) if q_mask is not None else None
    km = _as_mask(kv_mask, B, S, q16.device) if kv_mask is not None else None
    out = torch.empty(B, L, heads * d, dtype=torch.float32, device=q16.device)
    nws = lib.drg_attention_workspace_bytes(B, heads, L, S, int(d), int(nsplit))      # 0: this call keeps one CTA per (query tile, head)
    wsa = workspace(nws, q16.device, tag="attention_split") if nws else None
    check(lib.drg_attention_split16(q16.data_ptr(), k16.data_ptr(), vt16.data_ptr(), _ptr(qm), _ptr(km), B, heads, L, S, int(d),
                                    float(scale), out.data_ptr(), int(nsplit), _ptr(wsa), wsa.numel() if wsa is not None else 0, _stream()))
    return out


@_on_device
def layernorm(x, weight, bias, eps=1e-5, residual=None, pre_add=False, stage=False):
    """residual + LayerNorm(x) (pre_add=False; 4d transformer.py:88,92-94) or LayerNorm(x + residual) (pre_add=True; vision3d
    transformer.py:214,236) over the last dimension (drg_layernorm)."""
    _require_cuda(x, weight, bias, residual)
    lib = load_library()
    x = _f32c(x)
    C = x.shape[-1]
    out = torch.empty_like(x)
    w = _f32c(weight) if weight is not None else None
    b = _f32c(bias) if bias is not None else None
    r = _f32c(residual) if residual is not None else None
    # stage: the kernel also writes the result as the LEFT split operand of the next linear; it rides on the returned tensor and
    # prep_operand() hands it out instead of launching a staging kernel
    a16 = torch.empty(*x.shape[:-1], split_pitch(C), dtype=torch.int16, device=x.device) if (stage and C % 4 == 0 and C <= 1152) else None
    check(lib.drg_layernorm(x.data_ptr(), _ptr(w), _ptr(b), _ptr(r), int(bool(pre_add)), x.numel() // C, C, float(eps), out.data_ptr(),
                            _ptr(a16), _stream()))
    if a16 is not None:
        out._drg_a16 = a16
    return out


@_on_device
def fourier_embed(x, length, k0=0.0, use_pi=True, use_input=False, center=None):
    """vision3d FourierEmbedding.forward of (x - center): x [..., n] -> [..., n * (2 * length + use_input)] (drg_fourier_embed)."""
    _require_cuda(x, center)
    lib = load_library()
    x = _f32c(x)
    n = x.shape[-1]
    c = _f32c(center).reshape(-1) if center is not None else None
    if c is not None and c.numel() != n:
        raise ValueError("fourier_embed: center must hold one value per coordinate")
    out = torch.empty(*x.shape[:-1], n * (2 * int(length) + int(bool(use_input))), dtype=torch.float32, device=x.device)
    check(lib.drg_fourier_embed(x.data_ptr(), _ptr(c), x.numel() // n, n, int(length), float(k0), int(bool(use_pi)), int(bool(use_input)),
                                out.data_ptr(), _stream()))
    return out


@_on_device
def _match(x, mode, mutual, threshold, largest, want_mask, capacity=None):
    """capacity=None: read the match count on the host (as torch.nonzero() does) and return exact-size outputs.
    capacity=int: sync-free; returns (index [capacity,3], vals [capacity], mask, count) with the count on the device."""
    _require_cuda(x)
    lib = load_library()
    x = _f32c(x)
    B, N, M = x.shape
    dev = x.device
    nbytes = lib.drg_match_workspace_bytes(B, N, M)
    ws = workspace(nbytes, dev, "match")
    total = torch.empty(1, dtype=torch.int32, device=dev)
    has_thr = threshold is not None
    thr = float(threshold) if has_thr else 0.0
    args = (x.data_ptr(), B, N, M, int(mode), int(bool(mutual)), int(has_thr), thr, int(bool(largest)), ws.data_ptr(), ws.numel())
    check(lib.drg_match_count(*args, total.data_ptr(), _stream()))
    k = int(total.item()) if capacity is None else int(capacity)
    index = torch.empty(max(k, 1), 3, dtype=torch.int64, device=dev)
    vals = torch.empty(max(k, 1), dtype=torch.float32, device=dev)
    mask = torch.empty(B, N, M, dtype=torch.bool, device=dev) if want_mask else None
    check(lib.drg_match_write(*args, index.data_ptr(), vals.data_ptr(), max(k, 1), _ptr(mask), _stream()))
    if capacity is None:
        return index[:k], vals[:k], mask
    return index, vals, mask, total


@_on_device
def topk_select(score_mat, k, largest=True, threshold=None, mutual=True, row_masks=None, col_masks=None, want_mask=False):
    """(batch_)mutual_topk_select for k >= 1 on a [B,N,M] score tensor (drg_topk_match_*).
    Returns (index [K,3] int64 (b, row, col) in row-major order, scores [K], corr_mat [B,N,M] bool or None)."""
    _require_cuda(score_mat, row_masks, col_masks)
    lib = load_library()
    x = _f32c(score_mat)
    B, N, M = x.shape
    dev = x.device
    rm = _as_mask(row_masks) if row_masks is not None else None
    cm = _as_mask(col_masks) if col_masks is not None else None
    ws = workspace(lib.drg_match_workspace_bytes(B, N, M), dev, "match")
    total = torch.empty(1, dtype=torch.int32, device=dev)
    has_thr = threshold is not None
    args = (x.data_ptr(), B, N, M, int(k), int(bool(mutual)), int(has_thr), float(threshold) if has_thr else 0.0, int(bool(largest)),
            _ptr(rm), _ptr(cm), ws.data_ptr(), ws.numel())
    check(lib.drg_topk_match_count(*args, total.data_ptr(), _stream()))
    n = int(total.item())          # as torch.nonzero() in the reference: the number of hits is read on the host
    index = torch.empty(max(n, 1), 3, dtype=torch.int64, device=dev)
    vals = torch.empty(max(n, 1), dtype=torch.float32, device=dev)
    mask = torch.empty(B, N, M, dtype=torch.bool, device=dev) if want_mask else None
    check(lib.drg_topk_match_write(*args, index.data_ptr(), vals.data_ptr(), max(n, 1), _ptr(mask), _stream()))
    return index[:n], vals[:n], mask


def get_match(conf, thr=0.0, mutual=True, want_mask=True):
    """(index [K,3] int64, mconf [K], mask [B,N,M] bool) -- Matching.get_match."""
    return _match(conf, 0, mutual, thr, True, want_mask)


def top1_select(score_mat, largest=True, threshold=None, mutual=True):
    """mutual_topk_select with k = 1 on a 2-D score matrix: (row_idx [K], col_idx [K], scores [K])."""
    index, vals, _ = _match(score_mat.unsqueeze(0), 1, mutual, threshold, largest, False)
    return index[:, 1].contiguous(), index[:, 2].contiguous(), vals


@_on_device
def soft_procrustes(conf, src_pcd, tgt_pcd, src_mask, tgt_mask, sample_rate, max_condition_num, padded_lengths=False,
                    want_warped=False, want_selection=False):
    """SoftProcrustesLayer.forward on the device (drg_soft_procrustes).
    Returns dict(R, t, R_forwd, t_forwd, condition, solution_mask[, src_warped][, sel_w, sel_src, sel_tgt])."""
    _require_cuda(conf, src_pcd, tgt_pcd, src_mask, tgt_mask)
    lib = load_library()
    conf = _f32c(conf)
    src_pcd = _f32c(src_pcd)
    tgt_pcd = _f32c(tgt_pcd)
    B, N, M = conf.shape
    dev = conf.device
    sm = _as_mask(src_mask) if src_mask is not None else None
    tm = _as_mask(tgt_mask) if tgt_mask is not None else None
    out = dict(R=torch.empty(B, 3, 3, dtype=torch.float32, device=dev), t=torch.empty(B, 3, 1, dtype=torch.float32, device=dev),
               R_forwd=torch.empty(B, 3, 3, dtype=torch.float32, device=dev),
               t_forwd=torch.empty(B, 3, 1, dtype=torch.float32, device=dev),
               condition=torch.empty(B, dtype=torch.float64, device=dev),
               solution_mask=torch.empty(B, dtype=torch.bool, device=dev))
    if want_warped:
        out["src_warped"] = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    k_max = min(int(max(N, M) * float(sample_rate)) + 1, N * M)
    if want_selection:
        out["sel_w"] = torch.empty(B, k_max, dtype=torch.float32, device=dev)
        out["sel_src"] = torch.empty(B, k_max, dtype=torch.int32, device=dev)
        out["sel_tgt"] = torch.empty(B, k_max, dtype=torch.int32, device=dev)
    nbytes = lib.drg_soft_procrustes_workspace_bytes(B, N, M)
    ws = workspace(nbytes, dev, "procrustes")
    a = _lib.ProcrustesArgs(conf=_ptr(conf), src_pcd=_ptr(src_pcd), tgt_pcd=_ptr(tgt_pcd), src_mask=_ptr(sm), tgt_mask=_ptr(tm),
                            B=B, N=N, M=M, sample_rate=float(sample_rate), max_condition_num=float(max_condition_num),
                            padded_lengths=int(bool(padded_lengths)), R=_ptr(out["R"]), t=_ptr(out["t"]),
                            R_forwd=_ptr(out["R_forwd"]), t_forwd=_ptr(out["t_forwd"]), condition=_ptr(out["condition"]),
                            solution_mask=_ptr(out["solution_mask"]), src_warped=_ptr(out.get("src_warped")), K_max=k_max,
                            sel_w=_ptr(out.get("sel_w")), sel_src=_ptr(out.get("sel_src")), sel_tgt=_ptr(out.get("sel_tgt")))
    check(lib.drg_soft_procrustes(a, ws.data_ptr(), ws.numel(), _stream()))
    return out


@_on_device
def weighted_procrustes(X, Y, w, eps=1e-4):
    """batch_weighted_procrustes on the device: X, Y [B,K,3], w [B,K,1] -> (R [B,3,3], t [B,3,1], condition [B] fp64)."""
    _require_cuda(X, Y, w)
    lib = load_library()
    X = _f32c(X)
    Y = _f32c(Y)
    w = _f32c(w).reshape(X.shape[0], X.shape[1])
    B, K, _ = X.shape
    dev = X.device
    R = torch.empty(B, 3, 3, dtype=torch.float32, device=dev)
    t = torch.empty(B, 3, 1, dtype=torch.float32, device=dev)
    cond = torch.empty(B, dtype=torch.float64, device=dev)
    check(lib.drg_weighted_procrustes(X.data_ptr(), Y.data_ptr(), w.data_ptr(), B, K, float(eps), R.data_ptr(), t.data_ptr(),
                                      cond.data_ptr(), _stream()))
    return R, t, cond


@_on_device
def weighted_procrustes_backward(X, Y, w, R, grad_R, grad_t, eps=1e-4):
    """dL/d w [B,K,1] of weighted_procrustes given dL/d R [B,3,3] and dL/d t [B,3,1] (drg_weighted_procrustes_backward)."""
    _require_cuda(X, Y, w, R, grad_R, grad_t)
    lib = load_library()
    X, Y, R = _f32c(X), _f32c(Y), _f32c(R)
    B, K, _ = X.shape
    wf = _f32c(w).reshape(B, K)
    gR = _f32c(grad_R) if grad_R is not None else torch.zeros(B, 3, 3, dtype=torch.float32, device=X.device)
    gt = _f32c(grad_t) if grad_t is not None else torch.zeros(B, 3, 1, dtype=torch.float32, device=X.device)
    gw = torch.empty(B, K, dtype=torch.float32, device=X.device)
    check(lib.drg_weighted_procrustes_backward(X.data_ptr(), Y.data_ptr(), wf.data_ptr(), R.data_ptr(), gR.data_ptr(), gt.data_ptr(), B, K,
                                               float(eps), gw.data_ptr(), _stream()))
    return gw.view(w.shape)


@_on_device
def ransac_correspondence(src_pcd, tgt_pcd, match_pred, distance_threshold=0.05, ransac_n=3, max_iteration=50000, seed=0,
                          want_trials=False):
    """Correspondence RANSAC of every batch element in two launches (drg_ransac_correspondence).

    src_pcd [B,N,3], tgt_pcd [B,M,3], match_pred [C,3] int64 rows (b, i, j) grouped by b (get_match's output)
    -> dict(pose [B,4,4], fitness [B], inlier_rmse [B], best_trial [B] int32, inlier_count [B] int32
            [, trial_count [B,T] int32, trial_err2 [B,T]]); two launches, nothing synchronises with the host."""
    _require_cuda(src_pcd, tgt_pcd, match_pred)
    lib = load_library()
    src, tgt = _f32c(src_pcd), _f32c(tgt_pcd)
    if src.dim() != 3 or tgt.dim() != 3 or src.shape[-1] != 3 or tgt.shape[-1] != 3 or src.shape[0] != tgt.shape[0]:
        raise ValueError("src_pcd / tgt_pcd must be [B,N,3] / [B,M,3]")
    B, N, M = src.shape[0], src.shape[1], tgt.shape[1]
    dev = src.device
    match = match_pred.to(torch.int64).contiguous().view(-1, 3)
    T = int(max_iteration)
    fl = torch.empty(B * 18, dtype=torch.float32, device=dev)  # pose | fitness | rmse
    it = torch.empty(2, B, dtype=torch.int32, device=dev)
    out = {"pose": fl[:B * 16].view(B, 4, 4), "fitness": fl[B * 16:B * 17], "inlier_rmse": fl[B * 17:], "best_trial": it[0],
           "inlier_count": it[1]}
    tc = te = None
    if want_trials:
        tc = out["trial_count"] = torch.empty(B, T, dtype=torch.int32, device=dev)
        te = out["trial_err2"] = torch.empty(B, T, dtype=torch.float32, device=dev)
    nbytes = lib.drg_ransac_workspace_bytes(B, T)
    ws = workspace(nbytes, dev, "ransac")
    check(lib.drg_ransac_correspondence(src.data_ptr(), tgt.data_ptr(), B, N, M, match.data_ptr() if match.numel() else None,
                                        match.shape[0], None, float(distance_threshold), int(ransac_n), T, int(seed) & (2 ** 64 - 1),
                                        out["pose"].data_ptr(), out["fitness"].data_ptr(), out["inlier_rmse"].data_ptr(),
                                        out["best_trial"].data_ptr(), out["inlier_count"].data_ptr(),
                                        tc.data_ptr() if tc is not None else None, te.data_ptr() if te is not None else None,
                                        ws.data_ptr(), nbytes, _stream()))
    return out


@_on_device
def sigmoid(x):
    _require_cuda(x)
    x = _f32c(x)
    y = torch.empty_like(x)
    check(load_library().drg_sigmoid(x.data_ptr(), y.data_ptr(), x.numel(), _stream()))
    return y


@_on_device
def min_value(x):
    """Global minimum as a 1-element device tensor (no host read)."""
    _require_cuda(x)
    x = _f32c(x)
    out = torch.empty(1, dtype=torch.float32, device=x.device)
    scratch = torch.empty(1, dtype=torch.int32, device=x.device)
    check(load_library().drg_min_value(x.data_ptr(), x.numel(), out.data_ptr(), scratch.data_ptr(), _stream()))
    return out


@_on_device
def counter_add(counter, inc=1):
    """counter (1-element int64 CUDA tensor) += inc, on the stream."""
    _require_cuda(counter)
    check(load_library().drg_counter_add(counter.data_ptr(), int(inc), _stream()))


class ShardedSinkhornState:
    """The rank-local state of a row-sharded Sinkhorn (drg_sinkhorn_shard_*): local rows of the scores, local src mask,
    replicated tgt mask, and a private workspace holding u (local rows), v (replicated) and the constants."""

    def __init__(self, scores_local, alpha, src_mask_local, tgt_mask, apply_mask=False, shift=None):
        _require_cuda(scores_local, alpha, src_mask_local, tgt_mask, shift)
        self.lib = load_library()
        self.scores = _f32c(scores_local)
        self.B, self.N, self.M = self.scores.shape
        self.src_mask = _as_mask(src_mask_local)
        self.tgt_mask = _as_mask(tgt_mask)
        self.alpha = _f32c(alpha.detach().reshape(()))
        self.shift = shift
        self.apply_mask = bool(apply_mask)
        nbytes = self.lib.drg_sinkhorn_workspace_bytes(self.B, self.N, self.M)
        if nbytes == 0:
            raise _lib.DiffRegLibraryError(f"sharded sinkhorn: unsupported local shape {tuple(self.scores.shape)}")
        self.ws = torch.empty(nbytes, dtype=torch.uint8, device=self.scores.device)   # private: potentials live here
        self.partial = torch.empty(self.B, self.M + 1, 2, dtype=torch.float32, device=self.scores.device)

    def _args(self, out_mode=_lib.DRG_OUT_NONE, out=None):
        return SinkhornArgs(scores=_ptr(self.scores), src_mask=_ptr(self.src_mask), tgt_mask=_ptr(self.tgt_mask), alpha=_ptr(self.alpha),
                            shift=_ptr(self.shift), B=self.B, N=self.N, M=self.M, iters=1, apply_mask=int(self.apply_mask),
                            out_mode=out_mode, out=_ptr(out))

    def local_counts(self):
        """[B,2] int32: (valid src rows of THIS rank, valid tgt columns); sum column 0 over the ranks before begin()."""
        return torch.stack((self.src_mask.sum(dim=1), self.tgt_mask.sum(dim=1)), dim=1).to(torch.int32)

    def begin(self, global_counts):
        gc = global_counts.to(torch.int32).contiguous()
        check(self.lib.drg_sinkhorn_shard_begin(self._args(), gc.data_ptr(), self.ws.data_ptr(), self.ws.numel(), _stream()))

    def local(self):
        """Row pass over the local rows with the current v; returns this rank's column partials [B, M+1, 2]."""
        check(self.lib.drg_sinkhorn_shard_local(self._args(), self.ws.data_ptr(), self.ws.numel(), self.partial.data_ptr(), _stream()))
        return self.partial

    def local_exchange(self, comm):
        """Row pass + in-kernel peer-to-peer all-reduce of the column partials + update of v (one call per iteration;
        `comm` is distributed.P2PComm.handle)."""
        check(self.lib.drg_sinkhorn_shard_local_exchange(self._args(), self.ws.data_ptr(), self.ws.numel(), comm, _stream()))

    def iterate_exchange(self, comm, iters):
        """`iters` iterations of local_exchange enqueued by ONE library call."""
        check(self.lib.drg_sinkhorn_shard_iterate(self._args(), self.ws.data_ptr(), self.ws.numel(), comm, int(iters), _stream()))

    def update(self, reduced):
        reduced = _f32c(reduced)
        check(self.lib.drg_sinkhorn_shard_update(self._args(), self.ws.data_ptr(), self.ws.numel(), reduced.data_ptr(), _stream()))

    def final(self, out_mode="conf"):
        mode = {"log_full": _lib.DRG_OUT_LOG_FULL, "conf": _lib.DRG_OUT_CONF}[out_mode]
        shape = (self.B, self.N + 1, self.M + 1) if mode == _lib.DRG_OUT_LOG_FULL else (self.B, self.N, self.M)
        out = torch.empty(*shape, dtype=torch.float32, device=self.scores.device)
        check(self.lib.drg_sinkhorn_shard_final(self._args(mode, out), self.ws.data_ptr(), self.ws.numel(), _stream()))
        return out


@_on_device
def match_from_best(rowbest, colbest, M, threshold=None, capacity=None):
    """Mutual top-1 matches from the packed bests of sinkhorn(..., want_best=True) (drg_match_from_best).
    capacity=None reads the count on the host and returns exact-size (index [K,3], vals [K]); otherwise returns
    (index [capacity,3], vals [capacity], count) with the count on the device."""
    _require_cuda(rowbest, colbest)
    lib = load_library()
    B, N = rowbest.shape
    dev = rowbest.device
    cap = int(capacity) if capacity is not None else B * min(N, M)
    index = torch.empty(max(cap, 1), 3, dtype=torch.int64, device=dev)
    vals = torch.empty(max(cap, 1), dtype=torch.float32, device=dev)
    total = torch.empty(1, dtype=torch.int32, device=dev)
    has_thr = threshold is not None
    check(lib.drg_match_from_best(rowbest.data_ptr(), colbest.data_ptr(), B, N, M, int(has_thr), float(threshold) if has_thr else 0.0,
                                  index.data_ptr(), vals.data_ptr(), max(cap, 1), total.data_ptr(), _stream()))
    if capacity is None:
        k = int(total.item())
        return index[:k], vals[:k]
    return index, vals, total


@_on_device
def sinkhorn_soft_procrustes(scores, alpha, iters, src_mask, tgt_mask, src_pcd, tgt_pcd, sample_rate, max_condition_num,
                             padded_lengths=False, apply_mask=True, shift=None, want_warped=True):
    """get_warped_from_noising_matching in one call (drg_sinkhorn_soft_procrustes): Sinkhorn on the sampler state, then
    SoftProcrustes + warp straight from the potentials -- the confidence matrix is never written.
    Returns the same dict as soft_procrustes()."""
    _require_cuda(scores, alpha, src_mask, tgt_mask, src_pcd, tgt_pcd, shift)
    lib = load_library()
    scores = _f32c(scores)
    src_pcd = _f32c(src_pcd)
    tgt_pcd = _f32c(tgt_pcd)
    B, N, M = scores.shape
    dev = scores.device
    sm = _as_mask(src_mask)
    tm = _as_mask(tgt_mask)
    alpha = _f32c(alpha.detach().reshape(()))
    out = dict(R=torch.empty(B, 3, 3, dtype=torch.float32, device=dev), t=torch.empty(B, 3, 1, dtype=torch.float32, device=dev),
               R_forwd=torch.empty(B, 3, 3, dtype=torch.float32, device=dev),
               t_forwd=torch.empty(B, 3, 1, dtype=torch.float32, device=dev),
               condition=torch.empty(B, dtype=torch.float64, device=dev),
               solution_mask=torch.empty(B, dtype=torch.bool, device=dev))
    if want_warped:
        out["src_warped"] = torch.empty(B, N, 3, dtype=torch.float32, device=dev)
    n_s = lib.drg_sinkhorn_workspace_bytes(B, N, M)
    if n_s == 0:
        raise _lib.DiffRegLibraryError(f"sinkhorn: unsupported shape B={B} N={N} M={M}")
    ws_s = workspace(n_s, dev, "sinkhorn")
    ws_p = workspace(lib.drg_soft_procrustes_workspace_bytes(B, N, M), dev, "procrustes")
    s = SinkhornArgs(scores=_ptr(scores), src_mask=_ptr(sm), tgt_mask=_ptr(tm), alpha=_ptr(alpha), shift=_ptr(shift), B=B, N=N, M=M,
                     iters=int(iters), apply_mask=int(bool(apply_mask)), out_mode=_lib.DRG_OUT_NONE)
    k_max = min(int(max(N, M) * float(sample_rate)) + 1, N * M)
    a = _lib.ProcrustesArgs(conf=None, src_pcd=_ptr(src_pcd), tgt_pcd=_ptr(tgt_pcd), src_mask=_ptr(sm), tgt_mask=_ptr(tm), B=B, N=N,
                            M=M, sample_rate=float(sample_rate), max_condition_num=float(max_condition_num),
                            padded_lengths=int(bool(padded_lengths)), R=_ptr(out["R"]), t=_ptr(out["t"]),
                            R_forwd=_ptr(out["R_forwd"]), t_forwd=_ptr(out["t_forwd"]), condition=_ptr(out["condition"]),
                            solution_mask=_ptr(out["solution_mask"]), src_warped=_ptr(out.get("src_warped")), K_max=k_max)
    check(lib.drg_sinkhorn_soft_procrustes(s, a, ws_s.data_ptr(), ws_s.numel(), ws_p.data_ptr(), ws_p.numel(), _stream()))
    return out


_split_out_cache = {}


@_on_device
def project_pair_split(src_feats, tgt_feats, w_operand, out_dim, scale, want_plain=False):
    """Both projections of Matching.forward and the operand preparation of the similarity GEMM in two launches:
    drg_prep_operand_pair (16-bit hi/lo split of src | tgt features) + drg_project_split16 (tensor-core GEMM against the
    prepared weight, epilogue writes scale * (x W^T) already split: src rows as the left operand, tgt rows as the right one).
    src_feats [B,N,C], tgt_feats [B,M,C], w_operand = prep_operand(W, split=True, pattern=1) [C_out, split_pitch(C)].
    Returns (src_operand [B,N,split_pitch(C_out)], tgt_operand [B,M,split_pitch(C_out)] (int16), plain [B*(N+M), C_out] or None)."""
    _require_cuda(src_feats, tgt_feats, w_operand)
    lib = load_library()
    src_feats = _f32c(src_feats)
    tgt_feats = _f32c(tgt_feats)
    B, N, C = src_feats.shape
    M = tgt_feats.shape[1]
    dev = src_feats.device
    rows_a, rows_b = B * N, B * M
    a16 = torch.empty(rows_a + rows_b, split_pitch(C), dtype=torch.int16, device=dev)
    check(lib.drg_prep_operand_pair(src_feats.data_ptr(), rows_a, 0, tgt_feats.data_ptr(), rows_b, 0, C, 1.0, 1, a16.data_ptr(), _stream()))
    kc_out, pitch_out = split_cols(out_dim), split_pitch(out_dim)
    if kc_out == out_dim:
        split = torch.empty(rows_a + rows_b, pitch_out, dtype=torch.int16, device=dev)
    else:
        split = torch.zeros(rows_a + rows_b, pitch_out, dtype=torch.int16, device=dev)   # the epilogue leaves the padding columns alone
    plain = torch.empty(rows_a + rows_b, out_dim, dtype=torch.float32, device=dev) if want_plain else None
    check(lib.drg_project_split16(a16.data_ptr(), w_operand.data_ptr(), rows_a + rows_b, rows_a, out_dim, C, float(scale), _ptr(plain),
                                  split.data_ptr(), _stream()))
    return split[:rows_a].view(B, N, pitch_out), split[rows_a:].view(B, M, pitch_out), plain
