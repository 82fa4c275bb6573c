"""Tensor-level wrappers: torch CUDA tensors in, C-ABI calls out.  torch is used for device
memory, streams and nothing else."""
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


def workspace(nbytes, device, tag="default"):
    """Grow-only per-(device, stream, tag) scratch buffer handed to the C ABI."""
    key = (device.index, _stream(), tag)
    buf = _workspaces.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device)
        _workspaces[key] = buf
    return buf


def _as_mask(m):
    if m.dtype != torch.bool:
        m = m != 0
    return m.contiguous()


def _f32c(t):
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def sinkhorn(scores, alpha, iters, src_mask, tgt_mask, out_mode="log_full", apply_mask=False, shift=None,
             x_t=None, noise=None, k_x0=0.0, k_xt=0.0, sigma=0.0, want_conf=False, x_min=None, return_potentials=False):
    """Log-domain Sinkhorn with dustbins (drg_sinkhorn).

    out_mode: 'log_full' -> [B,N+1,M+1] log-assignment; 'conf' -> [B,N,M] exp()[:, :-1, :-1];
              'ddim' -> x_next [B,N,M] (and conf if want_conf); 'none' -> potentials only.
    """
    _require_cuda(scores, alpha, src_mask, tgt_mask, shift, x_t, noise, x_min)
    lib = load_library()
    scores = _f32c(scores)
    B, N, M = scores.shape
    dev = scores.device
    src_mask = _as_mask(src_mask)
    tgt_mask = _as_mask(tgt_mask)
    alpha = _f32c(alpha.detach().reshape(()))
    mode = {"log_full": _lib.DRG_OUT_LOG_FULL, "conf": _lib.DRG_OUT_CONF, "ddim": _lib.DRG_OUT_DDIM,
            "none": _lib.DRG_OUT_NONE}[out_mode]
    out = None
    if mode == _lib.DRG_OUT_LOG_FULL:
        out = torch.empty(B, N + 1, M + 1, dtype=torch.float32, device=dev)
    elif mode in (_lib.DRG_OUT_CONF, _lib.DRG_OUT_DDIM):
        out = torch.empty(B, N, M, dtype=torch.float32, device=dev)
    conf = torch.empty(B, N, M, dtype=torch.float32, device=dev) if (mode == _lib.DRG_OUT_DDIM and want_conf) else None
    u = torch.empty(B, N + 1, dtype=torch.float32, device=dev) if return_potentials else None
    v = torch.empty(B, M + 1, dtype=torch.float32, device=dev) if return_potentials else None
    if x_t is not None:
        x_t = _f32c(x_t)
    if noise is not None:
        noise = _f32c(noise)
    nbytes = lib.drg_sinkhorn_workspace_bytes(B, N, M)
    if nbytes == 0:
        raise _lib.DiffRegLibraryError(f"sinkhorn: unsupported shape B={B} N={N} M={M}")
    ws = workspace(nbytes, dev, "sinkhorn")
    a = SinkhornArgs(scores=_ptr(scores), src_mask=_ptr(src_mask), tgt_mask=_ptr(tgt_mask), alpha=_ptr(alpha),
                     shift=_ptr(shift), B=B, N=N, M=M, iters=int(iters), apply_mask=int(bool(apply_mask)),
                     out_mode=mode, out=_ptr(out), u=_ptr(u), v=_ptr(v), x_t=_ptr(x_t), noise=_ptr(noise),
                     conf=_ptr(conf), k_x0=float(k_x0), k_xt=float(k_xt), sigma=float(sigma), x_min=_ptr(x_min))
    check(lib.drg_sinkhorn(a, ws.data_ptr(), ws.numel(), _stream()))
    res = [out]
    if conf is not None:
        res.append(conf)
    if return_potentials:
        res += [u, v]
    return res[0] if len(res) == 1 else tuple(res)


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
