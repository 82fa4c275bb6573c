"""Shared helpers for the parity tests: golden loading and the north_star tolerances."""
import glob
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")

# BASELINE.json north_star tolerances
TOL_LOG = 1e-4      # abs, log-matrix and confidences, fp32
TOL_ROT = 1e-5      # rad
TOL_TRANS = 1e-5    # m
MARGIN = 1e-5       # top-1 margin above which indices must be bit-exact


def load(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    out = {}
    for k in z.files:
        a = z[k]
        out[k] = torch.from_numpy(a) if a.ndim > 0 and a.dtype.kind in "fbiu" else a[()] if a.ndim == 0 else a
    return out


def names(prefix):
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def rot_angle(Ra, Rb):
    """Geodesic angle between rotation matrices (rad), fp64."""
    D = Ra.double() @ Rb.double().transpose(-1, -2)
    # robust for tiny angles: use the skew part
    skew = 0.5 * (D - D.transpose(-1, -2))
    s = torch.sqrt(skew[..., 2, 1] ** 2 + skew[..., 0, 2] ** 2 + skew[..., 1, 0] ** 2)
    c = 0.5 * (D.diagonal(dim1=-2, dim2=-1).sum(-1) - 1.0)
    return torch.atan2(s, c)


def finite_close(a, b, tol):
    """max |a-b| over entries where either is finite; -inf must coincide."""
    a = a.double()
    b = b.double()
    same_inf = (torch.isinf(a) == torch.isinf(b)).all()
    m = torch.isfinite(a) & torch.isfinite(b)
    err = (a[m] - b[m]).abs().max().item() if m.any() else 0.0
    return bool(same_inf) and err <= tol, err


def check_top1_pairs(conf_ref, rows, cols, mutual, margin=MARGIN):
    """Top-1 row/column selection against a reference confidence matrix [N,M], honouring the north_star rule:
    indices must be identical wherever the reference's top-1 margin exceeds `margin`; inside the margin (ties,
    e.g. all-zero padded rows where torch.topk's pick is implementation defined) any near-maximal entry is allowed.
    Returns (ok, message)."""
    conf_ref = conf_ref.double()
    N, M = conf_ref.shape
    got = set(zip(rows.tolist(), cols.tolist()))
    rmax, rarg = conf_ref.max(dim=1)
    cmax, carg = conf_ref.max(dim=0)
    r2 = conf_ref.topk(min(2, M), dim=1)[0]
    c2 = conf_ref.topk(min(2, N), dim=0)[0]
    r_clear = (r2[:, 0] - r2[:, -1] > margin) if M > 1 else torch.ones(N, dtype=torch.bool)
    c_clear = (c2[0] - c2[-1] > margin) if N > 1 else torch.ones(M, dtype=torch.bool)
    near_row = conf_ref >= (rmax[:, None] - margin)
    near_col = conf_ref >= (cmax[None, :] - margin)
    row_hit = torch.zeros_like(near_row)
    row_hit[torch.arange(N), rarg] = True
    col_hit = torch.zeros_like(near_col)
    col_hit[carg, torch.arange(M)] = True
    if mutual:
        required = row_hit & col_hit & r_clear[:, None] & c_clear[None, :]
        allowed = near_row & near_col
    else:
        required = (row_hit & r_clear[:, None]) | (col_hit & c_clear[None, :])
        allowed = near_row | near_col
    req = set(map(tuple, required.nonzero().tolist()))
    alw = set(map(tuple, allowed.nonzero().tolist()))
    if not req <= got:
        return False, f"missing required pairs {sorted(req - got)[:5]}"
    if not got <= alw:
        return False, f"pairs outside the margin {sorted(got - alw)[:5]}"
    order = sorted(got)
    if order != list(zip(rows.tolist(), cols.tolist())):
        return False, "pairs are not in row-major order"
    return True, ""
