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
