"""Drop-in for the reference's ``models/position_encoding.py``.

``VolumetricPositionEncoding(config)`` reads the same attributes (``feature_dim, vol_bnds, voxel_size, pe_type``) and
returns the same tensors as Diff-Reg-4dmatch/models/position_encoding.py:5-87 -- rotary: ``[B,N,d,2]`` = (cos, sin),
sinusoidal: ``[B,N,d]`` -- computed by ``drg_position_code``; ``embed_rotary`` / ``embed_pos`` go through
``drg_prep_operand``.  Shadow the reference module with

    # models/position_encoding.py
    from diffreg_b200.position_encoding import VolumetricPositionEncoding      # noqa: F401

CUDA tensors only; forward-only (position codes carry no gradient in the reference either, :83-84)."""
import ctypes
import math

import torch
from torch import nn

from . import ops
from ._lib import check, load_library

_PE_CODE = {"rotary": 1, "sinusoidal": 2}


class VolumetricPositionEncoding(nn.Module):
    def __init__(self, config):
        super().__init__()
        self.feature_dim = config.feature_dim
        self.vol_bnds = config.vol_bnds
        self.voxel_size = config.voxel_size
        self.vol_origin = self.vol_bnds[0]
        self.pe_type = config.pe_type
        if self.pe_type not in _PE_CODE:
            raise KeyError(self.pe_type)
        if self.feature_dim % 6:
            raise ValueError("VolumetricPositionEncoding: feature_dim must be a multiple of 6 (x | y | z thirds of sin/cos pairs)")
        d3 = self.feature_dim // 3
        # position_encoding.py:58 -- computed once, on the CPU in fp32 exactly as the reference does on the fly
        div = torch.exp(torch.arange(0, d3, 2, dtype=torch.float) * (-math.log(10000.0) / d3))
        self.register_buffer("div_term", div, persistent=False)

    def voxelize(self, xyz):
        """position_encoding.py:16-24."""
        origin = torch.as_tensor(self.vol_origin, dtype=torch.float32, device=xyz.device).view(1, 1, -1)
        return (xyz - origin) / self.voxel_size

    @torch.no_grad()
    def forward(self, XYZ):
        if not XYZ.is_cuda:
            from ._lib import DiffRegLibraryError
            raise DiffRegLibraryError("diffreg_b200 operates on CUDA tensors only (no CPU fallback)")
        lib = load_library()
        xyz = XYZ.detach().float().contiguous()
        B, N, _ = xyz.shape
        d = self.feature_dim
        div = self.div_term.to(xyz.device)
        rotary = self.pe_type == "rotary"
        out = torch.empty((B, N, d, 2) if rotary else (B, N, d), dtype=torch.float32, device=xyz.device)
        origin = (ctypes.c_float * 3)(*[float(v) for v in self.vol_origin])
        check(lib.drg_position_code(xyz.data_ptr(), div.data_ptr(), B * N, d, origin, float(self.voxel_size), _PE_CODE[self.pe_type],
                                    out.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return out

    def lazy(self, XYZ):
        """The position code of XYZ [B,N,3] WITHOUT materialising it: an ops.LazyPositionCode that Matching.similarity /
        ops.prep_operand accept in place of forward()'s tensor; the operand staging of the similarity GEMM then computes
        cos / sin from the points itself (bit-identical to forward() + embed_pos, no [B,N,d,2] tensor in HBM)."""
        return ops.LazyPositionCode(XYZ.detach().float().contiguous(), self.div_term, self.vol_origin, self.voxel_size, self.pe_type,
                                    self.feature_dim)

    @staticmethod
    def embed_rotary(x, cos, sin):
        """position_encoding.py:26-35."""
        return ops.prep_operand(x, 1.0, False, 0, pe=torch.stack((cos, sin), dim=-1), pe_type="rotary")

    @staticmethod
    def embed_pos(pe_type, x, pe):
        """position_encoding.py:37-46."""
        if pe_type not in _PE_CODE:
            raise KeyError(pe_type)
        return ops.prep_operand(x, 1.0, False, 0, pe=pe, pe_type=pe_type)
