"""Golden vectors for the volumetric position encoding from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden_pe.py
Imports Diff-Reg-4dmatch/models/position_encoding.py as it lies, runs VolumetricPositionEncoding.forward and
embed_pos on seeded points / features on the CPU, and stores inputs + outputs as tests/golden/pe_*.npz."""
import importlib
import os
import sys
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, f"{REF}/Diff-Reg-4dmatch")
pe_mod = importlib.import_module("models.position_encoding")
torch.set_num_threads(1)

VOL_BNDS = [[-3.6, -2.4, 1.14], [1.093, 0.78, 2.92]]     # configs/test/4dmatch.yaml:49-50


@torch.no_grad()
def run(tag, B, N, C, pe_type, seed):
    g = torch.Generator().manual_seed(seed)
    lo, hi = torch.tensor(VOL_BNDS[0]), torch.tensor(VOL_BNDS[1])
    xyz = lo + (hi - lo) * torch.rand(B, N, 3, generator=g)
    x = torch.randn(B, N, C, generator=g)
    vol = pe_mod.VolumetricPositionEncoding(SimpleNamespace(feature_dim=C, vol_bnds=VOL_BNDS, voxel_size=0.04, pe_type=pe_type))
    code = vol(xyz)
    emb = pe_mod.VolumetricPositionEncoding.embed_pos(pe_type, x, code)
    out = dict(xyz=xyz, x=x, code=code, embedded=emb, feature_dim=C, voxel_size=0.04, vol_origin=np.asarray(VOL_BNDS[0], dtype=np.float32),
               pe_type=pe_type)
    np.savez_compressed(os.path.join(HERE, tag + ".npz"), **{k: (v.numpy() if isinstance(v, torch.Tensor) else np.asarray(v)) for k, v in out.items()})
    print("wrote", tag, tuple(code.shape))


run("pe_rotary_528", 1, 37, 528, "rotary", 1)           # 4DMatch / 3DMatch coarse feature width
run("pe_sinusoidal_432", 2, 19, 432, "sinusoidal", 2)
run("pe_rotary_36_b2", 2, 64, 36, "rotary", 3)
