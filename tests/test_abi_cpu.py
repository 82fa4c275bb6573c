"""The C-ABI shared library loads without a GPU and exports every entry point include/diffreg_b200.h declares; the
product modules refuse to run without the CUDA path (no CPU / eager fallback)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    text = open(os.path.join(ROOT, "include", "diffreg_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(drg_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = _declared()
    for needed in ("drg_sinkhorn", "drg_dual_softmax", "drg_gemm_nt_tf32", "drg_prep_operand", "drg_match_count", "drg_match_write",
                   "drg_soft_procrustes", "drg_weighted_procrustes", "drg_sinkhorn_shard_local", "drg_sinkhorn_shard_update",
                   "drg_sinkhorn_shard_local_exchange", "drg_sinkhorn_shard_iterate", "drg_p2p_handle_bytes", "drg_p2p_create", "drg_p2p_connect", "drg_p2p_status",
                   "drg_p2p_destroy", "drg_position_code", "drg_gemm_nt_split16", "drg_project_split16", "drg_prep_operand_xyz", "drg_topk_match_count"):
        assert needed in names


def test_library_exports_every_declared_symbol():
    import diffreg_b200
    lib = ctypes.CDLL(diffreg_b200.library_path())
    missing = [n for n in _declared() if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.drg_version() >= 100


def test_workspace_queries_need_no_gpu():
    import diffreg_b200
    lib = diffreg_b200.load_library()
    assert lib.drg_sinkhorn_workspace_bytes(1, 4096, 4096) > 0
    assert lib.drg_sinkhorn_workspace_bytes(1, 4, 20000) == 0          # unsupported width reports 0, never crashes
    assert lib.drg_match_workspace_bytes(2, 100, 50) > 0
    assert lib.drg_soft_procrustes_workspace_bytes(1, 64, 64) >= 2 * 4 * 64 * 64


def test_profile_slot_table_matches_the_library():
    import diffreg_b200
    from diffreg_b200 import _lib
    assert diffreg_b200.load_library().drg_profile_slots() == len(_lib.PROFILE_SLOTS)


def test_ctypes_mirrors_match_the_compiled_structs():
    """The ctypes mirrors of the argument structs (diffreg_b200/_lib.py, and the stub shown in INTEGRATION.md) have the size
    the library was compiled with: a shorter mirror would make the library read past its end."""
    import diffreg_b200
    from diffreg_b200 import _lib
    lib = diffreg_b200.load_library()
    lib.drg_sizeof_sinkhorn_args.restype = ctypes.c_size_t
    lib.drg_sizeof_procrustes_args.restype = ctypes.c_size_t
    assert ctypes.sizeof(_lib.SinkhornArgs) == lib.drg_sizeof_sinkhorn_args()
    assert ctypes.sizeof(_lib.ProcrustesArgs) == lib.drg_sizeof_procrustes_args()
    # the stub in INTEGRATION.md lists the same fields in the same order
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    stub = text[text.index("class SinkhornArgs(ctypes.Structure)"):]
    stub = stub[:stub.index("lib.drg_sinkhorn_workspace_bytes")]
    assert re.findall(r'\("([a-z_0-9A-Z]+)", ctypes', stub) == [f[0] for f in _lib.SinkhornArgs._fields_]


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the behaviour on a box without a GPU")
def test_no_cpu_fallback():
    import diffreg_b200
    ones = torch.ones(1, 4, dtype=torch.bool)
    with pytest.raises(diffreg_b200._lib.DiffRegLibraryError):
        diffreg_b200.log_optimal_transport(torch.zeros(1, 4, 4), torch.tensor(1.0), 3, ones, ones)
    with pytest.raises(diffreg_b200._lib.DiffRegLibraryError):
        diffreg_b200.ops.gemm_nt(torch.zeros(8, 8), torch.zeros(8, 8))


def test_sampler_schedule_matches_reference_golden():
    """Host-side schedule of the sampler (cosine alphas, time pairs, DDIM closed form) against the reference's values."""
    from helpers import load
    from diffreg_b200 import sampler
    g = load("schedule")
    ac = sampler.cosine_alphas_cumprod()
    assert torch.equal(ac, g["alphas_cumprod"])
    pairs = sampler.time_pairs(20)
    assert pairs[0] == (999, 949) and pairs[-1] == (49, 0) and len(pairs) == 20
    # closed form == pred_noise form (pipeline.py:180-190) on random numbers
    gen = torch.Generator().manual_seed(0)
    x_t, x0, nz = (torch.randn(50, generator=gen, dtype=torch.float64) for _ in range(3))
    for t, tn in pairs:
        k0, kt, sg = sampler.ddim_coefficients(ac, t, tn)
        a, an = ac[t], ac[tn]
        pred = (torch.sqrt(1 / a) * x_t - x0) / torch.sqrt(1 / a - 1)
        sigma = ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
        c = (1 - an - sigma ** 2).sqrt()
        want = x0 * an.sqrt() + c * pred + sigma * nz
        assert (k0 * x0 + kt * x_t + sg * nz - want).abs().max() < 1e-9
