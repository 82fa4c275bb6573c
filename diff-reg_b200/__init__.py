"""diffreg_b200 -- B200 (sm_100a) kernels for Diff-Reg's per-step coarse matching-matrix
update, behind the reference's own PyTorch module API.

Public surface (mirrors the reference):
    Matching, log_optimal_transport, mutual_topk_select,
    batch_mutual_topk_select                                 (matching.py)
    Matching2D3D                                             (2D-3D flavour head)
    SoftProcrustesLayer                                      (procrustes.py)
    VolumetricPositionEncoding                               (position_encoding.py; next-row widening)
    GeometryAttentionLayer, RepositioningTransformer         (transformer.py; the denoising transformer)
    CrossModalFusionModule                                   (fusion.py; the 2D-3D flavour's fusion / denoising transformer)
    DenoisingSampler                                         (fused per-step driver)
    HostStepPipeline                                         (host buffers in / results out, copies overlapped, graph replay)
    ransac_pose_estimation, ransac_regist_coarse             (registration.py; correspondence RANSAC)
    RowShardedSinkhorn, shard_rows, shard_units              (multi-GPU paths, distributed.py)
Everything computes through libdiffreg_b200.so (C ABI, include/diffreg_b200.h).  There is
no CPU or eager-PyTorch fallback: a missing library or a non-CUDA tensor raises.
"""
from . import _lib  # noqa: F401
from ._lib import library_path, load_library, launch_count  # noqa: F401

__all__ = ["library_path", "load_library", "launch_count"]


def __getattr__(name):
    # lazy: the modules below import torch
    if name in ("Matching", "Matching2D3D", "log_optimal_transport", "mutual_topk_select", "batch_mutual_topk_select"):
        from . import matching
        return getattr(matching, name)
    if name == "SoftProcrustesLayer":
        from . import procrustes
        return procrustes.SoftProcrustesLayer
    if name in ("DenoisingSampler",):
        from . import sampler
        return getattr(sampler, name)
    if name == "VolumetricPositionEncoding":
        from . import position_encoding
        return position_encoding.VolumetricPositionEncoding
    if name in ("GeometryAttentionLayer", "RepositioningTransformer"):
        from . import transformer
        return getattr(transformer, name)
    if name == "CrossModalFusionModule":
        from . import fusion
        return fusion.CrossModalFusionModule
    if name in ("ransac_pose_estimation", "ransac_regist_coarse"):
        from . import registration
        return getattr(registration, name)
    if name == "HostStepPipeline":
        from . import hostpipe
        return hostpipe.HostStepPipeline
    if name in ("RowShardedSinkhorn", "EmulatedRowShards", "P2PComm", "shard_rows", "shard_units", "lse_allreduce", "lse_combine"):
        from . import distributed
        return getattr(distributed, name)
    if name == "ops":
        import importlib
        return importlib.import_module(__name__ + ".ops")
    raise AttributeError(name)
