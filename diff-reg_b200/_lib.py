"""ctypes binding of libdiffreg_b200.so (the C ABI in include/diffreg_b200.h)."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libdiffreg_b200.so")
_lib = None

DRG_OUT_LOG_FULL, DRG_OUT_CONF, DRG_OUT_DDIM, DRG_OUT_NONE = 0, 1, 2, 3

c_void_p, c_int, c_float, c_size_t = ctypes.c_void_p, ctypes.c_int, ctypes.c_float, ctypes.c_size_t


class SinkhornArgs(ctypes.Structure):
    """struct drg_sinkhorn_args"""
    _fields_ = [
        ("scores", c_void_p), ("src_mask", c_void_p), ("tgt_mask", c_void_p), ("alpha", c_void_p), ("shift", c_void_p),
        ("B", c_int), ("N", c_int), ("M", c_int), ("iters", c_int), ("apply_mask", c_int),
        ("out_mode", c_int), ("out", c_void_p), ("u", c_void_p), ("v", c_void_p),
        ("x_t", c_void_p), ("xt_shift", c_void_p), ("noise", c_void_p), ("conf", c_void_p),
        ("k_x0", c_float), ("k_xt", c_float), ("sigma", c_float), ("x_min", c_void_p),
        ("gen_noise", c_int), ("noise_seed", ctypes.c_ulonglong), ("noise_offset", ctypes.c_ulonglong),
        ("noise_offset_dev", c_void_p), ("rowbest", c_void_p), ("colbest", c_void_p),
        ("has_best_floor", c_int), ("best_floor", c_float),
    ]


class ProcrustesArgs(ctypes.Structure):
    """struct drg_procrustes_args"""
    _fields_ = [
        ("conf", c_void_p), ("src_pcd", c_void_p), ("tgt_pcd", c_void_p), ("src_mask", c_void_p), ("tgt_mask", c_void_p),
        ("B", c_int), ("N", c_int), ("M", c_int), ("sample_rate", c_float), ("max_condition_num", c_float),
        ("padded_lengths", c_int),
        ("R", c_void_p), ("t", c_void_p), ("R_forwd", c_void_p), ("t_forwd", c_void_p), ("condition", c_void_p),
        ("solution_mask", c_void_p), ("src_warped", c_void_p),
        ("K_max", c_int), ("sel_w", c_void_p), ("sel_src", c_void_p), ("sel_tgt", c_void_p),
    ]


class DiffRegLibraryError(RuntimeError):
    pass


def library_path():
    return _LIB_PATH


def _declare(lib):
    lib.drg_version.restype = c_int
    lib.drg_last_error.restype = ctypes.c_char_p
    lib.drg_launch_count.restype = ctypes.c_ulonglong
    lib.drg_sinkhorn_workspace_bytes.restype = c_size_t
    lib.drg_sinkhorn_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.drg_sinkhorn.restype = c_int
    lib.drg_sinkhorn.argtypes = [ctypes.POINTER(SinkhornArgs), c_void_p, c_size_t, c_void_p]
    for name, extra in (("drg_sinkhorn_shard_begin", [c_void_p]), ("drg_sinkhorn_shard_local", None), ("drg_sinkhorn_shard_update", None),
                        ("drg_sinkhorn_shard_final", None)):
        fn = getattr(lib, name)
        fn.restype = c_int
    lib.drg_sinkhorn_shard_begin.argtypes = [ctypes.POINTER(SinkhornArgs), c_void_p, c_void_p, c_size_t, c_void_p]
    lib.drg_sinkhorn_shard_local.argtypes = [ctypes.POINTER(SinkhornArgs), c_void_p, c_size_t, c_void_p, c_void_p]
    lib.drg_sinkhorn_shard_update.argtypes = [ctypes.POINTER(SinkhornArgs), c_void_p, c_size_t, c_void_p, c_void_p]
    lib.drg_sinkhorn_shard_final.argtypes = [ctypes.POINTER(SinkhornArgs), c_void_p, c_size_t, c_void_p]
    lib.drg_sinkhorn_shard_local_exchange.restype = c_int
    lib.drg_sinkhorn_shard_local_exchange.argtypes = [ctypes.POINTER(SinkhornArgs), c_void_p, c_size_t, c_void_p, c_void_p]
    lib.drg_sinkhorn_shard_iterate.restype = c_int
    lib.drg_sinkhorn_shard_iterate.argtypes = [ctypes.POINTER(SinkhornArgs), c_void_p, c_size_t, c_void_p, c_int, c_void_p]
    lib.drg_p2p_handle_bytes.restype = c_size_t
    lib.drg_p2p_create.restype = c_int
    lib.drg_p2p_create.argtypes = [c_int, c_int, c_size_t, c_int, ctypes.POINTER(c_void_p), c_void_p]
    lib.drg_p2p_connect.restype = c_int
    lib.drg_p2p_connect.argtypes = [c_void_p, c_void_p]
    lib.drg_p2p_status.restype = c_int
    lib.drg_p2p_status.argtypes = [c_void_p]
    lib.drg_p2p_destroy.restype = c_int
    lib.drg_p2p_destroy.argtypes = [c_void_p]
    lib.drg_dual_softmax.restype = c_int
    lib.drg_dual_softmax.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p,
                                     c_size_t, c_void_p]
    c_ll = ctypes.c_longlong
    lib.drg_gemm_nt_tf32.restype = c_int
    lib.drg_gemm_nt_tf32.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]
    lib.drg_gemm_nt_split16.restype = c_int
    lib.drg_gemm_nt_split16.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]
    lib.drg_project_split16.restype = c_int
    lib.drg_project_split16.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]
    lib.drg_prep_operand.restype = c_int
    lib.drg_prep_operand.argtypes = [c_void_p, c_void_p, c_int, c_ll, c_int, c_float, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.drg_prep_operand_ext.restype = c_int
    lib.drg_prep_operand_ext.argtypes = [c_void_p, c_void_p, c_int, c_ll, c_int, c_float, c_int, c_int, c_int, c_int, c_int, c_void_p,
                                         c_void_p, c_void_p]
    lib.drg_attn_softmax.restype = c_int
    lib.drg_attn_softmax.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p]
    lib.drg_layernorm.restype = c_int
    lib.drg_layernorm.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_ll, c_int, c_float, c_void_p, c_void_p, c_void_p]
    lib.drg_gemm_nt_split16_bias.restype = c_int
    lib.drg_gemm_nt_split16_bias.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p]
    lib.drg_sinkhorn_backward_workspace_bytes.restype = ctypes.c_size_t
    lib.drg_sinkhorn_backward_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int]
    lib.drg_sinkhorn_backward.restype = c_int
    lib.drg_sinkhorn_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p,
                                          c_void_p, c_void_p, c_void_p, ctypes.c_size_t, c_void_p]
    lib.drg_dual_softmax_backward_workspace_bytes.restype = ctypes.c_size_t
    lib.drg_dual_softmax_backward_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.drg_dual_softmax_backward.restype = c_int
    lib.drg_dual_softmax_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                              ctypes.c_size_t, c_void_p]
    lib.drg_ransac_workspace_bytes.restype = c_size_t
    lib.drg_ransac_workspace_bytes.argtypes = [c_int, c_int]
    lib.drg_ransac_correspondence.restype = c_int
    lib.drg_ransac_correspondence.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_void_p, ctypes.c_longlong, c_void_p, c_float, c_int, c_int,
                                              ctypes.c_ulonglong, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p,
                                              c_void_p, c_void_p, c_size_t, c_void_p]
    lib.drg_weighted_procrustes_backward.restype = c_int
    lib.drg_weighted_procrustes_backward.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_float,
                                                     c_void_p, c_void_p]
    lib.drg_prep_vt_split16.restype = c_int
    lib.drg_prep_vt_split16.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.drg_attention_split16.restype = c_int
    lib.drg_attention_workspace_bytes.restype = ctypes.c_size_t
    lib.drg_attention_workspace_bytes.argtypes = [c_int, c_int, c_int, c_int, c_int, c_int]
    lib.drg_attention_split16.argtypes = [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                                          c_void_p, c_int, c_void_p, ctypes.c_size_t, c_void_p]
    lib.drg_fourier_embed.restype = c_int
    lib.drg_fourier_embed.argtypes = [c_void_p, c_void_p, c_ll, c_int, c_int, c_float, c_int, c_int, c_void_p, c_void_p]
    lib.drg_prep_operand_xyz.restype = c_int
    lib.drg_prep_operand_xyz.argtypes = [c_void_p, c_void_p, c_void_p, ctypes.POINTER(c_float), c_float, c_int, c_ll, c_int, c_float,
                                         c_int, c_int, c_void_p, c_void_p, c_void_p]
    lib.drg_position_code.restype = c_int
    lib.drg_position_code.argtypes = [c_void_p, c_void_p, c_ll, c_int, ctypes.POINTER(c_float), c_float, c_int, c_void_p, c_void_p]
    lib.drg_prep_operand_pair.restype = c_int
    lib.drg_prep_operand_pair.argtypes = [c_void_p, c_ll, c_int, c_void_p, c_ll, c_int, c_int, c_float, c_int, c_void_p, c_void_p]
    lib.drg_match_workspace_bytes.restype = c_size_t
    lib.drg_match_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.drg_match_count.restype = c_int
    lib.drg_match_count.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_size_t,
                                    c_void_p, c_void_p]
    lib.drg_match_write.restype = c_int
    lib.drg_match_write.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_size_t,
                                    c_void_p, c_void_p, c_ll, c_void_p, c_void_p]
    lib.drg_topk_match_count.restype = c_int
    lib.drg_topk_match_count.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                         c_void_p, c_size_t, c_void_p, c_void_p]
    lib.drg_topk_match_write.restype = c_int
    lib.drg_topk_match_write.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_void_p, c_void_p,
                                         c_void_p, c_size_t, c_void_p, c_void_p, c_ll, c_void_p, c_void_p]
    lib.drg_match_from_best.restype = c_int
    lib.drg_match_from_best.argtypes = [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_float, c_void_p, c_void_p, c_ll, c_void_p,
                                        c_void_p]
    lib.drg_soft_procrustes_workspace_bytes.restype = c_size_t
    lib.drg_soft_procrustes_workspace_bytes.argtypes = [c_int, c_int, c_int]
    lib.drg_soft_procrustes.restype = c_int
    lib.drg_soft_procrustes.argtypes = [ctypes.POINTER(ProcrustesArgs), c_void_p, c_size_t, c_void_p]
    lib.drg_sinkhorn_soft_procrustes.restype = c_int
    lib.drg_sinkhorn_soft_procrustes.argtypes = [ctypes.POINTER(SinkhornArgs), ctypes.POINTER(ProcrustesArgs), c_void_p, c_size_t,
                                                 c_void_p, c_size_t, c_void_p]
    lib.drg_weighted_procrustes.restype = c_int
    lib.drg_weighted_procrustes.argtypes = [c_void_p, c_void_p, c_void_p, c_int, c_int, c_float, c_void_p, c_void_p, c_void_p,
                                            c_void_p]
    lib.drg_profile_enable.restype = None
    lib.drg_profile_enable.argtypes = [c_int]
    lib.drg_profile_reset.restype = None
    lib.drg_profile_read.restype = c_int
    lib.drg_profile_read.argtypes = [c_int, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_longlong)]
    lib.drg_profile_slots.restype = c_int
    lib.drg_counter_add.restype = c_int
    lib.drg_counter_add.argtypes = [c_void_p, ctypes.c_ulonglong, c_void_p]
    lib.drg_sigmoid.restype = c_int
    lib.drg_sigmoid.argtypes = [c_void_p, c_void_p, c_ll, c_void_p]
    lib.drg_min_value.restype = c_int
    lib.drg_min_value.argtypes = [c_void_p, c_ll, c_void_p, c_void_p, c_void_p]
    return lib


def load_library():
    """Load the CUDA library.  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise DiffRegLibraryError(
                f"{_LIB_PATH} not found: build it with `python diff-reg_b200/build.py` "
                "(or __graft_entry__.build()); diffreg_b200 has no CPU / eager fallback")
        _lib = _declare(ctypes.CDLL(_LIB_PATH))
    return _lib


def check(status):
    if status != 0:
        msg = load_library().drg_last_error().decode()
        raise DiffRegLibraryError(f"diffreg_b200 error {status}: {msg}")


def launch_count():
    return int(load_library().drg_launch_count())


PROFILE_SLOTS = ["skh_iter", "skh_col", "skh_final", "skh_prep", "gemm", "prep_operand", "rowcol_best", "match_rows",
                 "topk_collect", "procr_solve", "topk_threshold", "procr_select", "skh_fused"]


def profile_enable(on=True):
    lib = load_library()
    lib.drg_profile_reset()
    lib.drg_profile_enable(1 if on else 0)


def profile_read():
    """{slot name: (total_ms, launches)} for the launches recorded since profile_enable()."""
    lib = load_library()
    out = {}
    for k, name in enumerate(PROFILE_SLOTS):
        ms, n = ctypes.c_double(0.0), ctypes.c_longlong(0)
        check(lib.drg_profile_read(k, ctypes.byref(ms), ctypes.byref(n)))
        out[name] = (ms.value, n.value)
    return out
