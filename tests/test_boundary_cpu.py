"""The drop-in boundary checked on the CPU box.

  * interface parity with the UNMODIFIED reference classes (imported from /root/reference, or from the copy under
    oracle/_ref made by oracle/make_ref.py; skipped when neither exists): constructor / forward / static-method
    signatures, and state_dict keys + shapes, for the 4DMatch, 3DMatch and 2D-3D flavours;
  * the autograd rule: the kernels are forward-only, so a call that autograd would record raises instead of silently
    detaching -- the check runs before any CUDA work, hence CPU tensors exercise it here;
  * the shim files under shims/ (what INTEGRATION.md tells a maintainer to drop in) export the reference's names.
"""
import importlib.util
import inspect
import os
import sys
from types import SimpleNamespace

import pytest
import torch

import diffreg_b200
from diffreg_b200 import matching as M, procrustes as P
from diffreg_b200._lib import DiffRegLibraryError
from oracle import ref_loader

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
needs_ref = pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present (/root/reference or oracle/_ref)")


def _cfg(C=32, match_type="sinkhorn"):
    return dict(match_type=match_type, confidence_threshold=0.2, feature_dim=C, entangled=True, dsmax_temperature=0.1,
                skh_init_bin_score=1.0, skh_iters=3, skh_prefilter=False)


def _params(fn):
    return [(n, p.kind, p.default) for n, p in inspect.signature(fn).parameters.items()]


def _assert_prefix(ref_fn, our_fn, what):
    """Every call the reference accepts is accepted with the same meaning: our callable takes the reference's parameters
    -- same names, order and kinds; the same default wherever the reference has one (one drop-in function serves copies
    of the reference that differ only in having defaults: 3d models/matching.py:6 vs vision3d/ops/mutual_topk_select.py:7)
    -- and may only ADD keyword parameters with defaults after them."""
    r, o = _params(ref_fn), _params(our_fn)
    assert len(o) >= len(r), f"{what}: reference {r} vs ours {o}"
    for (rn, rk, rd), (on, ok, od) in zip(r, o):
        assert (rn, rk) == (on, ok), f"{what}: reference {r} vs ours {o}"
        if rd is not inspect.Parameter.empty:
            assert od == rd, f"{what}: default of {rn}: reference {rd!r} vs ours {od!r}"
    for name, _, default in o[len(r):]:
        assert default is not inspect.Parameter.empty, f"{what}: extra parameter {name} has no default"


@needs_ref
@pytest.mark.parametrize("flavour", ["4d", "3d", "2d3d"])
def test_signatures_and_state_dict_match_the_reference(flavour):
    ref = ref_loader.load_flavour(flavour)
    try:
        RM, RP = ref.matching.Matching, ref.procrustes.SoftProcrustesLayer
        OM = M.Matching2D3D if flavour == "2d3d" else M.Matching
        OP = P.SoftProcrustesLayer3DMatch if flavour == "3d" else P.SoftProcrustesLayer
        _assert_prefix(RM.__init__, OM.__init__, "Matching.__init__")
        _assert_prefix(RM.forward, OM.forward, "Matching.forward")
        _assert_prefix(RM.get_match, OM.get_match, "Matching.get_match")
        if hasattr(RM, "get_topk_match"):
            _assert_prefix(RM.get_topk_match, OM.get_topk_match, "Matching.get_topk_match")
        if hasattr(RM, "forward1"):
            _assert_prefix(RM.forward1, OM.forward1, "Matching.forward1")
        _assert_prefix(ref.matching.log_optimal_transport, M.log_optimal_transport, "log_optimal_transport")
        _assert_prefix(RP.__init__, OP.__init__, "SoftProcrustesLayer.__init__")
        _assert_prefix(RP.forward, OP.forward, "SoftProcrustesLayer.forward")
        _assert_prefix(RP.batch_weighted_procrustes, OP.batch_weighted_procrustes, "batch_weighted_procrustes")
        mts = ref.mts.mutual_topk_select if flavour == "2d3d" else getattr(ref.matching, "mutual_topk_select", None)
        if mts is not None:
            _assert_prefix(mts, M.mutual_topk_select, "mutual_topk_select")
        for mt in ("sinkhorn", "dual_softmax"):
            rsd = RM(_cfg(48, mt)).state_dict()
            osd = OM(_cfg(48, mt)).state_dict()
            assert list(rsd.keys()) == list(osd.keys())
            for k in rsd:
                assert rsd[k].shape == osd[k].shape and rsd[k].dtype == osd[k].dtype, k
            # a reference checkpoint loads strictly into the drop-in (and back)
            OM(_cfg(48, mt)).load_state_dict(rsd, strict=True)
            RM(_cfg(48, mt)).load_state_dict(osd, strict=True)
        pc = SimpleNamespace(sample_rate=0.5, max_condition_num=30.0)
        rp, op = RP(pc), OP(pc)
        assert (rp.sample_rate, rp.max_condition_num) == (op.sample_rate, op.max_condition_num)
        assert list(rp.state_dict().keys()) == list(op.state_dict().keys()) == []
        # the attributes the reference's callers read off the module (pipeline.py:213, transformer.py:188-207)
        m = OM(_cfg(48))
        for attr in ("bin_score", "skh_iters", "confidence_threshold", "match_type", "src_proj", "tgt_proj", "entangled"):
            assert hasattr(m, attr), attr
    finally:
        ref_loader.unload()


@needs_ref
def test_position_encoding_interface_matches_the_reference():
    ref = ref_loader.load_flavour("4d")
    try:
        from diffreg_b200.position_encoding import VolumetricPositionEncoding as OV
        RV = ref.pe.VolumetricPositionEncoding
        for name in ("__init__", "forward", "voxelize", "embed_rotary", "embed_pos"):
            _assert_prefix(getattr(RV, name), getattr(OV, name), "VolumetricPositionEncoding." + name)
    finally:
        ref_loader.unload()


@pytest.mark.skipif(not ref_loader.fusion_available(), reason="reference sources of the 2D-3D fusion module not present")
def test_fusion_module_interface_and_state_dict_match_the_reference():
    """2D-3D flavour of 8f rank 2: CrossModalFusionModule / TransformerLayer / FourierEmbedding behind the reference's constructor
    and forward signatures, parameters under the reference's names (a reference checkpoint loads strictly, and back)."""
    ref = ref_loader.load_fusion()
    try:
        from diffreg_b200 import fusion as OF
        pairs = [(ref.fusion.CrossModalFusionModule, OF.CrossModalFusionModule), (ref.transformer.TransformerLayer, OF.TransformerLayer),
                 (ref.transformer.AttentionLayer, OF.AttentionLayer), (ref.transformer.AttentionOutput, OF.AttentionOutput),
                 (ref.transformer.MultiHeadAttention, OF.MultiHeadAttention), (ref.embedding.FourierEmbedding, OF.FourierEmbedding)]
        for rc, oc in pairs:
            _assert_prefix(rc.__init__, oc.__init__, rc.__name__ + ".__init__")
        for rc, oc in pairs[:2] + pairs[-1:]:
            _assert_prefix(rc.forward, oc.forward, rc.__name__ + ".forward")
        for name in ("create_2d_embedding", "create_3d_embedding"):
            _assert_prefix(getattr(pairs[0][0], name), getattr(pairs[0][1], name), name)
        blocks = ["self", "cross", "self", "cross", "self", "cross"]
        for use_emb in (True, False):
            rn = ref.fusion.CrossModalFusionModule(512, 512, 256, 256, 4, blocks, use_embedding=use_emb)   # config.py:135-141
            on = OF.CrossModalFusionModule(512, 512, 256, 256, 4, blocks, use_embedding=use_emb)
            rsd, osd = rn.state_dict(), on.state_dict()
            assert list(rsd.keys()) == list(osd.keys())
            assert all(rsd[k].shape == osd[k].shape and rsd[k].dtype == osd[k].dtype for k in rsd)
            on.load_state_dict(rsd, strict=True)
            rn.load_state_dict(osd, strict=True)
        # what the fusion module never uses stays with the reference module: refused, not computed differently
        with pytest.raises(NotImplementedError):
            OF.TransformerLayer(64, 4, dropout=0.1)
        with pytest.raises(NotImplementedError):
            OF.TransformerLayer(64, 4, act_cfg="GELU")
        with pytest.raises(NotImplementedError):
            OF.TransformerLayer(64, 4, qk_embed_proj=True)
        x = torch.randn(1, 5, 64, requires_grad=True)
        with pytest.raises(RuntimeError, match="forward-only"):
            OF.TransformerLayer(64, 4)(x, x, x)
    finally:
        ref_loader.unload()


@needs_ref
def test_transformer_interface_and_state_dict_match_the_reference():
    """GeometryAttentionLayer / RepositioningTransformer take the reference's config keys and arguments
    and hold the reference's parameters under the reference's names (a reference checkpoint loads strictly, and back)."""
    ref = ref_loader.load_flavour("4d")
    try:
        from diffreg_b200 import transformer as OT
        RT = ref.transformer
        for cls in ("GeometryAttentionLayer", "RepositioningTransformer"):
            for name in ("__init__", "forward"):
                _assert_prefix(getattr(getattr(RT, cls), name), getattr(getattr(OT, cls), name), f"{cls}.{name}")

        class Cfg(dict):
            __getattr__ = dict.__getitem__
        lc = Cfg(feature_dim=48, n_head=4, pe_type="rotary")
        rsd, osd = RT.GeometryAttentionLayer(lc).state_dict(), OT.GeometryAttentionLayer(lc).state_dict()
        assert list(rsd.keys()) == list(osd.keys())
        assert all(rsd[k].shape == osd[k].shape and rsd[k].dtype == osd[k].dtype for k in rsd)
        tc = Cfg(feature_dim=48, n_head=4, layer_types=["self", "cross", "positioning", "self", "cross"], positioning_type="procrustes",
                 pe_type="rotary", entangled=False, vol_bnds=[[-3.6, -2.4, 1.14], [1.093, 0.78, 2.92]], voxel_size=0.04,
                 feature_matching=_cfg(48), procrustes=Cfg(max_condition_num=40, sample_rate=1.0))
        rn, on = RT.RepositioningTransformer(tc), OT.RepositioningTransformer(tc)
        rsd, osd = rn.state_dict(), on.state_dict()
        assert list(rsd.keys()) == list(osd.keys())
        assert all(rsd[k].shape == osd[k].shape and rsd[k].dtype == osd[k].dtype for k in rsd)
        on.load_state_dict(rsd, strict=True)
        rn.load_state_dict(osd, strict=True)
        for attr in ("d_model", "nhead", "layer_types", "positioning_type", "pe_type", "entangled", "positional_encoding", "layers"):
            assert hasattr(on, attr), attr
    finally:
        ref_loader.unload()


# ---------------------------------------------------------------------------------------------------------------
def test_autograd_tracked_calls_raise_instead_of_detaching():
    """Every public entry point of the modules; CPU tensors are enough (the check precedes any device work)."""
    C = 16
    head = M.Matching(_cfg(C)).eval()
    head2 = M.Matching2D3D(_cfg(C)).eval()
    proc = P.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
    src = torch.randn(1, 6, C, requires_grad=True)
    tgt = torch.randn(1, 5, C)
    sm, tm = torch.ones(1, 6, dtype=torch.bool), torch.ones(1, 5, dtype=torch.bool)
    conf = torch.rand(1, 6, 5, requires_grad=True)
    pts_s, pts_t = torch.randn(1, 6, 3), torch.randn(1, 5, 3)
    msg = "forward-only"
    with pytest.raises(DiffRegLibraryError, match=msg):
        head(src, tgt, None, None, sm, tm, {})
    with pytest.raises(DiffRegLibraryError, match=msg):
        head.forward1(src, tgt, None, None, sm, tm, {})
    with pytest.raises(DiffRegLibraryError, match=msg):
        head2(src, tgt, sm, tm)
    with pytest.raises(DiffRegLibraryError, match=msg):
        proc(conf, pts_s, pts_t, sm, tm)
    with pytest.raises(DiffRegLibraryError, match=msg):
        proc(conf.detach(), pts_s.requires_grad_(), pts_t, sm, tm)
    with pytest.raises(DiffRegLibraryError, match=msg):
        P.SoftProcrustesLayer.batch_weighted_procrustes(pts_s.detach(), pts_s.detach(), torch.rand(1, 6, 1, requires_grad=True))
    # log_optimal_transport is differentiable (CUDA forward + backward kernels): with CPU tensors the call
    # stops at the CUDA-only check, tracked or not -- never at a detached result
    with pytest.raises(DiffRegLibraryError, match="CUDA"):
        M.log_optimal_transport(conf, torch.tensor(1.0), 3, sm, tm)
    with pytest.raises(DiffRegLibraryError, match="CUDA"):
        M.log_optimal_transport(conf.detach(), torch.tensor(1.0, requires_grad=True), 3, sm, tm)
    with pytest.raises(DiffRegLibraryError, match=msg):
        M.Matching.get_match(conf, 0.2)
    with pytest.raises(DiffRegLibraryError, match=msg):
        M.mutual_topk_select(conf[0], 1)
    # a module left in training mode with trainable weights would return a differentiable result in the reference
    train_head = M.Matching(_cfg(C))
    assert train_head.training
    with pytest.raises(DiffRegLibraryError, match=msg):
        train_head(src.detach(), tgt, None, None, sm, tm, {})


def test_no_grad_calls_pass_the_guard_and_then_need_cuda():
    """Under torch.no_grad() (how the reference's testers call, lib/tester.py:245) the guard lets the call through; on
    this CPU box it then stops at the CUDA-only check -- never at a silent fallback."""
    C = 16
    head = M.Matching(_cfg(C)).eval()
    src = torch.randn(1, 6, C, requires_grad=True)
    tgt = torch.randn(1, 5, C)
    sm, tm = torch.ones(1, 6, dtype=torch.bool), torch.ones(1, 5, dtype=torch.bool)
    with torch.no_grad():
        with pytest.raises(DiffRegLibraryError, match="CUDA tensors only"):
            head(src, tgt, None, None, sm, tm, {})
    # eval mode + inputs that do not require grad: allowed outside no_grad too (parameters alone do not trip the guard)
    with pytest.raises(DiffRegLibraryError, match="CUDA tensors only"):
        head(src.detach(), tgt, None, None, sm, tm, {})


# ---------------------------------------------------------------------------------------------------------------
SHIMS = {
    "Diff-Reg-4dmatch/models/matching.py": ["Matching", "log_optimal_transport"],
    "Diff-Reg-4dmatch/models/procrustes.py": ["SoftProcrustesLayer"],
    "Diff-Reg-4dmatch/models/position_encoding.py": ["VolumetricPositionEncoding"],
    "Diff-Reg-4dmatch/models/transformer.py": ["GeometryAttentionLayer", "RepositioningTransformer"],
    "Diff-Reg-3dmatch/models/matching.py": ["Matching", "log_optimal_transport", "mutual_topk_select"],
    "Diff-Reg-3dmatch/models/procrustes.py": ["SoftProcrustesLayer"],
    "Diff-Reg-2d3d/experiments/matching.py": ["Matching", "log_optimal_transport"],
    "Diff-Reg-2d3d/experiments/procrustes.py": ["SoftProcrustesLayer"],
    "Diff-Reg-2d3d/experiments/fusion_module.py": ["CrossModalFusionModule"],
}


@pytest.mark.parametrize("rel", sorted(SHIMS))
def test_shim_files_export_the_reference_names(rel):
    path = os.path.join(ROOT, "shims", rel)
    spec = importlib.util.spec_from_file_location("_shim_" + rel.replace("/", "_").replace("-", "_").replace(".", "_"), path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for name in SHIMS[rel]:
        assert hasattr(mod, name), f"{rel} does not export {name}"
    if rel.endswith("3dmatch/models/procrustes.py") and "3dmatch" in rel and "4dmatch" not in rel:
        assert mod.SoftProcrustesLayer.padded_lengths is True          # 3d/models/procrustes.py:61-62
    if "2d3d" in rel and rel.endswith("matching.py"):
        assert mod.Matching is M.Matching2D3D
