"""Generate golden input/output vectors from the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

It imports the reference's own modules from where they lie, runs them on small seeded
inputs on CPU and stores inputs + reference outputs as ``tests/golden/*.npz``.  The
reference cannot travel to the GPU box, these files can.  Shims:
  * ``mutual_topk_select`` hard-codes ``.cuda()``; ``torch.Tensor.cuda`` is patched to the
    identity while it runs so the reference code executes unchanged on CPU;
  * the 2D-3D ``matching.py`` imports ``vision3d.ops`` whose compiled extension is missing;
    a stub ``vision3d.ops`` exposing the reference's own pure-torch
    ``mutual_topk_select.py`` (loaded by file path) is registered first;
  * sampler loop bodies live inside model ``forward``s that need datasets/backbones, so the
    loop is driven here with the reference's own methods (``get_warped_from_noising_matching``,
    ``predict_noise_from_start``, ``Matching``, ``SoftProcrustesLayer``) bound to a bare
    namespace, the denoising transformer replaced by fixed features.
"""
import importlib
import importlib.util
import math
import os
import sys
import types
from types import SimpleNamespace

import numpy as np
import torch

REF = "/root/reference"
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import diffreg_oracle as O  # only for make_problem (input generation)

torch.set_num_threads(1)
torch.use_deterministic_algorithms(True)


def _purge(prefixes):
    for k in list(sys.modules):
        if any(k == p or k.startswith(p + ".") for p in prefixes):
            del sys.modules[k]


def load_flavour(name):
    """Import the reference's matching/procrustes/pipeline modules for one flavour."""
    _purge(["models", "matching", "procrustes", "position_encoding", "lib", "datasets", "vision3d"])
    for p in list(sys.path):
        if p.startswith(REF):
            sys.path.remove(p)
    ns = SimpleNamespace()
    if name == "4d":
        sys.path.insert(0, f"{REF}/Diff-Reg-4dmatch")
        ns.matching = importlib.import_module("models.matching")
        ns.procrustes = importlib.import_module("models.procrustes")
        ns.pipeline = importlib.import_module("models.pipeline")
        ns.pe = importlib.import_module("models.position_encoding")
    elif name == "3d":
        sys.path.insert(0, f"{REF}/Diff-Reg-3dmatch")
        ns.matching = importlib.import_module("models.matching")
        ns.procrustes = importlib.import_module("models.procrustes")
    elif name == "2d3d":
        exp = f"{REF}/Diff-Reg-2d3d/experiments/2d3dmatr.rgbdv2.stage4.level3.stage1"
        spec = importlib.util.spec_from_file_location(
            "_ref_mts", f"{REF}/Diff-Reg-2d3d/vision3d/ops/mutual_topk_select.py")
        mts = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mts)
        v3d = types.ModuleType("vision3d")
        ops = types.ModuleType("vision3d.ops")
        ops.mutual_topk_select = mts.mutual_topk_select
        v3d.ops = ops
        sys.modules["vision3d"] = v3d
        sys.modules["vision3d.ops"] = ops
        sys.path.insert(0, exp)
        ns.matching = importlib.import_module("matching")
        ns.procrustes = importlib.import_module("procrustes")
        ns.mts = mts
    return ns


class cpu_cuda:
    """Make ``tensor.cuda()`` a no-op so reference code with literal .cuda() runs on CPU."""
    def __enter__(self):
        self._orig = torch.Tensor.cuda
        torch.Tensor.cuda = lambda self, *a, **k: self
    def __exit__(self, *a):
        torch.Tensor.cuda = self._orig


def npz(name, **arrs):
    out = {}
    for k, v in arrs.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
    print("wrote", name, {k: (a.shape, str(a.dtype)) for k, a in out.items()})


def cfg_match(C, match_type="sinkhorn", entangled=True, iters=3):
    return dict(match_type=match_type, confidence_threshold=0.2, feature_dim=C, entangled=entangled,
                dsmax_temperature=0.1, skh_init_bin_score=1.0, skh_iters=iters, skh_prefilter=False)


def rand_scores(g, B, N, M, src_mask, tgt_mask, scale=1.0, dtype=torch.float32):
    s = (torch.randn(B, N, M, generator=g) * scale).to(dtype)
    s.masked_fill_(~(src_mask[..., None] & tgt_mask[:, None]), float("-inf"))
    return s


@torch.no_grad()
def main():
    # ---------------------------------------------------------------- Sinkhorn (a4)
    r4 = load_flavour("4d")
    lot = r4.matching.log_optimal_transport
    cases = [  # name, B, N, M, valid counts, iters, scale, dtype, alpha
        ("lot_small_full", 1, 7, 5, None, 3, 1.0, torch.float32, 1.0),
        ("lot_prefix_b3", 3, 33, 41, [(33, 20), (17, 41), (5, 6)], 3, 2.0, torch.float32, 1.0),
        ("lot_wide_scores", 2, 40, 24, [(40, 24), (31, 9)], 3, 25.0, torch.float32, -0.5),
        ("lot_iters100", 1, 48, 64, None, 100, 1.0, torch.float32, 1.0),
        ("lot_fp64_state", 1, 21, 30, [(19, 30)], 3, 3.0, torch.float64, 1.0),
        ("lot_one_by_one", 1, 1, 1, None, 3, 1.0, torch.float32, 1.0),
        ("lot_rect_129x257", 1, 129, 257, [(120, 250)], 3, 1.0, torch.float32, 2.0),
    ]
    for name, B, N, M, valid, iters, scale, dt, alpha in cases:
        g = torch.Generator().manual_seed(hash(name) % (2 ** 31) if False else sum(map(ord, name)))
        sm = torch.ones(B, N, dtype=torch.bool)
        tm = torch.ones(B, M, dtype=torch.bool)
        if valid is not None:
            for b, (a, c) in enumerate(valid):
                sm[b, a:] = False
                tm[b, c:] = False
        s = rand_scores(g, B, N, M, sm, tm, scale, dt)
        al = torch.tensor(alpha)
        out = lot(s.clone(), al, iters, sm, tm)
        npz(name, scores=s, alpha=al, iters=iters, src_mask=sm, tgt_mask=tm, out=out)

    # arbitrary (non-prefix) masks as in the 2D-3D flavour (Q3)
    g = torch.Generator().manual_seed(77)
    sm = torch.rand(1, 50, generator=g) > 0.2
    tm = torch.rand(1, 37, generator=g) > 0.2
    s = rand_scores(g, 1, 50, 37, sm, tm)
    npz("lot_arbitrary_masks", scores=s, alpha=torch.tensor(1.0), iters=3, src_mask=sm, tgt_mask=tm,
        out=lot(s.clone(), torch.tensor(1.0), 3, sm, tm))

    # ---------------------------------------------------------------- Matching.forward, 3D flavour (a1,a2,a3)
    def run_matching(mod, tag, match_type, B, N, M, C, valid=None, entangled=True, pe=False):
        prob = O.make_problem(sum(map(ord, tag)), B, N, M, C, prefix_valid=valid)
        m = mod.Matching(cfg_match(C, match_type, entangled))
        m.src_proj.weight.copy_(prob["W"])
        m.eval()
        data = {}
        src_pe = tgt_pe = None
        if pe:
            vol = r4.pe.VolumetricPositionEncoding(SimpleNamespace(
                feature_dim=C, vol_bnds=[[-3.6, -2.4, 1.14], [1.093, 0.78, 2.92]], voxel_size=0.04, pe_type="rotary"))
            src_pe = vol(prob["s_pcd"])
            tgt_pe = vol(prob["t_pcd"])
        conf, match = m(prob["src_feats"], prob["tgt_feats"], src_pe, tgt_pe, prob["src_mask"], prob["tgt_mask"], data)
        extra = {}
        if pe:
            extra = dict(src_pe=src_pe, tgt_pe=tgt_pe)
        npz(tag, src_feats=prob["src_feats"], tgt_feats=prob["tgt_feats"], W=prob["W"],
            src_mask=prob["src_mask"], tgt_mask=prob["tgt_mask"], conf=conf, match=match,
            data_src_feats=data["src_feats"], data_tgt_feats=data["tgt_feats"],
            data_src_feats_nopos=data["src_feats_nopos"], data_tgt_feats_nopos=data["tgt_feats_nopos"],
            bin_score=torch.tensor(1.0), **extra)

    run_matching(r4.matching, "match4d_sinkhorn", "sinkhorn", 1, 40, 36, 32)
    run_matching(r4.matching, "match4d_sinkhorn_prefix_b2", "sinkhorn", 2, 24, 30, 64, valid=[(24, 19), (11, 30)])
    run_matching(r4.matching, "match4d_dualsoftmax_b3", "dual_softmax", 3, 28, 22, 32, valid=[(28, 22), (20, 15), (9, 21)])
    run_matching(r4.matching, "match4d_sinkhorn_rotary", "sinkhorn", 1, 30, 26, 36, entangled=False, pe=True)

    # get_match at several thresholds on a matrix with exact ties
    g = torch.Generator().manual_seed(5)
    conf = torch.rand(2, 12, 10, generator=g)
    conf[0, 3, :] = conf[0, 3, 4]          # a constant row: ties
    conf[1, :, 2] = 0.9                    # a constant column
    for thr in (0.0, 0.2, 0.55):
        for mutual in (True, False):
            idx, mconf, mask = r4.matching.Matching.get_match(conf, thr, mutual)
            npz(f"getmatch_thr{thr}_mutual{int(mutual)}", conf=conf, thr=thr, mutual=mutual, index=idx, mconf=mconf, mask=mask)

    # ---------------------------------------------------------------- SoftProcrustes (a7, a8)
    def run_procrustes(mod, tag, B, N, M, valid, sample_rate, max_cond, degenerate=False):
        prob = O.make_problem(sum(map(ord, tag)), B, N, M, 8, prefix_valid=valid)
        if degenerate:
            prob["s_pcd"][..., 2] = 0.0    # planar source: ill-conditioned covariance
            prob["t_pcd"][..., 2] = 0.0
        g = torch.Generator().manual_seed(11)
        conf = torch.rand(B, N, M, generator=g) ** 8    # few large entries
        conf = conf * (prob["src_mask"][..., None] & prob["tgt_mask"][:, None])
        layer = mod.SoftProcrustesLayer(SimpleNamespace(sample_rate=sample_rate, max_condition_num=max_cond))
        R, t, Rf, tf, cond, ok = layer(conf, prob["s_pcd"], prob["t_pcd"], prob["src_mask"], prob["tgt_mask"])
        npz(tag, conf=conf, s_pcd=prob["s_pcd"], t_pcd=prob["t_pcd"], src_mask=prob["src_mask"], tgt_mask=prob["tgt_mask"],
            sample_rate=sample_rate, max_condition_num=max_cond, R=R, t=t, R_forwd=Rf, t_forwd=tf, condition=cond, solution_mask=ok)

    run_procrustes(r4.procrustes, "procrustes4d_b1", 1, 40, 32, None, 1.0, 40.0)
    run_procrustes(r4.procrustes, "procrustes4d_b3_prefix", 3, 30, 36, [(30, 36), (14, 20), (25, 9)], 1.0, 40.0)
    run_procrustes(r4.procrustes, "procrustes4d_rate05", 2, 26, 26, [(26, 26), (20, 22)], 0.5, 40.0)
    run_procrustes(r4.procrustes, "procrustes4d_degenerate", 1, 24, 24, None, 1.0, 40.0, degenerate=True)

    g = torch.Generator().manual_seed(3)
    X = torch.randn(4, 50, 3, generator=g)
    Rg = torch.stack([O.random_rotation(g) for _ in range(4)])
    Y = X @ Rg.transpose(1, 2) + torch.randn(4, 1, 3, generator=g) + 0.05 * torch.randn(4, 50, 3, generator=g)
    Y[3] = -Y[3]                                            # forces the reflection fix
    w = torch.rand(4, 50, 1, generator=g)
    R, t, cond = r4.procrustes.SoftProcrustesLayer.batch_weighted_procrustes(X, Y, w)
    npz("kabsch_b4", X=X, Y=Y, w=w, R=R, t=t, condition=cond)

    # ---------------------------------------------------------------- schedule + DDIM update (a10) and one 4d step (a9)
    P = r4.pipeline
    betas = P.cosine_beta_schedule(1000)
    ac = torch.cumprod(1.0 - betas, dim=0)
    fake = SimpleNamespace(alphas_cumprod=ac, sqrt_recip_alphas_cumprod=torch.sqrt(1.0 / ac),
                           sqrt_recipm1_alphas_cumprod=torch.sqrt(1.0 / ac - 1))
    npz("schedule", alphas_cumprod=ac)

    def ref_sampler(flavour, tag, N, M, C, steps, arbitrary=0.0, max_cond=40.0, pe=False):
        """The reference's loop body driven with fixed features (see module docstring).  pe=True: the head runs with
        entangled=False and receives, every step, the rotary position codes of the WARPED source points and of the target
        points -- what the denoising transformer hands to denoising_coarse_matching (pipeline.py:177-178)."""
        prob = O.make_problem(sum(map(ord, tag)), 1, N, M, C, arbitrary_invalid=arbitrary)
        vol = None
        if pe:
            vol = r4.pe.VolumetricPositionEncoding(SimpleNamespace(
                feature_dim=C, vol_bnds=[[-3.6, -2.4, 1.14], [1.093, 0.78, 2.92]], voxel_size=0.04, pe_type="rotary"))
        g = torch.Generator().manual_seed(99)
        if flavour == "2d3d":
            mods = load_flavour("2d3d")
            head = mods.matching.Matching(cfg_match(C), mutual=True)
        elif flavour == "3d":
            mods = load_flavour("3d")
            head = mods.matching.Matching(cfg_match(C))
        else:
            mods = r4
            head = mods.matching.Matching(cfg_match(C, entangled=not pe))
        head.src_proj.weight.copy_(prob["W"])
        head.eval()
        proc = mods.procrustes.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=max_cond))
        fake.denoising_coarse_matching = head
        fake.denoising_soft_procrustes = proc
        sm, tm = prob["src_mask"], prob["tgt_mask"]
        x_T = torch.randn(1, N, M, generator=g)
        noises = [torch.randn(1, N, M, generator=g) for _ in range(steps)]
        x = x_T.clone()
        times = torch.linspace(0, 999, steps=steps + 1)
        times = list(reversed(times.int().tolist()))
        rec = {}
        for k, (time, time_next) in enumerate(zip(times[:-1], times[1:])):
            if flavour == "3d":
                x = x - x.min()                                           # 3d/models/pipeline.py:239
            rec[f"x_in_{k}"] = x.clone()
            warped, _ = P.Pipeline.get_warped_from_noising_matching(fake, prob["s_pcd"], prob["t_pcd"], sm, tm, x)
            rec[f"warped_{k}"] = warped
            if flavour == "2d3d":
                with cpu_cuda():
                    x_start, _, _, _ = head(prob["src_feats"], prob["tgt_feats"], sm, tm, True)
            elif pe:
                x_start, _ = head(prob["src_feats"], prob["tgt_feats"], vol(warped), vol(prob["t_pcd"]), sm, tm, {}, pe_type="rotary")
            else:
                x_start, _ = head(prob["src_feats"], prob["tgt_feats"], None, None, sm, tm, {})
            rec[f"x0_{k}"] = x_start
            tc = torch.full((1,), time, dtype=torch.long)
            pred = P.Pipeline.predict_noise_from_start(fake, x, tc, x_start)
            a, an = ac[time], ac[time_next]
            sigma = 1.0 * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
            c = (1 - an - sigma ** 2).sqrt()
            if flavour == "4d":
                x = x_start * an.sqrt() + c * pred + sigma * noises[k]    # 4d/models/pipeline.py:190
            else:
                x = x_start * an.sqrt() + c * pred                        # 3d :256, 2d3d :678
            rec[f"x_out_{k}"] = x
        if flavour == "4d":
            rec["conf_matrix_pred"] = torch.sigmoid(x)
        else:
            sim = x - x.min() if flavour == "3d" else x
            sim.masked_fill_(~(sm[..., None] * tm[:, None]).bool(), float("-inf"))
            la = mods.matching.log_optimal_transport(sim, head.bin_score, head.skh_iters, sm, tm)
            conf = la.exp()[:, :-1, :-1].contiguous()
            with cpu_cuda():
                sel = mods.mts.mutual_topk_select if flavour == "2d3d" else mods.matching.mutual_topk_select
                r, c_, w = sel(conf.squeeze(0), 1, largest=True, threshold=None, mutual=False)
            rec["conf_matrix_pred"] = conf
            rec["match_pred"] = torch.stack((torch.zeros_like(r), r, c_), dim=-1)
            rec["match_weights"] = w
        npz(tag, steps=steps, flavour=flavour, max_condition_num=max_cond, x_T=x_T, noises=torch.stack(noises),
            src_feats=prob["src_feats"], tgt_feats=prob["tgt_feats"], W=prob["W"], s_pcd=prob["s_pcd"], t_pcd=prob["t_pcd"],
            src_mask=sm, tgt_mask=tm, **rec)

    ref_sampler("4d", "sampler4d_3steps", 28, 24, 32, 3)
    ref_sampler("3d", "sampler3d_3steps", 20, 26, 32, 3, max_cond=0.0)
    ref_sampler("2d3d", "sampler2d3d_3steps", 30, 22, 32, 3, arbitrary=0.1, max_cond=200.0)
    # every shipped 3DMatch / 4DMatch config has entangled: False with rotary codes (configs/test/4dmatch.yaml:1,46)
    ref_sampler("4d", "sampler4d_rotary_3steps", 26, 30, 36, 3, max_cond=0.0, pe=True)
    ref_sampler("4d", "sampler4d_rotary_gate_3steps", 26, 30, 36, 3, max_cond=1e9, pe=True)

    # ---------------------------------------------------------------- 2D-3D head (a1') and mutual_topk_select (a6)
    r2 = load_flavour("2d3d")
    for mutual in (True, False):
        tag = f"match2d3d_mutual{int(mutual)}"
        prob = O.make_problem(sum(map(ord, tag)), 1, 45, 33, 32, arbitrary_invalid=0.1)
        m = r2.matching.Matching(cfg_match(32), mutual=True)
        m.src_proj.weight.copy_(prob["W"])
        m.eval()
        with cpu_cuda():
            conf, si, ti, wts = m(prob["src_feats"], prob["tgt_feats"], prob["src_mask"], prob["tgt_mask"], mutual)
        npz(tag, src_feats=prob["src_feats"], tgt_feats=prob["tgt_feats"], W=prob["W"], src_mask=prob["src_mask"],
            tgt_mask=prob["tgt_mask"], mutual=mutual, conf=conf, src_indices=si, tgt_indices=ti, weights=wts)

    g = torch.Generator().manual_seed(8)
    score = torch.rand(19, 23, generator=g)
    for mutual in (True, False):
        for thr in (None, 0.9):
            with cpu_cuda():
                r, c, w = r2.mts.mutual_topk_select(score, 1, largest=True, threshold=thr, mutual=mutual)
            npz(f"mts_mutual{int(mutual)}_thr{thr}", score=score, mutual=mutual, threshold=(-1.0 if thr is None else thr),
                has_threshold=thr is not None, rows=r, cols=c, scores=w)

    # mutual_topk_select / batch_mutual_topk_select for k > 1 (the 2D-3D fine matching: model.py:738-746 uses k = 2, thr 0.75)
    spec = importlib.util.spec_from_file_location("_ref_mts2", f"{REF}/Diff-Reg-2d3d/vision3d/ops/mutual_topk_select.py")
    mts2 = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mts2)
    g = torch.Generator().manual_seed(18)
    score = torch.rand(37, 29, generator=g)
    for k in (2, 3):
        for mutual in (True, False):
            with cpu_cuda():
                r, c, w = mts2.mutual_topk_select(score, k, largest=True, threshold=0.5, mutual=mutual)
                cm = mts2.mutual_topk_select(score, k, largest=False, threshold=None, mutual=mutual, reduce_result=False)
            npz(f"mtsk_k{k}_mutual{int(mutual)}", score=score, k=k, mutual=mutual, threshold=0.5, rows=r, cols=c, scores=w, corr_smallest=cm)
    bscore = torch.rand(5, 20, 24, generator=g) * 2 - 1
    rmask = torch.rand(5, 20, generator=g) > 0.15
    cmask = torch.rand(5, 24, generator=g) > 0.15
    for mutual in (True, False):
        with cpu_cuda():
            bi, ri, ci, w = mts2.batch_mutual_topk_select(bscore, k=2, row_masks=rmask, col_masks=cmask, threshold=0.3, largest=True,
                                                          mutual=mutual)
        npz(f"bmts_k2_mutual{int(mutual)}", score=bscore, row_masks=rmask, col_masks=cmask, k=2, mutual=mutual, threshold=0.3,
            batch=bi, rows=ri, cols=ci, scores=w)

    # ---------------------------------------------------------------- forward1 of the 3d flavour (a1'')
    r3 = load_flavour("3d")
    prob = O.make_problem(123, 1, 31, 27, 32)
    m = r3.matching.Matching(cfg_match(32))
    m.src_proj.weight.copy_(prob["W"])
    m.eval()
    with cpu_cuda():
        conf, match = m.forward1(prob["src_feats"], prob["tgt_feats"], None, None, prob["src_mask"], prob["tgt_mask"], {}, mutual=False)
    npz("match3d_forward1", src_feats=prob["src_feats"], tgt_feats=prob["tgt_feats"], W=prob["W"],
        src_mask=prob["src_mask"], tgt_mask=prob["tgt_mask"], conf=conf, match=match)

    # 3d SoftProcrustes variant (padded lengths)
    prob = O.make_problem(321, 1, 22, 28, 8, prefix_valid=[(18, 28)])
    g = torch.Generator().manual_seed(12)
    conf = torch.rand(1, 22, 28, generator=g) ** 8
    layer = r3.procrustes.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
    R, t, Rf, tf, cond, ok = layer(conf, prob["s_pcd"], prob["t_pcd"], prob["src_mask"], prob["tgt_mask"])
    npz("procrustes3d_padded", conf=conf, s_pcd=prob["s_pcd"], t_pcd=prob["t_pcd"], src_mask=prob["src_mask"],
        tgt_mask=prob["tgt_mask"], sample_rate=1.0, max_condition_num=40.0, R=R, t=t, R_forwd=Rf, t_forwd=tf,
        condition=cond, solution_mask=ok)


if __name__ == "__main__":
    main()
