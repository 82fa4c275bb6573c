"""The reverse-diffusion loop body of Diff-Reg's samplers, chained sync-free on one CUDA stream.

Restates the loop bodies
    4d   Diff-Reg-4dmatch/models/pipeline.py:156-223   (sigma*noise term, final sigmoid)
    3d   Diff-Reg-3dmatch/models/pipeline.py:221-309   (x -= x.min() first, no noise, final Sinkhorn + top-1 union)
    2d3d Diff-Reg-2d3d/experiments/<exp>/model.py:637-694, 830-846
with the denoising transformer (out of scope) replaced by a caller-supplied
callable that maps the warped source points to the (src_feats, tgt_feats) of this step; the default
keeps the features fixed.  Per step:

    [mask + Sinkhorn(x_t) + exp]            drg_sinkhorn (DRG_OUT_CONF, fused mask / shift)
    [top-K + Kabsch + gate + warp]          drg_soft_procrustes
    [projection, 1/sqrt(C), similarity]     drg_prep_operand + drg_gemm_nt_tf32
    [mask + Sinkhorn(sim) + exp + DDIM]     drg_sinkhorn (DRG_OUT_DDIM: x_next, x0 and min(x_next) in the final pass)
    [mutual-NN matches of x0 at thr]        drg_match_count / drg_match_write (device-side count)

Nothing in a step reads back to the host, so a step (or a whole sample) can be captured in a CUDA graph.
The sampler state is fp32 (the reference drifts to fp64 after the first step).
"""
import math

import torch

from . import ops


def cosine_alphas_cumprod(timesteps=1000, s=0.008):
    """pipeline.py:24-34, 97-99 (fp64)."""
    x = torch.linspace(0, timesteps, timesteps + 1, dtype=torch.float64)
    f = torch.cos(((x / timesteps) + s) / (1 + s) * math.pi * 0.5) ** 2
    f = f / f[0]
    betas = torch.clip(1 - (f[1:] / f[:-1]), 0, 0.999)
    return torch.cumprod(1.0 - betas, dim=0)


def time_pairs(sampling_steps, timesteps=1000):
    """pipeline.py:166-169."""
    ts = torch.linspace(0, timesteps - 1, steps=sampling_steps + 1)
    ts = list(reversed(ts.int().tolist()))
    return list(zip(ts[:-1], ts[1:]))


def ddim_coefficients(ac, t, t_next, eta=1.0):
    """Closed form of predict_noise_from_start + the x update (pipeline.py:180-190, 201-205):
        x_next = k_x0 * x0 + k_xt * x_t + sigma * noise
    with k_x0 = sqrt(a_next) - c / sqrt(1/a - 1),  k_xt = c * sqrt(1/a) / sqrt(1/a - 1)."""
    a, an = float(ac[t]), float(ac[t_next])
    sigma = eta * math.sqrt((1 - a / an) * (1 - an) / (1 - a))
    c = math.sqrt(1 - an - sigma ** 2)
    r, rm1 = math.sqrt(1.0 / a), math.sqrt(1.0 / a - 1)
    return math.sqrt(an) - c / rm1, c * r / rm1, sigma


class DenoisingSampler:
    def __init__(self, flavour, matching, procrustes, steps, denoising_matching=None, eta=1.0, timesteps=1000,
                 extract_matches=True, noise_seed=0):
        """matching: the head producing x0 each step (Matching / Matching2D3D); denoising_matching: the module whose
        bin_score / skh_iters drive the Sinkhorn on the noisy state (defaults to `matching`, as in the reference where
        both are `denoising_coarse_matching`)."""
        assert flavour in ("4d", "3d", "2d3d")
        self.flavour = flavour
        self.matching = matching
        self.state_matching = denoising_matching if denoising_matching is not None else matching
        self.procrustes = procrustes
        self.steps = steps
        self.eta = eta
        self.ac = cosine_alphas_cumprod(timesteps)
        self.pairs = time_pairs(steps, timesteps)
        self.noise_calls = 0
        self.extract_matches = extract_matches
        self.noise_seed = noise_seed      # key of the in-kernel Philox stream used when no noise tensor is passed

    @torch.no_grad()
    def step(self, k, x, shift, src_feats, tgt_feats, s_pcd, t_pcd, src_mask, tgt_mask, noise=None,
             pose_tgt_pcd=None, pose_tgt_mask=None, feature_fn=None, x_out=None, noise_counter=None, want_x0=False,
             src_pe=None, tgt_pe=None, pe_type="rotary", x_min_out=None):
        """One reverse step.  x: [1,N,M] state (as stored: for the 3d flavour the true state is x - shift).
        x_out: optional preallocated [1,N,M] buffer for x_next; noise_counter: optional 1-element int64 device
        counter used as the Philox offset (and incremented) instead of the host-side call count.
        src_pe / tgt_pe / pe_type: position codes handed to the matching head exactly as the reference loop hands the
        denoising transformer's to `denoising_coarse_matching` (pipeline.py:177-178); with `entangled: False` (every
        shipped 3DMatch / 4DMatch config) the head applies them after the projection.  `feature_fn` may return
        (src_feats, tgt_feats) or (src_feats, tgt_feats, src_pe, tgt_pe).
        x_min_out: optional 1-element fp32 buffer receiving min(x_next) (3d flavour; a fresh tensor otherwise).
        Returns (x_next, shift_next, aux)."""
        t, t_next = self.pairs[k]
        k_x0, k_xt, sigma = ddim_coefficients(self.ac, t, t_next, self.eta)
        sm, pm = self.state_matching, self.procrustes
        p_pcd = t_pcd if pose_tgt_pcd is None else pose_tgt_pcd
        p_mask = tgt_mask if pose_tgt_mask is None else pose_tgt_mask
        # noisy matching -> pose -> warped source points        get_warped_from_noising_matching
        if want_x0:      # tracing: also return the intermediate confidence matrix
            conf_d = ops.sinkhorn(x, sm.bin_score, sm.skh_iters, src_mask, p_mask, out_mode="conf", apply_mask=True, shift=shift)
            pose = ops.soft_procrustes(conf_d, s_pcd, p_pcd, src_mask, p_mask, pm.sample_rate, pm.max_condition_num,
                                       padded_lengths=pm.padded_lengths, want_warped=True)
        else:            # one call, no confidence matrix in HBM
            conf_d = None
            pose = ops.sinkhorn_soft_procrustes(x, sm.bin_score, sm.skh_iters, src_mask, p_mask, s_pcd, p_pcd, pm.sample_rate,
                                                pm.max_condition_num, padded_lengths=pm.padded_lengths, apply_mask=True,
                                                shift=shift, want_warped=True)
        if feature_fn is not None:
            feats = feature_fn(pose["src_warped"], t_pcd, src_feats, tgt_feats)
            if len(feats) == 4:
                src_feats, tgt_feats, src_pe, tgt_pe = feats
            else:
                src_feats, tgt_feats = feats
        # x0 from the matching head, fused with the DDIM update
        m = self.matching
        if self.flavour == "2d3d":        # the 2D-3D head takes no position codes (2d3d matching.py:91)
            sim = m.similarity(src_feats, tgt_feats)
        else:
            sim = m.similarity(src_feats, tgt_feats, src_pe, tgt_pe, pe_type)
        gen = self.flavour == "4d" and noise is None      # throughput mode: draw the noise inside the final pass
        use_noise = self.flavour == "4d"
        x_min = None
        if self.flavour == "3d":
            if x_min_out is not None:
                x_min = x_min_out.fill_(float("inf"))
            else:
                x_min = torch.full((1,), float("inf"), dtype=torch.float32, device=x.device)
        B, N, M = sim.shape
        # the row / column bests come out of the tiled final pass, which needs 16-byte aligned rows and buffers
        fused_match = (self.extract_matches and M % 4 == 0 and all(t is None or t.data_ptr() % 16 == 0 for t in (x, x_out, noise))
                       and tgt_mask.data_ptr() % 4 == 0)
        res = ops.sinkhorn(sim, m.bin_score, m.skh_iters, src_mask, tgt_mask, out_mode="ddim", apply_mask=True,
                           x_t=x, xt_shift=shift, noise=noise if use_noise else None, k_x0=k_x0, k_xt=k_xt,
                           sigma=sigma if use_noise else 0.0, want_conf=want_x0 or (self.extract_matches and not fused_match),
                           x_min=x_min, noise_seed=self.noise_seed if gen else None,
                           noise_offset=0 if noise_counter is not None else self.noise_calls,
                           noise_offset_dev=noise_counter, out=x_out, want_best=fused_match,
                           best_floor=(None if self.flavour == "2d3d" else m.confidence_threshold) if fused_match else None)
        res = list(res) if isinstance(res, tuple) else [res]
        x_next = res.pop(0)
        x0 = res.pop(0) if (want_x0 or (self.extract_matches and not fused_match)) else None
        if noise_counter is not None:
            ops.counter_add(noise_counter, 1)       # device-side Philox offset: graph replays draw fresh noise
        self.noise_calls += 1
        aux = {"pose": pose, "x0": x0, "conf_d": conf_d}
        if self.extract_matches:
            # Mutual top-1 matches (index based: a row's best column whose best row is that row), thresholded for
            # the 3D flavours.  Equals get_match(conf, thr, mutual=True) (matching.py:71-88) except at exact value
            # ties, where the lowest index wins instead of every tied entry being reported.  With aligned rows the
            # row / column bests come out of the DDIM pass itself and x0 is neither stored nor re-read.
            thr = None if self.flavour == "2d3d" else m.confidence_threshold
            if fused_match:
                rowbest, colbest = res
                aux["match"] = ops.match_from_best(rowbest, colbest, M, thr, capacity=B * min(N, M))
                aux["match"] = (aux["match"][0], aux["match"][1], None, aux["match"][2])
            else:
                aux["match"] = ops._match(x0, 1, True, thr, True, False, capacity=B * min(N, M))
        return x_next, x_min, aux

    @torch.no_grad()
    def sample(self, x_T, src_feats, tgt_feats, s_pcd, t_pcd, src_mask, tgt_mask, noises=None, pose_tgt_pcd=None,
               pose_tgt_mask=None, feature_fn=None, trace=None, src_pe=None, tgt_pe=None, pe_type="rotary"):
        x = x_T
        shift = ops.min_value(x) if self.flavour == "3d" else None
        aux = None
        for k in range(self.steps):
            noise = noises[k] if (noises is not None and self.flavour == "4d") else None
            x_in = x
            x, shift, aux = self.step(k, x, shift, src_feats, tgt_feats, s_pcd, t_pcd, src_mask, tgt_mask, noise,
                                      pose_tgt_pcd, pose_tgt_mask, feature_fn, want_x0=trace is not None,
                                      src_pe=src_pe, tgt_pe=tgt_pe, pe_type=pe_type)
            if trace is not None:
                trace.append({"x_in": x_in, "x_out": x, **aux})
        out = {"x_final": x, "pose": aux["pose"] if aux else None}
        if self.flavour == "4d":
            out["conf_matrix_pred"] = ops.sigmoid(x)                      # pipeline.py:192
        else:
            sm = self.state_matching
            conf = ops.sinkhorn(x, sm.bin_score, sm.skh_iters, src_mask, tgt_mask, out_mode="conf", apply_mask=True,
                                shift=shift)
            r, c, w = ops.top1_select(conf.squeeze(0), True, None, False)  # 3d pipeline.py:275-277 / 2d3d model.py:692-694
            out["conf_matrix_pred"] = conf
            out["match_pred"] = torch.stack((torch.zeros_like(r), r, c), dim=-1)
            out["match_weights"] = w
        return out
