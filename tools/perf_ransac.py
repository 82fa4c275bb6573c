"""Time the correspondence RANSAC (drg_ransac_correspondence) at the evaluation's size: 50 000 trials (the reference's
RANSACConvergenceCriteria, Diff-Reg-4dmatch/models/loss.py:23) over C correspondences.  CUDA events, warm-up, median of 20.
The kernel is FP32-issue bound: one (trial, correspondence) pair costs 9 FFMA (R s + t) + 3 FADD + 3 FFMA/FMUL (d^2) +
compare / count / error sum ~ 18 issue slots and three shared-memory broadcast loads."""
import json
import sys

import torch

sys.path.insert(0, __file__.rsplit("/tools/", 1)[0])
from diffreg_b200 import ops  # noqa: E402


def main():
    res = []
    for B, C in ((1, 500), (1, 2000), (1, 4096), (8, 2000)):
        g = torch.Generator().manual_seed(C)
        N = M = 4096
        src = (torch.rand(B, N, 3, generator=g) * 2 - 1).cuda()
        tgt = (torch.rand(B, M, 3, generator=g) * 2 - 1).cuda()
        rows = []
        for b in range(B):
            i = torch.randperm(N, generator=g)[:C]
            j = torch.randperm(M, generator=g)[:C]
            rows.append(torch.stack([torch.full((C,), b, dtype=torch.int64), i, j], 1))
        match = torch.cat(rows).cuda()
        tgt[0, match[: C // 4, 2]] = src[0, match[: C // 4, 1]] + 0.25
        T = 50000
        for _ in range(3):
            out = ops.ransac_correspondence(src, tgt, match, 0.05, 3, T, seed=1)
        torch.cuda.synchronize()
        times = []
        for _ in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = ops.ransac_correspondence(src, tgt, match, 0.05, 3, T, seed=1)
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        times.sort()
        ms = times[len(times) // 2]
        pairs = B * T * C
        res.append({"B": B, "C": C, "trials": T, "ms": round(ms, 4), "trial_corr_pairs_per_s": pairs / ms * 1e3,
                    "fp32_issue_slots_per_s": 18 * pairs / ms * 1e3, "fitness0": float(out["fitness"][0])})
        print(json.dumps(res[-1]), flush=True)


if __name__ == "__main__":
    main()
