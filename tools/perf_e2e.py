"""End-to-end loop variants of HostStepPipeline at the headline shape (tuning tool): result copies on the compute stream
or on their own stream; the host waiting for step i before launching step i + 1, or running one step ahead."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
import diffreg_b200
import bench

dev = torch.device("cuda", 0)
n, c = 4096, 256
host = bench.make_inputs(3000, n, c)
d = {k: v.to(dev) for k, v in host.items()}
head = diffreg_b200.Matching(bench.MATCH_CFG).to(dev).eval()
with torch.no_grad():
    head.src_proj.weight.copy_(d["W"])
proc = diffreg_b200.SoftProcrustesLayer(SimpleNamespace(sample_rate=1.0, max_condition_num=40.0))
smp = diffreg_b200.DenoisingSampler("4d", head, proc, bench.SAMPLER_STEPS, noise_seed=1234)
K = 200
for results_stream in (False, True):
    pipe = diffreg_b200.HostStepPipeline(smp, n, n, c, dev, results_stream=results_stream)
    for slot in range(2):
        for k_, v_ in pipe.staging(slot).items():
            v_.copy_(host[k_])
    for ahead in (False, True):
        for copy_in in (True, False):
            pipe.reset(d["x_T"])
            def run(first, count):
                end = first + count
                if copy_in: pipe.prefetch(first)
                if ahead:
                    pipe.launch(first)
                    if copy_in: pipe.prefetch(first + 1)
                for i in range(first, end):
                    if ahead:
                        if i + 1 < end:
                            pipe.launch(i + 1)
                            if copy_in: pipe.prefetch(i + 2)
                    else:
                        pipe.launch(i)
                        if copy_in: pipe.prefetch(i + 1)
                    int(pipe.finish(i)["count"][0])
            run(0, 10)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            run(10, K)
            torch.cuda.synchronize()
            dt = time.perf_counter() - t0
            print(f"results_stream={results_stream} ahead={ahead} h2d={copy_in}: {K / dt:.0f} steps/s ({1e6 * dt / K:.0f} us/step)", flush=True)
